"""The host-side drop-in surface on a real GPU: seam B2, seam B1 and the CLI over the engine."""

import json
import os
import subprocess
import sys
import threading
from pathlib import Path

import pytest

from advspec_b200 import completion as comp
from advspec_b200 import models, runtime
from advspec_b200.tokenizer import SyntheticTokenizer, generate_spec

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(autouse=True)
def _small_generation(monkeypatch):
    monkeypatch.setenv("ADVSPEC_MAX_NEW_TOKENS", "12")
    monkeypatch.setenv("ADVSPEC_DEVICES", "0")
    yield
    runtime.POOL.close()


def _spec(n=200):
    return generate_spec(SyntheticTokenizer(1024), n, seed=3)


def test_heterogeneous_panel_through_seam_b2(cuda_device, monkeypatch):
    """Two replicas of one model share a prefill; a different model prefills its own prompt; a model
    string nobody serves fails alone (3 tries) without sinking the others."""
    monkeypatch.setattr(models.time, "sleep", lambda s: None)
    monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
    panel = ["b200/tiny-llama", "b200/tiny-qwen2", "b200/tiny-llama", "nobody/serves-this"]
    res = models.call_models_parallel(panel, _spec(), 1, "tech")
    by = {}
    for r in res:
        by.setdefault(r.model, []).append(r)
    assert len(res) == 4 and len(by["b200/tiny-llama"]) == 2
    for name in ("b200/tiny-llama", "b200/tiny-qwen2"):
        for r in by[name]:
            assert r.error is None and r.output_tokens == 12 and r.input_tokens > 200 and isinstance(r.response, str)
    a, b = by["b200/tiny-llama"]
    assert a.input_tokens == b.input_tokens and a.response != b.response, "same prompt, different seeds"
    bad = by["nobody/serves-this"][0]
    assert bad.error and "not a local B200 model" in bad.error and bad.response == ""
    assert models.cost_tracker.total_output_tokens == 36


def test_seam_b1_coalesces_on_the_gpu(cuda_device, monkeypatch):
    monkeypatch.setenv("ADVSPEC_COALESCE_MS", "200")
    msgs = [{"role": "system", "content": "You review specs."}, {"role": "user", "content": _spec(120)}]
    out = []

    def one():
        out.append(comp.completion(model="b200/tiny-llama", messages=msgs, max_tokens=8000, timeout=60,
                                   temperature=0.7))

    ts = [threading.Thread(target=one) for _ in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(out) == 3 and all(r.usage.completion_tokens == 12 for r in out)
    assert len({r.usage.prompt_tokens for r in out}) == 1
    eng = next(iter(runtime.POOL._engines.values())).engine
    assert eng.timing().decode_batch == 3, "the three calls must have been decoded as one batch"


def test_cli_critique_json_on_the_gpu(cuda_device, tmp_path):
    env = dict(os.environ, ADVSPEC_MAX_NEW_TOKENS="8", ADVSPEC_DEVICES="0", HOME=str(tmp_path))
    p = subprocess.run([sys.executable, str(ROOT / "adversarial-spec_b200" / "debate.py"), "critique", "--models",
                        "b200/tiny-llama,b200/tiny-llama", "--doc-type", "prd", "--json", "--round", "2"],
                       input=_spec(150), capture_output=True, text=True, env=env, cwd=tmp_path, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.loads(p.stdout)
    assert out["round"] == 2 and out["doc_type"] == "prd" and out["all_agreed"] is False
    assert [r["model"] for r in out["results"]] == ["b200/tiny-llama"] * 2
    assert all(r["output_tokens"] == 8 and r["error"] is None and r["cost"] == 0.0 for r in out["results"])
    assert out["cost"]["output_tokens"] == 16
    assert "Calling 2 model(s) (critiquing): b200/tiny-llama, b200/tiny-llama..." in p.stderr
