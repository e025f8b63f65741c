"""BASELINE configs[0]: the UNMODIFIED reference CLI -> `litellm` (oracle/litellm_shim) -> HF GPT-2 on the
host CPU.  Only runs where the reference tree exists (the build container)."""

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/skills/adversarial-spec/scripts/debate.py")


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the build container")
def test_reference_cli_over_cpu_gpt2(tmp_path):
    from advspec_b200.tokenizer import SyntheticTokenizer, generate_spec

    spec = generate_spec(SyntheticTokenizer(50257), 256, seed=1)
    env = {k: v for k, v in os.environ.items() if not k.endswith("_API_KEY")}
    env.update(PYTHONPATH=f"{ROOT / 'oracle' / 'litellm_shim'}:{ROOT}", HOME=str(tmp_path), ADVSPEC_MAX_NEW_TOKENS="8")
    p = subprocess.run([sys.executable, str(REF), "critique", "--models", "local/gpt2-124m", "--doc-type", "tech",
                        "--json"], input=spec, capture_output=True, text=True, env=env, cwd=tmp_path, timeout=600)
    assert p.returncode == 0, p.stderr[-1500:]
    out = json.loads(p.stdout)
    r = out["results"][0]
    assert out["models"] == ["local/gpt2-124m"] and r["error"] is None
    assert r["output_tokens"] == 8 and r["input_tokens"] > 256 and r["cost"] > 0  # unknown prefix: $5/$15 default
    assert out["all_agreed"] is False and "no [SPEC] tags found" in p.stderr
