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


def test_export_tasks_runs_on_the_engine(cuda_device, tmp_path):
    """`debate.py export-tasks` is the other `completion` call site (reference
    debate.py:688-736) — one sequence, temperature 0.3, no retry — here through the CUDA engine."""
    env = dict(os.environ, ADVSPEC_MAX_NEW_TOKENS="24", ADVSPEC_DEVICES="0", HOME=str(tmp_path))
    p = subprocess.run([sys.executable, str(ROOT / "adversarial-spec_b200" / "debate.py"), "export-tasks", "--models",
                        "b200/tiny-llama", "--doc-type", "prd", "--json"],
                       input=_spec(300), capture_output=True, text=True, env=env, cwd=tmp_path, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert json.loads(p.stdout) == {"tasks": []}, "a random-init model emits no [TASK] blocks"
    # the same call in-process: one sequence was decoded, sampled at T = 0.3 from the engine's own logits
    from advspec_b200 import envelope
    from oracle import sampling_ref

    prompt = envelope.EXPORT_TASKS_PROMPT.format(doc_type_name=envelope.get_doc_type_name("prd"), spec=_spec(300))
    r = comp.completion(model="b200/tiny-llama", messages=[{"role": "user", "content": prompt}], max_tokens=8000,
                        temperature=0.3)
    assert r.usage.completion_tokens == 12 and r.usage.prompt_tokens > 300
    res = next(iter(runtime.POOL._engines.values()))
    e = res.engine
    assert e.timing().decode_batch == 1
    ids = res.tok.encode(runtime.render_prompt(res.tok, "", prompt), bos=True)
    pid = e.prefill(ids)
    lg = e.get_logits(1)[0]
    seed = runtime.opponent_seed(0, (comp._call_counter - 1) * 16)
    tok0 = e.decode(e.fork(pid, [seed]), 1, temperature=0.3).tokens[0][0]
    want, gap = sampling_ref.sample(lg, 0.3, seed, 0)
    assert tok0 == want or gap < 1e-3


def test_prefix_reuse_across_calls_matches_a_full_prefill(cuda_device, diag, monkeypatch):
    """a second call on the same round and document that only adds a `--focus` / context
    section after the spec prefills its tail against the kept prefix KV (advspec_prefill_extend).  The
    logits after the extended prefill must equal a from-scratch prefill of the same prompt (same kernels,
    another chunking -> accumulation-order noise only) and the HF oracle's."""
    from oracle import hf_oracle
    from tests.gpu_util import TOL_MAX, TOL_RMS, make_engine, rel_errors
    import numpy as np

    spec, model, e = make_engine("tiny-llama-128", 31, max_prefix=2048)
    rng = np.random.default_rng(3)
    head = rng.integers(0, spec.vocab_size, 900).tolist()
    tail_a = rng.integers(0, spec.vocab_size, 40).tolist()
    tail_b = rng.integers(0, spec.vocab_size, 133).tolist()
    pid = e.prefill(head + tail_a)
    t_full = e.timing().prefill_ms
    pid2 = e.prefill_extend(pid, len(head), tail_b)  # keep 900 tokens (not a multiple of any tile), new tail
    t_ext = e.timing().prefill_ms
    got = e.get_logits(1)[0].copy()
    ref = hf_oracle.hf_logits(model, head + tail_b)[-1]
    mx, rms = rel_errors(got, ref)
    scratch_pid = e.prefill(head + tail_b)
    scratch = e.get_logits(1)[0].copy()
    mx2, rms2 = rel_errors(got, scratch)
    diag["prefix_extend/tiny-llama-128"] = {"vs_hf": {"max": mx, "rms": rms}, "vs_scratch": {"max": mx2, "rms": rms2},
                                            "full_ms": t_full, "extend_ms": t_ext}
    assert mx < TOL_MAX and rms < TOL_RMS and mx2 < 0.03, (mx, rms, mx2)
    # decode continues from an extended prefix exactly as from a fresh one
    pid3 = e.prefill_extend(scratch_pid, len(head), tail_b)
    ids = e.fork(pid3, [5, 6])
    e.decode_step(ids, [7, 9])
    step = e.get_logits(2)
    ref2 = hf_oracle.hf_logits(model, head + tail_b + [7])[-1]
    mx3, rms3 = rel_errors(step[0], ref2)
    assert mx3 < TOL_MAX and rms3 < TOL_RMS, (mx3, rms3)
    with pytest.raises(Exception):
        e.prefill_extend(pid, 10, [1, 2])  # a dead prefix id is refused
    e.close()

    # through the public API: call 2 (= call 1 + a focus section) prefills only its tail
    monkeypatch.setattr(models.time, "sleep", lambda s: None)
    doc = _spec(1500)
    runtime.PREFIXES.stats.update(full=0, extended=0, rearmed=0, tokens_reused=0, tokens_prefilled=0)
    r1 = models.call_models_parallel(["b200/tiny-llama"] * 2, doc, 3, "tech")
    r2 = models.call_models_parallel(["b200/tiny-llama"] * 2, doc, 3, "tech", focus="security")
    st = runtime.PREFIXES.stats
    assert all(r.error is None for r in r1 + r2)
    assert (st["full"], st["extended"]) == (1, 1) and st["tokens_reused"] > 1500
    assert r2[0].input_tokens > r1[0].input_tokens


@pytest.mark.parametrize("impl", ["chunk", "step"])
def test_per_opponent_tails_match_hf_and_share_one_prefill(cuda_device, diag, monkeypatch, impl):
    """opponents whose prompts differ only at the end (per-opponent personas) share ONE
    prefill of the common tokens; each opponent's own tail (37 / 1 / 12 tokens) is fed after the fork — as one
    prompt chunk per opponent whose K/V moves into the opponent's own KV (`advspec_append_tail`, default), or
    token by token through the batched decode step, right-aligned (the batch is 1, then 2, then 3 opponents
    wide).  Either way every opponent then sits at a different position, the logits it samples its first
    token from must match an HF forward of ITS full prompt, greedy decoding must continue from them within
    the stated tolerance, and the shared prefix must be untouched."""
    from oracle import hf_oracle
    from tests.gpu_util import TOL_MAX, TOL_RMS, make_engine, rel_errors
    import numpy as np

    monkeypatch.setenv("ADVSPEC_TAIL_IMPL", impl)
    spec, model, e = make_engine("tiny-llama-128", 41, max_prefix=2048, max_new=128)
    rng = np.random.default_rng(8)
    head = rng.integers(0, spec.vocab_size, 700).tolist()
    tails_in = [rng.integers(0, spec.vocab_size, n).tolist() for n in (37, 1, 12)]
    prompts = [head + t for t in tails_in]
    shared, tails = runtime.plan_tails(prompts, tail_max=64)
    assert shared == head and tails == tails_in
    pid = e.prefill(shared)
    ids = e.fork(pid, [1, 2, 3])
    runtime.feed_tails(e, ids, tails)
    lg = e.get_logits(3)
    worst = (0.0, 0.0)
    for i in range(3):
        mx, rms = rel_errors(lg[i], hf_oracle.hf_logits(model, prompts[i])[-1])
        worst = (max(worst[0], mx), max(worst[1], rms))
    diag[f"persona_tails/{impl}/tiny-llama-128"] = {"max": worst[0], "rms": worst[1]}
    assert worst[0] < TOL_MAX and worst[1] < TOL_RMS, worst
    if impl == "chunk":
        with pytest.raises(Exception):
            e.append_tail(ids[0], [1, 2])  # a tail directly follows the fork: this opponent already has one
    res = e.decode(ids, 6, temperature=0.0)
    assert res.lens == [6, 6, 6]
    for i in range(3):
        assert res.tokens[i][0] == int(lg[i].argmax())  # token 0 comes from the logits after the tail
        hist = list(prompts[i])
        for tok in res.tokens[i]:
            ref = hf_oracle.hf_logits(model, hist)[-1]
            assert ref.max() - ref[tok] <= TOL_MAX * ref.std(), (i, tok, int(ref.argmax()))
            hist.append(tok)
    # the prefix is as it was: a new opponent forked from it continues the SHARED prompt
    e.release_seqs(ids)
    (again,) = e.fork(pid, [9])
    e.decode_step([again], [5])
    mx, rms = rel_errors(e.get_logits(1)[0], hf_oracle.hf_logits(model, head + [5])[-1])
    assert mx < TOL_MAX and rms < TOL_RMS, (mx, rms)
    if impl == "chunk":
        (late,) = e.fork(pid, [10])
        with pytest.raises(Exception):
            e.append_tail(late, list(range(129)))  # beyond the opponent's own KV capacity (max_new 128)
        with pytest.raises(Exception):
            e.append_tail(7, [1])  # not forked
    e.close()
    if impl == "step":
        return

    # through seam B2: three personas, one prefill, one decode batch of three
    monkeypatch.setattr(models.time, "sleep", lambda s: None)
    monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
    runtime.PREFIXES.stats.update(full=0, extended=0, rearmed=0, tokens_reused=0, tokens_prefilled=0)
    panel = ["b200/tiny-llama@security-engineer", "b200/tiny-llama@oncall-engineer", "b200/tiny-llama"]
    res = models.call_models_parallel(panel, _spec(1200), 1, "prd")
    assert sorted(r.model for r in res) == sorted(panel) and all(r.error is None and r.output_tokens == 12 for r in res)
    by = {r.model: r for r in res}
    assert by[panel[0]].input_tokens > by[panel[2]].input_tokens and by[panel[1]].input_tokens > by[panel[2]].input_tokens
    st = runtime.PREFIXES.stats
    assert (st["full"], st["extended"]) == (1, 0) and st["tokens_prefilled"] < by[panel[2]].input_tokens
    eng = next(iter(runtime.POOL._engines.values())).engine
    assert eng.timing().decode_batch == 3, "three personas, ONE decode batch"


def test_resident_server_keeps_engines_across_cli_invocations(cuda_device, tmp_path):
    """the reference runs one `debate.py` process per round (`--session`, then
    `--resume`); with the resident server the second and third invocation find the engine loaded."""
    import time

    sock = str(tmp_path / "advspec.sock")
    env = dict(os.environ, ADVSPEC_MAX_NEW_TOKENS="8", ADVSPEC_DEVICES="0", HOME=str(tmp_path))
    srv = subprocess.Popen([sys.executable, str(ROOT / "adversarial-spec_b200" / "resident.py"), "serve", "--socket", sock],
                           env=env, cwd=tmp_path, stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(600):
            if os.path.exists(sock):
                break
            assert srv.poll() is None, srv.stderr.read()
            time.sleep(0.1)
        # (like the reference, `--resume` without `--models` needs a provider key for the default model)
        cli = [sys.executable, str(ROOT / "adversarial-spec_b200" / "debate.py"), "critique", "--json",
               "--models", "b200/tiny-llama,b200/tiny-llama"]
        cenv = dict(env, ADVSPEC_SERVER=sock)
        walls = []
        for i, extra in enumerate((["--session", "s1"], ["--resume", "s1"], ["--resume", "s1"])):
            t0 = time.perf_counter()
            p = subprocess.run(cli + extra, input=_spec(200) if i == 0 else "", capture_output=True, text=True,
                               env=cenv, cwd=tmp_path, timeout=300)
            walls.append(time.perf_counter() - t0)
            assert p.returncode == 0, p.stderr[-2000:]
            out = json.loads(p.stdout)
            assert out["round"] == i + 1 and out["session"] == "s1" and len(out["results"]) == 2
            assert all(r["output_tokens"] == 8 and r["error"] is None for r in out["results"])
        from advspec_b200 import resident

        stats = resident.request(sock, [], op="stats")
        assert stats == {"engines_resident": 1, "engines_created": 1}, stats
        sess = json.loads((tmp_path / ".config" / "adversarial-spec" / "sessions" / "s1.json").read_text())
        assert sess["round"] == 4 and len(sess["history"]) == 3
        assert sorted(p.name for p in (tmp_path / ".adversarial-spec-checkpoints").glob("*.md")) == \
            ["s1-round-1.md", "s1-round-2.md", "s1-round-3.md"]
        resident.request(sock, [], op="shutdown")
        srv.wait(timeout=60)
    finally:
        if srv.poll() is None:
            srv.terminate()
