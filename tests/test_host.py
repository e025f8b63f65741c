"""Host-side logic: tokenizer, envelope, placement, local-panel fan-out, B1 coalescing."""

import hashlib
import json
import threading

import numpy as np
from pathlib import Path
from unittest.mock import patch

import pytest

from advspec_b200 import completion as comp
from advspec_b200 import envelope, models, runtime
from advspec_b200.model_spec import REGISTRY, resolve
from advspec_b200.runtime import Generation
from advspec_b200.tokenizer import SyntheticTokenizer, generate_spec, render_chat

GOLDEN = Path(__file__).parent / "golden"
REF_SCRIPTS = Path("/root/reference/skills/adversarial-spec/scripts")


def test_tokenizer_roundtrip_and_exact_lengths():
    tok = SyntheticTokenizer(1024)
    text = "The system must store data.\n\n## Limits\n[AGREE] ünïcode [SPEC]x[/SPEC]"
    assert tok.decode(tok.encode(text)) == text
    for n in (1, 17, 256, 4096):
        s = generate_spec(tok, max(n, 8), seed=n)
        assert tok.count(s) == max(n, 8) and tok.count(s.strip()) == max(n, 8)
    big = SyntheticTokenizer(128256)
    ids = list(range(0, 128256, 997))
    assert isinstance(big.decode(ids), str)  # every id decodes
    assert generate_spec(big, 300, 5) == generate_spec(big, 300, 5)


def test_envelope_slot_order_matches_reference_capture():
    g = json.loads((GOLDEN / "reference_messages.json").read_text())
    for row in g["rows"]:
        sysm, usr = envelope.build_messages(g["spec"], g["round"], **row["kwargs"])
        assert row["roles"] == ["system", "user"] and row["max_tokens"] == 8000 and row["temperature"] == 0.7
        # same slot order as the reference: round line first, spec inside, instruction last
        assert usr.startswith(f"This is round {g['round']} of adversarial spec development.")
        assert usr[:40] == row["user_prefix"][:40]
        assert g["spec"] in usr
        if row["kwargs"].get("context"):
            assert usr.index(g["spec"]) < usr.index("ctx")
        if row["kwargs"].get("focus"):
            assert "CRITICAL FOCUS" in usr and usr.index(g["spec"]) < usr.index("CRITICAL FOCUS")
        if row["kwargs"].get("persona") == "Space Pirate":
            assert sysm.startswith("You are a Space Pirate participating in adversarial spec development")


@pytest.mark.skipif(not REF_SCRIPTS.exists(), reason="reference tree only exists in the build container")
def test_envelope_is_byte_identical_with_reference_prompts(monkeypatch):
    import importlib

    monkeypatch.setenv("ADVSPEC_REFERENCE_SCRIPTS", str(REF_SCRIPTS))
    env2 = importlib.reload(envelope)
    try:
        assert env2.SOURCE == "reference"
        g = json.loads((GOLDEN / "reference_messages.json").read_text())
        for row in g["rows"]:
            sysm, usr = env2.build_messages(g["spec"], g["round"], **row["kwargs"])
            assert hashlib.sha256(sysm.encode()).hexdigest() == row["system_sha256"]
            assert hashlib.sha256(usr.encode()).hexdigest() == row["user_sha256"]
    finally:
        monkeypatch.delenv("ADVSPEC_REFERENCE_SCRIPTS")
        importlib.reload(envelope)


def test_model_table_matches_survey():
    got = {n: round(resolve(n).n_params() / 1e9, 3) for n in
           ["llama-3-8b", "mistral-7b", "qwen2-7b", "phi-3-mini", "gemma-7b", "llama-3-70b"]}
    assert got == {"llama-3-8b": 8.030, "mistral-7b": 7.242, "qwen2-7b": 7.616, "phi-3-mini": 3.821,
                   "gemma-7b": 8.538, "llama-3-70b": 70.554}
    s = resolve("b200/llama-3-8b")
    assert s.kv_bytes_per_token == 131072
    assert round(s.decode_weight_bytes() / 1e9, 2) == 15.01
    assert round(s.prefill_flops(4096) / 1e12, 1) == 61.6
    assert round(s.decode_step_bytes(4096, [256]) / 1e9, 2) == 15.58


def test_placement_policies():
    names = ["b200/llama-3-8b"] * 3
    p = runtime.plan_placement(names, [0], "batch")
    assert len(p) == 1 and p[0].indices == [0, 1, 2] and p[0].device == 0
    p = runtime.plan_placement(names, [0, 1, 2, 3], "spread")
    assert sorted((x.device, tuple(x.indices)) for x in p) == [(0, (0,)), (1, (1,)), (2, (2,))]
    het = ["b200/llama-3-8b", "b200/mistral-7b", "b200/llama-3-8b", "b200/qwen2-7b"]
    p = runtime.plan_placement(het, [0, 1], "batch")
    assert {(x.spec.name, x.device, tuple(x.indices)) for x in p} == {
        ("llama-3-8b", 0, (0, 2)), ("mistral-7b", 1, (1,)), ("qwen2-7b", 0, (3,))}


def _fake_round(calls):
    def run_round(model_names, system_prompt, user_message, seeds, max_tokens, temperature, devices=None):
        calls.append((list(model_names), list(seeds), max_tokens, temperature, system_prompt, user_message))
        return [Generation(f"critique {i} [SPEC]s{i}[/SPEC]", 111, 7, [1] * 7) for i in range(len(model_names))]
    return run_round


def test_local_panel_is_one_batched_round(monkeypatch):
    calls = []
    monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
    with patch.object(runtime, "run_round", _fake_round(calls)):
        res = models.call_models_parallel(["b200/llama-3-8b"] * 3, "SPEC BODY", 2, "prd")
    assert len(calls) == 1, "same-weight opponents must share ONE engine round"
    names, seeds, max_tokens, temp, sysm, usr = calls[0]
    assert names == ["b200/llama-3-8b"] * 3 and len(set(seeds)) == 3
    assert (max_tokens, temp) == (8000, 0.7) and "SPEC BODY" in usr and usr.startswith("This is round 2")
    assert [r.model for r in res] == ["b200/llama-3-8b"] * 3
    assert all(r.input_tokens == 111 and r.output_tokens == 7 and r.error is None for r in res)
    assert [r.spec for r in res] == ["s0", "s1", "s2"]
    assert models.cost_tracker.total_output_tokens == 21


def test_local_group_failure_is_retried_then_reported_without_sinking_others(monkeypatch):
    attempts = {"n": 0}

    def run_round(model_names, *a, **k):
        attempts["n"] += 1
        return [RuntimeError("gpu on fire") if "mistral" in m else Generation("ok [AGREE] [SPEC]x[/SPEC]", 5, 2, [1, 2])
                for m in model_names]

    sleeps = []
    monkeypatch.setattr(models.time, "sleep", lambda s: sleeps.append(s))
    with patch.object(runtime, "run_round", run_round):
        res = models.call_models_parallel(["b200/llama-3-8b", "b200/mistral-7b"], "S", 1, "tech")
    by = {r.model: r for r in res}
    assert by["b200/llama-3-8b"].agreed and by["b200/llama-3-8b"].error is None
    assert by["b200/mistral-7b"].error == "gpu on fire" and by["b200/mistral-7b"].response == ""
    assert attempts["n"] == 3 and sleeps == [1.0, 2.0]  # the reference's schedule (models.py:46-47)


def test_b1_completion_coalesces_identical_concurrent_calls(monkeypatch):
    calls = []
    monkeypatch.setenv("ADVSPEC_COALESCE_MS", "150")
    msgs = [{"role": "system", "content": "sys"}, {"role": "user", "content": "usr"}]
    out = []
    with patch.object(runtime, "run_round", _fake_round(calls)):
        def one():
            out.append(comp.completion(model="b200/tiny-llama", messages=msgs, max_tokens=8000, timeout=60,
                                       temperature=0.7))
        ts = [threading.Thread(target=one) for _ in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    assert len(calls) == 1 and len(calls[0][0]) == 3, "three identical prompts -> one prefill"
    assert sorted(r.choices[0].message.content for r in out) == [f"critique {i} [SPEC]s{i}[/SPEC]" for i in range(3)]
    assert all(r.usage.prompt_tokens == 111 and r.usage.completion_tokens == 7 for r in out)


def test_non_local_model_without_litellm_raises():
    with pytest.raises(RuntimeError, match="not a local B200 model"):
        comp.completion(model="fake/a", messages=[], max_tokens=1, timeout=1, temperature=0.7)


def test_render_chat_and_budget_env(monkeypatch):
    assert render_chat("S", "U") == "[SYSTEM]\nS\n[USER]\nU\n[ASSISTANT]\n"
    monkeypatch.setenv("ADVSPEC_MAX_NEW_TOKENS", "64")
    assert runtime.effective_max_new(8000) == 64
    monkeypatch.delenv("ADVSPEC_MAX_NEW_TOKENS")
    assert runtime.effective_max_new(8000) == 8000
    assert len({runtime.opponent_seed(1, i) for i in range(8)}) == 8
    assert "tiny-llama" in REGISTRY


@pytest.mark.parametrize("tp", [2, 4])
def test_shard_blob_partitions_every_tensor_exactly(tp):
    """weights.shard_blob: the ranks' shares tile the whole model's tensors without overlap or loss
    (q/k/v rows per head range, wo / wd columns, gate/up row pairs, lm_head rows; norms and the
    embedding table replicated), byte for byte, in the layout csrc/engine.cu computes for a rank."""
    import ctypes as C

    from advspec_b200 import engine as eng, weights
    from advspec_b200.model_spec import ModelSpec

    spec = ModelSpec("shard-test", "llama", 2, 64, 8, 4, 64, 128, 96, qkv_bias=True)
    full_lay = weights.blob_layout(spec)
    rng = np.random.default_rng(0)
    full = rng.integers(0, 255, full_lay.total, dtype=np.uint8)

    def view(buf, lay, key):
        off, shape, kind = lay.offsets[key]
        n = int(np.prod(shape)) * (2 if kind == "bf16" else 4)
        return buf[off: off + n].view(np.uint16 if kind == "bf16" else np.uint32).reshape(shape)

    loc = weights.tp_local_spec(spec, tp)
    lay = weights.blob_layout(loc, spec.vocab_size)
    shards = [weights.shard_blob(full, spec, r, tp) for r in range(tp)]
    try:
        lib = eng.load_library()
        for r in range(tp):
            d = eng.make_desc(spec, 64, 16, 2, r, tp)
            assert lib.advspec_weight_blob_bytes(C.byref(d)) == shards[r].nbytes == lay.total
    except eng.EngineError:
        pass  # library not built: the layout agreement is covered by tests/test_abi.py
    dh, H, Hkv = spec.head_dim, spec.n_heads, spec.n_kv_heads
    for l in range(spec.n_layers):
        parts = [view(s, lay, (l, "wqkv")) for s in shards]
        hl, kl = loc.n_heads * dh, loc.n_kv_heads * dh
        q = np.concatenate([p[:hl] for p in parts]); k = np.concatenate([p[hl:hl + kl] for p in parts])
        v = np.concatenate([p[hl + kl:] for p in parts])
        assert np.array_equal(np.concatenate([q, k, v]), view(full, full_lay, (l, "wqkv")))
        b = [view(s, lay, (l, "bqkv")) for s in shards]
        assert np.array_equal(np.concatenate([np.concatenate([p[:hl] for p in b]), np.concatenate([p[hl:hl + kl] for p in b]),
                                              np.concatenate([p[hl + kl:] for p in b])]), view(full, full_lay, (l, "bqkv")))
        for name, axis in (("wo", 1), ("wd", 1), ("wgu", 0)):
            assert np.array_equal(np.concatenate([view(s, lay, (l, name)) for s in shards], axis=axis),
                                  view(full, full_lay, (l, name)))
        for name in ("attn_norm", "mlp_norm"):
            assert all(np.array_equal(view(s, lay, (l, name)), view(full, full_lay, (l, name))) for s in shards)
    assert np.array_equal(np.concatenate([view(s, lay, (-1, "lm_head")) for s in shards]), view(full, full_lay, (-1, "lm_head")))
    assert all(np.array_equal(view(s, lay, (-1, "embed")), view(full, full_lay, (-1, "embed"))) for s in shards)
    with pytest.raises(ValueError):
        weights.tp_local_spec(spec, 3)


def test_tp_world_needs_one_process_per_gpu(monkeypatch):
    monkeypatch.delenv("ADVSPEC_TP", raising=False)
    assert runtime.tp_world() == (0, 1)
    monkeypatch.setenv("ADVSPEC_TP", "4")
    monkeypatch.setenv("WORLD_SIZE", "1")
    with pytest.raises(RuntimeError, match="WORLD_SIZE=4"):
        runtime.tp_world()
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("RANK", "3")
    assert runtime.tp_world() == (3, 4)
