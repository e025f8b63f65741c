"""Host runtime logic that needs no GPU: leased engine pool, cross-call prefix reuse, per-opponent errors for
unknown local models, one retry back-off per attempt, adaptive B1 coalescing (a fake engine stands in)."""

import threading
import time

import numpy as np
import pytest

from advspec_b200 import completion as comp
from advspec_b200 import engine as eng
from advspec_b200 import models, providers, runtime
from advspec_b200.model_spec import resolve
from advspec_b200.runtime import Generation


class FakeEngine:
    """Records what the runtime asks of an engine; blows up if used after close()."""
    made = []

    def __init__(self, spec, device=0, max_prefix_tokens=0, max_new_tokens=0, max_seqs=8, **kw):
        self.spec, self.device = spec, device
        self.max_prefix, self.max_new = max_prefix_tokens, max_new_tokens
        self.closed = False
        self.calls = []
        self.live = 0
        self.busy = 0
        FakeEngine.made.append(self)

    def _alive(self):
        assert not self.closed, "engine used after the pool destroyed it"

    def init_weights_random(self, seed, std=0.02):
        self._alive()

    def close(self):
        assert self.busy == 0, "engine destroyed while a round is inside it"
        self.closed = True

    def prefill(self, tokens):
        self._alive()
        self.live += 1
        self.calls.append(("prefill", len(tokens)))
        return self.live

    def prefill_extend(self, pid, keep, tail):
        self._alive()
        if pid != self.live:
            raise eng.EngineError(4, "prefix is not live")
        self.live += 1
        self.calls.append(("extend", keep, len(tail)))
        return self.live

    def fork(self, pid, seeds):
        self._alive()
        assert pid == self.live
        return list(range(len(seeds)))

    def decode(self, ids, max_new, temperature=0.7, eos_id=-1):
        self._alive()
        self.busy += 1
        time.sleep(0.02)
        self.busy -= 1
        self._alive()
        return eng.DecodeResult([[5] * max_new for _ in ids], [max_new] * len(ids))

    def release_seqs(self, ids):
        self._alive()

    def timing(self):
        self._alive()
        return eng.Timing()


@pytest.fixture
def fake_engine(monkeypatch):
    FakeEngine.made = []
    monkeypatch.setattr(eng, "Engine", FakeEngine)
    monkeypatch.setattr(runtime, "POOL", runtime.EnginePool())
    monkeypatch.setattr(runtime, "PREFIXES", runtime.PrefixCache())
    monkeypatch.setenv("ADVSPEC_MAX_NEW_TOKENS", "4")
    yield
    runtime.POOL.close()


def test_pool_never_destroys_an_engine_under_a_concurrent_round(fake_engine):
    """two callers with different prompt lengths on one (model, device).  The short
    round is inside the engine when the long one arrives; the pool must wait for it, then replace."""
    spec = resolve("tiny-llama")
    errs = []

    def round_(n_prompt_words):
        try:
            for _ in range(6):
                out = runtime.generate_group(spec, 0, "sys", "word " * n_prompt_words, 2, [1, 2], 8000, 0.7)
                assert len(out) == 2 and all(g.completion_tokens == 4 for g in out)
        except BaseException as ex:  # noqa: BLE001 - surfaced below
            errs.append(ex)

    ts = [threading.Thread(target=round_, args=(n,)) for n in (50, 900, 120, 2000)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert runtime.POOL.resident_count() == 1
    assert sum(not e.closed for e in FakeEngine.made) == 1
    assert 1 <= len(FakeEngine.made) <= 4 and runtime.POOL.created == len(FakeEngine.made)
    # growth only: a replacement is never smaller than what it replaced
    caps = [e.max_prefix for e in FakeEngine.made]
    assert caps == sorted(caps)


def test_prefix_cache_extends_tails_rearms_identical_prompts_and_refuses_thin_overlaps(fake_engine):
    spec = resolve("tiny-llama")
    doc = "alpha beta gamma delta " * 200
    base = runtime.generate_group(spec, 0, "SYS", f"This is round 1.\n{doc}\nReview.", 1, [1], 8000, 0.7)
    e = FakeEngine.made[-1]
    n = base[0].prompt_tokens
    assert e.calls == [("prefill", n)]
    # same round and document, a focus section after the spec: only the tail is prefilled
    runtime.generate_group(spec, 0, "SYS", f"This is round 1.\n{doc}\n**CRITICAL FOCUS: SECURITY**\nReview.", 1, [1], 8000, 0.7)
    kind, keep, tail = e.calls[-1]
    assert kind == "extend" and keep > 0.9 * n and tail < 0.1 * n
    # the identical prompt again (a retry, or the second batch of a 16-opponent panel): no prefill at all
    runtime.generate_group(spec, 0, "SYS", f"This is round 1.\n{doc}\n**CRITICAL FOCUS: SECURITY**\nReview.", 1, [1], 8000, 0.7)
    assert e.calls[-1][0] == "extend" and e.calls[-1][2] == 0
    # the next ROUND diverges at the round number, a few tokens in: full prefill
    runtime.generate_group(spec, 0, "SYS", f"This is round 2.\n{doc}\nReview.", 1, [1], 8000, 0.7)
    assert e.calls[-1][0] == "prefill"
    st = runtime.PREFIXES.stats
    assert (st["full"], st["extended"], st["rearmed"]) == (2, 1, 1) and st["tokens_reused"] > n
    # someone used the engine behind the cache's back: the stale id is refused by the engine -> full prefill
    e.live += 7
    runtime.generate_group(spec, 0, "SYS", f"This is round 2.\n{doc}\nReview.", 1, [1], 8000, 0.7)
    assert e.calls[-1][0] == "prefill"


def test_prefix_cache_can_be_switched_off(fake_engine, monkeypatch):
    monkeypatch.setenv("ADVSPEC_PREFIX_CACHE", "0")
    spec = resolve("tiny-llama")
    for _ in range(2):
        runtime.generate_group(spec, 0, "S", "same prompt " * 50, 1, [1], 8000, 0.7)
    assert [c[0] for c in FakeEngine.made[-1].calls] == ["prefill", "prefill"]


def test_unknown_local_model_fails_alone(fake_engine, monkeypatch):
    """`b200/no-such-model` must become that opponent's .error after the reference's three
    tries (one back-off per attempt), with the other opponents served; the CLI pre-flight refuses it."""
    sleeps = []
    monkeypatch.setattr(models.time, "sleep", lambda s: sleeps.append(s))
    monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
    res = models.call_models_parallel(["b200/tiny-llama", "b200/no-such-model", "b200/also-missing"], "spec", 1, "tech")
    by = {r.model: r for r in res}
    assert by["b200/tiny-llama"].error is None and by["b200/tiny-llama"].output_tokens == 4
    for bad in ("b200/no-such-model", "b200/also-missing"):
        assert "unknown local model" in by[bad].error and by[bad].response == ""
    backoffs = [s for s in sleeps if s >= 1.0]  # (the fake engine's own 20 ms naps share the patched module)
    assert backoffs == [1.0, 2.0], "two failing opponents still back off once per attempt"
    valid, invalid = providers.validate_model_credentials(["b200/tiny-llama", "b200/no-such-model", "local/qwen2-7b"])
    assert valid == ["b200/tiny-llama", "local/qwen2-7b"] and invalid == ["b200/no-such-model"]
    assert "unknown local model" in providers.required_key_hint("b200/no-such-model")


def test_local_round_honours_the_call_timeout(monkeypatch):
    def slow_round(model_names, *a, **k):
        threading.Event().wait(0.5)
        return [Generation("late", 1, 1, [1]) for _ in model_names]

    monkeypatch.setattr(runtime, "run_round", slow_round)
    monkeypatch.setattr(models.time, "sleep", lambda s: None)
    monkeypatch.setattr(models, "MAX_RETRIES", 1)
    t0 = time.perf_counter()
    res = models.call_models_parallel(["b200/tiny-llama"], "s", 1, "tech", timeout=0.1)
    assert res[0].error and "timed out" in res[0].error and time.perf_counter() - t0 < 2.0


def test_b1_leader_does_not_sit_out_a_fixed_window(monkeypatch):
    """a lone `completion` call (export-tasks) returns after the quiet gap, not after the
    whole coalescing window; concurrent identical calls still share one round."""
    calls = []

    def run_round(model_names, system_prompt, user_message, seeds, max_tokens, temperature, devices=None):
        calls.append(len(model_names))
        return [Generation("x", 3, 1, [1]) for _ in model_names]

    monkeypatch.setattr(runtime, "run_round", run_round)
    monkeypatch.setenv("ADVSPEC_COALESCE_MS", "2000")
    monkeypatch.setenv("ADVSPEC_COALESCE_QUIET_MS", "20")
    msgs = [{"role": "user", "content": "solo"}]
    t0 = time.perf_counter()
    comp.completion(model="b200/tiny-llama", messages=msgs, max_tokens=8, timeout=5, temperature=0.3)
    assert time.perf_counter() - t0 < 0.5 and calls == [1]
    out = []
    ts = [threading.Thread(target=lambda: out.append(comp.completion(model="b200/tiny-llama", messages=msgs,
                                                                      max_tokens=8, timeout=5, temperature=0.3)))
          for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert calls == [1, 4] and len(out) == 4


def test_real_tokenizer_is_required_beside_a_weight_blob(monkeypatch, tmp_path):
    """a served blob without its tokenizer must fail loudly; with tokenizer.json + meta.json the
    runtime encodes with the model's own vocabulary, eos and chat template."""
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import Whitespace

    spec = resolve("tiny-llama")
    (tmp_path / "tiny-llama.blob").write_bytes(b"\0")
    monkeypatch.setenv("ADVSPEC_WEIGHTS_DIR", str(tmp_path))
    with pytest.raises(FileNotFoundError, match="needs its own tokenizer"):
        runtime.tokenizer_for(spec)
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2, "hello": 3, "spec": 4, "SYS": 5, "USR": 6}
    tk = Tokenizer(WordLevel(vocab, unk_token="<unk>"))
    tk.pre_tokenizer = Whitespace()
    tk.save(str(tmp_path / "tiny-llama.tokenizer.json"))
    (tmp_path / "tiny-llama.meta.json").write_text(
        '{"eos_token_id": 2, "bos_token_id": 1, "chat_template": "SYS {system} USR {user}"}')
    tok, meta = runtime.tokenizer_for(spec)
    assert tok.eos_id == 2 and tok.encode("hello spec", bos=True) == [1, 3, 4]
    assert tok.encode(runtime.render_prompt(tok, "hello", "spec"), bos=True) == [1, 5, 3, 6, 4]
    assert tok.decode([3, 4]) == "hello spec"
    monkeypatch.delenv("ADVSPEC_WEIGHTS_DIR")
    assert type(runtime.tokenizer_for(spec)[0]).__name__ == "SyntheticTokenizer"
