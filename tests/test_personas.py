"""Per-opponent personas: `b200/<model>@<persona>`.

Not in the reference — there `--persona` is one flag for the whole panel and replaces the system prompt
(debate.py:835, prompts.py:290-304), which would put the variable text FIRST and end prefix sharing.  Here the
persona text moves behind the document, the panel shares one prefill of the common tokens and each opponent's
own tail is stepped through the batched decode path.  Host logic on a recording engine (CPU)."""
import io
import json
import sys

import pytest

from advspec_b200 import debate, engine as eng, envelope, models, providers, runtime  # noqa: E402
from advspec_b200.model_spec import resolve, split_persona  # noqa: E402

from tests.test_host_runtime import FakeEngine  # noqa: E402


class StepEngine(FakeEngine):
    def fork(self, pid, seeds):
        ids = super().fork(pid, seeds)
        self.calls.append(("fork", len(ids)))
        return [10 + i for i in ids]  # ids that are not positions, so mix-ups show

    def decode_step(self, ids, forced):
        self._alive()
        assert len(ids) == len(forced) >= 1
        self.calls.append(("step", tuple(ids), tuple(forced)))

    def append_tail(self, sid, tokens):
        self._alive()
        assert len(tokens) >= 1
        self.calls.append(("tail", sid, len(tokens)))

    def decode(self, ids, max_new, temperature=0.7, eos_id=-1):
        self.calls.append(("decode", tuple(ids), max_new))
        return super().decode(ids, max_new, temperature, eos_id)


@pytest.fixture
def step_engine(monkeypatch):
    FakeEngine.made = []
    monkeypatch.setattr(eng, "Engine", StepEngine)
    monkeypatch.setattr(runtime, "POOL", runtime.EnginePool())
    monkeypatch.setattr(runtime, "PREFIXES", runtime.PrefixCache())
    monkeypatch.setenv("ADVSPEC_MAX_NEW_TOKENS", "4")
    yield
    runtime.POOL.close()


def test_model_suffix_is_parsed_only_on_local_models():
    assert split_persona("b200/llama-3-8b@security-engineer") == ("b200/llama-3-8b", "security-engineer")
    assert split_persona("local/tiny-llama@QA lead") == ("local/tiny-llama", "QA lead")
    assert split_persona("b200/llama-3-8b") == ("b200/llama-3-8b", None)
    assert split_persona("b200/llama-3-8b@") == ("b200/llama-3-8b", None)
    assert split_persona("vertex_ai/gemini-pro@001") == ("vertex_ai/gemini-pro@001", None)  # a provider's own '@'
    assert resolve("b200/tiny-llama@oncall-engineer") is resolve("b200/tiny-llama")
    assert providers.validate_model_credentials(["b200/tiny-llama@x", "b200/nope@x"]) == (["b200/tiny-llama@x"], ["b200/nope@x"])


def test_one_persona_for_everyone_is_the_reference_layout():
    ref = envelope.build_messages("DOC", 2, "tech", False, "security", "oncall-engineer", "CTX", True)
    system, users = envelope.build_panel_messages("DOC", 2, "tech", False, "security", ["oncall-engineer"] * 3, "CTX", True)
    assert (system, users) == (ref[0], [ref[1]] * 3)
    none = envelope.build_messages("DOC", 1, "prd")
    assert envelope.build_panel_messages("DOC", 1, "prd", False, None, [None, None]) == (none[0], [none[1]] * 2)


def test_different_personas_move_behind_the_document():
    base_sys, base_user = envelope.build_messages("DOC BODY", 1, "prd")
    system, users = envelope.build_panel_messages("DOC BODY", 1, "prd", False, None,
                                                  ["security-engineer", None, "night-shift nurse"])
    assert system == base_sys  # the doc type's default for every opponent: the prefix stays shared
    assert users[1] == base_user
    for u, p in ((users[0], "security-engineer"), (users[2], "night-shift nurse")):
        assert u.startswith(base_user) and u != base_user
        assert u.endswith(envelope.get_system_prompt("prd", p))  # the persona's own prompt text, verbatim
        assert envelope.PERSONA_TAIL_HEADER in u[len(base_user):]
    assert envelope.PERSONAS["security-engineer"] in users[0]
    assert "night-shift nurse" in users[2]  # unknown key -> the generic persona sentence (prompts.py:297)


def test_tail_plan_keeps_a_token_per_opponent_and_refuses_thin_or_long_tails():
    p = list(range(100))
    shared, tails = runtime.plan_tails([p + [1, 2, 3], p + [7], p + [1, 2, 9, 9]], tail_max=16)
    assert shared == p and tails == [[1, 2, 3], [7], [1, 2, 9, 9]]
    # one prompt is a prefix of the other: the shared part gives up its last token so that no tail is empty
    shared, tails = runtime.plan_tails([p, p + [5, 6]], tail_max=16)
    assert shared == p[:-1] and tails == [[99], [99, 5, 6]]
    assert runtime.plan_tails([p + [1] * 20, p + [2]], tail_max=16) is None  # tail too long
    assert runtime.plan_tails([[1, 2, 3] + [4] * 50, [1, 2, 3] + [5] * 50], tail_max=64) is None  # < half shared
    assert runtime.plan_tails([[1, 2], [3, 4]], tail_max=16) is None  # nothing shared


def test_tails_are_right_aligned_and_the_last_step_carries_everyone_in_fork_order():
    class Rec:
        def __init__(self):
            self.steps = []

        def decode_step(self, ids, forced):
            self.steps.append((list(ids), list(forced)))

    r = Rec()
    n = runtime.step_tails(r, [10, 11, 12], [[1, 2, 3], [7], [4, 5]])
    assert n == 3
    assert r.steps == [([10], [1]), ([10, 12], [2, 4]), ([10, 11, 12], [3, 7, 5])]


def test_panel_with_personas_feeds_each_tail_as_one_chunk_by_default(step_engine):
    spec = resolve("tiny-llama")
    doc = "alpha beta gamma delta " * 300
    system, users = envelope.build_panel_messages(doc, 1, "prd", False, None,
                                                  ["security-engineer", "oncall-engineer", None])
    out = runtime.generate_group(spec, 0, system, users, 3, [1, 2, 3], 8000, 0.7)
    e = FakeEngine.made[-1]
    kinds = [c[0] for c in e.calls]
    assert kinds == ["prefill", "fork", "tail", "tail", "tail", "decode"]
    shared = e.calls[0][1]
    assert [c[1:] for c in e.calls if c[0] == "tail"] == [(10 + i, g.prompt_tokens - shared) for i, g in enumerate(out)]
    assert e.calls[-1][:2] == ("decode", (10, 11, 12))
    assert e.max_prefix >= max(g.prompt_tokens for g in out)  # the chunk's K/V is staged behind the prefix
    assert e.max_new >= 4 + max(g.prompt_tokens for g in out) - shared


def test_panel_with_personas_is_one_prefill_one_batch(step_engine, monkeypatch):
    monkeypatch.setenv("ADVSPEC_TAIL_IMPL", "step")
    spec = resolve("tiny-llama")
    doc = "alpha beta gamma delta " * 300
    system, users = envelope.build_panel_messages(doc, 1, "prd", False, None,
                                                  ["security-engineer", "oncall-engineer", None])
    out = runtime.generate_group(spec, 0, system, users, 3, [1, 2, 3], 8000, 0.7)
    e = FakeEngine.made[-1]
    kinds = [c[0] for c in e.calls]
    assert kinds.count("prefill") == 1 and kinds.count("fork") == 1 and kinds.count("decode") == 1
    shared = e.calls[0][1]
    lens = [g.prompt_tokens for g in out]
    assert len(set(lens)) == 3 and shared < min(lens) and shared > 0.9 * min(lens)
    steps = [c for c in e.calls if c[0] == "step"]
    assert len(steps) == max(lens) - shared  # as many steps as the longest tail
    for i, g in enumerate(out):  # every opponent was fed exactly its own tail
        assert sum(1 for s in steps if 10 + i in s[1]) == g.prompt_tokens - shared
    assert steps[-1][1] == (10, 11, 12) and e.calls[-1][:2] == ("decode", (10, 11, 12))
    assert kinds.index("prefill") < kinds.index("fork") < kinds.index("step") < kinds.index("decode")
    # suffix KV is sized for the tail AND the generation
    assert e.max_new >= 4 + max(lens) - shared
    assert all(g.completion_tokens == 4 and g.tail_ms >= 0.0 for g in out)
    assert runtime.PREFIXES.stats["full"] == 1 and runtime.PREFIXES.stats["tokens_prefilled"] == shared


def test_prompts_that_differ_early_fall_back_to_a_prefill_each(step_engine):
    spec = resolve("tiny-llama")
    users = ["first document " * 80, "second one " * 80, "first document " * 80]
    out = runtime.generate_group(spec, 0, "SYS", users, 3, [1, 2, 3], 8000, 0.7)
    e = FakeEngine.made[-1]
    kinds = [c[0] for c in e.calls]
    assert "step" not in kinds and "tail" not in kinds and kinds.count("decode") == 2
    assert sorted(c[1] for c in e.calls if c[0] == "fork") == [1, 2]  # the two identical prompts share a batch
    assert out[0].prompt_tokens == out[2].prompt_tokens != out[1].prompt_tokens
    with pytest.raises(ValueError):
        runtime.generate_group(spec, 0, "SYS", users[:2], 3, [1, 2, 3], 8000, 0.7)


def test_seam_b2_routes_suffix_personas_and_prices_them_at_zero(step_engine, monkeypatch):
    seen = {}
    real = runtime.run_round

    def spy(names, system, user, seeds, max_tokens, temperature, devices=None):
        seen["names"], seen["system"], seen["user"] = list(names), system, user
        return real(names, system, user, seeds, max_tokens, temperature, devices)

    monkeypatch.setattr(runtime, "run_round", spy)
    monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
    names = ["b200/tiny-llama@security-engineer", "b200/tiny-llama", "b200/tiny-llama@junior-developer"]
    res = models.call_models_parallel(names, "spec text " * 200, 1, "tech", False, None, "qa-engineer")
    assert [r.model for r in res] == names and all(r.error is None for r in res)
    assert seen["system"] == envelope.get_system_prompt("tech", None)
    assert isinstance(seen["user"], list) and len(set(seen["user"])) == 3
    # the un-suffixed opponent keeps the panel's --persona, as a tail like the others
    assert seen["user"][1].endswith(envelope.get_system_prompt("tech", "qa-engineer"))
    assert seen["user"][0].endswith(envelope.PERSONAS["security-engineer"])
    assert models.cost_tracker.total_cost == 0.0 and set(models.cost_tracker.by_model) == set(names)
    # a panel where every opponent names the SAME persona is the reference's layout again: persona system prompt
    same = ["b200/tiny-llama@qa-engineer"] * 2
    models.call_models_parallel(same, "spec", 1, "tech")
    assert seen["system"] == envelope.PERSONAS["qa-engineer"] and isinstance(seen["user"], str)


def test_cli_accepts_persona_suffixes_and_echoes_the_model_strings(step_engine, monkeypatch, capsys):
    monkeypatch.setattr(sys, "stdin", io.StringIO("# Spec\n" + "requirement line\n" * 50))
    monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
    argv = ["critique", "--models", "b200/tiny-llama@security-engineer,b200/tiny-llama@oncall-engineer", "--json"]
    monkeypatch.setattr(sys, "argv", ["debate.py"] + argv)
    debate.main(_forward=False)
    out = json.loads(capsys.readouterr().out)
    assert out["models"] == ["b200/tiny-llama@security-engineer", "b200/tiny-llama@oncall-engineer"]
    assert [r["model"] for r in out["results"]] == out["models"] and out["persona"] is None
    assert all(r["error"] is None and r["cost"] == 0.0 for r in out["results"])
    e = FakeEngine.made[-1]
    assert [c[0] for c in e.calls].count("prefill") == 1


def test_seam_b1_coalesces_calls_whose_user_messages_differ(step_engine, monkeypatch):
    import threading

    from advspec_b200 import completion as comp

    doc = "shared specification body " * 200
    users = [doc + "\n\nperspective one", doc + "\n\nperspective two, a little longer", doc + "\n\nperspective one"]
    outs = [None] * 3

    def call(i):
        outs[i] = comp.completion(model="b200/tiny-llama", max_tokens=8000, temperature=0.7,
                                  messages=[{"role": "system", "content": "SYS"}, {"role": "user", "content": users[i]}])

    ts = [threading.Thread(target=call, args=(i,)) for i in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    e = FakeEngine.made[-1]
    kinds = [c[0] for c in e.calls]
    assert kinds.count("prefill") == 1 and kinds.count("decode") == 1 and kinds.count("tail") == 3
    assert all(o.usage.completion_tokens == 4 for o in outs)
    assert outs[0].usage.prompt_tokens == outs[2].usage.prompt_tokens < outs[1].usage.prompt_tokens


def test_bench_personas_section_reports_both_layouts(step_engine, monkeypatch):
    """bench.py's `personas` section on the recording engine: both cases run, the per-opponent case prefills
    once and steps the longest tail, nothing errors."""
    import types

    import bench

    class D:
        rank, world, local_rank = 0, 1, 0

        def barrier(self):
            pass

        def gather(self, obj):
            return [obj]

    monkeypatch.setattr(bench, "make_doc", lambda v, n, seed, title: "requirement sentence number one. " * 150)
    monkeypatch.setattr(StepEngine, "timing", lambda self: eng.Timing(prefill_ms=2.0, decode_ms=9.0, decode_steps=3,
                                                                      decode_batch=3), raising=False)
    rec = bench.section_personas(types.SimpleNamespace(gen=4), D(), {"hbm_gbs": 6489.9})
    row = rec["per_rank"][0]
    one, per = row["one_persona_for_the_panel"], row["persona_per_opponent"]
    assert one["errors"] == [] and per["errors"] == [] and one["tokens"] == per["tokens"] == 12
    assert len(set(one["input_tokens"])) == 1 and one["longest_tail"] == 0
    assert len(set(per["input_tokens"])) == 3 and per["longest_tail"] == per["input_tokens"][-1] - per["tokens_prefilled"] > 0
    assert per["prefills"] == 1 and per["decode_batch"] == 3
    kinds = [c[0] for c in FakeEngine.made[-1].calls]
    assert kinds.count("tail") == 6 and kinds.count("step") == 2 * per["longest_tail"]  # chunk case, then stepped
    assert row["persona_per_opponent_stepped"]["tokens"] == 12 and "ADVSPEC_TAIL_IMPL" not in __import__("os").environ
    assert rec["persona_per_opponent_tokens_per_s"] > 0 and rec["one_persona_for_the_panel_tokens_per_s"] > 0


def test_more_opponents_than_one_batch_share_the_prefix_across_batches(step_engine):
    """Ten opponents, two personas alternating: two decode batches (8 + 2); the second batch's shared prefix
    is the first one's, so it is re-armed, not prefilled again."""
    spec = resolve("tiny-llama")
    doc = "alpha beta gamma delta " * 300
    people = ["security-engineer", "oncall-engineer"] * 5
    system, users = envelope.build_panel_messages(doc, 1, "prd", False, None, people)
    out = runtime.generate_group(spec, 0, system, users, 10, list(range(10)), 8000, 0.7)
    e = FakeEngine.made[-1]
    kinds = [c[0] for c in e.calls]
    assert kinds.count("prefill") == 1 and [c[1] for c in e.calls if c[0] == "fork"] == [8, 2]
    assert kinds.count("tail") == 10 and kinds.count("decode") == 2
    ext = [c for c in e.calls if c[0] == "extend"]
    assert len(ext) == 1 and ext[0][2] == 0  # batch 2: same shared tokens -> re-armed, nothing prefilled
    assert len(out) == 10 and all(g is not None and g.completion_tokens == 4 for g in out)
    assert out[0].prompt_tokens == out[2].prompt_tokens != out[1].prompt_tokens
