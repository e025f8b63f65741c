"""The host mirror reproduces the UNMODIFIED reference CLI byte for byte on the golden cases
(tests/golden/reference_cli_cases.json, produced by oracle/make_golden.py from
/root/reference with a canned litellm stub)."""

import io
import json
import sys
import time
from contextlib import redirect_stderr, redirect_stdout
from pathlib import Path
from types import SimpleNamespace
from unittest.mock import patch

import pytest

from advspec_b200 import debate, models

GOLDEN = Path(__file__).parent / "golden"
CASES = json.loads((GOLDEN / "reference_cli_cases.json").read_text())


def _canned(case):
    def completion(**kw):
        spec = case["responses"][kw["model"]]
        time.sleep(spec.get("delay", 0.0))
        if spec.get("raise"):
            raise RuntimeError(spec["raise"])
        usage = None if spec.get("no_usage") else SimpleNamespace(prompt_tokens=spec["in"],
                                                                  completion_tokens=spec["out"])
        return SimpleNamespace(choices=[SimpleNamespace(message=SimpleNamespace(content=spec["content"]))],
                               usage=usage)
    return completion


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cli_matches_reference(case, monkeypatch, tmp_path):
    monkeypatch.chdir(tmp_path)
    for k in list(__import__("os").environ):
        if k.endswith("_API_KEY"):
            monkeypatch.delenv(k)
    monkeypatch.setattr(sys, "argv", ["debate.py", *case["argv"]])
    monkeypatch.setattr(sys, "stdin", io.StringIO(case["stdin"]))
    monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
    monkeypatch.setattr(debate, "cost_tracker", models.cost_tracker)
    out, err, rc = io.StringIO(), io.StringIO(), 0
    with patch.object(models, "completion", _canned(case)), patch.object(debate, "completion", _canned(case)), \
            patch.object(models.time, "sleep", lambda s, _real=time.sleep: _real(0.2 * s)), \
            redirect_stdout(out), redirect_stderr(err):
        try:
            debate.main()
        except SystemExit as e:
            rc = e.code or 0
    exp = case["expected"]
    assert rc == exp["returncode"]
    assert out.getvalue() == exp["stdout"]
    if case["name"] == "missing_key_exit2":
        assert "gpt-4o (requires OPENAI_API_KEY)" in err.getvalue()
    else:
        assert err.getvalue() == exp["stderr"]


def test_parsing_kats():
    """[AGREE]/[SPEC] semantics (reference models.py:149-160; KATs of its test_models.py:161-190)."""
    assert models.detect_agreement("fine\n[AGREE]\n")
    assert not models.detect_agreement("agree")
    assert models.extract_spec("a [SPEC]\n body \n[/SPEC] b") == "body"
    assert models.extract_spec("[SPEC]only open") is None
    assert models.extract_spec("only close[/SPEC]") is None
    assert models.extract_spec("[SPEC]a[/SPEC][SPEC]b[/SPEC]") == "a"
    assert models.extract_spec("[SPEC][/SPEC]") == ""


def test_cost_tracker_math():
    ct = models.CostTracker()
    c = ct.add("gpt-4o", 1_000_000, 500_000)
    assert c == pytest.approx(2.5 + 5.0)
    c2 = ct.add("totally/unknown", 1000, 2000)  # $5/$15 default
    assert c2 == pytest.approx(0.005 + 0.03)
    assert ct.add("b200/llama-3-8b", 10**6, 10**6) == 0.0
    assert ct.total_input_tokens == 2_001_000 and len(ct.by_model) == 3
    assert "By model:" in ct.summary()


def test_model_response_defaults():
    r = models.ModelResponse(model="m", response="r", agreed=False, spec=None)
    assert (r.error, r.input_tokens, r.output_tokens, r.cost) == (None, 0, 0, 0.0)


@pytest.mark.skipif(not Path("/root/reference/skills/adversarial-spec/scripts/models.py").exists(),
                    reason="reference tree only exists in the build container")
def test_extract_tasks_agrees_with_reference_on_random_blocks():
    """Property test in the build container: the restated task parser equals the reference's own
    function (models.py:163-247) on randomly assembled [TASK] blocks."""
    import importlib.util
    import random
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        Path(td, "litellm.py").write_text("suppress_debug_info=False\ndef completion(**k):\n    raise RuntimeError\n")
        sys.path[:0] = [td, "/root/reference/skills/adversarial-spec/scripts"]
        before = set(sys.modules)
        try:
            spec = importlib.util.spec_from_file_location("_ref_models", "/root/reference/skills/adversarial-spec/scripts/models.py")
            ref = importlib.util.module_from_spec(spec)
            sys.modules["_ref_models"] = ref  # dataclasses resolves the defining module through sys.modules
            spec.loader.exec_module(ref)
        finally:
            del sys.path[:2]
            for k in set(sys.modules) - before - {"_ref_models"}:
                del sys.modules[k]  # the stub `litellm` and the reference's prompts/providers must not leak
    rng = random.Random(0)
    lines = ["title: A", "title:", "type: bug", "priority: high", "description: d1", "more text", "- item", "- ",
             "acceptance_criteria:", "acceptance_criteria: inline", "  - indented", "", "random: x", "[TASK]", "[/TASK]"]
    for _ in range(400):
        txt = "\n".join(rng.choice(lines) for _ in range(rng.randint(0, 25)))
        assert models.extract_tasks(txt) == ref.extract_tasks(txt), txt


# ---------------------------------------------------------------- multi-step flows (session, resume, context)
FLOWS = json.loads((GOLDEN / "reference_cli_flows.json").read_text())
REF_SCRIPTS = Path("/root/reference/skills/adversarial-spec/scripts")


@pytest.mark.parametrize("flow", FLOWS, ids=[f["name"] for f in FLOWS])
def test_cli_flows_match_reference(flow, monkeypatch, tmp_path):
    """Session create -> resume, context files, checkpoints: stdout, stderr, exit codes and every file
    left behind (session JSON, round-N.md) equal what the UNMODIFIED reference CLI produced
    (tests/golden/reference_cli_flows.json, oracle/make_golden.py); with the reference's prompts.py
    available the messages put on the wire are compared as well."""
    import re
    from advspec_b200 import envelope, session

    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("HOME", str(tmp_path))
    monkeypatch.setattr(session, "SESSIONS_DIR", tmp_path / ".config" / "adversarial-spec" / "sessions")
    monkeypatch.setattr(session, "CHECKPOINTS_DIR", tmp_path / ".adversarial-spec-checkpoints")
    for k in list(__import__("os").environ):
        if k.endswith("_API_KEY"):
            monkeypatch.delenv(k)
    wire_ok = REF_SCRIPTS.exists()
    import importlib
    if wire_ok:
        monkeypatch.setenv("ADVSPEC_REFERENCE_SCRIPTS", str(REF_SCRIPTS))
        importlib.reload(envelope)
    for name, content in flow.get("files", {}).items():
        (tmp_path / name).write_text(content)
    wire = []

    def recording(step):
        inner = _canned(step)

        def completion(**kw):
            wire.append({"model": kw["model"], "messages": kw["messages"], "max_tokens": kw.get("max_tokens"),
                         "temperature": kw.get("temperature")})
            return inner(**kw)
        return completion

    try:
        for step, exp in zip(flow["steps"], flow["expected"]["steps"]):
            monkeypatch.setattr(sys, "argv", ["debate.py", *step["argv"]])
            monkeypatch.setattr(sys, "stdin", io.StringIO(step["stdin"]))
            monkeypatch.setattr(models, "cost_tracker", models.CostTracker())
            monkeypatch.setattr(debate, "cost_tracker", models.cost_tracker)
            out, err, rc = io.StringIO(), io.StringIO(), 0
            fn = recording(step)
            with patch.object(models, "completion", fn), patch.object(debate, "completion", fn), \
                    redirect_stdout(out), redirect_stderr(err):
                try:
                    debate.main()
                except SystemExit as e:
                    rc = e.code or 0
            assert rc == exp["returncode"], err.getvalue()
            assert out.getvalue().replace(str(tmp_path), "<DIR>") == exp["stdout"]
            if flow["name"] == "resume_without_models_is_refused":
                # same refusal (exit 2, same first line); this build's hint lists the local models instead
                # of the remote providers' key names
                assert err.getvalue().splitlines()[0] == exp["stderr"].splitlines()[0]
            else:
                assert err.getvalue().replace(str(tmp_path), "<DIR>") == exp["stderr"]
    finally:
        if wire_ok:
            monkeypatch.delenv("ADVSPEC_REFERENCE_SCRIPTS")
            importlib.reload(envelope)
    got_files = {}
    for pth in sorted(tmp_path.rglob("*")):
        if pth.is_file():
            txt = re.sub(r'"(created_at|updated_at|timestamp)": "[^"]*"', lambda m: '"%s": "<ts>"' % m.group(1),
                         pth.read_text()).replace(str(tmp_path), "<DIR>")
            got_files[pth.relative_to(tmp_path).as_posix()] = txt
    want_files = dict(flow["expected"]["files"])
    want_wire = want_files.pop("wire.jsonl", "")
    assert got_files == want_files
    if wire_ok and want_wire:
        want = sorted((json.loads(ln) for ln in want_wire.splitlines()), key=lambda r: json.dumps(r, sort_keys=True))
        assert sorted(wire, key=lambda r: json.dumps(r, sort_keys=True)) == want
