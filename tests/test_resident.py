"""The resident round driver (adversarial-spec_b200/resident.py): running the CLI inside a long-lived
process — directly or through the Unix-socket server — leaves exactly what the UNMODIFIED reference CLI
leaves (stdout, stderr, exit code, session JSON, checkpoints: tests/golden/reference_cli_flows.json)."""

import io
import json
import os
import re
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace
from unittest.mock import patch

import pytest

from advspec_b200 import debate, models, resident, runtime

GOLDEN = Path(__file__).parent / "golden"
FLOWS = json.loads((GOLDEN / "reference_cli_flows.json").read_text())


def _canned(step):
    def completion(**kw):
        spec = step["responses"][kw["model"]]
        if spec.get("raise"):
            raise RuntimeError(spec["raise"])
        usage = None if spec.get("no_usage") else SimpleNamespace(prompt_tokens=spec["in"],
                                                                  completion_tokens=spec["out"])
        return SimpleNamespace(choices=[SimpleNamespace(message=SimpleNamespace(content=spec["content"]))],
                               usage=usage)
    return completion


def _files(root: Path) -> dict:
    got = {}
    for pth in sorted(root.rglob("*")):
        if pth.is_file() and pth.suffix != ".sock":
            txt = re.sub(r'"(created_at|updated_at|timestamp)": "[^"]*"', lambda m: '"%s": "<ts>"' % m.group(1),
                         pth.read_text()).replace(str(root), "<DIR>")
            got[pth.relative_to(root).as_posix()] = txt
    return got


@pytest.mark.parametrize("via", ["run_cli", "socket"])
@pytest.mark.parametrize("flow", FLOWS, ids=[f["name"] for f in FLOWS])
def test_resident_driver_reproduces_the_reference_flows(flow, via, monkeypatch, tmp_path):
    for k in list(os.environ):
        if k.endswith("_API_KEY"):
            monkeypatch.delenv(k)
    monkeypatch.setattr(models.time, "sleep", lambda s: None)
    work = tmp_path / "work"
    work.mkdir()
    for name, content in flow.get("files", {}).items():
        (work / name).write_text(content)
    cwd0 = os.getcwd()
    sock = str(tmp_path / "s.sock")
    ready, stop = threading.Event(), threading.Event()
    th = None
    if via == "socket":
        th = threading.Thread(target=resident.serve, args=(sock, ready, stop), daemon=True)
        th.start()
        assert ready.wait(5)
    try:
        for step, exp in zip(flow["steps"], flow["expected"]["steps"]):
            fn = _canned(step)
            with patch.object(models, "completion", fn), patch.object(debate, "completion", fn):
                if via == "run_cli":
                    r = resident.run_cli(step["argv"], step["stdin"], cwd=str(work), home=str(work))
                    code, out, err = r.code, r.stdout, r.stderr
                else:
                    monkeypatch.chdir(work)
                    monkeypatch.setenv("HOME", str(work))
                    d = resident.request(sock, step["argv"], step["stdin"], timeout=30)
                    monkeypatch.chdir(cwd0)
                    code, out, err = d["code"], d["stdout"], d["stderr"]
            assert code == exp["returncode"], err
            assert out.replace(str(work), "<DIR>") == exp["stdout"]
            if flow["name"] == "resume_without_models_is_refused":
                assert err.splitlines()[0] == exp["stderr"].splitlines()[0]
            else:
                assert err.replace(str(work), "<DIR>") == exp["stderr"]
    finally:
        if th is not None:
            stop.set()
            th.join(5)
    assert os.getcwd() == cwd0, "run_cli must restore the server's working directory"
    want = dict(flow["expected"]["files"])
    want.pop("wire.jsonl", None)
    assert _files(work) == want


def test_cli_forwards_to_the_server_named_by_the_environment(monkeypatch, tmp_path, capsys):
    """ADVSPEC_SERVER=<socket>: `debate.main()` sends argv + stdin to the server and relays its answer; an
    unreachable server is an error (exit 2), never a silent local run."""
    sock = str(tmp_path / "s.sock")
    ready, stop = threading.Event(), threading.Event()
    th = threading.Thread(target=resident.serve, args=(sock, ready, stop), daemon=True)
    th.start()
    assert ready.wait(5)
    seen = {}

    def fake_round(model_names, system_prompt, user_message, seeds, max_tokens, temperature, devices=None):
        seen["thread"] = threading.current_thread().name
        return [runtime.Generation("[AGREE]\n[SPEC]ok[/SPEC]", 10, 3, [1, 2, 3]) for _ in model_names]

    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("HOME", str(tmp_path))
    monkeypatch.setenv("ADVSPEC_SERVER", sock)
    monkeypatch.setattr(sys, "argv", ["debate.py", "critique", "--models", "b200/tiny-llama,b200/tiny-llama", "--json"])
    monkeypatch.setattr(sys, "stdin", io.StringIO("a spec"))
    try:
        with patch.object(runtime, "run_round", fake_round), pytest.raises(SystemExit) as ex:
            debate.main()
    finally:
        stop.set()
        th.join(5)
    assert ex.value.code == 0
    out = json.loads(capsys.readouterr().out)
    assert out["all_agreed"] is True and [r["output_tokens"] for r in out["results"]] == [3, 3]
    assert seen["thread"] != threading.main_thread().name, "the round must have run inside the server"
    monkeypatch.setenv("ADVSPEC_SERVER", str(tmp_path / "nobody.sock"))
    monkeypatch.setattr(sys, "stdin", io.StringIO("a spec"))
    with pytest.raises(SystemExit) as ex:
        debate.main()
    assert ex.value.code == 2 and "no server answers" in capsys.readouterr().err
