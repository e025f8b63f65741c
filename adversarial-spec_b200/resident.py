"""Resident round driver: many `debate.py` invocations over ONE set of loaded engines.

The reference drives a debate as one CLI process per round — `debate.py critique --session S`, then
`debate.py critique --resume S` until every opponent answers [AGREE]
(skills/adversarial-spec/scripts/debate.py:739-795 load_or_resume_session, :855-878 checkpoint + session
update; skills/adversarial-spec/scripts/session.py:16-82 on-disk formats).  With remote providers a fresh
process per round costs nothing; with a local engine every process would re-create the engine and
re-materialise 16-140 GB of weights.  This module keeps the engines resident and runs each invocation
exactly as the CLI would:

  run_cli(argv, stdin_text, cwd, home)   one invocation in this process: same argv, same stdin, stdout /
                                         stderr / exit code captured, session and checkpoint files written
                                         under the CALLER's home and working directory, cost totals
                                         starting from zero like a fresh process
  serve(socket_path)                     a Unix-socket server around run_cli
  request(socket_path, argv, stdin_text) the client side; `debate.py` forwards to it when ADVSPEC_SERVER
                                         names a socket (see debate.main)

Server:  python adversarial-spec_b200/resident.py serve --socket /tmp/advspec.sock
Client:  ADVSPEC_SERVER=/tmp/advspec.sock python adversarial-spec_b200/debate.py critique --resume S --json
"""

from __future__ import annotations

import contextlib
import io
import json
import os
import socket
import struct
import sys
import threading
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence

if __package__ in (None, ""):  # executed as a script: make the hyphenated package importable
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import debate as _debate, models as _models, runtime as _runtime, session as _session
else:
    from . import debate as _debate, models as _models, runtime as _runtime, session as _session

_cli_mu = threading.Lock()  # argv, stdio, cwd and the session directories are process-global
FORWARDED_ENV = ("ADVSPEC_MAX_NEW_TOKENS", "ADVSPEC_SEED", "ADVSPEC_PLACEMENT", "OPENAI_API_KEY", "ANTHROPIC_API_KEY",
                 "GEMINI_API_KEY", "XAI_API_KEY", "MISTRAL_API_KEY", "GROQ_API_KEY", "DEEPSEEK_API_KEY",
                 "ZHIPUAI_API_KEY", "OPENROUTER_API_KEY", "OPENAI_API_BASE")


@dataclass
class CliResult:
    code: int
    stdout: str
    stderr: str
    wall_s: float = 0.0
    engines_resident: int = 0  # engines alive in this process after the call (1 per (model, GPU))
    engines_created: int = 0   # engines this call had to create (0 = every engine was already resident)


def _reset_cost_tracker() -> None:
    ct = _models.cost_tracker
    ct.total_input_tokens = ct.total_output_tokens = 0
    ct.total_cost = 0.0
    ct.by_model.clear()


def run_cli(argv: Sequence[str], stdin_text: str = "", cwd: Optional[str] = None, home: Optional[str] = None,
            env: Optional[dict] = None) -> CliResult:
    """One `debate.py <argv...>` invocation inside this process (engines stay loaded between calls)."""
    with _cli_mu:
        saved = (sys.argv, sys.stdin, os.getcwd(), _session.SESSIONS_DIR, _session.CHECKPOINTS_DIR)
        saved_env = {k: os.environ.get(k) for k in (env or {})}
        out, err = io.StringIO(), io.StringIO()
        created0 = _runtime.POOL.created
        t0 = time.perf_counter()
        code = 0
        try:
            if cwd:
                os.chdir(cwd)
            base_home = Path(home) if home else Path.home()
            _session.SESSIONS_DIR = base_home / ".config" / "adversarial-spec" / "sessions"
            _session.CHECKPOINTS_DIR = Path.cwd() / ".adversarial-spec-checkpoints"
            for k, v in (env or {}).items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            _reset_cost_tracker()
            sys.argv = ["debate.py", *argv]
            sys.stdin = io.StringIO(stdin_text)
            with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
                try:
                    _debate.main(_forward=False)
                except SystemExit as e:
                    code = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
                    if e.code is not None and not isinstance(e.code, int):
                        print(e.code, file=sys.stderr)
        finally:
            sys.argv, sys.stdin = saved[0], saved[1]
            os.chdir(saved[2])
            _session.SESSIONS_DIR, _session.CHECKPOINTS_DIR = saved[3], saved[4]
            for k, v in saved_env.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        return CliResult(code, out.getvalue(), err.getvalue(), time.perf_counter() - t0,
                         _runtime.POOL.resident_count(), _runtime.POOL.created - created0)


# ----------------------------------------------------------------------------- socket protocol
# One request per connection: 4-byte big-endian length + UTF-8 JSON, both ways.
def _send(sock: socket.socket, obj: dict) -> None:
    raw = json.dumps(obj).encode()
    sock.sendall(struct.pack(">I", len(raw)) + raw)


def _recv(sock: socket.socket) -> dict:
    def exactly(n: int) -> bytes:
        buf = b""
        while len(buf) < n:
            part = sock.recv(n - len(buf))
            if not part:
                raise ConnectionError("peer closed the connection mid-message")
            buf += part
        return buf

    (n,) = struct.unpack(">I", exactly(4))
    return json.loads(exactly(n).decode())


def serve(socket_path: str, ready: Optional[threading.Event] = None, stop: Optional[threading.Event] = None) -> None:
    """Serve CLI invocations until `stop` is set or a {"op": "shutdown"} request arrives."""
    path = Path(socket_path)
    if path.exists():
        path.unlink()
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(str(path))
    srv.listen(8)
    srv.settimeout(0.2)
    if ready is not None:
        ready.set()
    try:
        while not (stop is not None and stop.is_set()):
            try:
                conn, _ = srv.accept()
            except socket.timeout:
                continue
            with conn:
                try:
                    req = _recv(conn)
                    if req.get("op") == "shutdown":
                        _send(conn, {"ok": True})
                        break
                    if req.get("op") == "stats":
                        _send(conn, {"engines_resident": _runtime.POOL.resident_count(),
                                     "engines_created": _runtime.POOL.created})
                        continue
                    r = run_cli(req.get("argv", []), req.get("stdin", ""), req.get("cwd"), req.get("home"),
                                req.get("env"))
                    _send(conn, {"code": r.code, "stdout": r.stdout, "stderr": r.stderr, "wall_s": r.wall_s,
                                 "engines_resident": r.engines_resident, "engines_created": r.engines_created})
                except Exception as ex:  # a bad request must not take the resident engines down
                    with contextlib.suppress(Exception):
                        _send(conn, {"code": 70, "stdout": "", "stderr": f"advspec server error: {ex}\n"})
    finally:
        srv.close()
        with contextlib.suppress(FileNotFoundError):
            path.unlink()
        _runtime.POOL.close()


def request(socket_path: str, argv: Sequence[str], stdin_text: str = "", timeout: Optional[float] = None,
            op: Optional[str] = None) -> dict:
    """Client side: run `debate.py <argv>` in the server process; the caller's cwd, HOME and the
    benchmark / credential environment travel with the request."""
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
        s.settimeout(timeout)
        s.connect(socket_path)
        if op:
            _send(s, {"op": op})
        else:
            _send(s, {"argv": list(argv), "stdin": stdin_text, "cwd": os.getcwd(), "home": str(Path.home()),
                      "env": {k: os.environ.get(k) for k in FORWARDED_ENV if k in os.environ}})
        return _recv(s)


def main() -> None:
    import argparse

    ap = argparse.ArgumentParser(description="resident adversarial-spec engine server")
    ap.add_argument("action", choices=["serve", "shutdown", "stats"])
    ap.add_argument("--socket", default=os.environ.get("ADVSPEC_SERVER", "/tmp/advspec_b200.sock"))
    a = ap.parse_args()
    if a.action == "serve":
        print(f"advspec server listening on {a.socket}", file=sys.stderr, flush=True)
        serve(a.socket)
    else:
        print(json.dumps(request(a.socket, [], op=a.action)))


if __name__ == "__main__":
    main()
