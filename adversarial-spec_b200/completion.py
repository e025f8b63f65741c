"""Seam B1: a drop-in for ``litellm.completion`` as the reference calls it
(skills/adversarial-spec/scripts/models.py:614-628; debate.py:715 for export-tasks).

Only the fields the reference reads exist on the response:
``.choices[0].message.content`` and ``.usage.prompt_tokens/.completion_tokens``
(models.py:629, 639-640).  The reference enters this function from N threads at
once with identical ``messages`` when opponents share a model (models.py:699);
those calls are coalesced here — the first caller waits until arrivals go quiet (5 ms
without a new sibling, 25 ms at most), then ONE prefill serves them all.  Concurrent calls for the same
model and system prompt whose user messages differ join the same round: the engine shares whatever token
prefix they have in common (``runtime.generate_group``).  Callers that can see the whole
panel should use ``models.call_models_parallel`` (seam B2) instead.
"""

from __future__ import annotations

import hashlib
import os
import threading
import time
from dataclasses import dataclass
from typing import Any, Optional

from . import runtime
from .model_spec import is_local_model, resolve


@dataclass
class Message:
    content: str
    role: str = "assistant"


@dataclass
class Choice:
    message: Message
    index: int = 0
    finish_reason: str = "stop"


@dataclass
class Usage:
    prompt_tokens: int
    completion_tokens: int

    @property
    def total_tokens(self) -> int:
        return self.prompt_tokens + self.completion_tokens


@dataclass
class CompletionResponse:
    choices: list
    usage: Optional[Usage]
    model: str


def _split_messages(messages: list[dict]) -> tuple[str, str]:
    system = "\n".join(m.get("content", "") for m in messages if m.get("role") == "system")
    user = "\n".join(m.get("content", "") for m in messages if m.get("role") != "system")
    return system, user


class _Batch:
    def __init__(self):
        self.n = 0
        self.users: list[str] = []
        self.closed = False
        self.results: Optional[list] = None
        self.done = threading.Event()
        self.arrived = threading.Condition()


_pending: dict[tuple, _Batch] = {}
_pending_mu = threading.Lock()
_call_counter = 0


def _env_ms(name: str, default: float) -> float:
    try:
        return float(os.environ.get(name, default)) / 1000.0
    except ValueError:
        return default / 1000.0


def _gather_siblings(b: "_Batch") -> None:
    """The leader's wait for the sibling calls of one fan-out.  The reference starts its N threads within
    microseconds of each other (models.py:699-719), so the leader does not sit out a fixed window: it
    returns as soon as no new caller has joined for ADVSPEC_COALESCE_QUIET_MS (default 5 ms), and never
    waits longer than ADVSPEC_COALESCE_MS (default 25 ms) in total."""
    deadline = time.perf_counter() + _env_ms("ADVSPEC_COALESCE_MS", 25.0)
    quiet = _env_ms("ADVSPEC_COALESCE_QUIET_MS", 5.0)
    seen, last_change = b.n, time.perf_counter()
    while True:
        now = time.perf_counter()
        if now >= deadline:
            return
        with b.arrived:
            b.arrived.wait(timeout=min(quiet, deadline - now))
            n = b.n
        now = time.perf_counter()
        if n != seen:
            seen, last_change = n, now
        elif now - last_change >= quiet:
            return


def completion(*, model: str, messages: list[dict], max_tokens: int = 8000, timeout: Any = None,
               temperature: float = 1.0, **_ignored) -> CompletionResponse:
    global _call_counter
    if not is_local_model(model):
        try:
            import litellm  # the reference's own route for remote providers
        except ImportError:
            raise RuntimeError(
                f"model {model!r} is not a local B200 model (use b200/<name>) and litellm is not installed")
        kw = dict(model=model, messages=messages, max_tokens=max_tokens, timeout=timeout)
        if temperature is not None:
            kw["temperature"] = temperature
        return litellm.completion(**kw)

    spec = resolve(model)
    system, user = _split_messages(messages)
    digest = hashlib.sha256(system.encode()).hexdigest()
    key = (spec.name, digest, max_tokens, float(temperature))
    with _pending_mu:
        b = _pending.get(key)
        leader = b is None or b.closed
        if leader:
            b = _Batch()
            _pending[key] = b
        my = b.n
        b.n += 1
        b.users.append(user)
        base = _call_counter
        _call_counter += 1
    if not leader:
        with b.arrived:
            b.arrived.notify_all()
    if leader:
        _gather_siblings(b)
        with _pending_mu:
            b.closed = True
            if _pending.get(key) is b:
                del _pending[key]
            n = b.n
            users = list(b.users)
        seeds = [runtime.opponent_seed(0, base * 16 + i) for i in range(n)]
        try:
            b.results = runtime.run_round([model] * n, system, user if len(set(users)) == 1 else users, seeds,
                                          max_tokens, temperature)
        except Exception as ex:
            b.results = [ex] * n
        b.done.set()
    else:
        if not b.done.wait(timeout if isinstance(timeout, (int, float)) and timeout else None):
            raise TimeoutError(f"local completion for {model} timed out after {timeout}s")
    r = b.results[my]
    if isinstance(r, Exception):
        raise r
    return CompletionResponse([Choice(Message(r.text))], Usage(r.prompt_tokens, r.completion_tokens), model)
