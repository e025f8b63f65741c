"""CPU restatement of the reference's fan-out, used as the benchmark's CPU baseline.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).  Restates
``call_models_parallel`` (skills/adversarial-spec/scripts/models.py:681-722: one
thread per opponent, gather with as_completed) and the request half of
``call_single_model`` (:482-503, :614-628: system + user message, max_tokens
8000, temperature 0.7) over the CPU ``completion`` of oracle/litellm_shim.  When
/root/reference is present (this container only) tests/test_reference_pin.py
runs the reference's real ``models.call_models_parallel`` against the same shim
and checks this restatement returns the same token counts.
"""

from __future__ import annotations

import concurrent.futures
import sys
from pathlib import Path

_SHIM = Path(__file__).resolve().parent / "litellm_shim"


def _shim():
    if str(_SHIM) not in sys.path:
        sys.path.insert(0, str(_SHIM))
    import litellm  # the shim, unless a real litellm is installed ahead of it

    if not hasattr(litellm, "LAST_TIMINGS"):
        raise RuntimeError("a real litellm shadows oracle/litellm_shim; the CPU baseline needs the shim")
    return litellm


def cpu_call_models_parallel(models: list[str], system_prompt: str, user_message: str, timeout: int = 600):
    """Returns [(model, content, prompt_tokens, completion_tokens)] in completion order."""
    litellm = _shim()

    def one(model: str):
        r = litellm.completion(model=model,
                               messages=[{"role": "system", "content": system_prompt},
                                         {"role": "user", "content": user_message}],
                               max_tokens=8000, timeout=timeout, temperature=0.7)
        return (model, r.choices[0].message.content, r.usage.prompt_tokens, r.usage.completion_tokens)

    out = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(models)) as ex:
        futs = [ex.submit(one, m) for m in models]
        for f in concurrent.futures.as_completed(futs):
            out.append(f.result())
    return out


def take_timings() -> list:
    litellm = _shim()
    t = list(litellm.LAST_TIMINGS)
    litellm.LAST_TIMINGS.clear()
    return t
