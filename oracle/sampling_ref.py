"""numpy restatement of the engine's seeded Gumbel-max sampler (csrc/common.cuh
``mix64`` / ``uniform01`` and csrc/decode_kernels.cuh ``sample_kernel``).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference samples remotely at
temperature 0.7 (skills/adversarial-spec/scripts/models.py:626) and exposes no
seed, so there is nothing in the reference to pin the sampler against; the hash
is pinned bit-exactly by tests/test_sampling_ref.py against constants computed
independently in pure-Python integers.
"""

from __future__ import annotations

import numpy as np

_M = (1 << 64) - 1


def mix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


def _mix64_np(z: np.ndarray) -> np.ndarray:
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, step: int, vocab: int) -> np.ndarray:
    """u[v] in (0,1) for v < vocab, float32, identical bits to the device function."""
    with np.errstate(over="ignore"):
        idx = np.arange(vocab, dtype=np.uint64)
        inner = _mix64_np((np.uint64(step) << np.uint64(32)) | idx)
        h = _mix64_np(np.uint64(seed) ^ inner)
    m = (h >> np.uint64(40)).astype(np.float32)
    return (m + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def sample(logits: np.ndarray, temperature: float, seed: int, step: int) -> tuple[int, float]:
    """Returns (token, gap) where gap is the winner's margin over the runner-up in the
    perturbed score — a device/host ulp difference can only flip the result if gap ~ 0."""
    lg = np.asarray(logits, dtype=np.float64)
    if temperature <= 0:
        sc = lg
    else:
        u = uniform01(seed, step, lg.size).astype(np.float64)
        sc = lg / float(np.float32(temperature)) - np.log(-np.log(u))
    order = np.argsort(-sc, kind="stable")
    gap = float(sc[order[0]] - sc[order[1]]) if lg.size > 1 else float("inf")
    return int(order[0]), gap
