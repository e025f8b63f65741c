"""Shared helpers for the GPU parity tests: engine + HF oracle on the same weights."""

from __future__ import annotations

import functools

import numpy as np

from advspec_b200 import engine as eng
from advspec_b200 import model_spec
from oracle import hf_oracle


@functools.lru_cache(maxsize=8)
def oracle_and_blob(name: str, seed: int):
    spec = model_spec.resolve(name)
    model = hf_oracle.build_hf_model(spec, seed)
    blob = hf_oracle.export_blob(spec, model)
    return spec, model, blob, hf_oracle.rope_inv_freq(model)


def make_engine(name: str, seed: int, max_prefix: int = 1024, max_new: int = 64, max_seqs: int = 8):
    spec, model, blob, inv = oracle_and_blob(name, seed)
    e = eng.Engine(spec, 0, max_prefix, max_new, max_seqs)
    e.load_weights(blob)
    e.set_rope_inv_freq(inv)
    return spec, model, e


def rel_errors(got: np.ndarray, ref: np.ndarray) -> tuple[float, float]:
    sd = float(ref.std()) + 1e-12
    d = got.astype(np.float64) - ref.astype(np.float64)
    return float(np.abs(d).max() / sd), float(np.sqrt((d * d).mean()) / sd)


# Stated tolerance for engine-vs-HF logits (bf16 activations between kernels, fp32
# accumulation and residual stream): oracle/restate.py with engine_rounding=True
# predicts max 0.017-0.025 and rms 0.002-0.005 of the logit std on these shapes.
TOL_MAX = 0.08
TOL_RMS = 0.02
