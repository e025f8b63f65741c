"""The numeric oracle is pinned: transformers reproduces the committed golden logits, and the
from-scratch restatement (oracle/restate.py, reading the ENGINE's weight blob) agrees with it."""

import hashlib
from pathlib import Path

import numpy as np
import pytest

from advspec_b200 import model_spec
from oracle import hf_oracle, restate, sampling_ref

GOLDEN = Path(__file__).parent / "golden"
NAMES = ["tiny-llama", "tiny-llama-128", "tiny-qwen2", "tiny-gemma", "tiny-phi3", "tiny-gemma256", "tiny-mistral"]


@pytest.fixture(scope="module", params=NAMES)
def case(request):
    name = request.param
    g = np.load(GOLDEN / f"hf_logits_{name}.npz")
    spec = model_spec.resolve(name)
    model = hf_oracle.build_hf_model(spec, int(g["seed"]))
    return name, spec, model, g


def test_hf_reproduces_golden_logits(case):
    name, spec, model, g = case
    got = hf_oracle.hf_logits(model, g["tokens"].tolist())
    assert np.abs(got - g["logits"]).max() < 1e-4 * g["logits"].std()


def test_blob_is_bit_stable(case):
    name, spec, model, g = case
    blob = hf_oracle.export_blob(spec, model)
    assert hashlib.sha256(blob.tobytes()).digest() == g["blob_sha256"].tobytes()
    assert np.array_equal(hf_oracle.rope_inv_freq(model), g["inv_freq"])


def test_restatement_matches_hf_and_bounds_engine_rounding(case):
    name, spec, model, g = case
    bm = restate.BlobModel(spec, hf_oracle.export_blob(spec, model), g["inv_freq"])
    toks, ref = g["tokens"].tolist(), g["logits"]
    exact = bm.forward_logits(toks, engine_rounding=False)
    assert np.abs(exact - ref).max() < 1e-4 * ref.std()
    rounded = bm.forward_logits(toks, engine_rounding=True)
    mx = np.abs(rounded - ref).max() / ref.std()
    rms = np.sqrt(((rounded - ref) ** 2).mean()) / ref.std()
    # the GPU parity tests state TOL_MAX=0.08 / TOL_RMS=0.02: bf16 activation rounding alone stays well inside
    assert mx < 0.04 and rms < 0.01, (mx, rms)


def test_sampler_hash_is_pinned():
    # pure-Python big-int evaluation vs the vectorised numpy path
    for seed, step in [(0, 0), (12345, 7), (2**63 - 1, 4_000_000)]:
        u = sampling_ref.uniform01(seed, step, 64)
        for v in (0, 1, 63):
            h = sampling_ref.mix64(seed ^ sampling_ref.mix64((step << 32) | v))
            want = np.float32((np.float32(h >> 40) + np.float32(0.5)) * np.float32(1.0 / 16777216.0))
            assert u[v] == want
    assert sampling_ref.mix64(0) == 0xE220A8397B1DCDAF  # splitmix64's first output for state 0
    lg = np.array([0.1, 5.0, 0.2, 4.9], dtype=np.float32)
    assert sampling_ref.sample(lg, 0.0, 1, 0)[0] == 1
    counts = np.bincount([sampling_ref.sample(lg, 0.7, s, 0)[0] for s in range(400)], minlength=4)
    p = np.exp(lg / 0.7) / np.exp(lg / 0.7).sum()
    assert np.abs(counts / 400 - p).max() < 0.08
