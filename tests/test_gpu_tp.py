"""Tensor parallelism on real GPUs: needs >= 2 B200s on the box, so it is
skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_tp.py -m gpu`."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _n_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_tensor_parallel_matches_hf_and_the_unsplit_engine(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", TP_MODELS="tiny-gqa4,tiny-llama-128,llama-3-70b-2layer-v32k")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           str(ROOT / "tests" / "tp_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("TP_RESULT ")]
    assert p.returncode == 0 and line, (p.stdout[-2000:], p.stderr[-4000:])
    res = json.loads(line[-1][len("TP_RESULT "):])
    # the last one has Llama-3-70B's layer shape (d 8192, 64/8 heads, MLP 28672) at 2 layers: a rank's share at
    # TP=2 is 32 query / 4 KV heads and 14,336 MLP columns
    for name in ("tiny-gqa4", "tiny-llama-128", "llama-3-70b-2layer-v32k"):
        r = res[name]
        # same tolerance as the single-GPU parity tests (tests/test_gpu_engine.py): bf16 activations
        assert r["prefill_max_over_std"] < 0.08 and r["prefill_rms_over_std"] < 0.02, r
        assert r["decode_max_over_std"] < 0.08, r
        assert r["ranks_identical"], r
        # the all-reduce changes the summation order, so a near-tie may flip a token; most must agree
        assert r["greedy_token_agreement"] >= 0.75 and r["sampled_token_agreement"] >= 0.5, r
