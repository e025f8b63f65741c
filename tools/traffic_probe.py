"""Workload for bench.py's ncu pass: one prefill and two decode steps of the bench configuration, kernels
launched one by one (ADVSPEC_NO_GRAPH=1) so `ncu -k regex:gemv_mma_kernel -s 1 -c <4L+1>` sees exactly the
weight-streaming GEMV launches of ONE decode step (launch 0 is the prefill's own lm_head GEMV)."""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["ADVSPEC_NO_GRAPH"] = "1"
import numpy as np  # noqa: E402

import advspec_loader  # noqa: E402

advspec_loader.load()
from advspec_b200 import engine as eng, model_spec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama-3-8b")
ap.add_argument("--opponents", type=int, default=3)
ap.add_argument("--spec-tokens", type=int, default=4096)
a = ap.parse_args()
spec = model_spec.resolve(a.model)
n = a.spec_tokens + 972  # the PRD envelope of bench.py's workload
e = eng.Engine(spec, 0, (n + 255) // 256 * 256, 16, 8)
e.init_weights_random(0, 0.02)
pid = e.prefill(np.random.default_rng(0).integers(0, spec.vocab_size, n).tolist())
ids = e.fork(pid, list(range(1, a.opponents + 1)))
e.decode(ids, 3, temperature=0.7)
print("TRAFFIC_PROBE " + json.dumps({"gemv_launches_per_step": 4 * spec.n_layers + 1, "prompt_tokens": n,
                                     "opponents": a.opponents}), flush=True)
e.close()
