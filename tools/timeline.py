"""In-situ decode timeline on the bench workload: per-kernel cost inside the graph replay."""
import json, sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import advspec_loader; advspec_loader.load()
from advspec_b200 import engine as eng, model_spec

name = os.environ.get("TL_MODEL", "llama-3-8b")
b = int(os.environ.get("TL_B", "3"))
ptok = int(os.environ.get("TL_PROMPT", "5068"))
spec = model_spec.resolve(name)
e = eng.Engine(spec, 0, (ptok + 255) // 256 * 256, 320, 8)
e.init_weights_random(0, 0.02)
prompt = np.random.default_rng(0).integers(0, spec.vocab_size, ptok).tolist()
pid = e.prefill(prompt); ids = e.fork(pid, list(range(1, b + 1)))
e.decode(ids, 64, temperature=0.7)           # warm + graph build
tm0 = e.timing()
e.ktrace_enable(True)
e.decode(ids, 40, temperature=0.7)
tr = e.ktrace_read()
e.ktrace_enable(False)
tm = e.timing()
step_bytes, gemv_bytes = e.decode_step_bytes(ids)
ts = np.array([t for t, k in tr], dtype=np.int64); kinds = np.array([k for t, k in tr])
# steps are delimited by the sampler (kind 3)
samp = np.where(kinds == 3)[0]
per_step = []
for a, z in zip(samp[1:-1], samp[2:]):
    d = np.diff(ts[a:z + 1])            # cost of kernel i = start(i+1) - start(i)
    per_step.append((kinds[a:z], d))
L = spec.n_layers
kpl = (len(per_step[0][0]) - 2) // L    # kernels per layer
names = {1: "gemv", 2: "attn", 3: "sample", 4: "rope", 5: "combine"}
agg = {}
for ks, d in per_step:
    # layout: sample, then L x [kpl kernels], then lm_head gemv
    for i, (k, dt) in enumerate(zip(ks, d)):
        if i == 0: key = "sample+embed"
        elif i == len(ks) - 1: key = "gemv lm_head"
        else:
            pos = (i - 1) % kpl
            key = f"L[{pos}] {names[int(k)]}"
        agg.setdefault(key, []).append(float(dt))
tot = sum(np.sum(v) for v in agg.values()) / len(per_step)
out = {"model": name, "b": b, "prompt": ptok, "decode_ms_per_step_events": tm.decode_ms / max(tm.decode_steps, 1),
       "timeline_us_per_step": tot / 1e3, "kernels_per_step": len(per_step[0][0]), "steps_seen": len(per_step),
       "gemv_impl": os.environ.get("ADVSPEC_GEMV_IMPL", "3"), "rows": {}}
gemv_ns = 0.0
for k, v in agg.items():
    per = np.sum(v) / len(per_step)
    n = len(v) / len(per_step)
    out["rows"][k] = {"n_per_step": n, "us_each": per / n / 1e3, "us_per_step": per / 1e3, "share": per / tot}
    if "gemv" in k: gemv_ns += per
out["gemv_us_per_step"] = gemv_ns / 1e3
out["gemv_gbs"] = gemv_bytes / gemv_ns
out["step_gbs"] = step_bytes / tot
print(json.dumps(out, indent=1))
