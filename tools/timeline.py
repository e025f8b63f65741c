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
ph = e.ktrace_phases()
e.ktrace_enable(False)
tm = e.timing()
step_bytes, gemv_bytes = e.decode_step_bytes(ids)
from advspec_b200 import measure
tl = measure.summarize(tr, spec.n_layers)
out = {"model": name, "b": b, "prompt": ptok, "decode_ms_per_step_events": tm.decode_ms / max(tm.decode_steps, 1),
       "timeline_us_per_step": tl["us_per_step"], "kernels_per_step": tl["kernels_per_step"], "steps_seen": tl["steps"],
       "gemv_impl": os.environ.get("ADVSPEC_GEMV_IMPL", "3"), "rows": tl["rows"],
       "gemv_us_per_step": tl["gemv_us_per_step"], "gemv_gbs": gemv_bytes / (tl["gemv_us_per_step"] * 1e3),
       "step_gbs": step_bytes / (tl["us_per_step"] * 1e3)}
out["attn_phases_ns"] = [ph[i + 1] - ph[i] for i in range(5)]
if os.environ.get("TL_PHASES"):
    out["phases_rel_ns"] = [int(ph[i]) - int(ph[0]) for i in range(16)]
print(json.dumps(out, indent=1))
