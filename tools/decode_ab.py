"""A/B of engine knobs in ONE process on ONE GPU: production decode ms/step (CUDA events over the
graph replay) and prefill ms, several repetitions each."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import advspec_loader; advspec_loader.load()
from advspec_b200 import engine as eng, model_spec

spec = model_spec.resolve(os.environ.get("TL_MODEL", "llama-3-8b"))
ptok, b = 5068, 3
prompt = np.random.default_rng(0).integers(0, spec.vocab_size, ptok).tolist()
variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else [{}]
reps = int(os.environ.get("AB_REPS", "3"))
for rnd in range(2):
    for v in variants:
        for k in ("ADVSPEC_X_SMEM_MAX", "ADVSPEC_ATTN_MIN_SPLIT", "ADVSPEC_GEMV_IMPL", "ADVSPEC_ATTN_IMPL",
                  "ADVSPEC_GEMM_NARROW", "ADVSPEC_GEMM_SPLITK", "ADVSPEC_GEMM_BAND_MB", "ADVSPEC_NO_PDL", "ADVSPEC_NO_GRAPH",
                  "ADVSPEC_ATTN_PREFILL_TC", "ADVSPEC_L2_EVICT_FIRST", "ADVSPEC_PREFILL_CHUNK"):
            os.environ.pop(k, None)
        os.environ.update({k: str(x) for k, x in v.items()})
        e = eng.Engine(spec, 0, 5120, 300, 8)
        e.init_weights_random(0, 0.02)
        pre, dec = [], []
        for r in range(reps + 1):
            pid = e.prefill(prompt); ids = e.fork(pid, [1, 2, 3][:b])
            e.decode(ids, 200, temperature=0.7)
            tm = e.timing()
            if r > 0:
                pre.append(tm.prefill_ms); dec.append(tm.decode_ms / tm.decode_steps)
        e.close()
        print(f"round {rnd} {json.dumps(v):60s} prefill ms {min(pre):7.2f} (med {sorted(pre)[len(pre)//2]:7.2f})   decode ms/step {min(dec):.4f} (med {sorted(dec)[len(dec)//2]:.4f})", flush=True)
