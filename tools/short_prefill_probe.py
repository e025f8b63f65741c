"""How long does a SHORT prompt chunk take, and is it the GPU or the host?  Event-timed prefill_ms (which
includes any gap the host leaves between launches) for a 200-token prompt, a 200-token extension of a
5,068-token prefix and three 200-token per-opponent tails; run it again under
`ncu --metrics gpu__time_duration.sum` and sum the kernel durations of one chunk to get the GPU-busy part."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import advspec_loader; advspec_loader.load()
from advspec_b200 import engine as eng, model_spec

spec = model_spec.resolve(os.environ.get("TL_MODEL", "llama-3-8b"))
rng = np.random.default_rng(0)
long_p = rng.integers(0, spec.vocab_size, 5068).tolist()
short = rng.integers(0, spec.vocab_size, 200).tolist()
e = eng.Engine(spec, 0, 5632, 512, 8)
e.init_weights_random(0, 0.02)
reps = int(os.environ.get("PROBE_REPS", "3"))
for r in range(reps):
    t0 = time.perf_counter(); e.prefill(short); w = (time.perf_counter() - t0) * 1e3
    print(f"prefill 200 tokens: {e.timing().prefill_ms:.2f} ms on the stream, {w:.2f} ms wall", flush=True)
if os.environ.get("PROBE_SHORT_ONLY"):
    e.close(); sys.exit(0)
for r in range(reps):
    pid = e.prefill(long_p); base = e.timing().prefill_ms
    t0 = time.perf_counter(); pid = e.prefill_extend(pid, 5068, short); w = (time.perf_counter() - t0) * 1e3
    print(f"extend a 5,068-token prefix by 200: {e.timing().prefill_ms:.2f} ms on the stream, {w:.2f} ms wall "
          f"(full prefill {base:.2f})", flush=True)
for r in range(reps):
    pid = e.prefill(long_p); base = e.timing().prefill_ms
    ids = e.fork(pid, [1, 2, 3])
    t0 = time.perf_counter()
    for s in ids:
        e.append_tail(s, short)
    w = (time.perf_counter() - t0) * 1e3
    print(f"three 200-token tails: {e.timing().prefill_ms - base:.2f} ms on the stream, {w:.2f} ms wall", flush=True)
e.close()
