"""Debug aid: one prefill -> fork -> decode_step -> decode on a tiny model, verbose."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import advspec_loader; advspec_loader.load()
from advspec_b200 import engine as eng, model_spec

name = sys.argv[1] if len(sys.argv) > 1 else "tiny-llama"
spec = model_spec.resolve(name)
e = eng.Engine(spec, 0, 512, 64, 8)
e.init_weights_random(1, 0.05)
prompt = np.random.default_rng(0).integers(0, spec.vocab_size, 150).tolist()
t = time.time(); pid = e.prefill(prompt); print("prefill ok", time.time() - t, flush=True)
ids = e.fork(pid, [1, 2, 3]); print("fork ok", ids, flush=True)
e.decode_step(ids, [5, 6, 7]); print("decode_step ok", flush=True)
lg = e.get_logits(3); print("logits", lg.shape, float(np.abs(lg).max()), flush=True)
r = e.decode(ids, 8, temperature=0.7); print("decode ok", r.tokens, flush=True)
tm = e.timing(); print("timing", tm.prefill_ms, tm.decode_ms, tm.decode_steps, tm.kernel_launches)
