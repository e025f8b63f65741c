"""What bf16 activation rounding alone costs at FULL width (test infrastructure, CPU only).

oracle/restate.py re-implements the forward from the engine's blob and can round activations to bf16 at
exactly the points the CUDA path does.  On the 2-layer Llama-3-8B-width model of tests/test_gpu_fullwidth.py
this predicts the engine-vs-HF error the GPU parity tests should see; the stated tolerance
(tests/gpu_util.py: max 0.08, rms 0.02 of the logit std) has to sit above it.  Output of one run is kept in
profiles/r02_restate_fullwidth_rounding_prediction.txt.
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import advspec_loader; advspec_loader.load()
from advspec_b200 import model_spec
from oracle import hf_oracle, restate
spec = model_spec.resolve("llama-3-8b-2layer")
m = hf_oracle.build_hf_model_fast(spec, 77)
blob = hf_oracle.export_blob(spec, m)
toks = np.random.default_rng(1).integers(0, spec.vocab_size, 96).tolist()
ref = hf_oracle.hf_logits(m, toks)
bm = restate.BlobModel(spec, blob, hf_oracle.rope_inv_freq(m))
ex = bm.forward_logits(toks, engine_rounding=False)
rd = bm.forward_logits(toks, engine_rounding=True)
sd = ref.std()
print("exact max/rms", np.abs(ex-ref).max()/sd, np.sqrt(((ex-ref)**2).mean())/sd)
print("rounded max/rms", np.abs(rd-ref).max()/sd, np.sqrt(((rd-ref)**2).mean())/sd)
print("rounded last-row max/rms", np.abs(rd[-1]-ref[-1]).max()/sd, np.sqrt(((rd[-1]-ref[-1])**2).mean())/sd)
