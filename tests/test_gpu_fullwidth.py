"""Parity against the HF oracle at the shapes the benchmark runs.

Every model here has the FULL width of its family — d_model, heads, head_dim, MLP width and vocabulary
of SURVEY.md §8(d)'s table — and two layers, so the HF CPU forward (the oracle north_star prescribes)
finishes in seconds on the GPU box's host cores while every kernel runs at its production configuration:
the tcgen05 GEMMs at N = 4096 / 6144 / 28672 / 128256, the 8,192-token prefill chunk boundary, the
decode GEMV's ring at K = 4096 / 14336, the decode attention's 15-way prefix split with a 6-stage TMA
ring, the sampler's 128K-256K vocabulary scan.  Tolerance: tests/gpu_util.py (stated there).
"""

import copy

import numpy as np
import pytest
import torch

from advspec_b200 import engine as eng
from advspec_b200 import model_spec
from oracle import hf_oracle, sampling_ref
from tests.gpu_util import TOL_MAX, TOL_RMS, rel_errors

pytestmark = pytest.mark.gpu


def _family_spec(name: str):
    base = model_spec.resolve(name)
    return model_spec.with_layers(base, 2, f"{base.name}-2layer")


@pytest.fixture(scope="module")
def built():
    """One full-width oracle + blob at a time (6 GB of fp32 each): the previous one is dropped first."""
    state = {}

    def get(name: str, seed: int):
        if state.get("key") != (name, seed):
            state.clear()
            spec = _family_spec(name)
            model = hf_oracle.build_hf_model_fast(spec, seed)
            state.update(key=(name, seed), spec=spec, model=model, blob=hf_oracle.export_blob(spec, model),
                         inv=hf_oracle.rope_inv_freq(model))
        return state["spec"], state["model"], state["blob"], state["inv"]

    yield get
    state.clear()


def _engine(spec, blob, inv, max_prefix, max_new=32, max_seqs=8):
    e = eng.Engine(spec, 0, max_prefix, max_new, max_seqs)
    e.load_weights(blob)
    e.set_rope_inv_freq(inv)
    return e


def _tokens(spec, n, seed):
    return np.random.default_rng(seed).integers(0, spec.vocab_size, n).tolist()


@torch.no_grad()
def _hf_prefix_then_continuations(model, prompt, conts):
    """HF logits for prompt+cont_i at the positions of the continuation tokens: the prompt runs ONCE with
    a KV cache (fp32, transformers' own DynamicCache), each continuation then runs over a copy of it.
    Returns (last-prompt-position logits [V], [len(conts)][len(cont)][V])."""
    model.config._attn_implementation = "sdpa"
    try:
        r = model(input_ids=torch.tensor([list(prompt)]), use_cache=True, logits_to_keep=1)
        last = r.logits[0, -1].float().numpy()
        outs = []
        for c in conts:
            past = copy.deepcopy(r.past_key_values)
            rc = model(input_ids=torch.tensor([list(c)]), past_key_values=past, use_cache=True)
            outs.append(rc.logits[0].float().numpy())
    finally:
        model.config._attn_implementation = "eager"
    return last, outs


def _teacher_forced(e, model, spec, prompt, conts, diag, tag):
    ref_last, ref = _hf_prefix_then_continuations(model, prompt, conts)
    pid = e.prefill(prompt)
    got_last = e.get_logits(1)[0]
    mx, rms = rel_errors(got_last, ref_last)
    diag[f"fullwidth/{tag}/prefill_last"] = {"max": mx, "rms": rms, "prefill_ms": e.timing().prefill_ms}
    assert mx < TOL_MAX and rms < TOL_RMS, (tag, "prefill", mx, rms)
    b = len(conts)
    ids = e.fork(pid, list(range(1, b + 1)))
    worst = (0.0, 0.0)
    for t in range(len(conts[0])):
        e.decode_step(ids, [c[t] for c in conts])
        lg = e.get_logits(b)
        for i in range(b):
            m2, r2 = rel_errors(lg[i], ref[i][t])
            worst = (max(worst[0], m2), max(worst[1], r2))
    diag[f"fullwidth/{tag}/decode_b{b}"] = {"max": worst[0], "rms": worst[1], "steps": len(conts[0])}
    assert worst[0] < TOL_MAX and worst[1] < TOL_RMS, (tag, "decode", worst)
    e.release_prefix(pid)
    return got_last


def test_llama3_8b_width_bench_prompt_b3_and_b8(cuda_device, diag, built):
    """The benchmark's own configuration, two layers deep: 5,068 prompt tokens, 3 opponents (one 12-row MMA
    group per KV head, 15 prefix splits), then the full batch of 8 (two groups), 8 teacher-forced steps each
    with different continuations per opponent; the first sampled token follows the numpy sampler."""
    spec, model, blob, inv = built("llama-3-8b", 77)
    e = _engine(spec, blob, inv, 5120 + 4096)
    prompt = _tokens(spec, 5068, 1)
    last = _teacher_forced(e, model, spec, prompt, [_tokens(spec, 8, 10 + i) for i in range(3)], diag, "llama-3-8b/p5068")
    _teacher_forced(e, model, spec, prompt, [_tokens(spec, 8, 30 + i) for i in range(8)], diag, "llama-3-8b/p5068")
    # sampling at the real vocabulary size: token 0 of each opponent from the shared prefill logits
    pid = e.prefill(prompt)
    seeds = [101, 202, 303]
    res = e.decode(e.fork(pid, seeds), 4, temperature=0.7)
    for i, s in enumerate(seeds):
        want, gap = sampling_ref.sample(last, 0.7, s, 0)
        assert res.tokens[i][0] == want or gap < 1e-3, (i, res.tokens[i][0], want, gap)
    e.close()


@pytest.mark.parametrize("n", [8193, 9000])
def test_llama3_8b_width_crosses_the_real_chunk_boundary(cuda_device, diag, built, n):
    """Prompts longer than the 8,192-token prefill chunk at full width: the second chunk's queries attend
    to the first chunk's KV through the prefix cache (q_pos0 = 8192 in RoPE and attention).  8,193 puts a
    single row in the second chunk; 9,000 gives it 808.  Then decode continues from the 2-chunk prefix."""
    spec, model, blob, inv = built("llama-3-8b", 77)
    e = _engine(spec, blob, inv, 5120 + 4096)
    prompt = _tokens(spec, n, n)
    _teacher_forced(e, model, spec, prompt, [_tokens(spec, 3, 50 + i) for i in range(3)], diag, f"llama-3-8b/p{n}")
    e.close()


@pytest.mark.parametrize("name", ["mistral-7b", "qwen2-7b", "phi-3-mini", "gemma-7b"])
def test_other_families_at_full_width(cuda_device, diag, built, name):
    """Config 3's other four models at their real widths (2 layers): Mistral (the `mistral` branch of the
    HF config), Qwen2 (QKV bias, d = 3584, G = 7 -> two opponent groups per KV head, 152K vocabulary),
    Phi-3 (MHA, head_dim 96 in the padded 128-wide tiles), Gemma (MHA, head_dim 256, GeGLU, tied 256K
    lm_head, sqrt(d) embedding scale) on a 2,100-token prompt, 3 opponents, 4 teacher-forced steps."""
    spec, model, blob, inv = built(name, 5)
    e = _engine(spec, blob, inv, 2304)
    prompt = _tokens(spec, 2100, 3)
    _teacher_forced(e, model, spec, prompt, [_tokens(spec, 4, 20 + i) for i in range(3)], diag, f"{name}/p2100")
    e.close()
