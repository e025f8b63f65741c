"""HF-transformers CPU forward of seeded random-init weights — the numerical oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's arithmetic lives
in third-party services behind ``litellm.completion``
(skills/adversarial-spec/scripts/models.py:628; litellm==1.80.13 pinned in
requirements.txt:1, not vendored), so BASELINE.json's north_star prescribes
this oracle instead: "generated logits match an HF CPU forward of the same
random-init weights on the same input tokens".  No reference test pins a
numeric model output — parity against the *reference* is
pinned only for the text/fan-out contract (tests/golden/reference_cli_*.json);
the numeric side is pinned against transformers 5.5.0 itself, which is
importable both here and on the GPU box.

Per-family semantics followed (transformers/models/...):
  llama/modeling_llama.py   RMSNorm fp32 :53-66, rotate_half RoPE :138-146,
                            SwiGLU :183, GQA repeat_kv :187
  qwen2/modeling_qwen2.py   q/k/v bias :200-202
  phi3/modeling_phi3.py     fused qkv_proj :224-237, gate_up_proj chunk :54-61
  gemma/modeling_gemma.py   norm x*(1+w) :77, embed * sqrt(d) :382, GeGLU-tanh
  mistral/modeling_mistral.py  sliding_window set to None (full attention; stated in DESIGN.md)
"""

from __future__ import annotations

import numpy as np
import torch

import advspec_loader

advspec_loader.load()
from advspec_b200.model_spec import ModelSpec  # noqa: E402
from advspec_b200.weights import BlobWriter  # noqa: E402


def hf_config(spec: ModelSpec):
    import transformers as tf

    common = dict(
        vocab_size=spec.vocab_size, hidden_size=spec.d_model, intermediate_size=spec.d_ff,
        num_hidden_layers=spec.n_layers, num_attention_heads=spec.n_heads,
        num_key_value_heads=spec.n_kv_heads, max_position_embeddings=spec.max_position_embeddings,
        rms_norm_eps=spec.norm_eps, rope_theta=spec.rope_theta, tie_word_embeddings=spec.tied_lm_head,
        attention_dropout=0.0,
    )
    fam = spec.family
    if fam == "llama":
        cfg = tf.LlamaConfig(head_dim=spec.head_dim, attention_bias=False, mlp_bias=False, **common)
    elif fam == "mistral":
        cfg = tf.MistralConfig(head_dim=spec.head_dim, sliding_window=None, **common)
    elif fam == "qwen2":
        cfg = tf.Qwen2Config(use_sliding_window=False, **common)
    elif fam == "phi3":
        cfg = tf.Phi3Config(pad_token_id=0, **common)
    elif fam == "gemma":
        cfg = tf.GemmaConfig(head_dim=spec.head_dim, hidden_activation="gelu_pytorch_tanh",
                             hidden_act="gelu_pytorch_tanh", **common)
    else:
        raise ValueError(f"unknown family {fam}")
    cfg._attn_implementation = "eager"
    return cfg


def build_hf_model(spec: ModelSpec, seed: int):
    """Random-init HF model (fp32 compute) whose every parameter value is bf16-representable
    (norm/bias vectors stay fp32-exact: the blob stores them in fp32)."""
    import transformers as tf

    torch.manual_seed(seed)
    cfg = hf_config(spec)
    model = tf.AutoModelForCausalLM.from_config(cfg, attn_implementation="eager").float().eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:
                # make norm weights / biases non-trivial so a dropped one is visible
                base = 0.0 if (name.endswith("bias") or spec.family == "gemma") else 1.0
                p.copy_(base + 0.1 * torch.randn(p.shape, generator=g))
                if name.endswith("bias"):
                    p.mul_(0.2)
            else:
                p.copy_(p.bfloat16().float())
    return model


def build_hf_model_fast(spec: ModelSpec, seed: int, std: float = 0.02):
    """Same contract as build_hf_model for FULL-WIDTH shapes (billions of parameters): HF's own
    per-module initialisation is skipped and every matrix is drawn N(0, std) in one pass, then rounded
    to bf16 once (so the engine's blob holds exactly the oracle's values).  The forward is still
    transformers' own modeling code; only the weight VALUES differ from build_hf_model's."""
    import transformers as tf
    from transformers.initialization import no_init_weights

    cfg = hf_config(spec)
    with no_init_weights():
        model = tf.AutoModelForCausalLM.from_config(cfg, attn_implementation="eager").float().eval()
    g = torch.Generator().manual_seed(seed)
    chunk = 32 * 1024 * 1024  # draw and round through reused scratch buffers: no 6 GB temporaries
    scr, scb = torch.empty(chunk), torch.empty(chunk, dtype=torch.bfloat16)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:
                base = 0.0 if (name.endswith("bias") or spec.family == "gemma") else 1.0
                p.copy_(base + 0.1 * torch.randn(p.shape, generator=g))
                if name.endswith("bias"):
                    p.mul_(0.2)
            else:
                flat = p.view(-1)
                for o in range(0, flat.numel(), chunk):
                    n = min(chunk, flat.numel() - o)
                    scr[:n].normal_(0.0, std, generator=g)
                    scb[:n].copy_(scr[:n])
                    flat[o:o + n].copy_(scb[:n])
        if spec.tied_lm_head:
            model.tie_weights()
    return model


def rope_inv_freq(model) -> np.ndarray:
    return model.model.rotary_emb.inv_freq.detach().float().numpy().copy()


def export_blob(spec: ModelSpec, model) -> np.ndarray:
    """HF state dict -> the engine's weight blob (uint8 array)."""
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    w = BlobWriter(spec)
    w.put(-1, "embed", sd["model.embed_tokens.weight"])
    gemma = spec.family == "gemma"
    for l in range(spec.n_layers):
        p = f"model.layers.{l}."
        w.put(l, "attn_norm", sd[p + "input_layernorm.weight"] + (1.0 if gemma else 0.0))
        if spec.family == "phi3":
            qkv = sd[p + "self_attn.qkv_proj.weight"]
            gu = sd[p + "mlp.gate_up_proj.weight"]
            gate, up = gu[: spec.d_ff], gu[spec.d_ff:]
        else:
            qkv = torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"],
                             sd[p + "self_attn.v_proj.weight"]], dim=0)
            gate, up = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
        w.put(l, "wqkv", qkv)
        if spec.qkv_bias:
            w.put(l, "bqkv", torch.cat([sd[p + "self_attn.q_proj.bias"], sd[p + "self_attn.k_proj.bias"],
                                        sd[p + "self_attn.v_proj.bias"]]))
        w.put(l, "wo", sd[p + "self_attn.o_proj.weight"])
        w.put(l, "mlp_norm", sd[p + "post_attention_layernorm.weight"] + (1.0 if gemma else 0.0))
        # rows (gate_0, up_0, gate_1, up_1, ...): weights.interleave_gate_up, on torch tensors
        w.put(l, "wgu", torch.stack([gate, up], dim=1).reshape(2 * spec.d_ff, spec.d_model))
        w.put(l, "wd", sd[p + "mlp.down_proj.weight"])
    w.put(-1, "final_norm", sd["model.norm.weight"] + (1.0 if gemma else 0.0))
    if not spec.tied_lm_head:
        w.put(-1, "lm_head", sd["lm_head.weight"])
    return w.buf


@torch.no_grad()
def hf_logits(model, tokens) -> np.ndarray:
    """fp32 logits [n_tokens, vocab] for one sequence."""
    ids = torch.tensor([list(tokens)], dtype=torch.long)
    return model(input_ids=ids, use_cache=False).logits[0].float().numpy()


@torch.no_grad()
def hf_logits_last(model, tokens, keep: int, attn: str = "sdpa") -> np.ndarray:
    """fp32 logits of the LAST `keep` positions only ([keep, vocab]) — full-width vocabularies make
    all-position logits (n x 128K floats) pointless.  `attn="sdpa"` keeps the S x S score matrix out of
    memory for 5K-9K-token prompts (same fp32 math as the eager path; torch's CPU kernel)."""
    ids = torch.tensor([list(tokens)], dtype=torch.long)
    prev = model.config._attn_implementation
    model.config._attn_implementation = attn
    try:
        out = model(input_ids=ids, use_cache=False, logits_to_keep=keep).logits[0].float().numpy()
    finally:
        model.config._attn_implementation = prev
    return out


@torch.no_grad()
def hf_greedy(model, tokens, n_new: int) -> list[int]:
    ids = list(tokens)
    out = []
    past = None
    cur = torch.tensor([ids], dtype=torch.long)
    for _ in range(n_new):
        r = model(input_ids=cur, past_key_values=past, use_cache=True)
        past = r.past_key_values
        nxt = int(r.logits[0, -1].argmax())
        out.append(nxt)
        cur = torch.tensor([[nxt]], dtype=torch.long)
    return out
