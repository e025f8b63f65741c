"""Model shapes the local engine serves (SURVEY.md §8(d) model table).

The reference names an opponent by a litellm model string and never sees a
shape (skills/adversarial-spec/scripts/models.py:614-616).  Here an opponent
string such as ``b200/llama-3-8b`` resolves to a ``ModelSpec``; no pretrained
weights exist offline, so weights are seeded random-init of this shape.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Optional

LOCAL_PREFIXES = ("b200/", "local/")


@dataclass(frozen=True)
class ModelSpec:
    name: str
    family: str  # llama | mistral | qwen2 | phi3 | gemma
    n_layers: int
    d_model: int
    n_heads: int
    n_kv_heads: int
    head_dim: int
    d_ff: int
    vocab_size: int
    rope_theta: float = 10000.0
    norm_eps: float = 1e-5
    act: int = 0  # 0 SiLU-gated, 1 tanh-GELU-gated
    qkv_bias: bool = False
    tied_lm_head: bool = False
    embed_scale: float = 1.0
    max_position_embeddings: int = 32768 + 1024

    @property
    def qkv_dim(self) -> int:
        return (self.n_heads + 2 * self.n_kv_heads) * self.head_dim

    @property
    def kv_bytes_per_token(self) -> int:
        """bf16 K+V bytes per token over all layers (SURVEY.md §8(d) last column)."""
        return 2 * self.n_layers * self.n_kv_heads * self.head_dim * 2

    def n_params(self) -> int:
        d, hd = self.d_model, self.n_heads * self.head_dim
        per_layer = self.qkv_dim * d + d * hd + 3 * self.d_ff * d + 2 * d
        if self.qkv_bias:
            per_layer += self.qkv_dim
        total = self.n_layers * per_layer + d + self.vocab_size * d
        if not self.tied_lm_head:
            total += self.vocab_size * d
        return total

    def matmul_params(self) -> int:
        """P_mm of SURVEY.md §8(d): every matmul weight except embedding and lm_head."""
        d, hd = self.d_model, self.n_heads * self.head_dim
        return self.n_layers * (self.qkv_dim * d + d * hd + 3 * self.d_ff * d)

    def decode_weight_bytes(self) -> int:
        """W_read of SURVEY.md §8(d): bf16 bytes streamed per decode step."""
        return 2 * (self.matmul_params() + self.vocab_size * self.d_model)

    def prefill_flops(self, n_tokens: int) -> float:
        """flops_alg of SURVEY.md §8(d): causal attention at half the dense S^2,
        lm_head for the last position only."""
        s = float(n_tokens)
        return (
            2.0 * s * self.matmul_params()
            + self.n_layers * 2.0 * s * s * self.n_heads * self.head_dim
            + 2.0 * self.vocab_size * self.d_model
        )

    def decode_step_bytes(self, prefix_len: int, suffix_lens: list[int]) -> float:
        """bytes_alg of SURVEY.md §8(d) for one decode step of b opponents."""
        kvb = self.kv_bytes_per_token
        return self.decode_weight_bytes() + kvb * (prefix_len + sum(suffix_lens)) + kvb * len(suffix_lens)


def _llama(name, L, d, H, Hkv, ffn, V, theta=500000.0, eps=1e-5, **kw) -> ModelSpec:
    return ModelSpec(name, kw.pop("family", "llama"), L, d, H, Hkv, kw.pop("head_dim", d // H), ffn, V,
                     rope_theta=theta, norm_eps=eps, **kw)


REGISTRY: dict[str, ModelSpec] = {
    s.name: s
    for s in [
        _llama("llama-3-8b", 32, 4096, 32, 8, 14336, 128256),
        _llama("llama-3-70b", 80, 8192, 64, 8, 28672, 128256),
        _llama("mistral-7b", 32, 4096, 32, 8, 14336, 32000, theta=10000.0, family="mistral"),
        _llama("qwen2-7b", 28, 3584, 28, 4, 18944, 152064, theta=1000000.0, eps=1e-6,
               family="qwen2", qkv_bias=True),
        _llama("phi-3-mini", 32, 3072, 32, 32, 8192, 32064, theta=10000.0, family="phi3"),
        _llama("gemma-7b", 28, 3072, 16, 16, 24576, 256000, theta=10000.0, eps=1e-6, family="gemma",
               head_dim=256, act=1, tied_lm_head=True, embed_scale=math.sqrt(3072.0)),
        # small shapes for parity tests (same code paths, seconds on a CPU oracle)
        _llama("tiny-llama", 2, 256, 4, 2, 512, 1024, head_dim=64),
        _llama("tiny-llama-128", 2, 512, 4, 2, 1024, 2048, head_dim=128),
        _llama("tiny-qwen2", 2, 256, 4, 2, 512, 1024, head_dim=64, family="qwen2", qkv_bias=True,
               theta=1000000.0, eps=1e-6),
        _llama("tiny-gqa4", 2, 512, 8, 2, 768, 1024, head_dim=64),
        _llama("tiny-mistral", 2, 512, 8, 2, 1024, 2048, head_dim=64, family="mistral", theta=10000.0),
        _llama("tiny-phi3", 2, 192, 2, 2, 512, 1024, head_dim=96, family="phi3", theta=10000.0),
        _llama("tiny-gemma256", 2, 512, 2, 2, 512, 1024, head_dim=256, family="gemma", act=1,
               tied_lm_head=True, embed_scale=math.sqrt(512.0), theta=10000.0, eps=1e-6),
        _llama("tiny-gemma", 2, 256, 2, 2, 512, 1024, head_dim=128, family="gemma", act=1,
               tied_lm_head=True, embed_scale=16.0, theta=10000.0, eps=1e-6),
        # Llama-3-8B layer shape with 2 layers: the CPU-baseline sample (per-layer cost scales to 32)
        _llama("llama-3-8b-2layer", 2, 4096, 32, 8, 14336, 128256),
        # Llama-3-70B's layer shape (d 8192, 64/8 heads, MLP 28672), 2 layers, a 32K vocabulary: the
        # tensor-parallel parity case a CPU oracle can hold (the lm_head is not what TP=8 is about)
        _llama("llama-3-70b-2layer-v32k", 2, 8192, 64, 8, 28672, 32768),
    ]
}


def is_local_model(model: str) -> bool:
    return model.startswith(LOCAL_PREFIXES)


def split_persona(model: str) -> tuple[str, Optional[str]]:
    """``b200/<name>@<persona>`` -> (``b200/<name>``, persona): a per-opponent persona (the
    reference has one global ``--persona`` for the whole panel, debate.py:835).  Only local model strings carry
    the suffix — remote provider ids may contain ``@`` themselves and pass through untouched."""
    if is_local_model(model) and "@" in model:
        base, persona = model.split("@", 1)
        return base, (persona.strip() or None)
    return model, None


def resolve(model: str) -> ModelSpec:
    """``b200/<name>`` or ``local/<name>`` (or a bare registry name), with or without an ``@persona`` suffix
    -> ModelSpec."""
    key = split_persona(model)[0]
    model = key
    for p in LOCAL_PREFIXES:
        if model.startswith(p):
            key = model[len(p):]
    key = key.lower()
    if key not in REGISTRY:
        raise KeyError(f"unknown local model {model!r}; known: {', '.join(sorted(REGISTRY))}")
    return REGISTRY[key]


def with_layers(spec: ModelSpec, n_layers: int, name: Optional[str] = None) -> ModelSpec:
    return replace(spec, n_layers=n_layers, name=name or f"{spec.name}-{n_layers}layer")
