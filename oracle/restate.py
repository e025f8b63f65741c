"""Plain-torch restatement of the decoder forward, reading the engine's own blob.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Two uses:
  * ``forward_logits(..., engine_rounding=False)`` is a from-scratch fp32
    restatement of what HF computes (pinned against transformers itself in
    tests/test_oracle.py) — it proves the blob layout (fused QKV, interleaved
    gate/up, Gemma's folded 1+w) carries the same function as the HF modules.
  * ``engine_rounding=True`` rounds activations to bf16 at exactly the points
    the CUDA path does (norm output, qkv, rotated q/k, attention output, gated
    MLP activation; fp32 residual stream), which predicts the engine-vs-HF
    error and sets the tolerance the GPU parity tests state.

Semantics restated (transformers 5.5.0): RMSNorm modeling_llama.py:53-66;
rotate_half RoPE :138-146 with angles = float32(pos) * inv_freq; attention
softmax(q k^T / sqrt(Dh) + causal) v with GQA head h -> kv head h // (H/Hkv)
:187-220; SwiGLU down(act(gate) * up) :183; Gemma GeGLU-tanh and sqrt(d)
embedding scale gemma/modeling_gemma.py:382.
"""

from __future__ import annotations

import math

import numpy as np
import torch

import advspec_loader

advspec_loader.load()
from advspec_b200.model_spec import ModelSpec  # noqa: E402
from advspec_b200.weights import blob_layout, bf16_bits_to_f32  # noqa: E402


def _bf16(x: torch.Tensor, on: bool) -> torch.Tensor:
    return x.bfloat16().float() if on else x


class BlobModel:
    def __init__(self, spec: ModelSpec, blob: np.ndarray, inv_freq: np.ndarray | None = None,
                 embed_rows: int | None = None):
        """`embed_rows`: set when `spec`/`blob` are one tensor-parallel rank's share (weights.shard_blob)."""
        self.spec = spec
        lay = blob_layout(spec, embed_rows)
        assert blob.nbytes == lay.total
        self.t: dict = {}
        for (l, name), (off, shape, kind) in lay.offsets.items():
            n = int(np.prod(shape))
            if kind == "bf16":
                a = bf16_bits_to_f32(blob[off: off + 2 * n].view(np.uint16))
            else:
                a = blob[off: off + 4 * n].view(np.float32)
            self.t[(l, name)] = torch.from_numpy(a.reshape(shape).copy())
        half = spec.head_dim // 2
        if inv_freq is None:
            inv_freq = 1.0 / (spec.rope_theta ** (torch.arange(0, spec.head_dim, 2).float() / spec.head_dim))
            inv_freq = inv_freq.numpy()
        self.inv_freq = torch.from_numpy(np.asarray(inv_freq, dtype=np.float32).reshape(half))

    def rmsnorm(self, x, w):
        var = x.pow(2).mean(-1, keepdim=True)
        return x * torch.rsqrt(var + self.spec.norm_eps) * w

    def rope(self, x, pos):
        # x [n, heads, Dh]; angles in fp32 exactly as HF: float(pos) * inv_freq
        ang = pos.float()[:, None] * self.inv_freq[None, :]
        cos, sin = ang.cos()[:, None, :], ang.sin()[:, None, :]
        half = self.spec.head_dim // 2
        a, b = x[..., :half], x[..., half:]
        return torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)

    @torch.no_grad()
    def forward_logits(self, tokens, engine_rounding: bool = False, tp_rank: int = 0, allreduce=None) -> np.ndarray:
        """With `allreduce` (in-place sum over ranks of a tensor) this is ONE tensor-parallel rank's
        computation as include/advspec_engine.h states it: `self` holds the rank's share, the partial
        o-proj / down-proj products are summed across ranks with rank 0's partial carrying the residual,
        and the result is the rank's vocabulary shard of the logits."""
        s, r = self.spec, engine_rounding

        def row_split(x, partial):
            if allreduce is None:
                return x + partial
            y = (x + partial) if tp_rank == 0 else partial.clone()
            allreduce(y)
            return y

        n = len(tokens)
        ids = torch.tensor(list(tokens), dtype=torch.long)
        x = self.t[(-1, "embed")][ids] * np.float32(s.embed_scale)
        pos = torch.arange(n)
        H, Hkv, Dh = s.n_heads, s.n_kv_heads, s.head_dim
        mask = torch.full((n, n), float("-inf")).triu(1)
        for l in range(s.n_layers):
            xn = _bf16(self.rmsnorm(x, self.t[(l, "attn_norm")]), r)
            qkv = xn @ self.t[(l, "wqkv")].T
            if s.qkv_bias:
                qkv = qkv + self.t[(l, "bqkv")]
            qkv = _bf16(qkv, r)
            q = qkv[:, : H * Dh].reshape(n, H, Dh)
            k = qkv[:, H * Dh: (H + Hkv) * Dh].reshape(n, Hkv, Dh)
            v = qkv[:, (H + Hkv) * Dh:].reshape(n, Hkv, Dh)
            q, k = _bf16(self.rope(q, pos), r), _bf16(self.rope(k, pos), r)
            k = k.repeat_interleave(H // Hkv, dim=1)
            v = v.repeat_interleave(H // Hkv, dim=1)
            sc = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(Dh) + mask
            att = torch.einsum("hqk,khd->qhd", sc.softmax(-1), v).reshape(n, H * Dh)
            att = _bf16(att, r)
            x = row_split(x, att @ self.t[(l, "wo")].T)
            xn = _bf16(self.rmsnorm(x, self.t[(l, "mlp_norm")]), r)
            gu = xn @ self.t[(l, "wgu")].T
            g, u = gu[:, 0::2], gu[:, 1::2]
            act = torch.nn.functional.gelu(g, approximate="tanh") if s.act == 1 else torch.nn.functional.silu(g)
            h = _bf16(act * u, r)
            x = row_split(x, h @ self.t[(l, "wd")].T)
        xn = _bf16(self.rmsnorm(x, self.t[(-1, "final_norm")]), r)
        head = self.t[(-1, "embed")] if s.tied_lm_head else self.t[(-1, "lm_head")]
        return (xn @ head.T).numpy()
