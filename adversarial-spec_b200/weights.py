"""Weight blob shared byte-for-byte by the engine and the CPU oracle.

Layout (restates ``make_layout`` in csrc/engine.cu; tests check both agree):
every tensor starts on a 256-byte boundary, in this order —
  embed bf16 [V, d]
  per layer: attn_norm f32 [d]; wqkv bf16 [(H+2Hkv)*Dh, d] (q heads, k heads,
  v heads); bqkv f32 [(H+2Hkv)*Dh] (Qwen2 only); wo bf16 [d, H*Dh];
  mlp_norm f32 [d]; wgu bf16 [2*ffn, d] with rows interleaved (2i = gate_i,
  2i+1 = up_i) so one GEMM tile holds both halves of the gated activation;
  wd bf16 [d, ffn]
  final_norm f32 [d]; lm_head bf16 [V, d] (absent when tied to embed)
Matrices are row-major [out_features, in_features], i.e. K-major for both the
prefill GEMM and the decode GEMV.  Norm vectors stay fp32 so Gemma's (1 + w)
is exact.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .model_spec import ModelSpec

ALIGN = 256


def _align(v: int) -> int:
    return (v + ALIGN - 1) // ALIGN * ALIGN


@dataclass
class BlobLayout:
    offsets: dict  # (layer or -1, name) -> (byte offset, shape, "bf16" | "f32")
    total: int


def blob_layout(spec: ModelSpec) -> BlobLayout:
    off = 0
    offsets: dict = {}

    def take(key, shape, kind):
        nonlocal off
        n = int(np.prod(shape)) * (2 if kind == "bf16" else 4)
        offsets[key] = (off, tuple(shape), kind)
        off = _align(off + n)

    d, qkv, hd = spec.d_model, spec.qkv_dim, spec.n_heads * spec.head_dim
    take((-1, "embed"), (spec.vocab_size, d), "bf16")
    for l in range(spec.n_layers):
        take((l, "attn_norm"), (d,), "f32")
        take((l, "wqkv"), (qkv, d), "bf16")
        if spec.qkv_bias:
            take((l, "bqkv"), (qkv,), "f32")
        take((l, "wo"), (d, hd), "bf16")
        take((l, "mlp_norm"), (d,), "f32")
        take((l, "wgu"), (2 * spec.d_ff, d), "bf16")
        take((l, "wd"), (d, spec.d_ff), "bf16")
    take((-1, "final_norm"), (d,), "f32")
    if not spec.tied_lm_head:
        take((-1, "lm_head"), (spec.vocab_size, d), "bf16")
    return BlobLayout(offsets, off)


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    rounded = u + 0x7FFF + ((u >> 16) & 1)
    return (rounded >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


class BlobWriter:
    def __init__(self, spec: ModelSpec):
        self.spec = spec
        self.layout = blob_layout(spec)
        self.buf = np.zeros(self.layout.total, dtype=np.uint8)

    def put(self, layer: int, name: str, value: np.ndarray) -> None:
        off, shape, kind = self.layout.offsets[(layer, name)]
        v = np.asarray(value, dtype=np.float32)
        if tuple(v.shape) != shape:
            raise ValueError(f"{name}[{layer}]: shape {v.shape}, expected {shape}")
        if kind == "bf16":
            raw = f32_to_bf16_bits(v).reshape(-1).view(np.uint8)
        else:
            raw = np.ascontiguousarray(v).reshape(-1).view(np.uint8)
        self.buf[off: off + raw.size] = raw

    def get(self, layer: int, name: str) -> np.ndarray:
        off, shape, kind = self.layout.offsets[(layer, name)]
        n = int(np.prod(shape))
        if kind == "bf16":
            return bf16_bits_to_f32(self.buf[off: off + 2 * n].view(np.uint16)).reshape(shape)
        return self.buf[off: off + 4 * n].view(np.float32).reshape(shape).copy()


def interleave_gate_up(gate: np.ndarray, up: np.ndarray) -> np.ndarray:
    """[ffn, d] x2 -> [2*ffn, d] with rows (gate_0, up_0, gate_1, up_1, ...)."""
    out = np.empty((2 * gate.shape[0], gate.shape[1]), dtype=np.float32)
    out[0::2] = gate
    out[1::2] = up
    return out
