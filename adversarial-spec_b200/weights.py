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

from dataclasses import dataclass, replace

import numpy as np

from .model_spec import ModelSpec

ALIGN = 256


def _align(v: int) -> int:
    return (v + ALIGN - 1) // ALIGN * ALIGN


@dataclass
class BlobLayout:
    offsets: dict  # (layer or -1, name) -> (byte offset, shape, "bf16" | "f32")
    total: int


def tp_local_spec(spec: ModelSpec, tp_size: int) -> ModelSpec:
    """The shape ONE tensor-parallel rank computes with (restates `localize` in csrc/engine.cu):
    its share of the heads, of the MLP width and of the lm_head rows; d_model and depth are whole."""
    if tp_size == 1:
        return spec
    if spec.n_heads % tp_size or spec.n_kv_heads % tp_size or spec.d_ff % (8 * tp_size) \
            or spec.vocab_size % tp_size or spec.tied_lm_head:
        raise ValueError(f"{spec.name} cannot be split {tp_size} ways")
    return replace(spec, n_heads=spec.n_heads // tp_size, n_kv_heads=spec.n_kv_heads // tp_size,
                   d_ff=spec.d_ff // tp_size, vocab_size=spec.vocab_size // tp_size)


def blob_layout(spec: ModelSpec, embed_rows: int | None = None) -> BlobLayout:
    """`embed_rows`: rows of the embedding table when `spec` is a tensor-parallel share (the table
    is replicated whole on every rank)."""
    off = 0
    offsets: dict = {}

    def take(key, shape, kind):
        nonlocal off
        n = int(np.prod(shape)) * (2 if kind == "bf16" else 4)
        offsets[key] = (off, tuple(shape), kind)
        off = _align(off + n)

    d, qkv, hd = spec.d_model, spec.qkv_dim, spec.n_heads * spec.head_dim
    take((-1, "embed"), (embed_rows or spec.vocab_size, d), "bf16")
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
    def __init__(self, spec: ModelSpec, embed_rows: int | None = None):
        self.spec = spec
        self.layout = blob_layout(spec, embed_rows)
        self.buf = np.zeros(self.layout.total, dtype=np.uint8)

    def put(self, layer: int, name: str, value) -> None:
        """`value`: a numpy array, or a torch tensor (bf16 matrices then convert through torch's
        multi-threaded round-to-nearest-even cast — same bits, much faster for full-width models)."""
        off, shape, kind = self.layout.offsets[(layer, name)]
        if kind == "bf16" and hasattr(value, "bfloat16"):
            if tuple(value.shape) != shape:
                raise ValueError(f"{name}[{layer}]: shape {tuple(value.shape)}, expected {shape}")
            import torch

            raw = value.detach().contiguous().bfloat16().view(torch.int16).numpy().reshape(-1).view(np.uint8)
            self.buf[off: off + raw.size] = raw
            return
        if hasattr(value, "detach"):
            value = value.detach().float().numpy()
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


def shard_blob(full: np.ndarray, spec: ModelSpec, tp_rank: int, tp_size: int) -> np.ndarray:
    """Rank `tp_rank`'s weight blob cut out of the whole model's blob, byte for byte (no re-rounding):
    q/k/v rows of its heads, the matching wo columns, its gate/up row pairs and wd columns, its
    lm_head rows; norm vectors and the embedding table whole (include/advspec_engine.h)."""
    if tp_size == 1:
        return full
    loc = tp_local_spec(spec, tp_size)
    src, dst = blob_layout(spec), blob_layout(loc, spec.vocab_size)
    out = np.zeros(dst.total, dtype=np.uint8)

    def view(buf, lay, key):
        off, shape, kind = lay.offsets[key]
        n = int(np.prod(shape))
        dt = np.uint16 if kind == "bf16" else np.float32
        return buf[off: off + n * (2 if kind == "bf16" else 4)].view(dt).reshape(shape)

    r, dh = tp_rank, spec.head_dim
    hl, kl, fl, vl = loc.n_heads * dh, loc.n_kv_heads * dh, loc.d_ff, loc.vocab_size
    hq, hk = spec.n_heads * dh, spec.n_kv_heads * dh
    qkv_rows = np.concatenate([np.arange(r * hl, (r + 1) * hl), hq + np.arange(r * kl, (r + 1) * kl),
                               hq + hk + np.arange(r * kl, (r + 1) * kl)])
    view(out, dst, (-1, "embed"))[...] = view(full, src, (-1, "embed"))
    view(out, dst, (-1, "final_norm"))[...] = view(full, src, (-1, "final_norm"))
    view(out, dst, (-1, "lm_head"))[...] = view(full, src, (-1, "lm_head"))[r * vl:(r + 1) * vl]
    for l in range(spec.n_layers):
        for name in ("attn_norm", "mlp_norm"):
            view(out, dst, (l, name))[...] = view(full, src, (l, name))
        view(out, dst, (l, "wqkv"))[...] = view(full, src, (l, "wqkv"))[qkv_rows]
        if spec.qkv_bias:
            view(out, dst, (l, "bqkv"))[...] = view(full, src, (l, "bqkv"))[qkv_rows]
        view(out, dst, (l, "wo"))[...] = view(full, src, (l, "wo"))[:, r * hl:(r + 1) * hl]
        view(out, dst, (l, "wgu"))[...] = view(full, src, (l, "wgu"))[2 * r * fl:2 * (r + 1) * fl]
        view(out, dst, (l, "wd"))[...] = view(full, src, (l, "wd"))[:, r * fl:(r + 1) * fl]
    return out
