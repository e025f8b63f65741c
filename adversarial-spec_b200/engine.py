"""ctypes binding of include/advspec_engine.h — the only way Python reaches the GPU.

The library is built in-tree (``adversarial-spec_b200/libadvspec_b200.so``) by
``__graft_entry__.build()`` / ``csrc/Makefile``.  A missing library or a
missing CUDA device is an error: there is no CPU fallback on the product path
(the CPU oracle lives under ``oracle/`` and is only ever called by tests and
the benchmark's ``cpu_baseline`` leg).
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from .model_spec import ModelSpec

ABI_VERSION = 1
_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libadvspec_b200.so"

STATUS_NAMES = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "OOM", 4: "STATE", 5: "KERNEL"}


class EngineError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"advspec engine error {STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class ModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_layers", C.c_int32),
        ("d_model", C.c_int32),
        ("n_heads", C.c_int32),
        ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32),
        ("d_ff", C.c_int32),
        ("vocab_size", C.c_int32),
        ("act", C.c_int32),
        ("qkv_bias", C.c_int32),
        ("tied_lm_head", C.c_int32),
        ("max_prefix_tokens", C.c_int32),
        ("max_new_tokens", C.c_int32),
        ("max_seqs", C.c_int32),
        ("tp_rank", C.c_int32),
        ("tp_size", C.c_int32),
        ("rope_theta", C.c_float),
        ("norm_eps", C.c_float),
        ("embed_scale", C.c_float),
        ("reserved_f", C.c_float),
        ("reserved_i", C.c_int32 * 4),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("prefill_ms", C.c_float),
        ("decode_ms", C.c_float),
        ("decode_steps", C.c_int32),
        ("decode_batch", C.c_int32),
        ("kernel_launches", C.c_int64),
        ("gemv_ms", C.c_float),
        ("gemv_launches", C.c_int32),
        ("reserved", C.c_int32),
    ]


EXPORTED_SYMBOLS = [
    "advspec_weight_blob_bytes", "advspec_weight_offset", "advspec_engine_create",
    "advspec_engine_destroy", "advspec_last_error", "advspec_tp_unique_id", "advspec_tp_init",
    "advspec_tp_ipc_export", "advspec_tp_ipc_import",
    "advspec_load_weights",
    "advspec_init_weights_random", "advspec_set_rope_inv_freq", "advspec_prefill", "advspec_prefill_extend", "advspec_fork",
    "advspec_decode", "advspec_append_tail", "advspec_decode_step", "advspec_get_logits", "advspec_prefill_logits",
    "advspec_release_seqs", "advspec_release_prefix", "advspec_prefix_kv_region",
    "advspec_prefix_adopt", "advspec_get_timing", "advspec_profile_decode_step",
    "advspec_decode_step_bytes", "advspec_ktrace_enable", "advspec_ktrace_read", "advspec_ktrace_phases", "advspec_op_gemm", "advspec_op_gemm_check", "advspec_op_gemv",
    "advspec_op_attn_prefill", "advspec_op_attn_decode",
]

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load the in-tree shared library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("ADVSPEC_LIB", LIB_PATH))
    if not path.exists():
        raise EngineError(2, f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             f"or `make -C {_PKG_DIR / 'csrc'}` (no CPU fallback exists)")
    lib = C.CDLL(str(path))
    P = C.POINTER
    vp, i32, i64, sz, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float
    sig = {
        "advspec_weight_blob_bytes": (sz, [P(ModelDesc)]),
        "advspec_weight_offset": (sz, [P(ModelDesc), i32, C.c_char_p]),
        "advspec_engine_create": (i32, [P(ModelDesc), i32, P(vp)]),
        "advspec_engine_destroy": (None, [vp]),
        "advspec_last_error": (C.c_char_p, [vp]),
        "advspec_tp_unique_id": (i32, [P(C.c_uint8)]),
        "advspec_tp_init": (i32, [vp, P(C.c_uint8)]),
        "advspec_tp_ipc_export": (i32, [vp, P(C.c_uint8)]),
        "advspec_tp_ipc_import": (i32, [vp, P(C.c_uint8)]),
        "advspec_load_weights": (i32, [vp, vp, sz]),
        "advspec_init_weights_random": (i32, [vp, C.c_uint64, f32]),
        "advspec_set_rope_inv_freq": (i32, [vp, P(f32), i32]),
        "advspec_prefill": (i32, [vp, P(i32), i32, P(i32)]),
        "advspec_prefill_extend": (i32, [vp, i32, i32, P(i32), i32, P(i32)]),
        "advspec_fork": (i32, [vp, i32, i32, P(C.c_uint64), P(i32)]),
        "advspec_decode": (i32, [vp, P(i32), i32, i32, f32, i32, P(i32), P(i32)]),
        "advspec_append_tail": (i32, [vp, i32, P(i32), i32]),
        "advspec_decode_step": (i32, [vp, P(i32), i32, P(i32)]),
        "advspec_get_logits": (i32, [vp, i32, P(f32)]),
        "advspec_prefill_logits": (i32, [vp, P(i32), i32, P(f32)]),
        "advspec_release_seqs": (i32, [vp, P(i32), i32]),
        "advspec_release_prefix": (i32, [vp, i32]),
        "advspec_prefix_kv_region": (i32, [vp, i32, P(vp), P(sz)]),
        "advspec_prefix_adopt": (i32, [vp, i32, P(f32), P(i32)]),
        "advspec_get_timing": (i32, [vp, P(Timing)]),
        "advspec_profile_decode_step": (i32, [vp, P(i32), i32]),
        "advspec_decode_step_bytes": (i32, [vp, P(i32), i32, P(C.c_double), P(C.c_double)]),
        "advspec_ktrace_enable": (i32, [vp, i32]),
        "advspec_ktrace_read": (i32, [vp, P(C.c_uint64), i32, P(i32)]),
        "advspec_ktrace_phases": (i32, [vp, P(C.c_uint64)]),
        "advspec_op_gemm": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, i32, i32, i32, i32, i32]),
        "advspec_op_gemm_check": (i32, [i32, vp, i64, vp, i64, vp, i64, vp, i32, i32, i32, i32, i32]),
        "advspec_op_gemv": (i32, [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32]),
        "advspec_op_attn_prefill": (i32, [i32, vp, i64, vp, vp, i64, vp, i32, i32, i32, i32, i32, i32]),
        "advspec_op_attn_decode": (i32, [i32, vp, vp, vp, vp, vp, i64, i32, vp, vp, i64, P(i32), vp, i32, i32,
                                         i32, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def make_desc(spec: ModelSpec, max_prefix_tokens: int, max_new_tokens: int, max_seqs: int,
              tp_rank: int = 0, tp_size: int = 1) -> ModelDesc:
    """`spec` is always the WHOLE model; tp_rank/tp_size select the share one handle holds."""
    d = ModelDesc()
    d.abi_version = ABI_VERSION
    d.n_layers, d.d_model, d.n_heads, d.n_kv_heads = spec.n_layers, spec.d_model, spec.n_heads, spec.n_kv_heads
    d.head_dim, d.d_ff, d.vocab_size = spec.head_dim, spec.d_ff, spec.vocab_size
    d.act, d.qkv_bias, d.tied_lm_head = spec.act, int(spec.qkv_bias), int(spec.tied_lm_head)
    d.max_prefix_tokens, d.max_new_tokens, d.max_seqs = max_prefix_tokens, max_new_tokens, max_seqs
    d.tp_rank, d.tp_size = tp_rank, tp_size
    d.rope_theta, d.norm_eps, d.embed_scale = spec.rope_theta, spec.norm_eps, spec.embed_scale
    return d


def _i32(a: Sequence[int]) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


@dataclass
class DecodeResult:
    tokens: list[list[int]]  # per opponent, only the emitted tokens
    lens: list[int]


class Engine:
    """One model's weights + KV on one GPU.  Thread-safe (the C side serialises per handle;
    ctypes releases the GIL for the duration of each call)."""

    def __init__(self, spec: ModelSpec, device: int = 0, max_prefix_tokens: int = 4096 + 1024,
                 max_new_tokens: int = 1024, max_seqs: int = 8, tp_rank: int = 0, tp_size: int = 1):
        self.lib = load_library()
        self.spec = spec
        self.device = device
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.vocab_local = spec.vocab_size // tp_size  # columns of this handle's logits
        self.desc = make_desc(spec, max_prefix_tokens, max_new_tokens, max_seqs, tp_rank, tp_size)
        h = C.c_void_p()
        st = self.lib.advspec_engine_create(C.byref(self.desc), device, C.byref(h))
        if st != 0:
            raise EngineError(st, (self.lib.advspec_last_error(None) or b"").decode())
        self.h = h

    # -- helpers ---------------------------------------------------------
    def _check(self, st: int) -> None:
        if st != 0:
            raise EngineError(st, (self.lib.advspec_last_error(self.h) or b"").decode())

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.advspec_engine_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # -- tensor parallelism (one process per GPU; see include/advspec_engine.h) ----------
    @staticmethod
    def tp_unique_id() -> bytes:
        lib = load_library()
        buf = (C.c_uint8 * 128)()
        st = lib.advspec_tp_unique_id(buf)
        if st != 0:
            raise EngineError(st, (lib.advspec_last_error(None) or b"").decode())
        return bytes(buf)

    def tp_init(self, unique_id: bytes) -> None:
        """Join this handle's NCCL communicator; every rank calls it, concurrently."""
        if len(unique_id) != 128:
            raise ValueError("the NCCL unique id is 128 bytes")
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.advspec_tp_init(self.h, buf))

    def tp_ipc_export(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self.lib.advspec_tp_ipc_export(self.h, buf))
        return bytes(buf)

    def tp_ipc_import(self, handles: Sequence[bytes]) -> None:
        """`handles`: every rank's tp_ipc_export() in rank order."""
        raw = b"".join(handles)
        if len(raw) != 64 * self.tp_size:
            raise ValueError("one 64-byte handle per rank")
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        self._check(self.lib.advspec_tp_ipc_import(self.h, buf))

    # -- weights -----------------------------------------------------------
    def load_weights(self, blob: np.ndarray) -> None:
        blob = np.ascontiguousarray(blob)
        self._check(self.lib.advspec_load_weights(self.h, blob.ctypes.data_as(C.c_void_p), blob.nbytes))

    def init_weights_random(self, seed: int, std: float = 0.02) -> None:
        self._check(self.lib.advspec_init_weights_random(self.h, C.c_uint64(seed), C.c_float(std)))

    def set_rope_inv_freq(self, inv_freq: np.ndarray) -> None:
        a = np.ascontiguousarray(inv_freq, dtype=np.float32)
        self._check(self.lib.advspec_set_rope_inv_freq(self.h, _p(a, C.c_float), a.size))

    # -- hot path -----------------------------------------------------------
    def prefill(self, tokens: Sequence[int]) -> int:
        t = _i32(tokens)
        pid = C.c_int32()
        self._check(self.lib.advspec_prefill(self.h, _p(t, C.c_int32), t.size, C.byref(pid)))
        return pid.value

    def prefill_extend(self, prefix_id: int, keep_tokens: int, tail: Sequence[int]) -> int:
        """Keep the first `keep_tokens` tokens' KV of a live prefix and prefill `tail` after them."""
        t = _i32(tail) if len(tail) else np.zeros(1, dtype=np.int32)
        pid = C.c_int32()
        self._check(self.lib.advspec_prefill_extend(self.h, prefix_id, keep_tokens, _p(t, C.c_int32), len(tail),
                                                    C.byref(pid)))
        return pid.value

    def fork(self, prefix_id: int, seeds: Sequence[int]) -> list[int]:
        s = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        ids = np.zeros(len(seeds), dtype=np.int32)
        self._check(self.lib.advspec_fork(self.h, prefix_id, len(seeds), _p(s, C.c_uint64), _p(ids, C.c_int32)))
        return ids.tolist()

    def decode(self, seq_ids: Sequence[int], max_new: int, temperature: float = 0.7,
               eos_id: int = -1) -> DecodeResult:
        ids = _i32(seq_ids)
        out = np.full((len(ids), max_new), -1, dtype=np.int32)
        lens = np.zeros(len(ids), dtype=np.int32)
        self._check(self.lib.advspec_decode(self.h, _p(ids, C.c_int32), len(ids), max_new,
                                            C.c_float(temperature), eos_id, _p(out, C.c_int32),
                                            _p(lens, C.c_int32)))
        return DecodeResult([out[i, : lens[i]].tolist() for i in range(len(ids))], lens.tolist())

    def append_tail(self, seq_id: int, tokens: Sequence[int]) -> None:
        """Continue the shared prefix with `tokens` for ONE freshly forked opponent (per-opponent prompt tail)."""
        t = _i32(tokens)
        self._check(self.lib.advspec_append_tail(self.h, int(seq_id), _p(t, C.c_int32), t.size))

    def decode_step(self, seq_ids: Sequence[int], forced_tokens: Sequence[int]) -> None:
        ids, f = _i32(seq_ids), _i32(forced_tokens)
        self._check(self.lib.advspec_decode_step(self.h, _p(ids, C.c_int32), len(ids), _p(f, C.c_int32)))

    def get_logits(self, n: int = 1) -> np.ndarray:
        out = np.zeros((n, self.vocab_local), dtype=np.float32)
        self._check(self.lib.advspec_get_logits(self.h, n, _p(out, C.c_float)))
        return out

    def prefill_logits(self, tokens: Sequence[int]) -> np.ndarray:
        t = _i32(tokens)
        out = np.zeros((t.size, self.vocab_local), dtype=np.float32)
        self._check(self.lib.advspec_prefill_logits(self.h, _p(t, C.c_int32), t.size, _p(out, C.c_float)))
        return out

    def release_seqs(self, seq_ids: Sequence[int]) -> None:
        ids = _i32(seq_ids)
        self._check(self.lib.advspec_release_seqs(self.h, _p(ids, C.c_int32), len(ids)))

    def release_prefix(self, prefix_id: int) -> None:
        self._check(self.lib.advspec_release_prefix(self.h, prefix_id))

    # -- multi-GPU replicas --------------------------------------------------
    def prefix_kv_region(self, prefix_id: int) -> tuple[int, int]:
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.advspec_prefix_kv_region(self.h, prefix_id, C.byref(ptr), C.byref(n)))
        return int(ptr.value), int(n.value)

    def prefix_adopt(self, n_tokens: int, logits: np.ndarray) -> int:
        lg = np.ascontiguousarray(logits, dtype=np.float32).reshape(-1)
        pid = C.c_int32()
        self._check(self.lib.advspec_prefix_adopt(self.h, n_tokens, _p(lg, C.c_float), C.byref(pid)))
        return pid.value

    # -- measurement -----------------------------------------------------------
    def timing(self) -> Timing:
        t = Timing()
        self._check(self.lib.advspec_get_timing(self.h, C.byref(t)))
        return t

    def profile_decode_step(self, seq_ids: Sequence[int]) -> Timing:
        ids = _i32(seq_ids)
        self._check(self.lib.advspec_profile_decode_step(self.h, _p(ids, C.c_int32), len(ids)))
        return self.timing()

    def ktrace_enable(self, on: bool) -> None:
        self._check(self.lib.advspec_ktrace_enable(self.h, 1 if on else 0))

    def ktrace_read(self, cap: int = 8192) -> list[tuple[int, int]]:
        """[(timestamp_ns, kind)] of the decode kernels launched since the last read."""
        buf = np.zeros(cap, dtype=np.uint64)
        n = C.c_int32()
        self._check(self.lib.advspec_ktrace_read(self.h, _p(buf, C.c_uint64), cap, C.byref(n)))
        return [(int(v) >> 4, int(v) & 15) for v in buf[: n.value]]

    def ktrace_phases(self) -> list[int]:
        buf = np.zeros(16, dtype=np.uint64)
        self._check(self.lib.advspec_ktrace_phases(self.h, _p(buf, C.c_uint64)))
        return [int(v) for v in buf]

    def decode_step_bytes(self, seq_ids: Sequence[int]) -> tuple[float, float]:
        ids = _i32(seq_ids)
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.advspec_decode_step_bytes(self.h, _p(ids, C.c_int32), len(ids), C.byref(a),
                                                       C.byref(b)))
        return a.value, b.value
