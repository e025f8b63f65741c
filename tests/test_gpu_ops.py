"""Op-level parity of the sm_100a kernels through the C ABI (device pointers from torch)."""

import ctypes as C

import numpy as np
import pytest
import torch

from advspec_b200 import engine as eng

pytestmark = pytest.mark.gpu


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _err(lib):
    return (lib.advspec_last_error(None) or b"").decode()


def _act(x, act):
    return torch.nn.functional.gelu(x, approximate="tanh") if act == 1 else torch.nn.functional.silu(x)


def _gemm_ref(A, B, epi, act, bias, Cin):
    acc = A.float() @ B.float().T
    if epi == 0:
        return (acc + (bias if bias is not None else 0)).bfloat16()
    if epi == 1:
        return Cin + acc
    if epi == 2:
        return (_act(acc[:, 0::2], act) * acc[:, 1::2]).bfloat16()
    return acc


# the last two are the benchmark's own o-proj and down-proj (5,068 prompt rows): 640 tiles = 4.3 waves of 148,
# so the fp32 epilogues cut the last wave's 48 tiles three ways along K, and K = 14,336 walks A in 3 bands
GEMM_SHAPES = [(128, 256, 64), (256, 512, 512), (384, 768, 320), (200, 1000, 264), (70, 128, 128),
               (1024, 6144, 4096), (5068, 4096, 4096), (5068, 4096, 14336)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_tcgen05(cuda_device, diag, M, N, K, epi):
    lib = eng.load_library()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) if epi == 0 else None
    act = (M // 128) % 2
    if epi in (1, 3):
        Cbuf = torch.randn(M, N, device="cuda", generator=g)
    elif epi == 2:
        Cbuf = torch.zeros(M, N // 2, device="cuda", dtype=torch.bfloat16)
    else:
        Cbuf = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ref = _gemm_ref(A, B, epi, act, bias, Cbuf.clone() if epi == 1 else None)
    chk = Cbuf.clone()
    st = lib.advspec_op_gemm(0, _ptr(A), K, _ptr(B), K, _ptr(Cbuf), Cbuf.shape[1], _ptr(bias), M, N, K, epi, act)
    assert st == 0, _err(lib)
    st = lib.advspec_op_gemm_check(0, _ptr(A), K, _ptr(B), K, _ptr(chk), chk.shape[1], _ptr(bias), M, N, K, epi, act)
    assert st == 0, _err(lib)
    scale = float(ref.float().abs().max()) + 1e-6
    e_tc = float((Cbuf.float() - ref.float()).abs().max()) / scale
    e_ck = float((chk.float() - ref.float()).abs().max()) / scale
    diag[f"gemm/{M}x{N}x{K}/epi{epi}"] = {"tcgen05_rel_err": e_tc, "check_rel_err": e_ck}
    # fp32 accumulate of bf16 products; outputs bf16-rounded for epi 0/2 (2^-8 relative)
    tol = 1e-2 if epi in (0, 2) else 2e-3
    assert e_ck < tol, f"check kernel off by {e_ck}"
    assert e_tc < tol, f"tcgen05 kernel off by {e_tc} (check kernel {e_ck})"


@pytest.mark.parametrize("N,K", [(512, 256), (6144, 4096), (4096, 14336), (1000, 3584), (2048, 18944)])
@pytest.mark.parametrize("b", [1, 3, 5, 8])
@pytest.mark.parametrize("mode", [(0, 0), (1, 0), (0, 1), (1, 2), (1, 3)])
def test_gemv(cuda_device, diag, N, K, b, mode):
    in_mode, epi = mode
    lib = eng.load_library()
    g = torch.Generator(device="cuda").manual_seed(N + K + b)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    act, eps = 0, 1e-5
    if in_mode == 0:
        x = torch.randn(b, K, device="cuda", generator=g).bfloat16()
        xin, nw = x.float(), None
    else:
        x = torch.randn(b, K, device="cuda", generator=g) * 3.0
        nw = 1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)
        xin = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * nw).bfloat16().float()
    acc = xin @ W.float().T
    bias = torch.randn(N, device="cuda", generator=g) if epi == 0 else None
    if epi == 0:
        y = torch.zeros(b, N, device="cuda", dtype=torch.bfloat16)
        ref = (acc + bias).bfloat16()
    elif epi == 1:
        y = torch.randn(b, N, device="cuda", generator=g)
        ref = y + acc
    elif epi == 2:
        y = torch.zeros(b, N // 2, device="cuda", dtype=torch.bfloat16)
        ref = (_act(acc[:, 0::2], act) * acc[:, 1::2]).bfloat16()
    else:
        y = torch.zeros(b, N, device="cuda")
        ref = acc
    st = lib.advspec_op_gemv(0, _ptr(W), _ptr(x), _ptr(nw), _ptr(bias), _ptr(y), b, N, K, in_mode, epi, act,
                             C.c_float(eps))
    assert st == 0, _err(lib)
    scale = float(ref.float().abs().max()) + 1e-6
    e = float((y.float() - ref.float()).abs().max()) / scale
    diag[f"gemv/{N}x{K}/b{b}/in{in_mode}/epi{epi}"] = e
    assert e < (1e-2 if epi in (0, 2) else 2e-3), e


@pytest.mark.parametrize("M,N,K,epi", [(5068, 4096, 14336, 1), (5068, 4096, 4096, 3), (2100, 3584, 18944, 1),
                                       (9000, 8192, 3584, 1)])
def test_gemm_k_split_tail_is_deterministic_and_schedule_independent(cuda_device, diag, monkeypatch, M, N, K, epi):
    """The K-split of the last partial wave (long K only) adds its partial tiles in split order: two runs give
    the same bits, the banded raster gives the same bits, and the unsplit schedule differs only by fp32
    summation order."""
    lib = eng.load_library()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    C0 = torch.randn(M, N, device="cuda", generator=g)

    def run():
        Cb = C0.clone()
        st = lib.advspec_op_gemm(0, _ptr(A), K, _ptr(B), K, _ptr(Cb), N, None, M, N, K, epi, 0)
        assert st == 0, _err(lib)
        return Cb

    a, b = run(), run()
    assert torch.equal(a, b), "same schedule, same bits"
    monkeypatch.setenv("ADVSPEC_GEMM_SPLITK", "0")
    plain = run()
    monkeypatch.delenv("ADVSPEC_GEMM_SPLITK")
    monkeypatch.setenv("ADVSPEC_GEMM_BAND_MB", "48")
    banded = run()
    monkeypatch.delenv("ADVSPEC_GEMM_BAND_MB")
    # (another raster puts other tiles in the K-split tail: same values up to fp32 summation order)
    assert float((banded - a).abs().max()) / float(a.abs().max()) < 1e-4
    run()  # restores the defaults inside the library (the knobs are re-read per call)
    ref = A.float() @ B.float().T + (C0 if epi == 1 else 0)
    scale = float(ref.abs().max())
    e_split, e_plain = float((a - ref).abs().max()) / scale, float((plain - ref).abs().max()) / scale
    diag[f"gemm_ksplit/{M}x{N}x{K}/epi{epi}"] = {"split_rel_err": e_split, "plain_rel_err": e_plain,
                                                 "split_vs_plain": float((a - plain).abs().max()) / scale}
    assert e_split < 2e-3 and e_plain < 2e-3, (e_split, e_plain)


def _attn_ref(q, kc, vc, n_q, q_pos0, H, Hkv, Dh):
    # q [n_q, H*Dh(+...)] bf16; kc/vc [Hkv, stride, Dh]
    total = q_pos0 + n_q
    qh = q[:, : H * Dh].float().reshape(n_q, H, Dh)
    k = kc[:, :total].float().repeat_interleave(H // Hkv, dim=0)  # [H, total, Dh]
    v = vc[:, :total].float().repeat_interleave(H // Hkv, dim=0)
    sc = torch.einsum("qhd,hkd->hqk", qh, k) / (Dh ** 0.5)
    qpos = q_pos0 + torch.arange(n_q, device=q.device)
    mask = torch.arange(total, device=q.device)[None, :] > qpos[:, None]
    sc = sc.masked_fill(mask[None], float("-inf"))
    return torch.einsum("hqk,hkd->qhd", sc.softmax(-1), v).reshape(n_q, H * Dh)


@pytest.mark.parametrize("cfg", [(64, 0, 4, 2, 64), (200, 0, 4, 2, 128), (333, 100, 8, 2, 128),
                                 (1024, 0, 8, 8, 128), (130, 62, 2, 1, 64), (96, 0, 2, 2, 256),
                                 (128, 0, 2, 2, 128), (640, 256, 4, 1, 128), (4096, 0, 8, 2, 128),
                                 (300, 40, 4, 4, 96), (700, 0, 2, 2, 96), (200, 0, 2, 1, 256)])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_attn_prefill(cuda_device, diag, cfg, impl):
    n_q, q_pos0, H, Hkv, Dh = cfg
    if impl == 2 and Dh == 256:
        pytest.skip("the tcgen05 attention kernel serves head_dim 64, 96 and 128")
    lib = eng.load_library()
    g = torch.Generator(device="cuda").manual_seed(n_q + q_pos0 + H)
    stride = q_pos0 + n_q + 37
    ldq = (H + 2 * Hkv) * Dh
    q = torch.randn(n_q, ldq, device="cuda", generator=g).bfloat16()
    kc = torch.randn(Hkv, stride, Dh, device="cuda", generator=g).bfloat16()
    vc = torch.randn(Hkv, stride, Dh, device="cuda", generator=g).bfloat16()
    out = torch.zeros(n_q, H * Dh, device="cuda", dtype=torch.bfloat16)
    st = lib.advspec_op_attn_prefill(0, _ptr(q), ldq, _ptr(kc), _ptr(vc), stride, _ptr(out), n_q, q_pos0, H,
                                     Hkv, Dh, impl)
    assert st == 0, _err(lib)
    ref = _attn_ref(q, kc, vc, n_q, q_pos0, H, Hkv, Dh)
    e = float((out.float() - ref).abs().max())
    diag[f"attn_prefill/{cfg}/impl{impl}"] = e
    # outputs are O(1); bf16 output rounding + bf16 P in the tensor-core kernel
    assert e < 3e-2, e


def _rope(x, cos, sin):
    """rotate_half RoPE on [..., Dh] with per-row cos/sin [..., Dh/2] (fp32)."""
    half = x.shape[-1] // 2
    a, b = x[..., :half], x[..., half:]
    return torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)


# (prefix_len, b, H, Hkv, Dh, suffix lengths already stored per opponent): the benchmark's own
# configuration first (5,068-token prefix, 3 opponents, Llama-3-8B heads), then Qwen2 (G = 7: two
# opponent groups per KV head at b = 3), Phi-3 (MHA, head_dim 96 in the padded 128-wide tile), Gemma
# (MHA, head_dim 256), the full batch, a tensor-parallel rank's share of Llama-3-70B (one KV head, G = 8),
# and suffix lengths around the 64-key tile boundary.
ATTN_DECODE_CASES = [
    (5068, 3, 32, 8, 128, [0, 5, 17]),
    (5068, 3, 32, 8, 128, [63, 64, 65]),
    (5068, 8, 32, 8, 128, [1, 63, 64, 65, 0, 130, 200, 255]),
    (5068, 1, 32, 8, 128, [255]),
    (2100, 3, 28, 4, 128, [3, 64, 100]),
    (2100, 3, 32, 32, 96, [1, 63, 65]),
    (2100, 8, 32, 32, 96, [0, 1, 2, 3, 64, 65, 66, 127]),
    (2100, 3, 16, 16, 256, [0, 64, 129]),
    (9000, 1, 8, 1, 128, [31]),
    (300, 2, 4, 2, 64, [0, 70]),
    (100, 3, 8, 2, 128, [5, 5, 5]),
]


@pytest.mark.parametrize("case", ATTN_DECODE_CASES)
def test_attn_decode_op(cuda_device, diag, case):
    """advspec_op_attn_decode (RoPE of the new q/k, KV append, split-KV attention over the shared prefix and
    each opponent's suffix, merge) against an fp32 torch reference at the engine's production shapes."""
    prefix_len, b, H, Hkv, Dh, suf = case
    lib = eng.load_library()
    g = torch.Generator(device="cuda").manual_seed(prefix_len + 31 * b + H + Dh)
    pstride, sstride = prefix_len + 91, 320
    half = Dh // 2
    QKV = (H + 2 * Hkv) * Dh
    qkv = torch.randn(b, QKV, device="cuda", generator=g).bfloat16()
    pk = torch.randn(Hkv, pstride, Dh, device="cuda", generator=g).bfloat16()
    pv = torch.randn(Hkv, pstride, Dh, device="cuda", generator=g).bfloat16()
    pk[:, prefix_len:] = 0  # rows the engine never wrote hold zeros (TMA loads whole 64-key boxes)
    pv[:, prefix_len:] = 0
    sk = torch.zeros(b, Hkv, sstride, Dh, device="cuda", dtype=torch.bfloat16)
    sv = torch.zeros_like(sk)
    for i, t in enumerate(suf):
        sk[i, :, :t] = torch.randn(Hkv, t, Dh, device="cuda", generator=g).bfloat16()
        sv[i, :, :t] = torch.randn(Hkv, t, Dh, device="cuda", generator=g).bfloat16()
    max_pos = prefix_len + sstride
    inv = 1.0 / (10000.0 ** (torch.arange(0, Dh, 2, device="cuda").float() / Dh))
    ang = torch.arange(max_pos, device="cuda").float()[:, None] * inv[None, :]
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    pos = np.asarray([prefix_len + t for t in suf], dtype=np.int32)
    out = torch.zeros(b, H * Dh, device="cuda", dtype=torch.bfloat16)
    sk_before, sv_before = sk.clone(), sv.clone()

    st = lib.advspec_op_attn_decode(0, _ptr(qkv), _ptr(cos), _ptr(sin), _ptr(pk), _ptr(pv), pstride, prefix_len,
                                    _ptr(sk), _ptr(sv), sstride, pos.ctypes.data_as(C.POINTER(C.c_int32)),
                                    _ptr(out), b, H, Hkv, Dh)
    assert st == 0, _err(lib)

    worst, worst_k = 0.0, 0.0
    for i, t in enumerate(suf):
        c, s = cos[prefix_len + t], sin[prefix_len + t]
        q = _rope(qkv[i, : H * Dh].float().reshape(H, Dh), c, s).bfloat16().float()
        knew = _rope(qkv[i, H * Dh: (H + Hkv) * Dh].float().reshape(Hkv, Dh), c, s).bfloat16()
        vnew = qkv[i, (H + Hkv) * Dh:].reshape(Hkv, Dh)
        # side effect: the new k (rotated) and v rows are appended at suffix row t, nothing else changes
        worst_k = max(worst_k, float((sk[i, :, t].float() - knew.float()).abs().max()))
        assert torch.equal(sv[i, :, t], vnew)
        assert torch.equal(sk[i, :, :t], sk_before[i, :, :t]) and torch.equal(sv[i, :, :t], sv_before[i, :, :t])
        k = torch.cat([pk[:, :prefix_len], sk_before[i, :, :t], knew[:, None]], dim=1).float()
        v = torch.cat([pv[:, :prefix_len], sv_before[i, :, :t], vnew[:, None]], dim=1).float()
        k = k.repeat_interleave(H // Hkv, dim=0)
        v = v.repeat_interleave(H // Hkv, dim=0)
        sc = torch.einsum("hd,hkd->hk", q, k) / (Dh ** 0.5)
        ref = torch.einsum("hk,hkd->hd", sc.softmax(-1), v).reshape(H * Dh)
        worst = max(worst, float((out[i].float() - ref).abs().max()))
    diag[f"attn_decode_op/{case}"] = {"max_abs": worst, "k_append_max_abs": worst_k}
    # outputs are averages of N(0,1) values over thousands of keys (|o| ~ 0.05-1); bf16 P and bf16 output
    assert worst_k < 2e-2, worst_k
    assert worst < 1e-2, worst
