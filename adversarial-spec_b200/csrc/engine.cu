// engine.cu — the C ABI of include/advspec_engine.h: one model's weights, one
// shared-prefix KV region and up to 8 forked opponents on one B200.
//
// Path (stands in for the N concurrent `completion` calls of the reference,
// skills/adversarial-spec/scripts/models.py:628 under :681-722):
//   prefill  : embed -> L x [rmsnorm, QKV GEMM(tcgen05), RoPE+KV write, causal
//              attention, O GEMM(+residual), rmsnorm, gate/up GEMM(+act*up),
//              down GEMM(+residual)] -> final norm + lm_head of the last row
//   fork     : N opponent slots over the one prefix KV, no copy
//   decode   : CUDA-graph-replayed step: L x [norm+QKV GEMV, RoPE+KV append,
//              split-KV attention (prefix shared, suffix private), combine,
//              O GEMV(+res), norm+gate/up GEMV(+act*up), down GEMV(+res)]
//              -> norm+lm_head GEMV -> Gumbel-max sample + next embedding
// Tensor parallelism (one large opponent over tp_size GPUs, one process per GPU): a rank is the
// same engine over a NARROWER model — its share of the query/KV heads, of the MLP columns and of
// the vocabulary — plus three exchange points: an all-reduce of the residual stream after the
// o-proj and after the down-proj (rank 0's partial carries the residual, so the sum IS the new
// residual), and an all-gather of the per-rank sampler winners.  NCCL is bound at run time.
// No CPU fallback exists: every compute entry point needs a CUDA device.
#include "../../include/advspec_engine.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is dlopen'ed (advspec_tp_*)

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "attn.cuh"
#include "attn_decode_mma.cuh"
#include "attn_prefill_tc.cuh"
#include "common.cuh"
#include "decode_kernels.cuh"
#include "gemm_tcgen05.cuh"
#include "gemv_mma.cuh"
#include "tp_allreduce.cuh"

using namespace advspec;

namespace {

thread_local std::string g_create_error;

// NCCL bound at run time: a process that never asks for tensor parallelism never needs the library.
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string err;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
bool nccl_load(std::string* why) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return true;
  const char* env = getenv("ADVSPEC_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    if (why) *why = std::string("cannot load NCCL (set ADVSPEC_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
    return false;
  }
  auto sym = [&](const char* n) { return dlsym(h, n); };
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce ||
      !g_nccl.AllGather || !g_nccl.GetErrorString) {
    if (why) *why = "the NCCL library lacks a required symbol";
    return false;
  }
  g_nccl.lib = h;
  return true;
}

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v, size_t a = kAlign) { return (v + a - 1) / a * a; }

struct BlobLayout {
  struct LayerOff {
    size_t attn_norm, wqkv, bqkv, wo, mlp_norm, wgu, wd;
  };
  size_t embed = 0, final_norm = 0, lm_head = 0, total = 0;
  std::vector<LayerOff> layers;
};

inline int qkv_dim(const advspec_model_desc& d) { return (d.n_heads + 2 * d.n_kv_heads) * d.head_dim; }

// The shape one tensor-parallel rank computes with: its share of the heads, of the MLP width and of
// the vocabulary (lm_head rows); d_model, layers and the embedding table (v_full rows) are whole.
struct LocalDesc {
  advspec_model_desc d;
  int v_full;
};
LocalDesc localize(const advspec_model_desc& full) {
  LocalDesc r{full, full.vocab_size};
  const int tp = full.tp_size > 1 ? full.tp_size : 1;
  r.d.n_heads = full.n_heads / tp;
  r.d.n_kv_heads = full.n_kv_heads / tp;
  r.d.d_ff = full.d_ff / tp;
  r.d.vocab_size = full.vocab_size / tp;
  return r;
}

// Order and alignment restated in adversarial-spec_b200/weights.py (checked by tests).
BlobLayout make_layout(const advspec_model_desc& d, int v_full) {
  BlobLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes);
    return o;
  };
  const size_t dm = d.d_model, qkv = qkv_dim(d), hd = (size_t)d.n_heads * d.head_dim;
  L.embed = take((size_t)v_full * dm * 2);
  L.layers.resize(d.n_layers);
  for (int l = 0; l < d.n_layers; ++l) {
    auto& o = L.layers[l];
    o.attn_norm = take(dm * 4);
    o.wqkv = take(qkv * dm * 2);
    o.bqkv = d.qkv_bias ? take(qkv * 4) : (size_t)-1;
    o.wo = take(dm * hd * 2);
    o.mlp_norm = take(dm * 4);
    o.wgu = take((size_t)2 * d.d_ff * dm * 2);
    o.wd = take(dm * (size_t)d.d_ff * 2);
  }
  L.final_norm = take(dm * 4);
  L.lm_head = d.tied_lm_head ? (size_t)-1 : take((size_t)d.vocab_size * dm * 2);
  L.total = off;
  return L;
}

bool desc_ok(const advspec_model_desc* d, std::string* why) {
  auto bad = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (!d) return bad("null model desc");
  if (d->abi_version != ADVSPEC_ABI_VERSION) return bad("abi_version mismatch");
  if (d->n_layers < 1 || d->d_model < 8 || d->n_heads < 1 || d->n_kv_heads < 1 || d->d_ff < 8 ||
      d->vocab_size < 2)
    return bad("non-positive model dimension");
  if (d->n_heads % d->n_kv_heads) return bad("n_heads must be a multiple of n_kv_heads");
  if (d->head_dim != 64 && d->head_dim != 96 && d->head_dim != 128 && d->head_dim != 256)
    return bad("head_dim must be 64, 96, 128 or 256");
  if (d->d_model % 8 || d->d_ff % 8 || (d->n_heads * d->head_dim) % 8)
    return bad("d_model, d_ff and n_heads*head_dim must be multiples of 8");
  if (d->max_seqs < 1 || d->max_seqs > 8) return bad("max_seqs must be in 1..8");
  if (d->max_prefix_tokens < 1 || d->max_new_tokens < 1) return bad("bad KV capacities");
  if (d->tp_size != 1) {
    const int tp = d->tp_size;
    if (tp != 2 && tp != 4 && tp != 8) return bad("tp_size must be 1, 2, 4 or 8");
    if (d->tp_rank < 0 || d->tp_rank >= tp) return bad("tp_rank outside 0..tp_size-1");
    if (d->n_heads % tp || d->n_kv_heads % tp) return bad("tp_size must divide n_heads and n_kv_heads");
    if (d->d_ff % (8 * tp) || d->vocab_size % tp) return bad("tp_size must divide d_ff/8 and vocab_size");
    if (d->tied_lm_head) return bad("tensor parallelism with a tied lm_head is not supported");
  } else if (d->tp_rank != 0) {
    return bad("tp_rank must be 0 when tp_size is 1");
  }
  if (d->act != 0 && d->act != 1) return bad("act must be 0 (SiLU) or 1 (tanh GELU)");
  return true;
}

// ------------------------------------------------------------------ TMA maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// bf16 [rows][cols] row-major with row pitch ld (elements); box = 64 cols x box_rows, 128B swizzle.
bool make_tmap(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int g_num_sms = 0;
int num_sms(int device) {
  if (g_num_sms == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}

template <typename K>
cudaError_t set_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

int g_gemm_band_mb = 0;    // ADVSPEC_GEMM_BAND_MB=<n>: raster bands whose A rows total n MB (0 = one band; measured: 48 MB
                           // bands cut the down-proj's DRAM reads from 963 to 864 MB but not its time)
int g_gemm_splitk = 1;     // ADVSPEC_GEMM_SPLITK=0 turns the K-split of the last partial wave off
constexpr int kGemmSemWords = 256;  // ordering words a caller provides for the K-split tail

// Tile order and tail split of one launch (see gemm_tcgen05.cuh): bands of m-tiles whose A rows stay in L2;
// when the last wave of the persistent grid is partial and the epilogue accumulates in fp32, its tiles are
// cut along K among the idle SMs.
void gemm_schedule(GemmParams* p, int BN, int epi, int grid, unsigned int* sem) {
  const int num_m = (p->M + kGemmBM - 1) / kGemmBM, num_n = (p->N + BN - 1) / BN;
  const int num_kb = (p->K + kGemmBK - 1) / kGemmBK;
  const int tiles = num_m * num_n;
  int band = num_m;
  if (g_gemm_band_mb > 0) {
    const int64_t per_tile = (int64_t)kGemmBM * p->K * 2;
    const int bm_max = (int)std::max<int64_t>(4, ((int64_t)g_gemm_band_mb << 20) / std::max<int64_t>(per_tile, 1));
    const int n_bands = (num_m + bm_max - 1) / bm_max;
    band = (num_m + n_bands - 1) / n_bands;
  }
  p->band_m = std::max(1, band);
  p->full_items = tiles;
  p->split = 1;
  p->sem = nullptr;
  const int tail = tiles % grid;
  if (g_gemm_splitk && sem && (epi == EPI_RESADD_F32 || epi == EPI_F32) && tiles > grid && tail > 0 &&
      tail <= kGemmSemWords) {
    // a split must keep a long K loop (>= 48 k-blocks = 3,072 columns): shorter items are all pipeline fill and
    // drain plus serialised epilogues (measured: the o-proj, K = 4,096, went from 167 to 207 us with 21-block splits)
    const int split = std::min(std::min(grid / tail, 4), num_kb / 48);
    if (split >= 2) {
      p->full_items = tiles - tail;
      p->split = split;
      p->sem = sem;
    }
  }
  p->total_items = p->full_items + (tiles - p->full_items) * p->split;
}

template <int BN, int EPI>
cudaError_t launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int device,
                          cudaStream_t st, unsigned int* sem) {
  auto kern = gemm_tc_kernel<BN, EPI>;
  {  // function attributes are per device: set on every launch (host-side, microseconds)
    cudaError_t e = set_smem(kern, GemmCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
  }
  const int tiles = ((p.M + kGemmBM - 1) / kGemmBM) * ((p.N + BN - 1) / BN);
  const int grid = std::min(tiles, num_sms(device));
  gemm_schedule(&p, BN, EPI, grid, sem);
  kern<<<grid, kGemmThreads, GemmCfg<BN>::kSmemBytes, st>>>(ta, tb, p);
  return cudaGetLastError();
}

bool g_gemm_narrow = false;  // ADVSPEC_GEMM_NARROW=1 (experiment)
bool g_attn_prefill_tc = true;  // ADVSPEC_ATTN_PREFILL_TC=0 falls back to the mma.sync kernel (A/B)

// C = A[M,K] * B[N,K]^T on tcgen05.  A rows / B rows are the TMA extents.
cudaError_t launch_gemm(const void* A, int64_t lda, int64_t a_rows, const void* B, int64_t ldb,
                        const GemmParams& p, int epi, int device, cudaStream_t st, std::string* err,
                        unsigned int* sem = nullptr) {
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15)) {
    if (err) *err = "gemm: operands must be 16-byte aligned with pitches that are multiples of 8";
    return cudaErrorInvalidValue;
  }
  // Tile width: the persistent grid runs ceil(tiles / SMs) waves of tiles whose time is ~ BN, so
  // pick the BN in {128, 256} with the smaller waves x BN product (wave quantisation costs up to 14 %
  // on the N = 4096 / 6144 matrices with BN = 256).
  bool wide = p.N > 128;
  if (wide && g_gemm_narrow) {  // ADVSPEC_GEMM_NARROW=1; measured SLOWER in round 1 (prefill 94 -> 103 ms), off by default
    const int sms = num_sms(device);
    const int64_t mt = (p.M + kGemmBM - 1) / kGemmBM;
    const int64_t t256 = mt * ((p.N + 255) / 256), t128 = mt * ((p.N + 127) / 128);
    const int64_t c256 = (t256 + sms - 1) / sms * 256, c128 = (t128 + sms - 1) / sms * 128;
    wide = c256 <= c128 + c128 / 32;  // prefer the wide tile unless the narrow one wins by > 3 %
  }
  CUtensorMap ta, tb;
  if (!make_tmap(&ta, A, a_rows, p.K, lda, kGemmBM) ||
      !make_tmap(&tb, B, p.N, p.K, ldb, wide ? 256 : 128)) {
    if (err) *err = "cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
#define ADV_GEMM_CASE(E)                                                   \
  case E:                                                                  \
    return wide ? launch_gemm_t<256, E>(ta, tb, p, device, st, sem)       \
                : launch_gemm_t<128, E>(ta, tb, p, device, st, sem);
  switch (epi) {
    ADV_GEMM_CASE(EPI_BF16)
    ADV_GEMM_CASE(EPI_RESADD_F32)
    ADV_GEMM_CASE(EPI_GATED_BF16)
    ADV_GEMM_CASE(EPI_F32)
  }
#undef ADV_GEMM_CASE
  if (err) *err = "gemm: unknown epilogue";
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemm_check(const void* A, int64_t lda, const void* B, int64_t ldb,
                              const GemmParams& p, int epi, cudaStream_t st) {
  const int ncols = (epi == EPI_GATED_BF16) ? p.N / 2 : p.N;
  const int64_t total = (int64_t)p.M * ncols;
  const int grid = (int)((total + 255) / 256);
  auto a = reinterpret_cast<const __nv_bfloat16*>(A);
  auto b = reinterpret_cast<const __nv_bfloat16*>(B);
  switch (epi) {
    case EPI_BF16: gemm_check_kernel<EPI_BF16><<<grid, 256, 0, st>>>(a, lda, b, ldb, p); break;
    case EPI_RESADD_F32: gemm_check_kernel<EPI_RESADD_F32><<<grid, 256, 0, st>>>(a, lda, b, ldb, p); break;
    case EPI_GATED_BF16: gemm_check_kernel<EPI_GATED_BF16><<<grid, 256, 0, st>>>(a, lda, b, ldb, p); break;
    case EPI_F32: gemm_check_kernel<EPI_F32><<<grid, 256, 0, st>>>(a, lda, b, ldb, p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------- GEMV
bool g_use_pdl = true;
bool g_trace = false;
// ADVSPEC_TRACE=1: synchronise after every decode-path launch and name it on stderr (debugging aid)
#define ADV_TRACE(st, name)                                                      \
  do {                                                                           \
    if (g_trace) {                                                               \
      cudaError_t _te = cudaStreamSynchronize(st);                               \
      fprintf(stderr, "[advspec trace] %s -> %s\n", name, cudaGetErrorString(_te)); \
      fflush(stderr);                                                            \
    }                                                                            \
  } while (0)

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && g_use_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

int g_gemv_impl = 3;  // 3: bulk-async stream + tensor-core consumers (default); 1: register loads (fallback, A/B)

cudaError_t launch_gemv_v1(const GemvParams& p, int b, int device, cudaStream_t st, bool pdl) {
  const int pairs = (p.N + 1) / 2;
  const int grid = std::max(1, std::min(2 * num_sms(device), pairs));
  dim3 g(grid), blk(kGemvThreads);
  switch (b) {
    case 1: return launch_pdl(gemv_kernel<1, 2>, g, blk, 0, st, pdl, p);
    case 2: return launch_pdl(gemv_kernel<2, 2>, g, blk, 0, st, pdl, p);
    case 3: return launch_pdl(gemv_kernel<3, 2>, g, blk, 0, st, pdl, p);
    case 4: return launch_pdl(gemv_kernel<4, 2>, g, blk, 0, st, pdl, p);
    case 5: return launch_pdl(gemv_kernel<5, 1>, g, blk, 0, st, pdl, p);
    case 6: return launch_pdl(gemv_kernel<6, 1>, g, blk, 0, st, pdl, p);
    case 7: return launch_pdl(gemv_kernel<7, 1>, g, blk, 0, st, pdl, p);
    case 8: return launch_pdl(gemv_kernel<8, 1>, g, blk, 0, st, pdl, p);
  }
  return cudaErrorInvalidValue;
}

bool g_tp_ll = true;  // ADVSPEC_TP_AR=flag: the push / fence / flag / sum exchange instead of low-latency packets (A/B)
int g_attn_min_split = 256;  // ADVSPEC_ATTN_MIN_SPLIT: fewest prefix tokens worth a split of their own
size_t g_x_smem_max = 40000;  // ADVSPEC_X_SMEM_MAX: larger plain-bf16 inputs are read through L1 from global (measured faster)

template <int B>
cudaError_t launch_gemv_mma_t(const GemvParams& p, int device, cudaStream_t st, bool pdl) {
  constexpr int kMaxDyn = 232448 - 9 * 1024;  // 227 KB per CTA minus the kernel's 8.5 KB of static shared memory
  {  // function attributes are per device: set on every launch (host-side, microseconds)
    cudaError_t e = set_smem(gemv_mma_kernel<B>, kMaxDyn);
    if (e != cudaSuccess) return e;
  }
  const size_t xbytes = ((size_t)B * ((size_t)p.K * 2 + 16) + 127) / 128 * 128;
  int x_in_smem = 0;
  size_t x_smem = 0;
  if (p.in_mode == 1) {
    x_smem = xbytes;
  } else if (xbytes + 2 * (size_t)kGmStageBytes <= (size_t)kMaxDyn && xbytes <= g_x_smem_max) {
    x_in_smem = 1;
    x_smem = xbytes;
  }
  if (x_smem + 2 * (size_t)kGmStageBytes > (size_t)kMaxDyn) return cudaErrorInvalidValue;
  const int stages = std::min<int>(kGmMaxStages, (int)((kMaxDyn - x_smem) / kGmStageBytes));
  const int pairs = (p.N + 1) / 2;
  const int grid = std::max(1, std::min(num_sms(device), pairs));
  const size_t dyn = (size_t)stages * kGmStageBytes + x_smem;
  return launch_pdl(gemv_mma_kernel<B>, dim3(grid), dim3(kGmThreads), dyn, st, pdl, p, stages, x_in_smem);
}

cudaError_t launch_gemv_mma(const GemvParams& p, int b, int device, cudaStream_t st, bool pdl) {
  switch (b) {
    case 1: return launch_gemv_mma_t<1>(p, device, st, pdl);
    case 2: return launch_gemv_mma_t<2>(p, device, st, pdl);
    case 3: return launch_gemv_mma_t<3>(p, device, st, pdl);
    case 4: return launch_gemv_mma_t<4>(p, device, st, pdl);
    case 5: return launch_gemv_mma_t<5>(p, device, st, pdl);
    case 6: return launch_gemv_mma_t<6>(p, device, st, pdl);
    case 7: return launch_gemv_mma_t<7>(p, device, st, pdl);
    case 8: return launch_gemv_mma_t<8>(p, device, st, pdl);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemv(const GemvParams& p, int b, int device, cudaStream_t st, bool pdl) {
  // the bulk-async + tensor-core kernel serves every shape in the model table; the register-load kernel is the
  // fallback for K % 16 != 0 or a fused-RMSNorm input too large to sit beside a 2-stage ring
  // (ADVSPEC_GEMV_IMPL=1 forces it: the A/B the round-1 profiles refer to)
  const bool ok3 = g_gemv_impl == 3 && (p.K % 16 == 0) &&
                   (p.in_mode != 1 || ((size_t)b * ((size_t)p.K * 2 + 16) + 2 * (size_t)kGmStageBytes + 256 <= (size_t)(232448 - 9 * 1024)));
  if (ok3) return launch_gemv_mma(p, b, device, st, pdl);
  return launch_gemv_v1(p, b, device, st, pdl);
}

// -------------------------------------------------------------- attention
cudaError_t launch_attn_prefill(const AttnPrefillParams& p, int DH, int impl, cudaStream_t st,
                                std::string* err) {
  if (impl == 1 || DH % 8 != 0 || DH > 256) {
    const int64_t warps = (int64_t)p.n_q * p.H;
    const int grid = (int)((warps * 32 + 255) / 256);
    attn_prefill_check_kernel<<<grid, 256, 0, st>>>(p, DH);
    return cudaGetLastError();
  }
  const bool big = p.n_q >= 512;  // 128-row query tiles once there are enough tiles to fill the GPU
  auto launch = [&](auto kern, int bm, int dht) -> cudaError_t {
    const int smem = (bm + 4 * 64) * dht * 2;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    dim3 grid((p.n_q + bm - 1) / bm, p.H);
    kern<<<grid, bm * 2, smem, st>>>(p);
    return cudaSuccess;
  };
  cudaError_t le;
  if (DH <= 64) le = big ? launch(attn_prefill_kernel<64, 128>, 128, 64) : launch(attn_prefill_kernel<64, 64>, 64, 64);
  else if (DH <= 128) le = big ? launch(attn_prefill_kernel<128, 128>, 128, 128) : launch(attn_prefill_kernel<128, 64>, 64, 128);
  else le = launch(attn_prefill_kernel<256, 64>, 64, 256);  // Gemma: 256-wide heads
  if (le != cudaSuccess) return le;
  (void)err;
  return cudaGetLastError();
}

// Prompt attention on tcgen05 (head_dim 64 / 96 / 128): Q from [n_q][ldq] rows, K/V from [Hkv*kv_stride][dh] rows.
cudaError_t launch_attn_prefill_tc(const void* q, int64_t ldq, const void* kc, const void* vc, int64_t kv_stride,
                                   void* out, int n_q, int q_pos0, int H, int Hkv, int dh, cudaStream_t st,
                                   std::string* err) {
  CUtensorMap tq, tk, tv;
  const int64_t kv_rows = (int64_t)Hkv * kv_stride;
  // the Q map spans the whole row (ldq columns): the second 64-column box of the last head may reach into K's columns
  if (!make_tmap(&tq, q, n_q, ldq, ldq, kAtBM) || !make_tmap(&tk, kc, kv_rows, dh, dh, kAtBN) ||
      !make_tmap(&tv, vc, kv_rows, dh, dh, kAtBN)) {
    if (err) *err = "cuTensorMapEncodeTiled failed (attention)";
    return cudaErrorInvalidValue;
  }
  AttnPrefillTcParams p{reinterpret_cast<__nv_bfloat16*>(out), n_q, q_pos0, H, Hkv, (int)kv_stride,
                        1.0f / sqrtf((float)dh), dh};
  dim3 grid((n_q + kAtBM - 1) / kAtBM, H);
  {  // function attributes are per device: set on every launch (host-side, microseconds)
    cudaError_t e = set_smem(attn_prefill_tc_kernel, kAtSmem);
    if (e != cudaSuccess) return e;
    // two CTAs per SM need the whole shared-memory carve-out (the default heuristic sizes it for ONE block)
    e = cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  attn_prefill_tc_kernel<<<grid, kAtThreads, kAtSmem, st>>>(tq, tk, tv, p);
  return cudaGetLastError();
}

cudaError_t launch_attn_decode(const AttnDecodeParams& p, int n_items, int DH, cudaStream_t st, bool pdl) {
  dim3 g(n_items), blk(128);
  switch (DH) {
    case 64: return launch_pdl(attn_decode_kernel<64>, g, blk, 0, st, pdl, p);
    case 96: return launch_pdl(attn_decode_kernel<96>, g, blk, 0, st, pdl, p);
    case 128: return launch_pdl(attn_decode_kernel<128>, g, blk, 0, st, pdl, p);
    case 256: return launch_pdl(attn_decode_kernel<256>, g, blk, 0, st, pdl, p);
  }
  return cudaErrorInvalidValue;
}

int g_attn_impl = 2;  // 2: fused tensor-core decode attention (default); 1: scalar 3-kernel path (A/B, other head dims)

cudaError_t launch_attn_decode2(const AttnDecode2Params& p, int n_ctas, int DH, cudaStream_t st, bool pdl) {
  constexpr int NST = 6;
  dim3 g(n_ctas), blk(256);
  if (DH == 128 || DH == 96) {  // 96 (Phi-3): the 128-wide tile with zero padding, p.dh = 96
    const int smem = 16 * 128 * 2 + NST * 2 * 64 * 128 * 2 + 1024 + 128;  // + alignment slack + barriers
    {  // function attributes are per device: set on every launch (host-side, microseconds)
      cudaError_t e = set_smem(attn_decode_mma_kernel<128, NST>, smem);
      if (e != cudaSuccess) return e;
    }
    return launch_pdl(attn_decode_mma_kernel<128, NST>, g, blk, smem, st, pdl, p);
  }
  if (DH == 256) {  // Gemma: 64 KB per K+V tile pair, three stages
    constexpr int NS3 = 3;
    const int smem = 16 * 256 * 2 + NS3 * 2 * 64 * 256 * 2 + 1024 + 128;
    {
      cudaError_t e = set_smem(attn_decode_mma_kernel<256, NS3>, smem);
      if (e != cudaSuccess) return e;
    }
    return launch_pdl(attn_decode_mma_kernel<256, NS3>, g, blk, smem, st, pdl, p);
  }
  if (DH == 64) {
    const int smem = 16 * 64 * 2 + NST * 2 * 64 * 64 * 2 + 1024 + 128;
    {  // function attributes are per device: set on every launch (host-side, microseconds)
      cudaError_t e = set_smem(attn_decode_mma_kernel<64, NST>, smem);
      if (e != cudaSuccess) return e;
    }
    return launch_pdl(attn_decode_mma_kernel<64, NST>, g, blk, smem, st, pdl, p);
  }
  return cudaErrorInvalidValue;
}

__global__ void advance_kernel(const int* slots, int* suf_len) {
  pdl_wait();
  suf_len[slots[threadIdx.x]] += 1;
}

// advspec_append_tail: K/V rows [row0, row0 + n) of every (layer, K|V, KV head) of the prefix region
// ([L][2][Hkv][pstride][dh]) -> rows [0, n) of opponent `slot` in the suffix region ([L][2][slot][Hkv][sstride][dh]).
// grid (L * 2, Hkv); 16-byte vectors (dh * 2 bytes is a multiple of 16 for every head_dim served).
__global__ void kv_tail_to_suffix_kernel(const __nv_bfloat16* __restrict__ pkv, __nv_bfloat16* __restrict__ skv,
                                         size_t pkv_layer_elems, size_t skv_layer_elems, int Hkv, int64_t pstride,
                                         int64_t sstride, int dh, int slot, int row0, int n) {
  const int lk = blockIdx.x, hk = blockIdx.y;  // lk = layer * 2 + (0: K, 1: V)
  const uint4* src = reinterpret_cast<const uint4*>(pkv + (size_t)lk * pkv_layer_elems +
                                                    ((size_t)hk * pstride + row0) * dh);
  uint4* dst = reinterpret_cast<uint4*>(skv + (size_t)lk * skv_layer_elems +
                                        ((size_t)slot * Hkv + hk) * sstride * dh);
  const int vecs = n * dh / 8;
  for (int i = threadIdx.x; i < vecs; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

// ============================================================================
struct advspec_engine {
  advspec_model_desc d{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::string err;

  BlobLayout lay;
  uint8_t* w = nullptr;
  bool weights_ready = false;

  // RoPE tables
  float *inv_freq = nullptr, *rope_cos = nullptr, *rope_sin = nullptr;
  int max_pos = 0;

  // KV: prefix [L][2][Hkv][max_prefix][DH]; suffix [L][2][max_seqs][Hkv][max_new][DH]
  __nv_bfloat16 *pkv = nullptr, *skv = nullptr;
  size_t pkv_layer_elems = 0, skv_layer_elems = 0;

  // prefill workspaces (chunk of C tokens)
  int C = 0;
  int* p_tokens = nullptr;
  float* p_x = nullptr;
  __nv_bfloat16 *p_xn = nullptr, *p_qkv = nullptr, *p_attn = nullptr, *p_h = nullptr;
  float* prefill_logits = nullptr;  // [V]

  // decode workspaces (batch of max_seqs)
  float *dx = nullptr, *dx_save = nullptr, *dq = nullptr, *dlogits = nullptr;
  __nv_bfloat16 *dqkv = nullptr, *dattn = nullptr, *dh = nullptr;
  float *part_m = nullptr, *part_l = nullptr, *part_o = nullptr;
  unsigned int* gemm_sem = nullptr;   // ordering words of the prefill GEMM's K-split tail (zero between launches)
  AttnItem* items = nullptr;
  int items_cap = 0, n_items = 0, n_slots = 0;
  // fused tensor-core decode attention: work decomposition of the current batch
  int a2_opg = 1, a2_n_og = 1, a2_n_splits = 1, a2_ctas = 0;
  int h_slots[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int* s_pos = nullptr;     // [max_seqs] absolute position of each batch entry's next token
  std::vector<CUtensorMap> kv_maps;  // [L][2] prefix K / V tensor maps of each layer (passed by value per launch)
  bool attn_fused = false;  // this engine's shape is served by attn_decode_mma_kernel

  // opponent state (device arrays indexed by slot)
  int *s_slots = nullptr, *s_forced = nullptr;  // [max_seqs] batch -> slot / forced tokens
  uint64_t* s_seeds = nullptr;
  int *s_suf_len = nullptr, *s_n_out = nullptr, *s_done = nullptr, *s_cur_tok = nullptr,
      *s_out = nullptr;
  // partial winners of the sampler: [tp_size][2][max_seqs * kSampleChunks] 4-byte words (scores, then
  // token ids); a rank fills its own slice, tensor-parallel ranks all-gather the rest
  float* samp_best = nullptr;
  int* samp_idx = nullptr;  // = samp_best + max_seqs * kSampleChunks (this rank's slice)
  uint32_t* samp_pack = nullptr;

  // tensor parallelism
  int V_full = 0;  // rows of the embedding table / range of token ids (d.vocab_size is this rank's share)
  int tp_rank = 0, tp_size = 1;
  ncclComm_t comm = nullptr;
  // decode-step exchange over peer memory (tp_allreduce.cuh): this rank's region, the peers' mappings
  uint8_t* ar_region = nullptr;
  uint8_t* ar_peer[kArMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned int* ar_gen = nullptr;
  int64_t ar_max_elems = 0;
  bool ar_ready = false;

  // host-side bookkeeping
  int prefix_gen = 0;     // id of the live prefix (0 = none)
  int prefix_len = 0;
  bool slot_used[8] = {false, false, false, false, false, false, false, false};
  std::vector<int> h_suf_len = std::vector<int>(8, 0);
  bool logits_broadcast = true;  // current logits are the prefill's (shared) ones
  bool tails_open = false;       // advspec_append_tail is filling dlogits rows in logits_slots order
  std::vector<int> logits_slots;  // batch order of dlogits when !logits_broadcast

  // decode graph cache
  cudaGraphExec_t graph = nullptr;
  std::vector<int> graph_key;
  int64_t graph_launches = 0;  // kernels (and NCCL calls) one replay of `graph` launches

  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  advspec_timing tm{};
  int64_t launches = 0;
  int debug_flags = 0;  // bit0: check GEMM instead of tcgen05; bit1: scalar attention (tests only)
  bool use_graph = true;

  bool fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return false;
  }
  const uint8_t* wp(size_t off) const { return w + off; }
};

#define E_CUDA(e, call)                                                                 \
  do {                                                                                  \
    cudaError_t _err = (call);                                                          \
    if (_err != cudaSuccess) {                                                          \
      (e)->fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_err), __FILE__, __LINE__); \
      return (_err == cudaErrorMemoryAllocation) ? ADVSPEC_ERR_OOM : ADVSPEC_ERR_CUDA;  \
    }                                                                                   \
  } while (0)

namespace {

advspec_status check_watchdog(advspec_engine* e) {
  unsigned int code = 0;
  E_CUDA(e, cudaMemcpyFromSymbol(&code, g_watchdog_code, sizeof code));
  if (code != 0) {
    e->fail("device watchdog tripped: mbarrier wait timed out at site 0x%x", code & 0x7fffffffu);
    unsigned int zero = 0;
    cudaMemcpyToSymbol(g_watchdog_code, &zero, sizeof zero);
    return ADVSPEC_ERR_KERNEL;
  }
  return ADVSPEC_OK;
}

template <typename T>
cudaError_t dmalloc(T** p, size_t n) {
  return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
}

__nv_bfloat16* prefix_k(advspec_engine* e, int layer) { return e->pkv + (size_t)layer * 2 * e->pkv_layer_elems; }
__nv_bfloat16* prefix_v(advspec_engine* e, int layer) { return prefix_k(e, layer) + e->pkv_layer_elems; }
__nv_bfloat16* suffix_k(advspec_engine* e, int layer) { return e->skv + (size_t)layer * 2 * e->skv_layer_elems; }
__nv_bfloat16* suffix_v(advspec_engine* e, int layer) { return suffix_k(e, layer) + e->skv_layer_elems; }

struct LayerW {
  const float* attn_norm;
  const __nv_bfloat16* wqkv;
  const float* bqkv;
  const __nv_bfloat16* wo;
  const float* mlp_norm;
  const __nv_bfloat16* wgu;
  const __nv_bfloat16* wd;
};
LayerW layer_w(advspec_engine* e, int l) {
  const auto& o = e->lay.layers[l];
  LayerW r;
  r.attn_norm = reinterpret_cast<const float*>(e->wp(o.attn_norm));
  r.wqkv = reinterpret_cast<const __nv_bfloat16*>(e->wp(o.wqkv));
  r.bqkv = e->d.qkv_bias ? reinterpret_cast<const float*>(e->wp(o.bqkv)) : nullptr;
  r.wo = reinterpret_cast<const __nv_bfloat16*>(e->wp(o.wo));
  r.mlp_norm = reinterpret_cast<const float*>(e->wp(o.mlp_norm));
  r.wgu = reinterpret_cast<const __nv_bfloat16*>(e->wp(o.wgu));
  r.wd = reinterpret_cast<const __nv_bfloat16*>(e->wp(o.wd));
  return r;
}
const __nv_bfloat16* embed_w(advspec_engine* e) { return reinterpret_cast<const __nv_bfloat16*>(e->wp(e->lay.embed)); }
const __nv_bfloat16* lm_head_w(advspec_engine* e) {
  return e->d.tied_lm_head ? embed_w(e) : reinterpret_cast<const __nv_bfloat16*>(e->wp(e->lay.lm_head));
}
const float* final_norm_w(advspec_engine* e) { return reinterpret_cast<const float*>(e->wp(e->lay.final_norm)); }

// Tensor-parallel exchange: sum the ranks' partial residual streams in place (rank 0's partial
// already carries the residual, so the sum is the new residual on every rank).
advspec_status tp_allreduce(advspec_engine* e, float* buf, size_t count) {
  if (e->tp_size == 1) return ADVSPEC_OK;
  if (!e->comm) {
    e->fail("tensor-parallel engine used before advspec_tp_init");
    return ADVSPEC_ERR_STATE;
  }
  if (e->ar_ready && (int64_t)count <= e->ar_max_elems && count % 4 == 0) {
    // latency-bound size (the decode step): push/flag/sum over NVLink peer memory, one launch
    ArParams ap{};
    for (int r = 0; r < e->tp_size; ++r) ap.peer[r] = e->ar_peer[r];
    ap.data = buf;
    ap.n = (int)count;
    ap.tp = e->tp_size;
    ap.rank = e->tp_rank;
    ap.max_elems = e->ar_max_elems;
    ap.gen = e->ar_gen;
    if (g_tp_ll)
      E_CUDA(e, launch_pdl(tp_allreduce_ll_kernel, dim3(kArCtas), dim3(kArThreads), 0, e->stream, true, ap));
    else
      E_CUDA(e, launch_pdl(tp_allreduce_kernel, dim3(kArCtas), dim3(kArThreads), 0, e->stream, true, ap));
    e->launches++;
    return ADVSPEC_OK;
  }
  ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat32, ncclSum, e->comm, e->stream);
  if (r != ncclSuccess) {
    e->fail("ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
    return ADVSPEC_ERR_CUDA;
  }
  e->launches++;
  return ADVSPEC_OK;
}
// Epilogue of a row-split product (o-proj, down-proj): rank 0 adds into the residual, the other
// ranks overwrite their copy with the bare partial; tp_allreduce follows.
inline int tp_resadd_epi(const advspec_engine* e) {
  return (e->tp_size > 1 && e->tp_rank != 0) ? EPI_F32 : EPI_RESADD_F32;
}

// One GEMM of the prefill path (tcgen05 unless the test-only debug flag asks for the check kernel).
advspec_status prefill_gemm(advspec_engine* e, const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* B,
                            int64_t ldb, void* C, int64_t ldc, const float* bias, int M, int N, int K,
                            int epi) {
  GemmParams p{C, ldc, bias, M, N, K, e->d.act};
  if (e->debug_flags & 1) {
    E_CUDA(e, launch_gemm_check(A, lda, B, ldb, p, epi, e->stream));
  } else {
    std::string why;
    // the activation buffers hold C rows, so the TMA extent may cover whole 128-row tiles
    const int64_t a_rows = std::min<int64_t>((M + kGemmBM - 1) / kGemmBM * kGemmBM, e->C);
    cudaError_t r = launch_gemm(A, lda, a_rows, B, ldb, p, epi, e->device, e->stream, &why, e->gemm_sem);
    if (r != cudaSuccess) {
      e->fail("prefill gemm M=%d N=%d K=%d failed: %s %s", M, N, K, cudaGetErrorString(r), why.c_str());
      return ADVSPEC_ERR_CUDA;
    }
  }
  e->launches++;
  return ADVSPEC_OK;
}

// Runs the transformer layers over `m` prompt tokens at positions pos0.. (tokens
// already in p_tokens), leaving the fp32 residual stream in p_x.
advspec_status prefill_chunk(advspec_engine* e, int m, int pos0) {
  const auto& d = e->d;
  const int dm = d.d_model, QKV = qkv_dim(d), HD = d.n_heads * d.head_dim;
  embed_kernel<<<m, 256, 0, e->stream>>>(e->p_tokens, embed_w(e), e->p_x, dm, d.embed_scale);
  E_CUDA(e, cudaGetLastError());
  e->launches++;
  for (int l = 0; l < d.n_layers; ++l) {
    const LayerW w = layer_w(e, l);
    E_CUDA(e, launch_rmsnorm(e->p_x, w.attn_norm, e->p_xn, m, dm, d.norm_eps, e->stream));
    advspec_status s = prefill_gemm(e, e->p_xn, dm, w.wqkv, dm, e->p_qkv, QKV, w.bqkv, m, QKV, dm, EPI_BF16);
    if (s) return s;
    rope_prefill_kernel<<<m, 256, 0, e->stream>>>(e->p_qkv, QKV, prefix_k(e, l), prefix_v(e, l),
                                                   d.max_prefix_tokens, e->rope_cos, e->rope_sin, pos0,
                                                   d.n_heads, d.n_kv_heads, d.head_dim);
    E_CUDA(e, cudaGetLastError());
    AttnPrefillParams ap{e->p_qkv, QKV, prefix_k(e, l), prefix_v(e, l), d.max_prefix_tokens, e->p_attn,
                         m, pos0, d.n_heads, d.n_kv_heads, 1.0f / sqrtf((float)d.head_dim), d.head_dim};
    if (d.head_dim <= 128 && g_attn_prefill_tc && !(e->debug_flags & 2)) {
      std::string why;
      cudaError_t r = launch_attn_prefill_tc(e->p_qkv, QKV, prefix_k(e, l), prefix_v(e, l), d.max_prefix_tokens,
                                             e->p_attn, m, pos0, d.n_heads, d.n_kv_heads, d.head_dim, e->stream, &why);
      if (r != cudaSuccess) {
        e->fail("tcgen05 attention launch failed: %s %s", cudaGetErrorString(r), why.c_str());
        return ADVSPEC_ERR_CUDA;
      }
    } else {
      E_CUDA(e, launch_attn_prefill(ap, d.head_dim, (e->debug_flags & 2) ? 1 : 0, e->stream, nullptr));
    }
    s = prefill_gemm(e, e->p_attn, HD, w.wo, HD, e->p_x, dm, nullptr, m, dm, HD, tp_resadd_epi(e));
    if (s) return s;
    s = tp_allreduce(e, e->p_x, (size_t)m * dm);
    if (s) return s;
    E_CUDA(e, launch_rmsnorm(e->p_x, w.mlp_norm, e->p_xn, m, dm, d.norm_eps, e->stream));
    s = prefill_gemm(e, e->p_xn, dm, w.wgu, dm, e->p_h, d.d_ff, nullptr, m, 2 * d.d_ff, dm, EPI_GATED_BF16);
    if (s) return s;
    s = prefill_gemm(e, e->p_h, d.d_ff, w.wd, d.d_ff, e->p_x, dm, nullptr, m, dm, d.d_ff, tp_resadd_epi(e));
    if (s) return s;
    s = tp_allreduce(e, e->p_x, (size_t)m * dm);
    if (s) return s;
    e->launches += 4;
  }
  return ADVSPEC_OK;
}

void free_all(advspec_engine* e) {
  cudaSetDevice(e->device);
  if (e->graph) cudaGraphExecDestroy(e->graph);
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  e->comm = nullptr;
  for (int r = 0; r < kArMaxRanks; ++r)
    if (e->ar_peer[r] && r != e->tp_rank) cudaIpcCloseMemHandle(e->ar_peer[r]);
  if (e->ar_region) cudaFree(e->ar_region);
  if (e->ar_gen) cudaFree(e->ar_gen);
  void* ptrs[] = {e->w, e->inv_freq, e->rope_cos, e->rope_sin, e->pkv, e->skv, e->p_tokens, e->p_x,
                  e->p_xn, e->p_qkv, e->p_attn, e->p_h, e->prefill_logits, e->dx, e->dx_save, e->dq,
                  e->dlogits, e->dqkv, e->dattn, e->dh, e->part_m, e->part_l, e->part_o, e->gemm_sem, e->items, e->s_pos,
                  e->s_slots, e->s_forced, e->s_seeds, e->s_suf_len, e->s_n_out, e->s_done,
                  e->s_cur_tok, e->s_out, e->samp_pack};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
}

// Attention work items for a decode call of batch `slots` over the live prefix.
void build_items(advspec_engine* e, const std::vector<int>& slots, std::vector<AttnItem>* out,
                 int* n_slots) {
  const auto& d = e->d;
  const int b = (int)slots.size();
  const int G = d.n_heads / d.n_kv_heads;
  // rows of one KV head over the shared prefix: (batch, head) for every opponent
  const int rows = b * G;
  const int nrg = (rows + 3) / 4;
  int n_splits = std::max(1, (2 * num_sms(e->device)) / std::max(1, d.n_kv_heads * nrg));
  n_splits = std::min(n_splits, std::max(1, e->prefix_len / 256));
  *n_slots = n_splits + 1;
  out->clear();
  for (int hk = 0; hk < d.n_kv_heads; ++hk) {
    for (int rg = 0; rg < nrg; ++rg) {
      for (int s = 0; s < n_splits; ++s) {
        AttnItem it{};
        it.kv_head = hk;
        it.seq = -1;
        it.tok_begin = (int)((int64_t)e->prefix_len * s / n_splits);
        it.tok_end = (int)((int64_t)e->prefix_len * (s + 1) / n_splits);
        it.slot = s;
        it.n_rows = 0;
        for (int r = rg * 4; r < std::min(rows, rg * 4 + 4); ++r) {
          it.row_b[it.n_rows] = r / G;
          it.row_head[it.n_rows] = hk * G + r % G;
          it.n_rows++;
        }
        out->push_back(it);
      }
    }
  }
  for (int bi = 0; bi < b; ++bi) {
    for (int hk = 0; hk < d.n_kv_heads; ++hk) {
      for (int g0 = 0; g0 < G; g0 += 4) {
        AttnItem it{};
        it.kv_head = hk;
        it.seq = slots[bi];
        it.slot = n_splits;
        it.n_rows = 0;
        for (int g = g0; g < std::min(G, g0 + 4); ++g) {
          it.row_b[it.n_rows] = bi;
          it.row_head[it.n_rows] = hk * G + g;
          it.n_rows++;
        }
        out->push_back(it);
      }
    }
  }
}

// Decomposition of the fused decode attention for a batch: per KV head, opponents are grouped so that
// a group's query rows (opponents x G heads) fill one 16-row MMA tile; the prefix is cut into splits
// shared by the whole group (one CTA per SM in total), each opponent's suffix is its own CTA.
struct Attn2Plan {
  int opg, n_og, n_splits, ctas, n_slots;
};
Attn2Plan plan_attn2_shape(int b, int n_heads, int n_kv_heads, int prefix_len, int device) {
  const int G = n_heads / n_kv_heads;
  const int opg = std::max(1, 16 / G);
  const int n_og = (b + opg - 1) / opg;
  const int groups = n_kv_heads * n_og;
  // one wave of CTAs when that still cuts the prefix at least in two (GQA models); with many KV heads
  // (MHA: Phi-3 has 32) the prefix CTAs alone fill the wave and the short per-opponent suffix CTAs trail
  const int slots_left = std::max(groups, num_sms(device) - b * n_kv_heads);
  int n_splits = std::max(1, slots_left / std::max(1, groups));
  // (measured: applying this whenever it cuts the prefix finer — Gemma 9 splits instead of 6 — is slower,
  // 4.18 vs 4.08 ms/step: the trailing suffix CTAs cost more than the finer slices save)
  if (n_splits < 2) n_splits = std::max(1, num_sms(device) / std::max(1, groups));
  n_splits = std::min(n_splits, std::max(1, prefix_len / g_attn_min_split));
  n_splits = std::min(n_splits, 300);
  return Attn2Plan{opg, n_og, n_splits, groups * (n_splits + opg), n_splits + 1};
}
void plan_attn2(advspec_engine* e, const std::vector<int>& slots) {
  const auto& d = e->d;
  const int b = (int)slots.size();
  const Attn2Plan pl = plan_attn2_shape(b, d.n_heads, d.n_kv_heads, e->prefix_len, e->device);
  e->a2_opg = pl.opg;
  e->a2_n_og = pl.n_og;
  e->a2_n_splits = pl.n_splits;
  e->a2_ctas = pl.ctas;
  e->n_slots = pl.n_slots;
  for (int i = 0; i < 8; ++i) e->h_slots[i] = i < b ? slots[i] : 0;
}

// Enqueue one forward step (all layers + lm_head) for the batch in s_slots.
advspec_status enqueue_forward(advspec_engine* e, int b, float* gemv_ms_out) {
  const auto& d = e->d;
  const int dm = d.d_model, QKV = qkv_dim(d), HD = d.n_heads * d.head_dim;
  const bool prof = gemv_ms_out != nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
  auto gemv = [&](const GemvParams& gp) -> cudaError_t {
    cudaEvent_t a = nullptr, z = nullptr;
    if (prof) {
      cudaEventCreate(&a);
      cudaEventCreate(&z);
      cudaEventRecord(a, e->stream);
    }
    cudaError_t r = launch_gemv(gp, b, e->device, e->stream, !prof);
    if (prof) {
      cudaEventRecord(z, e->stream);
      evs.emplace_back(a, z);
    }
    e->launches++;
    return r;
  };
  for (int l = 0; l < d.n_layers; ++l) {
    const LayerW w = layer_w(e, l);
    GemvParams g1{w.wqkv, e->dx, w.attn_norm, w.bqkv, e->dqkv, QKV, dm, 1, EPI_BF16, d.act, d.norm_eps};
    E_CUDA(e, gemv(g1));
    ADV_TRACE(e->stream, "gemv qkv");
    if (e->attn_fused) {
      AttnDecode2Params a2{};
      a2.qkv = e->dqkv;
      a2.rope_cos = e->rope_cos;
      a2.rope_sin = e->rope_sin;
      a2.pk = prefix_k(e, l);
      a2.pv = prefix_v(e, l);
      a2.pstride = d.max_prefix_tokens;
      a2.sk = suffix_k(e, l);
      a2.sv = suffix_v(e, l);
      a2.sstride = d.max_new_tokens;
      a2.map_k = e->kv_maps[2 * l];
      a2.map_v = e->kv_maps[2 * l + 1];
      a2.pos_b = e->s_pos;
      for (int i = 0; i < 8; ++i) a2.slots[i] = e->h_slots[i];
      a2.prefix_len = e->prefix_len;
      a2.part_m = e->part_m;
      a2.part_l = e->part_l;
      a2.part_o = e->part_o;
      a2.b = b;
      a2.H = d.n_heads;
      a2.Hkv = d.n_kv_heads;
      a2.G = d.n_heads / d.n_kv_heads;
      a2.opg = e->a2_opg;
      a2.n_og = e->a2_n_og;
      a2.n_splits = e->a2_n_splits;
      a2.n_slots = e->n_slots;
      a2.scale = 1.0f / sqrtf((float)d.head_dim);
      a2.dh = d.head_dim;
      E_CUDA(e, launch_attn_decode2(a2, e->a2_ctas, d.head_dim, e->stream, true));
      ADV_TRACE(e->stream, "attn_decode_mma");
      E_CUDA(e, launch_pdl(attn_decode_combine2_kernel, dim3(b * d.n_heads), dim3(128), 0, e->stream, true,
                           (const float*)e->part_m, (const float*)e->part_l, (const float*)e->part_o, e->dattn,
                           e->n_slots, d.head_dim));
      ADV_TRACE(e->stream, "attn_combine2");
      e->launches += 2;  // attention (RoPE + append + split-KV) and the combine
    } else {
    E_CUDA(e, launch_pdl(rope_decode_kernel, dim3(b), dim3(256), 0, e->stream, true,
                           (const __nv_bfloat16*)e->dqkv, e->dq, suffix_k(e, l), suffix_v(e, l),
                           (int64_t)d.max_new_tokens, (const int*)e->s_slots, (const int*)e->s_suf_len,
                           e->prefix_len, (const float*)e->rope_cos, (const float*)e->rope_sin, d.n_heads,
                           d.n_kv_heads, d.head_dim));
      ADV_TRACE(e->stream, "rope_decode");
      AttnDecodeParams ap{};
      ap.items = e->items;
      ap.q = e->dq;
      ap.pk = prefix_k(e, l);
      ap.pv = prefix_v(e, l);
      ap.pstride = d.max_prefix_tokens;
      ap.sk = suffix_k(e, l);
      ap.sv = suffix_v(e, l);
      ap.sstride = d.max_new_tokens;
      ap.suf_len = e->s_suf_len;
      ap.part_m = e->part_m;
      ap.part_l = e->part_l;
      ap.part_o = e->part_o;
      ap.H = d.n_heads;
      ap.Hkv = d.n_kv_heads;
      ap.n_slots = e->n_slots;
      ap.scale = 1.0f / sqrtf((float)d.head_dim);
      E_CUDA(e, launch_attn_decode(ap, e->n_items, d.head_dim, e->stream, true));
      ADV_TRACE(e->stream, "attn_decode");
      E_CUDA(e, launch_pdl(attn_decode_combine_kernel, dim3(b * d.n_heads), dim3(128), 0, e->stream, true,
                           (const float*)e->part_m, (const float*)e->part_l, (const float*)e->part_o,
                           e->dattn, e->n_slots, d.head_dim));
      ADV_TRACE(e->stream, "attn_combine");
      e->launches += 3;  // rope, attention, combine
    }
    GemvParams g2{w.wo, e->dattn, nullptr, nullptr, e->dx, dm, HD, 0, tp_resadd_epi(e), d.act, d.norm_eps};
    E_CUDA(e, gemv(g2));
    ADV_TRACE(e->stream, "gemv o");
    if (advspec_status ts = tp_allreduce(e, e->dx, (size_t)b * dm)) return ts;
    GemvParams g3{w.wgu, e->dx, w.mlp_norm, nullptr, e->dh, 2 * d.d_ff, dm, 1, EPI_GATED_BF16, d.act, d.norm_eps};
    E_CUDA(e, gemv(g3));
    ADV_TRACE(e->stream, "gemv gate_up");
    GemvParams g4{w.wd, e->dh, nullptr, nullptr, e->dx, dm, d.d_ff, 0, tp_resadd_epi(e), d.act, d.norm_eps};
    E_CUDA(e, gemv(g4));
    if (advspec_status ts = tp_allreduce(e, e->dx, (size_t)b * dm)) return ts;
    ADV_TRACE(e->stream, "gemv down");
  }
  GemvParams gl{lm_head_w(e), e->dx, final_norm_w(e), nullptr, e->dlogits, d.vocab_size, dm, 1, EPI_F32,
                d.act, d.norm_eps};
  E_CUDA(e, gemv(gl));
  ADV_TRACE(e->stream, "gemv lm_head");
  if (prof) {
    E_CUDA(e, cudaStreamSynchronize(e->stream));
    float total = 0.f;
    for (auto& pr : evs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pr.first, pr.second);
      total += ms;
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    *gemv_ms_out = total;
    e->tm.gemv_launches = (int)evs.size();
  }
  return ADVSPEC_OK;
}

SampleParams make_sample_params(advspec_engine* e, float temperature, int eos_id, bool broadcast,
                                bool advance, const int* forced, bool record) {
  SampleParams sp{};
  sp.logits = broadcast ? e->prefill_logits : e->dlogits;
  sp.broadcast_logits = broadcast ? 1 : 0;
  sp.V = e->d.vocab_size;
  sp.v_off = e->tp_rank * e->d.vocab_size;
  sp.n_ranks = e->tp_size;
  sp.rank_stride = 2 * e->d.max_seqs * kSampleChunks;
  sp.temperature = temperature;
  sp.slots = e->s_slots;
  sp.seeds = e->s_seeds;
  sp.suf_len = e->s_suf_len;
  sp.n_out = e->s_n_out;
  sp.done = e->s_done;
  sp.out_tokens = record ? e->s_out : nullptr;
  sp.out_stride = e->d.max_new_tokens;
  sp.eos_id = eos_id;
  sp.advance = advance ? 1 : 0;
  sp.forced = forced;
  sp.pos_b = e->s_pos;
  sp.prefix_len = e->prefix_len;
  sp.part_best = forced ? nullptr : reinterpret_cast<const float*>(e->samp_pack);
  sp.part_idx = forced ? nullptr : reinterpret_cast<const int*>(e->samp_pack) + e->d.max_seqs * kSampleChunks;
  sp.cur_tok = e->s_cur_tok;
  sp.embed = embed_w(e);
  sp.x = e->dx;
  sp.d = e->d.d_model;
  sp.embed_scale = e->d.embed_scale;
  return sp;
}

// Sampler = vocabulary scan on 64 CTAs per opponent + a one-CTA merge that also advances the
// opponent and writes the next embedding.
cudaError_t launch_sampler(advspec_engine* e, const SampleParams& sp, int n, bool pdl) {
  if (sp.forced == nullptr) {
    cudaError_t r = launch_pdl(sample_partial_kernel, dim3(kSampleChunks, n), dim3(256), 0, e->stream, pdl, sp,
                               e->samp_best, e->samp_idx);
    if (r != cudaSuccess) return r;
    e->launches++;
    pdl = true;
    if (e->tp_size > 1) {
      // every rank scanned its share of the vocabulary: gather the winners, then all merge identically
      if (!e->comm) return cudaErrorNotReady;
      const size_t words = (size_t)sp.rank_stride;
      ncclResult_t nr = g_nccl.AllGather(e->samp_pack + (size_t)e->tp_rank * words, e->samp_pack, words, ncclUint32,
                                         e->comm, e->stream);
      if (nr != ncclSuccess) return cudaErrorUnknown;
      e->launches++;
    }
  }
  e->launches++;
  return launch_pdl(sample_kernel, dim3(n), dim3(1024), 0, e->stream, pdl, sp);
}

advspec_status setup_batch(advspec_engine* e, const int32_t* seq_ids, int n, std::vector<int>* slots) {
  if (!seq_ids || n < 1 || n > e->d.max_seqs) {
    e->fail("batch size %d outside 1..%d", n, e->d.max_seqs);
    return ADVSPEC_ERR_INVALID;
  }
  if (e->prefix_gen == 0) {
    e->fail("no live prefix: call advspec_prefill first");
    return ADVSPEC_ERR_STATE;
  }
  slots->clear();
  for (int i = 0; i < n; ++i) {
    const int s = seq_ids[i];
    if (s < 0 || s >= e->d.max_seqs || !e->slot_used[s]) {
      e->fail("seq id %d is not a live opponent", s);
      return ADVSPEC_ERR_STATE;
    }
    for (int j = 0; j < i; ++j)
      if (seq_ids[j] == s) {
        e->fail("seq id %d listed twice", s);
        return ADVSPEC_ERR_INVALID;
      }
    slots->push_back(s);
  }
  E_CUDA(e, cudaMemcpyAsync(e->s_slots, slots->data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  std::vector<AttnItem> items;
  build_items(e, *slots, &items, &e->n_slots);
  if ((int)items.size() > e->items_cap) {
    e->fail("internal: %zu attention items exceed capacity %d", items.size(), e->items_cap);
    return ADVSPEC_ERR_INVALID;
  }
  e->n_items = (int)items.size();
  E_CUDA(e, cudaMemcpyAsync(e->items, items.data(), items.size() * sizeof(AttnItem), cudaMemcpyHostToDevice,
                            e->stream));
  if (e->attn_fused) plan_attn2(e, *slots);
  E_CUDA(e, cudaStreamSynchronize(e->stream));  // items/slots vectors die with this scope
  return ADVSPEC_OK;
}

}  // namespace

// ============================================================================
extern "C" {

size_t advspec_weight_blob_bytes(const advspec_model_desc* desc) {
  if (!desc_ok(desc, nullptr)) return 0;
  const LocalDesc ld = localize(*desc);
  return make_layout(ld.d, ld.v_full).total;
}

size_t advspec_weight_offset(const advspec_model_desc* desc, int32_t layer, const char* name) {
  if (!desc_ok(desc, nullptr) || !name) return (size_t)-1;
  const LocalDesc ld = localize(*desc);
  const BlobLayout L = make_layout(ld.d, ld.v_full);
  const std::string n(name);
  if (layer < 0) {
    if (n == "embed") return L.embed;
    if (n == "final_norm") return L.final_norm;
    if (n == "lm_head") return L.lm_head;
    return (size_t)-1;
  }
  if (layer >= desc->n_layers) return (size_t)-1;
  const auto& o = L.layers[layer];
  if (n == "attn_norm") return o.attn_norm;
  if (n == "wqkv") return o.wqkv;
  if (n == "bqkv") return o.bqkv;
  if (n == "wo") return o.wo;
  if (n == "mlp_norm") return o.mlp_norm;
  if (n == "wgu") return o.wgu;
  if (n == "wd") return o.wd;
  return (size_t)-1;
}

const char* advspec_last_error(const advspec_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

advspec_status advspec_engine_create(const advspec_model_desc* desc, int32_t device, advspec_engine** out) {
  if (!out) return ADVSPEC_ERR_INVALID;
  *out = nullptr;
  std::string why;
  if (!desc_ok(desc, &why)) {
    g_create_error = why;
    return ADVSPEC_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0) {
    g_create_error = "no CUDA device " + std::to_string(device) + " (this engine has no CPU fallback)";
    return ADVSPEC_ERR_CUDA;
  }
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                     "; this library carries sm_100a code only";
    return ADVSPEC_ERR_CUDA;
  }
  advspec_engine* e = new advspec_engine();
  const LocalDesc ld = localize(*desc);
  e->d = ld.d;  // from here on every shape is this rank's share
  e->V_full = ld.v_full;
  e->tp_rank = desc->tp_size > 1 ? desc->tp_rank : 0;
  e->tp_size = desc->tp_size > 1 ? desc->tp_size : 1;
  e->device = device;
  e->lay = make_layout(ld.d, ld.v_full);
  const char* dbg = getenv("ADVSPEC_DEBUG_FLAGS");
  e->debug_flags = dbg ? atoi(dbg) : 0;
  e->use_graph = getenv("ADVSPEC_NO_GRAPH") == nullptr;
  g_use_pdl = getenv("ADVSPEC_NO_PDL") == nullptr;
  g_trace = getenv("ADVSPEC_TRACE") != nullptr;
  // process-wide A/B knobs are re-read at every create (unset = default), so one process can compare them
  const char* gi = getenv("ADVSPEC_GEMV_IMPL");
  g_gemv_impl = (gi && atoi(gi) == 1) ? 1 : 3;
  const char* ai = getenv("ADVSPEC_ATTN_IMPL");
  g_attn_impl = (ai && atoi(ai) == 1) ? 1 : 2;
  const char* xm = getenv("ADVSPEC_X_SMEM_MAX");
  g_x_smem_max = xm ? (size_t)atoll(xm) : 40000;
  g_gemm_narrow = getenv("ADVSPEC_GEMM_NARROW") != nullptr;
  {
    const char* ar = getenv("ADVSPEC_TP_AR");
    g_tp_ll = !(ar && std::string(ar) == "flag");
  }
  const char* gb = getenv("ADVSPEC_GEMM_BAND_MB");
  g_gemm_band_mb = gb ? std::max(0, atoi(gb)) : 0;
  const char* gs = getenv("ADVSPEC_GEMM_SPLITK");
  g_gemm_splitk = gs ? atoi(gs) != 0 : 1;
  const char* tc = getenv("ADVSPEC_ATTN_PREFILL_TC");
  g_attn_prefill_tc = tc ? atoi(tc) != 0 : true;
  const char* ms = getenv("ADVSPEC_ATTN_MIN_SPLIT");
  g_attn_min_split = ms ? std::max(64, atoi(ms)) : 256;

  auto boot = [&]() -> advspec_status {
    const auto& d = e->d;
    E_CUDA(e, cudaSetDevice(device));
    {
      const char* ef = getenv("ADVSPEC_L2_EVICT_FIRST");
      const int v = ef ? atoi(ef) : 1;
      E_CUDA(e, cudaMemcpyToSymbol(g_l2_evict_first, &v, sizeof v));
    }
    E_CUDA(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    E_CUDA(e, cudaEventCreate(&e->ev0));
    E_CUDA(e, cudaEventCreate(&e->ev1));
    E_CUDA(e, dmalloc(&e->w, e->lay.total));
    const int half = d.head_dim / 2;
    e->max_pos = d.max_prefix_tokens + d.max_new_tokens;
    E_CUDA(e, dmalloc(&e->inv_freq, half));
    E_CUDA(e, dmalloc(&e->rope_cos, (size_t)e->max_pos * half));
    E_CUDA(e, dmalloc(&e->rope_sin, (size_t)e->max_pos * half));
    e->pkv_layer_elems = (size_t)d.n_kv_heads * d.max_prefix_tokens * d.head_dim;
    e->skv_layer_elems = (size_t)d.max_seqs * d.n_kv_heads * d.max_new_tokens * d.head_dim;
    E_CUDA(e, dmalloc(&e->pkv, (size_t)d.n_layers * 2 * e->pkv_layer_elems));
    E_CUDA(e, dmalloc(&e->skv, (size_t)d.n_layers * 2 * e->skv_layer_elems));
    // whole 64/128-key boxes are loaded by TMA: rows never written must hold finite values (0 * NaN = NaN)
    E_CUDA(e, cudaMemsetAsync(e->pkv, 0, (size_t)d.n_layers * 2 * e->pkv_layer_elems * sizeof(__nv_bfloat16), e->stream));
    E_CUDA(e, cudaMemsetAsync(e->skv, 0, (size_t)d.n_layers * 2 * e->skv_layer_elems * sizeof(__nv_bfloat16), e->stream));
    // prefill chunk: up to 8192 tokens per pass (fewer, fuller waves of GEMM tiles than 4096 + remainder)
    e->C = std::min(8192, (d.max_prefix_tokens + 127) / 128 * 128);
    if (const char* pc = getenv("ADVSPEC_PREFILL_CHUNK"))  // tests: force multi-chunk prefill on small prompts
      e->C = std::max(128, std::min(e->C, atoi(pc) / 128 * 128));
    const size_t C = e->C, dm = d.d_model, QKV = qkv_dim(d), HD = (size_t)d.n_heads * d.head_dim;
    E_CUDA(e, dmalloc(&e->p_tokens, C));
    E_CUDA(e, dmalloc(&e->p_x, C * dm));
    E_CUDA(e, dmalloc(&e->p_xn, C * dm));
    E_CUDA(e, dmalloc(&e->p_qkv, C * QKV));
    E_CUDA(e, dmalloc(&e->p_attn, C * HD));
    E_CUDA(e, dmalloc(&e->p_h, C * (size_t)d.d_ff));
    E_CUDA(e, cudaMemsetAsync(e->p_xn, 0, C * dm * 2, e->stream));
    E_CUDA(e, cudaMemsetAsync(e->p_attn, 0, C * HD * 2, e->stream));
    E_CUDA(e, cudaMemsetAsync(e->p_h, 0, C * (size_t)d.d_ff * 2, e->stream));
    E_CUDA(e, dmalloc(&e->prefill_logits, (size_t)d.vocab_size));
    const size_t B = d.max_seqs;
    E_CUDA(e, dmalloc(&e->dx, B * dm));
    E_CUDA(e, dmalloc(&e->dx_save, B * dm));
    E_CUDA(e, dmalloc(&e->dq, B * HD));
    E_CUDA(e, dmalloc(&e->dlogits, B * (size_t)d.vocab_size));
    E_CUDA(e, dmalloc(&e->dqkv, B * QKV));
    E_CUDA(e, dmalloc(&e->dattn, B * HD));
    E_CUDA(e, dmalloc(&e->dh, B * (size_t)d.d_ff));
    const int G = d.n_heads / d.n_kv_heads;
    const int max_splits = 2 * 148 + 1;
    E_CUDA(e, dmalloc(&e->part_m, B * d.n_heads * (size_t)(max_splits + 1)));
    E_CUDA(e, dmalloc(&e->part_l, B * d.n_heads * (size_t)(max_splits + 1)));
    E_CUDA(e, dmalloc(&e->part_o, B * d.n_heads * (size_t)(max_splits + 1) * d.head_dim));
    e->items_cap = d.n_kv_heads * (((int)B * G + 3) / 4) * max_splits + (int)B * d.n_kv_heads * ((G + 3) / 4);
    E_CUDA(e, dmalloc(&e->items, (size_t)e->items_cap));
    E_CUDA(e, dmalloc(&e->gemm_sem, kGemmSemWords));
    E_CUDA(e, cudaMemsetAsync(e->gemm_sem, 0, kGemmSemWords * sizeof(unsigned int), e->stream));
    E_CUDA(e, dmalloc(&e->s_pos, B));
    E_CUDA(e, cudaMemsetAsync(e->s_pos, 0, B * sizeof(int), e->stream));
    e->attn_fused = g_attn_impl == 2 && (d.head_dim == 64 || d.head_dim == 96 || d.head_dim == 128 || d.head_dim == 256) && G <= 16 &&
                    d.n_heads <= 255;
    if (e->attn_fused) {
      e->kv_maps.resize((size_t)d.n_layers * 2);
      for (int l = 0; l < d.n_layers; ++l) {
        const int64_t rows = (int64_t)d.n_kv_heads * d.max_prefix_tokens;
        if (!make_tmap(&e->kv_maps[2 * l], prefix_k(e, l), rows, d.head_dim, d.head_dim, 64) ||
            !make_tmap(&e->kv_maps[2 * l + 1], prefix_v(e, l), rows, d.head_dim, d.head_dim, 64)) {
          e->fail("cuTensorMapEncodeTiled failed for the prefix KV of layer %d", l);
          return ADVSPEC_ERR_CUDA;
        }
      }
    }
    E_CUDA(e, dmalloc(&e->s_slots, B));
    E_CUDA(e, dmalloc(&e->s_forced, B));
    E_CUDA(e, dmalloc(&e->s_seeds, B));
    E_CUDA(e, dmalloc(&e->s_suf_len, B));
    E_CUDA(e, dmalloc(&e->s_n_out, B));
    E_CUDA(e, dmalloc(&e->s_done, B));
    E_CUDA(e, dmalloc(&e->s_cur_tok, B));
    E_CUDA(e, dmalloc(&e->s_out, B * (size_t)d.max_new_tokens));
    {
      const size_t b64 = B * (size_t)kSampleChunks;
      E_CUDA(e, dmalloc(&e->samp_pack, (size_t)e->tp_size * 2 * b64));
      e->samp_best = reinterpret_cast<float*>(e->samp_pack + (size_t)e->tp_rank * 2 * b64);
      e->samp_idx = reinterpret_cast<int*>(e->samp_best + b64);
    }
    E_CUDA(e, cudaMemsetAsync(e->s_suf_len, 0, B * sizeof(int), e->stream));
    E_CUDA(e, cudaMemsetAsync(e->s_n_out, 0, B * sizeof(int), e->stream));
    E_CUDA(e, cudaMemsetAsync(e->s_done, 0, B * sizeof(int), e->stream));
    // default RoPE frequencies: theta^(-2i/head_dim), fp32 like the oracle
    std::vector<float> inv(half);
    for (int i = 0; i < half; ++i)
      inv[i] = 1.0f / powf(d.rope_theta, (float)(2 * i) / (float)d.head_dim);
    E_CUDA(e, cudaMemcpyAsync(e->inv_freq, inv.data(), half * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    const int64_t tot = (int64_t)e->max_pos * half;
    rope_table_kernel<<<(int)((tot + 255) / 256), 256, 0, e->stream>>>(e->inv_freq, e->rope_cos, e->rope_sin,
                                                                      e->max_pos, half);
    E_CUDA(e, cudaGetLastError());
    E_CUDA(e, cudaStreamSynchronize(e->stream));
    return ADVSPEC_OK;
  };
  advspec_status s = boot();
  if (s != ADVSPEC_OK) {
    g_create_error = e->err;
    free_all(e);
    delete e;
    return s;
  }
  *out = e;
  return ADVSPEC_OK;
}

void advspec_engine_destroy(advspec_engine* e) {
  if (!e) return;
  {
    std::lock_guard<std::mutex> lk(e->mu);
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    free_all(e);
  }
  delete e;
}

advspec_status advspec_tp_unique_id(uint8_t* out128) {
  if (!out128) return ADVSPEC_ERR_INVALID;
  std::string why;
  if (!nccl_load(&why)) {
    g_create_error = why;
    return ADVSPEC_ERR_STATE;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "the ABI carries the NCCL unique id as 128 opaque bytes");
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) {
    g_create_error = std::string("ncclGetUniqueId failed: ") + g_nccl.GetErrorString(r);
    return ADVSPEC_ERR_CUDA;
  }
  memcpy(out128, &id, 128);
  return ADVSPEC_OK;
}

advspec_status advspec_tp_init(advspec_engine* e, const uint8_t* id128) {
  if (!e || !id128) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->tp_size == 1) return ADVSPEC_OK;  // nothing to join
  if (e->comm) {
    e->fail("advspec_tp_init called twice on one handle");
    return ADVSPEC_ERR_STATE;
  }
  std::string why;
  if (!nccl_load(&why)) {
    e->fail("%s", why.c_str());
    return ADVSPEC_ERR_STATE;
  }
  E_CUDA(e, cudaSetDevice(e->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&e->comm, e->tp_size, id, e->tp_rank);  // blocks until all ranks join
  if (r != ncclSuccess) {
    e->comm = nullptr;
    e->fail("ncclCommInitRank(rank %d of %d) failed: %s", e->tp_rank, e->tp_size, g_nccl.GetErrorString(r));
    return ADVSPEC_ERR_CUDA;
  }
  // one eager round of both collectives: NCCL's lazy setup (buffers, channels) must not happen for the
  // first time inside the decode step's stream capture
  const size_t words = (size_t)2 * e->d.max_seqs * kSampleChunks;
  E_CUDA(e, cudaMemsetAsync(e->dx, 0, (size_t)e->d.max_seqs * e->d.d_model * sizeof(float), e->stream));
  ncclResult_t r1 = g_nccl.AllReduce(e->dx, e->dx, (size_t)e->d.max_seqs * e->d.d_model, ncclFloat32, ncclSum,
                                     e->comm, e->stream);
  ncclResult_t r2 = g_nccl.AllGather(e->samp_pack + (size_t)e->tp_rank * words, e->samp_pack, words, ncclUint32,
                                     e->comm, e->stream);
  if (r1 != ncclSuccess || r2 != ncclSuccess) {
    e->fail("NCCL warm-up collectives failed: %s", g_nccl.GetErrorString(r1 != ncclSuccess ? r1 : r2));
    return ADVSPEC_ERR_CUDA;
  }
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  return ADVSPEC_OK;
}

advspec_status advspec_tp_ipc_export(advspec_engine* e, uint8_t* out64) {
  if (!e || !out64) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries the CUDA IPC handle as 64 opaque bytes");
  if (e->tp_size == 1) {
    e->fail("advspec_tp_ipc_export on a handle without tensor parallelism");
    return ADVSPEC_ERR_STATE;
  }
  E_CUDA(e, cudaSetDevice(e->device));
  if (!e->ar_region) {
    e->ar_max_elems = (int64_t)e->d.max_seqs * e->d.d_model;
    // sized for the low-latency packets (8 bytes per float); the flag variant uses the first half
    const size_t bytes = kArFlagBytes + (size_t)2 * e->tp_size * e->ar_max_elems * 2 * sizeof(float);
    E_CUDA(e, cudaMalloc(reinterpret_cast<void**>(&e->ar_region), bytes));
    E_CUDA(e, cudaMemset(e->ar_region, 0, bytes));
    E_CUDA(e, dmalloc(&e->ar_gen, kArCtas));
    E_CUDA(e, cudaMemset(e->ar_gen, 0, kArCtas * sizeof(unsigned int)));
    E_CUDA(e, cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  E_CUDA(e, cudaIpcGetMemHandle(&h, e->ar_region));
  memcpy(out64, &h, 64);
  return ADVSPEC_OK;
}

advspec_status advspec_tp_ipc_import(advspec_engine* e, const uint8_t* handles) {
  if (!e || !handles) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->ar_region || e->ar_ready) {
    e->fail("advspec_tp_ipc_import needs one prior advspec_tp_ipc_export on this handle");
    return ADVSPEC_ERR_STATE;
  }
  E_CUDA(e, cudaSetDevice(e->device));
  for (int r = 0; r < e->tp_size; ++r) {
    if (r == e->tp_rank) {
      e->ar_peer[r] = e->ar_region;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    cudaError_t ce = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (ce != cudaSuccess) {
      e->fail("cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(ce));
      return ADVSPEC_ERR_CUDA;
    }
    e->ar_peer[r] = static_cast<uint8_t*>(ptr);
  }
  e->ar_ready = getenv("ADVSPEC_TP_NCCL_ONLY") == nullptr;
  return ADVSPEC_OK;
}

advspec_status advspec_load_weights(advspec_engine* e, const void* host_blob, size_t bytes) {
  if (!e) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!host_blob || bytes != e->lay.total) {
    e->fail("weight blob is %zu bytes, expected %zu", bytes, e->lay.total);
    return ADVSPEC_ERR_INVALID;
  }
  E_CUDA(e, cudaSetDevice(e->device));
  E_CUDA(e, cudaMemcpyAsync(e->w, host_blob, bytes, cudaMemcpyHostToDevice, e->stream));
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  e->weights_ready = true;
  return ADVSPEC_OK;
}

advspec_status advspec_init_weights_random(advspec_engine* e, uint64_t seed, float std) {
  if (!e) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  E_CUDA(e, cudaSetDevice(e->device));
  const auto& d = e->d;
  auto rnd = [&](size_t off, size_t n, uint64_t tag) {
    init_normal_bf16_kernel<<<1184, 256, 0, e->stream>>>(reinterpret_cast<__nv_bfloat16*>(e->w + off),
                                                         (int64_t)n, mix64(seed ^ tag), std);
  };
  auto ones = [&](size_t off, size_t n, float v) {
    fill_f32_kernel<<<64, 256, 0, e->stream>>>(reinterpret_cast<float*>(e->w + off), (int64_t)n, v);
  };
  const size_t dm = d.d_model, QKV = qkv_dim(d), HD = (size_t)d.n_heads * d.head_dim;
  rnd(e->lay.embed, (size_t)e->V_full * dm, 1);  // replicated: the same table on every tensor-parallel rank
  const uint64_t rk = (uint64_t)e->tp_rank << 40;  // sharded tensors: an independent draw per rank
  for (int l = 0; l < d.n_layers; ++l) {
    const auto& o = e->lay.layers[l];
    const uint64_t t = (16 * (uint64_t)(l + 1)) ^ rk;
    ones(o.attn_norm, dm, 1.0f);
    rnd(o.wqkv, QKV * dm, t + 1);
    if (d.qkv_bias) ones(o.bqkv, QKV, 0.0f);
    rnd(o.wo, dm * HD, t + 2);
    ones(o.mlp_norm, dm, 1.0f);
    rnd(o.wgu, (size_t)2 * d.d_ff * dm, t + 3);
    rnd(o.wd, dm * (size_t)d.d_ff, t + 4);
  }
  ones(e->lay.final_norm, dm, 1.0f);
  if (!d.tied_lm_head) rnd(e->lay.lm_head, (size_t)d.vocab_size * dm, 2 ^ rk);
  E_CUDA(e, cudaGetLastError());
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  e->weights_ready = true;
  return ADVSPEC_OK;
}

advspec_status advspec_set_rope_inv_freq(advspec_engine* e, const float* inv_freq, int32_t n) {
  if (!e) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  const int half = e->d.head_dim / 2;
  if (!inv_freq || n != half) {
    e->fail("inv_freq must have head_dim/2 = %d entries", half);
    return ADVSPEC_ERR_INVALID;
  }
  E_CUDA(e, cudaSetDevice(e->device));
  E_CUDA(e, cudaMemcpyAsync(e->inv_freq, inv_freq, half * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  const int64_t tot = (int64_t)e->max_pos * half;
  rope_table_kernel<<<(int)((tot + 255) / 256), 256, 0, e->stream>>>(e->inv_freq, e->rope_cos, e->rope_sin,
                                                                    e->max_pos, half);
  E_CUDA(e, cudaGetLastError());
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  return ADVSPEC_OK;
}

// Prefill `n` tokens at positions pos_base.. against the KV of positions 0..pos_base-1 already in the
// prefix region (pos_base = 0: a fresh prompt).
static advspec_status prefill_impl(advspec_engine* e, const int32_t* tokens, int32_t n, float* all_logits_host,
                                   int32_t pos_base = 0) {
  const auto& d = e->d;
  if (!e->weights_ready) {
    e->fail("weights not loaded");
    return ADVSPEC_ERR_STATE;
  }
  if (!tokens || n < 1 || pos_base < 0 || (int64_t)pos_base + n > d.max_prefix_tokens) {
    e->fail("prompt of %d tokens at position %d outside 1..%d", n, pos_base, d.max_prefix_tokens);
    return ADVSPEC_ERR_INVALID;
  }
  for (int i = 0; i < n; ++i)
    if (tokens[i] < 0 || tokens[i] >= e->V_full) {
      e->fail("token %d at position %d outside the vocabulary", tokens[i], i);
      return ADVSPEC_ERR_INVALID;
    }
  E_CUDA(e, cudaSetDevice(e->device));
  // a new prompt replaces the live prefix and every opponent forked from it
  e->prefix_gen = 0;
  for (bool& u : e->slot_used) u = false;
  float* all_logits_dev = nullptr;
  if (all_logits_host) {
    const size_t bytes = (size_t)n * d.vocab_size * sizeof(float);
    if (bytes > ((size_t)4 << 30)) {
      e->fail("advspec_prefill_logits: %zu bytes of logits is beyond the diagnostic limit", bytes);
      return ADVSPEC_ERR_INVALID;
    }
    E_CUDA(e, cudaMalloc(reinterpret_cast<void**>(&all_logits_dev), bytes));
  }
  E_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  advspec_status st = ADVSPEC_OK;
  for (int c0 = 0; c0 < n && st == ADVSPEC_OK; c0 += e->C) {
    const int m = std::min(e->C, n - c0);
    cudaError_t ce = cudaMemcpyAsync(e->p_tokens, tokens + c0, m * sizeof(int), cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) {
      e->fail("token upload failed: %s", cudaGetErrorString(ce));
      st = ADVSPEC_ERR_CUDA;
      break;
    }
    st = prefill_chunk(e, m, pos_base + c0);
    if (st != ADVSPEC_OK) break;
    if (all_logits_dev) {
      launch_rmsnorm(e->p_x, final_norm_w(e), e->p_xn, m, d.d_model, d.norm_eps, e->stream);
      st = prefill_gemm(e, e->p_xn, d.d_model, lm_head_w(e), d.d_model,
                        all_logits_dev + (size_t)c0 * d.vocab_size, d.vocab_size, nullptr, m, d.vocab_size,
                        d.d_model, EPI_F32);
    }
    if (c0 + m == n && st == ADVSPEC_OK) {
      // next-token logits of the last prompt position: fused final-norm + lm_head GEMV
      GemvParams gl{lm_head_w(e), e->p_x + (size_t)(m - 1) * d.d_model, final_norm_w(e), nullptr,
                    e->prefill_logits, d.vocab_size, d.d_model, 1, EPI_F32, d.act, d.norm_eps};
      cudaError_t r = launch_gemv(gl, 1, e->device, e->stream, false);
      e->launches++;
      if (r != cudaSuccess) {
        e->fail("lm_head gemv failed: %s", cudaGetErrorString(r));
        st = ADVSPEC_ERR_CUDA;
      }
    }
  }
  if (st == ADVSPEC_OK) {
    cudaEventRecord(e->ev1, e->stream);
    cudaError_t r = cudaStreamSynchronize(e->stream);
    if (r != cudaSuccess) {
      e->fail("prefill failed on the device: %s", cudaGetErrorString(r));
      st = ADVSPEC_ERR_CUDA;
    } else {
      cudaEventElapsedTime(&e->tm.prefill_ms, e->ev0, e->ev1);
      st = check_watchdog(e);
    }
  }
  if (st == ADVSPEC_OK && all_logits_host) {
    cudaError_t r = cudaMemcpy(all_logits_host, all_logits_dev, (size_t)n * d.vocab_size * sizeof(float),
                               cudaMemcpyDeviceToHost);
    if (r != cudaSuccess) {
      e->fail("logits download failed: %s", cudaGetErrorString(r));
      st = ADVSPEC_ERR_CUDA;
    }
  }
  if (all_logits_dev) cudaFree(all_logits_dev);
  return st;
}

advspec_status advspec_prefill(advspec_engine* e, const int32_t* tokens, int32_t n_tokens, int32_t* prefix_id) {
  if (!e || !prefix_id) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  advspec_status st = prefill_impl(e, tokens, n_tokens, nullptr);
  if (st != ADVSPEC_OK) return st;
  static int next_gen = 1;
  e->prefix_gen = next_gen++;
  e->prefix_len = n_tokens;
  e->logits_broadcast = true;
  e->tails_open = false;
  *prefix_id = e->prefix_gen;
  return ADVSPEC_OK;
}

advspec_status advspec_prefill_extend(advspec_engine* e, int32_t prefix_id, int32_t keep_tokens,
                                      const int32_t* tokens, int32_t n_tokens, int32_t* new_prefix_id) {
  if (!e || !new_prefix_id) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->prefix_gen == 0 || prefix_id != e->prefix_gen) {
    e->fail("prefix %d is not live", prefix_id);
    return ADVSPEC_ERR_STATE;
  }
  if (keep_tokens < 0 || keep_tokens > e->prefix_len || n_tokens < 0 ||
      (n_tokens == 0 && keep_tokens != e->prefix_len) || (keep_tokens == 0 && n_tokens == 0)) {
    e->fail("prefill_extend: keep %d of a %d-token prefix with %d new tokens is not valid (keep the whole "
            "prefix to re-arm it, or append at least one token)", keep_tokens, e->prefix_len, n_tokens);
    return ADVSPEC_ERR_INVALID;
  }
  const int old_gen = e->prefix_gen;
  if (n_tokens > 0) {
    // the first keep_tokens rows of the prefix KV stay; rows past them are overwritten by the new tail
    advspec_status st = prefill_impl(e, tokens, n_tokens, nullptr, keep_tokens);
    if (st != ADVSPEC_OK) return st;  // (prefill_impl dropped the live prefix: a failed tail leaves none)
  } else {
    e->tm.prefill_ms = 0.f;
    for (bool& u : e->slot_used) u = false;
  }
  (void)old_gen;
  static int next_gen = 1 << 24;
  e->prefix_gen = next_gen++;
  e->prefix_len = keep_tokens + n_tokens;
  e->logits_broadcast = true;
  e->tails_open = false;
  *new_prefix_id = e->prefix_gen;
  return ADVSPEC_OK;
}

advspec_status advspec_prefill_logits(advspec_engine* e, const int32_t* tokens, int32_t n_tokens, float* out) {
  if (!e || !out) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  return prefill_impl(e, tokens, n_tokens, out);
}

advspec_status advspec_fork(advspec_engine* e, int32_t prefix_id, int32_t n_seqs, const uint64_t* seeds,
                            int32_t* seq_ids) {
  if (!e || !seeds || !seq_ids) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->prefix_gen == 0 || prefix_id != e->prefix_gen) {
    e->fail("prefix %d is not live", prefix_id);
    return ADVSPEC_ERR_STATE;
  }
  int free_slots = 0;
  for (int s = 0; s < e->d.max_seqs; ++s) free_slots += e->slot_used[s] ? 0 : 1;
  if (n_seqs < 1 || n_seqs > free_slots) {
    e->fail("cannot fork %d opponents: %d slots free", n_seqs, free_slots);
    return ADVSPEC_ERR_INVALID;
  }
  E_CUDA(e, cudaSetDevice(e->device));
  int k = 0;
  for (int s = 0; s < e->d.max_seqs && k < n_seqs; ++s) {
    if (e->slot_used[s]) continue;
    e->slot_used[s] = true;
    const int zero = 0;
    E_CUDA(e, cudaMemcpyAsync(e->s_seeds + s, &seeds[k], sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
    E_CUDA(e, cudaMemcpyAsync(e->s_suf_len + s, &zero, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    E_CUDA(e, cudaMemcpyAsync(e->s_n_out + s, &zero, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    E_CUDA(e, cudaMemcpyAsync(e->s_done + s, &zero, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    E_CUDA(e, cudaStreamSynchronize(e->stream));
    e->h_suf_len[s] = 0;
    seq_ids[k++] = s;
  }
  return ADVSPEC_OK;
}

advspec_status advspec_release_seqs(advspec_engine* e, const int32_t* seq_ids, int32_t n) {
  if (!e || !seq_ids) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  for (int i = 0; i < n; ++i)
    if (seq_ids[i] >= 0 && seq_ids[i] < e->d.max_seqs) e->slot_used[seq_ids[i]] = false;
  return ADVSPEC_OK;
}

advspec_status advspec_release_prefix(advspec_engine* e, int32_t prefix_id) {
  if (!e) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (prefix_id == e->prefix_gen) {
    e->prefix_gen = 0;
    for (bool& u : e->slot_used) u = false;
  }
  return ADVSPEC_OK;
}

advspec_status advspec_decode(advspec_engine* e, const int32_t* seq_ids, int32_t n, int32_t max_new,
                              float temperature, int32_t eos_id, int32_t* out_tokens, int32_t* out_lens) {
  if (!e || !out_tokens || !out_lens) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  const auto& d = e->d;
  E_CUDA(e, cudaSetDevice(e->device));
  std::vector<int> slots;
  advspec_status st = setup_batch(e, seq_ids, n, &slots);
  if (st != ADVSPEC_OK) return st;
  int room = d.max_new_tokens;
  for (int s : slots) room = std::min(room, d.max_new_tokens - e->h_suf_len[s]);
  if (max_new < 1 || max_new > room) {
    e->fail("max_new %d outside 1..%d (suffix KV capacity left)", max_new, room);
    return ADVSPEC_ERR_INVALID;
  }
  if (!e->logits_broadcast && e->logits_slots != slots) {
    e->fail("decode after decode_step / append_tail must use the same opponents in the same order");
    return ADVSPEC_ERR_STATE;
  }
  const int zero = 0;
  for (int s : slots) {
    E_CUDA(e, cudaMemcpyAsync(e->s_n_out + s, &zero, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    E_CUDA(e, cudaMemcpyAsync(e->s_done + s, &zero, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  }
  E_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  // token 0 comes from the logits already on the device (the shared prefill's, or the last step's)
  SampleParams sp0 = make_sample_params(e, temperature, eos_id, e->logits_broadcast, false, nullptr, true);
  E_CUDA(e, launch_sampler(e, sp0, n, false));

  const int steps = max_new - 1;
  const SampleParams sp = make_sample_params(e, temperature, eos_id, false, true, nullptr, true);
  if (steps > 0) {
    std::vector<int> key = slots;
    key.push_back(-1);
    key.push_back(e->prefix_len);
    key.push_back(e->prefix_gen);
    { int tbits; memcpy(&tbits, &temperature, sizeof tbits); key.push_back(tbits); }
    key.push_back(eos_id);
    if (e->use_graph) {
      if (!e->graph || e->graph_key != key) {
        if (e->graph) {
          cudaGraphExecDestroy(e->graph);
          e->graph = nullptr;
        }
        cudaGraph_t g = nullptr;
        E_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        const int64_t before = e->launches;
        advspec_status fs = enqueue_forward(e, n, nullptr);
        cudaError_t le = launch_sampler(e, sp, n, true);
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        e->graph_launches = e->launches - before;  // what one replay launches (counted while recording)
        e->launches = before;                      // capture only records; launches are counted per replay
        if (fs != ADVSPEC_OK || le != cudaSuccess || ce != cudaSuccess) {
          if (g) cudaGraphDestroy(g);
          if (fs == ADVSPEC_OK)
            e->fail("graph capture failed: %s", cudaGetErrorString(le != cudaSuccess ? le : ce));
          return ADVSPEC_ERR_CUDA;
        }
        cudaError_t ie = cudaGraphInstantiate(&e->graph, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) {
          e->graph = nullptr;
          e->fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
          return ADVSPEC_ERR_CUDA;
        }
        e->graph_key = key;
      }
    }
    for (int s = 0; s < steps; ++s) {
      if (e->use_graph) {
        E_CUDA(e, cudaGraphLaunch(e->graph, e->stream));
        e->launches += e->graph_launches;
      } else {
        advspec_status fs = enqueue_forward(e, n, nullptr);
        if (fs != ADVSPEC_OK) return fs;
        E_CUDA(e, launch_sampler(e, sp, n, true));
      }
      if (eos_id >= 0 && (s % 32) == 31) {
        int done_h[8];
        E_CUDA(e, cudaMemcpyAsync(done_h, e->s_done, d.max_seqs * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        E_CUDA(e, cudaStreamSynchronize(e->stream));
        bool all = true;
        for (int sl : slots) all = all && done_h[sl] != 0;
        if (all) {
          e->tm.decode_steps = s + 1;
          break;
        }
      }
      e->tm.decode_steps = s + 1;
    }
  } else {
    e->tm.decode_steps = 0;
  }
  E_CUDA(e, cudaEventRecord(e->ev1, e->stream));
  std::vector<int> h_out((size_t)d.max_seqs * d.max_new_tokens);
  int n_out_h[8], suf_h[8];
  E_CUDA(e, cudaMemcpyAsync(h_out.data(), e->s_out, h_out.size() * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  E_CUDA(e, cudaMemcpyAsync(n_out_h, e->s_n_out, d.max_seqs * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  E_CUDA(e, cudaMemcpyAsync(suf_h, e->s_suf_len, d.max_seqs * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  E_CUDA(e, cudaEventElapsedTime(&e->tm.decode_ms, e->ev0, e->ev1));
  e->tm.decode_batch = n;
  st = check_watchdog(e);
  if (st != ADVSPEC_OK) return st;
  for (int i = 0; i < n; ++i) {
    const int s = slots[i];
    out_lens[i] = n_out_h[s];
    e->h_suf_len[s] = suf_h[s];
    for (int t = 0; t < max_new; ++t)
      out_tokens[(size_t)i * max_new + t] = t < n_out_h[s] ? h_out[(size_t)s * d.max_new_tokens + t] : -1;
  }
  e->logits_broadcast = false;
  e->logits_slots = slots;
  e->tails_open = false;
  return ADVSPEC_OK;
}

advspec_status advspec_append_tail(advspec_engine* e, int32_t seq_id, const int32_t* tokens, int32_t n_tokens) {
  if (!e || !tokens) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  const auto& d = e->d;
  if (e->prefix_gen == 0) {
    e->fail("append_tail: no live prefix");
    return ADVSPEC_ERR_STATE;
  }
  if (seq_id < 0 || seq_id >= d.max_seqs || !e->slot_used[seq_id]) {
    e->fail("append_tail: opponent %d is not forked", seq_id);
    return ADVSPEC_ERR_INVALID;
  }
  if (e->h_suf_len[seq_id] != 0) {
    e->fail("append_tail: opponent %d already holds %d tokens of its own; a tail directly follows the fork", seq_id,
            e->h_suf_len[seq_id]);
    return ADVSPEC_ERR_STATE;
  }
  if (n_tokens < 1 || n_tokens > d.max_new_tokens) {
    e->fail("append_tail: %d tokens outside 1..%d (suffix KV capacity)", n_tokens, d.max_new_tokens);
    return ADVSPEC_ERR_INVALID;
  }
  if (!e->tails_open) e->logits_slots.clear();
  if ((int)e->logits_slots.size() >= d.max_seqs) {
    e->fail("append_tail: more tails than opponents");
    return ADVSPEC_ERR_STATE;
  }
  // The tail runs as a prompt chunk at positions prefix_len.. against the prefix KV; its own K/V lands in the
  // prefix region's rows PAST the live prefix (scratch until the next tail or extension), the prefix and its
  // forks stay as they are.
  const int gen = e->prefix_gen;
  bool used[8];
  std::copy(std::begin(e->slot_used), std::end(e->slot_used), used);
  const float prefix_ms = e->tm.prefill_ms;
  advspec_status st = prefill_impl(e, tokens, n_tokens, nullptr, e->prefix_len);
  e->prefix_gen = gen;
  std::copy(used, used + 8, std::begin(e->slot_used));
  if (st != ADVSPEC_OK) return st;
  e->tm.prefill_ms += prefix_ms;  // prompt processing of the round: shared prefix + tails
  kv_tail_to_suffix_kernel<<<dim3(d.n_layers * 2, d.n_kv_heads), 256, 0, e->stream>>>(
      e->pkv, e->skv, e->pkv_layer_elems, e->skv_layer_elems, d.n_kv_heads, (int64_t)d.max_prefix_tokens,
      (int64_t)d.max_new_tokens, d.head_dim, seq_id, e->prefix_len, n_tokens);
  E_CUDA(e, cudaGetLastError());
  e->launches++;
  const int row = (int)e->logits_slots.size();
  E_CUDA(e, cudaMemcpyAsync(e->dlogits + (size_t)row * d.vocab_size, e->prefill_logits,
                            (size_t)d.vocab_size * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
  e->h_suf_len[seq_id] = n_tokens;
  E_CUDA(e, cudaMemcpyAsync(e->s_suf_len + seq_id, &e->h_suf_len[seq_id], sizeof(int), cudaMemcpyHostToDevice,
                            e->stream));
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  e->logits_slots.push_back(seq_id);
  e->logits_broadcast = false;
  e->tails_open = true;
  return check_watchdog(e);
}

advspec_status advspec_decode_step(advspec_engine* e, const int32_t* seq_ids, int32_t n,
                                   const int32_t* forced_tokens) {
  if (!e || !forced_tokens) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  const auto& d = e->d;
  E_CUDA(e, cudaSetDevice(e->device));
  std::vector<int> slots;
  advspec_status st = setup_batch(e, seq_ids, n, &slots);
  if (st != ADVSPEC_OK) return st;
  for (int i = 0; i < n; ++i) {
    if (forced_tokens[i] < 0 || forced_tokens[i] >= e->V_full) {
      e->fail("forced token %d outside the vocabulary", forced_tokens[i]);
      return ADVSPEC_ERR_INVALID;
    }
    if (e->h_suf_len[slots[i]] >= d.max_new_tokens) {
      e->fail("opponent %d has no suffix KV capacity left", slots[i]);
      return ADVSPEC_ERR_INVALID;
    }
  }
  E_CUDA(e, cudaMemcpyAsync(e->s_forced, forced_tokens, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  SampleParams sp = make_sample_params(e, 0.f, -1, false, false, e->s_forced, false);
  E_CUDA(e, launch_sampler(e, sp, n, false));
  ADV_TRACE(e->stream, "sample(forced)");
  st = enqueue_forward(e, n, nullptr);
  if (st != ADVSPEC_OK) return st;
  advance_kernel<<<1, n, 0, e->stream>>>(e->s_slots, e->s_suf_len);
  E_CUDA(e, cudaGetLastError());
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  for (int s : slots) e->h_suf_len[s] += 1;
  e->logits_broadcast = false;
  e->logits_slots = slots;
  e->tails_open = false;
  return check_watchdog(e);
}

advspec_status advspec_get_logits(advspec_engine* e, int32_t n, float* out) {
  if (!e || !out || n < 1) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  E_CUDA(e, cudaSetDevice(e->device));
  const size_t V = e->d.vocab_size;
  if (e->logits_broadcast) {
    if (n != 1) {
      e->fail("prefill logits have one row");
      return ADVSPEC_ERR_INVALID;
    }
    E_CUDA(e, cudaMemcpy(out, e->prefill_logits, V * sizeof(float), cudaMemcpyDeviceToHost));
  } else {
    if (n > (int)e->logits_slots.size()) {
      e->fail("only %zu logit rows are available", e->logits_slots.size());
      return ADVSPEC_ERR_INVALID;
    }
    E_CUDA(e, cudaMemcpy(out, e->dlogits, (size_t)n * V * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return ADVSPEC_OK;
}

advspec_status advspec_prefix_kv_region(advspec_engine* e, int32_t prefix_id, void** dev_ptr, size_t* bytes) {
  if (!e || !dev_ptr || !bytes) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->prefix_gen == 0 || prefix_id != e->prefix_gen) {
    e->fail("prefix %d is not live", prefix_id);
    return ADVSPEC_ERR_STATE;
  }
  *dev_ptr = e->pkv;
  *bytes = (size_t)e->d.n_layers * 2 * e->pkv_layer_elems * sizeof(__nv_bfloat16);
  return ADVSPEC_OK;
}

advspec_status advspec_prefix_adopt(advspec_engine* e, int32_t n_tokens, const float* logits, int32_t* prefix_id) {
  if (!e || !logits || !prefix_id) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  if (n_tokens < 1 || n_tokens > e->d.max_prefix_tokens) {
    e->fail("adopted prefix of %d tokens outside 1..%d", n_tokens, e->d.max_prefix_tokens);
    return ADVSPEC_ERR_INVALID;
  }
  E_CUDA(e, cudaSetDevice(e->device));
  E_CUDA(e, cudaMemcpy(e->prefill_logits, logits, (size_t)e->d.vocab_size * sizeof(float), cudaMemcpyHostToDevice));
  for (bool& u : e->slot_used) u = false;
  static int next_gen = 1 << 20;
  e->prefix_gen = next_gen++;
  e->prefix_len = n_tokens;
  e->logits_broadcast = true;
  e->tails_open = false;
  *prefix_id = e->prefix_gen;
  return ADVSPEC_OK;
}

advspec_status advspec_get_timing(advspec_engine* e, advspec_timing* out) {
  if (!e || !out) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  e->tm.kernel_launches = e->launches;
  *out = e->tm;
  return ADVSPEC_OK;
}

advspec_status advspec_profile_decode_step(advspec_engine* e, const int32_t* seq_ids, int32_t n) {
  if (!e) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  E_CUDA(e, cudaSetDevice(e->device));
  std::vector<int> slots;
  advspec_status st = setup_batch(e, seq_ids, n, &slots);
  if (st != ADVSPEC_OK) return st;
  for (int s : slots)
    if (e->h_suf_len[s] >= e->d.max_new_tokens) {
      e->fail("opponent %d has no suffix KV capacity left for a profiled step", s);
      return ADVSPEC_ERR_INVALID;
    }
  const size_t xb = (size_t)e->d.max_seqs * e->d.d_model * sizeof(float);
  E_CUDA(e, cudaMemcpyAsync(e->dx_save, e->dx, xb, cudaMemcpyDeviceToDevice, e->stream));
  // warm pass, then the timed pass (suffix length is not advanced: the same KV row is rewritten)
  st = enqueue_forward(e, n, nullptr);
  if (st != ADVSPEC_OK) return st;
  E_CUDA(e, cudaMemcpyAsync(e->dx, e->dx_save, xb, cudaMemcpyDeviceToDevice, e->stream));
  float ms = 0.f;
  st = enqueue_forward(e, n, &ms);
  if (st != ADVSPEC_OK) return st;
  e->tm.gemv_ms = ms;
  E_CUDA(e, cudaMemcpyAsync(e->dx, e->dx_save, xb, cudaMemcpyDeviceToDevice, e->stream));
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  return check_watchdog(e);
}

advspec_status advspec_decode_step_bytes(advspec_engine* e, const int32_t* seq_ids, int32_t n,
                                         double* step_bytes, double* gemv_bytes) {
  if (!e || !seq_ids || !step_bytes || !gemv_bytes) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  const auto& d = e->d;
  const double dm = d.d_model, QKV = qkv_dim(d), HD = (double)d.n_heads * d.head_dim;
  // every matmul weight once, lm_head streamed, embedding table gathered (not streamed)
  const double w = 2.0 * (d.n_layers * (QKV * dm + dm * HD + 2.0 * d.d_ff * dm + dm * d.d_ff) +
                          (double)d.vocab_size * dm);
  const double kvB = 2.0 * d.n_layers * d.n_kv_heads * d.head_dim * 2.0;
  double toks = e->prefix_len;
  for (int i = 0; i < n; ++i) {
    const int s = seq_ids[i];
    if (s < 0 || s >= d.max_seqs) return ADVSPEC_ERR_INVALID;
    toks += e->h_suf_len[s];
  }
  *gemv_bytes = w;
  *step_bytes = w + kvB * toks + kvB * n;
  return ADVSPEC_OK;
}

advspec_status advspec_ktrace_enable(advspec_engine* e, int32_t on) {
  if (!e) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  E_CUDA(e, cudaSetDevice(e->device));
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  const int v = on ? 1 : 0;
  const unsigned int zero = 0;
  E_CUDA(e, cudaMemcpyToSymbol(g_ktrace_on, &v, sizeof v));
  E_CUDA(e, cudaMemcpyToSymbol(g_ktrace_n, &zero, sizeof zero));
  return ADVSPEC_OK;
}

advspec_status advspec_ktrace_read(advspec_engine* e, uint64_t* out, int32_t cap, int32_t* n) {
  if (!e || !out || !n || cap < 0) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  E_CUDA(e, cudaSetDevice(e->device));
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  unsigned int cnt = 0;
  E_CUDA(e, cudaMemcpyFromSymbol(&cnt, g_ktrace_n, sizeof cnt));
  const int m = (int)std::min<unsigned int>(std::min<unsigned int>(cnt, (unsigned int)kTraceCap), (unsigned int)cap);
  if (m > 0) E_CUDA(e, cudaMemcpyFromSymbol(out, g_ktrace, (size_t)m * sizeof(uint64_t)));
  const unsigned int zero = 0;
  E_CUDA(e, cudaMemcpyToSymbol(g_ktrace_n, &zero, sizeof zero));
  *n = m;
  return ADVSPEC_OK;
}

advspec_status advspec_ktrace_phases(advspec_engine* e, uint64_t* out16) {
  if (!e || !out16) return ADVSPEC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  E_CUDA(e, cudaSetDevice(e->device));
  E_CUDA(e, cudaStreamSynchronize(e->stream));
  E_CUDA(e, cudaMemcpyFromSymbol(out16, g_phase, 16 * sizeof(uint64_t)));
  return ADVSPEC_OK;
}

// ------------------------------------------------------------- op-level
static advspec_status op_begin(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    g_create_error = "no CUDA device for op-level call";
    return ADVSPEC_ERR_CUDA;
  }
  if (cudaSetDevice(device) != cudaSuccess) return ADVSPEC_ERR_CUDA;
  return ADVSPEC_OK;
}
static advspec_status op_end(const char* what) {
  cudaError_t r = cudaDeviceSynchronize();
  if (r != cudaSuccess) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(r);
    return ADVSPEC_ERR_CUDA;
  }
  unsigned int code = 0;
  cudaMemcpyFromSymbol(&code, g_watchdog_code, sizeof code);
  if (code) {
    char buf[128];
    snprintf(buf, sizeof buf, "%s: device watchdog tripped at site 0x%x", what, code & 0x7fffffffu);
    g_create_error = buf;
    unsigned int zero = 0;
    cudaMemcpyToSymbol(g_watchdog_code, &zero, sizeof zero);
    return ADVSPEC_ERR_KERNEL;
  }
  return ADVSPEC_OK;
}

advspec_status advspec_op_gemm(int32_t device, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                               int64_t ldc, const void* aux, int32_t M, int32_t N, int32_t K, int32_t epilogue,
                               int32_t act) {
  advspec_status s = op_begin(device);
  if (s) return s;
  GemmParams p{C, ldc, reinterpret_cast<const float*>(aux), M, N, K, act};
  std::string why;
  {  // the A/B knobs are re-read per call (unset = default)
    const char* gb = getenv("ADVSPEC_GEMM_BAND_MB");
    g_gemm_band_mb = gb ? std::max(0, atoi(gb)) : 0;
    const char* gs = getenv("ADVSPEC_GEMM_SPLITK");
    g_gemm_splitk = gs ? atoi(gs) != 0 : 1;
  }
  unsigned int* sem = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&sem), kGemmSemWords * sizeof(unsigned int)) != cudaSuccess) return ADVSPEC_ERR_OOM;
  cudaMemset(sem, 0, kGemmSemWords * sizeof(unsigned int));
  cudaError_t r = launch_gemm(A, lda, M, B, ldb, p, epilogue, device, 0, &why, sem);
  if (r != cudaSuccess) {
    cudaFree(sem);
    g_create_error = std::string("op_gemm launch: ") + cudaGetErrorString(r) + " " + why;
    return ADVSPEC_ERR_CUDA;
  }
  s = op_end("op_gemm");
  cudaFree(sem);
  return s;
}

advspec_status advspec_op_gemm_check(int32_t device, const void* A, int64_t lda, const void* B, int64_t ldb,
                                     void* C, int64_t ldc, const void* aux, int32_t M, int32_t N, int32_t K,
                                     int32_t epilogue, int32_t act) {
  advspec_status s = op_begin(device);
  if (s) return s;
  GemmParams p{C, ldc, reinterpret_cast<const float*>(aux), M, N, K, act};
  cudaError_t r = launch_gemm_check(A, lda, B, ldb, p, epilogue, 0);
  if (r != cudaSuccess) {
    g_create_error = std::string("op_gemm_check launch: ") + cudaGetErrorString(r);
    return ADVSPEC_ERR_CUDA;
  }
  return op_end("op_gemm_check");
}

advspec_status advspec_op_gemv(int32_t device, const void* W, const void* x, const void* norm_w,
                               const void* bias, void* y, int32_t b, int32_t N, int32_t K, int32_t in_mode,
                               int32_t epilogue, int32_t act, float eps) {
  advspec_status s = op_begin(device);
  if (s) return s;
  if (b < 1 || b > 8 || (K % 8) || (epilogue == EPI_GATED_BF16 && (N % 2))) {
    g_create_error = "op_gemv: need 1 <= b <= 8, K % 8 == 0, even N for the gated epilogue";
    return ADVSPEC_ERR_INVALID;
  }
  GemvParams p{reinterpret_cast<const __nv_bfloat16*>(W), x, reinterpret_cast<const float*>(norm_w),
               reinterpret_cast<const float*>(bias), y, N, K, in_mode, epilogue, act, eps};
  if (const char* gi = getenv("ADVSPEC_GEMV_IMPL")) g_gemv_impl = atoi(gi) == 1 ? 1 : 3;
  if (const char* ai = getenv("ADVSPEC_ATTN_IMPL")) g_attn_impl = atoi(ai) == 1 ? 1 : 2;
  if (const char* xm = getenv("ADVSPEC_X_SMEM_MAX")) g_x_smem_max = (size_t)atoll(xm);
  g_gemm_narrow = getenv("ADVSPEC_GEMM_NARROW") != nullptr;
  {
    const char* tc = getenv("ADVSPEC_ATTN_PREFILL_TC");
    g_attn_prefill_tc = tc ? atoi(tc) != 0 : true;
    }
  if (const char* ms = getenv("ADVSPEC_ATTN_MIN_SPLIT")) g_attn_min_split = std::max(64, atoi(ms));
  cudaError_t r = launch_gemv(p, b, device, 0, false);
  if (r != cudaSuccess) {
    g_create_error = std::string("op_gemv launch: ") + cudaGetErrorString(r);
    return ADVSPEC_ERR_CUDA;
  }
  return op_end("op_gemv");
}

advspec_status advspec_op_attn_prefill(int32_t device, const void* q, int64_t ldq, const void* kcache,
                                       const void* vcache, int64_t kv_stride, void* out, int32_t n_q,
                                       int32_t q_pos0, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim,
                                       int32_t impl) {
  advspec_status s = op_begin(device);
  if (s) return s;
  AttnPrefillParams p{reinterpret_cast<const __nv_bfloat16*>(q), ldq,
                      reinterpret_cast<const __nv_bfloat16*>(kcache),
                      reinterpret_cast<const __nv_bfloat16*>(vcache), kv_stride,
                      reinterpret_cast<__nv_bfloat16*>(out), n_q, q_pos0, n_heads, n_kv_heads,
                      1.0f / sqrtf((float)head_dim), head_dim};
  cudaError_t r;
  if (impl == 2) {
    if (head_dim != 128 && head_dim != 96 && head_dim != 64) {
      g_create_error = "op_attn_prefill: the tcgen05 kernel serves head_dim 64, 96 and 128";
      return ADVSPEC_ERR_INVALID;
    }
    std::string why;
    r = launch_attn_prefill_tc(q, ldq, kcache, vcache, kv_stride, out, n_q, q_pos0, n_heads, n_kv_heads, head_dim, 0,
                               &why);
  } else {
    r = launch_attn_prefill(p, head_dim, impl, 0, nullptr);
  }
  if (r != cudaSuccess) {
    g_create_error = std::string("op_attn_prefill launch: ") + cudaGetErrorString(r);
    return ADVSPEC_ERR_CUDA;
  }
  return op_end("op_attn_prefill");
}

advspec_status advspec_op_attn_decode(int32_t device, const void* qkv, const void* rope_cos, const void* rope_sin,
                                      const void* prefix_k, const void* prefix_v, int64_t prefix_stride,
                                      int32_t prefix_len, void* suffix_k, void* suffix_v, int64_t suffix_stride,
                                      const int32_t* pos_host, void* out, int32_t b, int32_t n_heads,
                                      int32_t n_kv_heads, int32_t head_dim) {
  advspec_status s = op_begin(device);
  if (s) return s;
  const int G = n_kv_heads > 0 ? n_heads / n_kv_heads : 0;
  if (b < 1 || b > 8 || n_kv_heads < 1 || n_heads % n_kv_heads || G > 16 || n_heads > 255 || !pos_host ||
      (head_dim != 64 && head_dim != 96 && head_dim != 128 && head_dim != 256) || prefix_len < 1 ||
      prefix_len > prefix_stride) {
    g_create_error = "op_attn_decode: need 1 <= b <= 8, heads/kv_heads <= 16, head_dim in {64,96,128,256}, "
                     "1 <= prefix_len <= prefix_stride";
    return ADVSPEC_ERR_INVALID;
  }
  for (int i = 0; i < b; ++i)
    if (pos_host[i] < prefix_len || pos_host[i] - prefix_len >= suffix_stride) {
      g_create_error = "op_attn_decode: pos[i] must be prefix_len + (suffix tokens already stored) < suffix capacity";
      return ADVSPEC_ERR_INVALID;
    }
  if (const char* ms = getenv("ADVSPEC_ATTN_MIN_SPLIT")) g_attn_min_split = std::max(64, atoi(ms));
  const Attn2Plan pl = plan_attn2_shape(b, n_heads, n_kv_heads, prefix_len, device);
  CUtensorMap hm[2];
  const int64_t rows = (int64_t)n_kv_heads * prefix_stride;
  if (!make_tmap(&hm[0], prefix_k, rows, head_dim, head_dim, 64) ||
      !make_tmap(&hm[1], prefix_v, rows, head_dim, head_dim, 64)) {
    g_create_error = "op_attn_decode: cuTensorMapEncodeTiled failed";
    return ADVSPEC_ERR_CUDA;
  }
  int* dpos = nullptr;
  float *pm = nullptr, *plv = nullptr, *po = nullptr;
  const size_t nrow = (size_t)b * n_heads * pl.n_slots;
  auto cleanup = [&]() {
    cudaFree(dpos);
    cudaFree(pm);
    cudaFree(plv);
    cudaFree(po);
  };
  if (cudaMalloc(reinterpret_cast<void**>(&dpos), 8 * sizeof(int)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&pm), nrow * 4) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&plv), nrow * 4) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&po), nrow * head_dim * 4) != cudaSuccess) {
    cleanup();
    g_create_error = "op_attn_decode: out of device memory";
    return ADVSPEC_ERR_OOM;
  }
  cudaMemcpy(dpos, pos_host, b * sizeof(int), cudaMemcpyHostToDevice);
  AttnDecode2Params a2{};
  a2.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  a2.rope_cos = reinterpret_cast<const float*>(rope_cos);
  a2.rope_sin = reinterpret_cast<const float*>(rope_sin);
  a2.pk = reinterpret_cast<const __nv_bfloat16*>(prefix_k);
  a2.pv = reinterpret_cast<const __nv_bfloat16*>(prefix_v);
  a2.pstride = prefix_stride;
  a2.sk = reinterpret_cast<__nv_bfloat16*>(suffix_k);
  a2.sv = reinterpret_cast<__nv_bfloat16*>(suffix_v);
  a2.sstride = suffix_stride;
  a2.map_k = hm[0];
  a2.map_v = hm[1];
  a2.pos_b = dpos;
  for (int i = 0; i < 8; ++i) a2.slots[i] = i < b ? i : 0;
  a2.prefix_len = prefix_len;
  a2.part_m = pm;
  a2.part_l = plv;
  a2.part_o = po;
  a2.b = b;
  a2.H = n_heads;
  a2.Hkv = n_kv_heads;
  a2.G = G;
  a2.opg = pl.opg;
  a2.n_og = pl.n_og;
  a2.n_splits = pl.n_splits;
  a2.n_slots = pl.n_slots;
  a2.scale = 1.0f / sqrtf((float)head_dim);
  a2.dh = head_dim;
  cudaError_t r = launch_attn_decode2(a2, pl.ctas, head_dim, 0, false);
  if (r == cudaSuccess)
    r = launch_pdl(attn_decode_combine2_kernel, dim3(b * n_heads), dim3(128), 0, (cudaStream_t)0, false,
                   (const float*)pm, (const float*)plv, (const float*)po, reinterpret_cast<__nv_bfloat16*>(out),
                   pl.n_slots, (int)head_dim);
  if (r != cudaSuccess) {
    cleanup();
    g_create_error = std::string("op_attn_decode launch: ") + cudaGetErrorString(r);
    return ADVSPEC_ERR_CUDA;
  }
  s = op_end("op_attn_decode");
  cleanup();
  return s;
}

}  // extern "C"
