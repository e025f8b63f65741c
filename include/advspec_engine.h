/*
 * advspec_engine.h — C ABI of the B200 opponent-critique engine.
 *
 * What this replaces.  The reference (zscole/adversarial-spec) has NO native
 * boundary: its fan-out ends in a Python name lookup,
 *     completion(model=, messages=, max_tokens=, timeout=, temperature=)
 *         skills/adversarial-spec/scripts/models.py:614-628   (seam B1)
 * called N times concurrently by
 *     call_models_parallel(models, spec, round_num, doc_type, ...)
 *         skills/adversarial-spec/scripts/models.py:681-722   (seam B2)
 * Everything below `completion` is remote third-party inference (litellm).
 * This header is the native surface a maintainer binds (ctypes, see
 * INTEGRATION.md) so that one `call_models_parallel` round becomes
 *     prefill(shared prompt) -> fork(N opponents) -> decode(batched)
 * on the local GPU.  Each entry point cites the reference behaviour it
 * stands in for.
 *
 * Conventions: extern "C"; plain pointers and sizes; every buffer is
 * caller-owned; functions return an advspec_status (0 = OK) and never throw;
 * `advspec_last_error` gives the message for the last failure on a handle
 * (or the create-time failure when the handle is NULL).  One engine = one
 * model's weights on one CUDA device.  Calls on one handle are serialised by
 * an internal mutex (the reference enters seam B1 from N threads at once,
 * models.py:699); different handles are independent.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with
 * ADVSPEC_ERR_CUDA.
 */
#ifndef ADVSPEC_ENGINE_H
#define ADVSPEC_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVSPEC_ABI_VERSION 1

typedef enum advspec_status {
  ADVSPEC_OK = 0,
  ADVSPEC_ERR_INVALID = 1,   /* bad argument / unsupported shape            */
  ADVSPEC_ERR_CUDA = 2,      /* CUDA runtime/driver failure, or no device   */
  ADVSPEC_ERR_OOM = 3,       /* device allocation failed                    */
  ADVSPEC_ERR_STATE = 4,     /* call out of order (no weights, bad ids ...) */
  ADVSPEC_ERR_KERNEL = 5     /* device-side watchdog / self-check tripped   */
} advspec_status;

typedef struct advspec_engine advspec_engine; /* opaque */

/* Decoder-only transformer description (Llama / Mistral / Qwen2 / Phi-3 /
 * Gemma shapes; SURVEY.md §8(d) model table).  Fixed 96-byte POD. */
typedef struct advspec_model_desc {
  int32_t abi_version;        /* ADVSPEC_ABI_VERSION */
  int32_t n_layers;
  int32_t d_model;
  int32_t n_heads;
  int32_t n_kv_heads;
  int32_t head_dim;           /* 64, 96, 128 or 256 */
  int32_t d_ff;
  int32_t vocab_size;
  int32_t act;                /* 0 = SiLU-gated (SwiGLU), 1 = tanh-GELU-gated (GeGLU) */
  int32_t qkv_bias;           /* Qwen2: 1 */
  int32_t tied_lm_head;       /* Gemma: 1 (lm_head == embedding table) */
  int32_t max_prefix_tokens;  /* capacity of one shared-prefix KV region */
  int32_t max_new_tokens;     /* capacity of one opponent's private KV suffix */
  int32_t max_seqs;           /* max opponents decoded together (<= 8) */
  int32_t tp_rank;            /* tensor parallelism: this handle's rank ... */
  int32_t tp_size;            /* ... of 1, 2, 4 or 8 (one process per GPU); see advspec_tp_init */
  float rope_theta;
  float norm_eps;
  float embed_scale;          /* 1.0, or sqrt(d_model) for Gemma */
  float reserved_f;
  int32_t reserved_i[4];
} advspec_model_desc;

/* Device-time breakdown of the most recent prefill / decode on a handle,
 * measured with CUDA events on the engine's own stream. */
typedef struct advspec_timing {
  float prefill_ms;        /* last advspec_prefill, H2D of tokens excluded */
  float decode_ms;         /* last advspec_decode, all steps */
  int32_t decode_steps;    /* steps executed by the last advspec_decode */
  int32_t decode_batch;    /* opponents decoded together */
  int64_t kernel_launches; /* engine kernels launched since create */
  float gemv_ms;           /* summed device time of the weight-streaming GEMV
                              kernel in the last *profiled* decode step */
  int32_t gemv_launches;   /* launches behind gemv_ms */
  int32_t reserved;
} advspec_timing;

/* ---- lifecycle ---------------------------------------------------------- */

/* Bytes of the weight blob `advspec_load_weights` expects for this shape
 * (layout: DESIGN.md "Weight blob").  Pure arithmetic; works without a GPU. */
size_t advspec_weight_blob_bytes(const advspec_model_desc *desc);

/* Byte offset of a named tensor inside the blob; layer = -1 for the globals
 * ("embed", "final_norm", "lm_head"), else "attn_norm", "wqkv", "bqkv", "wo",
 * "mlp_norm", "wgu", "wd".  Returns (size_t)-1 if absent.  No GPU needed. */
size_t advspec_weight_offset(const advspec_model_desc *desc, int32_t layer,
                             const char *name);

/* Stands in for "a provider exists for this model string" — the reference
 * resolves that remotely inside litellm (models.py:628).  Allocates weights,
 * KV regions and workspaces on `device`. */
advspec_status advspec_engine_create(const advspec_model_desc *desc,
                                     int32_t device, advspec_engine **out);
void advspec_engine_destroy(advspec_engine *e);
const char *advspec_last_error(const advspec_engine *e);

/* ---- tensor parallelism (one large opponent over tp_size GPUs) ----
 * The reference has no counterpart (a remote provider hides its own sharding behind
 * models.py:628).  `desc` always describes the WHOLE model; a handle created with
 * tp_size > 1 holds rank tp_rank's share: n_heads/tp query heads and n_kv_heads/tp KV
 * heads (wqkv rows, wo columns), d_ff/tp MLP columns (wgu rows, wd columns) and
 * vocab_size/tp rows of lm_head; norm vectors and the embedding table are whole.
 * advspec_weight_blob_bytes / advspec_weight_offset answer for that share.
 * advspec_get_logits / advspec_prefill_logits return the rank's vocab_size/tp columns.
 * Every rank makes the same calls with the same tokens and seeds; the ranks exchange
 * the residual stream (all-reduce after o-proj and down-proj) and the sampler's
 * per-rank winners (all-gather) over NCCL on the engine's stream, and return
 * identical tokens.
 *
 * advspec_tp_unique_id: rank 0 obtains the 128-byte NCCL id and hands it to the other
 * ranks by any means (the host layer uses torch.distributed).  advspec_tp_init: called
 * on every rank's handle, concurrently; blocks until all tp_size ranks have joined.
 * NCCL is loaded at run time (libnccl.so.2, or the path in ADVSPEC_NCCL_LIB). */
advspec_status advspec_tp_unique_id(uint8_t *out128);
advspec_status advspec_tp_init(advspec_engine *e, const uint8_t *id128);
/* Optional, after advspec_tp_init: the decode step's residual exchange (b x d_model floats, 2 per
 * layer per token: latency-bound) runs as ONE kernel over NVLink peer memory instead of through NCCL
 * (push to every peer, flag, sum in rank order; csrc/tp_allreduce.cuh).  advspec_tp_ipc_export
 * allocates this rank's exchange region and returns its 64-byte CUDA IPC handle; the host layer
 * gathers the tp_size handles in rank order and gives all of them to advspec_tp_ipc_import on every
 * rank (own entry ignored).  No rank may start a decode before every rank has imported.  Without
 * these two calls (or with ADVSPEC_TP_NCCL_ONLY set) every exchange goes through NCCL. */
advspec_status advspec_tp_ipc_export(advspec_engine *e, uint8_t *out64);
advspec_status advspec_tp_ipc_import(advspec_engine *e, const uint8_t *handles);

/* Copy a host weight blob (bf16 matrices, fp32 norm/bias vectors) to HBM. */
advspec_status advspec_load_weights(advspec_engine *e, const void *host_blob,
                                    size_t bytes);
/* Seeded N(0, std) weights generated on the device (norm weights = 1), for
 * full-size synthetic benchmarks where no checkpoint exists offline. */
advspec_status advspec_init_weights_random(advspec_engine *e, uint64_t seed,
                                           float std);
/* Optional: RoPE inverse frequencies computed by the caller exactly as the
 * oracle computes them (head_dim/2 floats).  Default: theta^(-2i/head_dim). */
advspec_status advspec_set_rope_inv_freq(advspec_engine *e,
                                         const float *inv_freq, int32_t n);

/* ---- the hot path -------------------------------------------------------- */

/* Prefill the prompt tokens ONCE and keep their KV as a shared prefix.
 * Replaces the prompt-processing half of every opponent's `completion` call
 * (models.py:628) for opponents with the same weights — the reference sends
 * the identical messages N times (SURVEY.md §8(b) "Threading"). */
advspec_status advspec_prefill(advspec_engine *e, const int32_t *tokens,
                               int32_t n_tokens, int32_t *prefix_id);

/* Continue a live prefix instead of starting over: keep the KV of its first keep_tokens tokens and
 * prefill `tokens[0..n_tokens)` at positions keep_tokens.. against it.  Serves the reference's prompt
 * variants that differ only AFTER the spec — `--context`, `--focus`, `--preserve-intent` sections follow
 * the document in the user message (prompts.py:233-241; models.py:130-146, 485-503) — so a second call on
 * the same round and document prefills only its tail.  n_tokens == 0 with keep_tokens == the prefix
 * length re-arms the prefix as is (identical prompt: retries, a panel larger than one batch).  Every
 * opponent forked from the old prefix is released; the old id dies; returns the new id. */
advspec_status advspec_prefill_extend(advspec_engine *e, int32_t prefix_id,
                                      int32_t keep_tokens, const int32_t *tokens,
                                      int32_t n_tokens, int32_t *new_prefix_id);

/* Fork n opponents over one prefix without copying KV.  seeds[i] drives the
 * sampler of opponent i (the reference gets per-call randomness from the
 * provider at temperature 0.7, models.py:626). */
advspec_status advspec_fork(advspec_engine *e, int32_t prefix_id,
                            int32_t n_seqs, const uint64_t *seeds,
                            int32_t *seq_ids);

/* Give ONE freshly forked opponent more prompt tokens of its own: `tokens[0..n_tokens)` continue the shared
 * prefix for this opponent only (a per-opponent persona or instruction behind the common document — the
 * reference has one `--persona` for the whole panel, debate.py:835).  The tail runs as a
 * prompt chunk (GEMM-shaped, one pass over the weights) at positions prefix_len.. against the prefix KV; its
 * K/V becomes the first n_tokens entries of the opponent's own KV and the next-token logits after it are the
 * ones `advspec_decode` samples this opponent's first token from.  Call it for every opponent of the coming
 * decode batch, in batch order, directly after the fork; the shared prefix and the other opponents are
 * untouched.  n_tokens counts against the opponent's max_new_tokens capacity; prefix_len + n_tokens must fit
 * max_prefix_tokens (the chunk's K/V is staged behind the prefix). */
advspec_status advspec_append_tail(advspec_engine *e, int32_t seq_id,
                                   const int32_t *tokens, int32_t n_tokens);

/* Batched autoregressive decode of the forked opponents: up to max_new tokens
 * each (`max_tokens`, models.py:620), sampling at `temperature` (0 = greedy),
 * stopping an opponent at eos_id (< 0: never).  out_tokens is [n][max_new]
 * row-major; out_lens[i] = tokens produced for opponent i (what the reference
 * reads back as usage.completion_tokens, models.py:640). */
advspec_status advspec_decode(advspec_engine *e, const int32_t *seq_ids,
                              int32_t n, int32_t max_new, float temperature,
                              int32_t eos_id, int32_t *out_tokens,
                              int32_t *out_lens);

/* One teacher-forced step: append forced_tokens[i] to opponent i and compute
 * the next-token logits (parity tests compare these with the HF oracle). */
advspec_status advspec_decode_step(advspec_engine *e, const int32_t *seq_ids,
                                   int32_t n, const int32_t *forced_tokens);

/* fp32 logits [n][vocab] of the last decode step (n = its batch), or of the
 * last prefill's final position (n = 1). */
advspec_status advspec_get_logits(advspec_engine *e, int32_t n, float *out);

/* Diagnostic: logits at EVERY prompt position ([n_tokens][vocab] fp32) through
 * the same prefill kernels.  Does not create a prefix. */
advspec_status advspec_prefill_logits(advspec_engine *e, const int32_t *tokens,
                                      int32_t n_tokens, float *out);

advspec_status advspec_release_seqs(advspec_engine *e, const int32_t *seq_ids,
                                    int32_t n);
advspec_status advspec_release_prefix(advspec_engine *e, int32_t prefix_id);

/* ---- multi-GPU replica placement ----------------------- */

/* Device address and size of a prefix's KV so rank 0 can NCCL-broadcast it to
 * replica GPUs instead of every GPU recomputing the prefill. */
advspec_status advspec_prefix_kv_region(advspec_engine *e, int32_t prefix_id,
                                        void **dev_ptr, size_t *bytes);
/* Receiver side: reserve a prefix of n_tokens whose KV will be filled by the
 * caller (broadcast) and whose last-position logits are `logits` (host fp32
 * [vocab]).  Returns its id. */
advspec_status advspec_prefix_adopt(advspec_engine *e, int32_t n_tokens,
                                    const float *logits, int32_t *prefix_id);

/* ---- measurement --------------------------------------------------------- */

advspec_status advspec_get_timing(advspec_engine *e, advspec_timing *out);
/* Run ONE extra decode step with per-kernel CUDA-event timing (not graph
 * replay) and fill gemv_ms/gemv_launches; state is rolled back afterwards. */
advspec_status advspec_profile_decode_step(advspec_engine *e,
                                           const int32_t *seq_ids, int32_t n);
/* Algorithmic HBM bytes of one decode step at the current lengths (SURVEY.md
 * §8(d): W_read + kvB*(S_p + sum t_i) + kvB*b), and of the GEMV launches only. */
advspec_status advspec_decode_step_bytes(advspec_engine *e,
                                         const int32_t *seq_ids, int32_t n,
                                         double *step_bytes,
                                         double *gemv_bytes);

/* In-situ kernel timeline of the decode path.  While enabled, block 0 of every decode
 * kernel stamps the GPU's global nanosecond timer on entry; consecutive stamps give each
 * kernel's real cost (run + launch gap) inside the CUDA-graph replay.  `advspec_ktrace_read`
 * copies up to cap entries ((ns << 4) | kind; kind 1 = weight-streaming GEMV, 2 = attention,
 * 3 = sampler, 4 = RoPE/append, 5 = combine) and resets the log. */
advspec_status advspec_ktrace_enable(advspec_engine *e, int32_t on);
advspec_status advspec_ktrace_read(advspec_engine *e, uint64_t *out, int32_t cap,
                                   int32_t *n);
/* Diagnostic: 16 phase stamps (ns) left by one CTA of the last instrumented kernel
 * launch while tracing was on (where inside the kernel the time goes). */
advspec_status advspec_ktrace_phases(advspec_engine *e, uint64_t *out16);

/* ---- op-level entry points (device pointers; used by tests/) ------------- */

/* C[M,N] = A[M,K] * B[N,K]^T on tcgen05; epilogue: 0 = bf16 store (+bias f32[N]
 * if aux != NULL), 1 = fp32 in-place residual add (C is f32), 2 = gated
 * activation over interleaved column pairs -> bf16 [M,N/2] (act as in desc),
 * 3 = fp32 store.  lda/ldb/ldc in elements. */
advspec_status advspec_op_gemm(int32_t device, const void *A, int64_t lda,
                               const void *B, int64_t ldb, void *C,
                               int64_t ldc, const void *aux, int32_t M,
                               int32_t N, int32_t K, int32_t epilogue,
                               int32_t act);
/* Same contract on plain CUDA cores (verification only, never on the path). */
advspec_status advspec_op_gemm_check(int32_t device, const void *A,
                                     int64_t lda, const void *B, int64_t ldb,
                                     void *C, int64_t ldc, const void *aux,
                                     int32_t M, int32_t N, int32_t K,
                                     int32_t epilogue, int32_t act);
/* y = epilogue(W[N,K] * x[b,K]) weight-streaming GEMV, b <= 8.
 * in_mode 0: x bf16 [b,K]; 1: x f32 [b,K] with fused RMSNorm (norm_w f32[K]).
 * epilogue as above but 0 writes bf16 [b,N], 1 adds into f32 y [b,N],
 * 2 writes bf16 [b,N/2], 3 writes f32 [b,N]. */
advspec_status advspec_op_gemv(int32_t device, const void *W, const void *x,
                               const void *norm_w, const void *bias, void *y,
                               int32_t b, int32_t N, int32_t K,
                               int32_t in_mode, int32_t epilogue, int32_t act,
                               float eps);
/* Causal attention for a prompt chunk.  q: bf16 [n_q][ldq] (head h at column
 * h*head_dim), K/V cache: bf16 [n_kv_heads][kv_stride][head_dim]; query i sits
 * at position q_pos0+i and sees keys 0..q_pos0+i.  out: bf16 [n_q][n_heads*
 * head_dim].  impl 0 = mma.sync kernel, 1 = scalar check kernel, 2 = tcgen05 kernel
 * (head_dim 64, 96 or 128). */
advspec_status advspec_op_attn_prefill(int32_t device, const void *q,
                                       int64_t ldq, const void *kcache,
                                       const void *vcache, int64_t kv_stride,
                                       void *out, int32_t n_q, int32_t q_pos0,
                                       int32_t n_heads, int32_t n_kv_heads,
                                       int32_t head_dim, int32_t impl);

/* One decode-attention step, the engine's kernels on caller-supplied device buffers (tests/ compare it
 * with an fp32 reference).  qkv: bf16 [b][(n_heads+2*n_kv_heads)*head_dim] raw projections of each
 * opponent's new token; rope_cos/rope_sin: f32 [>= max pos+1][head_dim/2]; prefix K/V: bf16
 * [n_kv_heads][prefix_stride][head_dim] of which prefix_len tokens are valid, shared by all opponents;
 * suffix K/V: bf16 [b][n_kv_heads][suffix_stride][head_dim], opponent i holding pos[i]-prefix_len tokens
 * already (RoPE applied); pos: HOST int32 [b], absolute position of each new token.  The kernel rotates
 * q and k, appends k/v at suffix row pos[i]-prefix_len (side effect, as in the engine) and writes
 * out: bf16 [b][n_heads*head_dim] = softmax(q k^T / sqrt(head_dim)) v over prefix + suffix + new token. */
advspec_status advspec_op_attn_decode(int32_t device, const void *qkv,
                                      const void *rope_cos, const void *rope_sin,
                                      const void *prefix_k, const void *prefix_v,
                                      int64_t prefix_stride, int32_t prefix_len,
                                      void *suffix_k, void *suffix_v,
                                      int64_t suffix_stride, const int32_t *pos,
                                      void *out, int32_t b, int32_t n_heads,
                                      int32_t n_kv_heads, int32_t head_dim);

#ifdef __cplusplus
}
#endif
#endif /* ADVSPEC_ENGINE_H */
