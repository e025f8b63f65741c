/* critique_panel.c — the engine through its C ABI alone (no Python, no torch): one shared-prefix
 * prefill, N opponents forked over it, batched sampling decode.  This is the sequence a native host
 * (or a cgo / JNI / N-API binding) makes where the reference makes N `completion` calls
 * (skills/adversarial-spec/scripts/models.py:681-722 -> :628).
 *
 * build: gcc -O2 -Iinclude examples/critique_panel.c -Ladversarial-spec_b200 -ladvspec_b200 \
 *            -Wl,-rpath,$PWD/adversarial-spec_b200 -Wl,--allow-shlib-undefined -o critique_panel
 * run  : ./critique_panel [n_opponents=3] [prompt_tokens=1024] [new_tokens=64]      (needs a B200) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "advspec_engine.h"

static void die(advspec_engine *e, const char *what, advspec_status st) {
  fprintf(stderr, "%s failed (status %d): %s\n", what, (int)st, advspec_last_error(e));
  exit(1);
}

int main(int argc, char **argv) {
  const int n_opp = argc > 1 ? atoi(argv[1]) : 3;
  const int n_prompt = argc > 2 ? atoi(argv[2]) : 1024;
  const int n_new = argc > 3 ? atoi(argv[3]) : 64;
  if (n_opp < 1 || n_opp > 8 || n_prompt < 1 || n_new < 1) {
    fprintf(stderr, "usage: %s [opponents 1..8] [prompt tokens] [new tokens]\n", argv[0]);
    return 2;
  }
  /* a 2-layer Llama-3-8B-shaped model: full-size matrices, quick to initialise */
  advspec_model_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = ADVSPEC_ABI_VERSION;
  d.n_layers = 2;
  d.d_model = 4096;
  d.n_heads = 32;
  d.n_kv_heads = 8;
  d.head_dim = 128;
  d.d_ff = 14336;
  d.vocab_size = 128256;
  d.act = 0;
  d.max_prefix_tokens = (n_prompt + 255) / 256 * 256;
  d.max_new_tokens = n_new + 16;
  d.max_seqs = 8;
  d.tp_rank = 0;
  d.tp_size = 1;
  d.rope_theta = 500000.0f;
  d.norm_eps = 1e-5f;
  d.embed_scale = 1.0f;

  advspec_engine *e = NULL;
  advspec_status st = advspec_engine_create(&d, 0, &e);
  if (st != ADVSPEC_OK) die(NULL, "advspec_engine_create", st);
  if ((st = advspec_init_weights_random(e, 0, 0.02f)) != ADVSPEC_OK) die(e, "advspec_init_weights_random", st);

  int32_t *prompt = malloc(sizeof(int32_t) * (size_t)n_prompt);
  uint64_t x = 88172645463325252ull;
  for (int i = 0; i < n_prompt; ++i) { /* xorshift token ids: a stand-in for the tokenised spec */
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    prompt[i] = (int32_t)(x % (uint64_t)d.vocab_size);
  }
  int32_t prefix = 0;
  if ((st = advspec_prefill(e, prompt, n_prompt, &prefix)) != ADVSPEC_OK) die(e, "advspec_prefill", st);

  uint64_t seeds[8];
  int32_t ids[8];
  for (int i = 0; i < n_opp; ++i) seeds[i] = 1000u + (uint64_t)i;
  if ((st = advspec_fork(e, prefix, n_opp, seeds, ids)) != ADVSPEC_OK) die(e, "advspec_fork", st);

  int32_t *out = malloc(sizeof(int32_t) * (size_t)n_opp * (size_t)n_new);
  int32_t lens[8];
  if ((st = advspec_decode(e, ids, n_opp, n_new, 0.7f, -1, out, lens)) != ADVSPEC_OK) die(e, "advspec_decode", st);

  advspec_timing tm;
  advspec_get_timing(e, &tm);
  for (int i = 0; i < n_opp; ++i) {
    printf("opponent %d: %d tokens:", i, lens[i]);
    for (int t = 0; t < lens[i] && t < 8; ++t) printf(" %d", out[(size_t)i * n_new + t]);
    printf(" ...\n");
  }
  printf("prefill %.2f ms (once, shared by %d opponents); decode %.3f ms/step x %d steps, batch %d\n",
         tm.prefill_ms, n_opp, tm.decode_ms / (tm.decode_steps > 0 ? tm.decode_steps : 1), tm.decode_steps,
         tm.decode_batch);
  advspec_engine_destroy(e);
  free(out);
  free(prompt);
  return 0;
}
