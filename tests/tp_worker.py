"""One rank of the tensor-parallel GPU check (launched by tests/test_gpu_tp.py under torchrun, one
process per GPU).  Parity: a small HF-initialised model split `world` ways must give the whole-model
HF CPU logits (prefill and teacher-forced decode) and — sampled with the same seeds — the tokens the
unsplit engine gives.  Optional throughput leg on a full-size shape with random weights (TP_PERF=model)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

import advspec_loader

advspec_loader.load()
from advspec_b200 import engine as eng, model_spec, runtime, weights  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    out = {"world": world}
    for name in [m for m in os.environ.get("TP_MODELS", "tiny-gqa4,tiny-llama-128").split(",") if m]:
        from oracle import hf_oracle  # the checker (test infrastructure)
        spec = model_spec.resolve(name)
        big = spec.n_params() > 200_000_000  # full-width shapes: one-pass init instead of HF's per-module init
        model = hf_oracle.build_hf_model_fast(spec, 11) if big else hf_oracle.build_hf_model(spec, 11)
        full = hf_oracle.export_blob(spec, model)
        rng = np.random.default_rng(5)
        prompt = rng.integers(0, spec.vocab_size, 200).tolist()
        forced = rng.integers(0, spec.vocab_size, 6).tolist()
        e = runtime.create_tp_engine(spec, dev, 256, 64, 4, rank, world)
        e.set_rope_inv_freq(hf_oracle.rope_inv_freq(model))
        e.load_weights(weights.shard_blob(full, spec, rank, world))

        def gather(shard):  # [n, V/world] per rank -> [n, V]
            parts = [torch.zeros(shard.shape) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(shard)))
            return torch.cat(parts, dim=1).numpy()

        lg = gather(e.prefill_logits(prompt))
        pid = e.prefill(prompt)
        ids = e.fork(pid, [1, 2])
        dec = []
        for t in forced:
            e.decode_step(ids, [t, t])
            dec.append(gather(e.get_logits(2)))
        pid = e.prefill(prompt)  # a fresh prefix: decode after teacher forcing must keep the same opponents
        ids = e.fork(pid, [7, 8, 9])
        sampled = e.decode(ids, 24, temperature=0.7).tokens
        pid = e.prefill(prompt)
        ids = e.fork(pid, [7])
        greedy = e.decode(ids, 24, temperature=0.0).tokens
        # per-opponent prompt tails (advspec_append_tail): the prompt's first 150 tokens are shared, opponent 0
        # continues with tokens 150..199, opponent 1 with 150..179 — the logits after each tail are those of
        # the same sequence at positions 199 / 179
        pid = e.prefill(prompt[:150])
        ids = e.fork(pid, [3, 4])
        e.append_tail(ids[0], prompt[150:200])
        e.append_tail(ids[1], prompt[150:180])
        tails = gather(e.get_logits(2))
        tail_tokens = e.decode(ids, 8, temperature=0.0).tokens
        e.close()
        res = {}
        if rank == 0:
            want = hf_oracle.hf_logits(model, prompt + forced)
            std = float(want.std())
            res["prefill_max_over_std"] = float(np.abs(lg - want[:200]).max() / std)
            res["prefill_rms_over_std"] = float(np.sqrt(((lg - want[:200]) ** 2).mean()) / std)
            dmax = 0.0
            for i in range(len(forced)):
                for b in range(2):
                    dmax = max(dmax, float(np.abs(dec[i][b] - want[200 + i]).max() / std))
            res["decode_max_over_std"] = dmax
            res["tails_max_over_std"] = float(max(np.abs(tails[0] - want[199]).max(),
                                                  np.abs(tails[1] - want[179]).max()) / std)
            res["tail_first_tokens_are_argmax"] = bool(tail_tokens[0][0] == int(tails[0].argmax()) and
                                                       tail_tokens[1][0] == int(tails[1].argmax()))
            # the unsplit engine on this GPU, same seeds
            e1 = eng.Engine(spec, dev, 256, 64, 4)
            e1.set_rope_inv_freq(hf_oracle.rope_inv_freq(model))
            e1.load_weights(full)
            pid1 = e1.prefill(prompt)
            s1 = e1.decode(e1.fork(pid1, [7, 8, 9]), 24, temperature=0.7).tokens
            pid1 = e1.prefill(prompt)
            g1 = e1.decode(e1.fork(pid1, [7]), 24, temperature=0.0).tokens
            e1.close()
            tot = sum(len(x) for x in s1)
            res["sampled_token_agreement"] = sum(a == b for x, y in zip(s1, sampled) for a, b in zip(x, y)) / tot
            res["greedy_token_agreement"] = sum(a == b for a, b in zip(g1[0], greedy[0])) / len(g1[0])
        # every rank must have produced the same tokens
        mine = torch.tensor([t for row in sampled + tail_tokens for t in row] + greedy[0], dtype=torch.int64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        res["ranks_identical"] = bool(all(torch.equal(allr[0], a) for a in allr))
        out[name] = res
    perf = os.environ.get("TP_PERF")
    if perf:
        spec = model_spec.resolve(perf)
        ptok, gen = int(os.environ.get("TP_PROMPT", "5068")), int(os.environ.get("TP_GEN", "64"))
        b = int(os.environ.get("TP_B", "1"))
        e = runtime.create_tp_engine(spec, dev, (ptok + 255) // 256 * 256, gen + 16, 8, rank, world)
        e.init_weights_random(0, 0.02)
        prompt = np.random.default_rng(0).integers(0, spec.vocab_size, ptok).tolist()
        best = None
        for rep in range(int(os.environ.get("TP_REPS", "3"))):
            pid = e.prefill(prompt)
            ids = e.fork(pid, list(range(1, b + 1)))
            r = e.decode(ids, gen, temperature=0.7)
            tm = e.timing()
            cur = {"prefill_ms": tm.prefill_ms, "decode_ms_per_step": tm.decode_ms / max(tm.decode_steps, 1)}
            if rep > 0 and (best is None or cur["decode_ms_per_step"] < best["decode_ms_per_step"]):
                best = cur
        step_bytes, gemv_bytes = e.decode_step_bytes(ids)
        e.close()
        t = torch.tensor([best["prefill_ms"], best["decode_ms_per_step"]], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["perf"] = {"model": perf, "tp": world, "prompt": ptok, "b": b, "new_tokens": gen,
                       "prefill_ms": float(t[0]), "decode_ms_per_step": float(t[1]),
                       "tokens_per_s_decode": b / (float(t[1]) * 1e-3),
                       "bytes_per_step_per_gpu": step_bytes,
                       "hbm_gbs_per_gpu": step_bytes / (float(t[1]) * 1e-3) / 1e9,
                       "prefill_tflops": spec.prefill_flops(ptok) / (float(t[0]) * 1e-3) / 1e12}
    if rank == 0:
        print("TP_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
