#!/usr/bin/env python3
"""bench.py — aggregate critique tokens/sec of one opponent-critique round on B200.

Metric (BASELINE.json): aggregate critique tokens/sec across the opponents of a
panel; prefill TFLOPS reported beside it.  Workload at N=1 is BASELINE.json
configs[1] on one GPU: 3 opponents x Llama-3-8B (seeded random-init weights,
synthetic tokenizer), 4,096-token PRD + prompt envelope, G new tokens each at
temperature 0.7 — one shared-prefix prefill, then batched decode (b=3).  One
"step" = one `call_models_parallel` round.  With --gpus N (torchrun, one process
per GPU) every rank hosts its own 3-opponent panel over the same document: weak
scaling, no data-path collective; times are the max over ranks.

  value  = sum of output tokens over all ranks / device time (CUDA events inside the
           engine around prefill and decode; token ids are the only input, 19 KB)
  e2e    = the same through the public host API `models.call_models_parallel` with
           host strings in and host strings out: tokenisation, H2D of token ids, D2H
           of generated ids and detokenisation are all inside the timed region
  roofline = the dominant kernel (weight-streaming GEMV): algorithmic bytes of one
           decode step's GEMV launches / their summed CUDA-event time
  cpu_baseline = the reference's fan-out restated over a CPU HF model (oracle/), on a
           bounded sample (2 of the 32 layers, few new tokens), scaled per layer

`--impl reference` times that CPU fan-out itself as the reference arm.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

_emit = lambda line: print(json.dumps(line), flush=True)
METRIC = "aggregate_critique_tokens_per_sec"
UNIT = "tokens/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--opponents", type=int, default=3)
    ap.add_argument("--spec-tokens", type=int, default=4096)
    ap.add_argument("--gen", type=int, default=256, help="new tokens per opponent (the CLI's 8000 cap, bounded)")
    ap.add_argument("--doc-type", default="prd")
    ap.add_argument("--cpu-sample-layers", type=int, default=1)
    ap.add_argument("--cpu-sample-gen", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args):
    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import envelope, model_spec
    from advspec_b200.tokenizer import SyntheticTokenizer, generate_spec, render_chat

    spec = model_spec.resolve(args.model)
    tok = SyntheticTokenizer(spec.vocab_size)
    doc = generate_spec(tok, args.spec_tokens, seed=2024, title="Synthetic PRD").strip()
    system_prompt, user_message = envelope.build_messages(doc, 1, args.doc_type)
    prompt_tokens = len(tok.encode(render_chat(system_prompt, user_message), bos=True))
    return spec, doc, system_prompt, user_message, prompt_tokens


def config_dict(args, spec, prompt_tokens, world):
    return {
        "workload": f"{args.opponents}-opponent replica panel, {spec.name} (random-init, synthetic tokenizer), "
                    f"{args.spec_tokens}-token {args.doc_type.upper()} (+envelope = {prompt_tokens} prompt tokens), "
                    f"{args.gen} new tokens per opponent, temperature 0.7, shared-prefix prefill + batched decode",
        "baseline_config": "configs[1] (3x Llama-3-8B, 4K PRD), opponents co-batched on each GPU",
        "opponents_per_gpu": args.opponents,
        "panels": world,
        "spec_tokens": args.spec_tokens,
        "prompt_tokens": prompt_tokens,
        "new_tokens_per_opponent": args.gen,
        "placement": "batch (same-weight opponents share one prefill and one weight stream per GPU)",
        "l2": "inputs larger than L2: 16 GB of weights are re-streamed every decode step (L2 is 126 MB)",
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the busiest half of the samples (idle gaps between steps excluded)
        busy = sm[len(sm) // 2:] if sm else []
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- peaks
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d.get("hbm_gbs"), "bf16_tflops": d.get("bf16_tflops"),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- CPU arm
def cpu_round(args, system_prompt, user_message, sample_layers, sample_gen):
    """One bounded sample of the reference's CPU fan-out; returns (estimated full-workload
    seconds, raw sample seconds, description).  The sample keeps the FULL prompt and layer
    shape but runs `sample_layers` of the model's layers and `sample_gen` new tokens; per-layer
    time (forward hooks) is scaled to all layers, per-token decode time to `args.gen` tokens."""
    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import model_spec
    from oracle import fanout_ref

    full = model_spec.resolve(args.model)
    layers = min(sample_layers, full.n_layers)
    name = f"{full.name}-{layers}layer"
    if name not in model_spec.REGISTRY:
        model_spec.REGISTRY[name] = model_spec.with_layers(full, layers, name)
    os.environ["ADVSPEC_MAX_NEW_TOKENS"] = str(sample_gen)
    os.environ["ADVSPEC_CPU_FAST_INIT"] = "1"
    fanout_ref.take_timings()
    t0 = time.perf_counter()
    out = fanout_ref.cpu_call_models_parallel([f"local/{name}"] * args.opponents, system_prompt, user_message)
    wall = time.perf_counter() - t0
    tm = fanout_ref.take_timings()
    scale = full.n_layers / layers
    n = len(tm)
    pre_wall = max(t["prefill_s"] for t in tm)
    dec_wall = max(t["decode_s"] for t in tm)
    pre_layer = sum(t["prefill_layer_s"] for t in tm) / n
    dec_layer = sum(t["decode_layer_s"] for t in tm) / n
    new_tok = sum(t["new_tokens"] for t in tm) / n
    est_prefill = pre_wall + (scale - 1.0) * pre_layer
    est_decode = (dec_wall + (scale - 1.0) * dec_layer) * (args.gen / max(new_tok, 1.0))
    desc = (f"{args.opponents} threads x HF CPU fp32 {name}: full {tm[0]['prompt_tokens']}-token prompt, "
            f"{int(new_tok)} new tokens, {layers} of {full.n_layers} layers; layer time x{scale:.0f}, "
            f"decode x{args.gen / max(new_tok, 1.0):.0f} to the full workload")
    return est_prefill + est_decode, wall, desc, sum(o[3] for o in out)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import torch

    spec, doc, system_prompt, user_message, prompt_tokens = workload(args)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(max(args.warmup, 0)):
        cpu_round(args, system_prompt, user_message, args.cpu_sample_layers, args.cpu_sample_gen)
    ests, walls, desc = [], [], ""
    for _ in range(args.steps):
        est, wall, desc, _ = cpu_round(args, system_prompt, user_message, args.cpu_sample_layers,
                                       args.cpu_sample_gen)
        ests.append(est)
        walls.append(wall)
    est = sum(ests) / len(ests)
    value = args.opponents * args.gen / est
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": est * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, spec, prompt_tokens, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "sample_wall_s": sum(walls) / len(walls)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference fan-out (models.py:681-722) restated over a CPU HF model; litellm and the "
                "reference tree do not exist on this box; whole-workload time is estimated from a bounded sample",
    }
    _emit(line)


# ----------------------------------------------------------------------------- B200 arm
def run_b200_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["ADVSPEC_DEVICES"] = str(local_rank)
    os.environ["ADVSPEC_MAX_NEW_TOKENS"] = str(args.gen)
    os.environ["ADVSPEC_PLACEMENT"] = "batch"

    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import models as amodels, runtime

    spec, doc, system_prompt, user_message, prompt_tokens = workload(args)
    names = [f"b200/{spec.name}"] * args.opponents

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_round(round_num):
        results = amodels.call_models_parallel(names, doc, round_num, args.doc_type)
        bad = [r.error for r in results if r.error]
        if bad:
            raise SystemExit(f"bench.py: opponent failed: {bad[0]}")
        return results

    def engine():
        return runtime.POOL.get(spec, local_rank, prompt_tokens, args.gen).engine

    # stderr noise from the reference-compatible warnings (random-init models never emit [SPEC])
    devnull = open(os.devnull, "w")
    real_stderr = sys.stderr
    sys.stderr = devnull
    try:
        for w in range(args.warmup):
            one_round(100 + w)
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        launches0 = engine().timing().kernel_launches
        dev_ms, out_tokens = 0.0, 0
        prefill_ms, decode_ms, decode_steps = [], [], []
        t0 = time.perf_counter()
        for k in range(args.steps):
            results = one_round(1 + k)
            tm = engine().timing()
            dev_ms += tm.prefill_ms + tm.decode_ms
            prefill_ms.append(tm.prefill_ms)
            decode_ms.append(tm.decode_ms)
            decode_steps.append(tm.decode_steps)
            out_tokens += sum(r.output_tokens for r in results)
        barrier()
        wall = time.perf_counter() - t0
        launches = engine().timing().kernel_launches - launches0
        clocks = sampler.stop()
    finally:
        sys.stderr = real_stderr

    # max over ranks of the time, sum over ranks of the tokens
    (dev_ms, wall), (out_tokens, launches) = runtime.reduce_round_stats([dev_ms, wall], [out_tokens, launches],
                                                                        device="cuda")
    out_tokens, launches = int(out_tokens), int(launches)

    # Extra, N > 1 only: the SAME 'world'-opponent panel spread one opponent per GPU (configs[1]'s literal
    # placement) with the prefix prefilled once on rank 0 and its KV broadcast over NVLink, against every
    # rank recomputing the prefill.  Reported beside the headline; SURVEY.md §8(e) asks for all three.
    spread = None
    if world > 1:
        prompt_ids = _prompt_ids(spec, system_prompt, user_message)
        e = engine()

        def spread_round(broadcast: bool):
            barrier()
            t0 = time.perf_counter()
            if broadcast:
                pid = runtime.replicate_prefix(e, prompt_ids, rank, src=0, device=local_rank)
            else:
                pid = e.prefill(prompt_ids)
            ids = e.fork(pid, [runtime.opponent_seed(7, rank)])
            res = e.decode(ids, args.gen, temperature=0.7)
            e.release_prefix(pid)
            barrier()
            return time.perf_counter() - t0, sum(res.lens)

        spread = {}
        for mode, bc in (("kv_broadcast", True), ("recompute", False)):
            spread_round(bc)  # warm (NCCL communicator, graph for b = 1)
            times, toks = zip(*[spread_round(bc) for _ in range(max(1, args.steps))])
            (tmax,), (tsum,) = runtime.reduce_round_stats([sum(times)], [sum(toks)], device="cuda")
            spread[mode] = {"tokens_per_s": tsum / tmax, "s_per_round": tmax / len(times), "opponents": world}

    line = None
    if rank == 0:
        peaks = measured_peaks()
        e = engine()
        # roofline of the dominant kernel, measured IN SITU: the engine stamps the GPU's global timer
        # at the start of every decode kernel inside the CUDA-graph replay; consecutive stamps give each
        # kernel's real cost (run + launch gap).  Host-side events cannot time 5-40 us kernels without
        # becoming CPU-bound, and ncu serialises them with cold caches.
        from advspec_b200 import measure

        pid = e.prefill(_prompt_ids(spec, system_prompt, user_message))
        ids = e.fork(pid, [1 + i for i in range(args.opponents)])
        e.decode(ids, max(2, args.gen // 2), temperature=0.7)
        step_bytes, gemv_bytes = e.decode_step_bytes(ids)
        e.ktrace_enable(True)
        e.decode(ids, min(40, max(4, args.gen // 4)), temperature=0.7)
        tl = measure.summarize(e.ktrace_read(), spec.n_layers)
        e.ktrace_enable(False)
        gemv_ms = tl.get("gemv_us_per_step", 0.0) / 1e3
        gemv_gbs = gemv_bytes / (gemv_ms * 1e-3) / 1e9 if gemv_ms > 0 else 0.0
        e.release_prefix(pid)
        mean_prefill = sum(prefill_ms) / len(prefill_ms)
        mean_decode = sum(decode_ms) / len(decode_ms)
        mean_steps = sum(decode_steps) / len(decode_steps)
        step_ms = mean_decode / max(mean_steps, 1)
        decode_gbs = step_bytes / (step_ms * 1e-3) / 1e9
        prefill_tflops = spec.prefill_flops(prompt_tokens) / (mean_prefill * 1e-3) / 1e12
        value = out_tokens / (dev_ms * 1e-3)
        e2e_value = out_tokens / wall
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args, spec, prompt_tokens, world),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": prompt_tokens * 4,
                    "d2h_bytes_per_step": args.opponents * args.gen * 4 + args.opponents * 12,
                    "api": "advspec_b200.models.call_models_parallel (seam B2), host strings in/out",
                    "wall_s": wall},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "gemv_mma_kernel (weight-streaming GEMV: every decode matmul)",
                         "achieved": gemv_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gemv_gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
                         "algorithmic_bytes_per_step": gemv_bytes,
                         "launches_per_step": tl.get("gemv_launches_per_step"),
                         "kernel_ms_per_step": gemv_ms, "traffic": _ncu_traffic_per_step(spec),
                         "how": "device global-timer stamps at kernel entry inside the graph replay; cost = run + launch gap",
                         "timeline": {k: round(v["us_each"], 2) for k, v in tl.get("rows", {}).items()}},
            "decode": {"ms_per_step": step_ms, "batch": args.opponents, "algorithmic_bytes_per_step": step_bytes,
                       "achieved_gbs": decode_gbs, "frac_of_hbm_peak": decode_gbs / peaks["hbm_gbs"],
                       "tokens_per_s_per_gpu": args.opponents / (step_ms * 1e-3)},
            "prefill": {"ms": mean_prefill, "tokens": prompt_tokens, "algorithmic_tflop": spec.prefill_flops(prompt_tokens) / 1e12,
                        "tflops": prefill_tflops, "frac_of_bf16_peak": prefill_tflops / peaks["bf16_tflops"],
                        "shared_by_opponents": args.opponents},
        }
        if spread is not None:
            line["replica_spread"] = dict(spread, note="one opponent per GPU, one panel of n_gpus opponents; wall "
                                          "time incl. host calls; headline `value` is the co-batched weak-scaling run")
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            try:
                est, wall_cpu, desc, _ = cpu_round(args, system_prompt, user_message, args.cpu_sample_layers,
                                                   args.cpu_sample_gen)
                line["cpu_baseline"] = {"value": args.opponents * args.gen / est, "unit": UNIT, "cores": cores,
                                        "kind": "port", "sample": desc, "sample_wall_s": wall_cpu}
            except Exception as ex:  # the baseline must never sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"failed: {ex}"}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def _ncu_traffic_per_step(spec):
    """DRAM bytes per decode step of the GEMV launches, from the committed ncu --set full capture
    (profiles/r01_gemv_traffic.json: read + write per launch for the four per-layer shapes); lm_head is
    taken at its algorithmic size.  None when the capture does not match the model."""
    p = ROOT / "profiles" / "r01_gemv_traffic.json"
    if not p.exists() or spec.name != "llama-3-8b":
        return None
    d = json.loads(p.read_text())
    return (d["per_layer_traffic_mb"] * spec.n_layers + spec.vocab_size * spec.d_model * 2 / 1e6) * 1e6


def _prompt_ids(spec, system_prompt, user_message):
    from advspec_b200.tokenizer import SyntheticTokenizer, render_chat

    return SyntheticTokenizer(spec.vocab_size).encode(render_chat(system_prompt, user_message), bos=True)


def main():
    # Libraries (NCCL's version banner, HF warnings) must not share stdout with the ONE JSON line the
    # driver parses: everything written to fd 1 goes to stderr, the JSON goes to the saved real stdout.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _emit
    _emit = lambda line: (real_stdout.write(json.dumps(line) + "\n"), real_stdout.flush())
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(args.gpus, 1) and world > 1:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
