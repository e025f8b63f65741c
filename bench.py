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
           decode step's GEMV launches / their in-situ cost; `traffic` = DRAM bytes of the
           same launches from an ncu pass run by this script (tools/traffic_probe.py)
  cpu_baseline = the reference's fan-out restated over a CPU HF model (oracle/), on a
           bounded sample (2 of the 32 layers, 16 new tokens, median of 3), scaled per layer

Beside the headline the same JSON line carries the other BASELINE configs, measured on the
GPUs of this run:
  strong   one 8-opponent Llama-3-8B panel spread over the N GPUs (strong scaling)
  hetero   config 3: the five-model panel, 8K spec, models round-robin over the N GPUs
           (N >= 5: one model per GPU), no collective
  converge config 4: 4 opponents x 6 rounds on a 16K tech spec through the resident CLI
           driver (`debate.py critique --session / --resume`), co-batched on one GPU; N >= 4
           adds the one-opponent-per-GPU placement
  tp       config 5 (N >= 2): one Llama-3-70B opponent tensor-parallel over the N GPUs,
           32K-token spec: decode HBM fraction per GPU, prefill TF/s per GPU
  personas SURVEY.md §8(f4): the headline panel with a different persona per opponent
           (`b200/llama-3-8b@<persona>`) against one persona for everyone, per rank
Each section is optional (`--sections`) and guarded: a failure or a time-out there is
reported in its own key and never costs the headline.

`--impl reference` times the CPU fan-out itself as the reference arm, on the same config.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
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
ALL_SECTIONS = ("strong", "personas", "hetero", "converge", "tp")
HETERO_PANEL = ["llama-3-8b", "mistral-7b", "qwen2-7b", "phi-3-mini", "gemma-7b"]


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
    ap.add_argument("--cpu-sample-layers", type=int, default=2)
    ap.add_argument("--cpu-sample-gen", type=int, default=16)
    ap.add_argument("--cpu-samples", type=int, default=3)
    ap.add_argument("--cpu-budget-s", type=float, default=240.0,
                    help="reference arm: wall-time budget for all (warmup + steps) samples")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sections", default="all",
                    help="comma list of extra sections (strong,personas,hetero,converge,tp), 'all' or 'none'")
    ap.add_argument("--sections-budget-s", type=float, default=420.0)
    ap.add_argument("--traffic", default="auto", choices=["auto", "off"],
                    help="auto: measure the GEMV's DRAM traffic with an ncu pass (N=1, when ncu is on PATH)")
    ap.add_argument("--extras-gen", type=int, default=64, help="new tokens per opponent in converge / tp")
    return ap.parse_args()


def sections_of(args) -> tuple:
    if args.sections == "all":
        return ALL_SECTIONS
    if args.sections in ("none", ""):
        return ()
    return tuple(s for s in args.sections.split(",") if s in ALL_SECTIONS)


def make_doc(vocab_size, n_tokens, seed, title):
    from advspec_b200.tokenizer import SyntheticTokenizer, generate_spec

    return generate_spec(SyntheticTokenizer(vocab_size), n_tokens, seed=seed, title=title).strip()


def workload(args):
    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import envelope, model_spec
    from advspec_b200.tokenizer import SyntheticTokenizer, render_chat

    spec = model_spec.resolve(args.model)
    tok = SyntheticTokenizer(spec.vocab_size)
    doc = make_doc(spec.vocab_size, args.spec_tokens, 2024, "Synthetic PRD")
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
        "prefix_cache": "no reuse between steps: consecutive rounds diverge at the round number (token ~400 of "
                        "5,068), below the cache's 50 % threshold, so every step prefills the whole prompt",
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
def cpu_sample(args, system_prompt, user_message, sample_layers, sample_gen, prompt_frac=1.0):
    """One bounded sample of the reference's CPU fan-out; returns (estimated full-workload seconds, raw
    sample seconds, description).  The sample keeps the layer shape but runs `sample_layers` of the model's
    layers and `sample_gen` new tokens; per-layer time (forward hooks) is scaled to all layers, per-token decode
    time to `args.gen` tokens.  `prompt_frac` < 1 (reference arm only, when even one layer of the full prompt
    does not fit a step's share of the budget) prefills a leading slice of the prompt and scales the prefill by
    the per-layer flop ratio of SURVEY.md §8(d) (linear GEMM term + quadratic attention term)."""
    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import model_spec
    from oracle import fanout_ref

    full = model_spec.resolve(args.model)
    layers = max(1, min(sample_layers, full.n_layers))
    name = f"{full.name}-{layers}layer"
    if name not in model_spec.REGISTRY:
        model_spec.REGISTRY[name] = model_spec.with_layers(full, layers, name)
    saved = os.environ.get("ADVSPEC_MAX_NEW_TOKENS")
    os.environ["ADVSPEC_MAX_NEW_TOKENS"] = str(sample_gen)
    os.environ["ADVSPEC_CPU_FAST_INIT"] = "1"
    user = user_message
    if prompt_frac < 1.0:
        user = user_message[: max(64, int(len(user_message) * prompt_frac))]
    try:
        fanout_ref.take_timings()
        t0 = time.perf_counter()
        fanout_ref.cpu_call_models_parallel([f"local/{name}"] * args.opponents, system_prompt, user)
        wall = time.perf_counter() - t0
        tm = fanout_ref.take_timings()
    finally:
        if saved is None:
            os.environ.pop("ADVSPEC_MAX_NEW_TOKENS", None)
        else:
            os.environ["ADVSPEC_MAX_NEW_TOKENS"] = saved
    scale = full.n_layers / layers
    n = len(tm)
    pre_wall = max(t["prefill_s"] for t in tm)
    dec_wall = max(t["decode_s"] for t in tm)
    pre_layer = sum(t["prefill_layer_s"] for t in tm) / n
    dec_layer = sum(t["decode_layer_s"] for t in tm) / n
    new_tok = sum(t["new_tokens"] for t in tm) / n
    est_prefill = pre_wall + (scale - 1.0) * pre_layer
    sliced = ""
    if prompt_frac < 1.0:
        p_s = tm[0]["prompt_tokens"]
        p_full = len(_prompt_ids(full, system_prompt, user_message))
        one = model_spec.with_layers(full, 1, "flops-1layer")
        per_layer = lambda t: one.prefill_flops(t) - 2.0 * one.vocab_size * one.d_model
        ratio = per_layer(p_full) / per_layer(p_s)
        est_prefill *= ratio
        sliced = f"; prompt slice {p_s} of {p_full} tokens, prefill x{ratio:.2f} by per-layer flops"
    # n new tokens cost n-1 decode forwards (the first comes from the prefill's logits)
    fwd = max(new_tok - 1.0, 1.0)
    est_decode = (dec_wall + (scale - 1.0) * dec_layer) * ((args.gen - 1) / fwd)
    desc = (f"{args.opponents} threads x HF CPU fp32 {name}: {tm[0]['prompt_tokens']}-token prompt, "
            f"{int(new_tok)} new tokens ({int(fwd)} decode forwards), {layers} of {full.n_layers} layers; layer time "
            f"x{scale:.0f}, decode forwards x{(args.gen - 1) / fwd:.0f} to the full workload{sliced}")
    return est_prefill + est_decode, wall, desc


def cpu_baseline(args, system_prompt, user_message, cores):
    """cpu_baseline of the B200 arm: `--cpu-samples` samples (default 3) of 2 layers x 16 new tokens, the
    median estimate and the spread; samples stop early once 90 s are spent (at least one runs)."""
    ests, walls, desc = [], [], ""
    t0 = time.perf_counter()
    for i in range(max(1, args.cpu_samples)):
        est, wall, desc = cpu_sample(args, system_prompt, user_message, args.cpu_sample_layers, args.cpu_sample_gen)
        ests.append(est)
        walls.append(wall)
        if time.perf_counter() - t0 > 90.0:
            break
    vals = sorted(args.opponents * args.gen / e for e in ests)
    return {"value": statistics.median(vals), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": desc + f"; median of {len(vals)} samples", "samples": len(vals),
            "spread": {"min": vals[0], "max": vals[-1]}, "sample_wall_s": sum(walls) / len(walls)}


def run_reference_arm(args, rank, world):
    """The CPU fan-out on this box's host cores, on the B200 arm's config.  A 'step' is one bounded sample
    sized so that warmup + steps samples fit `--cpu-budget-s`: the first (untimed) pass calibrates, then the
    layer count / new-token count of the sample shrink until a step fits its share of the budget.  N panels
    on the same host cores run back to back — tokens and time both scale by N — so the CPU value is the
    one-panel value at every N; the config (panels = N) is the B200 arm's."""
    if rank != 0:
        return
    import torch

    spec, doc, system_prompt, user_message, prompt_tokens = workload(args)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_panels = world if world > 1 else max(args.gpus, 1)
    n_samples = max(1, args.warmup) + args.steps
    share = args.cpu_budget_s / n_samples
    # calibration: the cheapest full-prompt sample first (1 layer, 2 new tokens; builds the model), then the
    # largest of {1, 2 layers} x {2, 16 tokens} whose wall time fits a step's share of the budget; if even the
    # cheapest does not fit, a leading slice of the prompt
    t0 = time.perf_counter()
    est, wall, desc = cpu_sample(args, system_prompt, user_message, 1, 2)
    first_wall = time.perf_counter() - t0
    est, wall, desc = cpu_sample(args, system_prompt, user_message, 1, 2)
    frac = 1.0
    layers = min(args.cpu_sample_layers, 2) if 2.2 * wall <= share else 1
    gen = args.cpu_sample_gen if 1.3 * wall * layers <= share else 2
    if wall > share:
        frac = max(0.1, 0.8 * share / wall)
        est, wall, desc = cpu_sample(args, system_prompt, user_message, 1, 2, frac)
    elif (layers, gen) != (1, 2):
        est, wall, desc = cpu_sample(args, system_prompt, user_message, layers, gen)  # (builds the 2-layer model)
    for _ in range(max(0, args.warmup - 2)):
        cpu_sample(args, system_prompt, user_message, layers, gen, frac)
    ests, walls = [], []
    for _ in range(args.steps):
        est, wall, desc = cpu_sample(args, system_prompt, user_message, layers, gen, frac)
        ests.append(est)
        walls.append(wall)
    vals = sorted(args.opponents * args.gen / e for e in ests)
    value = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_panels,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": args.opponents * args.gen / value * 1e3,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, spec, prompt_tokens, n_panels),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": desc + f"; median of {len(vals)} steps; {n_panels} panel(s) on the same host "
                                          f"cores run back to back, so tokens/s does not depend on the panel count",
                         "spread": {"min": vals[0], "max": vals[-1]}, "sample_wall_s": sum(walls) / len(walls),
                         "first_sample_wall_s_incl_model_build": first_wall},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference fan-out (models.py:681-722) restated over a CPU HF model; litellm and the "
                "reference tree do not exist on this box; whole-workload time is estimated from a bounded sample",
    }
    _emit(line)


# ----------------------------------------------------------------------------- helpers of the B200 arm
class Quiet:
    """stderr noise from the reference-compatible warnings (random-init models never emit [SPEC])."""

    def __enter__(self):
        self.real, self.null = sys.stderr, open(os.devnull, "w")
        sys.stderr = self.null

    def __exit__(self, *exc):
        sys.stderr = self.real
        self.null.close()
        return False


class Dist:
    """torch.distributed plumbing of the benchmark: barrier, max/sum reduction, object gather (gloo)."""

    def __init__(self, rank, world, local_rank):
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.gloo = None
        if world > 1:
            import torch
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.gloo = dist.new_group(backend="gloo")

    def barrier(self):
        import torch

        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
            torch.cuda.synchronize()

    def gather(self, obj) -> list:
        if self.world == 1:
            return [obj]
        import torch.distributed as dist

        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.gloo)
        return out

    def close(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()


def engine_stats(spec, device, peaks):
    """prefill ms / decode ms per step / HBM fraction of the resident engine's LAST round."""
    from advspec_b200 import runtime

    r = runtime.POOL._engines.get((spec.name, device))
    if r is None:
        return {}
    tm = r.engine.timing()
    steps = max(tm.decode_steps, 1)
    return {"prefill_ms": tm.prefill_ms, "decode_ms_per_step": tm.decode_ms / steps, "decode_batch": tm.decode_batch}


# ----------------------------------------------------------------------------- extra sections
def section_strong(args, D, peaks):
    """One 8-opponent Llama-3-8B panel over the N GPUs (strong scaling): each rank hosts 8/N opponents
    behind its own prefill of the same 4K PRD; no collective.  Decode streams the weights once per step
    whatever the batch, so this curve is flat by construction of the workload."""
    from advspec_b200 import models as amodels, model_spec, runtime

    spec = model_spec.resolve("llama-3-8b")
    doc = make_doc(spec.vocab_size, 4096, 2024, "Synthetic PRD")
    mine = runtime.shard_panels(8, D.rank, D.world)
    names = ["b200/llama-3-8b"] * len(mine)
    os.environ["ADVSPEC_SEED"] = str(1000 + D.rank)
    walls, toks = [], 0
    for rep in range(3):
        D.barrier()
        t0 = time.perf_counter()
        res = amodels.call_models_parallel(names, doc, 50 + rep, "prd") if names else []
        D.barrier()
        if rep > 0:
            walls.append(time.perf_counter() - t0)
            toks += sum(r.output_tokens for r in res)
    st = engine_stats(spec, D.local_rank, peaks) if names else {}
    rows = D.gather({"rank": D.rank, "opponents": len(mine), "tokens": toks, "wall_s": sum(walls), **st})
    total = sum(r["tokens"] for r in rows)
    wall = max(r["wall_s"] for r in rows)
    return {"workload": "one 8-opponent Llama-3-8B panel, 4,096-token PRD, %d new tokens each" % args.gen,
            "scaling": "strong", "opponents": 8, "tokens_per_s": total / wall if wall > 0 else None,
            "s_per_round": wall / 2, "per_rank": rows}


def section_personas(args, D, peaks):
    """the headline panel (3 x Llama-3-8B, 4,096-token PRD) with a persona per opponent
    (`b200/llama-3-8b@<persona>`: persona text behind the document, ONE prefill of the shared tokens, the
    per-opponent tails stepped through the batched decode path, one decode batch) beside the reference's
    layout (one `--persona` for the whole panel, in the system prompt).  Every rank runs its own panel."""
    from advspec_b200 import models as amodels, model_spec, runtime

    spec = model_spec.resolve("llama-3-8b")
    doc = make_doc(spec.vocab_size, 4096, 2024, "Synthetic PRD")
    os.environ["ADVSPEC_SEED"] = str(3000 + D.rank)
    people = ("security-engineer", "oncall-engineer", "junior-developer")
    per = [f"b200/llama-3-8b@{p}" for p in people]
    cases = (("one_persona_for_the_panel", ["b200/llama-3-8b"] * 3, people[0], None),
             ("persona_per_opponent", per, None, "chunk"),          # advspec_append_tail: one prompt chunk per tail
             ("persona_per_opponent_stepped", per, None, "step"))  # A/B: tails through the decode step
    row = {"rank": D.rank}
    for label, panel, persona, impl in cases:
        wall, res, before = 0.0, [], {}
        os.environ["ADVSPEC_TAIL_IMPL"] = impl or "chunk"
        for rep in range(2):  # the first round may re-create the engine (larger suffix KV) and captures the graph
            before = dict(runtime.PREFIXES.stats)
            D.barrier()
            t0 = time.perf_counter()
            res = amodels.call_models_parallel(panel, doc, 70 + rep, "prd", False, None, persona)
            wall = time.perf_counter() - t0
        st = engine_stats(spec, D.local_rank, peaks)
        after = runtime.PREFIXES.stats
        prefilled = after["tokens_prefilled"] - before.get("tokens_prefilled", 0)
        toks = sum(r.output_tokens for r in res)
        row[label] = {"wall_s": wall, "tokens": toks, "tokens_per_s": toks / wall if wall > 0 else None,
                      "input_tokens": sorted(r.input_tokens for r in res), "tokens_prefilled": prefilled,
                      "longest_tail": max(r.input_tokens for r in res) - prefilled if res else None,
                      "prefills": after["full"] - before.get("full", 0), "errors": [r.error for r in res if r.error],
                      **st}
    os.environ.pop("ADVSPEC_TAIL_IMPL", None)
    rows = D.gather(row)
    out = {"workload": "3 x Llama-3-8B, 4,096-token PRD, %d new tokens each; persona per opponent vs one for the "
                       "panel; one panel per rank" % args.gen, "per_rank": rows}
    for label, _, _, _ in cases:
        wall = max(r[label]["wall_s"] for r in rows)
        out[label + "_tokens_per_s"] = sum(r[label]["tokens"] for r in rows) / wall if wall > 0 else None
    return out


def section_hetero(args, D, peaks):
    """Config 3: Llama-3-8B / Mistral-7B / Qwen2-7B / Phi-3-mini / Gemma-7B, 8,192-token tech spec, models
    round-robin over the N GPUs (N >= 5: one model per GPU), each with its own prefill, no collective.  A
    rank hosting several models runs them one after another (clean per-model numbers)."""
    from advspec_b200 import models as amodels, model_spec, runtime

    mine = runtime.shard_panels(len(HETERO_PANEL), D.rank, D.world)
    doc = make_doc(32000, 8192, 11, "Synthetic Tech Spec")
    rows = []
    D.barrier()
    for i in mine:
        spec = model_spec.resolve(HETERO_PANEL[i])
        name = f"b200/{spec.name}"
        amodels.call_models_parallel([name], doc, 60, "tech")  # engine creation + weights + graph capture
        t0 = time.perf_counter()
        res = amodels.call_models_parallel([name], doc, 61, "tech")
        wall = time.perf_counter() - t0
        st = engine_stats(spec, D.local_rank, peaks)
        row = {"model": spec.name, "gpu": D.rank, "input_tokens": res[0].input_tokens,
               "output_tokens": res[0].output_tokens, "error": res[0].error, "wall_s": wall, **st}
        if st.get("decode_ms_per_step"):
            step_bytes = spec.decode_step_bytes(res[0].input_tokens, [res[0].output_tokens // 2])
            row["decode_hbm_frac"] = step_bytes / (st["decode_ms_per_step"] * 1e-3) / 1e9 / peaks["hbm_gbs"]
            row["prefill_tflops"] = spec.prefill_flops(res[0].input_tokens) / (st["prefill_ms"] * 1e-3) / 1e12
        rows.append(row)
    D.barrier()
    allrows = [r for part in D.gather(rows) for r in part]
    per_rank = {}
    for r in allrows:
        per_rank[r["gpu"]] = per_rank.get(r["gpu"], 0.0) + r["wall_s"]
    wall = max(per_rank.values()) if per_rank else 0.0
    total = sum(r["output_tokens"] for r in allrows)
    return {"workload": "5-model heterogeneous panel, 8,192-token tech spec, %d new tokens each, temperature 0.7; "
                        "Mistral sliding window off, Phi-3 position limit raised (DESIGN.md)" % args.gen,
            "placement": "one model per GPU" if D.world >= len(HETERO_PANEL) else
                         f"{len(HETERO_PANEL)} models round-robin over {D.world} GPU(s), sequential on a shared GPU",
            "collective": "none", "tokens_per_s": total / wall if wall > 0 else None, "round_wall_s": wall,
            "models": sorted(allrows, key=lambda r: HETERO_PANEL.index(r["model"]))}


def section_converge(args, D, peaks):
    """Config 4: 4 Llama-3-8B opponents x 6 rounds on a 16,384-token tech spec, driven as the reference
    drives it (`debate.py critique --session`, then `--resume` per round; session JSON and round-N.md
    checkpoints written) through the resident driver, so the engine stays loaded across the six
    invocations.  Rank 0: the co-batched placement.  N >= 4: also one opponent per GPU on ranks 0-3,
    every GPU prefilling the prompt itself (engine level, 6 rounds)."""
    import tempfile

    from advspec_b200 import model_spec, resident, runtime

    spec = model_spec.resolve("llama-3-8b")
    doc = make_doc(32000, 16384, 11, "Synthetic Tech Spec")
    gen = args.extras_gen
    out = {"workload": f"4 x Llama-3-8B, 6 rounds, 16,384-token tech spec, {gen} new tokens per opponent per round"}
    if D.rank == 0:
        td = tempfile.mkdtemp(prefix="advspec_bench_")
        panel = ",".join(["b200/llama-3-8b"] * 4)
        rounds = []
        for r in range(1, 7):
            argv = ["critique", "--models", panel, "--doc-type", "tech", "--json"] + \
                   (["--session", "cfg4"] if r == 1 else ["--resume", "cfg4"])
            res = resident.run_cli(argv, doc if r == 1 else "", cwd=td, home=td,
                                   env={"ADVSPEC_MAX_NEW_TOKENS": str(gen)})
            body = json.loads(res.stdout) if res.code == 0 and res.stdout else {}
            toks = sum(x["output_tokens"] for x in body.get("results", []))
            rounds.append({"round": body.get("round"), "rc": res.code, "wall_s": res.wall_s, "output_tokens": toks,
                           "engines_created": res.engines_created, **engine_stats(spec, D.local_rank, peaks)})
        steady = rounds[1:]
        out["batch"] = {
            "placement": "4 opponents co-batched on one GPU (one prefill, b = 4)",
            "driver": "adversarial-spec_b200/resident.py run_cli: debate.py critique --session / --resume",
            "rounds": rounds, "engines_created_after_round_1": sum(r["engines_created"] for r in steady),
            "tokens_per_s_rounds_2_6": sum(r["output_tokens"] for r in steady) / max(sum(r["wall_s"] for r in steady), 1e-9),
            "s_per_round_rounds_2_6": sum(r["wall_s"] for r in steady) / len(steady),
            "session_round_on_disk": json.loads((Path(td) / ".config" / "adversarial-spec" / "sessions" /
                                                 "cfg4.json").read_text())["round"],
            "checkpoints": len(list((Path(td) / ".adversarial-spec-checkpoints").glob("*.md")))}
    if D.world >= 4:
        from advspec_b200 import envelope
        from advspec_b200.tokenizer import SyntheticTokenizer, render_chat

        tok = SyntheticTokenizer(spec.vocab_size)
        active = D.rank < 4
        walls, toks = [], 0
        for r in range(0, 7):
            system_prompt, user_message = envelope.build_messages(doc, max(r, 1), "tech")
            ids = tok.encode(render_chat(system_prompt, user_message), bos=True)
            D.barrier()
            t0 = time.perf_counter()
            if active:
                with runtime.POOL.lease(spec, D.local_rank, len(ids), gen) as res, res.lock:
                    e = res.engine
                    pid = e.prefill(ids)
                    dec = e.decode(e.fork(pid, [runtime.opponent_seed(r, D.rank)]), gen, temperature=0.7)
                    e.release_prefix(pid)
                    runtime.PREFIXES.forget(e)
            D.barrier()
            if r > 0:
                walls.append(time.perf_counter() - t0)
                toks += sum(dec.lens) if active else 0
        rows = D.gather({"rank": D.rank, "tokens": toks, "wall_s": sum(walls),
                         **(engine_stats(spec, D.local_rank, peaks) if active else {})})
        wall = max(r["wall_s"] for r in rows)
        out["spread"] = {"placement": "one opponent per GPU on 4 GPUs, every GPU prefills the 17.5K-token prompt",
                         "collective": "none", "tokens_per_s": sum(r["tokens"] for r in rows) / wall,
                         "s_per_round": wall / 6, "per_rank": rows[:4]}
    return out


def section_tp(args, D, peaks):
    """Config 5: ONE Llama-3-70B opponent tensor-parallel over the N GPUs (N >= 2), 32,768-token tech spec,
    through `call_models_parallel` with ADVSPEC_TP=N (every rank makes the same call).  NVLink carries two
    residual-stream exchanges per layer and the sampler's winners."""
    from advspec_b200 import models as amodels, model_spec, runtime

    if D.world < 2:
        return {"skipped": "needs >= 2 GPUs (141 GB of bf16 weights)"}
    spec = model_spec.resolve("llama-3-70b")
    gen = args.extras_gen
    doc = make_doc(32000, 32768, 11, "Synthetic Tech Spec")
    runtime.POOL.close()  # the other sections' engines leave the GPUs: 70.5 GB of weights per GPU at TP=2
    saved = {k: os.environ.get(k) for k in ("ADVSPEC_TP", "ADVSPEC_MAX_NEW_TOKENS", "ADVSPEC_MIN_PREFIX", "ADVSPEC_MIN_NEW")}
    os.environ.update(ADVSPEC_TP=str(D.world), ADVSPEC_MAX_NEW_TOKENS=str(gen), ADVSPEC_MIN_PREFIX="0", ADVSPEC_MIN_NEW="16")
    try:
        rows = []
        for rep in range(3):
            D.barrier()
            t0 = time.perf_counter()
            res = amodels.call_models_parallel(["b200/llama-3-70b"], doc, 70 + rep, "tech")
            D.barrier()
            wall = time.perf_counter() - t0
            if res[0].error:
                return {"error": res[0].error}
            rows.append({"wall_s": wall, "input_tokens": res[0].input_tokens, "output_tokens": res[0].output_tokens,
                         **engine_stats(spec, D.local_rank, peaks)})
        best = min(rows[1:], key=lambda r: r["decode_ms_per_step"])
        stats = D.gather(best)
        pre = max(s["prefill_ms"] for s in stats)
        step = max(s["decode_ms_per_step"] for s in stats)
        ptoks = best["input_tokens"]
        step_bytes_gpu = spec.decode_step_bytes(ptoks, [gen // 2]) / D.world
        flops = spec.prefill_flops(ptoks)
        return {"workload": f"1 x Llama-3-70B TP={D.world}, 32,768-token tech spec (+envelope = {ptoks} prompt "
                            f"tokens), {gen} new tokens, temperature 0.7",
                "tp": D.world, "decode_ms_per_step": step, "tokens_per_s_decode": 1e3 / step,
                "decode_bytes_per_step_per_gpu": step_bytes_gpu,
                "decode_hbm_gbs_per_gpu": step_bytes_gpu / (step * 1e-3) / 1e9,
                "decode_hbm_frac_per_gpu": step_bytes_gpu / (step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "prefill_ms": pre, "prefill_tflops_aggregate": flops / (pre * 1e-3) / 1e12,
                "prefill_frac_of_bf16_peak_per_gpu": flops / (pre * 1e-3) / 1e12 / D.world / peaks["bf16_tflops"],
                "round_wall_s": min(r["wall_s"] for r in rows[1:]),
                "tokens_per_s_round": best["output_tokens"] / min(r["wall_s"] for r in rows[1:]),
                "first_round_wall_s_incl_engine_creation": rows[0]["wall_s"]}
    finally:
        runtime.POOL.close()
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


SECTION_FNS = {"strong": section_strong, "personas": section_personas, "hetero": section_hetero, "converge": section_converge, "tp": section_tp}


# ----------------------------------------------------------------------------- DRAM traffic (ncu pass)
def measure_gemv_traffic(args, spec):
    """DRAM bytes (read + write) of the GEMV launches of ONE decode step, from an ncu pass over
    tools/traffic_probe.py run by this script after the timed region (N = 1 only).  Returns
    (bytes per step or None, how)."""
    import shutil

    if args.traffic == "off":
        return None, "not measured (--traffic off)"
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if not ncu:
        return _committed_traffic(spec), "ncu not on PATH; committed capture profiles/r01_gemv_traffic.json"
    import csv
    import tempfile

    per_step = 4 * spec.n_layers + 1
    log = tempfile.NamedTemporaryFile(prefix="advspec_ncu_", suffix=".csv", delete=False)
    log.close()
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "-k", "regex:gemv_mma_kernel", "-s", "1", "-c", str(per_step), "--csv", "--log-file", log.name,
           sys.executable, str(ROOT / "tools" / "traffic_probe.py"), "--model", spec.name,
           "--opponents", str(args.opponents), "--spec-tokens", str(args.spec_tokens)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        rows = list(csv.reader(ln for ln in open(log.name) if ln.startswith('"')))
    except (subprocess.TimeoutExpired, OSError) as ex:
        return _committed_traffic(spec), f"ncu pass failed ({type(ex).__name__}); committed capture"
    finally:
        try:
            os.unlink(log.name)
        except OSError:
            pass
    if len(rows) < 2 or "TRAFFIC_PROBE " not in p.stdout + p.stderr:
        return _committed_traffic(spec), "ncu pass gave no rows; committed capture profiles/r01_gemv_traffic.json"
    head = rows[0]
    try:
        i_name, i_val, i_unit = head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
    except ValueError:
        return _committed_traffic(spec), "ncu csv not understood; committed capture"
    total, seen = 0.0, 0
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[1:]:
        if len(r) <= max(i_val, i_unit) or "dram__bytes" not in r[i_name]:
            continue
        total += float(r[i_val].replace(",", "")) * mult.get(r[i_unit], 1.0)
        seen += 1
    if seen != 2 * per_step:
        return _committed_traffic(spec), f"ncu saw {seen // 2} of {per_step} GEMV launches; committed capture"
    return total, (f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the {per_step} gemv_mma_kernel launches of one "
                   f"decode step of tools/traffic_probe.py, run by bench.py after the timed region")


def _committed_traffic(spec):
    p = ROOT / "profiles" / "r01_gemv_traffic.json"
    if not p.exists() or spec.name != "llama-3-8b":
        return None
    d = json.loads(p.read_text())
    return (d["per_layer_traffic_mb"] * spec.n_layers + spec.vocab_size * spec.d_model * 2 / 1e6) * 1e6


def _prompt_ids(spec, system_prompt, user_message):
    from advspec_b200.tokenizer import SyntheticTokenizer, render_chat

    return SyntheticTokenizer(spec.vocab_size).encode(render_chat(system_prompt, user_message), bos=True)


# ----------------------------------------------------------------------------- B200 arm
def run_b200_arm(args, rank, world, local_rank):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    D = Dist(rank, world, local_rank)
    os.environ["ADVSPEC_DEVICES"] = str(local_rank)
    os.environ["ADVSPEC_MAX_NEW_TOKENS"] = str(args.gen)
    os.environ["ADVSPEC_PLACEMENT"] = "batch"
    wanted = sections_of(args)
    if wanted:
        # one resident Llama-3-8B engine serves the headline AND the 16K-token convergence config
        os.environ["ADVSPEC_MIN_PREFIX"] = "17920"
        os.environ["ADVSPEC_MIN_NEW"] = str(max(args.gen, args.extras_gen) + 16)

    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import models as amodels, runtime

    spec, doc, system_prompt, user_message, prompt_tokens = workload(args)
    names = [f"b200/{spec.name}"] * args.opponents

    def one_round(round_num):
        results = amodels.call_models_parallel(names, doc, round_num, args.doc_type)
        bad = [r.error for r in results if r.error]
        if bad:
            raise SystemExit(f"bench.py: opponent failed: {bad[0]}")
        return results

    def engine():
        return runtime.POOL.get(spec, local_rank, prompt_tokens, args.gen).engine

    with Quiet():
        for w in range(args.warmup):
            one_round(100 + w)
        sampler = ClockSampler(local_rank)
        D.barrier()
        sampler.start()
        launches0 = engine().timing().kernel_launches
        pf0 = dict(runtime.PREFIXES.stats)
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
        D.barrier()
        wall = time.perf_counter() - t0
        launches = engine().timing().kernel_launches - launches0
        clocks = sampler.stop()
        pf1 = runtime.PREFIXES.stats
        reused = (pf1["extended"] - pf0["extended"]) + (pf1["rearmed"] - pf0["rearmed"])
        if reused or pf1["full"] - pf0["full"] != args.steps:
            raise SystemExit(f"bench.py: {reused} of the {args.steps} timed steps reused the previous step's prefix "
                             f"(every step must prefill its whole prompt)")

    # max over ranks of the time, sum over ranks of the tokens
    (dev_ms, wall), (out_tokens, launches) = runtime.reduce_round_stats([dev_ms, wall], [out_tokens, launches],
                                                                        device="cuda")
    out_tokens, launches = int(out_tokens), int(launches)

    # N > 1: the SAME 'world'-opponent panel spread one opponent per GPU (configs[1]'s literal placement) with the
    # prefix prefilled once on rank 0 and its KV broadcast over NVLink, against every rank recomputing the prefill.
    spread = None
    if world > 1:
        prompt_ids = _prompt_ids(spec, system_prompt, user_message)
        e = engine()
        runtime.PREFIXES.forget(e)

        def spread_round(broadcast: bool):
            D.barrier()
            t0 = time.perf_counter()
            if broadcast:
                pid = runtime.replicate_prefix(e, prompt_ids, rank, src=0, device=local_rank)
            else:
                pid = e.prefill(prompt_ids)
            ids = e.fork(pid, [runtime.opponent_seed(7, rank)])
            res = e.decode(ids, args.gen, temperature=0.7)
            e.release_prefix(pid)
            D.barrier()
            return time.perf_counter() - t0, sum(res.lens)

        spread = {}
        for mode, bc in (("kv_broadcast", True), ("recompute", False)):
            spread_round(bc)  # warm (NCCL communicator, graph for b = 1)
            times, toks = zip(*[spread_round(bc) for _ in range(max(1, min(args.steps, 3)))])
            (tmax,), (tsum,) = runtime.reduce_round_stats([sum(times)], [sum(toks)], device="cuda")
            spread[mode] = {"tokens_per_s": tsum / tmax, "s_per_round": tmax / len(times), "opponents": world}

    peaks = measured_peaks()
    line = None
    if rank == 0:
        e = engine()
        runtime.PREFIXES.forget(e)
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
                         "kernel_ms_per_step": gemv_ms, "traffic": None,
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

    # ---- the other BASELINE configs on the same GPUs (guarded: never at the headline's expense)
    done = threading.Event()

    def on_timeout():
        if done.is_set():
            return
        if rank == 0 and line is not None:
            line.setdefault("sections_note", f"a section exceeded the {args.sections_budget_s:.0f} s budget; "
                                             f"the line was emitted by the watchdog")
            _emit(line)
        os._exit(0)

    dog = threading.Timer(args.sections_budget_s, on_timeout)
    dog.daemon = True
    if wanted:
        dog.start()
    t_sections = time.perf_counter()
    for name in wanted:
        if time.perf_counter() - t_sections > args.sections_budget_s * 0.8:
            if line is not None:
                line[name] = {"skipped": "section budget spent"}
            continue
        t1 = time.perf_counter()
        try:
            with Quiet():
                rec = SECTION_FNS[name](args, D, peaks)
        except BaseException as ex:  # noqa: BLE001 - a section must never sink the headline
            rec = {"error": f"{type(ex).__name__}: {ex}"[:400]}
            if isinstance(ex, KeyboardInterrupt):
                raise
        if line is not None:
            rec["section_s"] = round(time.perf_counter() - t1, 2)
            line[name] = rec
    done.set()
    dog.cancel()

    if rank == 0:
        if world == 1:
            with Quiet():
                runtime.POOL.close()  # the ncu pass and the CPU baseline want the GPU memory and the cores
            traffic, how = measure_gemv_traffic(args, spec)
            line["roofline"]["traffic"] = traffic
            line["roofline"]["traffic_source"] = how
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            try:
                line["cpu_baseline"] = cpu_baseline(args, system_prompt, user_message, cores)
            except Exception as ex:  # the baseline must never sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"failed: {ex}"}
        _emit(line)
    D.close()
    return line


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
