"""Host-side runtime: resident engines, placement of opponents on GPUs, one batched round.

This is what replaces the provider router the reference reaches through
``litellm.completion`` (skills/adversarial-spec/scripts/models.py:628): opponents
that name the same local model share ONE prefill and one weight stream.
Knobs come from the environment so the reference's CLI surface stays unchanged
(SURVEY.md §5 "Config / flags"):
  ADVSPEC_MAX_NEW_TOKENS  cap on generated tokens per opponent (the CLI hard-codes 8000)
  ADVSPEC_SEED            base sampling seed (default 0)
  ADVSPEC_WEIGHT_SEED     seed of the synthetic weights (default 0)
  ADVSPEC_WEIGHTS_DIR     directory of <model>.blob files; absent -> seeded random init on device
  ADVSPEC_DEVICES         comma list of CUDA devices to use (default: the current/only one)
  ADVSPEC_PLACEMENT       "batch" (default: same-weight opponents share a GPU) | "spread"
  ADVSPEC_TP              tensor-parallel width for every local model: the process must run under
                          torchrun with WORLD_SIZE == ADVSPEC_TP, all ranks make the same calls
                          (one large opponent over 8 GPUs)
"""

from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import engine as eng
from .model_spec import ModelSpec, resolve
from .tokenizer import SyntheticTokenizer, render_chat

MAX_BATCH = 8


def _env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def visible_devices() -> list[int]:
    raw = os.environ.get("ADVSPEC_DEVICES")
    if raw:
        return [int(x) for x in raw.split(",") if x.strip() != ""]
    if "LOCAL_RANK" in os.environ and _env_int("WORLD_SIZE", 1) > 1:
        return [_env_int("LOCAL_RANK", 0)]  # one process per GPU under torchrun
    return [0]


def tp_world() -> tuple[int, int]:
    """(rank, size) of the tensor-parallel group this process belongs to; (0, 1) when ADVSPEC_TP is unset."""
    tp = _env_int("ADVSPEC_TP", 1)
    if tp <= 1:
        return 0, 1
    world = _env_int("WORLD_SIZE", 1)
    if world != tp:
        raise RuntimeError(f"ADVSPEC_TP={tp} needs torchrun with WORLD_SIZE={tp} (got {world}): one process per GPU")
    return _env_int("RANK", 0), tp


def create_tp_engine(spec: ModelSpec, device: int, max_prefix_tokens: int, max_new_tokens: int, max_seqs: int,
                     tp_rank: int, tp_size: int) -> eng.Engine:
    """This rank's handle of a tensor-parallel engine: rank 0 draws the NCCL id, torch.distributed
    (the plumbing; any backend) carries it to the others, every rank joins (include/advspec_engine.h)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group("gloo")  # 128 bytes of plumbing; the engine owns its NCCL communicator
    e = eng.Engine(spec, device, max_prefix_tokens, max_new_tokens, max_seqs, tp_rank=tp_rank, tp_size=tp_size)
    buf = torch.zeros(128, dtype=torch.uint8)
    if tp_rank == 0:
        buf = torch.frombuffer(bytearray(eng.Engine.tp_unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        buf = buf.cuda(device)
    dist.broadcast(buf, src=0)
    e.tp_init(bytes(buf.cpu().numpy().tobytes()))
    if os.environ.get("ADVSPEC_TP_NCCL_ONLY") is None:
        # decode-step exchange over NVLink peer memory: gather every rank's CUDA IPC handle
        mine = torch.frombuffer(bytearray(e.tp_ipc_export()), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            mine = mine.cuda(device)
        allh = [torch.zeros_like(mine) for _ in range(tp_size)]
        dist.all_gather(allh, mine)
        e.tp_ipc_import([bytes(h.cpu().numpy().tobytes()) for h in allh])
        dist.barrier()  # nobody pushes before everybody has mapped
    return e


@dataclass
class Generation:
    text: str
    prompt_tokens: int
    completion_tokens: int
    token_ids: list[int]
    prefill_ms: float = 0.0
    decode_ms: float = 0.0


@dataclass
class _Resident:
    engine: eng.Engine
    tok: SyntheticTokenizer
    lock: threading.Lock = field(default_factory=threading.Lock)
    max_prefix: int = 0
    max_new: int = 0


class EnginePool:
    """Engines stay resident across calls and rounds (model load would otherwise dominate)."""

    def __init__(self):
        self._engines: dict[tuple[str, int], _Resident] = {}
        self._mu = threading.Lock()

    def get(self, spec: ModelSpec, device: int, need_prefix: int, need_new: int) -> _Resident:
        key = (spec.name, device)
        with self._mu:
            r = self._engines.get(key)
            if r is not None and (r.max_prefix < need_prefix or r.max_new < need_new):
                r.engine.close()
                r = None
            if r is None:
                cap_prefix = max(need_prefix, _env_int("ADVSPEC_MIN_PREFIX", 0))
                cap_prefix = (cap_prefix + 255) // 256 * 256
                cap_new = max(need_new, 16)
                tp_rank, tp_size = tp_world()
                if tp_size > 1:
                    e = create_tp_engine(spec, device, cap_prefix, cap_new, MAX_BATCH, tp_rank, tp_size)
                else:
                    e = eng.Engine(spec, device, cap_prefix, cap_new, MAX_BATCH)
                wdir = os.environ.get("ADVSPEC_WEIGHTS_DIR")
                blob_path = os.path.join(wdir, f"{spec.name}.blob") if wdir else None
                if blob_path and os.path.exists(blob_path):
                    from .weights import shard_blob
                    e.load_weights(shard_blob(np.fromfile(blob_path, dtype=np.uint8), spec, tp_rank, tp_size))
                else:
                    e.init_weights_random(_env_int("ADVSPEC_WEIGHT_SEED", 0), 0.02)
                r = _Resident(e, SyntheticTokenizer(spec.vocab_size), max_prefix=cap_prefix, max_new=cap_new)
                self._engines[key] = r
            return r

    def close(self) -> None:
        with self._mu:
            for r in self._engines.values():
                r.engine.close()
            self._engines.clear()


POOL = EnginePool()


def effective_max_new(max_tokens: int) -> int:
    cap = _env_int("ADVSPEC_MAX_NEW_TOKENS", 0)
    n = max_tokens if cap <= 0 else min(max_tokens, cap)
    return max(1, n)


def opponent_seed(round_num: int, index: int) -> int:
    base = _env_int("ADVSPEC_SEED", 0)
    return (base * 1_000_003 + round_num * 1009 + index * 7919 + 1) & ((1 << 63) - 1)


def generate_group(spec: ModelSpec, device: int, system_prompt: str, user_message: str, n_opponents: int,
                   seeds: Sequence[int], max_tokens: int, temperature: float) -> list[Generation]:
    """One shared-prefix round for `n_opponents` opponents of one model on one GPU."""
    tok = SyntheticTokenizer(spec.vocab_size)
    prompt_ids = tok.encode(render_chat(system_prompt, user_message), bos=True)
    max_new = effective_max_new(max_tokens)
    res = POOL.get(spec, device, len(prompt_ids), max_new)
    out: list[Generation] = []
    with res.lock:
        e = res.engine
        for g0 in range(0, n_opponents, MAX_BATCH):
            batch_seeds = list(seeds[g0: g0 + MAX_BATCH])
            pid = e.prefill(prompt_ids)
            ids = e.fork(pid, batch_seeds)
            dec = e.decode(ids, max_new, temperature=temperature, eos_id=tok.eos_id)
            tm = e.timing()
            for toks in dec.tokens:
                body = toks[:-1] if (toks and toks[-1] == tok.eos_id) else toks
                out.append(Generation(res.tok.decode(body), len(prompt_ids), len(toks), toks,
                                      tm.prefill_ms, tm.decode_ms))
            e.release_prefix(pid)
    return out


@dataclass
class Placement:
    device: int
    spec: ModelSpec
    indices: list[int]  # positions in the caller's opponent list


def plan_placement(model_names: Sequence[str], devices: Sequence[int], policy: Optional[str] = None) -> list[Placement]:
    """Opponents -> GPUs.  Heterogeneous panels: one model per GPU round-robin.  Same-weight
    replicas: "batch" keeps them on one GPU behind a single prefill and one weight stream
    (same aggregate tokens/s as one-per-GPU); "spread" puts one per GPU."""
    policy = policy or os.environ.get("ADVSPEC_PLACEMENT", "batch")
    groups: dict[str, list[int]] = {}
    for i, m in enumerate(model_names):
        groups.setdefault(resolve(m).name, []).append(i)
    out: list[Placement] = []
    d = 0
    for name, idxs in groups.items():
        spec = resolve(name)
        if policy == "spread" and len(devices) > 1:
            for i in idxs:
                out.append(Placement(devices[d % len(devices)], spec, [i]))
                d += 1
        else:
            out.append(Placement(devices[d % len(devices)], spec, idxs))
            d += 1
    # merge placements that landed on the same (device, model)
    merged: dict[tuple[int, str], Placement] = {}
    for p in out:
        k = (p.device, p.spec.name)
        if k in merged:
            merged[k].indices.extend(p.indices)
        else:
            merged[k] = p
    return list(merged.values())


def run_round(model_names: Sequence[str], system_prompt: str, user_message: str, seeds: Sequence[int],
              max_tokens: int, temperature: float, devices: Optional[Sequence[int]] = None) -> list:
    """All local opponents of one critique round.  Returns, per opponent (input order), a
    Generation or the Exception its group raised."""
    devices = list(devices) if devices is not None else visible_devices()
    plan = plan_placement(model_names, devices)
    results: list = [None] * len(model_names)

    def work(p: Placement):
        try:
            gens = generate_group(p.spec, p.device, system_prompt, user_message, len(p.indices),
                                  [seeds[i] for i in p.indices], max_tokens, temperature)
            for i, g in zip(p.indices, gens):
                results[i] = g
        except Exception as ex:  # surfaced per opponent, like a failed provider call
            for i in p.indices:
                results[i] = ex

    if len(plan) == 1:
        work(plan[0])
    else:
        with ThreadPoolExecutor(max_workers=len(plan)) as pool:
            list(pool.map(work, plan))
    return results


def reduce_round_stats(max_values: Sequence[float], sum_values: Sequence[int], device: str = "cpu"):
    """Multi-GPU bookkeeping of a round: times are the MAX over ranks, token/launch counts the SUM.
    No data-path collective exists between panels; this is measurement only.
    Works with any initialised torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(max_values), list(sum_values)
    m = torch.tensor(list(max_values), dtype=torch.float64, device=device)
    s = torch.tensor(list(sum_values), dtype=torch.int64, device=device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return m.tolist(), s.tolist()


def shard_panels(n_panels: int, rank: int, world: int) -> list[int]:
    """Independent panels (or opponents of a heterogeneous panel) -> ranks, round-robin."""
    return [i for i in range(n_panels) if i % world == rank]


# ----------------------------------------------------------------------------- replica placement
class _DevMem:
    """Lets torch alias a raw device allocation of the engine (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def device_bytes_tensor(ptr: int, nbytes: int, device: int):
    import torch

    return torch.as_tensor(_DevMem(ptr, nbytes), device=f"cuda:{device}")


def replicate_prefix(engine, prompt_ids: Sequence[int], rank: int, src: int = 0, broadcast=None,
                     alias=None, device: int = 0) -> int:
    """Same-weight replicas one per GPU: rank `src` prefills the shared prompt ONCE,
    then its prefix KV region and last-position logits are broadcast (NCCL over NVLink) and the other
    ranks adopt them instead of recomputing the prefill.  This is the path's only real exchange step.
    Returns the local prefix id.  `broadcast`/`alias` are injectable for the gloo CPU test."""
    import torch
    import torch.distributed as dist

    broadcast = broadcast or (lambda t: dist.broadcast(t, src=src))
    alias = alias or (lambda ptr, n: device_bytes_tensor(ptr, n, device))
    vocab = engine.spec.vocab_size
    if rank == src:
        pid = engine.prefill(prompt_ids)
        logits = torch.from_numpy(engine.get_logits(1)[0].copy())
    else:
        pid = None
        logits = torch.empty(vocab, dtype=torch.float32)
    on_gpu = torch.cuda.is_available() and alias is not None and dist.is_initialized() and dist.get_backend() == "nccl"
    lg = logits.cuda(device) if on_gpu else logits
    broadcast(lg)
    if rank != src:
        pid = engine.prefix_adopt(len(prompt_ids), lg.cpu().numpy())
    ptr, nbytes = engine.prefix_kv_region(pid)
    kv = alias(ptr, nbytes)
    broadcast(kv)
    if on_gpu:
        torch.cuda.synchronize(device)
    return pid
