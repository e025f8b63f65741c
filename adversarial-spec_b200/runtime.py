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
import time
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


_plumbing = {"group": None}


def plumbing_group():
    """A gloo process group for the few bytes of host-side plumbing (NCCL id, CUDA IPC handles, the spec text):
    the default group when it already is gloo, else a gloo group created beside it (every rank must call this
    at the same point — the callers below are collective anyway).  Host tensors only, so it does not matter
    which thread or CUDA device the caller is on."""
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group("gloo")
    if _plumbing["group"] is None:
        _plumbing["group"] = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
    return _plumbing["group"]


def broadcast_object(obj, src: int = 0):
    """Plumbing for the tensor-parallel CLI: one Python object from rank `src` to every rank of the
    torchrun job."""
    import torch.distributed as dist

    box = [obj]
    dist.broadcast_object_list(box, src=src, group=plumbing_group())
    return box[0]


def create_tp_engine(spec: ModelSpec, device: int, max_prefix_tokens: int, max_new_tokens: int, max_seqs: int,
                     tp_rank: int, tp_size: int) -> eng.Engine:
    """This rank's handle of a tensor-parallel engine: rank 0 draws the NCCL id, torch.distributed
    (the plumbing; gloo, host tensors) carries it to the others, every rank joins (include/advspec_engine.h)."""
    import torch
    import torch.distributed as dist

    pg = plumbing_group()
    e = eng.Engine(spec, device, max_prefix_tokens, max_new_tokens, max_seqs, tp_rank=tp_rank, tp_size=tp_size)
    buf = torch.zeros(128, dtype=torch.uint8)
    if tp_rank == 0:
        buf = torch.frombuffer(bytearray(eng.Engine.tp_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(buf, src=0, group=pg)
    e.tp_init(bytes(buf.numpy().tobytes()))
    if os.environ.get("ADVSPEC_TP_NCCL_ONLY") is None:
        # decode-step exchange over NVLink peer memory: gather every rank's CUDA IPC handle
        mine = torch.frombuffer(bytearray(e.tp_ipc_export()), dtype=torch.uint8).clone()
        allh = [torch.zeros_like(mine) for _ in range(tp_size)]
        dist.all_gather(allh, mine, group=pg)
        e.tp_ipc_import([bytes(h.numpy().tobytes()) for h in allh])
        dist.barrier(group=pg)  # nobody pushes before everybody has mapped
    return e


@dataclass
class Generation:
    text: str
    prompt_tokens: int
    completion_tokens: int
    token_ids: list[int]
    prefill_ms: float = 0.0
    decode_ms: float = 0.0
    tail_ms: float = 0.0  # wall time of feeding the per-opponent prompt tails (per-opponent personas)


class HFTokenizerAdapter:
    """A model's REAL tokenizer (tokenizers-format `tokenizer.json`) behind the interface the runtime uses.
    Loaded when a weight blob is served (ADVSPEC_WEIGHTS_DIR/<model>.blob + <model>.tokenizer.json, optional
    <model>.meta.json with {"eos_token_id", "bos_token_id", "chat_template": "...{system}...{user}...",
    "rope_inv_freq": [...]}).  Random-init benchmark weights keep the synthetic tokenizer."""

    def __init__(self, tokenizer_json: str, meta: dict, vocab_size: int):
        from tokenizers import Tokenizer

        self.tk = Tokenizer.from_file(tokenizer_json)
        if self.tk.get_vocab_size() > vocab_size:
            raise ValueError(f"{tokenizer_json}: {self.tk.get_vocab_size()} tokens exceed the model's vocabulary "
                             f"of {vocab_size}")
        self.vocab_size = vocab_size
        self.eos_id = int(meta.get("eos_token_id", -1))
        self.bos_id = meta.get("bos_token_id")
        self.template = meta.get("chat_template")
        if self.eos_id < 0:
            raise ValueError("meta.json must give eos_token_id for a real tokenizer")

    def encode(self, text: str, bos: bool = False) -> list[int]:
        ids = self.tk.encode(text, add_special_tokens=False).ids
        return ([int(self.bos_id)] if bos and self.bos_id is not None else []) + ids

    def decode(self, ids) -> str:
        return self.tk.decode([int(t) for t in ids if int(t) >= 0], skip_special_tokens=True)

    def count(self, text: str) -> int:
        return len(self.encode(text))

    def render(self, system_prompt: str, user_message: str) -> str:
        if self.template:
            return self.template.replace("{system}", system_prompt).replace("{user}", user_message)
        return render_chat(system_prompt, user_message)


def tokenizer_for(spec: ModelSpec):
    """(tokenizer, meta) for one model: the real one when a weight blob is served, else synthetic.
    A blob WITHOUT its tokenizer is an error — feeding trained weights synthetic ids produces noise."""
    wdir = os.environ.get("ADVSPEC_WEIGHTS_DIR")
    if wdir and os.path.exists(os.path.join(wdir, f"{spec.name}.blob")):
        tj = os.path.join(wdir, f"{spec.name}.tokenizer.json")
        mj = os.path.join(wdir, f"{spec.name}.meta.json")
        if not os.path.exists(tj):
            raise FileNotFoundError(f"{wdir}/{spec.name}.blob is served but {tj} is missing: a weight blob needs "
                                    f"its own tokenizer (and {spec.name}.meta.json with eos_token_id)")
        import json

        meta = json.loads(open(mj).read()) if os.path.exists(mj) else {}
        return HFTokenizerAdapter(tj, meta, spec.vocab_size), meta
    return SyntheticTokenizer(spec.vocab_size), {}


def render_prompt(tok, system_prompt: str, user_message: str) -> str:
    return tok.render(system_prompt, user_message) if hasattr(tok, "render") else render_chat(system_prompt, user_message)


@dataclass
class _Resident:
    engine: eng.Engine
    tok: object
    lock: threading.Lock = field(default_factory=threading.Lock)
    max_prefix: int = 0
    max_new: int = 0
    users: int = 0  # callers holding this engine (EnginePool.lease); it is never closed while > 0


class EnginePool:
    """Engines stay resident across calls and rounds (model load would otherwise dominate).

    Callers `lease()` an engine for the duration of a round.  An engine that is too small for a new
    request is replaced only when nobody holds it: the pool waits for its users to drain, so a
    concurrent round on the same (model, device) can never be handed a destroyed handle."""

    def __init__(self):
        self._engines: dict[tuple[str, int], _Resident] = {}
        self._mu = threading.Lock()
        self._idle = threading.Condition(self._mu)
        self.created = 0  # engines ever created by this pool (resident.py reports it per invocation)

    def _build(self, spec: ModelSpec, device: int, need_prefix: int, need_new: int) -> _Resident:
        cap_prefix = max(need_prefix, _env_int("ADVSPEC_MIN_PREFIX", 0))
        cap_prefix = (cap_prefix + 255) // 256 * 256
        cap_new = max(need_new, _env_int("ADVSPEC_MIN_NEW", 16))
        tp_rank, tp_size = tp_world()
        tok, meta = tokenizer_for(spec)
        if tp_size > 1:
            e = create_tp_engine(spec, device, cap_prefix, cap_new, MAX_BATCH, tp_rank, tp_size)
        else:
            e = eng.Engine(spec, device, cap_prefix, cap_new, MAX_BATCH)
        wdir = os.environ.get("ADVSPEC_WEIGHTS_DIR")
        blob_path = os.path.join(wdir, f"{spec.name}.blob") if wdir else None
        if blob_path and os.path.exists(blob_path):
            from .weights import shard_blob
            e.load_weights(shard_blob(np.fromfile(blob_path, dtype=np.uint8), spec, tp_rank, tp_size))
            if meta.get("rope_inv_freq"):
                e.set_rope_inv_freq(np.asarray(meta["rope_inv_freq"], dtype=np.float32))
        else:
            e.init_weights_random(_env_int("ADVSPEC_WEIGHT_SEED", 0), 0.02)
        self.created += 1
        return _Resident(e, tok, max_prefix=cap_prefix, max_new=cap_new)

    def acquire(self, spec: ModelSpec, device: int, need_prefix: int, need_new: int) -> _Resident:
        key = (spec.name, device)
        with self._idle:
            while True:
                r = self._engines.get(key)
                if r is not None and r.max_prefix >= need_prefix and r.max_new >= need_new:
                    r.users += 1
                    return r
                if r is not None and r.users > 0:
                    self._idle.wait()  # too small, but somebody is inside it: replace it once they are done
                    continue
                if r is not None:
                    del self._engines[key]
                    PREFIXES.forget(r.engine)
                    r.engine.close()  # nobody holds it and nobody can get it any more
                    need_prefix, need_new = max(need_prefix, r.max_prefix), max(need_new, r.max_new)
                r = self._build(spec, device, need_prefix, need_new)
                r.users = 1
                self._engines[key] = r
                return r

    def release(self, r: _Resident) -> None:
        with self._idle:
            r.users -= 1
            self._idle.notify_all()

    def lease(self, spec: ModelSpec, device: int, need_prefix: int, need_new: int):
        pool = self

        class _Lease:
            def __enter__(self_inner):
                self_inner.r = pool.acquire(spec, device, need_prefix, need_new)
                return self_inner.r

            def __exit__(self_inner, *exc):
                pool.release(self_inner.r)
                return False

        return _Lease()

    def get(self, spec: ModelSpec, device: int, need_prefix: int, need_new: int) -> _Resident:
        """The resident engine without holding it (measurement code that owns the process)."""
        r = self.acquire(spec, device, need_prefix, need_new)
        self.release(r)
        return r

    def resident_count(self) -> int:
        with self._mu:
            return len(self._engines)

    def close(self) -> None:
        with self._idle:
            while any(r.users > 0 for r in self._engines.values()):
                self._idle.wait()
            for r in self._engines.values():
                PREFIXES.forget(r.engine)
                r.engine.close()
            self._engines.clear()


POOL = EnginePool()


class PrefixCache:
    """Cross-call reuse of the live prefix KV of each resident engine.

    The reference assembles system -> "This is round r" -> spec -> context -> focus -> instruction
    (prompts.py:233-241), so calls on the same round and document that differ in `--context`, `--focus` or
    `--preserve-intent` share everything up to the end of the spec.  The engine keeps ONE live prefix; this
    cache remembers its token ids and, when the next prompt shares at least ADVSPEC_PREFIX_REUSE_MIN (default
    0.5) of its length with it, prefills only the tail (`advspec_prefill_extend`).  An identical prompt
    (retry, a panel larger than one batch) is re-armed without any prefill.  Consecutive ROUNDS diverge at
    the round number, a few hundred tokens in, and are prefilled in full."""

    def __init__(self):
        self._live: dict[int, tuple[int, np.ndarray]] = {}  # id(engine) -> (prefix id, token ids)
        self.stats = {"full": 0, "extended": 0, "rearmed": 0, "tokens_reused": 0, "tokens_prefilled": 0}

    def prefill(self, e: eng.Engine, key, prompt_ids: Sequence[int]) -> int:
        ids = np.asarray(prompt_ids, dtype=np.int32)
        n = int(ids.size)
        live = self._live.get(id(e))
        try:
            min_frac = float(os.environ.get("ADVSPEC_PREFIX_REUSE_MIN", "0.5"))
        except ValueError:
            min_frac = 0.5
        if live is not None and os.environ.get("ADVSPEC_PREFIX_CACHE", "1") != "0":
            pid, old = live
            m = min(n, int(old.size))
            neq = np.nonzero(ids[:m] != old[:m])[0]
            lcp = int(neq[0]) if neq.size else m
            try:
                if lcp == n and n == old.size:
                    new = e.prefill_extend(pid, n, [])
                    self.stats["rearmed"] += 1
                    self.stats["tokens_reused"] += n
                    self._live[id(e)] = (new, ids)
                    return new
                keep = min(lcp, n - 1)  # at least the last token runs, to produce the next-token logits
                if keep >= 1 and keep >= min_frac * n:
                    new = e.prefill_extend(pid, keep, ids[keep:].tolist())
                    self.stats["extended"] += 1
                    self.stats["tokens_reused"] += keep
                    self.stats["tokens_prefilled"] += n - keep
                    self._live[id(e)] = (new, ids)
                    return new
            except eng.EngineError as ex:
                if ex.status != 4:  # STATE: someone replaced the live prefix behind the cache -> full prefill
                    raise
        self._live.pop(id(e), None)
        new = e.prefill(ids.tolist())
        self.stats["full"] += 1
        self.stats["tokens_prefilled"] += n
        self._live[id(e)] = (new, ids)
        return new

    def forget(self, e: eng.Engine) -> None:
        self._live.pop(id(e), None)


PREFIXES = PrefixCache()


def effective_max_new(max_tokens: int) -> int:
    cap = _env_int("ADVSPEC_MAX_NEW_TOKENS", 0)
    n = max_tokens if cap <= 0 else min(max_tokens, cap)
    return max(1, n)


def opponent_seed(round_num: int, index: int) -> int:
    base = _env_int("ADVSPEC_SEED", 0)
    return (base * 1_000_003 + round_num * 1009 + index * 7919 + 1) & ((1 << 63) - 1)


def _common_prefix_len(seqs: Sequence[Sequence[int]]) -> int:
    n = min(len(q) for q in seqs)
    first = np.asarray(seqs[0][:n], dtype=np.int64)
    for q in seqs[1:]:
        neq = np.nonzero(np.asarray(q[:n], dtype=np.int64) != first[:n])[0]
        if neq.size:
            n = int(neq[0])
    return n


def plan_tails(prompts: Sequence[Sequence[int]], tail_max: int, min_frac: float = 0.5):
    """Split the prompts of one batch into (tokens of the shared prefix, one tail per opponent), or None
    when they should not share: nothing in common, a tail longer than `tail_max`, or a shared part shorter
    than `min_frac` of the shortest prompt.  Every tail keeps at least one token — the step that consumes an
    opponent's last prompt token is the one that produces its first next-token logits."""
    keep = min(_common_prefix_len(prompts), min(len(q) for q in prompts) - 1)
    if keep < 1 or keep < min_frac * min(len(q) for q in prompts):
        return None
    tails = [list(q[keep:]) for q in prompts]
    if max(len(t) for t in tails) > tail_max:
        return None
    return list(prompts[0][:keep]), tails


def step_tails(e, seq_ids: Sequence[int], tails: Sequence[Sequence[int]]) -> int:
    """Feed each opponent the tokens its prompt has beyond the shared prefix, teacher-forced through the
    batched decode step (`advspec_decode_step`): the tails are RIGHT-aligned, so the opponents with the
    longest tails start alone and the final step carries every opponent in fork order — which is what
    `advspec_decode` needs to sample each opponent's first token from that step's logits.  Returns the
    number of steps."""
    longest = max(len(t) for t in tails)
    for t in range(longest):
        part = [(sid, tl[t - (longest - len(tl))]) for sid, tl in zip(seq_ids, tails) if t >= longest - len(tl)]
        e.decode_step([sid for sid, _ in part], [f for _, f in part])
    return longest


def feed_tails(e, seq_ids: Sequence[int], tails: Sequence[Sequence[int]]) -> None:
    """Give every freshly forked opponent the prompt tokens it has beyond the shared prefix.  Default: one
    GEMM-shaped prompt chunk per opponent (`advspec_append_tail`: one pass over the weights per TAIL, its K/V
    moved into the opponent's own KV).  ADVSPEC_TAIL_IMPL=step feeds the tails token by token through the
    batched decode step instead (one pass over the weights per TOKEN of the longest tail; the A/B)."""
    if os.environ.get("ADVSPEC_TAIL_IMPL", "chunk") == "step":
        step_tails(e, seq_ids, tails)
        return
    for sid, tail in zip(seq_ids, tails):  # fork order = the decode batch's order
        e.append_tail(sid, tail)


def generate_group(spec: ModelSpec, device: int, system_prompt: str, user_message, n_opponents: int,
                   seeds: Sequence[int], max_tokens: int, temperature: float) -> list[Generation]:
    """One round for `n_opponents` opponents of one model on one GPU.  `user_message` is one string for the
    whole panel (the reference's case: identical messages, one shared-prefix prefill) or one string per
    opponent (per-opponent personas): opponents whose prompts differ only at the end still
    share ONE prefill of the common tokens and decode as ONE batch; each opponent's own tail is fed after
    the fork (`feed_tails`).  Prompts that differ early (or by more than ADVSPEC_TAIL_MAX tokens, default
    512) fall back to one prefill + decode per distinct prompt."""
    tok, _ = tokenizer_for(spec)
    users = [user_message] * n_opponents if isinstance(user_message, str) else list(user_message)
    if len(users) != n_opponents:
        raise ValueError(f"{len(users)} user messages for {n_opponents} opponents")
    encoded: dict[str, list[int]] = {}
    for u in users:
        if u not in encoded:
            encoded[u] = tok.encode(render_prompt(tok, system_prompt, u), bos=True)
    prompts = [encoded[u] for u in users]
    max_new = effective_max_new(max_tokens)
    tail_max = _env_int("ADVSPEC_TAIL_MAX", 512)
    # batches of at most MAX_BATCH opponents, each: (indices, shared prefix tokens, tails or None)
    work: list[tuple[list[int], list[int], Optional[list[list[int]]]]] = []
    for g0 in range(0, n_opponents, MAX_BATCH):
        idx = list(range(g0, min(n_opponents, g0 + MAX_BATCH)))
        mine = [prompts[i] for i in idx]
        if all(q is mine[0] or q == mine[0] for q in mine):
            work.append((idx, mine[0], None))
            continue
        plan = plan_tails(mine, tail_max)
        if plan is not None:
            work.append((idx, plan[0], plan[1]))
            continue
        by_prompt: dict[str, list[int]] = {}
        for i in idx:
            by_prompt.setdefault(users[i], []).append(i)
        work.extend((ii, prompts[ii[0]], None) for ii in by_prompt.values())
    longest_tail = max([len(t) for _, _, tails in work if tails for t in tails] or [0])
    out: list[Optional[Generation]] = [None] * n_opponents
    with POOL.lease(spec, device, max(len(q) for q in prompts), max_new + longest_tail) as res, res.lock:
        e = res.engine
        for idx, prefix_ids, tails in work:
            pid = PREFIXES.prefill(e, (spec.name, device), prefix_ids)
            ids = e.fork(pid, [seeds[i] for i in idx])
            tail_ms = 0.0
            if tails is not None:
                t0 = time.perf_counter()
                feed_tails(e, ids, tails)
                tail_ms = (time.perf_counter() - t0) * 1e3
            dec = e.decode(ids, max_new, temperature=temperature, eos_id=tok.eos_id)
            tm = e.timing()
            for i, toks in zip(idx, dec.tokens):
                body = toks[:-1] if (toks and toks[-1] == tok.eos_id) else toks
                out[i] = Generation(res.tok.decode(body), len(prompts[i]), len(toks), toks,
                                    tm.prefill_ms, tm.decode_ms, tail_ms)
            e.release_seqs(ids)
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


def run_round(model_names: Sequence[str], system_prompt: str, user_message, seeds: Sequence[int],
              max_tokens: int, temperature: float, devices: Optional[Sequence[int]] = None) -> list:
    """All local opponents of one critique round.  `user_message`: one string for the panel, or one per
    opponent (per-opponent personas).  Returns, per opponent (input order), a Generation or the Exception
    its group raised."""
    if not isinstance(user_message, str) and len(user_message) != len(model_names):
        raise ValueError(f"{len(user_message)} user messages for {len(model_names)} opponents")
    devices = list(devices) if devices is not None else visible_devices()
    results: list = [None] * len(model_names)
    known: list[int] = []
    for i, m in enumerate(model_names):  # an unknown local name fails THAT opponent, like a bad provider id
        try:
            resolve(m)
            known.append(i)
        except KeyError as ex:
            results[i] = ValueError(str(ex.args[0]) if ex.args else str(ex))
    plan = plan_placement([model_names[i] for i in known], devices)
    for p in plan:
        p.indices = [known[j] for j in p.indices]

    def work(p: Placement):
        try:
            users = user_message if isinstance(user_message, str) else [user_message[i] for i in p.indices]
            gens = generate_group(p.spec, p.device, system_prompt, users, len(p.indices),
                                  [seeds[i] for i in p.indices], max_tokens, temperature)
            for i, g in zip(p.indices, gens):
                results[i] = g
        except Exception as ex:  # surfaced per opponent, like a failed provider call
            for i in p.indices:
                results[i] = ex

    if not plan:
        return results
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
