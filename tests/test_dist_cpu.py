"""The N>1 bookkeeping on CPU: gloo, world_size 2 (the data path itself has no collective)."""

import os

import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import advspec_loader

    advspec_loader.load()
    import torch.distributed as dist

    from advspec_b200 import runtime

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mx, sm = runtime.reduce_round_stats([10.0 + rank, 20.0 - rank], [100 * (rank + 1), 7])
    devs = runtime.visible_devices()
    shard = runtime.shard_panels(5, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mx, sm, devs, shard))


def test_reduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29631, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    for rank, mx, sm, devs, shard in got:
        assert mx == [11.0, 20.0] and sm == [300, 14]
        assert devs == [rank], "one process per GPU: LOCAL_RANK picks the device"
    assert got[0][4] == [0, 2, 4] and got[1][4] == [1, 3]


def test_single_process_is_identity():
    import advspec_loader

    advspec_loader.load()
    from advspec_b200 import runtime

    assert runtime.reduce_round_stats([1.5], [3]) == ([1.5], [3])
    assert runtime.shard_panels(3, 0, 1) == [0, 1, 2]


class _FakeEngine:
    """Host-memory stand-in with the engine's replica-placement surface (the real one needs a GPU)."""

    def __init__(self, vocab, kv_bytes):
        import numpy as np
        from types import SimpleNamespace

        self.spec = SimpleNamespace(vocab_size=vocab)
        self.kv = np.zeros(kv_bytes, dtype=np.uint8)
        self.logits = None
        self.adopted = None

    def prefill(self, ids):
        import numpy as np

        self.kv[:] = (np.arange(self.kv.size) * 7 + len(ids)) % 251
        self.logits = np.linspace(-1, 1, self.spec.vocab_size, dtype=np.float32)
        return 42

    def get_logits(self, n):
        return self.logits[None]

    def prefix_adopt(self, n_tokens, logits):
        self.adopted = (n_tokens, logits.copy())
        return 1048576

    def prefix_kv_region(self, pid):
        return self.kv.ctypes.data, self.kv.size


def _replica_worker(rank, world, port, q):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import advspec_loader

    advspec_loader.load()
    import numpy as np
    import torch
    import torch.distributed as dist

    from advspec_b200 import runtime

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = _FakeEngine(64, 4096)
    alias = lambda ptr, n: torch.from_numpy(e.kv)  # noqa: E731  (host memory instead of a device pointer)
    pid = runtime.replicate_prefix(e, list(range(37)), rank, src=0, alias=alias)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, pid, int(e.kv.astype(np.int64).sum()), None if e.adopted is None else (e.adopted[0], float(e.adopted[1].sum()))))


def test_replica_prefix_broadcast_world2():
    """Rank 0 prefills once; rank 1 adopts the broadcast KV bytes and logits instead of recomputing."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_replica_worker, args=(r, 2, 29633, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    (r0, pid0, sum0, ad0), (r1, pid1, sum1, ad1) = got
    assert pid0 == 42 and pid1 == 1048576
    assert sum0 == sum1 and sum0 > 0, "rank 1 must hold rank 0's KV bytes"
    assert ad0 is None and ad1[0] == 37 and abs(ad1[1]) < 1e-3


# ---------------------------------------------------------------- tensor parallelism (host + sharding)
def _tp_worker(rank, world, port, q):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import numpy as np
    import torch
    import torch.distributed as dist
    import advspec_loader
    advspec_loader.load()
    from advspec_b200 import model_spec, weights
    from oracle import hf_oracle, restate
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    spec = model_spec.resolve("tiny-gqa4")
    full = hf_oracle.export_blob(spec, hf_oracle.build_hf_model(spec, 3))
    tokens = np.random.default_rng(1).integers(0, spec.vocab_size, 24).tolist()
    loc = weights.tp_local_spec(spec, world)
    shard = weights.shard_blob(full, spec, rank, world)
    m = restate.BlobModel(loc, shard, embed_rows=spec.vocab_size)
    mine = m.forward_logits(tokens, tp_rank=rank, allreduce=lambda t: dist.all_reduce(t))
    parts = [torch.zeros(mine.shape) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(mine))
    got = torch.cat(parts, dim=1).numpy()
    if rank == 0:
        want = restate.BlobModel(spec, full).forward_logits(tokens)
        q.put((float(np.abs(got - want).max()), float(want.std())))
    dist.destroy_process_group()


def test_tensor_parallel_sharding_matches_whole_model_world2():
    """weights.shard_blob + the exchange points stated in include/advspec_engine.h (all-reduce of the
    residual after o-proj and down-proj, rank 0 carries the residual; logits = concatenated vocab
    shards) reproduce the whole-model forward: two gloo ranks, restatement oracle as the arithmetic."""
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 17
    ps = [ctx.Process(target=_tp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    err, std = q.get(timeout=120)
    for p in ps:
        p.join(30)
    assert err < 1e-4 * max(std, 1.0), (err, std)


def _tp_cli_worker(rank, world, port, q, stdin_path):
    """One rank of `ADVSPEC_TP=2 torchrun debate.py critique`: only rank 0 is given the spec on stdin."""
    import io
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import advspec_loader

    advspec_loader.load()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), ADVSPEC_TP=str(world))
    from advspec_b200 import debate, runtime

    sys.stdin = open(stdin_path) if rank == 0 else io.StringIO("")  # torchrun children share ONE stdin
    spec = debate.read_spec_from_stdin()
    # the same gloo group then carries the engine's NCCL id / IPC handles (runtime.create_tp_engine)
    import torch
    import torch.distributed as dist

    buf = torch.full((4,), float(rank))
    dist.broadcast(buf, src=0, group=runtime.plumbing_group())
    dist.barrier(group=runtime.plumbing_group())
    dist.destroy_process_group()
    q.put((rank, spec, buf.tolist()))


def test_tensor_parallel_cli_broadcasts_the_one_stdin(tmp_path):
    """under torchrun the ranks inherit one stdin; rank 0 reads the spec and every rank gets it."""
    text = "  # Spec\n\nThe service must store data.\n"
    p = tmp_path / "spec.md"
    p.write_text(text)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tp_cli_worker, args=(r, 2, 29647, q, str(p))) for r in range(2)]
    [pr.start() for pr in procs]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [pr.join(timeout=60) for pr in procs]
    assert [g[1] for g in got] == [text.strip()] * 2
    assert got[1][2] == [0.0] * 4
