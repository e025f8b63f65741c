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
