"""CPU, world_size 2 over gloo: the multi-GPU path of bench.py (scene sharding + max-over-ranks timing)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unidet3d_b200 import sharding
    import bench
    mine = sharding.shard_indices(10, rank, world)
    # each rank builds its own, different, batch (bench.make_workload seeds scenes by rank)
    cfg, scenes, names, preset = bench.make_workload("small_1", rank)
    fingerprint = float(scenes[0][0][:100].sum())
    t_dev, t_e2e = sharding.aggregate_times([1.0 + rank, 5.0 - rank])
    q.put((rank, mine, fingerprint, t_dev, t_e2e))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, s0, f0, a0, b0), (r1, s1, f1, a1, b1) = out
    assert sorted(s0 + s1) == list(range(10)) and not set(s0) & set(s1)       # disjoint cover
    assert f0 != f1                                                           # ranks work on different scenes
    assert a0 == a1 == 2.0 and b0 == b1 == 5.0                                # MAX over ranks, same on every rank
    from unidet3d_b200 import sharding
    assert sharding.whole_job_throughput(8, 2, 2.0) == 8.0
