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
    # SyncBatchNorm exchange of the train-mode backbone: each rank holds different rows; after ONE all-reduce of
    # [sum, sum of squares, count] every rank folds the statistics of the union (ops.sync_bn_sums)
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(300 + 200 * rank, 8, generator=g, dtype=torch.float64) * (1 + rank) + rank
    sums, count = ops.sync_bn_sums(torch.stack((x.sum(0), (x * x).sum(0))), float(x.shape[0]))
    mean = sums[0] / count
    var = sums[1] / count - mean * mean
    # the variant the GPU path uses: the global row count stays a tensor (read by the kernels on the device, no host wait)
    sums_d, local_count, count_t = ops.sync_bn_sums(torch.stack((x.sum(0), (x * x).sum(0))), float(x.shape[0]), device_count=True)
    assert local_count == float(x.shape[0]) and count_t.shape == (1,) and float(count_t) == count and torch.equal(sums_d, sums)
    # data-parallel gradient exchange of the training side (unidet3d_b200/train.py): bucketed all-reduce + average
    from unidet3d_b200 import train
    ps = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (7,), (2, 2, 2), (1000,))]
    for i, p_ in enumerate(ps):
        p_.grad = torch.full(p_.shape, float((rank + 1) * (i + 1)))
    n_coll = train.allreduce_gradients(ps, bucket_bytes=4 * 40)        # small buckets: several collectives
    avg = [float(p_.grad.mean()) for p_ in ps]
    q.put((rank, mine, fingerprint, t_dev, t_e2e, count, mean.tolist(), var.tolist(), n_coll, avg))
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
    (r0, s0, f0, a0, b0, c0, m0, v0, nc0, g0), (r1, s1, f1, a1, b1, c1, m1, v1, nc1, g1) = out
    assert nc0 == nc1 == 2 and g0 == g1 == [1.5 * (i + 1) for i in range(4)]       # mean of ranks' (rank + 1) * (i + 1)
    xs = [torch.randn(300 + 200 * r, 8, generator=torch.Generator().manual_seed(100 + r), dtype=torch.float64) * (1 + r) + r
          for r in range(2)]
    allx = torch.cat(xs)
    assert c0 == c1 == 800.0
    assert torch.allclose(torch.tensor(m0, dtype=torch.float64), allx.mean(0), atol=1e-12) and m0 == m1
    assert torch.allclose(torch.tensor(v0, dtype=torch.float64), allx.var(0, unbiased=False), atol=1e-10) and v0 == v1
    assert sorted(s0 + s1) == list(range(10)) and not set(s0) & set(s1)       # disjoint cover
    assert f0 != f1                                                           # ranks work on different scenes
    assert a0 == a1 == 2.0 and b0 == b1 == 5.0                                # MAX over ranks, same on every rank
    from unidet3d_b200 import sharding
    assert sharding.whole_job_throughput(8, 2, 2.0) == 8.0
