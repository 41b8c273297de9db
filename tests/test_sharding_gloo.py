"""CPU, world_size 2 over gloo: the N>1 host logic — contiguous pair sharding,
per-pair seeding independent of the sharding, and the final metric reduction."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions():
    from stitch_b200.pipeline import shard_range
    for n in (0, 1, 7, 16, 64, 129):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_pair_inputs_independent_of_sharding():
    from stitch_b200.pipeline import make_pair_batch
    whole = make_pair_batch(0, 4, size=64, iters=1)
    part = make_pair_batch(2, 2, size=64, iters=1)
    for a, b in zip(whole.tensors()[:7], part.tensors()[:7]):
        assert torch.equal(a[2:4], b)
    assert torch.equal(whole.coords[:, :, 2:4], part.coords)


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stitch_b200.pipeline import make_pair_batch, shard_range
    lo, hi = shard_range(n_pairs, rank, world)
    pb = make_pair_batch(lo, hi - lo, size=32, iters=1, channels=8)
    # the only collective of the path: [sum of per-pair metric, pairs] summed, elapsed max-reduced
    metric = pb.image1.double().mean(dim=(1, 2, 3)).sum()
    red = torch.tensor([metric.item(), float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(red, op=dist.ReduceOp.SUM)
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((red.tolist(), t.item()))
    dist.destroy_process_group()


def test_two_rank_reduction_matches_single_process():
    from stitch_b200.pipeline import make_pair_batch
    n_pairs, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    (red, tmax) = q.get()
    whole = make_pair_batch(0, n_pairs, size=32, iters=1, channels=8)
    ref = whole.image1.double().mean(dim=(1, 2, 3)).sum().item()
    assert abs(red[0] - ref) < 1e-9 and red[1] == n_pairs
    assert abs(tmax - 0.2) < 1e-12
