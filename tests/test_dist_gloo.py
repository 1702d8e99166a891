"""CPU, world_size 2 over gloo: the data-parallel decomposition used by ssdn.train.train_step.

Each rank takes an equal shard of the global batch, computes the gradient of the MEAN loss over its shard, the
gradients are summed by one all-reduce over a flat buffer and scaled by 1/world (folded into the Adam kernel on the
GPU).  This test proves on the oracle that the result equals the single-process gradient / update on the global batch."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ssdn_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    n = 4
    _, noisy = O.synthetic_batch(n, 1, 32, seed=3)
    sigma = torch.full((n, 1, 1, 1), 25 / 255)
    tr = O.CpuTrainer("ssdn", "known", 1, seed=0)
    lo, hi = rank * n // world, (rank + 1) * n // world
    out = tr.loss(noisy[lo:hi], sigma[lo:hi])
    out["loss"].mean().backward()
    flat = torch.cat([t.grad.reshape(-1) for t in tr.leaves])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                 # the single collective of the step
    flat /= world
    losses = [torch.zeros(n // world, 1) for _ in range(world)]
    dist.all_gather(losses, out["loss"].detach())
    if rank == 0:
        q.put((flat, torch.cat(losses)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_equals_global_gradient():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ssdn_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, losses = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 4
    _, noisy = O.synthetic_batch(n, 1, 32, seed=3)
    sigma = torch.full((n, 1, 1, 1), 25 / 255)
    tr = O.CpuTrainer("ssdn", "known", 1, seed=0)
    out = tr.loss(noisy, sigma)
    out["loss"].mean().backward()
    ref = torch.cat([t.grad.reshape(-1) for t in tr.leaves])
    assert torch.allclose(losses, out["loss"].detach(), rtol=1e-5, atol=1e-6)
    assert ((flat - ref).abs().max() / ref.abs().max()).item() < 1e-4


# ---------------------------------------------------------------------------------------------- trainer-side sharding
def _shard_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
    from ssdn.datasets import FixedLengthSampler, SamplingOrder
    from ssdn.train import RankShardSampler
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # ranks deliberately seeded DIFFERENTLY: the order must still be rank 0's
    data = list(range(10))
    base = FixedLengthSampler(data, num_samples=24, shuffled=True)
    mine = list(RankShardSampler(base, rank, world))
    order = base.last_iter().order
    # resume: a restored global order with the cursor at 8 images seen (by all ranks together)
    base2 = FixedLengthSampler(data, num_samples=24, shuffled=True)
    base2.for_next_iter(SamplingOrder(list(order), 8))
    resumed = list(RankShardSampler(base2, rank, world))
    q.put((rank, mine, order, resumed))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_shard_sampler_partitions_one_global_order():
    """DenoiserTrainer.train_data under data parallelism: the ranks read disjoint, interleaved slices of ONE order (rank 0's),
    together exactly the reference's sequence of global mini-batches; after a resume they continue from the global cursor."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        rank, mine, order, resumed = q.get(timeout=300)
        got[rank] = (mine, order, resumed)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    order = got[0][1]
    assert got[1][1] == order and len(order) == 24                 # one order on both ranks although they seeded differently
    assert got[0][0] == order[0::2] and got[1][0] == order[1::2]   # interleaved: global batch b = positions [b*B, (b+1)*B) split over ranks
    assert got[0][2] == order[8::2] and got[1][2] == order[9::2]   # resumed from the global cursor
    inter = [x for pair in zip(got[0][0], got[1][0]) for x in pair]
    assert inter == order
