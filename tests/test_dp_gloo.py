"""world_size-2 gloo (CPU) test of the data-parallel host logic: bucketed gradient all-reduce equals
the mean over ranks, batch shards tile the global batch, and per-sample Philox offsets make sampling
independent of the sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video_diffusion_nnx_b200.trainer import allreduce_mean_, plan_buckets, shard_range

    n = 2700
    slices = {"late": (0, 100), "downs.0": (100, 300), "downs.1": (300, 700), "mid": (700, 1500),
              "ups.0": (1500, 2500), "ups.1": (2500, 2600), "final": (2600, 2700)}
    order = ["final", "ups.1", "ups.0", "mid", "downs.1", "downs.0", "late"]
    buckets = plan_buckets(order, slices, bucket_elems=900)
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(n, generator=g)
    mine = grad.clone()
    allreduce_mean_(grad, buckets)
    gathered = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = torch.stack(gathered).mean(0)
    ok = torch.allclose(grad, want, atol=1e-6)
    lo, hi = shard_range(8, world, rank)
    # timesteps are drawn for the GLOBAL batch from the key and sliced per rank (trainer.TrainStep.draw): the shards
    # tile a single process's draw, and different ranks get different timesteps
    from video_diffusion_nnx_b200.gaussian_diffusion import Key, randint_from_key

    t_all = randint_from_key(Key(77, 3), 1000, 8)
    mine_t = t_all[lo:hi].clone()
    got_t = [torch.zeros(hi - lo, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(got_t, mine_t)
    ok = ok and torch.equal(torch.cat(got_t), t_all) and not torch.equal(got_t[0], got_t[1])
    # the NCCL unique id of the C-ABI communicator travels over any torch.distributed backend
    from video_diffusion_nnx_b200.trainer import Communicator

    box = [Communicator.new_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ids = [None] * world
    dist.all_gather_object(ids, box[0])
    ok = ok and len(box[0]) == 128 and ids[0] == ids[1]
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_bucketed_allreduce_is_the_mean_over_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert [r for _, _, r in res] == [(0, 4), (4, 8)]
