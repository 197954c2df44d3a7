"""Numerical check of the data-parallel training step on 2+ GPUs (reference: trainer.py:163,307-326 - batch sharded
over the 'data' axis, gradients averaged by the all-reduce GSPMD inserts):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp.py
Every rank runs TrainStep on its shard of a fixed global batch (bucketed NCCL all-reduce overlapped with backward,
1/world folded into Adam); rank 0 then repeats the same steps alone on the whole batch. Checks: replicas stay
bit-identical, the mean of the shard losses equals the full-batch loss, and the parameter update of the sharded run
equals the full-batch update."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion  # noqa: E402
from video_diffusion_nnx_b200.trainer import TrainStep, shard_range  # noqa: E402
from video_diffusion_nnx_b200.unet3d import Unet3D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
b, Fr, S, T, n_steps = 2, 4, 64, 1000, 3
G = b * world
rng = np.random.default_rng(5)
xs = [torch.from_numpy(rng.random((G, 1, Fr, S, S), dtype=np.float32)) for _ in range(n_steps)]
ts_ = [torch.from_numpy(rng.integers(0, T, (G,)).astype(np.int32)) for _ in range(n_steps)]
ns = [torch.from_numpy(rng.standard_normal((G, 1, Fr, S, S)).astype(np.float32)) for _ in range(n_steps)]


def run(batch, pg, lo, hi, graph):
    net = Unet3D(dim=32, channels=1, rngs=0)
    init = net.state_dict()
    gd = GaussianDiffusion(net, image_size=S, num_frames=Fr, channels=1, timesteps=T, loss_type="l2")
    step = TrainStep(gd, batch_size=batch, step_start_ema=0, update_ema_every=1, ema_decay=0.9, use_graph=graph,
                     process_group=pg, bucket_bytes=4 << 20)
    losses = []
    for i in range(n_steps):
        step.x.copy_(xs[i][lo:hi])
        step.t.copy_(ts_[i][lo:hi])
        step.noise.copy_(ns[i][lo:hi])
        losses.append(float(step.step_device(i).item()))
    torch.cuda.synchronize()
    return init, losses, net.store.flat.clone()


def draws_match():
    """TrainStep.step under DP draws t / noise for the GLOBAL batch and slices its shard (the reference's pjit step
    draws once and shards): the gathered shard draws must equal a single process's draws on the whole batch, bit for bit."""
    lo, hi = shard_range(G, world, rank)
    net = Unet3D(dim=32, channels=1, rngs=0)
    gd = GaussianDiffusion(net, image_size=S, num_frames=Fr, channels=1, timesteps=T, loss_type="l2")
    step = TrainStep(gd, batch_size=b, use_graph=False, process_group=dist.group.WORLD)
    step.draw(1234)
    torch.cuda.synchronize()
    t_all = [torch.empty_like(step.t) for _ in range(world)]
    n_all = [torch.empty_like(step.noise) for _ in range(world)]
    dist.all_gather(t_all, step.t)
    dist.all_gather(n_all, step.noise)
    good = True
    if rank == 0:
        net1 = Unet3D(dim=32, channels=1, rngs=0)
        gd1 = GaussianDiffusion(net1, image_size=S, num_frames=Fr, channels=1, timesteps=T, loss_type="l2")
        full = TrainStep(gd1, batch_size=G, use_graph=False)
        full.draw(1234)
        torch.cuda.synchronize()
        good = torch.equal(torch.cat(t_all), full.t) and torch.equal(torch.cat(n_all), full.noise)
        print(f"DP draws equal the full-batch draws: {good}; shard timesteps differ across ranks: "
              f"{not torch.equal(t_all[0], t_all[-1])}", flush=True)
    return good


ok = draws_match()
for graph in (False, True):
    lo, hi = shard_range(G, world, rank)
    init, losses, flat = run(b, dist.group.WORLD, lo, hi, graph)
    all_flat = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(all_flat, flat)
    lt = torch.tensor(losses, device="cuda", dtype=torch.float64)
    dist.all_reduce(lt)
    mean_losses = (lt / world).tolist()
    if rank == 0:
        same = all(torch.equal(all_flat[0], f) for f in all_flat[1:])
        _, full_losses, full_flat = run(G, None, 0, G, graph)
        net0 = Unet3D(dim=32, channels=1, rngs=0)
        net0._ensure_store(False)
        w0 = net0.store.flat
        d_dp, d_full = (flat - w0).double(), (full_flat - w0).double()
        rel = float((d_dp - d_full).norm() / d_full.norm())
        lerr = max(abs(a - c) / c for a, c in zip(mean_losses, full_losses))
        print(f"graph={graph}: replicas identical {same}; mean shard loss {mean_losses} vs full batch {full_losses} "
              f"(max rel {lerr:.2e}); update rel-L2 sharded vs full batch {rel:.3e}", flush=True)
        ok = ok and same and lerr < 5e-3 and rel < 0.1
    dist.barrier()
if rank == 0:
    print("DP CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.destroy_process_group()
