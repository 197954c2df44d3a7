"""The data-parallel training step of the reference (trainer.py:322-390) on the sm_100a kernels.

Semantics kept: loss = GaussianDiffusion.__call__ (random t, normalise, q_sample, Unet3D, L1/L2),
gradients of every Unet parameter, optax.adam defaults (b1 .9, b2 .999, eps 1e-8, bias corrected)
with the piecewise-cosine learning-rate schedule (trainer.py:138-147), EMA
`decay*ema + (1-decay)*p` iff step >= step_start_ema and step % update_ema_every == 0
(trainer.py:373-382); no gradient clipping, no accumulation (the reference implements neither).

Data parallelism: one process per GPU; each rank holds a full replica and a batch shard
(trainer.py:307-309 shards the batch over the `data` mesh axis). The gradient exchange that GSPMD
inserts implicitly in the reference is explicit here: the flat fp32 gradient is reduced in a few
contiguous slices with NCCL (torch.distributed) on a communication stream, each slice launched as
soon as the backward stage that completes it has been enqueued, overlapping the rest of backward.

Stated difference: the reference also differentiates and Adam-updates the ten
schedule tables; they are constants here.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from . import ops
from .gaussian_diffusion import GaussianDiffusion, Key, as_key


def piecewise_cosine_lr(step: int, init_value: float, decay_start: int, decay_steps: int, coeff: float) -> float:
    """optax.piecewise_interpolate_schedule('cosine', init, {start: 1.0, start+steps: coeff})."""
    b0, b1 = decay_start, decay_start + decay_steps
    if b1 == b0:  # the dict literal collapses to {b0: coeff}
        return init_value if step < b0 else init_value * coeff
    v0, v1 = init_value, init_value * coeff
    if step <= b0:
        return v0
    if step >= b1:
        return v1
    pct = (step - b0) / (b1 - b0)
    return v1 + (v0 - v1) / 2.0 * (math.cos(math.pi * pct) + 1.0)


def plan_buckets(stage_order, slices, bucket_elems: int):
    """Group consecutive backward stages into gradient buckets. stage_order: stage names in backward
    order; slices: {name: (begin, end)} element ranges of the flat gradient each stage completes.
    Returns [(stage_names, (lo, hi))]: a bucket closes once it holds >= bucket_elems elements (and at the
    last stage). Buckets are contiguous because the flat layout mirrors the completion order."""
    out, cur, lo, hi = [], [], None, None
    for i, name in enumerate(stage_order):
        b, e = slices[name]
        cur.append(name)
        lo = b if lo is None else min(lo, b)
        hi = e if hi is None else max(hi, e)
        if (hi - lo) >= bucket_elems or i == len(stage_order) - 1:
            out.append((cur, (lo, hi)))
            cur, lo, hi = [], None, None
    return out


def shard_range(global_batch: int, world: int, rank: int):
    """Batch shard of `rank` (trainer.py:163 requires divisibility; gaussian_diffusion.py:281)."""
    assert global_batch % world == 0, "batch_size must be divisible by number of devices"
    per = global_batch // world
    return rank * per, (rank + 1) * per


def allreduce_mean_(flat_grad, buckets, group=None, async_streams=None):
    """Sum-reduce each bucket slice of the flat gradient over the data-parallel group and scale by
    1/world (the GSPMD-implicit mean of trainer.py:363-364 made explicit). Used by the CPU gloo tests;
    TrainStep issues the same reductions per bucket on a side stream and folds the 1/world into Adam."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    for _, (lo, hi) in buckets:
        dist.all_reduce(flat_grad[lo:hi], group=group)
    flat_grad.mul_(1.0 / world)
    return flat_grad


class TrainStep:
    def __init__(self, diffusion: GaussianDiffusion, *, batch_size: int, train_lr: float = 1e-4,
                 lr_decay_start_step: int = 0, lr_decay_steps: int = 0, lr_decay_coeff: float = 1.0,
                 step_start_ema: int = 2000, update_ema_every: int = 10, ema_decay: float = 0.9999,
                 use_graph: bool = True, process_group=None, bucket_bytes: int = 8 << 20):
        self.gd = diffusion
        net = diffusion.denoise_fn
        self.net = net
        self.B = batch_size
        self.lr_args = (train_lr, lr_decay_start_step, lr_decay_steps, lr_decay_coeff)
        self.step_start_ema, self.update_ema_every, self.ema_decay = step_start_ema, update_ema_every, ema_decay
        self.use_graph = use_graph
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        net.train(True)
        self.eng = net.engine(batch_size, diffusion.num_frames, diffusion.image_size, diffusion.image_size, training=True)
        st = net.store
        dev = st.flat.device
        self.m = torch.zeros_like(st.flat)
        self.v = torch.zeros_like(st.flat)
        self.ema = st.flat.clone()
        self.hp = torch.zeros(16, dtype=torch.float32, device=dev)
        self.hp_host = torch.zeros(16, dtype=torch.float32).pin_memory()
        shape = (batch_size, diffusion.channels, diffusion.num_frames, diffusion.image_size, diffusion.image_size)
        self.x = torch.empty(shape, dtype=torch.float32, device=dev)
        self.noise = torch.empty(shape, dtype=torch.float32, device=dev)
        self.x_noisy = torch.empty(shape, dtype=torch.float32, device=dev)
        self.t = torch.empty(batch_size, dtype=torch.int32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.dpred = torch.empty((batch_size, diffusion.num_frames, diffusion.image_size, diffusion.image_size,
                                  diffusion.channels), dtype=torch.float32, device=dev)
        self.count = 0
        self._graphs = None
        self._segments = self._plan_segments(bucket_bytes)
        self.comm_stream = torch.cuda.Stream(priority=-1) if self.world > 1 else None  # reductions start as soon as issued
        self.launches_per_step = None

    # -- plan: group backward stages into segments, one gradient bucket per segment ----------
    def _plan_segments(self, bucket_bytes: int):
        stages = self.eng.backward_stages()
        self.eng._stages = stages
        fn_of = dict(stages)
        plan = plan_buckets([n for n, _ in stages], self.eng.grad_slices(), bucket_bytes // 4)
        segs = [([fn_of[n] for n in names], rng) for names, rng in plan]
        if self.world == 1:  # no exchange: one segment
            fns = [f for s, _ in segs for f in s]
            segs = [(fns, (0, self.net.store.total))]
        return segs

    # -- pieces ------------------------------------------------------------------------------
    def _fwd_loss(self):
        gd, eng = self.gd, self.eng
        ops.q_sample(self.x, self.noise, self.t, gd.table("sqrt_alphas_cumprod"),
                     gd.table("sqrt_one_minus_alphas_cumprod"), self.x_noisy, True)
        pred = eng.forward(self.x_noisy, self.t)
        B, C = self.x.shape[0], self.x.shape[1]
        ops.loss_fwd_bwd(pred, self.noise, self.loss, self.dpred, B, C, self.x.numel() // (B * C),
                         gd.loss_type == "l1")
        self.net.store.grad.zero_()
        self.eng._bw_state["dout"] = self.dpred

    def _optimizer(self):
        st = self.net.store
        ops.adam_ema(st.flat, st.grad, self.m, self.v, self.ema, self.hp)
        self.eng.repack()

    def _run_eager(self):
        self._fwd_loss()
        for i, (fns, (lo, hi)) in enumerate(self._segments):
            for f in fns:
                f()
            self._reduce(lo, hi)
        self._wait_comm()
        self._optimizer()

    def _reduce(self, lo, hi):
        if self.world == 1:
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            torch.distributed.all_reduce(self.net.store.grad[lo:hi], group=self.pg)

    def _wait_comm(self):
        if self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def _capture(self):
        """Capture [q_sample + fwd + loss + bwd segment 0], [bwd segment k]..., [adam + repack] as CUDA
        graphs; NCCL reductions are launched between the graph launches."""
        torch.cuda.synchronize()
        from ._lib import lib as _vdn
        n0 = _vdn.vdn_launch_count()
        graphs = []
        # The critical path is captured on a HIGH-priority stream (kernel nodes inherit the priority of the stream
        # they were captured on), the engine's side streams keep the default (lowest) priority: when both have
        # blocks ready, the dependency chain goes first and the weight-gradient GEMMs fill what is left.
        import os
        cap = torch.cuda.Stream(priority=-1) if os.environ.get("VDN_NO_PRIORITY") is None else None
        kw = {"stream": cap} if cap is not None else {}
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, **kw):
            self._fwd_loss()
            for f in self._segments[0][0]:
                f()
        graphs.append(g)
        for fns, _ in self._segments[1:]:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, **kw):
                for f in fns:
                    f()
            graphs.append(g)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, **kw):
            self._optimizer()
        graphs.append(g)
        self._graphs = graphs
        self.launches_per_step = int(_vdn.vdn_launch_count() - n0)  # kernels of this library in one replayed step

    def _run_graphs(self):
        for g, (_, (lo, hi)) in zip(self._graphs[:-1], self._segments):
            g.replay()
            self._reduce(lo, hi)
        self._wait_comm()
        self._graphs[-1].replay()

    # -- public ------------------------------------------------------------------------------
    def loss_and_grad(self, x, t, noise) -> torch.Tensor:
        """Loss of `GaussianDiffusion.__call__` for explicit (t, noise) and its gradient wrt every Unet
        parameter in net.store.grad (no optimizer update). Used by the parity tests."""
        self.x.copy_(x)
        self.t.copy_(t)
        self.noise.copy_(noise)
        self._fwd_loss()
        for fns, _ in self._segments:
            for f in fns:
                f()
        return self.loss

    def set_hyper(self, step: int) -> None:
        lr = piecewise_cosine_lr(step, *self.lr_args)
        c = self.count + 1
        do_ema = 1.0 if (step >= self.step_start_ema and step % self.update_ema_every == 0) else 0.0
        h = self.hp_host
        h[0], h[1], h[2], h[3] = lr, 0.9, 0.999, 1e-8
        h[4], h[5] = 1.0 - 0.9 ** c, 1.0 - 0.999 ** c
        h[6], h[7], h[8] = self.ema_decay, do_ema, 1.0 / self.world
        self.hp.copy_(h, non_blocking=True)

    def step_device(self, step: int) -> torch.Tensor:
        """One optimisation step on the batch already staged in self.x / self.t / self.noise."""
        self.set_hyper(step)
        if self.use_graph:
            if self._graphs is None:
                # warm-up pass (allocates every pooled buffer, sets kernel attributes) on a snapshot of the training
                # state: the first call must apply ONE optimizer update, like every other call
                st = self.net.store
                state = (st.flat, self.m, self.v, self.ema)
                snap = [t.clone() for t in state]
                self._run_eager()
                torch.cuda.synchronize()
                for dst, src in zip(state, snap):
                    dst.copy_(src)
                self.eng.repack()
                self._capture()
            self._run_graphs()
        else:
            self._run_eager()
        self.count += 1
        return self.loss

    def step(self, x_host_or_dev: torch.Tensor, key, step: int) -> torch.Tensor:
        """trainer.py:560-572: one training step on this rank's batch shard. `key` seeds t and the noise."""
        key = as_key(key)
        _, t_key, loss_key = key.split(3)
        _, noise_key, _ = loss_key.split(3)
        self.x.copy_(x_host_or_dev, non_blocking=True)
        g = torch.Generator(device="cpu").manual_seed((t_key.seed * 7919 + t_key.stream) % (2 ** 63))
        self.t.copy_(torch.randint(0, self.gd.num_timesteps, (self.B,), generator=g, dtype=torch.int32))
        ops.randn(self.noise, noise_key.seed, noise_key.stream)
        return self.step_device(step)
