"""The data-parallel training step of the reference (trainer.py:322-390) on the sm_100a kernels.

Semantics kept: loss = GaussianDiffusion.__call__ (random t, normalise, q_sample, Unet3D, L1/L2),
gradients of every Unet parameter, optax.adam defaults (b1 .9, b2 .999, eps 1e-8, bias corrected)
with the piecewise-cosine learning-rate schedule (trainer.py:138-147), EMA
`decay*ema + (1-decay)*p` iff step >= step_start_ema and step % update_ema_every == 0
(trainer.py:373-382); no gradient accumulation (the reference implements none). `max_grad_norm`
(accepted and ignored by the reference trainer, trainer.py:156) is honoured when given: the
global-norm clip of utils.py:127-152 folded into the fused Adam kernel.

Data parallelism: one process per GPU; each rank holds a full replica and a batch shard
(trainer.py:307-309 shards the batch over the `data` mesh axis). The gradient exchange that GSPMD
inserts implicitly in the reference is explicit here and goes through the C ABI
(vdn_comm_init / vdn_allreduce_bucket, include/vdn.h): the flat fp32 gradient is sum-reduced in a
few contiguous buckets on a communication stream, each bucket enqueued right after the backward
stages that complete it, overlapping the rest of backward. The collectives are CAPTURED in the
step's CUDA graph: forward + backward + bucket reductions + optimizer replay as ONE graph launch.
t and the noise are drawn for the GLOBAL batch from one key and sliced per rank (the reference's
pjit step draws them once and shards them), so a DP run consumes the same draws as one process
on the whole batch.

Stated difference: the reference also differentiates and Adam-updates the ten
schedule tables; they are constants here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import ops
from ._lib import VDN_BF16, VDN_F32, check, host_flag, lib
from .gaussian_diffusion import GaussianDiffusion, as_key, randint_from_key


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
    TrainStep issues the same reductions per bucket through the C ABI and folds the 1/world into Adam."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    for _, (lo, hi) in buckets:
        dist.all_reduce(flat_grad[lo:hi], group=group)
    flat_grad.mul_(1.0 / world)
    return flat_grad


class Communicator:
    """One NCCL communicator per process behind the C ABI (vdn_comm_*; reference: the collective GSPMD inserts for
    the pjit of trainer.py:322-326). The unique id travels out of band; `from_process_group` uses a
    torch.distributed group (any backend) for that one broadcast."""

    def __init__(self, rank: int, world: int, unique_id: bytes, max_ctas: int = 0):
        self.rank, self.world = rank, world
        h = C.c_void_p()
        check(lib.vdn_comm_init(C.byref(h), unique_id, rank, world, max_ctas), "vdn_comm_init")
        self._h = h

    @staticmethod
    def new_unique_id() -> bytes:
        buf = C.create_string_buffer(lib.vdn_comm_unique_id_bytes())
        check(lib.vdn_comm_unique_id(buf), "vdn_comm_unique_id")
        return buf.raw

    @classmethod
    def from_process_group(cls, group=None, max_ctas: int = 0) -> "Communicator":
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.new_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(rank, world, box[0], max_ctas)

    def all_reduce_sum_(self, t: torch.Tensor) -> None:
        """In-place sum over the ranks, enqueued on the current stream (graph capturable)."""
        assert t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16)
        check(lib.vdn_allreduce_bucket(self._h, t.data_ptr(), t.numel(), VDN_F32 if t.dtype == torch.float32 else VDN_BF16,
                                       torch.cuda.current_stream().cuda_stream), "vdn_allreduce_bucket")

    def destroy(self) -> None:
        if self._h is not None:
            lib.vdn_comm_destroy(self._h)
            self._h = None


class _PinnedRing:
    """Ring of pinned host staging buffers for small per-step host -> device copies. The copy of step i is ordered
    behind step i-1's graph and the host runs ahead, so ONE staging buffer would be overwritten before its DMA ran;
    each slot is guarded by an event recorded after its copy (a wait only happens if the host is `slots` steps ahead)."""

    def __init__(self, shape, dtype, slots: int = 16):
        self.bufs = [torch.zeros(shape, dtype=dtype).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.i = 0

    def push(self, dst: torch.Tensor, fill) -> None:
        i = self.i
        self.i = (i + 1) % len(self.bufs)
        if self.events[i] is not None:
            self.events[i].synchronize()
        fill(self.bufs[i])
        dst.copy_(self.bufs[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


class TrainStep:
    def __init__(self, diffusion: GaussianDiffusion, *, batch_size: int, train_lr: float = 1e-4,
                 lr_decay_start_step: int = 0, lr_decay_steps: int = 0, lr_decay_coeff: float = 1.0,
                 step_start_ema: int = 2000, update_ema_every: int = 10, ema_decay: float = 0.9999,
                 max_grad_norm: Optional[float] = None, use_graph: bool = True, process_group=None,
                 comm: Optional[Communicator] = None, comm_max_ctas: int = 4, bucket_bytes: int = 4 << 20):
        """batch_size is the PER-RANK batch (the shard of trainer.py:307-309). `process_group`: a torch.distributed
        group used only to exchange the NCCL unique id; `comm`: an existing Communicator instead."""
        self.gd = diffusion
        net = diffusion.denoise_fn
        self.net = net
        self.B = batch_size
        self.lr_args = (train_lr, lr_decay_start_step, lr_decay_steps, lr_decay_coeff)
        self.step_start_ema, self.update_ema_every, self.ema_decay = step_start_ema, update_ema_every, ema_decay
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.use_graph = use_graph
        if comm is None and process_group is not None and torch.distributed.get_world_size(process_group) > 1:
            comm = Communicator.from_process_group(process_group, comm_max_ctas)
        self.comm = comm
        self.world = comm.world if comm is not None else 1
        self.rank = comm.rank if comm is not None else 0
        net.train(True)
        self.eng = net.engine(batch_size, diffusion.num_frames, diffusion.image_size, diffusion.image_size, training=True)
        st = net.store
        dev = st.flat.device
        self.m = torch.zeros_like(st.flat)
        self.v = torch.zeros_like(st.flat)
        self.ema = st.flat.clone()
        self.hp = torch.zeros(16, dtype=torch.float32, device=dev)
        self._hp_ring = _PinnedRing((16,), torch.float32)
        self._t_ring = _PinnedRing((batch_size,), torch.int32)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        shape = (batch_size, diffusion.channels, diffusion.num_frames, diffusion.image_size, diffusion.image_size)
        self.x = torch.empty(shape, dtype=torch.float32, device=dev)
        self.noise = torch.empty(shape, dtype=torch.float32, device=dev)
        self.x_noisy = torch.empty(shape, dtype=torch.float32, device=dev)
        self.t = torch.empty(batch_size, dtype=torch.int32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.dpred = torch.empty((batch_size, diffusion.num_frames, diffusion.image_size, diffusion.image_size,
                                  diffusion.channels), dtype=torch.float32, device=dev)
        self.count = 0
        self._graph = None
        self._segments = self._plan_segments(bucket_bytes)
        self.comm_stream = torch.cuda.Stream(priority=-1) if self.world > 1 else None  # reductions start as soon as issued
        self.launches_per_step = None
        self.graph_launches_per_step = None

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
        B, C_ = self.x.shape[0], self.x.shape[1]
        ops.loss_fwd_bwd(pred, self.noise, self.loss, self.dpred, B, C_, self.x.numel() // (B * C_),
                         gd.loss_type == "l1")
        self.net.store.grad.zero_()
        self.eng._bw_state["dout"] = self.dpred

    def _optimizer(self):
        st = self.net.store
        if self.max_grad_norm > 0.0:
            ops.grad_sqnorm(st.grad, self.sqnorm)
            ops.adam_ema(st.flat, st.grad, self.m, self.v, self.ema, self.hp, sqnorm=self.sqnorm)
        else:
            ops.adam_ema(st.flat, st.grad, self.m, self.v, self.ema, self.hp)
        self.eng.repack()

    def _reduce(self, lo, hi):
        """Sum-reduce grad[lo:hi] over the ranks on the communication stream, ordered after everything enqueued so far
        on the current stream (fork edge; under capture it becomes a branch of the step graph)."""
        if self.world == 1:
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self.comm.all_reduce_sum_(self.net.store.grad[lo:hi])

    def _wait_comm(self):
        if self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def _step_body(self):
        self._fwd_loss()
        for fns, (lo, hi) in self._segments:
            for f in fns:
                f()
            self._reduce(lo, hi)
        self._wait_comm()
        self._optimizer()

    def _capture(self):
        """The whole step - q_sample, forward, loss, backward, the bucket reductions on the forked communication
        stream, Adam + EMA + operand repack - as ONE CUDA graph."""
        torch.cuda.synchronize()
        n0 = lib.vdn_launch_count()
        # The critical path is captured on a HIGH-priority stream (kernel nodes inherit the priority of the stream
        # they were captured on), the engine's side streams keep the default (lowest) priority: when both have
        # blocks ready, the dependency chain goes first and the weight-gradient GEMMs fill what is left.
        cap = torch.cuda.Stream(priority=-1) if host_flag("VDN_NO_PRIORITY") is None else None
        kw = {"stream": cap} if cap is not None else {}
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, **kw):
            self._step_body()
        self._graph = g
        self.launches_per_step = int(lib.vdn_launch_count() - n0)  # kernels of this library in one replayed step
        self.graph_launches_per_step = 1

    # -- public ------------------------------------------------------------------------------
    def loss_and_grad(self, x, t, noise) -> torch.Tensor:
        """Loss of `GaussianDiffusion.__call__` for explicit (t, noise) and its gradient wrt every Unet
        parameter in net.store.grad (no optimizer update, no exchange). Used by the parity tests."""
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
        vals = torch.tensor([lr, 0.9, 0.999, 1e-8, 1.0 - 0.9 ** c, 1.0 - 0.999 ** c, self.ema_decay, do_ema,
                             1.0 / self.world, self.max_grad_norm, 1e-6, 0, 0, 0, 0, 0], dtype=torch.float32)
        self._hp_ring.push(self.hp, lambda h: h.copy_(vals))

    def step_device(self, step: int) -> torch.Tensor:
        """One optimisation step on the batch already staged in self.x / self.t / self.noise."""
        self.set_hyper(step)
        if self.use_graph:
            if self._graph is None:
                # warm-up pass (allocates every pooled buffer, sets kernel attributes, opens the NCCL channels) on a
                # snapshot of the training state: the first call must apply ONE optimizer update, like every other call
                st = self.net.store
                state = (st.flat, self.m, self.v, self.ema)
                snap = [t.clone() for t in state]
                self._step_body()
                torch.cuda.synchronize()
                for dst, src in zip(state, snap):
                    dst.copy_(src)
                self.eng.repack()
                self._capture()
            self._graph.replay()
        else:
            self._step_body()
        self.count += 1
        # the master weights changed; this engine's operands were repacked inside the step, every other engine of the
        # same Unet3D (inference engines, cached samplers) repacks lazily before its next forward
        self.net.store.bump()
        self.eng.packed_version = self.net.store.version
        return self.loss

    def draw(self, key):
        """t and the noise of this rank's shard, drawn for the GLOBAL batch from `key` (gaussian_diffusion.py:493-496,
        :441-445) and sliced: a DP run consumes the same draws as a single process on the whole batch."""
        key = as_key(key)
        _, t_key, loss_key = key.split(3)
        _, noise_key, _ = loss_key.split(3)
        t_all = randint_from_key(t_key, self.gd.num_timesteps, self.world * self.B)
        lo, hi = shard_range(self.world * self.B, self.world, self.rank)
        self._t_ring.push(self.t, lambda h: h.copy_(t_all[lo:hi]))
        ops.randn(self.noise, noise_key.seed, noise_key.stream, elem_offset=lo * (self.noise.numel() // self.B))

    def step(self, x_host_or_dev: torch.Tensor, key, step: int) -> torch.Tensor:
        """trainer.py:560-572: one training step on this rank's batch shard. `key` seeds t and the noise."""
        self.x.copy_(x_host_or_dev, non_blocking=True)
        self.draw(key)
        return self.step_device(step)
