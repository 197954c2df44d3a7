"""End-to-end training steps (the reference's _pjit_train_step, trainer.py:337-390): p_losses value_and_grad, optax
Adam (b1 0.9, b2 0.999, eps 1e-8, piecewise-cosine learning rate), apply_updates and the EMA rule of :373-382,
through TrainStep on the GPU (eager and CUDA-graph replay) against the CPU oracle driven by torch.optim.Adam on the
same weights, clips, timesteps and noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_ours(p0, xs, ts_, ns, use_graph, lr, ema_start, ema_every, decay):
    from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion
    from video_diffusion_nnx_b200.trainer import TrainStep
    from video_diffusion_nnx_b200.unet3d import Unet3D

    net = Unet3D(dim=32, channels=1)
    net.load_state_dict({k: v.detach().numpy() for k, v in p0.items()})
    gd = GaussianDiffusion(net, image_size=64, num_frames=2, channels=1, timesteps=200, loss_type="l2")
    step = TrainStep(gd, batch_size=xs[0].shape[0], train_lr=lr, step_start_ema=ema_start, update_ema_every=ema_every,
                     ema_decay=decay, use_graph=use_graph)
    losses = []
    for i, (x, t, n) in enumerate(zip(xs, ts_, ns)):
        step.x.copy_(x)          # TrainStep normalises to [-1, 1] itself (gaussian_diffusion.py:492)
        step.t.copy_(t)
        step.noise.copy_(n)
        losses.append(float(step.step_device(i).item()))
    torch.cuda.synchronize()
    return losses, net.state_dict(), net.state_dict(flat=step.ema)


def test_three_adam_ema_steps_match_the_oracle():
    from oracle import diffusion_oracle as D
    from oracle import unet3d_oracle as U
    from video_diffusion_nnx_b200.trainer import piecewise_cosine_lr

    B, lr, ema_start, ema_every, decay, n_steps = 2, 1e-4, 1, 2, 0.9, 3
    rng = np.random.default_rng(7)
    p0 = U.init_params(32, 1, seed=3, perturb=0.05)
    xs = [torch.from_numpy(rng.random((B, 1, 2, 64, 64), dtype=np.float32)) for _ in range(n_steps)]
    ts_ = [torch.from_numpy(rng.integers(0, 200, (B,)).astype(np.int32)) for _ in range(n_steps)]
    ns = [torch.from_numpy(rng.standard_normal((B, 1, 2, 64, 64)).astype(np.float32)) for _ in range(n_steps)]

    # ---- oracle: fp32 autograd + torch Adam (same defaults as optax.adam) + the reference's EMA rule ----
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    gdo = D.GaussianDiffusionOracle(lambda xx, tt: U.unet3d_forward(p, xx, tt, 32), image_size=64, num_frames=2,
                                    channels=1, timesteps=200, loss_type="l2")
    opt = torch.optim.Adam(list(p.values()), lr=lr, betas=(0.9, 0.999), eps=1e-8)
    ema = {k: v.detach().clone() for k, v in p.items()}
    ref_losses = []
    for i in range(n_steps):
        for g in opt.param_groups:
            g["lr"] = piecewise_cosine_lr(i, lr, 0, 0, 1.0)
        opt.zero_grad()
        loss = gdo(xs[i], ts_[i], ns[i])
        loss.backward()
        for v in p.values():          # parameters the graph never touches (dead PreNorm / rel-pos bias): zero gradient
            if v.grad is None:
                v.grad = torch.zeros_like(v)
        opt.step()
        ref_losses.append(float(loss.item()))
        if i >= ema_start and i % ema_every == 0:   # trainer.py:373-382
            for k in ema:
                ema[k] = decay * ema[k] + (1 - decay) * p[k].detach()

    for use_graph in (False, True):
        losses, params, ema_ours = _run_ours(p0, xs, ts_, ns, use_graph, lr, ema_start, ema_every, decay)
        for a, b in zip(losses, ref_losses):
            assert abs(a - b) / b < 1.5e-2, (losses, ref_losses)   # bf16 tensor-core path vs fp32 oracle
        num = den = 0.0
        enum = eden = 0.0
        for k, v0 in p0.items():
            d_ref = (p[k].detach() - v0).double()
            d_got = torch.from_numpy(params[k]).double() - v0.double()
            num += (d_got - d_ref).pow(2).sum().item()
            den += d_ref.pow(2).sum().item()
            e_ref = (ema[k] - v0).double()
            e_got = torch.from_numpy(ema_ours[k]).double() - v0.double()
            enum += (e_got - e_ref).pow(2).sum().item()
            eden += e_ref.pow(2).sum().item()
        rel, erel = (num / den) ** 0.5, (enum / max(eden, 1e-30)) ** 0.5
        print(f"graph={use_graph}: losses {losses} vs {ref_losses}; update rel-L2 {rel:.3e}, EMA-update rel-L2 {erel:.3e}")
        # Adam normalises each gradient element, so elements whose gradient is bf16 noise move in a random direction:
        # the bound is on the whole update vector (measured 8-10 %). The optimizer arithmetic itself is checked to
        # 1e-3 of a step against torch's Adam on identical gradients in test_bench_path_parity_gpu.py.
        assert rel < 0.15 and erel < 0.15


def test_device_prefetcher_keeps_order_and_values():
    from video_diffusion_nnx_b200.data import DevicePrefetcher

    g = torch.Generator().manual_seed(3)
    host = [torch.rand(2, 1, 2, 64, 64, generator=g) for _ in range(7)]
    got = [b.clone() for b in DevicePrefetcher(iter(host), depth=3)]
    torch.cuda.synchronize()
    assert len(got) == 7 and all(b.is_cuda for b in got)
    assert all(torch.equal(b.cpu(), h) for b, h in zip(got, host))
