"""GPU parity of the host-visible diffusion helpers that mirror the reference class (gaussian_diffusion.py:101-228):
q_mean_variance, q_posterior, predict_start_from_noise, p_mean_variance (static and dynamic thresholding) against the
CPU oracle on the same weights / inputs, and their consistency with the fused p_sample kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(T=200, Fr=2, S=64, dyn=False):
    from oracle import diffusion_oracle as D
    from oracle import unet3d_oracle as U
    from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion
    from video_diffusion_nnx_b200.unet3d import Unet3D

    p = U.init_params(32, 1, seed=3, perturb=0.05)
    net = Unet3D(dim=32, channels=1)
    net.load_state_dict({k: v.numpy() for k, v in p.items()})
    net.train(False)
    gd = GaussianDiffusion(net, image_size=S, num_frames=Fr, channels=1, timesteps=T, loss_type="l2",
                           use_dynamic_thres=dyn, dynamic_thres_percentile=0.9)
    with torch.no_grad():
        gdo = D.GaussianDiffusionOracle(lambda xx, tt: U.unet3d_forward(p, xx, tt, 32), image_size=S, num_frames=Fr,
                                        channels=1, timesteps=T, loss_type="l2", use_dynamic_thres=dyn,
                                        dynamic_thres_percentile=0.9)
    return gd, gdo


def test_forward_posterior_helpers_match_oracle():
    gd, gdo = _pair()
    rng = np.random.default_rng(2)
    x0 = torch.from_numpy(rng.uniform(-1, 1, (3, 1, 2, 64, 64)).astype(np.float32))
    xt = torch.from_numpy(rng.standard_normal((3, 1, 2, 64, 64)).astype(np.float32))
    t = torch.tensor([0, 57, 199], dtype=torch.int32)
    for got, want in zip(gd.q_mean_variance(x0, t), gdo.q_mean_variance(x0, t)):
        assert torch.allclose(got.cpu(), want.expand_as(got.cpu()) if want.numel() != got.numel() else want, rtol=1e-6, atol=1e-7)
    for got, want in zip(gd.q_posterior(x0, xt, t), gdo.q_posterior(x0, xt, t)):
        assert torch.allclose(got.cpu(), want, rtol=1e-6, atol=1e-7)
    got = gd.predict_start_from_noise(xt, t, x0).cpu()
    assert torch.allclose(got, gdo.predict_start_from_noise(xt, t, x0), rtol=1e-6, atol=1e-6)
    # reference known answer (gaussian_diffusion_test.py:88-109): at t = 0 the mean is sqrt(alphas_cumprod[0]) * x0
    m0, v0, _ = gd.q_mean_variance(x0, torch.zeros(3, dtype=torch.int32))
    assert torch.allclose(m0.cpu(), x0 * float(gd.sqrt_alphas_cumprod[0]), atol=1e-6)
    assert torch.allclose(v0.cpu().flatten(), torch.full((3,), 1.0 - float(gd.alphas_cumprod[0])), atol=1e-7)


@pytest.mark.parametrize("dyn", [False, True])
def test_p_mean_variance_and_p_sample(dyn):
    gd, gdo = _pair(dyn=dyn)
    rng = np.random.default_rng(3)
    x = torch.from_numpy((2.0 * rng.standard_normal((2, 1, 2, 64, 64))).astype(np.float32))
    z = torch.from_numpy(rng.standard_normal((2, 1, 2, 64, 64)).astype(np.float32))
    t = torch.tensor([0, 120], dtype=torch.int32)
    mean, var, logvar = gd.p_mean_variance(x, t, clip_denoised=True)
    with torch.no_grad():
        mean_o, var_o, logvar_o = gdo.p_mean_variance(x, t, clip_denoised=True)
        samp_o = gdo.p_sample(x, t, z)
    scale = mean_o.abs().max().item()
    assert (mean.cpu() - mean_o).abs().max().item() < 2e-2 * scale   # bf16 Unet vs fp32 oracle
    assert torch.allclose(var.cpu(), var_o, rtol=1e-6) and torch.allclose(logvar.cpu(), logvar_o, rtol=1e-6)
    samp = gd.p_sample(x, t, z=z.cuda())
    assert (samp.cpu() - samp_o).abs().max().item() < 2e-2 * samp_o.abs().max().item()
    # the fused kernel with z = 0 IS the posterior mean of p_mean_variance (reference test :175-189 at t = 0), up to
    # the run-to-run noise of two Unet forwards (GroupNorm partial sums are accumulated with fp32 atomics)
    if not dyn:
        mean_k = gd.p_sample(x, t, z=torch.zeros_like(z).cuda())
        assert torch.allclose(mean_k, mean, atol=2e-3 * max(1.0, scale))
    assert torch.allclose(samp[0].cpu(), mean[0].cpu(), atol=2e-3 * max(1.0, scale))  # t = 0: no noise added


def test_dynamic_threshold_sampling_loop_runs_and_shards():
    gd, _ = _pair(T=6, dyn=True)
    full = gd.p_sample_loop((2,), 9)
    shard = gd.p_sample_loop((1,), 9, sample_offset=1)
    assert torch.isfinite(full).all() and full.shape == (2, 1, 2, 64, 64)
    assert torch.allclose(shard[0], full[1], atol=2e-2)
