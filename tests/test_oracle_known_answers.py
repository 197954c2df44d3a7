"""Pins the CPU oracle to every known answer the reference's own tests hold for this path
(SURVEY.md 8c): gaussian_diffusion_test.py:75-86 (table shapes), :88-109 (q_mean_variance@t=0),
:111-123 (x0 round trip, atol 1e-4), :135-158 (q_sample@t=0, atol 1e-6), :175-189 (p_sample@t=0 ==
posterior mean, atol 1e-5), :191-210 (loss 0 / 0.5 / 0.25 with a zero predictor); utils_test.py:102-110
(extract), :112-117 (beta range), :121-131 ((un)normalize). The reference tests omit the now-required
`key` argument (they are stale); the random draws are explicit inputs here. Same fixture as the
reference's setUp: image 8, frames 2, channels 3, T=10, batch 2, MockDenoiseFn returning zeros."""
import numpy as np
import torch

from oracle import diffusion_oracle as D


def _mock_denoise(x, t):  # gaussian_diffusion_test.py:18-33
    b, c, f, h, w = x.shape
    return torch.zeros(b, f, h, w, c, dtype=x.dtype)


def _setup(loss_type="l1"):
    gd = D.GaussianDiffusionOracle(_mock_denoise, image_size=8, num_frames=2, channels=3, timesteps=10,
                                   loss_type=loss_type)
    x_start = torch.ones(2, 3, 2, 8, 8)
    t = torch.tensor([0, 5], dtype=torch.int32)
    noise = torch.zeros_like(x_start)
    return gd, x_start, t, noise


def test_initialization_shapes():
    gd, *_ = _setup()
    assert gd.num_timesteps == 10
    for n in D.SCHEDULE_NAMES:
        assert tuple(getattr(gd, n).shape) == (10,)
        assert getattr(gd, n).dtype == torch.float32


def test_q_mean_variance_t0():
    gd, x_start, t, _ = _setup()
    mean, var, logvar = gd.q_mean_variance(x_start, t)
    assert mean.shape == x_start.shape and var.shape == (2, 1, 1, 1, 1) and logvar.shape == (2, 1, 1, 1, 1)
    t0 = torch.zeros(2, dtype=torch.int32)
    mean0, var0, _ = gd.q_mean_variance(x_start, t0)
    np.testing.assert_allclose(mean0, gd.sqrt_alphas_cumprod[0] * x_start, atol=1e-6)
    np.testing.assert_allclose(var0, torch.full((2, 1, 1, 1, 1), float(1.0 - gd.alphas_cumprod[0])), atol=1e-6)


def test_predict_start_from_noise_round_trip():
    gd, x_start, _, noise = _setup()
    t_mid = torch.full((2,), 5, dtype=torch.int32)
    x_t = gd.q_sample(x_start, t_mid, noise)
    np.testing.assert_allclose(gd.predict_start_from_noise(x_t, t_mid, noise), x_start, atol=1e-4)


def test_q_posterior_shapes():
    gd, x_start, _, noise = _setup()
    t_mid = torch.full((2,), 5, dtype=torch.int32)
    mean, var, logvar = gd.q_posterior(x_start, gd.q_sample(x_start, t_mid, noise), t_mid)
    assert mean.shape == x_start.shape and var.shape == (2, 1, 1, 1, 1) and logvar.shape == (2, 1, 1, 1, 1)


def test_q_sample_t0_closed_form():
    gd, x_start, _, _ = _setup()
    noise = torch.from_numpy(np.random.default_rng(42).standard_normal(tuple(x_start.shape)).astype(np.float32))
    t0 = torch.zeros(2, dtype=torch.int32)
    expected = gd.sqrt_alphas_cumprod[0] * x_start + gd.sqrt_one_minus_alphas_cumprod[0] * noise
    np.testing.assert_allclose(gd.q_sample(x_start, t0, noise), expected, atol=1e-6)


def test_p_sample_t0_is_posterior_mean():
    gd, x_start, _, _ = _setup()
    x_t = torch.zeros_like(x_start)
    t0 = torch.zeros(2, dtype=torch.int32)
    mean0, _, _ = gd.p_mean_variance(x_t, t0, clip_denoised=False)
    z = torch.randn(x_t.shape, generator=torch.Generator().manual_seed(0))
    np.testing.assert_allclose(gd.p_sample(x_t, t0, z), mean0, atol=1e-5)


def test_p_losses_known_values():
    gd, x_start, t, noise = _setup("l1")
    assert abs(gd.p_losses(x_start, t, noise).item() - 0.0) < 1e-6
    target = torch.ones_like(x_start) * 0.5
    assert abs(gd.p_losses(x_start, t, target).item() - 0.5) < 1e-6
    gd.loss_type = "l2"
    assert abs(gd.p_losses(x_start, t, target).item() - 0.25) < 1e-6


def test_extract_known_answer():
    out = D.extract(torch.arange(10), torch.tensor([1, 3, 5]), (3, 10, 10, 10))
    assert out.shape == (3, 1, 1, 1)
    assert out.flatten().tolist() == [1, 3, 5]


def test_cosine_beta_schedule_range():
    betas = D.cosine_beta_schedule(100)
    assert betas.shape == (100,) and betas.dtype == np.float32
    assert (betas >= 0).all() and (betas <= 1).all()


def test_normalize_unnormalize():
    np.testing.assert_allclose(D.unnormalize_img(torch.tensor([-1.0, 0.0, 1.0])), [0.0, 0.5, 1.0], atol=1e-6)
    np.testing.assert_allclose(D.normalize_img(torch.tensor([0.0, 0.5, 1.0])), [-1.0, 0.0, 1.0], atol=1e-6)


def test_product_schedule_is_bit_identical_to_oracle():
    """The product's host-side tables (no CUDA needed) match the oracle bit for bit."""
    from video_diffusion_nnx_b200.gaussian_diffusion import make_schedule

    for T in (10, 200, 1000):
        a, b = make_schedule(T), D.make_schedule(T)
        for k in D.SCHEDULE_NAMES:
            assert np.array_equal(a[k], b[k]), (T, k)


def test_dynamic_thresholding_matches_definition():
    """gaussian_diffusion.py:205-217 (Imagen dynamic thresholding, off by default): s = max(quantile(|x0|, p), 1) per
    sample with linear interpolation, x0 <- clip(x0, -s, s) / s; equals static clipping whenever the quantile <= 1."""
    import numpy as np

    from oracle import diffusion_oracle as D

    rng = np.random.default_rng(5)
    B, C, Fr, S, T = 3, 1, 2, 8, 50
    eps = torch.from_numpy(rng.standard_normal((B, Fr, S, S, C)).astype(np.float32))
    mk = lambda dyn: D.GaussianDiffusionOracle(lambda x, t: eps, image_size=S, num_frames=Fr, channels=C, timesteps=T,  # noqa: E731
                                               use_dynamic_thres=dyn, dynamic_thres_percentile=0.9)
    scale = torch.tensor([0.2, 1.0, 4.0]).view(B, 1, 1, 1, 1)  # sample 0: everything inside [-1, 1]
    x = torch.from_numpy(rng.standard_normal((B, C, Fr, S, S)).astype(np.float32)) * scale
    t = torch.tensor([3, 3, 3], dtype=torch.int32)
    gd_dyn, gd_st = mk(True), mk(False)
    x0 = gd_dyn.predict_start_from_noise(x, t, eps.permute(0, 4, 1, 2, 3)) * torch.tensor([0.1, 1.0, 1.0]).view(B, 1, 1, 1, 1)
    # feed x0 straight through q_posterior's inverse: compare the clipped x0 implied by the posterior mean
    m_dyn, _, _ = gd_dyn.p_mean_variance(x, t, clip_denoised=True)
    m_st, _, _ = gd_st.p_mean_variance(x, t, clip_denoised=True)
    c1 = D.extract(gd_dyn.posterior_mean_coef1, t, x.shape)
    c2 = D.extract(gd_dyn.posterior_mean_coef2, t, x.shape)
    x0_dyn = (m_dyn - c2 * x) / c1
    raw = gd_dyn.predict_start_from_noise(x, t, eps.permute(0, 4, 1, 2, 3))
    for b in range(B):
        s = max(float(np.quantile(np.abs(raw[b].numpy()).ravel(), 0.9)), 1.0)
        want = np.clip(raw[b].numpy(), -s, s) / s
        assert np.allclose(x0_dyn[b].numpy(), want, atol=2e-4)
        if s == 1.0:
            assert torch.allclose(m_dyn[b], m_st[b], atol=1e-6)
    assert float(x0_dyn.abs().max()) <= 1.0 + 1e-4
    del x0
