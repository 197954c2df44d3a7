"""ORACLE (test infrastructure only) - CPU restatement of the reference GaussianDiffusion math.

NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this, as the checker / CPU baseline.

Follows gaussian_diffusion.py:77-98 (schedule tables), :120-136 (predict_start_from_noise),
:139-159 (q_posterior), :162-228 (p_mean_variance), :231-261 (p_sample), :401-420 (q_sample),
:423-470 (p_losses), :473-502 (__call__); utils.py:225-238 (extract), :241-256
(cosine_beta_schedule), :259-280 ((un)normalize_img).

Pinned by the reference's own known answers (gaussian_diffusion_test.py:88-109,111-123,135-158,
175-189,191-210; utils_test.py:102-110,121-131) in tests/test_oracle_known_answers.py.
Also pinned by executing the reference's own gaussian_diffusion.py / utils.py over oracle/refshim
(tests/test_oracle_vs_reference_code.py: every method at 1e-9, the ten schedule tables bit-identical in
float32). For the Unet numerics see the PARITY PIN paragraph of oracle/unet3d_oracle.py.

jax.random (threefry) streams cannot be reproduced here, so every random draw (t, noise, the
per-step z of p_sample) is an explicit input.
"""
from __future__ import annotations

import math
from typing import Callable, Dict

import numpy as np
import torch

SCHEDULE_NAMES = (
    "alphas_cumprod",
    "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod",
    "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod",
    "posterior_variance",
    "posterior_log_variance_clipped",
    "posterior_mean_coef1",
    "posterior_mean_coef2",
)


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    """utils.py:241-256. The reference asks for float64 but runs with jax x64 disabled, so the
    arithmetic is float32; restated in float32 numpy."""
    f = np.float32
    steps = timesteps + 1
    x = np.linspace(0, timesteps, steps, dtype=f)
    ac = np.cos(((x / f(timesteps)) + f(s)) / f(1 + s) * f(math.pi) * f(0.5)).astype(f) ** 2
    ac = ac / ac[0]
    betas = f(1) - (ac[1:] / ac[:-1])
    return np.clip(betas, f(0), f(0.9999)).astype(f)


def make_schedule(timesteps: int) -> Dict[str, np.ndarray]:
    """gaussian_diffusion.py:77-98: the ten float32 tables of length T."""
    f = np.float32
    betas = cosine_beta_schedule(timesteps).astype(f)
    alphas = f(1) - betas
    ac = np.cumprod(alphas, axis=0, dtype=f)
    ac_prev = np.concatenate([np.ones(1, f), ac[:-1]])
    post_var = betas * (f(1) - ac_prev) / (f(1) - ac)
    return {
        "alphas_cumprod": ac,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(f(1) - ac),
        "log_one_minus_alphas_cumprod": np.log(f(1) - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(f(1) / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(f(1) / ac - f(1)),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.maximum(post_var, f(1e-20))),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (f(1) - ac),
        "posterior_mean_coef2": (f(1) - ac_prev) * np.sqrt(alphas) / (f(1) - ac),
    }


def extract(a: torch.Tensor, t: torch.Tensor, x_shape) -> torch.Tensor:
    """utils.py:225-238: take_along_axis(a, t) reshaped to (b,1,1,...)."""
    b = t.shape[0]
    out = torch.gather(a, -1, t.long())
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def normalize_img(t):  # utils.py:271-280
    return t * 2 - 1


def unnormalize_img(t):  # utils.py:259-268
    return (t + 1) * 0.5


class GaussianDiffusionOracle:
    """Same method names / argument meaning as the reference class; `denoise_fn(x_bcfhw, t)` must
    return (b,f,h,w,c) like Unet3D (unet3d.py:387)."""

    def __init__(self, denoise_fn: Callable, *, image_size: int, num_frames: int, channels: int = 3,
                 timesteps: int = 1000, loss_type: str = "l1", dtype=torch.float32, use_dynamic_thres: bool = False,
                 dynamic_thres_percentile: float = 0.9):
        self.denoise_fn = denoise_fn
        self.use_dynamic_thres, self.dynamic_thres_percentile = use_dynamic_thres, dynamic_thres_percentile  # :61-62
        self.image_size, self.num_frames, self.channels = image_size, num_frames, channels
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        self.dtype = dtype
        for k, v in make_schedule(self.num_timesteps).items():
            setattr(self, k, torch.from_numpy(v).to(dtype))

    def q_sample(self, x_start, t, noise):  # :401-420
        return (extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def q_mean_variance(self, x_start, t):  # :101-117
        mean = extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = extract(1.0 - self.alphas_cumprod, t, x_start.shape)
        log_variance = extract(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def predict_start_from_noise(self, x_t, t, noise):  # :120-136
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def q_posterior(self, x_start, x_t, t):  # :139-159
        mean = (extract(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        return (mean, extract(self.posterior_variance, t, x_t.shape),
                extract(self.posterior_log_variance_clipped, t, x_t.shape))

    def p_mean_variance(self, x, t, clip_denoised: bool = True):  # :162-228 (cond_scale 1)
        eps = self.denoise_fn(x, t).permute(0, 4, 1, 2, 3)  # 'b f h w c -> b c f h w' :197
        x_recon = self.predict_start_from_noise(x, t, eps)
        if clip_denoised:
            s = 1.0
            if getattr(self, "use_dynamic_thres", False):  # :205-217 (Imagen): per-sample quantile of |x0|, >= 1
                flat = x_recon.abs().reshape(x_recon.shape[0], -1)
                s = torch.quantile(flat, self.dynamic_thres_percentile, dim=-1)  # jnp.quantile default: linear
                s = torch.clamp(s, min=1.0).reshape(-1, 1, 1, 1, 1)
                x_recon = torch.maximum(torch.minimum(x_recon, s), -s) / s
            else:
                x_recon = x_recon.clamp(-s, s) / s
        return self.q_posterior(x_recon, x, t)

    def p_sample(self, x, t, z, clip_denoised: bool = True):  # :231-261, z = the N(0,1) draw of :254
        mean, _, log_var = self.p_mean_variance(x, t, clip_denoised)
        nonzero = (1.0 - (t == 0).to(x.dtype)).reshape(-1, 1, 1, 1, 1)
        return mean + nonzero * torch.exp(0.5 * log_var) * z

    def p_sample_loop(self, img, z_fn: Callable[[int], torch.Tensor]):  # :264-320
        b = img.shape[0]
        for i in reversed(range(self.num_timesteps)):
            t = torch.full((b,), i, dtype=torch.int32)
            img = self.p_sample(img, t, z_fn(i))
        return unnormalize_img(img)

    def p_losses(self, x_start, t, noise):  # :423-470
        x_noisy = self.q_sample(x_start, t, noise)
        pred = self.denoise_fn(x_noisy, t).permute(0, 4, 1, 2, 3)  # :460
        if self.loss_type == "l1":
            return (pred - noise).abs().mean()
        if self.loss_type == "l2":
            return ((pred - noise) ** 2).mean()
        raise ValueError(f"Unsupported loss type: {self.loss_type}")

    def __call__(self, x, t, noise):  # :473-502 with the random t / noise made explicit
        return self.p_losses(normalize_img(x), t, noise)
