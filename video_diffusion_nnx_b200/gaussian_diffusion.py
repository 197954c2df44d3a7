"""GaussianDiffusion with the reference's method surface (gaussian_diffusion.py:53-65 and
:120-502), executed by the sm_100a kernels behind include/vdn.h.

`key` arguments: the reference threads jax.random PRNG keys; here a key is an integer seed (or a
`Key`), consumed by a counter-based Philox normal generator (csrc/misc.cu) so that every draw is a
pure function of (seed, stream, element index) - independent of how a batch is sharded over GPUs.
JAX's threefry streams cannot be reproduced without JAX, so parity tests pass t / noise explicitly.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import ops

SCHEDULE_NAMES = ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                  "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                  "posterior_mean_coef2")


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    """utils.py:241-256, in float32 (the reference runs with jax x64 disabled)."""
    f = np.float32
    x = np.linspace(0, timesteps, timesteps + 1, dtype=f)
    ac = np.cos(((x / f(timesteps)) + f(s)) / f(1 + s) * f(math.pi) * f(0.5)).astype(f) ** 2
    ac = ac / ac[0]
    return np.clip(f(1) - (ac[1:] / ac[:-1]), f(0), f(0.9999)).astype(f)


def make_schedule(timesteps: int) -> Dict[str, np.ndarray]:
    """The ten float32 tables of gaussian_diffusion.py:77-98."""
    f = np.float32
    betas = cosine_beta_schedule(timesteps)
    alphas = f(1) - betas
    ac = np.cumprod(alphas, axis=0, dtype=f)
    ac_prev = np.concatenate([np.ones(1, f), ac[:-1]])
    pv = betas * (f(1) - ac_prev) / (f(1) - ac)
    return {
        "alphas_cumprod": ac,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(f(1) - ac),
        "log_one_minus_alphas_cumprod": np.log(f(1) - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(f(1) / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(f(1) / ac - f(1)),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": np.log(np.maximum(pv, f(1e-20))),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (f(1) - ac),
        "posterior_mean_coef2": (f(1) - ac_prev) * np.sqrt(alphas) / (f(1) - ac),
    }


class Key:
    """Minimal stand-in for a jax PRNG key: (seed, stream counter)."""

    def __init__(self, seed: int, stream: int = 0):
        self.seed, self.stream = int(seed), int(stream)

    def split(self, n: int = 2):
        return [Key(self.seed, self.stream * 1000003 + i + 1) for i in range(n)]


def as_key(key) -> Key:
    return key if isinstance(key, Key) else Key(int(key))


def randint_from_key(key: Key, high: int, n: int) -> torch.Tensor:
    """n int32 draws from [0, high) as a pure function of the key (stands in for jax.random.randint,
    gaussian_diffusion.py:496); host tensor. One seed-folding rule for every caller."""
    g = torch.Generator(device="cpu").manual_seed((key.seed * 7919 + key.stream) % (2 ** 63))
    return torch.randint(0, high, (n,), generator=g, dtype=torch.int32)


class GaussianDiffusion:
    def __init__(self, denoise_fn, *, image_size: int, num_frames: int, text_use_bert_cls: bool = False,
                 channels: int = 3, timesteps: int = 1000, loss_type: str = "l1", use_dynamic_thres: bool = False,
                 dynamic_thres_percentile: float = 0.9):
        if loss_type not in ("l1", "l2"):
            raise ValueError(f"Unsupported loss type: {loss_type}")
        self.denoise_fn = denoise_fn
        self.image_size, self.num_frames, self.channels = image_size, num_frames, channels
        self.loss_type = loss_type
        self.text_use_bert_cls = text_use_bert_cls
        self.use_dynamic_thres = use_dynamic_thres
        self.dynamic_thres_percentile = dynamic_thres_percentile
        self.num_timesteps = int(timesteps)
        self._host_tables = make_schedule(self.num_timesteps)
        self._dev_tables: Dict[str, torch.Tensor] = {}

    # schedule tables as attributes, like the reference's nnx.Variables
    def __getattr__(self, name):
        if name in SCHEDULE_NAMES:
            return self.table(name)
        raise AttributeError(name)

    def table(self, name: str) -> torch.Tensor:
        if name not in self._dev_tables:
            self._dev_tables[name] = torch.from_numpy(self._host_tables[name]).to(self.denoise_fn.device)
        return self._dev_tables[name]

    @property
    def device(self):
        return self.denoise_fn.device

    # ---- state exchange (the tree `nnx.split(GaussianDiffusion)` gives: trainer.py:136, utils.py:486) ----------
    def state_dict(self, flat: Optional[torch.Tensor] = None) -> Dict[str, np.ndarray]:
        """`denoise_fn.<Unet3D nnx path>` leaves (from the device store, or from another flat buffer of the same layout
        such as the EMA copy) + the ten schedule tables: what the reference checkpoints per tree."""
        from .checkpoint import diffusion_state

        return diffusion_state(self.denoise_fn.state_dict(flat), self._host_tables)

    def load_state_dict(self, state: Dict[str, "np.ndarray | torch.Tensor"]) -> None:
        """Loads a reference-named state: Unet3D leaves (with or without the `denoise_fn.` prefix) and, when present,
        the schedule tables - the reference Adam-updates them, so a trained checkpoint's tables differ
        from the closed form and sampling must use the checkpoint's."""
        from .checkpoint import split_diffusion_state

        state = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in state.items()}
        unet, sched = split_diffusion_state(state)
        self.denoise_fn.load_state_dict(unet)
        for name, arr in sched.items():
            if arr.shape != (self.num_timesteps,):
                raise ValueError(f"{name}: expected shape ({self.num_timesteps},), got {arr.shape}")
            self._host_tables[name] = arr.astype(np.float32)
            if name in self._dev_tables:  # in place: captured sampler / training graphs hold this address
                self._dev_tables[name].copy_(torch.from_numpy(self._host_tables[name]))

    def _f32(self, x):
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    def _i32(self, t):
        return t.to(device=self.device, dtype=torch.int32).contiguous()

    def _normal(self, shape, key: Key, elem_offset: int = 0) -> torch.Tensor:
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        ops.randn(out, key.seed, key.stream, elem_offset)
        return out

    # ---- forward process --------------------------------------------------------------
    def q_sample(self, x_start, t, key=None, noise=None, *, _normalize: bool = False):
        """gaussian_diffusion.py:401-420."""
        x_start = self._f32(x_start)
        if noise is None:
            assert key is not None, "A PRNGKey must be provided to q_sample if noise is not."
            noise = self._normal(x_start.shape, as_key(key))
        out = torch.empty_like(x_start)
        ops.q_sample(x_start, self._f32(noise), self._i32(t), self.table("sqrt_alphas_cumprod"),
                     self.table("sqrt_one_minus_alphas_cumprod"), out, _normalize)
        return out

    def p_losses(self, x_start, t, key=None, cond=None, noise=None, *, _normalize: bool = False, **kwargs):
        """gaussian_diffusion.py:423-470 (forward only; the training step in trainer.py adds backward)."""
        key = as_key(0 if key is None else key)
        _, noise_key, q_key = key.split(3)
        x_start = self._f32(x_start)
        if noise is None:
            noise = self._normal(x_start.shape, noise_key)
        noise = self._f32(noise)
        t = self._i32(t)
        x_noisy = self.q_sample(x_start, t, key=q_key, noise=noise, _normalize=_normalize)
        pred = self.denoise_fn(x_noisy, t, cond=cond, **kwargs)  # (b f h w c)
        B, C = x_start.shape[0], x_start.shape[1]
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        ops.loss_fwd_bwd(pred, noise, loss, None, B, C, x_start.numel() // (B * C), self.loss_type == "l1")
        return loss[0]

    def __call__(self, x, key, *args, **kwargs):
        """gaussian_diffusion.py:473-502: random t, normalise to [-1,1], p_losses."""
        B = x.shape[0]
        assert tuple(x.shape[1:]) == (self.channels, self.num_frames, self.image_size, self.image_size), \
            "expected (b, c, f, h, w) matching the configured channels / frames / size"
        key = as_key(key)
        _, t_key, loss_key = key.split(3)
        t = randint_from_key(t_key, self.num_timesteps, B)
        return self.p_losses(x, t, loss_key, *args, _normalize=True, **kwargs)

    def _extract(self, name: str, t) -> torch.Tensor:
        """utils.py:225-238: table[t] broadcast to (b, 1, 1, 1, 1)."""
        return self.table(name)[self._i32(t).long()].view(-1, 1, 1, 1, 1)

    def q_mean_variance(self, x_start, t):
        """gaussian_diffusion.py:101-117 (host-visible helper, not on the hot path)."""
        mean = self._extract("sqrt_alphas_cumprod", t) * self._f32(x_start)
        variance = 1.0 - self._extract("alphas_cumprod", t)
        return mean, variance, self._extract("log_one_minus_alphas_cumprod", t)

    # ---- reverse process --------------------------------------------------------------
    def q_posterior(self, x_start, x_t, t):
        """gaussian_diffusion.py:139-159 (host-visible helper; the sampling loop uses the fused kernel)."""
        mean = (self._extract("posterior_mean_coef1", t) * self._f32(x_start)
                + self._extract("posterior_mean_coef2", t) * self._f32(x_t))
        return mean, self._extract("posterior_variance", t), self._extract("posterior_log_variance_clipped", t)

    def _x0_from_eps(self, x, t, eps, clip_denoised: bool):
        """x0 estimate of p_mean_variance (gaussian_diffusion.py:197-220) from the Unet output (b f h w c)."""
        B = x.shape[0]
        eps = eps.reshape(B, self.num_frames, self.image_size, self.image_size, self.channels).permute(0, 4, 1, 2, 3)
        x_recon = self.predict_start_from_noise(x, t, eps.float())
        if clip_denoised:
            if self.use_dynamic_thres:  # :205-217: per-sample quantile of |x0| (linear interpolation), at least 1
                s = torch.quantile(x_recon.abs().reshape(B, -1), self.dynamic_thres_percentile, dim=-1)
                s = torch.clamp(s, min=1.0).view(-1, 1, 1, 1, 1)
                x_recon = torch.maximum(torch.minimum(x_recon, s), -s) / s
            else:
                x_recon = x_recon.clamp(-1.0, 1.0)
        return x_recon

    def p_mean_variance(self, x, t, clip_denoised: bool, cond=None, cond_scale: float = 1.0):
        """gaussian_diffusion.py:162-228 (host-visible helper: Unet forward on the CUDA engine, then the posterior
        of the x0 estimate; p_sample / p_sample_loop fuse the same arithmetic into one kernel)."""
        x, t = self._f32(x), self._i32(t)
        eps = self.denoise_fn.forward_with_cond_scale(x, t, cond=cond, cond_scale=cond_scale)
        return self.q_posterior(self._x0_from_eps(x, t, eps, clip_denoised), x, t)

    def predict_start_from_noise(self, x_t, t, noise):
        """gaussian_diffusion.py:120-136 (host-visible helper; the fused kernel is p_sample)."""
        rc = self.table("sqrt_recip_alphas_cumprod")[self._i32(t).long()].view(-1, 1, 1, 1, 1)
        rm = self.table("sqrt_recipm1_alphas_cumprod")[self._i32(t).long()].view(-1, 1, 1, 1, 1)
        return rc * self._f32(x_t) - rm * self._f32(noise)

    def p_sample(self, x, t, key=None, cond=None, cond_scale: float = 1.0, clip_denoised: bool = True, *, z=None):
        """gaussian_diffusion.py:231-261: one Unet forward + the fused posterior update kernel."""
        x = self._f32(x)
        t = self._i32(t)
        eps = self.denoise_fn.forward_with_cond_scale(x, t, cond=cond, cond_scale=cond_scale)
        if z is None:
            z = self._normal(x.shape, as_key(key))
        if self.use_dynamic_thres and clip_denoised:
            # dynamic thresholding (off in every reference config) needs a per-sample quantile between the Unet and
            # the posterior update: composed from the host-visible helpers instead of the fused kernel
            mean, _, log_var = self.q_posterior(self._x0_from_eps(x, t, eps, True), x, t)
            nonzero = (t != 0).to(torch.float32).view(-1, 1, 1, 1, 1)
            return mean + nonzero * torch.exp(0.5 * log_var) * self._f32(z)
        out = torch.empty_like(x)
        B, C = x.shape[0], x.shape[1]
        ops.p_sample(x, eps, self._f32(z), t, self.table("sqrt_recip_alphas_cumprod"),
                     self.table("sqrt_recipm1_alphas_cumprod"), self.table("posterior_mean_coef1"),
                     self.table("posterior_mean_coef2"), self.table("posterior_log_variance_clipped"), out, B, C,
                     x.numel() // (B * C), clip_denoised)
        return out

    def p_sample_loop(self, shape, key, cond=None, cond_scale: float = 1.0, *, sample_offset: int = 0,
                      use_graph: bool = True, timesteps: Optional[int] = None):
        """gaussian_diffusion.py:264-320. The T host iterations replay ONE captured CUDA graph
        (Unet forward + Philox z + posterior update); `sample_offset` is the global index of this
        shard's first sample, so a sharded batch draws the same noise as the unsharded one."""
        if cond is not None or cond_scale != 1.0:
            # the reference's loop ignores both (gaussian_diffusion.py:301,316 never pass them on); an unconditional
            # model has nothing to scale, so anything but the defaults is a caller error here
            raise NotImplementedError("conditioning is outside the accelerated hot path (cond must be None, cond_scale 1.0)")
        key = as_key(key)
        B = shape[0]
        shape = (B, self.channels, self.num_frames, self.image_size, self.image_size)
        per_sample = shape[1] * shape[2] * shape[3] * shape[4]
        if self.use_dynamic_thres:  # eager loop over p_sample (quantile between the Unet and the update); same draws
            loop_key, init_key = key.split(2)
            T = self.num_timesteps if timesteps is None else timesteps
            img = self._normal(shape, init_key, sample_offset * per_sample)
            z = torch.empty(shape, dtype=torch.float32, device=self.device)
            for i in reversed(range(T)):
                ops.randn(z, loop_key.seed, loop_key.stream * 4099 + i + 1, sample_offset * per_sample)
                img = self.p_sample(img, torch.full((B,), i, dtype=torch.int32, device=self.device), z=z)
            return (img + 1) * 0.5
        assert per_sample % 4 == 0
        loop_key, init_key = key.split(2)
        T = self.num_timesteps if timesteps is None else timesteps
        FHW = per_sample // shape[1]
        # Per-batch-size sampler state: persistent buffers + ONE captured graph of a whole timestep
        # (Philox z drawn from the device-resident t, Unet forward, posterior update, t -= 1). A timestep is a
        # single graph replay with no host-side argument update; the graph is reused across calls.
        st = self._samplers.get(B) if hasattr(self, "_samplers") else None
        if not hasattr(self, "_samplers"):
            self._samplers = {}
        if st is None:
            st = {"eng": self.denoise_fn.engine(B, self.num_frames, self.image_size, self.image_size, training=False),
                  "img": torch.empty(shape, dtype=torch.float32, device=self.device),
                  "t": torch.empty(B, dtype=torch.int32, device=self.device),
                  "z": torch.empty(shape, dtype=torch.float32, device=self.device),
                  "nxt": torch.empty(shape, dtype=torch.float32, device=self.device),
                  "prm": torch.zeros(3, dtype=torch.int64, device=self.device), "graph": None}
            self._samplers[B] = st
        eng, img, t_dev, z, nxt = st["eng"], st["img"], st["t"], st["z"], st["nxt"]
        eng.sync_weights()  # a training step / state upload since the last call: repack before replaying the graph
        img.copy_(self._normal(shape, init_key, sample_offset * per_sample))
        tabs = [self.table(n) for n in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                                        "posterior_mean_coef1", "posterior_mean_coef2",
                                        "posterior_log_variance_clipped")]
        # z_i = Philox(seed, stream*4099 + i + 1, global element offset): the three values live in device memory
        seed64 = loop_key.seed & (2 ** 64 - 1)
        prm_host = torch.tensor([seed64 - 2 ** 64 if seed64 >= 2 ** 63 else seed64, loop_key.stream * 4099 + 1,
                                 sample_offset * per_sample], dtype=torch.int64)
        st["prm"].copy_(prm_host)
        prm = st["prm"]

        def body():
            ops.randn_t(z, prm, t_dev)
            eps = eng.forward(img, t_dev)
            ops.p_sample(img, eps, z, t_dev, *tabs, nxt, B, shape[1], FHW, True)
            img.copy_(nxt)
            ops.countdown(t_dev)

        t_dev.fill_(T - 1)
        if not use_graph:
            for _ in range(T):
                body()
        else:
            if st["graph"] is None:
                body()  # one eager timestep (also warms every lazily-built operand)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    body()
                st["graph"] = graph
                # restart from the initial state: the eager + captured steps consumed two timesteps
                img.copy_(self._normal(shape, init_key, sample_offset * per_sample))
                t_dev.fill_(T - 1)
            for _ in range(T):
                st["graph"].replay()
        return (img + 1) * 0.5  # unnormalize_img, utils.py:259-268

    def sample(self, key, cond=None, cond_scale: float = 1.0, batch_size: int = 16, **kw):
        """gaussian_diffusion.py:323-357."""
        if cond is not None:
            raise NotImplementedError("conditioning is outside the accelerated hot path")
        shape = (batch_size, self.channels, self.num_frames, self.image_size, self.image_size)
        return self.p_sample_loop(shape, key, cond=cond, cond_scale=cond_scale, **kw)
