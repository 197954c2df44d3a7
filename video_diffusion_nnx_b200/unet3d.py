"""Unet3D with the reference's constructor / call surface (unet3d.py:58-75,254-270), executed by the
sm_100a kernels behind include/vdn.h.

Same names, argument meaning and output layout ((b f h w c), unet3d.py:387) as the reference class.
The state is exchanged through state_dict()/load_state_dict() keyed by the reference's nnx state
paths in flax layouts, so reference checkpoints map one to one.

Differences, stated: tensors are torch CUDA tensors (there is no JAX in this image; INTEGRATION.md
shows the jax.ffi registration of the same C ABI); text conditioning (cond_dim / use_bert_text_cond)
is out of scope and raises; `rngs` is an integer seed (or an object with a `.seed`).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .engine import DIM_HEAD, HD, HEADS, ParamStore, UnetEngine, internal_param_spec


def _seed_of(rngs) -> int:
    if rngs is None:
        return 0
    if isinstance(rngs, int):
        return rngs
    return int(getattr(rngs, "seed", 0))


class Unet3D:
    def __init__(self, dim: int, rngs=0, dim_mults=(1, 2, 4, 8), cond_dim=None, out_dim=None, channels: int = 3,
                 attn_heads: int = 8, attn_dim_head: int = 32, use_bert_text_cond: bool = False, init_dim=None,
                 init_kernel_size: int = 7, use_sparse_linear_attn: bool = True, block_type: str = "resnet",
                 resnet_groups: int = 8, log_dims: bool = False, device="cuda", precision: str = "bf16"):
        """Reference constructor arguments (unet3d.py:58-75) + `device` and `precision`: "bf16" = the tensor-core
        throughput path (bf16 activations / operands, fp32 accumulation), "fp32" = the fp32-grade forward path
        (engine_f32.py: fp32 activations, split-bf16 operands on the same tcgen05 kernels; forward only, used for
        evaluation / parity at the reference's own float32 precision)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        if cond_dim is not None or use_bert_text_cond:
            raise NotImplementedError("text conditioning is outside the accelerated hot path (SURVEY.md C7)")
        if attn_heads != HEADS or attn_dim_head != DIM_HEAD:
            raise NotImplementedError("the fused attention kernels are built for attn_heads=8, attn_dim_head=32 (every "
                                      "reference config; SpatialLinearAttention hard-codes D=32 too, unet3d.py:174,225)")
        if init_dim not in (None, dim):
            # the reference cannot run this either: final_conv's block is built for 2*dim input channels
            # (unet3d.py:249-250) while the concat at :377 delivers dim + init_dim
            raise NotImplementedError("init_dim != dim: the reference's own final block (2*dim inputs) rejects it")
        if init_kernel_size % 2 != 1 or dim % 32 != 0:
            raise NotImplementedError("init_kernel_size must be odd (unet3d.py:105) and dim a multiple of 32")
        if block_type != "resnet":
            raise NotImplementedError("block_type is accepted and ignored by the reference; only 'resnet' exists")
        # GroupNorm groups (modules.py:167): channels per group must be a power of two >= 2 at every level
        cmin = dim * min([1] + list(dim_mults))
        cpg = cmin // resnet_groups if resnet_groups > 0 and cmin % resnet_groups == 0 else 0
        if cpg < 2 or (cpg & (cpg - 1)) != 0 or resnet_groups not in (2, 4, 8):
            # the conv epilogues keep at most 8 (sum, sum-of-squares) pairs per column tile
            raise NotImplementedError("resnet_groups must be 2, 4 or 8 (every reference config uses 8)")
        self.resnet_groups = resnet_groups
        self.use_sparse_linear_attn = bool(use_sparse_linear_attn)
        self.dim, self.channels, self.dim_mults = dim, channels, tuple(dim_mults)
        self.out_dim = channels if out_dim is None else out_dim
        self.init_kernel_size = init_kernel_size
        self.log_dims = log_dims
        self.has_cond = False
        self.device = torch.device(device)
        self.spec = internal_param_spec(dim, channels, self.dim_mults, init_kernel_size, self.out_dim,
                                        self.use_sparse_linear_attn)
        self.store: Optional[ParamStore] = None
        self._engines: Dict[tuple, UnetEngine] = {}
        self._f32_engines: Dict[tuple, object] = {}
        self._host_state = self._init_state(_seed_of(rngs))
        self.training = False

    # ------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------
    def _init_state(self, seed: int) -> Dict[str, np.ndarray]:
        """flax default initialisers: lecun_normal kernels (truncated normal / sqrt(fan_in)), zero biases,
        unit norm scales, Embed normal(stddev 1/sqrt(features)). Not bit-identical to nnx.Rngs(seed)
        (JAX threefry is not reproducible here); benchmarks only need the distributions."""
        rng = np.random.default_rng(seed)
        out: Dict[str, np.ndarray] = {}

        def trunc_normal(shape, fan_in):
            x = rng.standard_normal(size=shape)
            bad = np.abs(x) > 2
            while bad.any():
                x[bad] = rng.standard_normal(size=int(bad.sum()))
                bad = np.abs(x) > 2
            return (x * (math.sqrt(1.0 / fan_in) / 0.87962566103423978)).astype(np.float32)

        for name, shape in self.reference_param_shapes().items():
            leaf = name.rsplit(".", 1)[1]
            if leaf == "kernel":
                if len(shape) == 3 and shape[0] == HEADS and shape[1] == DIM_HEAD:  # LinearGeneral out
                    fan_in = HD
                elif len(shape) == 3 and shape[1] == HEADS:  # LinearGeneral q/k/v
                    fan_in = shape[0]
                else:
                    fan_in = int(np.prod(shape[:-1]))
                out[name] = trunc_normal(shape, fan_in)
            elif leaf == "embedding":
                out[name] = (rng.standard_normal(size=shape) / math.sqrt(shape[0])).astype(np.float32)
            elif leaf == "scale":
                out[name] = np.ones(shape, np.float32)
            else:
                out[name] = np.zeros(shape, np.float32)
        return out

    def reference_param_shapes(self) -> Dict[str, tuple]:
        """nnx state paths and flax shapes of the reference Unet3D (unet3d.py:58-252)."""
        s: Dict[str, tuple] = {}
        dim, ch, td = self.dim, self.channels, self.dim * 4
        k = self.init_kernel_size

        def mha(p, c):
            s[p + ".fn.norm.scale"] = (c,)
            s[p + ".fn.norm.bias"] = (c,)
            for n in ("q", "k", "v"):
                s[f"{p}.fn.fn.fn.{n}.kernel"] = (c, HEADS, DIM_HEAD)
                s[f"{p}.fn.fn.fn.{n}.bias"] = (HEADS, DIM_HEAD)
            s[p + ".fn.fn.fn.out.kernel"] = (HEADS, DIM_HEAD, c)
            s[p + ".fn.fn.fn.out.bias"] = (c,)

        def sla(p, c):
            if not self.use_sparse_linear_attn:  # Identity() holds no state (unet3d.py:179-181)
                return
            s[p + ".fn.norm.scale"] = (c,)
            s[p + ".fn.norm.bias"] = (c,)
            for n in ("q", "k", "v"):
                s[f"{p}.fn.fn.{n}.kernel"] = (1, c, HD)
            s[p + ".fn.fn.to_out.kernel"] = (1, HD, c)

        def resnet(p, cin, cout, time=True):
            if time:
                s[p + ".mlp.layers.1.kernel"] = (td, 2 * cout)
                s[p + ".mlp.layers.1.bias"] = (2 * cout,)
            s[p + ".norm_1.scale"] = (2 * cout,)
            s[p + ".norm_1.bias"] = (2 * cout,)
            for b, ci in (("block_1", cin), ("block_2", cout)):
                s[f"{p}.{b}.proj.kernel"] = (1, 3, 3, ci, cout)
                s[f"{p}.{b}.proj.bias"] = (cout,)
                s[f"{p}.{b}.norm.scale"] = (cout,)
                s[f"{p}.{b}.norm.bias"] = (cout,)
            if cin != cout:
                s[p + ".res_conv.kernel"] = (1, cin, cout)
                s[p + ".res_conv.bias"] = (cout,)
            s[p + ".norm_2.scale"] = (cout,)
            s[p + ".norm_2.bias"] = (cout,)

        s["time_rel_pos_bias.relative_attention_bias.embedding"] = (32, HEADS)
        s["init_conv.kernel"] = (1, k, k, ch, dim)
        s["init_conv.bias"] = (dim,)
        mha("init_temporal_attn", dim)
        s["time_mlp.layers.1.kernel"] = (dim, td)
        s["time_mlp.layers.1.bias"] = (td,)
        s["time_mlp.layers.3.kernel"] = (td, td)
        s["time_mlp.layers.3.bias"] = (td,)
        dims = [dim] + [dim * m for m in self.dim_mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        n = len(in_out)
        for l, (ci, co) in enumerate(in_out):
            resnet(f"downs.{l}.0", ci, co)
            resnet(f"downs.{l}.1", co, co)
            sla(f"downs.{l}.2", co)
            mha(f"downs.{l}.3", co)
            if l < n - 1:
                s[f"downs.{l}.4.kernel"] = (1, 4, 4, co, co)
                s[f"downs.{l}.4.bias"] = (co,)
        mid = dims[-1]
        resnet("mid_block1", mid, mid)
        mha("mid_spatial_attn", mid)
        mha("mid_temporal_attn", mid)
        resnet("mid_block2", mid, mid)
        for i, (ci, co) in enumerate(reversed(in_out)):
            resnet(f"ups.{i}.0", co * 2, ci)
            resnet(f"ups.{i}.1", ci, ci)
            sla(f"ups.{i}.2", ci)
            mha(f"ups.{i}.3", ci)
            if i < n - 1:
                s[f"ups.{i}.4.kernel"] = (1, 4, 4, ci, ci)
                s[f"ups.{i}.4.bias"] = (ci,)
        resnet("final_conv.layers.0", dim * 2, dim, time=False)
        s["final_conv.layers.1.kernel"] = (1, dim, self.out_dim)
        s["final_conv.layers.1.bias"] = (self.out_dim,)
        return s

    # reference path -> (internal name, column slice of the fused qkv or None)
    @staticmethod
    def _to_internal(name: str):
        n = name
        for a, b in ((".mlp.layers.1.", ".mlp."), ("time_mlp.layers.", "time_mlp."), ("final_conv.layers.", "final_conv."),
                     ("time_rel_pos_bias.relative_attention_bias.", "time_rel_pos_bias.")):
            n = n.replace(a, b)
        n = n.replace(".fn.fn.fn.", ".").replace(".fn.fn.", ".").replace(".fn.norm.", ".norm.")
        for i, q in enumerate(("q", "k", "v")):
            for leaf in ("kernel", "bias"):
                if n.endswith(f".{q}.{leaf}"):
                    return n[: -len(f".{q}.{leaf}")] + f".qkv.{leaf}", i
        return n, None

    def load_state_dict(self, state: Dict[str, "np.ndarray | torch.Tensor"]) -> None:
        """Load reference-named arrays (flax layouts) into the flat device store. Keys may carry the `denoise_fn.`
        prefix of the reference's checkpoint tree (the trainer splits the GaussianDiffusion module, trainer.py:136);
        the schedule tables of such a tree are ignored here (GaussianDiffusion.load_state_dict takes them)."""
        from .checkpoint import split_diffusion_state

        state, _ = split_diffusion_state(dict(state))
        shapes = self.reference_param_shapes()
        missing = [k for k in shapes if k not in state]
        if missing:
            raise KeyError(f"state is missing {len(missing)} entries, e.g. {missing[:3]}")
        host = {}
        for k, shape in shapes.items():
            v = state[k]
            v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            if tuple(v.shape) != tuple(shape):
                raise ValueError(f"{k}: expected shape {shape}, got {tuple(v.shape)}")
            host[k] = v.astype(np.float32)
        self._host_state = host
        if self.store is not None:
            self._upload()

    def _upload(self) -> None:
        st = self.store
        bufs = {name: np.zeros(shape, np.float32) for name, shape in self.spec}
        for k, v in self._host_state.items():
            name, qi = self._to_internal(k)
            dst = bufs[name]
            if qi is None:
                dst[...] = v.reshape(dst.shape)
            elif name.endswith(".kernel"):
                dst[:, qi * HD:(qi + 1) * HD] = v.reshape(dst.shape[0], HD)
            else:
                dst[qi * HD:(qi + 1) * HD] = v.reshape(HD)
        flat = np.zeros(st.total, np.float32)
        for name, _ in self.spec:
            off, shape = st.offsets[name]
            flat[off:off + bufs[name].size] = bufs[name].ravel()
        st.flat.copy_(torch.from_numpy(flat))
        st.bump()  # every engine (and every sampler graph built on one) repacks before its next forward

    def state_dict(self, flat: Optional[torch.Tensor] = None) -> Dict[str, np.ndarray]:
        """Reference-named numpy arrays (flax layouts) read back from the device store (or from another
        flat buffer of the same layout, e.g. the EMA copy)."""
        if self.store is None:
            return dict(self._host_state)
        src = (self.store.flat if flat is None else flat).detach().cpu().numpy()
        out = {}
        for k, shape in self.reference_param_shapes().items():
            name, qi = self._to_internal(k)
            off, ishape = self.store.offsets[name]
            a = src[off:off + int(np.prod(ishape))].reshape(ishape)
            if qi is None:
                out[k] = a.reshape(shape).copy()
            elif name.endswith(".kernel"):
                out[k] = a[:, qi * HD:(qi + 1) * HD].reshape(shape).copy()
            else:
                out[k] = a[qi * HD:(qi + 1) * HD].reshape(shape).copy()
        return out

    # ------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------
    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def _ensure_store(self, with_grad: bool) -> None:
        if self.store is None:
            if not torch.cuda.is_available():
                raise RuntimeError("Unet3D needs a CUDA device (sm_100a); there is no CPU fallback")
            self.store = ParamStore(self.spec, self.device, with_grad)
            self._upload()
        elif with_grad:
            self.store.ensure_grad()  # in place: existing engines / sampler graphs keep pointing at the same store

    def engine(self, B: int, F: int, H: int, W: int, training: Optional[bool] = None) -> UnetEngine:
        training = self.training if training is None else training
        self._ensure_store(training)
        key = (B, F, H, W, training)
        if key not in self._engines:
            self._engines[key] = UnetEngine(self.store, dim=self.dim, channels=self.channels, dim_mults=self.dim_mults,
                                            init_kernel_size=self.init_kernel_size, B=B, F=F, H=H, W=W,
                                            training=training, out_dim=self.out_dim, groups=self.resnet_groups,
                                            use_sla=self.use_sparse_linear_attn)
        return self._engines[key]

    def __call__(self, x: torch.Tensor, time: torch.Tensor, cond=None, null_cond_prob: float = 0.0,
                 focus_present_mask=None, prob_focus_present: float = 0.0) -> torch.Tensor:
        """x (b c f h w) fp32, time (b,) int -> (b f h w c) fp32. focus_present_mask / pos_bias never reach
        the attention in the reference (PreNorm drops kwargs, modules.py:146-148), so they are ignored."""
        if cond is not None:
            raise NotImplementedError("conditioning is outside the accelerated hot path")
        assert x.dim() == 5 and x.shape[1] == self.channels, "expected (b, c, f, h, w)"
        B, _, Fr, H, W = x.shape
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        time = time.to(device=self.device, dtype=torch.int32).contiguous()
        if self.precision == "fp32" and not self.training:
            return self.forward_fp32(x, time)
        return self.engine(B, Fr, H, W).forward(x, time)

    def forward_fp32(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """The same forward at the reference's float32 precision (engine_f32.F32Engine): (b c f h w) -> (b f h w c)."""
        from .engine_f32 import F32Engine

        B, _, Fr, H, W = x.shape
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        time = time.to(device=self.device, dtype=torch.int32).contiguous()
        self._ensure_store(False)
        key = (B, Fr, H, W)
        if key not in self._f32_engines:
            self._f32_engines[key] = F32Engine(self, B, Fr, H, W)
        return self._f32_engines[key].forward(x, time)

    def forward_with_cond_scale(self, *args, cond_scale: float = 2.0, **kwargs):
        """unet3d.py:254-260: without conditioning this is exactly one forward."""
        return self.__call__(*args, null_cond_prob=0.0, **kwargs)
