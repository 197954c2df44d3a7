"""Stand-alone modules of the reference's modules.py that are callable on their own (test_modules.py:242-293):
`MultiheadAttention` with its optional `focus_present_mask` / `pos_bias` inputs (modules.py:247-326) and
`RelativePositionBias` (modules.py:330-390), executed by the sm_100a kernels behind include/vdn.h.

Inside `Unet3D` these inputs are dead - `PreNorm` drops every keyword argument (modules.py:146-148) and the bias
computed at unet3d.py:279 is never consumed - so the engine's fused attention kernels do not carry them
(SURVEY.md section 0.4). Here they are live, with the reference's literal semantics: mask and bias are applied AFTER
the softmax, masked entries become finfo(float32).min, and a mask that is true for every batch element returns
`out(v)`.

Same constructor arguments and call signatures as the reference classes; parameters in flax layouts under the
reference's names (`q/k/v.kernel (in, heads, dim)`, `out.kernel (heads, dim, in)`,
`relative_attention_bias.embedding (num_buckets, heads)`). Tensors are torch CUDA tensors (no JAX in this image).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from ._lib import VDN_BF16, check, lib, ptr, stream_ptr
from .ops import TAPS_1x1, VDN_TAP_UNIT


def _seed_of(rngs) -> int:
    if rngs is None:
        return 0
    if isinstance(rngs, int):
        return rngs
    return int(getattr(rngs, "seed", 0))


def _trunc_normal(rng, shape, fan_in):
    x = rng.standard_normal(size=shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(x) > 2
    return (x * (math.sqrt(1.0 / fan_in) / 0.87962566103423978)).astype(np.float32)


class RelativePositionBias:
    """modules.py:330-390. `__call__(n)` -> (heads, n, n) fp32. As in the reference, `__call__` buckets with the static
    defaults 32 / 128 whatever `num_buckets` / `max_distance` the constructor was given (modules.py:386; Unet3D passes
    max_distance=32, unet3d.py:98-100) - only the embedding table size follows `num_buckets`."""

    def __init__(self, rngs=0, heads: int = 8, num_buckets: int = 32, max_distance: int = 128, device="cuda"):
        if num_buckets < 32:
            raise ValueError("bucket ids reach 31 (static defaults of __call__): the table needs >= 32 rows")
        self.heads, self.num_buckets, self.max_distance = heads, num_buckets, max_distance
        self.device = torch.device(device)
        rng = np.random.default_rng(_seed_of(rngs))
        # nnx.Embed default init: variance_scaling(1.0, 'fan_in', 'normal', out_axis=0)
        self._host = (rng.standard_normal(size=(num_buckets, heads)) / math.sqrt(num_buckets)).astype(np.float32)
        self._dev: Optional[torch.Tensor] = None

    @property
    def embedding(self) -> torch.Tensor:
        if self._dev is None:
            self._dev = torch.from_numpy(self._host).to(self.device)
        return self._dev

    def state_dict(self) -> Dict[str, np.ndarray]:
        return {"relative_attention_bias.embedding": self._host.copy()}

    def load_state_dict(self, state) -> None:
        a = np.asarray(state["relative_attention_bias.embedding"], np.float32)
        if a.shape != self._host.shape:
            raise ValueError(f"embedding: expected {self._host.shape}, got {a.shape}")
        self._host = a.copy()
        self._dev = None

    def _run(self, n: int, want_buckets: bool):
        out = torch.empty((self.heads, n, n), dtype=torch.float32, device=self.device)
        buckets = torch.empty((n, n), dtype=torch.int32, device=self.device) if want_buckets else None
        check(lib.vdn_rel_pos_bias(ptr(self.embedding), n, self.heads, ptr(out), ptr(buckets), stream_ptr()),
              "vdn_rel_pos_bias")
        return out, buckets

    def __call__(self, n: int) -> torch.Tensor:
        return self._run(int(n), False)[0]

    def buckets(self, n: int) -> torch.Tensor:
        """int32 (n, n) bucket ids of rel = i - j (modules.py:351-378): integer work, bit-exact."""
        return self._run(int(n), True)[1]


class MultiheadAttention:
    """modules.py:247-326 on x (..., F, in_features): q/k/v = LinearGeneral(in -> (heads, dim)) + bias, q / sqrt(dim),
    softmax over keys, optional post-softmax mask / bias, out = LinearGeneral((heads, dim) -> in) + bias.
    The projections run on the tcgen05 tap-GEMM (bf16 operands, fp32 accumulation), the core on vdn_mha_core_ext_fwd."""

    def __init__(self, in_features: int, dim: int, num_heads: int, rngs=0, rotary_emb=None, device="cuda"):
        if rotary_emb is not None:
            raise NotImplementedError("rotary embeddings are a TODO in the reference as well (modules.py:296-300)")
        if in_features % 16 or (num_heads * dim) % 16 or dim not in (8, 16, 32, 64):
            raise NotImplementedError("in_features and heads*dim must be multiples of 16, dim in {8, 16, 32, 64}")
        self.in_features, self.dim, self.num_heads = in_features, dim, num_heads
        self.device = torch.device(device)
        rng = np.random.default_rng(_seed_of(rngs))
        hd = num_heads * dim
        self._host = {}
        for n in ("q", "k", "v"):
            self._host[f"{n}.kernel"] = _trunc_normal(rng, (in_features, num_heads, dim), in_features)
            self._host[f"{n}.bias"] = np.zeros((num_heads, dim), np.float32)
        self._host["out.kernel"] = _trunc_normal(rng, (num_heads, dim, in_features), hd)
        self._host["out.bias"] = np.zeros((in_features,), np.float32)
        self._packed = None

    def state_dict(self) -> Dict[str, np.ndarray]:
        return {k: v.copy() for k, v in self._host.items()}

    def load_state_dict(self, state) -> None:
        for k, cur in self._host.items():
            a = np.asarray(state[k].detach().cpu().numpy() if isinstance(state[k], torch.Tensor) else state[k], np.float32)
            if a.shape != cur.shape:
                raise ValueError(f"{k}: expected {cur.shape}, got {a.shape}")
            self._host[k] = a.copy()
        self._packed = None

    def _pack(self):
        if self._packed is None:
            C, hd, dev = self.in_features, self.num_heads * self.dim, self.device
            w = np.concatenate([self._host[f"{n}.kernel"].reshape(C, hd) for n in ("q", "k", "v")], axis=1)  # [C][3hd]
            b = np.concatenate([self._host[f"{n}.bias"].reshape(hd) for n in ("q", "k", "v")])
            w_dev = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
            wo_dev = torch.from_numpy(np.ascontiguousarray(self._host["out.kernel"].reshape(hd, C))).to(dev)
            wp = torch.empty(3 * hd, C, dtype=torch.bfloat16, device=dev)
            wop = torch.empty(C, hd, dtype=torch.bfloat16, device=dev)
            ops.pack_weight(w_dev, wp, 1, C, 3 * hd, 0)
            ops.pack_weight(wo_dev, wop, 1, hd, C, 0)
            self._packed = (wp, torch.from_numpy(b).to(dev), wop, torch.from_numpy(self._host["out.bias"]).to(dev))
        return self._packed

    def __call__(self, x: torch.Tensor, focus_present_mask: Optional[torch.Tensor] = None,
                 pos_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not torch.cuda.is_available():
            raise RuntimeError("MultiheadAttention needs a CUDA device (sm_100a); there is no CPU fallback")
        C, H, D = self.in_features, self.num_heads, self.dim
        hd = H * D
        assert x.shape[-1] == C and x.dim() >= 2
        S = x.shape[-2]
        P = x.numel() // C
        n_seq = P // S
        wp, b_qkv, wop, b_out = self._pack()
        # rows padded to whole 128-pixel tiles of a (tiles, 1, 128) grid for the tap-GEMM
        Pp = (P + 127) // 128 * 128
        xb = torch.zeros((Pp // 128, 1, 128, C), dtype=torch.bfloat16, device=self.device)
        xb.view(Pp, C)[:P].copy_(x.reshape(P, C))
        qkv = torch.empty((Pp // 128, 1, 128, 3 * hd), dtype=torch.bfloat16, device=self.device)
        ops.tapgemm(VDN_TAP_UNIT, [xb], wp, TAPS_1x1, bias=b_qkv, out=qkv)
        o = torch.zeros((Pp // 128, 1, 128, hd), dtype=torch.bfloat16, device=self.device)
        mask_dev, spb, copy_v = None, 1, 0
        if focus_present_mask is not None:
            m = torch.as_tensor(focus_present_mask).to(torch.bool).reshape(-1).cpu()
            nb = m.numel()
            assert x.shape[0] == nb and n_seq % nb == 0, "focus_present_mask has one entry per batch element (axis 0)"
            if bool(m.all()):  # modules.py:291-292: every sample focuses on the present -> out(v)
                copy_v = 1
            elif bool(m.any()):
                mask_dev, spb = m.to(torch.uint8).to(self.device), n_seq // nb
        pb = None
        if pos_bias is not None and not copy_v:
            pb = pos_bias.to(device=self.device, dtype=torch.float32).contiguous()
            assert tuple(pb.shape) == (H, S, S), f"pos_bias must be (heads, F, F) = {(H, S, S)}"
        check(lib.vdn_mha_core_ext_fwd(ptr(qkv), ptr(o), VDN_BF16, H, D, n_seq, S, 1, ptr(mask_dev), spb, ptr(pb),
                                       copy_v, stream_ptr()), "vdn_mha_core_ext_fwd")
        out = torch.empty((Pp // 128, 1, 128, C), dtype=torch.float32, device=self.device)
        ops.tapgemm(VDN_TAP_UNIT, [o], wop, TAPS_1x1, bias=b_out, out=out, out_dtype=torch.float32)
        return out.view(Pp, C)[:P].reshape(x.shape)
