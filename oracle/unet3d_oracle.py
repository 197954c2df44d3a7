"""ORACLE (test infrastructure only) - CPU restatement of the reference Unet3D forward.

This file is NOT part of the product. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker / CPU baseline.

PARITY PIN. The reference (maxsonate/video-diffusion-nnx) is flax.nnx / JAX code; jax / flax are not installable in
this image and the reference's own tests assert shapes only (test_unet3d.py:24,48; test_modules.py). What pins this
oracle:
  1. THE REFERENCE'S OWN PYTHON, EXECUTED HERE: tests/golden/make_ref_golden.py imports /root/reference/modules.py,
     unet3d.py, gaussian_diffusion.py and utils.py unmodified over oracle/refshim (a numpy restatement of the jax /
     flax.nnx calls they make) and records the Unet3D forward, the diffusion methods, the stand-alone modules and the
     state tree; tests/test_oracle_vs_reference_code.py holds this oracle to those fixtures at 1e-9 (float64). That
     pins graph wiring, argument plumbing, dead-code behaviours, the diffusion algebra, names / shapes / count of the
     parameters - everything a re-reading can get wrong.
  2. every closed-form known answer the reference's tests hold for this path (tests/test_oracle_known_answers.py);
  3. the flax layer semantics re-derived from definition in numpy (tests/test_oracle_semantics.py).
What remains UNPINNED: the numerics INSIDE the flax layers and jax.nn functions as XLA evaluates them (conv
accumulation order, transcendental implementations) - restated from the public flax / jax definitions, twice, by
different routes (torch calls here, explicit dilate / pad / shifted-slice matmuls in refshim).

Restates, in plain PyTorch on CPU (fp32 or fp64), the EFFECTIVE graph of
  unet3d.py:262-387   Unet3D.__call__
  modules.py:64-129   SpatialLinearAttention (q scale discarded, :108 vs :118)
  modules.py:132-148  PreNorm (norm computed and thrown away; kwargs dropped)
  modules.py:150-179  Block;  modules.py:182-243 ResnetBlock
  modules.py:247-326  MultiheadAttention; modules.py:330-390 RelativePositionBias
  modules.py:30-45    SinusoidalPosEmb
  utils.py:103-125    Upsample / Downsample
with the flax defaults hard-coded: SAME padding, GroupNorm/LayerNorm
eps 1e-6 with "fast variance" E[x^2]-E[x]^2, tanh-GELU, ConvTranspose with transpose_kernel=False.

Parameters live in a flat dict keyed by the nnx state path, in flax layouts:
  Conv kernel (1,kh,kw,in,out) | 1x1 Conv kernel (1,in,out) | Linear (in,out)
  LinearGeneral q/k/v (in,heads,dim), out (heads,dim,out) | norms scale/bias (C,)
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

HEADS = 8
DIM_HEAD = 32
SLA_D = 32  # unet3d.py:174,225 hard-codes D=32
GROUPS = 8
EPS = 1e-6


# --------------------------------------------------------------------------------------
# flax layer semantics
# --------------------------------------------------------------------------------------
def conv_khw(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1) -> torch.Tensor:
    """nnx.Conv with kernel (1,kh,kw,in,out), SAME padding, on (B,F,H,W,C). modules.py:162, unet3d.py:110,
    utils.py:125. SAME for k=3,s=1: pad 1; k=7: pad 3; k=4,s=2 (even H): pad (1,1)."""
    _, kh, kw, cin, cout = kernel.shape
    B, Fr, H, W, C = x.shape
    assert C == cin
    xx = x.reshape(B * Fr, H, W, C).permute(0, 3, 1, 2)
    w = kernel[0].permute(3, 2, 0, 1)  # (out,in,kh,kw); cross-correlation, no flip
    if stride == 1:
        pad = (kh - 1) // 2
    else:
        assert kh == 4 and stride == 2 and H % 2 == 0
        pad = 1
    y = F.conv2d(xx, w, bias, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).reshape(B, Fr, y.shape[2], y.shape[3], cout)


def conv_transpose_k4s2(x: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """nnx.ConvTranspose(dim,dim,(1,4,4),(1,2,2)) SAME, transpose_kernel=False (utils.py:113):
    conv_general_dilated(lhs_dilation=2, padding=(2,2), unflipped kernel). Equivalent torch call:
    conv_transpose2d with the spatially FLIPPED kernel, stride 2, padding 1."""
    _, kh, kw, cin, cout = kernel.shape
    B, Fr, H, W, C = x.shape
    xx = x.reshape(B * Fr, H, W, C).permute(0, 3, 1, 2)
    wt = kernel[0].flip(0, 1).permute(2, 3, 0, 1)  # (in,out,kh,kw)
    y = F.conv_transpose2d(xx, wt, bias, stride=2, padding=1)
    return y.permute(0, 2, 3, 1).reshape(B, Fr, 2 * H, 2 * W, cout)


def conv1x1(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """nnx.Conv(kernel_size=1): kernel (1,in,out); pointwise linear over the last axis."""
    y = x @ kernel[0]
    return y if bias is None else y + bias


def group_norm(x: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, groups: int = GROUPS) -> torch.Tensor:
    """nnx.GroupNorm(C, num_groups=8): statistics per (batch, group) over (F,H,W,C/groups),
    fast variance max(0, E[x^2]-E[x]^2), eps 1e-6. modules.py:167."""
    B = x.shape[0]
    C = x.shape[-1]
    g = x.reshape(B, -1, groups, C // groups)
    mean = g.mean(dim=(1, 3), keepdim=True)
    var = ((g * g).mean(dim=(1, 3), keepdim=True) - mean * mean).clamp_min(0.0)
    y = (g - mean) * torch.rsqrt(var + EPS)
    return y.reshape(x.shape) * scale + bias


def layer_norm(x: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """nnx.LayerNorm over the last axis, fast variance, eps 1e-6. modules.py:208,223."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x * x).mean(dim=-1, keepdim=True) - mean * mean).clamp_min(0.0)
    return (x - mean) * torch.rsqrt(var + EPS) * scale + bias


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """nnx.gelu == jax.nn.gelu(approximate=True). unet3d.py:131."""
    return F.gelu(x, approximate="tanh")


# --------------------------------------------------------------------------------------
# modules.py
# --------------------------------------------------------------------------------------
def sinusoidal_pos_emb(t: torch.Tensor, dim: int, dtype) -> torch.Tensor:
    """modules.py:35-45."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=dtype) * -e)
    ang = t.to(dtype)[:, None] * freq[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=-1)


def relative_position_bucket(rel: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """modules.py:351-378 (integer work; must be bit-exact). __call__ always uses the defaults 32/128
    (modules.py:386), whatever the constructor was given."""
    n = -rel
    num_buckets //= 2
    ret = (n < 0).to(torch.int32) * num_buckets
    n = n.abs()
    max_exact = num_buckets // 2
    is_small = n < max_exact
    # float32 log, as jnp does with x64 disabled; n == 0 gives -inf -> int cast is masked by is_small
    nf = n.to(torch.float32).clamp_min(1.0)
    val_large = max_exact + (torch.log(nf / max_exact) / math.log(max_distance / max_exact)
                             * (num_buckets - max_exact)).to(torch.int32)
    val_large = torch.minimum(val_large, torch.full_like(val_large, num_buckets - 1))
    return ret + torch.where(is_small, n.to(torch.int32), val_large)


def relative_position_bias(p: Params, prefix: str, n: int) -> torch.Tensor:
    """modules.py:380-390 -> (heads, n, n)."""
    pos = torch.arange(n, dtype=torch.int32)
    rel = pos[:, None] - pos[None, :]
    buckets = relative_position_bucket(rel)
    emb = p[prefix + ".relative_attention_bias.embedding"][buckets.long()]  # (n,n,heads)
    return emb.permute(2, 0, 1)


def block(p: Params, prefix: str, x: torch.Tensor, scale_shift=None, groups: int = GROUPS) -> torch.Tensor:
    """modules.py:171-179."""
    x = conv_khw(x, p[prefix + ".proj.kernel"], p[prefix + ".proj.bias"])
    x = group_norm(x, p[prefix + ".norm.scale"], p[prefix + ".norm.bias"], groups)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(p: Params, prefix: str, x: torch.Tensor, t_emb: Optional[torch.Tensor], groups: int = GROUPS) -> torch.Tensor:
    """modules.py:226-243."""
    scale_shift = None
    if (prefix + ".mlp.layers.1.kernel") in p:
        assert t_emb is not None
        e = F.silu(t_emb) @ p[prefix + ".mlp.layers.1.kernel"] + p[prefix + ".mlp.layers.1.bias"]
        e = layer_norm(e, p[prefix + ".norm_1.scale"], p[prefix + ".norm_1.bias"])
        e = e[:, None, None, None, :]
        scale_shift = torch.chunk(e, 2, dim=-1)
    h = block(p, prefix + ".block_1", x, scale_shift, groups)
    h = block(p, prefix + ".block_2", h, None, groups)
    if (prefix + ".res_conv.kernel") in p:
        s = conv1x1(x, p[prefix + ".res_conv.kernel"], p[prefix + ".res_conv.bias"])
    else:
        s = x
    return h + layer_norm(s, p[prefix + ".norm_2.scale"], p[prefix + ".norm_2.bias"])


def spatial_linear_attention(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """modules.py:94-129. q: softmax over the D features, NOT scaled (the scaled copy is dead code);
    k: softmax over the N = H*W tokens."""
    B, Fr, H, W, C = x.shape
    xx = x.reshape(B * Fr, H * W, C)
    q = (xx @ p[prefix + ".q.kernel"][0]).reshape(B * Fr, H * W, HEADS, SLA_D)
    k = (xx @ p[prefix + ".k.kernel"][0]).reshape(B * Fr, H * W, HEADS, SLA_D)
    v = (xx @ p[prefix + ".v.kernel"][0]).reshape(B * Fr, H * W, HEADS, SLA_D)
    q = q.softmax(dim=-1)
    k = k.softmax(dim=1)
    ctx = torch.einsum("bnhd,bnhe->bhde", k, v)
    out = torch.einsum("bhde,bnhd->bnhe", ctx, q).reshape(B * Fr, H * W, HEADS * SLA_D)
    out = out @ p[prefix + ".to_out.kernel"][0]
    return out.reshape(B, Fr, H, W, C)


def multihead_attention(p: Params, prefix: str, x: torch.Tensor, focus_present_mask=None, pos_bias=None):
    """modules.py:280-326 on x (..., S, C). Mask and bias are applied AFTER the softmax (C3)."""
    q = torch.einsum("...c,chd->...hd", x, p[prefix + ".q.kernel"]) + p[prefix + ".q.bias"]
    k = torch.einsum("...c,chd->...hd", x, p[prefix + ".k.kernel"]) + p[prefix + ".k.bias"]
    v = torch.einsum("...c,chd->...hd", x, p[prefix + ".v.kernel"]) + p[prefix + ".v.bias"]
    S = x.shape[-2]
    if focus_present_mask is not None and bool(focus_present_mask.all()):
        return torch.einsum("...hd,hdc->...c", v, p[prefix + ".out.kernel"]) + p[prefix + ".out.bias"]
    q = q / DIM_HEAD ** 0.5
    qk = torch.einsum("...ihd,...jhd->...hij", q, k)
    attn = qk.softmax(dim=-1)
    if focus_present_mask is not None and bool(focus_present_mask.any()):
        eye = torch.eye(S, dtype=torch.bool)
        mask = torch.where(focus_present_mask.reshape(-1, 1, 1, 1, 1, 1), eye, torch.ones_like(eye))
        attn = torch.where(mask, attn, torch.full_like(attn, torch.finfo(torch.float32).min))
    if pos_bias is not None:
        attn = attn + pos_bias
    o = torch.einsum("...hij,...jhd->...ihd", attn, v)
    return torch.einsum("...hd,hdc->...c", o, p[prefix + ".out.kernel"]) + p[prefix + ".out.bias"]


def temporal_attention(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Residual(PreNorm(EinopsToAndFrom('b f h w c','b (h w) f c', MHA))) with the PreNorm bug:
    f(x) + x, un-normalised input, no mask / bias (unet3d.py:86-96,118-120; modules.py:146-148)."""
    B, Fr, H, W, C = x.shape
    xs = x.permute(0, 2, 3, 1, 4).reshape(B, H * W, Fr, C)
    o = multihead_attention(p, prefix + ".fn.fn.fn", xs)
    return o.reshape(B, H, W, Fr, C).permute(0, 3, 1, 2, 4) + x


def spatial_attention(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """mid_spatial_attn: 'b f h w c -> b f (h w) c' (unet3d.py:196-205)."""
    B, Fr, H, W, C = x.shape
    o = multihead_attention(p, prefix + ".fn.fn.fn", x.reshape(B, Fr, H * W, C))
    return o.reshape(B, Fr, H, W, C) + x


def sla_residual(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return spatial_linear_attention(p, prefix + ".fn.fn", x) + x


# --------------------------------------------------------------------------------------
# unet3d.py
# --------------------------------------------------------------------------------------
def unet3d_forward(p: Params, x: torch.Tensor, time: torch.Tensor, dim: int, dim_mults=(1, 2, 4, 8),
                   resnet_groups: int = GROUPS, use_sparse_linear_attn: bool = True) -> torch.Tensor:
    """unet3d.py:262-387, unconditional (has_cond False). x (B,C,F,H,W), time (B,) int -> (B,F,H,W,C).
    resnet_groups -> every GroupNorm (unet3d.py:156); use_sparse_linear_attn=False puts Identity() in the spatial
    attention slots (unet3d.py:179-181,230-231)."""
    dtype = x.dtype

    def res(pp, prefix, xx, tt):
        return resnet_block(pp, prefix, xx, tt, resnet_groups)

    def sla(pp, prefix, xx):
        return sla_residual(pp, prefix, xx) if use_sparse_linear_attn else xx

    n_res = len(dim_mults)
    x = x.permute(0, 2, 3, 4, 1)  # :280
    x = conv_khw(x, p["init_conv.kernel"], p["init_conv.bias"])  # :282
    x = temporal_attention(p, "init_temporal_attn", x)  # :284
    r = x
    t = sinusoidal_pos_emb(time, dim, dtype)  # :288 / :128-133
    t = t @ p["time_mlp.layers.1.kernel"] + p["time_mlp.layers.1.bias"]
    t = gelu_tanh(t)
    t = t @ p["time_mlp.layers.3.kernel"] + p["time_mlp.layers.3.bias"]
    hs = []
    for l in range(n_res):  # :303-314
        x = res(p, f"downs.{l}.0", x, t)
        x = res(p, f"downs.{l}.1", x, t)
        x = sla(p, f"downs.{l}.2", x)
        x = temporal_attention(p, f"downs.{l}.3", x)
        hs.append(x)
        if l < n_res - 1:
            x = conv_khw(x, p[f"downs.{l}.4.kernel"], p[f"downs.{l}.4.bias"], stride=2)
    x = res(p, "mid_block1", x, t)  # :320
    x = spatial_attention(p, "mid_spatial_attn", x)  # :324
    x = temporal_attention(p, "mid_temporal_attn", x)  # :328
    x = res(p, "mid_block2", x, t)  # :334
    for i in range(n_res):  # :337-370
        x = torch.cat([x, hs.pop()], dim=-1)
        x = res(p, f"ups.{i}.0", x, t)
        x = res(p, f"ups.{i}.1", x, t)
        x = sla(p, f"ups.{i}.2", x)
        x = temporal_attention(p, f"ups.{i}.3", x)
        if i < n_res - 1:
            x = conv_transpose_k4s2(x, p[f"ups.{i}.4.kernel"], p[f"ups.{i}.4.bias"])
    x = torch.cat([x, r], dim=-1)  # :377
    x = res(p, "final_conv.layers.0", x, None)  # :250 (no time embedding)
    return conv1x1(x, p["final_conv.layers.1.kernel"], p["final_conv.layers.1.bias"])  # :251


# --------------------------------------------------------------------------------------
# parameter construction (flax init distributions; NOT bit-identical to nnx.Rngs(0))
# --------------------------------------------------------------------------------------
def param_shapes(dim: int, channels: int, dim_mults=(1, 2, 4, 8), init_kernel_size: int = 7,
                 use_sparse_linear_attn: bool = True) -> Dict[str, tuple]:
    """Names and shapes of the Unet3D state (unet3d.py:58-252), in creation order."""
    s: Dict[str, tuple] = {}
    time_dim = dim * 4
    hd = HEADS * DIM_HEAD

    def mha(prefix, c):
        s[prefix + ".fn.norm.scale"] = (c,)
        s[prefix + ".fn.norm.bias"] = (c,)
        for n in ("q", "k", "v"):
            s[f"{prefix}.fn.fn.fn.{n}.kernel"] = (c, HEADS, DIM_HEAD)
            s[f"{prefix}.fn.fn.fn.{n}.bias"] = (HEADS, DIM_HEAD)
        s[prefix + ".fn.fn.fn.out.kernel"] = (HEADS, DIM_HEAD, c)
        s[prefix + ".fn.fn.fn.out.bias"] = (c,)

    def sla(prefix, c):
        if not use_sparse_linear_attn:
            return
        s[prefix + ".fn.norm.scale"] = (c,)
        s[prefix + ".fn.norm.bias"] = (c,)
        for n in ("q", "k", "v"):
            s[f"{prefix}.fn.fn.{n}.kernel"] = (1, c, hd)
        s[prefix + ".fn.fn.to_out.kernel"] = (1, hd, c)

    def resnet(prefix, cin, cout, with_time=True):
        if with_time:
            s[prefix + ".mlp.layers.1.kernel"] = (time_dim, 2 * cout)
            s[prefix + ".mlp.layers.1.bias"] = (2 * cout,)
        s[prefix + ".norm_1.scale"] = (2 * cout,)
        s[prefix + ".norm_1.bias"] = (2 * cout,)
        for b, ci in (("block_1", cin), ("block_2", cout)):
            s[f"{prefix}.{b}.proj.kernel"] = (1, 3, 3, ci, cout)
            s[f"{prefix}.{b}.proj.bias"] = (cout,)
            s[f"{prefix}.{b}.norm.scale"] = (cout,)
            s[f"{prefix}.{b}.norm.bias"] = (cout,)
        if cin != cout:
            s[prefix + ".res_conv.kernel"] = (1, cin, cout)
            s[prefix + ".res_conv.bias"] = (cout,)
        s[prefix + ".norm_2.scale"] = (cout,)
        s[prefix + ".norm_2.bias"] = (cout,)

    s["time_rel_pos_bias.relative_attention_bias.embedding"] = (32, HEADS)
    k = init_kernel_size
    s["init_conv.kernel"] = (1, k, k, channels, dim)
    s["init_conv.bias"] = (dim,)
    mha("init_temporal_attn", dim)
    s["time_mlp.layers.1.kernel"] = (dim, time_dim)
    s["time_mlp.layers.1.bias"] = (time_dim,)
    s["time_mlp.layers.3.kernel"] = (time_dim, time_dim)
    s["time_mlp.layers.3.bias"] = (time_dim,)
    dims = [dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    n_res = len(in_out)
    for l, (ci, co) in enumerate(in_out):
        resnet(f"downs.{l}.0", ci, co)
        resnet(f"downs.{l}.1", co, co)
        sla(f"downs.{l}.2", co)
        mha(f"downs.{l}.3", co)
        if l < n_res - 1:
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
        if i < n_res - 1:
            s[f"ups.{i}.4.kernel"] = (1, 4, 4, ci, ci)
            s[f"ups.{i}.4.bias"] = (ci,)
    resnet("final_conv.layers.0", dim * 2, dim, with_time=False)
    s["final_conv.layers.1.kernel"] = (1, dim, channels)
    s["final_conv.layers.1.bias"] = (channels,)
    return s


def init_params(dim: int, channels: int, seed: int = 3, dtype=torch.float32, perturb: float = 0.0,
                dim_mults=(1, 2, 4, 8), use_sparse_linear_attn: bool = True) -> Params:
    """flax default initialisers: kernels lecun_normal (truncated normal, std sqrt(1/fan_in)),
    biases 0, norm scale 1 / bias 0, embedding normal(std 1/sqrt(features))... `perturb` > 0 adds
    N(0, perturb) to biases and norm parameters so that parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for name, shape in param_shapes(dim, channels, dim_mults, use_sparse_linear_attn=use_sparse_linear_attn).items():
        leaf = name.rsplit(".", 1)[1]
        if leaf == "kernel":
            if ".out.kernel" in name and len(shape) == 3 and shape[0] == HEADS:
                fan_in = shape[0] * shape[1]
            elif len(shape) == 3 and shape[1] == HEADS:  # LinearGeneral q/k/v (in, heads, dim)
                fan_in = shape[0]
            else:
                fan_in = 1
                for d in shape[:-1]:
                    fan_in *= d
            std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
            w = torch.empty(shape, dtype=torch.float64)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            p[name] = (w * std).to(dtype)
        elif leaf == "embedding":
            p[name] = (torch.randn(shape, generator=g, dtype=torch.float64) / math.sqrt(shape[0])).to(dtype)
        elif leaf == "scale":
            p[name] = torch.ones(shape, dtype=dtype)
            if perturb:
                p[name] = p[name] + perturb * torch.randn(shape, generator=g, dtype=torch.float64).to(dtype)
        else:  # bias
            p[name] = torch.zeros(shape, dtype=dtype)
            if perturb:
                p[name] = p[name] + perturb * torch.randn(shape, generator=g, dtype=torch.float64).to(dtype)
    return p
