"""Thin Python wrappers over the C ABI (include/vdn.h). Tensors are torch CUDA tensors used
purely as device-memory handles; all arithmetic happens in libvdn.so kernels."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import VDN_BF16, VDN_F32, VDN_TAP_DOWN, VDN_TAP_UNIT, VDN_TAP_UP, TapGemmDesc, check, lib, ptr, stream_ptr

GN_REPLICAS = 16  # kGnReplicas in csrc/vdn_common.cuh: GroupNorm partial sums are [R][B][G][2]
TAPS_1x1 = [(0, 0)]
TAPS_3x3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]  # kernel (kh,kw) row-major, SAME padding
TAPS_4x4 = [(ky, kx) for ky in range(4) for kx in range(4)]  # VDN_TAP_DOWN: kernel indices


def up_class_taps(py: int, px: int):
    """Taps of output-parity class (py,px) of the k=4,s=2 transposed conv (utils.py:113).

    out[2y+py, 2x+px] = sum_{a in {py,py+2}, b in {px,px+2}} x[y + (a+py)/2 - 1, x + (b+px)/2 - 1] w[a,b]
    Returns (shifts, kernel_tap_indices)."""
    shifts, kidx = [], []
    for a in (py, py + 2):
        for b in (px, px + 2):
            shifts.append(((a + py) // 2 - 1, (b + px) // 2 - 1))
            kidx.append(a * 4 + b)
    return shifts, kidx


def pack_weight(src: torch.Tensor, dst: torch.Tensor, taps: int, cin: int, cout: int, mode: int,
                perm: Optional[Sequence[int]] = None, n_off: int = 0, k_off: int = 0) -> None:
    """src fp32 [taps][cin][cout] (reference layout) -> dst bf16 [rows][ld] K-major operand."""
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.is_contiguous() and dst.is_contiguous()
    assert src.numel() >= taps * cin * cout
    ld = dst.shape[-1]
    p = None
    if perm is not None:
        p = (C.c_int * len(perm))(*perm)
    check(lib.vdn_pack_weight(ptr(src), ptr(dst), taps, cin, cout, mode, p, ld, n_off, k_off, stream_ptr()),
          "vdn_pack_weight")


# timing experiments only (tools/whatif.sh, which opts in with VDN_DEBUG=1): VDN_SKIP="gn_bwd,wgrad,..." turns the
# named wrappers into no-ops so that the step-time contribution of a kernel family can be read off (results are
# wrong, of course). Empty in every product / test / bench process.
_SKIP = set(filter(None, (_lib.host_flag("VDN_SKIP", "") or "").split(",")))


def tapgemm(kind: int, srcs: Sequence[torch.Tensor], wp: torch.Tensor, taps, *, bias=None, residual=None,
            residual2=None, out=None, out2=None, split_col: int = 0, gn_sums=None, gn_groups: int = 0, rows_per_sample: int = 0,
            py: int = 0, px: int = 0, out_dtype=torch.bfloat16, ref: bool = False, workspace=None) -> torch.Tensor:
    """`workspace`: optional scratch tensor of >= tapgemm_workspace_bytes(...) bytes (vdn_tapgemm_ws): lets small-M
    launches split their K loop over thread-block clusters. Not to be shared by launches that may run concurrently."""
    x0 = srcs[0]
    assert x0.dtype == torch.bfloat16 and x0.is_contiguous() and x0.dim() == 4
    n_img, Hs, Ws, Csrc = x0.shape
    for s in srcs[1:]:
        assert s.shape == x0.shape and s.dtype == x0.dtype and s.is_contiguous()
    H, W = (Hs // 2, Ws // 2) if kind == VDN_TAP_DOWN else (Hs, Ws)
    n_out = wp.shape[0]
    assert wp.dtype == torch.bfloat16 and wp.is_contiguous() and wp.shape[1] == len(taps) * len(srcs) * Csrc
    d = TapGemmDesc()
    d.kind, d.n_img, d.H, d.W = kind, n_img, H, W
    d.n_src, d.src_c, d.n_taps = len(srcs), Csrc, len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.n_out, d.py, d.px = n_out, py, px
    d.out_dtype = VDN_F32 if out_dtype == torch.float32 else VDN_BF16
    d.split_col = split_col
    d.gn_groups = gn_groups if gn_sums is not None else 0
    d.rows_per_sample = rows_per_sample
    if out is None:
        oh, ow = (2 * H, 2 * W) if kind == VDN_TAP_UP else (H, W)
        out = torch.empty((n_img, oh, ow, split_col if split_col else n_out), dtype=out_dtype, device=x0.device)
    if workspace is not None and not ref:
        check(lib.vdn_tapgemm_ws(C.byref(d), ptr(x0), ptr(srcs[1]) if len(srcs) > 1 else None, ptr(wp), ptr(bias),
                                 ptr(residual), ptr(residual2), ptr(out), ptr(out2), ptr(gn_sums), ptr(workspace),
                                 workspace.numel() * workspace.element_size(), stream_ptr()), "vdn_tapgemm_ws")
        return out
    fn = lib.vdn_tapgemm_ref if ref else lib.vdn_tapgemm
    check(fn(C.byref(d), ptr(x0), ptr(srcs[1]) if len(srcs) > 1 else None, ptr(wp), ptr(bias), ptr(residual),
             ptr(residual2), ptr(out), ptr(out2), ptr(gn_sums), stream_ptr()), "vdn_tapgemm")
    return out


def tapgemm_workspace_bytes(kind: int, n_img: int, H: int, W: int, n_src: int, src_c: int, taps, n_out: int, *,
                            split_col: int = 0, gn_groups: int = 0, rows_per_sample: int = 0) -> int:
    """Scratch bytes vdn_tapgemm_ws wants for this launch (0: the launch does not split its K loop)."""
    d = TapGemmDesc()
    d.kind, d.n_img, d.H, d.W = kind, n_img, H, W
    d.n_src, d.src_c, d.n_taps = n_src, src_c, len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.n_out, d.split_col, d.gn_groups, d.rows_per_sample = n_out, split_col, gn_groups, rows_per_sample
    return int(lib.vdn_tapgemm_workspace(C.byref(d)))


# ------------------------------------------------------------------------------------------
# wgrad
# ------------------------------------------------------------------------------------------
def wgrad(kind: int, srcs: Sequence[torch.Tensor], g: torch.Tensor, dw: torch.Tensor, taps, ref: bool = False,
          dbias: Optional[torch.Tensor] = None) -> None:
    """dw[tap][n_src*C][Cout] (fp32, reference kernel layout) += sum_pixels src[p + tap]^T g[p]."""
    if "wgrad" in _SKIP:
        return
    x0 = srcs[0]
    assert x0.dtype == torch.bfloat16 and g.dtype == torch.bfloat16 and dw.dtype == torch.float32
    assert x0.is_contiguous() and g.is_contiguous() and dw.is_contiguous()
    n_img, Hs, Ws, Cs = x0.shape
    cout = g.shape[-1]
    if kind == VDN_TAP_DOWN:
        H, W = Hs // 2, Ws // 2
        assert g.shape[1] == H and g.shape[2] == W
    elif kind == VDN_TAP_UP:
        H, W = Hs, Ws
        assert g.shape[1] == 2 * H and g.shape[2] == 2 * W
    else:
        H, W = Hs, Ws
        assert g.shape[1] == H and g.shape[2] == W
    assert dw.numel() == len(taps) * len(srcs) * Cs * cout
    dy = (C.c_int * len(taps))(*[t[0] for t in taps])
    dx = (C.c_int * len(taps))(*[t[1] for t in taps])
    if dbias is not None and not ref:
        check(lib.vdn_wgrad_bias(kind, ptr(x0), ptr(srcs[1]) if len(srcs) > 1 else None, ptr(g), ptr(dw), ptr(dbias),
                                 n_img, H, W, len(srcs), Cs, cout, len(taps), dy, dx, stream_ptr()), "vdn_wgrad_bias")
        return
    fn = lib.vdn_wgrad_ref if ref else lib.vdn_wgrad
    check(fn(kind, ptr(x0), ptr(srcs[1]) if len(srcs) > 1 else None, ptr(g), ptr(dw), n_img, H, W, len(srcs), Cs,
             cout, len(taps), dy, dx, stream_ptr()), "vdn_wgrad")


# ------------------------------------------------------------------------------------------
# norms
# ------------------------------------------------------------------------------------------
def gn_silu_fwd(x_raw, sums, gamma, beta, ss, out, B, rows, Cc, G=8):
    if "gn_fwd" in _SKIP:
        return
    ss_ld = ss.stride(0) if ss is not None else 0
    check(lib.vdn_gn_silu_fwd(ptr(x_raw), ptr(sums), ptr(gamma), ptr(beta), ptr(ss), ss_ld, ptr(out), B, rows, Cc, G,
                              stream_ptr()), "vdn_gn_silu_fwd")


def resblock_tail_fwd(b_raw, sums, gamma, beta, s, ln_g, ln_b, out, B, rows, Cc, G=8):
    if "tail_fwd" in _SKIP:
        return
    check(lib.vdn_resblock_tail_fwd(ptr(b_raw), ptr(sums), ptr(gamma), ptr(beta), ptr(s), ptr(ln_g), ptr(ln_b),
                                    ptr(out), B, rows, Cc, G, stream_ptr()), "vdn_resblock_tail_fwd")


def gn_silu_bwd(dy, x_raw, sums, gamma, beta, ss, T_ws, dx_raw, dgamma, dbeta, dss, B, rows, Cc, G=8, dconv_bias=None,
                prezeroed=False):
    """prezeroed: T_ws was zeroed by the caller (vdn_gn_silu_bwd_acc: no memset node in front of the kernels)."""
    if "gn_bwd" in _SKIP:
        return
    ss_ld = ss.stride(0) if ss is not None else 0
    dss_ld = dss.stride(0) if dss is not None else 0
    fn = lib.vdn_gn_silu_bwd_acc if prezeroed else lib.vdn_gn_silu_bwd
    check(fn(ptr(dy), ptr(x_raw), ptr(sums), ptr(gamma), ptr(beta), ptr(ss), ss_ld, ptr(T_ws),
                              ptr(dx_raw), ptr(dgamma), ptr(dbeta), ptr(dss), dss_ld, ptr(dconv_bias), B, rows, Cc, G, stream_ptr()),
          "vdn_gn_silu_bwd")


def ln_bwd(s, dy, ln_g, ds, dg, db, P, Cc):
    if "ln_bwd" in _SKIP:
        return
    check(lib.vdn_ln_bwd(ptr(s), ptr(dy), ptr(ln_g), ptr(ds), ptr(dg), ptr(db), C.c_long(P), Cc, stream_ptr()),
          "vdn_ln_bwd")


# ------------------------------------------------------------------------------------------
# attention cores
# ------------------------------------------------------------------------------------------
MHA_TEMPORAL, MHA_SPATIAL = 0, 1


def mha_core_fwd(qkv, o, lse, mode, B, F, HW):
    check(lib.vdn_mha_core_fwd(ptr(qkv), ptr(o), ptr(lse), mode, B, F, HW, stream_ptr()), "vdn_mha_core_fwd")


def mha_core_bwd(qkv, o, d_o, lse, D_ws, dqkv, mode, B, F, HW):
    check(lib.vdn_mha_core_bwd(ptr(qkv), ptr(o), ptr(d_o), ptr(lse), ptr(D_ws), ptr(dqkv), mode, B, F, HW,
                               stream_ptr()), "vdn_mha_core_bwd")


def sla_workspace_floats(n_img, N) -> int:
    return int(lib.vdn_sla_workspace_floats(n_img, N))


def sla_core_fwd(qkv, tok_out, ctx, kstat, ws, n_img, N):
    if "sla_fwd" in _SKIP:
        return
    check(lib.vdn_sla_core_fwd(ptr(qkv), ptr(tok_out), ptr(ctx), ptr(kstat), ptr(ws), n_img, N, stream_ptr()),
          "vdn_sla_core_fwd")


def sla_fused_fwd(x, w_qkv, w_out, out, ctx, kstat, ws, n_img, N, Cc):
    check(lib.vdn_sla_fused_fwd(ptr(x), ptr(w_qkv), ptr(w_out), ptr(out), ptr(ctx), ptr(kstat), ptr(ws), n_img, N, Cc,
                                stream_ptr()), "vdn_sla_fused_fwd")


def sla_core_bwd(qkv, d_tok, ctx, kstat, dctx, dqkv, n_img, N, prezeroed=False):
    if "sla_bwd" in _SKIP:
        return
    fn = lib.vdn_sla_core_bwd_acc if prezeroed else lib.vdn_sla_core_bwd
    check(fn(ptr(qkv), ptr(d_tok), ptr(ctx), ptr(kstat), ptr(dctx), ptr(dqkv), n_img, N,
                               stream_ptr()), "vdn_sla_core_bwd")


# ------------------------------------------------------------------------------------------
# small kernels
# ------------------------------------------------------------------------------------------
def init_conv_fwd(x, w, bias, out, B, Cin, F, H, W, Cout, ks):
    check(lib.vdn_init_conv_fwd(ptr(x), ptr(w), ptr(bias), ptr(out), B, Cin, F, H, W, Cout, ks, stream_ptr()),
          "vdn_init_conv_fwd")


def init_conv_wgrad(x, dy, dw, dbias, B, Cin, F, H, W, Cout, ks):
    check(lib.vdn_init_conv_wgrad(ptr(x), ptr(dy), ptr(dw), ptr(dbias), B, Cin, F, H, W, Cout, ks, stream_ptr()),
          "vdn_init_conv_wgrad")


def final_conv_fwd(h, w, bias, out, P, Cc, Co):
    check(lib.vdn_final_conv_fwd(ptr(h), ptr(w), ptr(bias), ptr(out), C.c_long(P), Cc, Co, stream_ptr()),
          "vdn_final_conv_fwd")


def final_conv_bwd(h, dout, w, dh, dw, db, P, Cc, Co):
    check(lib.vdn_final_conv_bwd(ptr(h), ptr(dout), ptr(w), ptr(dh), ptr(dw), ptr(db), C.c_long(P), Cc, Co,
                                 stream_ptr()), "vdn_final_conv_bwd")


def time_mlp_fwd(time, w1, b1, w2, b2, emb, h1, t_out, B, dim):
    check(lib.vdn_time_mlp_fwd(ptr(time), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(emb), ptr(h1), ptr(t_out), B, dim,
                               stream_ptr()), "vdn_time_mlp_fwd")


def time_mlp_bwd(dt, emb, h1, w2, dw1, db1, dw2, db2, dh1_ws, B, dim):
    check(lib.vdn_time_mlp_bwd(ptr(dt), ptr(emb), ptr(h1), ptr(w2), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2),
                               ptr(dh1_ws), B, dim, stream_ptr()), "vdn_time_mlp_bwd")


class TimeHead(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("ln_g", C.c_void_p), ("ln_b", C.c_void_p),
                ("dw", C.c_void_p), ("db", C.c_void_p), ("dln_g", C.c_void_p), ("dln_b", C.c_void_p),
                ("n_out", C.c_int), ("off", C.c_int)]


def make_time_head_table(entries, device) -> torch.Tensor:
    """entries: list of dicts with tensors w,b,ln_g,ln_b,(dw,db,dln_g,dln_b) and ints n_out, off.
    Returns a device uint8 tensor holding the vdn_time_head array."""
    arr = (TimeHead * len(entries))()
    for i, e in enumerate(entries):
        for k in ("w", "b", "ln_g", "ln_b", "dw", "db", "dln_g", "dln_b"):
            t = e.get(k)
            setattr(arr[i], k, t.data_ptr() if t is not None else None)
        arr[i].n_out, arr[i].off = e["n_out"], e["off"]
    raw = bytes(arr)
    host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    return host.to(device)


def time_heads_fwd(t, table, n_heads, e_pre, ss, B, td):
    check(lib.vdn_time_heads_fwd(ptr(t), ptr(table), n_heads, ptr(e_pre), ptr(ss), ss.stride(0), B, td, stream_ptr()),
          "vdn_time_heads_fwd")


def time_heads_bwd(t, table, n_heads, e_pre, dss, de_ws, dt, B, td):
    check(lib.vdn_time_heads_bwd(ptr(t), ptr(table), n_heads, ptr(e_pre), ptr(dss), dss.stride(0), ptr(de_ws),
                                 ptr(dt), B, td, stream_ptr()), "vdn_time_heads_bwd")


def q_sample(x_start, noise, t, sqrt_ac, sqrt_1mac, out, normalize: bool):
    B = x_start.shape[0]
    check(lib.vdn_q_sample(ptr(x_start), ptr(noise), ptr(t), ptr(sqrt_ac), ptr(sqrt_1mac), ptr(out), B,
                           C.c_long(x_start.numel() // B), int(normalize), stream_ptr()), "vdn_q_sample")


def loss_fwd_bwd(pred, noise, loss, dpred, B, Cc, FHW, l1: bool):
    check(lib.vdn_loss(ptr(pred), ptr(noise), ptr(loss), ptr(dpred), B, Cc, C.c_long(FHW), int(l1), stream_ptr()),
          "vdn_loss")


def p_sample(x, eps, z, t, recip, recipm1, coef1, coef2, logvar, out, B, Cc, FHW, clip: bool = True):
    check(lib.vdn_p_sample(ptr(x), ptr(eps), ptr(z), ptr(t), ptr(recip), ptr(recipm1), ptr(coef1), ptr(coef2),
                           ptr(logvar), ptr(out), B, Cc, C.c_long(FHW), int(clip), stream_ptr()), "vdn_p_sample")


def colsum(dy, db, P, Cc):
    check(lib.vdn_colsum(ptr(dy), ptr(db), C.c_long(P), Cc, stream_ptr()), "vdn_colsum")


def add_bf16(a, b, out):
    check(lib.vdn_add_bf16(ptr(a), ptr(b), ptr(out), C.c_long(a.numel()), stream_ptr()), "vdn_add_bf16")


def adam_ema(p, g, m, v, ema, hp, sqnorm=None):
    """Fused optax.adam + EMA over the flat state; with `sqnorm` (device scalar from grad_sqnorm) the global-norm
    clip of utils.py:127-152 is applied first (hp[9] = max_grad_norm, hp[10] = epsilon)."""
    if sqnorm is None:
        check(lib.vdn_adam_ema(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(hp), p.numel(), stream_ptr()), "vdn_adam_ema")
    else:
        check(lib.vdn_adam_ema_clip(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(hp), ptr(sqnorm), p.numel(),
                                    stream_ptr()), "vdn_adam_ema_clip")


def grad_sqnorm(g, out):
    """out[0] = sum g^2 (fp32 device scalar)."""
    check(lib.vdn_grad_sqnorm(ptr(g), g.numel(), ptr(out), stream_ptr()), "vdn_grad_sqnorm")


def randn(out, seed: int, subseq: int = 0, elem_offset: int = 0):
    """Fill `out` (fp32) with N(0,1) draws: element i = Philox(seed, subseq, elem_offset + i)."""
    check(lib.vdn_randn(ptr(out), C.c_long(out.numel()), C.c_ulonglong(seed & (2 ** 64 - 1)), C.c_ulonglong(subseq),
                        C.c_ulonglong(elem_offset), stream_ptr()), "vdn_randn")


def randn_t(out, params_dev, t_dev):
    """As randn() with (seed, subseq_add, elem_offset) = params_dev[0..2] (int64 device tensor) and
    subseq = subseq_add + t_dev[0], all read on the device (graph-capturable countdown)."""
    check(lib.vdn_randn_t(ptr(out), C.c_long(out.numel()), ptr(params_dev), ptr(t_dev), stream_ptr()), "vdn_randn_t")


def countdown(t_dev):
    check(lib.vdn_countdown(ptr(t_dev), t_dev.numel(), stream_ptr()), "vdn_countdown")


class PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("taps", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
                ("mode", C.c_int), ("ld", C.c_int), ("n_off", C.c_int), ("k_off", C.c_int), ("perm", C.c_int * 16),
                ("begin", C.c_longlong)]


def make_pack_table(jobs, device):
    """jobs: list of (src fp32 tensor, dst bf16 2-D tensor, taps, cin, cout, mode, perm|None). Returns
    (device table, n_jobs, total 32x32 tiles) for pack_batched()."""
    arr = (PackJob * len(jobs))()
    begin = 0
    for i, (src, dst, taps, cin, cout, mode, perm) in enumerate(jobs):
        a = arr[i]
        a.src, a.dst = src.data_ptr(), dst.data_ptr()
        a.taps, a.cin, a.cout, a.mode, a.ld, a.n_off, a.k_off = taps, cin, cout, mode, dst.shape[-1], 0, 0
        for t in range(16):
            a.perm[t] = perm[t] if (perm is not None and t < len(perm)) else t
        a.begin = begin
        begin += taps * ((cin + 31) // 32) * ((cout + 31) // 32)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device), len(jobs), begin


def pack_batched(table, n_jobs: int, total: int):
    check(lib.vdn_pack_batched(ptr(table), n_jobs, C.c_longlong(total), stream_ptr()), "vdn_pack_batched")


def qkv_headmajor_pack(w, bias, dst, bias_dst, Cc):
    check(lib.vdn_qkv_headmajor_pack(ptr(w), ptr(bias), ptr(dst), ptr(bias_dst), Cc, stream_ptr()),
          "vdn_qkv_headmajor_pack")


def mha_temporal_fused_fwd(x, w_hm, bias_hm, o, qkv, lse, B, F, H, W, Cc):
    if "mha_fwd" in _SKIP:
        return
    check(lib.vdn_mha_temporal_fused_fwd(ptr(x), ptr(w_hm), ptr(bias_hm), ptr(o), ptr(qkv), ptr(lse), B, F, H, W, Cc,
                                         stream_ptr()), "vdn_mha_temporal_fused_fwd")


def mha_temporal_bwd(qkv, o, d_o, lse, dqkv, B, F, H, W):
    check(lib.vdn_mha_temporal_bwd(ptr(qkv), ptr(o), ptr(d_o), ptr(lse), ptr(dqkv), B, F, H, W, stream_ptr()),
          "vdn_mha_temporal_bwd")


def mha_fold_pack(w_qkv, b_qkv, w_out, b_out, fa, fu, fm, fb):
    check(lib.vdn_mha_fold_pack(ptr(w_qkv), ptr(b_qkv), ptr(w_out), ptr(b_out), ptr(fa), ptr(fu), ptr(fm), ptr(fb),
                                stream_ptr()), "vdn_mha_fold_pack")


def mha_temporal_core_fwd(qkv, o, lse, B, F, H, W):
    check(lib.vdn_mha_temporal_core_fwd(ptr(qkv), ptr(o), ptr(lse), B, F, H, W, stream_ptr()), "vdn_mha_temporal_core_fwd")


def mha_temporal_folded_fwd(x, fa, fu, fm, fb, out, B, F, H, W, Cc):
    check(lib.vdn_mha_temporal_folded_fwd(ptr(x), ptr(fa), ptr(fu), ptr(fm), ptr(fb), ptr(out), B, F, H, W, Cc,
                                          stream_ptr()), "vdn_mha_temporal_folded_fwd")


def mha_tc_supported(F: int, Cc: int) -> bool:
    return bool(lib.vdn_mha_temporal_tc_supported(F, Cc))


def mha_temporal_tc_fwd(x, w_hm, bias_hm, o, qkv, lse, B, F, H, W, Cc):
    if "mha_fwd" in _SKIP:
        return
    check(lib.vdn_mha_temporal_tc_fwd(ptr(x), ptr(w_hm), ptr(bias_hm), ptr(o), ptr(qkv), ptr(lse), B, F, H, W, Cc,
                                      stream_ptr()), "vdn_mha_temporal_tc_fwd")


def mha_temporal_tc_bwd(qkv, d_o, lse, dqkv, B, F, H, W, dbias=None):
    if "mha_bwd" in _SKIP:
        return
    check(lib.vdn_mha_temporal_tc_bwd(ptr(qkv), ptr(d_o), ptr(lse), ptr(dqkv), ptr(dbias), B, F, H, W, stream_ptr()),
          "vdn_mha_temporal_tc_bwd")
