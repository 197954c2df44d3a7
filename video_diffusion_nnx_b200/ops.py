"""Thin Python wrappers over the C ABI (include/vdn.h). Tensors are torch CUDA tensors used
purely as device-memory handles; all arithmetic happens in libvdn.so kernels."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import VDN_BF16, VDN_F32, VDN_TAP_DOWN, VDN_TAP_UNIT, VDN_TAP_UP, TapGemmDesc, check, lib, ptr, stream_ptr

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


def tapgemm(kind: int, srcs: Sequence[torch.Tensor], wp: torch.Tensor, taps, *, bias=None, residual=None,
            out=None, out2=None, split_col: int = 0, gn_sums=None, gn_groups: int = 0, rows_per_sample: int = 0,
            py: int = 0, px: int = 0, out_dtype=torch.bfloat16, ref: bool = False) -> torch.Tensor:
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
    fn = lib.vdn_tapgemm_ref if ref else lib.vdn_tapgemm
    check(fn(C.byref(d), ptr(x0), ptr(srcs[1]) if len(srcs) > 1 else None, ptr(wp), ptr(bias), ptr(residual),
             ptr(out), ptr(out2), ptr(gn_sums), stream_ptr()), "vdn_tapgemm")
    return out
