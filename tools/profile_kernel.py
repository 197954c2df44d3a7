"""Runs one hot kernel in isolation for `ncu --set full` captures.
  python tools/profile_kernel.py conv_l0 | qkv_l0 | wgrad_l0 | conv_l3
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "conv_l0"
dev = "cuda"
B, F = 4, 10


def bf(*s):
    return torch.randn(*s, device=dev).to(torch.bfloat16)


if what in ("conv_l0", "conv_l3", "qkv_l0", "conv_v23x_l0", "conv_v23x_l1"):
    if what.startswith("conv_v23x"):  # config_v2_3 scale-up: 16 frames, 128x128, dim 128 (slab kernel; VDN_NO_SLABCONV=1: generic)
        F = 16
    H, C, N, taps = {"conv_l0": (64, 32, 32, ops.TAPS_3x3), "conv_l3": (8, 256, 256, ops.TAPS_3x3),
                     "qkv_l0": (64, 32, 768, ops.TAPS_1x1), "conv_v23x_l0": (128, 128, 128, ops.TAPS_3x3),
                     "conv_v23x_l1": (64, 256, 256, ops.TAPS_3x3)}[what]
    x = bf(B * F, H, H, C)
    w = torch.randn(len(taps), C, N, device=dev) * (len(taps) * C) ** -0.5
    wp = torch.empty(N, len(taps) * C, dtype=torch.bfloat16, device=dev)
    ops.pack_weight(w, wp, len(taps), C, N, 0)
    out = torch.empty(B * F, H, H, N, dtype=torch.bfloat16, device=dev)
    bias = torch.zeros(N, device=dev)
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device=dev)
    for _ in range(5):
        if what == "qkv_l0":
            ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, taps, bias=bias, out=out)
        else:
            ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, taps, bias=bias, out=out, gn_sums=sums, gn_groups=8,
                        rows_per_sample=F * H * H)
elif what == "wgrad_qkv_l0":  # weight + bias gradient of a q|k|v projection at the 64x64 level (the bench's `roofline`)
    x, g = bf(B * F, 64, 64, 32), bf(B * F, 64, 64, 768)
    dw, db = torch.zeros(1, 32, 768, device=dev), torch.zeros(768, device=dev)
    for _ in range(5):
        ops.wgrad(ops.VDN_TAP_UNIT, [x], g, dw, ops.TAPS_1x1, dbias=db)
elif what == "wgrad_l0":
    x, g = bf(B * F, 64, 64, 32), bf(B * F, 64, 64, 32)
    dw = torch.zeros(9, 32, 32, device=dev)
    for _ in range(5):
        ops.wgrad(ops.VDN_TAP_UNIT, [x], g, dw, ops.TAPS_3x3)
torch.cuda.synchronize()
print("done")
if what == "mha_fused_l0":
    C = 32
    x = bf(B, F, 64, 64, C)
    w = torch.randn(C, 768, device=dev) / C ** 0.5
    bias = torch.zeros(768, device=dev)
    w_hm = torch.empty(768, C, dtype=torch.bfloat16, device=dev)
    b_hm = torch.empty(768, device=dev)
    ops.qkv_headmajor_pack(w, bias, w_hm, b_hm, C)
    P = B * F * 64 * 64
    o = torch.empty(P, 256, dtype=torch.bfloat16, device=dev)
    for _ in range(4):
        ops.mha_temporal_fused_fwd(x, w_hm, b_hm, o, None, None, B, F, 64, 64, C)
    torch.cuda.synchronize()
    print("done")
