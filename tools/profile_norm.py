"""Runs the GroupNorm / LayerNorm elementwise kernels at one level's shape (config_v2_2, B=4), for ncu.
  python tools/profile_norm.py [level]   (level 0 = 64x64 C=32 ... 3 = 8x8 C=256)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402

lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, Fr = 4, 10
H = 64 >> lvl
C = 32 << lvl
rows = Fr * H * H
dev = "cuda"


def bf(*s):
    return torch.randn(*s, device=dev).to(torch.bfloat16)


x, dy, s = bf(B, rows, C), bf(B, rows, C), bf(B, rows, C)
out, dx, ds = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device=dev)
xf = x.float().view(B, rows, 8, C // 8)
sums[0, :, :, 0] = xf.sum(dim=(1, 3))
sums[0, :, :, 1] = (xf * xf).sum(dim=(1, 3))
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
ss = torch.randn(B, 2 * C, device=dev) * 0.1
T = torch.zeros(B, C, 2, device=dev)
dg, db, dss, dcb = (torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(B, 2 * C, device=dev),
                    torch.zeros(C, device=dev))
for _ in range(3):
    ops.gn_silu_fwd(x, sums, gamma, beta, ss, out, B, rows, C)
    ops.resblock_tail_fwd(x, sums, gamma, beta, s, gamma, beta, out, B, rows, C)
    T.zero_()
    ops.gn_silu_bwd(dy, x, sums, gamma, beta, ss, T, dx, dg, db, dss, B, rows, C, dconv_bias=dcb)
    ops.ln_bwd(s, dy, gamma, ds, dg, db, B * rows, C)
    ops.colsum(dy, dcb, B * rows, C)
torch.cuda.synchronize()
print("done")
