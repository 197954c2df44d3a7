"""Producer / MMA / epilogue clock64 timeline of CTA (0,0) of the generic tap-GEMM kernel on a small-M conv
(default: the 8x8 level of config_v2_2, 256 -> 256 channels, 36 K steps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200._lib import lib  # noqa: E402

dev = "cuda"
H = int(sys.argv[1]) if len(sys.argv) > 1 else 8
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n_img = 40
x = torch.randn(n_img, H, H, C, device=dev).to(torch.bfloat16)
w = torch.randn(9, C, C, device=dev) * (9 * C) ** -0.5
wp = torch.empty(C, 9 * C, dtype=torch.bfloat16, device=dev)
ops.pack_weight(w, wp, 9, C, C, 0)
out = torch.empty(n_img, H, H, C, dtype=torch.bfloat16, device=dev)
bias = torch.zeros(C, device=dev)
sums = torch.zeros(ops.GN_REPLICAS, 4, 8, 2, device=dev)
trace = torch.zeros(1024, dtype=torch.int64, device=dev)
for rep in range(3):
    trace.zero_()
    lib.vdn_debug_tapgemm_trace(trace.data_ptr())
    ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, bias=bias, out=out, gn_sums=sums, gn_groups=8, rows_per_sample=10 * H * H)
    torch.cuda.synchronize()
    lib.vdn_debug_tapgemm_trace(None)
tr = trace[:256].view(4, 64).cpu()
t0 = int(min(v for v in tr.flatten().tolist() if v))
p0 = [(i, int(v) - t0) for i, v in enumerate(tr[0].tolist()) if v]
p1 = [(i, int(v) - t0) for i, v in enumerate(tr[1].tolist()) if v]
mm = [int(v) - t0 for v in tr[2].tolist() if v]
print("producer0 (step, cycles after its empty-wait):", p0)
print("producer1:", p1)
print("mma (cycles after its full-wait, per step):", mm)
print("mma step deltas:", [b - a for a, b in zip(mm, mm[1:])])
ep = [int(v) - t0 for v in tr[3].tolist()[:7]]
print("epilogue: wait start, accumulator ready, epilogue done, CTA exit | column loop done, staged barrier passed, write-out done:", ep)
