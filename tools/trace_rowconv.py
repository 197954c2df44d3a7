"""Prints the producer / MMA / epilogue clock64 timeline of the row-ring conv kernel (first CTAs)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200._lib import lib  # noqa: E402

dev = "cuda"
n_img, H, W, c, N = 40, 64, 64, 32, 32
x = torch.randn(n_img, H, W, c, device=dev).to(torch.bfloat16)
w = torch.randn(9, c, N, device=dev) * (9 * c) ** -0.5
wp = torch.empty(N, 9 * c, dtype=torch.bfloat16, device=dev)
ops.pack_weight(w, wp, 9, c, N, 0)
out = torch.empty(n_img, H, W, N, dtype=torch.bfloat16, device=dev)
bias = torch.zeros(N, device=dev)
sums = torch.zeros(ops.GN_REPLICAS, 4, 8, 2, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
trace = torch.zeros(8 * 3 * 64 + 4 * 1024, dtype=torch.int64, device=dev)
for rep in range(3):
    flush.zero_()
    trace.zero_()
    lib.vdn_debug_rowconv_trace(C.c_void_p(trace.data_ptr()))
    ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, bias=bias, out=out, gn_sums=sums, gn_groups=8,
                rows_per_sample=10 * H * W)
    torch.cuda.synchronize()
    lib.vdn_debug_rowconv_trace(None)
tr = trace[:8 * 3 * 64].view(8, 3, 64).cpu()
cta = trace[8 * 3 * 64:].view(1024, 4).cpu()
cta = cta[cta[:, 0] != 0]
t0g = int(cta[:, 0].min())
import numpy as np
st = (cta[:, 0] - t0g).numpy(); en = (cta[:, 1] - t0g).numpy()
print("ctas", len(cta), "start ns: min/med/max", st.min(), np.median(st), st.max(), " end ns: min/med/max", en.min(), np.median(en), en.max(), " dur med/max", np.median(en - st), (en - st).max())
print("distinct SMs", len(set(cta[:, 2].tolist())))
for cta in (0, 1, 5):
    t0 = int(tr[cta, 0, 0])
    for role, name in enumerate(("producer", "mma", "epilogue")):
        vals = [int(v) - t0 for v in tr[cta, role] if int(v) != 0]
        print(f"cta {cta} {name:9s}", vals)
