"""Times the (1,3,3) conv shapes of the 64x64 / 32x32 levels (config_v2_2, B=4) through vdn_tapgemm with
the row-ring kernel's tuning knobs (VDN_RC_S stages, VDN_RC_CPS CTAs per SM), CUDA events, rotating over
buffers larger than L2.   python tools/bench_rowconv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import _lib, ops  # noqa: E402

dev = "cuda"


def bench(n_img, H, W, n_src, N, split, nbuf=12, iters=48):
    c = 32
    xs = [[torch.randn(n_img, H, W, c, device=dev).to(torch.bfloat16) for _ in range(n_src)] for _ in range(nbuf)]
    outs = [torch.empty(n_img, H, W, 32 if split else N, dtype=torch.bfloat16, device=dev) for _ in range(nbuf)]
    outs2 = [torch.empty(n_img, H, W, 32, dtype=torch.bfloat16, device=dev) for _ in range(nbuf)] if split else None
    w = torch.randn(9, n_src * c, N, device=dev) * (9 * n_src * c) ** -0.5
    wp = torch.empty(N, 9 * n_src * c, dtype=torch.bfloat16, device=dev)
    ops.pack_weight(w, wp, 9, n_src * c, N, 0)
    bias = torch.zeros(N, device=dev)
    sums = torch.zeros(ops.GN_REPLICAS, 4, 8, 2, device=dev)

    def run(i):
        if split:
            ops.tapgemm(ops.VDN_TAP_UNIT, xs[i % nbuf], wp, ops.TAPS_3x3, out=outs[i % nbuf], out2=outs2[i % nbuf],
                        split_col=32)
        else:
            ops.tapgemm(ops.VDN_TAP_UNIT, xs[i % nbuf], wp, ops.TAPS_3x3, bias=bias, out=outs[i % nbuf], gn_sums=sums,
                        gn_groups=8, rows_per_sample=n_img // 4 * H * W)

    for i in range(6):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    M = n_img * H * W
    byts = 2.0 * M * (n_src * c + N) + 2 * 9 * n_src * c * N
    return us, byts / us / 1e3


shapes = [("L0 32->32", 40, 64, 64, 1, 32, False), ("L0 64->32", 40, 64, 64, 2, 32, False),
          ("L0 dgrad 32->64 split", 40, 64, 64, 1, 64, True), ("L1 32->64", 40, 32, 32, 1, 64, False)]
for name, *shp in shapes:
    row = []
    for cfg in [("off", None, None), ("def", None, None), ("S2c2", 2, 2), ("S3c2", 3, 2), ("S3c1", 3, 1), ("S4c1", 4, 1),
                ("S2c3", 2, 3)]:
        tag, S, cps = cfg
        for k in ("VDN_RC_S", "VDN_RC_CPS", "VDN_BN"):
            os.environ.pop(k, None)
        if tag == "off":
            os.environ["VDN_RC_S"] = "2"  # placeholder; generic kernel is selected below by VDN_NO_ROWCONV at import
            continue
        if S:
            os.environ["VDN_RC_S"], os.environ["VDN_RC_CPS"] = str(S), str(cps)
        _lib.apply_env_switches()
        try:
            us, gbs = bench(*shp)
            row.append(f"{tag}: {us:6.1f} us {gbs:6.0f} GB/s")
        except Exception as e:  # smem overflow for this combination
            row.append(f"{tag}: n/a")
    print(f"{name:24s} " + " | ".join(row), flush=True)
