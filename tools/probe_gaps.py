"""Kernel-to-kernel gaps of back-to-back small-M tap-GEMM launches inside one CUDA graph: every CTA of 8 consecutive
launches stamps the global timer when its prologue is done, when its predecessor kernel is complete (pdl_wait returns)
and when it exits.

  python tools/probe_gaps.py [H C] > profiles/r2_probe_gaps.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200._lib import lib, debug_switches  # noqa: E402

dev = "cuda"
n_img, NL = 40, 8


def run(H, C, sw):
    w = torch.randn(9, C, C, device=dev) * (9 * C) ** -0.5
    wp = torch.empty(C, 9 * C, dtype=torch.bfloat16, device=dev)
    ops.pack_weight(w, wp, 9, C, C, 0)
    xs = [torch.randn(n_img, H, H, C, device=dev).to(torch.bfloat16) for _ in range(NL)]
    outs = [torch.empty(n_img, H, H, C, dtype=torch.bfloat16, device=dev) for _ in range(NL)]
    bias = torch.zeros(C, device=dev)
    sums = torch.zeros(ops.GN_REPLICAS, 4, 8, 2, device=dev)
    traces = torch.zeros(NL, 1024, dtype=torch.int64, device=dev)

    def launch(i, traced):
        lib.vdn_debug_tapgemm_trace(traces[i].data_ptr() if traced else None)
        ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i]], wp, ops.TAPS_3x3, bias=bias, out=outs[i], gn_sums=sums, gn_groups=8,
                    rows_per_sample=10 * H * H)
        lib.vdn_debug_tapgemm_trace(None)

    with debug_switches(**sw):
        for i in range(NL):
            launch(i, False)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for i in range(NL):
                    launch(i, True)
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
    t = traces[:, 256:].view(NL, 256, 3).cpu()
    n_cta = int((t[0, :, 2] != 0).sum())
    t0 = int(t[0, :n_cta, 0].min())
    print(f"conv {C}->{C} @{H}x{H} switches {sw}: {n_cta} CTAs traced per launch; ns relative to the first stamp")
    prev_end = None
    for i in range(NL):
        a = t[i, :n_cta] - t0
        pro, dep, ex = a[:, 0], a[:, 1], a[:, 2]
        line = (f"  launch {i}: prologue done {int(pro.min()):6d}..{int(pro.max()):6d} | predecessor complete "
                f"{int(dep.min()):6d}..{int(dep.max()):6d} | exit {int(ex.min()):6d}..{int(ex.max()):6d} "
                f"(CTA run {int((ex - dep).min()):5d}..{int((ex - dep).max()):5d}, median {int((ex - dep).median()):5d})")
        if prev_end is not None:
            line += f" | gap last exit -> first start {int(dep.min()) - prev_end:5d}"
        prev_end = int(ex.max())
        print(line)


shapes = [(8, 256), (16, 128)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[1]), int(sys.argv[2]))]
for (H, C) in shapes:
    run(H, C, {})
    run(H, C, {"VDN_NO_PDL": 1})
