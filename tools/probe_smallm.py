"""Small-M tap-GEMM probe: per-launch time (CUDA graph, rotating buffers) and the CTA (0,0) phase timeline of the
(1,3,3) convs of the 8x8 / 16x16 / 32x32 levels of config_v2_2 at several N-tile widths (VDN_BN).

  python tools/probe_smallm.py > profiles/r2_probe_smallm.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200._lib import lib, debug_switches  # noqa: E402

dev = "cuda"
n_img = 40


def timeit(fns, n=48):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(n):
                fns[i % len(fns)]()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (4 * n) * 1e3


def case(H, C, gn=True, nb=24, splitk=False):
    w = torch.randn(9, C, C, device=dev) * (9 * C) ** -0.5
    wps = []
    for _ in range(nb):
        wp = torch.empty(C, 9 * C, dtype=torch.bfloat16, device=dev)
        ops.pack_weight(w, wp, 9, C, C, 0)
        wps.append(wp)
    xs = [torch.randn(n_img, H, H, C, device=dev).to(torch.bfloat16) for _ in range(nb)]
    outs = [torch.empty(n_img, H, H, C, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
    bias = torch.zeros(C, device=dev)
    sums = torch.zeros(ops.GN_REPLICAS, 4, 8, 2, device=dev)
    kw = dict(gn_sums=sums, gn_groups=8, rows_per_sample=10 * H * H) if gn else {}
    if splitk:
        nbytes = ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, n_img, H, H, 1, C, ops.TAPS_3x3, C,
                                             gn_groups=8 if gn else 0, rows_per_sample=10 * H * H if gn else 0)
        if nbytes == 0:
            return None
        kw["workspace"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    fns = [(lambda i=i: ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i]], wps[i], ops.TAPS_3x3, bias=bias, out=outs[i], **kw))
           for i in range(nb)]
    return fns


def trace_once(fn):
    trace = torch.zeros(1024, dtype=torch.int64, device=dev)
    for _ in range(3):
        trace.zero_()
        lib.vdn_debug_tapgemm_trace(trace.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.vdn_debug_tapgemm_trace(None)
    tr = trace[:256].view(4, 64).cpu()
    vals = [v for v in tr.flatten().tolist() if v]
    if not vals:
        return "no trace (another kernel took the launch)"
    t0 = min(vals)
    mm = [int(v) - t0 for v in tr[2].tolist() if v]
    ep = [int(v) - t0 for v in tr[3].tolist()[:7]]
    return (f"mma first {mm[0]} last {mm[-1]} steps {len(mm)} ({(mm[-1] - mm[0]) / max(1, len(mm) - 1):.0f}/step) | acc ready {ep[1]} "
            f"cols done {ep[4]} staged {ep[5]} written {ep[6]} epi done {ep[2]} exit {ep[3]}")


for (H, C) in ((8, 256), (16, 128), (32, 64)):
    for gn in (True, False):
        M = n_img * H * H
        fl = 2.0 * M * 9 * C * C
        for sbn in (0, 128, 64):
            with debug_switches(VDN_SPLITK=1):
                fns = case(H, C, gn, splitk=True)
            if fns is None:
                break
            with debug_switches(VDN_SPLITK=1, **({"VDN_SPLITK_BN": sbn} if sbn else {})):
                us = timeit(fns)
                tr = trace_once(fns[0])
            print(f"conv {C}->{C} @{H}x{H} gn={int(gn)} split-K BN={sbn or 'auto':>4}: {us:6.2f} us {fl / us / 1e6:6.1f} TF/s | {tr} "
                  "(cols done = partial dumped, staged = cluster barrier passed, written = slice finalised)", flush=True)
        fns = case(H, C, gn)
        for bn in (0, 64, 128, 256):
            if bn > C:
                continue
            sw = {"VDN_BN": bn} if bn else {}
            if H == 32:
                sw["VDN_NO_ROWCONV"] = 1
            with debug_switches(**sw):
                us = timeit(fns)
                tr = trace_once(fns[0])
            M = n_img * H * H
            fl = 2.0 * M * 9 * C * C
            print(f"conv {C}->{C} @{H}x{H} gn={int(gn)} BN={bn or 'auto':>4}: {us:6.2f} us {fl / us / 1e6:6.1f} TF/s | {tr}", flush=True)
