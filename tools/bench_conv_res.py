"""(1,3,3) convs of config_v2_2 with / without a residual operand (the dgrad of conv1 adds the LayerNorm-branch
gradient): CUDA-graph timed, rotating buffers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402

dev = "cuda"


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


for (H, C, nb) in ((64, 32, 12), (32, 64, 12), (16, 128, 24), (8, 256, 24)):
    n_img = 40
    w = torch.randn(9, C, C, device=dev) * (9 * C) ** -0.5
    wp = torch.empty(C, 9 * C, dtype=torch.bfloat16, device=dev)
    ops.pack_weight(w, wp, 9, C, C, 0)
    xs = [torch.randn(n_img, H, H, C, device=dev).to(torch.bfloat16) for _ in range(nb)]
    rs = [torch.randn(n_img, H, H, C, device=dev).to(torch.bfloat16) for _ in range(nb)]
    outs = [torch.empty(n_img, H, H, C, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
    for res in (False, True):
        fns = [(lambda i=i: ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i]], wp, ops.TAPS_3x3, out=outs[i],
                                        residual=rs[i] if res else None)) for i in range(nb)]
        print(f"conv {C}->{C} @{H}x{H} residual={int(res)}: {timeit(fns):6.2f} us", flush=True)
