"""Times the slab (1,3,3) conv (csrc/conv3x3_slab.cu) on one layer shape under the kernel's experiment switches
(VDN_SLAB_DBG: 1 no epilogue, 2 no MMAs, 4 no slab loads, 8 no weight loads; VDN_SLAB_S stages), CUDA-graph timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import _lib, ops  # noqa: E402


def time_conv(n_img, H, W, C, N, nbuf=8, n=32, gn=False):
    xs = [torch.randn(n_img, H, W, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    outs = [torch.empty(n_img, H, W, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    w = torch.randn(9, C, N, device="cuda") * (9 * C) ** -0.5
    wp = torch.empty(N, 9 * C, dtype=torch.bfloat16, device="cuda")
    ops.pack_weight(w, wp, 9, C, N, 0)
    bias = torch.zeros(N, device="cuda")
    sums = torch.zeros(ops.GN_REPLICAS, 4, 8, 2, device="cuda")

    def run(i):
        if gn:
            ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i % nbuf]], wp, ops.TAPS_3x3, bias=bias, out=outs[i % nbuf], gn_sums=sums,
                        gn_groups=8, rows_per_sample=n_img // 4 * H * W)
        else:
            ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i % nbuf]], wp, ops.TAPS_3x3, bias=bias, out=outs[i % nbuf])

    for i in range(4):
        run(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(n):
                run(i)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    tf = 2.0 * n_img * H * W * C * 9 * N / (us * 1e-6) / 1e12
    return us, tf


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gn":
        for gn, dbg in ((False, "0"), (True, "0"), (True, "16"), (True, "32"), (True, "48")):
            os.environ["VDN_SLAB_DBG"] = dbg
            _lib.apply_env_switches()
            us, tf = time_conv(64, 128, 128, 128, 128, gn=gn)
            print(f"gn={gn} dbg={dbg} (16: no atomics, 32: no accumulation): {us:8.1f} us {tf:7.1f} TF/s", flush=True)
        sys.exit(0)
    shapes = [(64, 128, 128, 128, 128), (64, 64, 64, 256, 256), (64, 32, 32, 512, 512), (64, 16, 16, 1024, 1024)]
    if len(sys.argv) > 1:
        shapes = shapes[: int(sys.argv[1])]
    for shp in shapes:
        for env in ({"VDN_SLAB_MIN_ITEMS": "1000000000"}, {}, {"VDN_SLAB_DBG": "1"}, {"VDN_SLAB_DBG": "2"},
                    {"VDN_SLAB_DBG": "3"}, {"VDN_SLAB_BK": "16"}, {"VDN_SLAB_BK": "16", "VDN_SLAB_DBG": "1"},
                    {"VDN_SLAB_BK": "16", "VDN_SLAB_DBG": "2"}, {"VDN_SLAB_BK": "16", "VDN_SLAB_DBG": "3"},
                    {"VDN_SLAB_BK": "16", "VDN_SLAB_S": "3"}):
            for k in ("VDN_SLAB_BK", "VDN_SLAB_DBG", "VDN_SLAB_S", "VDN_SLAB_MIN_ITEMS", "VDN_SLAB_GRID"):
                os.environ.pop(k, None)
            os.environ.update(env)
            _lib.apply_env_switches()
            us, tf = time_conv(*shp)
            print(f"{shp} {str(env):40s} {us:8.1f} us {tf:7.1f} TF/s", flush=True)
