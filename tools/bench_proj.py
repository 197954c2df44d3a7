"""Times 1x1 projections (tap-GEMM with one tap) of the attention blocks at one level, CUDA-graph timed.
Run twice (VDN_NO_PERSIST=1 / unset) to compare the one-tile-per-CTA and the persistent kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import _lib, ops  # noqa: E402


def time_proj(n_img, H, W, C, N, res, nbuf=4, n=16):
    xs = [torch.randn(n_img, H, W, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    outs = [torch.empty(n_img, H, W, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    rs = [torch.randn(n_img, H, W, N, device="cuda").to(torch.bfloat16) for _ in range(nbuf)] if res else None
    w = torch.randn(1, C, N, device="cuda") * C ** -0.5
    wp = torch.empty(N, C, dtype=torch.bfloat16, device="cuda")
    ops.pack_weight(w, wp, 1, C, N, 0)

    def run(i):
        ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i % nbuf]], wp, ops.TAPS_1x1, out=outs[i % nbuf],
                    residual=rs[i % nbuf] if res else None)

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
    P = n_img * H * W
    gb = 2.0 * P * (C + N * (2 if res else 1)) / 1e9
    return us, 2.0 * P * C * N / (us * 1e-6) / 1e12, gb / (us * 1e-6) / 1e3


if __name__ == "__main__":
    tag = "generic" if os.environ.get("VDN_NO_PERSIST") else "persist"
    _lib.apply_env_switches()
    for name, n_img, HW, shapes in (("v2_3x L0", 64, 128, [(128, 768, 0), (256, 128, 1), (768, 128, 1), (128, 256, 0), (256, 128, 0)]),
                                    ("v2_3x L1", 64, 64, [(256, 768, 0), (256, 256, 1), (768, 256, 1)]),
                                    ("v2_2 L0", 40, 64, [(32, 768, 0), (256, 32, 1), (768, 32, 1), (32, 256, 0)]),
                                    ("v2_2 L1", 40, 32, [(64, 768, 0), (256, 64, 1), (768, 64, 1)])):
        for C, N, res in shapes:
            us, tf, tbs = time_proj(n_img, HW, HW, C, N, bool(res))
            print(f"{tag:8s} {name:9s} K={C:4d} N={N:4d} res={res}  {us:8.1f} us {tf:7.1f} TF/s {tbs:5.2f} TB/s", flush=True)
