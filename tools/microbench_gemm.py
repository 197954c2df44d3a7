"""CUDA-event timings of tapgemm / wgrad over the shapes of config_v2_2 (B=4): kernel microbench sweep."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402

dev = "cuda"
B, F = 4, 10


def bf(*s):
    return torch.randn(*s, device=dev).to(torch.bfloat16)


def timeit(fn, n=20):
    """GPU time per launch, measured on a CUDA graph of n launches (no host launch overhead)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3


def gemm_case(H, C, N, taps, n_src=1, gn=False, res=False):
    xs = [bf(B * F, H, H, C) for _ in range(n_src)]
    nt = len(taps)
    wp = bf(N, nt * n_src * C)
    out = torch.empty(B * F, H, H, N, dtype=torch.bfloat16, device=dev)
    bias = torch.zeros(N, device=dev)
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device=dev) if gn else None
    r = bf(B * F, H, H, N) if res else None
    us = timeit(lambda: ops.tapgemm(ops.VDN_TAP_UNIT, xs, wp, taps, bias=bias, out=out, gn_sums=sums, gn_groups=8,
                                    rows_per_sample=F * H * H, residual=r))
    M = B * F * H * H
    fl = 2.0 * M * N * nt * n_src * C
    by = (M * n_src * C + M * N) * 2
    print(f"tapgemm H={H:3d} C={C:4d}x{n_src} N={N:4d} taps={nt:2d} gn={int(gn)} res={int(res)}: {us:8.1f} us  "
          f"{fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s(alg)")


def wgrad_case(H, C, N, taps, n_src=1):
    xs = [bf(B * F, H, H, C) for _ in range(n_src)]
    g = bf(B * F, H, H, N)
    dw = torch.zeros(len(taps), n_src * C, N, device=dev)
    us = timeit(lambda: ops.wgrad(ops.VDN_TAP_UNIT, xs, g, dw, taps))
    M = B * F * H * H
    fl = 2.0 * M * N * len(taps) * n_src * C
    print(f"wgrad   H={H:3d} C={C:4d}x{n_src} N={N:4d} taps={len(taps):2d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")


T3, T1 = ops.TAPS_3x3, ops.TAPS_1x1
for args in [(64, 32, 32, T1), (64, 32, 32, T3), (64, 32, 32, T3, 1, True), (64, 32, 32, T3, 2, True), (64, 64, 32, T3),
             (64, 32, 768, T1), (64, 256, 32, T1, 1, False, True), (64, 768, 32, T1, 1, False, True),
             (32, 64, 64, T3, 1, True), (32, 64, 768, T1), (16, 128, 128, T3, 1, True), (16, 128, 768, T1),
             (8, 256, 256, T3, 1, True), (8, 256, 256, T3, 2, True), (8, 256, 768, T1), (8, 256, 256, T1)]:
    gemm_case(*args)
for args in [(64, 32, 32, T3), (64, 32, 768, T1), (64, 256, 32, T1), (32, 64, 64, T3), (16, 128, 128, T3),
             (8, 256, 256, T3), (8, 256, 256, T3, 2), (8, 256, 768, T1)]:
    wgrad_case(*args)
