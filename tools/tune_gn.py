"""Times the GroupNorm+SiLU backward pair (and forward / tail / LayerNorm backward) at the four resolution levels of
config_v2_2 (batch 4) for the chunk-count switch VDN_GN_ITERS and the single-launch variant VDN_GN_FUSED, CUDA-graph
timed over buffers rotating beyond L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import _lib, ops  # noqa: E402


def graph_time_us(run, n=32, warm=4):
    for i in range(warm):
        run(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(n):
                run(i)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B, Fr = 4, 10
for lvl, (S, C) in enumerate([(64, 32), (32, 64), (16, 128), (8, 256)]):
    rows = Fr * S * S
    P = B * rows
    nb = max(3, int(300e6 // (P * C * 2 * 3)) + 1)
    nb = min(nb, 48)
    bf = torch.bfloat16
    xr = [torch.randn(B * Fr, S, S, C, device="cuda").to(bf) for _ in range(nb)]
    dy = [torch.randn(B * Fr, S, S, C, device="cuda").to(bf) for _ in range(nb)]
    dx = [torch.empty(B * Fr, S, S, C, dtype=bf, device="cuda") for _ in range(nb)]
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device="cuda")
    sums[0, :, :, 1] = rows * (C // 8)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    T = torch.zeros(B, C, 2, device="cuda")
    dg, db, dcb = (torch.zeros(C, device="cuda") for _ in range(3))

    def bwd(i):
        ops.gn_silu_bwd(dy[i % nb], xr[i % nb], sums, gamma, beta, None, T, dx[i % nb], dg, db, None, B, rows, C, dconv_bias=dcb)

    def fwd(i):
        ops.gn_silu_fwd(xr[i % nb], sums, gamma, beta, None, dx[i % nb], B, rows, C)

    def tail(i):
        ops.resblock_tail_fwd(xr[i % nb], sums, gamma, beta, dy[i % nb], gamma, beta, dx[i % nb], B, rows, C)

    def lnb(i):
        ops.ln_bwd(xr[i % nb], dy[i % nb], gamma, dx[i % nb], dg, db, P, C)

    row = [f"L{lvl} {S}x{S} C={C} ({P * C * 2 / 1e6:.1f} MB/tensor):"]
    row.append(f"fwd {graph_time_us(fwd):.1f}")
    row.append(f"tail {graph_time_us(tail):.1f}")
    row.append(f"ln_bwd {graph_time_us(lnb):.1f}")
    for it in (0, 1, 2, 3, 4, 6):
        if it:
            _lib.debug_set("VDN_GN_ITERS", it)
        row.append(f"bwd iters={it or 'auto'} {graph_time_us(bwd):.1f}")
        _lib.debug_clear("VDN_GN_ITERS")
    _lib.debug_set("VDN_GN_FUSED", 1)
    row.append(f"bwd fused {graph_time_us(bwd):.1f}")
    _lib.debug_clear("VDN_GN_FUSED")
    print(" | ".join(row), flush=True)
