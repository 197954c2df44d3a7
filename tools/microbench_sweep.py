"""Kernel microbench sweep (BASELINE.json configs[4]): (1,3,3) conv, SpatialLinearAttention block and temporal
attention across batch / frames / resolution, plus the GroupNorm kernels, each timed as a CUDA graph of repeated
launches (buffers rotate so that the working set exceeds L2 where it can) and reported against the measured peaks
(MEASURED_PEAKS.json): tensor TFLOP/s for the contractions, HBM GB/s for everything (algorithmic bytes).

  python tools/microbench_sweep.py > profiles/r1_microbench_sweep.txt
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_diffusion_nnx_b200 import ops  # noqa: E402

dev = "cuda"
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TF = pk["hbm_gbs"], pk["bf16_tflops"]


def bf(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(torch.bfloat16)


def timeit(fns, n=24):
    """us per launch of the rotating list of closures `fns`, from a CUDA graph of n launches."""
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
    for _ in range(2):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * n) * 1e3


def row(name, us, flops, byts):
    tf, gbs = flops / us / 1e6, byts / us / 1e3
    print(f"{name:58s} {us:8.1f} us {tf:8.1f} TF/s ({100 * tf / TF:5.1f}%) {gbs:8.0f} GB/s ({100 * gbs / HBM:5.1f}%)", flush=True)


def conv_case(B, Fr, H, C, N, n_src=1):
    n_img, nb = B * Fr, 6
    w = torch.randn(9, n_src * C, N, device=dev) * (9 * n_src * C) ** -0.5
    wp = torch.empty(N, 9 * n_src * C, dtype=torch.bfloat16, device=dev)
    ops.pack_weight(w, wp, 9, n_src * C, N, 0)
    bias = torch.zeros(N, device=dev)
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device=dev)
    xs = [[bf(n_img, H, H, C) for _ in range(n_src)] for _ in range(nb)]
    outs = [torch.empty(n_img, H, H, N, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
    fns = [(lambda i=i: ops.tapgemm(ops.VDN_TAP_UNIT, xs[i], wp, ops.TAPS_3x3, bias=bias, out=outs[i], gn_sums=sums,
                                    gn_groups=8, rows_per_sample=Fr * H * H)) for i in range(nb)]
    us = timeit(fns)
    P = n_img * H * H
    row(f"conv(1,3,3) {n_src * C:3d}->{N:3d} @{H}x{H} B={B} F={Fr}", us, 2.0 * P * 9 * n_src * C * N,
        2.0 * P * (n_src * C + N) + 2 * 9 * n_src * C * N)


def sla_case(B, Fr, H, C):
    n_img, N, nb = B * Fr, H * H, 4
    P = n_img * N
    wq = torch.randn(1, C, 768, device=dev) * C ** -0.5
    wo = torch.randn(1, 256, C, device=dev) / 16
    w_qkv = torch.empty(768, C, dtype=torch.bfloat16, device=dev)
    w_out = torch.empty(C, 256, dtype=torch.bfloat16, device=dev)
    ops.pack_weight(wq, w_qkv, 1, C, 768, 0)
    ops.pack_weight(wo, w_out, 1, 256, C, 0)
    xs = [bf(n_img, H, H, C) for _ in range(nb)]
    out = torch.empty(n_img, H, H, C, dtype=torch.bfloat16, device=dev)
    ctx = torch.empty(n_img, 8, 32, 32, device=dev)
    kstat = torch.empty(n_img, 8, 2, 32, device=dev)
    ws = torch.empty(ops.sla_workspace_floats(n_img, N), device=dev)
    flops = 2.0 * P * C * 768 + 4.0 * P * 256 * 32 + 2.0 * P * 256 * C
    if C == 32:
        fns = [(lambda i=i: ops.sla_fused_fwd(xs[i], w_qkv, w_out, out, ctx, kstat, ws, n_img, N, C)) for i in range(nb)]
        row(f"SLA block fwd fused      C={C:3d} @{H}x{H} B={B} F={Fr}", timeit(fns), flops, 2.0 * P * C * 3)
    qkv = torch.empty(P, 768, dtype=torch.bfloat16, device=dev)
    tok = torch.empty(P, 256, dtype=torch.bfloat16, device=dev)

    def unfused(i):
        ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i]], w_qkv, ops.TAPS_1x1, out=qkv.view(n_img, H, H, 768))
        ops.sla_core_fwd(qkv, tok, ctx, kstat, ws, n_img, N)
        ops.tapgemm(ops.VDN_TAP_UNIT, [tok.view(n_img, H, H, 256)], w_out, ops.TAPS_1x1, residual=xs[i], out=out)

    fns = [(lambda i=i: unfused(i)) for i in range(nb)]
    row(f"SLA block fwd unfused    C={C:3d} @{H}x{H} B={B} F={Fr}", timeit(fns, n=12), flops, 2.0 * P * C * 3)
    dtok, dqkv, dctx = bf(P, 256), torch.empty(P, 768, dtype=torch.bfloat16, device=dev), torch.empty(n_img, 8, 32, 32, device=dev)
    ops.sla_core_fwd(qkv, tok, ctx, kstat, ws, n_img, N)
    fns = [lambda: ops.sla_core_bwd(qkv, dtok, ctx, kstat, dctx, dqkv, n_img, N)]
    row(f"SLA core bwd (dctx + tokens)      @{H}x{H} B={B} F={Fr}", timeit(fns, n=12), 12.0 * P * 256 * 32 / 2 * 2,
        2.0 * P * (768 * 2 + 256 + 768))


def mha_case(B, Fr, H, C):
    P, nb = B * Fr * H * H, 4
    w = torch.randn(C, 768, device=dev) / C ** 0.5
    bias = 0.1 * torch.randn(768, device=dev)
    w_hm = torch.empty(768, C, dtype=torch.bfloat16, device=dev)
    b_hm = torch.empty(768, device=dev)
    ops.qkv_headmajor_pack(w, bias, w_hm, b_hm, C)
    xs = [bf(B, Fr, H, H, C) for _ in range(nb)]
    o = torch.empty(P, 256, dtype=torch.bfloat16, device=dev)
    qkv = torch.empty(P, 768, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(P, 8, device=dev)
    flops = 2.0 * P * C * 768 + 4.0 * P * Fr * 256
    if ops.mha_tc_supported(Fr, C):
        fns = [(lambda i=i: ops.mha_temporal_tc_fwd(xs[i], w_hm, b_hm, o, qkv, lse, B, Fr, H, H, C)) for i in range(nb)]
        row(f"temporal MHA fwd (proj+core, train) C={C:3d} @{H}x{H} B={B} F={Fr}", timeit(fns, n=12), flops,
            2.0 * P * (C + 768 + 256))
        fns = [(lambda i=i: ops.mha_temporal_tc_fwd(xs[i], w_hm, b_hm, o, None, None, B, Fr, H, H, C)) for i in range(nb)]
        row(f"temporal MHA fwd (proj+core, infer) C={C:3d} @{H}x{H} B={B} F={Fr}", timeit(fns, n=12), flops, 2.0 * P * (C + 256))
    if C == 32:
        wo, bo = torch.randn(256, C, device=dev) / 16, torch.zeros(C, device=dev)
        fa, fm = (torch.empty(8, 32, 32, dtype=torch.bfloat16, device=dev) for _ in range(2))
        fu, fb = torch.empty(8, 32, device=dev), torch.empty(32, device=dev)
        ops.mha_fold_pack(w, bias, wo, bo, fa, fu, fm, fb)
        out = torch.empty(B, Fr, H, H, C, dtype=torch.bfloat16, device=dev)
        fns = [(lambda i=i: ops.mha_temporal_folded_fwd(xs[i], fa, fu, fm, fb, out, B, Fr, H, H, C)) for i in range(nb)]
        row(f"temporal MHA BLOCK folded (infer)   C={C:3d} @{H}x{H} B={B} F={Fr}", timeit(fns, n=12),
            flops + 2.0 * P * 256 * C, 2.0 * P * C * 2)
    do, dq = bf(P, 256), torch.empty(P, 768, dtype=torch.bfloat16, device=dev)
    fns = [lambda: ops.mha_temporal_tc_bwd(qkv, do, lse, dq, B, Fr, H, H)]
    row(f"temporal MHA core bwd                       @{H}x{H} B={B} F={Fr}", timeit(fns, n=12), 10.0 * P * Fr * 256,
        2.0 * P * (768 + 256 + 768))


def norm_case(B, Fr, H, C):
    rows = Fr * H * H
    x, dy, s = bf(B, rows, C), bf(B, rows, C), bf(B, rows, C)
    out, dx, ds = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device=dev)
    xf = x.float().view(B, rows, 8, C // 8)
    sums[0, :, :, 0], sums[0, :, :, 1] = xf.sum(dim=(1, 3)), (xf * xf).sum(dim=(1, 3))
    gam, bet, ss = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.randn(B, 2 * C, device=dev) * 0.1
    T = torch.zeros(B, C, 2, device=dev)
    dg, db, dss, dcb = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(B, 2 * C, device=dev), torch.zeros(C, device=dev)
    E = 2.0 * B * rows * C
    tag = f"C={C:3d} @{H}x{H} B={B} F={Fr}"
    row(f"gn_silu_fwd            {tag}", timeit([lambda: ops.gn_silu_fwd(x, sums, gam, bet, ss, out, B, rows, C)]), 0, 2 * E)
    row(f"resblock_tail_fwd      {tag}", timeit([lambda: ops.resblock_tail_fwd(x, sums, gam, bet, s, gam, bet, out, B, rows, C)]), 0, 3 * E)
    row(f"gn_silu_bwd (2 passes) {tag}", timeit([lambda: ops.gn_silu_bwd(dy, x, sums, gam, bet, ss, T, dx, dg, db, dss, B, rows, C, dconv_bias=dcb)]), 0, 5 * E)
    row(f"ln_bwd                 {tag}", timeit([lambda: ops.ln_bwd(s, dy, gam, ds, dg, db, B * rows, C)]), 0, 3 * E)


print(f"# peaks: HBM {HBM} GB/s, bf16 {TF} TFLOP/s (MEASURED_PEAKS.json); CUDA-graph timing, us per launch")
print("# --- (1,3,3) convolutions (fwd + bias + GroupNorm partial sums) ---")
for B_, Fr_, H_, C_, N_, ns_ in [(4, 10, 64, 32, 32, 1), (16, 10, 64, 32, 32, 1), (4, 10, 64, 32, 32, 2), (4, 10, 32, 64, 64, 1),
                                 (4, 10, 16, 128, 128, 1), (4, 10, 8, 256, 256, 1), (4, 16, 32, 128, 128, 1), (2, 16, 128, 32, 32, 1),
                                 # the four levels of the v2_3x workload (slab kernel at widths >= 32) and its concat conv
                                 (4, 16, 128, 128, 128, 1), (4, 16, 64, 256, 256, 1), (4, 16, 32, 512, 512, 1),
                                 (4, 16, 16, 1024, 1024, 1), (4, 16, 128, 128, 128, 2)]:
    conv_case(B_, Fr_, H_, C_, N_, ns_)
print("# --- SpatialLinearAttention ---")
for B_, Fr_, H_, C_ in [(4, 10, 64, 32), (16, 10, 64, 32), (4, 10, 32, 64), (4, 10, 16, 128), (4, 16, 32, 128)]:
    sla_case(B_, Fr_, H_, C_)
print("# --- temporal attention ---")
for B_, Fr_, H_, C_ in [(4, 10, 64, 32), (16, 10, 64, 32), (4, 2, 64, 32), (4, 16, 64, 32), (4, 10, 32, 64), (4, 16, 32, 128)]:
    mha_case(B_, Fr_, H_, C_)
print("# --- GroupNorm / LayerNorm elementwise kernels (algorithmic bytes / time) ---")
for B_, Fr_, H_, C_ in [(4, 10, 64, 32), (16, 10, 64, 32), (4, 10, 32, 64), (4, 10, 8, 256)]:
    norm_case(B_, Fr_, H_, C_)
