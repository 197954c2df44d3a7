"""Fused SpatialLinearAttention forward (inference, C = 32): apply pass on mma.sync vs tcgen05; 160 frames of 64 x 64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200._lib import debug_switches  # noqa: E402

dev = "cuda"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for n_img, H in ((160, 64), (40, 64)):
    N = H * H
    x = torch.randn(n_img, H, H, 32, device=dev).to(torch.bfloat16)
    wq = (torch.randn(32, 768, device=dev) * 32 ** -0.5).to(torch.bfloat16)
    wo = (torch.randn(256, 32, device=dev) / 16).to(torch.bfloat16)
    w_qkv = torch.empty(768, 32, dtype=torch.bfloat16, device=dev)
    w_out = torch.empty(32, 256, dtype=torch.bfloat16, device=dev)
    ops.pack_weight(wq.float().reshape(1, 32, 768).contiguous(), w_qkv, 1, 32, 768, 0)
    ops.pack_weight(wo.float().reshape(1, 256, 32).contiguous(), w_out, 1, 256, 32, 0)
    ctx = torch.empty(n_img, 8, 32, 32, device=dev)
    kstat = torch.empty(n_img, 8, 2, 32, device=dev)
    ws = torch.empty(ops.sla_workspace_floats(n_img, N), device=dev)
    outs = {}
    for name, sw in (("mma.sync", {"VDN_SLA_APPLY_MMA": 1}), ("tcgen05", {})):
        out = torch.empty_like(x)
        with debug_switches(**sw):
            us = timeit(lambda: ops.sla_fused_fwd(x, w_qkv, w_out, out, ctx, kstat, ws, n_img, N, 32))
        outs[name] = out.float()
        print(f"n_img={n_img} {H}x{H} whole fused forward (ctx pass + merge + apply), apply on {name:9s}: {us:7.1f} us", flush=True)
    d = (outs["tcgen05"] - outs["mma.sync"]).abs().max().item() / outs["mma.sync"].abs().max().item()
    print(f"   max |tcgen05 - mma.sync| / max |mma.sync| = {d:.2e}")
