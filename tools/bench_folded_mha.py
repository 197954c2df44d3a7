"""Folded temporal attention block (inference, C = 32): all-mma.sync kernel vs the tcgen05 version, 16 x 10 x 64 x 64."""
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


for (B, Fr, H, W) in ((16, 10, 64, 64), (4, 10, 64, 64), (16, 16, 64, 64)):
    x = torch.randn(B, Fr, H, W, 32, device=dev).to(torch.bfloat16)
    wqkv = torch.randn(32, 768, device=dev) / 32 ** 0.5
    bqkv = 0.1 * torch.randn(768, device=dev)
    wo = torch.randn(256, 32, device=dev) / 16.0
    bo = 0.1 * torch.randn(32, device=dev)
    fa, fm = (torch.empty(8, 32, 32, dtype=torch.bfloat16, device=dev) for _ in range(2))
    fu, fb = torch.empty(8, 32, device=dev), torch.empty(32, device=dev)
    ops.mha_fold_pack(wqkv, bqkv, wo, bo, fa, fu, fm, fb)
    outs = {}
    for name, sw in (("mma.sync", {"VDN_MHA_FOLDED_MMA": 1}), ("tcgen05", {})):
        out = torch.empty_like(x)
        with debug_switches(**sw):
            us = timeit(lambda: ops.mha_temporal_folded_fwd(x, fa, fu, fm, fb, out, B, Fr, H, W, 32))
        outs[name] = out.float()
        print(f"B={B} F={Fr} {H}x{W} {name:9s}: {us:7.1f} us", flush=True)
    d = (outs["tcgen05"] - outs["mma.sync"]).abs().max().item() / outs["mma.sync"].abs().max().item()
    print(f"   max |tcgen05 - mma.sync| / max |mma.sync| = {d:.2e}")

# phase timeline of CTA 0 of the tcgen05 kernel (clock64 deltas, cycles)
from video_diffusion_nnx_b200._lib import lib  # noqa: E402
trace = torch.zeros(1024, dtype=torch.int64, device=dev)
lib.vdn_debug_trace_buffer(trace.data_ptr())
ops.mha_temporal_folded_fwd(x, fa, fu, fm, fb, out, B, Fr, H, W, 32)
torch.cuda.synchronize()
lib.vdn_debug_trace_buffer(None)
tr = trace.view(128, 8)[:, :7].cpu()
rows = [r for r in tr.tolist() if r[0]]
names = ["wait x", "Y = X A", "y -> smem", "core", "O = Z M", "store"]
for i, r in enumerate(rows[:6] + rows[-2:]):
    print("tile", i, {n: r[k + 1] - r[k] for k, n in enumerate(names)}, "total", r[6] - r[0])
import statistics
print("median per phase over", len(rows), "tiles:", {n: int(statistics.median(r[k + 1] - r[k] for r in rows)) for k, n in enumerate(names)},
      "tile period", int(statistics.median(b[0] - a[0] for a, b in zip(rows, rows[1:]))))
