"""Temporal attention forward of the training engines at C = 32 (writes q|k|v, o, lse): warp-MMA kernel vs the version
with the projection on tcgen05; CUDA-graph timed, q|k|v outputs rotating over sets larger than L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200._lib import debug_switches  # noqa: E402

dev, bf = "cuda", torch.bfloat16
B, Fr, S = 4, 10, 64
P = B * Fr * S * S


def gtime(run, n=12):
    for i in range(3):
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
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * n) * 1e3


x = torch.randn(B, Fr, S, S, 32, device=dev).to(bf)
w = torch.randn(32, 768, device=dev) / 32 ** 0.5
bias = 0.1 * torch.randn(768, device=dev)
w_hm = torch.empty(768, 32, dtype=bf, device=dev)
b_hm = torch.empty(768, device=dev)
ops.qkv_headmajor_pack(w, bias, w_hm, b_hm, 32)
nb = 3
qkvs = [torch.empty(P, 768, dtype=bf, device=dev) for _ in range(nb)]
os_ = [torch.empty(P, 256, dtype=bf, device=dev) for _ in range(nb)]
lse = torch.empty(P, 8, device=dev)
res = {}
for name, sw in (("mma.sync", {"VDN_MHA_TRAIN_MMA": 1}), ("tcgen05 projection", {})):
    with debug_switches(**sw):
        us = gtime(lambda i: ops.mha_temporal_tc_fwd(x, w_hm, b_hm, os_[i % nb], qkvs[i % nb], lse, B, Fr, S, S, 32))
    res[name] = (qkvs[0].float().clone(), os_[0].float().clone(), lse.clone())
    print(f"{name:20s}: {us:7.1f} us  ({(P * 2 * (32 + 768 + 256) + P * 32) / us / 1e6:.2f} TB/s of algorithmic traffic)", flush=True)
a, b = res["mma.sync"], res["tcgen05 projection"]
for nm, u, v in zip(("qkv", "o", "lse"), a, b):
    print(f"   {nm}: max |diff| / max |ref| = {(u - v).abs().max().item() / u.abs().max().item():.2e}")

# phase timeline of CTA 0 (clock64 deltas per tile: wait x | per pass: GEMM wait, convert + q|k|v store, core + o store)
from video_diffusion_nnx_b200._lib import lib  # noqa: E402
import statistics  # noqa: E402
trace = torch.zeros(1024, dtype=torch.int64, device=dev)
lib.vdn_debug_trace_buffer(trace.data_ptr())
ops.mha_temporal_tc_fwd(x, w_hm, b_hm, os_[0], qkvs[0], lse, B, Fr, S, S, 32)
torch.cuda.synchronize()
lib.vdn_debug_trace_buffer(None)
rows = [r for r in trace.view(64, 16)[:, :14].cpu().tolist() if r[0]]
d = [[r[k + 1] - r[k] for k in range(13)] for r in rows]
med = [int(statistics.median(c)) for c in zip(*d)]
print("median cycles over", len(rows), "tiles: wait x", med[0], "| passes (GEMM, convert, core):", [tuple(med[1 + 3 * p: 4 + 3 * p]) for p in range(4)],
      "| tile period", int(statistics.median(b[0] - a[0] for a, b in zip(rows, rows[1:]))))
