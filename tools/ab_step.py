"""A/B of the config_v2_2 training step (CUDA-graph replay, as bench.py times it) under sets of library switches:

  python tools/ab_step.py "" "VDN_PERSIST_NARROW_RES=1" "VDN_WG_CEIL=1,VDN_WG_COLSUM=1"

Every set builds its own TrainStep (dispatch decisions are baked into the captured graph)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200._lib import debug_switches, set_host_flag  # noqa: E402
from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion  # noqa: E402
from video_diffusion_nnx_b200.trainer import TrainStep  # noqa: E402
from video_diffusion_nnx_b200.unet3d import Unet3D  # noqa: E402

B = int(os.environ.get("AB_BATCH", "4"))


def measure(sw, n=30):
    with debug_switches(**sw):
        torch.manual_seed(0)
        net = Unet3D(dim=32, channels=1)
        gd = GaussianDiffusion(net, image_size=64, num_frames=10, channels=1, timesteps=1000, loss_type="l2")
        ts = TrainStep(gd, batch_size=B, use_graph=True, step_start_ema=0)
        ts.x.copy_(torch.rand(ts.x.shape))
        ts.t.copy_(torch.randint(0, 1000, (B,), dtype=torch.int32))
        ops.randn(ts.noise, 1, 0)
        for i in range(4):
            ts.step_device(i)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                ts.step_device(10 + i)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / n)
        loss = float(ts.loss.item()) if hasattr(ts, "loss") else float("nan")
    del ts, gd, net
    torch.cuda.empty_cache()
    return best, loss


sets = sys.argv[1:] or [""]
for spec in sets:
    sw, host = {}, {}
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        if k.startswith("HOST_"):  # Python-side engine switch (e.g. HOST_VDN_DEFER_JOINS=0)
            host[k[5:]] = v
        else:
            sw[k] = int(v)
    for k, v in host.items():
        set_host_flag(k, v)
    ms, loss = measure(sw)
    for k in host:
        set_host_flag(k, None)
    print(f"{spec or '(defaults)':60s} {ms:7.3f} ms/step  {B / ms * 1e3:7.1f} clips/s  loss {loss:.4f}", flush=True)
