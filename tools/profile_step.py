"""One eager (un-graphed) config_v2_2 training step between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [train|sample]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion  # noqa: E402
from video_diffusion_nnx_b200.trainer import TrainStep  # noqa: E402
from video_diffusion_nnx_b200.unet3d import Unet3D  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
net = Unet3D(dim=32, channels=1)
gd = GaussianDiffusion(net, image_size=64, num_frames=10, channels=1, timesteps=1000, loss_type="l2")
if mode == "train":
    ts = TrainStep(gd, batch_size=B, use_graph=False, step_start_ema=0)
    ts.x.copy_(torch.rand(ts.x.shape))
    ts.t.copy_(torch.randint(0, 1000, (B,), dtype=torch.int32))
    for i in range(2):
        ops.randn(ts.noise, 1, i)
        ts.step_device(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.randn(ts.noise, 1, 2)
    ts.step_device(2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", ts.loss.item())
else:
    eng = net.engine(B, 10, 64, 64, training=False)
    x = torch.randn(B, 1, 10, 64, 64, device="cuda")
    t = torch.full((B,), 500, dtype=torch.int32, device="cuda")
    for i in range(2):
        eng.forward(x, t)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.forward(x, t)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ok")
