"""Times the phases of the config_v2_2 training step separately (each replayed from its own CUDA graph):
forward + loss, backward, optimizer + repack. VDN_NO_OVERLAP=1 runs everything on one stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_nnx_b200 import ops  # noqa: E402
from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion  # noqa: E402
from video_diffusion_nnx_b200.trainer import TrainStep  # noqa: E402
from video_diffusion_nnx_b200.unet3d import Unet3D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = Unet3D(dim=32, channels=1)
gd = GaussianDiffusion(net, image_size=64, num_frames=10, channels=1, timesteps=1000, loss_type="l2")
ts = TrainStep(gd, batch_size=B, use_graph=False, step_start_ema=0)
ts.x.copy_(torch.rand(ts.x.shape))
ts.t.copy_(torch.randint(0, 1000, (B,), dtype=torch.int32))
for i in range(2):
    ops.randn(ts.noise, 1, i)
    ts.step_device(i)
torch.cuda.synchronize()


def graph_of(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def bwd():
    for fns, _ in ts._segments:
        for f in fns:
            f()


def timeit(g, n=20):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


gf, gb, go = graph_of(ts._fwd_loss), None, graph_of(ts._optimizer)
gf.replay()
gb = graph_of(bwd)


def both():
    ts._fwd_loss()
    bwd()
    ts._optimizer()


ga = graph_of(both)
print(f"overlap={'off' if os.environ.get('VDN_NO_OVERLAP') else 'on'} B={B}: forward+loss {timeit(gf):.3f} ms, backward {timeit(gb):.3f} ms, "
      f"optimizer+repack {timeit(go):.3f} ms, whole step {timeit(ga):.3f} ms")
