"""The fp32-grade forward path (engine_f32.py / csrc/fp32_path.cu: fp32 activations, split-bf16 operands on the tcgen05
tap-GEMM) against the oracle at the tolerance BASELINE.json's north_star states for fp32: loss and predicted noise
within 1e-3 relative. The reference computes in float32 throughout (modules.py has no dtype=), so the gate is against
the float64 oracle (the float32 oracle itself sits ~1e-6 from it)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("dim,B,Fr,S,T", [(32, 2, 2, 64, 200), (32, 1, 10, 64, 1000), (64, 1, 4, 32, 1000)])
def test_fp32_path_loss_and_eps_within_1e_3(dim, B, Fr, S, T):
    from oracle import diffusion_oracle as D
    from oracle import unet3d_oracle as U
    from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion
    from video_diffusion_nnx_b200.unet3d import Unet3D

    p32 = U.init_params(dim, 1, seed=3, perturb=0.05)
    p64 = {k: v.double() for k, v in p32.items()}
    rng = np.random.default_rng(17)
    x = torch.from_numpy(rng.random((B, 1, Fr, S, S), dtype=np.float32))
    t = torch.from_numpy(rng.integers(0, T, (B,)).astype(np.int32))
    noise = torch.from_numpy(rng.standard_normal((B, 1, Fr, S, S)).astype(np.float32))
    cap = {}

    def fwd(xx, tt):
        cap["eps"] = U.unet3d_forward(p64, xx, tt, dim)
        return cap["eps"]

    gdo = D.GaussianDiffusionOracle(fwd, image_size=S, num_frames=Fr, channels=1, timesteps=T, loss_type="l2", dtype=torch.float64)
    loss_ref = float(gdo(x.double(), t, noise.double()).item())
    eps_ref = cap["eps"]

    net = Unet3D(dim=dim, channels=1, precision="fp32")
    net.load_state_dict({k: v.numpy() for k, v in p32.items()})
    gd = GaussianDiffusion(net, image_size=S, num_frames=Fr, channels=1, timesteps=T, loss_type="l2")
    xn = gd.q_sample(x.cuda() * 2 - 1, t.cuda(), noise=noise.cuda())
    eps = net(xn, t.cuda())
    loss = float(gd.p_losses(x.cuda() * 2 - 1, t.cuda(), noise=noise.cuda()).item())
    e_eps, e_loss = _rel_l2(eps, eps_ref), abs(loss - loss_ref) / loss_ref
    # the bf16 throughput path on the same inputs, for the record
    net.precision = "bf16"
    e_bf16 = _rel_l2(net(xn, t.cuda()), eps_ref)
    print(f"dim {dim} B{B} F{Fr} {S}x{S}: fp32-grade eps rel-L2 {e_eps:.2e}, loss rel {e_loss:.2e} (bf16 path eps {e_bf16:.2e})")
    assert eps.dtype == torch.float32 and eps.shape == (B, Fr, S, S, 1)
    assert e_eps < TOL and e_loss < TOL
    assert e_bf16 > 3 * e_eps  # the switch really selects a different arithmetic


def test_fp32_path_tracks_weight_updates():
    """The split operands are packed per store version: a state upload is picked up by the next forward."""
    from oracle import unet3d_oracle as U
    from video_diffusion_nnx_b200.unet3d import Unet3D

    net = Unet3D(dim=32, channels=1, precision="fp32")
    x = torch.randn(1, 1, 2, 64, 64, device="cuda")
    t = torch.tensor([5], dtype=torch.int32, device="cuda")
    a = net(x, t).clone()
    p = U.init_params(32, 1, seed=9, perturb=0.05)
    net.load_state_dict({k: v.numpy() for k, v in p.items()})
    b = net(x, t)
    ref = U.unet3d_forward({k: v.double() for k, v in p.items()}, x.cpu().double(), t.cpu(), 32)
    assert _rel_l2(a, b) > 0.1 and _rel_l2(b, ref) < TOL
