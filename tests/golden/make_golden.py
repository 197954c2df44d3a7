"""Generates the committed golden vectors from the CPU oracle (oracle/). The reference itself cannot
be imported in this image (no jax/flax), so these pin the ORACLE against drift; the oracle in turn is
pinned to the reference's known answers by tests/test_oracle_known_answers.py.

Run from the repo root:  python tests/golden/make_golden.py
Weights are NOT stored (40 MB): they are regenerated from oracle.init_params(seed=3, perturb=0.05).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import diffusion_oracle as D  # noqa: E402
from oracle import unet3d_oracle as U  # noqa: E402


def case(name, B, Fr, HW, dim, T, loss_type):
    torch.set_num_threads(os.cpu_count())
    rng = np.random.default_rng(0)
    x = rng.random((B, 1, Fr, HW, HW), dtype=np.float32)                       # clips in [0,1)
    t = np.random.default_rng(1).integers(0, T, (B,)).astype(np.int32)
    noise = np.random.default_rng(2).standard_normal((B, 1, Fr, HW, HW)).astype(np.float32)
    p = U.init_params(dim, 1, seed=3, perturb=0.05)
    p64 = {k: v.double() for k, v in p.items()}
    out = {}
    for tag, pp, dt in (("f32", p, torch.float32), ("f64", p64, torch.float64)):
        gd = D.GaussianDiffusionOracle(lambda xx, tt: U.unet3d_forward(pp, xx, tt, dim), image_size=HW, num_frames=Fr,
                                       channels=1, timesteps=T, loss_type=loss_type, dtype=dt)
        xt = torch.from_numpy(x).to(dt)
        nz = torch.from_numpy(noise).to(dt)
        tt = torch.from_numpy(t)
        x_noisy = gd.q_sample(D.normalize_img(xt), tt, nz)
        eps = U.unet3d_forward(pp, x_noisy, tt, dim)
        loss = gd(xt, tt, nz)
        out[f"eps_{tag}"] = eps.numpy().astype(np.float32)
        out[f"loss_{tag}"] = np.array(loss.item(), np.float64)
        if tag == "f32":
            out["x_noisy"] = x_noisy.numpy()
    print(name, "loss f32", out["loss_f32"], "f64", out["loss_f64"],
          "eps f32-vs-f64 rel", np.abs(out["eps_f32"] - out["eps_f64"]).max() / np.abs(out["eps_f64"]).max())
    np.savez_compressed(os.path.join(os.path.dirname(__file__), name + ".npz"), x=x, t=t, noise=noise, **out)


if __name__ == "__main__":
    case("v1_0_b2_seed3", B=2, Fr=2, HW=64, dim=32, T=200, loss_type="l2")   # configs/config_v1_0.yaml shapes
