"""Generates tests/golden/ref_code_*.npz|json by EXECUTING THE REFERENCE'S OWN PYTHON (/root/reference/modules.py,
unet3d.py, gaussian_diffusion.py, utils.py - imported unmodified, read-only) over oracle/refshim, the numpy
restatement of the jax / flax.nnx calls they make (oracle/refshim/README.md says what that pins and what it does not).
Runs only where /root/reference exists (the build container); the fixtures it writes travel with the repo and are what
tests/test_oracle_vs_reference_code.py holds the oracle to.

    python tests/golden/make_ref_golden.py
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(1, REF)
sys.path.insert(2, ROOT)
warnings.filterwarnings("ignore", category=RuntimeWarning)  # log(0) in the bucket function's masked branch

import torch  # noqa: E402
from flax import nnx  # noqa: E402  (refshim)
import jax  # noqa: E402  (refshim)

import gaussian_diffusion as ref_gd  # noqa: E402  (the reference)
import modules as ref_modules  # noqa: E402
import unet3d as ref_unet  # noqa: E402
import utils as ref_utils  # noqa: E402

from oracle import unet3d_oracle as U  # noqa: E402

F64 = np.float64


def load_params(module, params, prefix=""):
    """Overwrite every Variable of a reference module with the array of the same nnx path."""
    st = nnx.state_paths(module)
    for path, var in st.items():
        key = prefix + path
        if key in params:
            a = np.asarray(params[key], F64)
            assert a.shape == var.shape, (key, a.shape, var.shape)
            var.value = a.copy()
    return st


def main():
    rng = np.random.default_rng(2024)
    out = {}

    # ---- 1. the state tree of the reference's Unet3D: names and shapes straight from its constructor ----
    dim, ch = 32, 1
    net = ref_unet.Unet3D(dim=dim, rngs=nnx.Rngs(0), channels=ch)
    st = nnx.state_paths(net)
    tree = {k: list(v.shape) for k, v in st.items()}
    gd = ref_gd.GaussianDiffusion(net, image_size=32, num_frames=3, channels=ch, timesteps=200, loss_type="l2")
    gd_tree = {k: list(v.shape) for k, v in nnx.state_paths(gd).items()}
    json.dump({"unet3d_dim32_ch1": tree, "gaussian_diffusion_T200": gd_tree,
               "n_params_unet": int(sum(int(np.prod(s)) for s in tree.values()))},
              open(os.path.join(HERE, "ref_code_state_tree.json"), "w"), indent=0, sort_keys=True)

    # ---- 2. Unet3D forward + diffusion methods on the oracle's seeded parameters (float64) ----
    p = {k: v.numpy().astype(F64) for k, v in U.init_params(dim, ch, seed=3, perturb=0.05, dtype=torch.float64).items()}
    load_params(net, p)
    B, Fr, S, T = 2, 3, 32, 200
    x = rng.random((B, ch, Fr, S, S))
    t = np.array([57, 3], dtype=np.int32)
    noise = rng.standard_normal((B, ch, Fr, S, S))
    out.update(unet_x=x, unet_t=t, unet_noise=noise)
    out["unet_eps"] = net(x * 2 - 1, t)
    out["unet_eps_fwcs"] = net.forward_with_cond_scale(x * 2 - 1, t, cond_scale=2.0)  # has_cond False: one forward
    key = jax.random.PRNGKey(5)
    out["q_sample"] = gd.q_sample(x * 2 - 1, t, key, noise=noise)
    out["p_losses_l2"] = np.asarray(gd.p_losses(x * 2 - 1, t, key, noise=noise))
    gd.loss_type = "l1"
    out["p_losses_l1"] = np.asarray(gd.p_losses(x * 2 - 1, t, key, noise=noise))
    gd.loss_type = "l2"
    x0 = gd.predict_start_from_noise(noise, t, x)
    out["predict_start"] = x0
    pm, pv, plv = gd.q_posterior(x0, noise, t)
    out.update(q_posterior_mean=pm, q_posterior_var=np.asarray(pv), q_posterior_logvar=np.asarray(plv))
    qm, qv, qlv = gd.q_mean_variance(x, t)
    out.update(q_mean=qm, q_var=np.asarray(qv), q_logvar=np.asarray(qlv))
    for name, tt in (("hi", np.array([199, 120], np.int32)), ("zero", np.array([0, 0], np.int32))):
        k = jax.random.PRNGKey(11)
        out[f"p_sample_z_{name}"] = jax.random.normal(k, shape=noise.shape, dtype=noise.dtype)  # the draw p_sample makes
        out[f"p_sample_{name}"] = gd.p_sample(noise, tt, k)
        out[f"p_sample_t_{name}"] = tt
    gd.use_dynamic_thres = True
    out["p_sample_dyn"] = gd.p_sample(noise * 3, np.array([199, 120], np.int32), jax.random.PRNGKey(11))
    gd.use_dynamic_thres = False
    for n in ref_gd.GaussianDiffusion.__init__.__code__.co_names:
        pass
    for n, v in nnx.state_paths(gd).items():
        if "." not in n:
            out["sched_" + n] = np.asarray(v.value)
    out["extract"] = ref_utils.extract(np.arange(10.0) * 1.5, np.array([2, 7], np.int32), (2, 1, 1, 1, 1))

    # ---- 3. stand-alone modules, called the way test_modules.py calls them ----
    def mod_params(m, tag):
        stm = nnx.state_paths(m)
        for path, var in stm.items():
            v = rng.standard_normal(var.shape) * (0.3 if path.endswith("kernel") or path.endswith("embedding") else 0.1) \
                + (1.0 if path.endswith("scale") else 0.0)
            out[f"{tag}__{path}"] = v.astype(np.float32)  # stored (and used) at float32 precision: half the fixture size
            var.value = out[f"{tag}__{path}"].astype(F64)

    mha = ref_modules.MultiheadAttention(in_features=32, dim=8, num_heads=4, rngs=nnx.Rngs(1))
    mod_params(mha, "mha")
    xm = rng.standard_normal((2, 6, 7, 5, 32))
    bias = rng.standard_normal((4, 5, 5))
    out.update(mha_x=xm, mha_bias=bias)
    out["mha_plain"] = mha(xm)
    out["mha_bias_out"] = mha(xm, pos_bias=bias)
    out["mha_allfocus"] = mha(xm, focus_present_mask=np.array([True, True]))
    with np.errstate(all="ignore"):
        out["mha_mixed"] = mha(xm, focus_present_mask=np.array([True, False]))
        out["mha_mixed_bias"] = mha(xm, focus_present_mask=np.array([False, True]), pos_bias=bias)

    rpb = ref_modules.RelativePositionBias(rngs=nnx.Rngs(2), heads=8, num_buckets=32, max_distance=32)
    mod_params(rpb, "rpb")
    out["rpb_10"] = rpb(10)
    out["rpb_40"] = rpb(40)
    rel = np.arange(-300, 301, dtype=np.int32)
    out["rpb_rel"] = rel
    out["rpb_buckets_f64log"] = np.asarray(ref_modules.RelativePositionBias._relative_position_bucket(rel))
    # the same formula with the float32 log the reference evaluates under jax's default x64-off mode
    n = -rel
    ret = (n < 0).astype(np.int32) * 16
    n = np.abs(n)
    with np.errstate(all="ignore"):
        large = 8 + (np.log(n.astype(np.float32) / np.float32(8)) / np.float32(np.log(128 / 8)) * np.float32(8)).astype(np.int32)
    out["rpb_buckets_f32log"] = ret + np.where(n < 8, n, np.minimum(large, 15))

    sla = ref_modules.SpatialLinearAttention(dim=32, heads=8, D=32, rngs=nnx.Rngs(3))
    mod_params(sla, "sla")
    xs = rng.standard_normal((1, 2, 8, 8, 32))
    out.update(sla_x=xs, sla_out=sla(xs))

    rb = ref_modules.ResnetBlock(32, 64, nnx.Rngs(4), time_emb_dim=128, groups=8)
    mod_params(rb, "rb")
    xr, te = rng.standard_normal((2, 2, 8, 8, 32)), rng.standard_normal((2, 128))
    out.update(rb_x=xr, rb_t=te, rb_out=rb(xr, te))
    rb2 = ref_modules.ResnetBlock(64, 64, nnx.Rngs(5), time_emb_dim=None, groups=8)
    mod_params(rb2, "rb2")
    out["rb2_out"] = rb2(out["rb_out"])

    pre = ref_modules.Residual(ref_modules.PreNorm(32, sla, rngs=nnx.Rngs(6)))
    out["residual_prenorm_sla"] = pre(xs)  # == sla(xs) + xs: the LayerNorm result is discarded (modules.py:146-148)

    out["sinusoidal"] = ref_modules.SinusoidalPosEmb(32)(np.array([0, 7, 199], np.int32))
    down, up = ref_utils.Downsample(32, nnx.Rngs(7)), ref_utils.Upsample(32, nnx.Rngs(8))
    mod_params(down, "down")
    mod_params(up, "up")
    xd = rng.standard_normal((1, 2, 8, 8, 32))
    out.update(updown_x=xd, down_out=down(xd), up_out=up(xd))

    np.savez_compressed(os.path.join(HERE, "ref_code_golden.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print(f"wrote {len(out)} arrays; unet leaves {len(tree)}, params {sum(int(np.prod(s)) for s in tree.values())}")


if __name__ == "__main__":
    main()
