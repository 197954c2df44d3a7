"""Reference-side checkpoint converter (runs where the REFERENCE runs: needs jax, flax, orbax - none of which exist in
the image this repo was built in, so this script is untested there; it only uses the calls utils.py:431-508 itself makes).

  orbax -> npz:  python tools/convert_reference_checkpoint.py to-npz  <ckpt_dir> <step> out.npz
  npz -> orbax:  python tools/convert_reference_checkpoint.py to-orbax in.npz <ckpt_dir> [--config configs/config_v2_2.yaml]

The .npz layout is video_diffusion_nnx_b200/checkpoint.py's: key "<model|ema_params>/<nnx state path with '/'>", i.e. the
tree the reference saves (`{'model': state, 'ema_params': state}` of nnx.split(GaussianDiffusion), utils.py:445-448):
`denoise_fn/...` leaves plus the ten schedule tables. Load it here with
`video_diffusion_nnx_b200.checkpoint.load_checkpoint(path, load_ema_params=...)` -> `GaussianDiffusion.load_state_dict`.
"""
import argparse
import sys

import numpy as np


def _flatten(tree, prefix=()):
    out = {}
    if isinstance(tree, dict):
        for k, v in tree.items():
            out.update(_flatten(v, prefix + (str(k),)))
    elif hasattr(tree, "value") and not isinstance(tree, np.ndarray):  # nnx.VariableState
        out["/".join(prefix)] = np.asarray(tree.value)
    else:
        out["/".join(prefix)] = np.asarray(tree)
    return out


def to_npz(ckpt_dir, step, out_path):
    import orbax.checkpoint as ocp

    mgr = ocp.CheckpointManager(ckpt_dir, options=ocp.CheckpointManagerOptions())
    tree = mgr.restore(step)  # the raw pytree: {'model': {...}, 'ema_params': {...}}
    flat = {}
    for top in ("model", "ema_params"):
        sub = tree[top]
        sub = sub.to_pure_dict() if hasattr(sub, "to_pure_dict") else sub
        for k, v in _flatten(sub).items():
            flat[f"{top}/{k[:-len('/value')] if k.endswith('/value') else k}"] = v
    np.savez(out_path, __step__=np.asarray(step, np.int64), **flat)
    print(f"wrote {len(flat)} arrays to {out_path}")


def to_orbax(npz_path, ckpt_dir, config):
    import jax
    import orbax.checkpoint as ocp
    import yaml
    from flax import nnx
    from orbax.checkpoint import args as ocp_args

    sys.path.insert(0, ".")
    from gaussian_diffusion import GaussianDiffusion  # the reference's modules (run from the reference checkout)
    from unet3d import Unet3D

    cfg = yaml.safe_load(open(config))
    net = Unet3D(rngs=nnx.Rngs(0), **{k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg["unet"].items()})
    model = GaussianDiffusion(net, **cfg["diffusion"])
    _, state = nnx.split(model)
    with np.load(npz_path) as z:
        flat = {k: z[k] for k in z.files if k != "__step__"}
        step = int(z["__step__"]) if "__step__" in z.files else 0

    def fill(top):
        pure = state.to_pure_dict()

        def rec(node, path):
            for k, v in node.items():
                p = path + (str(k),)
                if isinstance(v, dict):
                    rec(v, p)
                else:
                    node[k] = jax.numpy.asarray(flat[top + "/" + "/".join(p)])

        rec(pure, ())
        st = nnx.State(jax.tree.map(lambda x: x, state))
        st.replace_by_pure_dict(pure)
        return st

    mgr = ocp.CheckpointManager(ckpt_dir, options=ocp.CheckpointManagerOptions(create=True))
    mgr.save(step, args=ocp_args.StandardSave({"model": fill("model"), "ema_params": fill("ema_params")}), force=True)
    mgr.wait_until_finished()
    print(f"wrote orbax checkpoint step {step} under {ckpt_dir}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("to-npz")
    a.add_argument("ckpt_dir")
    a.add_argument("step", type=int)
    a.add_argument("out")
    b = sub.add_parser("to-orbax")
    b.add_argument("npz")
    b.add_argument("ckpt_dir")
    b.add_argument("--config", default="configs/config_v2_2.yaml")
    args = ap.parse_args()
    if args.cmd == "to-npz":
        to_npz(args.ckpt_dir, args.step, args.out)
    else:
        to_orbax(args.npz, args.ckpt_dir, args.config)
