"""Checkpoint bridge.

What the reference checkpoints: ONE pytree `{'model': state, 'ema_params': state}` written through orbax
(utils.py:431-456), where `state` is `nnx.split(GaussianDiffusion(...))[1]` (trainer.py:136, 287, 600; sample.py:97
restores against the same split, utils.py:484-505). Its leaves are therefore
  * `denoise_fn/<Unet3D nnx path>`  - every Unet3D parameter, and
  * the ten schedule Variables of gaussian_diffusion.py:85-98 (`alphas_cumprod`, `sqrt_alphas_cumprod`, ...), which the
    reference trainer Adam-updates like any other leaf, so a trained checkpoint carries CHANGED tables.
orbax is not in this image, so the same tree travels as a flat `.npz`:
  key   = "<'model'|'ema_params'>/<state path joined with '/'>"   (a trailing "/value" - how some flax versions spell
          the VariableState leaf - is accepted on load and optionally written on save)
  value = the array in flax layout.
`tools/convert_reference_checkpoint.py` is the reference-side converter (orbax directory <-> this .npz); it needs
jax/flax/orbax and is therefore run where the reference runs. Pure host IO; nothing here touches the GPU path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

TREES = ("model", "ema_params")
UNET_PREFIX = "denoise_fn."
SCHEDULE_NAMES = ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                  "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                  "posterior_mean_coef2")


def diffusion_state(unet_state: Dict[str, np.ndarray], schedule: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """The state of `nnx.split(GaussianDiffusion)`: Unet3D leaves under `denoise_fn.` + the ten schedule tables."""
    out = {UNET_PREFIX + k: np.asarray(v) for k, v in unet_state.items()}
    for n in SCHEDULE_NAMES:
        out[n] = np.asarray(schedule[n], np.float32)
    return out


def split_diffusion_state(state: Dict[str, np.ndarray]) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """Inverse of diffusion_state(): (Unet3D state keyed by bare nnx paths, schedule tables present in `state`).
    A bare Unet3D state (no `denoise_fn.` prefix, no tables) passes through unchanged."""
    unet, sched = {}, {}
    for k, v in state.items():
        if k in SCHEDULE_NAMES:
            sched[k] = np.asarray(v, np.float32)
        elif k.startswith(UNET_PREFIX):
            unet[k[len(UNET_PREFIX):]] = np.asarray(v)
        else:
            unet[k] = np.asarray(v)
    return unet, sched


def flatten_tree(tree: Dict[str, Dict[str, np.ndarray]], value_suffix: bool = False) -> Dict[str, np.ndarray]:
    """{'model': {'denoise_fn.downs.0.0.block_1.proj.kernel': a, ...}, 'ema_params': {...}}
    -> {'model/denoise_fn/downs/0/0/...': a}."""
    out = {}
    for top in TREES:
        for path, arr in tree.get(top, {}).items():
            out[top + "/" + path.replace(".", "/") + ("/value" if value_suffix else "")] = np.asarray(arr)
    return out


def unflatten_tree(flat: Dict[str, np.ndarray]) -> Dict[str, Dict[str, np.ndarray]]:
    tree: Dict[str, Dict[str, np.ndarray]] = {t: {} for t in TREES}
    for key, arr in flat.items():
        top, _, rest = key.partition("/")
        if top in tree:
            if rest.endswith("/value"):
                rest = rest[: -len("/value")]
            tree[top][rest.replace("/", ".")] = np.asarray(arr)
    return tree


def nest(state: Dict[str, np.ndarray]) -> dict:
    """Dotted nnx paths -> nested dicts (the shape of `nnx.State.to_pure_dict()` / an orbax StandardSave item)."""
    root: dict = {}
    for path, arr in state.items():
        node = root
        parts = path.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = arr
    return root


def unnest(tree: dict, prefix: str = "") -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for k, v in tree.items():
        path = f"{prefix}.{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(unnest(v, path))
        else:
            out[path] = np.asarray(v)
    return out


def save_checkpoint(path: str, model_state: Dict[str, np.ndarray], ema_state: Optional[Dict[str, np.ndarray]] = None,
                    step: int = 0, value_suffix: bool = False) -> None:
    """utils.py:431-456: both trees in one file (EMA defaults to the model parameters, as at initialisation).
    Pass the FULL diffusion state (`GaussianDiffusion.state_dict()` / `diffusion_state()`), which is what the
    reference's StandardRestore template (utils.py:486-496) requires."""
    flat = flatten_tree({"model": model_state, "ema_params": model_state if ema_state is None else ema_state},
                        value_suffix)
    np.savez(path, __step__=np.asarray(step, np.int64), **flat)


def load_checkpoint(path: str, load_ema_params: bool = False) -> Tuple[Dict[str, np.ndarray], int]:
    """utils.py:460-505: returns the state to merge into the model (`load_ema_params` picks the EMA tree, the
    `sample.py --load-ema-params` switch) and the saved step. Feed the result to `GaussianDiffusion.load_state_dict`
    (Unet3D leaves + trained schedule tables) or, for a bare Unet3D state, to `Unet3D.load_state_dict`."""
    with np.load(path) as z:
        flat = {k: z[k] for k in z.files if k != "__step__"}
        step = int(z["__step__"]) if "__step__" in z.files else 0
    tree = unflatten_tree(flat)
    chosen = tree["ema_params"] if load_ema_params else tree["model"]
    if not chosen:
        raise KeyError(f"{path}: no {'ema_params' if load_ema_params else 'model'} tree in the checkpoint")
    return chosen, step
