"""Checkpoint bridge: the reference saves ONE pytree `{'model': state, 'ema_params': state}`
through orbax (utils.py:445-455) and restores it into `nnx.split(model)` templates (utils.py:484-505). orbax is not
in this image, so the same tree travels as a flat `.npz`: key = "<'model'|'ema_params'>/<nnx state path with '/'>",
value = the array in flax layout - exactly what `jax.tree_util.tree_flatten_with_path` gives for the orbax tree, so a
reference-side converter is `{'/'.join(path): leaf}` in one direction and `nnx.State` from nested dicts in the other.
Pure host IO; nothing here touches the GPU path."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

TREES = ("model", "ema_params")


def flatten_tree(tree: Dict[str, Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    """{'model': {'downs.0.0.block_1.proj.kernel': a, ...}, 'ema_params': {...}} -> {'model/downs/0/0/...': a}."""
    out = {}
    for top in TREES:
        for path, arr in tree.get(top, {}).items():
            out[top + "/" + path.replace(".", "/")] = np.asarray(arr)
    return out


def unflatten_tree(flat: Dict[str, np.ndarray]) -> Dict[str, Dict[str, np.ndarray]]:
    tree: Dict[str, Dict[str, np.ndarray]] = {t: {} for t in TREES}
    for key, arr in flat.items():
        top, _, rest = key.partition("/")
        if top in tree:
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
                    step: int = 0) -> None:
    """utils.py:425-458: both trees in one file (EMA defaults to the model parameters, as at initialisation)."""
    flat = flatten_tree({"model": model_state, "ema_params": model_state if ema_state is None else ema_state})
    np.savez(path, __step__=np.asarray(step, np.int64), **flat)


def load_checkpoint(path: str, load_ema_params: bool = False) -> Tuple[Dict[str, np.ndarray], int]:
    """utils.py:460-505: returns the state to merge into the model (`load_ema_params` picks the EMA tree, the
    `sample.py --load-ema-params` switch) and the saved step. Feed the result to `Unet3D.load_state_dict`."""
    with np.load(path) as z:
        flat = {k: z[k] for k in z.files if k != "__step__"}
        step = int(z["__step__"]) if "__step__" in z.files else 0
    tree = unflatten_tree(flat)
    chosen = tree["ema_params"] if load_ema_params else tree["model"]
    if not chosen:
        raise KeyError(f"{path}: no {'ema_params' if load_ema_params else 'model'} tree in the checkpoint")
    return chosen, step
