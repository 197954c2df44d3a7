"""refshim: the slice of the `jax` namespace the reference touches, on numpy (see ../README.md)."""
import numpy as _np

Array = _np.ndarray

from . import numpy  # noqa: E402,F401
from . import random  # noqa: E402,F401
from . import nn  # noqa: E402,F401
from . import tree_util  # noqa: E402,F401
from . import sharding  # noqa: E402,F401
from . import experimental  # noqa: E402,F401
from . import tree  # noqa: E402,F401


def local_devices():
    return ["cpu:0"]


def local_device_count():
    return 1


def devices():
    return ["cpu:0"]


def device_put_replicated(x, devices):
    return x


def device_put(x, *a, **k):
    return x


def device_get(x):
    return x


def jit(f=None, **kw):
    return f if f is not None else (lambda g: g)
