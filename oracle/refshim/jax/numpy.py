"""refshim jax.numpy = numpy (2.x has concat / astype / bool); dtype follows the operands (no x64 switch)."""
import numpy as _np
from numpy import *  # noqa: F401,F403

ndarray = _np.ndarray
bool = _np.bool_  # noqa: A001
pi = _np.pi


def astype(x, dtype):
    return _np.asarray(x).astype(dtype)


def concat(arrays, axis=0):
    return _np.concatenate(arrays, axis=axis)


def __getattr__(name):
    return getattr(_np, name)


def linspace(start, stop, num=50, endpoint=True, dtype=None, **kw):
    """jax runs the reference with x64 disabled: a float64 request silently yields float32 (utils.py:252 asks for
    float64). Everything computed from the result stays float32 under numpy's weak-scalar promotion."""
    if dtype is not None and _np.dtype(dtype) == _np.float64:
        dtype = _np.float32
    return _np.linspace(start, stop, num, endpoint=endpoint, dtype=dtype, **kw)
