"""refshim jax.nn (public definitions: softmax = exp(x - max) / sum; gelu default approximate=True = tanh form)."""
import numpy as _np


def softmax(x, axis=-1):
    x = _np.asarray(x)
    e = _np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def sigmoid(x):
    return 1.0 / (1.0 + _np.exp(-_np.asarray(x)))


def silu(x):
    return x * sigmoid(x)


swish = silu


def gelu(x, approximate=True):
    x = _np.asarray(x)
    if approximate:
        return 0.5 * x * (1.0 + _np.tanh(_np.sqrt(2.0 / _np.pi) * (x + 0.044715 * x ** 3)))
    from math import erf

    return 0.5 * x * (1.0 + _np.vectorize(erf)(x / _np.sqrt(2.0)))
