"""refshim jax.random: deterministic functions of the key (NOT threefry: draws differ from JAX's; the golden script
records every draw it consumes, so nothing downstream depends on the stream)."""
import numpy as _np


def PRNGKey(seed):
    return _np.array([0, int(seed) & 0xFFFFFFFF], dtype=_np.uint32)


key = PRNGKey


def _rng(k):
    return _np.random.default_rng([int(v) for v in _np.asarray(k).ravel()])


def split(k, num=2):
    base = [int(v) for v in _np.asarray(k).ravel()]
    return _np.stack([_np.array(_np.random.SeedSequence(base + [i]).generate_state(2), dtype=_np.uint32) for i in range(num)])


def normal(k, shape=(), dtype=_np.float64):
    return _rng(k).standard_normal(size=tuple(shape)).astype(dtype)


def uniform(key=None, shape=(), dtype=_np.float64, minval=0.0, maxval=1.0):
    return (_rng(key).random(size=tuple(shape)) * (maxval - minval) + minval).astype(dtype)


def randint(k, shape, minval, maxval, dtype=_np.int32):
    return _rng(k).integers(minval, maxval, size=tuple(shape)).astype(dtype)


def choice(k, a, shape=()):
    a = _np.asarray(a)
    return a[_rng(k).integers(0, a.shape[0], size=tuple(shape))]
