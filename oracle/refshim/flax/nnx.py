"""refshim flax.nnx: the layers / containers the reference instantiates, on numpy, from the public flax semantics
(see ../README.md). Channels-last everywhere. Defaults as in flax: Conv / ConvTranspose padding='SAME', use_bias=True,
cross-correlation with kernel (*k, in, out); ConvTranspose transpose_kernel=False (kernel NOT flipped); GroupNorm /
LayerNorm epsilon 1e-6 with use_fast_variance=True (var = max(0, E[x^2] - E[x]^2)); Linear kernel (in, out);
LinearGeneral kernel (*in_dims, *out_dims); Embed = take; Sequential.layers."""
import math

import numpy as np

from jax import nn as _jnn


# ---------------------------------------------------------------------------------------------
# containers
# ---------------------------------------------------------------------------------------------
class Variable:
    def __init__(self, value):
        self.value = np.asarray(value)

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)

    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return self.value.dtype

    def __getitem__(self, i):
        return self.value[i]

    def _v(self, o):
        return o.value if isinstance(o, Variable) else o

    def __add__(self, o): return self.value + self._v(o)
    def __radd__(self, o): return self._v(o) + self.value
    def __sub__(self, o): return self.value - self._v(o)
    def __rsub__(self, o): return self._v(o) - self.value
    def __mul__(self, o): return self.value * self._v(o)
    def __rmul__(self, o): return self._v(o) * self.value
    def __truediv__(self, o): return self.value / self._v(o)
    def __rtruediv__(self, o): return self._v(o) / self.value
    def __matmul__(self, o): return self.value @ self._v(o)
    def __rmatmul__(self, o): return self._v(o) @ self.value
    def __neg__(self): return -self.value


class Param(Variable):
    pass


class Module:
    pass


class Rngs:
    def __init__(self, seed=0, **kw):
        self.seed = int(seed) if not kw else int(next(iter(kw.values())))
        self._n = 0
        self._rng = np.random.default_rng(self.seed)

    def params(self):
        self._n += 1
        return np.array([self.seed, self._n], dtype=np.uint32)

    def __getattr__(self, name):  # any other stream name
        return self.params

    # initialisers (flax defaults; the golden script overwrites every parameter with the oracle's)
    def lecun_normal(self, shape, fan_in):
        return self._rng.standard_normal(size=shape) * math.sqrt(1.0 / fan_in)


class Sequential(Module):
    def __init__(self, *fns):
        self.layers = list(fns)

    def __call__(self, *args, **kwargs):
        out = self.layers[0](*args, **kwargs)
        for f in self.layers[1:]:
            out = f(out)
        return out


def state_paths(module, prefix=()):
    """{dotted path: Variable} of every Variable reachable from `module`, with the path convention of nnx.State:
    attribute names, list indices, Sequential's `layers`."""
    out = {}

    def visit(obj, path):
        if isinstance(obj, Variable):
            out[".".join(str(p) for p in path)] = obj
        elif isinstance(obj, Module):
            for k, v in vars(obj).items():
                visit(v, path + (k,))
        elif isinstance(obj, (list, tuple)):
            for i, v in enumerate(obj):
                visit(v, path + (i,))

    visit(module, tuple(prefix))
    return out


def split(module, *a):
    return None, state_paths(module)


def merge(graphdef, state):
    raise NotImplementedError("refshim: call module methods directly")


# ---------------------------------------------------------------------------------------------
# activations
# ---------------------------------------------------------------------------------------------
silu = _jnn.silu
swish = _jnn.silu
gelu = _jnn.gelu
softmax = _jnn.softmax
sigmoid = _jnn.sigmoid


# ---------------------------------------------------------------------------------------------
# lax.conv_general_dilated from its definition (channels-last, kernel (*k, I, O), cross-correlation)
# ---------------------------------------------------------------------------------------------
def _same_pads(in_size, k, stride):
    """lax.padtype_to_pads('SAME')."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return total // 2, total - total // 2


def _conv_transpose_pads(k, s):
    """jax.lax._conv_transpose_padding(k, s, 'SAME')."""
    pad_len = k + s - 2
    pad_a = k - 1 if s > k - 1 else int(math.ceil(pad_len / 2))
    return pad_a, pad_len - pad_a


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None):
    """lhs (N, *spatial, I), rhs (*k, I, O) -> (N, *out_spatial, O).  1. dilate lhs (insert d-1 zeros between
    elements), 2. pad explicitly, 3. out[n, o..., :] = sum_taps lhs[n, o*stride + tap, :] @ rhs[tap]."""
    nd = rhs.ndim - 2
    ks = rhs.shape[:nd]
    lhs_dilation = lhs_dilation or (1,) * nd
    x = lhs
    for ax, d in enumerate(lhs_dilation):
        if d > 1:
            n = x.shape[1 + ax]
            shp = list(x.shape)
            shp[1 + ax] = (n - 1) * d + 1
            y = np.zeros(shp, dtype=x.dtype)
            idx = [slice(None)] * x.ndim
            idx[1 + ax] = slice(0, None, d)
            y[tuple(idx)] = x
            x = y
    x = np.pad(x, [(0, 0)] + [tuple(p) for p in padding] + [(0, 0)])
    out_sp = [(x.shape[1 + a] - ks[a]) // window_strides[a] + 1 for a in range(nd)]
    out = np.zeros([x.shape[0]] + out_sp + [rhs.shape[-1]], dtype=np.result_type(x.dtype, rhs.dtype))
    for tap in np.ndindex(*ks):
        idx = [slice(None)]
        for a in range(nd):
            idx.append(slice(tap[a], tap[a] + (out_sp[a] - 1) * window_strides[a] + 1, window_strides[a]))
        out += x[tuple(idx)] @ rhs[tap]
    return out


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class Conv(Module):
    def __init__(self, in_features, out_features, kernel_size, strides=1, *, padding="SAME", use_bias=True, rngs=None, **kw):
        ks = _tup(kernel_size, 1)
        self.kernel_size, self.strides, self.padding, self.use_bias = ks, _tup(strides, len(ks)), padding, use_bias
        self.in_features, self.out_features = in_features, out_features
        fan_in = in_features * int(np.prod(ks))
        self.kernel = Param(rngs.lecun_normal(ks + (in_features, out_features), fan_in))
        self.bias = Param(np.zeros((out_features,))) if use_bias else None

    def __call__(self, x):
        nd = len(self.kernel_size)
        x = np.asarray(x)
        lead = x.shape[: x.ndim - nd - 1]  # flax flattens every leading batch dimension
        xf = x.reshape((-1,) + x.shape[x.ndim - nd - 1:])
        assert self.padding == "SAME"
        pads = [_same_pads(xf.shape[1 + a], self.kernel_size[a], self.strides[a]) for a in range(nd)]
        y = conv_general_dilated(xf, self.kernel.value, self.strides, pads)
        if self.bias is not None:
            y = y + self.bias.value
        return y.reshape(lead + y.shape[1:])


class ConvTranspose(Module):
    def __init__(self, in_features, out_features, kernel_size, strides=None, *, padding="SAME", use_bias=True,
                 transpose_kernel=False, rngs=None, **kw):
        ks = _tup(kernel_size, 1)
        self.kernel_size, self.strides = ks, _tup(strides if strides is not None else 1, len(ks))
        assert padding == "SAME" and not transpose_kernel
        fan_in = in_features * int(np.prod(ks))
        self.kernel = Param(rngs.lecun_normal(ks + (in_features, out_features), fan_in))
        self.bias = Param(np.zeros((out_features,))) if use_bias else None

    def __call__(self, x):
        nd = len(self.kernel_size)
        x = np.asarray(x)
        lead = x.shape[: x.ndim - nd - 1]
        xf = x.reshape((-1,) + x.shape[x.ndim - nd - 1:])
        # jax.lax.conv_transpose: stride-1 conv of the lhs-dilated input with the UNFLIPPED kernel
        pads = [_conv_transpose_pads(self.kernel_size[a], self.strides[a]) for a in range(nd)]
        y = conv_general_dilated(xf, self.kernel.value, (1,) * nd, pads, lhs_dilation=self.strides)
        if self.bias is not None:
            y = y + self.bias.value
        return y.reshape(lead + y.shape[1:])


class Linear(Module):
    def __init__(self, in_features, out_features, *, use_bias=True, rngs=None, **kw):
        self.kernel = Param(rngs.lecun_normal((in_features, out_features), in_features))
        self.bias = Param(np.zeros((out_features,))) if use_bias else None

    def __call__(self, x):
        y = np.asarray(x) @ self.kernel.value
        return y + self.bias.value if self.bias is not None else y


class LinearGeneral(Module):
    def __init__(self, in_features, out_features, *, axis=-1, use_bias=True, rngs=None, **kw):
        self.in_dims, self.out_dims = _tup(in_features, 1), _tup(out_features, 1)
        self.axis = _tup(axis, 1)
        assert len(self.axis) == len(self.in_dims)
        self.kernel = Param(rngs.lecun_normal(self.in_dims + self.out_dims, int(np.prod(self.in_dims))))
        self.bias = Param(np.zeros(self.out_dims)) if use_bias else None

    def __call__(self, x):
        x = np.asarray(x)
        n = len(self.in_dims)
        assert tuple(a % x.ndim for a in self.axis) == tuple(range(x.ndim - n, x.ndim)), "contracted axes must be trailing"
        assert x.shape[x.ndim - n:] == self.in_dims
        y = np.tensordot(x, self.kernel.value, axes=(list(range(x.ndim - n, x.ndim)), list(range(n))))
        return y + self.bias.value if self.bias is not None else y


class Embed(Module):
    def __init__(self, num_embeddings, features, *, rngs=None, **kw):
        self.embedding = Param(rngs._rng.standard_normal(size=(num_embeddings, features)) / math.sqrt(num_embeddings))

    def __call__(self, ids):
        return self.embedding.value[np.asarray(ids)]


def _stats(x, axes):
    mean = x.mean(axis=axes, keepdims=True)
    mean2 = (x * x).mean(axis=axes, keepdims=True)
    return mean, np.maximum(0.0, mean2 - mean * mean)  # use_fast_variance=True


class LayerNorm(Module):
    def __init__(self, num_features, *, epsilon=1e-6, rngs=None, **kw):
        self.epsilon = epsilon
        self.scale = Param(np.ones((num_features,)))
        self.bias = Param(np.zeros((num_features,)))

    def __call__(self, x):
        x = np.asarray(x)
        mean, var = _stats(x, (-1,))
        return (x - mean) * (1.0 / np.sqrt(var + self.epsilon) * self.scale.value) + self.bias.value


class GroupNorm(Module):
    def __init__(self, num_features, num_groups=32, *, epsilon=1e-6, rngs=None, **kw):
        self.num_groups, self.epsilon = num_groups, epsilon
        self.scale = Param(np.ones((num_features,)))
        self.bias = Param(np.zeros((num_features,)))

    def __call__(self, x):
        x = np.asarray(x)
        C, G = x.shape[-1], self.num_groups
        g = x.reshape(x.shape[:-1] + (G, C // G))
        axes = tuple(range(1, x.ndim - 1)) + (g.ndim - 1,)  # every non-batch axis + the channels inside a group
        mean, var = _stats(g, axes)
        mean = np.broadcast_to(mean, g.shape[:1] + (1,) * (x.ndim - 2) + (G, C // G)).reshape((x.shape[0],) + (1,) * (x.ndim - 2) + (C,))
        var = np.broadcast_to(var, g.shape[:1] + (1,) * (x.ndim - 2) + (G, C // G)).reshape((x.shape[0],) + (1,) * (x.ndim - 2) + (C,))
        return (x - mean) * (1.0 / np.sqrt(var + self.epsilon) * self.scale.value) + self.bias.value
