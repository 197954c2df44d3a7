"""jax.ffi registration of the C ABI (include/vdn.h) for hosts that have jaxlib - the binding a maintainer of the
reference adds so that the kernels run INSIDE the reference's own pjit'd train step (trainer.py:337-361) and
`pjit_step` of p_sample_loop (gaussian_diffusion.py:299-301) as XLA custom calls.

NOT runnable in the image this repo was developed in (no jax / jaxlib / XLA headers): importing
this module is safe everywhere, `register()` raises with the reason when jax or the compiled shim is missing. The C side
is ffi/vdn_ffi.cc (one XLA_FFI handler symbol `<name>_ffi` per entry point); build it where jaxlib is installed:

    g++ -O2 -std=c++17 -shared -fPIC ffi/vdn_ffi.cc -Iinclude -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \\
        -I/usr/local/cuda/include -Lvideo_diffusion_nnx_b200 -lvdn -Wl,-rpath,'$ORIGIN' \\
        -o video_diffusion_nnx_b200/libvdn_ffi.so

Usage inside the reference (modules.py), e.g. Block.__call__'s GroupNorm + scale/shift + SiLU (modules.py:171-179):

    from video_diffusion_nnx_b200 import jax_ffi
    jax_ffi.register()
    y = jax_ffi.gn_silu(x_raw, sums, gamma, beta, scale_shift, groups=8)     # differentiable (custom_vjp)

Buffers are owned by XLA; optional operands of the C ABI are passed as zero-element arrays; scratch is an extra result
sized by the vdn_*_workspace() queries (exposed here as `workspace_bytes`).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB_PATH = os.path.join(_HERE, "libvdn_ffi.so")
_registered = False


def _targets(lib):
    lib.vdn_ffi_targets.restype = ctypes.POINTER(ctypes.c_char_p)
    arr = lib.vdn_ffi_targets()
    out, i = [], 0
    while arr[i]:
        out.append(arr[i].decode())
        i += 1
    return out


def register(lib_path: str = FFI_LIB_PATH):
    """Registers every `<name>_ffi` handler of libvdn_ffi.so as the CUDA FFI target `<name>`. Idempotent."""
    global _registered
    if _registered:
        return
    try:
        import jax
    except ImportError as e:  # pragma: no cover - the development image has no jax
        raise RuntimeError("jax is not installed: the jax.ffi binding needs a host with jaxlib (use the ctypes binding "
                           "video_diffusion_nnx_b200._lib / ops here)") from e
    if not os.path.exists(lib_path):
        raise RuntimeError(f"{lib_path} not found: build ffi/vdn_ffi.cc against jaxlib's headers (see this module's docstring)")
    lib = ctypes.CDLL(lib_path)
    lib.vdn_ffi_available.restype = ctypes.c_int
    if not lib.vdn_ffi_available():
        raise RuntimeError("libvdn_ffi.so was compiled without xla/ffi/api/ffi.h (stub build): rebuild it where jaxlib is installed")
    for name in _targets(lib):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name + "_ffi")), platform="CUDA")
    _registered = True


def workspace_bytes(op: str, *args) -> int:
    """Scratch size of entry point `op` (e.g. "sla_core_fwd", n_img, N): allocate it as an extra uint8 result."""
    from ._lib import lib

    return int(getattr(lib, f"vdn_{op}_workspace")(*args))


# ------------------------------------------------------------------------------------------
# differentiable wrappers (jax.custom_vjp pairing forward / backward targets); built lazily so that importing this
# module never needs jax
# ------------------------------------------------------------------------------------------
def _build_wrappers():
    import jax
    import jax.numpy as jnp
    import numpy as np

    def _empty(dtype=jnp.float32):
        return jnp.zeros((0,), dtype)

    def call(name, out_types, *operands, **attrs):
        return jax.ffi.ffi_call(name, out_types)(*operands, **attrs)

    # ---- GroupNorm + scale/shift + SiLU (modules.py:171-179) ----
    @jax.custom_vjp
    def gn_silu(x_raw, sums, gamma, beta, scale_shift, groups=8):
        return _gn_silu_fwd(x_raw, sums, gamma, beta, scale_shift, groups)[0]

    def _gn_silu_fwd(x_raw, sums, gamma, beta, scale_shift, groups):
        B, C = x_raw.shape[0], x_raw.shape[-1]
        rows = int(np.prod(x_raw.shape[1:-1]))
        ss = _empty() if scale_shift is None else scale_shift
        y = call("vdn_gn_silu_fwd", jax.ShapeDtypeStruct(x_raw.shape, jnp.bfloat16), x_raw, sums, gamma, beta, ss,
                 B=B, rows=rows, C=C, G=groups, ss_ld=0 if scale_shift is None else scale_shift.shape[-1])
        return y, (x_raw, sums, gamma, beta, scale_shift, groups)

    def _gn_silu_bwd(res, dy):
        x_raw, sums, gamma, beta, scale_shift, groups = res
        B, C = x_raw.shape[0], x_raw.shape[-1]
        rows = int(np.prod(x_raw.shape[1:-1]))
        ss = _empty() if scale_shift is None else scale_shift
        f32 = jnp.float32
        outs = (jax.ShapeDtypeStruct((B, C, 2), f32), jax.ShapeDtypeStruct(x_raw.shape, jnp.bfloat16),
                jax.ShapeDtypeStruct((C,), f32), jax.ShapeDtypeStruct((C,), f32),
                jax.ShapeDtypeStruct((B, 2 * C) if scale_shift is not None else (0,), f32), jax.ShapeDtypeStruct((C,), f32))
        _, dx, dgamma, dbeta, dss, _ = call("vdn_gn_silu_bwd", outs, dy, x_raw, sums, gamma, beta, ss, B=B, rows=rows, C=C,
                                           G=groups, ss_ld=0 if scale_shift is None else scale_shift.shape[-1], dss_ld=2 * C)
        # the statistics are produced by the conv epilogue and consumed here: their cotangent flows through dx
        return dx, jnp.zeros_like(sums), dgamma, dbeta, (None if scale_shift is None else dss), None

    gn_silu.defvjp(lambda *a: _gn_silu_fwd(*a), _gn_silu_bwd)

    # ---- the tap-GEMM: conv (1,3,3) / 1x1 / Linear / Downsample / Upsample forward (modules.py:162-165 ...) ----
    def tapgemm(srcs, wp, bias, taps, *, kind=0, n_out, out_dtype=jnp.bfloat16, residual=None, gn_groups=0,
                rows_per_sample=0, py=0, px=0):
        x0 = srcs[0]
        n_img, Hs, Ws, C = x0.shape
        H, W = (Hs // 2, Ws // 2) if kind == 1 else (Hs, Ws)
        oh, ow = (2 * H, 2 * W) if kind == 2 else (H, W)
        B = max(1, (n_img * H * W) // max(rows_per_sample, 1))
        outs = (jax.ShapeDtypeStruct((n_img, oh, ow, n_out), out_dtype), jax.ShapeDtypeStruct((0,), out_dtype),
                jax.ShapeDtypeStruct((16, B, gn_groups, 2) if gn_groups else (0,), jnp.float32))
        return call("vdn_tapgemm", outs, x0, srcs[1] if len(srcs) > 1 else _empty(x0.dtype), wp,
                    _empty() if bias is None else bias, _empty(out_dtype) if residual is None else residual, _empty(out_dtype),
                    kind=kind, n_img=n_img, H=H, W=W, n_src=len(srcs), src_c=C,
                    tap_dy=np.asarray([t[0] for t in taps], np.int32), tap_dx=np.asarray([t[1] for t in taps], np.int32),
                    n_out=n_out, py=py, px=px, out_dtype=1 if out_dtype == jnp.float32 else 0, split_col=0,
                    gn_groups=gn_groups, rows_per_sample=rows_per_sample)

    return {"gn_silu": gn_silu, "tapgemm": tapgemm, "call": call}


_wrappers = None


def __getattr__(name):  # PEP 562: jax_ffi.gn_silu / jax_ffi.tapgemm / jax_ffi.call resolve on first use
    global _wrappers
    if name in ("gn_silu", "tapgemm", "call"):
        if _wrappers is None:
            register()
            _wrappers = _build_wrappers()
        return _wrappers[name]
    raise AttributeError(name)
