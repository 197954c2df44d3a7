"""ctypes binding of the C ABI declared in include/vdn.h.

This is the harness-side binding used in this image (no jaxlib here); INTEGRATION.md shows the
jax.ffi registration a maintainer of the reference would add for the same symbols. There is no
CPU or eager fallback: if libvdn.so is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdn.so")

VDN_TAP_UNIT, VDN_TAP_DOWN, VDN_TAP_UP = 0, 1, 2
VDN_BF16, VDN_F32 = 0, 1


class VdnError(RuntimeError):
    pass


class TapGemmDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("n_img", C.c_int),
        ("H", C.c_int),
        ("W", C.c_int),
        ("n_src", C.c_int),
        ("src_c", C.c_int),
        ("n_taps", C.c_int),
        ("tap_dy", C.c_int * 16),
        ("tap_dx", C.c_int * 16),
        ("n_out", C.c_int),
        ("py", C.c_int),
        ("px", C.c_int),
        ("out_dtype", C.c_int),
        ("split_col", C.c_int),
        ("gn_groups", C.c_int),
        ("rows_per_sample", C.c_int),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise VdnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no fallback path)"
        )
    return C.CDLL(LIB_PATH)


lib = _load()
lib.vdn_last_error.restype = C.c_char_p
lib.vdn_version.restype = C.c_int
lib.vdn_launch_count.restype = C.c_ulonglong


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise VdnError(f"{what} failed with status {rc}: {lib.vdn_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
