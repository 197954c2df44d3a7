"""ctypes binding of the C ABI declared in include/vdn.h.

This is the harness-side binding used in this image (no jaxlib here); ffi/vdn_ffi.cc + jax_ffi.py are the
jax.ffi registration a maintainer of the reference would add for the same symbols. There is no
CPU or eager fallback: if libvdn.so is missing the import fails loudly.

Every exported function gets its `argtypes` / `restype` from the prototypes in include/vdn.h (parsed once at import),
so a call with the wrong number or kind of arguments fails in ctypes instead of corrupting the stack.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdn.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vdn.h")

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


_SCALARS = {
    "int": C.c_int, "long": C.c_long, "long long": C.c_longlong, "unsigned long long": C.c_ulonglong,
    "size_t": C.c_size_t, "float": C.c_float, "double": C.c_double, "unsigned": C.c_uint, "unsigned int": C.c_uint,
    "int64_t": C.c_int64, "uint64_t": C.c_uint64,
}


def _ctype_of(decl: str):
    """C parameter / return declaration -> ctypes type. Every pointer is passed as an address (c_void_p), except
    `const char*` (c_char_p)."""
    d = decl.strip()
    if "*" in d:
        base = d[: d.index("*")].replace("const", "").strip()
        return C.c_char_p if base == "char" else C.c_void_p
    words = [w for w in d.replace("const", "").split() if w]
    for n in (len(words), len(words) - 1):  # with or without a trailing parameter name
        t = " ".join(words[:n])
        if t in _SCALARS:
            return _SCALARS[t]
    if d == "void":
        return None
    raise VdnError(f"include/vdn.h: cannot map C declaration {decl!r} to ctypes")


def parse_prototypes(header_text: str):
    """{name: (restype, [argtypes])} of every `vdn_*` function prototype in the header."""
    src = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(vdn_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        argt = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        out[name] = (_ctype_of(ret), argt)
    return out


def _bind(lib):
    protos = parse_prototypes(open(HEADER_PATH).read())
    missing = []
    for name, (ret, argt) in protos.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = ret
        fn.argtypes = argt
    if missing:
        raise VdnError(f"libvdn.so does not export {missing}: rebuild it (include/vdn.h is newer than the library)")
    return protos


lib = _load()
PROTOTYPES = _bind(lib)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise VdnError(f"{what} failed with status {rc}: {lib.vdn_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------
# experiment / test switches: never read from the environment by the product
# ------------------------------------------------------------------------------------------
def debug_set(name: str, value: int = 1) -> None:
    check(lib.vdn_debug_set(name.encode(), int(value), 1), "vdn_debug_set")


def debug_clear(name: str) -> None:
    check(lib.vdn_debug_set(name.encode(), 0, 0), "vdn_debug_set")


@contextlib.contextmanager
def debug_switches(**kv):
    """with debug_switches(VDN_SLAB_MIN_ITEMS=1, VDN_SLAB_GRID=3): ...  (tests and tools only)."""
    for k, v in kv.items():
        debug_set(k, v)
    try:
        yield
    finally:
        for k in kv:
            debug_clear(k)


_APPLIED = set()


def apply_env_switches() -> None:
    """tools/ only: mirror every integer-valued VDN_* environment variable into the library's switch table (and clear
    the ones a previous call had set). The library itself never looks at the environment."""
    now = {k: v for k, v in os.environ.items() if k.startswith("VDN_") and v.lstrip("-").isdigit()}
    for k in _APPLIED - set(now):
        debug_clear(k)
    for k, v in now.items():
        debug_set(k, int(v))
    _APPLIED.clear()
    _APPLIED.update(now)


# host-side (Python) switches used by tools/: plain module state, seeded from the environment only when the
# process opts in with VDN_DEBUG=1
_HOST_FLAGS = {}


def host_flag(name: str, default=None):
    if name in _HOST_FLAGS:
        return _HOST_FLAGS[name]
    if os.environ.get("VDN_DEBUG") == "1":
        return os.environ.get(name, default)
    return default


def set_host_flag(name: str, value) -> None:
    _HOST_FLAGS[name] = value
