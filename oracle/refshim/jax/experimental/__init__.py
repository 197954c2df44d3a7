from . import mesh_utils, pjit  # noqa: F401
