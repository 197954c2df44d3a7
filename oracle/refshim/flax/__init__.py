from . import nnx  # noqa: F401
