from .tree_util import tree_leaves as leaves  # noqa: F401
from .tree_util import tree_map as map  # noqa: F401,A001
