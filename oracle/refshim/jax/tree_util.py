"""refshim jax.tree_util on nested dict / list / tuple pytrees."""


def tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    return f(tree, *rest)


def tree_leaves(tree):
    if isinstance(tree, dict):
        return [l for v in tree.values() for l in tree_leaves(v)]
    if isinstance(tree, (list, tuple)):
        return [l for v in tree for l in tree_leaves(v)]
    return [tree]


def tree_reduce(f, tree, init):
    acc = init
    for l in tree_leaves(tree):
        acc = f(acc, l)
    return acc


class DictKey:
    def __init__(self, key):
        self.key = key


class SequenceKey:
    def __init__(self, idx):
        self.idx = idx


def tree_leaves_with_path(tree, prefix=()):
    if isinstance(tree, dict):
        return [x for k, v in tree.items() for x in tree_leaves_with_path(v, prefix + (DictKey(k),))]
    if isinstance(tree, (list, tuple)):
        return [x for i, v in enumerate(tree) for x in tree_leaves_with_path(v, prefix + (SequenceKey(i),))]
    return [(prefix, tree)]
