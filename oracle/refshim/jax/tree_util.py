"""jax.tree_util subset over dict / list / tuple pytrees."""


def tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return type(tree)((k, tree_map(f, v, *[r[k] for r in rest])) for k, v in tree.items())
    if isinstance(tree, (list, tuple)):
        out = [tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree)]
        return type(tree)(out) if not hasattr(tree, "_fields") else type(tree)(*out)
    if tree is None:
        return None
    return f(tree, *rest)


tree_multimap = tree_map


def tree_leaves(tree):
    out = []
    tree_map(out.append, tree)
    return out


def tree_flatten(tree):
    return tree_leaves(tree), None


def tree_reduce(f, tree, initializer=None):
    import functools
    leaves = tree_leaves(tree)
    return functools.reduce(f, leaves) if initializer is None else functools.reduce(f, leaves, initializer)
