def _walk(tree, prefix=""):
    for k, v in tree.items():
        if isinstance(v, dict) or hasattr(v, "items"):
            yield from _walk(v, f"{prefix}{k}/")
        else:
            yield f"{prefix}{k}", tuple(getattr(v, "shape", ()))


def get_parameter_overview(params, **kw):
    return "\n".join(f"{n}: {s}" for n, s in _walk(params))


def log_parameter_overview(params, **kw):
    pass
