class FrozenDict(dict):
    pass


def freeze(x):
    return FrozenDict((k, freeze(v) if isinstance(v, dict) else v) for k, v in x.items())


def unfreeze(x):
    return {k: unfreeze(v) if isinstance(v, dict) else v for k, v in x.items()}
