def replicate(x, *a, **k):
    return x


def unreplicate(x):
    return x
