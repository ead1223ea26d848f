"""jax.random subset.  Keys are small uint32 arrays; draws come from numpy's PCG64 seeded by the key, NOT from
threefry -- every comparison that involves randomness must feed its draws as inputs (see `override`)."""
import contextlib
import zlib

import numpy as np

_overrides = {}


@contextlib.contextmanager
def override(**fns):
    """`with jax.random.override(uniform=lambda key, shape, ...: arr):` replaces a sampler for the block."""
    old = dict(_overrides)
    _overrides.update(fns)
    try:
        yield
    finally:
        _overrides.clear()
        _overrides.update(old)


def PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], np.uint32)


def _gen(key):
    return np.random.default_rng([int(v) & 0xFFFFFFFFFFFFFFFF for v in np.asarray(key).reshape(-1)])


def split(key, num=2):
    k = np.asarray(key).reshape(-1)
    return np.stack([np.array([zlib.crc32(bytes(k.tobytes()) + bytes([i & 255, i >> 8])) & 0xFFFFFFFF, (int(k[-1]) + i + 1) & 0xFFFFFFFF],
                              np.uint32) for i in range(num)], 0)


def fold_in(key, data):
    k = np.asarray(key).reshape(-1)
    return np.array([zlib.crc32(k.tobytes() + int(data).to_bytes(8, "little", signed=True)) & 0xFFFFFFFF, int(k[-1]) & 0xFFFFFFFF],
                    np.uint32)


def normal(key, shape=(), dtype=np.float32):
    if "normal" in _overrides:
        return _overrides["normal"](key, shape, dtype)
    return _gen(key).standard_normal(shape).astype(dtype)


def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    if "uniform" in _overrides:
        return _overrides["uniform"](key, shape, dtype, minval, maxval)
    return (_gen(key).random(shape) * (maxval - minval) + minval).astype(dtype)


def truncated_normal(key, lower, upper, shape=(), dtype=np.float32):
    g = _gen(key)
    x = g.standard_normal(shape)
    bad = (x < lower) | (x > upper)
    while bad.any():
        x[bad] = g.standard_normal(int(bad.sum()))
        bad = (x < lower) | (x > upper)
    return x.astype(dtype)


def choice(key, a, shape=(), replace=True, p=None):
    if "choice" in _overrides:
        return _overrides["choice"](key, a, shape, replace, p)
    return _gen(key).choice(a, size=tuple(shape), replace=replace, p=p).astype(np.int32)


def randint(key, shape, minval, maxval, dtype=np.int32):
    return _gen(key).integers(minval, maxval, size=shape).astype(dtype)


def bernoulli(key, p=0.5, shape=()):
    return _gen(key).random(shape) < p


def _wrap_public(ns, names):
    import functools
    from ._array import unwrap, wrap

    def deco(fn):
        @functools.wraps(fn)
        def f(*a, **k):
            return wrap(fn(*unwrap(a), **unwrap(k)))
        return f
    for n in names:
        ns[n] = deco(ns[n])


_wrap_public(globals(), ["PRNGKey", "split", "fold_in", "normal", "uniform", "truncated_normal", "choice", "randint", "bernoulli"])
