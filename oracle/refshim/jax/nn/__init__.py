"""jax.nn subset."""
import numpy as np

from . import initializers  # noqa: F401


def _f(x):
    x = np.asarray(x)
    return x.astype(np.float32) if x.dtype.itemsize < 4 and x.dtype.kind not in "iub" else x


def one_hot(x, num_classes, dtype=np.float32, axis=-1):
    """Out-of-range indices (e.g. -1) give an all-zero row, as jax.nn.one_hot."""
    x = np.asarray(x)
    return (x[..., None] == np.arange(num_classes)).astype(dtype)


def softmax(x, axis=-1):
    x0 = np.asarray(x)
    xf = _f(x0)
    e = np.exp(xf - xf.max(axis=axis, keepdims=True))
    return (e / e.sum(axis=axis, keepdims=True)).astype(x0.dtype)


def logsumexp(x, axis=None, keepdims=False):
    x0 = np.asarray(x)
    xf = _f(x0)
    m = xf.max(axis=axis, keepdims=True)
    out = np.log(np.exp(xf - m).sum(axis=axis, keepdims=True)) + m
    if not keepdims:
        out = np.squeeze(out, axis=axis) if axis is not None else out.reshape(())
    return out.astype(x0.dtype)


def log_softmax(x, axis=-1):
    x0 = np.asarray(x)
    return (_f(x0) - _f(logsumexp(x0, axis=axis, keepdims=True))).astype(x0.dtype)


def sigmoid(x):
    x0 = np.asarray(x)
    return (1.0 / (1.0 + np.exp(-_f(x0)))).astype(x0.dtype)


def gelu(x, approximate=True):
    x0 = np.asarray(x)
    xf = _f(x0)
    return (0.5 * xf * (1 + np.tanh(np.sqrt(2 / np.pi) * (xf + 0.044715 * xf ** 3)))).astype(x0.dtype)


def _wrap_public(ns, names):
    import functools
    from .._array import unwrap, wrap

    def deco(fn):
        @functools.wraps(fn)
        def f(*a, **k):
            return wrap(fn(*unwrap(a), **unwrap(k)))
        return f
    for n in names:
        ns[n] = deco(ns[n])


_wrap_public(globals(), ["one_hot", "softmax", "logsumexp", "log_softmax", "sigmoid", "gelu"])
