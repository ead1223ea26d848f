"""`Array`: an ndarray subclass that gives numpy jax's default "x64 disabled" type behaviour.

In jax (default config) float64 / int64 / uint64 do not exist: a np.float64 scalar such as `np.log(4)` multiplied
into a float32 array gives float32, integer reductions stay int32.  Plain numpy (NEP 50) would promote those to
64 bit, which changes e.g. where fp32 ties fall in pretrain_model.py:220-230.  Every array the shim hands to the
reference code is an `Array`; ufuncs and numpy API functions on it first demote 64-bit operands and then demote the
result, and return `Array` again.  Python scalars stay weakly typed, as in jax.
"""
import numpy as np

_DEMOTE = {np.dtype(np.float64): np.float32, np.dtype(np.int64): np.int32, np.dtype(np.uint64): np.uint32,
           np.dtype(np.complex128): np.complex64}


def _demote_value(x):
    if isinstance(x, np.ndarray):
        x = x.view(np.ndarray)
        t = _DEMOTE.get(x.dtype)
        return x.astype(t) if t is not None else x
    if isinstance(x, np.generic):
        t = _DEMOTE.get(x.dtype)
        return t(x) if t is not None else x
    return x


def unwrap(x):
    """Array / 64-bit numpy values -> plain demoted ndarray (containers handled recursively)."""
    if isinstance(x, (np.ndarray, np.generic)):
        return _demote_value(x)
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(*[unwrap(v) for v in x])
    if isinstance(x, (list, tuple)):
        return type(x)(unwrap(v) for v in x)
    if isinstance(x, dict):
        return {k: unwrap(v) for k, v in x.items()}
    return x


def wrap(x):
    if isinstance(x, np.ndarray):
        return _demote_value(x).view(Array)
    if isinstance(x, np.generic):
        return np.asarray(_demote_value(x)).view(Array)
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(*[wrap(v) for v in x])
    if isinstance(x, (list, tuple)):
        return type(x)(wrap(v) for v in x)
    return x


class Array(np.ndarray):
    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if out is not None:                                   # in-place (x += y): write through, numpy semantics
            outs = tuple(o.view(np.ndarray) if isinstance(o, np.ndarray) else o for o in out)
            res = getattr(ufunc, method)(*unwrap(inputs), out=outs, **unwrap(kwargs))
            return out[0] if len(out) == 1 else out
        return wrap(getattr(ufunc, method)(*unwrap(inputs), **unwrap(kwargs)))

    def __array_function__(self, func, types, args, kwargs):
        return wrap(func(*unwrap(args), **unwrap(kwargs)))

    # jax arrays are immutable: `x += y` rebinds instead of mutating the operand the caller may still hold
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def __iand__(self, o):
        return self & o

    def __ior__(self, o):
        return self | o

    def astype(self, dtype, *a, **k):
        return wrap(self.view(np.ndarray).astype(dtype, *a, **k))

    def argmax(self, *a, **k):
        return wrap(self.view(np.ndarray).argmax(*a, **k))

    def argmin(self, *a, **k):
        return wrap(self.view(np.ndarray).argmin(*a, **k))

    def block_until_ready(self):
        return self
