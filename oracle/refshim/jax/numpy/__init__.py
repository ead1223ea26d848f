"""`jax.numpy` stand-in: numpy with the same function names / axis semantics (TEST INFRASTRUCTURE, see ../../README.md).

Differences from plain numpy that matter to the reference files are handled here:
* `bfloat16` (ml_dtypes) exists and contractions on it accumulate in fp32 and round once, as XLA does;
* `einsum` / `matmul` / `dot` run on torch-CPU (multi-threaded) so Base-size clips finish in seconds;
* float `arange` honours sub-fp32 dtypes;
* every result is a `jax._array.Array` (jax's "x64 disabled" promotion: no float64 / int64 results, immutable `+=`).
Everything else falls through to numpy (`__getattr__`), wrapped the same way.
"""
import functools as _functools

import numpy as _np

from .._array import Array, unwrap as _unwrap, wrap as _wrap

try:
    import ml_dtypes as _ml
    bfloat16 = _ml.bfloat16
except Exception:  # pragma: no cover
    bfloat16 = None

ndarray = _np.ndarray


def _jaxlike(fn):
    @_functools.wraps(fn)
    def f(*a, **k):
        return _wrap(fn(*_unwrap(a), **_unwrap(k)))
    return f


float32, float16, float64 = _np.float32, _np.float16, _np.float64
int32, int64, uint32, uint8, int8 = _np.int32, _np.int64, _np.uint32, _np.uint8, _np.int8
bool_ = _np.bool_
pi, inf, newaxis = _np.pi, _np.inf, None


def _is_bf16(*xs):
    return bfloat16 is not None and any(getattr(x, "dtype", None) == bfloat16 for x in xs)


def _torch_contract(fn, *ops):
    import torch
    ops = [_np.asarray(o) for o in ops]
    out_dt = _np.result_type(*[o.dtype for o in ops]) if not _is_bf16(*ops) else bfloat16
    ts = [torch.from_numpy(_np.ascontiguousarray(o.astype(_np.float32) if o.dtype != _np.float64 else o)) for o in ops]
    if any(t.dtype == torch.float64 for t in ts):
        ts = [t.double() for t in ts]
    return fn(*ts).numpy().astype(out_dt)


@_jaxlike
def einsum(subscripts, *operands, **kw):
    import torch
    return _torch_contract(lambda *ts: torch.einsum(subscripts, *ts), *operands)


@_jaxlike
def matmul(a, b, **kw):
    import torch
    return _torch_contract(torch.matmul, a, b)


@_jaxlike
def dot(a, b, **kw):
    a, b = _np.asarray(a), _np.asarray(b)
    if a.ndim <= 2 and b.ndim <= 2:
        return matmul(a, b)
    return _np.dot(a, b)


@_jaxlike
def asarray(x, dtype=None):
    return _np.asarray(x, dtype=dtype)


@_jaxlike
def array(x, dtype=None, **kw):
    a = _np.array(x, dtype=dtype)
    if dtype is None and a.dtype == _np.int64:
        a = a.astype(_np.int32)
    if dtype is None and a.dtype == _np.float64:
        a = a.astype(_np.float32)
    return a


@_jaxlike
def arange(*args, dtype=None):
    if dtype is not None and _np.dtype(dtype).kind not in "iub" and _np.dtype(dtype).itemsize < 4:
        return _np.arange(*args, dtype=_np.float32).astype(dtype)
    a = _np.arange(*args, dtype=dtype)
    return a.astype(_np.int32) if dtype is None and a.dtype == _np.int64 else a


@_jaxlike
def zeros(shape, dtype=float32):
    return _np.zeros(shape, dtype)


@_jaxlike
def ones(shape, dtype=float32):
    return _np.ones(shape, dtype)


@_jaxlike
def full(shape, fill_value, dtype=None):
    if dtype is None:
        dtype = _np.float32 if isinstance(fill_value, float) else (_np.int32 if isinstance(fill_value, int) else None)
    return _np.full(shape, fill_value, dtype)


@_jaxlike
def logspace(start, stop, num=50, endpoint=True, base=10.0, dtype=None, axis=0):
    """jax: linspace in `dtype`, then power(base, .) in `dtype`."""
    dt = _np.float32 if dtype is None else dtype
    lin = _np.linspace(start, stop, num, endpoint=endpoint).astype(dt)
    return _np.power(_np.asarray(base, dt), lin).astype(dt)


@_jaxlike
def cumsum(x, axis=None, dtype=None):
    x = _np.asarray(x)
    out = _np.cumsum(x, axis=axis, dtype=dtype)
    return out.astype(x.dtype) if dtype is None and x.dtype.kind in "iu" else out


@_jaxlike
def sum(x, axis=None, dtype=None, keepdims=False):  # noqa: A001
    x = _np.asarray(x)
    out = _np.sum(x, axis=axis, dtype=dtype, keepdims=keepdims)
    return out.astype(x.dtype) if dtype is None and x.dtype.kind in "iu" else out


@_jaxlike
def argsort(x, axis=-1, **kw):
    """jax.numpy.argsort is stable."""
    return _np.argsort(_np.asarray(x), axis=axis, kind="stable").astype(_np.int32)


@_jaxlike
def nan_to_num(x, nan=0.0, posinf=None, neginf=None, **kw):
    return _np.nan_to_num(x, nan=nan, posinf=posinf, neginf=neginf)


@_jaxlike
def clip(a, a_min=None, a_max=None, **kw):
    a = _np.asarray(a)
    out = _np.clip(a, a_min, a_max)
    return out.astype(a.dtype)


def __getattr__(name):
    obj = getattr(_np, name)
    if callable(obj) and not isinstance(obj, type):
        return _jaxlike(obj)
    return obj
