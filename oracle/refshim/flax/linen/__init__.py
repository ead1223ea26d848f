"""`flax.linen` stand-in for executing the reference's model files without flax (TEST INFRASTRUCTURE, see
oracle/refshim/README.md).

The Module system (dataclass fields, `setup` vs `@compact`, explicit / attribute / `ClassName_i` auto-naming, scope =
nested dict of parameters, `init` / `apply(..., method=)`) follows flax 0.3.4's documented behaviour; the layer
bodies restate flax 0.3.4 `linen/linear.py`, `linen/normalization.py`, `linen/attention.py` from library knowledge
(flax itself is absent from /root/reference and from this image).
"""
import dataclasses
import functools
import threading
from typing import Any, Callable, Optional, Sequence, Tuple, Union

import numpy as np

import jax
import jax.numpy as jnp
from jax.nn import initializers  # noqa: F401  (nn.initializers.*)
from jax.nn import sigmoid, softmax, gelu  # noqa: F401

from . import attention  # noqa: F401,E402  (nn.attention.dot_product_attention_weights)
from .attention import dot_product_attention, dot_product_attention_weights  # noqa: F401,E402

_tls = threading.local()


def _stack():
    if not hasattr(_tls, "stack"):
        _tls.stack = []
    return _tls.stack


def _current():
    s = _stack()
    return s[-1] if s else None


def compact(fn):
    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        self._require_bound()
        object.__setattr__(self, "_autonames", {})
        _stack().append((self, False))
        try:
            return fn(self, *args, **kwargs)
        finally:
            _stack().pop()
    wrapped._is_compact = True
    return wrapped


def remat(cls, **kw):
    return cls


class Module:
    """Fields are the subclass's annotations (+ `parent`, `name`); a module is *bound* once it has a scope (a dict of
    its own parameters and its children's scopes)."""
    name: Optional[str] = None
    parent: Any = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann.pop("parent", None)
        ann.pop("name", None)
        ann["parent"] = Any
        ann["name"] = Optional[str]
        cls.__annotations__ = ann
        cls.parent = None
        cls.name = None
        dataclasses.dataclass(cls, eq=False, repr=False)

    # -- construction / binding ---------------------------------------------------------------------------------
    def __post_init__(self):
        d = self.__dict__
        d.update(_scope=None, _mode=None, _rng=None, _autonames={}, _setup_done=False, _path=())
        cur = _current()
        if cur is None or d.get("parent") is not None:
            return
        owner, in_setup = cur
        if in_setup and self.name is None:
            return                                       # named by the attribute it is assigned to (__setattr__)
        if self.name is None:
            prefix = type(self).__name__
            i = owner._autonames.get(prefix, 0)
            owner._autonames[prefix] = i + 1
            object.__setattr__(self, "name", f"{prefix}_{i}")
        self._bind(owner, self.name)

    def _bind(self, owner, name):
        d = self.__dict__
        d["parent"], d["_mode"], d["_rng"] = owner, owner._mode, owner._rng
        d["_path"] = owner._path + (name,)
        if owner._mode == "init":
            d["_scope"] = owner._scope.setdefault(name, {})
        else:
            d["_scope"] = owner._scope.get(name, {})
        self._run_setup()

    def _require_bound(self):
        if self.__dict__.get("_scope") is None:
            raise RuntimeError(f"{type(self).__name__} is not bound: call it through .apply() / .init()")

    def _run_setup(self):
        if self._setup_done:
            return
        object.__setattr__(self, "_setup_done", True)
        _stack().append((self, True))
        try:
            self.setup()
        finally:
            _stack().pop()

    def setup(self):
        pass

    def __setattr__(self, key, val):
        if isinstance(val, Module) and val.__dict__.get("_scope") is None and self.__dict__.get("_scope") is not None:
            if val.name is None:
                object.__setattr__(val, "name", key)
            val._bind(self, val.name)
        object.__setattr__(self, key, val)

    # -- parameters ---------------------------------------------------------------------------------------------
    def param(self, name, init_fn, *init_args):
        self._require_bound()
        if name in self._scope:
            return jnp.asarray(self._scope[name])
        if self._mode != "init":
            raise KeyError("parameter '" + "/".join(self._path + (name,)) + "' is missing from the supplied tree")
        key = self._rng
        for part in self._path + (name,):
            key = jax.random.fold_in(key, sum(part.encode()) * 131 + len(part))
        val = jnp.asarray(init_fn(key, *init_args))
        self._scope[name] = val
        return val

    def make_rng(self, name):
        return jax.random.fold_in(self._rng if self._rng is not None else jax.random.PRNGKey(0), hash(name) & 0xFFFF)

    # -- top-level entry points -----------------------------------------------------------------------------------
    def clone(self, **updates):
        fields = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.init}
        fields.update(parent=None)
        fields.update(updates)
        return type(self)(**fields)

    def _run(self, scope, mode, rng, args, kwargs, method):
        saved = list(_stack())
        _stack().clear()                                  # a top-level module has no enclosing context
        try:
            top = self.clone()
            top.__dict__.update(_scope=scope, _mode=mode, _rng=rng, _path=())
            top._run_setup()
            if method is None:
                fn = type(top).__call__
            else:
                fn = getattr(method, "__func__", method)
            return fn(top, *args, **kwargs)
        finally:
            _stack().clear()
            _stack().extend(saved)

    def apply(self, variables, *args, rngs=None, method=None, mutable=False, **kwargs):
        rng = None if not rngs else next(iter(rngs.values()))
        return self._run(variables["params"], "apply", rng, args, kwargs, method)

    def init(self, rngs, *args, method=None, **kwargs):
        rng = rngs if not isinstance(rngs, dict) else next(iter(rngs.values()))
        scope = {}
        self._run(scope, "init", rng, args, kwargs, method)
        return {"params": scope}

    def init_with_output(self, rngs, *args, method=None, **kwargs):
        rng = rngs if not isinstance(rngs, dict) else next(iter(rngs.values()))
        scope = {}
        out = self._run(scope, "init", rng, args, kwargs, method)
        return out, {"params": scope}


# ---------------------------------------------------------------------------------------------------------------
# Layers (flax 0.3.4 bodies restated; dtype handling: cast inputs and parameters to `dtype`, contract with fp32
# accumulation, add the bias in `dtype`)
# ---------------------------------------------------------------------------------------------------------------

default_kernel_init = initializers.lecun_normal()


def _tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x,)


class Dense(Module):
    """flax/linen/linear.py `Dense`: y = asarray(x, dtype) . asarray(kernel, dtype) (+ asarray(bias, dtype))."""
    features: int = None
    use_bias: bool = True
    dtype: Any = jnp.float32
    precision: Any = None
    kernel_init: Callable = default_kernel_init
    bias_init: Callable = initializers.zeros

    @compact
    def __call__(self, inputs):
        inputs = jnp.asarray(inputs, self.dtype)
        kernel = jnp.asarray(self.param("kernel", self.kernel_init, (inputs.shape[-1], self.features)), self.dtype)
        y = jnp.matmul(inputs, kernel)
        if self.use_bias:
            y = y + jnp.asarray(self.param("bias", self.bias_init, (self.features,)), self.dtype)
        return y


class DenseGeneral(Module):
    """flax/linen/linear.py `DenseGeneral` without batch_dims: kernel [*in_shape, *features], bias [*features];
    the initializer sees the flattened 2-D shape (kernel_init_wrap)."""
    features: Union[int, Sequence[int]] = None
    axis: Union[int, Sequence[int]] = -1
    batch_dims: Sequence[int] = ()
    use_bias: bool = True
    dtype: Any = jnp.float32
    kernel_init: Callable = default_kernel_init
    bias_init: Callable = initializers.zeros
    precision: Any = None

    @compact
    def __call__(self, inputs):
        assert not self.batch_dims
        features, axis = _tuple(self.features), _tuple(self.axis)
        inputs = jnp.asarray(inputs, self.dtype)
        nd = inputs.ndim
        axis = tuple(a % nd for a in axis)
        assert axis == tuple(range(nd - len(axis), nd)), "refshim: contraction axes must be the trailing ones"
        in_shape = tuple(inputs.shape[a] for a in axis)
        n_in, n_out = int(np.prod(in_shape)), int(np.prod(features))

        def init_wrap(key, shape, dtype=jnp.float32):
            return np.reshape(self.kernel_init(key, (n_in, n_out), dtype), shape)

        kernel = jnp.asarray(self.param("kernel", init_wrap, in_shape + features), self.dtype)
        lead = inputs.shape[:nd - len(axis)]
        out = jnp.matmul(inputs.reshape(lead + (n_in,)), kernel.reshape(n_in, n_out)).reshape(lead + features)
        if self.use_bias:
            out = out + jnp.asarray(self.param("bias", self.bias_init, features), self.dtype)
        return out


class LayerNorm(Module):
    """flax/linen/normalization.py `LayerNorm` (0.3.4): statistics in fp32 with var = E[x^2] - E[x]^2; scale and
    bias cast to `dtype` before use; output cast to `dtype`."""
    epsilon: float = 1e-6
    dtype: Any = jnp.float32
    use_bias: bool = True
    use_scale: bool = True
    bias_init: Callable = initializers.zeros
    scale_init: Callable = initializers.ones

    @compact
    def __call__(self, x):
        x = jnp.asarray(x, jnp.float32)
        features = x.shape[-1]
        mean = jnp.mean(x, axis=-1, keepdims=True)
        mean2 = jnp.mean(jax.lax.square(x), axis=-1, keepdims=True)
        var = mean2 - jax.lax.square(mean)
        mul = jax.lax.rsqrt(var + np.float32(self.epsilon))
        if self.use_scale:
            mul = mul * jnp.asarray(self.param("scale", self.scale_init, (features,)), self.dtype)
        y = (x - mean) * mul
        if self.use_bias:
            y = y + jnp.asarray(self.param("bias", self.bias_init, (features,)), self.dtype)
        return jnp.asarray(y, self.dtype)


class Conv(Module):
    """flax/linen/linear.py `Conv`, 1-D, padding SAME.  Only the non-overlapping case the reference uses
    (stride == kernel size on a length that divides evenly, so SAME adds no padding) is implemented: it is the
    reshape-GEMM [B, L/k, k*C] x [k*C, F]."""
    features: int = None
    kernel_size: Sequence[int] = None
    strides: Optional[Sequence[int]] = None
    padding: Any = "SAME"
    use_bias: bool = True
    dtype: Any = jnp.float32
    kernel_init: Callable = default_kernel_init
    bias_init: Callable = initializers.zeros
    precision: Any = None

    @compact
    def __call__(self, inputs):
        inputs = jnp.asarray(inputs, self.dtype)
        (k,), (s,) = tuple(self.kernel_size), tuple(self.strides or (1,))
        B, L, C = inputs.shape
        assert k == s and L % k == 0, "refshim Conv: only stride == kernel_size on an evenly divisible length"
        kernel = jnp.asarray(self.param("kernel", self.kernel_init, (k, C, self.features)), self.dtype)
        y = jnp.matmul(inputs.reshape(B, L // k, k * C), kernel.reshape(k * C, self.features))
        if self.use_bias:
            y = y + jnp.asarray(self.param("bias", self.bias_init, (self.features,)), self.dtype)
        return y


default_embed_init = initializers.variance_scaling(1.0, "fan_in", "normal")


class Embed(Module):
    """flax/linen/linear.py `Embed`: the table is a setup() parameter named 'embedding'; lookup by integer index."""
    num_embeddings: int = None
    features: int = None
    dtype: Any = jnp.float32
    embedding_init: Callable = default_embed_init

    def setup(self):
        self.embedding = self.param("embedding", self.embedding_init, (self.num_embeddings, self.features))

    def __call__(self, inputs):
        self._require_bound()
        inputs = jnp.asarray(inputs)
        assert inputs.dtype.kind in "iu", "Input type must be an integer or unsigned integer."
        return jnp.asarray(self.embedding)[inputs]


class MultiHeadDotProductAttention(Module):
    """flax/linen/attention.py `MultiHeadDotProductAttention` (no dropout, no cache): query/key/value
    DenseGeneral(features=(num_heads, head_dim)) with bias, dot_product_attention, `out` DenseGeneral over
    (heads, head_dim) with bias."""
    num_heads: int = None
    dtype: Any = jnp.float32
    qkv_features: Optional[int] = None
    out_features: Optional[int] = None
    broadcast_dropout: bool = True
    dropout_rate: float = 0.0
    deterministic: Optional[bool] = None
    precision: Any = None
    kernel_init: Callable = default_kernel_init
    bias_init: Callable = initializers.zeros
    use_bias: bool = True
    decode: bool = False

    @compact
    def __call__(self, inputs_q, inputs_kv, mask=None, deterministic=None):
        assert self.dropout_rate == 0.0 and not self.decode
        features = self.out_features or inputs_q.shape[-1]
        qkv_features = self.qkv_features or inputs_q.shape[-1]
        assert qkv_features % self.num_heads == 0
        head_dim = qkv_features // self.num_heads
        dense = functools.partial(DenseGeneral, axis=-1, features=(self.num_heads, head_dim), dtype=self.dtype,
                                  kernel_init=self.kernel_init, bias_init=self.bias_init, use_bias=self.use_bias)
        query = dense(name="query")(inputs_q)
        key = dense(name="key")(inputs_kv)
        value = dense(name="value")(inputs_kv)
        bias = None
        if mask is not None:
            bias = jax.lax.select(mask > 0, jnp.full(mask.shape, 0.).astype(self.dtype),
                                  jnp.full(mask.shape, -1e10).astype(self.dtype))
        x = dot_product_attention(query, key, value, bias=bias, dtype=self.dtype)
        return DenseGeneral(features=features, axis=(-2, -1), kernel_init=self.kernel_init, bias_init=self.bias_init,
                            use_bias=self.use_bias, dtype=self.dtype, name="out")(x)


SelfAttention = MultiHeadDotProductAttention
