"""jax.nn.initializers subset (numpy generators keyed by jax.random keys; distributions as published, draws are
NOT threefry)."""
import numpy as np

from .. import random as _random


def zeros(key, shape, dtype=np.float32):
    return np.zeros(shape, dtype)


def ones(key, shape, dtype=np.float32):
    return np.ones(shape, dtype)


def normal(stddev=1e-2, dtype=np.float32):
    def init(key, shape, dtype=dtype):
        return (_random.normal(key, shape, np.float32) * stddev).astype(dtype)
    return init


def _fans(shape):
    receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    return shape[-2] * receptive, shape[-1] * receptive


def variance_scaling(scale, mode, distribution, dtype=np.float32):
    def init(key, shape, dtype=dtype):
        fan_in, fan_out = _fans(shape)
        denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
        var = scale / denom
        if distribution == "truncated_normal":
            return (_random.truncated_normal(key, -2, 2, shape, np.float32) * np.sqrt(var) / .87962566103423978).astype(dtype)
        if distribution == "normal":
            return (_random.normal(key, shape, np.float32) * np.sqrt(var)).astype(dtype)
        return (_random.uniform(key, shape, np.float32, -1.0, 1.0) * np.sqrt(3 * var)).astype(dtype)
    return init


def lecun_normal(dtype=np.float32):
    return variance_scaling(1.0, "fan_in", "truncated_normal", dtype)


def xavier_uniform(dtype=np.float32):
    return variance_scaling(1.0, "fan_avg", "uniform", dtype)


glorot_uniform = xavier_uniform
