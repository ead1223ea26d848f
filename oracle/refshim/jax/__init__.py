"""numpy-backed stand-in for the subset of `jax` the reference's model files call.  TEST INFRASTRUCTURE ONLY
(see oracle/refshim/README.md): lets /root/reference/mreserve/modeling.py etc. be imported and executed unmodified.
"""
import threading
import types

import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, nn, random, tree_util  # noqa: F401
from .tree_util import tree_map  # noqa: F401

__version__ = "0.0-refshim"


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def checkpoint(fn, **kw):
    return fn


remat = checkpoint


class _Device:
    platform = "cpu"
    id = 0

    def __repr__(self):
        return "CpuDevice(id=0)"


def local_devices():
    return [_Device()]


devices = local_devices


def device_count():
    return 1


local_device_count = device_count


def process_index():
    return 0


def process_count():
    return 1


def device_get(x):
    return x


def device_put(x, *a, **k):
    return x


def _tree_index(tree, i):
    return tree_map(lambda a: a[i], tree)


def _tree_stack(trees):
    return tree_util.tree_multimap(lambda *xs: _np.stack([_np.asarray(x) for x in xs], 0), *trees)


def vmap(fn, in_axes=0, out_axes=0):
    """Python loop over the leading axis of every argument, outputs stacked on axis 0."""
    def mapped(*args):
        n = len(tree_util.tree_leaves(args)[0])
        return _tree_stack([fn(*_tree_index(args, i)) for i in range(n)])
    return mapped


def pmap(fn, axis_name=None, **kw):
    """One thread per leading-axis entry; lax.all_gather / pmean rendezvous on a barrier (rank-major results)."""
    def mapped(*args):
        n = len(tree_util.tree_leaves(args)[0])
        ctx = lax._AxisCtx(axis_name, n)
        outs, errs = [None] * n, []

        def run(i):
            lax._tls.ctx, lax._tls.rank = ctx, i
            try:
                outs[i] = fn(*_tree_index(args, i))
            except BaseException as e:  # noqa: BLE001
                errs.append(e)
                ctx.barrier.abort()
            finally:
                lax._tls.ctx = None

        ts = [threading.Thread(target=run, args=(i,)) for i in range(n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
        return _tree_stack(outs)
    return mapped


def value_and_grad(*a, **k):
    raise NotImplementedError("refshim has no autodiff: gradients are pinned through the oracle's torch autograd")


grad = value_and_grad
