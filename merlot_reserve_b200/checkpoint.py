"""Reader / writer for the reference's checkpoint files (mirrors mreserve/checkpoint.py:29-119).

The reference saves a flax `TrainState` with `flax.training.checkpoints.save_checkpoint` (checkpoint.py:99): one
file `ckpt_<step>` holding the msgpack encoding of the state dict
`{'step': int, 'params': {...}, 'opt_state': ... | None}`.  flax (pinned 0.3.4, requirements.txt:24; not in the
tree and not installable here) encodes every ndarray leaf as msgpack ExtType 1 whose payload is the msgpack tuple
`(shape, dtype.name, raw C-order bytes)`, numpy scalars as ExtType 3 (same payload), python complex as ExtType 2,
and splits arrays above 2**30 bytes into `{'__msgpack_chunked_array__': True, 'shape': ..., 'chunks': {'0': ...}}`.
That published format is restated here with `msgpack` + numpy only; no flax, no jax.

Reference behaviour kept:
  * fp32 leaves are stored as fp16 (`_compress_state`, checkpoint.py:29-36) and widened back to fp32 on load
    (`_decompress_state`, :39-40); bf16 leaves (the Adam moments of pretrain/optimization.py:36-51) stay bf16;
  * `load_checkpoint(path, step=None)` accepts the file itself (what `PretrainedMerlotReserve.from_pretrained`
    passes, modeling.py:991) or a directory, in which case the highest `ckpt_<step>` wins (flax
    `restore_checkpoint` natural-sort rule); `use_bfloat16_weights` casts the params to bf16 (:117-118);
  * `save_checkpoint(state, path, keep=, overwrite=, no_optimizer=)` writes `ckpt_<step>` and prunes old files.

bf16 arrays are returned as `ml_dtypes.bfloat16` numpy arrays (the dtype jax itself uses).
"""
from __future__ import annotations

import os
import re
from typing import Any, Callable, Dict, Optional

import msgpack
import numpy as np

try:  # the numpy extension dtype jax uses for bfloat16
    import ml_dtypes
    _BF16 = np.dtype(ml_dtypes.bfloat16)
except Exception:  # pragma: no cover - ml_dtypes ships with this image
    ml_dtypes = None
    _BF16 = None

_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
MAX_CHUNK_SIZE = 2 ** 30
_CHUNK_KEY = "__msgpack_chunked_array__"


# ------------------------------------------------------------------------------------------------
# flax.serialization wire format
# ------------------------------------------------------------------------------------------------
def _dtype_from_name(name) -> np.dtype:
    if isinstance(name, bytes):
        name = name.decode()
    if name == "bfloat16":
        if _BF16 is None:
            raise TypeError("checkpoint holds bfloat16 leaves but ml_dtypes is not importable")
        return _BF16
    return np.dtype(name)


def _ndarray_to_bytes(arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)
    if arr.dtype.hasobject or arr.dtype.isalignedstruct:
        raise ValueError("object and structured dtypes are not supported")
    return msgpack.packb((list(arr.shape), arr.dtype.name, arr.tobytes("C")), use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
    return np.frombuffer(buf, dtype=_dtype_from_name(dtype_name)).reshape(tuple(shape))


def _ext_pack(x):
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, _ndarray_to_bytes(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(_EXT_COMPLEX, msgpack.packb((x.real, x.imag)))
    raise TypeError(f"cannot serialise leaf of type {type(x)}")


def _ext_unpack(code: int, data: bytes):
    if code == _EXT_NDARRAY:
        return _ndarray_from_bytes(data)
    if code == _EXT_NPSCALAR:
        return _ndarray_from_bytes(data)[()]
    if code == _EXT_COMPLEX:
        re_, im_ = msgpack.unpackb(data)
        return complex(re_, im_)
    return msgpack.ExtType(code, data)


def _chunk(arr: np.ndarray, max_bytes: int) -> Dict:
    per = max(1, int(max_bytes // arr.dtype.itemsize))
    flat = arr.reshape(-1)
    # flax.serialization._chunk stores the shape through _tuple_to_dict: {'0': d0, '1': d1, ...}
    return {_CHUNK_KEY: True, "shape": {str(i): int(s) for i, s in enumerate(arr.shape)},
            "chunks": {str(i): flat[o:o + per] for i, o in enumerate(range(0, flat.size, per))}}


def _chunk_leaves(tree, max_bytes: int):
    if isinstance(tree, dict):
        return {k: _chunk_leaves(v, max_bytes) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):  # flax state dicts index sequences by str(position)
        return {str(i): _chunk_leaves(v, max_bytes) for i, v in enumerate(tree)}
    if isinstance(tree, np.ndarray) and tree.size * tree.dtype.itemsize > max_bytes:
        return _chunk(tree, max_bytes)
    return tree


def _unchunk_leaves(tree):
    if isinstance(tree, dict):
        if tree.get(_CHUNK_KEY):
            n = len(tree["chunks"])
            flat = np.concatenate([np.asarray(tree["chunks"][str(i)]).reshape(-1) for i in range(n)])
            shape = tree["shape"]
            if isinstance(shape, dict):          # flax's layout; a plain list (files written by round-1 code) also reads
                shape = [shape[str(i)] for i in range(len(shape))]
            return flat.reshape(tuple(int(x) for x in shape))
        return {k: _unchunk_leaves(v) for k, v in tree.items()}
    return tree


def msgpack_serialize(tree, max_chunk_bytes: int = MAX_CHUNK_SIZE) -> bytes:
    """flax.serialization.msgpack_serialize of a tree of dicts / numpy arrays / python scalars / None."""
    return msgpack.packb(_chunk_leaves(tree, max_chunk_bytes), default=_ext_pack, strict_types=True)


def msgpack_restore(encoded: bytes):
    """flax.serialization.msgpack_restore: bytes -> nested dict with numpy leaves."""
    return _unchunk_leaves(msgpack.unpackb(encoded, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


# ------------------------------------------------------------------------------------------------
# mreserve/checkpoint.py surface
# ------------------------------------------------------------------------------------------------
def tree_map(f: Callable, tree):
    if isinstance(tree, dict):
        return {k: tree_map(f, v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v) for v in tree)
    return f(tree)


def _treemap_cast(from_dtype, to_dtype, tree):
    """checkpoint.py:10-26: leaves of `from_dtype` become `to_dtype`; ints, None and other dtypes are untouched."""
    from_dtype = np.dtype(from_dtype)

    def cast(x):
        if not hasattr(x, "dtype"):
            return x
        return x.astype(to_dtype) if x.dtype == from_dtype else x

    return tree_map(cast, tree)


def _compress_state(state):
    return _treemap_cast(np.float32, np.float16, state)      # checkpoint.py:29-36


def _decompress_state(state):
    return _treemap_cast(np.float16, np.float32, state)      # checkpoint.py:39-40


def bf16_to_f32(params):
    return _treemap_cast(_BF16, np.float32, params)          # checkpoint.py:43-49


def f32_to_bf16(params):
    return _treemap_cast(np.float32, _BF16, params)          # checkpoint.py:52-58


_CKPT_RE = re.compile(r"^ckpt_(\d+)$")


def _resolve(path: str, step: Optional[int]) -> Optional[str]:
    if os.path.isfile(path):
        return path
    if not os.path.isdir(path):
        return None
    if step is not None:
        cand = os.path.join(path, f"ckpt_{step}")
        if not os.path.isfile(cand):
            raise ValueError(f"Matching checkpoint not found: {cand}")
        return cand
    steps = sorted(int(m.group(1)) for m in (_CKPT_RE.match(n) for n in os.listdir(path)) if m)
    return os.path.join(path, f"ckpt_{steps[-1]}") if steps else None


def _restore_into(target, loaded, where="state"):
    """flax `from_state_dict` for plain trees: the loaded tree must have exactly the target's keys."""
    if isinstance(target, dict):
        if not isinstance(loaded, dict) or set(map(str, target)) != set(map(str, loaded)):
            raise ValueError(f"checkpoint does not match the target structure at '{where}'")
        return {k: _restore_into(v, loaded[k if k in loaded else str(k)], f"{where}/{k}") for k, v in target.items()}
    return loaded


def load_checkpoint(path: str, state: Optional[Dict] = None, step: Optional[int] = None,
                    use_bfloat16_weights: bool = False) -> Dict:
    """checkpoint.py:102-119.  Returns the state dict ('step', 'params', 'opt_state') with fp32 parameters.  When
    nothing is found at `path` the (decompressed) `state` argument comes back unchanged, as flax does."""
    fn = _resolve(path, step)
    if fn is None:
        if state is None:
            raise FileNotFoundError(f"no checkpoint at {path}")
        loaded = _compress_state(state)
    else:
        with open(fn, "rb") as f:
            loaded = msgpack_restore(f.read())
        if state is not None:
            loaded = _restore_into(_compress_state(state), loaded)
    loaded = _decompress_state(loaded)
    if use_bfloat16_weights:
        loaded = dict(loaded, params=f32_to_bf16(loaded["params"]))
    return loaded


def save_checkpoint(state: Dict, path: str, keep: Optional[int] = None, overwrite: bool = True,
                    no_optimizer: bool = False, max_chunk_bytes: int = MAX_CHUNK_SIZE) -> str:
    """checkpoint.py:61-99 for a host-side state dict {'step', 'params', 'opt_state'} (the replica axis of the
    reference's pmap state does not exist here: one process per GPU holds unreplicated parameters)."""
    step = int(np.asarray(state["step"]).reshape(-1)[0])
    keep = 100000000 if keep is None else keep
    if no_optimizer:
        state = dict(state, opt_state=None)
    state = _compress_state(dict(state, step=step))
    os.makedirs(path, exist_ok=True)
    fn = os.path.join(path, f"ckpt_{step}")
    existing = sorted(int(m.group(1)) for m in (_CKPT_RE.match(n) for n in os.listdir(path)) if m)
    if not overwrite and existing and existing[-1] >= step:
        raise ValueError(f"Trying to save an outdated checkpoint at step: {step} (newest on disk: {existing[-1]})")
    tmp = fn + ".tmp"
    with open(tmp, "wb") as f:
        f.write(msgpack_serialize(state, max_chunk_bytes))
    os.replace(tmp, fn)
    if overwrite:                                            # newer files than the one just written are stale
        for s in existing:
            if s > step:
                os.remove(os.path.join(path, f"ckpt_{s}"))
    kept = sorted(int(m.group(1)) for m in (_CKPT_RE.match(n) for n in os.listdir(path)) if m)
    for s in kept[:-keep] if keep < len(kept) else []:
        os.remove(os.path.join(path, f"ckpt_{s}"))
    return fn


def flatten_dict(tree: Dict, prefix: str = "") -> Dict[str, Any]:
    """clu.parameter_overview.flatten_dict: nested keys joined by '/' (used by tree_map_nested_keys, :143-155)."""
    out = {}
    for k, v in tree.items():
        key = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten_dict(v, key))
        else:
            out[key] = v
    return out


def tree_map_nested_keys(f: Callable[[str, Any], Any], params: Dict) -> Dict:
    """checkpoint.py:143-155: tree map that also hands `f` the '/'-joined key of the leaf."""
    def rec(node, prefix):
        if isinstance(node, dict):
            return {k: rec(v, f"{prefix}/{k}" if prefix else str(k)) for k, v in node.items()}
        return f(prefix, node)
    return rec(params, "")


def log_param_shapes(params: Dict) -> int:
    """checkpoint.py:122-140: prints a parameter overview and returns the number of parameters."""
    flat = flatten_dict(params)
    width = max((len(k) for k in flat), default=4)
    for k in sorted(flat):
        v = np.asarray(flat[k])
        print(f"{k:<{width}}  {str(tuple(v.shape)):<20} {v.size:>12,d}")
    total = int(sum(np.asarray(v).size for v in flat.values()))
    print(f"Total: {total:,d}")
    return total
