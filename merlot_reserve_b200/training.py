"""Forward + backward of the contrastive pretraining step on the B200 kernels (pretrain/pretrain_model.py:306-340).

The reference gets its gradients from `jax.value_and_grad` over `MerlotReservePretrainer.__call__` +
`loss_fn_given_preds`; here the same chain rule is written out over the C-ABI kernels: every GEMM gradient is
`mr_gemm_epilogue` with transposed (MN-major) operands and split-K, attention uses `mr_attention_bwd`, the row ops
have their `_bwd` entries, and the gather / pool / assembly steps scatter their gradients back with the `_bwd`
kernels of include/mreserve_b200.h.  torch only owns memory, prepares integer indices, and runs the collectives.

`PretrainStep.loss_and_grads(batch)` returns the loss, the loss_info dict and the parameter gradients as a flat
dict keyed by the reference's flax parameter paths ('vision_encoder/transformer/layer_00/mlp_layer/out/kernel', ...),
each in the flax layout and fp32, so that they compare one to one with `jax.grad` (tests use torch autograd over the
CPU oracle).  Training runs the un-folded LayerNorm path (the folded weights of the inference path are functions of
gamma / beta and would have to be re-packed every step).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from .modeling import LTOVPOOL, MASK, MASKAUDIO, PADDING, _AUDIO_K_PAD, _PackedEncoder, _PackedPool, _dev
from .pretrain_model import MerlotReservePretrainer, gather_contrastive_targets

BF16 = torch.bfloat16


def _np_dtype(dt):
    if dt == BF16:
        import ml_dtypes
        return np.dtype(ml_dtypes.bfloat16)
    return np.dtype(np.float32)


def _nu_decode(code: np.ndarray) -> np.ndarray:
    """Host twin of the device decoder (pretrain/optimization.py:36-39), for the three contrastive temperatures."""
    a = np.abs(code.astype(np.float32))
    return np.cbrt(np.where(code.astype(np.float32) >= 0, a, a * np.float32(1 + 1 / 512))).astype(np.float32)


def _nu_encode(v: np.ndarray) -> np.ndarray:
    """pretrain/optimization.py:42-50."""
    bf = _np_dtype(BF16)
    v_pow = (v * v * v).astype(np.float32)
    c = v_pow.astype(bf).astype(np.float32)
    return np.where(np.abs(c - v_pow) < np.abs(c * np.float32(1 + 1 / 512) - v_pow), c, -c).astype(bf)


def _split_k(rows: int) -> int:
    """The round-1 rule (kept for A/B, MR_WGRAD_PLAN=legacy): enough K parts to occupy the SMs, ~2k rows per part."""
    return int(max(1, min(16, rows // 2048)))


_LEGACY_WGRAD = os.environ.get("MR_WGRAD_PLAN", "") == "legacy"
_FUSED_DGELU = os.environ.get("MR_FUSED_DGELU", "1") != "0"
_OVERLAP_ALLREDUCE = os.environ.get("MR_OVERLAP_ALLREDUCE", "1") != "0"
# Timing probe of bench.py only: run the step WITHOUT the gradient all-reduce (the result is then not a training
# step); the difference to the full step is the all-reduce time that the backward pass does not hide.
_SKIP_GRAD_ALLREDUCE = False
# The collectives of the step go through the library's own entries (mr_all_gather / mr_reduce_scatter / mr_all_reduce on
# an ncclComm_t made by mr_nccl_comm_init) when the group's backend is NCCL; MR_NATIVE_COLLECTIVES=0 keeps
# torch.distributed's wrappers around the same NCCL calls (and is what a gloo / CPU group uses).
_NATIVE_COLLECTIVES = os.environ.get("MR_NATIVE_COLLECTIVES", "1") != "0"
_COMMUNICATORS: Dict[int, tuple] = {}
# MR_ATTN_BWD_FINISH=0: the three-pass tail (mr_attention_bwd, mr_rotary_bwd, mr_colsum) for same-box A/B runs
_FUSED_QKV_TAIL = os.environ.get("MR_ATTN_BWD_FINISH", "1") != "0"


def _wgrad(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW in the packed [out, in] layout = dy^T [out, rows] . x [rows, in]; tile width and split-K from the
    wave model of `_lib.wgrad_plan` (tiles x splits a multiple of the SM count, few fp32 partials)."""
    if _LEGACY_WGRAD:
        return L.gemm_bwd(dy, x, trans_a=True, trans_b=True, split_k=_split_k(dy.shape[0]))
    tile_n, split_k = L.wgrad_plan(dy.shape[1], x.shape[1], dy.shape[0])
    return L.gemm_bwd(dy, x, trans_a=True, trans_b=True, split_k=split_k, tile_n=tile_n)


def _dgrad(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dx bf16 [rows, in] = dy [rows, out] . W [out, in] (the forward's packed weight, consumed MN-major)."""
    return L.gemm_bwd(dy, w, trans_b=True, out_f32=False)


def arena_layout(names: List[str], numel: Dict[str, int]) -> Tuple[Dict[str, int], int]:
    """Element offsets of the flat optimizer / gradient arenas: tensors in sorted-name order, 8-element (16-byte bf16)
    alignment, and every top-level module ('audio_encoder', 'head', 'joint_transformer', ...) AND every transformer
    layer ('.../layer_07') starting on a 1024-element boundary, so that the buckets of the overlapped gradient
    all-reduce are whole rows of the bf16 cast (rows of 1024) and never share a row with a module or layer that is
    still differentiating.  -> (offsets, total)."""
    assert list(names) == sorted(names)
    offs, n, group = {}, 0, None
    for name in names:
        g = _bucket_group(name)
        if g != group:
            group = g
            n = (n + 1023) // 1024 * 1024
        offs[name] = n
        n += (numel[name] + 7) // 8 * 8
    return offs, (n + 1023) // 1024 * 1024


def _bucket_group(name: str) -> str:
    """Alignment group of a parameter: its transformer layer ('a/b/layer_07') if it has one, else its top-level module."""
    parts = name.split("/")
    for i, part in enumerate(parts):
        if part.startswith("layer_"):
            return "/".join(parts[:i + 1])
    return parts[0]


# The groups of top-level modules whose gradients become final together during `loss_and_grads`' backward pass, in
# that order; train_step starts one asynchronous all-reduce per group (`on_ready`).
ALLREDUCE_BUCKETS = (("head", "span_encoder", "joint_transformer", "token_encoder"), ("audio_encoder",), ("vision_encoder",))
# The vision encoder is differentiated last, so its bucket would only start when the backward pass ends.  Its layers
# are final one by one (11 -> 0): after the layers listed here an extra bucket with the finished upper layers starts,
# so that only the last few layers' gradients (plus embedding / pool) are reduced behind the backward pass.
VISION_LAYER_SPLITS = tuple(int(v) for v in os.environ.get("MR_VISION_BUCKET_SPLITS", "8,4").split(",") if v != "")


class _Grads:
    """Accumulates parameter gradients (fp32, packed layout) under flax-style path names."""

    def __init__(self):
        self.g: Dict[str, torch.Tensor] = {}

    def add(self, name: str, t: torch.Tensor):
        t = t.float() if t.dtype != torch.float32 else t
        if name in self.g:
            self.g[name] += t
        else:
            self.g[name] = t


class PretrainStep:
    def __init__(self, config: Dict, params: Dict):
        self.model = MerlotReservePretrainer.from_config(config)
        self.model.fuse_layernorm = False
        self.model.bind(params)
        self.params = params
        self.cfg = self.model.config
        self.H = self.model.hidden_size
        self._state = None

    # --------------------------------------------------------------------------------------------
    # train_step (pretrain_model.py:306-340): grads -> nan_to_num -> mean over the data-parallel ranks (bf16, as the
    # reference's pmean) -> AdamW (pretrain/optimization.py:158-195) on fp32 masters -> refreshed bf16 weights
    # --------------------------------------------------------------------------------------------
    def _registry(self) -> Dict[str, Tuple[torch.Tensor, bool]]:
        """flax parameter path -> (device tensor the kernels read, decays?).  bf16 tensors get an fp32 master."""
        p, H = self.model._p, self.H
        reg: Dict[str, Tuple[torch.Tensor, bool]] = {}
        for prefix, enc in (("vision_encoder/transformer", p.vis), ("audio_encoder/transformer", p.aud),
                            ("span_encoder/transformer", p.span), ("joint_transformer", p.joint)):
            reg[f"{prefix}/pre_ln/scale"], reg[f"{prefix}/pre_ln/bias"] = (enc.pre_ln[0], False), (enc.pre_ln[1], False)
            reg[f"{prefix}/final_ln/scale"], reg[f"{prefix}/final_ln/bias"] = (enc.final_ln[0], False), (enc.final_ln[1], False)
            if enc.add_cls:
                reg[f"{prefix}/cls"] = (enc.cls, False)
                reg[f"{prefix}/cls_proj/kernel"], reg[f"{prefix}/cls_proj/bias"] = (enc.cls_proj[0], True), (enc.cls_proj[1], False)
            for i, lw in enumerate(enc.layers):
                q = f"{prefix}/layer_{i:02d}"
                reg[f"{q}/pre_attn_ln/scale"], reg[f"{q}/pre_attn_ln/bias"] = (lw["ln1"][0], False), (lw["ln1"][1], False)
                reg[f"{q}/pre_mlp_ln/scale"], reg[f"{q}/pre_mlp_ln/bias"] = (lw["ln2"][0], False), (lw["ln2"][1], False)
                reg[f"{q}/attention_layer/qkv/kernel"], reg[f"{q}/attention_layer/qkv/bias"] = (lw["wqkv"], True), (lw["bqkv"], True)
                reg[f"{q}/attention_layer/attn_proj/kernel"] = (lw["wo"], True)
                reg[f"{q}/mlp_layer/intermediate/kernel"], reg[f"{q}/mlp_layer/intermediate/bias"] = (lw["w1"], True), (lw["b1"], False)
                reg[f"{q}/mlp_layer/out/kernel"] = (lw["w2"], True)
        for prefix, pool in (("vision_encoder/seq_attnpool", p.vis_pool), ("audio_encoder/seq_attnpool", p.aud_pool)):
            reg[f"{prefix}/query/kernel"], reg[f"{prefix}/query/bias"] = (pool.wq, True), (pool.bq, True)
            reg[f"{prefix}/key/kernel"], reg[f"{prefix}/value/kernel"] = (pool.wkv[:H], True), (pool.wkv[H:], True)
            reg[f"{prefix}/key/bias"], reg[f"{prefix}/value/bias"] = (pool.bkv[:H], True), (pool.bkv[H:], True)
            reg[f"{prefix}/out/kernel"], reg[f"{prefix}/out/bias"] = (pool.wo, True), (pool.bo, False)
        reg["vision_encoder/embedding/kernel"], reg["vision_encoder/embedding/bias"] = (p.vis_emb[0], True), (p.vis_emb[1], False)
        reg["audio_encoder/embedding/kernel"], reg["audio_encoder/embedding/bias"] = (p.aud_emb[0], True), (p.aud_emb[1], False)
        reg["token_encoder/Embed_0/embedding"] = (p.tok_emb, True)
        reg["head/kernel"], reg["head/bias"] = (p.head[0], True), (p.head[1], False)
        return reg

    # default optimizer block = pretrain/configs/base.yaml:81-90
    DEFAULT_OPT = {"learning_rate": 4e-4, "num_train_steps": 750000, "num_warmup_steps": 3750, "weight_decay_rate": 0.1,
                   "beta_2": 0.98, "eps": 1e-6, "use_bfloat16_adam": True}

    def _init_state(self, opt: Optional[Dict] = None):
        """Optimizer state as flat arenas (mr_adamw_multi): fp32 master parameters, the two Adam moments in the
        reference's storage (bf16 mu, sign-coded bf16 of nu^3; optimization.py:36-51,72) or fp32 when
        `use_bfloat16_adam` is off, and the chunk table that maps arena elements to the tensors the kernels read.
        Decay mask = flax ndim > 1 (optimization.py:182-184): the qkv / pool biases are [heads, 64] in flax."""
        opt = opt or self.DEFAULT_OPT
        reg = self._registry()
        names = sorted(reg)
        dev = next(iter(reg.values()))[0].device
        offs, n_total = arena_layout(names, {k: reg[k][0].numel() for k in names})
        mdt = BF16 if opt.get("use_bfloat16_adam", True) else torch.float32
        p = torch.zeros(n_total, dtype=torch.float32, device=dev)
        slots = []
        from .checkpoint import flatten_dict
        flat_src = flatten_dict(self.params)
        for name in names:
            t, decay = reg[name]
            assert t.is_contiguous(), name
            dst = p[offs[name]:offs[name] + t.numel()].view(t.shape)
            # fp32 masters come from the caller's fp32 tree (pretrain_model.py:321-323 keeps fp32 TrainState.params and
            # casts only a copy to bf16), not from the bf16 working copies the kernels read
            dst.copy_(self._from_flax(name, np.asarray(flat_src[name], np.float32), dst))
            slots.append((offs[name], t.numel(), t, decay))
        self._state = {"names": names, "offs": offs, "n": n_total, "p": p,
                       "m": torch.zeros(n_total, dtype=mdt, device=dev), "v": torch.zeros(n_total, dtype=mdt, device=dev),
                       "table": L.adamw_chunk_table(slots, dev),
                       # gradient arena (+ 8 trailing slots: the 3 contrastive temperatures ride in the same all-reduce)
                       "g": torch.zeros(n_total + 1024, dtype=torch.float32, device=dev),
                       "g16": torch.empty(n_total + 1024, dtype=BF16, device=dev)}
        self._scale_state = {"m": np.zeros(3, _np_dtype(mdt)), "v": np.zeros(3, _np_dtype(mdt))}
        self.step_count = 0

    def _step_scalars(self, opt: Dict):
        """(lr, bias correction on?, t) of the step about to be taken."""
        lr = opt["learning_rate"] * self.lr_scale(self.step_count, opt)
        # bias correction: off by default with the bf16 moments (optimization.py:165), always on for optax.scale_by_adam
        corr = opt.get("do_bias_correction", False) if opt.get("use_bfloat16_adam", True) else True
        return lr, corr, self.step_count + 1

    def _update_temperatures(self, g_scales: np.ndarray, opt: Dict, world: int) -> None:
        """Adam on the three contrastive temperatures.  They live on the host (they parameterise kernel launches);
        ndim 1: no weight decay (optimization.py:182-184)."""
        b1, b2, eps = opt.get("beta_1", 0.9), opt.get("beta_2", 0.98), opt.get("eps", 1e-8)
        lr, corr, t1 = self._step_scalars(opt)
        gs = np.nan_to_num(g_scales) / world
        ss, p = self._scale_state, self.model._p
        coded = ss["m"].dtype != np.float32
        f32 = np.float32
        mh = f32(1 - b1) * gs + f32(b1) * ss["m"].astype(f32)
        vh = f32(1 - b2) * gs * gs + f32(b2) * (_nu_decode(ss["v"]) if coded else ss["v"])
        ss["m"], ss["v"] = (mh.astype(ss["m"].dtype), _nu_encode(vh)) if coded else (mh, vh)
        if corr:
            mh, vh = mh / f32(1 - b1 ** t1), vh / f32(1 - b2 ** t1)
        p.contrastive_scales = (p.contrastive_scales - f32(lr) * mh / (np.sqrt(vh) + f32(eps))).astype(np.float32)
        self.step_count += 1

    # --------------------------------------------------------------------------------------------
    # CUDA-graph form of train_step: the ~1 700 kernel launches and the index glue of one step are captured once and
    # replayed with one graph launch.  Everything that changes from step to step enters through device memory:
    # the batch (static input buffers), the learning rate / bias corrections (mr_adamw_multi reads f32 [3]) and the
    # three temperatures (mr_unit_normalize(_bwd) read exp(min(s, log 100) / 2) from a device scalar).  The Adam
    # update of the three temperatures stays on the host, between replays.
    # --------------------------------------------------------------------------------------------
    def communicator(self, group=None):
        """The library-side NCCL communicator of `group` (created collectively on first use, one per group), or None
        when the collectives stay with torch.distributed (single rank, non-NCCL backend, MR_NATIVE_COLLECTIVES=0)."""
        import torch.distributed as dist
        if not (_NATIVE_COLLECTIVES and dist.is_available() and dist.is_initialized()):
            return None
        if dist.get_world_size(group) < 2 or dist.get_backend(group) != "nccl":
            return None
        # one communicator per process group for the whole process (ncclCommInitRank is collective and takes ~1 s), held
        # together with the group object so that its id cannot be reused while the entry lives
        key = id(group) if group is not None else 0
        if key not in _COMMUNICATORS:
            _COMMUNICATORS[key] = (L.Communicator.from_group(group), group)
        if self.__dict__.get("_comm_stream") is None:
            self._comm_stream = torch.cuda.Stream(priority=-1)
        return _COMMUNICATORS[key][0]

    def capture_train_step(self, batch: Dict, opt: Optional[Dict] = None, group=None):
        """Captures train_step for batches shaped like `batch` (which must carry the '_rng/*' draws: the numpy draw of
        the eager path cannot be captured).  Returns `run(batch) -> (loss, loss_info)` (0-d CUDA tensors, valid until
        the next call).  With more than one rank the NCCL collectives of the step are captured too."""
        import torch.distributed as dist
        opt = opt or self.DEFAULT_OPT
        if self._state is None:
            self._init_state(opt)
        for k in ("_rng/split_audio2text", "_rng/split_text2audio", "_rng/gumbel_u"):
            if k not in batch:
                raise ValueError(f"capture_train_step needs the random draws as inputs: '{k}' is missing from the batch")
        dev = next(iter(self._registry().values()))[0].device
        static = {k: _dev(v).clone() for k, v in batch.items()}
        self._graph_scalars = {"temp": torch.ones(3, dtype=torch.float32, device=dev),
                               "unclipped": torch.ones(3, dtype=torch.float32, device=dev),
                               "hyper": torch.ones(3, dtype=torch.float32, device=dev)}
        host = torch.empty(9, dtype=torch.float32).pin_memory()
        out_host = torch.empty(3, dtype=torch.float32).pin_memory()
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

        def set_scalars():
            ls_raw = self.model._p.contrastive_scales.astype(np.float32)
            b1, b2 = opt.get("beta_1", 0.9), opt.get("beta_2", 0.98)
            lr, corr, t1 = self._step_scalars(opt)
            host[0:3] = torch.from_numpy(np.exp(np.minimum(ls_raw, np.float32(math.log(100.0))) / 2.0).astype(np.float32))
            host[3:6] = torch.from_numpy((ls_raw <= math.log(100.0)).astype(np.float32))
            host[6], host[7], host[8] = lr, (1 - b1 ** t1) if corr else 1.0, (1 - b2 ** t1) if corr else 1.0
            gs = self._graph_scalars
            gs["temp"].copy_(host[0:3], non_blocking=True)
            gs["unclipped"].copy_(host[3:6], non_blocking=True)
            gs["hyper"].copy_(host[6:9], non_blocking=True)

        # warm-up on a side stream (lazy kernel attributes, NCCL communicator set-up, allocator pools), restoring the state
        snap = {k: self._state[k].clone() for k in ("p", "m", "v")}
        working = [(t, t.clone()) for t, _ in self._registry().values()]
        saved_scales = (self.model._p.contrastive_scales.copy(), {k: v.copy() for k, v in self._scale_state.items()},
                        self.step_count)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            set_scalars()
            self.train_step(static, opt, group=group)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for k, v in snap.items():
            self._state[k].copy_(v)
        for t, v in working:
            t.copy_(v)
        self.model._p.contrastive_scales, self._scale_state, self.step_count = saved_scales[0], saved_scales[1], saved_scales[2]
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss, info = self.train_step(static, opt, group=group)
        g_scales = self._graph_scalars["g_scales"]

        def run(new_batch: Optional[Dict] = None):
            if new_batch is not None and new_batch is not batch:
                for k, dst in static.items():
                    src = new_batch[k]
                    src = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(src))
                    dst.copy_(src.reshape(dst.shape).to(dst.dtype), non_blocking=True)
            set_scalars()
            graph.replay()
            out_host.copy_(g_scales.float(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self._update_temperatures(out_host.numpy().copy(), opt, world)
            return loss, info

        run.graph, run.static = graph, static
        return run

    def lr_scale(self, step: int, opt: Dict) -> float:
        """lr_scale_linearwarmup_cosinedecay (optimization.py:117-137)."""
        nw, nt, final = opt["num_warmup_steps"], opt["num_train_steps"], opt.get("final_lr_scale", 0.02)
        if step < nw:
            return step / nw
        s = min((step - nw) / (nt - nw + 1.0), 1.0)
        return final + (1.0 - final) * (1.0 - (1.0 - math.cos(math.pi * s)) / 2.0)

    def optimizer_state(self) -> Dict[str, Dict[str, torch.Tensor]]:
        """Views of the arenas per flax parameter path (packed layout): {'p': .., 'm': .., 'v': ..}."""
        st, reg = self._state, self._registry()
        out = {}
        for name in st["names"]:
            o, k, shp = st["offs"][name], reg[name][0].numel(), reg[name][0].shape
            out[name] = {key: st[key][o:o + k].view(shp) for key in ("p", "m", "v")}
        return out

    def train_step(self, batch: Dict, opt: Optional[Dict] = None, group=None):
        """One optimisation step; returns (loss, loss_info) of this rank (0-d CUDA tensors)."""
        import torch.distributed as dist
        opt = opt or self.DEFAULT_OPT
        if self._state is None:
            self._init_state(opt)
        st = self._state
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        n, g = st["n"], st["g"]
        # The reference differentiates bf16 parameters, so its gradients ARE bf16, and averages them in bf16
        # (pretrain_model.py:323-333): one flat bf16 arena that the optimizer kernel reads as is.  With more than one
        # rank the arena is reduced in three buckets AS THE BACKWARD PASS FINISHES THEM (the order of
        # `loss_and_grads`: heads + span + joint + token table, then the audio encoder, then the vision encoder), each
        # an asynchronous NCCL all-reduce that runs under the rest of the backward pass; MR_OVERLAP_ALLREDUCE=0 keeps
        # the single blocking all-reduce of round 1p.
        pending, done = [], set()
        self.last_bucket_bytes = []
        comm = self.communicator(group) if world > 1 else None

        def start_all_reduce(buf):
            """Asynchronous sum of one bf16 bucket: NCCL on the communication stream, behind everything the current
            stream has enqueued so far; joined before the optimizer kernel."""
            if comm is None:
                pending.append(dist.all_reduce(buf, group=group, async_op=True))
                return
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                comm.all_reduce(buf)
            pending.append(None)

        def flush(names, reduce_now):
            """Copy finished gradients into the arena, cast their 1024-element rows to bf16, start their all-reduce."""
            names = sorted((k for k in names if k not in done), key=lambda k: st["offs"][k])
            if not names:
                return
            done.update(names)
            for k in names:
                t = grads_ref[k]
                g[st["offs"][k]:st["offs"][k] + t.numel()].copy_(t.reshape(-1))
            # maximal runs of whole 1024-element rows (a bucket that was split by layers leaves a hole in the middle)
            runs, lo, hi = [], None, None
            for k in names:
                a = st["offs"][k] // 1024
                b = -(-(st["offs"][k] + grads_ref[k].numel()) // 1024)
                if lo is None:
                    lo, hi = a, b
                elif a <= hi:
                    hi = max(hi, b)
                else:
                    runs.append((lo, hi))
                    lo, hi = a, b
            runs.append((lo, hi))
            for lo, hi in runs:
                L.cast_pad(g.view(-1, 1024)[lo:hi], hi - lo, 1024, 1024, out=st["g16"].view(-1, 1024)[lo:hi], sanitize=True)
                if reduce_now:
                    self.last_bucket_bytes.append((hi - lo) * 2048)
                    if not _SKIP_GRAD_ALLREDUCE:
                        start_all_reduce(st["g16"][lo * 1024:hi * 1024])

        overlap = world > 1 and _OVERLAP_ALLREDUCE
        grads_ref: Dict[str, torch.Tensor] = {}

        def on_ready(grads_so_far, prefixes, name_filter=None):
            grads_ref.update(grads_so_far)
            flush([k for k in st["names"] if k.split("/")[0] in prefixes and k in grads_so_far
                   and (name_filter is None or name_filter(k))], True)

        loss, info, grads = self.loss_and_grads(batch, group=group, on_ready=on_ready if overlap else None, comm=comm)
        grads_ref.update(grads)
        if overlap:
            g[n:n + 3].copy_(grads["contrastive_scales"].reshape(-1))
            L.cast_pad(g.view(-1, 1024)[n // 1024:], 1, 1024, 1024, out=st["g16"].view(-1, 1024)[n // 1024:], sanitize=True)
            if not _SKIP_GRAD_ALLREDUCE:
                start_all_reduce(st["g16"][n:])
            left = [k for k in st["names"] if k not in done]
            assert not left, f"gradients outside the all-reduce buckets: {left[:4]}"
            for w in pending:                                # the current stream waits for NCCL's stream
                if w is not None:
                    w.wait()
            if comm is not None and pending:
                torch.cuda.current_stream().wait_stream(self._comm_stream)
            g16 = st["g16"]
        else:
            for name in st["names"]:
                t = grads[name]
                g[st["offs"][name]:st["offs"][name] + t.numel()].copy_(t.reshape(-1))
            g[n:n + 3].copy_(grads["contrastive_scales"].reshape(-1))
            g16 = L.cast_pad(g.view(-1, 1024), (n + 1024) // 1024, 1024, 1024, out=st["g16"].view(-1, 1024), sanitize=True).view(-1)
            if world > 1 and not _SKIP_GRAD_ALLREDUCE:
                comm.all_reduce(g16) if comm is not None else dist.all_reduce(g16, group=group)
        b1, b2, eps, wd = opt.get("beta_1", 0.9), opt.get("beta_2", 0.98), opt.get("eps", 1e-8), opt["weight_decay_rate"]
        lr, corr, t1 = self._step_scalars(opt)
        gm = getattr(self, "_graph_scalars", None)
        L.adamw_multi(g16, st["table"], st["p"], st["m"], st["v"], lr=lr, beta1=b1, beta2=b2, eps=eps, weight_decay=wd,
                      grad_scale=1.0 / world, bias_corr1=(1 - b1 ** t1) if corr else 1.0,
                      bias_corr2=(1 - b2 ** t1) if corr else 1.0, hyper_dev=None if gm is None else gm["hyper"])
        if gm is not None:                 # captured: the host part below runs after every replay (train_step_graphed)
            gm["g_scales"] = g16[n:n + 3]
            return loss, info
        self._update_temperatures(g16[n:n + 3].float().cpu().numpy(), opt, world)
        return loss, info

    # --------------------------------------------------------------------------------------------
    # transformer encoder with a tape (modeling.py:283-376)
    # --------------------------------------------------------------------------------------------
    def _enc_fwd(self, x0: torch.Tensor, enc: _PackedEncoder, B: int, T: int, rot, valid=None, group=None):
        H = enc.H
        x = L.layernorm(x0, *enc.pre_ln)
        tape = {"x0": x0, "layers": [], "B": B, "T": T, "rot": rot, "valid": valid, "group": group}
        for lw in enc.layers:
            h1 = L.layernorm(x, *lw["ln1"])
            qkv = L.gemm(h1, lw["wqkv"], bias=lw["bqkv"], rot=rot, hidden=H, q_scale=0.125)
            lse = torch.empty((B, H // 64, T), dtype=torch.float32, device=x.device)
            ctx = L.attention(qkv, B, T, H, valid=valid, group=group, lse=lse)
            xm = L.gemm(ctx, lw["wo"], residual=x)
            h2 = L.layernorm(xm, *lw["ln2"])
            # bias + my_gelu in the epilogue, with the pre-activation as a second output (MR_EPI_GELU_DUAL): the
            # backward needs both (gelu' reads `up`, the MLP-out wgrad reads `u`)
            up = torch.empty((h2.shape[0], lw["w1"].shape[0]), dtype=BF16, device=h2.device)
            u = L.gemm(h2, lw["w1"], bias=lw["b1"], gelu=True, pre_out=up)
            xn = L.gemm(u, lw["w2"], residual=xm)
            tape["layers"].append((x, h1, qkv, lse, ctx, xm, h2, up, u))
            x = xn
        return x, tape

    def _enc_bwd(self, dx: torch.Tensor, tape, enc: _PackedEncoder, G: _Grads, prefix: str, on_layer=None) -> torch.Tensor:
        """on_layer(i): called when the gradients of layers i .. last are final (layer i has just been differentiated)."""
        H, B, T = enc.H, tape["B"], tape["T"]
        for i in reversed(range(len(enc.layers))):
            lw = enc.layers[i]
            x, h1, qkv, lse, ctx, xm, h2, up, u = tape["layers"][i]
            p = f"{prefix}/layer_{i:02d}"
            G.add(f"{p}/mlp_layer/out/kernel", _wgrad(dx, u))
            if _FUSED_DGELU:          # d(up) = (dx . W2) * gelu'(up) in the dgrad's epilogue (MR_EPI_DGELU)
                dup = L.gemm_bwd(dx, lw["w2"], trans_b=True, out_f32=False, dgelu_pre=up)
            else:
                dup = L.gelu(up, _dgrad(dx, lw["w2"]))
            dh2 = _dgrad(dup, lw["w1"])
            G.add(f"{p}/mlp_layer/intermediate/kernel", _wgrad(dup, h2))
            G.add(f"{p}/mlp_layer/intermediate/bias", L.colsum(dup))
            dxm, dg, db = L.layernorm_bwd(dh2, xm, lw["ln2"][0], dres=dx)
            G.add(f"{p}/pre_mlp_ln/scale", dg)
            G.add(f"{p}/pre_mlp_ln/bias", db)
            dctx = _dgrad(dxm, lw["wo"])
            G.add(f"{p}/attention_layer/attn_proj/kernel", _wgrad(dxm, ctx))
            # d(q, k, v) -> backward of the qkv epilogue (rotary, 1/sqrt(d)) -> bias gradient: one call, one tail pass
            if _FUSED_QKV_TAIL:
                dqkv, dbqkv = L.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=tape["valid"], group=tape["group"],
                                              rot=tape["rot"], q_scale=0.125, bias_grad=True)
            else:
                dqkv = L.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=tape["valid"], group=tape["group"])
                L.rotary_bwd(dqkv, tape["rot"], H)
                dbqkv = L.colsum(dqkv)
            dh1 = _dgrad(dqkv, lw["wqkv"])
            G.add(f"{p}/attention_layer/qkv/kernel", _wgrad(dqkv, h1))
            G.add(f"{p}/attention_layer/qkv/bias", dbqkv)
            dx, dg, db = L.layernorm_bwd(dh1, x, lw["ln1"][0], dres=dxm)
            G.add(f"{p}/pre_attn_ln/scale", dg)
            G.add(f"{p}/pre_attn_ln/bias", db)
            if on_layer is not None:
                on_layer(i)
        dx0, dg, db = L.layernorm_bwd(dx, tape["x0"], enc.pre_ln[0])
        G.add(f"{prefix}/pre_ln/scale", dg)
        G.add(f"{prefix}/pre_ln/bias", db)
        return dx0

    # --- final_ln + cls_proj (modeling.py:366-372) ---
    def _final_fwd(self, x_last, enc: _PackedEncoder, B, T, want_cls):
        xf = L.layernorm(x_last, *enc.final_ln)
        cls_rows = cls_out = None
        if want_cls:
            cls_rows = xf.view(B, T, enc.H)[:, 0].contiguous()
            cls_out = L.gemm(cls_rows, enc.cls_proj[0], bias=enc.cls_proj[1])
        return xf, cls_rows, cls_out

    def _final_bwd(self, dxf, d_cls_out, x_last, cls_rows, enc, B, T, G: _Grads, prefix):
        """dxf: gradient w.r.t. the final_ln output [B*T, H] bf16 (modified in place); d_cls_out bf16 [B, H] or None."""
        if d_cls_out is not None:
            G.add(f"{prefix}/cls_proj/kernel", _wgrad(d_cls_out, cls_rows))
            G.add(f"{prefix}/cls_proj/bias", L.colsum(d_cls_out))
            d_rows = _dgrad(d_cls_out, enc.cls_proj[0])
            v = dxf.view(B, T, enc.H)
            v[:, 0] = L.add(v[:, 0].contiguous(), d_rows)
        dx_last, dg, db = L.layernorm_bwd(dxf, x_last, enc.final_ln[0])
        G.add(f"{prefix}/final_ln/scale", dg)
        G.add(f"{prefix}/final_ln/bias", db)
        return dx_last

    # --- attention pools (modeling.py:413-429, :464-475) ---
    def _pool_fwd(self, xf, pool: _PackedPool, B, T, geom):
        H = self.H
        q_in = L.window_mean(xf, B, T, H, *geom)
        q = L.gemm(q_in, pool.wq, bias=pool.bq)
        kv = L.gemm(xf, pool.wkv, bias=pool.bkv)
        ctxp = L.window_attention(q, kv, B, T, H, *geom)
        return L.gemm(ctxp, pool.wo, bias=pool.bo), (q_in, q, kv, ctxp)

    def _pool_bwd(self, d_pooled, xf, saved, pool: _PackedPool, B, T, geom, G: _Grads, prefix) -> torch.Tensor:
        H = self.H
        q_in, q, kv, ctxp = saved
        G.add(f"{prefix}/out/kernel", _wgrad(d_pooled, ctxp))
        G.add(f"{prefix}/out/bias", L.colsum(d_pooled))
        d_ctxp = _dgrad(d_pooled, pool.wo)
        dq, dkv = L.window_attention_bwd(q, kv, d_ctxp, B, T, H, *geom)
        dwkv = _wgrad(dkv, xf)                       # [2H, H]: key rows then value rows
        dbkv = L.colsum(dkv)
        G.add(f"{prefix}/key/kernel", dwkv[:H])
        G.add(f"{prefix}/value/kernel", dwkv[H:])
        G.add(f"{prefix}/key/bias", dbkv[:H])
        G.add(f"{prefix}/value/bias", dbkv[H:])
        dxf = _dgrad(dkv, pool.wkv)
        G.add(f"{prefix}/query/kernel", _wgrad(dq, q_in))
        G.add(f"{prefix}/query/bias", L.colsum(dq))
        dq_in = _dgrad(dq, pool.wq)
        L.window_mean_bwd(dq_in, dxf, B, T, H, *geom)
        return dxf

    # --- embedding front ends ---
    def _embed_rows_bwd(self, dx0, a, B, T, G: _Grads, prefix):
        """Backward of `x0[:, 1:] = a . W_emb + b`, `x0[:, 0] = cls` (modeling.py:402,316-320)."""
        H = self.H
        v = dx0.view(B, T, H)
        G.add(f"{prefix}/transformer/cls", v[:, 0].float().sum(0))
        dxe = v[:, 1:].reshape(-1, H).contiguous()
        G.add(f"{prefix}/embedding/kernel", _wgrad(dxe, a))          # [H, K (padded)]
        G.add(f"{prefix}/embedding/bias", L.colsum(dxe))

    # --------------------------------------------------------------------------------------------
    def loss_and_grads(self, batch: Dict, group=None, on_ready=None, comm=None):
        """on_ready(grads, prefixes): called during the backward pass each time every gradient under the given
        top-level parameter prefixes is final (train_step starts their all-reduce under the remaining backward)."""
        import torch.distributed as dist
        m, p, H = self.model, self.model._p, self.H
        cfg, data = self.cfg, self.cfg["data"]
        G_, nsub = data["num_segment_groups"], data["num_audio_subsegments"]
        Ltxt, Lseq = data["lang_seq_len"], data["seq_len"]
        n_a2t, n_t2a = data["num_audio2text_seqs"], data["num_text2audio_seqs"]
        gh, gw = m.output_grid_h, m.output_grid_w
        hw, atl = gh * gw, cfg["audio_token_length"]
        G = _Grads()

        images = _dev(batch["images"], torch.float32)
        dev = images.device
        b, nseg_hw, pp3 = images.shape
        nseg = nseg_hw // hw
        per_group, n_aud = nseg // G_, nseg * nsub
        i32 = lambda k, cols: _dev(batch[k], torch.int32).reshape(-1, cols).contiguous()

        # ================================ forward ================================
        # ---- vision (modeling.py:391-430) ----
        Bv, Tv = b * nseg, hw + 1
        av = L.cast_pad(images.reshape(Bv * hw, pp3), Bv * hw, pp3, pp3)
        xv0 = torch.empty((Bv * Tv, H), dtype=BF16, device=dev)
        L.gemm(av, p.vis_emb[0], bias=p.vis_emb[1], out=xv0, seq_in=hw, seq_out=Tv, row_off=1)
        L.fill_cls(xv0, p.vis.cls, Bv, Tv, H)
        xv_last, tape_v = self._enc_fwd(xv0, p.vis, Bv, Tv, m._table("vision"))
        xvf, v_cls_rows, v_cls = self._final_fwd(xv_last, p.vis, Bv, Tv, True)
        geom_v = (1, 4, (gh // 2) * (gw // 2), gh, gw)
        v_pooled, v_pool_saved = self._pool_fwd(xvf, p.vis_pool, Bv, Tv, geom_v)
        hw1 = hw // 4
        vis_len = per_group * hw1
        imgs_seq = v_pooled.view(b * G_, vis_len, H)

        # ---- audio (modeling.py:443-476) ----
        audio = _dev(batch["audio_clips"], torch.float32).reshape(b * n_aud, cfg["audio_seq_length"], -1)
        Ba, na = audio.shape[0], cfg["audio_seq_length"] // cfg["audio_patch_size"]
        Ta = na + 1
        aa = L.cast_pad(audio.reshape(Ba * na, -1), Ba * na, cfg["audio_patch_size"] * 65, _AUDIO_K_PAD)
        xa0 = torch.empty((Ba * Ta, H), dtype=BF16, device=dev)
        L.gemm(aa, p.aud_emb[0], bias=p.aud_emb[1], out=xa0, seq_in=na, seq_out=Ta, row_off=1)
        L.fill_cls(xa0, p.aud.cls, Ba, Ta, H)
        xa_last, tape_a = self._enc_fwd(xa0, p.aud, Ba, Ta, m._table("audio"))
        xaf, a_cls_rows, audio_cls = self._final_fwd(xa_last, p.aud, Ba, Ta, True)
        r = m.audio_pooling_ratio
        geom_a = (0, r, na // r, 0, 0)
        a_pooled, a_pool_saved = self._pool_fwd(xaf, p.aud_pool, Ba, Ta, geom_a)
        audio_seq = a_pooled.view(b, n_aud, atl, H)

        # ---- joint inputs (pretrain_model.py:97-144) ----
        tok = {k: i32(k, Ltxt) for k in ("text2audio", "text2audio/audio_ptr", "text2audio/text_ptr", "audio2text",
                                         "audio2text/audio_ptr", "audio2text/text_ptr")}
        tok.update({k: i32(k, Lseq) for k in ("random_text", "random_text/text_ptr", "audio_text_matching",
                                              "audio_text_matching/audio_ptr")})
        text_spans = i32("text_spans", cfg["text_span_length"])
        n_spans = text_spans.shape[0] // b
        vsi = _dev(batch["video_src_index"], torch.int64).reshape(b, G_, per_group)
        rng = m._draw_rng(batch, b * G_ * max(n_a2t, n_t2a), per_group, b, n_spans, dev)
        rows = {"audio2text": b * G_ * n_a2t, "audio_text_matching": b, "random_text": b, "text2audio": b * G_ * n_t2a}
        keys = sorted(rows)
        R = sum(rows.values())
        xj0 = torch.empty((R * Lseq, H), dtype=BF16, device=dev)
        rotj = torch.empty((R * Lseq, 32), dtype=torch.float32, device=dev)
        validj = torch.empty((R * Lseq,), dtype=torch.uint8, device=dev)
        groupj = torch.zeros((R, Lseq), dtype=torch.int32, device=dev)
        start, r0 = {}, 0
        for k in keys:
            start[k] = r0
            r0 += rows[k]
        sl = lambda k: slice(start[k] * Lseq, (start[k] + rows[k]) * Lseq)
        views = lambda k: (xj0[sl(k)], rotj[sl(k)], validj[sl(k)])
        hp, wp = gh // 2, gw // 2
        vis_seg = torch.arange(per_group, device=dev).repeat_interleave(hw1)

        def video_src(seg_idx, n_seqs, split):
            v = vsi.repeat(1, n_seqs, 1).reshape(-1, per_group)
            v = m._augment_video_src_idx(v, split.to(torch.int64))
            parts = [torch.gather(v, 1, seg_idx.long()), torch.gather(v, 1, vis_seg[None].expand(v.shape[0], -1))]
            pad = Lseq - Ltxt - vis_len
            if pad > 0:
                parts.append(torch.full((v.shape[0], pad), -1, dtype=torch.int64, device=dev))
            return torch.cat(parts, 1).to(torch.int32)

        asm = {}          # per key: kwargs shared by the forward assembly and its backward
        for k, n_seqs, use_audio in (("audio2text", n_a2t, True), ("text2audio", n_t2a, False)):
            ptr = tok[k + "/audio_ptr"]
            seg = torch.remainder(torch.div(ptr, nsub, rounding_mode="floor"), per_group).to(torch.int32).contiguous()
            asm[k] = dict(aud_ptr=ptr if use_audio else None, n_aud=n_aud if use_audio else 0, n_vis=vis_len, T=Lseq,
                          audio_token_length=atl, aud_div=G_ * n_seqs, vis_group=G_, vis_tile=n_seqs)
            L.embed_assemble(tok[k], p.tok_emb, H=H, seg_idx=seg, audio=audio_seq if use_audio else None, vision=imgs_seq,
                             pool_h=hp, pool_w=wp, out=views(k), **asm[k])
            groupj[start[k]:start[k] + rows[k]] = video_src(seg, n_seqs, rng["_rng/split_" + k])
        atm = tok["audio_text_matching"]
        is_pool = atm == LTOVPOOL
        pool_cum = torch.cumsum(is_pool.to(torch.int32), -1).to(torch.int32)
        asm["audio_text_matching"] = dict(aud_ptr=tok["audio_text_matching/audio_ptr"], n_aud=n_aud, T=Lseq,
                                          audio_token_length=atl)
        L.embed_assemble(atm, p.tok_emb, H=H, seg_idx=pool_cum.contiguous(), audio=audio_seq,
                         out=views("audio_text_matching"), **asm["audio_text_matching"])
        asm["random_text"] = dict(T=Lseq)
        L.embed_assemble(tok["random_text"], p.tok_emb, H=H, out=views("random_text"), **asm["random_text"])

        xj_last, tape_j = self._enc_fwd(xj0, p.joint, R, Lseq, rotj, valid=validj, group=groupj.reshape(-1))
        xjf, _, _ = self._final_fwd(xj_last, p.joint, R, Lseq, False)
        mo_all = L.gemm(xjf, p.head[0], bias=p.head[1])                      # joint_proj on every row (:151), bf16
        mo = {k: mo_all[sl(k)].view(rows[k], Lseq, H) for k in keys}

        # ---- pools and selections (pretrain_model.py:160-233); text2audio / audio2text pool over all 640 rows of a
        #      sequence with do_pool false beyond lang_seq_len, which equals the reference's [:, :lang_seq_len] slice ----
        def pad_to_seq(t, fill):
            return torch.nn.functional.pad(t, (0, Lseq - t.shape[1]), value=fill).contiguous()

        pools = []        # (do_pool u8 [b, l2], idx [b, l2], key, S, which pooled gradient it receives)

        def pool(do_pool, idx, k, S, acc=None, tag=None):
            l2 = (rows[k] * Lseq) // b
            dp = do_pool.to(torch.uint8).reshape(b, l2).contiguous()
            ix = idx.reshape(b, l2).contiguous()
            pools.append((dp, ix, k, S, tag))
            return L.segment_pool(dp, ix, mo[k].reshape(b, l2, H), S, accumulate_into=acc)

        a2v, _ = pool(is_pool, pool_cum - 1, "audio_text_matching", nseg, tag="a2v")
        t2a_tok = pad_to_seq(tok["text2audio"], 0)
        t2a_x, is_selected = pool(t2a_tok == MASKAUDIO, pad_to_seq(tok["text2audio/audio_ptr"], -1), "text2audio", n_aud,
                                  tag="t2a")
        n_trg = int(n_aud * data["mask_rate"]) * n_t2a
        idx_sort = torch.sort(-is_selected, dim=-1, stable=True)[1]
        base = (torch.arange(b, device=dev) * n_aud)[:, None]
        best_rows = (idx_sort[:, :n_trg] + base).reshape(-1).to(torch.int32).contiguous()
        extra_rows = (idx_sort[:, n_trg:] + base).reshape(-1).to(torch.int32).contiguous()
        counts, sel = {}, None
        for k in ("text2audio", "audio2text", "random_text"):
            tk = tok[k] if tok[k].shape[1] == Lseq else pad_to_seq(tok[k], 0)
            pk = tok[k + "/text_ptr"] if tok[k].shape[1] == Lseq else pad_to_seq(tok[k + "/text_ptr"], -1)
            if sel is None:
                sel = pool(tk == MASK, pk, k, n_spans, tag="t2sp")
                counts[k] = sel[1].clone()
            else:
                before = sel[1].clone()
                pool(tk == MASK, pk, k, n_spans, acc=sel, tag="t2sp")
                counts[k] = sel[1] - before
        t2sp_x, t2sp_ct = sel
        src = torch.stack([torch.zeros_like(t2sp_ct), counts["text2audio"], counts["audio2text"],
                           counts["random_text"]], -1).argmax(-1) - 1
        is_valid = (text_spans != PADDING).any(-1).reshape(b, n_spans) & (t2sp_ct > 0)
        logits_for_pred = is_valid.float() * 1e6 + math.log(4.0) * (counts["text2audio"] + counts["audio2text"])
        z = -torch.log(-torch.log(rng["_rng/gumbel_u"].float()))
        n_inc = data["num_text_spans_to_include"]
        best = torch.sort(-(logits_for_pred + z).reshape(-1), stable=True)[1][:n_inc * b]
        best32 = best.to(torch.int32).contiguous()

        # ---- span encoder (modeling.py:488-504) on the selected spans ----
        spans_sel = text_spans[best].contiguous()
        Bs, Ts = spans_sel.shape[0], spans_sel.shape[1] + 1
        xs0, _, valid_s = L.embed_assemble(spans_sel, p.tok_emb, H=H, cls_offset=1)
        L.fill_cls(xs0, p.span.cls, Bs, Ts, H)
        xs_last, tape_s = self._enc_fwd(xs0, p.span, Bs, Ts, m._table("span"), valid=valid_s)
        xsf, s_cls_rows, sp_cls = self._final_fwd(xs_last, p.span, Bs, Ts, True)

        # ---- heads: unit_normalize * exp(scale / 2) -> bf16 (pretrain_model.py:236-257) ----
        ls_raw = p.contrastive_scales.astype(np.float32)
        ls = np.minimum(ls_raw, np.float32(math.log(100.0)))
        temp = [float(np.exp(v / 2.0)) for v in ls]
        gm = getattr(self, "_graph_scalars", None)     # CUDA-graph mode: temperatures live in device memory
        tkw = (lambda hi: dict(scale=temp[hi])) if gm is None else (lambda hi: dict(scale=1.0, scale_dev=gm["temp"][hi:hi + 1]))
        un = lambda v, hi, idx=None: L.unit_normalize(v.reshape(-1, H), out_f32=False, row_idx=idx, **tkw(hi))
        heads = {                       # name -> {field: (normalised output, source tensor, row gather, temperature)}
            "imgs_to_audio": {"x": (a2v, None), "y": (v_cls, None)},
            "text_to_audio": {"x": (t2a_x, best_rows), "y": (audio_cls, best_rows), "y_extra": (audio_cls, extra_rows)},
            "stuff_to_span": {"x": (t2sp_x, best32), "y": (sp_cls, None)},
        }
        preds = {}
        for hi, (name, fields) in enumerate(heads.items()):
            preds[name] = {f: un(srct, hi, idx) for f, (srct, idx) in fields.items()}
        preds["stuff_to_span"]["_sources"] = src.reshape(-1)[best].to(torch.int32).contiguous()

        # ---- loss forward (pretrain_model.py:262-303) ----
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        if comm is None and world > 1:
            comm = self.communicator(group)
        jobs = gather_contrastive_targets(preds, group, comm=comm)
        info, lses = {}, []
        for name, x, pos_off, y_all in jobs:
            lse, numer = L.contrastive_lse(x, y_all, pos_off)
            lses.append(lse)
            red = L.contrastive_reduce(lse, numer, preds[name].get("_sources"))
            info[name] = info.get(name, 0.0) + red[0] / 2.0
            if "_sources" in preds[name]:
                for i, k in enumerate(("text2audio", "audio2text", "random_text")):
                    key = f"_{name}_from_{k}"
                    info[key] = info.get(key, 0.0) + red[1 + i] / 2.0
        loss = sum(v for k, v in info.items() if not k.startswith("_"))

        # ================================ backward ================================
        # ---- loss: d(preds) (fp32).  The gradients w.r.t. the gathered targets go back to the ranks that own the rows
        # by ONE packed reduce-scatter, the transpose of the packed all-gather (pretrain_model.py:290): every
        # direction's dy_all GEMM writes its rows rank-major into one [world, S, H] buffer (row remap in the epilogue).
        dpred = {name: {f: torch.zeros(t.shape, dtype=torch.float32, device=dev) for f, t in d.items() if f != "_sources"}
                 for name, d in preds.items()}
        n_owns = [y_all.shape[0] // world for _, _, _, y_all in jobs]
        S = sum(n_owns)
        packed = torch.empty((world, S, H), dtype=torch.float32, device=dev)
        dxs, off = [], 0
        for (name, x, pos_off, y_all), lse, n_own in zip(jobs, lses, n_owns):
            dx, _ = L.contrastive_lse_bwd(x, y_all, lse, pos_off, 1.0 / (2.0 * x.shape[0]), dy_out=packed,
                                          dy_seq_in=n_own, dy_seq_out=S, dy_row_off=off)
            dxs.append(dx)
            off += n_own
        if world > 1:
            own_all = torch.empty((S, H), dtype=torch.float32, device=dev)
            if comm is not None:
                comm.reduce_scatter(packed, own_all)
            else:
                dist.reduce_scatter_tensor(own_all, packed, group=group)
        else:
            own_all = packed[0]
        off = 0
        for ji, ((name, x, pos_off, y_all), dx, n_own) in enumerate(zip(jobs, dxs, n_owns)):
            k1, k2 = ("x", "y") if ji % 2 == 0 else ("y", "x")
            dpred[name][k1] += dx
            own = own_all[off:off + n_own]
            off += n_own
            n2 = preds[name][k2].shape[0]
            dpred[name][k2] += own[:n2]
            if (k2 + "_extra") in preds[name]:
                dpred[name][k2 + "_extra"] += own[n2:]

        # ---- unit_normalize backward (+ scatter through the row gathers); d(contrastive_scales) ----
        dsrc: Dict[int, torch.Tensor] = {}
        dscale = torch.zeros(3, dtype=torch.float32, device=dev)
        for hi, (name, fields) in enumerate(heads.items()):
            for f, (srct, idx) in fields.items():
                g_out = dpred[name][f]
                d = L.unit_normalize_bwd(srct.reshape(-1, H), g_out, row_idx=idx, **tkw(hi))
                key = id(srct)
                dsrc[key] = dsrc[key] + d if key in dsrc else d
                # y = exp(s / 2) * unit(x):  dL/ds = 0.5 * sum(dy * y), zero where the clip at log(100) is active
                if gm is not None:
                    dscale[hi] += 0.5 * (g_out * preds[name][f].float()).sum() * gm["unclipped"][hi]
                elif ls_raw[hi] <= math.log(100.0):
                    dscale[hi] += 0.5 * (g_out * preds[name][f].float()).sum()
        G.add("contrastive_scales", dscale)
        d_a2v, d_vcls, d_t2ax = dsrc[id(a2v)], dsrc[id(v_cls)], dsrc[id(t2a_x)]
        d_acls, d_t2sp, d_spcls = dsrc[id(audio_cls)], dsrc[id(t2sp_x)], dsrc[id(sp_cls)]

        # ---- span encoder ----
        dxsf = torch.zeros_like(xsf)
        dxs_last = self._final_bwd(dxsf, d_spcls.to(BF16), xs_last, s_cls_rows, p.span, Bs, Ts, G, "span_encoder/transformer")
        dxs0 = self._enc_bwd(dxs_last, tape_s, p.span, G, "span_encoder/transformer")
        G.add("span_encoder/transformer/cls", dxs0.view(Bs, Ts, H)[:, 0].float().sum(0))
        d_tok_emb = torch.zeros(p.tok_emb.shape, dtype=torch.float32, device=dev)
        L.embed_assemble_bwd(spans_sel, dxs0, H=H, vocab=p.tok_emb.shape[0], d_emb=d_tok_emb, cls_offset=1)

        # ---- pools -> joint_proj -> joint transformer ----
        dmo_all = torch.zeros_like(mo_all)
        dpooled = {"a2v": d_a2v.view(b, nseg, H), "t2a": d_t2ax.view(b, n_aud, H), "t2sp": d_t2sp.view(b, n_spans, H)}
        for dp, ix, k, S, tag in pools:
            l2 = dp.shape[1]
            dv = L.segment_pool_bwd(dp, ix, dpooled[tag].contiguous(), l2).view(-1, H)
            L.add(dmo_all[sl(k)], dv, out=dmo_all[sl(k)])
        G.add("head/kernel", _wgrad(dmo_all, xjf))
        G.add("head/bias", L.colsum(dmo_all))
        dxjf = _dgrad(dmo_all, p.head[0])
        dxj_last = self._final_bwd(dxjf, None, xj_last, None, p.joint, R, Lseq, G, "joint_transformer")
        dxj0 = self._enc_bwd(dxj_last, tape_j, p.joint, G, "joint_transformer")
        d_audio_seq = torch.zeros(audio_seq.shape, dtype=torch.float32, device=dev)
        d_imgs_seq = torch.zeros(imgs_seq.shape, dtype=torch.float32, device=dev)
        for k in keys:
            kw = asm[k]
            L.embed_assemble_bwd(tok[k], dxj0[sl(k)], H=H, vocab=p.tok_emb.shape[0], d_emb=d_tok_emb,
                                 d_audio=d_audio_seq if kw.get("n_aud", 0) else None,
                                 d_vision=d_imgs_seq if kw.get("n_vis", 0) else None, **kw)
        G.add("token_encoder/Embed_0/embedding", d_tok_emb)
        if on_ready is not None:
            on_ready(G.g, ALLREDUCE_BUCKETS[0])

        # ---- audio encoder ----
        dxaf = self._pool_bwd(d_audio_seq.view(-1, H).to(BF16), xaf, a_pool_saved, p.aud_pool, Ba, Ta, geom_a, G,
                              "audio_encoder/seq_attnpool")
        dxa_last = self._final_bwd(dxaf, d_acls.to(BF16), xa_last, a_cls_rows, p.aud, Ba, Ta, G, "audio_encoder/transformer")
        dxa0 = self._enc_bwd(dxa_last, tape_a, p.aud, G, "audio_encoder/transformer")
        self._embed_rows_bwd(dxa0, aa, Ba, Ta, G, "audio_encoder")
        if on_ready is not None:
            on_ready(G.g, ALLREDUCE_BUCKETS[1])

        # ---- vision encoder ----
        dxvf = self._pool_bwd(d_imgs_seq.view(-1, H).to(BF16), xvf, v_pool_saved, p.vis_pool, Bv, Tv, geom_v, G,
                              "vision_encoder/seq_attnpool")
        dxv_last = self._final_bwd(dxvf, d_vcls.to(BF16), xv_last, v_cls_rows, p.vis, Bv, Tv, G, "vision_encoder/transformer")
        def vision_layers_done(i):
            if on_ready is not None and i in VISION_LAYER_SPLITS and i > 0:
                upper = tuple(f"vision_encoder/transformer/layer_{j:02d}/" for j in range(i, len(p.vis.layers)))
                on_ready(G.g, ALLREDUCE_BUCKETS[2], name_filter=lambda k: k.startswith(upper))

        dxv0 = self._enc_bwd(dxv_last, tape_v, p.vis, G, "vision_encoder/transformer", on_layer=vision_layers_done)
        self._embed_rows_bwd(dxv0, av, Bv, Tv, G, "vision_encoder")
        if on_ready is not None:
            on_ready(G.g, ALLREDUCE_BUCKETS[2])
        if getattr(self, "_keep_locals", False):
            self._dbg = dict(locals())
        return loss, info, G.g

    # --------------------------------------------------------------------------------------------
    # TrainState <-> checkpoint (mreserve/checkpoint.py:61-119): parameters and moments in the flax tree / layouts
    # --------------------------------------------------------------------------------------------
    def _flax_shape(self, path: str):
        node = self.params
        for part in path.split("/"):
            node = node[part]
        return np.asarray(node).shape

    def _to_flax(self, name: str, t: torch.Tensor) -> np.ndarray:
        """Packed layout (K-major [out, in(+pad)] kernels) -> the flax leaf; bf16 comes back as ml_dtypes.bfloat16."""
        shp = self._flax_shape(name)
        if name.endswith("kernel"):
            t = t.t()
            t = t[: int(np.prod(shp)) // t.shape[1]]         # drops the audio embedding's K padding rows
        a = t.contiguous().reshape(shp).cpu()
        return a.view(torch.int16).numpy().view(_np_dtype(BF16)) if a.dtype == BF16 else a.numpy()

    def _from_flax(self, name: str, a: np.ndarray, like: torch.Tensor) -> torch.Tensor:
        """The inverse of `_to_flax`: a tensor shaped and typed like the packed slot `like`."""
        a = np.array(a, copy=True)                           # msgpack hands out read-only buffers
        t = torch.from_numpy(a.view(np.int16)).view(BF16) if a.dtype == _np_dtype(BF16) else torch.from_numpy(a)
        if name.endswith("kernel"):
            out_f = like.shape[0]
            t = t.reshape(-1, out_f).t()
            if t.shape[1] < like.shape[1]:
                t = torch.nn.functional.pad(t, (0, like.shape[1] - t.shape[1]))
        return t.reshape(like.shape).to(dtype=like.dtype, device=like.device)

    @staticmethod
    def _nest(flat: Dict[str, np.ndarray]) -> Dict:
        tree: Dict = {}
        for k, v in flat.items():
            node = tree
            parts = k.split("/")
            for part in parts[:-1]:
                node = node.setdefault(part, {})
            node[parts[-1]] = v
        return tree

    def state_dict(self) -> Dict:
        """The reference's TrainState as `save_checkpoint` sees it (checkpoint.py:61-99): {'step', 'params',
        'opt_state'} with fp32 parameters and the Adam moments in their stored form (bf16 mu, coded bf16 nu).
        `opt_state` follows flax's state dict of the optax chain of optimization.py:173-193 -- a tuple becomes
        {'0': ..} -- i.e. '0' = ScaleByAdamState{count, mu, nu}, '1' = MaskedState{inner_state: {}} of
        add_decayed_weights, '2' = ScaleByScheduleState{count}, '3' = ScaleState {} (optax 0.0.9 layout restated
        from library knowledge; not verifiable offline)."""
        if self._state is None:
            self._init_state()
        views = self.optimizer_state()
        trees = {key: {n: self._to_flax(n, d[key]) for n, d in views.items()} for key in ("p", "m", "v")}
        trees["p"]["contrastive_scales"] = np.asarray(self.model._p.contrastive_scales, np.float32).copy()
        trees["m"]["contrastive_scales"] = self._scale_state["m"].copy()
        trees["v"]["contrastive_scales"] = self._scale_state["v"].copy()
        count = np.int32(self.step_count)
        return {"step": int(self.step_count), "params": self._nest(trees["p"]),
                "opt_state": {"0": {"count": count, "mu": self._nest(trees["m"]), "nu": self._nest(trees["v"])},
                              "1": {"inner_state": {}}, "2": {"count": count}, "3": {}}}

    def load_state_dict(self, state: Dict, opt: Optional[Dict] = None) -> None:
        """Restores parameters, moments and the step from `checkpoint.load_checkpoint(...)`; `opt_state` may be None
        (checkpoints written with no_optimizer, checkpoint.py:84-90): the moments then restart from zero."""
        from .checkpoint import flatten_dict
        if self._state is None:
            self._init_state(opt)
        views, st = self.optimizer_state(), self._state
        flat_p = flatten_dict(state["params"])
        adam = (state.get("opt_state") or {}).get("0")
        flat_m = flatten_dict(adam["mu"]) if adam else None
        flat_v = flatten_dict(adam["nu"]) if adam else None
        reg = self._registry()
        for name, d in views.items():
            d["p"].copy_(self._from_flax(name, np.asarray(flat_p[name], np.float32), d["p"]))
            live = reg[name][0]
            live.copy_(d["p"].to(live.dtype))
            for key, flat in (("m", flat_m), ("v", flat_v)):
                if flat is None:
                    d[key].zero_()
                else:
                    d[key].copy_(self._from_flax(name, np.asarray(flat[name]), d[key]))
        self.model._p.contrastive_scales = np.asarray(flat_p["contrastive_scales"], np.float32).copy()
        sdt = self._scale_state["m"].dtype
        self._scale_state = {"m": np.asarray(flat_m["contrastive_scales"]).astype(sdt) if flat_m else np.zeros(3, sdt),
                             "v": np.asarray(flat_v["contrastive_scales"]).astype(sdt) if flat_v else np.zeros(3, sdt)}
        self.step_count = int(np.asarray(state["step"]).reshape(-1)[0])
        del st

    def grads_to_flax(self, g: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
        """Packed-layout gradients -> the flax shapes of the reference's parameter tree (numpy, fp32)."""
        out = {}
        H = self.H

        def shape_of(path):
            node = self.params
            for part in path.split("/"):
                node = node[part]
            return np.asarray(node).shape

        for name, t in g.items():
            shp = shape_of(name)
            t = t.detach().float()
            if name.endswith("kernel"):
                t = t.t()                                    # packed [out, in] -> flax [in, out]
                if name == "audio_encoder/embedding/kernel":
                    t = t[: shp[0] * shp[1]]                 # drop the K padding rows
            out[name] = t.reshape(shp).cpu().numpy()
        return out
