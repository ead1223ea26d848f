"""ctypes binding of libmreserve_b200.so (include/mreserve_b200.h).

torch is used here only as the owner of device memory and of the CUDA stream: every wrapper passes raw
device pointers + the current stream handle through the C ABI.  There is NO fallback: if the shared
library is missing or the device is not a B200 the import / first call fails loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Sequence, Tuple

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libmreserve_b200.so"

EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_ROTARY_QKV, EPI_OUT_F32, EPI_LN_IN, EPI_STATS_OUT, EPI_DGELU = 1, 2, 4, 8, 16, 32, 64, 128
EPI_GELU_DUAL = 256
ALGO_TCGEN05, ALGO_SIMT_DEBUG = 0, 1


class MrError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("lda", C.c_int32), ("ldd", C.c_int32),
                ("epilogue", C.c_int32), ("seq_in", C.c_int32), ("seq_out", C.c_int32), ("row_off", C.c_int32),
                ("rot_rows", C.c_int32), ("hidden", C.c_int32), ("algo", C.c_int32), ("q_scale", C.c_float),
                ("num_sms", C.c_int32), ("ln_eps", C.c_float), ("algo_pair", C.c_int32),
                ("trans_a", C.c_int32), ("trans_b", C.c_int32), ("ldb", C.c_int32), ("split_k", C.c_int32),
                ("tile_n", C.c_int32)]


class AttentionDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("algo", C.c_int32), ("emit_lse", C.c_int32),
                ("finish", C.c_int32), ("rot_rows", C.c_int32), ("q_scale", C.c_float)]


class LayerNormDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("H", C.c_int32), ("ldx", C.c_int32), ("seq_in", C.c_int32),
                ("seq_out", C.c_int32), ("row_off", C.c_int32), ("eps", C.c_float), ("emit_stats", C.c_int32)]


class EncoderDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("num_layers", C.c_int32),
                ("rot_rows", C.c_int32), ("ln_eps", C.c_float)]


class CastDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int32), ("ld_out", C.c_int32), ("sanitize", C.c_int32),
                ("reserved_", C.c_int32)]


class PatchifyDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("grid_h", C.c_int32), ("grid_w", C.c_int32), ("patch", C.c_int32)]


class LogmelDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("n_samples", C.c_int32), ("playback_speed", C.c_int32), ("pad_size", C.c_int32),
                ("split", C.c_int32), ("build_tables", C.c_int32)]


class ResizePatchifyDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("grid_h", C.c_int32), ("grid_w", C.c_int32),
                ("out_f32", C.c_int32)]


class FillClsDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("H", C.c_int32)]


class WindowDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("mode", C.c_int32), ("win", C.c_int32),
                ("n_win", C.c_int32), ("grid_h", C.c_int32), ("grid_w", C.c_int32)]


class AssembleDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("n_aud", C.c_int32), ("n_vis", C.c_int32),
                ("pool_h", C.c_int32), ("pool_w", C.c_int32), ("T", C.c_int32), ("vocab", C.c_int32),
                ("audio_token_length", C.c_int32), ("cls_offset", C.c_int32), ("aud_div", C.c_int32),
                ("vis_group", C.c_int32), ("vis_tile", C.c_int32), ("seg_div", C.c_int32)]


class EncoderWeights(C.Structure):
    _fields_ = [("pre_ln_g", C.c_void_p), ("pre_ln_b", C.c_void_p), ("final_ln_g", C.c_void_p), ("final_ln_b", C.c_void_p),
                ("cls", C.c_void_p), ("cls_proj_w", C.c_void_p), ("cls_proj_b", C.c_void_p),
                ("layers", C.POINTER(C.c_void_p)), ("num_layers", C.c_int32), ("reserved_", C.c_int32)]


class PoolWeights(C.Structure):
    _fields_ = [("wq", C.c_void_p), ("bq", C.c_void_p), ("wkv", C.c_void_p), ("bkv", C.c_void_p), ("wo", C.c_void_p),
                ("bo", C.c_void_p)]


class VisionDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("grid_h", C.c_int32), ("grid_w", C.c_int32), ("H", C.c_int32), ("in_bf16", C.c_int32),
                ("ln_eps", C.c_float), ("enc", EncoderWeights), ("pool", PoolWeights), ("emb_w", C.c_void_p),
                ("emb_b", C.c_void_p), ("rot", C.c_void_p)]


class AudioDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("seq_len", C.c_int32), ("patch", C.c_int32), ("pool_ratio", C.c_int32), ("H", C.c_int32),
                ("ln_eps", C.c_float), ("enc", EncoderWeights), ("pool", PoolWeights), ("emb_w", C.c_void_p),
                ("emb_b", C.c_void_p), ("rot", C.c_void_p)]


class SpanDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("vocab", C.c_int32), ("normalize", C.c_int32),
                ("ln_eps", C.c_float), ("enc", EncoderWeights), ("tok_emb", C.c_void_p), ("rot", C.c_void_p)]


class JointDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("n_aud", C.c_int32), ("n_vis", C.c_int32),
                ("pool_h", C.c_int32), ("pool_w", C.c_int32), ("T", C.c_int32), ("vocab", C.c_int32),
                ("audio_token_length", C.c_int32), ("aud_div", C.c_int32), ("vis_group", C.c_int32),
                ("vis_tile", C.c_int32), ("seg_div", C.c_int32), ("n_keep", C.c_int32), ("head", C.c_int32),
                ("ln_eps", C.c_float), ("reserved_", C.c_int32), ("enc", EncoderWeights), ("tok_emb", C.c_void_p),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p)]


class EmbedVideoDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("nseg", C.c_int32), ("L", C.c_int32), ("reserved_", C.c_int32),
                ("vision", VisionDesc), ("audio", AudioDesc), ("joint", JointDesc)]


class RotaryDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("ndims", C.c_int32)]


class NormalizeDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("H", C.c_int32), ("in_f32", C.c_int32), ("out_f32", C.c_int32),
                ("scale", C.c_float), ("scale_from_device", C.c_int32)]


class SegmentPoolDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("S", C.c_int32), ("v_f32", C.c_int32),
                ("accumulate", C.c_int32)]


class ContrastiveDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_all", C.c_int32), ("H", C.c_int32), ("pos_off", C.c_int32),
                ("grad_scale", C.c_float), ("dy_seq_in", C.c_int32), ("dy_seq_out", C.c_int32), ("dy_row_off", C.c_int32)]


class ElementwiseDesc(C.Structure):
    _fields_ = [("n", C.c_int64), ("mode", C.c_int32)]


class ColsumDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("ld", C.c_int32)]


class AdamWDesc(C.Structure):
    _fields_ = [("n", C.c_int64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("grad_scale", C.c_float)]


class AdamWChunk(C.Structure):
    _fields_ = [("offset", C.c_int64), ("w", C.c_void_p), ("count", C.c_int32), ("flags", C.c_int32)]


class AdamWMultiDesc(C.Structure):
    _fields_ = [("n_chunks", C.c_int32), ("moments_bf16", C.c_int32), ("grad_bf16", C.c_int32), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("grad_scale", C.c_float), ("bias_corr1", C.c_float), ("bias_corr2", C.c_float),
                ("one_minus_beta1", C.c_float), ("one_minus_beta2", C.c_float), ("hyper_from_device", C.c_int32)]


ADAMW_DECAY, ADAMW_W_F32, ADAMW_CHUNK = 1, 2, 2048


class RotaryBwdDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("hidden", C.c_int32), ("rot_rows", C.c_int32), ("q_scale", C.c_float)]


class RowDotDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("H", C.c_int32)]


class CollectiveDesc(C.Structure):
    _fields_ = [("comm", C.c_void_p), ("count", C.c_int64), ("dtype", C.c_int32), ("average", C.c_int32)]


# every compute entry point has the same signature
ENTRY_POINTS = {
    "mr_gemm_epilogue": GemmDesc,
    "mr_attention_fwd": AttentionDesc,
    "mr_attention_bwd": AttentionDesc,
    "mr_layernorm_fwd": LayerNormDesc,
    "mr_transformer_encode": EncoderDesc,
    "mr_vision_encode": VisionDesc,
    "mr_audio_encode": AudioDesc,
    "mr_span_encode": SpanDesc,
    "mr_joint_encode": JointDesc,
    "mr_embed_video_batched": EmbedVideoDesc,
    "mr_cast_pad": CastDesc,
    "mr_patchify_u8": PatchifyDesc,
    "mr_logmel": LogmelDesc,
    "mr_resize_patchify": ResizePatchifyDesc,
    "mr_fill_cls": FillClsDesc,
    "mr_window_mean": WindowDesc,
    "mr_window_attention": WindowDesc,
    "mr_embed_assemble": AssembleDesc,
    "mr_rotary_table": RotaryDesc,
    "mr_unit_normalize": NormalizeDesc,
    "mr_segment_pool": SegmentPoolDesc,
    "mr_contrastive_lse_fwd": ContrastiveDesc,
    "mr_contrastive_reduce": ContrastiveDesc,
    "mr_row_dot": RowDotDesc,
    "mr_layernorm_bwd": LayerNormDesc,
    "mr_gelu": ElementwiseDesc,
    "mr_colsum": ColsumDesc,
    "mr_rotary_bwd": RotaryBwdDesc,
    "mr_contrastive_lse_bwd": ContrastiveDesc,
    "mr_unit_normalize_bwd": NormalizeDesc,
    "mr_segment_pool_bwd": SegmentPoolDesc,
    "mr_embed_assemble_bwd": AssembleDesc,
    "mr_window_attention_bwd": WindowDesc,
    "mr_window_mean_bwd": WindowDesc,
    "mr_add": ElementwiseDesc,
    "mr_adamw": AdamWDesc,
    "mr_adamw_multi": AdamWMultiDesc,
    "mr_all_gather": CollectiveDesc,
    "mr_reduce_scatter": CollectiveDesc,
    "mr_all_reduce": CollectiveDesc,
}
WORKSPACE_FNS = {
    "mr_gemm_epilogue_workspace_bytes": GemmDesc,
    "mr_attention_fwd_workspace_bytes": AttentionDesc,
    "mr_attention_bwd_workspace_bytes": AttentionDesc,
    "mr_transformer_encode_workspace_bytes": EncoderDesc,
    "mr_logmel_workspace_bytes": LogmelDesc,
    "mr_resize_patchify_workspace_bytes": ResizePatchifyDesc,
    "mr_vision_encode_workspace_bytes": VisionDesc,
    "mr_audio_encode_workspace_bytes": AudioDesc,
    "mr_span_encode_workspace_bytes": SpanDesc,
    "mr_joint_encode_workspace_bytes": JointDesc,
    "mr_embed_video_batched_workspace_bytes": EmbedVideoDesc,
    "mr_layernorm_bwd_workspace_bytes": LayerNormDesc,
    "mr_colsum_workspace_bytes": ColsumDesc,
    "mr_contrastive_lse_bwd_workspace_bytes": ContrastiveDesc,
    "mr_contrastive_lse_fwd_workspace_bytes": ContrastiveDesc,
}
MISC_FNS = ["mr_check_device", "mr_abi_version", "mr_status_string", "mr_launch_count", "mr_debug_attention_trace",
            "mr_logmel_tables_bytes", "mr_nccl_version", "mr_nccl_unique_id", "mr_nccl_comm_init", "mr_nccl_comm_destroy",
            "mr_nccl_comm_info"]

# MrGemmDesc.algo_pair used when a call does not say: 0 auto, 1 force CTA pairs, 2 never (MR_GEMM_PAIR overrides)
import os as _os
GEMM_PAIR_DEFAULT = int(_os.environ.get("MR_GEMM_PAIR", "0"))

ABI_VERSION = 17  # must equal mr_abi_version() of the loaded library (struct layouts above)

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (once).  Raises MrError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise MrError(f"{_LIB_PATH} is missing: build it with `python -m merlot_reserve_b200.build` "
                      f"(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, desc in ENTRY_POINTS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int32
        fn.argtypes = [C.c_void_p, C.POINTER(desc), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t]
    for name, desc in WORKSPACE_FNS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_size_t
        fn.argtypes = [C.POINTER(desc)]
    lib.mr_check_device.restype = C.c_int32
    lib.mr_check_device.argtypes = [C.c_int32]
    lib.mr_abi_version.restype = C.c_int32
    lib.mr_status_string.restype = C.c_char_p
    lib.mr_status_string.argtypes = [C.c_int32]
    lib.mr_launch_count.restype = C.c_int64
    lib.mr_logmel_tables_bytes.restype = C.c_size_t
    lib.mr_logmel_tables_bytes.argtypes = []
    lib.mr_nccl_version.restype = C.c_int32
    lib.mr_nccl_unique_id.restype = C.c_int32
    lib.mr_nccl_unique_id.argtypes = [C.c_void_p]
    lib.mr_nccl_comm_init.restype = C.c_int32
    lib.mr_nccl_comm_init.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p]
    lib.mr_nccl_comm_destroy.restype = C.c_int32
    lib.mr_nccl_comm_destroy.argtypes = [C.c_void_p]
    lib.mr_nccl_comm_info.restype = C.c_int32
    lib.mr_nccl_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    if lib.mr_abi_version() != ABI_VERSION:
        raise MrError(f"{_LIB_PATH} has ABI version {lib.mr_abi_version()}, this package needs {ABI_VERSION}: "
                      f"rebuild with `python -m merlot_reserve_b200.build --force`")
    _lib = lib
    return lib


def launch_count() -> int:
    return int(load().mr_launch_count())


def require_b200(device: int = 0) -> None:
    if not torch.cuda.is_available():
        raise MrError("no CUDA device: the MERLOT Reserve B200 path has no CPU fallback")
    st = load().mr_check_device(device)
    if st != 0:
        raise MrError(f"device {device}: {load().mr_status_string(st).decode()}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise MrError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise MrError("expected a contiguous tensor")
    return t.data_ptr()


def _call(name: str, desc, ins: Sequence[Optional[torch.Tensor]], outs: Sequence[Optional[torch.Tensor]],
          workspace: Optional[torch.Tensor] = None) -> None:
    lib = load()
    in_arr = (C.c_void_p * len(ins))(*[_ptr(t) for t in ins])
    out_arr = (C.c_void_p * len(outs))(*[_ptr(t) for t in outs])
    stream = torch.cuda.current_stream().cuda_stream
    ws_ptr = _ptr(workspace)
    ws_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    st = getattr(lib, name)(C.c_void_p(stream), C.byref(desc), in_arr, out_arr, C.c_void_p(ws_ptr), ws_bytes)
    if prof is not None:
        e1.record()
        prof.append((name, _work(name, desc), e0, e1))
    if st != 0:
        raise MrError(f"{name} failed ({st}): {lib.mr_status_string(st).decode()}")


# Optional per-launch timing (bench.py): set PROFILE = [] to collect (entry, (tag, flops, bytes), ev0, ev1)
# for every call; CUDA events are recorded on the launching stream around the enqueue.
PROFILE = None


def _work(name: str, d):
    """Algorithmic work of one call: (kernel tag, FLOPs, bytes) -- SURVEY.md §8(d) definitions."""
    if name == "mr_gemm_epilogue":
        tag = "gemm"
        if d.epilogue & EPI_ROTARY_QKV:
            tag = "gemm_qkv_rotary"
        elif d.epilogue & EPI_GELU:
            tag = "gemm_bias_gelu"
        elif d.epilogue & EPI_RESIDUAL:
            tag = "gemm_residual"
        return (tag, 2.0 * d.M * d.N * d.K, 2.0 * (d.M * d.K + d.N * d.K + d.M * d.N))
    if name == "mr_attention_fwd":
        return ("attention", 4.0 * d.B * d.T * d.T * d.H, 8.0 * d.B * d.T * d.H)
    if name == "mr_layernorm_fwd":
        return ("layernorm", 0.0, 4.0 * d.rows * d.H)
    return (name[3:], 0.0, 0.0)


# ------------------------------------------------------------------------------------------------
# typed wrappers
# ------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, residual=None, rot=None, out=None, gelu=False,
         out_f32=False, hidden=0, q_scale=1.0, seq_in=0, seq_out=0, row_off=0, out_rows=None, M=None, lda=None,
         algo=ALGO_TCGEN05, ln_stats=None, ln_colsum=None, ln_eps=1e-5, stats_out=None, pair=None,
         pre_out=None) -> torch.Tensor:
    """D = epilogue(A @ W^T).  a: bf16 [M, K] (or strided rows via M / lda), w: bf16 [N, K].
    ln_stats / ln_colsum: fold LayerNorm(a) into the GEMM (w must be the gamma-scaled weight, bias = beta W + b);
    stats_out: f32 [N/64, M, 2] buffer receiving the row statistics of the (bf16) output."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    if pair is None:
        pair = GEMM_PAIR_DEFAULT
    N, K = w.shape
    if M is None:
        M = a.shape[0]
    if lda is None:
        lda = a.shape[-1]
    rows = out_rows if out_rows is not None else M
    if out is None:
        out = torch.empty((rows, N), dtype=torch.float32 if out_f32 else torch.bfloat16, device=a.device)
    epi = 0
    if bias is not None:
        epi |= EPI_BIAS
    if gelu:
        epi |= EPI_GELU
    if residual is not None:
        epi |= EPI_RESIDUAL
    if rot is not None:
        epi |= EPI_ROTARY_QKV
    if out.dtype == torch.float32:
        epi |= EPI_OUT_F32
    if ln_stats is not None:
        epi |= EPI_LN_IN
    if stats_out is not None:
        epi |= EPI_STATS_OUT
    if pre_out is not None:          # bias + GELU with the pre-activation as a second output (MR_EPI_GELU_DUAL)
        assert gelu and bias is not None and pre_out.shape == out.shape and pre_out.dtype == torch.bfloat16
        epi |= EPI_GELU_DUAL
    d = GemmDesc(M=M, N=N, K=K, lda=lda, ldd=out.shape[-1], epilogue=epi, seq_in=seq_in, seq_out=seq_out,
                 row_off=row_off, rot_rows=(rot.shape[0] if rot is not None else 0), hidden=hidden, algo=algo,
                 q_scale=q_scale, num_sms=0, ln_eps=ln_eps, algo_pair=pair)
    _call("mr_gemm_epilogue", d, [a, w, bias, residual, rot, ln_stats, ln_colsum], [out, stats_out, pre_out])
    return out


def transformer_encode(x: torch.Tensor, rot: torch.Tensor, pre_ln, layer_ptrs: Sequence[torch.Tensor], B: int, T: int,
                       H: int, valid=None, group=None, eps=1e-5) -> torch.Tensor:
    """In-place layer stack (mr_transformer_encode).  layer_ptrs: 8 tensors per layer, in the header's order."""
    n_layers = len(layer_ptrs) // 8
    d = EncoderDesc(B=B, T=T, H=H, num_layers=n_layers, rot_rows=rot.shape[0], ln_eps=eps)
    nbytes = load().mr_transformer_encode_workspace_bytes(C.byref(d))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    _call("mr_transformer_encode", d, [x, rot, valid, group, pre_ln[0], pre_ln[1], *layer_ptrs], [x], ws)
    return x


def encoder_weights(enc) -> EncoderWeights:
    """`MrEncoderWeights` over a packed encoder (modeling._PackedEncoder with the LayerNorms folded in).  The returned
    struct references `enc`'s tensors by address: keep `enc` alive as long as the struct is used."""
    n = len(enc.layer_ptrs)
    arr = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in enc.layer_ptrs])
    w = EncoderWeights(pre_ln_g=enc.pre_ln[0].data_ptr(), pre_ln_b=enc.pre_ln[1].data_ptr(),
                       final_ln_g=enc.final_ln[0].data_ptr(), final_ln_b=enc.final_ln[1].data_ptr(),
                       cls=enc.cls.data_ptr() if enc.add_cls else None,
                       cls_proj_w=enc.cls_proj[0].data_ptr() if enc.add_cls else None,
                       cls_proj_b=enc.cls_proj[1].data_ptr() if enc.add_cls else None,
                       layers=C.cast(arr, C.POINTER(C.c_void_p)), num_layers=n // 8)
    w._keep = arr          # the host pointer array must outlive the struct
    return w


def pool_weights(p) -> PoolWeights:
    return PoolWeights(wq=p.wq.data_ptr(), bq=p.bq.data_ptr(), wkv=p.wkv.data_ptr(), bkv=p.bkv.data_ptr(),
                       wo=p.wo.data_ptr(), bo=p.bo.data_ptr())


_WS_CACHE = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """A per-device scratch buffer that only grows (composite entries: the caller owns the workspace).  It is shared by
    every composite call of this process on that device, i.e. it assumes ONE compute stream per device (the facade's
    convention); callers that run composites on several streams at once pass their own `ws=` to `composite`."""
    key = torch.device(device).index
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf


def composite_workspace_bytes(name: str, desc) -> int:
    return int(getattr(load(), name + "_workspace_bytes")(C.byref(desc)))


def composite(name: str, desc, ins, outs, device, ws=None) -> None:
    """One composite entry (mr_vision_encode / mr_audio_encode / mr_span_encode / mr_joint_encode /
    mr_embed_video_batched).  `ws` = caller-owned workspace (CUDA-graph captures own theirs); default = the growing
    per-device scratch buffer."""
    if ws is None:
        ws = workspace(composite_workspace_bytes(name, desc), device)
    _call(name, desc, ins, outs, ws)


ATTN_AUTO, ATTN_MMA_SYNC, ATTN_TCGEN05 = 0, 1, 2
ATTN_TCGEN05_2TILE, ATTN_TCGEN05_3STREAM = 3, 4    # forward only: force one of the two tensor-memory kernels


def attention(qkv: torch.Tensor, B: int, T: int, H: int, valid=None, group=None, out=None, algo=ATTN_AUTO, lse=None) -> torch.Tensor:
    """lse: optional f32 [B, H/64, T] buffer that receives the row log-sum-exp (needed by attention_bwd)."""
    if out is None:
        out = torch.empty((B * T, H), dtype=torch.bfloat16, device=qkv.device)
    _call("mr_attention_fwd", AttentionDesc(B=B, T=T, H=H, algo=algo, emit_lse=int(lse is not None)), [qkv, valid, group],
          [out, lse])
    return out


ATTN_BWD_EPILOGUE, ATTN_BWD_BIAS_GRAD = 1, 2


def attention_bwd(qkv: torch.Tensor, dctx: torch.Tensor, ctx: torch.Tensor, lse: torch.Tensor, B: int, T: int, H: int,
                  valid=None, group=None, algo=ATTN_AUTO, rot=None, q_scale: Optional[float] = None, bias_grad: bool = False):
    """d(q, k, v) of mr_attention_fwd.  With `q_scale` (and optionally `rot` f32 [rot_rows, 32]) the result is the
    gradient w.r.t. the qkv GEMM's pre-epilogue output (mr_rotary_bwd folded in); with `bias_grad` the call returns
    (dqkv, dbias f32 [3H]) -- the column sums mr_colsum would compute -- from the same pass."""
    finish = (ATTN_BWD_EPILOGUE if q_scale is not None else 0) | (ATTN_BWD_BIAS_GRAD if bias_grad else 0)
    if rot is not None and q_scale is None:
        raise MrError("attention_bwd: rot needs q_scale (the backward of the qkv epilogue is applied as a whole)")
    d = AttentionDesc(B=B, T=T, H=H, algo=algo, finish=finish, rot_rows=0 if rot is None else rot.shape[0],
                      q_scale=q_scale or 0.0)
    dqkv = torch.empty_like(qkv)
    dbias = torch.empty((3 * H,), dtype=torch.float32, device=qkv.device) if bias_grad else None
    _call("mr_attention_bwd", d, [qkv, dctx, ctx, lse, valid, group, rot], [dqkv, dbias],
          _ws("mr_attention_bwd_workspace_bytes", d, qkv.device))
    return (dqkv, dbias) if bias_grad else dqkv


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, rows=None, H=None, ldx=None, seq_in=0,
              seq_out=0, row_off=0, eps=1e-5, out=None, stats_out=None) -> torch.Tensor:
    H = H or x.shape[-1]
    rows = rows if rows is not None else x.numel() // H
    ldx = ldx or H
    if out is None:
        out = torch.empty((rows, H), dtype=torch.bfloat16, device=x.device)
    d = LayerNormDesc(rows=rows, H=H, ldx=ldx, seq_in=seq_in, seq_out=seq_out, row_off=row_off, eps=eps,
                      emit_stats=int(stats_out is not None))
    _call("mr_layernorm_fwd", d, [x, gamma, beta], [out, stats_out])
    return out


def cast_pad(x: torch.Tensor, rows: int, cols: int, ld_out: int, out=None, sanitize: bool = False) -> torch.Tensor:
    assert x.dtype == torch.float32
    if out is None:
        out = torch.empty((rows, ld_out), dtype=torch.bfloat16, device=x.device)
    _call("mr_cast_pad", CastDesc(rows=rows, cols=cols, ld_out=ld_out, sanitize=int(sanitize)), [x], [out])
    return out


def patchify_u8(frames: torch.Tensor, grid_h: int, grid_w: int, out=None) -> torch.Tensor:
    """uint8 NHWC frames [B, grid_h*16, grid_w*16, 3] -> bf16 patches [B*grid_h*grid_w, 768] (preprocess.py:464-481
    without the resize: convert_image_dtype, space_to_depth(16), cast to bf16)."""
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_contiguous()
    B = frames.shape[0]
    if tuple(frames.shape[1:]) != (grid_h * 16, grid_w * 16, 3):
        raise MrError(f"frames must be [B, {grid_h * 16}, {grid_w * 16}, 3] uint8, got {tuple(frames.shape)}")
    if out is None:
        out = torch.empty((B * grid_h * grid_w, 768), dtype=torch.bfloat16, device=frames.device)
    if B:
        _call("mr_patchify_u8", PatchifyDesc(B=B, grid_h=grid_h, grid_w=grid_w, patch=16), [frames], [out])
    return out


_LOGMEL_TABLES = {}


def logmel(waveform: torch.Tensor, pad_size: int = 2, split: bool = True) -> torch.Tensor:
    """make_spectrogram on the device: f32 [B, n_samples] -> [B, 3, 60, 65] (split) or [B, n_frames, 65]."""
    assert waveform.dtype == torch.float32 and waveform.dim() == 2
    B, n = waveform.shape
    dev = waveform.device
    key = dev.index
    build = key not in _LOGMEL_TABLES
    if build:
        _LOGMEL_TABLES[key] = torch.empty((int(load().mr_logmel_tables_bytes()),), dtype=torch.uint8, device=dev)
    d = LogmelDesc(B=B, n_samples=n, playback_speed=1, pad_size=pad_size, split=int(split), build_tables=int(build))
    n_frames = 1 + n // 588
    out = torch.empty((B, 3, 60, 65) if split else (B, n_frames, 65), dtype=torch.float32, device=dev)
    _call("mr_logmel", d, [waveform.contiguous(), _LOGMEL_TABLES[key]], [out], _ws("mr_logmel_workspace_bytes", d, dev))
    return out


def resize_patchify(frames: torch.Tensor, grid_h: int, grid_w: int, out_f32: bool = False) -> torch.Tensor:
    """preprocess_image_to_patches on the device: uint8 [B, h, w, 3] (any size) -> [B, grid_h * grid_w, 768]."""
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3
    B, h, w, _ = frames.shape
    d = ResizePatchifyDesc(B=B, in_h=h, in_w=w, grid_h=grid_h, grid_w=grid_w, out_f32=int(out_f32))
    out = torch.empty((B, grid_h * grid_w, 768), dtype=torch.float32 if out_f32 else torch.bfloat16, device=frames.device)
    _call("mr_resize_patchify", d, [frames.contiguous()], [out], _ws("mr_resize_patchify_workspace_bytes", d, frames.device))
    return out


def fill_cls(x: torch.Tensor, cls: torch.Tensor, B: int, T: int, H: int) -> None:
    _call("mr_fill_cls", FillClsDesc(B=B, T=T, H=H), [cls], [x])


def _window_desc(B, T, H, mode, win, n_win, grid_h=0, grid_w=0):
    return WindowDesc(B=B, T=T, H=H, mode=mode, win=win, n_win=n_win, grid_h=grid_h, grid_w=grid_w)


def window_mean(x, B, T, H, mode, win, n_win, grid_h=0, grid_w=0) -> torch.Tensor:
    out = torch.empty((B * n_win, H), dtype=torch.bfloat16, device=x.device)
    _call("mr_window_mean", _window_desc(B, T, H, mode, win, n_win, grid_h, grid_w), [x], [out])
    return out


def window_attention(q, kv, B, T, H, mode, win, n_win, grid_h=0, grid_w=0) -> torch.Tensor:
    out = torch.empty((B * n_win, H), dtype=torch.bfloat16, device=q.device)
    _call("mr_window_attention", _window_desc(B, T, H, mode, win, n_win, grid_h, grid_w), [q, kv], [out])
    return out


def embed_assemble(tokens, emb, *, H, seg_idx=None, aud_ptr=None, audio=None, vision=None, tok_emb=None, n_aud=0,
                   n_vis=0, pool_h=0, pool_w=0, T=None, audio_token_length=6, cls_offset=0, aud_div=0, vis_group=0,
                   vis_tile=0, out=None):
    """`out` = (x, rot, valid) views to fill (rows of a larger joint batch); allocated when None."""
    B, L = tokens.shape
    T = T if T is not None else L + n_vis
    dev = tokens.device
    rows = B * (T + cls_offset)
    if out is None:
        x = torch.empty((rows, H), dtype=torch.bfloat16, device=dev)
        rot = torch.empty((rows, 32), dtype=torch.float32, device=dev)
        valid = torch.empty((rows,), dtype=torch.uint8, device=dev)
    else:
        x, rot, valid = out
        assert x.numel() == rows * H and rot.numel() == rows * 32 and valid.numel() == rows
    d = AssembleDesc(B=B, L=L, H=H, n_aud=n_aud, n_vis=n_vis, pool_h=pool_h, pool_w=pool_w, T=T,
                     vocab=(emb.shape[0] if emb is not None else 1), audio_token_length=audio_token_length,
                     cls_offset=cls_offset, aud_div=aud_div, vis_group=vis_group, vis_tile=vis_tile)
    _call("mr_embed_assemble", d, [tokens, seg_idx, aud_ptr, emb, audio, vision, tok_emb], [x, rot, valid])
    return x, rot, valid


def rotary_table(coords: torch.Tensor) -> torch.Tensor:
    rows, nd = coords.shape
    out = torch.empty((rows, 32), dtype=torch.float32, device=coords.device)
    _call("mr_rotary_table", RotaryDesc(rows=rows, ndims=nd), [coords], [out])
    return out


def unit_normalize(x: torch.Tensor, scale: float = 1.0, out_f32: bool = True, row_idx=None, out=None, scale_dev=None) -> torch.Tensor:
    """y[r] = x[row_idx[r]] / sqrt(|x[row_idx[r]]|^2 + 1e-5) * scale (row_idx None: identity); `scale_dev`: optional
    f32 [1] device tensor multiplied into `scale` at run time."""
    H = x.shape[-1]
    rows = x.shape[0] if row_idx is None else row_idx.numel()
    if out is None:
        out = torch.empty((rows, H), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    if rows == 0:
        return out
    d = NormalizeDesc(rows=rows, H=H, in_f32=int(x.dtype == torch.float32), out_f32=int(out.dtype == torch.float32),
                      scale=scale, scale_from_device=int(scale_dev is not None))
    _call("mr_unit_normalize", d, [x, row_idx, scale_dev], [out])
    return out


def segment_pool(do_pool: torch.Tensor, idx: torch.Tensor, v: torch.Tensor, S: int, accumulate_into=None):
    """one_hot_pool; `accumulate_into` = (pooled, count) of an earlier call to add to."""
    B, L, H = v.shape
    if accumulate_into is None:
        pooled = torch.empty((B, S, H), dtype=torch.float32, device=v.device)
        count = torch.empty((B, S), dtype=torch.float32, device=v.device)
    else:
        pooled, count = accumulate_into
    d = SegmentPoolDesc(B=B, L=L, H=H, S=S, v_f32=int(v.dtype == torch.float32),
                        accumulate=int(accumulate_into is not None))
    _call("mr_segment_pool", d, [do_pool, idx, v], [pooled, count])
    return pooled, count


def contrastive_lse(x: torch.Tensor, y_all: torch.Tensor, pos_off: int = 0):
    """(lse [n], numer [n]) of one loss direction; row i's positive is row pos_off + i of y_all.  One fused kernel."""
    n, H = x.shape
    d = ContrastiveDesc(n=n, n_all=y_all.shape[0], H=H, pos_off=pos_off)
    lse = torch.empty((n,), dtype=torch.float32, device=x.device)
    numer = torch.empty((n,), dtype=torch.float32, device=x.device)
    _call("mr_contrastive_lse_fwd", d, [x, y_all], [lse, numer])
    return lse, numer


def contrastive_reduce(lse: torch.Tensor, numer: torch.Tensor, sources=None) -> torch.Tensor:
    """[mean(lse - numer), per-source means x3] (pretrain_model.py:295-300)."""
    out = torch.empty((4,), dtype=torch.float32, device=lse.device)
    _call("mr_contrastive_reduce", ContrastiveDesc(n=lse.numel(), n_all=0, H=0), [lse, numer, sources], [out])
    return out


def gemm_bwd(a: torch.Tensor, b: torch.Tensor, *, trans_a=False, trans_b=False, out_f32=True, split_k=0, out=None,
             bias=None, tile_n=0, dgelu_pre=None) -> torch.Tensor:
    """Backward GEMMs on transposed (MN-major) operands, D [M, N] = op(a) . op(b):
       trans_a: a is [K, M] (else [M, K]);  trans_b: b is [K, N] (else [N, K]).
    dgrad: gemm_bwd(dY, W_packed, trans_b=True);  wgrad (packed layout): gemm_bwd(dY, X, trans_a=True, trans_b=True).
    dgelu_pre: bf16 [M, N] pre-activation of my_gelu; the result is D * gelu'(dgelu_pre) (MR_EPI_DGELU, bf16 dgrad).
    tile_n: 0 = the library's rule, 128 / 256 = force the tile width (wgrad plans, `wgrad_plan`)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    K, M = a.shape if trans_a else (a.shape[1], a.shape[0])
    Kb, N = b.shape if trans_b else (b.shape[1], b.shape[0])
    assert K == Kb, (a.shape, b.shape)
    if dgelu_pre is not None:
        assert trans_b and not trans_a and not out_f32 and bias is None and split_k <= 1
        assert dgelu_pre.dtype == torch.bfloat16 and dgelu_pre.is_contiguous() and tuple(dgelu_pre.shape) == (M, N)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if out_f32 else torch.bfloat16, device=a.device)
    epi = (EPI_OUT_F32 if out.dtype == torch.float32 else 0) | (EPI_BIAS if bias is not None else 0)
    if dgelu_pre is not None:
        epi = EPI_DGELU
    d = GemmDesc(M=M, N=N, K=K, lda=a.shape[1], ldd=out.shape[-1], epilogue=epi, algo=ALGO_TCGEN05, q_scale=1.0,
                 trans_a=int(trans_a), trans_b=int(trans_b), ldb=b.shape[1], split_k=split_k, tile_n=tile_n)
    ws = None
    if split_k > 1:
        nbytes = load().mr_gemm_epilogue_workspace_bytes(C.byref(d))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=a.device)
    _call("mr_gemm_epilogue", d, [a, b, bias, dgelu_pre, None, None, None], [out, None], ws)
    return out


def wgrad_plan(out_features: int, in_features: int, rows: int, sms: int = 148) -> Tuple[int, int]:
    """(tile_n, split_k) for dW [out, in] = dY^T . X reduced over `rows` token rows.

    The kernel walks tiles x splits round-robin over one persistent CTA per SM, so the time is
    waves x (k-blocks per part x t_kb + t_epilogue) + the second-stage sum over the fp32 partials.  Constants in us,
    fitted to same-box timings on a B200 (profiles/r01q_probe.jsonl): a 64-row k-block of a 128 x 256 tile costs
    0.40 (1.55 PFLOP/s over 148 SMs); a 128 x 128 tile with both operands MN-major costs 0.32, i.e. 1.6x more per
    FLOP (two 8 KB slab boxes per operand and k-block, shared-memory read limit), so the wide tile wins whenever
    in_features % 256 == 0; the epilogue stores the fp32 tile from registers (4 / 2.5); the reduce streams
    split_k + 1 fp32 copies of dW at ~5 TB/s plus a launch."""
    kb_total = -(-rows // 64)
    m_tiles = -(-out_features // 128)
    best = None
    for bn in (256, 128):
        if bn == 256 and in_features % 256 != 0:
            continue
        tiles = m_tiles * -(-in_features // bn)
        t_kb, t_epi = (0.40, 4.0) if bn == 256 else (0.32, 2.5)
        for s in range(1, 17):
            if s > kb_total:
                break
            waves = -(-(tiles * s) // sms)
            t = waves * (-(-kb_total // s) * t_kb + t_epi)
            if s > 1:
                t += 3.0 + (s + 1) * out_features * in_features * 4 / 5.0e6
            if best is None or t < best[0] - 1e-9:
                best = (t, bn, s)
    return best[1], (best[2] if best[2] > 1 else 0)


def row_dot(x: torch.Tensor, w: torch.Tensor, row_idx=None, rows=None) -> torch.Tensor:
    H = x.shape[-1]
    rows = rows if rows is not None else (row_idx.numel() if row_idx is not None else x.shape[0])
    out = torch.empty((rows,), dtype=torch.float32, device=x.device)
    _call("mr_row_dot", RowDotDesc(rows=rows, H=H), [x, w, row_idx], [out])
    return out


# ------------------------------------------------------------------------------------------------
# backward wrappers
# ------------------------------------------------------------------------------------------------
def _ws(fn_name: str, desc, device) -> torch.Tensor:
    n = getattr(load(), fn_name)(C.byref(desc))
    return torch.empty((max(int(n), 16),), dtype=torch.uint8, device=device)


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, dres=None, eps=1e-5):
    """-> (dx bf16 [rows, H] (+ dres), dgamma f32 [H], dbeta f32 [H])."""
    rows, H = x.shape
    d = LayerNormDesc(rows=rows, H=H, ldx=H, eps=eps)
    dx = torch.empty_like(x)
    dg = torch.empty((H,), dtype=torch.float32, device=x.device)
    db = torch.empty((H,), dtype=torch.float32, device=x.device)
    _call("mr_layernorm_bwd", d, [dy, x, gamma, dres], [dx, dg, db], _ws("mr_layernorm_bwd_workspace_bytes", d, x.device))
    return dx, dg, db


def gelu(x: torch.Tensor, dy=None) -> torch.Tensor:
    """dy None: x sigmoid(1.702 x); else dy * gelu'(x)."""
    out = torch.empty_like(x)
    _call("mr_gelu", ElementwiseDesc(n=x.numel(), mode=int(dy is not None)), [x, dy], [out])
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    rows, cols = x.shape
    d = ColsumDesc(rows=rows, cols=cols, ld=cols)
    out = torch.empty((cols,), dtype=torch.float32, device=x.device)
    _call("mr_colsum", d, [x], [out], _ws("mr_colsum_workspace_bytes", d, x.device))
    return out


def rotary_bwd(dqkv: torch.Tensor, rot: torch.Tensor, hidden: int, q_scale: float = 0.125) -> torch.Tensor:
    _call("mr_rotary_bwd", RotaryBwdDesc(rows=dqkv.shape[0], hidden=hidden, rot_rows=rot.shape[0], q_scale=q_scale),
          [rot], [dqkv])
    return dqkv


def contrastive_lse_bwd(x: torch.Tensor, y_all: torch.Tensor, lse: torch.Tensor, pos_off: int, grad_scale: float,
                        dy_out=None, dy_seq_in: int = 0, dy_seq_out: int = 0, dy_row_off: int = 0):
    """dx [n, H], dy_all.  With `dy_out` (a [world, S, H] fp32 buffer shared by the six directions of a step) the rows
    of dy_all are written rank-major at (rank, dy_row_off + i): the layout one reduce-scatter returns to the owners."""
    n, H = x.shape
    d = ContrastiveDesc(n=n, n_all=y_all.shape[0], H=H, pos_off=pos_off, grad_scale=grad_scale,
                        dy_seq_in=dy_seq_in, dy_seq_out=dy_seq_out, dy_row_off=dy_row_off)
    dx = torch.empty((n, H), dtype=torch.float32, device=x.device)
    dy = torch.empty((y_all.shape[0], H), dtype=torch.float32, device=x.device) if dy_out is None else dy_out
    _call("mr_contrastive_lse_bwd", d, [x, y_all, lse], [dx, dy], _ws("mr_contrastive_lse_bwd_workspace_bytes", d, x.device))
    return dx, dy


def unit_normalize_bwd(x: torch.Tensor, dy: torch.Tensor, scale: float = 1.0, row_idx=None, scale_dev=None) -> torch.Tensor:
    """`scale_dev`: optional f32 [1] device tensor multiplied into `scale` at run time (CUDA-graph-safe temperature)."""
    H = x.shape[-1]
    rows = dy.shape[0]
    dx = torch.zeros((x.shape[0], H), dtype=torch.float32, device=x.device)
    d = NormalizeDesc(rows=rows, H=H, in_f32=int(x.dtype == torch.float32), out_f32=int(dy.dtype == torch.float32), scale=scale,
                      scale_from_device=int(scale_dev is not None))
    _call("mr_unit_normalize_bwd", d, [x, dy, row_idx, scale_dev], [dx])
    return dx


def segment_pool_bwd(do_pool: torch.Tensor, idx: torch.Tensor, dpooled: torch.Tensor, L: int) -> torch.Tensor:
    B, S, H = dpooled.shape
    dv = torch.empty((B, L, H), dtype=torch.bfloat16, device=dpooled.device)
    _call("mr_segment_pool_bwd", SegmentPoolDesc(B=B, L=L, H=H, S=S), [do_pool, idx, dpooled], [dv])
    return dv


def embed_assemble_bwd(tokens, dx, *, H, vocab, d_emb=None, aud_ptr=None, d_audio=None, d_vision=None, n_aud=0, n_vis=0,
                       T=None, audio_token_length=6, cls_offset=0, aud_div=0, vis_group=0, vis_tile=0) -> None:
    """Accumulates (atomics) into d_emb [vocab, H], d_audio, d_vision (all f32, zero-initialised by the caller)."""
    B, L = tokens.shape
    T = T if T is not None else L + n_vis
    d = AssembleDesc(B=B, L=L, H=H, n_aud=n_aud if d_audio is not None else 0, n_vis=n_vis if d_vision is not None else 0,
                     pool_h=1, pool_w=1, T=T, vocab=vocab, audio_token_length=audio_token_length, cls_offset=cls_offset,
                     aud_div=aud_div, vis_group=vis_group, vis_tile=vis_tile)
    if d_vision is None and n_vis:
        d.n_vis = 0
        d.T = T            # rows beyond L are simply skipped
    _call("mr_embed_assemble_bwd", d, [tokens, aud_ptr, dx], [d_emb, d_audio, d_vision])


def window_attention_bwd(q, kv, dctx, B, T, H, mode, win, n_win, grid_h=0, grid_w=0):
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    _call("mr_window_attention_bwd", _window_desc(B, T, H, mode, win, n_win, grid_h, grid_w), [q, kv, dctx], [dq, dkv])
    return dq, dkv


def window_mean_bwd(dq_in, dx, B, T, H, mode, win, n_win, grid_h=0, grid_w=0) -> None:
    _call("mr_window_mean_bwd", _window_desc(B, T, H, mode, win, n_win, grid_h, grid_w), [dq_in], [dx])


def add(a: torch.Tensor, b: torch.Tensor, out=None) -> torch.Tensor:
    """bf16 a + (bf16 | f32) b -> bf16."""
    out = torch.empty_like(a) if out is None else out
    _call("mr_add", ElementwiseDesc(n=a.numel(), mode=int(b.dtype == torch.float32)), [a, b], [out])
    return out


def adamw(g, p, m, v, w=None, *, lr, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.0, grad_scale=1.0) -> None:
    d = AdamWDesc(n=p.numel(), lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale)
    _call("mr_adamw", d, [g], [p, m, v, w])


def adamw_chunk_table(slots, device) -> torch.Tensor:
    """Chunk table of mr_adamw_multi as a device uint8 tensor.  `slots` = [(arena_offset, numel, working-copy tensor
    or None, decay?)], arena offsets multiples of 8 elements; the working copy must be contiguous (bf16 or fp32)."""
    import numpy as np
    rows = []
    for off, n, w, decay in slots:
        assert off % 8 == 0 and (w is None or (w.is_contiguous() and w.numel() == n))
        flags = (ADAMW_DECAY if decay else 0) | (ADAMW_W_F32 if (w is not None and w.dtype == torch.float32) else 0)
        esz = 0 if w is None else w.element_size()
        for c0 in range(0, n, ADAMW_CHUNK):
            rows.append((off + c0, 0 if w is None else w.data_ptr() + c0 * esz, min(ADAMW_CHUNK, n - c0), flags))
    arr = np.array(rows, dtype=np.dtype([("offset", "<i8"), ("w", "<u8"), ("count", "<i4"), ("flags", "<i4")]))
    assert arr.dtype.itemsize == C.sizeof(AdamWChunk)
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(device)


def adamw_multi(g, table, p, m, v, *, lr, beta1=0.9, beta2=0.98, eps=1e-8, weight_decay=0.0, grad_scale=1.0,
                bias_corr1=1.0, bias_corr2=1.0, hyper_dev=None) -> None:
    """One launch for all parameters (include/mreserve_b200.h: mr_adamw_multi).  Moment dtype selects the storage:
    bf16 = the reference's compressed moments (pretrain/optimization.py:36-51), fp32 = plain Adam."""
    assert m.dtype == v.dtype and m.dtype in (torch.bfloat16, torch.float32) and p.dtype == torch.float32
    assert g.dtype in (torch.bfloat16, torch.float32), g.dtype
    d = AdamWMultiDesc(n_chunks=table.numel() // C.sizeof(AdamWChunk), moments_bf16=int(m.dtype == torch.bfloat16),
                       grad_bf16=int(g.dtype == torch.bfloat16), lr=lr, beta1=beta1, beta2=beta2, eps=eps,
                       weight_decay=weight_decay, grad_scale=grad_scale, bias_corr1=bias_corr1, bias_corr2=bias_corr2,
                       one_minus_beta1=1.0 - beta1, one_minus_beta2=1.0 - beta2, hyper_from_device=int(hyper_dev is not None))
    _call("mr_adamw_multi", d, [g, table, hyper_dev], [p, m, v])


# ------------------------------------------------------------------------------------------------
# collective entries (csrc/collectives.cu): ncclComm_t by handle, enqueued on the current stream
# ------------------------------------------------------------------------------------------------
_DT = {torch.bfloat16: 0, torch.float32: 1}


def _check(st: int, what: str) -> None:
    if st != 0:
        raise MrError(f"{what} failed ({st}): {load().mr_status_string(st).decode()}")


def nccl_version() -> int:
    """ncclGetVersion() of the libnccl the library resolved (0 when none could be loaded)."""
    return int(load().mr_nccl_version())


class Communicator:
    """An ncclComm_t owned by this process and used only through the library's collective entries.  Created
    collectively by `Communicator.from_group(group)`: rank 0 draws the unique id, the torch.distributed group (any
    backend) carries its 128 bytes to the other ranks, every rank calls ncclCommInitRank.  `Communicator(handle, ...)`
    wraps a communicator made elsewhere (e.g. ProcessGroupNCCL._comm_ptr())."""

    def __init__(self, handle: int, world: int, rank: int, owned: bool = False):
        self.handle, self.world, self.rank, self._owned = handle, world, rank, owned

    @classmethod
    def from_group(cls, group=None) -> "Communicator":
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        lib = load()
        ident = C.create_string_buffer(128)
        if rank == 0:
            _check(lib.mr_nccl_unique_id(ident), "mr_nccl_unique_id")
        box = [bytes(ident.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = C.create_string_buffer(box[0], 128)
        comm = C.c_void_p()
        _check(lib.mr_nccl_comm_init(C.byref(comm), world, rank, ident), "mr_nccl_comm_init")
        w, r = C.c_int32(), C.c_int32()
        _check(lib.mr_nccl_comm_info(comm, C.byref(w), C.byref(r)), "mr_nccl_comm_info")
        assert (w.value, r.value) == (world, rank)
        return cls(comm.value, world, rank, owned=True)

    def destroy(self) -> None:
        """ncclCommDestroy.  Release every CUDA graph that captured this communicator's collectives first: NCCL keeps
        the communicator alive (and this call waiting) while a captured graph still references it."""
        if self._owned and self.handle:
            _check(load().mr_nccl_comm_destroy(C.c_void_p(self.handle)), "mr_nccl_comm_destroy")
        self.handle = None

    def _desc(self, t: torch.Tensor, count: int, average: bool) -> CollectiveDesc:
        if t.dtype not in _DT:
            raise MrError(f"collectives move bf16 or f32, not {t.dtype}")
        if self.handle is None:
            raise MrError("communicator was destroyed")
        return CollectiveDesc(comm=self.handle, count=count, dtype=_DT[t.dtype], average=int(average))

    def all_gather(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None) -> torch.Tensor:
        """recv [world, *send.shape]: rank-major, i.e. jax.lax.all_gather(send, 'batch') (pretrain_model.py:290)."""
        if recv is None:
            recv = torch.empty((self.world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        assert recv.numel() == self.world * send.numel() and recv.dtype == send.dtype
        _call("mr_all_gather", self._desc(send, send.numel(), False), [send], [recv])
        return recv

    def reduce_scatter(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None, average: bool = False) -> torch.Tensor:
        """send [world, ...] rank-major; recv [...] = sum over ranks of block `rank`."""
        if recv is None:
            recv = torch.empty(tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        assert send.numel() == self.world * recv.numel() and recv.dtype == send.dtype
        _call("mr_reduce_scatter", self._desc(send, recv.numel(), average), [send], [recv])
        return recv

    def all_reduce(self, buf: torch.Tensor, average: bool = False) -> torch.Tensor:
        """In place sum (or mean = jax.lax.pmean, pretrain_model.py:329,333) over the ranks."""
        _call("mr_all_reduce", self._desc(buf, buf.numel(), average), [buf], [buf])
        return buf
