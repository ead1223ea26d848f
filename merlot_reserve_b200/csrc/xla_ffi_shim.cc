// XLA-FFI handlers over the C ABI of include/mreserve_b200.h, so the reference's Flax modules can call the sm_100a
// kernels through jax.ffi.ffi_call (INTEGRATION.md shows the Python side).
//
// Every compute entry of the library has ONE signature,
//     int32_t fn(stream, const Desc*, const void* const* in, void* const* out, void* workspace, size_t workspace_bytes)
// so ONE handler template wraps all of them: operands -> in[], results -> out[], the descriptor travels as the
// attribute `desc` (the struct's bytes viewed as int32: np.frombuffer(bytes(desc), np.int32); descriptors that
// carry device pointers -- the packed weight blobs of the composites, the ncclComm_t of the collectives -- are
// built once outside jit, where those addresses are known), and when the attribute `workspace` is 1 the LAST
// result is the scratch buffer (sized by the entry's *_workspace_bytes function).  A zero-element operand or
// result stands for an optional pointer that is NULL.
//
// STATUS: needs jaxlib's header tree (xla/ffi/api/ffi.h).  This image has no jax / jaxlib (no wheels, no network), so
// the file compiles to nothing here and has never been built or run -- it is the documented shape of the binding,
// written against the XLA FFI API as published with jax 0.4.31+.  Where jaxlib exists:
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       xla_ffi_shim.cc -L../lib -lmreserve_b200 -o libmreserve_b200_xla.so
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define MR_HAVE_XLA_FFI 1
#endif
#endif

#ifdef MR_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mreserve_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

template <class Desc, int32_t (*Fn)(mr_stream_t, const Desc*, const void* const*, void* const*, void*, size_t)>
ffi::Error Call(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, ffi::Span<const int32_t> desc,
                int32_t workspace, const char* name) {
  if (desc.size() * sizeof(int32_t) != sizeof(Desc))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, std::string(name) + ": `desc` has the wrong size");
  alignas(8) Desc d;
  memcpy(&d, desc.begin(), sizeof(Desc));
  std::vector<const void*> in(args.size() + 1, nullptr);
  for (size_t i = 0; i < args.size(); ++i) {
    auto b = args.get<ffi::AnyBuffer>(i);
    if (!b.has_value()) return b.error();
    in[i] = b->element_count() ? b->untyped_data() : nullptr;
  }
  size_t n_out = rets.size() - (workspace ? 1 : 0);
  std::vector<void*> out(n_out + 1, nullptr);
  for (size_t i = 0; i < n_out; ++i) {
    auto b = rets.get<ffi::AnyBuffer>(i);
    if (!b.has_value()) return b.error();
    out[i] = (*b)->element_count() ? (*b)->untyped_data() : nullptr;
  }
  void* ws = nullptr;
  size_t ws_bytes = 0;
  if (workspace) {
    auto b = rets.get<ffi::AnyBuffer>(n_out);
    if (!b.has_value()) return b.error();
    ws = (*b)->untyped_data();
    ws_bytes = (*b)->size_bytes();
  }
  int32_t st = Fn(stream, &d, in.data(), out.data(), ws, ws_bytes);
  if (st == MR_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(name) + ": " + mr_status_string(st));
}

}  // namespace

#define MR_FFI(Symbol, Desc, fn)                                                                              \
  static ffi::Error Symbol##Impl(cudaStream_t s, ffi::RemainingArgs a, ffi::RemainingRets r,                  \
                                 ffi::Span<const int32_t> desc, int32_t workspace) {                          \
    return Call<Desc, fn>(s, a, r, desc, workspace, #fn);                                                     \
  }                                                                                                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(Symbol, Symbol##Impl,                                                         \
                                ffi::Ffi::Bind()                                                              \
                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()                                 \
                                    .RemainingArgs()                                                          \
                                    .RemainingRets()                                                          \
                                    .Attr<ffi::Span<const int32_t>>("desc")                                   \
                                    .Attr<int32_t>("workspace"))

// the hot path (mreserve/modeling.py:188-280, 283-376) and its composites (:379-504, :806-846)
MR_FFI(MrGemmEpilogue, MrGemmDesc, mr_gemm_epilogue);
MR_FFI(MrAttentionFwd, MrAttentionDesc, mr_attention_fwd);
MR_FFI(MrLayerNormFwd, MrLayerNormDesc, mr_layernorm_fwd);
MR_FFI(MrTransformerEncode, MrEncoderDesc, mr_transformer_encode);
MR_FFI(MrVisionEncode, MrVisionDesc, mr_vision_encode);
MR_FFI(MrAudioEncode, MrAudioDesc, mr_audio_encode);
MR_FFI(MrSpanEncode, MrSpanDesc, mr_span_encode);
MR_FFI(MrJointEncode, MrJointDesc, mr_joint_encode);
MR_FFI(MrEmbedVideoBatched, MrEmbedVideoDesc, mr_embed_video_batched);
// input assembly, pools, heads (modeling.py:311-326, 419-430, 541-578, 651-758)
MR_FFI(MrCastPad, MrCastDesc, mr_cast_pad);
MR_FFI(MrFillCls, MrFillClsDesc, mr_fill_cls);
MR_FFI(MrWindowMean, MrWindowDesc, mr_window_mean);
MR_FFI(MrWindowAttention, MrWindowDesc, mr_window_attention);
MR_FFI(MrEmbedAssemble, MrAssembleDesc, mr_embed_assemble);
MR_FFI(MrRotaryTable, MrRotaryDesc, mr_rotary_table);
MR_FFI(MrUnitNormalize, MrNormalizeDesc, mr_unit_normalize);
MR_FFI(MrSegmentPool, MrSegmentPoolDesc, mr_segment_pool);
MR_FFI(MrRowDot, MrRowDotDesc, mr_row_dot);
// preprocessing (mreserve/preprocess.py:121-156, 464-481; pretrain/data_utils.py:42-147)
MR_FFI(MrPatchifyU8, MrPatchifyDesc, mr_patchify_u8);
MR_FFI(MrLogmel, MrLogmelDesc, mr_logmel);
MR_FFI(MrResizePatchify, MrResizePatchifyDesc, mr_resize_patchify);
// loss (pretrain/pretrain_model.py:262-303) and the backward / optimizer entries (:306-340)
MR_FFI(MrContrastiveLseFwd, MrContrastiveDesc, mr_contrastive_lse_fwd);
MR_FFI(MrContrastiveReduce, MrContrastiveDesc, mr_contrastive_reduce);
MR_FFI(MrContrastiveLseBwd, MrContrastiveDesc, mr_contrastive_lse_bwd);
MR_FFI(MrAttentionBwd, MrAttentionDesc, mr_attention_bwd);
MR_FFI(MrLayerNormBwd, MrLayerNormDesc, mr_layernorm_bwd);
MR_FFI(MrGelu, MrElementwiseDesc, mr_gelu);
MR_FFI(MrAdd, MrElementwiseDesc, mr_add);
MR_FFI(MrColsum, MrColsumDesc, mr_colsum);
MR_FFI(MrRotaryBwd, MrRotaryBwdDesc, mr_rotary_bwd);
MR_FFI(MrUnitNormalizeBwd, MrNormalizeDesc, mr_unit_normalize_bwd);
MR_FFI(MrSegmentPoolBwd, MrSegmentPoolDesc, mr_segment_pool_bwd);
MR_FFI(MrEmbedAssembleBwd, MrAssembleDesc, mr_embed_assemble_bwd);
MR_FFI(MrWindowAttentionBwd, MrWindowDesc, mr_window_attention_bwd);
MR_FFI(MrWindowMeanBwd, MrWindowDesc, mr_window_mean_bwd);
MR_FFI(MrAdamW, MrAdamWDesc, mr_adamw);
MR_FFI(MrAdamWMulti, MrAdamWMultiDesc, mr_adamw_multi);
// collectives (pretrain_model.py:288-290, 329, 333): the descriptor carries the ncclComm_t
MR_FFI(MrAllGather, MrCollectiveDesc, mr_all_gather);
MR_FFI(MrReduceScatter, MrCollectiveDesc, mr_reduce_scatter);
MR_FFI(MrAllReduce, MrCollectiveDesc, mr_all_reduce);
#endif  // MR_HAVE_XLA_FFI
