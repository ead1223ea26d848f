// XLA-FFI handlers over the C ABI of include/mreserve_b200.h, so the reference's Flax modules can call the
// sm_100a kernels through jax.ffi.ffi_call.
//
// This translation unit needs jaxlib's header tree (xla/ffi/api/ffi.h).  The build image has no jax/jaxlib
// (no wheels, no network), so here it compiles to nothing and has NEVER been built or run; it is the
// documented shape of the binding, kept next to the entries it wraps.  Build where jaxlib exists with:
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

#include "../../include/mreserve_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error to_error(int32_t st, const char* what) {
  if (st == MR_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + mr_status_string(st));
}

// D = epilogue(A . W^T): flax nn.Dense / nn.DenseGeneral (mreserve/modeling.py:228,235,252,255,402,453).
// Optional operands (bias / residual / rot) are passed as zero-element buffers when unused.
ffi::Error GemmImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> a, ffi::Buffer<ffi::BF16> w,
                    ffi::Buffer<ffi::F32> bias, ffi::Buffer<ffi::BF16> residual, ffi::Buffer<ffi::F32> rot,
                    ffi::ResultBuffer<ffi::BF16> d, int32_t epilogue, int32_t hidden, float q_scale,
                    int32_t seq_in, int32_t seq_out, int32_t row_off) {
  const auto ad = a.dimensions();
  const auto wd = w.dimensions();
  MrGemmDesc g{};
  g.K = static_cast<int32_t>(ad.back());
  g.M = static_cast<int32_t>(a.element_count() / g.K);
  g.N = static_cast<int32_t>(wd.front());
  g.lda = g.K;
  g.ldd = g.N;
  g.epilogue = epilogue;
  g.hidden = hidden;
  g.q_scale = q_scale;
  g.seq_in = seq_in;
  g.seq_out = seq_out;
  g.row_off = row_off;
  g.rot_rows = rot.element_count() ? static_cast<int32_t>(rot.element_count() / 32) : 0;
  const void* in[5] = {a.untyped_data(), w.untyped_data(),
                       bias.element_count() ? bias.untyped_data() : nullptr,
                       residual.element_count() ? residual.untyped_data() : nullptr,
                       rot.element_count() ? rot.untyped_data() : nullptr};
  void* out[1] = {d->untyped_data()};
  return to_error(mr_gemm_epilogue(stream, &g, in, out, nullptr, 0), "mr_gemm_epilogue");
}

// softmax(q k^T + mask) v: apply_attention (mreserve/modeling.py:188-200) with validity / video_src masks.
ffi::Error AttentionImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> qkv, ffi::Buffer<ffi::U8> valid,
                         ffi::Buffer<ffi::S32> group, ffi::ResultBuffer<ffi::BF16> ctx, int32_t B, int32_t T,
                         int32_t H) {
  MrAttentionDesc d{B, T, H, MR_ATTN_ALGO_AUTO};
  const void* in[3] = {qkv.untyped_data(), valid.element_count() ? valid.untyped_data() : nullptr,
                       group.element_count() ? group.untyped_data() : nullptr};
  void* out[2] = {ctx->untyped_data(), nullptr};
  return to_error(mr_attention_fwd(stream, &d, in, out, nullptr, 0), "mr_attention_fwd");
}

// flax LayerNorm (mreserve/modeling.py:272,277,360,366).
ffi::Error LayerNormImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::F32> gamma,
                         ffi::Buffer<ffi::F32> beta, ffi::ResultBuffer<ffi::BF16> y, float eps) {
  MrLayerNormDesc d{};
  d.H = static_cast<int32_t>(x.dimensions().back());
  d.rows = static_cast<int32_t>(x.element_count() / d.H);
  d.ldx = d.H;
  d.eps = eps;
  const void* in[3] = {x.untyped_data(), gamma.untyped_data(), beta.untyped_data()};
  void* out[1] = {y->untyped_data()};
  return to_error(mr_layernorm_fwd(stream, &d, in, out, nullptr, 0), "mr_layernorm_fwd");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(MrGemmEpilogue, GemmImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("epilogue")
                                  .Attr<int32_t>("hidden")
                                  .Attr<float>("q_scale")
                                  .Attr<int32_t>("seq_in")
                                  .Attr<int32_t>("seq_out")
                                  .Attr<int32_t>("row_off"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(MrAttentionFwd, AttentionImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("B")
                                  .Attr<int32_t>("T")
                                  .Attr<int32_t>("H"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(MrLayerNormFwd, LayerNormImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<float>("eps"));
#endif  // MR_HAVE_XLA_FFI
