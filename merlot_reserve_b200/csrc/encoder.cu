// mr_transformer_encode: the layer loop of TransformerEncoder (modeling.py:283-376) as ONE C-ABI call.
//
// pre_ln, then num_layers x { qkv GEMM (LayerNorm folded in, rotary scale + 1/sqrt(d) in the epilogue) -> fused
// attention -> attn_proj GEMM (+ residual, emits the row statistics of the next LayerNorm) -> MLP-up GEMM
// (LayerNorm folded in, sigmoid-GELU) -> MLP-down GEMM (+ residual, emits statistics) }.  Host code only: every
// step enqueues one of the kernels behind mr_layernorm_fwd / mr_gemm_epilogue / mr_attention_fwd on the caller's
// stream; nothing is allocated and nothing synchronises.  This is the unit a jax.ffi custom call binds per encoder
// (INTEGRATION.md): the four encoders of the model differ only in B, T, the rotary table and the masks.
#include "common.cuh"

namespace {

inline size_t align256(size_t n) { return (n + 255) / 256 * 256; }

struct Carve {
  size_t qkv, ctx, u, stats, total;
};

Carve carve(const MrEncoderDesc* d) {
  const size_t M = static_cast<size_t>(d->B) * d->T, H = d->H;
  Carve c;
  c.qkv = 0;
  c.ctx = c.qkv + align256(M * 3 * H * 2);
  c.u = c.ctx + align256(M * H * 2);
  c.stats = c.u + align256(M * 4 * H * 2);
  c.total = c.stats + align256((H / 64) * M * 2 * sizeof(float));
  return c;
}

}  // namespace

extern "C" size_t mr_transformer_encode_workspace_bytes(const MrEncoderDesc* d) {
  if (!d || d->B <= 0 || d->T <= 0 || d->H <= 0) return 0;
  return carve(d).total;
}

extern "C" int32_t mr_transformer_encode(mr_stream_t stream, const MrEncoderDesc* d, const void* const* in,
                                         void* const* out, void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->T > 0 && d->H > 0 && d->H % 64 == 0 && d->num_layers >= 0 && d->rot_rows > 0,
             MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  const int64_t M64 = static_cast<int64_t>(d->B) * d->T;
  MR_REQUIRE(M64 < (int64_t(1) << 31), MR_ERR_BAD_SHAPE);
  const int M = static_cast<int>(M64), H = d->H;
  const void* x_in = in[0];
  const void* rot = in[1];
  const void* valid = in[2];
  const void* group = in[3];
  void* x = out[0];
  MR_REQUIRE(x_in && rot && in[4] && in[5] && x, MR_ERR_NULL_POINTER);
  const Carve c = carve(d);
  MR_REQUIRE(workspace && workspace_bytes >= c.total, MR_ERR_WORKSPACE);
  MR_REQUIRE(mr::aligned16(workspace), MR_ERR_MISALIGNED);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* qkv = ws + c.qkv;
  void* ctx = ws + c.ctx;
  void* u = ws + c.u;
  void* stats = ws + c.stats;

  // pre_ln (modeling.py:360): x = LN(x_in); also emits the statistics layer 0's folded LayerNorm needs
  MrLayerNormDesc ln{};
  ln.rows = M;
  ln.H = H;
  ln.ldx = H;
  ln.eps = d->ln_eps;
  ln.emit_stats = 1;
  {
    const void* lin[3] = {x_in, in[4], in[5]};
    void* lout[2] = {x, stats};
    const int32_t st = mr_layernorm_fwd(stream, &ln, lin, lout, nullptr, 0);
    if (st != MR_OK) return st;
  }

  MrGemmDesc g{};
  g.M = M;
  g.ln_eps = d->ln_eps;
  MrAttentionDesc a{};
  a.B = d->B;
  a.T = d->T;
  a.H = H;
  for (int l = 0; l < d->num_layers; ++l) {
    const void* const* w = in + 6 + 8 * l;   // wqkv', bqkv', sqkv, wo, w1', b1', s1, w2
    for (int i = 0; i < 8; ++i) MR_REQUIRE(w[i], MR_ERR_NULL_POINTER);
    int32_t st;
    {  // qkv = rotary(LN(x) Wqkv + b)            modeling.py:228,116-144,272
      g.N = 3 * H;
      g.K = H;
      g.lda = H;
      g.ldd = 3 * H;
      g.epilogue = MR_EPI_BIAS | MR_EPI_ROTARY_QKV | MR_EPI_LN_IN;
      g.rot_rows = d->rot_rows;
      g.hidden = H;
      g.q_scale = 0.125f;
      const void* gin[7] = {x, w[0], w[1], nullptr, rot, stats, w[2]};
      void* gout[2] = {qkv, nullptr};
      if ((st = mr_gemm_epilogue(stream, &g, gin, gout, nullptr, 0)) != MR_OK) return st;
    }
    {  // ctx = softmax(q k^T + mask) v             modeling.py:188-200
      const void* ain[3] = {qkv, valid, group};
      void* aout[2] = {ctx, nullptr};
      if ((st = mr_attention_fwd(stream, &a, ain, aout, nullptr, 0)) != MR_OK) return st;
    }
    {  // x += ctx Wo                                modeling.py:235,275
      g.N = H;
      g.K = H;
      g.lda = H;
      g.ldd = H;
      g.epilogue = MR_EPI_RESIDUAL | MR_EPI_STATS_OUT;
      g.rot_rows = 0;
      const void* gin[7] = {ctx, w[3], nullptr, x, nullptr, nullptr, nullptr};
      void* gout[2] = {x, stats};
      if ((st = mr_gemm_epilogue(stream, &g, gin, gout, nullptr, 0)) != MR_OK) return st;
    }
    {  // u = gelu(LN(x) W1 + b1)                    modeling.py:252-254,277
      g.N = 4 * H;
      g.K = H;
      g.lda = H;
      g.ldd = 4 * H;
      g.epilogue = MR_EPI_BIAS | MR_EPI_GELU | MR_EPI_LN_IN;
      const void* gin[7] = {x, w[4], w[5], nullptr, nullptr, stats, w[6]};
      void* gout[2] = {u, nullptr};
      if ((st = mr_gemm_epilogue(stream, &g, gin, gout, nullptr, 0)) != MR_OK) return st;
    }
    {  // x += u W2                                  modeling.py:255,279
      g.N = H;
      g.K = 4 * H;
      g.lda = 4 * H;
      g.ldd = H;
      g.epilogue = MR_EPI_RESIDUAL | MR_EPI_STATS_OUT;
      const void* gin[7] = {u, w[7], nullptr, x, nullptr, nullptr, nullptr};
      void* gout[2] = {x, stats};
      if ((st = mr_gemm_epilogue(stream, &g, gin, gout, nullptr, 0)) != MR_OK) return st;
    }
  }
  return MR_OK;
}
