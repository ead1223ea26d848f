// Composite entries: one C-ABI call per sub-module of the reference model and one for the whole
// zero-shot forward.  Host code only -- each call enqueues the kernels behind the per-op entries on the caller's
// stream, carves its temporaries out of the caller's workspace, allocates nothing and never synchronises, so a call
// (or a sequence of calls) can be captured into a CUDA graph as is.
//
//   mr_vision_encode        VisionTransformer.__call__            mreserve/modeling.py:391-430
//   mr_audio_encode         AudioTransformer.__call__             mreserve/modeling.py:443-476
//   mr_span_encode          TokenEmbedder + SpanTransformer       mreserve/modeling.py:488-504,516-538 (+ unit_normalize :767-774)
//   mr_joint_encode         prepare_multimodal_inputs + joint_transformer + head + unit_normalize   :651-758,:841-842
//   mr_embed_video_batched  batch_embed_video (vmap of embed_video)                                  :806-846
#include "common.cuh"

namespace {

using mr::aligned16;

inline size_t align256(size_t n) { return (n + 255) / 256 * 256; }

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* p) : base(static_cast<uint8_t*>(p)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  }
};

constexpr int kAudioKPad = 192;   // 2 x 65 = 130 input features, zero padded to a multiple of the 64-wide K block

#define MR_TRY(expr)                        \
  do {                                      \
    const int32_t st__ = (expr);            \
    if (st__ != MR_OK) return st__;         \
  } while (0)

int32_t gemm(mr_stream_t s, const void* a, const void* w, const void* bias, void* out, int M, int N, int K, int lda,
             int ldd, bool out_f32, int seq_in = 0, int seq_out = 0, int row_off = 0) {
  MrGemmDesc g{};
  g.M = M;
  g.N = N;
  g.K = K;
  g.lda = lda;
  g.ldd = ldd;
  g.epilogue = (bias ? MR_EPI_BIAS : 0) | (out_f32 ? MR_EPI_OUT_F32 : 0);
  g.seq_in = seq_in;
  g.seq_out = seq_out;
  g.row_off = row_off;
  const void* in[7] = {a, w, bias, nullptr, nullptr, nullptr, nullptr};
  void* o[2] = {out, nullptr};
  return mr_gemm_epilogue(s, &g, in, o, nullptr, 0);
}

bool enc_ok(const MrEncoderWeights& e, bool cls) {
  if (!e.pre_ln_g || !e.pre_ln_b || !e.final_ln_g || !e.final_ln_b || e.num_layers < 0) return false;
  if (e.num_layers > 0 && !e.layers) return false;
  if (cls && (!e.cls || !e.cls_proj_w || !e.cls_proj_b)) return false;
  return true;
}

// pre_ln + layer stack, in place on x [B*T, H]
int32_t run_stack(mr_stream_t s, const MrEncoderWeights& e, void* x, const void* rot, int rot_rows, const void* valid,
                  const void* group, int B, int T, int H, float eps, void* ws, size_t ws_bytes) {
  MrEncoderDesc d{};
  d.B = B;
  d.T = T;
  d.H = H;
  d.num_layers = e.num_layers;
  d.rot_rows = rot_rows;
  d.ln_eps = eps;
  const int n_in = 6 + 8 * e.num_layers;
  const void* in[6 + 8 * 64];
  if (e.num_layers > 64) return MR_ERR_BAD_SHAPE;
  in[0] = x;
  in[1] = rot;
  in[2] = valid;
  in[3] = group;
  in[4] = e.pre_ln_g;
  in[5] = e.pre_ln_b;
  for (int i = 0; i < 8 * e.num_layers; ++i) in[6 + i] = e.layers[i];
  (void)n_in;
  void* out[1] = {x};
  return mr_transformer_encode(s, &d, in, out, ws, ws_bytes);
}

int32_t layernorm(mr_stream_t s, const void* x, const void* g, const void* b, void* y, int rows, int H, float eps,
                  int seq_in = 0, int seq_out = 0, int row_off = 0) {
  MrLayerNormDesc d{};
  d.rows = rows;
  d.H = H;
  d.ldx = H;
  d.seq_in = seq_in;
  d.seq_out = seq_out;
  d.row_off = row_off;
  d.eps = eps;
  const void* in[3] = {x, g, b};
  void* out[2] = {y, nullptr};
  return mr_layernorm_fwd(s, &d, in, out, nullptr, 0);
}

// flax MultiHeadDotProductAttention with one mean query per window (modeling.py:413-429, :464-475)
struct PoolTmp {
  void *q_in, *q, *kv, *ctx;
};
int32_t attn_pool(mr_stream_t s, const MrPoolWeights& p, const void* xf, void* out, const PoolTmp& t, int B, int T, int H,
                  int mode, int win, int n_win, int gh, int gw) {
  MrWindowDesc w{};
  w.B = B;
  w.T = T;
  w.H = H;
  w.mode = mode;
  w.win = win;
  w.n_win = n_win;
  w.grid_h = gh;
  w.grid_w = gw;
  {
    const void* in[1] = {xf};
    void* o[1] = {t.q_in};
    MR_TRY(mr_window_mean(s, &w, in, o, nullptr, 0));
  }
  MR_TRY(gemm(s, t.q_in, p.wq, p.bq, t.q, B * n_win, H, H, H, H, false));
  MR_TRY(gemm(s, xf, p.wkv, p.bkv, t.kv, B * T, 2 * H, H, H, 2 * H, false));
  {
    const void* in[2] = {t.q, t.kv};
    void* o[1] = {t.ctx};
    MR_TRY(mr_window_attention(s, &w, in, o, nullptr, 0));
  }
  return gemm(s, t.ctx, p.wo, p.bo, out, B * n_win, H, H, H, H, false);
}

bool pool_ok(const MrPoolWeights& p) { return p.wq && p.bq && p.wkv && p.bkv && p.wo && p.bo; }

size_t stack_ws(int B, int T, int H) {
  MrEncoderDesc d{};
  d.B = B;
  d.T = T;
  d.H = H;
  return mr_transformer_encode_workspace_bytes(&d);
}

// ---------------------------------------------------------------------------------------------------------------
// vision / audio share one body: embed GEMM into rows 1.. of [B, 1 + n, H], CLS fill, stack, final LN, pool, cls_proj
// ---------------------------------------------------------------------------------------------------------------
struct TowerShape {
  int B, n, K, Kpad, H, mode, win, n_win, gh, gw;
};

size_t tower_carve(const TowerShape& t, bool in_bf16, bool need_xf, Bump& b, void** a, void** x, void** xf, PoolTmp* pt,
                   void** stack, size_t* stack_bytes) {
  const size_t M = static_cast<size_t>(t.B) * (t.n + 1), rows_in = static_cast<size_t>(t.B) * t.n;
  *a = in_bf16 ? nullptr : b.take(rows_in * t.Kpad * 2);
  *x = b.take(M * t.H * 2);
  *xf = need_xf ? b.take(M * t.H * 2) : nullptr;
  pt->q_in = b.take(static_cast<size_t>(t.B) * t.n_win * t.H * 2);
  pt->q = b.take(static_cast<size_t>(t.B) * t.n_win * t.H * 2);
  pt->kv = b.take(M * 2 * t.H * 2);
  pt->ctx = b.take(static_cast<size_t>(t.B) * t.n_win * t.H * 2);
  *stack_bytes = stack_ws(t.B, t.n + 1, t.H);
  *stack = b.take(*stack_bytes);
  return b.off;
}

int32_t tower_run(mr_stream_t s, const TowerShape& t, const MrEncoderWeights& enc, const MrPoolWeights& pool,
                  const void* emb_w, const void* emb_b, const void* rot, float eps, const void* input, bool in_bf16,
                  void* out_pool, void* out_cls, void* out_xf, void* ws, size_t ws_bytes) {
  const int T = t.n + 1, M = t.B * T;
  Bump b(ws);
  void *a, *x, *xf, *stack;
  PoolTmp pt;
  size_t stack_bytes;
  const size_t need = tower_carve(t, in_bf16, out_xf == nullptr, b, &a, &x, &xf, &pt, &stack, &stack_bytes);
  MR_REQUIRE(ws && ws_bytes >= need, MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(ws), MR_ERR_MISALIGNED);
  if (out_xf) xf = out_xf;
  const void* a_in = input;
  if (!in_bf16) {   // cast to the module dtype (modeling.py:402 / :453), zero padding K to the TMA-legal pitch
    MrCastDesc c{};
    c.rows = static_cast<int64_t>(t.B) * t.n;
    c.cols = t.K;
    c.ld_out = t.Kpad;
    const void* in[1] = {input};
    void* o[1] = {a};
    MR_TRY(mr_cast_pad(s, &c, in, o, nullptr, 0));
    a_in = a;
  }
  // embedding rows land at 1.. of every sequence; row 0 = CLS (modeling.py:311-320)
  MR_TRY(gemm(s, a_in, emb_w, emb_b, x, t.B * t.n, t.H, t.Kpad, t.Kpad, t.H, false, t.n, T, 1));
  {
    MrFillClsDesc f{t.B, T, t.H};
    const void* in[1] = {enc.cls};
    void* o[1] = {x};
    MR_TRY(mr_fill_cls(s, &f, in, o, nullptr, 0));
  }
  MR_TRY(run_stack(s, enc, x, rot, T, nullptr, nullptr, t.B, T, t.H, eps, stack, stack_bytes));
  MR_TRY(layernorm(s, x, enc.final_ln_g, enc.final_ln_b, xf, M, t.H, eps));
  MR_TRY(attn_pool(s, pool, xf, out_pool, pt, t.B, T, t.H, t.mode, t.win, t.n_win, t.gh, t.gw));
  if (out_cls)   // cls_proj on row 0 of every sequence (modeling.py:370-372): strided A operand
    MR_TRY(gemm(s, xf, enc.cls_proj_w, enc.cls_proj_b, out_cls, t.B, t.H, t.H, T * t.H, t.H, false));
  return MR_OK;
}

bool vision_shape(const MrVisionDesc* d, TowerShape* t) {
  if (!d || d->B < 0 || d->grid_h <= 0 || d->grid_w <= 0 || (d->grid_h & 1) || (d->grid_w & 1) || d->H <= 0 || d->H % 64)
    return false;
  const int hw = d->grid_h * d->grid_w;
  *t = TowerShape{d->B, hw, 768, 768, d->H, 1, 4, hw / 4, d->grid_h, d->grid_w};
  return true;
}

bool audio_shape(const MrAudioDesc* d, TowerShape* t) {
  if (!d || d->B < 0 || d->seq_len <= 0 || d->patch <= 0 || d->seq_len % d->patch || d->pool_ratio <= 0 || d->H <= 0 ||
      d->H % 64 || d->patch * 65 > kAudioKPad)
    return false;
  const int n = d->seq_len / d->patch;
  if (n % d->pool_ratio) return false;
  *t = TowerShape{d->B, n, d->patch * 65, kAudioKPad, d->H, 0, d->pool_ratio, n / d->pool_ratio, 0, 0};
  return true;
}

size_t tower_ws(const TowerShape& t, bool in_bf16, bool need_xf) {
  Bump b(nullptr);
  void *a, *x, *xf, *stack;
  PoolTmp pt;
  size_t sb;
  return tower_carve(t, in_bf16, need_xf, b, &a, &x, &xf, &pt, &stack, &sb);
}

// ---------------------------------------------------------------------------------------------------------------
// joint
// ---------------------------------------------------------------------------------------------------------------
struct JointTmp {
  void *x, *rot, *valid, *kept, *proj, *stack;
  size_t stack_bytes;
};

size_t joint_carve(const MrJointDesc* d, Bump& b, JointTmp* t) {
  const size_t M = static_cast<size_t>(d->B) * d->T, keep = static_cast<size_t>(d->B) * d->n_keep;
  t->x = b.take(M * d->H * 2);
  t->rot = b.take(M * 32 * 4);
  t->valid = b.take(M);
  t->kept = b.take(keep * d->H * 2);
  t->proj = b.take(keep * d->H * 4);
  t->stack_bytes = stack_ws(d->B, d->T, d->H);
  t->stack = b.take(t->stack_bytes);
  return b.off;
}

bool joint_shape_ok(const MrJointDesc* d) {
  return d && d->B >= 0 && d->L > 0 && d->H > 0 && d->H % 64 == 0 && d->n_aud >= 0 && d->n_vis >= 0 &&
         d->T >= d->L + d->n_vis && d->n_keep > 0 && d->n_keep <= d->T && d->vocab > 0 && d->seg_div >= 0;
}

int32_t joint_run(mr_stream_t s, const MrJointDesc* d, const void* tokens, const void* seg, const void* aud_ptr,
                  const void* audio, const void* vision, const void* group, const void* tok_embs, void* out, void* ws,
                  size_t ws_bytes) {
  Bump b(ws);
  JointTmp t;
  const size_t need = joint_carve(d, b, &t);
  MR_REQUIRE(ws && ws_bytes >= need, MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(ws), MR_ERR_MISALIGNED);
  MrAssembleDesc a{};
  a.B = d->B;
  a.L = d->L;
  a.H = d->H;
  a.n_aud = (audio && aud_ptr) ? d->n_aud : 0;
  a.n_vis = vision ? d->n_vis : 0;
  a.pool_h = d->pool_h;
  a.pool_w = d->pool_w;
  a.T = d->T;
  a.vocab = d->vocab;
  a.audio_token_length = d->audio_token_length;
  a.aud_div = d->aud_div;
  a.vis_group = d->vis_group;
  a.vis_tile = d->vis_tile;
  a.seg_div = d->seg_div;
  {
    const void* in[7] = {tokens, seg, aud_ptr, d->tok_emb, audio, vision, tok_embs};
    void* o[3] = {t.x, t.rot, t.valid};
    MR_TRY(mr_embed_assemble(s, &a, in, o, nullptr, 0));
  }
  MR_TRY(run_stack(s, d->enc, t.x, t.rot, d->B * d->T, t.valid, group, d->B, d->T, d->H, d->ln_eps, t.stack, t.stack_bytes));
  // final_ln on the first n_keep rows of every sequence only (the rest is never read, modeling.py:842)
  void* kept = d->head ? t.kept : out;
  MR_TRY(layernorm(s, t.x, d->enc.final_ln_g, d->enc.final_ln_b, kept, d->B * d->n_keep, d->H, d->ln_eps, d->T, d->n_keep, 0));
  if (!d->head) return MR_OK;
  MR_TRY(gemm(s, kept, d->head_w, d->head_b, t.proj, d->B * d->n_keep, d->H, d->H, d->H, d->H, true));
  MrNormalizeDesc n{};
  n.rows = d->B * d->n_keep;
  n.H = d->H;
  n.in_f32 = 1;
  n.out_f32 = 1;
  n.scale = 1.0f;
  const void* in[2] = {t.proj, nullptr};
  void* o[1] = {out};
  return mr_unit_normalize(s, &n, in, o, nullptr, 0);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t mr_vision_encode_workspace_bytes(const MrVisionDesc* d) {
  TowerShape t;
  if (!vision_shape(d, &t) || d->B == 0) return 0;
  return tower_ws(t, d->in_bf16 != 0, true);
}

extern "C" int32_t mr_vision_encode(mr_stream_t stream, const MrVisionDesc* d, const void* const* in, void* const* out,
                                    void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  TowerShape t;
  MR_REQUIRE(vision_shape(d, &t), MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && out[0] && d->emb_w && d->emb_b && d->rot && enc_ok(d->enc, true) && pool_ok(d->pool), MR_ERR_NULL_POINTER);
  return tower_run(stream, t, d->enc, d->pool, d->emb_w, d->emb_b, d->rot, d->ln_eps, in[0], d->in_bf16 != 0, out[0], out[1],
                   out[2], workspace, workspace_bytes);
}

extern "C" size_t mr_audio_encode_workspace_bytes(const MrAudioDesc* d) {
  TowerShape t;
  if (!audio_shape(d, &t) || d->B == 0) return 0;
  return tower_ws(t, false, true);
}

extern "C" int32_t mr_audio_encode(mr_stream_t stream, const MrAudioDesc* d, const void* const* in, void* const* out,
                                   void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  TowerShape t;
  MR_REQUIRE(audio_shape(d, &t), MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && out[0] && d->emb_w && d->emb_b && d->rot && enc_ok(d->enc, true) && pool_ok(d->pool), MR_ERR_NULL_POINTER);
  return tower_run(stream, t, d->enc, d->pool, d->emb_w, d->emb_b, d->rot, d->ln_eps, in[0], false, out[0], out[1], nullptr,
                   workspace, workspace_bytes);
}

namespace {
struct SpanTmp {
  void *x, *rot, *valid, *cls_ln, *proj, *stack;
  size_t stack_bytes;
};
size_t span_carve(const MrSpanDesc* d, Bump& b, SpanTmp* t) {
  const size_t M = static_cast<size_t>(d->B) * (d->L + 1);
  t->x = b.take(M * d->H * 2);
  t->rot = b.take(M * 32 * 4);
  t->valid = b.take(M);
  t->cls_ln = b.take(static_cast<size_t>(d->B) * d->H * 2);
  t->proj = b.take(static_cast<size_t>(d->B) * d->H * 4);
  t->stack_bytes = stack_ws(d->B, d->L + 1, d->H);
  t->stack = b.take(t->stack_bytes);
  return b.off;
}
bool span_shape_ok(const MrSpanDesc* d) {
  return d && d->B >= 0 && d->L > 0 && d->L < 16 && d->H > 0 && d->H % 64 == 0 && d->vocab > 0;   // modeling.py:495
}
}  // namespace

extern "C" size_t mr_span_encode_workspace_bytes(const MrSpanDesc* d) {
  if (!span_shape_ok(d) || d->B == 0) return 0;
  Bump b(nullptr);
  SpanTmp t;
  return span_carve(d, b, &t);
}

extern "C" int32_t mr_span_encode(mr_stream_t stream, const MrSpanDesc* d, const void* const* in, void* const* out,
                                  void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  MR_REQUIRE(span_shape_ok(d), MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && out[0] && d->tok_emb && d->rot && enc_ok(d->enc, true), MR_ERR_NULL_POINTER);
  Bump b(workspace);
  SpanTmp t;
  const size_t need = span_carve(d, b, &t);
  MR_REQUIRE(workspace && workspace_bytes >= need, MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(workspace), MR_ERR_MISALIGNED);
  const int T = d->L + 1;
  MrAssembleDesc a{};
  a.B = d->B;
  a.L = d->L;
  a.H = d->H;
  a.T = d->L;
  a.vocab = d->vocab;
  a.audio_token_length = 6;
  a.cls_offset = 1;
  a.pool_h = a.pool_w = 1;
  {
    const void* ain[7] = {in[0], nullptr, nullptr, d->tok_emb, nullptr, nullptr, in[1]};
    void* o[3] = {t.x, t.rot, t.valid};
    MR_TRY(mr_embed_assemble(stream, &a, ain, o, nullptr, 0));
  }
  {
    MrFillClsDesc f{d->B, T, d->H};
    const void* fin[1] = {d->enc.cls};
    void* o[1] = {t.x};
    MR_TRY(mr_fill_cls(stream, &f, fin, o, nullptr, 0));
  }
  // coordinates (1..L) / 16 with the CLS row at zero: the caller's static table (modeling.py:498)
  MR_TRY(run_stack(stream, d->enc, t.x, d->rot, T, t.valid, nullptr, d->B, T, d->H, d->ln_eps, t.stack, t.stack_bytes));
  MR_TRY(layernorm(stream, t.x, d->enc.final_ln_g, d->enc.final_ln_b, t.cls_ln, d->B, d->H, d->ln_eps, T, 1, 0));
  void* proj = d->normalize ? t.proj : out[0];
  MR_TRY(gemm(stream, t.cls_ln, d->enc.cls_proj_w, d->enc.cls_proj_b, proj, d->B, d->H, d->H, d->H, d->H, true));
  if (!d->normalize) return MR_OK;
  MrNormalizeDesc n{};
  n.rows = d->B;
  n.H = d->H;
  n.in_f32 = 1;
  n.out_f32 = 1;
  n.scale = 1.0f;
  const void* nin[2] = {t.proj, nullptr};
  void* o[1] = {out[0]};
  return mr_unit_normalize(stream, &n, nin, o, nullptr, 0);
}

extern "C" size_t mr_joint_encode_workspace_bytes(const MrJointDesc* d) {
  if (!joint_shape_ok(d) || d->B == 0) return 0;
  Bump b(nullptr);
  JointTmp t;
  return joint_carve(d, b, &t);
}

extern "C" int32_t mr_joint_encode(mr_stream_t stream, const MrJointDesc* d, const void* const* in, void* const* out,
                                   void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  MR_REQUIRE(joint_shape_ok(d), MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && out[0] && d->tok_emb && enc_ok(d->enc, false), MR_ERR_NULL_POINTER);
  MR_REQUIRE(!d->head || (d->head_w && d->head_b), MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n_vis == 0 || in[4], MR_ERR_NULL_POINTER);                  // vision rows announced but not given
  MR_REQUIRE(!(in[2] && d->n_aud > 0) || in[3], MR_ERR_NULL_POINTER);        // audio pointers without audio spans
  return joint_run(stream, d, in[0], in[1], in[2], in[3], in[4], in[5], in[6], out[0], workspace, workspace_bytes);
}

namespace {
struct VideoTmp {
  void *vis_out, *aud_out, *sub;
  size_t sub_bytes;
};
bool video_descs(const MrEmbedVideoDesc* d, TowerShape* tv, TowerShape* ta) {
  if (!d || d->B < 0 || d->nseg <= 0 || d->L <= 0) return false;
  if (!vision_shape(&d->vision, tv) || !audio_shape(&d->audio, ta) || !joint_shape_ok(&d->joint)) return false;
  const int per = tv->n_win;
  return d->vision.B == d->B * d->nseg && d->audio.B == d->B * 3 * d->nseg && d->joint.B == d->B && d->joint.L == d->L &&
         d->joint.n_vis == d->nseg * per && d->joint.n_aud == 3 * d->nseg && d->joint.audio_token_length == ta->n_win &&
         d->joint.H == d->vision.H && d->joint.H == d->audio.H && d->joint.n_keep == d->L && d->joint.head == 1 &&
         d->joint.seg_div == 3 && d->joint.T == d->L + d->joint.n_vis;
}
size_t video_carve(const MrEmbedVideoDesc* d, const TowerShape& tv, const TowerShape& ta, Bump& b, VideoTmp* t) {
  t->vis_out = b.take(static_cast<size_t>(tv.B) * tv.n_win * tv.H * 2);
  t->aud_out = b.take(static_cast<size_t>(ta.B) * ta.n_win * ta.H * 2);
  // the three stages run one after the other: they share one scratch region
  size_t sub = tower_ws(tv, d->vision.in_bf16 != 0, true);
  const size_t sa = tower_ws(ta, false, true);
  if (sa > sub) sub = sa;
  const size_t sj = mr_joint_encode_workspace_bytes(&d->joint);
  if (sj > sub) sub = sj;
  t->sub_bytes = sub;
  t->sub = b.take(sub);
  return b.off;
}
}  // namespace

extern "C" size_t mr_embed_video_batched_workspace_bytes(const MrEmbedVideoDesc* d) {
  TowerShape tv, ta;
  if (!video_descs(d, &tv, &ta) || d->B == 0) return 0;
  Bump b(nullptr);
  VideoTmp t;
  return video_carve(d, tv, ta, b, &t);
}

extern "C" int32_t mr_embed_video_batched(mr_stream_t stream, const MrEmbedVideoDesc* d, const void* const* in,
                                          void* const* out, void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  TowerShape tv, ta;
  MR_REQUIRE(video_descs(d, &tv, &ta), MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && in[1] && in[2] && in[3] && out[0], MR_ERR_NULL_POINTER);
  Bump b(workspace);
  VideoTmp t;
  const size_t need = video_carve(d, tv, ta, b, &t);
  MR_REQUIRE(workspace && workspace_bytes >= need, MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(workspace), MR_ERR_MISALIGNED);
  {  // imgs_enc = vision_encoder(images)['seq_attnpool']            modeling.py:828-829
    const void* vin[1] = {in[0]};
    void* vout[3] = {t.vis_out, nullptr, nullptr};
    MR_TRY(mr_vision_encode(stream, &d->vision, vin, vout, t.sub, t.sub_bytes));
  }
  {  // audio_enc = audio_encoder(audio_clips)['seq_attnpool']       modeling.py:832
    const void* ain[1] = {in[1]};
    void* aout[2] = {t.aud_out, nullptr};
    MR_TRY(mr_audio_encode(stream, &d->audio, ain, aout, t.sub, t.sub_bytes));
  }
  // joint: token_segment_idx = subseg_idxs // 3, audio_pointers = subseg_idxs   modeling.py:834-843
  const void* jin[7] = {in[2], in[3], in[3], t.aud_out, t.vis_out, nullptr, nullptr};
  void* jout[1] = {out[0]};
  return mr_joint_encode(stream, &d->joint, jin, jout, t.sub, t.sub_bytes);
}
