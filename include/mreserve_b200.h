/*
 * libmreserve_b200 -- flat C ABI of the B200-native MERLOT Reserve encoder + contrastive hot path.
 *
 * The reference (rowanz/merlot_reserve) is pure Python/JAX/Flax and has no FFI of its own; the
 * boundary it exposes is the Flax module call (mreserve/modeling.py) whose arithmetic XLA emits.
 * Each entry below replaces the XLA-lowered op group of one reference call site (cited per entry)
 * and is what a jax.ffi custom call for that site would bind.
 *
 * Conventions (every entry):
 *   - `stream` is a cudaStream_t passed as void*.  The library only ENQUEUES work on it; it never
 *     synchronises, never allocates device memory and keeps no mutable global state.
 *   - `in` / `out` are arrays of DEVICE pointers (entries may be NULL where marked optional).
 *   - `workspace` is caller-allocated device scratch of at least <entry>_workspace_bytes(desc).
 *   - return value: MR_OK (0) or a negative MrStatus.  CUDA launch errors are returned as
 *     MR_ERR_CUDA_BASE - cudaError_t.
 *   - all matrices are row-major; "bf16" is __nv_bfloat16; row strides are in ELEMENTS.
 */
#ifndef MRESERVE_B200_H_
#define MRESERVE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mr_stream_t;

typedef enum MrStatus {
  MR_OK = 0,
  MR_ERR_BAD_SHAPE = -1,
  MR_ERR_MISALIGNED = -2,
  MR_ERR_UNSUPPORTED_ARCH = -3,
  MR_ERR_NULL_POINTER = -4,
  MR_ERR_WORKSPACE = -5,
  MR_ERR_DRIVER = -6,
  MR_ERR_NCCL_UNAVAILABLE = -7, /* libnccl.so.2 could not be loaded (collective entries only) */
  MR_ERR_CUDA_BASE = -1000,     /* MR_ERR_CUDA_BASE - cudaError_t */
  MR_ERR_NCCL_BASE = -2000      /* MR_ERR_NCCL_BASE - ncclResult_t */
} MrStatus;

/* Library / device probe.  Returns MR_OK when device `device` is compute capability 10.x. */
int32_t mr_check_device(int32_t device);
/* ABI version: bumped when any struct below changes. */
int32_t mr_abi_version(void);
/* Human readable text for a status code (static storage). */
const char* mr_status_string(int32_t status);
/* Number of kernels this library has launched since load (per process, all threads). */
int64_t mr_launch_count(void);
/* Diagnostics (not part of the reference-facing path): device buffer of 2*64*8 int64 that receives clock64()
 * stamps of the softmax phases of CTA 0 in every following tcgen05 attention launch; NULL switches it off. */
int32_t mr_debug_attention_trace(void* device_buffer);

/* ------------------------------------------------------------------------------------------
 * mr_gemm_epilogue: D = epilogue(A . W^T)      [tcgen05 tensor cores, TMEM accumulators, TMA]
 *
 * Replaces every flax nn.Dense / nn.DenseGeneral / nn.Conv(k=s=2) on the path:
 *   patch embed modeling.py:402, audio embed :453, qkv :228 (+ apply_rotary :116-144 and the
 *   1/sqrt(d) query scale of flax dot_product_attention_weights), attn_proj :235 (+ residual
 *   :275), MLP intermediate + my_gelu :252-254,:240, MLP out :255 (+ residual :279), cls_proj
 *   :371, head :631, attention-pool q/k/v/out :424,:468.
 *
 *   in[0]  A        bf16 [M, lda]      activations (K-major)
 *   in[1]  W        bf16 [N, K]        weight, TRANSPOSED flax kernel (flax stores [K, N])
 *   in[2]  bias     f32  [N]           optional (MR_EPI_BIAS)
 *   in[3]  residual bf16 [rows, ldd]   optional (MR_EPI_RESIDUAL), indexed like the output
 *   in[4]  rot      f32  [rot_rows,32] optional (MR_EPI_ROTARY_QKV) per-position scale table
 *   out[0] D        bf16 or f32 [rows, ldd]
 *
 * Epilogue order: acc -> +bias -> rotary/q-scale -> gelu -> +residual -> cast.
 * Output row of input row r:  seq_in == 0 ? r : (r / seq_in) * seq_out + (r % seq_in) + row_off
 * (used to drop the patch-embed result into rows 1.. of the [B, 1+hw, H] token buffer).
 * MR_EPI_ROTARY_QKV: N == 3*hidden; for columns c < 2*hidden with (c % 64) < 32 the value is
 * multiplied by rot[(r % rot_rows) * 32 + (c % 64)]; columns c < hidden are then multiplied by
 * q_scale.
 * Constraints: lda % 8 == 0, ldb % 8 == 0, ldd % 8 == 0 (row pitches; M, N, K are arbitrary), pointers 16-byte aligned.
 *
 * LayerNorm folding (TransformerLayer, modeling.py:272-279: `Dense(LN(x))`).  With W' = gamma (.) W,
 *   LN(x) W + b = rstd_r (x W' - mu_r colsum(W')) + (beta W + b),
 * so the pre-LN GEMMs (qkv, MLP intermediate) read the residual stream x directly and apply the row
 * statistics in the epilogue; the LayerNorm kernel, its [M, H] output and the re-read of it disappear.
 *   MR_EPI_LN_IN:     in[1] = W' (bf16 [N, K]), in[2] = beta W + b (f32 [N]), in[5] = stats f32 [K/64][M][2],
 *                     in[6] = colsum(W') f32 [N]; mu = sum/K, var = sumsq/K - mu^2 (flax LayerNorm), eps = ln_eps.
 *   MR_EPI_STATS_OUT: out[1] = stats f32 [N/64][M][2]: for each output row and each 64-column slab the sum and
 *                     the sum of squares of the bf16-ROUNDED outputs (N % 64 == 0).  The residual GEMMs
 *                     (attn_proj, MLP out) produce x, so they produce the statistics of the next LayerNorm.
 * Both need the bf16, un-remapped output path (no MR_EPI_OUT_F32, seq_in == 0).
 * ------------------------------------------------------------------------------------------ */
enum {
  MR_EPI_BIAS = 1,
  MR_EPI_GELU = 2,       /* x * sigmoid(1.702 x)  (modeling.py:240) */
  MR_EPI_RESIDUAL = 4,
  MR_EPI_ROTARY_QKV = 8,
  MR_EPI_OUT_F32 = 16,
  MR_EPI_LN_IN = 32,     /* A is the UN-normalised residual stream; LayerNorm is folded into the GEMM (below) */
  MR_EPI_STATS_OUT = 64, /* also emit per-row (sum, sum of squares) partials of the bf16 output (below) */
  MR_EPI_GELU_DUAL = 256,/* with MR_EPI_BIAS | MR_EPI_GELU (bf16 staged output, nothing else): out[2] (bf16, same shape and
                            pitch as out[0]) also receives the PRE-activation A W^T + b.  The training forward of the MLP
                            keeps both (pretrain_model.py:320 differentiates my_gelu): no separate GELU pass over HBM. */
  MR_EPI_DGELU = 128     /* backward of my_gelu fused into the dgrad that produces its input gradient:
                            D = acc * gelu'(in[3]), in[3] = the forward's PRE-activation, bf16, indexed like D
                            (gelu'(p) = s (1 + 1.702 p (1 - s)), s = sigmoid(1.702 p)).  bf16 staged output with
                            trans_b = 1, trans_a = 0 only; excludes every other epilogue flag. */
};

enum { MR_GEMM_ALGO_TCGEN05 = 0, MR_GEMM_ALGO_SIMT_DEBUG = 1 };

typedef struct MrGemmDesc {
  int32_t M, N, K;
  int32_t lda;
  int32_t ldd;
  int32_t epilogue; /* bit-or of MR_EPI_* */
  int32_t seq_in, seq_out, row_off;
  int32_t rot_rows;
  int32_t hidden;
  int32_t algo;     /* MR_GEMM_ALGO_* ; SIMT_DEBUG is a slow CUDA-core kernel kept for bisecting */
  float q_scale;
  int32_t num_sms;  /* 0 = use every SM of the current device */
  float ln_eps;     /* MR_EPI_LN_IN */
  int32_t algo_pair; /* 0 = auto, 1 = force CTA pairs (cta_group::2, 256 x 256 tiles), 2 = never; pairs need the
                        staged bf16 output path and a 256-wide tile, otherwise the field is ignored */
  /* Backward GEMMs: operands given TRANSPOSED (MN-major for the tensor core, no transposed copies in HBM):
       trans_a: in[0] is A^T, bf16 [K, lda >= M]        trans_b: in[1] is W^T, bf16 [K, ldb >= N]
     dgrad  dX[M, K_f] = dY[M, N_f] . W[N_f, K_f]:  A = dY, in[1] = the forward's packed W with trans_b = 1;
     wgrad  dW[N_f, K_f] = dY^T . X:                in[0] = dY with trans_a = 1, in[1] = X with trans_b = 1
     (the reduction then runs over the M_f token rows: use split_k).  Plain epilogue only (bias on the fp32 path).
     split_k > 1: the K range is cut in split_k parts, fp32 partials go to the workspace
     (mr_gemm_epilogue_workspace_bytes) and are summed in a fixed order by a second kernel. */
  int32_t trans_a, trans_b;
  int32_t ldb;       /* row pitch of in[1] in elements; 0 = dense (K, or N when trans_b) */
  int32_t split_k;
  int32_t tile_n;    /* 0 = auto (256 columns unless that leaves SMs without a tile), 128 or 256 = force the tile
                        width (256 needs N % 256 == 0).  The auto rule counts output tiles only; a split-K caller
                        that knows tiles x split_k fills the SMs asks for 256 here. */
} MrGemmDesc;

int32_t mr_gemm_epilogue(mr_stream_t stream, const MrGemmDesc* d, const void* const* in, void* const* out,
                         void* workspace, size_t workspace_bytes);
size_t mr_gemm_epilogue_workspace_bytes(const MrGemmDesc* d);

/* ------------------------------------------------------------------------------------------
 * mr_attention_fwd: softmax(q k^T + mask) v per (sequence, head), flash style, never
 * materialising [T, T].  Replaces apply_attention modeling.py:188-200 (after the rotary/q-scale
 * already folded into the qkv GEMM epilogue) and the mask construction of :343-358 / :743-756.
 *
 *   in[0]  qkv    bf16 [B*T, 3*H]   columns: q heads | k heads | v heads, 64 per head
 *   in[1]  valid  u8   [B*T]        optional; 0 = padding position (modeling.py:726)
 *   in[2]  group  i32  [B*T]        optional video_src id per position (modeling.py:746-756)
 *   out[0] ctx    bf16 [B*T, H]
 * Mask(q,k) = valid[q] & valid[k] & (group[q] == group[k]).  A query row whose mask is all
 * false sees a constant -1e10 bias in the reference, i.e. attends UNIFORMLY to all T keys; that
 * behaviour is reproduced (it matters for padding rows returned by embed_video :842).
 * ------------------------------------------------------------------------------------------ */
enum { MR_ATTN_ALGO_AUTO = 0, MR_ATTN_ALGO_MMA_SYNC = 1, MR_ATTN_ALGO_TCGEN05 = 2,
       /* forward only: force one of the two tensor-memory kernels (TCGEN05 / AUTO choose by sequence length) */
       MR_ATTN_ALGO_TCGEN05_2TILE = 3, MR_ATTN_ALGO_TCGEN05_3STREAM = 4 };
/* mr_attention_bwd `finish` flags: fold the next steps of the backward pass into the pass that converts the fp32 dQ
 * accumulator (one sweep over d(qkv) instead of mr_attention_bwd + mr_rotary_bwd + mr_colsum):
 *   EPILOGUE   out[0] becomes the gradient w.r.t. the qkv GEMM's PRE-epilogue output: dq * q_scale * rot, dk * rot on
 *              the first 32 dims of every head (in[6] = rot f32 [rot_rows, 32], NULL = no rotary, only q_scale)
 *   BIAS_GRAD  out[1] = f32 [3H] column sums of out[0] (the qkv bias gradient, modeling.py:228) */
enum { MR_ATTN_BWD_EPILOGUE = 1, MR_ATTN_BWD_BIAS_GRAD = 2 };

typedef struct MrAttentionDesc {
  int32_t B, T, H; /* H = hidden = 64 * heads */
  int32_t algo;    /* AUTO: tcgen05/TMEM kernel (128-query tiles) when T > 64; mma.sync kernels below that: one
                      32 x 32 tile per (sequence, head) in a two-warp CTA when T <= 32 (audio T=31, span T=16),
                      64-query tiles otherwise.  MMA_SYNC forces the 64-tile mma.sync kernel, TCGEN05 the
                      tensor-memory kernel (both directions).  Forward, tensor-memory path: two 128-query tiles per SM
                      on 128- / 96-key blocks (attention_tc.cu) up to T = 256 and for calls with group ids, three
                      independent tile streams on 64-key blocks (attention_tc3.cu) above; TCGEN05_2TILE /
                      TCGEN05_3STREAM force one of them. */
  int32_t emit_lse; /* fwd: 1 = also write out[1] = lse f32 [B, H/64, T], the natural-log row log-sum-exp of the
                       masked scores (what mr_attention_bwd needs to recompute the probabilities) */
  int32_t finish;   /* bwd: MR_ATTN_BWD_* flags, 0 = plain d(q, k, v) */
  int32_t rot_rows; /* bwd, MR_ATTN_BWD_EPILOGUE: rows of the rotary table in[6] (row r uses table row r % rot_rows) */
  float q_scale;    /* bwd, MR_ATTN_BWD_EPILOGUE: the 1/sqrt(64) the qkv epilogue multiplied q by */
} MrAttentionDesc;

int32_t mr_attention_fwd(mr_stream_t stream, const MrAttentionDesc* d, const void* const* in, void* const* out,
                         void* workspace, size_t workspace_bytes);
size_t mr_attention_fwd_workspace_bytes(const MrAttentionDesc* d);

/* mr_attention_bwd: gradients of mr_attention_fwd w.r.t. the (rotary- and 1/sqrt(d)-scaled) q, k, v it consumed.
 *   in[0] qkv bf16 [B*T, 3H] (the forward's input); in[1] dctx bf16 [B*T, H]; in[2] ctx bf16 [B*T, H] (the forward's
 *   output); in[3] lse f32 [B, H/64, T] (emit_lse); in[4] valid u8 optional; in[5] group i32 optional
 *   out[0] dqkv bf16 [B*T, 3H] (dq | dk | dv, same layout as qkv); apply mr_rotary_bwd next, then the qkv dgrad/wgrad
 *   -- or set d->finish (see MR_ATTN_BWD_* above; in[6] and out[1] are read only when their flag is set).
 *   workspace: mr_attention_bwd_workspace_bytes (fp32 dq accumulator + row sums of dctx . ctx).  AUTO with T <= 32
 *   runs the one-block kernel, which writes dq in place and leaves the accumulator part of the workspace untouched. */
int32_t mr_attention_bwd(mr_stream_t stream, const MrAttentionDesc* d, const void* const* in, void* const* out,
                         void* workspace, size_t workspace_bytes);
size_t mr_attention_bwd_workspace_bytes(const MrAttentionDesc* d);

/* ------------------------------------------------------------------------------------------
 * mr_layernorm_fwd: flax LayerNorm(eps) (modeling.py:272,277,360,366): fp32 statistics,
 * var = E[x^2] - E[x]^2, y = (x - mean) * rsqrt(var + eps) * gamma + beta.
 *   in[0] x bf16 [rows_in, ldx]; in[1] gamma f32 [H]; in[2] beta f32 [H]; out[0] y bf16 [rows, H]
 *   out[1] (only when emit_stats) stats f32 [H/64][rows][2] of the bf16-rounded y, in the MR_EPI_STATS_OUT layout (slab 0
 *   carries the whole row, the other slabs are zero): lets `pre_ln` (modeling.py:360) feed layer 0's folded LN.
 * Row map: output row r reads input row  seq_out == 0 ? r : (r / seq_out) * seq_in + (r % seq_out) + row_off
 * (used to normalise only the first L rows of every joint sequence / only the CLS row).
 * ------------------------------------------------------------------------------------------ */
typedef struct MrLayerNormDesc {
  int32_t rows; /* output rows */
  int32_t H;
  int32_t ldx;
  int32_t seq_in, seq_out, row_off;
  float eps;
  int32_t emit_stats; /* 1: out[1] is written (see above) */
} MrLayerNormDesc;

int32_t mr_layernorm_fwd(mr_stream_t stream, const MrLayerNormDesc* d, const void* const* in, void* const* out,
                         void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_transformer_encode: the whole layer stack of TransformerEncoder (modeling.py:283-376, without the CLS
 * prepend before it and the final_ln / cls_proj after it) as one call: pre_ln, then per layer the qkv GEMM
 * (LayerNorm folded in, rotary + 1/sqrt(d) epilogue), fused attention, attn_proj (+ residual), MLP-up (LayerNorm
 * folded in, GELU), MLP-down (+ residual).  The vision, audio, span and joint encoders are this call with their own
 * B, T, rotary table and masks -- it is what one jax.ffi custom call per encoder binds.
 *
 *   in[0] x       bf16 [B*T, H]     input tokens (CLS row already in place), before pre_ln
 *   in[1] rot     f32  [rot_rows, 32]  rotary scale table (rot_rows == T for the static tables, B*T per-row)
 *   in[2] valid   u8   [B*T]        optional (modeling.py:343-358)
 *   in[3] group   i32  [B*T]        optional video_src ids (modeling.py:743-756)
 *   in[4], in[5]  pre_ln gamma, beta f32 [H]
 *   in[6 + 8 l + i], layer l:  i = 0 wqkv' bf16 [3H, H] (gamma-scaled, transposed), 1 bqkv' f32 [3H] (beta W + b),
 *                 2 colsum(wqkv') f32 [3H], 3 wo bf16 [H, H], 4 w1' bf16 [4H, H], 5 b1' f32 [4H],
 *                 6 colsum(w1') f32 [4H], 7 w2 bf16 [H, 4H]      (see MR_EPI_LN_IN for the primed quantities)
 *   out[0] x_out  bf16 [B*T, H]     residual stream after the last layer (may alias in[0])
 *   workspace: mr_transformer_encode_workspace_bytes(desc)  (qkv, ctx, MLP hidden, row statistics)
 * ------------------------------------------------------------------------------------------ */
typedef struct MrEncoderDesc {
  int32_t B, T, H;
  int32_t num_layers;
  int32_t rot_rows;
  float ln_eps;
} MrEncoderDesc;
int32_t mr_transformer_encode(mr_stream_t stream, const MrEncoderDesc* d, const void* const* in, void* const* out,
                              void* workspace, size_t workspace_bytes);
size_t mr_transformer_encode_workspace_bytes(const MrEncoderDesc* d);

/* ------------------------------------------------------------------------------------------
 * Composite entries: one call per sub-module of the reference model, and one for the whole zero-shot forward
 * (what `jax.jit(model.apply(..., method=embed_video))` is in the reference: ONE call, modeling.py:1017-1019).
 * Host code only: each enqueues the kernels behind the per-op entries on the caller's stream and carves its
 * temporaries out of the caller's workspace; nothing is allocated, nothing synchronises, every call is CUDA-graph
 * capturable.  Weights are DEVICE pointers in the packed layouts of the per-op entries, grouped in the POD structs
 * below (owned by the caller; `layers` is a HOST array of device pointers).
 * ------------------------------------------------------------------------------------------ */
typedef struct MrEncoderWeights {
  const void* pre_ln_g;   const void* pre_ln_b;    /* f32 [H]   modeling.py:360 */
  const void* final_ln_g; const void* final_ln_b;  /* f32 [H]   modeling.py:366 */
  const void* cls;                                  /* f32 [H]   modeling.py:316; NULL for the joint encoder */
  const void* cls_proj_w; const void* cls_proj_b;   /* bf16 [H, H], f32 [H]   modeling.py:371; NULL likewise */
  const void* const* layers; /* host array of 8 * num_layers device pointers, in mr_transformer_encode's order */
  int32_t num_layers;
  int32_t reserved_;
} MrEncoderWeights;

typedef struct MrPoolWeights {   /* seq_attnpool (flax MultiHeadDotProductAttention), modeling.py:424,468 */
  const void* wq;  const void* bq;    /* bf16 [H, H], f32 [H] */
  const void* wkv; const void* bkv;   /* bf16 [2H, H] (key rows, then value rows), f32 [2H] */
  const void* wo;  const void* bo;    /* bf16 [H, H], f32 [H] */
} MrPoolWeights;

/* mr_vision_encode: VisionTransformer.__call__ (modeling.py:391-430): cast, patch embed into rows 1.. of
 * [B, 1 + hw, H], CLS, layer stack, final_ln, 2x2 attention pool, optional cls_proj.
 *   in[0] images f32 (or bf16 when in_bf16) [B, hw, 768]
 *   out[0] seq_attnpool bf16 [B, hw/4, H]; out[1] cls bf16 [B, H] optional; out[2] seq incl. CLS row, after
 *   final_ln, bf16 [B*(1+hw), H] optional */
typedef struct MrVisionDesc {
  int32_t B, grid_h, grid_w, H;
  int32_t in_bf16;
  float ln_eps;
  MrEncoderWeights enc;
  MrPoolWeights pool;
  const void* emb_w; const void* emb_b;   /* bf16 [H, 768], f32 [H]   modeling.py:402 */
  const void* rot;                        /* f32 [1 + hw, 32] scale table of get_rotary_coordinates_2d, CLS row first */
} MrVisionDesc;
int32_t mr_vision_encode(mr_stream_t stream, const MrVisionDesc* d, const void* const* in, void* const* out,
                         void* workspace, size_t workspace_bytes);
size_t mr_vision_encode_workspace_bytes(const MrVisionDesc* d);

/* mr_audio_encode: AudioTransformer.__call__ (modeling.py:443-476).
 *   in[0] audio f32 [B, seq_len, 65]; out[0] seq_attnpool bf16 [B, seq_len/patch/pool_ratio, H]; out[1] cls bf16 [B, H] optional */
typedef struct MrAudioDesc {
  int32_t B, seq_len, patch, pool_ratio, H;
  float ln_eps;
  MrEncoderWeights enc;
  MrPoolWeights pool;
  const void* emb_w; const void* emb_b;   /* bf16 [H, 192] (K = patch * 65 = 130 zero padded to 192), f32 [H]   :453 */
  const void* rot;                        /* f32 [1 + seq_len/patch, 32] */
} MrAudioDesc;
int32_t mr_audio_encode(mr_stream_t stream, const MrAudioDesc* d, const void* const* in, void* const* out,
                        void* workspace, size_t workspace_bytes);
size_t mr_audio_encode_workspace_bytes(const MrAudioDesc* d);

/* mr_span_encode: TokenEmbedder + SpanTransformer (modeling.py:488-504, 516-538); with `normalize` also the
 * unit_normalize of embed_text_spans_only (:767-774).
 *   in[0] tokens i32 [B, L] (0 = padding); in[1] embedded spans bf16 [B, L, H] optional (replaces the table lookup;
 *   in[0] then only carries validity); out[0] f32 [B, H] */
typedef struct MrSpanDesc {
  int32_t B, L, H, vocab;
  int32_t normalize;
  float ln_eps;
  MrEncoderWeights enc;
  const void* tok_emb;                    /* bf16 [vocab, H] */
  const void* rot;                        /* f32 [1 + L, 32] */
} MrSpanDesc;
int32_t mr_span_encode(mr_stream_t stream, const MrSpanDesc* d, const void* const* in, void* const* out,
                       void* workspace, size_t workspace_bytes);
size_t mr_span_encode_workspace_bytes(const MrSpanDesc* d);

/* mr_joint_encode: prepare_multimodal_inputs (modeling.py:651-758) + joint_transformer (:297-376) + final_ln on the
 * first n_keep rows of every sequence; with `head` also joint_proj + unit_normalize (:842).
 *   in[0] tokens i32 [B, L]; in[1] seg_idx i32 [B, L] optional; in[2] aud_ptr i32 [B, L] optional;
 *   in[3] audio bf16 [B', n_aud, audio_token_length, H] optional; in[4] vision bf16 [B', n_vis, H] optional;
 *   in[5] group i32 [B*T] optional video_src ids (:746-756); in[6] token_embs bf16 [B, L, H] optional
 *   out[0] head ? f32 [B*n_keep, H] unit rows : bf16 [B*n_keep, H] final_ln rows
 * Field meanings as in MrAssembleDesc. */
typedef struct MrJointDesc {
  int32_t B, L, H;
  int32_t n_aud, n_vis, pool_h, pool_w, T, vocab, audio_token_length;
  int32_t aud_div, vis_group, vis_tile, seg_div;
  int32_t n_keep;
  int32_t head;
  float ln_eps;
  int32_t reserved_;
  MrEncoderWeights enc;
  const void* tok_emb;                    /* bf16 [vocab, H] */
  const void* head_w; const void* head_b; /* bf16 [H, H], f32 [H]   modeling.py:631 */
} MrJointDesc;
int32_t mr_joint_encode(mr_stream_t stream, const MrJointDesc* d, const void* const* in, void* const* out,
                        void* workspace, size_t workspace_bytes);
size_t mr_joint_encode_workspace_bytes(const MrJointDesc* d);

/* mr_embed_video_batched: batch_embed_video (modeling.py:806-846) for B clips of nseg segments in ONE call.
 *   in[0] images f32 [B*nseg, hw, 768]; in[1] audio_clips f32 [B*3*nseg, seq_len, 65]; in[2] tokens i32 [B, L];
 *   in[3] subseg_idxs i32 [B, L];  out[0] f32 [B, L, H] unit rows
 * The three sub-descriptors must describe exactly this composition (vision.B = B nseg, audio.B = 3 B nseg,
 * joint: B, L, n_vis = nseg hw/4, n_aud = 3 nseg, T = L + n_vis, n_keep = L, head = 1, seg_div = 3). */
typedef struct MrEmbedVideoDesc {
  int32_t B, nseg, L;
  int32_t reserved_;
  MrVisionDesc vision;
  MrAudioDesc audio;
  MrJointDesc joint;
} MrEmbedVideoDesc;
int32_t mr_embed_video_batched(mr_stream_t stream, const MrEmbedVideoDesc* d, const void* const* in, void* const* out,
                               void* workspace, size_t workspace_bytes);
size_t mr_embed_video_batched_workspace_bytes(const MrEmbedVideoDesc* d);

/* ------------------------------------------------------------------------------------------
 * mr_cast_pad: f32 [rows, cols] -> bf16 [rows, ld_out] (zero padded columns).  Stages the
 * fp32 pixel patches (modeling.py:402 casts to module dtype) and the [B,30,130] audio patches
 * (:453; K=130 is padded to a TMA-legal pitch).
 *   in[0] x f32 [rows, cols]; out[0] y bf16 [rows, ld_out]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrCastDesc {
  int64_t rows;
  int32_t cols;
  int32_t ld_out;
  int32_t sanitize; /* 1: jnp.nan_to_num before the cast (NaN -> 0, +-inf -> +-largest finite bf16): the gradient arena,
                       sanitised per replica BEFORE the all-reduce as pretrain_model.py:327-329 orders it */
  int32_t reserved_;
} MrCastDesc;
int32_t mr_cast_pad(mr_stream_t stream, const MrCastDesc* d, const void* const* in, void* const* out,
                    void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_patchify_u8: uint8 frames, already at the model's resolution -> bf16 patch rows, i.e. the tail of
 * preprocess_image_to_patches (mreserve/preprocess.py:464-481) on the device:
 *   tf.image.convert_image_dtype(uint8 -> float32)   = float32(u8) * float32(1/255)              (:471)
 *   tf.nn.space_to_depth(img, 16, NHWC) + reshape    = patch (by, bx) row-major, features (py, px, c)   (:478-479)
 * followed by the cast to the module dtype that modeling.py:402 applies (bf16, round to nearest even), so the
 * output is bit-identical to mr_cast_pad of the reference's fp32 patches.  Host-to-device traffic per frame is
 * 1 byte per sample instead of 4.  (The bilinear resize + pad in front of it, preprocess.py:474-477, is TensorFlow
 * code with no restatement here: frames must arrive at grid_h*16 x grid_w*16.)
 *   in[0] frames u8 [B, grid_h*patch, grid_w*patch, 3] (NHWC); out[0] patches bf16 [B*grid_h*grid_w, patch*patch*3]
 * patch must be 16 (768 features); pointers 16-byte aligned.
 * ------------------------------------------------------------------------------------------ */
typedef struct MrPatchifyDesc {
  int32_t B, grid_h, grid_w, patch;
} MrPatchifyDesc;
int32_t mr_patchify_u8(mr_stream_t stream, const MrPatchifyDesc* d, const void* const* in, void* const* out,
                       void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_logmel: make_spectrogram (mreserve/preprocess.py:121-156) on the device: librosa.feature.melspectrogram
 * (sr 22050, n_fft 1536, hop 588, 64 Slaney mels 20..11025 Hz, symmetric Hann, centred / reflect padded, power 2),
 * log(mel + 0.1) - log(0.1), the playback-speed column.  The STFT is one tcgen05 GEMM against the windowed DFT
 * matrix with hi / lo split bf16 operands (three products along K, fp32 accumulate).
 *   in[0] waveform f32 [B, n_samples]
 *   in[1] tables   persistent device buffer of mr_logmel_tables_bytes() (DFT matrix + mel basis); written when
 *                  build_tables = 1 (first call), only read afterwards
 *   out[0] split = 0: f32 [B, n_frames, 65], n_frames = 1 + n_samples / 588
 *          split = 1: f32 [B, 3, 60, 65], the three sub-clips with `pad_size` gap frames (:151-155); needs
 *                     n_frames == 180 + 4 * pad_size (n_samples = 110250 = 5 s for pad_size 2)
 * ------------------------------------------------------------------------------------------ */
typedef struct MrLogmelDesc {
  int32_t B, n_samples;
  int32_t playback_speed; /* 1 (the only value the reference's callers pass, preprocess.py:424,448) */
  int32_t pad_size, split;
  int32_t build_tables;
} MrLogmelDesc;
int32_t mr_logmel(mr_stream_t stream, const MrLogmelDesc* d, const void* const* in, void* const* out,
                  void* workspace, size_t workspace_bytes);
size_t mr_logmel_workspace_bytes(const MrLogmelDesc* d);
size_t mr_logmel_tables_bytes(void);

/* ------------------------------------------------------------------------------------------
 * mr_resize_patchify: preprocess_image_to_patches (mreserve/preprocess.py:464-481) for uint8 frames of ANY size:
 * convert_image_dtype, resize_and_pad (pretrain/data_utils.py:42-147 with do_random_scale = False,
 * do_flip_if_vertical = False: aspect-preserving scale = min(dw / w, dh / h), tf.image.resize BILINEAR with
 * antialias = True, clip to [0, 1], zero pad bottom / right), space_to_depth(16), cast to the module dtype.
 *   in[0] frames u8 [B, in_h, in_w, 3];  out[0] patches bf16 (f32 when out_f32) [B, grid_h * grid_w, 768]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrResizePatchifyDesc {
  int32_t B, in_h, in_w, grid_h, grid_w;
  int32_t out_f32;
} MrResizePatchifyDesc;
int32_t mr_resize_patchify(mr_stream_t stream, const MrResizePatchifyDesc* d, const void* const* in, void* const* out,
                           void* workspace, size_t workspace_bytes);
size_t mr_resize_patchify_workspace_bytes(const MrResizePatchifyDesc* d);

/* ------------------------------------------------------------------------------------------
 * mr_fill_cls: writes the CLS parameter into row 0 of every sequence of x [B, T, H]
 * (modeling.py:316-320).   in[0] cls f32 [H]; out[0] x bf16 [B*T, H]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrFillClsDesc {
  int32_t B, T, H;
} MrFillClsDesc;
int32_t mr_fill_cls(mr_stream_t stream, const MrFillClsDesc* d, const void* const* in, void* const* out,
                    void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * Attention pooling (flax MultiHeadDotProductAttention with a single mean query per window;
 * modeling.py:413-429 vision 2x2 windows, :464-475 audio windows of 5).
 *
 * mr_window_mean: q_in[w] = mean over the window's tokens of x.
 *   in[0] x bf16 [B, T, H] (row 0 = CLS, skipped);  out[0] q_in bf16 [B*n_win, H]
 * mr_window_attention: ctx[w] = softmax_j(q[w].k[w,j] / 8) v[w,j], per head.
 *   in[0] q bf16 [B*n_win, H]; in[1] kv bf16 [B*T, 2H] (k | v, computed on every token row);
 *   out[0] ctx bf16 [B*n_win, H]
 * Window w of sequence b covers token rows 1 + idx(w, j):
 *   mode 0 (1-D, audio):  idx = w * win + j
 *   mode 1 (2-D, vision): w = (y, x) on the (grid_h/2, grid_w/2) grid, j = (dy, dx):
 *                         idx = (2y + dy) * grid_w + 2x + dx          (reshape/swapaxes :419-421)
 * ------------------------------------------------------------------------------------------ */
typedef struct MrWindowDesc {
  int32_t B, T, H;
  int32_t mode;
  int32_t win;            /* tokens per window (4 for 2x2, 5 for audio) */
  int32_t n_win;          /* windows per sequence */
  int32_t grid_h, grid_w; /* mode 1 only */
} MrWindowDesc;
int32_t mr_window_mean(mr_stream_t stream, const MrWindowDesc* d, const void* const* in, void* const* out,
                       void* workspace, size_t workspace_bytes);
int32_t mr_window_attention(mr_stream_t stream, const MrWindowDesc* d, const void* const* in, void* const* out,
                            void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_embed_assemble: builds the joint-encoder input (prepare_multimodal_inputs,
 * modeling.py:651-758, incl. TokenEmbedder :507-538 and the audio splice :685-695).
 *   in[0] tokens   i32  [B, L]
 *   in[1] seg_idx  i32  [B, L]     token_segment_idx (already // 3 where the caller does that); optional
 *   in[2] aud_ptr  i32  [B, L]     audio_pointers; optional (no audio splice when NULL)
 *   in[3] emb      bf16 [vocab, H] token embedding table
 *   in[4] audio    bf16 [B, n_aud, 6, H] audio_spans; optional
 *   in[5] vision   bf16 [B, n_vis, H]    vision_input; optional
 *   in[6] tok_emb  bf16 [B, L, H]        optional pre-computed token_embs (replaces the lookup)
 *   out[0] x     bf16 [B, T (+1 if cls_offset), H]   T = max(L + n_vis, pad_len)
 *   out[1] rot   f32  [B, T, 32]  rotary scale table (construct_rotary_sinusoids + apply_rotary,
 *                                 :81-144, 4 coordinate axes x 4 frequencies)
 *   out[2] valid u8   [B, T]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrAssembleDesc {
  int32_t B, L, H;
  int32_t n_aud;          /* audio spans per sequence (0 = none) */
  int32_t n_vis;          /* vision rows per sequence (0 = none) */
  int32_t pool_h, pool_w; /* pooled vision grid (output_grid / vit_pooling_ratio) */
  int32_t T;              /* total rows per sequence incl. right padding */
  int32_t vocab;
  int32_t audio_token_length; /* 6 */
  int32_t cls_offset;     /* 1: leave row 0 of every sequence free for a CLS token (span encoder,
                             modeling.py:311-326); the output then has T + 1 rows per sequence,
                             valid[b, 0] = 1 and rot[b, 0] is the zero-coordinate row */
  /* Source-row maps for the pretraining forward (pretrain_model.py:97-135), which feeds the SAME encoder
     outputs to several joint sequences; they avoid materialising the reference's repeat / tile copies:
       audio row of sequence b  = b / aud_div                         (jnp.repeat(audio_seq, aud_div, axis=0))
       vision row of sequence b = (b / (vis_group * vis_tile)) * vis_group + b % vis_group
                                                                       (jnp.tile(imgs_seq, [1, vis_tile, 1, 1]))
     0 in aud_div / vis_tile means identity (row b). */
  int32_t aud_div;
  int32_t vis_group, vis_tile;
  int32_t seg_div;        /* > 1: in[1] holds raw sub-segment indices and token_segment_idx = floor(in[1] / seg_div)
                             (`subseg_idxs // 3`, modeling.py:836); 0 / 1: in[1] is used as given */
} MrAssembleDesc;
int32_t mr_embed_assemble(mr_stream_t stream, const MrAssembleDesc* d, const void* const* in, void* const* out,
                          void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_rotary_table: scale table for coordinates given explicitly (vision / audio / span encoders,
 * whose coordinates are trace-time constants in the reference: modeling.py:405,457,498).
 *   in[0] coords f32 [rows, ndims]  (ndims in {1, 2, 4});  out[0] rot f32 [rows, 32]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrRotaryDesc {
  int32_t rows;
  int32_t ndims;
} MrRotaryDesc;
int32_t mr_rotary_table(mr_stream_t stream, const MrRotaryDesc* d, const void* const* in, void* const* out,
                        void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_unit_normalize: y = x / sqrt(sum(x^2) + 1e-5) * scale in fp32 (modeling.py:570-578; the
 * optional scale is exp(log_scale/2) of pretrain_model.py:251-253).
 *   in[0] x f32 or bf16 [rows_in, H]; in[1] row_idx i32 [rows] optional gather (y[r] = f(x[row_idx[r]]), the
 *   `x[batch_indexer, best_idxs]` selections of pretrain_model.py:176-182,229-230); out[0] y f32 or bf16 [rows, H]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrNormalizeDesc {
  int32_t rows, H;
  int32_t in_f32, out_f32;
  float scale;
  int32_t scale_from_device; /* 1: the scale is desc.scale * *(const float*)in[2] (forward) / in[3] (backward), read on
                                the device at run time -- the learned temperature of a CUDA-graph-captured train step */
} MrNormalizeDesc;
int32_t mr_unit_normalize(mr_stream_t stream, const MrNormalizeDesc* d, const void* const* in, void* const* out,
                          void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_segment_pool: one_hot_pool (modeling.py:541-567) as a gather/segment-sum.
 *   in[0] do_pool u8  [B, L]; in[1] idx i32 [B, L]; in[2] v bf16 or f32 [B, L, H]
 *   out[0] pooled f32 [B, S, H] (zero where no position points at the segment)
 *   out[1] count  f32 [B, S]    (idx_oh.sum(1))
 * ------------------------------------------------------------------------------------------ */
typedef struct MrSegmentPoolDesc {
  int32_t B, L, H, S;
  int32_t v_f32;
  int32_t accumulate; /* 1: add to pooled / count instead of overwriting (the t2sp sum, pretrain_model.py:190-191) */
} MrSegmentPoolDesc;
int32_t mr_segment_pool(mr_stream_t stream, const MrSegmentPoolDesc* d, const void* const* in, void* const* out,
                        void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_contrastive_lse_fwd: one direction of loss_fn_given_preds (pretrain_model.py:276,288-295) in ONE kernel
 * (csrc/contrastive_tc.cu): tcgen05 logit tiles in TMEM, online row log-sum-exp, positive logit from the diagonal
 * block, column tiles spread over a thread-block cluster whose partial (max, sum) pairs are merged through
 * distributed shared memory.  The [n, n_all] logit matrix is never written to memory; no workspace.
 *   numer[i] = x[i] . y_all[pos_off + i]        (= sum_h x[i,h] * y[i,h], the positive logit, :276)
 *   lse[i]   = logsumexp_j( x[i] . y_all[j] )   in fp32 (:293-294)
 *   in[0] x bf16 [n, H]; in[1] y_all bf16 [n_all, H] (rank-major all-gather of cat(y, y_extra), :288-290)
 *   out[0] lse f32 [n]; out[1] numer f32 [n]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrContrastiveDesc {
  int32_t n, n_all, H;
  int32_t pos_off;   /* column of y_all holding row 0's positive (rank * rows-per-rank of the rank-major all-gather) */
  float grad_scale;  /* _bwd: upstream gradient times 1 / (2 n) (pretrain_model.py:295) */
  /* _bwd: row remap of out[1] = dy_all so that the six directions of a step land in ONE rank-major buffer that a
   * single reduce-scatter returns to the owners: row r -> (r / dy_seq_in) * dy_seq_out + r % dy_seq_in + dy_row_off
   * (dy_seq_in = rows per rank of this direction; 0 = plain [n_all, H]) */
  int32_t dy_seq_in, dy_seq_out, dy_row_off;
} MrContrastiveDesc;
int32_t mr_contrastive_lse_fwd(mr_stream_t stream, const MrContrastiveDesc* d, const void* const* in,
                               void* const* out, void* workspace, size_t workspace_bytes);
size_t mr_contrastive_lse_fwd_workspace_bytes(const MrContrastiveDesc* d);

/* ------------------------------------------------------------------------------------------
 * mr_contrastive_reduce: the scalar tail of one direction of loss_fn_given_preds
 * (pretrain_model.py:295-300):
 *   out[0] = mean_i(lse[i] - numer[i])
 *   out[1 + s] = sum_i((lse[i] - numer[i]) * [sources[i] == s]) / (count_s + 1e-5),  s = 0, 1, 2
 *   in[0] lse f32 [n]; in[1] numer f32 [n]; in[2] sources i32 [n] optional; out[0] f32 [4]
 * ------------------------------------------------------------------------------------------ */
int32_t mr_contrastive_reduce(mr_stream_t stream, const MrContrastiveDesc* d, const void* const* in,
                              void* const* out, void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * Collective entries (csrc/collectives.cu): the exchange steps of the data-parallel pretraining step
 * (SURVEY.md 8e), enqueued on `stream` through a communicator the CALLER owns.  `comm` is an ncclComm_t (from the
 * host's own ncclCommInitRank, or from mr_nccl_comm_init below); the library resolves libnccl.so.2 with dlopen at the
 * first call and prefers the copy already loaded in the process.  One host thread per GPU; like every other entry
 * these only enqueue (stream-ordered, CUDA-graph capturable) and never synchronise.
 *   mr_all_gather      in[0] send [count]           -> out[0] recv [world * count], rank-major: replaces
 *                      jax.lax.all_gather(y, 'batch') of the packed contrastive targets (pretrain_model.py:288-290)
 *   mr_reduce_scatter  in[0] send [world * count]   -> out[0] recv [count]: the transpose of that all-gather in the
 *                      backward pass (dL/dy_all returned to the ranks that own the rows)
 *   mr_all_reduce      in[0] send [count]           -> out[0] recv [count] (may alias): jax.lax.pmean of the bf16
 *                      gradients and of the metrics (pretrain_model.py:329,333); `average` = 0 leaves the sum
 * ------------------------------------------------------------------------------------------ */
enum { MR_DT_BF16 = 0, MR_DT_F32 = 1 };
typedef struct MrCollectiveDesc {
  void* comm;      /* ncclComm_t */
  int64_t count;   /* elements: sent per rank (all_gather), received per rank (reduce_scatter), total (all_reduce) */
  int32_t dtype;   /* MR_DT_BF16 | MR_DT_F32 */
  int32_t average; /* reduce_scatter / all_reduce: 1 = mean over ranks (ncclAvg), 0 = sum */
} MrCollectiveDesc;
int32_t mr_all_gather(mr_stream_t stream, const MrCollectiveDesc* d, const void* const* in, void* const* out,
                      void* workspace, size_t workspace_bytes);
int32_t mr_reduce_scatter(mr_stream_t stream, const MrCollectiveDesc* d, const void* const* in, void* const* out,
                          void* workspace, size_t workspace_bytes);
int32_t mr_all_reduce(mr_stream_t stream, const MrCollectiveDesc* d, const void* const* in, void* const* out,
                      void* workspace, size_t workspace_bytes);
/* Communicator helpers for hosts without their own NCCL binding (thin wrappers over ncclGetUniqueId /
 * ncclCommInitRank / ncclCommDestroy / ncclCommCount + ncclCommUserRank).  mr_nccl_comm_init is collective over the
 * `world` ranks and must be called with the device of this rank current; id128 = the 128 bytes rank 0 obtained
 * from mr_nccl_unique_id and shared out of band.  mr_nccl_version returns ncclGetVersion's code (0 = unavailable). */
int32_t mr_nccl_version(void);
int32_t mr_nccl_unique_id(void* id128);
int32_t mr_nccl_comm_init(void** comm, int32_t world, int32_t rank, const void* id128);
int32_t mr_nccl_comm_destroy(void* comm);
int32_t mr_nccl_comm_info(void* comm, int32_t* world, int32_t* rank);

/* ------------------------------------------------------------------------------------------
 * Backward entries (pretrain_model.py:306-340 differentiates the whole forward with jax.value_and_grad; these are
 * the op-level gradients the forward kernels above need).  GEMM gradients are mr_gemm_epilogue with trans_a /
 * trans_b / split_k.
 *
 * mr_layernorm_bwd: in[0] dy bf16 [rows, H]; in[1] x bf16 [rows, H] (the LayerNorm INPUT; statistics are
 *   recomputed); in[2] gamma f32 [H]; in[3] dres bf16 [rows, H] optional, added to dx (gradient of the residual
 *   branch that bypasses the LayerNorm); out[0] dx bf16 [rows, H]; out[1] dgamma f32 [H]; out[2] dbeta f32 [H].
 *   Uses MrLayerNormDesc {rows, H, ldx = H, eps}; workspace mr_layernorm_bwd_workspace_bytes.
 * mr_gelu: mode 0: out[0] = x sigmoid(1.702 x) (modeling.py:240); mode 1: out[0] = dy * d/dx of that.
 *   in[0] x bf16 [n]; in[1] dy bf16 [n] (mode 1); out[0] bf16 [n]
 * mr_colsum: out[0] f32 [cols] = column sums of in[0] bf16 [rows, ld] (bias gradients).
 * mr_rotary_bwd: backward of the qkv epilogue's element-wise factors, in place on out[0] = d(qkv) bf16
 *   [rows, 3 hidden]; in[0] = the forward's rotary table f32 [rot_rows, 32].
 * mr_contrastive_lse_bwd: one direction of the loss gradient.  in[0] x bf16 [n, H]; in[1] y_all bf16 [n_all, H];
 *   in[2] lse f32 [n] (forward); out[0] dx f32 [n, H]; out[1] dy_all f32 [n_all, H] (to be reduce-scattered to
 *   the ranks that own the rows).  d(logits) = grad_scale (softmax - onehot(pos_off + i)) is formed in bf16 and both
 *   products run on the tensor cores.
 * mr_unit_normalize_bwd: in[0] x (forward input, f32 if in_f32 else bf16) [rows_in, H]; in[1] dy (f32 if out_f32
 *   else bf16) [rows, H]; in[2] row_idx i32 [rows] optional (the forward's gather); out[0] dx f32 [rows_in, H]
 *   (rows that are not referenced keep their previous contents: zero the buffer first when gathering).
 * ------------------------------------------------------------------------------------------ */
typedef struct MrElementwiseDesc {
  int64_t n;
  int32_t mode;
} MrElementwiseDesc;
typedef struct MrColsumDesc {
  int32_t rows, cols, ld;
} MrColsumDesc;
typedef struct MrRotaryBwdDesc {
  int64_t rows;
  int32_t hidden, rot_rows;
  float q_scale;
} MrRotaryBwdDesc;
int32_t mr_layernorm_bwd(mr_stream_t stream, const MrLayerNormDesc* d, const void* const* in, void* const* out,
                         void* workspace, size_t workspace_bytes);
size_t mr_layernorm_bwd_workspace_bytes(const MrLayerNormDesc* d);
int32_t mr_gelu(mr_stream_t stream, const MrElementwiseDesc* d, const void* const* in, void* const* out,
                void* workspace, size_t workspace_bytes);
int32_t mr_colsum(mr_stream_t stream, const MrColsumDesc* d, const void* const* in, void* const* out, void* workspace,
                  size_t workspace_bytes);
size_t mr_colsum_workspace_bytes(const MrColsumDesc* d);
int32_t mr_rotary_bwd(mr_stream_t stream, const MrRotaryBwdDesc* d, const void* const* in, void* const* out,
                      void* workspace, size_t workspace_bytes);
int32_t mr_contrastive_lse_bwd(mr_stream_t stream, const MrContrastiveDesc* d, const void* const* in, void* const* out,
                               void* workspace, size_t workspace_bytes);
size_t mr_contrastive_lse_bwd_workspace_bytes(const MrContrastiveDesc* d);
int32_t mr_unit_normalize_bwd(mr_stream_t stream, const MrNormalizeDesc* d, const void* const* in, void* const* out,
                              void* workspace, size_t workspace_bytes);
/* mr_segment_pool_bwd: in[0] do_pool u8 [B, L]; in[1] idx i32 [B, L]; in[2] dpooled f32 [B, S, H]; out[0] dv bf16 [B, L, H].
 * mr_embed_assemble_bwd (same descriptor as the forward): in[0] tokens; in[1] aud_ptr (optional); in[2] dx bf16
 *   [B, T (+1), H]; out[0] d_emb f32 [vocab, H] (optional), out[1] d_audio f32 [B / aud_div, n_aud, 6, H] (optional),
 *   out[2] d_vision f32 [vision rows, n_vis, H] (optional) -- all ACCUMULATED with atomics: zero them first.
 * mr_window_attention_bwd: in[0] q, in[1] kv (forward inputs), in[2] dctx bf16 [B*n_win, H];
 *   out[0] dq bf16 [B*n_win, H]; out[1] dkv bf16 [B*T, 2H] (every row written; CLS rows zero).
 * mr_window_mean_bwd: out[0] dx bf16 [B*T, H] += in[0] dq_in bf16 [B*n_win, H] / win on the window's token rows.
 * mr_add: out[0] bf16 [n] = in[0] bf16 + in[1] (bf16, or f32 when mode == 1). */
int32_t mr_segment_pool_bwd(mr_stream_t stream, const MrSegmentPoolDesc* d, const void* const* in, void* const* out,
                            void* workspace, size_t workspace_bytes);
int32_t mr_embed_assemble_bwd(mr_stream_t stream, const MrAssembleDesc* d, const void* const* in, void* const* out,
                              void* workspace, size_t workspace_bytes);
int32_t mr_window_attention_bwd(mr_stream_t stream, const MrWindowDesc* d, const void* const* in, void* const* out,
                                void* workspace, size_t workspace_bytes);
int32_t mr_window_mean_bwd(mr_stream_t stream, const MrWindowDesc* d, const void* const* in, void* const* out,
                           void* workspace, size_t workspace_bytes);
int32_t mr_add(mr_stream_t stream, const MrElementwiseDesc* d, const void* const* in, void* const* out, void* workspace,
               size_t workspace_bytes);

/* mr_adamw: the parameter update of train_step (pretrain_model.py:340 `apply_gradients`; optimizer of
 * pretrain/optimization.py:158-195 = Adam moments, decoupled weight decay on ndim > 1 parameters, schedule * lr),
 * with fp32 moments (the reference's bf16-compressed moments, optimization.py:36-51, are a "next" row).
 *   in[0] g f32 [n]; out[0] p f32 [n] (master, updated in place); out[1] m, out[2] v f32 [n];
 *   out[3] w bf16 [n] optional working copy (same layout as the master) refreshed from the new p. */
typedef struct MrAdamWDesc {
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay;
  float grad_scale; /* e.g. 1 / world_size after a SUM all-reduce */
} MrAdamWDesc;
int32_t mr_adamw(mr_stream_t stream, const MrAdamWDesc* d, const void* const* in, void* const* out, void* workspace,
                 size_t workspace_bytes);

/* mr_adamw_multi: the whole parameter update of train_step in ONE launch, with the reference's optimizer state
 * (pretrain/optimization.py:54-114 `scale_by_bfloat16_adam`, :158-195 `construct_train_state`).
 *   Arenas (same element order, tensor slots start at multiples of 8 elements):
 *   in[0]  g      f32 or bf16 [n]   flat gradient (bf16 = straight out of the reference's bf16 pmean, pretrain_model.py:329)
 *   in[1]  table  MrAdamWChunk [n_chunks]  (device memory; built once per model)
 *   out[0] p      f32 [n]           master parameters, updated in place
 *   out[1] m      bf16 | f32 [n]    first moment; bf16 when moments_bf16 (optimization.py:72,98)
 *   out[2] v      bf16 | f32 [n]    second moment; moments_bf16: sign-coded bf16 of nu^3 (optimization.py:36-51)
 *   Each chunk covers `count` <= MR_ADAMW_CHUNK consecutive elements of ONE tensor starting at arena element `offset`;
 *   `w` points at the same elements inside the working copy the forward kernels read (bf16, or f32 with
 *   MR_ADAMW_W_F32; NULL = none) and is refreshed from the new p.  MR_ADAMW_DECAY = flax ndim > 1 (optimization.py:182-184).
 *   update: g' = nan_to_num(g) grad_scale; m = (1-b1) g' + b1 m; v = (1-b2) g'^2 + b2 decode(v);
 *           p -= lr ( (m / bias_corr1) / (sqrt(v / bias_corr2) + eps) + wd p )   with the UN-rounded m, v (optimization.py:101-111);
 *   bias_corr = 1 - beta^t when the reference's do_bias_correction is set, else 1 (its default, optimization.py:165). */
enum { MR_ADAMW_DECAY = 1, MR_ADAMW_W_F32 = 2 };
enum { MR_ADAMW_CHUNK = 2048 };
typedef struct MrAdamWChunk {
  int64_t offset; /* first arena element of the chunk (multiple of 8) */
  void* w;        /* working copy of these elements, or NULL */
  int32_t count;  /* 1 .. MR_ADAMW_CHUNK */
  int32_t flags;  /* MR_ADAMW_DECAY | MR_ADAMW_W_F32 */
} MrAdamWChunk;
typedef struct MrAdamWMultiDesc {
  int32_t n_chunks;
  int32_t moments_bf16; /* 1 = reference storage (bf16 mu, coded bf16 nu), 0 = fp32 moments */
  int32_t grad_bf16;    /* dtype of in[0] */
  float lr, beta1, beta2, eps, weight_decay, grad_scale, bias_corr1, bias_corr2;
  float one_minus_beta1, one_minus_beta2; /* (float)(1.0 - (double)beta): rounded once, as a traced python constant is */
  int32_t hyper_from_device; /* 1: in[2] = f32 [3] {lr, bias_corr1, bias_corr2} in device memory replaces the three
                                fields above at run time (the step-dependent scalars of a CUDA-graph-captured step) */
} MrAdamWMultiDesc;
int32_t mr_adamw_multi(mr_stream_t stream, const MrAdamWMultiDesc* d, const void* const* in, void* const* out,
                       void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * mr_row_dot: logits[r] = x[r] . w  (finetune `proj` head H -> 1, no bias;
 * finetune/vcr/qa_qar_joint_finetune.py:163-170), with an optional row gather.
 *   in[0] x bf16 [rows_in, H]; in[1] w f32 [H]; in[2] row_idx i32 [rows] optional; out[0] f32 [rows]
 * ------------------------------------------------------------------------------------------ */
typedef struct MrRowDotDesc {
  int32_t rows, H;
} MrRowDotDesc;
int32_t mr_row_dot(mr_stream_t stream, const MrRowDotDesc* d, const void* const* in, void* const* out,
                   void* workspace, size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MRESERVE_B200_H_ */
