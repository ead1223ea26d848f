// HBM-bound row kernels of the backward pass: LayerNorm backward, sigmoid-GELU forward/backward, column sums
// (bias gradients), rotary-scale backward, softmax-probability materialisation for the contrastive loss gradient,
// unit_normalize backward.  Same conventions as rowwise.cu: one warp per row, 16-byte accesses, fp32 arithmetic.
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kWarps = 8;
constexpr int kThreadsRow = kWarps * 32;
constexpr int kMaxVec = 4;   // H <= 1024: at most 4 uint4 (8 bf16) vectors per lane

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8b(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward (flax LayerNorm, modeling.py:272,277,360,366):  y = (x - mu) rstd gamma + beta
//   g = dy gamma;  dx = rstd (g - mean(g) - xhat mean(g xhat)) [+ dres];  dgamma = sum_r dy xhat;  dbeta = sum_r dy
// Persistent blocks; every lane keeps the dgamma / dbeta partials of its own columns in registers, the 8 warps of a
// block are combined through shared memory, blocks through `partial` + layernorm_bwd_reduce_kernel.
// ---------------------------------------------------------------------------------------------
// NV = uint4 vectors per lane (H <= 256 NV): sized exactly so that two 256-thread blocks fit an SM (16 rows in
// flight per SM; one row = 3 or 4 streams of H bf16, all requested before the first reduction).
template <int NV>
__global__ void __launch_bounds__(kThreadsRow, NV <= 3 ? 2 : 1)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                     const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ partial, int rows, int H,
                     float eps) {
  extern __shared__ float red[];   // [kWarps][2][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H / 8;
  const float inv_h = 1.0f / static_cast<float>(H);
  float dg[NV][8], db[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[i][j] = 0.f, db[i][j] = 0.f;
  }
  for (int row = blockIdx.x * kWarps + warp; row < rows; row += gridDim.x * kWarps) {
    const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * H);
    const uint4* ds = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * H);
    const uint4* rs = reinterpret_cast<const uint4*>(dres + static_cast<size_t>(row) * H);
    uint4 xu[NV], du[NV], ru[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {             // every load of the row is in flight before anything is consumed
      const int v = lane + 32 * i;
      if (v < nvec) {
        xu[i] = __ldg(xs + v);
        du[i] = __ldg(ds + v);
        if (dres != nullptr) ru[i] = __ldg(rs + v);
      }
    }
    float xv[NV][8], dv[NV][8];
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
        unpack8b(xu[i], xv[i]);
        unpack8b(du[i], dv[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s += xv[i][j];
          s2 += xv[i][j] * xv[i][j];
        }
      }
    }
    s = wsum(s);
    s2 = wsum(s2);
    const float mean = s * inv_h;
    const float rstd = rsqrtf(s2 * inv_h - mean * mean + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v);       // L1-resident
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[i][j] - mean) * rstd;
          const float g = dv[i][j] * gm[j];
          dg[i][j] += dv[i][j] * xh;
          db[i][j] += dv[i][j];
          xv[i][j] = xh;
          dv[i][j] = g;
          sg += g;
          sgx += g * xh;
        }
      }
    }
    sg = wsum(sg) * inv_h;
    sgx = wsum(sgx) * inv_h;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dv[i][j] - sg - xv[i][j] * sgx);
        if (dres != nullptr) {
          float r[8];
          unpack8b(ru[i], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * H)[v] = pack8b(o);
      }
    }
  }
  // block reduction of the parameter-gradient partials
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red[(warp * 2 + 0) * H + v * 8 + j] = dg[i][j];
        red[(warp * 2 + 1) * H + v * 8 + j] = db[i][j];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) a += red[w * 2 * H + c];
    partial[static_cast<size_t>(blockIdx.x) * 2 * H + c] = a;
  }
}

// Sums the per-block partials [nblk][2H]: 32 columns x 8 row groups per block, so a launch is ~50 blocks of coalesced
// 128-byte row reads instead of 2H threads walking nblk rows one after the other.
__global__ void __launch_bounds__(256)
layernorm_bwd_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dgamma, float* __restrict__ dbeta,
                            int nblk, int H) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a = 0.f;
  if (c < 2 * H)
    for (int b = ry; b < nblk; b += 8) a += partial[static_cast<size_t>(b) * 2 * H + c];
  red[ry][cx] = a;
  __syncthreads();
  if (ry == 0 && c < 2 * H) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][cx];   // fixed order: deterministic
    if (c < H)
      dgamma[c] = t;
    else
      dbeta[c - H] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// my_gelu (modeling.py:240): u = p sigmoid(1.702 p);  du/dp = s (1 + 1.702 p (1 - s)), s = sigmoid(1.702 p)
// mode 0: y = gelu(x);  mode 1: y = dy * gelu'(x)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gelu_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ y, int64_t nvec, int mode) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[8], o[8];
    unpack8b(__ldg(reinterpret_cast<const uint4*>(x) + i), p);
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = p[j] / (1.0f + __expf(-1.702f * p[j]));
    } else {
      float d[8];
      unpack8b(__ldg(reinterpret_cast<const uint4*>(dy) + i), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = 1.0f / (1.0f + __expf(-1.702f * p[j]));
        o[j] = d[j] * s * (1.0f + 1.702f * p[j] * (1.0f - s));
      }
    }
    reinterpret_cast<uint4*>(y)[i] = pack8b(o);
  }
}

// ---------------------------------------------------------------------------------------------
// Column sums of a bf16 [rows, cols] matrix (bias gradients): partial per block, then a second stage.
// Block = 256 threads = 32 column-vectors (8 columns each) x 8 row lanes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int rows, int cols, int ld,
                      int rows_per_block) {
  const int cv = blockIdx.x * 32 + (threadIdx.x & 31);   // column vector index
  const int rl = threadIdx.x >> 5;                        // 0..7
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, rows);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cv * 8 < cols) {
    for (int r = r0 + rl; r < r1; r += 8) {
      float f[8];
      unpack8b(__ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ld) + cv), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
  __shared__ float red[8][32][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][threadIdx.x & 31][j] = acc[j];
  __syncthreads();
  if (rl == 0 && cv * 8 < cols) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x & 31][j];
      if (cv * 8 + j < cols) partial[static_cast<size_t>(blockIdx.y) * cols + cv * 8 + j] = a;
    }
  }
}

__global__ void colsum_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int nblk, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float a = 0.f;
  for (int b = 0; b < nblk; ++b) a += partial[static_cast<size_t>(b) * cols + c];
  out[c] = a;
}

// ---------------------------------------------------------------------------------------------
// Backward of the qkv epilogue (rotary scale on the first 32 dims of every q / k head, 1/sqrt(d) on q): the same
// element-wise factors applied to d(qkv), in place.  Only the vectors that change are visited -- all of q and the
// lower half of every k head, (H/8 + H/16) 16-byte vectors per row -- so every lane of every warp has a load in
// flight (the first version walked all 3H/8 vectors and idled on v and the upper k halves: 2.4 TB/s).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rotary_bwd_kernel(bf16* __restrict__ dqkv, const float* __restrict__ rot, int64_t rows, int hidden, int rot_rows,
                  float q_scale) {
  const int q_vecs = hidden / 8;
  const int per_row = q_vecs + hidden / 16;
  const int64_t total = rows * per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / per_row;
    const int v = static_cast<int>(i - r * per_row);
    const bool is_q = v < q_vecs;
    // k: vector kv of the row's active k vectors = head kv / 4, 8-column group kv % 4 of its lower half
    const int c = is_q ? v * 8 : hidden + ((v - q_vecs) >> 2) * 64 + ((v - q_vecs) & 3) * 8;
    const int in_head = c & 63;
    uint4* cell = reinterpret_cast<uint4*>(dqkv + r * 3 * hidden + c);
    float f[8];
    unpack8b(*cell, f);
    const float base = is_q ? q_scale : 1.0f;
    if (in_head < 32) {
      const float4* rp = reinterpret_cast<const float4*>(rot + static_cast<size_t>(r % rot_rows) * 32 + in_head);
      const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
      f[0] *= base * r0.x; f[1] *= base * r0.y; f[2] *= base * r0.z; f[3] *= base * r0.w;
      f[4] *= base * r1.x; f[5] *= base * r1.y; f[6] *= base * r1.z; f[7] *= base * r1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= base;
    }
    *cell = pack8b(f);
  }
}

// ---------------------------------------------------------------------------------------------
// Tail of mr_attention_bwd in `finish` mode: ONE pass over d(qkv) that (1) turns the fp32 dQ accumulator into bf16,
// (2) applies the backward of the qkv epilogue (rotary scale on the first 32 dims of every q / k head, 1/sqrt(d) on q)
// and (3) sums every column for the qkv bias gradient -- what dq_convert_kernel + rotary_bwd_kernel +
// colsum_partial_kernel did in three passes (24 bytes per element of [rows, H] moved; this kernel moves 11).
// Block = 32 column vectors (8 columns each) x 8 row lanes over `rows_per_block` rows, like colsum_partial_kernel.
// ---------------------------------------------------------------------------------------------
template <bool kAcc>
__global__ void __launch_bounds__(256)
dqkv_finish_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv, const float* __restrict__ rot,
                   float* __restrict__ partial, int rows, int H, int rot_rows, float q_scale, int rows_per_block) {
  const int cv = blockIdx.x * 32 + (threadIdx.x & 31);
  const int c = cv * 8;
  const int rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, rows);
  const int region = c < H ? 0 : (c < 2 * H ? 1 : 2);                 // q | k | v
  const int in_head = c & 63;
  const bool rotated = rot != nullptr && region < 2 && in_head < 32;
  const bool from_acc = kAcc && region == 0;
  const bool changes = from_acc || rotated || (region == 0 && q_scale != 1.0f);
  const float base = region == 0 ? q_scale : 1.0f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < 3 * H && (partial != nullptr || changes)) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 8) {
      uint4* cell = reinterpret_cast<uint4*>(dqkv + static_cast<size_t>(r) * 3 * H + c);
      float f[8];
      if (from_acc) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(dq_acc + static_cast<size_t>(r) * H + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(dq_acc + static_cast<size_t>(r) * H + c) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      } else {
        unpack8b(*cell, f);
      }
      if (rotated) {
        const float4* rp = reinterpret_cast<const float4*>(rot + static_cast<size_t>(r % rot_rows) * 32 + in_head);
        const float4 s0 = __ldg(rp), s1 = __ldg(rp + 1);
        f[0] *= base * s0.x; f[1] *= base * s0.y; f[2] *= base * s0.z; f[3] *= base * s0.w;
        f[4] *= base * s1.x; f[5] *= base * s1.y; f[6] *= base * s1.z; f[7] *= base * s1.w;
      } else if (region == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= base;
      }
      if (changes) *cell = pack8b(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
  if (partial == nullptr) return;
  __shared__ float red[8][32][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][threadIdx.x & 31][j] = acc[j];
  __syncthreads();
  if (rl == 0 && c < 3 * H) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x & 31][j];
      partial[static_cast<size_t>(blockIdx.y) * 3 * H + c + j] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Contrastive loss backward, step 1: dlogits[i, j] = w (exp(logit[i, j] - lse[i]) - [j == pos(i)]) as bf16
// (pretrain_model.py:293-295; w = 1 / (2 n), the weight of one direction's mean).  The positive of row i is column
// pos_off + i of y_all (this rank's block of the rank-major all-gather).  One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsRow)
contrastive_dlogits_kernel(const float* __restrict__ logits, const float* __restrict__ lse, bf16* __restrict__ dl_a,
                           bf16* __restrict__ dl_b, int n, int n_all, int ld, int pos_off, float w) {
  // d(logits) is a difference of nearly equal numbers whenever the embeddings are close to each other (e.g. right
  // after initialisation): a single bf16 would lose the gradient, so it is kept as hi + lo bf16 pairs.
  //   dl_a [n, 2 ld]: (hi | lo) side by side  -> dx     = dl_a . [Y; Y]
  //   dl_b [2 n, ld]: hi rows then lo rows     -> dy_all = dl_b^T . [X; X]
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* lr = logits + static_cast<size_t>(row) * ld;
  const float l = lse[row];
  for (int c = lane; c < ld; c += 32) {
    float v = 0.f;
    if (c < n_all) {
      v = __expf(lr[c] - l);
      if (c == pos_off + row) v -= 1.0f;
      v *= w;
    }
    const bf16 hi = __float2bfloat16(v);
    const bf16 lo = __float2bfloat16(v - __bfloat162float(hi));
    dl_a[static_cast<size_t>(row) * 2 * ld + c] = hi;
    dl_a[static_cast<size_t>(row) * 2 * ld + ld + c] = lo;
    dl_b[static_cast<size_t>(row) * ld + c] = hi;
    dl_b[static_cast<size_t>(n + row) * ld + c] = lo;
  }
}

// ---------------------------------------------------------------------------------------------
// unit_normalize backward (modeling.py:570-578, y = s x / r, r = sqrt(sum x^2 + 1e-5)):
//   dx = (s / r) (dy - x (x . dy) / r^2).   x f32 or bf16, dy bf16 or f32, dx f32.  Optional scatter through
//   row_idx (the gathered rows of the forward: dx[row_idx[i]] += ...; duplicates are not expected).
// ---------------------------------------------------------------------------------------------
template <bool kXF32, bool kDyF32>
__global__ void __launch_bounds__(kThreadsRow)
unit_normalize_bwd_kernel(const void* __restrict__ x_, const void* __restrict__ dy_, const int32_t* __restrict__ row_idx,
                          float* __restrict__ dx, int rows, int H, float scale, const float* __restrict__ scale_dev) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  if (scale_dev != nullptr) scale *= __ldg(scale_dev);
  const size_t xr = row_idx ? static_cast<size_t>(row_idx[row]) : static_cast<size_t>(row);
  auto X = [&](int c) {
    return kXF32 ? static_cast<const float*>(x_)[xr * H + c] : __bfloat162float(static_cast<const bf16*>(x_)[xr * H + c]);
  };
  auto DY = [&](int c) {
    return kDyF32 ? static_cast<const float*>(dy_)[static_cast<size_t>(row) * H + c]
                  : __bfloat162float(static_cast<const bf16*>(dy_)[static_cast<size_t>(row) * H + c]);
  };
  float ss = 0.f, dot = 0.f;
  for (int c = lane; c < H; c += 32) {
    const float xv = X(c);
    ss += xv * xv;
    dot += xv * DY(c);
  }
  ss = wsum(ss);
  dot = wsum(dot);
  const float r2 = ss + 1e-5f;
  const float a = scale * rsqrtf(r2);
  const float b = dot / r2;
  for (int c = lane; c < H; c += 32) dx[xr * H + c] = a * (DY(c) - X(c) * b);
}


// ---------------------------------------------------------------------------------------------
// one_hot_pool backward (modeling.py:541-567): dv[b, l] = do_pool[b, l] ? dpooled[b, idx[b, l]] : 0.  Warp per (b, l).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsRow)
segment_pool_bwd_kernel(const uint8_t* __restrict__ do_pool, const int32_t* __restrict__ idx,
                        const float* __restrict__ dpooled, bf16* __restrict__ dv, MrSegmentPoolDesc d) {
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= static_cast<int64_t>(d.B) * d.L) return;
  const int b = static_cast<int>(gw / d.L);
  const int s = idx[gw];
  const bool hit = do_pool[gw] != 0 && s >= 0 && s < d.S;
  const float* src = dpooled + (static_cast<size_t>(b) * d.S + (hit ? s : 0)) * d.H;
  bf16* dst = dv + static_cast<size_t>(gw) * d.H;
  for (int c = lane; c < d.H; c += 32) dst[c] = __float2bfloat16(hit ? src[c] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// Joint-input assembly backward (prepare_multimodal_inputs, modeling.py:651-758): every row of dx goes to where the
// forward took the row from -- the token embedding table, an audio span row or a (possibly shared) vision row.
// fp32 atomics: several sequences share audio / vision rows (pretrain_model.py:104-109) and tokens repeat.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsRow)
embed_assemble_bwd_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ aud_ptr,
                          const bf16* __restrict__ dx, float* __restrict__ d_emb, float* __restrict__ d_audio,
                          float* __restrict__ d_vision, MrAssembleDesc d) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int t_out = d.T + d.cls_offset;
  if (row >= static_cast<int64_t>(d.B) * t_out) return;
  const int b = static_cast<int>(row / t_out), t = static_cast<int>(row % t_out) - d.cls_offset;
  float* dst = nullptr;
  if (t < 0) return;                       // CLS slot: handled by the caller (gradient of the cls parameter)
  if (t < d.L) {
    const int32_t* trow = tokens + static_cast<size_t>(b) * d.L;
    const int tok = trow[t];
    if (aud_ptr != nullptr && d_audio != nullptr && tok == 5) {
      int cnt = 0;
      for (int base = 0; base <= t; base += 32) {
        const int i = base + lane;
        cnt += __popc(__ballot_sync(0xffffffffu, (i <= t) && (trow[i] == 5)));
      }
      const int sub = max(cnt - 1, 0) % d.audio_token_length;
      const int ptr = max(aud_ptr[static_cast<size_t>(b) * d.L + t], 0);
      const int ab = d.aud_div > 1 ? b / d.aud_div : b;
      dst = d_audio + ((static_cast<size_t>(ab) * d.n_aud + ptr) * d.audio_token_length + sub) * d.H;
    } else if (d_emb != nullptr) {
      dst = d_emb + static_cast<size_t>(min(max(tok, 0), d.vocab - 1)) * d.H;
    }
  } else if (t < d.L + d.n_vis && d_vision != nullptr) {
    const int vb = d.vis_tile > 0 ? (b / (d.vis_group * d.vis_tile)) * d.vis_group + b % d.vis_group : b;
    dst = d_vision + (static_cast<size_t>(vb) * d.n_vis + (t - d.L)) * d.H;
  }
  if (dst == nullptr) return;
  const bf16* src = dx + static_cast<size_t>(row) * d.H;
  for (int c = lane; c < d.H; c += 32) atomicAdd(dst + c, __bfloat162float(src[c]));
}

// ---------------------------------------------------------------------------------------------
// Attention-pool backward (flax MultiHeadDotProductAttention with one mean query per window; modeling.py:413-429,
// :464-475).  One warp per (window, head), lane owns 2 of the 64 head dims; the softmax over <= 8 keys is recomputed.
//   dq[w] and dkv[token] (k | v) are written by exactly one warp each; CLS rows of dkv are zeroed by window 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int win_token(const MrWindowDesc& d, int w, int j) {
  if (d.mode == 0) return w * d.win + j;
  const int w2 = d.grid_w / 2;
  const int y = w / w2, x = w % w2;
  return (2 * y + (j >> 1)) * d.grid_w + 2 * x + (j & 1);
}

__global__ void __launch_bounds__(kThreadsRow)
window_attention_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kv, const bf16* __restrict__ dctx,
                            bf16* __restrict__ dq, bf16* __restrict__ dkv, MrWindowDesc d) {
  const int nh = d.H / 64;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= static_cast<int64_t>(d.B) * d.n_win * nh) return;
  const int h = static_cast<int>(gw % nh);
  const int64_t bw = gw / nh;
  const int b = static_cast<int>(bw / d.n_win), w = static_cast<int>(bw % d.n_win);
  const uint32_t qq = reinterpret_cast<const uint32_t*>(q + static_cast<size_t>(bw) * d.H + h * 64)[lane];
  const uint32_t gg = reinterpret_cast<const uint32_t*>(dctx + static_cast<size_t>(bw) * d.H + h * 64)[lane];
  const float q0 = bf16_lo(qq) * 0.125f, q1 = bf16_hi(qq) * 0.125f;
  const float g0 = bf16_lo(gg), g1 = bf16_hi(gg);
  float sc[8], k0[8], k1[8], v0[8], v1[8], dp[8];
  float m = -INFINITY;
  for (int j = 0; j < d.win; ++j) {
    const size_t row = static_cast<size_t>(b) * d.T + 1 + win_token(d, w, j);
    const bf16* kr = kv + row * (2 * d.H) + h * 64;
    const uint32_t kk = reinterpret_cast<const uint32_t*>(kr)[lane];
    const uint32_t vv = reinterpret_cast<const uint32_t*>(kr + d.H)[lane];
    k0[j] = bf16_lo(kk); k1[j] = bf16_hi(kk);
    v0[j] = bf16_lo(vv); v1[j] = bf16_hi(vv);
    sc[j] = wsum(q0 * k0[j] + q1 * k1[j]);
    dp[j] = wsum(g0 * v0[j] + g1 * v1[j]);
    m = fmaxf(m, sc[j]);
  }
  float l = 0.f;
  for (int j = 0; j < d.win; ++j) {
    sc[j] = __expf(sc[j] - m);
    l += sc[j];
  }
  const float inv = 1.0f / l;
  float dsum = 0.f;
  for (int j = 0; j < d.win; ++j) {
    sc[j] *= inv;                 // p_j
    dsum += sc[j] * dp[j];
  }
  float dq0 = 0.f, dq1 = 0.f;
  for (int j = 0; j < d.win; ++j) {
    const float ds = sc[j] * (dp[j] - dsum);
    dq0 += ds * k0[j];
    dq1 += ds * k1[j];
    const size_t row = static_cast<size_t>(b) * d.T + 1 + win_token(d, w, j);
    bf16* o = dkv + row * (2 * d.H) + h * 64;
    reinterpret_cast<uint32_t*>(o)[lane] = pack_bf16(ds * q0, ds * q1);              // dk = ds * (q / 8)
    reinterpret_cast<uint32_t*>(o + d.H)[lane] = pack_bf16(sc[j] * g0, sc[j] * g1);  // dv = p * dctx
  }
  reinterpret_cast<uint32_t*>(dq + static_cast<size_t>(bw) * d.H + h * 64)[lane] = pack_bf16(dq0 * 0.125f, dq1 * 0.125f);
  if (w == 0) {                    // CLS row of this sequence takes no part in the pool
    bf16* o = dkv + static_cast<size_t>(b) * d.T * (2 * d.H) + h * 64;
    reinterpret_cast<uint32_t*>(o)[lane] = 0u;
    reinterpret_cast<uint32_t*>(o + d.H)[lane] = 0u;
  }
}

// dx[b, 1 + tok(w, j)] += dq_in[b, w] / win  (backward of the window mean); one thread per 8 columns of a token row.
__global__ void __launch_bounds__(256)
window_mean_bwd_kernel(const bf16* __restrict__ dq_in, bf16* __restrict__ dx, MrWindowDesc d) {
  const int vec_per_row = d.H / 8;
  const int64_t total = static_cast<int64_t>(d.B) * d.n_win * d.win * vec_per_row;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % vec_per_row);
  int64_t r = i / vec_per_row;
  const int j = static_cast<int>(r % d.win);
  r /= d.win;
  const int w = static_cast<int>(r % d.n_win), b = static_cast<int>(r / d.n_win);
  float a[8], g[8];
  unpack8b(__ldg(reinterpret_cast<const uint4*>(dq_in + (static_cast<size_t>(b) * d.n_win + w) * d.H) + v), g);
  uint4* cell = reinterpret_cast<uint4*>(dx + (static_cast<size_t>(b) * d.T + 1 + win_token(d, w, j)) * d.H) + v;
  unpack8b(*cell, a);
  const float inv = 1.0f / static_cast<float>(d.win);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] += g[k] * inv;
  *cell = pack8b(a);
}

// y (bf16) = a (bf16) + b (f32 or bf16): gradient accumulation of two branches.  One thread per 8 elements.
template <bool kBF32>
__global__ void __launch_bounds__(256)
add_kernel(const bf16* __restrict__ a, const void* __restrict__ b_, bf16* __restrict__ y, int64_t nvec) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float f[8], g[8];
    unpack8b(__ldg(reinterpret_cast<const uint4*>(a) + i), f);
    if (kBF32) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(b_) + 2 * i);
      const float4 w = __ldg(reinterpret_cast<const float4*>(b_) + 2 * i + 1);
      g[0] = u.x; g[1] = u.y; g[2] = u.z; g[3] = u.w; g[4] = w.x; g[5] = w.y; g[6] = w.z; g[7] = w.w;
    } else {
      unpack8b(__ldg(reinterpret_cast<const uint4*>(b_) + i), g);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += g[k];
    reinterpret_cast<uint4*>(y)[i] = pack8b(f);
  }
}


// ---------------------------------------------------------------------------------------------
// AdamW step of pretrain/optimization.py:54-114,158-195 with fp32 moments: g' = nan_to_num(g) * grad_scale;
// m = b1 m + (1 - b1) g'; v = b2 v + (1 - b2) g'^2; p -= lr (m / (sqrt(v) + eps) + wd p)  (no bias correction, the
// reference's default); optionally refreshes the bf16 working copy the GEMMs read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adamw_kernel(const float* __restrict__ g, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ w, MrAdamWDesc d) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < d.n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float gi = g[i];
    if (gi != gi) gi = 0.f;                                   // nan_to_num (pretrain_model.py:327)
    gi = fminf(fmaxf(gi, -3.4028235e38f), 3.4028235e38f) * d.grad_scale;
    const float mi = d.beta1 * m[i] + (1.0f - d.beta1) * gi;
    const float vi = d.beta2 * v[i] + (1.0f - d.beta2) * gi * gi;
    float pi = p[i];
    pi -= d.lr * (mi / (sqrtf(vi) + d.eps) + d.weight_decay * pi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (w != nullptr) w[i] = __float2bfloat16(pi);
  }
}


// ---------------------------------------------------------------------------------------------
// Multi-tensor AdamW: ONE launch updates every parameter of the model.  Parameters, moments and gradients live in
// flat arenas (same element order); a chunk table says, for every kChunk elements, where they start, how many
// there are, whether weight decay applies (flax ndim > 1, optimization.py:182-184) and where the working copy the
// GEMMs read lives.  Moments follow the reference's storage (pretrain/optimization.py:36-51,86-99):
//   mu  bf16 (round to nearest even of the fp32 value),
//   nu  the "unsigned bfloat16" code: bf16(nu^3), with the SIGN BIT as an extra mantissa bit -- a negative code c
//       decodes to cbrt(|c| (1 + 2^-9)), a non-negative one to cbrt(c); the encoder picks whichever is closer.
// The update uses the un-rounded fp32 moments of this step (optimization.py:101-111), optional bias correction,
// then decoupled weight decay and the scheduled learning rate (optimization.py:179-195).
// ---------------------------------------------------------------------------------------------
constexpr int kAdamChunk = 2048;   // elements per block: 256 threads x 8
static_assert(kAdamChunk == MR_ADAMW_CHUNK && sizeof(MrAdamWChunk) == 24, "chunk table layout is part of the ABI");

struct AdamElem {
  float p, m, v;
};

template <bool kBf16Moments>
__device__ __forceinline__ float adam_decode_v(float code) {
  if (!kBf16Moments) return code;
  const float a = fabsf(code);
  return cbrtf(code >= 0.f ? a : a * (1.0f + 1.0f / 512.0f));          // optimization.py:36-39
}

__device__ __forceinline__ bf16 adam_encode_v(float v) {
  const float v_pow = v * v * v;                                         // optimization.py:42-50
  const bf16 v_bf = __float2bfloat16(v_pow);
  const float v_bf32 = __bfloat162float(v_bf);
  const float err0 = fabsf(v_bf32 - v_pow);
  const float err1 = fabsf(v_bf32 * (1.0f + 1.0f / 512.0f) - v_pow);
  return err0 < err1 ? v_bf : __float2bfloat16(-v_bf32);
}

template <bool kBf16Moments, bool kBf16Grad>
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const void* __restrict__ g_, const MrAdamWChunk* __restrict__ table, float* __restrict__ p_,
                   void* __restrict__ m_, void* __restrict__ v_, MrAdamWMultiDesc d, const float* __restrict__ hyper) {
  if (hyper != nullptr) {   // step-dependent scalars from device memory: the launch can sit in a CUDA graph
    d.lr = __ldg(hyper);
    d.bias_corr1 = __ldg(hyper + 1);
    d.bias_corr2 = __ldg(hyper + 2);
  }
  const MrAdamWChunk ch = table[blockIdx.x];
  const int i0 = threadIdx.x * 8;
  if (i0 >= ch.count) return;
  const int64_t base = ch.offset + i0;
  const int n = min(8, ch.count - i0);
  const float wd = (ch.flags & MR_ADAMW_DECAY) ? d.weight_decay : 0.f;
  float g[8], p[8], m[8], v[8];
  if (n == 8) {   // offsets are multiples of 8 elements: 16-byte (bf16) / 32-byte (fp32) vector accesses
    if (kBf16Grad) {
      const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(g_) + base);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) g[2 * k] = bf16_lo(w[k]), g[2 * k + 1] = bf16_hi(w[k]);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(g_) + base);
      const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(g_) + base + 4);
      g[0] = a.x, g[1] = a.y, g[2] = a.z, g[3] = a.w, g[4] = b.x, g[5] = b.y, g[6] = b.z, g[7] = b.w;
    }
    const float4 pa = *reinterpret_cast<const float4*>(p_ + base), pb = *reinterpret_cast<const float4*>(p_ + base + 4);
    p[0] = pa.x, p[1] = pa.y, p[2] = pa.z, p[3] = pa.w, p[4] = pb.x, p[5] = pb.y, p[6] = pb.z, p[7] = pb.w;
    if (kBf16Moments) {
      const uint4 um = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(m_) + base);
      const uint4 uv = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(v_) + base);
      const uint32_t wm[4] = {um.x, um.y, um.z, um.w}, wv[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        m[2 * k] = bf16_lo(wm[k]), m[2 * k + 1] = bf16_hi(wm[k]);
        v[2 * k] = bf16_lo(wv[k]), v[2 * k + 1] = bf16_hi(wv[k]);
      }
    } else {
      const float* mf = static_cast<const float*>(m_) + base;
      const float* vf = static_cast<const float*>(v_) + base;
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = mf[k], v[k] = vf[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < n) {
        g[k] = kBf16Grad ? __bfloat162float(static_cast<const bf16*>(g_)[base + k]) : static_cast<const float*>(g_)[base + k];
        p[k] = p_[base + k];
        m[k] = kBf16Moments ? __bfloat162float(static_cast<const bf16*>(m_)[base + k]) : static_cast<const float*>(m_)[base + k];
        v[k] = kBf16Moments ? __bfloat162float(static_cast<const bf16*>(v_)[base + k]) : static_cast<const float*>(v_)[base + k];
      } else {
        g[k] = p[k] = m[k] = v[k] = 0.f;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float gi = g[k];
    if (gi != gi) gi = 0.f;                                             // nan_to_num (pretrain_model.py:327)
    gi = fminf(fmaxf(gi, -3.4028235e38f), 3.4028235e38f) * d.grad_scale;
    // products and sum rounded separately (no FMA contraction), (1 - beta) rounded once from double, as the
    // reference's XLA graph does: the bf16 rounding of mu sits on exact ties for bf16 inputs surprisingly often
    const float mi = __fadd_rn(__fmul_rn(d.one_minus_beta1, gi), __fmul_rn(d.beta1, m[k]));       // optimization.py:86-90
    const float vi = __fadd_rn(__fmul_rn(d.one_minus_beta2, __fmul_rn(gi, gi)),
                               __fmul_rn(d.beta2, adam_decode_v<kBf16Moments>(v[k])));           // :92-95
    const float upd = (mi / d.bias_corr1) / (sqrtf(vi / d.bias_corr2) + d.eps);                  // :101-111
    p[k] -= d.lr * (upd + wd * p[k]);
    m[k] = mi;
    v[k] = vi;
  }
  bf16* w16 = (ch.flags & MR_ADAMW_W_F32) ? nullptr : static_cast<bf16*>(ch.w);
  float* w32 = (ch.flags & MR_ADAMW_W_F32) ? static_cast<float*>(ch.w) : nullptr;
  if (n == 8) {
    *reinterpret_cast<float4*>(p_ + base) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(p_ + base + 4) = make_float4(p[4], p[5], p[6], p[7]);
    if (kBf16Moments) {
      uint4 um, uv;
      um.x = pack_bf16(m[0], m[1]), um.y = pack_bf16(m[2], m[3]), um.z = pack_bf16(m[4], m[5]), um.w = pack_bf16(m[6], m[7]);
      bf16 c[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) c[k] = adam_encode_v(v[k]);
      uv = *reinterpret_cast<const uint4*>(c);
      *reinterpret_cast<uint4*>(static_cast<bf16*>(m_) + base) = um;
      *reinterpret_cast<uint4*>(static_cast<bf16*>(v_) + base) = uv;
    } else {
      float* mf = static_cast<float*>(m_) + base;
      float* vf = static_cast<float*>(v_) + base;
#pragma unroll
      for (int k = 0; k < 8; ++k) mf[k] = m[k], vf[k] = v[k];
    }
    if (w16 != nullptr) {
      uint4 uw;
      uw.x = pack_bf16(p[0], p[1]), uw.y = pack_bf16(p[2], p[3]), uw.z = pack_bf16(p[4], p[5]), uw.w = pack_bf16(p[6], p[7]);
      *reinterpret_cast<uint4*>(w16 + i0) = uw;
    }
    if (w32 != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) w32[i0 + k] = p[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < n) {
        p_[base + k] = p[k];
        if (kBf16Moments) {
          static_cast<bf16*>(m_)[base + k] = __float2bfloat16(m[k]);
          static_cast<bf16*>(v_)[base + k] = adam_encode_v(v[k]);
        } else {
          static_cast<float*>(m_)[base + k] = m[k];
          static_cast<float*>(v_)[base + k] = v[k];
        }
        if (w16 != nullptr) w16[i0 + k] = __float2bfloat16(p[k]);
        if (w32 != nullptr) w32[i0 + k] = p[k];
      }
    }
  }
}

}  // namespace

size_t dqkv_finish_partial_bytes(int64_t rows, int H) {
  return static_cast<size_t>(ceil_div64(rows, 512)) * 3 * H * sizeof(float);
}

// dq_acc: fp32 [rows, H] accumulator of dQ, or nullptr when dQ is already bf16 inside dqkv; rot / partial + dbias optional
int32_t launch_dqkv_finish(cudaStream_t stream, const float* dq_acc, bf16* dqkv, const float* rot, int rot_rows,
                           float q_scale, int64_t rows, int H, float* partial, float* dbias) {
  if (rows >= (int64_t(1) << 31)) return MR_ERR_BAD_SHAPE;
  const int nblk = static_cast<int>(ceil_div64(rows, 512));
  dim3 grid(ceil_div(3 * H / 8, 32), nblk);
  float* part = dbias ? partial : nullptr;
  if (dq_acc)
    dqkv_finish_kernel<true><<<grid, 256, 0, stream>>>(dq_acc, dqkv, rot, part, static_cast<int>(rows), H, rot_rows, q_scale, 512);
  else
    dqkv_finish_kernel<false><<<grid, 256, 0, stream>>>(nullptr, dqkv, rot, part, static_cast<int>(rows), H, rot_rows, q_scale, 512);
  int32_t st = after_launch();
  if (st != MR_OK || !dbias) return st;
  colsum_reduce_kernel<<<ceil_div(3 * H, 256), 256, 0, stream>>>(partial, dbias, nblk, 3 * H);
  return after_launch();
}
}  // namespace mr

using namespace mr;

extern "C" size_t mr_layernorm_bwd_workspace_bytes(const MrLayerNormDesc* d) {
  if (!d || d->H <= 0) return 0;
  return static_cast<size_t>(148 * 2) * 2 * d->H * sizeof(float);
}

extern "C" int32_t mr_layernorm_bwd(mr_stream_t stream_, const MrLayerNormDesc* d, const void* const* in, void* const* out,
                                    void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && in[2] && out[0] && out[1] && out[2], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && d->H > 0 && d->H % 8 == 0 && d->H <= 8 * 32 * kMaxVec && d->seq_out == 0 && d->ldx == d->H,
             MR_ERR_BAD_SHAPE);
  MR_REQUIRE(workspace && workspace_bytes >= mr_layernorm_bwd_workspace_bytes(d), MR_ERR_WORKSPACE);
  for (int i = 0; i < 4; ++i) MR_REQUIRE(!in[i] || aligned16(in[i]), MR_ERR_MISALIGNED);
  MR_REQUIRE(aligned16(out[0]) && aligned16(workspace), MR_ERR_MISALIGNED);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int nblk = ceil_div(d->rows, kWarps);
  if (nblk > 148 * 2) nblk = 148 * 2;
  if (nblk < 1) nblk = 1;
  const size_t smem = static_cast<size_t>(kWarps) * 2 * d->H * sizeof(float);
  const int nv = ceil_div(d->H / 8, 32);
  auto launch = [&](auto kern) -> int32_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4);
    if (e != cudaSuccess) return cuda_status(e);
    kern<<<nblk, kThreadsRow, smem, stream>>>(
        static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]), static_cast<const float*>(in[2]),
        static_cast<const bf16*>(in[3]), static_cast<bf16*>(out[0]), static_cast<float*>(workspace), d->rows, d->H, d->eps);
    return MR_OK;
  };
  int32_t lst = nv <= 1 ? launch(layernorm_bwd_kernel<1>) : nv == 2 ? launch(layernorm_bwd_kernel<2>)
              : nv == 3 ? launch(layernorm_bwd_kernel<3>) : launch(layernorm_bwd_kernel<4>);
  if (lst != MR_OK) return lst;
  int32_t st = after_launch();
  if (st != MR_OK) return st;
  layernorm_bwd_reduce_kernel<<<ceil_div(2 * d->H, 32), 256, 0, stream>>>(
      static_cast<const float*>(workspace), static_cast<float*>(out[1]), static_cast<float*>(out[2]), nblk, d->H);
  return after_launch();
}

extern "C" int32_t mr_gelu(mr_stream_t stream_, const MrElementwiseDesc* d, const void* const* in, void* const* out, void*,
                           size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n >= 0 && d->n % 8 == 0 && (d->mode == 0 || (d->mode == 1 && in[1])), MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(in[0]) && aligned16(out[0]) && (!in[1] || aligned16(in[1])), MR_ERR_MISALIGNED);
  if (d->n == 0) return MR_OK;
  const int64_t nvec = d->n / 8;
  int64_t blocks = ceil_div64(nvec, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gelu_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]), static_cast<bf16*>(out[0]), nvec, d->mode);
  return after_launch();
}

extern "C" size_t mr_colsum_workspace_bytes(const MrColsumDesc* d) {
  if (!d || d->rows <= 0 || d->cols <= 0) return 0;
  const int nblk = ceil_div(d->rows, 512);
  return static_cast<size_t>(nblk) * d->cols * sizeof(float);
}

extern "C" int32_t mr_colsum(mr_stream_t stream_, const MrColsumDesc* d, const void* const* in, void* const* out,
                             void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows > 0 && d->cols > 0 && d->cols % 8 == 0 && d->ld % 8 == 0 && d->ld >= d->cols, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(workspace && workspace_bytes >= mr_colsum_workspace_bytes(d), MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(in[0]), MR_ERR_MISALIGNED);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int nblk = ceil_div(d->rows, 512);
  dim3 grid(ceil_div(d->cols / 8, 32), nblk);
  colsum_partial_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(in[0]), static_cast<float*>(workspace),
                                                  d->rows, d->cols, d->ld, 512);
  int32_t st = after_launch();
  if (st != MR_OK) return st;
  colsum_reduce_kernel<<<ceil_div(d->cols, 256), 256, 0, stream>>>(static_cast<const float*>(workspace),
                                                                   static_cast<float*>(out[0]), nblk, d->cols);
  return after_launch();
}

extern "C" int32_t mr_rotary_bwd(mr_stream_t stream_, const MrRotaryBwdDesc* d, const void* const* in, void* const* out,
                                 void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && d->hidden > 0 && d->hidden % 64 == 0 && d->rot_rows > 0, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(out[0]) && aligned16(in[0]), MR_ERR_MISALIGNED);
  if (d->rows == 0) return MR_OK;
  const int64_t total = static_cast<int64_t>(d->rows) * (d->hidden / 8 + d->hidden / 16);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  rotary_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<bf16*>(out[0]), static_cast<const float*>(in[0]), d->rows, d->hidden, d->rot_rows, d->q_scale);
  return after_launch();
}

namespace {
struct LseBwdCarve {
  size_t logits, dl_a, dl_b, y2, x2, total;
};
LseBwdCarve lse_bwd_carve(const MrContrastiveDesc* d) {
  const size_t ld = (static_cast<size_t>(d->n_all) + 7) / 8 * 8, n = d->n, H = d->H;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  LseBwdCarve c;
  c.logits = 0;
  c.dl_a = c.logits + al(n * ld * sizeof(float));
  c.dl_b = c.dl_a + al(n * 2 * ld * sizeof(bf16));
  c.y2 = c.dl_b + al(2 * n * ld * sizeof(bf16));
  c.x2 = c.y2 + al(2 * ld * H * sizeof(bf16));
  c.total = c.x2 + al(2 * n * H * sizeof(bf16));
  return c;
}
}  // namespace

extern "C" size_t mr_contrastive_lse_bwd_workspace_bytes(const MrContrastiveDesc* d) {
  if (!d || d->n <= 0 || d->n_all <= 0 || d->H <= 0) return 0;
  return lse_bwd_carve(d).total;
}

// One direction of the loss: dx = w (softmax(x y_all^T) y_all - y_pos'), dy_all = w (softmax^T x) with the positive
// columns corrected, w = upstream / (2 n).  in[0] x, in[1] y_all, in[2] lse (from the forward); out[0] dx f32 [n, H],
// out[1] dy_all f32 [n_all, H].  d->pos_off = column of row 0's positive in y_all.
extern "C" int32_t mr_contrastive_lse_bwd(mr_stream_t stream_, const MrContrastiveDesc* d, const void* const* in,
                                          void* const* out, void* workspace, size_t workspace_bytes) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && in[2] && out[0] && out[1], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n > 0 && d->n_all > 0 && d->H > 0 && d->H % 8 == 0 && d->pos_off >= 0 && d->pos_off + d->n <= d->n_all,
             MR_ERR_BAD_SHAPE);
  MR_REQUIRE(workspace && workspace_bytes >= mr_contrastive_lse_bwd_workspace_bytes(d), MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(workspace), MR_ERR_MISALIGNED);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int ld = (d->n_all + 7) / 8 * 8;
  const LseBwdCarve c = lse_bwd_carve(d);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* logits = reinterpret_cast<float*>(ws + c.logits);
  bf16* dl_a = reinterpret_cast<bf16*>(ws + c.dl_a);
  bf16* dl_b = reinterpret_cast<bf16*>(ws + c.dl_b);
  bf16* y2 = reinterpret_cast<bf16*>(ws + c.y2);
  bf16* x2 = reinterpret_cast<bf16*>(ws + c.x2);
  const size_t ybytes = static_cast<size_t>(d->n_all) * d->H * sizeof(bf16), xbytes = static_cast<size_t>(d->n) * d->H * sizeof(bf16);
  cudaError_t e = cudaMemsetAsync(y2, 0, static_cast<size_t>(2) * ld * d->H * sizeof(bf16), stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(y2, in[1], ybytes, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(y2 + static_cast<size_t>(ld) * d->H, in[1], ybytes, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(x2, in[0], xbytes, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(x2 + static_cast<size_t>(d->n) * d->H, in[0], xbytes, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return cuda_status(e);
  // logits = x . y_all^T (recomputed: cheaper than keeping [n, n_all] fp32 alive since the forward)
  MrGemmDesc g{};
  g.M = d->n;
  g.N = d->n_all;
  g.K = d->H;
  g.lda = d->H;
  g.ldd = ld;
  g.epilogue = MR_EPI_OUT_F32;
  {
    const void* gin[7] = {in[0], in[1], nullptr, nullptr, nullptr, nullptr, nullptr};
    void* gout[2] = {logits, nullptr};
    int32_t st = mr_gemm_epilogue(stream_, &g, gin, gout, nullptr, 0);
    if (st != MR_OK) return st;
  }
  contrastive_dlogits_kernel<<<ceil_div(d->n, kWarps), kThreadsRow, 0, stream>>>(
      logits, static_cast<const float*>(in[2]), dl_a, dl_b, d->n, d->n_all, ld, d->pos_off, d->grad_scale);
  int32_t st = after_launch();
  if (st != MR_OK) return st;
  // dx[n, H] = (hi | lo)[n, 2 ld] . [Y; Y][2 ld, H]      (dgrad form: B given as [K, N])
  MrGemmDesc gx{};
  gx.M = d->n;
  gx.N = d->H;
  gx.K = 2 * ld;
  gx.lda = 2 * ld;
  gx.ldd = d->H;
  gx.ldb = d->H;
  gx.trans_b = 1;
  gx.epilogue = MR_EPI_OUT_F32;
  {
    const void* gin[7] = {dl_a, y2, nullptr, nullptr, nullptr, nullptr, nullptr};
    void* gout[2] = {out[0], nullptr};
    st = mr_gemm_epilogue(stream_, &gx, gin, gout, nullptr, 0);
    if (st != MR_OK) return st;
  }
  // dy_all[n_all, H] = [hi; lo]^T [n_all, 2 n] . [X; X][2 n, H]   (wgrad form: both operands given as [K, *])
  MrGemmDesc gy{};
  gy.M = d->n_all;
  gy.N = d->H;
  gy.K = 2 * d->n;
  gy.lda = ld;
  gy.ldd = d->H;
  gy.ldb = d->H;
  gy.trans_a = 1;
  gy.trans_b = 1;
  gy.epilogue = MR_EPI_OUT_F32;
  if (d->dy_seq_in > 0) {   // rows go straight into the step's packed rank-major reduce-scatter buffer
    MR_REQUIRE(d->n_all % d->dy_seq_in == 0 && d->dy_seq_out >= d->dy_seq_in + d->dy_row_off && d->dy_row_off >= 0,
               MR_ERR_BAD_SHAPE);
    gy.seq_in = d->dy_seq_in;
    gy.seq_out = d->dy_seq_out;
    gy.row_off = d->dy_row_off;
  }
  const void* gin[7] = {dl_b, x2, nullptr, nullptr, nullptr, nullptr, nullptr};
  void* gout[2] = {out[1], nullptr};
  return mr_gemm_epilogue(stream_, &gy, gin, gout, nullptr, 0);
}

extern "C" int32_t mr_unit_normalize_bwd(mr_stream_t stream_, const MrNormalizeDesc* d, const void* const* in,
                                         void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && d->H > 0, MR_ERR_BAD_SHAPE);
  if (d->rows == 0) return MR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = ceil_div(d->rows, kWarps);
  const int32_t* ridx = static_cast<const int32_t*>(in[2]);
  float* dx = static_cast<float*>(out[0]);
  const float* sdev = d->scale_from_device ? static_cast<const float*>(in[3]) : nullptr;
  MR_REQUIRE(!d->scale_from_device || sdev, MR_ERR_NULL_POINTER);
  // in_f32: dtype of x (the forward's input); out_f32: dtype of dy (the forward's output)
  if (d->in_f32 && d->out_f32)
    unit_normalize_bwd_kernel<true, true><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale, sdev);
  else if (d->in_f32)
    unit_normalize_bwd_kernel<true, false><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale, sdev);
  else if (d->out_f32)
    unit_normalize_bwd_kernel<false, true><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale, sdev);
  else
    unit_normalize_bwd_kernel<false, false><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale, sdev);
  return after_launch();
}

extern "C" int32_t mr_segment_pool_bwd(mr_stream_t stream_, const MrSegmentPoolDesc* d, const void* const* in,
                                       void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && in[2] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->L > 0 && d->H > 0 && d->S > 0, MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  const int64_t warps = static_cast<int64_t>(d->B) * d->L;
  segment_pool_bwd_kernel<<<static_cast<unsigned>(ceil_div64(warps, kWarps)), kThreadsRow, 0,
                            static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint8_t*>(in[0]), static_cast<const int32_t*>(in[1]), static_cast<const float*>(in[2]),
      static_cast<bf16*>(out[0]), *d);
  return after_launch();
}

extern "C" int32_t mr_embed_assemble_bwd(mr_stream_t stream_, const MrAssembleDesc* d, const void* const* in,
                                         void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[2], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->L > 0 && d->H > 0 && d->T >= d->L + d->n_vis && d->vocab > 0, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->cls_offset == 0 || d->cls_offset == 1, MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  const int64_t rows = static_cast<int64_t>(d->B) * (d->T + d->cls_offset);
  embed_assemble_bwd_kernel<<<static_cast<unsigned>(ceil_div64(rows, kWarps)), kThreadsRow, 0,
                              static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const int32_t*>(in[0]), d->n_aud > 0 ? static_cast<const int32_t*>(in[1]) : nullptr,
      static_cast<const bf16*>(in[2]), static_cast<float*>(out[0]), d->n_aud > 0 ? static_cast<float*>(out[1]) : nullptr,
      d->n_vis > 0 ? static_cast<float*>(out[2]) : nullptr, *d);
  return after_launch();
}

extern "C" int32_t mr_window_attention_bwd(mr_stream_t stream_, const MrWindowDesc* d, const void* const* in,
                                           void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && in[2] && out[0] && out[1], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->T > 0 && d->H > 0 && d->H % 64 == 0 && d->win > 0 && d->win <= 8 && d->n_win > 0 &&
                 (d->mode == 0 || (d->mode == 1 && d->win == 4)),
             MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  const int64_t warps = static_cast<int64_t>(d->B) * d->n_win * (d->H / 64);
  window_attention_bwd_kernel<<<static_cast<unsigned>(ceil_div64(warps, kWarps)), kThreadsRow, 0,
                                static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]), static_cast<const bf16*>(in[2]),
      static_cast<bf16*>(out[0]), static_cast<bf16*>(out[1]), *d);
  return after_launch();
}

extern "C" int32_t mr_window_mean_bwd(mr_stream_t stream_, const MrWindowDesc* d, const void* const* in, void* const* out,
                                      void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->T > 0 && d->H > 0 && d->H % 8 == 0 && d->win > 0 && d->n_win > 0, MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  const int64_t total = static_cast<int64_t>(d->B) * d->n_win * d->win * (d->H / 8);
  window_mean_bwd_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const bf16*>(in[0]), static_cast<bf16*>(out[0]), *d);
  return after_launch();
}

extern "C" int32_t mr_add(mr_stream_t stream_, const MrElementwiseDesc* d, const void* const* in, void* const* out, void*,
                          size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n >= 0 && d->n % 8 == 0 && (d->mode == 0 || d->mode == 1), MR_ERR_BAD_SHAPE);
  if (d->n == 0) return MR_OK;
  const int64_t nvec = d->n / 8;
  int64_t blocks = ceil_div64(nvec, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (d->mode == 1)
    add_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(in[0]), in[1],
                                                                        static_cast<bf16*>(out[0]), nvec);
  else
    add_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(in[0]), in[1],
                                                                         static_cast<bf16*>(out[0]), nvec);
  return after_launch();
}

extern "C" int32_t mr_adamw(mr_stream_t stream_, const MrAdamWDesc* d, const void* const* in, void* const* out, void*,
                            size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0] && out[1] && out[2], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n >= 0, MR_ERR_BAD_SHAPE);
  if (d->n == 0) return MR_OK;
  int64_t blocks = ceil_div64(d->n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const float*>(in[0]), static_cast<float*>(out[0]), static_cast<float*>(out[1]),
      static_cast<float*>(out[2]), static_cast<bf16*>(out[3]), *d);
  return after_launch();
}

extern "C" int32_t mr_adamw_multi(mr_stream_t stream_, const MrAdamWMultiDesc* d, const void* const* in, void* const* out,
                                  void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && out[0] && out[1] && out[2], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n_chunks >= 0 && d->bias_corr1 > 0.f && d->bias_corr2 > 0.f, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(in[0]) && aligned16(out[0]) && aligned16(out[1]) && aligned16(out[2]), MR_ERR_MISALIGNED);
  if (d->n_chunks == 0) return MR_OK;
  const MrAdamWChunk* table = static_cast<const MrAdamWChunk*>(in[1]);
  const float* hyper = d->hyper_from_device ? static_cast<const float*>(in[2]) : nullptr;
  MR_REQUIRE(!d->hyper_from_device || hyper, MR_ERR_NULL_POINTER);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned grid = static_cast<unsigned>(d->n_chunks);
  if (d->moments_bf16) {
    if (d->grad_bf16)
      adamw_multi_kernel<true, true><<<grid, 256, 0, stream>>>(in[0], table, static_cast<float*>(out[0]), out[1], out[2], *d, hyper);
    else
      adamw_multi_kernel<true, false><<<grid, 256, 0, stream>>>(in[0], table, static_cast<float*>(out[0]), out[1], out[2], *d, hyper);
  } else {
    if (d->grad_bf16)
      adamw_multi_kernel<false, true><<<grid, 256, 0, stream>>>(in[0], table, static_cast<float*>(out[0]), out[1], out[2], *d, hyper);
    else
      adamw_multi_kernel<false, false><<<grid, 256, 0, stream>>>(in[0], table, static_cast<float*>(out[0]), out[1], out[2], *d, hyper);
  }
  return after_launch();
}
