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
__global__ void __launch_bounds__(kThreadsRow)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                     const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ partial, int rows, int H,
                     float eps) {
  extern __shared__ float red[];   // [kWarps][2][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = H / 8;
  const float inv_h = 1.0f / static_cast<float>(H);
  float dg[kMaxVec][8], db[kMaxVec][8], gm[kMaxVec][8];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dg[i][j] = 0.f;
      db[i][j] = 0.f;
      gm[i][j] = v < nvec ? gamma[v * 8 + j] : 0.f;
    }
  }
  for (int row = blockIdx.x * kWarps + warp; row < rows; row += gridDim.x * kWarps) {
    const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * H);
    const uint4* ds = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * H);
    float xv[kMaxVec][8], dv[kMaxVec][8];
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        unpack8b(__ldg(xs + v), xv[i]);
        unpack8b(__ldg(ds + v), dv[i]);
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
    for (int i = 0; i < kMaxVec; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[i][j] - mean) * rstd;
          const float g = dv[i][j] * gm[i][j];
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
    for (int i = 0; i < kMaxVec; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dv[i][j] - sg - xv[i][j] * sgx);
        if (dres != nullptr) {
          float r[8];
          unpack8b(__ldg(reinterpret_cast<const uint4*>(dres + static_cast<size_t>(row) * H) + v), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * H)[v] = pack8b(o);
      }
    }
  }
  // block reduction of the parameter-gradient partials
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
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

__global__ void layernorm_bwd_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int nblk, int H) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * H) return;
  float a = 0.f;
  for (int b = 0; b < nblk; ++b) a += partial[static_cast<size_t>(b) * 2 * H + c];
  if (c < H)
    dgamma[c] = a;
  else
    dbeta[c - H] = a;
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
// element-wise factors applied to d(qkv), in place.  One thread per 8 columns.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rotary_bwd_kernel(bf16* __restrict__ dqkv, const float* __restrict__ rot, int64_t rows, int hidden, int rot_rows,
                  float q_scale) {
  const int vec_per_row = 3 * hidden / 8;
  const int64_t total = rows * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    if (c >= 2 * hidden) continue;              // v columns: untouched
    const bool is_q = c < hidden;
    const int in_head = c & 63;
    if (!is_q && in_head >= 32) continue;       // k, upper half of the head: factor 1
    uint4* cell = reinterpret_cast<uint4*>(dqkv + r * 3 * hidden + c);
    float f[8];
    unpack8b(*cell, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = is_q ? q_scale : 1.0f;
      if (in_head < 32) s *= rot[static_cast<size_t>(r % rot_rows) * 32 + in_head + j];
      f[j] *= s;
    }
    *cell = pack8b(f);
  }
}

// ---------------------------------------------------------------------------------------------
// Contrastive loss backward, step 1: dlogits[i, j] = w (exp(logit[i, j] - lse[i]) - [j == pos(i)]) as bf16
// (pretrain_model.py:293-295; w = 1 / (2 n), the weight of one direction's mean).  The positive of row i is column
// pos_off + i of y_all (this rank's block of the rank-major all-gather).  One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsRow)
contrastive_dlogits_kernel(const float* __restrict__ logits, const float* __restrict__ lse, bf16* __restrict__ dl, int n,
                           int n_all, int ld_logits, int ld_out, int pos_off, float w) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* lr = logits + static_cast<size_t>(row) * ld_logits;
  const float l = lse[row];
  bf16* o = dl + static_cast<size_t>(row) * ld_out;
  for (int c = lane; c < ld_out; c += 32) {
    float v = 0.f;
    if (c < n_all) {
      v = __expf(lr[c] - l);
      if (c == pos_off + row) v -= 1.0f;
      v *= w;
    }
    o[c] = __float2bfloat16(v);
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
                          float* __restrict__ dx, int rows, int H, float scale) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
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

}  // namespace
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
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  layernorm_bwd_kernel<<<nblk, kThreadsRow, smem, stream>>>(
      static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]), static_cast<const float*>(in[2]),
      static_cast<const bf16*>(in[3]), static_cast<bf16*>(out[0]), static_cast<float*>(workspace), d->rows, d->H, d->eps);
  int32_t st = after_launch();
  if (st != MR_OK) return st;
  layernorm_bwd_reduce_kernel<<<ceil_div(2 * d->H, 256), 256, 0, stream>>>(
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
  MR_REQUIRE(aligned16(out[0]), MR_ERR_MISALIGNED);
  if (d->rows == 0) return MR_OK;
  const int64_t total = static_cast<int64_t>(d->rows) * (3 * d->hidden / 8);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  rotary_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<bf16*>(out[0]), static_cast<const float*>(in[0]), d->rows, d->hidden, d->rot_rows, d->q_scale);
  return after_launch();
}

extern "C" size_t mr_contrastive_lse_bwd_workspace_bytes(const MrContrastiveDesc* d) {
  if (!d) return 0;
  const size_t ld = (static_cast<size_t>(d->n_all) + 7) / 8 * 8;
  return static_cast<size_t>(d->n) * ld * (sizeof(float) + sizeof(bf16));
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int ld = (d->n_all + 7) / 8 * 8;
  float* logits = static_cast<float*>(workspace);
  bf16* dl = reinterpret_cast<bf16*>(logits + static_cast<size_t>(d->n) * ld);
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
      logits, static_cast<const float*>(in[2]), dl, d->n, d->n_all, ld, ld, d->pos_off, d->grad_scale);
  int32_t st = after_launch();
  if (st != MR_OK) return st;
  // dx[n, H] = dl[n, n_all] . y_all[n_all, H]            (dgrad form: B given as [K, N])
  MrGemmDesc gx{};
  gx.M = d->n;
  gx.N = d->H;
  gx.K = d->n_all;      // the < 8 padding columns of dl are beyond K: clipped by the tensor map
  gx.lda = ld;
  gx.ldd = d->H;
  gx.ldb = d->H;
  gx.trans_b = 1;
  gx.epilogue = MR_EPI_OUT_F32;
  {
    const void* gin[7] = {dl, in[1], nullptr, nullptr, nullptr, nullptr, nullptr};
    void* gout[2] = {out[0], nullptr};
    st = mr_gemm_epilogue(stream_, &gx, gin, gout, nullptr, 0);
    if (st != MR_OK) return st;
  }
  // dy_all[n_all, H] = dl^T[n_all, n] . x[n, H]            (wgrad form: both operands given as [K, *])
  MrGemmDesc gy{};
  gy.M = d->n_all;
  gy.N = d->H;
  gy.K = d->n;
  gy.lda = ld;
  gy.ldd = d->H;
  gy.ldb = d->H;
  gy.trans_a = 1;
  gy.trans_b = 1;
  gy.epilogue = MR_EPI_OUT_F32;
  const void* gin[7] = {dl, in[0], nullptr, nullptr, nullptr, nullptr, nullptr};
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
  // in_f32: dtype of x (the forward's input); out_f32: dtype of dy (the forward's output)
  if (d->in_f32 && d->out_f32)
    unit_normalize_bwd_kernel<true, true><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale);
  else if (d->in_f32)
    unit_normalize_bwd_kernel<true, false><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale);
  else if (d->out_f32)
    unit_normalize_bwd_kernel<false, true><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale);
  else
    unit_normalize_bwd_kernel<false, false><<<grid, kThreadsRow, 0, stream>>>(in[0], in[1], ridx, dx, d->rows, d->H, d->scale);
  return after_launch();
}
