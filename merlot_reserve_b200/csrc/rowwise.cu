// HBM-bound row kernels of the path: LayerNorm, casts, CLS fill, window mean / window attention
// (the attention pools), joint-input assembly, rotary scale tables, unit_normalize, segment pooling,
// row-LSE and row dot.  All are "one warp per row (or per row-head)", 16-byte vector accesses,
// warp-shuffle reductions, fp32 arithmetic.
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kRowThreads = kWarpsPerBlock * 32;
constexpr float kPi = 3.14159265358979323846f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]);
  u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]);
  u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm.  kVec = uint4 vectors (8 bf16) per lane; H = 256 * kVec.  kVec == 0: generic H % 8 == 0.
// ---------------------------------------------------------------------------------------------
template <int kVec>
__global__ void __launch_bounds__(kRowThreads)
layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 bf16* __restrict__ y, float2* __restrict__ stats, MrLayerNormDesc d) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= d.rows) return;
  const int in_row = d.seq_out == 0 ? row : (row / d.seq_out) * d.seq_in + (row % d.seq_out) + d.row_off;
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<size_t>(in_row) * d.ldx);
  uint4* dst = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * d.H);
  const float inv_h = 1.0f / static_cast<float>(d.H);

  if constexpr (kVec > 0) {
    uint4 raw[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) raw[i] = __ldg(src + lane + 32 * i);
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += f[j];
        s2 += f[j] * f[j];
      }
    }
    s = warp_sum(s);
    s2 = warp_sum(s2);
    const float mean = s * inv_h;
    const float var = s2 * inv_h - mean * mean;
    const float rstd = rsqrtf(var + d.eps);
    float os = 0.f, os2 = 0.f;   // statistics of the rounded output (folded LayerNorm of the next GEMM)
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      float f[8];
      unpack8(raw[i], f);
      const int c = (lane + 32 * i) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * (rstd * g[j]) + b[j];
      const uint4 o = pack8(f);
      dst[lane + 32 * i] = o;
      if (stats != nullptr) {
        unpack8(o, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          os += f[j];
          os2 = fmaf(f[j], f[j], os2);
        }
      }
    }
    if (stats != nullptr) {
      os = warp_sum(os);
      os2 = warp_sum(os2);
      const int parts = d.H >> 6;
      if (lane < parts)
        stats[static_cast<size_t>(lane) * d.rows + row] = lane == 0 ? make_float2(os, os2) : make_float2(0.f, 0.f);
    }
  } else {
    const int nvec = d.H / 8;
    float s = 0.f, s2 = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      unpack8(__ldg(src + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += f[j];
        s2 += f[j] * f[j];
      }
    }
    s = warp_sum(s);
    s2 = warp_sum(s2);
    const float mean = s * inv_h;
    const float var = s2 * inv_h - mean * mean;
    const float rstd = rsqrtf(var + d.eps);
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      unpack8(__ldg(src + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * (rstd * gamma[v * 8 + j]) + beta[v * 8 + j];
      const uint4 o = pack8(f);
      dst[v] = o;
    }
    if (stats != nullptr) {
      // generic-H path: second pass over the rounded output just written by this warp
      __syncwarp();
      float os = 0.f, os2 = 0.f;
      for (int v = lane; v < nvec; v += 32) {
        float f[8];
        unpack8(dst[v], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          os += f[j];
          os2 = fmaf(f[j], f[j], os2);
        }
      }
      os = warp_sum(os);
      os2 = warp_sum(os2);
      const int parts = d.H >> 6;
      for (int q = lane; q < parts; q += 32)
        stats[static_cast<size_t>(q) * d.rows + row] = q == 0 ? make_float2(os, os2) : make_float2(0.f, 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 cast with zero column padding.  One thread per 8 output columns.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_pad_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t rows, int cols, int ld_out, int sanitize) {
  const int vec_per_row = ld_out / 8;
  const int64_t total = rows * vec_per_row;
  const bool fast = (cols % 4) == 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    float f[8];
    const float* src = x + r * cols + c;
    if (fast && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
      f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (c + j < cols) ? __ldg(src + j) : 0.f;
    }
    if (sanitize) {   // jnp.nan_to_num: NaN -> 0, +-inf -> +-largest finite value of the target type (bf16)
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] != f[j]) ? 0.f : fminf(fmaxf(f[j], -3.3895314e38f), 3.3895314e38f);
    }
    *reinterpret_cast<uint4*>(y + r * ld_out + c) = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// uint8 NHWC frames -> bf16 patch rows (convert_image_dtype + space_to_depth + cast, preprocess.py:471,478-479).
// One thread per 8 output features: for a 16 x 16 patch the 48 features of one patch row (px, c) are 48 contiguous
// bytes of the image row, so every thread reads one aligned 8-byte word and writes one 16-byte vector.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ frames, bf16* __restrict__ y, int64_t n_patches, int grid_h, int grid_w) {
  constexpr int kP = 16, kFeat = kP * kP * 3, kVec = kFeat / 8;   // 768 features, 96 vectors per patch
  const int64_t total = n_patches * kVec;
  const int img_w3 = grid_w * kP * 3;                             // bytes per image row
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t patch = i / kVec;
    const int f0 = static_cast<int>(i - patch * kVec) * 8;        // first feature of this vector
    const int py = f0 / (kP * 3), rest = f0 - py * (kP * 3);      // rest = px * 3 + c, a multiple of 8
    const int bx = static_cast<int>(patch % grid_w);
    const int64_t t = patch / grid_w;
    const int by = static_cast<int>(t % grid_h);
    const int64_t b = t / grid_h;
    const uint8_t* src = frames + ((b * grid_h + by) * kP + py) * static_cast<int64_t>(img_w3) + bx * (kP * 3) + rest;
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(src));
    float f[8];
    const float s = 1.0f / 255.0f;                                // float32(1/255), as TF's convert_image_dtype
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[j] = static_cast<float>((w.x >> (8 * j)) & 0xffu) * s;
      f[4 + j] = static_cast<float>((w.y >> (8 * j)) & 0xffu) * s;
    }
    *reinterpret_cast<uint4*>(y + patch * kFeat + f0) = pack8(f);
  }
}

__global__ void fill_cls_kernel(const float* __restrict__ cls, bf16* __restrict__ x, int B, int T, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, c = idx % H;
  x[static_cast<size_t>(b) * T * H + c] = __float2bfloat16(cls[c]);
}

// ---------------------------------------------------------------------------------------------
// Attention pooling helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int window_token(const MrWindowDesc& d, int w, int j) {
  if (d.mode == 0) return w * d.win + j;
  const int w2 = d.grid_w / 2;
  const int y = w / w2, x = w % w2;
  const int dy = j >> 1, dx = j & 1;
  return (2 * y + dy) * d.grid_w + 2 * x + dx;
}

// out[b*n_win + w, :] = mean_j x[b, 1 + tok(w, j), :]; one thread per 8 columns.
__global__ void __launch_bounds__(256)
window_mean_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, MrWindowDesc d) {
  const int vec_per_row = d.H / 8;
  const int64_t total = static_cast<int64_t>(d.B) * d.n_win * vec_per_row;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % vec_per_row);
  const int64_t bw = i / vec_per_row;
  const int b = static_cast<int>(bw / d.n_win), w = static_cast<int>(bw % d.n_win);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < d.win; ++j) {
    const int t = 1 + window_token(d, w, j);
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * d.T + t) * d.H) + v), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  const float inv = 1.0f / static_cast<float>(d.win);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] *= inv;
  reinterpret_cast<uint4*>(out + static_cast<size_t>(bw) * d.H)[v] = pack8(acc);
}

// One warp per (window, head): lane owns 2 of the 64 head dims.
__global__ void __launch_bounds__(kRowThreads)
window_attention_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kv, bf16* __restrict__ out,
                        MrWindowDesc d) {
  const int nh = d.H / 64;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(d.B) * d.n_win * nh;
  if (gw >= total) return;
  const int h = static_cast<int>(gw % nh);
  const int64_t bw = gw / nh;
  const int b = static_cast<int>(bw / d.n_win), w = static_cast<int>(bw % d.n_win);
  const uint32_t qq = __ldg(reinterpret_cast<const uint32_t*>(q + static_cast<size_t>(bw) * d.H + h * 64) + lane);
  const float q0 = bf16_lo(qq) * 0.125f, q1 = bf16_hi(qq) * 0.125f;
  float s[8], v0[8], v1[8];
  float m = -INFINITY;
  for (int j = 0; j < d.win; ++j) {
    const size_t row = static_cast<size_t>(b) * d.T + 1 + window_token(d, w, j);
    const bf16* kr = kv + row * (2 * d.H) + h * 64;
    const uint32_t kk = __ldg(reinterpret_cast<const uint32_t*>(kr) + lane);
    const uint32_t vv = __ldg(reinterpret_cast<const uint32_t*>(kr + d.H) + lane);
    s[j] = warp_sum(q0 * bf16_lo(kk) + q1 * bf16_hi(kk));
    v0[j] = bf16_lo(vv);
    v1[j] = bf16_hi(vv);
    m = fmaxf(m, s[j]);
  }
  float l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < d.win; ++j) {
    const float p = expf(s[j] - m);
    l += p;
    o0 += p * v0[j];
    o1 += p * v1[j];
  }
  const float inv = 1.0f / l;
  reinterpret_cast<uint32_t*>(out + static_cast<size_t>(bw) * d.H + h * 64)[lane] = pack_bf16(o0 * inv, o1 * inv);
}

// ---------------------------------------------------------------------------------------------
// Rotary scale table.  theta_i = coord[axis(i)] * f_{j(i)} * pi, i = axis * d + j, d = 16 / ndims,
// f_j = 2^(j * log2(5) / (d - 1));  S[2i] = sin - cos, S[2i+1] = sin + cos.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rot_freq(int j, int dexp) {
  return exp2f(static_cast<float>(j) * (2.3219280948873622f / static_cast<float>(dexp - 1)));
}

__global__ void rotary_table_kernel(const float* __restrict__ coords, float* __restrict__ rot, int rows, int ndims) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 16) return;
  const int r = idx >> 4, i = idx & 15;
  const int dexp = 16 / ndims;
  const int axis = i / dexp, j = i % dexp;
  const float theta = coords[r * ndims + axis] * rot_freq(j, dexp) * kPi;
  float sn, cs;
  sincosf(theta, &sn, &cs);
  rot[r * 32 + 2 * i] = sn - cs;
  rot[r * 32 + 2 * i + 1] = sn + cs;
}

// centred coordinate of get_rotary_coordinates (modeling.py:21-35): [-L/2 .. -1, 1 .. L - L/2]
__device__ __forceinline__ float centred_coord(int i, int L) {
  const int sl0 = L / 2;
  return i < sl0 ? static_cast<float>(i - sl0) : static_cast<float>(i - sl0 + 1);
}

// ---------------------------------------------------------------------------------------------
// Joint input assembly: one warp per output row (b, t).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowThreads)
embed_assemble_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ seg_idx,
                      const int32_t* __restrict__ aud_ptr, const bf16* __restrict__ emb,
                      const bf16* __restrict__ audio, const bf16* __restrict__ vision,
                      const bf16* __restrict__ tok_emb, bf16* __restrict__ x, float* __restrict__ rot,
                      uint8_t* __restrict__ valid, MrAssembleDesc d) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int t_out = d.T + d.cls_offset;
  if (row >= static_cast<int64_t>(d.B) * t_out) return;
  const int b = static_cast<int>(row / t_out), t = static_cast<int>(row % t_out) - d.cls_offset;
  const int nvec = d.H / 8;
  const uint4* src = nullptr;
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  uint8_t ok = 0;
  if (t < 0) {
    // CLS slot: the embedding itself is written by mr_fill_cls; coordinates are all zero
    ok = 1;
    if (lane < 16) reinterpret_cast<float2*>(rot + static_cast<size_t>(row) * 32)[lane] = make_float2(-1.f, 1.f);
    if (lane == 0) valid[row] = ok;
    return;
  } else if (t < d.L) {
    const int32_t* trow = tokens + static_cast<size_t>(b) * d.L;
    const int tok = trow[t];
    ok = tok != 0;
    bool from_audio = false;
    if (aud_ptr != nullptr && audio != nullptr && tok == 5) {  // AUDIOSPAN (lowercase_encoder.py:14)
      // subpos = (max(cumsum(tokens == AUDIOSPAN)[t] - 1, 0)) % audio_token_length
      int cnt = 0;
      for (int base = 0; base <= t; base += 32) {
        const int i = base + lane;
        const bool is_a = (i <= t) && (trow[i] == 5);
        cnt += __popc(__ballot_sync(0xffffffffu, is_a));
      }
      const int sub = max(cnt - 1, 0) % d.audio_token_length;
      const int ptr = max(aud_ptr[static_cast<size_t>(b) * d.L + t], 0);
      const int ab = d.aud_div > 1 ? b / d.aud_div : b;
      src = reinterpret_cast<const uint4*>(
          audio + ((static_cast<size_t>(ab) * d.n_aud + ptr) * d.audio_token_length + sub) * d.H);
      from_audio = true;
    }
    if (!from_audio) {
      if (tok_emb != nullptr)
        src = reinterpret_cast<const uint4*>(tok_emb + (static_cast<size_t>(b) * d.L + t) * d.H);
      else
        src = reinterpret_cast<const uint4*>(emb + static_cast<size_t>(min(max(tok, 0), d.vocab - 1)) * d.H);
    }
    if (seg_idx) {
      int sv = seg_idx[static_cast<size_t>(b) * d.L + t];
      if (d.seg_div > 1) sv = sv >= 0 ? sv / d.seg_div : -((-sv + d.seg_div - 1) / d.seg_div);   // floor division
      c[2] = static_cast<float>(sv) / 16.0f;
    }
    c[3] = static_cast<float>(1 + t) / 1024.0f;
  } else if (t < d.L + d.n_vis) {
    const int v = t - d.L;
    ok = 1;
    const int vb = d.vis_tile > 0 ? (b / (d.vis_group * d.vis_tile)) * d.vis_group + b % d.vis_group : b;
    src = reinterpret_cast<const uint4*>(vision + (static_cast<size_t>(vb) * d.n_vis + v) * d.H);
    const int per = d.pool_h * d.pool_w;
    const int pos = v % per;
    const float base_scale = 1.0f / (static_cast<float>(max(d.pool_h, d.pool_w)) + 1.0f);
    c[0] = base_scale * centred_coord(pos / d.pool_w, d.pool_h);
    c[1] = base_scale * centred_coord(pos % d.pool_w, d.pool_w);
    c[2] = static_cast<float>(v / per) / 16.0f;
  }
  uint4* dst = reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * d.H);
  for (int v = lane; v < nvec; v += 32) dst[v] = src ? __ldg(src + v) : make_uint4(0, 0, 0, 0);
  if (lane < 16) {
    const int axis = lane >> 2, j = lane & 3;
    const float theta = c[axis] * rot_freq(j, 4) * kPi;
    float sn, cs;
    sincosf(theta, &sn, &cs);
    reinterpret_cast<float2*>(rot + static_cast<size_t>(row) * 32)[lane] = make_float2(sn - cs, sn + cs);
  }
  if (lane == 0) valid[row] = ok;
}

// ---------------------------------------------------------------------------------------------
// unit_normalize: one warp per row
// ---------------------------------------------------------------------------------------------
template <bool kInF32, bool kOutF32>
__global__ void __launch_bounds__(kRowThreads)
unit_normalize_kernel(const void* __restrict__ x_, const int32_t* __restrict__ row_idx, void* __restrict__ y_,
                      int rows, int H, float scale, const float* __restrict__ scale_dev) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  if (scale_dev != nullptr) scale *= __ldg(scale_dev);
  const size_t in_row = row_idx ? static_cast<size_t>(row_idx[row]) : static_cast<size_t>(row);
  float ss = 0.f;
  for (int c = lane; c < H; c += 32) {
    const float v = kInF32 ? static_cast<const float*>(x_)[in_row * H + c]
                           : __bfloat162float(static_cast<const bf16*>(x_)[in_row * H + c]);
    ss += v * v;
  }
  ss = warp_sum(ss);
  const float inv = scale / sqrtf(ss + 1e-5f);
  for (int c = lane; c < H; c += 32) {
    const float v = kInF32 ? static_cast<const float*>(x_)[in_row * H + c]
                           : __bfloat162float(static_cast<const bf16*>(x_)[in_row * H + c]);
    if (kOutF32)
      static_cast<float*>(y_)[static_cast<size_t>(row) * H + c] = v * inv;
    else
      static_cast<bf16*>(y_)[static_cast<size_t>(row) * H + c] = __float2bfloat16(v * inv);
  }
}

// ---------------------------------------------------------------------------------------------
// one_hot_pool as a segment gather: one warp per (b, s)
// ---------------------------------------------------------------------------------------------
template <bool kVF32>
__global__ void __launch_bounds__(kRowThreads)
segment_pool_kernel(const uint8_t* __restrict__ do_pool, const int32_t* __restrict__ idx, const void* __restrict__ v_,
                    float* __restrict__ pooled, float* __restrict__ count, MrSegmentPoolDesc d) {
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= d.B * d.S) return;
  const int b = gw / d.S, s = gw % d.S;
  float* dst = pooled + static_cast<size_t>(gw) * d.H;
  if (!d.accumulate)
    for (int c = lane; c < d.H; c += 32) dst[c] = 0.f;
  int cnt = 0;
  for (int base = 0; base < d.L; base += 32) {
    const int l = base + lane;
    const bool hit = l < d.L && do_pool[static_cast<size_t>(b) * d.L + l] && idx[static_cast<size_t>(b) * d.L + l] == s;
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    cnt += __popc(mask);
    while (mask) {
      const int src_l = base + __ffs(mask) - 1;
      mask &= mask - 1;
      const size_t off = (static_cast<size_t>(b) * d.L + src_l) * d.H;
      for (int c = lane; c < d.H; c += 32)
        dst[c] += kVF32 ? static_cast<const float*>(v_)[off + c] : __bfloat162float(static_cast<const bf16*>(v_)[off + c]);
    }
  }
  if (lane == 0) count[gw] = (d.accumulate ? count[gw] : 0.f) + static_cast<float>(cnt);
}

// ---------------------------------------------------------------------------------------------
// Row log-sum-exp over fp32 logits + positive logit: one warp per row
// ---------------------------------------------------------------------------------------------
// One block: out[0] = mean(lse - numer); out[1 + s] = sum((lse - numer) [src == s]) / (count_s + 1e-5).
__global__ void __launch_bounds__(256)
contrastive_reduce_kernel(const float* __restrict__ lse, const float* __restrict__ numer,
                          const int32_t* __restrict__ sources, float* __restrict__ out, int n) {
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // total, per-source sums [3], per-source counts [3]
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = lse[i] - numer[i];
    acc[0] += v;
    if (sources) {
      const int s = sources[i];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (s == k) {
          acc[1 + k] += v;
          acc[4 + k] += 1.f;
        }
      }
    }
  }
  __shared__ float red[7][8];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const float w = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[7];
    for (int k = 0; k < 7; ++k) {
      tot[k] = 0.f;
      for (int w = 0; w < 8; ++w) tot[k] += red[k][w];
    }
    out[0] = tot[0] / static_cast<float>(n);
    for (int k = 0; k < 3; ++k) out[1 + k] = tot[1 + k] / (tot[4 + k] + 1e-5f);
  }
}

__global__ void __launch_bounds__(kRowThreads)
row_dot_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const int32_t* __restrict__ row_idx,
               float* __restrict__ out, int rows, int H) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t src = row_idx ? static_cast<size_t>(row_idx[row]) : static_cast<size_t>(row);
  float dot = 0.f;
  for (int c = lane; c < H; c += 32) dot += __bfloat162float(x[src * H + c]) * w[c];
  dot = warp_sum(dot);
  if (lane == 0) out[row] = dot;
}

}  // namespace
}  // namespace mr

using namespace mr;

extern "C" int32_t mr_layernorm_fwd(mr_stream_t stream_, const MrLayerNormDesc* d, const void* const* in,
                                    void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && in[2] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && d->H > 0 && d->H % 8 == 0 && d->ldx % 8 == 0 && d->ldx >= d->H, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(in[0]) && aligned16(in[1]) && aligned16(in[2]) && aligned16(out[0]), MR_ERR_MISALIGNED);
  if (d->rows == 0) return MR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bf16* x = static_cast<const bf16*>(in[0]);
  const float* g = static_cast<const float*>(in[1]);
  const float* b = static_cast<const float*>(in[2]);
  bf16* y = static_cast<bf16*>(out[0]);
  float2* stats = d->emit_stats ? static_cast<float2*>(out[1]) : nullptr;
  if (d->emit_stats) MR_REQUIRE(stats && d->H % 64 == 0 && aligned16(stats), MR_ERR_BAD_SHAPE);
  const int grid = ceil_div(d->rows, kWarpsPerBlock);
  if (d->H == 768)
    layernorm_kernel<3><<<grid, kRowThreads, 0, stream>>>(x, g, b, y, stats, *d);
  else if (d->H == 1024)
    layernorm_kernel<4><<<grid, kRowThreads, 0, stream>>>(x, g, b, y, stats, *d);
  else
    layernorm_kernel<0><<<grid, kRowThreads, 0, stream>>>(x, g, b, y, stats, *d);
  return after_launch();
}

extern "C" int32_t mr_cast_pad(mr_stream_t stream_, const MrCastDesc* d, const void* const* in, void* const* out,
                               void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && d->cols > 0 && d->ld_out >= d->cols && d->ld_out % 8 == 0, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(in[0]) && aligned16(out[0]), MR_ERR_MISALIGNED);
  if (d->rows == 0) return MR_OK;
  const int64_t total = d->rows * (d->ld_out / 8);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > 148 * 64) blocks = 148 * 64;
  cast_pad_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const float*>(in[0]), static_cast<bf16*>(out[0]), d->rows, d->cols, d->ld_out, d->sanitize);
  return after_launch();
}

extern "C" int32_t mr_patchify_u8(mr_stream_t stream_, const MrPatchifyDesc* d, const void* const* in, void* const* out,
                                  void*, size_t) {
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->grid_h > 0 && d->grid_w > 0 && d->patch == 16, MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(aligned16(in[0]) && aligned16(out[0]), MR_ERR_MISALIGNED);
  const int64_t n_patches = static_cast<int64_t>(d->B) * d->grid_h * d->grid_w;
  int64_t blocks = ceil_div64(n_patches * 96, 256);
  if (blocks > 148 * 64) blocks = 148 * 64;
  patchify_u8_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint8_t*>(in[0]), static_cast<bf16*>(out[0]), n_patches, d->grid_h, d->grid_w);
  return after_launch();
}

extern "C" int32_t mr_fill_cls(mr_stream_t stream_, const MrFillClsDesc* d, const void* const* in, void* const* out,
                               void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->T > 0 && d->H > 0, MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  const int total = d->B * d->H;
  fill_cls_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const float*>(in[0]), static_cast<bf16*>(out[0]), d->B, d->T, d->H);
  return after_launch();
}

static int32_t check_window(const MrWindowDesc* d) {
  MR_REQUIRE(d->B >= 0 && d->T > 0 && d->H > 0 && d->H % 64 == 0, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->win > 0 && d->win <= 8 && d->n_win > 0, MR_ERR_BAD_SHAPE);
  if (d->mode == 0) {
    MR_REQUIRE(d->n_win * d->win + 1 <= d->T, MR_ERR_BAD_SHAPE);
  } else if (d->mode == 1) {
    MR_REQUIRE(d->win == 4 && d->grid_h % 2 == 0 && d->grid_w % 2 == 0, MR_ERR_BAD_SHAPE);
    MR_REQUIRE(d->n_win == (d->grid_h / 2) * (d->grid_w / 2) && d->grid_h * d->grid_w + 1 <= d->T, MR_ERR_BAD_SHAPE);
  } else {
    return MR_ERR_BAD_SHAPE;
  }
  return MR_OK;
}

extern "C" int32_t mr_window_mean(mr_stream_t stream_, const MrWindowDesc* d, const void* const* in,
                                  void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  int32_t st = check_window(d);
  if (st != MR_OK) return st;
  MR_REQUIRE(aligned16(in[0]) && aligned16(out[0]), MR_ERR_MISALIGNED);
  if (d->B == 0) return MR_OK;
  const int64_t total = static_cast<int64_t>(d->B) * d->n_win * (d->H / 8);
  window_mean_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const bf16*>(in[0]), static_cast<bf16*>(out[0]), *d);
  return after_launch();
}

extern "C" int32_t mr_window_attention(mr_stream_t stream_, const MrWindowDesc* d, const void* const* in,
                                       void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && out[0], MR_ERR_NULL_POINTER);
  int32_t st = check_window(d);
  if (st != MR_OK) return st;
  MR_REQUIRE(aligned16(in[0]) && aligned16(in[1]) && aligned16(out[0]), MR_ERR_MISALIGNED);
  if (d->B == 0) return MR_OK;
  const int64_t warps = static_cast<int64_t>(d->B) * d->n_win * (d->H / 64);
  window_attention_kernel<<<static_cast<unsigned>(ceil_div64(warps, kWarpsPerBlock)), kRowThreads, 0,
                            static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]), static_cast<bf16*>(out[0]), *d);
  return after_launch();
}

extern "C" int32_t mr_rotary_table(mr_stream_t stream_, const MrRotaryDesc* d, const void* const* in,
                                   void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && (d->ndims == 1 || d->ndims == 2 || d->ndims == 4), MR_ERR_BAD_SHAPE);
  if (d->rows == 0) return MR_OK;
  rotary_table_kernel<<<ceil_div(d->rows * 16, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const float*>(in[0]), static_cast<float*>(out[0]), d->rows, d->ndims);
  return after_launch();
}

extern "C" int32_t mr_embed_assemble(mr_stream_t stream_, const MrAssembleDesc* d, const void* const* in,
                                     void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0] && out[1] && out[2], MR_ERR_NULL_POINTER);
  MR_REQUIRE(in[3] || in[6], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->L > 0 && d->H > 0 && d->H % 8 == 0 && d->T >= d->L + d->n_vis, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->n_vis == 0 || (in[5] && d->pool_h > 0 && d->pool_w > 0 && d->n_vis % (d->pool_h * d->pool_w) == 0),
             MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->n_aud == 0 || (in[4] && d->audio_token_length > 0), MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->vocab > 0, MR_ERR_BAD_SHAPE);
  for (int i = 3; i <= 6; ++i) MR_REQUIRE(!in[i] || aligned16(in[i]), MR_ERR_MISALIGNED);
  MR_REQUIRE(aligned16(out[0]) && aligned16(out[1]), MR_ERR_MISALIGNED);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(d->cls_offset == 0 || d->cls_offset == 1, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->aud_div >= 0 && d->vis_tile >= 0 && (d->vis_tile == 0 || d->vis_group > 0), MR_ERR_BAD_SHAPE);
  const int64_t rows = static_cast<int64_t>(d->B) * (d->T + d->cls_offset);
  embed_assemble_kernel<<<static_cast<unsigned>(ceil_div64(rows, kWarpsPerBlock)), kRowThreads, 0,
                          static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const int32_t*>(in[0]), static_cast<const int32_t*>(in[1]),
      d->n_aud > 0 ? static_cast<const int32_t*>(in[2]) : nullptr, static_cast<const bf16*>(in[3]),
      d->n_aud > 0 ? static_cast<const bf16*>(in[4]) : nullptr, static_cast<const bf16*>(in[5]),
      static_cast<const bf16*>(in[6]), static_cast<bf16*>(out[0]), static_cast<float*>(out[1]),
      static_cast<uint8_t*>(out[2]), *d);
  return after_launch();
}

extern "C" int32_t mr_unit_normalize(mr_stream_t stream_, const MrNormalizeDesc* d, const void* const* in,
                                     void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && d->H > 0, MR_ERR_BAD_SHAPE);
  if (d->rows == 0) return MR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = ceil_div(d->rows, kWarpsPerBlock);
  const int32_t* ridx = static_cast<const int32_t*>(in[1]);
  const float* sdev = d->scale_from_device ? static_cast<const float*>(in[2]) : nullptr;
  MR_REQUIRE(!d->scale_from_device || sdev, MR_ERR_NULL_POINTER);
  if (d->in_f32 && d->out_f32)
    unit_normalize_kernel<true, true><<<grid, kRowThreads, 0, stream>>>(in[0], ridx, out[0], d->rows, d->H, d->scale, sdev);
  else if (d->in_f32)
    unit_normalize_kernel<true, false><<<grid, kRowThreads, 0, stream>>>(in[0], ridx, out[0], d->rows, d->H, d->scale, sdev);
  else if (d->out_f32)
    unit_normalize_kernel<false, true><<<grid, kRowThreads, 0, stream>>>(in[0], ridx, out[0], d->rows, d->H, d->scale, sdev);
  else
    unit_normalize_kernel<false, false><<<grid, kRowThreads, 0, stream>>>(in[0], ridx, out[0], d->rows, d->H, d->scale, sdev);
  return after_launch();
}

extern "C" int32_t mr_segment_pool(mr_stream_t stream_, const MrSegmentPoolDesc* d, const void* const* in,
                                   void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && in[2] && out[0] && out[1], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->L > 0 && d->H > 0 && d->S > 0, MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = ceil_div(d->B * d->S, kWarpsPerBlock);
  if (d->v_f32)
    segment_pool_kernel<true><<<grid, kRowThreads, 0, stream>>>(
        static_cast<const uint8_t*>(in[0]), static_cast<const int32_t*>(in[1]), in[2], static_cast<float*>(out[0]),
        static_cast<float*>(out[1]), *d);
  else
    segment_pool_kernel<false><<<grid, kRowThreads, 0, stream>>>(
        static_cast<const uint8_t*>(in[0]), static_cast<const int32_t*>(in[1]), in[2], static_cast<float*>(out[0]),
        static_cast<float*>(out[1]), *d);
  return after_launch();
}

extern "C" int32_t mr_contrastive_reduce(mr_stream_t stream_, const MrContrastiveDesc* d, const void* const* in,
                                         void* const* out, void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n > 0, MR_ERR_BAD_SHAPE);
  contrastive_reduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const float*>(in[0]), static_cast<const float*>(in[1]), static_cast<const int32_t*>(in[2]),
      static_cast<float*>(out[0]), d->n);
  return after_launch();
}

extern "C" int32_t mr_row_dot(mr_stream_t stream_, const MrRowDotDesc* d, const void* const* in, void* const* out,
                              void*, size_t) {
  MR_REQUIRE(d && in && out && in[0] && in[1] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->rows >= 0 && d->H > 0, MR_ERR_BAD_SHAPE);
  if (d->rows == 0) return MR_OK;
  row_dot_kernel<<<ceil_div(d->rows, kWarpsPerBlock), kRowThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const bf16*>(in[0]), static_cast<const float*>(in[1]), static_cast<const int32_t*>(in[2]),
      static_cast<float*>(out[0]), d->rows, d->H);
  return after_launch();
}
