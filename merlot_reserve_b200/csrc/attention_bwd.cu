// mr_attention_bwd: gradient of the fused attention (apply_attention, modeling.py:188-200) w.r.t. q, k, v.
//
// Flash-attention style: P is recomputed from q, k and the forward's row log-sum-exp, never stored.
//   prep kernel : D[b, h, t] = sum_d dO . O                                  (one warp per (row, head))
//   main kernel : one CTA per (sequence, head, 64-key block j), loop over 64-query blocks i:
//                   S = Q_i K_j^T ; P = exp(S - lse_i) (masked) ; dP = dO_i V_j^T ; dS = P (dP - D_i)
//                   dV_j += P^T dO_i ; dK_j += dS^T Q_i            (fp32 register accumulators, written once)
//                   dQ_i += dS K_j                                 (fp32 atomics into the workspace)
//   convert     : dQ fp32 -> bf16 into the q columns of d(qkv)
// mma.sync m16n8k16 bf16 with fp32 accumulation (the tcgen05 version of this kernel is future work; the forward
// for T > 64 already runs on tcgen05).  q arrives pre-scaled by 1/sqrt(d) and rotary-scaled (qkv GEMM epilogue), so
// the gradients produced here are w.r.t. those scaled values; mr_rotary_bwd applies the same factors afterwards.
// Mask semantics follow the forward: key validity + group ids; a query row whose mask is all false attended
// uniformly to all T keys, and its gradient is the one of that uniform distribution (S := 0, lse = ln T).
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kD = 64;
constexpr int kInfoOob = static_cast<int>(0x80000000u);
constexpr int kInfoInvalid = static_cast<int>(0x80000001u);

__device__ __forceinline__ int swz(int row, int chunk) { return row * kD + ((chunk ^ (row & 7)) << 3); }

// [kBlk x 64] bf16 tile; the CTA has 2 * kBlk threads, four 16-byte chunks each.
template <int kBlk>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* __restrict__ src, size_t seq_base, int row0, int T, int ld,
                                          int col0, int tid) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = tid + 2 * kBlk * k;
    const int row = i >> 3, ch = i & 7;
    const int t = row0 + row;
    const bool ok = t < T;
    cp_async_16(s + swz(row, ch), src + (seq_base + static_cast<size_t>(ok ? t : T - 1)) * ld + col0 + ch * 8, ok);
  }
}

// D[b, h, t] = sum_d dO . O.  One warp per token row: a lane owns 8 consecutive features (one 16-byte load of dO and
// of O), 8 lanes cover one head, so a warp step covers 4 heads and a three-step shuffle finishes each of them.
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ out, float* __restrict__ Dsum, int B, int T,
                     int H) {
  const int nh = H / kD;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);   // b * T + t
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<int64_t>(B) * T) return;
  const int b_ = static_cast<int>(row / T), t = static_cast<int>(row % T);
  const uint4* a4 = reinterpret_cast<const uint4*>(dout + static_cast<size_t>(row) * H);
  const uint4* b4 = reinterpret_cast<const uint4*>(out + static_cast<size_t>(row) * H);
  for (int c = lane; c < H / 8; c += 32) {          // H % 64 == 0: whole groups of 8 lanes are in range together
    const uint4 a = __ldg(a4 + c), b = __ldg(b4 + c);
    float v = bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x);
    v = fmaf(bf16_lo(a.y), bf16_lo(b.y), fmaf(bf16_hi(a.y), bf16_hi(b.y), v));
    v = fmaf(bf16_lo(a.z), bf16_lo(b.z), fmaf(bf16_hi(a.z), bf16_hi(b.z), v));
    v = fmaf(bf16_lo(a.w), bf16_lo(b.w), fmaf(bf16_hi(a.w), bf16_hi(b.w), v));
    const unsigned grp = 0xffu << (lane & 24);      // the 8 lanes of this head
    v += __shfl_xor_sync(grp, v, 4);
    v += __shfl_xor_sync(grp, v, 2);
    v += __shfl_xor_sync(grp, v, 1);
    if ((lane & 7) == 0) Dsum[(static_cast<size_t>(b_) * nh + (c >> 3)) * T + t] = v;
  }
}

// kBlk = 64: 4 warps, any T (the key blocks of a sequence live in different CTAs: dQ through fp32 atomics).
// kBlk = 32: T <= 32 (audio T = 31, span T = 16): the whole (sequence, head) is ONE 32 x 32 block, so the CTA is
// 2 warps with 24 KB of shared memory (64 KB before: ~3x more items in flight per SM), and dQ is complete inside
// the CTA: written as bf16 straight into the q columns of d(qkv), no memset / atomics / convert pass.
template <int kBlk>
__global__ void __launch_bounds__(2 * kBlk)
attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                     const float* __restrict__ Dsum, const uint8_t* __restrict__ valid, const int32_t* __restrict__ group,
                     bf16* __restrict__ dqkv, float* __restrict__ dq_acc, int B, int T, int H) {
  static_assert(kBlk == 64 || kBlk == 32, "16 rows per warp x (2 or 4) warps");
  constexpr int kTile = kBlk * kD;            // elements
  constexpr int kQBufs = kBlk == 64 ? 2 : 1;  // the one-block kernel has nothing to prefetch
  constexpr bool kDirectDQ = kBlk == 32;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);
  bf16* sV = sK + kTile;
  bf16* sQ = sV + kTile;               // [kQBufs]
  bf16* sdO = sQ + kQBufs * kTile;     // [kQBufs]
  bf16* sP = sdO + kQBufs * kTile;
  bf16* sdS = sP + kTile;
  float* sLse = reinterpret_cast<float*>(sdS + kTile);   // [2][kBlk]
  float* sD = sLse + 2 * kBlk;                           // [2][kBlk]
  int* sKInfo = reinterpret_cast<int*>(sD + 2 * kBlk);   // [kBlk]
  int* sQInfo = sKInfo + kBlk;                           // [2][kBlk]: group id, or kInfoInvalid / kInfoOob

  const int nh = H / kD;
  const int nkb = (T + kBlk - 1) / kBlk;
  const int jb = blockIdx.x % nkb;
  const int h = (blockIdx.x / nkb) % nh;
  const int b = blockIdx.x / (nkb * nh);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int ld = 3 * H;
  const size_t seq_base = static_cast<size_t>(b) * T;
  const int k0 = jb * kBlk;
  const size_t stat_base = (static_cast<size_t>(b) * nh + h) * T;

  auto load_q_side = [&](int buf, int qb) {
    load_tile<kBlk>(sQ + buf * kTile, qkv, seq_base, qb * kBlk, T, ld, h * kD, tid);
    load_tile<kBlk>(sdO + buf * kTile, dout, seq_base, qb * kBlk, T, H, h * kD, tid);
    if (tid < kBlk) {
      const int t = qb * kBlk + tid;
      float l = 0.f, dsum = 0.f;
      int info = kInfoOob;
      if (t < T) {
        l = lse[stat_base + t];
        dsum = Dsum[stat_base + t];
        const bool v = valid ? (valid[seq_base + t] != 0) : true;
        info = v ? (group ? group[seq_base + t] : 0) : kInfoInvalid;
      }
      sLse[buf * kBlk + tid] = l;
      sD[buf * kBlk + tid] = dsum;
      sQInfo[buf * kBlk + tid] = info;
    }
  };

  load_tile<kBlk>(sK, qkv, seq_base, k0, T, ld, H + h * kD, tid);
  load_tile<kBlk>(sV, qkv, seq_base, k0, T, ld, 2 * H + h * kD, tid);
  if (tid < kBlk) {
    const int kidx = k0 + tid;
    int info = kInfoOob;
    if (kidx < T) {
      const bool v = valid ? (valid[seq_base + kidx] != 0) : true;
      info = v ? (group ? group[seq_base + kidx] : 0) : kInfoInvalid;
    }
    sKInfo[tid] = info;
  }
  load_q_side(0, 0);
  cp_async_commit();

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dk[i][j] = 0.f;
      dv[i][j] = 0.f;
    }

  const int nqb = nkb;
  for (int qb = 0; qb < nqb; ++qb) {
    const int buf = qb & (kQBufs - 1);
    if (kQBufs == 2 && qb + 1 < nqb) {
      load_q_side(buf ^ 1, qb + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* q_s = sQ + buf * kTile;
    const bf16* do_s = sdO + buf * kTile;

    // ---- S = Q K^T and dP = dO V^T for this warp's 16 query rows ----
    float s[kBlk / 8][4], dp[kBlk / 8][4];
#pragma unroll
    for (int i = 0; i < kBlk / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = 0.f;
        dp[i][j] = 0.f;
      }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t qf[4], of[4];
      {
        const int row = warp * 16 + (lane & 15);
        const int ch = 2 * ks + (lane >> 4);
        ldmatrix_x4(smem_u32(q_s + swz(row, ch)), qf[0], qf[1], qf[2], qf[3]);
        ldmatrix_x4(smem_u32(do_s + swz(row, ch)), of[0], of[1], of[2], of[3]);
      }
#pragma unroll
      for (int n16 = 0; n16 < kBlk / 16; ++n16) {
        const int mat = lane >> 3;
        const int row = n16 * 16 + (mat >> 1) * 8 + (lane & 7);
        const int ch = 2 * ks + (mat & 1);
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4(smem_u32(sK + swz(row, ch)), r0, r1, r2, r3);
        mma_bf16_16816(s[2 * n16], qf, r0, r1);
        mma_bf16_16816(s[2 * n16 + 1], qf, r2, r3);
        ldmatrix_x4(smem_u32(sV + swz(row, ch)), r0, r1, r2, r3);
        mma_bf16_16816(dp[2 * n16], of, r0, r1);
        mma_bf16_16816(dp[2 * n16 + 1], of, r2, r3);
      }
    }
    // ---- P = exp(S - lse) under the mask, dS = P (dP - D); both to shared memory as bf16 [q][key] ----
    {
      const int rl[2] = {warp * 16 + g, warp * 16 + g + 8};
      float lrow[2], drow[2];
      int qinfo[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        lrow[r] = sLse[buf * kBlk + rl[r]];
        drow[r] = sD[buf * kBlk + rl[r]];
        qinfo[r] = sQInfo[buf * kBlk + rl[r]];
      }
#pragma unroll
      for (int nt = 0; nt < kBlk / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const int kinfo = sKInfo[nt * 8 + tq * 2 + (e & 1)];
          float p;
          if (qinfo[r] == kInfoOob || kinfo == kInfoOob) {
            p = 0.f;
          } else if (qinfo[r] == kInfoInvalid) {
            p = __expf(-lrow[r]);                      // uniform row: S := 0 for every key of the sequence
          } else {
            p = (kinfo == qinfo[r]) ? __expf(s[nt][e] - lrow[r]) : 0.f;
          }
          s[nt][e] = p;
          dp[nt][e] = p * (dp[nt][e] - drow[r]);
        }
        const int c = tq * 2;
        *reinterpret_cast<uint32_t*>(sP + swz(rl[0], nt) + c) = pack_bf16(s[nt][0], s[nt][1]);
        *reinterpret_cast<uint32_t*>(sP + swz(rl[1], nt) + c) = pack_bf16(s[nt][2], s[nt][3]);
        *reinterpret_cast<uint32_t*>(sdS + swz(rl[0], nt) + c) = pack_bf16(dp[nt][0], dp[nt][1]);
        *reinterpret_cast<uint32_t*>(sdS + swz(rl[1], nt) + c) = pack_bf16(dp[nt][2], dp[nt][3]);
      }
    }
    // ---- dQ_i (this warp's 16 rows) = dS K_j : A from registers, B = K[key][d] transposed on load ----
    {
      float dq[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
#pragma unroll
      for (int kk = 0; kk < kBlk / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
        pa[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
        pa[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
        pa[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
        for (int d16 = 0; d16 < 4; ++d16) {
          const int mat = lane >> 3;
          const int row = kk * 16 + (mat & 1) * 8 + (lane & 7);
          const int ch = 2 * d16 + (mat >> 1);
          uint32_t r0, r1, r2, r3;
          ldmatrix_x4_trans(smem_u32(sK + swz(row, ch)), r0, r1, r2, r3);
          mma_bf16_16816(dq[2 * d16], pa, r0, r1);
          mma_bf16_16816(dq[2 * d16 + 1], pa, r2, r3);
        }
      }
      const int t0 = qb * kBlk + warp * 16 + g, t1 = t0 + 8;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const int col = h * kD + dt * 8 + tq * 2;
        if constexpr (kDirectDQ) {
          // the only key block of the sequence: dQ is final
          if (t0 < T) *reinterpret_cast<uint32_t*>(dqkv + (seq_base + t0) * ld + col) = pack_bf16(dq[dt][0], dq[dt][1]);
          if (t1 < T) *reinterpret_cast<uint32_t*>(dqkv + (seq_base + t1) * ld + col) = pack_bf16(dq[dt][2], dq[dt][3]);
        } else {
        if (t0 < T) {
          float* dst = dq_acc + (seq_base + t0) * H + col;
          atomicAdd(dst, dq[dt][0]);
          atomicAdd(dst + 1, dq[dt][1]);
        }
        if (t1 < T) {
          float* dst = dq_acc + (seq_base + t1) * H + col;
          atomicAdd(dst, dq[dt][2]);
          atomicAdd(dst + 1, dq[dt][3]);
        }
        }
      }
    }
    __syncthreads();   // P / dS of all kBlk query rows are in shared memory
    // ---- dV_j += P^T dO_i ; dK_j += dS^T Q_i  (this warp's 16 keys; A = transposed load of [q][key]) ----
#pragma unroll
    for (int ks = 0; ks < kBlk / 16; ++ks) {   // 16 query rows per step
      uint32_t pa[4], sa[4];
      {
        const int mat = lane >> 3;
        const int row = ks * 16 + (mat >> 1) * 8 + (lane & 7);   // query index (the k of this product)
        const int ch = 2 * warp + (mat & 1);                     // key chunk (the m of this product)
        ldmatrix_x4_trans(smem_u32(sP + swz(row, ch)), pa[0], pa[1], pa[2], pa[3]);
        ldmatrix_x4_trans(smem_u32(sdS + swz(row, ch)), sa[0], sa[1], sa[2], sa[3]);
      }
#pragma unroll
      for (int d16 = 0; d16 < 4; ++d16) {
        const int mat = lane >> 3;
        const int row = ks * 16 + (mat & 1) * 8 + (lane & 7);
        const int ch = 2 * d16 + (mat >> 1);
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4_trans(smem_u32(do_s + swz(row, ch)), r0, r1, r2, r3);
        mma_bf16_16816(dv[2 * d16], pa, r0, r1);
        mma_bf16_16816(dv[2 * d16 + 1], pa, r2, r3);
        ldmatrix_x4_trans(smem_u32(q_s + swz(row, ch)), r0, r1, r2, r3);
        mma_bf16_16816(dk[2 * d16], sa, r0, r1);
        mma_bf16_16816(dk[2 * d16 + 1], sa, r2, r3);
      }
    }
    __syncthreads();   // before the next iteration overwrites sP / sdS and prefetches into this buffer's sibling
  }

  // ---- write dK_j, dV_j (bf16) into the k / v columns of d(qkv) ----
  const int kr0 = k0 + warp * 16 + g, kr1 = kr0 + 8;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = h * kD + dt * 8 + tq * 2;
    if (kr0 < T) {
      bf16* row = dqkv + (seq_base + kr0) * ld;
      *reinterpret_cast<uint32_t*>(row + H + col) = pack_bf16(dk[dt][0], dk[dt][1]);
      *reinterpret_cast<uint32_t*>(row + 2 * H + col) = pack_bf16(dv[dt][0], dv[dt][1]);
    }
    if (kr1 < T) {
      bf16* row = dqkv + (seq_base + kr1) * ld;
      *reinterpret_cast<uint32_t*>(row + H + col) = pack_bf16(dk[dt][2], dk[dt][3]);
      *reinterpret_cast<uint32_t*>(row + 2 * H + col) = pack_bf16(dv[dt][2], dv[dt][3]);
    }
  }
}

__global__ void __launch_bounds__(256)
dq_convert_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv, int64_t rows, int H) {
  const int vec_per_row = H / 8;
  const int64_t total = rows * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    const float4 a = *reinterpret_cast<const float4*>(dq_acc + r * H + c);
    const float4 b = *reinterpret_cast<const float4*>(dq_acc + r * H + c + 4);
    uint4 o;
    o.x = pack_bf16(a.x, a.y);
    o.y = pack_bf16(a.z, a.w);
    o.z = pack_bf16(b.x, b.y);
    o.w = pack_bf16(b.z, b.w);
    *reinterpret_cast<uint4*>(dqkv + r * 3 * H + c) = o;
  }
}

template <int kBlk>
constexpr int bwd_smem_bytes() {
  return (4 + 2 * (kBlk == 64 ? 2 : 1)) * kBlk * kD * 2 + 4 * kBlk * 4 + 3 * kBlk * 4;
}

}  // namespace
}  // namespace mr

namespace mr {
size_t dqkv_finish_partial_bytes(int64_t rows, int H);   // backward_rows.cu
int32_t launch_dqkv_finish(cudaStream_t stream, const float* dq_acc, bf16* dqkv, const float* rot, int rot_rows,
                           float q_scale, int64_t rows, int H, float* partial, float* dbias);
int32_t launch_attention_bwd_tc(cudaStream_t stream, const bf16* qkv, const bf16* dout, const float* lse, const float* Dsum,
                                const uint8_t* valid, const int32_t* group, bf16* dqkv, float* dq_acc, int B, int T, int H);
}

extern "C" size_t mr_attention_bwd_workspace_bytes(const MrAttentionDesc* d) {
  if (!d || d->B <= 0 || d->T <= 0 || d->H <= 0) return 0;
  const size_t rows = static_cast<size_t>(d->B) * d->T;
  size_t bytes = rows * d->H * sizeof(float) + rows * (d->H / 64) * sizeof(float);
  if (d->finish & MR_ATTN_BWD_BIAS_GRAD) bytes += mr::dqkv_finish_partial_bytes(static_cast<int64_t>(rows), d->H) + 16;
  return bytes;
}

extern "C" int32_t mr_attention_bwd(mr_stream_t stream_, const MrAttentionDesc* d, const void* const* in, void* const* out,
                                    void* workspace, size_t workspace_bytes) {
  using namespace mr;
  MR_REQUIRE(d && in && out && in[0] && in[1] && in[2] && in[3] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->T > 0 && d->H > 0 && d->H % 64 == 0, MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(workspace && workspace_bytes >= mr_attention_bwd_workspace_bytes(d), MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(in[0]) && aligned16(in[1]) && aligned16(in[2]) && aligned16(out[0]) && aligned16(workspace),
             MR_ERR_MISALIGNED);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int nh = d->H / 64;
  const int64_t rows = static_cast<int64_t>(d->B) * d->T;
  float* dq_acc = static_cast<float*>(workspace);
  float* Dsum = dq_acc + rows * d->H;
  // finish mode: the backward of the qkv epilogue and / or the bias gradient fused into the dQ conversion pass
  const bool finish = d->finish != 0;
  const float* rot = nullptr;
  float* dbias = nullptr;
  float* partial = nullptr;
  if (finish) {
    MR_REQUIRE((d->finish & ~(MR_ATTN_BWD_EPILOGUE | MR_ATTN_BWD_BIAS_GRAD)) == 0, MR_ERR_BAD_SHAPE);
    if (d->finish & MR_ATTN_BWD_EPILOGUE) {
      MR_REQUIRE(d->q_scale > 0.f, MR_ERR_BAD_SHAPE);
      rot = static_cast<const float*>(in[6]);
      MR_REQUIRE(!rot || d->rot_rows > 0, MR_ERR_BAD_SHAPE);
      MR_REQUIRE(!rot || aligned16(rot), MR_ERR_MISALIGNED);
    }
    if (d->finish & MR_ATTN_BWD_BIAS_GRAD) {
      dbias = static_cast<float*>(out[1]);
      MR_REQUIRE(dbias, MR_ERR_NULL_POINTER);
      partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Dsum + rows * nh) + 15) & ~uintptr_t(15));
    }
  }
  const float q_scale = (d->finish & MR_ATTN_BWD_EPILOGUE) ? d->q_scale : 1.0f;
  MR_REQUIRE(d->algo >= MR_ATTN_ALGO_AUTO && d->algo <= MR_ATTN_ALGO_TCGEN05, MR_ERR_BAD_SHAPE);
  const bool use_tc = d->algo == MR_ATTN_ALGO_TCGEN05 || (d->algo == MR_ATTN_ALGO_AUTO && d->T > 64);
  // T <= 32 under AUTO: the one-block kernel, dQ written in place (no fp32 accumulator to clear or convert)
  const bool one_block = !use_tc && d->T <= 32 && d->algo == MR_ATTN_ALGO_AUTO;
  cudaError_t e = cudaSuccess;
  if (!one_block) e = cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(rows) * d->H * sizeof(float), stream);
  if (e != cudaSuccess) return cuda_status(e);
  attn_bwd_prep_kernel<<<static_cast<unsigned>(ceil_div64(rows, 8)), 256, 0, stream>>>(
      static_cast<const bf16*>(in[1]), static_cast<const bf16*>(in[2]), Dsum, d->B, d->T, d->H);
  int32_t st = after_launch();
  if (st != MR_OK) return st;
  // T > 64: tcgen05 / TMEM kernel (attention_bwd_tc.cu); short sequences (audio, span) and algo = MMA_SYNC: mma.sync
  if (use_tc) {
    st = launch_attention_bwd_tc(stream, static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]),
                                 static_cast<const float*>(in[3]), Dsum, static_cast<const uint8_t*>(in[4]),
                                 static_cast<const int32_t*>(in[5]), static_cast<bf16*>(out[0]), dq_acc, d->B, d->T, d->H);
    if (st != MR_OK) return st;
    if (finish)
      return launch_dqkv_finish(stream, dq_acc, static_cast<bf16*>(out[0]), rot, d->rot_rows, q_scale, rows, d->H, partial, dbias);
    int64_t cb = ceil_div64(rows * (d->H / 8), 256);
    if (cb > 148 * 16) cb = 148 * 16;
    dq_convert_kernel<<<static_cast<unsigned>(cb), 256, 0, stream>>>(dq_acc, static_cast<bf16*>(out[0]), rows, d->H);
    return after_launch();
  }
  if (one_block) {
    const int64_t blocks = static_cast<int64_t>(d->B) * nh;
    MR_REQUIRE(blocks < (int64_t(1) << 31), MR_ERR_BAD_SHAPE);
    attention_bwd_kernel<32><<<static_cast<unsigned>(blocks), 64, bwd_smem_bytes<32>(), stream>>>(
        static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]), static_cast<const float*>(in[3]), Dsum,
        static_cast<const uint8_t*>(in[4]), static_cast<const int32_t*>(in[5]), static_cast<bf16*>(out[0]), dq_acc, d->B,
        d->T, d->H);
    st = after_launch();
    if (st != MR_OK || !finish) return st;
    return launch_dqkv_finish(stream, nullptr, static_cast<bf16*>(out[0]), rot, d->rot_rows, q_scale, rows, d->H, partial, dbias);
  }
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(attention_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem_bytes<64>());
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  const int nkb = ceil_div(d->T, 64);
  const int64_t blocks = static_cast<int64_t>(d->B) * nh * nkb;
  MR_REQUIRE(blocks < (int64_t(1) << 31), MR_ERR_BAD_SHAPE);
  attention_bwd_kernel<64><<<static_cast<unsigned>(blocks), 128, bwd_smem_bytes<64>(), stream>>>(
      static_cast<const bf16*>(in[0]), static_cast<const bf16*>(in[1]), static_cast<const float*>(in[3]), Dsum,
      static_cast<const uint8_t*>(in[4]), static_cast<const int32_t*>(in[5]), static_cast<bf16*>(out[0]), dq_acc, d->B,
      d->T, d->H);
  st = after_launch();
  if (st != MR_OK) return st;
  if (finish)
    return launch_dqkv_finish(stream, dq_acc, static_cast<bf16*>(out[0]), rot, d->rot_rows, q_scale, rows, d->H, partial, dbias);
  int64_t cblocks = ceil_div64(rows * (d->H / 8), 256);
  if (cblocks > 148 * 16) cblocks = 148 * 16;
  dq_convert_kernel<<<static_cast<unsigned>(cblocks), 256, 0, stream>>>(dq_acc, static_cast<bf16*>(out[0]), rows, d->H);
  return after_launch();
}
