// mr_attention_fwd: fused softmax(q k^T + mask) v, one CTA per (sequence, head, 64-query block).
//
// Replaces apply_attention (mreserve/modeling.py:188-200) + flax dot_product_attention_weights and the
// O(T^2) boolean mask / -1e10 bias of modeling.py:343-358,743-756, which is passed here as per-position
// validity bytes and video_src group ids instead.  The 1/sqrt(64) query scale and the rotary scale are
// already folded into the qkv GEMM epilogue.
//
// v1 data path: cp.async double-buffered K/V tiles (XOR-swizzled, conflict-free ldmatrix), mma.sync
// m16n8k16 bf16 with fp32 accumulation, online softmax in fp32 (base-2), P kept in registers.
//
// Two instantiations.  <64, 2>: 64-query x 64-key tiles, 4 warps, double-buffered keys (any T).  <32, 1>: the
// audio (T = 31) and span (T = 16) sequences fit ONE 32 x 32 tile, so the CTA is 2 warps with 12 KB of shared
// memory (no second key buffer) instead of 4 warps (two of them idle) with 40 KB: ~3.5x more (sequence, head)
// items in flight per SM for a kernel whose time is the load -> mma -> softmax -> mma -> store latency chain.
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kD = 64;
constexpr int kInfoOob = static_cast<int>(0x80000000u);      // key index >= T
constexpr int kInfoInvalid = static_cast<int>(0x80000001u);  // padding key
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ int swz(int row, int chunk) { return row * kD + ((chunk ^ (row & 7)) << 3); }

// Loads a [kTile x 64] bf16 tile (rows row0.. of sequence b, column offset col0 in the qkv matrix); the CTA has
// 2 * kTile threads, so every thread moves four 16-byte chunks.
template <int kTile>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* __restrict__ qkv, size_t seq_base, int row0, int T,
                                          int ld, int col0, int tid) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = tid + 2 * kTile * k;
    const int row = i >> 3, ch = i & 7;
    const int t = row0 + row;
    const bool ok = t < T;
    const bf16* src = qkv + (seq_base + static_cast<size_t>(ok ? t : T - 1)) * ld + col0 + ch * 8;
    cp_async_16(s + swz(row, ch), src, ok);
  }
}

template <int kTile, int kBufs>
__global__ void __launch_bounds__(2 * kTile, kTile == 32 ? 10 : 1)
attention_fwd_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ valid, const int32_t* __restrict__ group,
                     bf16* __restrict__ out, float* __restrict__ lse_out, int B, int T, int H) {
  constexpr int kBlockM = kTile, kBlockN = kTile;
  static_assert(kTile == 64 || kTile == 32, "tile = 16 rows per warp x (2 or 4) warps");
  static_assert(kBufs == 1 || kBufs == 2, "kBufs == 1 requires T <= kTile (a single key block)");
  __shared__ __align__(128) bf16 sQ[kBlockM * kD];
  __shared__ __align__(128) bf16 sK[kBufs][kBlockN * kD];
  __shared__ __align__(128) bf16 sV[kBufs][kBlockN * kD];
  __shared__ int sInfo[kBufs][kBlockN];

  const int nh = H / kD;
  const int nqb = (T + kBlockM - 1) / kBlockM;
  const int qb = blockIdx.x % nqb;
  const int h = (blockIdx.x / nqb) % nh;
  const int b = blockIdx.x / (nqb * nh);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld = 3 * H;
  const size_t seq_base = static_cast<size_t>(b) * T;
  const int q0 = qb * kBlockM;
  const int nkb = (T + kBlockN - 1) / kBlockN;

  auto write_info = [&](int buf, int kb) {
    if (tid < kBlockN) {
      const int kidx = kb * kBlockN + tid;
      int info = kInfoOob;
      if (kidx < T) {
        const bool v = valid ? (valid[seq_base + kidx] != 0) : true;
        info = v ? (group ? group[seq_base + kidx] : 0) : kInfoInvalid;
      }
      sInfo[buf][tid] = info;
    }
  };

  load_tile<kTile>(sQ, qkv, seq_base, q0, T, ld, h * kD, tid);
  load_tile<kTile>(sK[0], qkv, seq_base, 0, T, ld, H + h * kD, tid);
  load_tile<kTile>(sV[0], qkv, seq_base, 0, T, ld, 2 * H + h * kD, tid);
  write_info(0, 0);
  cp_async_commit();

  // query-side mask info for the two rows this thread owns
  const int g = lane >> 2, tq = lane & 3;
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  bool q_uniform[2];
  int q_group[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const bool in = qrow[r] < T;
    const bool v = in ? (valid ? (valid[seq_base + qrow[r]] != 0) : true) : false;
    q_uniform[r] = !v;
    q_group[r] = (v && group) ? group[seq_base + qrow[r]] : 0;
  }
  const bool warp_active = (q0 + warp * 16) < T;

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & (kBufs - 1);
    if (kBufs == 2 && kb + 1 < nkb) {
      const int nxt = (kb + 1) & (kBufs - 1);
      load_tile<kTile>(sK[nxt], qkv, seq_base, (kb + 1) * kBlockN, T, ld, H + h * kD, tid);
      load_tile<kTile>(sV[nxt], qkv, seq_base, (kb + 1) * kBlockN, T, ld, 2 * H + h * kD, tid);
      write_info(nxt, kb + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (warp_active) {
      if (kb == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int row = warp * 16 + (lane & 15);
          const int ch = 2 * ks + (lane >> 4);
          ldmatrix_x4(smem_u32(sQ + swz(row, ch)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        }
      }
      // ---- S = Q K^T (16 x 64 per warp) ----
      float s[kBlockN / 8][4];
#pragma unroll
      for (int i = 0; i < kBlockN / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int n16 = 0; n16 < kBlockN / 16; ++n16) {
          const int mat = lane >> 3;
          const int row = n16 * 16 + (mat >> 1) * 8 + (lane & 7);
          const int ch = 2 * ks + (mat & 1);
          uint32_t r0, r1, r2, r3;
          ldmatrix_x4(smem_u32(sK[buf] + swz(row, ch)), r0, r1, r2, r3);
          mma_bf16_16816(s[2 * n16], qf[ks], r0, r1);
          mma_bf16_16816(s[2 * n16 + 1], qf[ks], r2, r3);
        }
      }
      // ---- mask + online softmax ----
      float blk_max[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < kBlockN / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const int info = sInfo[buf][nt * 8 + tq * 2 + (e & 1)];
          float v = s[nt][e];
          if (q_uniform[r]) {
            v = (info != kInfoOob) ? 0.f : -INFINITY;
          } else if (info != q_group[r]) {
            v = -INFINITY;
          }
          s[nt][e] = v;
          blk_max[r] = fmaxf(blk_max[r], v);
        }
      }
      float scale_o[2], m_use[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float bm = blk_max[r];
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
        const float m_new = fmaxf(m_run[r], bm);
        m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
        scale_o[r] = exp2f((m_run[r] - m_use[r]) * kLog2e);  // m_run = -inf -> 0
        m_run[r] = m_new;
      }
      float blk_sum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < kBlockN / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const float p = exp2f((s[nt][e] - m_use[r]) * kLog2e);
          s[nt][e] = p;
          blk_sum[r] += p;
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * scale_o[r] + blk_sum[r];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        o[dt][0] *= scale_o[0];
        o[dt][1] *= scale_o[0];
        o[dt][2] *= scale_o[1];
        o[dt][3] *= scale_o[1];
      }
      // ---- O += P V ----
#pragma unroll
      for (int kk = 0; kk < kBlockN / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int d16 = 0; d16 < 4; ++d16) {
          const int mat = lane >> 3;
          const int row = kk * 16 + (mat & 1) * 8 + (lane & 7);
          const int ch = 2 * d16 + (mat >> 1);
          uint32_t r0, r1, r2, r3;
          ldmatrix_x4_trans(smem_u32(sV[buf] + swz(row, ch)), r0, r1, r2, r3);
          mma_bf16_16816(o[2 * d16], pa, r0, r1);
          mma_bf16_16816(o[2 * d16 + 1], pa, r2, r3);
        }
      }
    }
    __syncthreads();
  }

  // ---- finalise: O / l, stage through sQ (each warp only touches its own 16 rows), 16-byte stores ----
  if (warp_active) {
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float l = l_run[r];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      inv[r] = l > 0.f ? 1.0f / l : 0.f;
      // row log-sum-exp of the masked scores, kept for the backward pass (mr_attention_bwd)
      if (lse_out != nullptr && tq == 0 && qrow[r] < T)
        lse_out[(static_cast<size_t>(b) * nh + h) * T + qrow[r]] = (m_run[r] == -INFINITY ? 0.f : m_run[r]) + (l > 0.f ? logf(l) : 0.f);
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int col = dt * 8 + tq * 2;
      const int r0 = warp * 16 + g, r1 = r0 + 8;
      *reinterpret_cast<uint32_t*>(sQ + swz(r0, dt) + (col & 7)) = pack_bf16(o[dt][0] * inv[0], o[dt][1] * inv[0]);
      *reinterpret_cast<uint32_t*>(sQ + swz(r1, dt) + (col & 7)) = pack_bf16(o[dt][2] * inv[1], o[dt][3] * inv[1]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane + 32 * k;
      const int row = warp * 16 + (i >> 3), ch = i & 7;
      const int t = q0 + row;
      if (t < T) {
        const uint4 v = *reinterpret_cast<const uint4*>(sQ + swz(row, ch));
        *reinterpret_cast<uint4*>(out + (seq_base + t) * H + h * kD + ch * 8) = v;
      }
    }
  }
}

}  // namespace
}  // namespace mr

namespace mr {
int32_t launch_attention_tc(cudaStream_t stream, const bf16* qkv, const uint8_t* valid, const int32_t* group,
                            bf16* out, float* lse, int B, int T, int H, int variant);
}

extern "C" size_t mr_attention_fwd_workspace_bytes(const MrAttentionDesc*) { return 0; }

extern "C" int32_t mr_attention_fwd(mr_stream_t stream_, const MrAttentionDesc* d, const void* const* in,
                                    void* const* out, void*, size_t) {
  using namespace mr;
  MR_REQUIRE(d && in && out && in[0] && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->B >= 0 && d->T > 0 && d->H > 0 && d->H % 64 == 0, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(in[0]) && aligned16(out[0]), MR_ERR_MISALIGNED);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(d->algo >= MR_ATTN_ALGO_AUTO && d->algo <= MR_ATTN_ALGO_TCGEN05_3STREAM, MR_ERR_BAD_SHAPE);
  float* lse = d->emit_lse ? static_cast<float*>(out[1]) : nullptr;
  MR_REQUIRE(!d->emit_lse || lse, MR_ERR_NULL_POINTER);
  if (d->algo >= MR_ATTN_ALGO_TCGEN05 || (d->algo == MR_ATTN_ALGO_AUTO && d->T > 64)) {
    const int variant = d->algo == MR_ATTN_ALGO_TCGEN05_2TILE ? 2 : (d->algo == MR_ATTN_ALGO_TCGEN05_3STREAM ? 3 : 0);
    return launch_attention_tc(static_cast<cudaStream_t>(stream_), static_cast<const bf16*>(in[0]),
                               static_cast<const uint8_t*>(in[1]), static_cast<const int32_t*>(in[2]),
                               static_cast<bf16*>(out[0]), lse, d->B, d->T, d->H, variant);
  }
  const int nh = d->H / 64;
  // T <= 32 (audio T = 31, span T = 16): one 32 x 32 tile per (sequence, head), 2-warp CTAs, 12 KB smem
  const bool small = d->T <= 32 && d->algo == MR_ATTN_ALGO_AUTO;   // MR_ATTN_ALGO_MMA_SYNC keeps the 64 x 64 tiles
  const int tile = small ? 32 : 64;
  const int nqb = ceil_div(d->T, tile);
  const int64_t blocks = static_cast<int64_t>(d->B) * nh * nqb;
  MR_REQUIRE(blocks < (int64_t(1) << 31), MR_ERR_BAD_SHAPE);
  if (small)
    attention_fwd_kernel<32, 1><<<static_cast<unsigned>(blocks), 64, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const bf16*>(in[0]), static_cast<const uint8_t*>(in[1]), static_cast<const int32_t*>(in[2]),
        static_cast<bf16*>(out[0]), lse, d->B, d->T, d->H);
  else
    attention_fwd_kernel<64, 2><<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const bf16*>(in[0]), static_cast<const uint8_t*>(in[1]), static_cast<const int32_t*>(in[2]),
        static_cast<bf16*>(out[0]), lse, d->B, d->T, d->H);
  return after_launch();
}
