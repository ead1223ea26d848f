// Tensor-core (tcgen05) flash attention forward for head dim 64, used when T > 64.
//
// One CTA per (sequence, head, 128-query block); two CTAs co-reside on an SM (256 TMEM columns and
// ~82 KB of shared memory each) so that one CTA's softmax overlaps the other's MMAs.
//
//   warp 0     TMA producer : Q tile once, then K / V tiles [128 keys x 64] into a 2-deep ring (128B swizzle)
//   warp 1     MMA issuer   : S = Q K^T  (SS form: 128 x 128 x 64, accumulator in TMEM cols [0,128))
//                             O += P V   (TS form: P read from TMEM cols [128,192) as bf16 pairs, V is the
//                             MN-major B operand straight from the TMA tile; O in TMEM cols [192,256))
//   warps 2-5  softmax      : thread = query row.  tcgen05.ld S -> mask -> exp2 (base-2, fp32) -> row sum ->
//                             bf16 P -> tcgen05.st.  The running max is only re-based ("lazy rescale") when a
//                             block's row max exceeds it by more than 2^8, in which case O is corrected in
//                             TMEM; any fixed shift within that range gives the same softmax in fp32/bf16.
//
// Masks follow mr_attention_fwd's contract: key validity + video_src group ids; a query row whose mask is
// all-false attends uniformly over all T keys (the reference's -1e10 bias behaviour).
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kQ = 128;
constexpr int kKV = 128;
constexpr int kD = 64;
constexpr int kTcThreads = 192;
constexpr int kTileBytes = 128 * kD * 2;  // 16 KB
constexpr int kTmemCols = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;
constexpr int kInfoOob = static_cast<int>(0x80000000u);
constexpr int kInfoInvalid = static_cast<int>(0x80000001u);
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int kSmemBytes = 5 * kTileBytes + 2 * kKV * 4 + 256 + 1024;

struct AttnParams {
  const uint8_t* valid;
  const int32_t* group;
  bf16* out;
  int B, T, H, nh, nqb, nkb;
};

template <bool kMasked>
__global__ void __launch_bounds__(kTcThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;        // 2 stages
  uint8_t* sV = smem + 3 * kTileBytes;    // 2 stages
  int* sInfo = reinterpret_cast<int*>(smem + 5 * kTileBytes);  // [2][kKV]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * kTileBytes + 2 * kKV * 4);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* s_free = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* pv_done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x % p.nqb;
  const int h = (blockIdx.x / p.nqb) % p.nh;
  const int b = blockIdx.x / (p.nqb * p.nh);
  const int seq_base = b * p.T;
  const int q0 = qb * kQ;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_2d(&tmap_qkv, q_full, sQ, h * kD, seq_base + q0, kEvictFirst);
      for (int j = 0; j < p.nkb; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kTileBytes);
        tma_load_2d(&tmap_qkv, &k_full[s], sK + s * kTileBytes, p.H + h * kD, seq_base + j * kKV, kEvictLast);
        mbar_arrive_expect_tx(&v_full[s], kTileBytes);
        tma_load_2d(&tmap_qkv, &v_full[s], sV + s * kTileBytes, 2 * p.H + h * kD, seq_base + j * kKV, kEvictLast);
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32_ex(kQ, kKV, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16_f32_ex(kQ, kD, 1);  // B (= V) is MN-major
      const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK + (j & 1) * kTileBytes));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16_ss(tmem_base + kColS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < p.nkb; ++j) {
        const int s = j & 1;
        if (j + 1 < p.nkb) {
          // S_{j+1} may overwrite the S columns as soon as the softmax warps have pulled S_j into registers
          mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          mbar_wait(s_free, j & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t vdesc = umma_desc_k_sw128(smem_u32(sV + s * kTileBytes));
#pragma unroll
        for (int kk = 0; kk < kKV / 16; ++kk) {
          // 16 keys = 16 rows x 128 B = 2048 B further into the V tile; P advances 8 TMEM columns
          umma_bf16_ts(tmem_base + kColO, tmem_base + kColP + 8 * kk, vdesc + 128 * kk, idesc_o,
                       (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(pv_done);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (128 threads, thread = query row) =====================
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const int st = threadIdx.x - 64;  // 0..127
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int q = q0 + row;
    const bool q_in = q < p.T;
    bool q_uniform = false;
    int q_group = 0;
    if (kMasked) {
      const bool v = q_in ? (p.valid ? p.valid[seq_base + q] != 0 : true) : false;
      q_uniform = !v;
      q_group = (v && p.group) ? p.group[seq_base + q] : 0;
    }
    float m2 = 0.f;  // reference max, log2 domain
    float l = 0.f;

    for (int j = 0; j < p.nkb; ++j) {
      const int nvk = min(kKV, p.T - j * kKV);  // keys of this block that exist
      if (kMasked) {
        const int kidx = j * kKV + st;
        int info = kInfoOob;
        if (kidx < p.T) {
          const bool v = p.valid ? p.valid[seq_base + kidx] != 0 : true;
          info = v ? (p.group ? p.group[seq_base + kidx] : 0) : kInfoInvalid;
        }
        sInfo[(j & 1) * kKV + st] = info;
        named_bar_sync(1, 128);
      }
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float sc[kKV];
#pragma unroll
      for (int c = 0; c < kKV / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + kColS + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sc[32 * c + i] = __uint_as_float(v[i]) * kLog2e;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);

      // ---- masking ----
      if (kMasked) {
        const int* info = sInfo + (j & 1) * kKV;
#pragma unroll
        for (int c4 = 0; c4 < kKV / 4; ++c4) {
          const int4 f = *reinterpret_cast<const int4*>(info + 4 * c4);
          const int ff[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = 4 * c4 + i;
            if (q_uniform)
              sc[c] = (ff[i] != kInfoOob) ? 0.f : -INFINITY;
            else if (ff[i] != q_group)
              sc[c] = -INFINITY;
          }
        }
      } else if (nvk < kKV) {
#pragma unroll
        for (int c = 0; c < kKV; ++c)
          if (c >= nvk) sc[c] = -INFINITY;
      }
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kKV; ++c) mx = fmaxf(mx, sc[c]);

      // ---- lazy re-basing of the reference max ----
      float factor = 1.f;
      if (j == 0) {
        m2 = (mx == -INFINITY) ? 0.f : mx;
      } else if (mx > m2 + kRescaleThreshold) {
        factor = ex2_approx(m2 - mx);
        m2 = mx;
      }
      // P and O may only be touched once PV_{j-1} has retired
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, factor != 1.f)) {
          l *= factor;
#pragma unroll
          for (int c = 0; c < kD / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + kColO + 32 * c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * factor);
            tmem_st_32x32(lane_addr + kColO + 32 * c, v);
          }
        }
      }
      // ---- P = exp2(s - m), row sum, bf16 pack, store to TMEM ----
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < kKV / 64; ++c) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p0 = ex2_approx(sc[64 * c + 2 * i] - m2);
          const float p1 = ex2_approx(sc[64 * c + 2 * i + 1] - m2);
          sum += p0 + p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st_32x32(lane_addr + kColP + 32 * c, pk);
      }
      l += sum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(pv_done, (p.nkb - 1) & 1);
    tc_fence_after();
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    uint4 o16[kD / 8];
#pragma unroll
    for (int c = 0; c < kD / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(lane_addr + kColO + 32 * c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o16[4 * c + i].x = pack_bf16(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
        o16[4 * c + i].y = pack_bf16(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
        o16[4 * c + i].z = pack_bf16(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
        o16[4 * c + i].w = pack_bf16(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
      }
    }
    if (q_in) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(seq_base) + q) * p.H + h * kD);
#pragma unroll
      for (int i = 0; i < kD / 8; ++i) dst[i] = o16[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// Called by mr_attention_fwd (attention.cu) when T is long enough for 128-wide tiles.
int32_t launch_attention_tc(cudaStream_t stream, const bf16* qkv, const uint8_t* valid, const int32_t* group,
                            bf16* out, int B, int T, int H) {
  CUtensorMap tm;
  if (!make_tmap_2d(&tm, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, 128, kD)) return MR_ERR_DRIVER;
  AttnParams p{};
  p.valid = valid;
  p.group = group;
  p.out = out;
  p.B = B;
  p.T = T;
  p.H = H;
  p.nh = H / kD;
  p.nqb = ceil_div(T, kQ);
  p.nkb = ceil_div(T, kKV);
  const int64_t blocks = static_cast<int64_t>(B) * p.nh * p.nqb;
  if (blocks >= (int64_t(1) << 31)) return MR_ERR_BAD_SHAPE;
  const bool masked = valid != nullptr || group != nullptr;
  auto kern = masked ? attention_tc_kernel<true> : attention_tc_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[masked]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set[masked] = true;
  }
  kern<<<static_cast<unsigned>(blocks), kTcThreads, kSmemBytes, stream>>>(tm, p);
  return after_launch();
}

}  // namespace mr
