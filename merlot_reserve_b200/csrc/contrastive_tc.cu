// mr_contrastive_lse_fwd: one direction of loss_fn_given_preds (pretrain/pretrain_model.py:276,288-295) in ONE kernel.
//
//   logits[i, j] = x[i] . y_all[j]          tcgen05.mma, bf16 x bf16 -> fp32 in TMEM, 128 x 128 tiles
//   lse[i]       = logsumexp_j logits[i, j] online (running max / running sum per row), fp32
//   numer[i]     = logits[i, pos_off + i]   the positive logit, picked out of the diagonal block
//
// The [n, n_all] logit matrix never exists in memory: a tile lives in TMEM (double buffered), is read once by the
// row-per-thread softmax warpgroup and is gone.  Work split: blockIdx.y = 128-row tile of x, and the 128-column tiles of
// y_all are dealt round-robin to the CTAs of a thread-block CLUSTER along x (up to 8 CTAs); every CTA keeps a running
// (max, sum) per row over its tiles, ships the pair to the cluster's rank-0 CTA through distributed shared memory
// (st.shared::cluster), and rank 0 merges them after one cluster barrier.  No workspace, no atomics, no second launch.
//
// Roles per CTA (256 threads): warp 0 = TMA producer (x and y_all k-blocks into a 4-stage ring, 128B swizzle),
// warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..7 = softmax warpgroup (thread = row = TMEM lane).
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64, kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2, kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 256;
constexpr int kMaxCluster = 8;
constexpr int kPartBytes = kMaxCluster * kBM * 8;                       // float2 (max, sum) per row per cluster rank
constexpr int kSmemBytes = kStages * kStageBytes + kPartBytes + 256 + 1024;
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

struct Params {
  int n, n_all, H, pos_off;
  float* lse;
  float* numer;
};

__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

__global__ void __launch_bounds__(kThreads, 1)
contrastive_lse_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                          const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array: an integer round trip makes the pointer generic and
  // every access through it LD.E / ST.E instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float2* part = reinterpret_cast<float2*>(smem + kStages * kStageBytes);          // [cluster rank][row]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kPartBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const int row_tile = blockIdx.y;
  const int col_tiles = (p.n_all + kBN - 1) / kBN;
  const int kblocks = (p.H + kBK - 1) / kBK;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, 2 * kBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ct = static_cast<int>(crank); ct < col_tiles; ct += static_cast<int>(csize)) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(&tmap_x, &full_bar[stage], sa, kb * kBK, row_tile * kBM, kEvictLast);
          tma_load_2d(&tmap_y, &full_bar[stage], sa + kABytes, kb * kBK, ct * kBN, kEvictNormal);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(kBM, kBN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int ct = static_cast<int>(crank); ct < col_tiles; ct += static_cast<int>(csize)) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != 0 || k != 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax warpgroup: thread = row =====================
    const int quarter = warp_idx & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int row = row_tile * kBM + row_in_tile;
    const int target = p.pos_off + row;                       // column of this row's positive
    const int warp_t0 = p.pos_off + row_tile * kBM + quarter * 32;   // targets of this warp: [warp_t0, warp_t0 + 32)
    float m = -INFINITY, s = 0.f;                             // running max (natural units) and sum of exp(logit - m)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int ct = static_cast<int>(crank); ct < col_tiles; ct += static_cast<int>(csize)) {
      const int col0 = ct * kBN;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      uint32_t v[128];
      tmem_ld_32x128_wait(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kBN), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);       // the MMA warp may overwrite this accumulator stage
      const int ncols = min(kBN, p.n_all - col0);             // >= 1
      if (ncols < kBN) {
#pragma unroll
        for (int j = 0; j < kBN; ++j)
          if (j >= ncols) v[j] = __float_as_uint(-INFINITY);
      }
      if (warp_t0 < col0 + kBN && warp_t0 + 32 > col0) {      // warp-uniform: the diagonal crosses this tile
        const int tj = target - col0;
        float pos = 0.f;
        bool hit = false;
#pragma unroll
        for (int j = 0; j < kBN; ++j)
          if (j == tj) {
            pos = __uint_as_float(v[j]);
            hit = true;
          }
        if (hit && row < p.n && tj < ncols) p.numer[row] = pos;
      }
      float tmax = __uint_as_float(v[0]);
#pragma unroll
      for (int j = 1; j + 1 < kBN; j += 2) tmax = fmax3(tmax, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
      tmax = fmaxf(tmax, __uint_as_float(v[kBN - 1]));
      const float m_new = fmaxf(m, tmax);
      const float off = m_new * kLog2e;
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kBN; j += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc4[q] += ex2_approx(fmaf(__uint_as_float(v[j + q]), kLog2e, -off));
      }
      s = s * ex2_approx((m - m_new) * kLog2e) + ((acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
      m = m_new;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    // ship (m, s) to the rank-0 CTA of the cluster
    const uint32_t dst = mapa_shared(smem_u32(&part[crank * kBM + row_in_tile]), 0);
    st_cluster_f32x2(dst, m, s);
  }

  tc_fence_before();
  cluster_sync_all();                                          // release / acquire: the partials are visible in rank 0
  if (crank == 0 && warp_idx >= 4) {
    const int row_in_tile = (warp_idx & 3) * 32 + lane;
    const int row = row_tile * kBM + row_in_tile;
    float M = -INFINITY;
    for (uint32_t c = 0; c < csize; ++c) M = fmaxf(M, part[c * kBM + row_in_tile].x);
    float S = 0.f;
    for (uint32_t c = 0; c < csize; ++c) {
      const float2 t = part[c * kBM + row_in_tile];
      S += t.y * ex2_approx((t.x - M) * kLog2e);
    }
    if (row < p.n) p.lse[row] = M + __log2f(S) * kLn2;
  }
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kBN);
  }
}

}  // namespace
}  // namespace mr

extern "C" size_t mr_contrastive_lse_fwd_workspace_bytes(const MrContrastiveDesc*) { return 0; }

extern "C" int32_t mr_contrastive_lse_fwd(mr_stream_t stream_, const MrContrastiveDesc* d, const void* const* in,
                                          void* const* out, void*, size_t) {
  using namespace mr;
  MR_REQUIRE(d && in && out && in[0] && in[1] && out[0] && out[1], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->n >= 0 && d->n_all > 0 && d->H > 0 && d->H % 8 == 0, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->pos_off >= 0 && d->pos_off + d->n <= d->n_all, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(in[0]) && aligned16(in[1]), MR_ERR_MISALIGNED);
  if (d->n == 0) return MR_OK;
  CUtensorMap mx, my;
  MR_REQUIRE(make_tmap_2d(&mx, in[0], d->n, d->H, d->H, kBM, kBK), MR_ERR_DRIVER);
  MR_REQUIRE(make_tmap_2d(&my, in[1], d->n_all, d->H, d->H, kBN, kBK), MR_ERR_DRIVER);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(contrastive_lse_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  const int col_tiles = ceil_div(d->n_all, kBN);
  const int cluster = col_tiles < kMaxCluster ? col_tiles : kMaxCluster;
  Params p{d->n, d->n_all, d->H, d->pos_off, static_cast<float*>(out[0]), static_cast<float*>(out[1])};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(cluster), static_cast<unsigned>(ceil_div(d->n, kBM)));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = static_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, contrastive_lse_tc_kernel, mx, my, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cuda_status(e != cudaSuccess ? e : cudaGetLastError());
}
