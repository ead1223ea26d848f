// mr_gemm_epilogue: D = epilogue(A . W^T) on the sm_100a tensor cores.
//
// Persistent, warp-specialised kernel (one CTA per SM):
//   warp 0      TMA producer   : cp.async.bulk.tensor (128B swizzle) of A[128 x 64] and W[BN x 64] tiles
//                                into a kStages-deep shared-memory ring, signalled through mbarriers
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32) with the
//                                accumulator in TMEM; tcgen05.commit releases ring slots / publishes tiles
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       : tcgen05.ld the accumulator (double buffered in TMEM, so the epilogue of
//                                tile i overlaps the main loop of tile i+1), apply
//                                bias / rotary scale / sigmoid-GELU / residual, cast and store.
//
// Two epilogue data paths:
//   kDirect = false (bf16 output, no row remap -- every GEMM of the transformer layers): each epilogue warp
//     owns a 32-row x 64-column box staged in shared memory with the 128-byte TMA swizzle.  The residual
//     box is TMA-loaded into it while the MMAs of the tile are still running, the fused result is written
//     over it in place and leaves with one TMA store.  A thread owns one accumulator ROW (that is how
//     TMEM lanes map to threads), so direct global accesses would touch 32 different 128-byte lines per
//     warp instruction; ncu showed exactly that as the limiter of the K=768 GEMMs (profiles/ncu_gemm_r01b.md).
//   kDirect = true (fp32 output or row-remapped output): per-thread 16-byte global accesses (rare, small).
//
// The arithmetic this replaces is flax Dense/DenseGeneral (+ the elementwise ops around it) at the
// reference call sites listed in include/mreserve_b200.h.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace mr {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;

struct GemmParams {
  int M, N, K;
  int ldd;
  int bias_on, out_f32;
  int seq_in, seq_out, row_off;
  int rot_rows, hidden;
  float q_scale;
  const float* bias;
  const bf16* residual;
  const float* rot;
  void* D;
  int num_n_tiles, num_tiles, num_k_blocks;
  // LayerNorm folded into the GEMM (see the header): consumer side reads per-row partial (sum, sum of squares),
  // producer side writes them for its own (bf16-rounded) output rows.
  const float2* stats_in;   // [K / 64][M]
  const float* ln_s;        // [N] column sums of the gamma-scaled weight
  float2* stats_out;        // [ceil(N / 64)][M]
  float ln_eps;
  // split-K (direct fp32 path only): work item = (tile, split); split s reduces k-blocks [s * kb_per_split, ...) and
  // writes its partial to D + s * split_stride elements
  int splits, kb_per_split;
  size_t split_stride;
};

constexpr int kBoxBytes = 32 * 128;  // one epilogue staging box: 32 rows x 64 bf16, 128B swizzle

// kCtas = 2: CTA pair (cta_group::2).  The pair owns a 256 x BN tile: each CTA stages its own 128 rows of A and
// HALF of the W tile (BN / 2 rows); the leader issues 256 x BN x 16 MMAs that read both CTAs' shared memory and
// write each CTA's half of the accumulator into its own TMEM.  Per CTA and k-block that is 32 KB of L2 -> SMEM
// traffic instead of 48 KB, and the smaller stages allow a 6-deep ring.
template <int BN, int kCtas = 1, bool kDual = false>
struct Cfg {
  // kDual (two output boxes per epilogue warp): one ring stage less pays for the second set of staging boxes
  static constexpr int kStages = ((BN == 256 && kCtas == 1) ? 4 : 6) - (kDual ? 1 : 0);
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / kCtas) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
  static constexpr int kOutBytes = kEpiWarps * kBoxBytes * (kDual ? 2 : 1);
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kBarBytes + 1024;  // +1024: manual alignment
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

__device__ __forceinline__ float gelu_sigmoid(float x) {
  // x * sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x)): ONE SFU op (tanh.approx, rel. error 2^-11, far below
  // the bf16 rounding of the result) instead of ex2 + rcp -- the SFU was the second limiter of this epilogue.
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

__device__ __forceinline__ float gelu_sigmoid_grad(float x) {
  // d/dx [x sigmoid(1.702 x)] = s (1 + 1.702 x (1 - s)), s = sigmoid(1.702 x) = 0.5 (1 + tanh(0.851 x))
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float sg = fmaf(0.5f, t, 0.5f);
  return sg * fmaf(1.702f * x, 1.0f - sg, 1.0f);
}

__device__ __forceinline__ int map_row(int r, int seq_in, int seq_out, int row_off) {
  return seq_in == 0 ? r : (r / seq_in) * seq_out + (r % seq_in) + row_off;
}

// Loads 8 consecutive fp32 column values (bias / colsum vectors).  kFull: the whole 64-column box is in range,
// two 16-byte loads that every lane of the warp shares (one L1 wavefront each); otherwise element-wise guarded.
template <bool kFull>
__device__ __forceinline__ void load_cols8(const float* __restrict__ v, int c, int n, float (&o)[8]) {
  if constexpr (kFull) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(v + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(v + c + 4));
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = (c + q < n) ? __ldg(v + c + q) : 0.f;
  }
}

// kBias / kLn / kStats are compile-time on the staged path so that the unrolled epilogue body is branch free and
// ptxas can hoist the (L1-resident) per-column vector loads ahead of their uses; with run-time flags every 8-column
// step sat behind its own reconvergence point and the epilogue became latency bound (profiles/ncu_gemm_r01j.md).
// The direct path keeps run-time p.bias_on (small GEMMs only).
template <int BN, bool kGelu, bool kResidual, bool kRotary, bool kDirect, bool kBias = false, bool kLn = false,
          bool kStats = false, int kCtas = 1, int kTrans = 0, bool kDual = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const GemmParams p) {
  using C = Cfg<BN, kCtas, kDual>;
  constexpr bool kPair = kCtas == 2;
  // MR_EPI_GELU_DUAL (training forward of the MLP-up projection): the bf16 PRE-activation goes out through a second
  // staging box and the `tmap_res` tensor map next to the activation -- the backward needs both, and a separate GELU
  // kernel would read the pre-activation back from HBM.
  static_assert(!kDual || (kGelu && !kResidual && !kDirect && !kStats && kTrans == 0), "dual output: staged bias+GELU only");
  static_assert(!(kPair && kDirect) && (!kPair || BN == 256), "CTA pairs: staged 256-wide tiles only");
  // kTrans bit 0: A is MN-major (given as [K, M], the backward's "A^T"); bit 1: B is MN-major (given as [K, N]).
  // An MN-major operand is staged as 64-column slabs of [64 k-rows x 128 B] (one TMA box each, 128B swizzle):
  // slab stride = LBO = 8 KB, 8-row groups inside a slab = SBO = 1 KB, one K=16 MMA step = 2 KB further.
  constexpr bool kAmn = (kTrans & 1) != 0, kBmn = (kTrans & 2) != 0;
  // MR_EPI_DGELU (the dgrad of the MLP "out" projection): the "residual" box carries the forward's pre-activation
  // and the accumulator is MULTIPLIED by gelu'(box) instead of added to it.  Spelled gelu + residual on kTrans == 2.
  constexpr bool kDGelu = kGelu && kResidual && kTrans == 2;
  static_assert(!kDGelu || !kDirect, "the fused GELU backward lives on the staged path");
  static_assert(!(kPair && kTrans != 0), "MN-major operands: single-CTA tiles only");
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs of the pair)
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array: an integer round trip makes the pointer generic and
  // every access through it LD.E / ST.E instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kOutBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_bar = tmem_empty_bar + 2;  // [kEpiWarps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kEpiWarps);
  uint8_t* smem_out = smem + C::kStages * C::kStageBytes;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (!kDirect) {
      tma_prefetch_desc(&tmap_out);
      if constexpr (kResidual || kDual) tma_prefetch_desc(&tmap_res);
    }
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kEpiWarps * kCtas);   // pair: the leader's barrier collects both CTAs' epilogues
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(&res_bar[w], 1);
    fence_mbar_init();
  }
  if (warp_idx == 2) {
    if constexpr (kPair) {
      tmem_alloc_pair(tmem_slot, C::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (kPair)
    cluster_sync_all();       // both CTAs' barriers are initialised before any remote arrive / TMA credit
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Static schedule over tiles in m-major order.  Default: round-robin (tile = cta, cta + grid, ...), so the
  // CTAs that share an A row block run at the same time and a long-K A block is fetched from HBM once.
  // Rotary (QKV) GEMMs: every CTA owns a CONTIGUOUS range instead, so consecutive tiles of a CTA share the row
  // block and the per-row rotary scales stay in registers (K = hidden there: 148 live A blocks fit the L2).
  // CTA pairs: the schedule is over PAIR tiles (256 rows); CTA `cta_rank` owns row block 2 * m_pair + cta_rank.
  // split-K exists on the direct path only; elsewhere these fold to constants
  const int splits = kDirect ? p.splits : 1;
  const int kb_per_split = kDirect ? p.kb_per_split : p.num_k_blocks;
  const int num_workers = static_cast<int>(gridDim.x) / kCtas;
  const int worker = static_cast<int>(blockIdx.x) / kCtas;
  const int tiles_per_cta = (p.num_tiles + num_workers - 1) / num_workers;
  // LN-consuming GEMMs (K = hidden as well) use the contiguous schedule too: the row statistics stay in registers.
  constexpr bool contiguous = kRotary || kLn;
  const int tile_step = contiguous ? 1 : num_workers;
  const int tile_begin = contiguous ? worker * tiles_per_cta : worker;
  const int tile_end = contiguous ? min(tile_begin + tiles_per_cta, p.num_tiles) : p.num_tiles;

  // Register re-distribution (staged path): the producer / MMA warpgroup runs single-thread loops, the two epilogue
  // warpgroups hold 64 accumulator values plus the hoisted per-column vectors per thread.  64 + 2 x 216 registers
  // per thread-slot fit the 64 K file of the CTA's launch allocation (384 x 168).
  // (setmaxnreg must dominate the role code it applies to, hence the nesting below.)
  if (warp_idx < 4) {
  if constexpr (!kDirect) setmaxnreg_dec<64>();
  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int mn = tile / splits, split = tile % splits;
        const int m_blk = (mn / p.num_n_tiles) * kCtas + static_cast<int>(cta_rank);
        const int n_blk = mn % p.num_n_tiles;
        const int kb_end = min((split + 1) * kb_per_split, p.num_k_blocks);
        for (int kb = split * kb_per_split; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          // activations are streamed once per N-sweep; weights are re-read by every M block
          if constexpr (kPair) {
            // both CTAs load into their own ring slot; all bytes are credited to the leader's barrier
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
            tma_load_2d_pair(&tmap_a, &full_bar[stage], sa, kb * kBK, m_blk * kBM, kEvictNormal);
            tma_load_2d_pair(&tmap_b, &full_bar[stage], sb, kb * kBK, n_blk * BN + static_cast<int>(cta_rank) * (BN / 2),
                             kEvictLast);
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          if constexpr (kAmn) {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i)
              tma_load_2d(&tmap_a, &full_bar[stage], sa + i * 8192, m_blk * kBM + 64 * i, kb * kBK, kEvictNormal);
          } else {
            tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * kBK, m_blk * kBM, kEvictNormal);
          }
          if constexpr (kBmn) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(&tmap_b, &full_bar[stage], sb + i * 8192, n_blk * BN + 64 * i, kb * kBK, kEvictLast);
          } else {
            tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * kBK, n_blk * BN, kEvictLast);
          }
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(kBM * kCtas, BN) | (kAmn ? (1u << 15) : 0u) | (kBmn ? (1u << 16) : 0u);
      constexpr uint64_t kLboMn = static_cast<uint64_t>((8192 >> 4) - 1) << 16;   // LBO field holds 1 for K-major
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        const int split = tile % splits;
        const int kb_begin = split * kb_per_split, kb_end = min(kb_begin + kb_per_split, p.num_k_blocks);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          const uint64_t adesc = umma_desc_k_sw128(sa) + (kAmn ? kLboMn : 0);
          const uint64_t bdesc = umma_desc_k_sw128(sb) + (kBmn ? kLboMn : 0);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // one K=16 step: K-major = 32 bytes along the 128-byte swizzle row (+2 in 16-byte units);
            // MN-major = 16 k-rows x 128 B = 2 KB further into every slab (+128)
            const uint64_t ad = adesc + (kAmn ? 128 : 2) * k, bd = bdesc + (kBmn ? 128 : 2) * k;
            const uint32_t accum = (kb != kb_begin || k != 0) ? 1u : 0u;
            if constexpr (kPair)
              umma_bf16_ss_pair(d_tmem, ad, bd, idesc, accum);
            else
              umma_bf16_ss(d_tmem, ad, bd, idesc, accum);
          }
          // frees the ring slot (in both CTAs of a pair) once these MMAs have read it
          if constexpr (kPair) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete (each CTA's epilogue reads its own half from its own TMEM)
        if constexpr (kPair) umma_commit_pair(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  }
  } else {
    if constexpr (!kDirect) setmaxnreg_inc<216>();
    // ===================== epilogue =====================
    const int ew = warp_idx - 4;
    const int quarter = warp_idx & 3;          // TMEM lane quarter this warp may access
    const int col_half = ew >> 2;              // which half of the BN columns this warp owns
    constexpr int kChunks = BN / 64;           // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (!kDirect) {
      // ---------- staged path: TMEM -> registers -> swizzled smem box -> TMA store ----------
      constexpr int kBoxes = BN / 128;         // 64-column boxes per warp
      uint8_t* sbuf = smem_out + ew * kBoxBytes;
      uint8_t* my_row = sbuf + lane * 128;     // this thread's 128-byte row of the box
      uint8_t* sbuf2 = smem_out + (kEpiWarps + ew) * kBoxBytes;   // kDual: the pre-activation box
      uint8_t* my_row2 = sbuf2 + lane * 128;
      const int sw = lane & 7;                 // 128B swizzle: 16-byte chunk j lives at chunk (j ^ (row & 7))
      uint64_t* rbar = &res_bar[ew];
      uint32_t res_phase = 0;
      float rot[kRotary ? 32 : 1];
      int rot_m = -1;
      float ln_mu = 0.f, ln_rstd = 1.f;        // LayerNorm statistics of this thread's row (LN-consuming GEMMs)
      int ln_m = -1;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int mn = tile;                     // no split-K on the staged path
        const int m_blk = (mn / p.num_n_tiles) * kCtas + static_cast<int>(cta_rank);
        const int n_blk = mn % p.num_n_tiles;
        const int row0 = m_blk * kBM + quarter * 32;
        const int r = row0 + lane;
        const int col_base = n_blk * BN + col_half * (BN / 2);
        if (kLn && m_blk != ln_m) {
          ln_m = m_blk;
          float s1 = 0.f, s2 = 0.f;
          if (r < p.M) {
            const int parts = p.K >> 6;
            for (int q = 0; q < parts; ++q) {
              const float2 t = __ldg(p.stats_in + static_cast<size_t>(q) * p.M + r);
              s1 += t.x;
              s2 += t.y;
            }
          }
          const float inv_k = 1.0f / static_cast<float>(p.K);
          ln_mu = s1 * inv_k;
          ln_rstd = rsqrtf(s2 * inv_k - ln_mu * ln_mu + p.ln_eps);
        }

        // Work that does not depend on the accumulator goes first: residual box 0 and the rotary row.
        if (lane == 0) {
          tma_store_wait_read<0>();            // the previous tile's store has finished reading the box
          if constexpr (kResidual) {
            if (col_base < p.N) {
              mbar_arrive_expect_tx(rbar, kBoxBytes);
              tma_load_2d(&tmap_res, rbar, sbuf, col_base, row0, kEvictFirst);
            }
          }
        }
        __syncwarp();
        if constexpr (kRotary) {
          if (m_blk != rot_m) {                // same row block as the previous tile: keep the registers
            rot_m = m_blk;
            const float4* src = reinterpret_cast<const float4*>(p.rot + static_cast<size_t>(r < p.M ? (r % p.rot_rows) : 0) * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 t = __ldg(src + j);
              rot[4 * j + 0] = t.x;
              rot[4 * j + 1] = t.y;
              rot[4 * j + 2] = t.z;
              rot[4 * j + 3] = t.w;
            }
          }
        }

        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();

#pragma unroll
        for (int box = 0; box < kBoxes; ++box) {
          const int col0 = col_base + 64 * box;
          const bool box_ok = col0 < p.N;      // warp-uniform
          if (box > 0) {
            if (lane == 0) {
              tma_store_wait_read<0>();
              if constexpr (kResidual) {
                if (box_ok) {
                  mbar_arrive_expect_tx(rbar, kBoxBytes);
                  tma_load_2d(&tmap_res, rbar, sbuf, col0, row0, kEvictFirst);
                }
              }
            }
            __syncwarp();
          }
          uint32_t v[64];
          {
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN + col_half * (BN / 2) + box * 64);
            uint32_t lo[32], hi[32];
            tmem_ld_32x32(taddr, lo);
            tmem_ld_32x32(taddr + 32, hi);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = lo[j];
              v[32 + j] = hi[j];
            }
          }
          if (box == kBoxes - 1) {
            // all TMEM reads of this accumulator stage are done: hand it back to the (leader's) MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (kPair)
                mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));
              else
                mbar_arrive(&tmem_empty_bar[acc]);
            }
          }
          if (!box_ok) continue;
          if constexpr (kResidual) {
            mbar_wait(rbar, res_phase);
            res_phase ^= 1;
          }
          const bool full_box = col0 + 64 <= p.N;
          float st_sum = 0.f, st_sq = 0.f;
          auto box_body = [&](auto full_tag) {
            constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float f[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(v[8 * j + q]);
              const int c = col0 + 8 * j;
              if constexpr (kLn) {
                // LN(x) W = rstd (x W' - mu colsum(W')): W' = gamma-scaled weight, beta W + b is the bias below
                float sv[8];
                load_cols8<kFull>(p.ln_s, c, p.N, sv);
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = (f[q] - ln_mu * sv[q]) * ln_rstd;
              }
              if constexpr (kBias) {
                float bv[8];
                load_cols8<kFull>(p.bias, c, p.N, bv);
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] += bv[q];
              }
              if constexpr (kRotary) {
                // a box is exactly one head: [q heads | k heads | v heads]; first 32 dims of q/k heads are scaled
                if (j < 4 && col0 < 2 * p.hidden) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) f[q] *= rot[8 * j + q];
                }
                if (col0 < p.hidden) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) f[q] *= p.q_scale;
                }
              }
              if constexpr (kDual) {
                uint4 o2;
                o2.x = pack_bf16(f[0], f[1]);
                o2.y = pack_bf16(f[2], f[3]);
                o2.z = pack_bf16(f[4], f[5]);
                o2.w = pack_bf16(f[6], f[7]);
                *reinterpret_cast<uint4*>(my_row2 + ((j ^ sw) << 4)) = o2;
              }
              if constexpr (kGelu && !kDGelu) {
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = gelu_sigmoid(f[q]);
              }
              uint4* cell = reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4));
              if constexpr (kDGelu) {
                const uint4 rr = *cell;
                const float pre[8] = {bf16_lo(rr.x), bf16_hi(rr.x), bf16_lo(rr.y), bf16_hi(rr.y),
                                      bf16_lo(rr.z), bf16_hi(rr.z), bf16_lo(rr.w), bf16_hi(rr.w)};
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] *= gelu_sigmoid_grad(pre[q]);
              } else if constexpr (kResidual) {
                const uint4 rr = *cell;
                f[0] += bf16_lo(rr.x); f[1] += bf16_hi(rr.x);
                f[2] += bf16_lo(rr.y); f[3] += bf16_hi(rr.y);
                f[4] += bf16_lo(rr.z); f[5] += bf16_hi(rr.z);
                f[6] += bf16_lo(rr.w); f[7] += bf16_hi(rr.w);
              }
              uint4 o;
              o.x = pack_bf16(f[0], f[1]);
              o.y = pack_bf16(f[2], f[3]);
              o.z = pack_bf16(f[4], f[5]);
              o.w = pack_bf16(f[6], f[7]);
              *cell = o;
              if constexpr (kStats) {
                // statistics of the ROUNDED values: exactly what the consuming GEMM will read back as x
                const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float a = bf16_lo(ow[q]), b = bf16_hi(ow[q]);
                  st_sum += a + b;
                  st_sq = fmaf(a, a, fmaf(b, b, st_sq));
                }
              }
            }
          };
          if (full_box)
            box_body(std::true_type{});
          else
            box_body(std::false_type{});
          if constexpr (kStats) {
            if (r < p.M) p.stats_out[static_cast<size_t>(col0 >> 6) * p.M + r] = make_float2(st_sum, st_sq);
          }
          fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, sbuf, col0, row0);   // rows >= M and columns >= N are clipped by the tensor map
            if constexpr (kDual) tma_store_2d(&tmap_res, sbuf2, col0, row0);
            tma_store_commit();
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (lane == 0) tma_store_wait<0>();      // all bulk stores complete before the CTA retires its smem
    } else
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      const int mn = tile / splits, split = tile % splits;
      const int m_blk = mn / p.num_n_tiles;
      const int n_blk = mn % p.num_n_tiles;
      const int r = m_blk * kBM + quarter * 32 + lane;
      const bool row_ok = r < p.M;
      const int out_row = map_row(r, p.seq_in, p.seq_out, p.row_off);
      const int col_base = n_blk * BN + col_half * (BN / 2);

      // Issue the global loads this row needs BEFORE waiting for the accumulator.
      uint4 res[kResidual ? kChunks : 1][4];
      if constexpr (kResidual) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = col_base + c * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) res[c][j] = make_uint4(0, 0, 0, 0);
          if (row_ok && col0 + 32 <= p.N) {
            const uint4* src = reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(out_row) * p.ldd + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) res[c][j] = __ldg(src + j);
          }
        }
      }
      float rot[kRotary ? 32 : 1];
      if constexpr (kRotary) {
        const float4* src = reinterpret_cast<const float4*>(p.rot + static_cast<size_t>(row_ok ? (r % p.rot_rows) : 0) * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 t = __ldg(src + j);
          rot[4 * j + 0] = t.x;
          rot[4 * j + 1] = t.y;
          rot[4 * j + 2] = t.z;
          rot[4 * j + 3] = t.w;
        }
      }

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();

#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * BN + col_half * (BN / 2) + c * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        if (c == kChunks - 1) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int col0 = col_base + c * 32;
        if (!row_ok || col0 >= p.N) continue;

        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const bool full_chunk = col0 + 32 <= p.N;
        if (p.bias_on) {
          if (full_chunk) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 t = __ldg(b4 + j);
              f[4 * j + 0] += t.x;
              f[4 * j + 1] += t.y;
              f[4 * j + 2] += t.z;
              f[4 * j + 3] += t.w;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
          }
        }
        if constexpr (kRotary) {
          // columns: [q heads | k heads | v heads]; first 32 dims of every q/k head are scaled
          if (col0 < 2 * p.hidden && (col0 & 63) == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= rot[j];
          }
          if (col0 < p.hidden) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= p.q_scale;
          }
        }
        if constexpr (kGelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_sigmoid(f[j]);
        }
        if constexpr (kResidual) {
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w[4] = {res[c][j].x, res[c][j].y, res[c][j].z, res[c][j].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                f[8 * j + 2 * q + 0] += bf16_lo(w[q]);
                f[8 * j + 2 * q + 1] += bf16_hi(w[q]);
              }
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N)
                f[j] += __bfloat162float(p.residual[static_cast<size_t>(out_row) * p.ldd + col0 + j]);
          }
        }
        if (p.out_f32) {
          float* dst = reinterpret_cast<float*>(p.D) + static_cast<size_t>(split) * p.split_stride +
                       static_cast<size_t>(out_row) * p.ldd + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(dst)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) dst[j] = f[j];
          }
        } else {
          bf16* dst = reinterpret_cast<bf16*>(p.D) + static_cast<size_t>(out_row) * p.ldd + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
              reinterpret_cast<uint4*>(dst)[j] = o;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) dst[j] = __float2bfloat16(f[j]);
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  if constexpr (kPair)
    cluster_sync_all();       // the leader's MMAs read the peer's shared memory: neither CTA may retire early
  else
    __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    if constexpr (kPair)
      tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else
      tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Slow CUDA-core kernel with the same epilogue contract, kept ONLY to bisect the tensor-core
// path in tests (MR_GEMM_ALGO_SIMT_DEBUG).  One thread per output element.
// ------------------------------------------------------------------------------------------
__global__ void gemm_simt_debug_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W,
                                       GemmParams p, int gelu, int residual, int rotary) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(p.M) * p.N) return;
  const int r = static_cast<int>(idx / p.N);
  const int c = static_cast<int>(idx % p.N);
  float acc = 0.f;
  const bf16* a = A + static_cast<size_t>(r) * lda;
  const bf16* w = W + static_cast<size_t>(c) * p.K;
  for (int k = 0; k < p.K; ++k) acc += __bfloat162float(a[k]) * __bfloat162float(w[k]);
  if (p.bias_on) acc += p.bias[c];
  if (rotary) {
    if (c < 2 * p.hidden && (c & 63) < 32) acc *= p.rot[static_cast<size_t>(r % p.rot_rows) * 32 + (c & 63)];
    if (c < p.hidden) acc *= p.q_scale;
  }
  if (gelu) acc = acc / (1.0f + expf(-1.702f * acc));
  const int out_row = map_row(r, p.seq_in, p.seq_out, p.row_off);
  const size_t o = static_cast<size_t>(out_row) * p.ldd + c;
  if (residual) acc += __bfloat162float(p.residual[o]);
  if (p.out_f32)
    reinterpret_cast<float*>(p.D)[o] = acc;
  else
    reinterpret_cast<bf16*>(p.D)[o] = __float2bfloat16(acc);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Maps {
  CUtensorMap a, b, out, res;
};

template <int BN, bool G, bool R, bool Q, bool D, bool B = false, bool L = false, bool S = false, int P = 1, int T = 0,
          bool U = false>
int32_t launch_one(cudaStream_t stream, const Maps& m, const GemmParams& p, int grid) {
  auto kern = gemm_tcgen05_kernel<BN, G, R, Q, D, B, L, S, P, T, U>;
  static bool attr_set = false;  // benign race: the attribute is idempotent
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, P, U>::kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  if constexpr (P == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg<BN, P, U>::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, m.a, m.b, m.out, m.res, p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_status(e != cudaSuccess ? e : cudaGetLastError());
  } else {
    kern<<<grid, kThreads, Cfg<BN, 1, U>::kSmemBytes, stream>>>(m.a, m.b, m.out, m.res, p);
    return after_launch();
  }
}

struct Epi {
  bool gelu, residual, rotary, bias, ln, stats;
};

// MN-major ("transposed") operands: the backward GEMMs.  Plain epilogue only (optional run-time bias on the direct
// path); trans = 1 (A^T given), 2 (B^T given, i.e. dgrad on the forward's packed weights), 3 (both: wgrad).
template <int BN>
int32_t launch_trans(cudaStream_t stream, const Maps& m, const GemmParams& p, int grid, int trans, bool direct) {
  if (direct) {
    if (trans == 1) return launch_one<BN, false, false, false, true, false, false, false, 1, 1>(stream, m, p, grid);
    if (trans == 2) return launch_one<BN, false, false, false, true, false, false, false, 1, 2>(stream, m, p, grid);
    return launch_one<BN, false, false, false, true, false, false, false, 1, 3>(stream, m, p, grid);
  }
  if (trans == 1) return launch_one<BN, false, false, false, false, false, false, false, 1, 1>(stream, m, p, grid);
  if (trans == 2) return launch_one<BN, false, false, false, false, false, false, false, 1, 2>(stream, m, p, grid);
  return launch_one<BN, false, false, false, false, false, false, false, 1, 3>(stream, m, p, grid);
}

// out[r, c] = sum_s partial[s][r, c]  (split-K second stage; deterministic order)
template <bool kOutF32>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, void* __restrict__ out, int M, int N, int ldd, int splits,
                     size_t split_stride) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n4 = (N + 3) / 4;
  if (i >= static_cast<int64_t>(M) * n4) return;
  const int r = static_cast<int>(i / n4), c = static_cast<int>(i % n4) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < splits; ++s) {
    const float* src = part + static_cast<size_t>(s) * split_stride + static_cast<size_t>(r) * ldd + c;
    if (c + 4 <= N) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    } else {
      for (int q = 0; q < 4; ++q)
        if (c + q < N) acc[q] += src[q];
    }
  }
  for (int q = 0; q < 4; ++q) {
    if (c + q >= N) break;
    if (kOutF32)
      static_cast<float*>(out)[static_cast<size_t>(r) * ldd + c + q] = acc[q];
    else
      static_cast<bf16*>(out)[static_cast<size_t>(r) * ldd + c + q] = __float2bfloat16(acc[q]);
  }
}

// direct path (fp32 or row-remapped output): bias is a run-time flag, no LN fold / statistics
template <int BN>
int32_t launch_direct(cudaStream_t stream, const Maps& m, const GemmParams& p, int grid, const Epi& e) {
  if (e.ln || e.stats) return MR_ERR_BAD_SHAPE;
  if (e.rotary && !e.gelu && !e.residual) return launch_one<BN, false, false, true, true>(stream, m, p, grid);
  if (e.gelu && !e.residual && !e.rotary) return launch_one<BN, true, false, false, true>(stream, m, p, grid);
  if (e.residual && !e.gelu && !e.rotary) return launch_one<BN, false, true, false, true>(stream, m, p, grid);
  if (!e.gelu && !e.residual && !e.rotary) return launch_one<BN, false, false, false, true>(stream, m, p, grid);
  if (e.gelu && e.residual && !e.rotary) return launch_one<BN, true, true, false, true>(stream, m, p, grid);
  return MR_ERR_BAD_SHAPE;  // rotary never combines with gelu/residual on this path
}

// staged path (bf16 output through shared memory + TMA store): every flag is compile time
template <int BN, bool G, bool R, bool Q, int P>
int32_t launch_staged_bls(cudaStream_t stream, const Maps& m, const GemmParams& p, int grid, const Epi& e) {
  if (e.ln && e.stats) return MR_ERR_BAD_SHAPE;
  if (e.ln) {
    if (!e.bias || R) return MR_ERR_BAD_SHAPE;   // the folded LayerNorm always carries beta W (+ b) as its bias
    if constexpr (!R) return launch_one<BN, G, R, Q, false, true, true, false, P>(stream, m, p, grid);
  }
  if (e.stats) {
    if (G || Q) return MR_ERR_BAD_SHAPE;
    if constexpr (!G && !Q) {
      if (e.bias) return launch_one<BN, G, R, Q, false, true, false, true, P>(stream, m, p, grid);
      return launch_one<BN, G, R, Q, false, false, false, true, P>(stream, m, p, grid);
    }
  }
  if (e.bias) return launch_one<BN, G, R, Q, false, true, false, false, P>(stream, m, p, grid);
  return launch_one<BN, G, R, Q, false, false, false, false, P>(stream, m, p, grid);
}

template <int BN, int P = 1>
int32_t launch_staged(cudaStream_t stream, const Maps& m, const GemmParams& p, int grid, const Epi& e) {
  if (e.rotary && !e.gelu && !e.residual) return launch_staged_bls<BN, false, false, true, P>(stream, m, p, grid, e);
  if (e.gelu && !e.residual && !e.rotary) return launch_staged_bls<BN, true, false, false, P>(stream, m, p, grid, e);
  if (e.residual && !e.gelu && !e.rotary) return launch_staged_bls<BN, false, true, false, P>(stream, m, p, grid, e);
  if (!e.gelu && !e.residual && !e.rotary) return launch_staged_bls<BN, false, false, false, P>(stream, m, p, grid, e);
  if (e.gelu && e.residual && !e.rotary) return launch_staged_bls<BN, true, true, false, P>(stream, m, p, grid, e);
  return MR_ERR_BAD_SHAPE;
}

}  // namespace

}  // namespace mr

extern "C" size_t mr_gemm_epilogue_workspace_bytes(const MrGemmDesc* d) {
  if (!d || d->split_k <= 1) return 0;
  return static_cast<size_t>(d->split_k) * static_cast<size_t>(d->M) * static_cast<size_t>(d->ldd) * sizeof(float);
}

extern "C" int32_t mr_gemm_epilogue(mr_stream_t stream_, const MrGemmDesc* d, const void* const* in,
                                    void* const* out, void* workspace, size_t workspace_bytes) {
  using namespace mr;
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bf16* A = static_cast<const bf16*>(in[0]);
  const bf16* W = static_cast<const bf16*>(in[1]);
  MR_REQUIRE(A && W && out[0], MR_ERR_NULL_POINTER);
  MR_REQUIRE(d->M >= 0 && d->N > 0 && d->K > 0, MR_ERR_BAD_SHAPE);
  if (d->M == 0) return MR_OK;
  const int trans = (d->trans_a ? 1 : 0) | (d->trans_b ? 2 : 0);
  const int ldb = d->ldb > 0 ? d->ldb : (d->trans_b ? d->N : d->K);
  // only the row PITCHES must be multiples of 8 elements (16 bytes, TMA); M, N, K themselves are arbitrary
  MR_REQUIRE(d->lda % 8 == 0 && d->lda >= (d->trans_a ? d->M : d->K), MR_ERR_BAD_SHAPE);
  MR_REQUIRE(ldb % 8 == 0 && ldb >= (d->trans_b ? d->N : d->K), MR_ERR_BAD_SHAPE);
  MR_REQUIRE(d->ldd % 8 == 0 && d->ldd >= d->N, MR_ERR_BAD_SHAPE);
  MR_REQUIRE(aligned16(A) && aligned16(W) && aligned16(out[0]), MR_ERR_MISALIGNED);
  const bool bias = d->epilogue & MR_EPI_BIAS, gelu = d->epilogue & MR_EPI_GELU;
  const bool residual = d->epilogue & MR_EPI_RESIDUAL, rotary = d->epilogue & MR_EPI_ROTARY_QKV;
  const bool dgelu = d->epilogue & MR_EPI_DGELU;
  if (dgelu) {
    // D = acc * gelu'(pre-activation): the dgrad of the MLP "out" projection, staged bf16 path
    MR_REQUIRE(d->epilogue == MR_EPI_DGELU && d->trans_b && !d->trans_a && d->seq_in == 0 && d->split_k <= 1 &&
                   d->algo == MR_GEMM_ALGO_TCGEN05, MR_ERR_BAD_SHAPE);
    MR_REQUIRE(in[3] && aligned16(in[3]), in[3] ? MR_ERR_MISALIGNED : MR_ERR_NULL_POINTER);
  }
  if (bias) MR_REQUIRE(in[2] && aligned16(in[2]), in[2] ? MR_ERR_MISALIGNED : MR_ERR_NULL_POINTER);
  if (residual) MR_REQUIRE(in[3] && aligned16(in[3]), in[3] ? MR_ERR_MISALIGNED : MR_ERR_NULL_POINTER);
  if (rotary) {
    MR_REQUIRE(in[4] && aligned16(in[4]), in[4] ? MR_ERR_MISALIGNED : MR_ERR_NULL_POINTER);
    MR_REQUIRE(d->rot_rows > 0 && d->hidden > 0 && d->hidden % 64 == 0 && d->N == 3 * d->hidden, MR_ERR_BAD_SHAPE);
  }
  if (d->seq_in != 0) MR_REQUIRE(d->seq_in > 0 && d->seq_out >= d->seq_in + d->row_off && d->row_off >= 0, MR_ERR_BAD_SHAPE);

  GemmParams p{};
  p.M = d->M;
  p.N = d->N;
  p.K = d->K;
  p.ldd = d->ldd;
  p.bias_on = bias;
  p.out_f32 = (d->epilogue & MR_EPI_OUT_F32) ? 1 : 0;
  p.seq_in = d->seq_in;
  p.seq_out = d->seq_out;
  p.row_off = d->row_off;
  p.rot_rows = d->rot_rows > 0 ? d->rot_rows : 1;
  p.hidden = d->hidden;
  p.q_scale = d->q_scale;
  p.bias = static_cast<const float*>(in[2]);
  p.residual = static_cast<const bf16*>(in[3]);
  p.rot = static_cast<const float*>(in[4]);
  p.D = out[0];
  p.splits = 1;
  const bool ln_in = d->epilogue & MR_EPI_LN_IN, stats_out = d->epilogue & MR_EPI_STATS_OUT;
  const bool direct_path = p.out_f32 || d->seq_in != 0;
  if (ln_in) {
    MR_REQUIRE(in[5] && in[6], MR_ERR_NULL_POINTER);
    MR_REQUIRE(aligned16(in[5]) && aligned16(in[6]), MR_ERR_MISALIGNED);
    MR_REQUIRE(d->K % 64 == 0 && d->N % 8 == 0 && !direct_path && d->algo == MR_GEMM_ALGO_TCGEN05, MR_ERR_BAD_SHAPE);
    p.stats_in = static_cast<const float2*>(in[5]);
    p.ln_s = static_cast<const float*>(in[6]);
    p.ln_eps = d->ln_eps;
  }
  if (stats_out) {
    MR_REQUIRE(out[1], MR_ERR_NULL_POINTER);
    MR_REQUIRE(aligned16(out[1]), MR_ERR_MISALIGNED);
    MR_REQUIRE(d->N % 64 == 0 && !direct_path && d->algo == MR_GEMM_ALGO_TCGEN05, MR_ERR_BAD_SHAPE);
    p.stats_out = static_cast<float2*>(out[1]);
  }

  if (d->algo == MR_GEMM_ALGO_SIMT_DEBUG) {
    const int64_t total = static_cast<int64_t>(d->M) * d->N;
    const int threads = 256;
    gemm_simt_debug_kernel<<<static_cast<unsigned>(ceil_div64(total, threads)), threads, 0, stream>>>(
        A, d->lda, W, p, gelu, residual, rotary);
    return after_launch();
  }
  MR_REQUIRE(d->algo == MR_GEMM_ALGO_TCGEN05, MR_ERR_BAD_SHAPE);

  // Tile width: 256 columns normally; 128 when that leaves SMs idle on a single wave.
  int sms = d->num_sms > 0 ? d->num_sms : device_sm_count();
  MR_REQUIRE(sms > 0, MR_ERR_UNSUPPORTED_ARCH);
  const int m_tiles = ceil_div(d->M, kBM);
  int bn = 256;
  if (d->N % 256 != 0 || static_cast<int64_t>(m_tiles) * ceil_div(d->N, 256) < sms) bn = 128;
  if (d->tile_n != 0) {
    MR_REQUIRE(d->tile_n == 128 || (d->tile_n == 256 && d->N % 256 == 0), MR_ERR_BAD_SHAPE);
    bn = d->tile_n;
  }
  p.num_n_tiles = ceil_div(d->N, bn);
  p.num_tiles = m_tiles * p.num_n_tiles;
  p.num_k_blocks = ceil_div(d->K, kBK);
  p.kb_per_split = p.num_k_blocks;
  const int splits = d->split_k > 1 ? (d->split_k < p.num_k_blocks ? d->split_k : p.num_k_blocks) : 1;
  if (splits > 1) {
    // partial sums go to the workspace (fp32, [splits][M][ldd]); a second kernel adds them up in a fixed order
    MR_REQUIRE(!(gelu || residual || rotary || bias || ln_in || stats_out), MR_ERR_BAD_SHAPE);
    MR_REQUIRE(d->seq_in == 0, MR_ERR_BAD_SHAPE);
    MR_REQUIRE(workspace && workspace_bytes >= mr_gemm_epilogue_workspace_bytes(d), MR_ERR_WORKSPACE);
    MR_REQUIRE(aligned16(workspace), MR_ERR_MISALIGNED);
    p.splits = splits;
    p.kb_per_split = ceil_div(p.num_k_blocks, splits);
    p.split_stride = static_cast<size_t>(d->M) * d->ldd;
    p.D = workspace;
    p.out_f32 = 1;
    p.num_tiles *= splits;
  }
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;

  Maps m;
  if (d->trans_a)
    MR_REQUIRE(make_tmap_2d(&m.a, A, d->K, d->M, d->lda, 64, 64), MR_ERR_DRIVER);
  else
    MR_REQUIRE(make_tmap_2d(&m.a, A, d->M, d->K, d->lda, kBM, kBK), MR_ERR_DRIVER);
  if (d->trans_b)
    MR_REQUIRE(make_tmap_2d(&m.b, W, d->K, d->N, ldb, 64, 64), MR_ERR_DRIVER);
  else
    MR_REQUIRE(make_tmap_2d(&m.b, W, d->N, d->K, ldb, bn, kBK), MR_ERR_DRIVER);
  MR_REQUIRE(!dgelu || (trans == 2 && !direct_path), MR_ERR_BAD_SHAPE);
  if (trans != 0) {
    MR_REQUIRE(!(gelu || residual || rotary || ln_in || stats_out) && (direct_path || !bias), MR_ERR_BAD_SHAPE);
    int32_t st;
    if (direct_path || splits > 1) {
      m.out = m.a;
      m.res = m.a;
      st = bn == 256 ? launch_trans<256>(stream, m, p, grid, trans, true) : launch_trans<128>(stream, m, p, grid, trans, true);
    } else {
      MR_REQUIRE(make_tmap_2d(&m.out, out[0], d->M, d->N, d->ldd, 32, 64), MR_ERR_DRIVER);
      m.res = m.out;
      if (dgelu) {
        MR_REQUIRE(make_tmap_2d(&m.res, in[3], d->M, d->N, d->ldd, 32, 64), MR_ERR_DRIVER);
        return bn == 256 ? launch_one<256, true, true, false, false, false, false, false, 1, 2>(stream, m, p, grid)
                         : launch_one<128, true, true, false, false, false, false, false, 1, 2>(stream, m, p, grid);
      }
      st = bn == 256 ? launch_trans<256>(stream, m, p, grid, trans, false) : launch_trans<128>(stream, m, p, grid, trans, false);
    }
    if (st != MR_OK || splits == 1) return st;
    const int64_t total = static_cast<int64_t>(d->M) * ((d->N + 3) / 4);
    const unsigned blocks = static_cast<unsigned>(ceil_div64(total, 256));
    if (d->epilogue & MR_EPI_OUT_F32)
      splitk_reduce_kernel<true><<<blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), out[0], d->M, d->N,
                                                             d->ldd, splits, p.split_stride);
    else
      splitk_reduce_kernel<false><<<blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), out[0], d->M, d->N,
                                                              d->ldd, splits, p.split_stride);
    return after_launch();
  }
  MR_REQUIRE(splits == 1, MR_ERR_BAD_SHAPE);   // split-K is offered for the backward (transposed-operand) GEMMs
  // fp32 output and row-remapped output keep per-thread global accesses; everything else is staged
  // through shared memory and moved by TMA
  const bool direct = direct_path;
  const Epi epi{gelu, residual, rotary, bias, ln_in, stats_out};
  if (direct) {
    m.out = m.a;
    m.res = m.a;
    if (bn == 256) return launch_direct<256>(stream, m, p, grid, epi);
    return launch_direct<128>(stream, m, p, grid, epi);
  }
  MR_REQUIRE(make_tmap_2d(&m.out, out[0], d->M, d->N, d->ldd, 32, 64), MR_ERR_DRIVER);
  m.res = m.out;
  if (residual) MR_REQUIRE(make_tmap_2d(&m.res, in[3], d->M, d->N, d->ldd, 32, 64), MR_ERR_DRIVER);
  // CTA pairs (cta_group::2, 256 x 256 pair tiles) whenever every pair of SMs gets at least one pair tile
  const int pair_tiles = ceil_div(d->M, 2 * kBM) * p.num_n_tiles;
  if (d->epilogue & MR_EPI_GELU_DUAL) {
    // bias + GELU with the pre-activation as a second bf16 output (out[2]); staged path, no other epilogue flag
    MR_REQUIRE(gelu && bias && !residual && !rotary && !ln_in && !stats_out && trans == 0, MR_ERR_BAD_SHAPE);
    MR_REQUIRE(out[2] && aligned16(out[2]), out[2] ? MR_ERR_MISALIGNED : MR_ERR_NULL_POINTER);
    MR_REQUIRE(make_tmap_2d(&m.res, out[2], d->M, d->N, d->ldd, 32, 64), MR_ERR_DRIVER);
    if (bn == 256 && d->algo_pair != 2 && sms >= 2 && (d->algo_pair == 1 || pair_tiles >= sms / 2)) {
      p.num_tiles = pair_tiles;
      const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
      MR_REQUIRE(make_tmap_2d(&m.b, W, d->N, d->K, d->K, bn / 2, kBK), MR_ERR_DRIVER);
      return launch_one<256, true, false, false, false, true, false, false, 2, 0, true>(stream, m, p, 2 * pairs);
    }
    if (bn == 256) return launch_one<256, true, false, false, false, true, false, false, 1, 0, true>(stream, m, p, grid);
    return launch_one<128, true, false, false, false, true, false, false, 1, 0, true>(stream, m, p, grid);
  }
  if (bn == 256 && trans == 0 && d->algo_pair != 2 && sms >= 2 && (d->algo_pair == 1 || pair_tiles >= sms / 2)) {
    p.num_tiles = pair_tiles;
    const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
    MR_REQUIRE(make_tmap_2d(&m.b, W, d->N, d->K, d->K, bn / 2, kBK), MR_ERR_DRIVER);
    return launch_staged<256, 2>(stream, m, p, 2 * pairs, epi);
  }
  if (bn == 256) return launch_staged<256>(stream, m, p, grid, epi);
  return launch_staged<128>(stream, m, p, grid, epi);
}
