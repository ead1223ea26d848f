// Tensor-core (tcgen05) flash attention forward for head dim 64, used when T > 64.
//
// Persistent kernel, one CTA per SM.  A work item is one (sequence, head, PAIR of 128-query tiles); the two
// query tiles of an item share every K/V tile and are processed by two independent softmax warpgroups, so
// the tensor core always has the other tile's MMAs to run while one warpgroup is in its exp2 phase (the
// SFU, not the tensor core, is the bound of this kernel at head dim 64: 128 x 128 exp2 per 2 x 2^21 MACs).
//
//   warp 8      TMA producer : Q pair (double buffered across items) and a 3-stage ring of K|V tiles
//                              [128 keys x 64] x 2, 128B swizzle; runs ahead across items
//   warps 9,10  MMA issuers  : one per query tile w, independent of each other (a late softmax of one tile
//                              never delays the other tile's MMAs).
//                              S_w = Q_w K^T   (SS form, 128 x 128 x 64, TMEM cols S0 [0,128) / S1 [128,256))
//                              O_w += P_w V    (TS form: P_w read from TMEM as bf16 pairs, cols [256,320) /
//                              [320,384); V is the MN-major B operand straight from the TMA tile; O_w in TMEM
//                              cols [384,448) / [448,512)).  The key blocks of all items form one stream and
//                              S of block g+1 is always issued before PV of block g (also across items).
//   warps 0-3   softmax, tile 0: thread = query row.  tcgen05.ld S -> (mask) -> row max -> one FFMA + ex2 per
//   warps 4-7   softmax, tile 1  score (base-2 domain, fp32) -> row sum -> bf16 P -> tcgen05.st.  The running
//                              max is only re-based ("lazy rescale") when a block's row max exceeds it by more
//                              than 2^8, in which case O is corrected in TMEM; any fixed shift within that
//                              range gives the same softmax in fp32/bf16.  Epilogue: O / l -> bf16 -> HBM.
//
// Register budget: a softmax thread holds a full 128-column score row, so the third warpgroup (producer, two
// MMA issuers, one idle warp) gives its registers away with setmaxnreg (88 vs 208 per thread; the pool is the
// CTA's launch allocation, 384 x 168).
//
// Measured alternatives (profiles/attention_r01_notes.md): strict SFU turn-taking between the two tiles, and
// two threads per row (16 softmax warps), were both slower than this layout.
//
// Masks follow mr_attention_fwd's contract: key validity + video_src group ids; a query row whose mask is
// all-false attends uniformly over all T keys (the reference's -1e10 bias behaviour, modeling.py:343-358).
// Key blocks whose 128 keys are all valid (and no group ids are given) take a mask-free fast path; there an
// invalid query row simply uses a zero score scale, which yields the uniform distribution.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kQ = 128;
constexpr int kKV = 128;
constexpr int kD = 64;
constexpr int kStagesKV = 3;
constexpr int kTcThreads = 384;            // 8 softmax warps + 1 warpgroup {producer, 2 MMA issuers, 1 idle}
constexpr int kProducerWarp = 8, kMmaWarp = 9;
constexpr int kTileBytes = 128 * kD * 2;   // 16 KB
constexpr int kTmemCols = 512;
constexpr uint32_t kColS = 0, kColP = 256, kColO = 384;   // + 128 / 64 / 64 per warpgroup
constexpr int kInfoOob = static_cast<int>(0x80000000u);
constexpr int kInfoInvalid = static_cast<int>(0x80000001u);
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr uint32_t kBarInfo = 1;           // named barriers 1, 2: mask info of tile 0 / 1 (128 threads)

// shared memory map (offsets from the 1024-aligned base)
constexpr int kOffQ = 0;                                   // [2 item buffers][2 tiles]
constexpr int kOffKV = 4 * kTileBytes;                     // [kStagesKV][K, V]
constexpr int kOffInfo = kOffKV + kStagesKV * 2 * kTileBytes;   // [2 wg][2][kKV] int
constexpr int kOffXch = kOffInfo + 2 * 2 * kKV * 4;         // [2 tiles][2 parity][2 halves][128 rows] float
constexpr int kOffBar = kOffXch + 2 * 2 * 2 * kQ * 4;
constexpr int kNumBars = 4 + 2 * kStagesKV + 10;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;

struct AttnParams {
  const uint8_t* valid;
  const int32_t* group;
  bf16* out;
  float* lse;   // optional [B, nh, T]: natural-log row log-sum-exp of the masked scores (for mr_attention_bwd)
  int B, T, H, nh, nqb, npairs, nkb;
  int num_items;
};

struct Item {
  int b, h, qp;
};
__device__ __forceinline__ Item decode_item(int item, const AttnParams& p) {
  Item it;
  it.qp = item % p.npairs;
  const int bh = item / p.npairs;
  it.h = bh % p.nh;
  it.b = bh / p.nh;
  return it;
}

template <bool kMasked>
__global__ void __launch_bounds__(kTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sKV = smem + kOffKV;
  int* sInfo = reinterpret_cast<int*>(smem + kOffInfo);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;                    // [2]
  uint64_t* q_empty = bars + 2;                   // [2]
  uint64_t* kv_full = bars + 4;                   // [kStagesKV]
  uint64_t* kv_empty = kv_full + kStagesKV;       // [kStagesKV]
  uint64_t* s_full = kv_empty + kStagesKV;        // [2] per warpgroup
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* pv_done = p_full + 2;
  uint64_t* o_free = pv_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_free[i], 4);
    }
    for (int s = 0; s < kStagesKV; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    fence_mbar_init();
  }
  if (warp_idx == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx >= 8) {
  setmaxnreg_dec<88>();
  if (warp_idx == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t n_item = 0, n_kv = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n_item) {
        const Item it = decode_item(item, p);
        const int seq_base = it.b * p.T;
        const int q0 = it.qp * 2 * kQ;
        const int ntiles = (q0 + kQ < p.T) ? 2 : 1;
        const int ib = n_item & 1;
        mbar_wait(&q_empty[ib], ((n_item >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[ib], ntiles * kTileBytes);
        for (int w = 0; w < ntiles; ++w)
          tma_load_2d(&tmap_qkv, &q_full[ib], sQ + (ib * 2 + w) * kTileBytes, it.h * kD, seq_base + q0 + w * kQ,
                      kEvictFirst);
        for (int j = 0; j < p.nkb; ++j, ++n_kv) {
          const int s = n_kv % kStagesKV;
          mbar_wait(&kv_empty[s], ((n_kv / kStagesKV) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * kTileBytes);
          uint8_t* dst = sKV + s * 2 * kTileBytes;
          tma_load_2d(&tmap_qkv, &kv_full[s], dst, p.H + it.h * kD, seq_base + j * kKV, kEvictLast);
          tma_load_2d(&tmap_qkv, &kv_full[s], dst + kTileBytes, 2 * p.H + it.h * kD, seq_base + j * kKV, kEvictLast);
        }
      }
    }
  } else if (warp_idx == kMmaWarp || warp_idx == kMmaWarp + 1) {
    // ===================== MMA issuers: one warp per query tile =====================
    // Two independent issuers, so a late softmax of one tile never delays the other tile's MMAs (a single
    // in-order issuer kept both tiles in lockstep, i.e. both in their ex2 phase at the same time).  Shared
    // resources are released by BOTH: kv_empty / q_empty count two arrivals; for items without a second
    // tile the tile-1 issuer only follows the producer (wait full, arrive empty).
    if (lane == 0) {
      const int w = warp_idx - kMmaWarp;
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32_ex(kQ, kKV, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16_f32_ex(kQ, kD, 1);  // B (= V) is MN-major
      const uint32_t tS = tmem_base + kColS + w * 128, tP = tmem_base + kColP + w * 64, tO = tmem_base + kColO + w * 64;
      uint32_t ns = 0, np = 0, ni = 0;   // S MMAs issued, PV MMAs issued, items with a tile for this warp
      // The key blocks of all items of this CTA form one stream; S of block g+1 is always issued before PV of
      // block g, also across item boundaries, so the softmax warps never wait for a cold S at an item start.
      struct Cursor {
        int item, j;
        uint32_t n_item, n_kv;
        bool mine;   // this warp's tile exists in the item
      };
      auto has_tile = [&](int item) { return w == 0 || (item % p.npairs) * 2 * kQ + kQ < p.T; };
      auto issue_s_block = [&](const Cursor& c) {
        if (c.j == 0) mbar_wait(&q_full[c.n_item & 1], (c.n_item >> 1) & 1);
        const uint32_t s = c.n_kv % kStagesKV;
        mbar_wait(&kv_full[s], (c.n_kv / kStagesKV) & 1);
        if (!c.mine) return;
        if (ns > 0) mbar_wait(&s_free[w], (ns - 1) & 1);   // softmax has S_{prev} in registers
        tc_fence_after();
        const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ + ((c.n_item & 1) * 2 + w) * kTileBytes));
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sKV + s * 2 * kTileBytes));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[w]);
        ++ns;
      };
      Cursor cur{static_cast<int>(blockIdx.x), 0, 0u, 0u, false};
      if (cur.item < p.num_items) {
        cur.mine = has_tile(cur.item);
        issue_s_block(cur);
      }
      while (cur.item < p.num_items) {
        Cursor nxt = cur;
        ++nxt.n_kv;
        if (++nxt.j == p.nkb) {
          nxt.j = 0;
          nxt.item += static_cast<int>(gridDim.x);
          ++nxt.n_item;
          nxt.mine = nxt.item < p.num_items ? has_tile(nxt.item) : false;
        }
        if (nxt.item < p.num_items) issue_s_block(nxt);
        const uint32_t s = cur.n_kv % kStagesKV;
        if (cur.mine) {
          const uint64_t vdesc = umma_desc_k_sw128(smem_u32(sKV + s * 2 * kTileBytes + kTileBytes));
          mbar_wait(&p_full[w], np & 1);
          if (cur.j == 0 && ni > 0) mbar_wait(&o_free[w], (ni - 1) & 1);   // previous item's O was read
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kKV / 16; ++kk) {
            // 16 keys = 16 rows x 128 B = 2048 B further into the V tile; P advances 8 TMEM columns
            umma_bf16_ts(tO, tP + 8 * kk, vdesc + 128 * kk, idesc_o, (cur.j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&pv_done[w]);
          ++np;
          umma_commit(&kv_empty[s]);   // this tile's S and PV MMAs on the stage have retired
          if (cur.j == p.nkb - 1) {
            umma_commit(&q_empty[cur.n_item & 1]);
            ++ni;
          }
        } else {
          mbar_arrive(&kv_empty[s]);     // nothing of this tile reads the stage
          if (cur.j == p.nkb - 1) mbar_arrive(&q_empty[cur.n_item & 1]);
        }
        cur = nxt;
      }
    }
  }
  } else {
    setmaxnreg_inc<208>();
    // ===================== softmax / correction / epilogue warpgroups (thread = query row) =====================
    const int wg = warp_idx >> 2;                  // query tile of the pair
    const int quarter = warp_idx & 3;              // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t colS = kColS + wg * 128, colP = kColP + wg * 64, colO = kColO + wg * 64;
    int* myInfo = sInfo + wg * 2 * kKV;
    uint32_t n_blk = 0;    // key blocks processed so far by this tile (phase of s_full / p_full / pv_done)

    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const Item it = decode_item(item, p);
      const int seq_base = it.b * p.T;
      const int q0 = it.qp * 2 * kQ + wg * kQ;
      if (q0 >= p.T) continue;                     // this pair has a single tile (warpgroup-uniform)
      const int q = q0 + row;
      const bool q_in = q < p.T;
      bool q_uniform = false;
      int q_group = 0;
      if (kMasked) {
        const bool v = q_in ? (p.valid ? p.valid[seq_base + q] != 0 : true) : false;
        q_uniform = !v;
        q_group = (v && p.group) ? p.group[seq_base + q] : 0;
      }
      const float cq = q_uniform ? 0.f : kLog2e;   // fast-path score scale: 0 -> uniform attention
      float m2 = 0.f;  // reference max, log2 domain
      float l = 0.f;

      for (int j = 0; j < p.nkb; ++j, ++n_blk) {
        const int nvk = min(kKV, p.T - j * kKV);   // keys of this block that exist
        bool fast = (nvk == kKV);
        if (kMasked) {
          const int kidx = j * kKV + row;
          int info = kInfoOob;
          if (kidx < p.T) {
            const bool v = p.valid ? p.valid[seq_base + kidx] != 0 : true;
            info = v ? (p.group ? p.group[seq_base + kidx] : 0) : kInfoInvalid;
          }
          myInfo[(j & 1) * kKV + row] = info;
          // fast path: every key in range and valid, and no group ids in play
          const int ok = (p.group == nullptr && info == 0) ? 1 : 0;
          fast = __syncthreads_and_named(ok, kBarInfo + wg, 128);
        }
        mbar_wait(&s_full[wg], n_blk & 1);
        tc_fence_after();
        float sc[kKV];
#pragma unroll
        for (int c = 0; c < kKV / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + colS + 32 * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sc[32 * c + i] = __uint_as_float(v[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[wg]);

        float cs = cq;   // score scale used for this block
        if (!fast) {
          if (kMasked) {
            // general path: per-key info words (group id / invalid / out of range)
            cs = kLog2e;
            const int* info = myInfo + (j & 1) * kKV;
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
          } else {
            // unmasked kernel: only the tail of the last key block is out of range; 32-column chunks that are
            // entirely in range are left alone (warp-uniform test)
#pragma unroll
            for (int c = 0; c < kKV / 32; ++c) {
              if (32 * c + 32 <= nvk) continue;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (32 * c + i >= nvk) sc[32 * c + i] = -INFINITY;
            }
          }
        }
        float mxp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains
#pragma unroll
        for (int c = 0; c < kKV / 4; ++c) {
#pragma unroll
          for (int i = 0; i < 4; ++i) mxp[i] = fmaxf(mxp[i], sc[32 * i + c]);
        }
        const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
        const float mcand = (mx == -INFINITY) ? -INFINITY : mx * cs;

        // ---- lazy re-basing of the reference max ----
        float factor = 1.f;
        if (j == 0) {
          m2 = (mcand == -INFINITY) ? 0.f : mcand;
        } else if (mcand > m2 + kRescaleThreshold) {
          factor = ex2_approx(m2 - mcand);
          m2 = mcand;
        }
        // ---- P = exp2(s * cs - m), row sum, bf16 pack ----
        const float neg_m = -m2;
        // Software pipelined in batches of 8: the ex2 results of one batch are only summed / packed while the
        // next batch's ex2 are in flight (the SFU latency is ~40 cycles; a consumer right behind its ex2 stalls
        // the warp, and with two softmax warps per scheduler nothing else hides it).
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
        uint32_t pk[kKV / 2];
#pragma unroll
        for (int c = 0; c < kKV; ++c) sc[c] = fmaf(sc[c], cs, neg_m);
#pragma unroll
        for (int b = 0; b <= kKV / 8; ++b) {
          if (b < kKV / 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sc[8 * b + i] = ex2_approx_v(sc[8 * b + i]);
          }
          if (b > 0) {
            const int o = 8 * (b - 1);
            sum0 += sc[o + 0] + sc[o + 4];
            sum1 += sc[o + 1] + sc[o + 5];
            sum2 += sc[o + 2] + sc[o + 6];
            sum3 += sc[o + 3] + sc[o + 7];
#pragma unroll
            for (int i = 0; i < 4; ++i) pk[4 * (b - 1) + i] = pack_bf16(sc[o + 2 * i], sc[o + 2 * i + 1]);
          }
        }
        l = l * factor + ((sum0 + sum1) + (sum2 + sum3));
        // P and O may only be touched once PV_{j-1} has retired; waiting only now keeps that MMA's latency
        // under the exp2 phase
        if (j > 0) {
          mbar_wait(&pv_done[wg], (n_blk - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, factor != 1.f)) {
#pragma unroll
            for (int c = 0; c < kD / 32; ++c) {
              uint32_t v[32];
              tmem_ld_32x32(lane_addr + colO + 32 * c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * factor);
              tmem_st_32x32(lane_addr + colO + 32 * c, v);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < kKV / 64; ++c) {
          uint32_t (&pc)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pk[32 * c]);
          tmem_st_32x32(lane_addr + colP + 32 * c, pc);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[wg]);
      }

      // ---- epilogue: O / l -> bf16 -> global ----
      mbar_wait(&pv_done[wg], (n_blk - 1) & 1);
      tc_fence_after();
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      uint4 o16[kD / 8];
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + colO + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o16[4 * c + i].x = pack_bf16(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          o16[4 * c + i].y = pack_bf16(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          o16[4 * c + i].z = pack_bf16(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          o16[4 * c + i].w = pack_bf16(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[wg]);     // the next item's first PV may overwrite O
      if (p.lse != nullptr && q_in)
        p.lse[(static_cast<size_t>(it.b) * p.nh + it.h) * p.T + q] = (m2 + (l > 0.f ? log2f(l) : 0.f)) * 0.6931471805599453f;
      if (q_in) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(seq_base) + q) * p.H + it.h * kD);
#pragma unroll
        for (int i = 0; i < kD / 8; ++i) dst[i] = o16[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// Called by mr_attention_fwd (attention.cu) when T is long enough for 128-wide tiles.
int32_t launch_attention_tc(cudaStream_t stream, const bf16* qkv, const uint8_t* valid, const int32_t* group,
                            bf16* out, float* lse, int B, int T, int H) {
  CUtensorMap tm;
  if (!make_tmap_2d(&tm, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, 128, kD)) return MR_ERR_DRIVER;
  AttnParams p{};
  p.valid = valid;
  p.group = group;
  p.out = out;
  p.lse = lse;
  p.B = B;
  p.T = T;
  p.H = H;
  p.nh = H / kD;
  p.nqb = ceil_div(T, kQ);
  p.npairs = ceil_div(p.nqb, 2);
  p.nkb = ceil_div(T, kKV);
  const int64_t items = static_cast<int64_t>(B) * p.nh * p.npairs;
  if (items >= (int64_t(1) << 31)) return MR_ERR_BAD_SHAPE;
  p.num_items = static_cast<int>(items);
  const int sms = device_sm_count();
  if (sms <= 0) return MR_ERR_UNSUPPORTED_ARCH;
  const int grid = p.num_items < sms ? p.num_items : sms;
  const bool masked = valid != nullptr || group != nullptr;
  auto kern = masked ? attention_tc_kernel<true> : attention_tc_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[masked]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set[masked] = true;
  }
  kern<<<grid, kTcThreads, kSmemBytes, stream>>>(tm, p);
  return after_launch();
}

}  // namespace mr
