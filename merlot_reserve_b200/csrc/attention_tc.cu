// Tensor-core (tcgen05) flash attention forward for head dim 64, used when T > 64.
//
// Persistent kernel, one CTA (16 warps) per SM.  A work item is one (sequence, head, PAIR of 128-query tiles); the
// two query tiles of an item share every K/V tile and are processed by two independent softmax warpgroups.  The
// SFU, not the tensor core, is the bound of this kernel at head dim 64 (128 x 128 exp2 per 2 x 2^21 MACs), so the
// design goal is that one of the two tiles is always in its ex2 phase.
//
//   warps 0-3   softmax, tile 0: thread = query row.  tcgen05.ld S (one x128 load) -> (mask) -> row max -> one FFMA
//   warps 4-7   softmax, tile 1  + ex2 per score (base-2 domain, fp32) -> row sum -> bf16 P -> tcgen05.st.  The
//                              running max is only re-based ("lazy rescale") when a block's row max exceeds it by
//                              more than 2^8, in which case O is corrected in TMEM.  After its last key block the
//                              thread publishes (l, m) in smem and moves on to the next item: it never waits for
//                              the last PV MMA nor touches global memory.
//   warps 8-11  epilogue      : per item and tile, once the last PV has retired: tcgen05.ld O -> O / l -> bf16 ->
//                              swizzled smem rows -> one TMA store per warp (3-D map [B, T, H], rows past T are
//                              clipped).  Runs under the softmax of the next item's first key block (that block's
//                              PV is the first MMA that needs O again).
//   warp 12     TMA producer : Q pair (double buffered across items) and a 3-stage ring of K|V tiles
//                              [128 keys x 64] x 2, 128B swizzle; runs ahead across items
//   warps 13,14 MMA issuers  : one per query tile w, independent of each other.
//                              S_w = Q_w K^T   (SS form, 128 x 128 x 64, TMEM cols S0 [0,128) / S1 [128,256))
//                              O_w += P_w V    (TS form: P_w read from TMEM as bf16 pairs, cols [256,320) /
//                              [320,384); V is the MN-major B operand straight from the TMA tile; O_w in TMEM
//                              cols [384,448) / [448,512)).  The key blocks of all items form one stream and
//                              S of block g+1 is always issued before PV of block g (also across items).
//
// SFU arbitration: the softmax warps (w, q) and (1 - w, q) sit on the same SM sub-partition and share its SFU.
// Explicit hand-over schemes (strict per-warp turn-taking with named barriers, a per-sub-partition lock, and a
// CTA-wide token taken by a whole tile per key block) were all measured SLOWER than letting the two warps
// contend freely: the pipe interleaves their ex2 at full rate, and every scheme serialises the ~200-cycle
// hand-over into the critical path (profiles/attention_r01_notes.md).  What is kept is a one-sided STAGGER: in
// the first key block of every two-tile item, tile 1 may only start its ex2 phase once tile 0 has issued all of
// its own (a monotonic smem counter).  The phase offset between two free-running tiles is neutrally stable --
// contention slows both equally -- so one nudge per item keeps tile 1 half a block behind tile 0: its ex2 run
// under tile 0's TMEM round trip, row max, PV wait and P store, and vice versa.
//
// Register budget (setmaxnreg; pool = 512 x 128): softmax 192, epilogue 88, producer / MMA issuers 40.
//
// Masks follow mr_attention_fwd's contract: key validity + video_src group ids; a query row whose mask is
// all-false attends uniformly over all T keys (the reference's -1e10 bias behaviour, modeling.py:343-358).
// Key blocks whose 128 keys are all valid (and no group ids are given) take a mask-free fast path; there an
// invalid query row simply uses a zero score scale, which yields the uniform distribution.  Otherwise, without
// group ids, the mask is an additive 0 / -inf per-key bias folded into one FFMA per score.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kQ = 128;
constexpr int kKVMax = 128;                // keys per block: template parameter kKV = 128, or 96 when that pads T less
constexpr int kD = 64;
constexpr int kStagesKV = 3;
constexpr int kTcThreads = 512;            // 8 softmax warps + 4 epilogue warps + {producer, 2 MMA issuers, 1 idle}
constexpr int kEpiWarp0 = 8, kProducerWarp = 12, kMmaWarp = 13;
constexpr int kTileBytes = 128 * kD * 2;   // 16 KB
constexpr int kTmemCols = 512;
constexpr uint32_t kColS = 0, kColP = 256, kColO = 384;   // + 128 / 64 / 64 per warpgroup
constexpr int kInfoOob = static_cast<int>(0x80000000u);
constexpr int kInfoInvalid = static_cast<int>(0x80000001u);
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

// shared memory map (offsets from the 1024-aligned base)
constexpr int kOffQ = 0;                                        // [2 item buffers][2 tiles]
constexpr int kOffKV = 4 * kTileBytes;                          // [kStagesKV][K, V]
constexpr int kOffO = kOffKV + kStagesKV * 2 * kTileBytes;      // [2 tiles][4 epilogue warps][32 rows x 128 B]
constexpr int kOffInfo = kOffO + 2 * kTileBytes;                // [8 softmax warps][kKVMax] {int info, float bias_v, float bias_e}
constexpr int kOffLM = kOffInfo + 8 * 3 * kKVMax * 4;           // [2 item parity][2 tiles][128 rows] float2 (l, m)
constexpr int kOffBar = kOffLM + 2 * 2 * kQ * 8;
constexpr int kNumBars = 4 + 2 * kStagesKV + 14;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;   // + tmem slot

struct AttnParams {
  const uint8_t* valid;
  const int32_t* group;
  bf16* out;
  float* lse;   // optional [B, nh, T]: natural-log row log-sum-exp of the masked scores (for mr_attention_bwd)
  int B, T, H, nh, nqb, npairs, nkb;
  int num_items;
  int stagger;        // 1: keep the two tiles of an item out of phase (see the header)
  long long* trace;   // diagnostics (mr_debug_attention_trace): clock64 stamps of CTA 0
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

// Items are visited with stride gridDim.x; (b, h, qp) advance by a fixed mixed-radix step, no divisions.
struct ItemWalker {
  Item it;
  int step_qp, step_h, step_b, npairs, nh;
  __device__ __forceinline__ ItemWalker(const AttnParams& p) {
    it = decode_item(static_cast<int>(blockIdx.x) < p.num_items ? static_cast<int>(blockIdx.x) : 0, p);
    npairs = p.npairs;
    nh = p.nh;
    step_qp = static_cast<int>(gridDim.x) % npairs;
    step_h = (static_cast<int>(gridDim.x) / npairs) % nh;
    step_b = (static_cast<int>(gridDim.x) / npairs) / nh;
  }
  __device__ __forceinline__ void advance() {
    it.qp += step_qp;
    const int c0 = it.qp >= npairs ? 1 : 0;
    it.qp -= c0 * npairs;
    it.h += step_h + c0;
    const int c1 = it.h >= nh ? 1 : 0;
    it.h -= c1 * nh;
    it.b += step_b + c1;
  }
};

// kKV = keys per block.  128 normally; 96 when ceil(T / 96) * 96 is at least 10 % below ceil(T / 128) * 128 (e.g. the
// multiple-choice joint sequences, T = 144 + 144 = 288 = 3 x 96: a 128-key tail block would spend a quarter of the
// kernel's exponentials on padding).  `tmap_kv` is the qkv map with a box of kKV rows.
// kGroups (implies kMasked): the call carries video_src group ids.  A separate instantiation, so that the validity-only
// kernel keeps its schedule (the group logic costs it 5-13 % when it is merely compiled in).
template <bool kMasked, int kKV, bool kGroups = false>
__global__ void __launch_bounds__(kTcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_out, const AttnParams p) {
  static_assert(kKV == 128 || kKV == 96, "key block of 128 or 96");
  constexpr int kKvBytes = kKV * kD * 2;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array: an integer round trip makes the pointer generic and
  // every access through it LD.E / ST.E instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sKV = smem + kOffKV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;                    // [2]
  uint64_t* q_empty = bars + 2;                   // [2]
  uint64_t* kv_full = bars + 4;                   // [kStagesKV]
  uint64_t* kv_empty = kv_full + kStagesKV;       // [kStagesKV]
  uint64_t* s_full = kv_empty + kStagesKV;        // [2] per tile from here on
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* pv_done = p_full + 2;
  uint64_t* o_full = pv_done + 2;                 // last PV of an item retired: O is final
  uint64_t* o_free = o_full + 2;                  // epilogue has read O
  uint64_t* lm_full = o_free + 2;                 // softmax published (l, m) of the item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lm_full + 2);
  volatile int* stagger = reinterpret_cast<volatile int*>(tmem_slot + 2);   // monotonic: 4 per two-tile item

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4);
      mbar_init(&lm_full[i], 4);
    }
    for (int s = 0; s < kStagesKV; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    *stagger = 0;
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

  if (warp_idx >= kProducerWarp) {
  setmaxnreg_dec<40>();
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
          mbar_arrive_expect_tx(&kv_full[s], 2 * kKvBytes);
          uint8_t* dst = sKV + s * 2 * kTileBytes;
          tma_load_2d(&tmap_kv, &kv_full[s], dst, p.H + it.h * kD, seq_base + j * kKV, kEvictLast);
          tma_load_2d(&tmap_kv, &kv_full[s], dst + kTileBytes, 2 * p.H + it.h * kD, seq_base + j * kKV, kEvictLast);
        }
      }
    }
  } else if (warp_idx == kMmaWarp || warp_idx == kMmaWarp + 1) {
    // ===================== MMA issuers: one warp per query tile =====================
    // Two independent issuers, so a late softmax of one tile never delays the other tile's MMAs.  Shared
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
            umma_commit(&o_full[w]);
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
  } else if (warp_idx >= kEpiWarp0) {
    setmaxnreg_dec<88>();
    // ===================== epilogue warpgroup: O / l -> bf16 -> smem -> TMA store =====================
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t n_tile[2] = {0, 0};   // items finished per tile (phase of o_full / lm_full)
    uint32_t n_item = 0;
    ItemWalker walk(p);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, walk.advance(), ++n_item) {
      const Item it = walk.it;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int q0 = it.qp * 2 * kQ + w * kQ;
        if (q0 >= p.T) continue;
        const int q = q0 + row;
        uint8_t* sOw = smem + kOffO + (w * 4 + quarter) * (32 * 128);
        mbar_wait(&lm_full[w], n_tile[w] & 1);
        const float2 lm = *reinterpret_cast<const float2*>(smem + kOffLM + (((n_item & 1) * 2 + w) * kQ + row) * 8);
        mbar_wait(&o_full[w], n_tile[w] & 1);
        tc_fence_after();
        ++n_tile[w];
        const float inv = lm.x > 0.f ? rcp_approx(lm.x) : 0.f;
        uint4 o16[kD / 8];
#pragma unroll
        for (int c = 0; c < kD / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + kColO + w * 64 + 32 * c, v);
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
        if (lane == 0) mbar_arrive(&o_free[w]);      // the next item's first PV may overwrite O
        if (p.lse != nullptr && q < p.T)
          p.lse[(static_cast<size_t>(it.b) * p.nh + it.h) * p.T + q] =
              (lm.y + (lm.x > 0.f ? log2f(lm.x) : 0.f)) * 0.6931471805599453f;
        if (lane == 0) tma_store_wait_read<1>();     // the store that last used these staging rows has read them
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kD / 8; ++i)
          *reinterpret_cast<uint4*>(sOw + lane * 128 + ((i ^ (lane & 7)) << 4)) = o16[i];   // 128B swizzle, conflict-free
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (q0 + quarter * 32 < p.T) tma_store_3d(&tmap_out, sOw, it.h * kD, q0 + quarter * 32, it.b);
          tma_store_commit();                        // (possibly empty) group: keeps wait_read<1> = "two stores ago"
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();              // smem must outlive the last store
  } else {
    setmaxnreg_inc<192>();
    // ===================== softmax warpgroups (thread = query row) =====================
    const int wg = warp_idx >> 2;                  // query tile of the pair
    const int quarter = warp_idx & 3;              // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t colS = kColS + wg * 128, colP = kColP + wg * 64, colO = kColO + wg * 64;
    int* myInfo = reinterpret_cast<int*>(smem + kOffInfo) + warp_idx * 3 * kKVMax;   // per warp: no cross-warp barrier
    float* myBiasV = reinterpret_cast<float*>(myInfo + kKVMax);                       // 0 / -inf: key valid and in range
    float* myBiasE = myBiasV + kKVMax;                                                // 0 / -inf: key in range
    uint32_t n_blk = 0;    // key blocks processed so far by this tile (phase of s_full / p_full / pv_done)
    uint32_t n_item = 0;   // items of this CTA so far (parity of the (l, m) hand-over buffer)
    int n_pair = 0;        // two-tile items so far

    long long* tr = (p.trace != nullptr && blockIdx.x == 0 && row == 0) ? p.trace + wg * 64 * 8 : nullptr;
    auto stamp = [&](int k) {
      if (tr != nullptr && n_blk < 64) tr[n_blk * 8 + k] = clock64();
    };

    // Raw mask bytes / group ids of the 4 keys (128 j + 4 lane + i) this lane fetches for its warp.  They are
    // requested one key block ahead and only decoded at the top of the next block, so the loads stay in flight.
    struct Raw {
      uint8_t v[4];
      int g[4];
    };
    auto load_raw = [&](int seq_base, int j) -> Raw {
      Raw r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kidx = min(j * kKV + min(lane * 4 + i, kKV - 1), p.T - 1);
        r.v[i] = p.valid ? p.valid[seq_base + kidx] : uint8_t(1);
        r.g[i] = kGroups ? p.group[seq_base + kidx] : 0;
      }
      return r;
    };

    ItemWalker walk(p);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, walk.advance(), ++n_item) {
      const Item it = walk.it;
      const int seq_base = it.b * p.T;
      const int q0 = it.qp * 2 * kQ + wg * kQ;
      if (q0 >= p.T) continue;                     // this pair has a single tile (warpgroup-uniform)
      const bool two_tiles = p.stagger != 0 && (it.qp * 2 * kQ + kQ < p.T);
      const int q = q0 + row;
      const bool q_in = q < p.T;
      bool q_uniform = false;
      int q_group = 0;
      Raw raw_next{};
      if (kMasked) {
        const bool v = q_in ? (p.valid ? p.valid[seq_base + q] != 0 : true) : false;
        q_uniform = !v;
        q_group = (kGroups && v) ? p.group[seq_base + q] : 0;
        raw_next = load_raw(seq_base, 0);
      }
      // Group ids: the 32 query rows of a warp almost always share ONE video_src id (ids change a few times per
      // sequence).  Such a warp folds "key valid AND key group == g0" into the additive per-key bias and keeps the
      // one-FFMA path; only warps that straddle a change compare per score.
      int g0 = 0;
      bool warp_uni = true;
      if (kGroups) {
        const unsigned vm = __ballot_sync(0xffffffffu, !q_uniform);
        g0 = vm != 0u ? __shfl_sync(0xffffffffu, q_group, __ffs(vm) - 1) : 0;
        warp_uni = __all_sync(0xffffffffu, q_uniform || q_group == g0);
      }
      const bool bias_path = !kGroups || warp_uni;
      const float cq = q_uniform ? 0.f : kLog2e;   // fast-path score scale: 0 -> uniform attention
      const float qs = q_uniform ? 0.f : 1.f;
      float m2 = 0.f;  // reference max, log2 domain
      float l = 0.f;

      for (int j = 0; j < p.nkb; ++j, ++n_blk) {
        const int nvk = min(kKV, p.T - j * kKV);   // keys of this block that exist
        bool fast = (nvk == kKV);
        if (kMasked) {
          int info[4];
          float bv[4], be[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool in_range = j * kKV + lane * 4 + i < p.T;
            info[i] = !in_range ? kInfoOob : (raw_next.v[i] != 0 ? raw_next.g[i] : kInfoInvalid);
            bv[i] = info[i] == g0 ? 0.f : -INFINITY;
            be[i] = in_range ? 0.f : -INFINITY;
          }
          if (j + 1 < p.nkb) raw_next = load_raw(seq_base, j + 1);
          const bool mine = lane * 4 < kKV;          // kKV = 96: lanes 24..31 hold no key of the block
          if (mine) {
            *reinterpret_cast<int4*>(myInfo + 4 * lane) = make_int4(info[0], info[1], info[2], info[3]);
            *reinterpret_cast<float4*>(myBiasV + 4 * lane) = make_float4(bv[0], bv[1], bv[2], bv[3]);
            *reinterpret_cast<float4*>(myBiasE + 4 * lane) = make_float4(be[0], be[1], be[2], be[3]);
          }
          // fast path: every key in range, valid and of the warp's group
          const bool ok = !mine || (bias_path && info[0] == g0 && info[1] == g0 && info[2] == g0 && info[3] == g0);
          fast = __all_sync(0xffffffffu, ok);      // also orders the smem writes above before the reads below
        }
        stamp(0);
        mbar_wait(&s_full[wg], n_blk & 1);
        tc_fence_after();
        stamp(1);
        float sc[kKV];
        {
          uint32_t v[kKV];
          if constexpr (kKV == 128)
            tmem_ld_32x128_wait(lane_addr + colS, v);
          else
            tmem_ld_32x96_wait(lane_addr + colS, v);
#pragma unroll
          for (int i = 0; i < kKV; ++i) sc[i] = __uint_as_float(v[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[wg]);
        stamp(2);

        float cs = cq;   // score scale used for this block
        if (!fast) {
          if (kMasked) {
            cs = kLog2e;
            if (bias_path) {
              // additive per-key bias (0 / -inf) folded into one FFMA per score.  A valid query row reads the
              // "key valid and in range" vector with scale 1; an invalid query row (it attends uniformly to every
              // existing key, modeling.py:343-358) reads the "key in range" vector with scale 0.  The choice is per
              // THREAD (a pointer), not per score: at most two broadcast addresses per warp and no selects.
              const float* myBias = q_uniform ? myBiasE : myBiasV;
#pragma unroll
              for (int c4 = 0; c4 < kKV / 4; ++c4) {
                const float4 b = *reinterpret_cast<const float4*>(myBias + 4 * c4);
                sc[4 * c4 + 0] = fmaf(sc[4 * c4 + 0], qs, b.x);
                sc[4 * c4 + 1] = fmaf(sc[4 * c4 + 1], qs, b.y);
                sc[4 * c4 + 2] = fmaf(sc[4 * c4 + 2], qs, b.z);
                sc[4 * c4 + 3] = fmaf(sc[4 * c4 + 3], qs, b.w);
              }
            } else {
              // group ids (video_src): per-thread comparison against the query's own group
#pragma unroll
              for (int c4 = 0; c4 < kKV / 4; ++c4) {
                const int4 f = *reinterpret_cast<const int4*>(myInfo + 4 * c4);
                const int ff[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int c = 4 * c4 + i;
                  const bool same = ff[i] == q_group, exists = ff[i] != kInfoOob;
                  const bool keep = q_uniform ? exists : same;
                  sc[c] = keep ? sc[c] * qs : -INFINITY;
                }
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
        for (int c = 0; c < kKV / 8; ++c) {
#pragma unroll
          for (int i = 0; i < 4; ++i) mxp[i] = fmax3(mxp[i], sc[(kKV / 4) * i + 2 * c], sc[(kKV / 4) * i + 2 * c + 1]);
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
        stamp(3);
        const float neg_m = -m2;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
        uint32_t pk[64];                       // kKV / 2 packed pairs; the tail of a 96-key block's store is never read
        if constexpr (kKV < 128) {
#pragma unroll
          for (int c = kKV / 2; c < 64; ++c) pk[c] = 0u;
        }
#pragma unroll
        for (int c = 0; c < kKV; ++c) sc[c] = fmaf(sc[c], cs, neg_m);
        // Phase stagger (see the header): in the first key block of an item, tile 1 starts its ex2 phase only after
        // tile 0 has issued all of its own.  Already staggered tiles pass without waiting.
        if (two_tiles && j == 0 && wg == 1) {
          if (lane == 0) {
            while (*stagger < 4 * (n_pair + 1)) __nanosleep(20);
          }
          __syncwarp();
        }
        // Software pipelined in batches of 8: the ex2 results of one batch are only summed / packed while the
        // next batch's ex2 are in flight (the SFU latency is ~40 cycles; a consumer right behind its ex2 stalls
        // the warp).
#pragma unroll
        for (int b = 0; b <= kKV / 8; ++b) {
          if (b < kKV / 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sc[8 * b + i] = ex2_approx_v(sc[8 * b + i]);
          }
          if (b == kKV / 8 && two_tiles && j == 0 && wg == 0) {
            __syncwarp();
            if (lane == 0) atomicAdd(const_cast<int*>(stagger), 1);
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
        stamp(4);
        // P (and O) may only be touched once the previous PV of this tile has retired -- also across items, the
        // softmax no longer waits for an item's last PV anywhere else.  Waiting only now keeps that MMA's latency
        // under the ex2 phase.
        if (n_blk > 0) {
          mbar_wait(&pv_done[wg], (n_blk - 1) & 1);
          tc_fence_after();
          stamp(5);
          if (j > 0 && __any_sync(0xffffffffu, factor != 1.f)) {
            uint32_t v[kD];
            tmem_ld_32x64(lane_addr + colO, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < kD; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * factor);
            tmem_st_32x64(lane_addr + colO, v);
          }
        }
        stamp(6);
        tmem_st_32x64(lane_addr + colP, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[wg]);
        stamp(7);
      }
      if (two_tiles) ++n_pair;
      // hand (l, m) to the epilogue warpgroup and move on
      *reinterpret_cast<float2*>(smem + kOffLM + (((n_item & 1) * 2 + wg) * kQ + row) * 8) = make_float2(l, m2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&lm_full[wg]);
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

long long* g_attention_trace = nullptr;

int32_t launch_attention_tc3(cudaStream_t stream, const bf16* qkv, const uint8_t* valid, bf16* out, float* lse, int B,
                             int T, int H);   // attention_tc3.cu

// Called by mr_attention_fwd (attention.cu) when T is long enough for 128-wide tiles.
// variant: 0 = choose; 2 = this file's two-tile kernel; 3 = the three-stream kernel of attention_tc3.cu (calls with
// group ids always take the two-tile kernel).  Same-box probe (profiles/attn_tc3_r02.txt, us, two-tile -> three-stream):
// T = 241: 346 -> 369;  T = 288: 705 -> 457;  T = 577: 697 -> 644;  T = 640: 132 -> 117;  T = 1120: 493 -> 422;
// T = 2464: 793 -> 665.  Three streams win from five 64-key blocks per tile on; below that the per-tile hand-overs
// (Q refill, one epilogue warpgroup for three streams) cost more than the third softmax warp per scheduler gains.
int32_t launch_attention_tc(cudaStream_t stream, const bf16* qkv, const uint8_t* valid, const int32_t* group,
                            bf16* out, float* lse, int B, int T, int H, int variant) {
  static const int tiles_env = []() {   // MR_ATTN_TILES=2 / 3 overrides the choice for A/B runs
    const char* e = getenv("MR_ATTN_TILES");
    return e ? atoi(e) : 0;
  }();
  if (variant == 0) variant = tiles_env == 2 || tiles_env == 3 ? tiles_env : (T > 256 ? 3 : 2);
  if (variant == 3 && group == nullptr) return launch_attention_tc3(stream, qkv, valid, out, lse, B, T, H);
  // key block: 96 where that pads the sequence at least 10 % less than 128 does (T = 288 -> 3 x 96 instead of 3 x 128)
  const int pad128 = ceil_div(T, 128) * 128, pad96 = ceil_div(T, 96) * 96;
  static const int kv_env = []() {
    const char* e = getenv("MR_ATTN_KV");
    return e ? atoi(e) : 0;
  }();
  const int kv = kv_env == 96 || kv_env == 128 ? kv_env : (pad96 * 10 <= pad128 * 9 ? 96 : 128);
  CUtensorMap tm, tm_kv;
  if (!make_tmap_2d(&tm, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, 128, kD)) return MR_ERR_DRIVER;
  if (!make_tmap_2d(&tm_kv, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, kv, kD)) return MR_ERR_DRIVER;
  CUtensorMap tm_out;   // [B, T, H]: a query tile that overhangs its sequence is clipped at T by the TMA store
  if (!make_tmap_3d(&tm_out, out, H, T, B, H, static_cast<int64_t>(T) * H, 32, kD)) return MR_ERR_DRIVER;
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
  p.nkb = ceil_div(T, kv);
  const int64_t items = static_cast<int64_t>(B) * p.nh * p.npairs;
  if (items >= (int64_t(1) << 31)) return MR_ERR_BAD_SHAPE;
  p.num_items = static_cast<int>(items);
  // measured (tools/attn_probe.py, same box): the stagger pays for two-key-block items (vision T = 241: -7 %), is
  // neutral for 5..20 key blocks; MR_ATTN_STAGGER=0/1 overrides for experiments
  static const int stagger_env = []() {
    const char* e = getenv("MR_ATTN_STAGGER");
    return e ? atoi(e) : -1;
  }();
  p.stagger = stagger_env >= 0 ? stagger_env : (p.nkb <= 2 ? 1 : 0);
  p.trace = g_attention_trace;
  const int sms = device_sm_count();
  if (sms <= 0) return MR_ERR_UNSUPPORTED_ARCH;
  const int grid = p.num_items < sms ? p.num_items : sms;
  const bool masked = valid != nullptr || group != nullptr;
  const bool groups = group != nullptr;
  auto kern = kv == 96 ? (groups ? attention_tc_kernel<true, 96, true> : masked ? attention_tc_kernel<true, 96> : attention_tc_kernel<false, 96>)
                       : (groups ? attention_tc_kernel<true, 128, true> : masked ? attention_tc_kernel<true, 128> : attention_tc_kernel<false, 128>);
  static bool attr_set[6] = {false, false, false, false, false, false};
  const int slot = (kv == 96 ? 3 : 0) + (groups ? 2 : (masked ? 1 : 0));
  if (!attr_set[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set[slot] = true;
  }
  kern<<<grid, kTcThreads, kSmemBytes, stream>>>(tm, tm_kv, tm_out, p);
  return after_launch();
}

}  // namespace mr

// Diagnostics only: when set to a device buffer of 2 * 64 * 8 int64, CTA 0 of every following attention_tc launch
// records clock64() stamps of its softmax phases (tools/attn_trace.py prints the breakdown).  NULL switches it off.
extern "C" int32_t mr_debug_attention_trace(void* device_buffer) {
  mr::g_attention_trace = static_cast<long long*>(device_buffer);
  return MR_OK;
}
