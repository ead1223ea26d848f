// Tensor-core (tcgen05 / TMEM) attention backward for head dim 64, used when T > 64 (mr_attention_bwd).
//
// Work item = (sequence, head, block j of 128 keys); an item loops over the blocks i of 128 queries.  The kernel is
// PERSISTENT (one CTA per SM walks items blockIdx.x, blockIdx.x + gridDim.x, ...): a CTA that lives for a single item
// pays ~8 us of launch, TMEM allocation, cold K / V / Q loads and dK / dV write-out against 2.5 us per block
// (profiles/attn_bwd_blocks_r02w.txt: equal work at 2 / 4 / 8 / 16 blocks per CTA), which is most of the time at
// T = 241 (two blocks per item).  Here K / V are double buffered ACROSS items, every barrier runs on a block counter
// that does not reset between items, and the dK / dV write-out of item n (drain warps, TMA stores) runs under the
// first block of item n + 1.
//
// Everything is kept TRANSPOSED -- keys on the TMEM lanes, queries on the columns -- so that the probabilities and
// their gradients are the TMEM (A) operands of the two products that accumulate over the query blocks:
//
//   S^T  = K_j Q_i^T            SS  A = K_j  [128 keys x 64]  K-major      B = Q_i   K-major     -> TMEM [0,128)
//   dP^T = V_j dO_i^T           SS  A = V_j                                B = dO_i              -> TMEM [128,256)
//   P^T  = exp2(S^T log2e - lse_i) under the mask;   dS^T = P^T (dP^T - D_i)       (compute warps, thread = key)
//   dV_j += P^T  dO_i           TS  A = P^T  bf16 in TMEM [256,320)        B = dO_i  MN-major    -> TMEM [320,384)
//   dK_j += dS^T Q_i            SS  A = dS^T staged in smem, K-major       B = Q_i   MN-major    -> TMEM [384,448)
//   dQ_i  = dS   K_j            SS  A = the same smem tile as MN-major A   B = K_j   MN-major    -> TMEM [448,512)
//
// Pipelining inside an item (r02): S^T and dP^T are produced in two 64-query HALVES (N = 64 MMAs) with their own
// barriers.  As soon as the compute warps have pulled half h of block g into registers they release it, and the
// issuer immediately computes half h of block g+1 into the same columns -- under the compute of the other half -- so
// the compute warps never wait for a cold S^T / dP^T and the dV / dK / dQ MMAs of block g run under the compute of
// block g+1.  dS^T lives only in a double-buffered shared-memory tile that both of its consumers read.  dQ_i leaves
// through TMA reduce-add boxes into the fp32 workspace that mr_attention_bwd converts to bf16 afterwards, because
// the key blocks of a sequence are different items.  lse_i and D_i are per-COLUMN quantities in this orientation: the
// statistics warps turn them (and the query side of the mask) into five broadcast vectors per query block:
//     p = exp2( s a[c] + cb[c] - 1e30 nb[c] |ktag - qtag[c]| ),   dS = p (dP - D[c])
//   a  = log2e (valid query) or 0 (invalid query: uniform attention, S := 0);  cb = -lse log2e, or -inf past T;
//   nb = 1 for valid queries (mask applies) / 0 for invalid ones;  ktag = the key's group id, 1e38 if the key is
//   padding (never equals a query tag; 1e38 x 1e30 overflows to -inf, and x 0 stays 0 for the uniform rows); the
//   statistics warps also fetch ktag (a sixth vector, per item) so that no compute warp ever waits on global memory.
//
//   warps 0-3   compute, half 0 (queries 0..63 of every block): thread = key.  tcgen05.ld S^T / dP^T of the half, P^T
//   warps 12-15 compute, half 1 (queries 64..127)           back to TMEM as bf16 pairs, dS^T to shared memory (per
//               block two [128 keys x 64 queries] slabs, 128B swizzle, double buffered).  Two warpgroups = two warps
//               per SM sub-partition: one warp alone cannot keep a sub-partition's issue slots busy.
//   warps 4-7   drain: dQ_i TMEM -> smem -> TMA reduce-add; after an item's last
//               block dK_j / dV_j TMEM -> bf16 -> smem -> TMA store
//   warp 8      TMA producer: (K_j, V_j) double buffered per item; (Q_i, dO_i) double buffered per block
//   warp 9      MMA issuer B (one thread): dV_j, dK_j, dQ_i
//   warp 10     statistics: the five per-query vectors of every block (and ktag once per item), two blocks ahead
//   warp 11     MMA issuer A (one thread): S^T / dP^T halves, refilled the moment the compute warps release them
// The mask semantics are those of attention_bwd.cu (key validity, video_src groups, uniform rows).
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kB = 128;                    // keys per item, queries per block
constexpr int kD = 64;
constexpr int kThreadsTc = 512;
constexpr int kProducerWarp = 8, kMmaWarp = 9, kMmaWarpS = 11;
constexpr int kTileBytes = kB * kD * 2;    // 16 KB
constexpr uint32_t kColS = 0, kColDP = 128, kColP = 256, kColDV = 320, kColDK = 384, kColDQ = 448;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kBig = 1.0e30f;
constexpr float kKeyMasked = 1.0e38f;
constexpr int kStatVecs = 6;               // a, cb, nb, qtag, D per query; ktag per key (written with an item's first block)

constexpr int kOffKV = 0;                             // [2 items]{K_j, V_j}
constexpr int kOffQ = 4 * kTileBytes;                 // [2 blocks]{Q_i, dO_i}
constexpr int kOffDS = kOffQ + 4 * kTileBytes;        // [2 blocks][2 slabs][128 keys][64 queries] bf16
constexpr int kOffStat = kOffDS + 4 * kTileBytes;     // [2 blocks][6][128] float
constexpr int kOffDQ = kOffStat + 2 * kStatVecs * kB * 4;   // [4 drain warps][32 rows x 32 floats] fp32, 128B swizzle
constexpr int kOffBar = kOffDQ + 4 * 4096;
constexpr int kNumBars = 26;
constexpr int kSmemBwdTc = kOffBar + kNumBars * 8 + 16 + 1024;
static_assert(kOffDQ % 1024 == 0 && kSmemBwdTc <= 227 * 1024, "shared memory layout");

struct BwdParams {
  const float* lse;
  const float* Dsum;
  const uint8_t* valid;
  const int32_t* group;
  int B, T, H, nh, nblk, n_items;
};

struct Item {
  int b, h, jb;
};
__device__ __forceinline__ Item decode_item(int it, int nblk, int nh) {
  Item x;
  x.jb = it % nblk;
  const int t = it / nblk;
  x.h = t % nh;
  x.b = t / nh;
  return x;
}

__global__ void __launch_bounds__(kThreadsTc, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dkv,
                        const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array: an integer round trip makes the pointer generic and
  // every access through it LD.E / ST.E instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sStat = reinterpret_cast<float*>(smem + kOffStat);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* kv_full = bars + 0;     // [2] per item buffer: K_j, V_j have landed
  uint64_t* kv_free = bars + 2;     // [2] every MMA of the item that read that buffer has completed
  uint64_t* q_full = bars + 4;      // [2]
  uint64_t* q_empty = bars + 6;     // [2]
  uint64_t* stat_full = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;     // [2] per 64-query half: S^T and dP^T of the half are in TMEM
  uint64_t* sdp_free = bars + 12;   // [2] per half: the compute warps hold S^T / dP^T of the half in registers
  uint64_t* ds_free = bars + 14;    // [2] per dS buffer: dK / dQ MMAs of the block have read it
  uint64_t* p_full = bars + 16;
  uint64_t* ds_full = bars + 17;
  uint64_t* dq_full = bars + 18;
  uint64_t* dq_free = bars + 19;
  uint64_t* pv_done = bars + 20;    // dV MMAs of a block have read P^T
  uint64_t* acc_full = bars + 21;   // every MMA of an item has completed: dK_j, dV_j are final, K_j / V_j smem is dead
  uint64_t* acc_free = bars + 22;   // the drain warps have pulled dK_j, dV_j out of TMEM
  uint64_t* stat_free = bars + 23;  // [2] the compute warps have finished reading the block's broadcast vectors
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = p.nblk;
  const int my_items = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int total = my_items * nqb;          // blocks this CTA processes; g = n * nqb + i counts them across items

  if (warp_idx == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq);
    tma_prefetch_desc(&tmap_dkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_free[i], 1);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&stat_full[i], 1);
      mbar_init(&stat_free[i], 8);
      mbar_init(&s_full[i], 1);
      mbar_init(&sdp_free[i], 4);
      mbar_init(&ds_free[i], 1);
    }
    mbar_init(p_full, 8);
    mbar_init(ds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_init(pv_done, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 4);
    fence_mbar_init();
  }
  if (warp_idx == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx >= 8 && warp_idx < 12) {
    setmaxnreg_dec<64>();
    if (warp_idx == kProducerWarp && lane == 0) {
      // ===================== TMA producer =====================
      int g = 0;
      for (int n = 0; n < my_items; ++n) {
        const Item it = decode_item(blockIdx.x + n * gridDim.x, p.nblk, p.nh);
        const int seq_base = it.b * p.T;
        uint8_t* kv = smem + kOffKV + (n & 1) * 2 * kTileBytes;
        mbar_wait(&kv_free[n & 1], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[n & 1], 2 * kTileBytes);
        tma_load_2d(&tmap_qkv, &kv_full[n & 1], kv, p.H + it.h * kD, seq_base + it.jb * kB, kEvictLast);
        tma_load_2d(&tmap_qkv, &kv_full[n & 1], kv + kTileBytes, 2 * p.H + it.h * kD, seq_base + it.jb * kB, kEvictLast);
        for (int i = 0; i < nqb; ++i, ++g) {
          const int buf = g & 1;
          const int qi = (i + it.jb) % nqb;     // rotated visiting order (see the drain warps)
          mbar_wait(&q_empty[buf], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[buf], 2 * kTileBytes);
          uint8_t* dst = smem + kOffQ + buf * 2 * kTileBytes;
          tma_load_2d(&tmap_qkv, &q_full[buf], dst, it.h * kD, seq_base + qi * kB, kEvictNormal);
          tma_load_2d(&tmap_do, &q_full[buf], dst + kTileBytes, it.h * kD, seq_base + qi * kB, kEvictNormal);
        }
      }
    } else if (warp_idx == kMmaWarp && lane == 0) {
      // ===================== MMA issuer B: dV_j, dK_j, dQ_i =====================
      constexpr uint32_t idesc_kv = umma_idesc_bf16_f32(kB, kD) | (1u << 16);                // B MN-major
      constexpr uint32_t idesc_dq = umma_idesc_bf16_f32(kB, kD) | (1u << 15) | (1u << 16);   // A and B MN-major
      // MN-major A of two 64-query slabs 16 KB apart: the descriptor's leading-byte-offset (16-byte units)
      constexpr uint64_t kLboSlab = static_cast<uint64_t>((2 * 8192 >> 4) - 1) << 16;
      const uint32_t tP = tmem_base + kColP;
      const uint32_t tDV = tmem_base + kColDV, tDK = tmem_base + kColDK, tDQ = tmem_base + kColDQ;
      int g = 0;
      for (int n = 0; n < my_items; ++n) {
        const int nb = n & 1;
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(smem + kOffKV + nb * 2 * kTileBytes));
        for (int i = 0; i < nqb; ++i, ++g) {
          const int buf = g & 1;
          const uint64_t qdesc = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes));
          const uint64_t dodesc = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes + kTileBytes));
          const uint32_t ds_addr = smem_u32(smem + kOffDS + buf * 2 * kTileBytes);
          // dV_j += P^T dO_i : K = 128 queries in 8 steps; P advances 8 TMEM columns, dO 16 rows (2 KB) per step
          mbar_wait(p_full, g & 1);
          if (i == 0 && n > 0) mbar_wait(acc_free, (n - 1) & 1);     // the previous item's dK / dV have left TMEM
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kB / 16; ++kk)
            umma_bf16_ts(tDV, tP + 8 * kk, dodesc + 128 * kk, idesc_kv, (i | kk) != 0 ? 1u : 0u);
          umma_commit(pv_done);
          // dK_j += dS^T Q_i (A = the smem dS^T tile read K-major: 16 queries = 32 B along the 128 B row, second
          // 64-query slab 16 KB further).  Q_i and dO_i are dead after it: the producer may refill the buffer.
          mbar_wait(ds_full, g & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kB / 16; ++kk) {
            const uint64_t ad = umma_desc_k_sw128(ds_addr + (kk >> 2) * kTileBytes) + 2 * (kk & 3);
            umma_bf16_ss(tDK, ad, qdesc + 128 * kk, idesc_kv, (i | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&q_empty[buf]);
          // dQ_i = dS K_j (the same dS^T tile read MN-major)
          if (g > 0) mbar_wait(dq_free, (g - 1) & 1);       // the drain warps have read the previous block's dQ
          tc_fence_after();
          const uint64_t dsdesc = umma_desc_k_sw128(ds_addr) + kLboSlab;
#pragma unroll
          for (int kk = 0; kk < kB / 16; ++kk)
            umma_bf16_ss(tDQ, dsdesc + 128 * kk, kdesc + 128 * kk, idesc_dq, kk != 0 ? 1u : 0u);
          umma_commit(dq_full);
          umma_commit(&ds_free[buf]);
          if (i + 1 == nqb) {
            umma_commit(acc_full);
            umma_commit(&kv_free[nb]);                    // K_j / V_j are dead: the producer may load item n + 2 here
          }
        }
      }
    } else if (warp_idx == kMmaWarpS && lane == 0) {
      // ===================== MMA issuer A: S^T, dP^T, one 64-query half at a time =====================
      // Its own thread: a half is refilled the moment the compute warps release it, whatever issuer B is waiting for
      // (one thread issuing everything in program order left the compute warps waiting for S^T 40 % of the time).
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kB, 64);                              // K-major A and B, one half
      const uint32_t tS = tmem_base + kColS, tDP = tmem_base + kColDP;
      int g = 0;
      for (int n = 0; n < my_items; ++n) {
        const int nb = n & 1;
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(smem + kOffKV + nb * 2 * kTileBytes));
        const uint64_t vdesc = umma_desc_k_sw128(smem_u32(smem + kOffKV + nb * 2 * kTileBytes + kTileBytes));
        mbar_wait(&kv_full[nb], (n >> 1) & 1);
        for (int i = 0; i < nqb; ++i, ++g) {
          const int buf = g & 1;
          mbar_wait(&q_full[buf], (g >> 1) & 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // B = rows [64 h, 64 h + 64) of the Q_i / dO_i tiles (8 KB further); the columns were block g-1's
            if (g > 0) mbar_wait(&sdp_free[h], (g - 1) & 1);
            tc_fence_after();
            const uint64_t qd = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes + h * 8192));
            const uint64_t dod = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes + kTileBytes + h * 8192));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tS + 64 * h, kdesc + 2 * k, qd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tDP + 64 * h, vdesc + 2 * k, dod + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&s_full[h]);
          }
        }
      }
    } else if (warp_idx == 10) {
      // ===================== statistics warp: the broadcast vectors of every block, four rows per thread =====================
      // The item of key block j visits the query blocks in the rotated order j, j+1, ... (mod nqb): at any moment the
      // CTAs working on one (sequence, head) then add into DIFFERENT rows of dQ -- the L2 atomic unit serialises per address.
      int g = 0;
      for (int n = 0; n < my_items; ++n) {
        const Item it = decode_item(blockIdx.x + n * gridDim.x, p.nblk, p.nh);
        const int seq_base = it.b * p.T;
        const size_t stat_base = (static_cast<size_t>(it.b) * p.nh + it.h) * p.T;
        for (int i = 0; i < nqb; ++i, ++g) {
          float vals[4][kStatVecs];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = r * 32 + lane;
            const int q = ((i + it.jb) % nqb) * kB + row;
            float a = 0.f, cb = -INFINITY, nb = 0.f, qtag = 0.f, dsum = 0.f, ktag = kKeyMasked;
            if (q < p.T) {
              const bool v = p.valid ? p.valid[seq_base + q] != 0 : true;
              a = v ? kLog2e : 0.f;
              cb = -p.lse[stat_base + q] * kLog2e;
              nb = v ? kBig : 0.f;
              qtag = (v && p.group) ? static_cast<float>(p.group[seq_base + q]) : 0.f;
              dsum = p.Dsum[stat_base + q];
            }
            const int key = it.jb * kB + row;
            if (i == 0 && key < p.T) {                   // the key side of the mask, once per item
              const bool v = p.valid ? p.valid[seq_base + key] != 0 : true;
              if (v) ktag = p.group ? static_cast<float>(p.group[seq_base + key]) : 0.f;
            }
            vals[r][0] = a; vals[r][1] = cb; vals[r][2] = nb; vals[r][3] = qtag; vals[r][4] = dsum; vals[r][5] = ktag;
          }
          mbar_wait(&stat_free[g & 1], ((g >> 1) & 1) ^ 1);
          float* st = sStat + (g & 1) * kStatVecs * kB;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = r * 32 + lane;
#pragma unroll
            for (int k = 0; k < 5; ++k) st[k * kB + row] = vals[r][k];
            if (i == 0) st[5 * kB + row] = vals[r][5];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&stat_full[g & 1]);
        }
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 8) {
    setmaxnreg_dec<96>();
    // ===================== drain warpgroup: dQ_i -> global, dK_j / dV_j -> global =====================
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;                 // query within the block = TMEM lane of dQ (key lane of dK / dV)
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* stage = smem + kOffDQ + quarter * 4096;
    int g = 0;
    for (int n = 0; n < my_items; ++n) {
      const Item it = decode_item(blockIdx.x + n * gridDim.x, p.nblk, p.nh);
      for (int i = 0; i < nqb; ++i, ++g) {
        mbar_wait(dq_full, g & 1);
        tc_fence_after();
        uint32_t v[kD];
        tmem_ld_32x64(lane_addr + kColDQ, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
        // dQ_i leaves as two TMA reduce-add boxes per warp (32 rows x 32 floats, 128B swizzle) through ONE 4 KB stage
        const int q0 = ((i + it.jb) % nqb) * kB + quarter * 32;
#pragma unroll
        for (int box = 0; box < 2; ++box) {
          if (lane == 0) tma_store_wait_read<0>();       // the previous box has been read
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                make_uint4(v[32 * box + 4 * c], v[32 * box + 4 * c + 1], v[32 * box + 4 * c + 2], v[32 * box + 4 * c + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (q0 < p.T) tma_reduce_add_3d(&tmap_dq, stage, it.h * kD + 32 * box, q0, it.b);   // rows past T are clipped
            tma_store_commit();
          }
        }
        if (i + 1 == nqb) {
          // ---- dK_j, dV_j: TMEM -> bf16 -> the warp's 4 KB stage -> TMA store into d(qkv) ----
          mbar_wait(acc_full, n & 1);
          tc_fence_after();
          const int r0 = it.jb * kB + quarter * 32;
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            tmem_ld_32x64(lane_addr + (which == 0 ? kColDK : kColDV), v);
            tmem_ld_wait();
            if (which == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_free);
            }
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint4 o;
              o.x = pack_bf16(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1]));
              o.y = pack_bf16(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3]));
              o.z = pack_bf16(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5]));
              o.w = pack_bf16(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7]));
              *reinterpret_cast<uint4*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) = o;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (r0 < p.T)                               // rows past T are clipped by the [B, T, 3H] tensor map
                tma_store_3d(&tmap_dkv, stage, (which == 0 ? p.H : 2 * p.H) + it.h * kD, r0, it.b);
              tma_store_commit();
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();                  // reductions and stores are complete before the CTA retires
  } else {
    setmaxnreg_inc<176>();
    // ===================== compute warpgroups: thread = key, one 64-query half each =====================
    const int half = warp_idx >= 12 ? 1 : 0;
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool plain = p.valid == nullptr && p.group == nullptr;
    int g = 0;
    for (int n = 0; n < my_items; ++n) {
      const Item it = decode_item(blockIdx.x + n * gridDim.x, p.nblk, p.nh);
      const bool key_exists = it.jb * kB + row < p.T;
      float ktag = kKeyMasked;
      for (int i = 0; i < nqb; ++i, ++g) {
      const float* st = sStat + (g & 1) * kStatVecs * kB;
      uint8_t* sDS = smem + kOffDS + (g & 1) * 2 * kTileBytes;
      mbar_wait(&stat_full[g & 1], (g >> 1) & 1);
      if (i == 0) ktag = st[5 * kB + row];
      {
        float pr[64];
        uint32_t dpv[64];
        mbar_wait(&s_full[half], g & 1);
        tc_fence_after();
        {
          uint32_t v[64];
          tmem_ld_32x64(lane_addr + kColS + 64 * half, v);
          tmem_ld_32x64(lane_addr + kColDP + 64 * half, dpv);
          tmem_ld_wait();
          // both are in registers: the issuer may refill these columns with block g+1
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sdp_free[half]);
          if (plain) {
            // no validity / group inputs (the vision encoder): every query is valid, a = log2e, nothing to mask but
            // the keys past T -- one FFMA and one broadcast vector per score instead of three and four
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
              const float4 cb = *reinterpret_cast<const float4*>(st + 1 * kB + 64 * half + 4 * c4);
              const float cc[4] = {cb.x, cb.y, cb.z, cb.w};
              // rows of keys past T are not masked here: their P^T only reaches dV_j rows that the TMA store clips, and
              // their dS^T is zeroed below (it would reach dQ)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                pr[4 * c4 + e] = ex2_approx(fmaf(__uint_as_float(v[4 * c4 + e]), kLog2e, cc[e]));
            }
          } else {
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
              const int c = 64 * half + 4 * c4;
              const float4 a = *reinterpret_cast<const float4*>(st + 0 * kB + c);
              const float4 cb = *reinterpret_cast<const float4*>(st + 1 * kB + c);
              const float4 nb = *reinterpret_cast<const float4*>(st + 2 * kB + c);
              const float4 qt = *reinterpret_cast<const float4*>(st + 3 * kB + c);
              const float aa[4] = {a.x, a.y, a.z, a.w}, cc[4] = {cb.x, cb.y, cb.z, cb.w};
              const float nn[4] = {nb.x, nb.y, nb.z, nb.w}, qq[4] = {qt.x, qt.y, qt.z, qt.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float t = fmaf(__uint_as_float(v[4 * c4 + e]), aa[e], cc[e]);
                t = fmaf(fabsf(ktag - qq[e]), -nn[e], t);
                pr[4 * c4 + e] = key_exists ? ex2_approx(t) : 0.f;
              }
            }
          }
        }
        // P^T of block g-1 has been read by its dV MMAs; the dS buffer by the dK / dQ MMAs of block g-2
        if (g > 0) mbar_wait(pv_done, (g - 1) & 1);
        if (g > 1) mbar_wait(&ds_free[g & 1], ((g - 2) >> 1) & 1);
        tc_fence_after();
        uint32_t pk[32];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {          // 8 queries at a time: dS^T goes straight to shared memory
          uint32_t dk[4];
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int c4 = 2 * ch + gg;
            const float4 d4 = *reinterpret_cast<const float4*>(st + 4 * kB + 64 * half + 4 * c4);
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
            float ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float pe = pr[4 * c4 + e];
              // an exactly-zero probability must give an exactly-zero gradient even where dP is garbage (rows past T)
              ds[e] = (pe == 0.f || !key_exists) ? 0.f : pe * (__uint_as_float(dpv[4 * c4 + e]) - dd[e]);
            }
            pk[2 * c4] = pack_bf16(pr[4 * c4], pr[4 * c4 + 1]);
            pk[2 * c4 + 1] = pack_bf16(pr[4 * c4 + 2], pr[4 * c4 + 3]);
            dk[2 * gg] = pack_bf16(ds[0], ds[1]);
            dk[2 * gg + 1] = pack_bf16(ds[2], ds[3]);
          }
          // slab `half`, row = key, chunk of 8 queries, 128B swizzle
          *reinterpret_cast<uint4*>(sDS + half * kTileBytes + row * 128 + ((ch ^ (row & 7)) << 4)) =
              make_uint4(dk[0], dk[1], dk[2], dk[3]);
        }
        // bf16 pairs: 64 queries = 32 TMEM columns
        tmem_st_32x32(lane_addr + kColP + 32 * half, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(ds_full);
        mbar_arrive(&stat_free[g & 1]);
      }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Called by mr_attention_bwd (attention_bwd.cu) after the prep kernel (D = rowsum(dO . O)) and the dq_acc memset.
int32_t launch_attention_bwd_tc(cudaStream_t stream, const bf16* qkv, const bf16* dout, const float* lse, const float* Dsum,
                                const uint8_t* valid, const int32_t* group, bf16* dqkv, float* dq_acc, int B, int T, int H) {
  CUtensorMap tm_qkv, tm_do, tm_dq, tm_dkv;
  if (!make_tmap_2d(&tm_qkv, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, kB, kD)) return MR_ERR_DRIVER;
  if (!make_tmap_2d(&tm_do, dout, static_cast<int64_t>(B) * T, H, H, kB, kD)) return MR_ERR_DRIVER;
  if (!make_tmap_3d_f32(&tm_dq, dq_acc, H, T, B, H, static_cast<int64_t>(T) * H, 32, 32)) return MR_ERR_DRIVER;
  if (!make_tmap_3d(&tm_dkv, dqkv, 3 * H, T, B, 3 * H, static_cast<int64_t>(T) * 3 * H, 32, kD)) return MR_ERR_DRIVER;
  BwdParams p{};
  p.lse = lse;
  p.Dsum = Dsum;
  p.valid = valid;
  p.group = group;
  p.B = B;
  p.T = T;
  p.H = H;
  p.nh = H / kD;
  p.nblk = ceil_div(T, kB);
  const int64_t items = static_cast<int64_t>(B) * p.nh * p.nblk;
  if (items >= (int64_t(1) << 31)) return MR_ERR_BAD_SHAPE;
  p.n_items = static_cast<int>(items);
  const int sms = device_sm_count();
  if (sms <= 0) return MR_ERR_DRIVER;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwdTc);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
  attention_bwd_tc_kernel<<<grid, kThreadsTc, kSmemBwdTc, stream>>>(tm_qkv, tm_do, tm_dq, tm_dkv, p);
  return after_launch();
}

}  // namespace mr
