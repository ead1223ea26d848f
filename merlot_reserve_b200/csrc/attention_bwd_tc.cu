// Tensor-core (tcgen05 / TMEM) attention backward for head dim 64, used when T > 64 (mr_attention_bwd).
//
// One CTA per (sequence, head, block j of 128 keys); it loops over the blocks i of 128 queries.  Everything is kept
// TRANSPOSED -- keys on the TMEM lanes, queries on the columns -- so that the probabilities and their gradients
// are the TMEM (A) operands of the two products that accumulate over the query blocks:
//
//   S^T  = K_j Q_i^T            SS  A = K_j  [128 keys x 64]  K-major      B = Q_i   K-major     -> TMEM [0,128)
//   dP^T = V_j dO_i^T           SS  A = V_j                                B = dO_i              -> TMEM [128,256)
//   P^T  = exp2(S^T log2e - lse_i) under the mask;   dS^T = P^T (dP^T - D_i)       (compute warps, thread = key)
//   dV_j += P^T  dO_i           TS  A = P^T  bf16 in TMEM [256,320)        B = dO_i  MN-major    -> TMEM [320,384)
//   dK_j += dS^T Q_i            SS  A = dS^T staged in smem, K-major       B = Q_i   MN-major    -> TMEM [384,448)
//   dQ_i  = dS   K_j            SS  A = the same smem tile as MN-major A   B = K_j   MN-major    -> TMEM [448,512)
//
// Pipelining (r02): S^T and dP^T are produced in two 64-query HALVES (N = 64 MMAs) with their own barriers.  As soon
// as the compute warps have pulled half h of block i into registers they release it, and the issuer immediately
// computes half h of block i+1 into the same columns -- under the compute of the other half -- so the compute warps
// never wait for a cold S^T / dP^T and the dV / dK / dQ MMAs of block i run under the compute of block i+1.  dS^T
// therefore no longer lives in TMEM (it used to overwrite dP^T, which is being refilled now): both of its consumers
// read the double-buffered shared-memory copy.  dQ_i leaves through fp32 vector reductions
// (red.global.add.v4.f32) into the workspace that mr_attention_bwd converts to bf16 afterwards, because the key
// blocks of a sequence live in different CTAs.  lse_i and D_i are per-COLUMN quantities in this orientation: the
// drain warpgroup turns them (and the query side of the mask) into five broadcast vectors per query block:
//     p = exp2( s a[c] + cb[c] - 1e30 nb[c] |ktag - qtag[c]| ),   dS = p (dP - D[c])
//   a  = log2e (valid query) or 0 (invalid query: uniform attention, S := 0);  cb = -lse log2e, or -inf past T;
//   nb = 1 for valid queries (mask applies) / 0 for invalid ones;  ktag = the key's group id, 1e38 if the key is
//   padding (never equals a query tag; 1e38 x 1e30 overflows to -inf, and x 0 stays 0 for the uniform rows).
//
//   warps 0-3   compute, half 0 (queries 0..63 of every block): thread = key.  tcgen05.ld S^T / dP^T of the half, P^T
//   warps 12-15 compute, half 1 (queries 64..127)           back to TMEM as bf16 pairs, dS^T to shared memory (per
//               block two [128 keys x 64 queries] slabs, 128B swizzle, double buffered).  Two warpgroups = two warps
//               per SM sub-partition: one warp alone cannot keep a sub-partition's issue slots busy.
//   warps 4-7   drain: per-query vectors of block i+1, then dQ_i TMEM -> red.global.add.v4.f32
//   warp 8      TMA producer: K_j, V_j once; (Q_i, dO_i) double buffered
//   warp 9      MMA issuer (one thread); warps 10-11 idle
// The mask semantics are those of attention_bwd.cu (key validity, video_src groups, uniform rows).
#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kB = 128;                    // keys per CTA, queries per step
constexpr int kD = 64;
constexpr int kThreadsTc = 512;
constexpr int kProducerWarp = 8, kMmaWarp = 9;
constexpr int kTileBytes = kB * kD * 2;    // 16 KB
constexpr uint32_t kColS = 0, kColDP = 128, kColP = 256, kColDV = 320, kColDK = 384, kColDQ = 448;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kBig = 1.0e30f;
constexpr float kKeyMasked = 1.0e38f;

constexpr int kOffK = 0;
constexpr int kOffV = kTileBytes;
constexpr int kOffQ = 2 * kTileBytes;                 // [2 buffers]{Q_i, dO_i}
constexpr int kOffDS = kOffQ + 4 * kTileBytes;        // [2 buffers][2 slabs][128 keys][64 queries] bf16
constexpr int kOffStat = kOffDS + 4 * kTileBytes;     // [2 buffers][5][128] float: a, cb, nb, qtag, D
constexpr int kOffDQ = kOffStat + 2 * 5 * kB * 4;      // [4 drain warps][2 boxes][32 rows x 32 floats] fp32, 128B swizzle
constexpr int kOffBar = kOffDQ + 4 * 2 * 4096;
constexpr int kNumBars = 20;
constexpr int kSmemBwdTc = kOffBar + kNumBars * 8 + 16 + 1024;

struct BwdParams {
  const float* lse;
  const float* Dsum;
  const uint8_t* valid;
  const int32_t* group;
  bf16* dqkv;
  float* dq_acc;
  int B, T, H, nh, nblk;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreadsTc, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_dq, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sStat = reinterpret_cast<float*>(smem + kOffStat);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;      // [2]
  uint64_t* q_empty = bars + 3;     // [2]
  uint64_t* stat_full = bars + 5;   // [2]
  uint64_t* s_full = bars + 7;      // [2] per 64-query half: S^T and dP^T of the half are in TMEM
  uint64_t* p_full = bars + 9;
  uint64_t* ds_full = bars + 10;
  uint64_t* dq_full = bars + 11;
  uint64_t* dq_free = bars + 12;
  uint64_t* kv_done = bars + 13;
  uint64_t* sdp_free = bars + 14;   // [2] per half: the compute warps hold S^T / dP^T of the half in registers
  uint64_t* pv_done = bars + 16;    // dV MMAs of a block have read P^T
  uint64_t* ds_free = bars + 17;    // [2] per dS buffer: dK / dQ MMAs of the block have read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = blockIdx.x % p.nblk;
  const int h = (blockIdx.x / p.nblk) % p.nh;
  const int b = blockIdx.x / (p.nblk * p.nh);
  const int seq_base = b * p.T;
  const int k0 = jb * kB;
  const int nqb = p.nblk;

  if (warp_idx == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&stat_full[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&sdp_free[i], 4);
      mbar_init(&ds_free[i], 1);
    }
    mbar_init(p_full, 8);
    mbar_init(ds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_init(kv_done, 1);
    mbar_init(pv_done, 1);
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
      mbar_arrive_expect_tx(kv_full, 2 * kTileBytes);
      tma_load_2d(&tmap_qkv, kv_full, smem + kOffK, p.H + h * kD, seq_base + k0, kEvictLast);
      tma_load_2d(&tmap_qkv, kv_full, smem + kOffV, 2 * p.H + h * kD, seq_base + k0, kEvictLast);
      for (int i = 0; i < nqb; ++i) {
        const int buf = i & 1;
        const int qi = (i + jb) % nqb;     // rotated visiting order (see the drain warps)
        mbar_wait(&q_empty[buf], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[buf], 2 * kTileBytes);
        uint8_t* dst = smem + kOffQ + buf * 2 * kTileBytes;
        tma_load_2d(&tmap_qkv, &q_full[buf], dst, h * kD, seq_base + qi * kB, kEvictNormal);
        tma_load_2d(&tmap_do, &q_full[buf], dst + kTileBytes, h * kD, seq_base + qi * kB, kEvictNormal);
      }
    } else if (warp_idx == kMmaWarp && lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kB, 64);                              // K-major A and B, one half
      constexpr uint32_t idesc_kv = umma_idesc_bf16_f32(kB, kD) | (1u << 16);                // B MN-major
      constexpr uint32_t idesc_dq = umma_idesc_bf16_f32(kB, kD) | (1u << 15) | (1u << 16);   // A and B MN-major
      // MN-major A of two 64-query slabs 16 KB apart: the descriptor's leading-byte-offset (16-byte units)
      constexpr uint64_t kLboSlab = static_cast<uint64_t>((2 * 8192 >> 4) - 1) << 16;
      const uint64_t kdesc = umma_desc_k_sw128(smem_u32(smem + kOffK));
      const uint64_t vdesc = umma_desc_k_sw128(smem_u32(smem + kOffV));
      const uint32_t tS = tmem_base + kColS, tDP = tmem_base + kColDP, tP = tmem_base + kColP;
      const uint32_t tDV = tmem_base + kColDV, tDK = tmem_base + kColDK, tDQ = tmem_base + kColDQ;
      // S^T / dP^T of one 64-query half of block i: B = rows [64 h, 64 h + 64) of the Q_i / dO_i tiles (8 KB further)
      auto issue_half = [&](int i, int h) {
        const int buf = i & 1;
        const uint64_t qd = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes + h * 8192));
        const uint64_t dod = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes + kTileBytes + h * 8192));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tS + 64 * h, kdesc + 2 * k, qd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tDP + 64 * h, vdesc + 2 * k, dod + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[h]);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_half(0, 0);
      issue_half(0, 1);
      for (int i = 0; i < nqb; ++i) {
        const int buf = i & 1;
        const uint64_t qdesc = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes));
        const uint64_t dodesc = umma_desc_k_sw128(smem_u32(smem + kOffQ + buf * 2 * kTileBytes + kTileBytes));
        const uint32_t ds_addr = smem_u32(smem + kOffDS + buf * 2 * kTileBytes);
        if (i + 1 < nqb) {
          // halves of block i+1 go into the columns the compute warps have just emptied (arrival order: half 0, half 1)
          mbar_wait(&q_full[buf ^ 1], ((i + 1) >> 1) & 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&sdp_free[h], i & 1);
            tc_fence_after();
            issue_half(i + 1, h);
          }
        }
        // dV_j += P^T dO_i : K = 128 queries in 8 steps; P advances 8 TMEM columns, dO 16 rows (2 KB) per step
        mbar_wait(p_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kB / 16; ++kk)
          umma_bf16_ts(tDV, tP + 8 * kk, dodesc + 128 * kk, idesc_kv, (i | kk) != 0 ? 1u : 0u);
        umma_commit(pv_done);
        // dK_j += dS^T Q_i (A = the smem dS^T tile read K-major: 16 queries = 32 B along the 128 B row, second
        // 64-query slab 16 KB further); dQ_i = dS K_j (the same tile read MN-major)
        mbar_wait(ds_full, i & 1);
        if (i > 0) mbar_wait(dq_free, (i - 1) & 1);       // the drain warps have read dQ_{i-1}
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kB / 16; ++kk) {
          const uint64_t ad = umma_desc_k_sw128(ds_addr + (kk >> 2) * kTileBytes) + 2 * (kk & 3);
          umma_bf16_ss(tDK, ad, qdesc + 128 * kk, idesc_kv, (i | kk) != 0 ? 1u : 0u);
        }
        const uint64_t dsdesc = umma_desc_k_sw128(ds_addr) + kLboSlab;
#pragma unroll
        for (int kk = 0; kk < kB / 16; ++kk)
          umma_bf16_ss(tDQ, dsdesc + 128 * kk, kdesc + 128 * kk, idesc_dq, kk != 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&q_empty[buf]);
        umma_commit(&ds_free[buf]);
      }
      umma_commit(kv_done);
    }
  } else if (warp_idx >= 4 && warp_idx < 8) {
    setmaxnreg_dec<96>();
    // ===================== drain warpgroup: per-query vectors, dQ_i -> global =====================
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;                 // query within the block = TMEM lane of dQ
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const size_t stat_base = (static_cast<size_t>(b) * p.nh + h) * p.T;
    // The CTA of key block j visits the query blocks in the rotated order j, j+1, ... (mod nqb): at any moment the
    // CTAs of one (sequence, head) then add into DIFFERENT rows of dQ -- the L2 atomic unit serialises per address.
    auto load_stats = [&](int i) {
      const int q = ((i + jb) % nqb) * kB + row;
      float a = 0.f, cb = -INFINITY, nb = 0.f, qtag = 0.f, dsum = 0.f;
      if (q < p.T) {
        const bool v = p.valid ? p.valid[seq_base + q] != 0 : true;
        a = v ? kLog2e : 0.f;
        cb = -p.lse[stat_base + q] * kLog2e;
        nb = v ? kBig : 0.f;
        qtag = (v && p.group) ? static_cast<float>(p.group[seq_base + q]) : 0.f;
        dsum = p.Dsum[stat_base + q];
      }
      float* st = sStat + (i & 1) * 5 * kB;
      st[0 * kB + row] = a;
      st[1 * kB + row] = cb;
      st[2 * kB + row] = nb;
      st[3 * kB + row] = qtag;
      st[4 * kB + row] = dsum;
      __syncwarp();
      if (lane == 0) mbar_arrive(&stat_full[i & 1]);
    };
    load_stats(0);
    for (int i = 0; i < nqb; ++i) {
      if (i + 1 < nqb) load_stats(i + 1);                // block i-1's readers finished before dq_full(i-1)
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t v[kD];
      tmem_ld_32x64(lane_addr + kColDQ, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      // dQ_i leaves as two TMA reduce-add boxes per warp (32 rows x 32 floats, 128B swizzle): per-lane REDG of the
      // 8 192 floats of a block kept the SM's load/store unit busy for longer than the block's whole compute phase
      const int q0 = ((i + jb) % nqb) * kB + quarter * 32;
      uint8_t* stage = smem + kOffDQ + quarter * 2 * 4096;
      if (lane == 0) tma_store_wait_read<0>();           // the previous block's boxes have been read
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 16; ++c)
        *reinterpret_cast<uint4*>(stage + (c >> 3) * 4096 + lane * 128 + (((c & 7) ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (q0 < p.T) {                                  // rows past T are clipped by the [B, T, H] tensor map
          tma_reduce_add_3d(&tmap_dq, stage, h * kD, q0, b);
          tma_reduce_add_3d(&tmap_dq, stage + 4096, h * kD + 32, q0, b);
        }
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();                  // the reductions are complete before the CTA retires
  } else {
    setmaxnreg_inc<176>();
    // ===================== compute warpgroups: thread = key, one 64-query half each =====================
    const int half = warp_idx >= 12 ? 1 : 0;
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int key = k0 + row;
    const bool key_exists = key < p.T;
    float ktag = kKeyMasked;
    if (key_exists) {
      const bool v = p.valid ? p.valid[seq_base + key] != 0 : true;
      if (v) ktag = p.group ? static_cast<float>(p.group[seq_base + key]) : 0.f;
    }
    for (int i = 0; i < nqb; ++i) {
      const float* st = sStat + (i & 1) * 5 * kB;
      uint8_t* sDS = smem + kOffDS + (i & 1) * 2 * kTileBytes;
      mbar_wait(&stat_full[i & 1], (i >> 1) & 1);
      {
        float pr[64];
        uint32_t dpv[64];
        mbar_wait(&s_full[half], i & 1);
        tc_fence_after();
        {
          uint32_t v[64];
          tmem_ld_32x64(lane_addr + kColS + 64 * half, v);
          tmem_ld_32x64(lane_addr + kColDP + 64 * half, dpv);
          tmem_ld_wait();
          // both are in registers: the issuer may refill these columns with block i+1
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sdp_free[half]);
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
        // P^T of block i-1 has been read by its dV MMAs; the dS buffer by the dK / dQ MMAs of block i-2
        if (i > 0) mbar_wait(pv_done, (i - 1) & 1);
        if (i > 1) mbar_wait(&ds_free[i & 1], ((i - 2) >> 1) & 1);
        tc_fence_after();
        uint32_t pk[32];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {          // 8 queries at a time: dS^T goes straight to shared memory
          uint32_t dk[4];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int c4 = 2 * ch + g;
            const float4 d4 = *reinterpret_cast<const float4*>(st + 4 * kB + 64 * half + 4 * c4);
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
            float ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float pe = pr[4 * c4 + e];
              // an exactly-zero probability must give an exactly-zero gradient even where dP is garbage (rows past T)
              ds[e] = pe == 0.f ? 0.f : pe * (__uint_as_float(dpv[4 * c4 + e]) - dd[e]);
            }
            pk[2 * c4] = pack_bf16(pr[4 * c4], pr[4 * c4 + 1]);
            pk[2 * c4 + 1] = pack_bf16(pr[4 * c4 + 2], pr[4 * c4 + 3]);
            dk[2 * g] = pack_bf16(ds[0], ds[1]);
            dk[2 * g + 1] = pack_bf16(ds[2], ds[3]);
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
      if (lane == 0) mbar_arrive(ds_full);
    }
    // ---- dK_j, dV_j: TMEM -> bf16 -> the k / v columns of d(qkv) ----
    mbar_wait(kv_done, 0);
    tc_fence_after();
    {
      const int which = half;      // warpgroup 0 writes dK_j, warpgroup 1 writes dV_j
      uint32_t v[kD];
      tmem_ld_32x64(lane_addr + (which == 0 ? kColDK : kColDV), v);
      tmem_ld_wait();
      if (key_exists) {
        bf16* dst = p.dqkv + (static_cast<size_t>(seq_base) + key) * 3 * p.H + (which == 0 ? p.H : 2 * p.H) + h * kD;
#pragma unroll
        for (int c = 0; c < kD; c += 8) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(v[c + 0]), __uint_as_float(v[c + 1]));
          o.y = pack_bf16(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
          o.z = pack_bf16(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5]));
          o.w = pack_bf16(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7]));
          *reinterpret_cast<uint4*>(dst + c) = o;
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
  CUtensorMap tm_qkv, tm_do, tm_dq;
  if (!make_tmap_2d(&tm_qkv, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, kB, kD)) return MR_ERR_DRIVER;
  if (!make_tmap_2d(&tm_do, dout, static_cast<int64_t>(B) * T, H, H, kB, kD)) return MR_ERR_DRIVER;
  if (!make_tmap_3d_f32(&tm_dq, dq_acc, H, T, B, H, static_cast<int64_t>(T) * H, 32, 32)) return MR_ERR_DRIVER;
  BwdParams p{};
  p.lse = lse;
  p.Dsum = Dsum;
  p.valid = valid;
  p.group = group;
  p.dqkv = dqkv;
  p.dq_acc = dq_acc;
  p.B = B;
  p.T = T;
  p.H = H;
  p.nh = H / kD;
  p.nblk = ceil_div(T, kB);
  const int64_t blocks = static_cast<int64_t>(B) * p.nh * p.nblk;
  if (blocks >= (int64_t(1) << 31)) return MR_ERR_BAD_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwdTc);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  attention_bwd_tc_kernel<<<static_cast<unsigned>(blocks), kThreadsTc, kSmemBwdTc, stream>>>(tm_qkv, tm_do, tm_dq, p);
  return after_launch();
}

}  // namespace mr
