// Tensor-core (tcgen05) flash attention forward, head dim 64: THREE independent query-tile streams per SM.
//
// attention_tc.cu runs two 128-query tiles per SM (one softmax warp per tile and SM sub-partition).  At head dim 64
// the SFU bounds the kernel (128 x 128 exp2 per 2 x 2^21 MACs), and ncu shows the two softmax warps of a scheduler
// issuing on ~21 % of their cycles each: the SFU idles whenever both are between ex2 phases (TMEM round trip, row
// max, PV wait, P store), 42 % of the time on the Base vision shape.  Two tiles of 128-key blocks fill TMEM
// (2 x (128 S + 64 P + 64 O) = 512 columns); with 64-key blocks a tile needs 64 + 32 + 64 = 160 columns and THREE fit
// (480).  So: three softmax warpgroups = three warps per scheduler, each on its own (sequence, head, query tile)
// stream with its own Q tile, K/V ring, TMEM columns, MMA issuer and producer -- nothing is shared between the
// streams, so no stream ever waits for another.
//
//   warps 0-11   softmax, stream s = warp / 4: thread = query row; tcgen05.ld S (64 columns) -> (key-validity bias)
//                -> row max (lazy re-basing, threshold 2^8) -> FFMA + ex2 per score -> row sum -> bf16 P ->
//                tcgen05.st; after a tile's last key block (l, m) go to smem and the warp starts its next tile
//   warps 12-15  epilogue for all three streams: polls o_full[s]; tcgen05.ld O -> O / l -> bf16 -> swizzled smem ->
//                one TMA store per warp (3-D map, rows past T clipped); optional lse
//   warp 16      TMA producer (one thread) for all three streams, a polling state machine: per stream the Q tile
//                (refilled as soon as the tile's last S MMA retired) and a 3-stage ring of K | V blocks [64 keys x 64]
//   warps 17-19  MMA issuer of stream s: S = Q K^T (128 x 64 x 64, SS) and O += P V (TS, P from TMEM, V MN-major);
//                S of block g+1 is issued before PV of block g, also across tiles
// Registers: setmaxnreg moves registers inside the CTA's own pool = threads x launch registers = 640 x 96 = 61 440:
// softmax 120, epilogue 80, producer / issuers 40 (384 x 120 + 128 x 80 + 128 x 40 = 61 440).
//
// Masks: key validity only (the inference paths: embed_video, multiple-choice scoring); calls with video_src group
// ids (pretraining) stay on attention_tc.cu.  A query row whose mask is all-false attends uniformly over all T keys
// (the reference's -1e10 bias behaviour, modeling.py:343-358): score scale 0 and the "key in range" bias vector.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace mr {
namespace {

constexpr int kQ = 128, kKV = 64, kD = 64;
constexpr int kStreams = 3;
constexpr int kStagesKV = 3;
constexpr int kThreads3 = 640;
constexpr int kEpiWarp0 = 12, kProducerWarp = 16, kMmaWarp0 = 17;
constexpr int kQBytes = kQ * kD * 2;       // 16 KB
constexpr int kKvBytes = kKV * kD * 2;     // 8 KB
constexpr uint32_t kColStride = 160, kColS = 0, kColP = 64, kColO = 96;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

// shared memory map (offsets from the 1024-aligned base)
constexpr int kOffQ = 0;                                            // [3 streams]
constexpr int kOffKV = kStreams * kQBytes;                          // [3 streams][kStagesKV]{K, V}
constexpr int kOffO = kOffKV + kStreams * kStagesKV * 2 * kKvBytes; // [4 epilogue warps][32 rows x 128 B]
constexpr int kOffBias = kOffO + 4 * 4096;                          // [12 softmax warps]{bias_v[64], bias_e[64]} float
constexpr int kOffLM = kOffBias + 12 * 2 * kKV * 4;                 // [3 streams][2 tile parity][128 rows] float2 (l, m)
constexpr int kOffBar = kOffLM + kStreams * 2 * kQ * 8;
constexpr int kBarsPerStream = 9 + 2 * kStagesKV;
constexpr int kSmem3 = kOffBar + kStreams * kBarsPerStream * 8 + 16 + 1024;
static_assert(kOffO % 1024 == 0 && kSmem3 <= 227 * 1024, "shared memory layout");

struct Attn3Params {
  const uint8_t* valid;
  float* lse;   // optional [B, nh, T]
  int B, T, H, nh, nqb, nkb;
  int num_tiles;
};

struct Tile {
  int b, h, qt;
};
__device__ __forceinline__ Tile decode_tile(int u, const Attn3Params& p) {
  Tile t;
  t.qt = u % p.nqb;
  const int bh = u / p.nqb;
  t.h = bh % p.nh;
  t.b = bh / p.nh;
  return t;
}

// non-blocking probe of an mbarrier phase (the epilogue warps watch three streams at once)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

template <bool kMasked>
__global__ void __launch_bounds__(kThreads3, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_out, const Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  // per stream s: bars + s * kBarsPerStream
  auto B_ = [&](int s, int k) { return bars + s * kBarsPerStream + k; };
  enum { Q_FULL = 0, Q_EMPTY, S_FULL, S_FREE, P_FULL, PV_DONE, O_FULL, O_FREE, LM_FULL, KV_FULL0, KV_EMPTY0 = KV_FULL0 + kStagesKV };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStreams * kBarsPerStream);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tiles of stream s of this CTA: u = blockIdx.x * 3 + s, stride 3 * gridDim.x
  const int stride = kStreams * static_cast<int>(gridDim.x);

  if (warp_idx == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < kStreams; ++s) {
      mbar_init(B_(s, Q_FULL), 1);
      mbar_init(B_(s, Q_EMPTY), 1);
      mbar_init(B_(s, S_FULL), 1);
      mbar_init(B_(s, S_FREE), 4);
      mbar_init(B_(s, P_FULL), 4);
      mbar_init(B_(s, PV_DONE), 1);
      mbar_init(B_(s, O_FULL), 1);
      mbar_init(B_(s, O_FREE), 4);
      mbar_init(B_(s, LM_FULL), 4);
      for (int k = 0; k < kStagesKV; ++k) {
        mbar_init(B_(s, KV_FULL0 + k), 1);
        mbar_init(B_(s, KV_EMPTY0 + k), 1);
      }
    }
    fence_mbar_init();
  }
  if (warp_idx == kMmaWarp0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx >= kProducerWarp) {
    setmaxnreg_dec<40>();
    if (warp_idx == kProducerWarp) {
      // ===================== TMA producer: one thread, three streams, never blocks on one of them =====================
      if (lane == 0) {
        int u[kStreams], j[kStreams];               // next tile / next key block (j = -1: the tile's Q comes first)
        uint32_t n_t[kStreams], n_kv[kStreams];
        int live = 0;
#pragma unroll
        for (int s = 0; s < kStreams; ++s) {
          u[s] = blockIdx.x * kStreams + s;
          j[s] = -1;
          n_t[s] = n_kv[s] = 0;
          live += u[s] < p.num_tiles ? 1 : 0;
        }
        while (live > 0) {
          bool progress = false;
#pragma unroll
          for (int s = 0; s < kStreams; ++s) {
            if (u[s] >= p.num_tiles) continue;
            const Tile t = decode_tile(u[s], p);
            const int seq_base = t.b * p.T;
            if (j[s] < 0) {
              if (!mbar_test(B_(s, Q_EMPTY), (n_t[s] & 1) ^ 1)) continue;
              mbar_arrive_expect_tx(B_(s, Q_FULL), kQBytes);
              tma_load_2d(&tmap_q, B_(s, Q_FULL), smem + kOffQ + s * kQBytes, t.h * kD, seq_base + t.qt * kQ, kEvictFirst);
              j[s] = 0;
              progress = true;
            }
            const int st = n_kv[s] % kStagesKV;
            if (!mbar_test(B_(s, KV_EMPTY0 + st), ((n_kv[s] / kStagesKV) & 1) ^ 1)) continue;
            mbar_arrive_expect_tx(B_(s, KV_FULL0 + st), 2 * kKvBytes);
            uint8_t* dst = smem + kOffKV + (s * kStagesKV + st) * 2 * kKvBytes;
            tma_load_2d(&tmap_kv, B_(s, KV_FULL0 + st), dst, p.H + t.h * kD, seq_base + j[s] * kKV, kEvictLast);
            tma_load_2d(&tmap_kv, B_(s, KV_FULL0 + st), dst + kKvBytes, 2 * p.H + t.h * kD, seq_base + j[s] * kKV, kEvictLast);
            ++n_kv[s];
            progress = true;
            if (++j[s] == p.nkb) {
              j[s] = -1;
              u[s] += stride;
              ++n_t[s];
              if (u[s] >= p.num_tiles) --live;
            }
          }
          if (!progress) __nanosleep(100);
        }
      }
    } else if (warp_idx >= kMmaWarp0 && warp_idx < kMmaWarp0 + kStreams) {
      // ===================== MMA issuer of stream s =====================
      if (lane == 0) {
        const int s = warp_idx - kMmaWarp0;
        const uint32_t sQ = smem_u32(smem + kOffQ + s * kQBytes);
        const uint32_t sKV = smem_u32(smem + kOffKV + s * kStagesKV * 2 * kKvBytes);
        constexpr uint32_t idesc_s = umma_idesc_bf16_f32_ex(kQ, kKV, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16_f32_ex(kQ, kD, 1);   // B (= V) is MN-major
        const uint32_t tS = tmem_base + s * kColStride + kColS, tP = tmem_base + s * kColStride + kColP;
        const uint32_t tO = tmem_base + s * kColStride + kColO;
        const uint64_t qdesc = umma_desc_k_sw128(sQ);
        uint32_t ns = 0, np = 0, nd = 0;    // S blocks issued, PV blocks issued, tiles finished
        struct Cursor {
          int u, j;
          uint32_t n_t, n_kv;
        };
        auto issue_s = [&](const Cursor& c) {
          if (c.j == 0) mbar_wait(B_(s, Q_FULL), c.n_t & 1);
          const uint32_t st = c.n_kv % kStagesKV;
          mbar_wait(B_(s, KV_FULL0 + st), (c.n_kv / kStagesKV) & 1);
          if (ns > 0) mbar_wait(B_(s, S_FREE), (ns - 1) & 1);     // the softmax warps hold the previous S in registers
          tc_fence_after();
          const uint64_t kdesc = umma_desc_k_sw128(sKV + st * 2 * kKvBytes);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(B_(s, S_FULL));
          if (c.j == p.nkb - 1) umma_commit(B_(s, Q_EMPTY));       // the tile's last read of Q
          ++ns;
        };
        Cursor cur{static_cast<int>(blockIdx.x) * kStreams + s, 0, 0u, 0u};
        if (cur.u < p.num_tiles) issue_s(cur);
        while (cur.u < p.num_tiles) {
          Cursor nxt = cur;
          ++nxt.n_kv;
          if (++nxt.j == p.nkb) {
            nxt.j = 0;
            nxt.u += stride;
            ++nxt.n_t;
          }
          if (nxt.u < p.num_tiles) issue_s(nxt);
          const uint32_t st = cur.n_kv % kStagesKV;
          const uint64_t vdesc = umma_desc_k_sw128(sKV + st * 2 * kKvBytes + kKvBytes);
          mbar_wait(B_(s, P_FULL), np & 1);
          if (cur.j == 0 && nd > 0) mbar_wait(B_(s, O_FREE), (nd - 1) & 1);   // the previous tile's O was read
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kKV / 16; ++kk)   // 16 keys = 16 rows x 128 B further into V; P advances 8 TMEM columns
            umma_bf16_ts(tO, tP + 8 * kk, vdesc + 128 * kk, idesc_o, (cur.j | kk) != 0 ? 1u : 0u);
          umma_commit(B_(s, PV_DONE));
          ++np;
          umma_commit(B_(s, KV_EMPTY0 + st));
          if (cur.j == p.nkb - 1) {
            umma_commit(B_(s, O_FULL));
            ++nd;
          }
          cur = nxt;
        }
      }
    }
  } else if (warp_idx >= kEpiWarp0) {
    setmaxnreg_dec<80>();
    // ===================== epilogue warps: every warp serves its 32-row quarter of all three streams =====================
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* sO = smem + kOffO + quarter * 4096;
    uint32_t done[kStreams] = {0, 0, 0};
    int next_u[kStreams];
    int left = 0;
#pragma unroll
    for (int s = 0; s < kStreams; ++s) {
      next_u[s] = blockIdx.x * kStreams + s;
      if (next_u[s] < p.num_tiles) left += (p.num_tiles - next_u[s] + stride - 1) / stride;
    }
    int s = 0, misses = 0;
    while (left > 0) {
      // round-robin probe (warp-uniform: lane 0 tests, the result is broadcast)
      bool ready = false;
      if (next_u[s] < p.num_tiles) {
        int r = 0;
        if (lane == 0) r = mbar_test(B_(s, O_FULL), done[s] & 1) ? 1 : 0;
        ready = __shfl_sync(0xffffffffu, r, 0) != 0;
      }
      if (!ready) {
        s = s + 1 == kStreams ? 0 : s + 1;
        if (++misses >= kStreams) {                   // a whole round found nothing: leave the issue slots to the softmax warps
          __nanosleep(200);
          misses = 0;
        }
        continue;
      }
      misses = 0;
      const Tile t = decode_tile(next_u[s], p);
      const int q0 = t.qt * kQ;
      const int q = q0 + row;
      mbar_wait(B_(s, LM_FULL), done[s] & 1);
      const float2 lm = *reinterpret_cast<const float2*>(smem + kOffLM + ((s * 2 + (done[s] & 1)) * kQ + row) * 8);
      tc_fence_after();
      const float inv = lm.x > 0.f ? rcp_approx(lm.x) : 0.f;
      uint4 o16[kD / 8];
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + s * kColStride + kColO + 32 * c, v);
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
      if (lane == 0) mbar_arrive(B_(s, O_FREE));     // the stream's next tile may overwrite O
      if (p.lse != nullptr && q < p.T)
        p.lse[(static_cast<size_t>(t.b) * p.nh + t.h) * p.T + q] =
            (lm.y + (lm.x > 0.f ? log2f(lm.x) : 0.f)) * 0.6931471805599453f;
      if (lane == 0) tma_store_wait_read<0>();       // the previous store has read the staging rows
      __syncwarp();
#pragma unroll
      for (int i = 0; i < kD / 8; ++i)
        *reinterpret_cast<uint4*>(sO + lane * 128 + ((i ^ (lane & 7)) << 4)) = o16[i];   // 128B swizzle
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (q0 + quarter * 32 < p.T) tma_store_3d(&tmap_out, sO, t.h * kD, q0 + quarter * 32, t.b);
        tma_store_commit();
      }
      ++done[s];
      next_u[s] += stride;
      --left;
      s = s + 1 == kStreams ? 0 : s + 1;
    }
    if (lane == 0) tma_store_wait<0>();              // smem must outlive the last store
  } else {
    setmaxnreg_inc<120>();
    // ===================== softmax warpgroups (thread = query row), one per stream =====================
    const int s = warp_idx >> 2;
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t colS = s * kColStride + kColS, colP = s * kColStride + kColP, colO = s * kColStride + kColO;
    float* myBiasV = reinterpret_cast<float*>(smem + kOffBias) + warp_idx * 2 * kKV;   // 0 / -inf: key valid and in range
    float* myBiasE = myBiasV + kKV;                                                    // 0 / -inf: key in range
    uint32_t n_blk = 0, n_t = 0;

    // the two keys (64 j + 2 lane + i) whose validity this lane fetches for its warp, one key block ahead
    auto load_raw = [&](int seq_base, int j) -> uint32_t {
      uint32_t r = 0;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kidx = min(j * kKV + lane * 2 + i, p.T - 1);
        r |= static_cast<uint32_t>(p.valid[seq_base + kidx] != 0 ? 1 : 0) << i;
      }
      return r;
    };

    for (int u = blockIdx.x * kStreams + s; u < p.num_tiles; u += stride, ++n_t) {
      const Tile t = decode_tile(u, p);
      const int seq_base = t.b * p.T;
      const int q = t.qt * kQ + row;
      bool q_uniform = false;
      uint32_t raw_next = 0;
      if (kMasked) {
        q_uniform = !(q < p.T && p.valid[seq_base + q] != 0);
        raw_next = load_raw(seq_base, 0);
      }
      const float cq = q_uniform ? 0.f : kLog2e;   // fast-path score scale: 0 -> uniform attention
      const float qs = q_uniform ? 0.f : 1.f;
      float m2 = 0.f;  // reference max, log2 domain
      float l = 0.f;

      for (int j = 0; j < p.nkb; ++j, ++n_blk) {
        const int nvk = min(kKV, p.T - j * kKV);   // keys of this block that exist
        bool fast = (nvk == kKV);
        if (kMasked) {
          float bv[2], be[2];
          bool all_ok = true;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const bool in_range = j * kKV + lane * 2 + i < p.T;
            const bool ok = in_range && ((raw_next >> i) & 1u) != 0;
            bv[i] = ok ? 0.f : -INFINITY;
            be[i] = in_range ? 0.f : -INFINITY;
            all_ok = all_ok && ok;
          }
          if (j + 1 < p.nkb) raw_next = load_raw(seq_base, j + 1);
          *reinterpret_cast<float2*>(myBiasV + 2 * lane) = make_float2(bv[0], bv[1]);
          *reinterpret_cast<float2*>(myBiasE + 2 * lane) = make_float2(be[0], be[1]);
          fast = __all_sync(0xffffffffu, all_ok);   // also orders the smem writes above before the reads below
        }
        mbar_wait(B_(s, S_FULL), n_blk & 1);
        tc_fence_after();
        // x32 loads: ptxas checks an instruction's register block against the LAUNCH bound (96 at 640 threads), not
        // against the setmaxnreg budget of the warpgroup
        float sc[kKV];
        {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(lane_addr + colS, v0);
          tmem_ld_32x32(lane_addr + colS + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            sc[i] = __uint_as_float(v0[i]);
            sc[32 + i] = __uint_as_float(v1[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B_(s, S_FREE));

        float cs = cq;   // score scale used for this block
        if (!fast) {
          if (kMasked) {
            // additive per-key bias (0 / -inf) folded into one FFMA per score; the vector is chosen per THREAD
            cs = kLog2e;
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
            // unmasked: only the tail of the last key block is out of range
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
        // ---- P = exp2(s * cs - m), row sum, bf16 pack (ex2 in batches of 8, sums / packs one batch behind) ----
        const float neg_m = -m2;
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
        // P (and O) may only be touched once the previous PV of this stream has retired -- also across tiles
        if (n_blk > 0) {
          mbar_wait(B_(s, PV_DONE), (n_blk - 1) & 1);
          tc_fence_after();
          if (j > 0 && __any_sync(0xffffffffu, factor != 1.f)) {
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
        tmem_st_32x32(lane_addr + colP, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B_(s, P_FULL));
      }
      // hand (l, m) to the epilogue warps and move on
      *reinterpret_cast<float2*>(smem + kOffLM + ((s * 2 + (n_t & 1)) * kQ + row) * 8) = make_float2(l, m2);
      __syncwarp();
      if (lane == 0) mbar_arrive(B_(s, LM_FULL));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == kMmaWarp0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Called by launch_attention_tc (attention_tc.cu) for calls without group ids when MR_ATTN_TILES selects it.
int32_t launch_attention_tc3(cudaStream_t stream, const bf16* qkv, const uint8_t* valid, bf16* out, float* lse, int B,
                             int T, int H) {
  CUtensorMap tm_q, tm_kv, tm_out;
  if (!make_tmap_2d(&tm_q, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, kQ, kD)) return MR_ERR_DRIVER;
  if (!make_tmap_2d(&tm_kv, qkv, static_cast<int64_t>(B) * T, 3 * H, 3 * H, kKV, kD)) return MR_ERR_DRIVER;
  if (!make_tmap_3d(&tm_out, out, H, T, B, H, static_cast<int64_t>(T) * H, 32, kD)) return MR_ERR_DRIVER;
  Attn3Params p{};
  p.valid = valid;
  p.lse = lse;
  p.B = B;
  p.T = T;
  p.H = H;
  p.nh = H / kD;
  p.nqb = ceil_div(T, kQ);
  p.nkb = ceil_div(T, kKV);
  const int64_t tiles = static_cast<int64_t>(B) * p.nh * p.nqb;
  if (tiles >= (int64_t(1) << 31)) return MR_ERR_BAD_SHAPE;
  p.num_tiles = static_cast<int>(tiles);
  const int sms = device_sm_count();
  if (sms <= 0) return MR_ERR_UNSUPPORTED_ARCH;
  const int want = static_cast<int>((tiles + kStreams - 1) / kStreams);
  const int grid = want < sms ? want : sms;
  const bool masked = valid != nullptr;
  auto kern = masked ? attention_tc3_kernel<true> : attention_tc3_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[masked ? 1 : 0]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem3);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set[masked ? 1 : 0] = true;
  }
  kern<<<grid, kThreads3, kSmem3, stream>>>(tm_q, tm_kv, tm_out, p);
  return after_launch();
}

}  // namespace mr
