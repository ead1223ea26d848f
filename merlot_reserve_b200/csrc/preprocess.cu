// Device preprocessing in front of the encoders: what the reference does on the host with
// librosa / TensorFlow, as B200 kernels that write the model's input tensors directly.
//
//   mr_logmel           make_spectrogram                 mreserve/preprocess.py:121-156
//   mr_resize_patchify  preprocess_image_to_patches      mreserve/preprocess.py:464-481 (+ pretrain/data_utils.py:42-147)
//
// mr_logmel: the STFT (n_fft 1536, hop 588, symmetric Hann, centred with reflect padding) is ONE tcgen05 GEMM of the
// frame matrix against the windowed DFT matrix [cos | -sin]: frames [M, 1536] x W [1536, 2 * 769].  bf16 operands
// alone would put the leakage floor of a loud tone above the 0.1 offset of the log, so both operands are split into
// bf16 hi + lo parts and the three significant products ride in one GEMM along K: A' = [hi | lo | hi] (K = 4608),
// W' = [Whi ; Whi ; Wlo] -- an fp32 accumulate of ~16-bit operands.  Then |X|^2, the 64 Slaney mel filters
// (<= 2 filters per bin: a sparse gather), log(mel + 0.1) - log(0.1), and the playback-speed column.
//
// mr_resize_patchify: uint8 frames of any size -> convert_image_dtype -> antialiased bilinear resize (TensorFlow's
// ScaleAndTranslate with the triangle kernel, columns then rows, fp32) to the aspect-preserving size -> clip ->
// zero pad to the model's resolution -> space_to_depth(16) -> bf16 patch rows [B, h1*w1, 768].  One fused pass:
// every output pixel reads its (taps_y x taps_x) source window; nothing but the patches is written.
#include "common.cuh"

namespace mr {
namespace {

constexpr int kNfft = 1536, kHop = 588, kBins = kNfft / 2 + 1, kMels = 64;
constexpr int kDftN = 2 * kBins;                  // 1538 output columns: re | im
constexpr int kDftLd = (kDftN + 7) / 8 * 8;       // 1544
constexpr int kK3 = 3 * kNfft;                    // 4608
constexpr double kPiD = 3.14159265358979323846;

inline size_t align256(size_t n) { return (n + 255) / 256 * 256; }

// layout of the persistent tables buffer
struct TablesLayout {
  size_t w, fb, rng, total;   // W' bf16 [kDftN, kK3]; mel basis f32 [kMels, kBins]; non-zero bin range int [kMels][2]
};
TablesLayout tables_layout() {
  TablesLayout t;
  t.w = 0;
  t.fb = align256(static_cast<size_t>(kDftN) * kK3 * sizeof(bf16));
  t.rng = t.fb + align256(static_cast<size_t>(kMels) * kBins * sizeof(float));
  t.total = t.rng + align256(static_cast<size_t>(kMels) * 2 * sizeof(int));
  return t;
}

// W'[c, k']: c < 769: hann[n] cos(2 pi c n / N); c >= 769: -hann[n] sin(2 pi (c - 769) n / N); k' = part * N + n with
// part 0, 1 = hi(bf16), part 2 = lo.  Double precision (computed once per process).
__global__ void __launch_bounds__(256)
dft_table_kernel(bf16* __restrict__ w) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(kDftN) * kNfft) return;
  const int c = static_cast<int>(i / kNfft), n = static_cast<int>(i % kNfft);
  const int k = c < kBins ? c : c - kBins;
  const double win = 0.5 - 0.5 * cospi(2.0 * n / (kNfft - 1));                 // scipy hann, sym=True
  const int ph = static_cast<int>((static_cast<int64_t>(k) * n) % kNfft);      // exact phase reduction
  double sn, cs;
  sincospi(2.0 * ph / kNfft, &sn, &cs);
  const float v = static_cast<float>(c < kBins ? win * cs : -win * sn);
  const bf16 hi = __float2bfloat16(v);
  const bf16 lo = __float2bfloat16(v - __bfloat162float(hi));
  bf16* row = w + static_cast<size_t>(c) * kK3;
  row[n] = hi;
  row[kNfft + n] = hi;
  row[2 * kNfft + n] = lo;
}

__device__ __forceinline__ double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
__device__ __forceinline__ double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// librosa.filters.mel(sr 22050, n_fft 1536, n_mels 64, fmin 20, fmax 11025, htk=False, norm='slaney') -> f32 [64, 769]
__global__ void __launch_bounds__(256)
mel_table_kernel(float* __restrict__ fb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kMels * kBins) return;
  const int m = i / kBins, f = i % kBins;
  const double m_lo = hz_to_mel(20.0), m_hi = hz_to_mel(11025.0), step = (m_hi - m_lo) / (kMels + 1);
  const double f0 = mel_to_hz(m_lo + step * m), f1 = mel_to_hz(m_lo + step * (m + 1)), f2 = mel_to_hz(m_lo + step * (m + 2));
  const double freq = 11025.0 * f / (kBins - 1);
  const double lower = (freq - f0) / (f1 - f0), upper = (f2 - freq) / (f2 - f1);
  const double wv = fmax(0.0, fmin(lower, upper)) * (2.0 / (f2 - f0));
  fb[i] = static_cast<float>(wv);
}

// [first, last + 1) non-zero bin of every mel filter (the triangles are contiguous)
__global__ void mel_range_kernel(const float* __restrict__ fb, int* __restrict__ rng) {
  const int m = threadIdx.x;
  if (m >= kMels) return;
  int lo = kBins, hi = 0;
  for (int f = 0; f < kBins; ++f)
    if (fb[m * kBins + f] != 0.f) {
      lo = min(lo, f);
      hi = f + 1;
    }
  rng[2 * m] = lo < hi ? lo : 0;
  rng[2 * m + 1] = lo < hi ? hi : 0;
}

// A'[m, :] for frame m = (b, t): samples ypad[t * hop + n], ypad = reflect-padded waveform (n_fft / 2 each side)
__global__ void __launch_bounds__(256)
logmel_frames_kernel(const float* __restrict__ wav, bf16* __restrict__ a, int b0, int nb, int n_samples, int n_frames) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(nb) * n_frames * (kNfft / 2);
  if (i >= total) return;
  const int n = static_cast<int>(i % (kNfft / 2)) * 2;
  const int64_t m = i / (kNfft / 2);
  const int t = static_cast<int>(m % n_frames), b = b0 + static_cast<int>(m / n_frames);
  float v[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int j = t * kHop + n + q - kNfft / 2;              // index into the un-padded waveform
    if (j < 0) j = -j;                                 // np.pad(mode='reflect'): no edge repeat
    if (j >= n_samples) j = 2 * (n_samples - 1) - j;
    v[q] = wav[static_cast<size_t>(b) * n_samples + j];
  }
  const bf16 h0 = __float2bfloat16(v[0]), h1 = __float2bfloat16(v[1]);
  const bf16 l0 = __float2bfloat16(v[0] - __bfloat162float(h0)), l1 = __float2bfloat16(v[1] - __bfloat162float(h1));
  bf16* row = a + static_cast<size_t>(m) * kK3 + n;
  const __nv_bfloat162 hh = __halves2bfloat162(h0, h1), ll = __halves2bfloat162(l0, l1);
  *reinterpret_cast<__nv_bfloat162*>(row) = hh;
  *reinterpret_cast<__nv_bfloat162*>(row + kNfft) = ll;
  *reinterpret_cast<__nv_bfloat162*>(row + 2 * kNfft) = hh;
}

// one warp per frame: |X|^2 -> mel -> log(mel + 0.1) - log(0.1); lane l owns mel filters l and l + 32
__global__ void __launch_bounds__(256)
logmel_mel_kernel(const float* __restrict__ spec, const float* __restrict__ fb, const int* __restrict__ rng,
                  float* __restrict__ out, int b0, int nb, int n_frames, int pad_size, int split, float playback) {
  __shared__ float pw[8][kBins + 3];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = static_cast<int64_t>(blockIdx.x) * 8 + wid;
  if (m >= static_cast<int64_t>(nb) * n_frames) return;
  const float* row = spec + static_cast<size_t>(m) * kDftLd;
  for (int f = lane; f < kBins; f += 32) {
    const float re = row[f], im = row[kBins + f];
    pw[wid][f] = re * re + im * im;
  }
  __syncwarp();
  const int t = static_cast<int>(m % n_frames), b = b0 + static_cast<int>(m / n_frames);
  float* dst;
  if (split) {   // three 60-frame sub-clips with `pad_size` gaps (preprocess.py:151-155); gap frames are dropped
    const int seq = 60, u = t - pad_size, blk = u / (seq + pad_size), r = u % (seq + pad_size);
    if (u < 0 || blk >= 3 || r >= seq) return;
    dst = out + ((static_cast<size_t>(b) * 3 + blk) * seq + r) * 65;
  } else {
    dst = out + (static_cast<size_t>(b) * n_frames + t) * 65;
  }
  const float log_eps = logf(0.1f);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int mel = lane + 32 * h;
    const float* w = fb + static_cast<size_t>(mel) * kBins;
    float acc = 0.f;
    const int f_end = __ldg(rng + 2 * mel + 1);
    for (int f = __ldg(rng + 2 * mel); f < f_end; ++f) acc = fmaf(__ldg(w + f), pw[wid][f], acc);
    dst[mel] = logf(acc + 0.1f) - log_eps;
  }
  if (lane == 0) dst[64] = playback;
}

// ---------------------------------------------------------------------------------------------------------------
// resize + pad + patchify
// ---------------------------------------------------------------------------------------------------------------
struct Spans {
  int taps_x, taps_y, sw, sh;
};

__host__ __device__ inline int span_taps(int in_size, int out_size) {
  const float inv_scale = 1.0f / (static_cast<float>(out_size) / static_cast<float>(in_size));
  const float ks = inv_scale > 1.0f ? inv_scale : 1.0f;
  const int t = 2 * static_cast<int>(ceilf(ks)) + 1;
  return t < in_size ? t : in_size;
}

// TensorFlow ScaleAndTranslate spans (triangle kernel, antialias): one thread per output coordinate of one axis.
// Every operation is an explicitly rounded fp32 op in the order of oracle/preprocess_oracle.py:resize_spans.
__global__ void __launch_bounds__(128)
resize_spans_kernel(int in_size, int out_size, int taps, int* __restrict__ starts, float* __restrict__ weights) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= out_size) return;
  const float scale = __fdiv_rn(static_cast<float>(out_size), static_cast<float>(in_size));
  const float inv_scale = __fdiv_rn(1.0f, scale);
  const float ks = fmaxf(inv_scale, 1.0f);
  const float sample_f = __fmul_rn(__fadd_rn(static_cast<float>(x), 0.5f), inv_scale);
  int lo = static_cast<int>(ceilf(__fadd_rn(__fadd_rn(sample_f, -ks), -0.5f)));
  int hi = static_cast<int>(floorf(__fadd_rn(__fadd_rn(sample_f, ks), -0.5f)));
  lo = min(max(lo, 0), in_size - 1);
  hi = min(max(hi, 0), in_size - 1) + 1;
  starts[x] = lo;
  float total = 0.f;
  float* w = weights + static_cast<size_t>(x) * taps;
  for (int j = 0; j < taps; ++j) {
    float wv = 0.f;
    if (lo + j < hi) {
      const float pos = __fdiv_rn(__fadd_rn(__fadd_rn(static_cast<float>(lo + j), 0.5f), -sample_f), ks);
      wv = fmaxf(0.f, __fadd_rn(1.0f, -fabsf(pos)));
      total = __fadd_rn(total, wv);
    }
    w[j] = wv;
  }
  if (total > 0.f)
    for (int j = 0; j < taps; ++j) w[j] = __fdiv_rn(w[j], total);
}

template <bool kOutF32>
__global__ void __launch_bounds__(256)
resize_patchify_kernel(const uint8_t* __restrict__ frames, void* __restrict__ out_, const int* __restrict__ sx,
                       const float* __restrict__ wx, const int* __restrict__ sy, const float* __restrict__ wy, int B,
                       int in_h, int in_w, int grid_h, int grid_w, Spans s) {
  const int out_h = grid_h * 16, out_w = grid_w * 16;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * out_h * out_w) return;
  // thread order = output memory order: (b, by, bx, py, px)
  const int px = static_cast<int>(i % 16), py = static_cast<int>((i / 16) % 16);
  const int bx = static_cast<int>((i / 256) % grid_w), by = static_cast<int>((i / (256 * grid_w)) % grid_h);
  const int b = static_cast<int>(i / (static_cast<int64_t>(256) * grid_w * grid_h));
  const int y = by * 16 + py, x = bx * 16 + px;
  float o[3] = {0.f, 0.f, 0.f};
  if (y < s.sh && x < s.sw) {
    const uint8_t* img = frames + static_cast<size_t>(b) * in_h * in_w * 3;
    const int x0 = sx[x], y0 = sy[y];
    const float* wxr = wx + static_cast<size_t>(x) * s.taps_x;
    const float* wyr = wy + static_cast<size_t>(y) * s.taps_y;
    for (int j = 0; j < s.taps_y; ++j) {
      const int r = min(y0 + j, in_h - 1);
      const uint8_t* rowp = img + static_cast<size_t>(r) * in_w * 3;
      float h[3] = {0.f, 0.f, 0.f};
      for (int k = 0; k < s.taps_x; ++k) {       // columns first (GatherColumns), taps in increasing source order
        const int c = min(x0 + k, in_w - 1);
        const float w = wxr[k];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float v = __fmul_rn(static_cast<float>(rowp[c * 3 + ch]), 1.0f / 255.0f);   // convert_image_dtype
          h[ch] = __fadd_rn(h[ch], __fmul_rn(w, v));
        }
      }
      const float w = wyr[j];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) o[ch] = __fadd_rn(o[ch], __fmul_rn(w, h[ch]));            // then rows
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) o[ch] = fminf(fmaxf(o[ch], 0.f), 1.f);                      // clip_by_value
  }
  const size_t off = static_cast<size_t>(i) * 3;
  if (kOutF32) {
    float* dst = static_cast<float*>(out_) + off;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
  } else {
    bf16* dst = static_cast<bf16*>(out_) + off;
    dst[0] = __float2bfloat16(o[0]); dst[1] = __float2bfloat16(o[1]); dst[2] = __float2bfloat16(o[2]);
  }
}

struct ResizeCarve {
  size_t sx, wx, sy, wy, total;
};
bool resize_geometry(const MrResizePatchifyDesc* d, Spans* s) {
  if (!d || d->B < 0 || d->in_h <= 0 || d->in_w <= 0 || d->grid_h <= 0 || d->grid_w <= 0 || d->grid_h > 24 || d->grid_w > 32)
    return false;                                           // preprocess.py:474-475
  const float dh = static_cast<float>(d->grid_h * 16), dw = static_cast<float>(d->grid_w * 16);
  const float h = static_cast<float>(d->in_h), w = static_cast<float>(d->in_w);
  const float sc_x = dw / w, sc_y = dh / h;
  const float image_scale = sc_x < sc_y ? sc_x : sc_y;       // data_utils.py:103-105
  s->sh = static_cast<int>(h * image_scale);
  s->sw = static_cast<int>(w * image_scale);
  if (s->sh <= 0 || s->sw <= 0) return false;
  s->taps_x = span_taps(d->in_w, s->sw);
  s->taps_y = span_taps(d->in_h, s->sh);
  return true;
}
ResizeCarve resize_carve(const Spans& s) {
  ResizeCarve c;
  c.sx = 0;
  c.wx = c.sx + align256(static_cast<size_t>(s.sw) * 4);
  c.sy = c.wx + align256(static_cast<size_t>(s.sw) * s.taps_x * 4);
  c.wy = c.sy + align256(static_cast<size_t>(s.sh) * 4);
  c.total = c.wy + align256(static_cast<size_t>(s.sh) * s.taps_y * 4);
  return c;
}

constexpr int kLogmelChunk = 64;    // audio segments per GEMM (bounds the workspace: 64 x 188 frames)

struct LogmelCarve {
  size_t a, spec, total;
};
bool logmel_geometry(const MrLogmelDesc* d, int* n_frames) {
  if (!d || d->B < 0 || d->n_samples < kNfft / 2 + 1 || d->playback_speed != 1) return false;
  *n_frames = 1 + d->n_samples / kHop;                       // librosa stft, center=True
  if (d->split && *n_frames != 180 + 4 * d->pad_size) return false;   // preprocess.py:148-149
  return d->pad_size >= 0;
}
LogmelCarve logmel_carve(int B, int n_frames) {
  const size_t M = static_cast<size_t>(B < kLogmelChunk ? B : kLogmelChunk) * n_frames;
  LogmelCarve c;
  c.a = 0;
  c.spec = align256(M * kK3 * sizeof(bf16));
  c.total = c.spec + align256(M * kDftLd * sizeof(float));
  return c;
}

}  // namespace
}  // namespace mr

extern "C" size_t mr_logmel_tables_bytes(void) { return mr::tables_layout().total; }

extern "C" size_t mr_logmel_workspace_bytes(const MrLogmelDesc* d) {
  int nf;
  if (!mr::logmel_geometry(d, &nf) || d->B == 0) return 0;
  return mr::logmel_carve(d->B, nf).total;
}

extern "C" int32_t mr_logmel(mr_stream_t stream_, const MrLogmelDesc* d, const void* const* in, void* const* out,
                             void* workspace, size_t workspace_bytes) {
  using namespace mr;
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  int nf;
  MR_REQUIRE(logmel_geometry(d, &nf), MR_ERR_BAD_SHAPE);
  MR_REQUIRE(in[1] && aligned16(in[1]), in[1] ? MR_ERR_MISALIGNED : MR_ERR_NULL_POINTER);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const TablesLayout tl = tables_layout();
  uint8_t* tables = static_cast<uint8_t*>(const_cast<void*>(in[1]));
  bf16* w = reinterpret_cast<bf16*>(tables + tl.w);
  float* fb = reinterpret_cast<float*>(tables + tl.fb);
  int* rng = reinterpret_cast<int*>(tables + tl.rng);
  if (d->build_tables) {
    dft_table_kernel<<<static_cast<unsigned>(ceil_div64(static_cast<int64_t>(kDftN) * kNfft, 256)), 256, 0, stream>>>(w);
    int32_t st = after_launch();
    if (st != MR_OK) return st;
    mel_table_kernel<<<ceil_div(kMels * kBins, 256), 256, 0, stream>>>(fb);
    if ((st = after_launch()) != MR_OK) return st;
    mel_range_kernel<<<1, kMels, 0, stream>>>(fb, rng);
    if ((st = after_launch()) != MR_OK) return st;
  }
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && out[0], MR_ERR_NULL_POINTER);
  const LogmelCarve c = logmel_carve(d->B, nf);
  MR_REQUIRE(workspace && workspace_bytes >= c.total, MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(workspace), MR_ERR_MISALIGNED);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  bf16* a = reinterpret_cast<bf16*>(ws + c.a);
  float* spec = reinterpret_cast<float*>(ws + c.spec);
  for (int b0 = 0; b0 < d->B; b0 += kLogmelChunk) {
    const int nb = d->B - b0 < kLogmelChunk ? d->B - b0 : kLogmelChunk;
    const int M = nb * nf;
    logmel_frames_kernel<<<static_cast<unsigned>(ceil_div64(static_cast<int64_t>(M) * (kNfft / 2), 256)), 256, 0, stream>>>(
        static_cast<const float*>(in[0]), a, b0, nb, d->n_samples, nf);
    int32_t st = after_launch();
    if (st != MR_OK) return st;
    MrGemmDesc g{};
    g.M = M;
    g.N = kDftN;
    g.K = kK3;
    g.lda = kK3;
    g.ldd = kDftLd;
    g.epilogue = MR_EPI_OUT_F32;
    const void* gin[7] = {a, w, nullptr, nullptr, nullptr, nullptr, nullptr};
    void* gout[2] = {spec, nullptr};
    if ((st = mr_gemm_epilogue(stream_, &g, gin, gout, nullptr, 0)) != MR_OK) return st;
    logmel_mel_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(spec, fb, rng, static_cast<float*>(out[0]), b0, nb, nf, d->pad_size,
                                                          d->split, static_cast<float>(d->playback_speed));
    if ((st = after_launch()) != MR_OK) return st;
  }
  return MR_OK;
}

extern "C" size_t mr_resize_patchify_workspace_bytes(const MrResizePatchifyDesc* d) {
  mr::Spans s;
  if (!mr::resize_geometry(d, &s)) return 0;
  return mr::resize_carve(s).total;
}

extern "C" int32_t mr_resize_patchify(mr_stream_t stream_, const MrResizePatchifyDesc* d, const void* const* in,
                                      void* const* out, void* workspace, size_t workspace_bytes) {
  using namespace mr;
  MR_REQUIRE(d && in && out, MR_ERR_NULL_POINTER);
  Spans s;
  MR_REQUIRE(resize_geometry(d, &s), MR_ERR_BAD_SHAPE);
  if (d->B == 0) return MR_OK;
  MR_REQUIRE(in[0] && out[0], MR_ERR_NULL_POINTER);
  const ResizeCarve c = resize_carve(s);
  MR_REQUIRE(workspace && workspace_bytes >= c.total, MR_ERR_WORKSPACE);
  MR_REQUIRE(aligned16(workspace), MR_ERR_MISALIGNED);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* sx = reinterpret_cast<int*>(ws + c.sx);
  float* wx = reinterpret_cast<float*>(ws + c.wx);
  int* sy = reinterpret_cast<int*>(ws + c.sy);
  float* wy = reinterpret_cast<float*>(ws + c.wy);
  resize_spans_kernel<<<ceil_div(s.sw, 128), 128, 0, stream>>>(d->in_w, s.sw, s.taps_x, sx, wx);
  int32_t st = after_launch();
  if (st != MR_OK) return st;
  resize_spans_kernel<<<ceil_div(s.sh, 128), 128, 0, stream>>>(d->in_h, s.sh, s.taps_y, sy, wy);
  if ((st = after_launch()) != MR_OK) return st;
  const int64_t total = static_cast<int64_t>(d->B) * d->grid_h * d->grid_w * 256;
  const unsigned blocks = static_cast<unsigned>(ceil_div64(total, 256));
  if (d->out_f32)
    resize_patchify_kernel<true><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(in[0]), out[0], sx, wx, sy, wy, d->B,
                                                            d->in_h, d->in_w, d->grid_h, d->grid_w, s);
  else
    resize_patchify_kernel<false><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(in[0]), out[0], sx, wx, sy, wy, d->B,
                                                             d->in_h, d->in_w, d->grid_h, d->grid_w, s);
  return after_launch();
}
