// Micro-probe: SFU throughput of ex2.approx.ftz.f32 vs ex2.approx.ftz.bf16x2 (and FFMA for scale) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>
__global__ void probe(float* out, int iters, float seed) {
  float a[8];
  uint32_t b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + 0.001f * (threadIdx.x + i);
    b[i] = 0x3c003c00u + threadIdx.x + i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b[i]));
      if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b[i]));
      if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (MODE == 6) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b[i]) : "f"(a[i]), "f"(a[i - 1]));
      }
      if (MODE == 7) {
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 4) & 7]) : "f"(a[i]));
      }
      if (MODE == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(b[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  const int iters = 4096, blocks = 148 * 2, threads = 1024;
  probe<MODE><<<blocks, threads>>>(out, 16, 0.5f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = double(blocks) * (threads / 32) * iters * 8;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %.3f ms  %.2f Gwarp-instr/s  => %.2f warp-instr/clk/SM at %d MHz (nominal)\n", name, ms, warp_instr / ms / 1e6,
         warp_instr / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32");
  run<1>("ex2.approx.ftz.bf16x2");
  run<4>("ex2.approx.f16x2");
  run<3>("tanh.approx.f32");
  run<2>("fma.rn.f32");
  run<5>("cvt.rn.bf16x2.f32");
  run<6>("ex2 + 0.5 cvt.bf16x2");
  run<7>("fma+ex2+add");
  run<8>("max.f32");
  return 0;
}
