// Shared host-side plumbing for libmreserve_b200: status codes, launch accounting, argument checks.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/mreserve_b200.h"

namespace mr {

using bf16 = __nv_bfloat16;

extern std::atomic<int64_t> g_launches;

inline int32_t cuda_status(cudaError_t e) {
  return e == cudaSuccess ? MR_OK : (MR_ERR_CUDA_BASE - static_cast<int32_t>(e));
}

// Call right after a <<<>>> / cudaLaunchKernelEx: counts the launch and converts launch errors.
inline int32_t after_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cuda_status(cudaGetLastError());
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int device_sm_count();

// 2-D bf16 tensor map over [rows, cols] with row pitch `ld` elements, box [box_rows, box_cols],
// 128-byte swizzle (box_cols * 2 bytes must be 128), zero fill out of bounds.
bool make_tmap_2d(CUtensorMap* m, const void* base, int64_t rows, int cols, int64_t ld, int box_rows, int box_cols);

// 3-D bf16 tensor map over [batch, rows, cols] (row pitch `ld`, batch pitch `batch_stride` elements), box
// [1, box_rows, box_cols], 128-byte swizzle.  Rows past `rows` are clipped by the TMA unit, so a tile that
// overhangs the end of one sequence never touches the next sequence of a concatenated [batch * rows, cols] buffer.
bool make_tmap_3d(CUtensorMap* m, const void* base, int cols, int rows, int batch, int64_t ld, int64_t batch_stride,
                  int box_rows, int box_cols);

bool make_tmap_3d_f32(CUtensorMap* m, const void* base, int cols, int rows, int batch, int64_t ld, int64_t batch_stride,
                      int box_rows, int box_cols);

}  // namespace mr

#define MR_REQUIRE(cond, code) \
  do {                         \
    if (!(cond)) return (code); \
  } while (0)
