// Library-level entry points: device probe, status strings, launch accounting.
#include "common.cuh"

namespace mr {

std::atomic<int64_t> g_launches{0};

int device_sm_count() {
  // cached per device; cudaGetDevice is cheap and keeps the library free of global mutable state
  // beyond this read-only cache
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cache[dev] = n;
  }
  return cache[dev];
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // resolved through the runtime so the library does not link against libcuda directly
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

bool make_tmap_2d(CUtensorMap* m, const void* base, int64_t rows, int cols, int64_t ld, int box_rows, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool make_tmap_3d(CUtensorMap* m, const void* base, int cols, int rows, int batch, int64_t ld, int64_t batch_stride,
                  int box_rows, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(batch_stride) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// fp32 flavour of make_tmap_3d (the dQ accumulator of the attention backward: TMA reduce-add boxes of 32 floats x box_rows)
bool make_tmap_3d_f32(CUtensorMap* m, const void* base, int cols, int rows, int batch, int64_t ld, int64_t batch_stride,
                      int box_rows, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, static_cast<cuuint64_t>(batch_stride) * 4};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

const char* nccl_error_string(int code);  // collectives.cu

}  // namespace mr

extern "C" int32_t mr_abi_version(void) { return 17; }

extern "C" int64_t mr_launch_count(void) { return mr::g_launches.load(std::memory_order_relaxed); }

extern "C" int32_t mr_check_device(int32_t device) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return mr::cuda_status(e);
  e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e != cudaSuccess) return mr::cuda_status(e);
  return major == 10 ? MR_OK : MR_ERR_UNSUPPORTED_ARCH;
}

extern "C" const char* mr_status_string(int32_t status) {
  switch (status) {
    case MR_OK: return "ok";
    case MR_ERR_BAD_SHAPE: return "bad shape or unsupported argument combination";
    case MR_ERR_MISALIGNED: return "pointer not 16-byte aligned";
    case MR_ERR_UNSUPPORTED_ARCH: return "device is not sm_100 (B200)";
    case MR_ERR_NULL_POINTER: return "required pointer is NULL";
    case MR_ERR_WORKSPACE: return "workspace too small";
    case MR_ERR_DRIVER: return "CUDA driver call failed (tensor map encode)";
    case MR_ERR_NCCL_UNAVAILABLE: return "libnccl.so.2 could not be loaded";
    default: break;
  }
  if (status <= MR_ERR_NCCL_BASE) return mr::nccl_error_string(MR_ERR_NCCL_BASE - status);
  if (status <= MR_ERR_CUDA_BASE) return cudaGetErrorString(static_cast<cudaError_t>(MR_ERR_CUDA_BASE - status));
  return "unknown status";
}
