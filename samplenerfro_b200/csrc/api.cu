// Error reporting, launch accounting and ABI version of librnerf_b200.so.
#include <stdarg.h>
#include <atomic>
#include "umma.cuh"

namespace rnerf {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int make_rows_tmap(CUtensorMap* out, const void* base, int64_t n_rows, int n_layers) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver (%s)", cudaGetErrorString(e));
      return e != cudaSuccess ? (int)e : RNERF_E_ARCH;
    }
    encode = (EncodeFn)fn;
  }
  if (n_rows <= 0 || n_rows >= (1ll << 31) || n_layers <= 0 || (reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
    set_error("make_rows_tmap: bad activation dump (rows %lld, layers %d)", (long long)n_rows, n_layers);
    return RNERF_E_SHAPE;
  }
  const cuuint64_t dims[3] = {256, (cuuint64_t)n_rows, (cuuint64_t)n_layers};
  const cuuint64_t strides[2] = {512, (cuuint64_t)n_rows * 512};      // bytes, dims 1 and 2
  const cuuint32_t box[3] = {64, 32, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return RNERF_E_SHAPE;
  }
  return 0;
}

}  // namespace rnerf

extern "C" {

int rnerf_abi_version(void) { return RNERF_ABI_VERSION; }
const char* rnerf_last_error(void) { return rnerf::g_err; }
uint64_t rnerf_launch_count(void) { return rnerf::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
