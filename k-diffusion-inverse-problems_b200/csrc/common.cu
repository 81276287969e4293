// Host-side plumbing of libkdip: thread-local error message, device checks, TMA descriptor encoding.
#include <cudaTypedefs.h>
#include <stdarg.h>

#include <atomic>

#include "kdip_common.cuh"

namespace kdip {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return KDIP_ECUDA;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int encode_tmap_bf16_4d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint32_t b0,
                        uint32_t b1, uint32_t b2, uint32_t b3) {
  auto fn = get_encode_fn();
  KDIP_REQUIRE(fn != nullptr, KDIP_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {d0 * 2, d0 * d1 * 2, d0 * d1 * d2 * 2};
  cuuint32_t box[4] = {b0, b1, b2, b3};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KDIP_REQUIRE(r == CUDA_SUCCESS, KDIP_ECUDA,
               "cuTensorMapEncodeTiled(4d) failed: %d dims=(%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u)", (int)r,
               (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)d3, b0, b1, b2, b3);
  return KDIP_OK;
}

int encode_tmap_bf16_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t bcols, uint32_t brows) {
  auto fn = get_encode_fn();
  KDIP_REQUIRE(fn != nullptr, KDIP_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {bcols, brows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KDIP_REQUIRE(r == CUDA_SUCCESS, KDIP_ECUDA, "cuTensorMapEncodeTiled(2d) failed: %d dims=(%llu,%llu) box=(%u,%u)", (int)r,
               (unsigned long long)cols, (unsigned long long)rows, bcols, brows);
  return KDIP_OK;
}

}  // namespace kdip

extern "C" const char* kdip_last_error(void) { return kdip::g_err; }
extern "C" int kdip_version(void) { return 1; }
extern "C" unsigned long long kdip_launch_count(void) { return kdip::g_launches.load(std::memory_order_relaxed); }
extern "C" void kdip_launch_count_add(size_t n) { kdip::g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

extern "C" int kdip_device_check(int* sm_count) {
  int dev = 0;
  KDIP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  KDIP_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  KDIP_REQUIRE(prop.major == 10, KDIP_EINVAL, "libkdip is built for sm_100a only; device is sm_%d%d (%s)", prop.major,
               prop.minor, prop.name);
  return KDIP_OK;
}
