// Library runtime: error plumbing, launch counter, device queries, TMA tensor-map encoding.
#include "msp_common.cuh"
#include "../../include/msp_b200.h"
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void msp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void msp_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* msp_last_error(void) { return g_err; }
extern "C" int msp_version(void) { return MSP_ABI_VERSION; }
extern "C" long long msp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

bool msp_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MSP_PDL");
    on = (e && e[0] == '1') ? 1 : 0;  // opt-in: inside the step's CUDA graph it measured no gain (22.67 vs 22.43 ms / step)
  }
  return on != 0;
}

int msp_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      sms = 148;
  }
  return sms;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int msp_encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box,
                         const uint32_t* elem_strides, int swizzle_bytes) {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || p == nullptr || qres != cudaDriverEntryPointSuccess) {
      msp_set_error("cuTensorMapEncodeTiled entry point unavailable (cuda error %d)", (int)e);
      return MSP_ERR_DRIVER;
    }
    fn = (PFN_encodeTiled)p;
  }
  if (((uintptr_t)base & 15u) != 0) {
    msp_set_error("tensor map: base address %p not 16-byte aligned", base);
    return MSP_ERR_ARG;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides[i];
    if (box[i] == 0 || box[i] > 256) {
      msp_set_error("tensor map: box[%d]=%u out of range", i, box[i]);
      return MSP_ERR_UNSUPPORTED;
    }
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gs[i] = strides_bytes[i];
    if (gs[i] % 16 != 0) {
      msp_set_error("tensor map: stride[%d]=%llu not a multiple of 16 bytes", i,
                    (unsigned long long)gs[i]);
      return MSP_ERR_ARG;
    }
  }
  const CUtensorMapSwizzle sw = swizzle_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd,
                  gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    msp_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu,%llu box "
                  "%u,%u,%u,%u)",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
                  rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return MSP_ERR_DRIVER;
  }
  return MSP_OK;
}
