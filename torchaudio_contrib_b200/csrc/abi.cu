// C-ABI glue: error strings, device queries.  The entry points themselves live next to their
// kernels (mulaw.cu, pointwise.cu, stft.cu, melbank.cu, pipeline.cu).
#include <stdarg.h>

#include "tac_common.cuh"

namespace tac {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace tac

extern "C" int tac_version(void) { return TAC_ABI_VERSION; }

extern "C" const char* tac_last_error(void) { return tac::last_error_buffer(); }

extern "C" int tac_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  using namespace tac;
  int dev = 0;
  TAC_CUDA_OK(cudaGetDevice(&dev));
  int sms = 0, major = 0, minor = 0;
  TAC_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TAC_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TAC_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  return TAC_OK;
}
