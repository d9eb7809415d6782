// C-ABI glue: error strings, device queries.  The entry points themselves live next to their
// kernels (mulaw.cu, pointwise.cu, stft.cu, melbank.cu, pipeline.cu).
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

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

static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_profile{0};
struct ProbeRecord {
  int kind;
  cudaEvent_t start, stop;
};
static std::mutex g_probe_mutex;
static std::vector<ProbeRecord> g_probe_log;

LaunchProbe::LaunchProbe(int kind, cudaStream_t stream) : kind_(kind), stream_(stream), slot_(-1) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_profile.load(std::memory_order_relaxed)) return;
  ProbeRecord r;
  r.kind = kind;
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  cudaEventRecord(r.start, stream);
  std::lock_guard<std::mutex> lock(g_probe_mutex);
  g_probe_log.push_back(r);
  slot_ = (int)g_probe_log.size() - 1;
}

LaunchProbe::~LaunchProbe() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lock(g_probe_mutex);
  if (slot_ < (int)g_probe_log.size()) cudaEventRecord(g_probe_log[slot_].stop, stream_);
}

}  // namespace tac

extern "C" int64_t tac_launch_count(void) { return tac::g_launches.load(); }

extern "C" int tac_profile_enable(int on) {
  tac::g_profile.store(on ? 1 : 0);
  return TAC_OK;
}

extern "C" int tac_profile_read(double* ms_by_kind, int64_t* launches_by_kind) {
  using namespace tac;
  std::lock_guard<std::mutex> lock(g_probe_mutex);
  for (int k = 0; k < 4; ++k) {
    if (ms_by_kind) ms_by_kind[k] = 0.0;
    if (launches_by_kind) launches_by_kind[k] = 0;
  }
  for (ProbeRecord& r : g_probe_log) {
    float ms = 0.f;
    cudaEventSynchronize(r.stop);
    if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess && r.kind >= 0 && r.kind < 4) {
      if (ms_by_kind) ms_by_kind[r.kind] += ms;
      if (launches_by_kind) launches_by_kind[r.kind] += 1;
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  g_probe_log.clear();
  return TAC_OK;
}

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
