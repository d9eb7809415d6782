// NVSwitch multicast memory for the kernel-side gather: every rank's full output buffer is one
// replica of a CUDA multicast object, so ONE store to the multicast address lands in the buffers of all GPUs of the box
// (the switch replicates it) instead of one unicast store per peer (peers.cu: 7x NVLink egress at 8 GPUs).
//
//   tac_mc_supported     CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED of the current device
//   tac_mc_create        rank 0: cuMulticastCreate + export as a POSIX file descriptor (the host side passes the
//                        descriptor to the other processes over a unix socket, SCM_RIGHTS)
//   tac_mc_import        other ranks: cuMemImportFromShareableHandle
//   tac_mc_add_device    every rank adds its device (all must have done so before anyone binds: host barrier)
//   tac_mc_bind          cuMemCreate of this rank's replica, cuMulticastBindMem, and two mappings: the replica itself
//                        (local pointer: what this rank reads) and the multicast object (what the kernels store to)
//   tac_mc_barrier       stream-ordered flag barrier: one multimem store of the epoch into flags[rank] of every replica,
//                        then wait until this rank's replica shows every rank's epoch (bounded by a timeout)
//   tac_mc_free          unmap, release
// Layout of a replica: TAC_PEER_HEADER_BYTES of flags (as peers.cu), then the payload.
#include <cuda.h>

#include "tac_common.cuh"

namespace tac {

struct McObject {
  CUmemGenericAllocationHandle mc = 0, mem = 0;
  size_t size = 0, gran = 0;
  CUdeviceptr local_va = 0, mc_va = 0;
  int dev = -1, n_devices = 0;
  bool have_mc = false, have_mem = false, bound = false, mapped_local = false, mapped_mc = false;
};

// Driver entry points are looked up through the runtime (cudaGetDriverEntryPoint), so the library has no link-time
// dependency on libcuda.so.1 and still loads on a machine without a driver (the CPU-side checks of tests/).
#define TAC_DRIVER_FUNCS(X)                                                                                          \
  X(cuGetErrorString) X(cuDeviceGet) X(cuDeviceGetAttribute) X(cuMulticastGetGranularity) X(cuMulticastCreate)         \
  X(cuMemExportToShareableHandle) X(cuMemImportFromShareableHandle) X(cuMulticastAddDevice) X(cuMemCreate)             \
  X(cuMulticastBindMem) X(cuMemAddressReserve) X(cuMemMap) X(cuMemSetAccess) X(cuMemUnmap) X(cuMemAddressFree)         \
  X(cuMulticastUnbind) X(cuMemRelease)
struct DriverApi {
#define X(name) decltype(&::name) name = nullptr;
  TAC_DRIVER_FUNCS(X)
#undef X
  bool ok = false;
  const char* missing = nullptr;
};
static const DriverApi& driver() {
  static DriverApi api = [] {
    DriverApi a;
    a.ok = true;
#define X(name)                                                                                          \
  {                                                                                                      \
    void* fn = nullptr;                                                                                  \
    cudaDriverEntryPointQueryResult q;                                                                   \
    if (cudaGetDriverEntryPoint(#name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) { \
      a.ok = false;                                                                                      \
      if (!a.missing) a.missing = #name;                                                                 \
    }                                                                                                    \
    a.name = reinterpret_cast<decltype(&::name)>(fn);                                                    \
  }
    TAC_DRIVER_FUNCS(X)
#undef X
    cudaGetLastError();
    return a;
  }();
  return api;
}
#define TAC_NEED_DRIVER()                                                                                                \
  do {                                                                                                                   \
    if (!driver().ok) return fail(TAC_ERR_UNSUPPORTED, "multicast: driver entry point %s not available", driver().missing); \
  } while (0)

static int cu_fail(CUresult r, const char* what) {
  const char* s = nullptr;
  if (driver().cuGetErrorString) driver().cuGetErrorString(r, &s);
  return fail(TAC_ERR_CUDA, "%s: %s", what, s ? s : "unknown driver error");
}
#define TAC_CU_OK(expr)                                 \
  do {                                                  \
    CUresult r__ = (expr);                              \
    if (r__ != CUDA_SUCCESS) return cu_fail(r__, #expr); \
  } while (0)

static CUmulticastObjectProp mc_prop(int n_devices, size_t size) {
  CUmulticastObjectProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.numDevices = (unsigned)n_devices;
  prop.size = size;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

static int mc_begin(int64_t bytes, int n_devices, McObject** out) {
  TAC_REQUIRE(bytes >= 0 && n_devices >= 1 && n_devices <= 16 && out, TAC_ERR_INVALID, "multicast: bad arguments");
  TAC_CUDA_OK(cudaFree(nullptr));
  TAC_NEED_DRIVER();                       // make sure the primary context exists and is current
  McObject* o = new McObject();
  TAC_CUDA_OK(cudaGetDevice(&o->dev));
  o->n_devices = n_devices;
  CUmulticastObjectProp prop = mc_prop(n_devices, 0);
  size_t gran = 0;
  CUresult r = driver().cuMulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS || gran == 0) {
    delete o;
    return cu_fail(r, "cuMulticastGetGranularity");
  }
  o->gran = gran;
  const size_t want = (size_t)bytes + TAC_PEER_HEADER_BYTES;
  o->size = (want + gran - 1) / gran * gran;
  *out = o;
  return TAC_OK;
}

__global__ void mc_barrier_kernel(uint32_t* mc_flags, const uint32_t* local_flags, uint32_t* timed_out, int n, int rank, uint32_t epoch,
                                  long long timeout_cycles) {
  const int q = threadIdx.x;
  if (q == 0) {
    __threadfence_system();                                 // this rank's earlier multicast stores before the flag
    asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(mc_flags + rank), "r"(epoch) : "memory");
  }
  if (q >= n) return;
  const long long t0 = clock64();
  for (;;) {
    uint32_t seen;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local_flags + q) : "memory");
    if ((int32_t)(seen - epoch) >= 0) break;
    if (clock64() - t0 > timeout_cycles) {
      *timed_out = 1u;
      break;
    }
    __nanosleep(64);
  }
}

}  // namespace tac

extern "C" int tac_mc_supported(int* supported) {
  using namespace tac;
  TAC_REQUIRE(supported, TAC_ERR_INVALID, "mc_supported: null argument");
  *supported = 0;
  TAC_CUDA_OK(cudaFree(nullptr));
  if (!driver().ok) return TAC_OK;
  int dev = 0;
  TAC_CUDA_OK(cudaGetDevice(&dev));
  CUdevice cu;
  TAC_CU_OK(driver().cuDeviceGet(&cu, dev));
  int v = 0;
  TAC_CU_OK(driver().cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cu));
  *supported = v;
  return TAC_OK;
}

extern "C" int tac_mc_create(int64_t bytes, int n_devices, void** obj, int* fd_out) {
  using namespace tac;
  TAC_REQUIRE(obj && fd_out, TAC_ERR_INVALID, "mc_create: null argument");
  TAC_REQUIRE(n_devices >= 2, TAC_ERR_UNSUPPORTED, "mc_create: a multicast object needs at least two devices (the driver refuses one)");
  McObject* o = nullptr;
  int rc = mc_begin(bytes, n_devices, &o);
  if (rc != TAC_OK) return rc;
  CUmulticastObjectProp prop = mc_prop(n_devices, o->size);
  CUresult r = driver().cuMulticastCreate(&o->mc, &prop);
  if (r != CUDA_SUCCESS) {
    delete o;
    return cu_fail(r, "cuMulticastCreate");
  }
  o->have_mc = true;
  int fd = -1;
  r = driver().cuMemExportToShareableHandle(&fd, o->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    driver().cuMemRelease(o->mc);
    delete o;
    return cu_fail(r, "cuMemExportToShareableHandle");
  }
  *fd_out = fd;
  *obj = o;
  return TAC_OK;
}

extern "C" int tac_mc_import(int fd, int64_t bytes, int n_devices, void** obj) {
  using namespace tac;
  TAC_REQUIRE(obj && fd >= 0, TAC_ERR_INVALID, "mc_import: bad arguments");
  McObject* o = nullptr;
  int rc = mc_begin(bytes, n_devices, &o);
  if (rc != TAC_OK) return rc;
  CUresult r = driver().cuMemImportFromShareableHandle(&o->mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) {
    delete o;
    return cu_fail(r, "cuMemImportFromShareableHandle");
  }
  o->have_mc = true;
  *obj = o;
  return TAC_OK;
}

extern "C" int tac_mc_add_device(void* obj) {
  using namespace tac;
  McObject* o = static_cast<McObject*>(obj);
  TAC_REQUIRE(o && o->have_mc, TAC_ERR_INVALID, "mc_add_device: no multicast object");
  CUdevice cu;
  TAC_CU_OK(driver().cuDeviceGet(&cu, o->dev));
  TAC_CU_OK(driver().cuMulticastAddDevice(o->mc, cu));
  return TAC_OK;
}

extern "C" int tac_mc_bind(void* obj, void** local_ptr, void** mc_ptr) {
  using namespace tac;
  McObject* o = static_cast<McObject*>(obj);
  TAC_REQUIRE(o && o->have_mc && local_ptr && mc_ptr, TAC_ERR_INVALID, "mc_bind: bad arguments");
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = o->dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  TAC_CU_OK(driver().cuMemCreate(&o->mem, o->size, &ap, 0));
  o->have_mem = true;
  TAC_CU_OK(driver().cuMulticastBindMem(o->mc, 0, o->mem, 0, o->size, 0));
  o->bound = true;
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof(ad));
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = o->dev;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  TAC_CU_OK(driver().cuMemAddressReserve(&o->local_va, o->size, o->gran, 0, 0));
  TAC_CU_OK(driver().cuMemMap(o->local_va, o->size, 0, o->mem, 0));
  o->mapped_local = true;
  TAC_CU_OK(driver().cuMemSetAccess(o->local_va, o->size, &ad, 1));
  TAC_CU_OK(driver().cuMemAddressReserve(&o->mc_va, o->size, o->gran, 0, 0));
  TAC_CU_OK(driver().cuMemMap(o->mc_va, o->size, 0, o->mc, 0));
  o->mapped_mc = true;
  TAC_CU_OK(driver().cuMemSetAccess(o->mc_va, o->size, &ad, 1));
  TAC_CUDA_OK(cudaMemset((void*)o->local_va, 0, TAC_PEER_HEADER_BYTES));
  TAC_CUDA_OK(cudaDeviceSynchronize());
  *local_ptr = (void*)o->local_va;
  *mc_ptr = (void*)o->mc_va;
  return TAC_OK;
}

extern "C" int tac_mc_barrier(void* obj, int rank, uint32_t epoch, double timeout_s, void* stream) {
  using namespace tac;
  McObject* o = static_cast<McObject*>(obj);
  TAC_REQUIRE(o && o->mapped_mc && o->mapped_local && rank >= 0 && rank < o->n_devices, TAC_ERR_INVALID,
              "mc_barrier: object not bound or bad rank %d", rank);
  if (!(timeout_s > 0.0) || timeout_s > 60.0) timeout_s = 60.0;
  const long long cycles = (long long)(timeout_s * 2.0e9);
  uint32_t* local = reinterpret_cast<uint32_t*>(o->local_va);
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  mc_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<uint32_t*>(o->mc_va), local, local + 16, o->n_devices, rank, epoch, cycles);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_mc_timed_out(void* obj, int* timed_out) {
  using namespace tac;
  McObject* o = static_cast<McObject*>(obj);
  TAC_REQUIRE(o && o->mapped_local && timed_out, TAC_ERR_INVALID, "mc_timed_out: bad arguments");
  uint32_t v = 0;
  TAC_CUDA_OK(cudaMemcpy(&v, reinterpret_cast<const uint32_t*>(o->local_va) + 16, sizeof(v), cudaMemcpyDeviceToHost));
  *timed_out = (int)v;
  return TAC_OK;
}

extern "C" int tac_mc_free(void* obj) {
  using namespace tac;
  McObject* o = static_cast<McObject*>(obj);
  if (!o) return TAC_OK;
  cudaDeviceSynchronize();
  if (o->mapped_mc) driver().cuMemUnmap(o->mc_va, o->size);
  if (o->mc_va) driver().cuMemAddressFree(o->mc_va, o->size);
  if (o->mapped_local) driver().cuMemUnmap(o->local_va, o->size);
  if (o->local_va) driver().cuMemAddressFree(o->local_va, o->size);
  if (o->bound) {
    CUdevice cu;
    if (driver().cuDeviceGet(&cu, o->dev) == CUDA_SUCCESS) driver().cuMulticastUnbind(o->mc, cu, 0, o->size);
  }
  if (o->have_mem) driver().cuMemRelease(o->mem);
  if (o->have_mc) driver().cuMemRelease(o->mc);
  delete o;
  return TAC_OK;
}
