// Peer memory for the batch-sharded path (BASELINE config 4): one process per GPU, every
// rank's output buffer mapped into every other rank's address space, so the mel kernel's epilogue can store a
// frame's bands into all of them over NVLink (stft.cu OUT_MEL_FUSED_PEERS) instead of a trailing all-gather.
//
//   tac_peer_alloc   cudaMalloc + cudaIpcGetMemHandle (the 64-byte handle travels through torch.distributed)
//   tac_peer_open    cudaIpcOpenMemHandle in another process of the same box (enables peer access lazily)
//   tac_peer_barrier one tiny kernel: publish "my stores of round `epoch` are done" to every rank's flag row,
//                    then wait until every rank has published the same to mine.  Stream-ordered after the mel
//                    kernel, so when it retires the local buffer holds every rank's frames of that round.
#include "tac_common.cuh"

namespace tac {

// flag block at the start of a peer allocation: flags[r] = last round rank r has completed towards this rank
struct PeerHeader {
  uint32_t flags[16];
  uint32_t timed_out;      // set by a barrier that gave up waiting (a dead peer must not hang the GPU)
  uint32_t pad[15];
};
static_assert(sizeof(PeerHeader) == TAC_PEER_HEADER_BYTES, "header size is part of the ABI");

struct BarrierArgs {
  uint32_t* flags[16];     // flag rows of all ranks (peer-mapped), index = rank
};

__global__ void peer_barrier_kernel(BarrierArgs a, int n_peers, int rank, uint32_t epoch, long long timeout_cycles) {
  const int q = threadIdx.x;
  if (q >= n_peers) return;
  __threadfence_system();                                  // this rank's earlier stores (previous kernel) before the flag
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flags[q] + rank), "r"(epoch) : "memory");
  const uint32_t* mine = a.flags[rank] + q;                // rank q's progress towards me
  const long long t0 = clock64();
  for (;;) {
    uint32_t seen;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
    if ((int32_t)(seen - epoch) >= 0) break;
    if (clock64() - t0 > timeout_cycles) {
      reinterpret_cast<PeerHeader*>(a.flags[rank])->timed_out = 1u;
      break;
    }
    __nanosleep(64);
  }
}

}  // namespace tac

extern "C" int tac_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle_out) {
  using namespace tac;
  TAC_REQUIRE(bytes >= 0 && dev_ptr && handle_out, TAC_ERR_INVALID, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == TAC_PEER_HANDLE_BYTES, "handle size is part of the ABI");
  void* p = nullptr;
  TAC_CUDA_OK(cudaMalloc(&p, (size_t)bytes + TAC_PEER_HEADER_BYTES));
  cudaError_t err = cudaMemset(p, 0, TAC_PEER_HEADER_BYTES);
  cudaIpcMemHandle_t h;
  if (err == cudaSuccess) err = cudaIpcGetMemHandle(&h, p);
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    cudaFree(p);
    return fail(TAC_ERR_CUDA, "peer_alloc: %s", cudaGetErrorString(err));
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return TAC_OK;
}

extern "C" int tac_peer_open(const unsigned char* handle, void** dev_ptr) {
  using namespace tac;
  TAC_REQUIRE(handle && dev_ptr, TAC_ERR_INVALID, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  TAC_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return TAC_OK;
}

extern "C" int tac_peer_close(void* dev_ptr) {
  using namespace tac;
  if (dev_ptr) TAC_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return TAC_OK;
}

extern "C" int tac_peer_free(void* dev_ptr) {
  using namespace tac;
  if (dev_ptr) TAC_CUDA_OK(cudaFree(dev_ptr));
  return TAC_OK;
}

extern "C" int tac_peer_barrier(void* const* peer_base, int n_peers, int rank, uint32_t epoch, double timeout_s, void* stream) {
  using namespace tac;
  TAC_REQUIRE(peer_base && n_peers >= 1 && n_peers <= 16 && rank >= 0 && rank < n_peers, TAC_ERR_INVALID,
              "peer_barrier: bad rank %d of %d", rank, n_peers);
  BarrierArgs a;
  for (int q = 0; q < 16; ++q) a.flags[q] = nullptr;
  for (int q = 0; q < n_peers; ++q) {
    TAC_REQUIRE(peer_base[q], TAC_ERR_INVALID, "peer_barrier: null base pointer for rank %d", q);
    a.flags[q] = static_cast<PeerHeader*>(peer_base[q])->flags;
  }
  if (!(timeout_s > 0.0) || timeout_s > 60.0) timeout_s = 60.0;      // bounded: never spin for ever on a dead peer
  const long long cycles = (long long)(timeout_s * 2.0e9);
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(a, n_peers, rank, epoch, cycles);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_peer_timed_out(const void* own_base, int* timed_out) {
  using namespace tac;
  TAC_REQUIRE(own_base && timed_out, TAC_ERR_INVALID, "peer_timed_out: bad arguments");
  uint32_t v = 0;
  TAC_CUDA_OK(cudaMemcpy(&v, &static_cast<const PeerHeader*>(own_base)->timed_out, sizeof(v), cudaMemcpyDeviceToHost));
  *timed_out = (int)v;
  return TAC_OK;
}
