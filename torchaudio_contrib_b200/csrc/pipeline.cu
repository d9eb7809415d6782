// Pipeline drivers.
//
//  tac_melspec_f32      device pointers: STFT+|.|^p (stft.cu) -> frame-major power rows in an L2-sized
//                       workspace -> tensor-core filterbank + dB (melbank.cu), slice by slice so that the
//                       rows written by one kernel are read by the next while still L2 resident.
//  tac_pipeline_*       host pointers: H2D -> the same two kernels -> D2H in double-buffered slices on
//                       two streams, so copies in both directions overlap compute.
#include <stdlib.h>

#include <math.h>

#include "bandplan.cuh"
#include "stft_params.cuh"
#include "tac_common.cuh"

namespace tac {

// melbank.cu
int launch_melbank_tiles(const float* tiles_ws, int64_t n_rows, int tile_rows, int64_t g_base, int64_t frames, int n_bins,
                         const void* plan_dev, int n_bands, int to_db, float ref, float amin, float* out,
                         cudaStream_t stream);
int balanced_tile_rows(int64_t rows);

// rows per slice: a multiple of (SMs x 16 warps) so that the STFT kernel's last wave is full, sized so the
// slice (rows x kpad x 4 B) stays well inside the 126 MB L2 together with the input it is computed from.
static int64_t slice_rows(int kpad) {
  static int64_t override_rows = -1;
  if (override_rows < 0) {
    const char* e = getenv("TAC_MELSPEC_SLICE_ROWS");
    override_rows = e ? atoll(e) : 0;
  }
  if (override_rows > 0) return override_rows;
  const int64_t wave = (int64_t)sm_count() * 16;
  const int64_t budget = (int64_t)96 << 20;                  // bytes of power tiles per slice (L2 is 126 MB)
  int64_t waves = budget / (wave * kpad * 4);
  if (waves < 1) waves = 1;
  return waves * wave;
}

static int run_melspec(StftParams sp, float power, const void* plan_dev, int n_bands, int to_db, float ref, float amin,
                       float* workspace, int64_t workspace_bytes, float* out, cudaStream_t stream) {
  const int64_t total = sp.n_seq * sp.frames;
  if (total == 0) return TAC_OK;
  const int64_t row_bytes = (int64_t)sp.kpad * 4;
  int64_t rows = slice_rows(sp.kpad);
  if (rows > total) rows = total;
  if (power_tile_bytes(rows, sp.kpad) > workspace_bytes) rows = (workspace_bytes / (row_bytes * 128) - 1) * 128;   // whole tiles + slack
  TAC_REQUIRE(rows >= 1 && workspace, TAC_ERR_WORKSPACE, "melspec: workspace of %lld bytes cannot hold one row of %lld bytes",
              (long long)workspace_bytes, (long long)row_bytes);
  sp.out = workspace;
  sp.out_mode = OUT_POWER_ROWS;
  sp.power = power;
  sp.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
  for (int64_t g0 = 0; g0 < total; g0 += rows) {
    sp.g0 = g0;
    sp.g1 = (g0 + rows < total) ? g0 + rows : total;
    int rc = launch_stft(sp, stream);
    if (rc != TAC_OK) return rc;
    rc = launch_melbank_tiles(workspace, sp.g1 - sp.g0, balanced_tile_rows(sp.g1 - sp.g0), g0, sp.frames, sp.bins, plan_dev,
                              n_bands, to_db, ref, amin, out, stream);
    if (rc != TAC_OK) return rc;
  }
  return TAC_OK;
}

// TAC_MELSPEC_FUSED=0 keeps the host pipeline on the two-kernel path (A/B timing)
static bool fused_enabled() {
  const char* e = getenv("TAC_MELSPEC_FUSED");       // read per pipeline creation, not cached
  return !(e && atoi(e) == 0);
}

// Fused path: STFT + |.|^p + two-band filterbank [+ dB] in ONE kernel, the spectrum never leaves the SM
// (stft.cu, OUT_MEL_FUSED).  n_fft = 2048 and a plan that carries a band plan (tac_fbplan_band_handle != 0).
static int run_melspec_banded(StftParams sp, float power, const void* plan_dev, int64_t band_handle, int n_bands, int to_db,
                              float ref, float amin, float* out, int frame_major, cudaStream_t stream,
                              float* const* peer_out = nullptr, int n_peers = 0, int64_t peer_seq0 = 0, int multicast = 0) {
  TAC_REQUIRE((reinterpret_cast<uintptr_t>(plan_dev) & 15) == 0, TAC_ERR_INVALID, "melspec_banded: plan must be 16-byte aligned");
  if (sp.n_fft != 2048) {
    // n_fft = 256 / 512 / 1024 / 4096: the warp kernels with the range-plan epilogue (stft_multi.cu); handle = offset | bytes << 32
    TAC_REQUIRE(sp.n_fft == 256 || sp.n_fft == 512 || sp.n_fft == 1024 || sp.n_fft == 4096, TAC_ERR_UNSUPPORTED,
                "melspec_banded: no one-kernel path for n_fft = %d", sp.n_fft);
    TAC_REQUIRE(n_peers == 0, TAC_ERR_UNSUPPORTED, "melspec_banded_peers: n_fft = 2048 only");
    const int64_t roff = band_handle & 0xffffffffLL, rbytes = band_handle >> 32;
    TAC_REQUIRE(roff > 0 && (roff & 15) == 0 && rbytes > 32 && (rbytes & 15) == 0, TAC_ERR_INVALID, "melspec_banded: bad range handle");
    if (sp.n_seq * sp.frames == 0) return TAC_OK;
    sp.out = out;
    sp.out_mode = OUT_MEL_RANGE;
    sp.power = power;
    sp.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
    sp.band_plan = static_cast<const unsigned char*>(plan_dev) + roff;
    sp.band_cmax = (int)((rbytes - 32) / 16);                          // uint4 count of meta + weights
    sp.n_bands = n_bands;
    sp.n_bands_pad = (n_bands + 31) / 32 * 32;
    sp.to_db = to_db ? 1 : 0;
    sp.amin = amin;
    sp.log10_ref = log10f(ref);
    if (frame_major) {
      sp.out_seq_stride = sp.frames * n_bands;
      sp.out_t_stride = n_bands;
      sp.out_band_stride = 1;
    } else {
      sp.out_seq_stride = (int64_t)n_bands * sp.frames;
      sp.out_t_stride = 1;
      sp.out_band_stride = sp.frames;
    }
    return launch_stft(sp, stream);
  }
  const int64_t off = band_handle & (((int64_t)1 << 48) - 1);
  const int cmax = (int)(band_handle >> 48);
  TAC_REQUIRE(off > 0 && (off & 15) == 0 && cmax >= 1 && cmax <= 16, TAC_ERR_INVALID, "melspec_banded: bad band handle");
  if (sp.n_seq * sp.frames == 0) return TAC_OK;
  sp.out = out;
  sp.out_mode = OUT_MEL_FUSED;
  sp.power = power;
  sp.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
  sp.band_plan = static_cast<const unsigned char*>(plan_dev) + off;
  sp.band_cmax = cmax;
  sp.n_bands = n_bands;
  sp.n_bands_pad = (n_bands + 31) / 32 * 32;
  sp.band_fast = (n_bands <= 128 && cmax <= 4) ? 1 : 0;             // the per-lane 4 x 4 list form follows `comb`
  sp.band_off_fast = (kBandOffComb + cmax * sp.n_bands_pad * 2 + 15) & ~15;
  sp.to_db = to_db ? 1 : 0;
  sp.amin = amin;
  sp.log10_ref = log10f(ref);
  if (frame_major) {                     // (n_seq, frames, n_bands): the memory order of the reference's transposed view
    sp.out_seq_stride = sp.frames * n_bands;
    sp.out_t_stride = n_bands;
    sp.out_band_stride = 1;
  } else {                               // (n_seq, n_bands, frames) contiguous
    sp.out_seq_stride = (int64_t)n_bands * sp.frames;
    sp.out_t_stride = 1;
    sp.out_band_stride = sp.frames;
  }
  if (n_peers > 0) {                     // every rank's full output takes this rank's frames (stores over NVLink)
    sp.out_mode = OUT_MEL_FUSED_PEERS;
    sp.out = nullptr;
    sp.n_peers = n_peers;
    sp.peer_multicast = multicast;
    sp.peer_seq0 = peer_seq0;
    for (int q = 0; q < n_peers; ++q) sp.peer_out[q] = peer_out[q];
  }
  return launch_stft(sp, stream);
}

}  // namespace tac

extern "C" int tac_melspec_banded_peers_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                                            const float* window, int n_fft, int hop, int center, int pad_mode, int normalized,
                                            float power, const void* plan_dev, int64_t band_handle, int n_bands, int to_db,
                                            float ref, float amin, float* const* peer_out, int n_peers, int64_t seq_offset,
                                            int frame_major, void* stream) {
  using namespace tac;
  StftParams sp;
  const int rc = fill_stft_params(sp, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, 1);
  if (rc != TAC_OK) return rc;
  TAC_REQUIRE(plan_dev && n_bands > 0, TAC_ERR_INVALID, "melspec_banded_peers: missing filterbank plan");
  TAC_REQUIRE(peer_out && n_peers >= 1 && n_peers <= kMaxPeers, TAC_ERR_INVALID,
              "melspec_banded_peers: %d output buffers given, 1..%d supported (the GPUs of one box)", n_peers, kMaxPeers);
  TAC_REQUIRE(seq_offset >= 0, TAC_ERR_INVALID, "melspec_banded_peers: negative sequence offset");
  for (int q = 0; q < n_peers; ++q)
    TAC_REQUIRE(peer_out[q] || sp.g1 == 0, TAC_ERR_INVALID, "melspec_banded_peers: null output pointer for rank %d", q);
  return run_melspec_banded(sp, power, plan_dev, band_handle, n_bands, to_db, ref, amin, nullptr, frame_major, as_stream(stream),
                            peer_out, n_peers, seq_offset);
}

extern "C" int tac_melspec_banded_mc_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride, const float* window,
                                         int n_fft, int hop, int center, int pad_mode, int normalized, float power,
                                         const void* plan_dev, int64_t band_handle, int n_bands, int to_db, float ref, float amin,
                                         float* mc_out, int64_t seq_offset, int frame_major, void* stream) {
  using namespace tac;
  StftParams sp;
  const int rc = fill_stft_params(sp, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, 1);
  if (rc != TAC_OK) return rc;
  TAC_REQUIRE(plan_dev && n_bands > 0, TAC_ERR_INVALID, "melspec_banded_mc: missing filterbank plan");
  TAC_REQUIRE(mc_out || sp.g1 == 0, TAC_ERR_INVALID, "melspec_banded_mc: null multicast output pointer");
  TAC_REQUIRE(seq_offset >= 0, TAC_ERR_INVALID, "melspec_banded_mc: negative sequence offset");
  float* const one[1] = {mc_out};
  return run_melspec_banded(sp, power, plan_dev, band_handle, n_bands, to_db, ref, amin, nullptr, frame_major, as_stream(stream),
                            one, 1, seq_offset, 1);
}

extern "C" int tac_melspec_banded_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride, const float* window,
                                      int n_fft, int hop, int center, int pad_mode, int normalized, float power,
                                      const void* plan_dev, int64_t band_handle, int n_bands, int to_db, float ref, float amin,
                                      float* out, int frame_major, void* stream) {
  using namespace tac;
  StftParams sp;
  const int rc = fill_stft_params(sp, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, 1);
  if (rc != TAC_OK) return rc;
  TAC_REQUIRE(plan_dev && n_bands > 0, TAC_ERR_INVALID, "melspec_banded: missing filterbank plan");
  TAC_REQUIRE(out || sp.g1 == 0, TAC_ERR_INVALID, "melspec_banded: null output pointer");
  return run_melspec_banded(sp, power, plan_dev, band_handle, n_bands, to_db, ref, amin, out, frame_major, as_stream(stream));
}

extern "C" int64_t tac_melspec_workspace_bytes(int64_t n_seq, int64_t n_samples, int n_fft, int hop, int center) {
  using namespace tac;
  if (n_fft <= 0) return 0;
  const int kpad = kpad_for_bins(n_fft / 2 + 1);
  const int64_t total = n_seq * tac_stft_num_frames(n_samples, n_fft, hop, center);
  int64_t rows = slice_rows(kpad);
  if (rows > total) rows = total;
  if (rows < 1) rows = 1;
  return power_tile_bytes(rows, kpad);
}

extern "C" int tac_melspec_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride, const float* window,
                               int n_fft, int hop, int center, int pad_mode, int normalized, float power,
                               const void* plan_dev, int n_bands, int to_db, float ref, float amin, void* workspace,
                               int64_t workspace_bytes, float* out, void* stream) {
  using namespace tac;
  StftParams sp;
  const int rc = fill_stft_params(sp, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, 1);
  if (rc != TAC_OK) return rc;
  TAC_REQUIRE(plan_dev && n_bands > 0, TAC_ERR_INVALID, "melspec: missing filterbank plan");
  TAC_REQUIRE(out || sp.g1 == 0, TAC_ERR_INVALID, "melspec: null output pointer");
  return run_melspec(sp, power, plan_dev, n_bands, to_db, ref, amin, static_cast<float*>(workspace), workspace_bytes, out,
                     as_stream(stream));
}

// ---------------------------------------------------------------------------------------------
// host-buffer pipeline
// ---------------------------------------------------------------------------------------------
// Host pipeline: three streams -- one carries every H2D copy, in order and back to back (PCIe is the bottleneck:
// measured 55 GB/s for one large pinned copy, less when copies of different streams interleave), one the kernels, one
// the D2H copies -- and kHostSlots buffer sets handed round between them with events:
//   H2D(i) waits until the kernel that last read slot's input is done;  kernel(i) waits for H2D(i) and for the D2H that
//   last read the slot's output;  D2H(i) waits for kernel(i).
// (The first version gave every slot its own stream running H2D -> kernel -> D2H; its H2D copies interleaved and the
// step took 0.87 ms for 41 MB in, against 0.74 ms for the bare copy.)
constexpr int kHostSlots = 8;

struct tac_pipeline {
  tac_pipeline_config cfg;
  int device;
  float* d_window;
  void* d_plan;
  int64_t band_handle;                  // non-zero: the one-kernel path applies (n_fft = 2048, two-band matrix)
  cudaStream_t s_in, s_in2, s_run, s_out;     // s_in2: second H2D stream, TAC_HOST_H2D_STREAMS=2 alternates the copies (measured
                                              // slower: 0.905 against 0.859 ms per config-2 step, profiles/r02_e2e_h2d_streams.txt)
  cudaEvent_t ev_in[kHostSlots], ev_run[kHostSlots], ev_out[kHostSlots];
  float* d_x[kHostSlots];
  float* d_out[kHostSlots];
  float* d_ws[kHostSlots];
  int64_t cap_x[kHostSlots], cap_out[kHostSlots], cap_ws[kHostSlots];       // bytes each slot's buffers hold
};

// Grow one device buffer; on failure the slot is left empty with capacity 0, so the next call allocates it again
// (capacities used to be per pipeline: a failed cudaMalloc left a null slot that the next call at the old size used).
static int grow_buffer(float** buf, int64_t* cap, int64_t bytes) {
  using namespace tac;
  if (bytes <= *cap) return TAC_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  TAC_CUDA_OK(cudaMalloc(buf, (size_t)bytes));
  *cap = bytes;
  return TAC_OK;
}

static int pipeline_reserve(tac_pipeline* p, int64_t x_bytes, int64_t out_bytes, int64_t ws_bytes) {
  using namespace tac;
  for (int i = 0; i < kHostSlots; ++i) {
    int rc = grow_buffer(&p->d_x[i], &p->cap_x[i], x_bytes);
    if (rc == TAC_OK) rc = grow_buffer(&p->d_out[i], &p->cap_out[i], out_bytes);
    if (rc == TAC_OK) rc = grow_buffer(&p->d_ws[i], &p->cap_ws[i], ws_bytes);
    if (rc != TAC_OK) return rc;
  }
  return TAC_OK;
}

extern "C" int tac_pipeline_destroy(tac_pipeline* p);

// everything that can fail after the handle exists; the caller destroys the half-built handle on any error
static int pipeline_build(tac_pipeline* p, const tac_pipeline_config* cfg, const float* window_host, const float* fb_host) {
  using namespace tac;
  TAC_CUDA_OK(cudaGetDevice(&p->device));
  TAC_CUDA_OK(cudaMalloc(&p->d_window, sizeof(float) * cfg->n_fft));
  TAC_CUDA_OK(cudaMemcpy(p->d_window, window_host, sizeof(float) * cfg->n_fft, cudaMemcpyHostToDevice));
  if (cfg->n_bands > 0) {
    const int64_t cap = tac_fbplan_bytes(cfg->n_bins, cfg->n_bands);
    void* host = malloc((size_t)cap);
    TAC_REQUIRE(host, TAC_ERR_INVALID, "pipeline_create: out of host memory");
    int64_t used = 0;
    int rc = tac_fbplan_build_host(fb_host, cfg->n_bins, cfg->n_bands, host, cap, &used);
    if (rc == TAC_OK) {
      cudaError_t e = cudaMalloc(&p->d_plan, (size_t)used);
      if (e == cudaSuccess) e = cudaMemcpy(p->d_plan, host, (size_t)used, cudaMemcpyHostToDevice);
      if (fused_enabled()) p->band_handle = tac_fbplan_fused_handle(host, cfg->n_fft);
      if (e != cudaSuccess) rc = fail(TAC_ERR_CUDA, "pipeline_create: plan upload failed: %s", cudaGetErrorString(e));
    }
    free(host);
    if (rc != TAC_OK) return rc;
  }
  TAC_CUDA_OK(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  TAC_CUDA_OK(cudaStreamCreateWithFlags(&p->s_in2, cudaStreamNonBlocking));
  TAC_CUDA_OK(cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking));
  TAC_CUDA_OK(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < kHostSlots; ++i) {
    TAC_CUDA_OK(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
    TAC_CUDA_OK(cudaEventCreateWithFlags(&p->ev_run[i], cudaEventDisableTiming));
    TAC_CUDA_OK(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
  }
  return TAC_OK;
}

extern "C" int tac_pipeline_create(const tac_pipeline_config* cfg, const float* window_host, const float* fb_host,
                                   tac_pipeline** out) {
  using namespace tac;
  TAC_REQUIRE(cfg && window_host && out, TAC_ERR_INVALID, "pipeline_create: null pointer");
  *out = nullptr;
  TAC_REQUIRE(cfg->n_fft >= 2 && cfg->n_fft <= 8192, TAC_ERR_UNSUPPORTED, "pipeline_create: n_fft=%d outside [2, 8192]", cfg->n_fft);
  TAC_REQUIRE(cfg->n_bands == 0 || (fb_host && cfg->n_bins == cfg->n_fft / 2 + 1), TAC_ERR_INVALID,
              "pipeline_create: filterbank must have n_fft/2+1 = %d rows (got %d)", cfg->n_fft / 2 + 1, cfg->n_bins);
  tac_pipeline* p = static_cast<tac_pipeline*>(calloc(1, sizeof(tac_pipeline)));
  TAC_REQUIRE(p, TAC_ERR_INVALID, "pipeline_create: out of host memory");
  p->cfg = *cfg;
  const int rc = pipeline_build(p, cfg, window_host, fb_host);
  if (rc != TAC_OK) {
    char saved[512];                                         // destroy must not overwrite the reason
    strncpy(saved, last_error_buffer(), sizeof(saved) - 1);
    saved[sizeof(saved) - 1] = 0;
    tac_pipeline_destroy(p);
    strncpy(last_error_buffer(), saved, 511);
    return rc;
  }
  *out = p;
  return TAC_OK;
}

extern "C" int tac_pipeline_run_host(tac_pipeline* p, const float* x_host, int64_t n_seq, int64_t n_samples,
                                     float* out_host) {
  using namespace tac;
  TAC_REQUIRE(p && x_host && out_host, TAC_ERR_INVALID, "pipeline_run: null pointer");
  TAC_REQUIRE(n_seq >= 0 && n_samples > 0, TAC_ERR_INVALID, "pipeline_run: bad shape");
  if (n_seq == 0) return TAC_OK;
  const tac_pipeline_config& c = p->cfg;
  TAC_CUDA_OK(cudaSetDevice(p->device));
  const int64_t frames = tac_stft_num_frames(n_samples, c.n_fft, c.hop, c.center);
  const int out_rows = c.n_bands > 0 ? c.n_bands : c.n_fft / 2 + 1;
  // Slice size: small enough that the tail after the last H2D (that slice's kernel and D2H, which nothing overlaps) is
  // short, large enough that a slice's kernel is still efficient and the per-copy overhead stays small: about one
  // eighth of the batch, between 2 and 8 MB of input (TAC_HOST_SLICE_MB overrides); the last slice is halved again and
  // again down to ~1 MB.
  static int64_t slice_override = -1;
  if (slice_override < 0) {
    const char* e = getenv("TAC_HOST_SLICE_MB");
    slice_override = (e && atoi(e) > 0) ? ((int64_t)atoi(e) << 20) : 0;
  }
  int64_t slice_bytes = slice_override;
  if (slice_bytes == 0) {
    slice_bytes = (n_seq * n_samples * 4 + 7) / 8;
    if (slice_bytes < ((int64_t)2 << 20)) slice_bytes = (int64_t)2 << 20;
    if (slice_bytes > ((int64_t)8 << 20)) slice_bytes = (int64_t)8 << 20;
  }
  int64_t per = (slice_bytes + n_samples * 4 - 1) / (n_samples * 4);
  if (per < 1) per = 1;
  if (per > n_seq) per = n_seq;
  const int64_t x_bytes = per * n_samples * 4;
  const int64_t o_bytes = per * out_rows * frames * 4;
  const int64_t ws_bytes = (c.n_bands > 0 && !p->band_handle) ? tac_melspec_workspace_bytes(per, n_samples, c.n_fft, c.hop, c.center) : 0;
  int rc = pipeline_reserve(p, x_bytes, o_bytes, ws_bytes);
  if (rc != TAC_OK) return rc;
  const int64_t min_tail = ((int64_t)1 << 20) / (n_samples * 4) + 1;      // sequences in ~1 MB
  int64_t used[kHostSlots] = {0};                                         // slices a slot has carried in this call
  static int h2d_streams = -1;
  if (h2d_streams < 0) {
    const char* e = getenv("TAC_HOST_H2D_STREAMS");
    h2d_streams = (e && atoi(e) == 2) ? 2 : 1;
  }
  int64_t slice_no = 0;
  int slot = 0;
  for (int64_t s0 = 0, ns = 0; s0 < n_seq; s0 += ns, slot = (slot + 1) % kHostSlots) {
    const int64_t left = n_seq - s0;
    ns = left < per ? left : per;
    if (left <= per && left > 2 * min_tail) ns = (left + 1) / 2;
    cudaStream_t sin = (h2d_streams == 2 && (slice_no & 1)) ? p->s_in2 : p->s_in;
    ++slice_no;
    if (used[slot]) TAC_CUDA_OK(cudaStreamWaitEvent(sin, p->ev_run[slot], 0));      // slot's input consumed
    TAC_CUDA_OK(cudaMemcpyAsync(p->d_x[slot], x_host + s0 * n_samples, (size_t)(ns * n_samples * 4), cudaMemcpyHostToDevice, sin));
    TAC_CUDA_OK(cudaEventRecord(p->ev_in[slot], sin));
    TAC_CUDA_OK(cudaStreamWaitEvent(p->s_run, p->ev_in[slot], 0));
    if (used[slot]) TAC_CUDA_OK(cudaStreamWaitEvent(p->s_run, p->ev_out[slot], 0));     // slot's output copied out
    cudaStream_t st = p->s_run;
    StftParams sp;
    rc = fill_stft_params(sp, p->d_x[slot], ns, n_samples, n_samples, p->d_window, c.n_fft, c.hop, c.center, c.pad_mode,
                          c.normalized, 1);
    if (rc != TAC_OK) return rc;
    if (c.n_bands > 0 && p->band_handle) {
      rc = run_melspec_banded(sp, c.power, p->d_plan, p->band_handle, c.n_bands, c.to_db, c.ref, c.amin, p->d_out[slot], 0, st);
    } else if (c.n_bands > 0) {
      rc = run_melspec(sp, c.power, p->d_plan, c.n_bands, c.to_db, c.ref, c.amin, p->d_ws[slot], p->cap_ws[slot], p->d_out[slot], st);
    } else {
      sp.out = p->d_out[slot];
      sp.out_mode = OUT_POWER_PUBLIC;
      sp.power = c.power;
      sp.power_mode = c.power == 2.0f ? 2 : (c.power == 1.0f ? 1 : 0);
      rc = launch_stft(sp, st);
    }
    if (rc != TAC_OK) return rc;
    TAC_CUDA_OK(cudaEventRecord(p->ev_run[slot], p->s_run));
    TAC_CUDA_OK(cudaStreamWaitEvent(p->s_out, p->ev_run[slot], 0));
    TAC_CUDA_OK(cudaMemcpyAsync(out_host + s0 * out_rows * frames, p->d_out[slot], (size_t)(ns * out_rows * frames * 4),
                                cudaMemcpyDeviceToHost, p->s_out));
    TAC_CUDA_OK(cudaEventRecord(p->ev_out[slot], p->s_out));
    ++used[slot];
  }
  TAC_CUDA_OK(cudaStreamSynchronize(p->s_out));          // every D2H follows its kernel, every kernel its H2D
  TAC_CUDA_OK(cudaStreamSynchronize(p->s_run));
  TAC_CUDA_OK(cudaStreamSynchronize(p->s_in));
  TAC_CUDA_OK(cudaStreamSynchronize(p->s_in2));
  return TAC_OK;
}

extern "C" int tac_pipeline_destroy(tac_pipeline* p) {
  using namespace tac;
  if (!p) return TAC_OK;
  cudaSetDevice(p->device);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_in2) cudaStreamDestroy(p->s_in2);
  if (p->s_run) cudaStreamDestroy(p->s_run);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  for (int i = 0; i < kHostSlots; ++i) {
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_run[i]) cudaEventDestroy(p->ev_run[i]);
    if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
    if (p->d_x[i]) cudaFree(p->d_x[i]);
    if (p->d_out[i]) cudaFree(p->d_out[i]);
    if (p->d_ws[i]) cudaFree(p->d_ws[i]);
  }
  if (p->d_window) cudaFree(p->d_window);
  if (p->d_plan) cudaFree(p->d_plan);
  free(p);
  return TAC_OK;
}
