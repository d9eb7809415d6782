/* tac_b200.h -- C ABI of the B200-native mel-spectrogram / mu-law path.
 *
 * The reference (keunwoochoi/torchaudio-contrib) has no FFI of its own: its operator
 * surface is the Python functions in torchaudio_contrib/functional.py, each of which hands
 * the arithmetic to a torch operator.  Every entry point below replaces one of those call
 * sites (cited as functional.py:LINE); INTEGRATION.md shows the ctypes stub that binds it.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / CUDA types in the signatures
 *     (`stream` is a cudaStream_t passed as void*; NULL = the legacy default stream);
 *   - "device" entry points take DEVICE pointers owned by the caller, enqueue work on
 *     `stream` and return immediately; nothing is allocated that the caller must free;
 *   - "host" entry points (tac_pipeline_*) take HOST pointers and do H2D, compute, D2H;
 *   - return value: TAC_OK or a negative TAC_ERR_*; tac_last_error() gives the message for
 *     the calling thread; nothing throws across the boundary;
 *   - fp32 everywhere; mu-law codes are int64 (functional.py:334 `.long()`).
 */
#ifndef TAC_B200_H_
#define TAC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAC_ABI_VERSION 1

enum {
  TAC_OK = 0,
  TAC_ERR_INVALID = -1,      /* bad argument (shape, size, null pointer)                  */
  TAC_ERR_UNSUPPORTED = -2,  /* valid for the reference, not implemented by these kernels */
  TAC_ERR_CUDA = -3,         /* CUDA runtime / driver error (message has the cuda string) */
  TAC_ERR_WORKSPACE = -4     /* workspace / plan buffer too small                         */
};

/* torch.nn.functional.pad modes accepted by torch.stft (functional.py:104) */
enum { TAC_PAD_REFLECT = 0, TAC_PAD_CONSTANT = 1, TAC_PAD_REPLICATE = 2, TAC_PAD_CIRCULAR = 3 };

int tac_version(void);
const char* tac_last_error(void);
/* Fills SM count and compute capability of the current device. */
int tac_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* frames = 1 + (T + 2*(n_fft/2)*center - n_fft) / hop   (torch.stft; tests/test_functional.py:14-15) */
int64_t tac_stft_num_frames(int64_t n_samples, int n_fft, int hop, int center);

/* ---- a1: stft (functional.py:48-113, call site torch.stft :99-107) ---------------------
 * x: (n_seq, n_samples) rows `seq_stride` floats apart.  window: n_fft floats, already
 * centre-padded from win_length (torch.stft semantics).  n_fft: any value in 2..8192 (powers of two
 * on the tuned kernels, everything else on the direct-DFT kernel and its adjoint).
 * out: (n_seq, bins, frames, 2) contiguous, bins = n_fft/2+1 (onesided) or n_fft. */
int tac_stft_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                 const float* window, int n_fft, int hop, int center, int pad_mode,
                 int normalized, int onesided, float* out, void* stream);

/* ---- a1+a2: Spectrogram = stft then complex_norm(power) (layers.py:294-304) -------------
 * out: (n_seq, bins, frames) contiguous. */
int tac_spectrogram_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                        const float* window, int n_fft, int hop, int center, int pad_mode,
                        int normalized, int onesided, float power, float* out, void* stream);

/* ---- a2: complex_norm (functional.py:116-128): z (n, 2) -> out (n) = |z|^power ---------- */
int tac_complex_norm_f32(const float* z, int64_t n, float power, float* out, void* stream);

/* ---- a5: amplitude_to_db (functional.py:277-296): 10*(log10(max(x^2, amin)) - log10(ref)) */
int tac_amplitude_to_db_f32(const float* x, int64_t n, float ref, float amin, float* out,
                            void* stream);

/* ---- N4: db_to_amplitude (functional.py:299-314): sqrt(10^(x/10 + log10(ref))) -------------- */
int tac_db_to_amplitude_f32(const float* x, int64_t n, float ref, float* out, void* stream);

/* ---- N4: angle / magphase (functional.py:187-201): z (n, 2) -> phase (n) = atan2(im, re) and,
 * when mag != NULL, mag (n) = |z|^power in the same pass. */
int tac_magphase_f32(const float* z, int64_t n, float power, float* mag, float* phase, void* stream);

/* ---- N2: phase_vocoder (functional.py:204-274) ---------------------------------------------
 * spec: (n_seq, n_bins, n_in, 2); out: (n_seq, n_bins, n_out, 2), n_out = ceil(n_in / rate).
 * idx0[j] = long(j*rate), idx1[j] = long(j*rate + 1), alpha[j] = frac(j*rate) for j < n_out are device
 * tables the caller builds with the reference's own expressions (:239-243, :250-253); an index >= n_in
 * reads the zero padding (:247-248).  advance: (n_bins) expected phase advance per bin.
 * Angles, the wrap, the running phase sum and sin/cos are evaluated in float64 in both variants. */
int tac_phase_vocoder_f32(const float* spec, int64_t n_seq, int n_bins, int64_t n_in,
                          const int32_t* idx0, const int32_t* idx1, const double* alpha,
                          const float* advance, int64_t n_out, float* out, void* stream);
int tac_phase_vocoder_f64(const double* spec, int64_t n_seq, int n_bins, int64_t n_in,
                          const int32_t* idx0, const int32_t* idx1, const double* alpha,
                          const double* advance, int64_t n_out, double* out, void* stream);
/* Backward of the phase vocoder w.r.t. the spectrogram (round 2).  range0 / range1: (n_in, 2) int32 tables, the
 * contiguous ranges [lo, hi) of output steps j with idx0[j] == i / idx1[j] == i (idx0 and idx1 are monotone);
 * workspace: 16 bytes per (row, output step); grad_spec: same shape as spec, overwritten.  float64 inside
 * like the forward kernel; a gather per input frame, no atomics (deterministic). */
int tac_phase_vocoder_backward_f32(const float* spec, const float* grad_out, int64_t n_seq, int n_bins,
                                   int64_t n_in, const int32_t* idx0, const int32_t* idx1,
                                   const double* alpha, const float* advance, int64_t n_out,
                                   const int32_t* range0, const int32_t* range1, void* workspace,
                                   int64_t workspace_bytes, float* grad_spec, void* stream);
int tac_phase_vocoder_backward_f64(const double* spec, const double* grad_out, int64_t n_seq, int n_bins,
                                   int64_t n_in, const int32_t* idx0, const int32_t* idx1,
                                   const double* alpha, const double* advance, int64_t n_out,
                                   const int32_t* range0, const int32_t* range1, void* workspace,
                                   int64_t workspace_bytes, double* grad_spec, void* stream);

/* ---- a3: apply_filterbank (functional.py:172-184) on tcgen05 tensor cores ---------------
 * The (n_bins, n_bands) row-major matrix is first turned into a "plan": per 32-bin K slice
 * the range of non-zero bands, plus the tf32 hi/lo split of that block laid out as the
 * 128B-swizzled K-major UMMA operand image.  Built on the host once per matrix, then copied
 * to the device by the caller (any 16-byte aligned device buffer). */
int64_t tac_fbplan_bytes(int n_bins, int n_bands);                 /* upper bound, bytes   */
int tac_fbplan_build_host(const float* fb_host, int n_bins, int n_bands,
                          void* plan_host, int64_t plan_capacity, int64_t* plan_bytes_used);

/* Non-zero when the plan also carries the matrix as a "band plan": every row has at most two
 * non-zeros, in adjacent columns, and they chain from band to band (any triangular filterbank,
 * e.g. create_mel_filter functional.py:131-169, n_bins = 1025).  The opaque handle is what
 * tac_melspec_banded_f32 takes; 0 means only the tensor-core path applies. */
int64_t tac_fbplan_band_handle(const void* plan_host);
/* Handle to pass to tac_melspec_banded_f32 for this plan at this fft length, 0 when the one-kernel path does not apply
 * (dense matrix, other fft length): n_fft = 2048 -> the band plan's handle, 256 / 512 / 1024 -> the range plan's. */
int64_t tac_fbplan_fused_handle(const void* plan_host, int n_fft);

/* spec: (n_seq, n_bins, frames) real, or (n_seq, n_bins, frames, 2) complex when is_complex.
 * Computes |.|^power first when is_complex (a2), contracts over bins with the plan (a3) and,
 * when to_db, applies amplitude_to_db(ref, amin) in the epilogue (a5).
 * out: (n_seq, n_bands, frames) contiguous. */
int tac_power_mel_f32(const float* spec, int is_complex, float power,
                      int64_t n_seq, int64_t frames, int n_bins,
                      const void* plan_dev, int n_bands,
                      int to_db, float ref, float amin, float* out, void* stream);

/* ---- a6: the Melspectrogram pipeline (layers.py:307-347 [+ AmplitudeToDb :350-381]) ------
 * stft -> |.|^power -> filterbank [-> dB] without materialising the complex spectrum: the
 * power spectrum goes through an L2-sized workspace in frame-major layout.
 * workspace: device buffer of at least tac_melspec_workspace_bytes(...) bytes. */
int64_t tac_melspec_workspace_bytes(int64_t n_seq, int64_t n_samples, int n_fft, int hop, int center);
int tac_melspec_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                    const float* window, int n_fft, int hop, int center, int pad_mode,
                    int normalized, float power,
                    const void* plan_dev, int n_bands, int to_db, float ref, float amin,
                    void* workspace, int64_t workspace_bytes, float* out, void* stream);

/* Same chain in ONE kernel for n_fft = 2048 and a plan with a band plan: each warp takes a
 * frame from its bulk-copied samples to its n_bands outputs; the spectrum never leaves the SM
 * and HBM sees the input once and the output once (SURVEY 2a "K3").  Replaces torch.stft
 * (functional.py:99-107), torch.norm/.pow (:126-128), the transpose-matmul-transpose (:183-184)
 * and the dB chain (:291-296).  out: (n_seq, n_bands, frames) contiguous, or when frame_major
 * (n_seq, frames, n_bands) -- the memory order behind the reference's transposed view. */
int tac_melspec_banded_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                           const float* window, int n_fft, int hop, int center, int pad_mode,
                           int normalized, float power,
                           const void* plan_dev, int64_t band_handle, int n_bands,
                           int to_db, float ref, float amin,
                           float* out, int frame_major, void* stream);

/* ---- row (e) / N3: batch split over the GPUs of one box, output gathered by the kernel ---
 * The reference has no multi-GPU path; SURVEY 8(e) shards the flattened batch
 * (functional.py:89-91) over one process per GPU and gathers the output.  Instead of a trailing
 * all-gather, tac_melspec_banded_peers_f32 stores every frame's bands straight into the output
 * buffers of ALL ranks (its own and the peer-mapped ones) over NVLink.
 *
 * Peer memory: tac_peer_alloc returns a device allocation of TAC_PEER_HEADER_BYTES + bytes and
 * its 64-byte IPC handle; the payload starts TAC_PEER_HEADER_BYTES after *dev_ptr.  Other
 * processes of the box map it with tac_peer_open (and unmap with tac_peer_close); the owner
 * releases it with tac_peer_free after the peers have closed it.
 * tac_peer_barrier (stream-ordered): tell every rank that this rank's stores of round `epoch`
 * (epochs must increase) are complete, then wait for the same from every rank; gives up after
 * timeout_s (<= 60) and records it -- tac_peer_timed_out reads that flag (synchronises). */
#define TAC_PEER_HANDLE_BYTES 64
#define TAC_PEER_HEADER_BYTES 128
int tac_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle_out /*[64]*/);
int tac_peer_open(const unsigned char* handle /*[64]*/, void** dev_ptr);
int tac_peer_close(void* dev_ptr);
int tac_peer_free(void* dev_ptr);
int tac_peer_barrier(void* const* peer_base /* host array [n_peers], index = rank */, int n_peers,
                     int rank, uint32_t epoch, double timeout_s, void* stream);
int tac_peer_timed_out(const void* own_base, int* timed_out);
/* As tac_melspec_banded_f32 for this rank's n_seq sequences; peer_out: host array of n_peers
 * (<= 8) device pointers, the payload of every rank's FULL output ((total_seq, n_bands, frames),
 * or (total_seq, frames, n_bands) when frame_major); this rank's rows start at seq_offset. */
int tac_melspec_banded_peers_f32(const float* x, int64_t n_seq, int64_t n_samples,
                                 int64_t seq_stride, const float* window, int n_fft, int hop,
                                 int center, int pad_mode, int normalized, float power,
                                 const void* plan_dev, int64_t band_handle, int n_bands,
                                 int to_db, float ref, float amin,
                                 float* const* peer_out, int n_peers, int64_t seq_offset,
                                 int frame_major, void* stream);

/* ---- N3 with NVSwitch multicast (csrc/multicast.cu) ------------------------------------------
 * The peer stores above cost one NVLink store per peer (7x egress on an 8-GPU box).  With a CUDA
 * multicast object every rank's full output buffer is one replica, and ONE store to the multicast
 * address is delivered to all of them by the switch.  Set-up (all ranks, in this order):
 *   rank 0: tac_mc_create -> obj + a POSIX file descriptor; the host passes the descriptor to the
 *           other processes (unix socket, SCM_RIGHTS); they call tac_mc_import;
 *   all:    tac_mc_add_device; host barrier (every device added); tac_mc_bind -> local_ptr (this
 *           rank's replica: TAC_PEER_HEADER_BYTES of flags, then the payload) and mc_ptr (same
 *           layout, multicast address: stores only);
 *   per step: tac_melspec_banded_mc_f32 (mc_out = mc_ptr + TAC_PEER_HEADER_BYTES), then
 *           tac_mc_barrier (stream-ordered; epoch increasing) -- afterwards local_ptr holds every
 *           rank's frames.  tac_mc_timed_out reads the give-up flag of a barrier (synchronises).
 * tac_mc_supported: 1 when the current device can take part (NVSwitch + driver support). */
int tac_mc_supported(int* supported);
int tac_mc_create(int64_t bytes, int n_devices, void** obj, int* fd_out);
int tac_mc_import(int fd, int64_t bytes, int n_devices, void** obj);
int tac_mc_add_device(void* obj);
int tac_mc_bind(void* obj, void** local_ptr, void** mc_ptr);
int tac_mc_barrier(void* obj, int rank, uint32_t epoch, double timeout_s, void* stream);
int tac_mc_timed_out(void* obj, int* timed_out);
int tac_mc_free(void* obj);
int tac_melspec_banded_mc_f32(const float* x, int64_t n_seq, int64_t n_samples,
                              int64_t seq_stride, const float* window, int n_fft, int hop,
                              int center, int pad_mode, int normalized, float power,
                              const void* plan_dev, int64_t band_handle, int n_bands,
                              int to_db, float ref, float amin,
                              float* mc_out, int64_t seq_offset, int frame_major, void* stream);

/* ---- N4 / H7: backward passes (the reference is differentiable w.r.t. the waveform) ---------
 * tac_stft_backward_f32: adjoint of tac_stft_f32.  grad_out: (n_seq, bins, frames, 2) contiguous;
 * grad_x: (n_seq, n_samples) contiguous, overwritten.
 * tac_spectrogram_backward_f32: adjoint of tac_spectrogram_f32 (stft then |.|^power); the spectrum
 * is recomputed from x.  grad_out: (n_seq, bins, frames) contiguous.
 * tac_filterbank_backward_f32: adjoint of the contraction of functional.py:183-184 w.r.t. its
 * input: grad_spec[s,k,t] = sum_m grad_y[s,m,t] fb[k,m]; grad_y is addressed through element
 * strides (so the reference's transposed view needs no copy), fb_dev: (n_bins, n_bands) row-major,
 * grad_spec: (n_seq, n_bins, frames) contiguous.
 * tac_amplitude_to_db_backward_f32 / tac_complex_norm_backward_f32: pointwise adjoints of
 * functional.py:291-296 and :126-128 (x, z: the forward inputs). */
/* workspace: device buffer of tac_stft_backward_workspace_bytes(...) bytes (the windowed frame
 * gradients, n_fft floats per frame; summed into grad_x without atomics, so the result is
 * deterministic). */
int64_t tac_stft_backward_workspace_bytes(int64_t n_seq, int64_t n_samples, int n_fft, int hop, int center);
int tac_stft_backward_f32(const float* grad_out, int64_t n_seq, int64_t n_samples,
                          const float* window, int n_fft, int hop, int center, int pad_mode,
                          int normalized, int onesided, float* grad_x,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* Gradient w.r.t. the window of tac_stft_f32 (round 2): run tac_stft_backward_f32 with a window of ONES first -- its
 * workspace then holds, per frame, scale * (gradient of the windowed frame before the window multiply) -- and hand
 * that workspace in as frames_ws; grad_window: n_fft floats (the centre-padded window), scratch: 256 * n_fft bytes.
 * Deterministic (two-step reduction in fixed order). */
int tac_window_grad_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                        const float* frames_ws, int n_fft, int hop, int center, int pad_mode,
                        float* grad_window, void* scratch, int64_t scratch_bytes, void* stream);
int tac_spectrogram_backward_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                                 const float* window, int n_fft, int hop, int center, int pad_mode,
                                 int normalized, int onesided, float power,
                                 const float* grad_out, float* grad_x,
                                 void* workspace, int64_t workspace_bytes, void* stream);
int tac_filterbank_backward_f32(const float* grad_y, int64_t stride_seq, int64_t stride_band,
                                int64_t stride_frame, const float* fb_dev, int64_t n_seq,
                                int64_t frames, int n_bins, int n_bands, float* grad_spec,
                                void* stream);
/* Backward of the Melspectrogram chain in one call: filterbank adjoint, stft + |.|^power adjoint
 * (spectrum recomputed from x) and overlap-add.  grad_y is addressed through element strides
 * like tac_filterbank_backward_f32; fb_dev: (n_fft/2+1, n_bands) row-major; workspace:
 * tac_melspec_backward_workspace_bytes(...) bytes.  n_fft = 2048 runs one warp per frame with the
 * forward kernel's register FFT; other sizes use the generic kernels. */
int64_t tac_melspec_backward_workspace_bytes(int64_t n_seq, int64_t n_samples, int n_fft, int hop, int center);
int tac_melspec_backward_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                             const float* window, int n_fft, int hop, int center, int pad_mode,
                             int normalized, float power, const float* fb_dev, int n_bands,
                             const float* grad_y, int64_t stride_seq, int64_t stride_band,
                             int64_t stride_frame, float* grad_x,
                             void* workspace, int64_t workspace_bytes, void* stream);
int tac_amplitude_to_db_backward_f32(const float* x, const float* grad_out, int64_t n, float amin,
                                     float* grad_x, void* stream);
int tac_complex_norm_backward_f32(const float* z, const float* grad_out, int64_t n, float power,
                                  float* grad_z, void* stream);

/* ---- a7: mu_law_encoding (functional.py:317-335) ----------------------------------------
 * The quantiser is evaluated as a table of decision levels: thresholds[j] is the smallest
 * float whose code is >= idx_min + j (thresholds[0] = -inf); |x| > x_limit or NaN gives
 * INT64_MIN (what the reference's float->int64 conversion yields after overflow to inf). */
int tac_mulaw_encode_f32_i64(const float* x, int64_t n, int n_quantize,
                             const float* thresholds_dev, int n_thresholds, int idx_min,
                             float x_limit, int64_t* out, void* stream);
/* The tables, on the host (no torch needed): thresholds (capacity floats; NULL only queries the
 * count), idx_min, x_limit for the encoder and decoded[n_quantize] for the decoder.  n_quantize =
 * 256 returns the levels SHIPPED with the library -- found with the reference's own fp32 torch CPU
 * chain, bit-exact with it for every float (*exact = 1); other values are bisected here with the
 * host's libm log1pf / expf (*exact = 0: equal to the reference wherever libm and torch agree). */
int tac_mulaw_tables_host(int n_quantize, float* thresholds, int capacity, int* n_thresholds,
                          int* idx_min, float* x_limit, float* decoded, int* exact);

/* ---- a8: mu_law_decoding (functional.py:338-354) ----------------------------------------
 * lut_dev[i] = decoded value of code i for 0 <= i < n_quantize; other codes are evaluated
 * with the closed form in the kernel. */
int tac_mulaw_decode_i64_f32(const int64_t* codes, int64_t n, int n_quantize,
                             const float* lut_dev, float* out, void* stream);
int tac_mulaw_decode_f32_f32(const float* codes, int64_t n, int n_quantize,
                             const float* lut_dev, float* out, void* stream);

/* ---- float64 operators (csrc/f64_path.cu) -------------------------------------------------
 * The reference computes in the dtype it is given (functional.py:99, :126-128, :183, :291-296,
 * :310-314, :331-334, :349-353); these are the same operators on double tensors, correct-first
 * kernels with double arithmetic throughout (forward only).  Layouts as the float entries;
 * tac_stft_f64: any n_fft in [2, 4096], window of n_fft doubles. */
int tac_stft_f64(const double* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                 const double* window, int n_fft, int hop, int center, int pad_mode,
                 int normalized, int onesided, double* out, void* stream);
int tac_complex_norm_f64(const double* z, int64_t n, double power, double* out, void* stream);
int tac_magphase_f64(const double* z, int64_t n, double power, double* mag /* may be NULL */,
                     double* phase, void* stream);
int tac_amplitude_to_db_f64(const double* x, int64_t n, double ref, double amin, double* out, void* stream);
int tac_db_to_amplitude_f64(const double* x, int64_t n, double ref, double* out, void* stream);
/* spec: (n_seq, n_bins, frames), fb_dev: (n_bins, n_bands) row-major, out: (n_seq, n_bands, frames) */
int tac_apply_filterbank_f64(const double* spec, const double* fb_dev, int64_t n_seq, int64_t frames,
                             int n_bins, int n_bands, double* out, void* stream);
int tac_mulaw_decode_i64_f64(const int64_t* codes, int64_t n, int n_quantize, const double* lut_dev,
                             double* out, void* stream);
int tac_mulaw_encode_f64_i64(const double* x, int64_t n, int n_quantize, int64_t* out, void* stream);

/* ---- harmonic-percussive separation (beta_hpss.py:37-129; a beta module the reference does not export) ----
 * mag: (n_seq, n_freq, n_time) magnitudes; median filters of `kernel_size` (odd, 3..63) along frequency
 * (percussive) and time (harmonic) with reflect padding, ^power, soft masks (eps 1e-6) or hard masks
 * (1.0 / 0.0); out_harm / out_perc = mag * mask (may be NULL with mask_only).  One kernel; medians are
 * selections, so the result is bit-exact with the reference for power 1 and 2. */
int tac_hpss_f32(const float* mag, int64_t n_seq, int n_freq, int n_time, int kernel_size, float power,
                 int hard, int mask_only, float* out_harm, float* out_perc, float* mask_harm,
                 float* mask_perc, void* stream);

/* ---- host-buffer plugin surface (what a reference-side caller with CPU tensors binds) ---
 * A pipeline owns its device staging buffers, plan, streams and events. */
typedef struct tac_pipeline tac_pipeline;

typedef struct tac_pipeline_config {
  int n_fft, hop, center, pad_mode, normalized;
  float power;           /* ComplexNorm power; Melspectrogram uses 2.0 (layers.py:346)      */
  int n_bins, n_bands;   /* filterbank shape; n_bands == 0: Spectrogram only (no filterbank) */
  int to_db;             /* append AmplitudeToDb(ref, amin)                                  */
  float ref, amin;
} tac_pipeline_config;

/* window_host: n_fft floats (centre-padded); fb_host: (n_bins, n_bands) row-major or NULL. */
int tac_pipeline_create(const tac_pipeline_config* cfg, const float* window_host,
                        const float* fb_host, tac_pipeline** out);
/* x_host: (n_seq, n_samples) contiguous; out_host: (n_seq, n_bands|bins, frames).
 * Copies in, computes and copies out in overlapped slices; returns when out_host is valid. */
int tac_pipeline_run_host(tac_pipeline* p, const float* x_host, int64_t n_seq, int64_t n_samples,
                          float* out_host);
int tac_pipeline_destroy(tac_pipeline* p);

/* Adjoints of the pointwise operators the reference differentiates through torch:
 *   op 0  db_to_amplitude  (functional.py:299-314): a = forward output y, g1 = dL/dy        -> out = dL/dx      (n)
 *   op 1  magphase / angle (functional.py:187-201): a = z (n x 2), g1 = dL/d|z|^p0 or NULL,
 *                                                   g2 = dL/dphase or NULL                   -> out = dL/dz      (n x 2)
 *   op 2  mu_law_decoding of float codes (functional.py:349-354): a = codes, g1 = dL/dy, p0 = mu -> out = dL/dcodes (n) */
int tac_pointwise_backward_f32(int op, const float* a, const float* g1, const float* g2, int64_t n, float p0,
                               float* out, void* stream);

/* ---- instrumentation (used by bench.py; off by default) -----------------------------------
 * tac_launch_count: kernels launched by this library in this process so far.
 * tac_profile_enable(1): bracket every kernel launch with CUDA events on its own stream;
 * tac_profile_read: synchronise those events, return the summed milliseconds and launch counts
 * per kernel kind (0 = stft, 1 = filterbank, 2 = mu-law, 3 = pointwise) and clear the log. */
int64_t tac_launch_count(void);
int tac_profile_enable(int on);
int tac_profile_read(double* ms_by_kind /*[4]*/, int64_t* launches_by_kind /*[4]*/);
/* Which kernel serves the one-kernel mel path (tac_melspec_banded_f32 and everything built on it):
 * 0 (default) = two frames per warp in packed fp32 pairs (csrc/stft_pair.cu), 1 = one frame per warp
 * (csrc/stft.cu, round 1), 2 = the pair kernel with its second 32-point FFT pass on the tcgen05 tensor cores
 * (3xTF32, operands through tensor memory; csrc/stft_pair_tc.cu).  0 and 1 agree bit for bit, 2 to fp32 rounding
 * (the transform is summed in a different order); the switch exists for A/B timing and the parity tests.
 * Returns the previous value; a negative argument only queries.  Environment: TAC_MEL_VARIANT=0|1|2. */
int tac_mel_kernel_variant(int variant);

#ifdef __cplusplus
}
#endif
#endif /* TAC_B200_H_ */
