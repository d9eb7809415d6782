// N2: phase vocoder (functional.py:204-274) -- time-stretch of a complex STFT by `rate` without changing pitch.
//
// Per (sequence, bin) row of the (*, bins, T, 2) input and output step j (T_out = ceil(T / rate) of them):
//   i0 = long(j * rate), i1 = long(j * rate + 1), alpha = frac(j * rate)   (functional.py:239-243, :250-253; the three
//        tables come from the host, built with the reference's own torch ops in its dtype, so they are exact)
//   c0 = spec[i0], c1 = spec[i1]                         (two zero frames appended, :247-248)
//   dphi = angle(c1) - angle(c0) - advance[bin];  dphi -= 2 pi round(dphi / 2 pi);  dphi += advance[bin]   (:261-265)
//   acc_j = angle(spec[0]) + sum_{i < j} dphi_i          (:266-267, cumsum of [phase_0, dphi_0 ... dphi_{n-2}])
//   out_j = (alpha |c1| + (1 - alpha) |c0|) (cos acc_j, sin acc_j)                                          (:269-272)
//
// The accumulated phase grows to ~pi * hop * T (1e5 .. 1e6 rad), so a float32 running sum loses the output
// after a few hundred frames; that is why the reference's own value test runs in float64
// (tests/test_functional.py:85-88).  This kernel always evaluates in float64 (angles, wrap, prefix sum,
// sin / cos) and only the loads and stores have the tensor's dtype.
//
// One warp per row: lanes take 32 consecutive output steps, the prefix sum is a warp shuffle scan with a
// carried total, so reads along T and writes along T_out are coalesced.  HBM-bound: ~16 / rate + 8 bytes per
// output value pair in float32.
#include <math.h>

#include "tac_common.cuh"

namespace tac {

constexpr int kPvWarps = 8;
constexpr int kPvThreads = kPvWarps * 32;

template <typename T>
struct Pair;
template <>
struct Pair<float> {
  using type = float2;
};
template <>
struct Pair<double> {
  using type = double2;
};

template <typename T>
__global__ void __launch_bounds__(kPvThreads)
phase_vocoder_kernel(const typename Pair<T>::type* __restrict__ spec, int64_t n_rows, int n_bins, int64_t n_in,
                     const int32_t* __restrict__ idx0, const int32_t* __restrict__ idx1, const double* __restrict__ alpha,
                     const T* __restrict__ advance,
                     int64_t n_out, typename Pair<T>::type* __restrict__ out) {
  using P = typename Pair<T>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kPvWarps + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * kPvWarps;
  constexpr double kTwoPi = 6.283185307179586476925286766559;
  for (int64_t row = warp_global; row < n_rows; row += warp_stride) {
    const P* src = spec + row * n_in;
    P* dst = out + row * n_out;
    const double adv = (double)advance[row % n_bins];
    const P first = src[0];
    double carry = atan2((double)first.y, (double)first.x);            // phase_0
    for (int64_t base = 0; base < n_out; base += 32) {
      const int64_t j = base + lane;
      double dphi = 0.0, mag = 0.0;
      if (j < n_out) {
        const int64_t i0 = idx0[j], i1 = idx1[j];
        const double a = alpha[j];
        P c0, c1;
        c0.x = c0.y = c1.x = c1.y = (T)0;
        if (i0 < n_in) c0 = src[i0];
        if (i1 < n_in) c1 = src[i1];
        const double r0 = (double)c0.x, q0 = (double)c0.y, r1 = (double)c1.x, q1 = (double)c1.y;
        dphi = atan2(q1, r1) - atan2(q0, r0) - adv;
        dphi = dphi - kTwoPi * rint(dphi / kTwoPi);                      // torch.round: half to even
        dphi = dphi + adv;
        mag = a * sqrt(r1 * r1 + q1 * q1) + (1.0 - a) * sqrt(r0 * r0 + q0 * q0);
      }
      // inclusive scan of dphi over the 32 lanes
      double incl = dphi;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      const double acc = carry + (incl - dphi);                          // exclusive: sum of the steps before j
      carry += __shfl_sync(0xffffffffu, incl, 31);
      if (j < n_out) {
        double sn, cs;
        sincos(acc, &sn, &cs);
        P o;
        o.x = (T)(mag * cs);
        o.y = (T)(mag * sn);
        dst[j] = o;
      }
    }
  }
}

template <typename T>
static int launch_phase_vocoder(const void* spec, int64_t n_seq, int n_bins, int64_t n_in, const int32_t* idx0,
                                const int32_t* idx1, const double* alpha, const void* advance, int64_t n_out, void* out, cudaStream_t stream) {
  TAC_REQUIRE(n_seq >= 0 && n_bins > 0 && n_in > 0 && n_out >= 0, TAC_ERR_INVALID,
              "phase_vocoder: bad shape n_seq=%lld bins=%d time=%lld -> %lld", (long long)n_seq, n_bins, (long long)n_in,
              (long long)n_out);
  const int64_t n_rows = n_seq * n_bins;
  if (n_rows == 0 || n_out == 0) return TAC_OK;
  TAC_REQUIRE(spec && idx0 && idx1 && alpha && advance && out, TAC_ERR_INVALID, "phase_vocoder: null pointer");
  TAC_REQUIRE(n_in + 2 < ((int64_t)1 << 31), TAC_ERR_UNSUPPORTED, "phase_vocoder: %lld frames exceed the 2^31 the index table holds",
              (long long)n_in);
  const int64_t want = (n_rows + kPvWarps - 1) / kPvWarps;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want < cap ? want : cap);
  using P = typename Pair<T>::type;
  LaunchProbe probe(KIND_POINTWISE, stream);
  phase_vocoder_kernel<T><<<grid, kPvThreads, 0, stream>>>(static_cast<const P*>(spec), n_rows, n_bins, n_in, idx0, idx1, alpha,
                                                           static_cast<const T*>(advance), n_out, static_cast<P*>(out));
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

// ---- backward (round 2) ------------------------------------------------------------------------------------------
// With g_j = dL/d out_j:  dL/d mag_j = g_j . (cos acc_j, sin acc_j) =: dm_j,   dL/d acc_j = mag_j g_j . (-sin, cos) =: h_j,
// dL/d phase_0 = H = sum_j h_j,   dL/d dphi_j = S_j = sum_{i > j} h_i   (adjoint of the cumulative sum; the phase wrap is
// piecewise constant: no gradient), and per input frame c = spec[i]:
//   grad c = [ sum_{j: i0_j = i} (1 - a_j) dm_j + sum_{j: i1_j = i} a_j dm_j ] c / |c|
//          + [ sum_{j: i1_j = i} S_j - sum_{j: i0_j = i} S_j + (i == 0) H ] (-q, r) / (r^2 + q^2)
// (|c| = 0: torch.norm's subgradient 0; the angle term is torch.atan2's formula as it stands).  i0 and i1 are monotone in
// j, so the steps that read frame i are two contiguous ranges [lo, hi) handed in as tables (built on the host with
// searchsorted): a gather per input frame, no atomics, deterministic.  One warp per row, two phases: A walks the output
// steps as the forward kernel does (float64 throughout) and leaves (dm_j, P_j = sum_{i <= j} h_i) in a per-row workspace;
// B walks the input frames.  Workspace: 16 bytes per output step.
template <typename T>
__global__ void __launch_bounds__(kPvThreads)
phase_vocoder_backward_kernel(const typename Pair<T>::type* __restrict__ spec, const typename Pair<T>::type* __restrict__ grad_out, int64_t n_rows,
                              int n_bins, int64_t n_in, const int32_t* __restrict__ idx0, const int32_t* __restrict__ idx1,
                              const double* __restrict__ alpha, const T* __restrict__ advance, int64_t n_out,
                              const int32_t* __restrict__ range0, const int32_t* __restrict__ range1, double2* __restrict__ ws,
                              typename Pair<T>::type* __restrict__ grad_spec) {
  using P = typename Pair<T>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kPvWarps + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * kPvWarps;
  constexpr double kTwoPi = 6.283185307179586476925286766559;
  for (int64_t row = warp_global; row < n_rows; row += warp_stride) {
    const P* src = spec + row * n_in;
    const P* gsrc = grad_out + row * n_out;
    double2* w = ws + row * n_out;
    const double adv = (double)advance[row % n_bins];
    const P first = src[0];
    double carry = atan2((double)first.y, (double)first.x);
    double carry_h = 0.0;
    for (int64_t base = 0; base < n_out; base += 32) {                  // phase A: forward recomputation, dm_j and prefix of h
      const int64_t j = base + lane;
      double dphi = 0.0, mag = 0.0;
      if (j < n_out) {
        const int64_t i0 = idx0[j], i1 = idx1[j];
        const double a = alpha[j];
        P c0, c1;
        c0.x = c0.y = c1.x = c1.y = (T)0;
        if (i0 < n_in) c0 = src[i0];
        if (i1 < n_in) c1 = src[i1];
        const double r0 = (double)c0.x, q0 = (double)c0.y, r1 = (double)c1.x, q1 = (double)c1.y;
        dphi = atan2(q1, r1) - atan2(q0, r0) - adv;
        dphi = dphi - kTwoPi * rint(dphi / kTwoPi);
        dphi = dphi + adv;
        mag = a * sqrt(r1 * r1 + q1 * q1) + (1.0 - a) * sqrt(r0 * r0 + q0 * q0);
      }
      double incl = dphi;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      const double acc = carry + (incl - dphi);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      double dm = 0.0, h = 0.0;
      if (j < n_out) {
        double sn, cs;
        sincos(acc, &sn, &cs);
        const P g = gsrc[j];
        dm = (double)g.x * cs + (double)g.y * sn;
        h = mag * ((double)g.y * cs - (double)g.x * sn);
      }
      double hin = h;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, hin, d);
        if (lane >= d) hin += up;
      }
      if (j < n_out) w[j] = make_double2(dm, carry_h + hin);
      carry_h += __shfl_sync(0xffffffffu, hin, 31);
    }
    __syncwarp();
    const double H = carry_h;
    P* gdst = grad_spec + row * n_in;
    for (int64_t base = 0; base < n_in; base += 32) {                   // phase B: gather per input frame
      const int64_t i = base + lane;
      if (i < n_in) {
        const P c = src[i];
        const double r = (double)c.x, q = (double)c.y;
        double gmag = 0.0, gang = (i == 0) ? H : 0.0;
        for (int j = range0[2 * i]; j < range0[2 * i + 1]; ++j) {
          const double2 e = w[j];
          gmag += (1.0 - alpha[j]) * e.x;
          gang -= H - e.y;
        }
        for (int j = range1[2 * i]; j < range1[2 * i + 1]; ++j) {
          const double2 e = w[j];
          gmag += alpha[j] * e.x;
          gang += H - e.y;
        }
        const double n2 = r * r + q * q, n1 = sqrt(n2);
        const double sr = n1 > 0.0 ? r / n1 : 0.0, sq = n1 > 0.0 ? q / n1 : 0.0;
        P o;
        o.x = (T)(gmag * sr + (gang != 0.0 ? gang * (-q / n2) : 0.0));
        o.y = (T)(gmag * sq + (gang != 0.0 ? gang * (r / n2) : 0.0));
        gdst[i] = o;
      }
    }
    __syncwarp();
  }
}

template <typename T>
static int launch_phase_vocoder_backward(const void* spec, const void* grad_out, int64_t n_seq, int n_bins, int64_t n_in, const int32_t* idx0,
                                         const int32_t* idx1, const double* alpha, const void* advance, int64_t n_out,
                                         const int32_t* range0, const int32_t* range1, void* workspace, int64_t workspace_bytes,
                                         void* grad_spec, cudaStream_t stream) {
  TAC_REQUIRE(n_seq >= 0 && n_bins > 0 && n_in > 0 && n_out >= 0, TAC_ERR_INVALID, "phase_vocoder_backward: bad shape");
  const int64_t n_rows = n_seq * n_bins;
  if (n_rows == 0) return TAC_OK;
  TAC_REQUIRE(spec && grad_spec && idx0 && idx1 && alpha && advance && range0 && range1 && (grad_out || n_out == 0), TAC_ERR_INVALID,
              "phase_vocoder_backward: null pointer");
  TAC_REQUIRE(n_in + 2 < ((int64_t)1 << 31), TAC_ERR_UNSUPPORTED, "phase_vocoder_backward: too many frames");
  TAC_REQUIRE(workspace_bytes >= n_rows * n_out * 16 && (workspace || n_out == 0), TAC_ERR_WORKSPACE,
              "phase_vocoder_backward: workspace of %lld bytes needed", (long long)(n_rows * n_out * 16));
  const int64_t want = (n_rows + kPvWarps - 1) / kPvWarps;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want < cap ? want : cap);
  using P = typename Pair<T>::type;
  LaunchProbe probe(KIND_POINTWISE, stream);
  phase_vocoder_backward_kernel<T><<<grid, kPvThreads, 0, stream>>>(static_cast<const P*>(spec), static_cast<const P*>(grad_out), n_rows, n_bins, n_in,
                                                                    idx0, idx1, alpha, static_cast<const T*>(advance), n_out, range0, range1,
                                                                    static_cast<double2*>(workspace), static_cast<P*>(grad_spec));
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

}  // namespace tac

extern "C" int tac_phase_vocoder_backward_f32(const float* spec, const float* grad_out, int64_t n_seq, int n_bins, int64_t n_in,
                                              const int32_t* idx0, const int32_t* idx1, const double* alpha, const float* advance,
                                              int64_t n_out, const int32_t* range0, const int32_t* range1, void* workspace,
                                              int64_t workspace_bytes, float* grad_spec, void* stream) {
  return tac::launch_phase_vocoder_backward<float>(spec, grad_out, n_seq, n_bins, n_in, idx0, idx1, alpha, advance, n_out, range0, range1,
                                                   workspace, workspace_bytes, grad_spec, tac::as_stream(stream));
}

extern "C" int tac_phase_vocoder_backward_f64(const double* spec, const double* grad_out, int64_t n_seq, int n_bins, int64_t n_in,
                                              const int32_t* idx0, const int32_t* idx1, const double* alpha, const double* advance,
                                              int64_t n_out, const int32_t* range0, const int32_t* range1, void* workspace,
                                              int64_t workspace_bytes, double* grad_spec, void* stream) {
  return tac::launch_phase_vocoder_backward<double>(spec, grad_out, n_seq, n_bins, n_in, idx0, idx1, alpha, advance, n_out, range0, range1,
                                                    workspace, workspace_bytes, grad_spec, tac::as_stream(stream));
}

extern "C" int tac_phase_vocoder_f32(const float* spec, int64_t n_seq, int n_bins, int64_t n_in, const int32_t* idx0,
                                     const int32_t* idx1, const double* alpha, const float* advance, int64_t n_out,
                                     float* out, void* stream) {
  return tac::launch_phase_vocoder<float>(spec, n_seq, n_bins, n_in, idx0, idx1, alpha, advance, n_out, out, tac::as_stream(stream));
}

extern "C" int tac_phase_vocoder_f64(const double* spec, int64_t n_seq, int n_bins, int64_t n_in, const int32_t* idx0,
                                     const int32_t* idx1, const double* alpha, const double* advance, int64_t n_out,
                                     double* out, void* stream) {
  return tac::launch_phase_vocoder<double>(spec, n_seq, n_bins, n_in, idx0, idx1, alpha, advance, n_out, out, tac::as_stream(stream));
}
