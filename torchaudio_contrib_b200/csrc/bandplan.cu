// Host-side construction of the band plan (bandplan.cuh) from a (n_bins, n_bands) row-major matrix.
#include "bandplan.cuh"

#include <string.h>

#include <vector>

#include "tac_common.cuh"

namespace tac {

int64_t build_band_plan(const float* fb, int n_bins, int n_bands, unsigned char* dst, int64_t capacity) {
  if (n_bins != kBandBins || n_bands < 1 || n_bands > 4096) return 0;
  if (capacity < band_plan_capacity(n_bands)) return 0;

  // ---- every bin -> (band b, w0 into b, w1 into b + 1) -------------------------------------------------
  std::vector<int> band(n_bins, 0);
  std::vector<float> w0(n_bins, 0.0f), w1(n_bins, 0.0f);
  int prev = 0;
  for (int k = 0; k < n_bins; ++k) {
    const float* row = fb + (size_t)k * n_bands;
    int first = -1, last = -1, nnz = 0;
    for (int b = 0; b < n_bands; ++b)
      if (row[b] != 0.0f) {              // NaN != 0 too: it has to reach the output, as in the reference's matmul
        if (first < 0) first = b;
        last = b;
        ++nnz;
      }
    if (nnz > 2 || (nnz == 2 && last != first + 1)) return 0;
    if (nnz == 0) {
      band[k] = prev;
    } else if (nnz == 2) {
      band[k] = first;
      w0[k] = row[first];
      w1[k] = row[first + 1];
    } else if (first == prev + 1) {      // continue the running segment: the single weight is its upper band
      band[k] = prev;
      w1[k] = row[first];
    } else {
      band[k] = first;
      w0[k] = row[first];
    }
    prev = band[k];
  }

  // ---- per lane: segment mask and the floats it stores --------------------------------------------------
  // lane l walks bins 32 l .. 32 l + 31 (lane 31: .. 1024).  Bit i: the band steps after local bin i (bit 31:
  // between bins 1023 and 1024, lane 31 only): u goes to column i of the lane's stash row, then u = v, v = 0.
  std::vector<std::vector<uint16_t>> lists(n_bands);
  auto record = [&](int b, int pos, bool used) {
    if (used && b >= 0 && b < n_bands) lists[b].push_back((uint16_t)pos);
  };
  uint32_t mask[32], end_pos[32];
  int n_stored = 0;
  for (int l = 0; l < 32; ++l) {
    mask[l] = 0;
    end_pos[l] = (l == 31) ? (uint32_t)kStashEnd31 : (uint32_t)(l * kStashStride + 31);
    const int k_begin = 32 * l, k_end = (l == 31) ? n_bins : 32 * l + 32;     // exclusive
    bool u_used = false, v_used = false;
    for (int k = k_begin; k < k_end; ++k) {
      u_used |= (w0[k] != 0.0f);
      v_used |= (w1[k] != 0.0f);
      if (k + 1 == k_end) {
        record(band[k], (int)end_pos[l], u_used);
        record(band[k] + 1, (int)end_pos[l] + 1, v_used);
        n_stored += 2;
      } else if (band[k + 1] != band[k]) {
        if (band[k + 1] != band[k] + 1 && v_used) return 0;     // v would be orphaned: not a chain
        mask[l] |= 1u << (k - k_begin);
        record(band[k], l * kStashStride + (k - k_begin), u_used);
        ++n_stored;
        u_used = (band[k + 1] == band[k] + 1) ? v_used : false;  // u = v carries the upper band's sum on
        v_used = false;
      }
    }
  }
  int cmax = 1;
  for (int b = 0; b < n_bands; ++b)
    if ((int)lists[b].size() > cmax) cmax = (int)lists[b].size();
  if (cmax > kBandMaxComb) return 0;
  const int pad = (n_bands + 31) / 32 * 32;

  BandPlanHeader hdr;
  memset(&hdr, 0, sizeof(hdr));
  hdr.magic = kBandPlanMagic;
  hdr.n_bins = n_bins;
  hdr.n_bands = n_bands;
  hdr.n_stored = n_stored;
  hdr.cmax = cmax;
  hdr.n_bands_pad = pad;
  hdr.zero_idx = kStashZero;
  memcpy(dst, &hdr, sizeof(hdr));

  float* w = reinterpret_cast<float*>(dst + kBandOffW);
  for (int j = 0; j < 16; ++j)
    for (int l = 0; l < 32; ++l) {
      const int k = 32 * l + 2 * j;
      float* q = w + ((size_t)j * 32 + l) * 4;
      q[0] = w0[k];
      q[1] = w1[k];
      q[2] = w0[k + 1];
      q[3] = w1[k + 1];
    }
  uint32_t* meta = reinterpret_cast<uint32_t*>(dst + kBandOffMeta);
  for (int l = 0; l < 32; ++l) {
    meta[4 * l] = mask[l];
    meta[4 * l + 1] = end_pos[l];
    float e0 = 0.0f, e1 = 0.0f;
    if (l == 31) {
      e0 = w0[n_bins - 1];
      e1 = w1[n_bins - 1];
    }
    memcpy(&meta[4 * l + 2], &e0, 4);
    memcpy(&meta[4 * l + 3], &e1, 4);
  }
  uint16_t* comb = reinterpret_cast<uint16_t*>(dst + kBandOffComb);
  for (int c = 0; c < cmax; ++c)
    for (int b = 0; b < pad; ++b)
      comb[(size_t)c * pad + b] = (b < n_bands && c < (int)lists[b].size()) ? lists[b][c] : (uint16_t)kStashZero;
  int64_t used = kBandOffComb + (int64_t)cmax * pad * 2;
  used = (used + 15) & ~(int64_t)15;
  // fast form: lane l sums bands l, l + 32, l + 64, l + 96, four entries each, as byte offsets into the stash
  if (n_bands <= 128 && cmax <= 4) {
    uint32_t* fastp = reinterpret_cast<uint32_t*>(dst + used);
    for (int l = 0; l < 32; ++l)
      for (int j = 0; j < 4; ++j)
        for (int c = 0; c < 4; ++c) {
          const int b = l + 32 * j;
          const int idx = (b < n_bands && c < (int)lists[b].size()) ? lists[b][c] : kStashZero;
          fastp[(l * 4 + j) * 4 + c] = (uint32_t)idx * 4u;
        }
    hdr.reserved = (int32_t)used;
    memcpy(dst, &hdr, sizeof(hdr));
    used += 32 * 16 * 4;
    // the same lists for the frame-major output layout: lane l sums the four ADJACENT bands 4 l .. 4 l + 3, so a frame's
    // bands leave the lane as one 16-byte store (512 contiguous bytes per warp instruction)
    uint32_t* quadp = reinterpret_cast<uint32_t*>(dst + used);
    for (int l = 0; l < 32; ++l)
      for (int j = 0; j < 4; ++j)
        for (int c = 0; c < 4; ++c) {
          const int b = 4 * l + j;
          const int idx = (b < n_bands && c < (int)lists[b].size()) ? lists[b][c] : kStashZero;
          quadp[(l * 4 + j) * 4 + c] = (uint32_t)idx * 4u;
        }
    used += 32 * 16 * 4;
  }
  return used;
}

int64_t build_range_plan(const float* fb, int n_bins, int n_bands, unsigned char* dst, int64_t capacity) {
  if (n_bins < 1 || n_bins > 65535 || n_bands < 1 || n_bands > 4096) return 0;
  if (capacity < range_plan_capacity(n_bins, n_bands)) return 0;
  const int pad = (n_bands + 31) / 32 * 32;
  std::vector<int> lo(pad, 0), len(pad, 0);
  int64_t total = 0;
  int max_len = 0;
  for (int b = 0; b < n_bands; ++b) {
    int first = -1, last = -1;
    for (int k = 0; k < n_bins; ++k)
      if (fb[(size_t)k * n_bands + b] != 0.0f) {          // NaN != 0 too: it reaches the output as in the reference's matmul
        if (first < 0) first = k;
        last = k;
      }
    if (first >= 0) {
      lo[b] = first;
      len[b] = last - first + 1;
      total += len[b];
      if (len[b] > max_len) max_len = len[b];
    }
  }
  if (total > (int64_t)4 * n_bins || max_len > 32767) return 0;   // dense: the tensor-core path is the right one
  const int nnz_pad = (int)((total + 3) / 4 * 4);
  RangePlanHeader* hdr = reinterpret_cast<RangePlanHeader*>(dst);
  memset(hdr, 0, sizeof(*hdr));
  hdr->magic = kRangePlanMagic;
  hdr->n_bins = n_bins;
  hdr->n_bands = n_bands;
  hdr->n_bands_pad = pad;
  hdr->nnz_pad = nnz_pad;
  hdr->max_len = max_len;
  int32_t* meta = reinterpret_cast<int32_t*>(dst + 32);
  float* w = reinterpret_cast<float*>(dst + 32 + (size_t)pad * 8);
  int off = 0;
  for (int b = 0; b < pad; ++b) {
    meta[2 * b] = lo[b] | (len[b] << 16);
    meta[2 * b + 1] = off;
    for (int i = 0; i < len[b]; ++i) w[off + i] = fb[(size_t)(lo[b] + i) * n_bands + b];
    off += len[b];
  }
  for (int i = off; i < nnz_pad; ++i) w[i] = 0.0f;
  return 32 + (int64_t)pad * 8 + (int64_t)nnz_pad * 4;
}

}  // namespace tac
