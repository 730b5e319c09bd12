// stats.cuh — per-frame statistics of a magnitude spectrum, computed by one warp from a row in shared memory.
//
// Covers the reductions librosa applies to `_spectrogram` output (librosa/feature/spectral.py):
//   spectral_centroid  (:46-191)   sum_k f_k S_k / sum_k S_k         (util.normalize(S, norm=1): no division
//                                                                    when the sum is below tiny(float32))
//   spectral_bandwidth (:194-352)  (sum_k S_k |f_k - centroid|^p)^(1/p), S normalised when norm=True
//   spectral_rolloff   (:535-684)  min f_k over the bins whose running sum has reached roll_percent * total
//   spectral_flatness  (:687-803)  exp(mean log max(amin, S^power)) / mean max(amin, S^power)
//   rms(S=...)         (:806-916)  sqrt(2 sum' S^2 / frame_length^2), DC (and Nyquist for even lengths) halved
// Lane l owns the contiguous bins [l*chunk, (l+1)*chunk) with `chunk` odd, so a warp-wide access
// row[l*chunk + i] touches 32 distinct banks, and the running sum needed by the roll-off is a lane-local
// walk on top of an exclusive warp scan of the chunk totals.
#pragma once
#include "common.cuh"

namespace b2l {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Returns the statistic for row `lane` (lanes >= N_STATS return 0); *negative is set when an entry is < 0.
__device__ inline float frame_stats(const float* row, const float* freq, int F, int lane, const StatsParams& sp,
                                    bool* negative) {
  const int chunk = ((F + 31) / 32) | 1;
  const int k0 = min(F, lane * chunk), k1 = min(F, k0 + chunk);
  float m0 = 0.0f, m1 = 0.0f, e2 = 0.0f, am = 0.0f, lg = 0.0f;
  bool neg = false;
  const int fmode = sp.flat_power == 2.0f ? 2 : (sp.flat_power == 1.0f ? 1 : 0);
  // warp-uniform switches: a caller that wants one row does not pay for the others
  const bool need_m1 = sp.want & ((1 << STAT_CENTROID) | (1 << STAT_BANDWIDTH));
  const bool need_flat = sp.want & (1 << STAT_FLATNESS);
  const bool need_e2 = sp.want & (1 << STAT_RMS);
  const bool need_pass2 = sp.want & ((1 << STAT_BANDWIDTH) | (1 << STAT_ROLLOFF));
  for (int k = k0; k < k1; ++k) {
    const float s = row[k];
    neg |= s < 0.0f;
    m0 += s;
    if (need_m1) m1 = fmaf(freq[k], s, m1);
    if (need_e2 | need_flat) {
      const float p2 = s * s;
      e2 += p2;
      if (need_flat) {
        float th = fmode == 2 ? p2 : (fmode == 1 ? s : powf(s, sp.flat_power));
        th = fmaxf(sp.flat_amin, th);
        am += th;
        lg += __log2f(th);
      }
    }
  }
  *negative = __any_sync(0xffffffffu, neg);
  // exclusive scan of the chunk totals
  float incl = m0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  float run = incl - m0;
  const float total = __shfl_sync(0xffffffffu, incl, 31);
  const float m1t = warp_sum(m1), e2t = warp_sum(e2), amt = warp_sum(am), lgt = warp_sum(lg);
  const float norm = total < 1.17549435e-38f ? 1.0f : total;     // util.normalize threshold = tiny(float32)
  const float centroid = m1t / norm;
  const float thr = sp.roll_percent * total;
  const bool p_is_2 = sp.bw_p == 2.0f;
  float bw = 0.0f, rmin = INFINITY;
  if (need_pass2) {
    for (int k = k0; k < k1; ++k) {
      const float s = row[k], fk = freq[k];
      run += s;
      if (run >= thr) rmin = fminf(rmin, fk);
      const float d = fabsf(fk - centroid);
      bw = fmaf(s, p_is_2 ? d * d : powf(d, sp.bw_p), bw);
    }
  }
  float bwt = warp_sum(bw);
  if (sp.bw_norm) bwt /= norm;
  const float bandwidth = p_is_2 ? sqrtf(bwt) : powf(bwt, 1.0f / sp.bw_p);
  rmin = warp_min(rmin);
  if (rmin == INFINITY) rmin = freq[F - 1];   // rounding left the last running sum a hair under the threshold
  const float invF = 1.0f / (float)F;
  const float flat = exp2f(lgt * invF) / (amt * invF);
  float e2a = e2t - 0.5f * row[0] * row[0];
  if ((sp.frame_length & 1) == 0) e2a -= 0.5f * row[F - 1] * row[F - 1];
  const float L = (float)sp.frame_length;
  const float rms = sqrtf(2.0f * e2a / (L * L));
  float r = 0.0f;
  if (lane == STAT_CENTROID) r = centroid;
  if (lane == STAT_BANDWIDTH) r = bandwidth;
  if (lane == STAT_ROLLOFF) r = rmin;
  if (lane == STAT_FLATNESS) r = flat;
  if (lane == STAT_RMS) r = rms;
  if (lane == STAT_TOTAL) r = total;
  return r;
}

}  // namespace b2l
