// mr_kernel.cuh — mixed-radix real FFT frames for even n_fft whose half is 5-smooth (2^a 3^b 5^c): the 400 / 320 /
// 480 / 800 / 960 / 1200-sample frames of speech front ends (25 ms at 16 kHz = 400 samples, ...).
//
// librosa.stft hands such frames to scipy.fft.rfft, whose ducc plans are mixed-radix Cooley-Tukey as well
// (librosa/core/spectrum.py:388); round 1 ran them through Bluestein's chirp-z transform on a power-of-two engine
// (czt_kernel.cuh: two 1024-point complex transforms per pair of 400-sample frames — five times the work of the
// n_fft = 512 kernel).  Here ONE WARP owns a frame:
//     z[e] = (x[2e] w[2e], x[2e+1] w[2e+1]),  e < M = n_fft / 2      packed real input, window folded in
//     Stockham passes of radix 5 / 3 / 8 / 4 / 2 between two shared-memory buffers of the warp (autosort, natural
//     order in and out; butterflies dealt to the lanes; pass twiddles from a shared table built in double precision)
//     X[k], X[M-k] from Z[k], Z[M-k]                                  real-FFT un-mix (r2c_pair)
// and the epilogues of the chirp-z kernel it replaces (complex STFT rows, |X|^power rows) plus a fused band-sparse
// mel projection (librosa/feature/spectral.py:2160) with the optional dB conversion and per-clip maximum that mfcc
// needs (librosa/core/spectrum.py:1866-1881).  Frames are read straight from global memory through the np.pad index
// map — consecutive warps take consecutive frames, so the n_fft / hop overlap is served by L1 / L2 and HBM sees
// every sample once.  Odd radices run first: with sub-length p = 1 a lane writes R consecutive values, and an odd
// R keeps the 64-bit stores of a half warp in distinct banks.
#pragma once
#include "common.cuh"
#include "fft_engine.cuh"
#include "fwd_kernel.cuh"   // load_padded, power_from_sq, sqmag, db10, sqrt_approx

namespace b2l {

constexpr int kMrMaxPass = 12;

struct MrArgs {
  const float* y;
  long long clip_stride;
  int n, n_clips;
  int L, M, hop, pad, pad_mode, n_frames, n_bins;   // L = n_fft, M = L / 2, n_bins = M + 1
  int n_pass;
  int radix[kMrMaxPass];
  int tw_off[kMrMaxPass];   // pass s (sub-length p_s > 1): tw[tw_off[s] + (r-1) p_s + k] = exp(-2 pi i r k / (p_s R_s))
  int tw_count;
  const float* win;         // [L] window * 1/2
  const float2* tw;
  const float2* twn;        // [M/2 + 1] exp(-2 pi i k / L)
  float2* out_c;            // mode 0: complex64 [clip][frame][bin]
  float* out_r;             // mode 1: |X|^power [clip][frame][bin];  mode 2: mel [clip][mel][frame]
  int mode, power_mode;
  float power;
  int* status;
  // mode 2
  const MelBand* band;
  const float* mel_w;
  int n_mels, mel_w_count;
  int log_mode;
  float amin, db_sub;
  unsigned int* clip_max;
};

// ---- radix-R DFTs in registers, natural order in and out (forward sign)
__device__ __forceinline__ void mr_dft3(float2 (&u)[3]) {
  constexpr float hs3 = 0.86602540378443864676f;   // sin(2 pi / 3)
  const float2 t1 = add2(u[1], u[2]);
  const float2 m1 = fma2(bc2(-0.5f), t1, u[0]);
  const float2 d = sub2(u[1], u[2]);
  const float2 m2 = mul2(mulni2(d), bc2(hs3));      // -i sin(2 pi/3) (u1 - u2)
  u[0] = add2(u[0], t1);
  u[1] = add2(m1, m2);
  u[2] = sub2(m1, m2);
}
__device__ __forceinline__ void mr_dft5(float2 (&u)[5]) {
  constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2 pi/5), cos(4 pi/5)
  constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2 pi/5), sin(4 pi/5)
  const float2 t1 = add2(u[1], u[4]), t2 = add2(u[2], u[3]);
  const float2 t3 = sub2(u[1], u[4]), t4 = sub2(u[2], u[3]);
  const float2 a1 = fma2(bc2(c2), t2, fma2(bc2(c1), t1, u[0]));
  const float2 a2 = fma2(bc2(c1), t2, fma2(bc2(c2), t1, u[0]));
  const float2 b1 = fma2(bc2(s2), t4, mul2(bc2(s1), t3));
  const float2 b2 = fma2(bc2(-s1), t4, mul2(bc2(s2), t3));
  u[0] = add2(u[0], add2(t1, t2));
  u[1] = add2(a1, mulni2(b1));   // a1 - i b1
  u[4] = add2(a1, muli2(b1));
  u[2] = add2(a2, mulni2(b2));
  u[3] = add2(a2, muli2(b2));
}

// One Stockham pass of radix R over M points: butterflies i = lane, lane + 32, ... < T = M / R.
template <int R>
__device__ __forceinline__ void mr_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                        const float2* __restrict__ tw, int p, int T, int lane) {
  constexpr bool POW2 = (R & (R - 1)) == 0;
  constexpr int LOGR = ilog2c(R);
  int k = lane % p;
  const int kstep = 32 % p;
  for (int i = lane; i < T; i += 32) {
    float2 u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 x = src[i + r * T];
      if (r > 0 && p > 1) x = cmul(x, tw[(r - 1) * p + k]);
      u[POW2 ? bitrevc(r, LOGR) : r] = x;
    }
    if constexpr (R == 3) mr_dft3(u);
    else if constexpr (R == 5) mr_dft5(u);
    else dft_reg<R, 0>(u);
    const int j = (i - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) dst[j + q * p] = u[q];
    k += kstep;
    if (k >= p) k -= p;
  }
}

// Dynamic shared memory: window [L] | pass twiddles [tw_count] | un-mix twiddles [M/2+1] | mel bands | mel weights |
// per warp: two exchange buffers of M complex values.
__host__ __device__ inline size_t mr_table_bytes(int L, int tw_count, int n_mels, int mel_w_count) {
  size_t b = (size_t)((L + 3) & ~3) * 4 + (size_t)((tw_count + 1) & ~1) * 8 + (size_t)((L / 4 + 2) & ~1) * 8;
  b += (size_t)n_mels * sizeof(MelBand) + (size_t)((mel_w_count + 3) & ~3) * 4;
  return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(512, 2) mr_kernel(const MrArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int L = a.L, M = a.M;
  float* s_win = reinterpret_cast<float*>(smem);
  float2* s_tw = reinterpret_cast<float2*>(s_win + ((L + 3) & ~3));
  float2* s_twn = s_tw + ((a.tw_count + 1) & ~1);
  MelBand* s_band = reinterpret_cast<MelBand*>(s_twn + ((L / 4 + 2) & ~1));
  float* s_melw = reinterpret_cast<float*>(s_band + a.n_mels);
  float2* s_x = reinterpret_cast<float2*>(smem + mr_table_bytes(L, a.tw_count, a.n_mels, a.mel_w_count));
  for (int i = tid; i < L; i += blockDim.x) s_win[i] = a.win[i];
  for (int i = tid; i < a.tw_count; i += blockDim.x) s_tw[i] = a.tw[i];
  for (int i = tid; i <= M / 2; i += blockDim.x) s_twn[i] = a.twn[i];
  if (a.mode == 2) {
    for (int i = tid; i < a.n_mels; i += blockDim.x) s_band[i] = a.band[i];
    for (int i = tid; i < a.mel_w_count; i += blockDim.x) s_melw[i] = a.mel_w[i];
  }
  __syncthreads();
  float2* buf0 = s_x + (size_t)warp * 2 * M;
  float2* buf1 = buf0 + M;

  const long long total = (long long)a.n_clips * a.n_frames;
  const long long stride = (long long)gridDim.x * nwarps;
  const bool vec_ok = ((a.clip_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 7) == 0);
  float wmax = -INFINITY;
  int wmax_clip = -1;
  auto flush_max = [&]() {   // per-clip maximum of the dB values this warp produced (log_mode)
    if (wmax_clip >= 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
      if (lane == 0 && wmax > -INFINITY) atomicMax(a.clip_max + wmax_clip, float_to_key(wmax));
    }
    wmax = -INFINITY;
  };
  // (clip, frame) of this warp's frames without a division per frame: advanced by the constant warp stride
  const long long f0 = (long long)blockIdx.x * nwarps + warp;
  int clip = (int)(f0 / a.n_frames), frame = (int)(f0 - (long long)clip * a.n_frames);
  const int step_c = (int)(stride / a.n_frames), step_f = (int)(stride - (long long)step_c * a.n_frames);
  for (long long f = f0; f < total; f += stride, clip += step_c, frame += step_f) {
    if (frame >= a.n_frames) {
      frame -= a.n_frames;
      ++clip;
    }
    const float* yc = a.y + (long long)clip * a.clip_stride;
    const long long s0 = (long long)frame * a.hop - a.pad;
    // ---- packed, windowed input
    if (s0 >= 0 && s0 + L <= a.n && vec_ok && (s0 & 1) == 0) {
      const float2* y2 = reinterpret_cast<const float2*>(yc + s0);
      const float2* w2 = reinterpret_cast<const float2*>(s_win);
      for (int e = lane; e < M; e += 32) buf0[e] = mul2(__ldg(y2 + e), w2[e]);
    } else {
      for (int e = lane; e < M; e += 32) {
        const float x0 = load_padded(yc, a.n, s0 + 2 * e, a.pad_mode, a.pad);
        const float x1 = load_padded(yc, a.n, s0 + 2 * e + 1, a.pad_mode, a.pad);
        buf0[e] = make_float2(x0 * s_win[2 * e], x1 * s_win[2 * e + 1]);
      }
    }
    __syncwarp();
    // ---- Stockham passes
    float2* src = buf0;
    float2* dst = buf1;
    int p = 1;
    for (int s = 0; s < a.n_pass; ++s) {
      const int R = a.radix[s], T = M / R;
      const float2* tw = s_tw + a.tw_off[s];
      switch (R) {
        case 5: mr_pass<5>(src, dst, tw, p, T, lane); break;
        case 3: mr_pass<3>(src, dst, tw, p, T, lane); break;
        case 8: mr_pass<8>(src, dst, tw, p, T, lane); break;
        case 4: mr_pass<4>(src, dst, tw, p, T, lane); break;
        default: mr_pass<2>(src, dst, tw, p, T, lane); break;
      }
      __syncwarp();
      float2* tmp = src;
      src = dst;
      dst = tmp;
      p *= R;
    }
    // ---- real-FFT un-mix and epilogue: src holds Z[0 .. M), dst is free
    const long long orow = ((long long)clip * a.n_frames + frame) * a.n_bins;
    float* prow = reinterpret_cast<float*>(dst);
    bool bad = false;
    for (int k = lane; k <= M / 2; k += 32) {
      const float2 A = src[k], B = src[k == 0 ? 0 : M - k];
      bad = bad || !(fabsf(A.x) + fabsf(A.y) <= 3.0e38f);
      float2 xa, xb;
      r2c_pair(A, B, s_twn[k], xa, xb);
      const bool two = (M - k) != k;
      if (a.mode == 0) {
        a.out_c[orow + k] = xa;
        if (two) a.out_c[orow + M - k] = xb;
      } else {
        float pa = sqmag(xa), pb = sqmag(xb);
        if (a.power_mode == 1) {
          pa = sqrt_approx(pa);
          pb = sqrt_approx(pb);
        } else if (a.power_mode != 2) {
          pa = power_from_sq(pa, a.power_mode, a.power);
          pb = power_from_sq(pb, a.power_mode, a.power);
        }
        if (a.mode == 1) {
          a.out_r[orow + k] = pa;
          if (two) a.out_r[orow + M - k] = pb;
        } else {
          prow[k] = pa;
          if (two) prow[M - k] = pb;
        }
      }
    }
    if (bad) *a.status = 1;   // util.valid_audio (librosa/util/utils.py:303-306): a non-finite sample poisons every bin
    if (a.mode == 2) {
      __syncwarp();
      if (a.log_mode && clip != wmax_clip) {
        flush_max();
        wmax_clip = clip;
      }
      float* o = a.out_r + (long long)clip * a.n_mels * a.n_frames + frame;
      for (int m = lane; m < a.n_mels; m += 32) {
        const MelBand b = s_band[m];
        const float* w = s_melw + b.off;
        const float* x = prow + b.lo;
        float acc0 = 0.0f, acc1 = 0.0f;
        int i = 0;
        for (; i + 1 < b.len; i += 2) {
          acc0 = fmaf(w[i], x[i], acc0);
          acc1 = fmaf(w[i + 1], x[i + 1], acc1);
        }
        if (i < b.len) acc0 = fmaf(w[i], x[i], acc0);
        float v = acc0 + acc1;
        if (a.log_mode) {
          v = db10(fmaxf(a.amin, v)) - a.db_sub;
          wmax = fmaxf(wmax, v);
        }
        o[(long long)m * a.n_frames] = v;
      }
    }
    __syncwarp();   // the rows are consumed before the next frame overwrites the buffers
  }
  if (a.mode == 2 && a.log_mode) flush_max();
}

}  // namespace b2l
