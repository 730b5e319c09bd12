// aux_kernels.cuh — the small kernels around the FFT path: clamp + DCT for mfcc, mel projection and
// power_to_db for S= inputs, and a batched transpose used as a layout adapter.
#pragma once
#include "common.cuh"

namespace b2l {

// ------------------------------------------------------------------ clamp + DCT (mfcc pass B)
// in  L   [n_clips][n_mels][T]   log-mel (dB), not yet clamped
// out C   [n_clips][n_mfcc][T]   C[k][t] = sum_m dct[k][m] * max(L[m][t], clipmax - top_db)
// Reference: np.maximum(log_spec, log_spec.max(...) - top_db) (librosa/core/spectrum.py:1881) followed by
// scipy.fft.dct(S, axis=-2, type, norm)[..., :n_mfcc, :] (* lifter) (librosa/feature/spectral.py:2005-2015);
// the DCT (any type / norm, lifter folded in) arrives as an explicit [n_mfcc][n_mels] matrix.
// Block: KG warps; warp w owns coefficients 8w .. 8w+7, lane owns frames lane + 32 i (i = 0..3) of a
// 128-frame tile that is staged (clamped) in shared memory; DCT rows are read as warp-uniform float4.
constexpr int DCT_TILE = 128;

__global__ void dct_clamp_kernel(const float* __restrict__ L, const float* __restrict__ dct,
                                 const unsigned int* __restrict__ clip_max, float top_db, int n_mels,
                                 int n_mfcc, int T, int tiles_per_clip, float* __restrict__ C) {
  extern __shared__ __align__(16) float s_dyn[];
  const int KG = blockDim.x >> 5;
  float* s_tile = s_dyn;                       // [n_mels][DCT_TILE]
  float* s_dct = s_dyn + n_mels * DCT_TILE;    // [n_mels][8*KG]  (transposed, zero padded)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KP = 8 * KG;
  for (int i = tid; i < n_mels * KP; i += blockDim.x) {
    int m = i / KP, k = i % KP;
    s_dct[i] = k < n_mfcc ? dct[k * n_mels + m] : 0.0f;
  }
  const int clip = blockIdx.x / tiles_per_clip;
  const int t0 = (blockIdx.x % tiles_per_clip) * DCT_TILE;
  float floor_v = -INFINITY;
  if (clip_max != nullptr && top_db >= 0.0f) floor_v = key_to_float(clip_max[clip]) - top_db;
  const float* Lc = L + (long long)clip * n_mels * T;
  for (int i = tid; i < n_mels * DCT_TILE; i += blockDim.x) {
    int m = i / DCT_TILE, x = i % DCT_TILE;
    float val = 0.0f;
    if (t0 + x < T) val = fmaxf(Lc[(long long)m * T + t0 + x], floor_v);
    s_tile[i] = val;
  }
  __syncthreads();
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
  for (int m = 0; m < n_mels; ++m) {
    const float4 d0 = *reinterpret_cast<const float4*>(s_dct + m * KP + 8 * warp);
    const float4 d1 = *reinterpret_cast<const float4*>(s_dct + m * KP + 8 * warp + 4);
    const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    float xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = s_tile[m * DCT_TILE + lane + 32 * i];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(dv[j], xv[i], acc[j][i]);
  }
  float* Cc = C + (long long)clip * n_mfcc * T;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = 8 * warp + j;
    if (k < n_mfcc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tt = t0 + lane + 32 * i;
        if (tt < T) Cc[(long long)k * T + tt] = acc[j][i];
      }
    }
  }
}

// ------------------------------------------------------------------ mel projection of a given spectrogram
// S [n_clips][T][F] (bins contiguous) -> mel [n_clips][n_mels][T]; band-sparse rows as in the fused kernel.
// One warp per (clip, frame-tile of 32); lanes own frames, band weights are warp-uniform.
__global__ void mel_project_kernel(const float* __restrict__ S, const float* __restrict__ mel_w,
                                   const MelBand* __restrict__ band, int n_mels, int F, int T,
                                   int tiles_per_clip, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_tile[];   // [F][33]
  const int clip = blockIdx.x / tiles_per_clip;
  const int t0 = (blockIdx.x % tiles_per_clip) * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NWp = blockDim.x >> 5;
  const float* Sc = S + ((long long)clip * T + t0) * F;
  for (int f = warp; f < 32; f += NWp) {
    const bool ok = t0 + f < T;
    for (int k = lane; k < F; k += 32) s_tile[k * 33 + f] = ok ? Sc[(long long)f * F + k] : 0.0f;
  }
  __syncthreads();
  for (int m = warp; m < n_mels; m += NWp) {
    const MelBand b = band[m];
    float acc = 0.0f;
    for (int kx = 0; kx < b.len; ++kx) acc = fmaf(__ldg(mel_w + b.off + kx), s_tile[(b.lo + kx) * 33 + lane], acc);
    if (t0 + lane < T) out[((long long)clip * n_mels + m) * T + t0 + lane] = acc;
  }
}

// ------------------------------------------------------------------ power_to_db
__global__ void db_kernel(const float* __restrict__ in, long long per_clip, float amin, float db_sub,
                          unsigned int* __restrict__ clip_max, float* __restrict__ out) {
  const int clip = blockIdx.y;
  const float* ic = in + (long long)clip * per_clip;
  float* oc = out + (long long)clip * per_clip;
  float mx = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_clip;
       i += (long long)gridDim.x * blockDim.x) {
    float v = 10.0f * log10f(fmaxf(amin, ic[i])) - db_sub;
    oc[i] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomicMax(clip_max + clip, float_to_key(mx));
}

__global__ void db_clamp_kernel(float* __restrict__ x, long long per_clip, const unsigned int* __restrict__ clip_max,
                                float top_db) {
  const int clip = blockIdx.y;
  float* xc = x + (long long)clip * per_clip;
  const float floor_v = key_to_float(clip_max[clip]) - top_db;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_clip;
       i += (long long)gridDim.x * blockDim.x)
    xc[i] = fmaxf(xc[i], floor_v);
}

// ------------------------------------------------------------------ finite scan of samples no frame reads
// y [n_clips][stride]; checks samples [begin, n) of every clip (the uncovered tail when the last frame
// ends before the clip does, or the whole clip when hop > n_fft leaves gaps).
__global__ void finite_scan_kernel(const float* __restrict__ y, long long stride, int n, int begin, int* status) {
  const float* yc = y + (long long)blockIdx.y * stride;
  bool bad = false;
  for (long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    bad |= !(fabsf(yc[i]) <= 3.0e38f);
  if (bad) *status = 1;
}

// ------------------------------------------------------------------ batched transpose
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, int rows, int cols, T* __restrict__ out) {
  __shared__ T tile[32][33];
  const long long base = (long long)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[base + (long long)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[base + (long long)c * rows + r] = tile[threadIdx.x][j];
  }
}

}  // namespace b2l
