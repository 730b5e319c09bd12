// fwd_kernel.cuh — fused frame + pad + window + real FFT kernel with three epilogues
// (complex STFT / power spectrogram / band-sparse mel projection with optional dB + per-clip max).
//
// Replaces, per frame, the reference's  util.frame -> float64 window product -> scipy.fft.rfft
// (librosa/core/spectrum.py:341-390),  np.abs(.)**power (:3000-3013) and the mel einsum
// (librosa/feature/spectral.py:2160).
//
// Work decomposition: a persistent CTA walks tiles of FT consecutive frames of one clip.  The
// contiguous sample span of a tile, (FT-1)*hop + n_fft floats, is staged once into shared memory
// (frames overlap *inside* shared memory) — by a single 1-D TMA bulk copy (cp.async.bulk + mbarrier)
// for interior tiles, prefetched one tile ahead, or by a cooperative gather with the pad-mode index
// map for tiles that touch the clip edges.  Each group of TPF threads then transforms one frame in
// registers (fft_engine.cuh).
#pragma once
#include "common.cuh"
#include "fft_engine.cuh"
#include "stats.cuh"

namespace b2l {

// ------------------------------------------------------------------ mbarrier / TMA (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------ padded sample fetch
// Virtual sample j of a clip of n samples under np.pad semantics (librosa/core/spectrum.py:252-328,
// equivalence to np.pad(y, n_fft//2, mode) per SURVEY Appendix A.2).
__device__ __forceinline__ float load_padded(const float* __restrict__ y, int n, long long j, int mode, int pad) {
  if (j >= 0 && j < n) return __ldg(y + j);
  switch (mode) {
    case PAD_EDGE:
      return __ldg(y + (j < 0 ? 0 : n - 1));
    case PAD_REFLECT: {
      if (n == 1) return __ldg(y);
      long long P = 2LL * (n - 1);
      long long m = j % P;
      if (m < 0) m += P;
      if (m >= n) m = P - m;
      return __ldg(y + m);
    }
    case PAD_SYMMETRIC: {
      long long P = 2LL * n;
      long long m = j % P;
      if (m < 0) m += P;
      if (m >= n) m = P - 1 - m;
      return __ldg(y + m);
    }
    case PAD_LINEAR_RAMP: {
      long long d = j < 0 ? -j : j - (n - 1);          // distance from the edge sample, 1..pad
      float edge = __ldg(y + (j < 0 ? 0 : n - 1));
      long long i = pad - d;                           // np.linspace(0, edge, pad, endpoint=False)[i]
      if (i <= 0) return 0.0f;
      return (float)((double)i * ((double)edge / (double)pad));
    }
    default:
      return 0.0f;
  }
}

// |X|^power from |X|^2 for power != 2 (kept out of line: the default power-2 path never executes it)
static __device__ __noinline__ float power_from_sq(float p2, int power_mode, float power) {
  float mag = sqrtf(p2);
  return power_mode == 1 ? mag : powf(mag, power);
}
__device__ __forceinline__ float sqmag(float2 x) { return fmaf(x.x, x.x, x.y * x.y); }
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Exchange-buffer slot of Z[M - k] for k = t + TPF*c (the partner of bin k in the real-FFT un-mix).
// For warp-multiple groups the padded index is affine in the lane with a compile-time offset.
template <int M, int TPF, int C>
__device__ __forceinline__ int partner_slot(int t) {
  if constexpr (TPF % 32 == 0) {
    constexpr int WPG = TPF / 32;                 // warps per group
    const int tl = t & 31, tw = t >> 5;
    const int row = tw + WPG * C;                 // k = tl + 32*row
    const int rows = M / 32;
    // tl != 0: M-k = 32*(rows-1-row) + (32-tl);  tl == 0: M-k = 32*((rows-row) mod rows)
    const int a = 33 * (rows - 1 - row) + 32 - tl;
    const int b = 33 * ((rows - row) & (rows - 1));
    return tl == 0 ? b : a;
  } else {
    const int k = t + TPF * C;
    return xphys((M - k) & (M - 1));
  }
}

// ------------------------------------------------------------------ the kernel
// NW warps per CTA.  With NSPLIT > 1 the CTA is NSPLIT independent parts of NW/NSPLIT warps ("virtual CTAs",
// called halves below): each has
// its own staging buffer, exchange area, mbarrier, named barrier and tile sequence, and only the constant
// tables are shared.  The halves drift apart, so the shared-memory-bound phases of one (operand fetch,
// exchange, mel gather) overlap the FP32-bound butterflies of the other instead of all warps of the SM
// hitting the same pipe at once.
template <int LOG2M, int TPF, int NW, int MODE, int NSPLIT>
__global__ void __launch_bounds__(NW * 32, 1) fwd_kernel(const FwdArgs a) {
  using Cfg = FftCfg<LOG2M, TPF>;
  constexpr int M = Cfg::M, N = 2 * M, PPT = Cfg::PPT;
  constexpr int NT = NW * 32;
  constexpr int NH = NSPLIT;                   // independent parts ("halves" when 2)
  constexpr bool DUAL = NSPLIT > 1;
  constexpr int HT = NT / NH;                  // threads per half
  constexpr int HW = NW / NH;                  // warps per half
  constexpr int FT = HT / TPF;                 // frames per tile == frame groups per half
  static_assert(NW % NSPLIT == 0, "the warps are split evenly");
  static_assert(TPF <= 32 || NH + NT / TPF <= 15, "named barriers: 1..NH for the halves, then one per frame group");
  static_assert(FT >= 1 && FT <= 32, "tile must hold 1..32 frames");
  using ML = MelLayout<M, FT>;
  constexpr int H = ML::H;                     // mel rows handled concurrently by one warp (common.cuh)
  constexpr int NPAIR = PPT / 2;               // bin pairs (k, M-k) per thread

  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int half = DUAL ? tid / HT : 0;
  const int htid = tid - half * HT;
  float* s_win = reinterpret_cast<float*>(smem + a.off_win);
  float2* s_tw = reinterpret_cast<float2*>(smem + a.off_tw);
  float* s_melw = reinterpret_cast<float*>(smem + a.off_melw);
  MelRow* s_row = reinterpret_cast<MelRow*>(smem + a.off_melband);
  float* s_in = reinterpret_cast<float*>(smem + a.off_in + half * a.in_stride);
  float2* s_xall = reinterpret_cast<float2*>(smem + a.off_xbuf + half * a.xbuf_stride);
  float* s_p = reinterpret_cast<float*>(s_xall);               // P tile aliases the exchange area
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + a.off_bar) + half;

  const int grp = htid / TPF;                  // frame group == local frame index
  const int t = htid % TPF;
  const int gbar = 1 + NH + half * FT + grp;   // named barrier of this frame group (used when TPF > 32)
  float2* xbuf = s_xall + grp * Cfg::XBUF_F2;

  auto half_sync = [&]() {
    if constexpr (DUAL) asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "n"(HT) : "memory");   // ids 1 .. NH
    else __syncthreads();
  };

  // ---- one-time table staging (whole CTA)
  for (int i = tid; i < N; i += NT) s_win[i] = a.window[i];
  for (int i = tid; i < Cfg::TW_COUNT; i += NT) s_tw[i] = a.tw[i];
  if constexpr (MODE == MODE_MEL) {
    for (int i = tid; i < a.mel_w_count; i += NT) s_melw[i] = a.mel_w[i];
    for (int i = tid; i < a.n_mel_rows; i += NT) s_row[i] = a.mel_rows[i];
  }
  if constexpr (MODE == MODE_STATS) {
    for (int i = tid; i < a.mel_w_count; i += NT) s_melw[i] = a.mel_w[i];   // bin frequencies
  }
  if (htid == 0) {
    mbar_init(s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int span = a.in_floats;
  const bool hop_even = (a.hop & 1) == 0;
  // un-mix twiddle of bin k = t + TPF*c:  W_N^k = W_N^t * W_(2*PPT)^c  (register x compile-time constant)
  const float2 wt = __ldg(a.twn + t);
  auto unmix_tw = [&](auto C) -> float2 {
    constexpr int c = decltype(C)::value;
    if constexpr (c == 0) return wt;
    else return cmul(wt, make_float2(TwC<c, 2 * PPT>::re, TwC<c, 2 * PPT>::im));
  };

  auto tile_origin = [&](long long tile, int& clip, int& t0, long long& s0) {
    clip = (int)(tile / a.tiles_per_clip);
    t0 = (int)(tile % a.tiles_per_clip) * FT;
    s0 = (long long)t0 * a.hop - a.pad;
  };
  // How a tile's span gets into shared memory:
  //   TILE_TMA      entirely inside the clip and 16-byte aligned -> one bulk copy
  //   TILE_TMA_ZERO zero ("constant") padding, aligned: bulk-copy the in-range part, threads zero the rest
  //   TILE_GATHER   anything else (reflect / edge / ... padding, unaligned clips): per-sample gather
  enum { TILE_GATHER = 0, TILE_TMA = 1, TILE_TMA_ZERO = 2 };
  auto tile_kind = [&](long long tile, int& lead, int& valid) -> int {
    int clip, t0;
    long long s0;
    tile_origin(tile, clip, t0, s0);
    lead = s0 < 0 ? (int)(-s0) : 0;                                   // floats before the clip starts
    long long end = s0 + span;
    valid = (int)((end > a.n ? (long long)a.n : end) - (s0 + lead));  // in-range floats
    if (!a.tma_ok || ((s0 + lead) & 3) != 0) return TILE_GATHER;
    if (lead == 0 && valid == span) return TILE_TMA;
    if (a.pad_mode == PAD_CONSTANT && valid > 0 && (lead & 3) == 0 && (valid & 3) == 0) return TILE_TMA_ZERO;
    return TILE_GATHER;
  };
  auto issue_tma = [&](long long tile, int lead, int valid) {
    int clip, t0;
    long long s0;
    tile_origin(tile, clip, t0, s0);
    fence_proxy_async();
    mbar_expect_tx(s_bar, (uint32_t)valid * 4u);
    tma_load_1d(s_in + lead, a.y + (long long)clip * a.clip_stride + s0 + lead, (uint32_t)valid * 4u, s_bar);
  };
  // Called by the whole half after the staging buffer has been released (B0): start the copy of `tile`.
  auto prefetch = [&](long long tile) {
    if (tile >= a.total_tiles) return;
    int lead, valid;
    const int kind = tile_kind(tile, lead, valid);
    if (kind == TILE_GATHER) return;
    if (htid == 0) issue_tma(tile, lead, valid);
    if (kind == TILE_TMA_ZERO) {
      for (int i = htid; i < lead; i += HT) s_in[i] = 0.0f;
      for (int i = lead + valid + htid; i < span; i += HT) s_in[i] = 0.0f;
    }
  };

  const long long tile_step = (long long)gridDim.x * NH;
  long long tile = (long long)blockIdx.x * NH + half;
  uint32_t phase = 0;
  prefetch(tile);

  for (; tile < a.total_tiles; tile += tile_step) {
    int clip, t0;
    long long s0;
    tile_origin(tile, clip, t0, s0);
    // ---------------- stage the tile's sample span
    {
      int lead, valid;
      const int kind = tile_kind(tile, lead, valid);
      if (kind == TILE_GATHER) {
        const float* yc = a.y + (long long)clip * a.clip_stride;
        for (int i = htid; i < span; i += HT) s_in[i] = load_padded(yc, a.n, s0 + i, a.pad_mode, a.pad);
        half_sync();
      } else {
        mbar_wait(s_bar, phase);
        phase ^= 1;
        if (kind == TILE_TMA_ZERO) half_sync();   // zeros written by other threads
      }
    }

    // ---------------- windowed frame -> registers (pass-0 operands)
    float2 v[PPT];
    {
      const float* fr = s_in + grp * a.hop;
      if (hop_even) {
        load_pass0<Cfg>(v, t, [&](int e) {
          float2 x = *reinterpret_cast<const float2*>(fr + 2 * e);
          float2 w = *reinterpret_cast<const float2*>(s_win + 2 * e);
          return make_float2(x.x * w.x, x.y * w.y);
        });
      } else {
        load_pass0<Cfg>(v, t, [&](int e) {
          float2 w = *reinterpret_cast<const float2*>(s_win + 2 * e);
          return make_float2(fr[2 * e] * w.x, fr[2 * e + 1] * w.y);
        });
      }
    }
    half_sync();   // B0: staging buffer consumed -> prefetch the next tile behind the math
    prefetch(tile + tile_step);

    // ---------------- M-point complex FFT
    fft_forward<Cfg>(v, t, gbar, xbuf, s_tw);
    if constexpr (Cfg::NPASS > 1) group_sync<TPF>(gbar);
    // Bin pair (k, M-k), k = t + TPF*c < M/2: Z[k] is already in one of this thread's registers; only the
    // upper half of the spectrum (indices >= M/2) goes through shared memory to reach its partner.
    static_for<0, PPT>([&](auto S) {
      constexpr int slot = decltype(S)::value;
      if constexpr (spectrum_offset<Cfg>(slot) >= M / 2) xbuf[xphys(t + spectrum_offset<Cfg>(slot))] = v[slot];
    });
    group_sync<TPF>(gbar);
    auto pair_operands = [&](auto C, float2& A, float2& B) {
      constexpr int c = decltype(C)::value;
      constexpr int sa = slot_of_pair<Cfg>(c);
      static_assert(sa >= 0, "pair operand must be register resident");
      A = v[sa];
      B = xbuf[partner_slot<M, TPF, c>(t)];
      if constexpr (c == 0) {
        if (t == 0) B = A;   // k = 0 pairs with itself (Z[M] == Z[0])
      }
    };

    const int frame = t0 + grp;
    const bool frame_ok = frame < a.n_frames;
    // util.valid_audio (librosa/util/utils.py:303-306) on the device: one NaN / Inf sample makes every
    // bin of every frame that reads it non-finite, so testing one spectrum value per thread and frame
    // catches it without a separate pass over the input.
    if (frame_ok && !(fabsf(v[0].x) + fabsf(v[0].y) <= 3.0e38f)) *a.status = 1;

    if constexpr (MODE == MODE_STFT) {
      float2* orow = a.out_c + ((long long)clip * a.n_frames + frame) * (M + 1);
      static_for<0, NPAIR>([&](auto C) {
        const int k = t + TPF * decltype(C)::value;
        float2 A, B, xa, xb;
        pair_operands(C, A, B);
        r2c_pair(A, B, unmix_tw(C), xa, xb);
        if (frame_ok) {
          orow[k] = xa;
          orow[M - k] = xb;
        }
      });
      if (t == 0) {
        float2 xa, xb;
        float2 zc = xbuf[xphys(M / 2)];
        r2c_pair(zc, zc, make_float2(0.0f, -1.0f), xa, xb);   // W_N^(M/2) = -i
        if (frame_ok) orow[M / 2] = xa;
      }
      group_sync<TPF>(gbar);   // pair reads done before the next tile's exchange writes
    } else {
      float pw[PPT + 1];
      static_for<0, NPAIR>([&](auto C) {
        constexpr int c = decltype(C)::value;
        float2 A, B, xa, xb;
        pair_operands(C, A, B);
        r2c_pair(A, B, unmix_tw(C), xa, xb);
        pw[2 * c] = sqmag(xa);
        pw[2 * c + 1] = sqmag(xb);
      });
      pw[PPT] = 0.0f;
      if (t == 0) {
        float2 xa, xb;
        float2 zc = xbuf[xphys(M / 2)];
        r2c_pair(zc, zc, make_float2(0.0f, -1.0f), xa, xb);
        pw[PPT] = sqmag(xa);
      }
      if (MODE == MODE_STATS || a.power_mode == 1) {
        // |X| (power = 1; the statistics of MODE_STATS are defined on it): MUFU-based square root, inline.
        // sqrt.approx.f32 is accurate to 2^-23 relative; 33 calls of an out-of-line IEEE square root per
        // thread and frame would cost as much as half the FFT.
        static_for<0, PPT + 1>([&](auto S) { pw[decltype(S)::value] = sqrt_approx(pw[decltype(S)::value]); });
      } else if (a.power_mode != 2) {   // warp-uniform, cold: general exponent through powf
        static_for<0, PPT + 1>([&](auto S) {
          pw[decltype(S)::value] = power_from_sq(pw[decltype(S)::value], a.power_mode, a.power);
        });
      }
      if constexpr (MODE == MODE_SPEC) {
        float* orow = a.out_r + ((long long)clip * a.n_frames + frame) * (M + 1);
        if (frame_ok) {
          static_for<0, NPAIR>([&](auto C) {
            constexpr int c = decltype(C)::value;
            const int k = t + TPF * c;
            orow[k] = pw[2 * c];
            orow[M - k] = pw[2 * c + 1];
          });
          if (t == 0) orow[M / 2] = pw[PPT];
        }
        group_sync<TPF>(gbar);
      } else {
        // ---------------- band-sparse mel projection over the tile
        half_sync();   // B1: every group finished reading its Z
        // P[f][k] frame-major with a bank-skewed row stride (MelLayout, common.cuh); bins M+1 .. M+3 of
        // every row are kept at zero so the 4-bin groups of the mel loop may run past the Nyquist bin.
        constexpr int RS = ML::RS;
        float* prow = s_p + grp * RS;
        static_for<0, NPAIR>([&](auto C) {
          constexpr int c = decltype(C)::value;
          const int k = t + TPF * c;
          prow[k] = pw[2 * c];
          prow[M - k] = pw[2 * c + 1];
        });
        if (t == 0) prow[M / 2] = pw[PPT];
        if (htid < 3 * FT) s_p[(htid / 3) * RS + M + 1 + (htid % 3)] = 0.0f;
        half_sync();   // B2
        if constexpr (MODE == MODE_STATS) {
          // one warp per frame of the tile: statistics of the magnitude row (stats.cuh)
          const int hwarp = htid >> 5, lane = htid & 31;
          for (int f = hwarp; f < FT; f += HW) {
            bool negative;
            const float r = frame_stats(s_p + f * RS, s_melw, M + 1, lane, a.stats, &negative);
            if (lane < N_STATS && t0 + f < a.n_frames)
              a.out_r[((long long)clip * N_STATS + lane) * a.n_frames + t0 + f] = r;
          }
        } else {
          // Work item = H adjacent mel rows; lane (fp, j) accumulates row item*H + j for the frame pair
          // (fp, fp + FP) over that row's padded band (host-built MelRow table: the rows of an item share
          // one trip count, start bins are congruent to j mod H so the skewed tile reads conflict-free,
          // weights are zero padded and 16-byte aligned).  One weight fetch feeds both frames; no
          // cross-lane reduction, one short loop per item.
          constexpr int FP = ML::FP;
          constexpr bool PAIR = ML::PAIR;
          const int hwarp = htid >> 5, lane = htid & 31;
          const int fp = lane & (FP - 1), j = lane / FP;
          const bool ok_a = (t0 + fp) < a.n_frames;
          const bool ok_b = PAIR && (t0 + fp + FP) < a.n_frames;
          float wmax = -INFINITY;
          const int n_items = a.n_mel_rows / H;
          const float* pbase = s_p + fp * RS;
          // output row stride / base: the public [clip][mel][frame] layout, or the tiled mfcc scratch whose
          // 64-frame tiles are contiguous 32 KB blocks for dct_clamp_kernel (t0 + fp and t0 + fp + FP share a tile)
          const long long orow = a.out_tiled ? 64 : a.n_frames;
          float* obase = a.out_tiled
                             ? a.out_r + ((long long)clip * ((a.n_frames + 63) >> 6) + (t0 >> 6)) * a.n_mels * 64 + (t0 & 63) + fp
                             : a.out_r + (long long)clip * a.n_mels * a.n_frames + t0 + fp;
          for (int item = hwarp; item < n_items; item += HW) {
            const int m = item * H + j;
            const MelRow row = s_row[m];
            const float4* wp = reinterpret_cast<const float4*>(s_melw + row.off);
            const float* pa = pbase + row.lo;
            const float4* wend = wp + row.quads;
            float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
            if constexpr (PAIR) {
              const float* pb = pa + FP * RS;
#pragma unroll 2
              for (; wp != wend; ++wp, pa += 4, pb += 4) {
                const float4 w = *wp;
                a0 = fmaf(w.x, pa[0], a0);
                b0 = fmaf(w.x, pb[0], b0);
                a1 = fmaf(w.y, pa[1], a1);
                b1 = fmaf(w.y, pb[1], b1);
                a0 = fmaf(w.z, pa[2], a0);
                b0 = fmaf(w.z, pb[2], b0);
                a1 = fmaf(w.w, pa[3], a1);
                b1 = fmaf(w.w, pb[3], b1);
              }
            } else {
#pragma unroll 2
              for (; wp != wend; ++wp, pa += 4) {
                const float4 w = *wp;
                a0 = fmaf(w.x, pa[0], a0);
                a1 = fmaf(w.y, pa[1], a1);
                a0 = fmaf(w.z, pa[2], a0);
                a1 = fmaf(w.w, pa[3], a1);
              }
            }
            float va = a0 + a1, vb = b0 + b1;
            if (m < a.n_mels) {
              if (a.log_mode) {
                va = 10.0f * log10f(fmaxf(a.amin, va)) - a.db_sub;
                if (ok_a) wmax = fmaxf(wmax, va);
                if constexpr (PAIR) {
                  vb = 10.0f * log10f(fmaxf(a.amin, vb)) - a.db_sub;
                  if (ok_b) wmax = fmaxf(wmax, vb);
                }
              }
              float* o = obase + (long long)m * orow;
              if (ok_a) o[0] = va;
              if (ok_b) o[FP] = vb;
            }
          }
          if (a.log_mode) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
            if (lane == 0 && wmax > -INFINITY) atomicMax(a.clip_max + clip, float_to_key(wmax));
          }
        }
        half_sync();   // B3: P reads done before the next tile's exchange writes
      }
    }
  }
}

}  // namespace b2l
