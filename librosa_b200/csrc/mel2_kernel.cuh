// mel2_kernel.cuh — fused mel spectrogram for hop = n_fft / 4 with autonomous frame groups.
//
// Same arithmetic as fwd_kernel<…, MODE_MEL> (frame, window, packed real FFT, |.|^2, band-sparse mel rows —
// librosa/core/spectrum.py:341-390, :3000-3013, librosa/feature/spectral.py:2160), different data movement:
//
//   * A frame group (one warp, n_fft = 2048) owns a run of CONSECUTIVE frames of a clip.  With hop = n_fft / 4
//     consecutive frames share three of their four hop-sized blocks, and — because hop / 2 is a multiple of the
//     group size — the samples thread t needs from a block are the same for every frame that contains the block.
//     The raw samples of the last four blocks therefore live in a thread-private ring in Tensor Memory
//     (4 x 16 columns per warp): per frame a thread loads ONE new block from global memory (eight coalesced
//     8-byte loads, every sample read from HBM exactly once, no shared-memory staging, no TMA, no tile
//     bookkeeping) and fetches the other three with tcgen05.ld.  fwd_kernel's operand fetch — 64 shared-memory
//     wavefronts per frame, the staging barrier and the mbarrier wait — disappears.
//   * The power row of a frame goes to a row tile that is separate from the exchange regions, and the hand-over
//     to the mel phase uses two mbarriers per half-CTA (rows full / rows free) instead of CTA barriers: a warp
//     never waits for the others except to read rows they have not written yet.
//   * The mel phase is fwd_kernel's (MelLayout, 16-byte loads, frame-pair lanes); the eight rows of a tile now
//     belong to eight different runs, so every frame lane carries its own output offset.
// Used for MODE_MEL without the dB epilogue, n_fft = 2048 (one warp per frame), hop = n_fft / 4; everything
// else stays on fwd_kernel.  The window, twiddle and un-mix tables are in Tensor Memory as in fwd_kernel.
#pragma once
#include "common.cuh"
#include "fft_engine.cuh"
#include "fwd_kernel.cuh"

namespace b2l {

__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, uint32_t tb, float (&a)[8], float (&b)[8]) {
  uint32_t q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%17];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(q[i]);
    b[i] = __uint_as_float(q[8 + i]);
  }
}
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])),
        "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])),
        "r"(__float_as_uint(r[7])), "r"(__float_as_uint(r[8])), "r"(__float_as_uint(r[9])), "r"(__float_as_uint(r[10])),
        "r"(__float_as_uint(r[11])), "r"(__float_as_uint(r[12])), "r"(__float_as_uint(r[13])),
        "r"(__float_as_uint(r[14])), "r"(__float_as_uint(r[15]))
      : "memory");
}

template <int LOG2M>
__global__ void __launch_bounds__(512, 1) mel2_kernel(const FwdArgs a) {
  constexpr int TPF = 32, NW = 16;
  using Cfg = FftCfg<LOG2M, TPF>;
  using Tab = TmemTab<Cfg>;
  constexpr int M = Cfg::M, N = 2 * M, PPT = Cfg::PPT, NPAIR = PPT / 2;
  constexpr int NH = 2, HW = NW / NH, FT = HW;            // 8 frame groups (warps) per half = 8 rows per tile
  using ML = MelLayout<M, FT>;
  constexpr int H = ML::H, FP = ML::FP;
  static_assert(PPT == 32 && ML::PAIR, "one warp per frame, frame-pair mel lanes");
  constexpr int PRS = ((M + 4 + 31) / 32) * 32 + 8;       // row stride of the power tile: >= M + 4, == 8 (mod 32)
  static_assert(PRS % 32 == ML::RSM % 32 && PRS >= M + 4, "power tile stride");
  constexpr int RING0 = Tab::NCOLS;                       // sample ring: 4 blocks x 16 columns per warp
  constexpr int NCOLS = RING0 + 4 * 64;
  static_assert(NCOLS <= 512, "Tensor Memory columns");

  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, t = tid & 31;
  const int half = warp / HW, hwarp = warp % HW;          // hwarp = frame group inside the half = row of the tile
  float* s_melw = reinterpret_cast<float*>(smem + a.off_melw);
  MelRow* s_row = reinterpret_cast<MelRow*>(smem + a.off_melband);
  const unsigned short* s_order = reinterpret_cast<const unsigned short*>(smem + a.off_melorder);
  float2* xbuf = reinterpret_cast<float2*>(smem + a.off_xbuf) + warp * Cfg::XBUF_F2;
  float* s_p = reinterpret_cast<float*>(smem + a.off_in) + half * (FT * PRS);          // power tile of this half
  long long* s_out = reinterpret_cast<long long*>(smem + a.off_win) + half * FT;       // output offset per row (-1: none)
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + a.off_bar) + 2 * half;        // rows written
  uint64_t* s_free = s_full + 1;                                                       // rows consumed
  uint32_t* s_taddr = reinterpret_cast<uint32_t*>(smem + a.off_bar + 32);

  // ---- tables (Tensor Memory), mel tables (shared memory), barriers
  if (tid < 32) tmem_alloc<NCOLS>(s_taddr);
  for (int i = tid; i < a.mel_w_count; i += 512) s_melw[i] = a.mel_w[i];
  for (int i = tid; i < a.n_mel_rows; i += 512) s_row[i] = a.mel_rows[i];
  {
    unsigned short* so = reinterpret_cast<unsigned short*>(smem + a.off_melorder);
    for (int i = tid; i < a.mel_list_len * HW; i += 512) so[i] = a.mel_order[i];
  }
  if (tid < NH) {
    mbar_init(reinterpret_cast<uint64_t*>(smem + a.off_bar) + 2 * tid, HW);
    mbar_init(reinterpret_cast<uint64_t*>(smem + a.off_bar) + 2 * tid + 1, HW);
    fence_mbar_init();
  }
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tbase = *s_taddr + ((uint32_t)((warp & 3) * 32) << 16);
  if (tid < 128) {
    // window in ELEMENT order: columns 2r, 2r+1 = window[2 (t + 32 r)], [.. + 1]; twiddles as in fwd_kernel
    for (int col = 0; col < 64; ++col) tmem_store1(tbase + col, a.window[2 * (t + TPF * (col >> 1)) + (col & 1)]);
    for (int sp = 1; sp < Cfg::NPASS; ++sp) {
      const int R = Cfg::radix(sp), p = Cfg::sublen(sp);
      for (int f = 0; f < PPT; ++f) {
        const int b = f / R, r = f % R, k = (t + TPF * b) & (p - 1);
        const float2 w = r == 0 ? make_float2(1.0f, 0.0f) : a.tw[Cfg::tw_offset(sp) + (r - 1) * p + k];
        tmem_store1(tbase + 64 * sp + 2 * f, w.x);
        tmem_store1(tbase + 64 * sp + 2 * f + 1, w.y);
      }
    }
    for (int cp = 0; cp < NPAIR; ++cp) {
      const float2 w = a.twn[t + TPF * cp];
      tmem_store1(tbase + Tab::UNMIX_COL + 2 * cp, w.x);
      tmem_store1(tbase + Tab::UNMIX_COL + 2 * cp + 1, w.y);
    }
    tmem_wait_st();
  }
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  Tab tab;
  tab.taddr = tbase;
  const uint32_t ring = tbase + RING0 + (uint32_t)(warp >> 2) * 64;
  const float2 wt = make_float2(0.0f, 0.0f);

  // ---- this group's run of (clip, frame) pairs; every group of the CTA takes the same number of steps
  const long long total = (long long)a.n_clips * a.n_frames;
  const long long gidx = (long long)blockIdx.x * NW + warp;
  const int steps = a.tiles_per_clip;                     // host: frames per group
  long long gpos = gidx * (long long)steps;
  const long long gend = min(total, gpos + steps);
  const int hop = a.hop;
  int clip = -1, frame = 0;                               // current position; clip < 0: ring not primed
  const bool vec_ok = a.tma_ok != 0;

  // samples 2 e, 2 e + 1 (e = t + 32 i, i = 0 .. 7) of block j of the current clip -> nb[2 i], nb[2 i + 1]
  auto load_block = [&](const float* yc, int j, float (&nb)[16]) {
    const long long p0 = (long long)j * hop - a.pad;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long pos = p0 + 2 * (t + TPF * i);
      if (vec_ok && pos >= 0 && pos + 1 < a.n) {
        const float2 v2 = __ldg(reinterpret_cast<const float2*>(yc + pos));
        nb[2 * i] = v2.x;
        nb[2 * i + 1] = v2.y;
      } else {
        nb[2 * i] = load_padded(yc, a.n, pos, a.pad_mode, a.pad);
        nb[2 * i + 1] = load_padded(yc, a.n, pos + 1, a.pad_mode, a.pad);
      }
    }
  };

  for (int step = 0; step < steps; ++step) {
    const uint32_t parity = (uint32_t)(step & 1);
    const bool live = gpos < gend;
    float pw[PPT + 1];
    long long out_off = -1;
    if (live) {
      const int c = (int)(gpos / a.n_frames);
      const int f = (int)(gpos - (long long)c * a.n_frames);
      const float* yc = a.y + (long long)c * a.clip_stride;
      float nb[16];
      tmem_wait_st();                                     // the block stored by the previous frame is in place
      if (c != clip || f != frame) {                      // start of a run or of a new clip: prime blocks f .. f+2
        clip = c;
        frame = f;
#pragma unroll 1
        for (int u = 0; u < 3; ++u) {
          load_block(yc, f + u, nb);
          tmem_st16f(ring + 16 * ((f + u) & 3), nb);
        }
        tmem_wait_st();
      }
      load_block(yc, f + 3, nb);                          // the one new block of this frame
      // ---- windowed operands with the first butterfly stage fused in (load_pass0_windowed): element r pairs with
      // element r + 16, i.e. block u with block u + 2; blocks 0 .. 2 come from the ring, block 3 is `nb`
      float2 v[PPT];
      static_for<0, 2>([&](auto U) {
        constexpr int u = 1 - decltype(U)::value;         // blocks (1, 3) first: `nb` dies early
        static_for<0, 2>([&](auto Hh) {
          constexpr int h = decltype(Hh)::value;          // half of the block: elements i = 4h .. 4h+3
          float xa[8], xb[8], wa[8], wb[8];
          if constexpr (u == 1) {
            float dummy[8];
            tmem_ld8x2(ring + 16 * ((f + 1) & 3) + 8 * h, tbase + 16 * 1 + 8 * h, xa, wa);
            tmem_ld8x2(tbase + 16 * 3 + 8 * h, tbase + 16 * 3 + 8 * h, wb, dummy);
#pragma unroll
            for (int q = 0; q < 8; ++q) xb[q] = nb[8 * h + q];
          } else {
            tmem_ld8x2(ring + 16 * (f & 3) + 8 * h, tbase + 8 * h, xa, wa);
            tmem_ld8x2(ring + 16 * ((f + 2) & 3) + 8 * h, tbase + 16 * 2 + 8 * h, xb, wb);
          }
          static_for<0, 4>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            constexpr int ra = 8 * u + 4 * h + q;         // element of block u; its partner is ra + 16
            constexpr int sa = bitrevc(ra, 5);            // pass-0 slot of element ra (even), partner in sa + 1
            static_assert((sa & 1) == 0 && bitrevc(sa + 1, 5) == ra + 16, "first-stage pair");
            const float pr = xa[2 * q] * wa[2 * q], pi = xa[2 * q + 1] * wa[2 * q + 1];
            v[sa] = make_float2(fmaf(xb[2 * q], wb[2 * q], pr), fmaf(xb[2 * q + 1], wb[2 * q + 1], pi));
            v[sa + 1] = make_float2(fmaf(-xb[2 * q], wb[2 * q], pr), fmaf(-xb[2 * q + 1], wb[2 * q + 1], pi));
          });
        });
      });
      tmem_st16f(ring + 16 * ((f + 3) & 3), nb);          // the new block replaces block f - 1 in the ring
      ++frame;
      // ---- M-point complex FFT, un-mix, |.|^2
      fft_forward_tab<Cfg, true>(v, t, 0, xbuf, tab);
      __syncwarp();
      static_for<0, PPT>([&](auto S) {
        constexpr int slot = decltype(S)::value;
        if constexpr (spectrum_offset<Cfg>(slot) >= M / 2) xbuf[xphys(t + spectrum_offset<Cfg>(slot))] = v[slot];
      });
      tab.begin_unmix();
      __syncwarp();
      if (!(fabsf(v[0].x) + fabsf(v[0].y) <= 3.0e38f)) *a.status = 1;     // util.valid_audio on the device
      static_for<0, NPAIR>([&](auto C) {
        constexpr int cc = decltype(C)::value;
        constexpr int sa = slot_of_pair<Cfg>(cc);
        float2 A = v[sa], B = xbuf[partner_slot<M, TPF, cc>(t)], xa2, xb2;
        if constexpr (cc == 0) {
          if (t == 0) B = A;
        }
        r2c_pair(A, B, tab.template unmix<cc>(wt), xa2, xb2);
        pw[2 * cc] = sqmag(xa2);
        pw[2 * cc + 1] = sqmag(xb2);
      });
      pw[PPT] = 0.0f;
      if (t == 0) {
        float2 xa2, xb2;
        const float2 zc = xbuf[xphys(M / 2)];
        r2c_pair(zc, zc, make_float2(0.0f, -1.0f), xa2, xb2);
        pw[PPT] = sqmag(xa2);
      }
      if (a.power_mode == 1) {
        static_for<0, PPT + 1>([&](auto S) { pw[decltype(S)::value] = sqrt_approx(pw[decltype(S)::value]); });
      } else if (a.power_mode != 2) {
        static_for<0, PPT + 1>([&](auto S) {
          pw[decltype(S)::value] = power_from_sq(pw[decltype(S)::value], a.power_mode, a.power);
        });
      }
      __syncwarp();                                       // pair reads done before the next frame's exchange writes
      out_off = (long long)c * a.n_mels * a.n_frames + f;
      ++gpos;
    }
    // ---- hand the power row to the tile
    if (step > 0) mbar_wait(s_free, parity ^ 1u);         // every warp of the half is done with the previous tile
    if (live) {
      float* prow = s_p + hwarp * PRS;
      static_for<0, NPAIR>([&](auto C) {
        constexpr int cc = decltype(C)::value;
        const int k = t + TPF * cc;
        prow[k] = pw[2 * cc];
        prow[M - k] = pw[2 * cc + 1];
      });
      if (t == 0) {
        prow[M / 2] = pw[PPT];
        prow[M + 1] = 0.0f;
        prow[M + 2] = 0.0f;
        prow[M + 3] = 0.0f;
      }
    }
    if (t == 0) s_out[hwarp] = out_off;
    __syncwarp();
    if (t == 0) mbar_arrive(s_full);
    mbar_wait(s_full, parity);
    // ---- mel rows of the tile (fwd_kernel's loop; every frame lane has its own output offset)
    {
      const int fp = t & (FP - 1), j = t / FP;
      const long long oa = s_out[fp], ob = s_out[fp + FP];
      const float* pbase = s_p + fp * PRS;
      for (int li = 0; li < a.mel_list_len; ++li) {
        const int item = s_order[li * HW + hwarp];
        if (item == 0xffff) break;
        const int m = item * H + j;
        const MelRow row = s_row[m];
        const float4* wp = reinterpret_cast<const float4*>(s_melw + row.off);
        const float4* pa = reinterpret_cast<const float4*>(pbase + row.lo);
        const float4* pb = pa + (FP * PRS) / 4;
        const float4* wend = wp + row.quads;
        float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll 2
        for (; wp != wend; ++wp, ++pa, ++pb) {
          const float4 w = *wp, x = *pa, y = *pb;
          a0 = fmaf(w.x, x.x, a0);
          b0 = fmaf(w.x, y.x, b0);
          a1 = fmaf(w.y, x.y, a1);
          b1 = fmaf(w.y, y.y, b1);
          a0 = fmaf(w.z, x.z, a0);
          b0 = fmaf(w.z, y.z, b0);
          a1 = fmaf(w.w, x.w, a1);
          b1 = fmaf(w.w, y.w, b1);
        }
        if (m < a.n_mels) {
          if (oa >= 0) a.out_r[oa + (long long)m * a.n_frames] = a0 + a1;
          if (ob >= 0) a.out_r[ob + (long long)m * a.n_frames] = b0 + b1;
        }
      }
    }
    __syncwarp();
    if (t == 0) mbar_arrive(s_free);
  }

  tmem_fence_before_sync();
  __syncthreads();
  if (tid < 32) {
    tmem_fence_after_sync();
    tmem_free<NCOLS>(tab.taddr);
  }
}

}  // namespace b2l
