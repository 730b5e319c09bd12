// inv2_kernel.cuh — istft for hop = n_fft / R: every frame group walks its own run of consecutive frames and
// keeps the overlap-add state in Tensor Memory.
//
// Replaces librosa.istft's per-block  win * scipy.fft.irfft(D)  (librosa/core/spectrum.py:566, :598), the numba
// __overlap_add loop (:629-643) and the window-sum-square division (:606-624), like inv_kernel.cuh, for the
// hop lengths librosa uses by default (hop = n_fft / 4) and their neighbours (R = 2, 4, 8).
//
// Why a second kernel.  inv_kernel.cuh transforms G frames per round, parks them in shared memory and lets the
// whole half-CTA gather the output samples: two CTA-wide barriers per round, an index-heavy gather (a third of
// its instructions are integer bookkeeping) and spectrum rows loaded by warps that all wait at the same time.
// Here a frame group (TPF threads = one or two warps) is autonomous:
//   * it owns a run of consecutive (clip, frame) pairs and takes them one after the other — the only
//     synchronisation is inside the group, so the load latency of one group hides behind the arithmetic of the
//     fifteen others;
//   * after the inverse FFT thread t holds the windowed samples 2e, 2e+1 of its frame for e = t + TPF*i,
//     i = 0 .. 31.  With hop = n_fft / R a frame is R chunks of hop samples and chunk j of frame f belongs to
//     output chunk f + j; because hop/2 is a multiple of TPF, the thread that holds a sample of frame f holds
//     the samples of frames f-1, f-2, ... that land on the same output position too.  The partial sums of the
//     R-1 unfinished output chunks are therefore THREAD-PRIVATE and live in Tensor Memory (one 32-bit column
//     per value, 16-column tcgen05.ld / tcgen05.st per chunk), not in shared memory;
//   * chunk f is complete once frame f has been added (frames are added in increasing order — the order of
//     the reference's loop, so results are bit-identical to inv_kernel): it is scaled by 1/wss and written
//     straight from registers with coalesced 8-byte stores.
// A run that starts in the middle of a clip first replays the R-1 frames before it (without emitting) to
// rebuild the state.  The window (1/n_fft folded in), the inter-pass twiddles and the un-mix twiddles come
// from Tensor Memory as in fwd_kernel (TmemTab).
#pragma once
#include "common.cuh"
#include "fft_engine.cuh"
#include "fwd_kernel.cuh"   // TMEM helpers

namespace b2l {

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(q[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16]) {
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

// R = n_fft / hop.  PPT / R elements (sample pairs) of a thread fall into one chunk: 8 for R = 4 — one
// 16-column TMEM access per chunk; R = 2 and R = 8 split / merge chunks into the same 16-column units.
template <int LOG2M, int TPF, int NW, int R>
__global__ void __launch_bounds__(NW * 32, 1) inv2_kernel(const InvArgs a) {
  using Cfg = FftCfg<LOG2M, TPF>;
  using Tab = TmemTab<Cfg>;
  constexpr int M = Cfg::M, PPT = Cfg::PPT, NT = NW * 32, NG = NT / TPF, NPAIR = PPT / 2;
  static_assert(PPT == 32 && TPF >= 32 && (128 % TPF) == 0, "one or more whole warps per frame, 32 points per thread");
  static_assert(R == 2 || R == 4 || R == 8, "hop = n_fft / 2, / 4 or / 8");
  // A frame = 4 units of 8 elements (16 TMEM columns) per thread; the output advances by ADV = 4 / R units per
  // frame, so NLIVE = 4 - ADV units of unfinished output are carried from frame to frame: slot j holds the
  // partial sums of output unit (first unit of the next frame) + j.  Unit u of the current frame is added to
  // slot u; the first ADV sums are final (emitted), the others move down to slot u - ADV — in place, because
  // slot u - ADV was consumed ADV units earlier.  (R = 8 advances by half a unit: served by inv_kernel.)
  static_assert(R == 2 || R == 4, "R = 8 is served by inv_kernel");
  constexpr int UNITS = 4;
  constexpr int ADV = UNITS / R;                             // units finished per frame: 1 (R = 4) or 2 (R = 2)
  constexpr int NLIVE = UNITS - ADV;                         // 3 or 2
  constexpr int ACC_COLS = 16 * NLIVE;                       // per warp
  constexpr int ACC0 = Tab::NCOLS;                           // first accumulator column
  constexpr int NCOLS = ACC0 + 4 * ACC_COLS;                 // four warps share a lane quarter
  static_assert(NCOLS <= 512, "Tensor Memory columns");

  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int grp = tid / TPF, t = tid % TPF;
  const int gbar = 1 + grp;                                  // named barrier of the group (TPF > 32)
  float2* xbuf = reinterpret_cast<float2*>(smem + a.off_xbuf) + grp * Cfg::XBUF_F2;
  uint32_t* s_taddr = reinterpret_cast<uint32_t*>(smem + a.off_acc);

  // ---- Tensor Memory: tables (as fwd_kernel, window in element order i = 0 .. 31) + accumulator columns
  if (tid < 32) tmem_alloc<NCOLS>(s_taddr);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tbase = *s_taddr + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
  if (tid < 128) {
    const int tt = tid % TPF;
    for (int col = 0; col < 64; ++col) tmem_store1(tbase + col, a.window[2 * (tt + TPF * (col >> 1)) + (col & 1)]);
    for (int sp = 1; sp < Cfg::NPASS; ++sp) {
      const int Rr = Cfg::radix(sp), p = Cfg::sublen(sp);
      for (int f = 0; f < PPT; ++f) {
        const int b = f / Rr, r = f % Rr, k = (tt + TPF * b) & (p - 1);
        const float2 w = r == 0 ? make_float2(1.0f, 0.0f) : a.tw[Cfg::tw_offset(sp) + (r - 1) * p + k];
        tmem_store1(tbase + 64 * sp + 2 * f, w.x);
        tmem_store1(tbase + 64 * sp + 2 * f + 1, w.y);
      }
    }
    for (int cp = 0; cp < NPAIR; ++cp) {
      const float2 w = a.twn[tt + TPF * cp];
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
  const uint32_t acc_base = tbase + ACC0 + (uint32_t)(tid >> 7) * ACC_COLS;   // warp index / 4 within the quarter
  const float2 wt = make_float2(0.0f, 0.0f);                 // (SmemTab only)

  // ---- this group's run of (clip, frame) pairs
  const long long total = (long long)a.n_clips * a.n_frames;
  const long long gidx = (long long)blockIdx.x * NG + grp;
  long long gpos = gidx * (long long)a.frames_per_slot;
  const long long gend = min(total, gpos + (long long)a.frames_per_slot);
  const int hop = a.hop;                                     // == 2 * M / R
  const bool vec2 = a.vec4 != 0;                             // 8-byte stores allowed (host-checked alignment)

  while (gpos < gend) {
    const int clip = (int)(gpos / a.n_frames);
    const int fs = (int)(gpos - (long long)clip * a.n_frames);
    const int fe = (int)min((long long)a.n_frames, fs + (gend - gpos));
    gpos += fe - fs;
    const float2* Dclip = a.D + (long long)clip * a.d_clip_stride;
    float* yclip = a.y + (long long)clip * a.y_clip_stride;
    uint32_t slot[NLIVE];
#pragma unroll
    for (int u = 0; u < NLIVE; ++u) slot[u] = acc_base + 16 * u;
    {
      float z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0.0f;
#pragma unroll
      for (int u = 0; u < NLIVE; ++u) tmem_st16(slot[u], z);
      tmem_wait_st();
    }
    // emit one finished unit: 8 sample pairs of this thread, output positions o0 + 2 * (t + TPF * i)
    auto emit_unit = [&](const float (&val)[16], long long o0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long o = o0 + 2 * (t + TPF * i);
        if (o >= 0 && o + 1 < a.out_len) {
          if (vec2) {
            const float2 s = __ldg(reinterpret_cast<const float2*>(a.inv_wss + o));
            *reinterpret_cast<float2*>(yclip + o) = make_float2(val[2 * i] * s.x, val[2 * i + 1] * s.y);
          } else {
            yclip[o] = val[2 * i] * __ldg(a.inv_wss + o);
            yclip[o + 1] = val[2 * i + 1] * __ldg(a.inv_wss + o + 1);
          }
        } else {
          if (o >= 0 && o < a.out_len) yclip[o] = val[2 * i] * __ldg(a.inv_wss + o);
          if (o + 1 >= 0 && o + 1 < a.out_len) yclip[o + 1] = val[2 * i + 1] * __ldg(a.inv_wss + o + 1);
        }
      }
    };

    const int fh = max(0, fs - (R - 1));                     // replayed frames rebuild the state
    auto prefetch_row = [&](int frame) {
      const char* row = reinterpret_cast<const char*>(Dclip + (long long)frame * (M + 1));
      for (int off = t * 128; off < (M + 1) * 8; off += TPF * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
    };
    const int ahead = a.acc_floats;                          // rows prefetched ahead of the one being transformed (host: 1 or 2)
    prefetch_row(fh);
    if (ahead > 1 && fh + 1 < fe) prefetch_row(fh + 1);
    for (int frame = fh; frame < fe; ++frame) {
      const float2* Drow = Dclip + (long long)frame * (M + 1);
      if (frame + ahead < fe) prefetch_row(frame + ahead);   // next row(s) -> L2 while this frame is transformed
      // ---- bin pairs -> packed spectrum Z (re/im swapped: the forward engine then computes the inverse)
      float2 v[PPT];
      tab.begin_unmix();
      static_for<0, NPAIR>([&](auto C) {
        constexpr int c = decltype(C)::value;
        const int k = t + TPF * c;
        float2 xa = __ldg(Drow + k), xb = __ldg(Drow + M - k);
        if (k == 0) { xa.y = 0.0f; xb.y = 0.0f; }            // irfft ignores Im of DC and Nyquist
        const float2 w = tab.template unmix<c>(wt);
        float2 A, B;
        c2r_pair(xa, xb, w, A, B);
        constexpr int sa = pass0_slot_of_pair<Cfg>(c);
        static_assert(sa >= 0, "lower-half element must be a pass-0 operand of the same thread");
        v[sa] = make_float2(A.y, A.x);
        if (k != 0) xbuf[xphys(M - k)] = make_float2(B.y, B.x);
      });
      if (t == 0) {
        float2 xc = __ldg(Drow + M / 2), A, B;
        c2r_pair(xc, xc, make_float2(0.0f, -1.0f), A, B);
        xbuf[xphys(M / 2)] = make_float2(A.y, A.x);
      }
      group_sync<TPF>(gbar);
      static_for<0, PPT>([&](auto S) {
        constexpr int sl = decltype(S)::value;
        if constexpr (pass0_offset<Cfg>(sl) >= M / 2) v[sl] = xbuf[xphys(t + pass0_offset<Cfg>(sl))];
      });
      group_sync<TPF>(gbar);                                 // operands fetched before the exchange area is reused
      fft_forward_tab<Cfg, false>(v, t, gbar, xbuf, tab);
      group_sync<TPF>(gbar);                                 // last-pass reads done before the next frame's stores

      // ---- window + overlap-add, unit by unit (unit u = elements i = 8u .. 8u+7 of this thread)
      const bool own = frame >= fs;
      const long long o_frame = (long long)frame * hop - a.start;   // output index of the frame's sample 0
      tmem_wait_st();
      static_for<0, UNITS>([&](auto U) {
        constexpr int u = decltype(U)::value;
        float wv[16], val[16];
        tmem_ld16(tbase + 16 * u, wv);
        static_for<0, 8>([&](auto I) {
          constexpr int i = decltype(I)::value;
          constexpr int sl = slot_of_pair<Cfg>(8 * u + i);   // register that holds element t + TPF*(8u+i)
          static_assert(sl >= 0, "spectrum element must be register resident");
          val[2 * i] = v[sl].y * wv[2 * i];                  // un-swap: real part of the inverse is .y
          val[2 * i + 1] = v[sl].x * wv[2 * i + 1];
        });
        if constexpr (u < NLIVE) {
          float acc[16];
          tmem_ld16(slot[u], acc);
#pragma unroll
          for (int i = 0; i < 16; ++i) val[i] = acc[i] + val[i];
        }
        if constexpr (u < ADV) {
          if (own) emit_unit(val, o_frame + (long long)u * (2 * M / UNITS));
        } else {
          tmem_st16(slot[u - ADV], val);   // live unit u - ADV of the next frame (the last ADV units start fresh)
        }
      });
    }
    // ---- tail of the clip: the units after the last frame, then zero fill up to out_len
    if (fe == a.n_frames) {
      tmem_wait_st();
      const long long o_next = (long long)fe * hop - a.start;
#pragma unroll
      for (int u = 0; u < NLIVE; ++u) {
        float acc[16];
        tmem_ld16(slot[u], acc);
        emit_unit(acc, o_next + (long long)u * (2 * M / UNITS));
      }
      for (long long o = o_next + (long long)NLIVE * (2 * M / UNITS) + t; o < a.out_len; o += TPF)
        if (o >= 0) yclip[o] = 0.0f;
    }
  }

  tmem_fence_before_sync();
  __syncthreads();
  if (tid < 32) {
    tmem_fence_after_sync();
    tmem_free<NCOLS>(tab.taddr);
  }
}

}  // namespace b2l
