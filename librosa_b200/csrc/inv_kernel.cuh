// inv_kernel.cuh — fused inverse real FFT + window + gather-form overlap-add + WOLA normalisation.
//
// Replaces librosa.istft's per-block  win * scipy.fft.irfft(D)  (librosa/core/spectrum.py:566, :598),
// the numba __overlap_add loop (:629-643) and the window-sum-square division (:606-624).
//
// A half-CTA (see DUAL in fwd_kernel.cuh) owns one slot of consecutive (clip, frame) pairs and walks it clip
// by clip ("pieces").  Inside a piece it takes its frames G at
// a time (one frame per thread group): each group rebuilds the packed half-length spectrum from bin
// pairs (c2r_pair) — the lower half lands directly in the registers that feed the first FFT pass, the
// upper half crosses shared memory to its partner thread — runs the same register FFT as the forward
// path with re/im swapped (== inverse transform), multiplies by the window (which carries 1/n_fft) and
// parks the frame in its exchange region.  The half then *gathers*: every output sample sums the frames
// that cover it in increasing frame order — the order the reference adds them — plus the carry of the
// previous rounds, and is written exactly once, already divided by the window-sum-square.  No atomics,
// bit-reproducible.
#pragma once
#include "common.cuh"
#include "fft_engine.cuh"

namespace b2l {

template <int LOG2M, int TPF, int NW, bool DUAL>
__global__ void __launch_bounds__(NW * 32, 1) inv_kernel(const InvArgs a) {
  using Cfg = FftCfg<LOG2M, TPF>;
  constexpr int M = Cfg::M, N = 2 * M, PPT = Cfg::PPT;
  constexpr int NT = NW * 32;
  constexpr int NH = DUAL ? 2 : 1;
  constexpr int HT = NT / NH;
  constexpr int G = HT / TPF;                  // frames per round of one half
  constexpr int NPAIR = PPT / 2;
  static_assert(TPF <= 32 || NH + NT / TPF <= 15, "named barriers: 1..NH for the halves, then one per frame group");

  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int half = DUAL ? tid / HT : 0;
  const int htid = tid - half * HT;
  float* s_win = reinterpret_cast<float*>(smem + a.off_win);
  float2* s_tw = reinterpret_cast<float2*>(smem + a.off_tw);
  float2* s_xall = reinterpret_cast<float2*>(smem + a.off_xbuf + half * a.xbuf_stride);
  float* s_carry = reinterpret_cast<float*>(smem + a.off_acc + half * a.acc_stride);   // 2 x clen floats

  const int grp = htid / TPF;
  const int t = htid % TPF;
  const int gbar = 1 + NH + half * G + grp;
  float2* xbuf = s_xall + grp * Cfg::XBUF_F2;
  auto half_sync = [&]() {
    if constexpr (DUAL) asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "n"(HT) : "memory");
    else __syncthreads();
  };

  for (int i = tid; i < N; i += NT) s_win[i] = a.window[i];
  for (int i = tid; i < Cfg::TW_COUNT; i += NT) s_tw[i] = a.tw[i];
  const int clen = a.n_fft > a.hop ? a.n_fft - a.hop : 0;
  __syncthreads();
  const float2 wt = __ldg(a.twn + t);                         // W_N^t, see unmix twiddle in fwd_kernel.cuh
  const int overlap = (a.n_fft + a.hop - 1) / a.hop;          // frames covering one sample
  const int hop_shift = (a.hop & (a.hop - 1)) == 0 ? 31 - __clz(a.hop) : -1;   // log2(hop) when hop is a power of two

  // Work distribution: the (clip, frame) pairs of the whole batch form one sequence of n_clips * n_frames
  // frames, cut into equal slots — one per resident half-CTA, so every SM finishes at the same time (a
  // (clip, segment) grid leaves the last wave mostly empty).  A slot is walked piece by piece: a piece is
  // the part of the slot that lies inside one clip; it starts from an empty carry and first rebuilds it
  // from the overlap - 1 halo frames before its first frame (transformed again, not emitted).
  const long long slot = (long long)blockIdx.x * NH + half;
  const long long total_frames = (long long)a.n_clips * a.n_frames;
  const long long slot_end = min(total_frames, (slot + 1) * (long long)a.frames_per_slot);
  for (long long gpos = slot * (long long)a.frames_per_slot; gpos < slot_end;) {
  const int clip = (int)(gpos / a.n_frames);
  const int fs = (int)(gpos % a.n_frames);
  const int fe = (int)min((long long)a.n_frames, fs + (slot_end - gpos));
  gpos += fe - fs;
  for (int i = htid; i < 2 * clen; i += HT) s_carry[i] = 0.0f;
  half_sync();
  const bool last_seg = (fe == a.n_frames);
  const int fh = max(0, fs - (overlap - 1));                  // halo frames rebuild the carry
  const long long emit_lo = (long long)fs * a.hop;
  const long long emit_hi = last_seg ? (1LL << 62) : (long long)fe * a.hop;

  const float2* Dclip = a.D + (long long)clip * a.d_clip_stride;
  float* yclip = a.y + (long long)clip * a.y_clip_stride;
  float* carry_cur = s_carry;
  float* carry_nxt = s_carry + clen;

  // The spectrum row of the next round's frame is pulled into L2 while this round is gathered, so the
  // operand loads at the top of the next round see L2 latency instead of HBM latency (the registers
  // are all taken by the FFT, so a register prefetch would spill).
  auto prefetch_row = [&](int first_frame) {
    const int frame = min(first_frame + grp, fe - 1);
    const char* row = reinterpret_cast<const char*>(Dclip + (long long)frame * (M + 1));
    for (int off = t * 128; off < (M + 1) * 8; off += TPF * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
  };

  int fr0 = fh;
  for (; fr0 < fe; fr0 += G) {
    // Every group transforms a frame in every round (groups past the end redo the last frame and are
    // ignored by the gather): sub-warp groups share a warp, so skipping would diverge at __syncwarp.
    const int frame = min(fr0 + grp, fe - 1);
    {
      // ---- bin pairs -> packed spectrum Z (re/im swapped for the inverse transform)
      const float2* Drow = Dclip + (long long)frame * (M + 1);
      float2 v[PPT];
      static_for<0, NPAIR>([&](auto C) {
        constexpr int c = decltype(C)::value;
        const int k = t + TPF * c;
        float2 xa = __ldg(Drow + k), xb = __ldg(Drow + M - k);
        if (k == 0) { xa.y = 0.0f; xb.y = 0.0f; }    // irfft ignores Im of DC and Nyquist
        float2 w;
        if constexpr (c == 0) w = wt;
        else w = cmul(wt, make_float2(TwC<c, 2 * PPT>::re, TwC<c, 2 * PPT>::im));
        float2 A, B;
        c2r_pair(xa, xb, w, A, B);
        constexpr int sa = pass0_slot_of_pair<Cfg>(c);
        static_assert(sa >= 0, "lower-half element must be a pass-0 operand of the same thread");
        v[sa] = make_float2(A.y, A.x);                                 // Z[k] stays in registers
        if (k != 0) xbuf[xphys(M - k)] = make_float2(B.y, B.x);        // Z[M-k] goes to its owner
      });
      if (t == 0) {
        float2 xc = __ldg(Drow + M / 2), A, B;
        c2r_pair(xc, xc, make_float2(0.0f, -1.0f), A, B);
        xbuf[xphys(M / 2)] = make_float2(A.y, A.x);
      }
      group_sync<TPF>(gbar);
      static_for<0, PPT>([&](auto S) {
        constexpr int slot = decltype(S)::value;
        if constexpr (pass0_offset<Cfg>(slot) >= M / 2) v[slot] = xbuf[xphys(t + pass0_offset<Cfg>(slot))];
      });
      group_sync<TPF>(gbar);          // operands fetched before the exchange area is overwritten
      fft_forward<Cfg>(v, t, gbar, xbuf, s_tw);
      if constexpr (Cfg::NPASS > 1) group_sync<TPF>(gbar);
      // ---- un-swap, window (carries 1/n_fft), park the frame: ybuf[j], j = 0 .. n_fft-1
      float2* ybuf = xbuf;
      static_for<0, PPT>([&](auto S) {
        constexpr int slot = decltype(S)::value;
        const int e = spectrum_index<Cfg>(t, slot);
        const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * e);
        ybuf[e] = make_float2(v[slot].y * w.x, v[slot].x * w.y);
      });
    }
    if (fr0 + G < fe) prefetch_row(fr0 + G);        // next round's rows -> L2 while this round is gathered
    half_sync();

    // ---- gather: emit positions [fr0*hop, (fr0+G)*hop), then rebuild the carry.
    // Position x = c*hop + i is covered by the round's frames g = c-q .. c (clipped to [0, ng)), where
    // q = (n_fft-1-i)/hop depends only on the offset i inside the hop; frames are added in increasing g,
    // the order in which the reference accumulates them (librosa/core/spectrum.py:629-643).
    const int ng = min(G, fe - fr0);                 // valid frames in this round
    const int emit_n = G * a.hop;
    const int n_chunks = G + (clen + a.hop - 1) / a.hop;      // chunks of hop positions incl. the carry zone
    const int o0 = (int)((long long)fr0 * a.hop - a.start);   // output index of x = 0 (fits: out_len < 2^31)
    // chunks whose positions this segment owns: emit_lo <= u < emit_hi
    const int c_lo = (int)max(0LL, (emit_lo - (long long)fr0 * a.hop + a.hop - 1) / a.hop);
    const int c_hi = (int)min((long long)G, (emit_hi - (long long)fr0 * a.hop) / a.hop);
    // Work items are (chunk c, offset i) pairs spread over every thread of the half.
    if (a.vec4) {
      // 4 consecutive samples per item: 16-byte shared loads, one 16-byte store (hop, n_fft - hop,
      // start and the row strides are multiples of 4 and the buffers 16-byte aligned; host-checked)
      const int per = a.hop >> 2;
      for (int item = htid; item < per * n_chunks; item += HT) {
        // (chunk, offset) of the item and the number of earlier frames covering the offset: shifts when the
        // hop is a power of two (the usual case) instead of two integer divisions per item
        int c, q;
        if (hop_shift >= 2) c = item >> (hop_shift - 2);
        else c = item / per;
        const int i = (item - c * per) << 2;
        const int x = c * a.hop + i;
        if (x >= emit_n + clen) continue;
        if (i >= a.n_fft) q = -1;
        else if (hop_shift >= 0) q = (a.n_fft - 1 - i) >> hop_shift;
        else q = (a.n_fft - 1 - i) / a.hop;
        float4 val = (x < clen) ? *reinterpret_cast<const float4*>(carry_cur + x) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int g_lo = max(0, c - q), g_hi = min(c, ng - 1);
        const float* yb = reinterpret_cast<const float*>(s_xall + g_lo * Cfg::XBUF_F2) + (c - g_lo) * a.hop + i;
        for (int g = g_lo; g <= g_hi; ++g) {
          const float4 f = *reinterpret_cast<const float4*>(yb);
          val.x += f.x; val.y += f.y; val.z += f.z; val.w += f.w;
          yb += 2 * Cfg::XBUF_F2 - a.hop;          // next frame's buffer, one hop earlier inside it
        }
        if (x < emit_n) {
          const int o = o0 + x;
          if (c >= c_lo && c < c_hi && o + 3 >= 0 && o < a.out_len) {
            if (o >= 0 && o + 3 < a.out_len) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.inv_wss + o));
              *reinterpret_cast<float4*>(yclip + o) = make_float4(val.x * s4.x, val.y * s4.y, val.z * s4.z, val.w * s4.w);
            } else {
              const float vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (o + e >= 0 && o + e < a.out_len) yclip[o + e] = vv[e] * __ldg(a.inv_wss + o + e);
            }
          }
        } else {
          *reinterpret_cast<float4*>(carry_nxt + (x - emit_n)) = val;
        }
      }
    } else {
      for (int item = htid; item < a.hop * n_chunks; item += HT) {
        const int c = hop_shift >= 0 ? item >> hop_shift : item / a.hop;
        const int i = item - c * a.hop;
        const int x = c * a.hop + i;
        if (x >= emit_n + clen) continue;
        int q;                                                        // hop > n_fft: gap positions see no frame
        if (i >= a.n_fft) q = -1;
        else if (hop_shift >= 0) q = (a.n_fft - 1 - i) >> hop_shift;
        else q = (a.n_fft - 1 - i) / a.hop;
        float val = (x < clen) ? carry_cur[x] : 0.0f;
        const int g_lo = max(0, c - q), g_hi = min(c, ng - 1);
        const float* yb = reinterpret_cast<const float*>(s_xall + g_lo * Cfg::XBUF_F2) + (c - g_lo) * a.hop + i;
        for (int g = g_lo; g <= g_hi; ++g) {
          val += *yb;
          yb += 2 * Cfg::XBUF_F2 - a.hop;
        }
        if (x < emit_n) {
          const int o = o0 + x;
          if (c >= c_lo && c < c_hi && o >= 0 && o < a.out_len) yclip[o] = val * __ldg(a.inv_wss + o);
        } else {
          carry_nxt[x - emit_n] = val;
        }
      }
    }
    half_sync();
    float* tmp = carry_cur; carry_cur = carry_nxt; carry_nxt = tmp;
  }
  // ---- flush: samples past the last round that only the carry reaches, then zero-fill the rest
  if (last_seg) {
    const long long u0 = (long long)fr0 * a.hop;     // fr0 == first frame index past the last round
    for (int x = htid; x < clen; x += HT) {
      const long long o = u0 + x - a.start;
      if (o >= 0 && o < a.out_len) yclip[o] = carry_cur[x] * __ldg(a.inv_wss + o);
    }
    for (long long o = u0 + clen - a.start + htid; o < a.out_len; o += HT)
      if (o >= 0) yclip[o] = 0.0f;
  }
  half_sync();   // the piece's last carry reads are done before the next piece clears the carry
  }  // pieces of this slot
}

}  // namespace b2l
