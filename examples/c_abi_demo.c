/* c_abi_demo.c — libb2l.so used from plain C, no Python: what a non-Python host (or a C extension of librosa)
 * would do.  Computes the STFT, the magnitude spectrogram and the spectral centroid of a 1 kHz tone and
 * checks them against closed-form answers.
 *
 *   gcc -std=c99 -I include examples/c_abi_demo.c -L librosa_b200/csrc -lb2l -lm -o /tmp/c_abi_demo
 *   LD_LIBRARY_PATH=librosa_b200/csrc /tmp/c_abi_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b2l.h"

#define CHECK(call)                                                            \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ != B2L_OK) {                                                       \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, b2l_last_error());   \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main(void) {
  const int sr = 22050, n = 22050, n_fft = 2048, hop = 512, bins = 1 + n_fft / 2;
  const double pi = 3.14159265358979323846, tone = 1000.0;
  float* y = (float*)malloc(sizeof(float) * n);
  double* window = (double*)malloc(sizeof(double) * n_fft);
  float* freqs = (float*)malloc(sizeof(float) * bins);
  for (int i = 0; i < n; ++i) y[i] = (float)(0.5 * sin(2.0 * pi * tone * i / sr));
  for (int i = 0; i < n_fft; ++i) window[i] = 0.5 - 0.5 * cos(2.0 * pi * i / n_fft);   /* periodic Hann */
  for (int k = 0; k < bins; ++k) freqs[k] = (float)((double)k * sr / n_fft);

  b2l_ctx* ctx = NULL;
  CHECK(b2l_ctx_create(0, &ctx));
  b2l_plan_desc desc;
  desc.n_fft = n_fft;
  desc.hop_length = hop;
  desc.center = 1;
  desc.pad_mode = B2L_PAD_CONSTANT;
  desc.h_window = window;
  desc.n_mels = 0;
  desc.h_mel_basis = NULL;
  desc.power = 1.0f;
  desc.n_mfcc = 0;
  desc.h_dct_basis = NULL;
  desc.amin = 1e-10f;
  desc.ref_value = 1.0f;
  desc.top_db = 80.0f;
  b2l_plan* plan = NULL;
  CHECK(b2l_plan_create(ctx, &desc, &plan));
  int64_t T = 0;
  CHECK(b2l_plan_n_frames(plan, n, &T));

  void *d_y = NULL, *d_D = NULL, *d_S = NULL, *d_f = NULL, *d_stats = NULL;
  CHECK(b2l_alloc(ctx, sizeof(float) * n, &d_y));
  CHECK(b2l_alloc(ctx, sizeof(float) * 2 * bins * T, &d_D));
  CHECK(b2l_alloc(ctx, sizeof(float) * bins * T, &d_S));
  CHECK(b2l_alloc(ctx, sizeof(float) * bins, &d_f));
  CHECK(b2l_alloc(ctx, sizeof(float) * B2L_N_STATS * T, &d_stats));
  CHECK(b2l_status_reset(ctx));
  CHECK(b2l_h2d(ctx, d_y, y, sizeof(float) * n));
  CHECK(b2l_h2d(ctx, d_f, freqs, sizeof(float) * bins));
  CHECK(b2l_stft(ctx, plan, (const float*)d_y, 1, n, n, d_D));
  CHECK(b2l_spectrogram(ctx, plan, (const float*)d_y, 1, n, n, (float*)d_S));
  b2l_stats_desc sd;
  sd.roll_percent = 0.85f;
  sd.flat_amin = 1e-10f;
  sd.flat_power = 2.0f;
  sd.bw_p = 2.0f;
  sd.bw_norm = 1;
  sd.frame_length = n_fft;
  sd.want = 0;
  CHECK(b2l_spectral_stats(ctx, plan, &sd, (const float*)d_y, 1, n, n, (const float*)d_f, (float*)d_stats));

  float* D = (float*)malloc(sizeof(float) * 2 * bins * T);
  float* S = (float*)malloc(sizeof(float) * bins * T);
  float* stats = (float*)malloc(sizeof(float) * B2L_N_STATS * T);
  CHECK(b2l_d2h(ctx, D, d_D, sizeof(float) * 2 * bins * T));
  CHECK(b2l_d2h(ctx, S, d_S, sizeof(float) * bins * T));
  CHECK(b2l_d2h(ctx, stats, d_stats, sizeof(float) * B2L_N_STATS * T));
  int flag = 0;
  CHECK(b2l_status_read(ctx, &flag));            /* synchronises the stream */

  /* middle frame: |X| peaks at the bin nearest 1 kHz with height ~ amplitude * sum(window) / 2 = 0.5 * 1024 / 2 */
  const int64_t t = T / 2;
  int peak = 0;
  for (int k = 1; k < bins; ++k)
    if (S[t * bins + k] > S[t * bins + peak]) peak = k;
  const double mag = hypot(D[2 * (t * bins + peak)], D[2 * (t * bins + peak) + 1]);
  const double centroid = stats[0 * T + t], rms = stats[4 * T + t];
  printf("frames=%lld peak_bin=%d (%.1f Hz) |X|=%.2f S=%.2f centroid=%.1f Hz rms=%.4f finite_flag=%d\n", (long long)T, peak,
         peak * (double)sr / n_fft, mag, S[t * bins + peak], centroid, rms, flag);
  int ok = (abs(peak - (int)lround(tone * n_fft / sr)) <= 1) && fabs(mag - S[t * bins + peak]) < 1e-3 * mag &&
           mag > 200.0 && mag < 260.0 && fabs(centroid - tone) < 15.0 && flag == 0;
  /* Parseval with the Hann window: rms(S) = amplitude / sqrt(2) * sqrt(mean(window^2)) = 0.5 / 1.4142 * 0.6124 */
  ok = ok && fabs(rms - 0.5 / sqrt(2.0) * sqrt(0.375)) < 2e-3;

  b2l_free(ctx, d_y); b2l_free(ctx, d_D); b2l_free(ctx, d_S); b2l_free(ctx, d_f); b2l_free(ctx, d_stats);
  b2l_plan_destroy(plan);
  b2l_ctx_destroy(ctx);
  free(y); free(window); free(freqs); free(D); free(S); free(stats);
  puts(ok ? "C ABI demo: OK" : "C ABI demo: FAILED");
  return ok ? 0 : 2;
}
