// Throughput of the sm_100 packed FP32 instructions (FFMA2 / FADD2 / FMUL2) against their scalar forms.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
// Prints, per variant, the warp-instruction issue rate per SM sub-partition per clock and the FP32 lane rate.
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ACC = 12;       // independent dependency chains per thread
constexpr int ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) rate(float2* out, float2 c, float2 d) {
  float2 acc[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
      if constexpr (MODE == 0) {          // scalar FFMA x2
        acc[i].x = fmaf(acc[i].x, c.x, d.x);
        acc[i].y = fmaf(acc[i].y, c.y, d.y);
      } else if constexpr (MODE == 1) {   // FFMA2
        acc[i] = __ffma2_rn(acc[i], c, d);
      } else if constexpr (MODE == 2) {   // scalar FADD x2
        acc[i].x = acc[i].x + d.x;
        acc[i].y = acc[i].y + d.y;
      } else if constexpr (MODE == 3) {   // FADD2
        acc[i] = __fadd2_rn(acc[i], d);
      } else if constexpr (MODE == 4) {   // FADD2 with the swap / negate swizzle (the -i butterfly)
        acc[i] = __fadd2_rn(d, make_float2(acc[i].y, -acc[i].x));
      } else if constexpr (MODE == 5) {   // FFMA2 with a broadcast immediate
        acc[i] = __ffma2_rn(acc[i], make_float2(0.92387953f, 0.92387953f), d);
      } else if constexpr (MODE == 6) {   // FMUL2
        acc[i] = __fmul2_rn(acc[i], c);
      } else {                            // scalar FMUL x2
        acc[i].x = acc[i].x * c.x;
        acc[i].y = acc[i].y * c.y;
      }
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < ACC; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float2* out, int sms, double clk_ghz) {
  const int blocks = sms * 8;
  rate<MODE><<<blocks, 256>>>(out, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) rate<MODE><<<blocks, 256>>>(out, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f));
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  const double lane_ops = double(blocks) * 256 * ITER * ACC * 2;           // FP32 lane operations
  const bool packed = (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6);
  const double warp_instr = lane_ops / 32 / (packed ? 2 : 1);
  const double cycles = ms * 1e-3 * clk_ghz * 1e9;
  printf("%-28s %8.3f ms  %7.2f T lane-ops/s  %5.3f warp-instr/clk/sub-partition  %6.1f lanes/clk/SM\n", name, ms,
         lane_ops / ms * 1e-9, warp_instr / cycles / (sms * 4), lane_ops / cycles / sms);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz * 1e-6;
  printf("%s, %d SMs, nominal %.3f GHz (rates below assume this clock)\n", p.name, p.multiProcessorCount, ghz);
  float2* out; cudaMalloc(&out, size_t(p.multiProcessorCount) * 8 * 256 * sizeof(float2));
  run<0>("FFMA  x2 (scalar)", out, p.multiProcessorCount, ghz);
  run<1>("FFMA2", out, p.multiProcessorCount, ghz);
  run<5>("FFMA2 broadcast immediate", out, p.multiProcessorCount, ghz);
  run<2>("FADD  x2 (scalar)", out, p.multiProcessorCount, ghz);
  run<3>("FADD2", out, p.multiProcessorCount, ghz);
  run<4>("FADD2 swap+negate swizzle", out, p.multiProcessorCount, ghz);
  run<7>("FMUL  x2 (scalar)", out, p.multiProcessorCount, ghz);
  run<6>("FMUL2", out, p.multiProcessorCount, ghz);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
