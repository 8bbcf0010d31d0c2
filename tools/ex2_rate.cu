// MUFU.EX2 throughput on sm_100a: fp32 ex2.approx vs packed ex2.approx.f16x2 / bf16x2 (two exponentials per
// instruction?), plus the cvt.rn.f16x2.f32 that feeds the packed form.  Prints exponentials per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ex2_rate tools/ex2_rate.cu && tools/ex2_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kIlp = 8;

template <int MODE>
__global__ void __launch_bounds__(512) rate_kernel(float* out, float seed) {
  float x[kIlp];
  uint32_t h[kIlp];
  for (int i = 0; i < kIlp; ++i) { x[i] = seed * (threadIdx.x + i) * 1e-6f - 1.f; h[i] = 0xb800b800u + threadIdx.x + i; }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kIlp; ++i) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 3) {   // cvt pair + packed exp (the softmax inner step)
        asm volatile("{ cvt.rn.f16x2.f32 %0, %1, %1; ex2.approx.f16x2 %0, %0; }" : "=r"(h[i]) : "f"(x[i]));
        x[i] += __uint_as_float(h[i] & 0x007fffffu);
      } else {                  // fp32 exp + cvt pack of two results (today's inner step, per pair: 2 ex2 + 1 cvt)
        float a = x[i], b = x[i] * 0.5f;
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a), "f"(b));
        x[i] += __uint_as_float(h[i] & 0x007fffffu);
      }
    }
  }
  float acc = 0.f;
  for (int i = 0; i < kIlp; ++i) acc += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run(const char* name, int exps_per_op, int sms, float mhz) {
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 2 * 512);
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  rate_kernel<MODE><<<sms * 2, 512>>>(out, 1.f);
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  rate_kernel<MODE><<<sms * 2, 512>>>(out, 1.f);
  cudaEventRecord(e);
  cudaEventSynchronize(e);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, s, e);
  const double ops = (double)sms * 2 * 512 * kIters * kIlp * exps_per_op;
  printf("%-34s %8.3f ms  %6.2f exponentials / clk / SM (at %.0f MHz)\n", name, ms, ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
  cudaFree(out);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float mhz = khz / 1000.f;
  run<0>("ex2.approx.f32", 1, prop.multiProcessorCount, mhz);
  run<1>("ex2.approx.f16x2", 2, prop.multiProcessorCount, mhz);
  run<2>("ex2.approx.bf16x2", 2, prop.multiProcessorCount, mhz);
  run<3>("cvt.f16x2 + ex2.f16x2 (per pair)", 2, prop.multiProcessorCount, mhz);
  run<4>("2 x ex2.f32 + cvt.f16x2 (per pair)", 2, prop.multiProcessorCount, mhz);
  return 0;
}
