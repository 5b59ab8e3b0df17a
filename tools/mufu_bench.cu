// Throughput of tanh.approx in f32, f16x2 and bf16x2 form (elements per clock per SM), to decide the gate epilogue's arithmetic.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  unsigned ua = __float_as_uint(a), ub = __float_as_uint(b), uc = __float_as_uint(c), ud = __float_as_uint(d);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d));
    } else if (MODE == 1) {
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ua));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ub));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(uc));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ud));
    } else {
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ua));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ub));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(uc));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ud));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ua ^ ub ^ uc ^ ud);
}
template <int MODE>
void run(const char* name, int per) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  k<MODE><<<148, 1024>>>(out, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double elems = 1024.0 * iters * 4 * per, clocks = ms * 1e-3 * khz * 1e3;
  printf("%-22s %.3f ms  %.1f elements/clk/SM (at %d MHz nominal)\n", name, ms, elems / clocks, khz / 1000);
}
int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("tanh.approx.f16x2", 2);
  run<2>("tanh.approx.bf16x2", 2);
  return 0;
}
