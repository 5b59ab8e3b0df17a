// mma.sync m16n8k16 f16->f32 latency / issue rate on sm_100a (timing experiment, not part of the product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k(long long* out, float* sink, int iters) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3, 7, 9};
  float c[CHAINS][4] = {};
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) mma(c[j], a, 0x3c003c00u, 0x3c003c00u);
  }
  float s = 0;
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = s;
}
// fadd dependent latency, lds latency, bar.sync latency
__global__ void k2(long long* out, float* sink, int iters) {
  __shared__ float sh[1024];
  sh[threadIdx.x] = threadIdx.x;
  __syncthreads();
  float x = sink[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = x + 1.0f;
  long long t1 = clock64();
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) idx = (int)sh[idx & 1023] & 1023;
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t3 = clock64();
  for (int i = 0; i < iters; ++i) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); x = y; }
  long long t4 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; }
  sink[threadIdx.x] = x + idx;
}
int main() {
  long long* d; float* s;
  cudaMalloc(&d, 64); cudaMalloc(&s, 4096); cudaMemset(s, 0, 4096);
  long long h[4];
  const int it = 1000;
  for (int threads : {32, 128, 256, 512}) {
#define RUN(C) k<C><<<1, threads>>>(d, s, it); k<C><<<1, threads>>>(d, s, it); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); \
    printf("threads %d chains %d: %.1f cycles per mma per warp-chain step, %.2f cycles per mma per SMSP\n", threads, C, (double)h[0] / it, (double)h[0] / it / C / ((threads + 127) / 128));
    RUN(1) RUN(2) RUN(4) RUN(8)
  }
  k2<<<1, 256>>>(d, s, it); cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("fadd %.1f  lds(+cvt) %.1f  bar.sync(256) %.1f  tanh %.1f\n", (double)h[0] / it, (double)h[1] / it, (double)h[2] / it, (double)h[3] / it);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
