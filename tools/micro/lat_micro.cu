// dependent-latency probes for the instructions on the generation ring's critical path (timing experiment)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__global__ void k(long long* out, uint32_t* sink, int iters) {
  __shared__ uint32_t sh[256];
  uint32_t x = threadIdx.x * 2654435761u;
  sh[threadIdx.x] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %0;" : "+r"(x));
  long long t1 = clock64();
  // STS -> BAR -> LDS round trip (one exchange phase)
  for (int i = 0; i < iters; ++i) {
    sh[threadIdx.x] = x;
    __syncthreads();
    x += sh[(threadIdx.x + 32) & 255];
  }
  long long t2 = clock64();
  // same with a WAR barrier too (what a single-buffered exchange costs)
  for (int i = 0; i < iters; ++i) {
    sh[threadIdx.x] = x;
    __syncthreads();
    x += sh[(threadIdx.x + 32) & 255];
    __syncthreads();
  }
  long long t3 = clock64();
  float f = __uint_as_float(x & 0x3fffffffu);
  for (int i = 0; i < iters; ++i) { __half2 h = __floats2half2_rn(f, f); f = __half2float(h.x) + 1.f; }
  long long t4 = clock64();
  for (int i = 0; i < iters; ++i) x = __shfl_xor_sync(0xffffffffu, x, 4) + 1;
  long long t5 = clock64();
  for (int i = 0; i < iters; ++i) x = __reduce_max_sync(0xffffffffu, x) + threadIdx.x;
  long long t6 = clock64();
  for (int i = 0; i < iters; ++i) x = __reduce_max_sync(0x11111111u << (threadIdx.x & 3), x) + threadIdx.x;
  long long t7 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5; out[6] = t7 - t6; }
  sink[threadIdx.x] = x + (uint32_t)f;
}
int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 64); cudaMalloc(&s, 4096);
  long long h[7];
  const int it = 1000;
  k<<<1, 256>>>(d, s, it); k<<<1, 256>>>(d, s, it);
  cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
  printf("movmatrix %.1f  sts+bar+lds %.1f  sts+bar+lds+bar %.1f  cvt.f16x2+cvt+fadd %.1f  shfl+iadd %.1f  redux(full)+iadd %.1f  redux(partial)+iadd %.1f   %s\n",
         (double)h[0] / it, (double)h[1] / it, (double)h[2] / it, (double)h[3] / it, (double)h[4] / it, (double)h[5] / it, (double)h[6] / it, cudaGetErrorString(cudaDeviceSynchronize()));
}
