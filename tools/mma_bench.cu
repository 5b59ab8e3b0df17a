// mma.sync.m16n8k16 bf16 on sm_100a: latency of a dependent chain and issue rate with independent accumulators.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&c)[4], const uint4& a, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
template <int NACC>
__global__ void k(float* out, int iters, long long* cyc) {
  float c[NACC][4] = {};
  uint4 a = make_uint4(threadIdx.x, 1, 2, 3);
  unsigned b0 = threadIdx.x, b1 = 7;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma(c[j], a, b0, b1);
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC>
void run(int threads) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<NACC><<<148, threads>>>(out, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("warps/SM %2d, independent accumulators %d: %.1f cycles per MMA per warp, %.2f MMA/clk/SM\n", threads / 32, NACC,
         (double)h / (iters * NACC), (double)iters * NACC * (threads / 32) / h);
}
int main() {
  run<1>(32); run<2>(32); run<4>(32); run<8>(32);
  run<1>(128); run<4>(128); run<1>(256); run<2>(256); run<4>(256); run<8>(256); run<4>(512);
  return 0;
}
