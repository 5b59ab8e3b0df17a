// tcgen05.mma issue rate by operand layout: a chain of M=128 x N x K=16 bf16 UMMAs on one accumulator, operands in shared
// memory (128B-swizzled tiles, contents irrelevant), K-major vs MN-major A and B.  One CTA per SM on every SM (the rate that
// matters is the one with the whole chip busy).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/umma_rate_bench
// tools/umma_rate_bench.cu ; run under gpurun.  Every wait is bounded.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../music_b200/csrc/tc05.cuh"
using namespace tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// mode bit 0: A MN-major, bit 1: B MN-major
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int n_mma, long long* __restrict__ cycles, int* __restrict__ flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8 * 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // finite bf16 values
  if (tid == 0) {
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s, sbase = smem_u32(sm);
  if (tid == 0) {
    const int a_mn = mode & 1, b_mn = (mode >> 1) & 1;
    const uint32_t idesc = idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t sa = sbase, sb = sbase + 4 * 16384;          // four adjacent tiles each (MN-major: N = 256 = 4 x 64 columns)
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3;                                       // cycle through the 4 (K-major) / first 4 of 8 (MN-major) k-steps
      const uint64_t ad = a_mn ? desc_mnmajor(sa, k, 16384) : desc_kmajor(sa, k);
      const uint64_t bd = b_mn ? desc_mnmajor(sb, k, 16384) : desc_kmajor(sb, k);
      umma_bf16(tmem, ad, bd, idesc, i > 0);
    }
    umma_commit(&done);
    bool ok = false;
    for (int i = 0; i < (1 << 24); ++i)
      if (mbar_test_wait(&done, 0)) { ok = true; break; }
    const long long t1 = clock64();
    if (!ok) atomicMax(flag, 1);
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int N>
int run(int sms, long long* d_cyc, int* d_flag) {
  const int smem = 8 * 16384 + 1024, n_mma = 2048;
  CK(cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[4] = {"A K-major,  B K-major ", "A MN-major, B K-major ", "A K-major,  B MN-major", "A MN-major, B MN-major"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<N><<<sms, 128, smem>>>(mode, n_mma, d_cyc, d_flag);
      CK(cudaDeviceSynchronize());
    }
    long long h[256];
    int flag = 0;
    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    long long mx = 0, mn = 1ll << 60;
    for (int i = 0; i < sms; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    const double per = (double)mx / n_mma, flops = 2.0 * 128 * N * 16;
    printf("M=128 N=%3d K=16  %s : %7.1f cycles per MMA (min CTA %.1f) = %6.0f FLOP/clk/SM%s\n", N, names[mode], per, (double)mn / n_mma,
           flops / per, flag ? "  [TIMEOUT]" : "");
  }
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  long long* d_cyc;
  int* d_flag;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 256));
  CK(cudaMalloc(&d_flag, sizeof(int)));
  CK(cudaMemset(d_flag, 0, sizeof(int)));
  printf("%s, %d SMs, all SMs busy\n", prop.name, sms);
  if (run<128>(sms, d_cyc, d_flag)) return 1;
  if (run<64>(sms, d_cyc, d_flag)) return 1;
  if (run<256>(sms, d_cyc, d_flag)) return 1;
  return 0;
}
