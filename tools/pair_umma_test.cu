// Exploration for round 2: a CTA PAIR (cta_group::2) computing D[256 x 128] = A[256 x 64] B[128 x 64]^T with one
// tcgen05.mma stream issued by the leader CTA.  Each CTA loads its own 128 rows of A and HALF of B (64 rows), so the B
// traffic per SM halves - the lever for the L2 -> SM-bound GEMMs of the head.  Exact arithmetic
// (small integers in bf16), every wait bounded.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/pair_umma_test
// tools/pair_umma_test.cu -lcuda ; run under gpurun.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../music_b200/csrc/tc05.cuh"
using namespace tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool bounded_wait(uint64_t* bar, uint32_t parity, int* flag, int code) {
  for (int i = 0; i < (1 << 22); ++i)
    if (mbar_test_wait(bar, parity)) return true;
  atomicMax(flag, code);
  return false;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int* __restrict__ flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full, peer_ready, mma_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(&full, 1);
    mbar_init(&peer_ready, 1);
    mbar_init(&mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint8_t* sA = sm;                 // [128 rows][64 k] SW128
  uint8_t* sB = sm + 16384;         // [64 rows of this CTA's half of N][64 k] SW128
  if (tid == 0) {
    mbar_expect_tx(&full, 16384 + 8192);
    tma_load_2d(sA, &tmA, &full, 0, 128 * (int)rank);
    tma_load_2d(sB, &tmB, &full, 0, 64 * (int)rank);
    if (bounded_wait(&full, 0, flag, 1)) {
      if (rank == 1) {              // tell the leader that this CTA's operands are in place
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(&peer_ready)));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
      } else if (bounded_wait(&peer_ready, 0, flag, 2)) {
        tc_fence_after();
        // instruction descriptor: bf16 x bf16 -> fp32, K-major operands, N = 128, M = 256 (the pair)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = desc_kmajor(smem_u32(sA), k), db = desc_kmajor(smem_u32(sB), k);
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
              "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(k > 0))
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(&mma_done)),
                     "h"((uint16_t)3)
                     : "memory");
      }
    }
  }
  __syncthreads();
  if (bounded_wait(&mma_done, 0, flag, 3)) {
    tc_fence_after();
    const int row = 128 * (int)rank + warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_addr(tmem, warp * 32, c * 32), v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(int64_t)row * 128 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeFn enc, CUtensorMap* m, void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  return 0;
}

int main() {
  CK(cudaSetDevice(0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
  std::vector<__nv_bfloat16> hA(256 * 64), hB(128 * 64);
  for (int i = 0; i < 256 * 64; ++i) hA[i] = __float2bfloat16((float)((i * 7 + i / 64) % 5 - 2));
  for (int i = 0; i < 128 * 64; ++i) hB[i] = __float2bfloat16((float)((i * 3 + i / 64) % 7 - 3));
  __nv_bfloat16 *A, *B;
  float* D;
  int* flag;
  CK(cudaMalloc(&A, hA.size() * 2)); CK(cudaMalloc(&B, hB.size() * 2)); CK(cudaMalloc(&D, 256 * 128 * 4)); CK(cudaMalloc(&flag, 4));
  CK(cudaMemcpy(A, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(D, 0xFF, 256 * 128 * 4)); CK(cudaMemset(flag, 0, 4));
  CUtensorMap tmA, tmB;
  if (make_map(enc, &tmA, A, 64, 256, 128) || make_map(enc, &tmB, B, 64, 128, 64)) return 1;
  const int smem = 16384 + 8192 + 1024;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pair_kernel<<<2, 128, smem>>>(tmA, tmB, D, flag);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int hflag = 0;
  std::vector<float> hD(256 * 128);
  CK(cudaMemcpy(&hflag, flag, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hD.data(), D, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 128; ++n) {
      float ref = 0;
      for (int k = 0; k < 64; ++k) ref += __bfloat162float(hA[m * 64 + k]) * __bfloat162float(hB[n * 64 + k]);
      const double e = fabs((double)hD[m * 128 + n] - ref);
      if (!(e <= 0)) { if (bad < 5) printf("mismatch D[%d][%d] = %g, expected %g\n", m, n, hD[m * 128 + n], ref); ++bad; }
      if (e > maxerr) maxerr = e;
    }
  printf("pair UMMA (cta_group::2, M=256 N=128 K=64): timeout flag %d, mismatches %d, max |err| %g -> %s\n", hflag, bad, maxerr,
         (hflag == 0 && bad == 0) ? "EXACT" : "FAILED");
  return (hflag == 0 && bad == 0) ? 0 : 2;
}
