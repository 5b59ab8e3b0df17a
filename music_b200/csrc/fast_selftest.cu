// fast_selftest.cu - wn_selftest_umma(): exercises the TMA / UMMA / TMEM building blocks of tc05.cuh in
// exactly the operand layouts the kernels rely on, against an exact SIMT product (inputs are small
// multiples of 1/4 so fp32 sums are exact: expected max error is 0).  Test facility only: it is the one
// place in the library that allocates device memory itself.
//   case 0  K-major A[128x64] x K-major B[128x64]^T           (block UMMA #1 shape, one tap)
//   case 1  K = 128 in two chunks, N = 256                     (skip/head GEMM shape)
//   case 2  A tile written by threads with the manual 128B swizzle (epilogue -> next UMMA operand)
//   case 3  MN-major A (two [128 t x 64] tiles, LBO) and MN-major B: D = A^T B over 128 time rows (wgrad)
//   case 4  TMA store of a swizzled tile                        (round trip)
//   case 5  M128 N64 K64 K-major                                (dense / dgrad shape)
//   case 6  K-major A x MN-major B ([K rows][N cols] weights used untransposed)
//   case 7  3-D TMA load with a negative row coordinate (zero fill, no bleed from the previous batch)
//   case 8  3-D TMA store with rows clipped at the end of a batch (negative store coordinates TRAP: never used)
#include "fast.cuh"
#include "fast_layout.cuh"
#include "tc05.cuh"

namespace wn {
using namespace tc;
namespace {

constexpr int NCASE = 9;

__global__ void fill_kernel(__nv_bfloat16* p, int64_t n, uint32_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = __float2bfloat16((float)((int)(h % 9) - 4) * 0.25f);
  }
}

// operands in global memory:  A: [256 rows][128 cols] bf16, B: [256 rows][128 cols] bf16 (row-major, pitch 128)
// out: fp32 [128][256]
__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB256,
                const __grid_constant__ CUtensorMap tmB64, const __grid_constant__ CUtensorMap tmOut,
                const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmOut3,
                const __nv_bfloat16* __restrict__ gA, int mode, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_ld, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA0 = sm;                 // 16K
  uint8_t* sA1 = sm + 16384;         // 16K
  uint8_t* sB0 = sm + 32768;         // 32K
  uint8_t* sB1 = sm + 65536;         // 32K
  if (tid == 0) {
    mbar_init(&bar_ld, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a0 = smem_u32(sA0), a1 = smem_u32(sA1), b0 = smem_u32(sB0), b1 = smem_u32(sB1);
  int N = 128;

  if (mode == 2) {
    // manual swizzled write of A[0:128, 0:64]
    for (int c = 0; c < 8; ++c) {
      uint4 v = *reinterpret_cast<const uint4*>(gA + (int64_t)tid * 128 + c * 8);
      *reinterpret_cast<uint4*>(sA0 + sw128_chunk(tid, c)) = v;
    }
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (tid == 0) {
    switch (mode) {
      case 0: case 2:
        mbar_expect_tx(&bar_ld, (mode == 0 ? 16384 : 0) + 16384);
        if (mode == 0) tma_load_2d(sA0, &tmA, &bar_ld, 0, 0);
        tma_load_2d(sB0, &tmB64, &bar_ld, 0, 0);          // B rows 0..63
        tma_load_2d(sB0 + 8192, &tmB64, &bar_ld, 0, 64);  // B rows 64..127 (contiguous 8-row groups)
        break;
      case 1:
        mbar_expect_tx(&bar_ld, 2 * 16384 + 2 * 32768);
        tma_load_2d(sA0, &tmA, &bar_ld, 0, 0);
        tma_load_2d(sA1, &tmA, &bar_ld, 64, 0);
        tma_load_2d(sB0, &tmB256, &bar_ld, 0, 0);
        tma_load_2d(sB1, &tmB256, &bar_ld, 64, 0);
        break;
      case 3:
        mbar_expect_tx(&bar_ld, 3 * 16384);
        tma_load_2d(sA0, &tmA, &bar_ld, 0, 0);            // A[t, 0:64]
        tma_load_2d(sA1, &tmA, &bar_ld, 64, 0);           // A[t, 64:128]
        tma_load_2d(sB0, &tmA, &bar_ld, 0, 128);          // B := A rows 128..255, cols 0..63
        break;
      case 4:
        mbar_expect_tx(&bar_ld, 16384);
        tma_load_2d(sA0, &tmA, &bar_ld, 64, 128);
        break;
      case 7:
        mbar_expect_tx(&bar_ld, 16384);
        tma_load_3d(sA0, &tmA3, &bar_ld, 64, -8, 1);
        break;
      case 8:
        mbar_expect_tx(&bar_ld, 16384);
        tma_load_2d(sA0, &tmA, &bar_ld, 64, 128);
        break;
      case 5:
        mbar_expect_tx(&bar_ld, 16384 + 8192);
        tma_load_2d(sA0, &tmA, &bar_ld, 0, 0);
        tma_load_2d(sB0, &tmB64, &bar_ld, 0, 0);
        break;
      case 6:
        mbar_expect_tx(&bar_ld, 16384 + 8192);
        tma_load_2d(sA0, &tmA, &bar_ld, 0, 0);            // A[128 x 64] K-major
        tma_load_2d(sB0, &tmB64, &bar_ld, 0, 0);          // B stored [K=64 rows][N=64 cols]
        break;
    }
    mbar_wait(&bar_ld, 0);
    tc_fence_after();
    if (mode == 0 || mode == 2) {
      const uint32_t id = idesc_bf16(128, 128, 0, 0);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(a0, k), desc_kmajor(b0, k), id, k > 0);
    } else if (mode == 1) {
      const uint32_t id = idesc_bf16(128, 256, 0, 0);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(a0, k), desc_kmajor(b0, k), id, k > 0);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(a1, k), desc_kmajor(b1, k), id, true);
    } else if (mode == 3) {
      const uint32_t id = idesc_bf16(128, 64, 1, 1);
      for (int k = 0; k < 8; ++k) umma_bf16(tmem, desc_mnmajor(a0, k, 16384), desc_mnmajor(b0, k, 16384), id, k > 0);
    } else if (mode == 5) {
      const uint32_t id = idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(a0, k), desc_kmajor(b0, k), id, k > 0);
    } else if (mode == 6) {
      const uint32_t id = idesc_bf16(128, 64, 0, 1);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(a0, k), desc_mnmajor(b0, k, 8192), id, k > 0);
    }
    umma_commit(&bar_mma);
    if (mode == 4 || mode == 7) {
      tma_store_2d(&tmOut, sA0, 0, 0);
      tma_store_commit();
      tma_store_wait_all();
    }
    if (mode == 8) {
      tma_store_3d(&tmOut3, sA0, 0, 8, 1);       // rows 8..135 of batch 1 (64 rows): tile rows 0..55 land
      tma_store_commit();
      tma_store_wait_all();
    }
  }
  __syncwarp();
  if (mode == 1) N = 256;
  if (mode == 3 || mode == 5 || mode == 6) N = 64;
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  if (mode != 4 && mode < 7) {
    for (int c = 0; c < N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_addr(tmem, warp * 32, c * 32), v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(int64_t)tid * 256 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// exact reference + comparison; err[0] = max |diff|
__global__ void ref_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, int mode,
                           const float* __restrict__ out, const __nv_bfloat16* __restrict__ out_bf, float* __restrict__ err) {
  const int m = blockIdx.x, n = threadIdx.x;
  auto a = [&](int r, int c) { return __bfloat162float(A[(int64_t)r * 128 + c]); };
  auto bq = [&](int r, int c) { return __bfloat162float(B[(int64_t)r * 128 + c]); };
  float ref = 0.f, got = 0.f;
  bool active = true;
  if (mode == 0 || mode == 2) {
    if (n >= 128) active = false;
    else for (int k = 0; k < 64; ++k) ref += a(m, k) * bq(n, k);
  } else if (mode == 1) {
    for (int k = 0; k < 128; ++k) ref += a(m, k) * bq(n, k);
  } else if (mode == 3) {
    if (n >= 64) active = false;
    else for (int t = 0; t < 128; ++t) ref += a(t, m) * a(128 + t, n);
  } else if (mode == 4) {
    if (n >= 64) active = false;
    else { ref = a(128 + m, 64 + n); got = __bfloat162float(out_bf[(int64_t)m * 64 + n]); }
  } else if (mode == 7) {
    if (n >= 64) active = false;
    else { ref = m < 8 ? 0.f : a(128 + m - 8, 64 + n); got = __bfloat162float(out_bf[(int64_t)m * 64 + n]); }
  } else if (mode == 8) {
    if (n >= 64) active = false;
    else {
      // out_bf viewed as (2 batches, 64 rows, 64 cols); batch 1 row r >= 8 <- tile row r - 8; the rest untouched (0xFFFF)
      const unsigned short bits = reinterpret_cast<const unsigned short*>(out_bf)[(int64_t)m * 64 + n];
      if (m < 64 + 8) { ref = 0.f; got = bits == 0xFFFFu ? 0.f : 1.f; }
      else { ref = a(128 + (m - 64) - 8, 64 + n); got = __bfloat162float(out_bf[(int64_t)m * 64 + n]); }
    }
  } else if (mode == 5) {
    if (n >= 64) active = false;
    else for (int k = 0; k < 64; ++k) ref += a(m, k) * bq(n, k);
  } else if (mode == 6) {
    if (n >= 64) active = false;
    else for (int k = 0; k < 64; ++k) ref += a(m, k) * bq(k, n);
  }
  if (!active) return;
  if (mode != 4 && mode < 7) got = out[(int64_t)m * 256 + n];
  float d = fabsf(got - ref);
  if (!(d == d)) d = 1e30f;
  atomicMax(reinterpret_cast<int*>(err), __float_as_int(d));
}

}  // namespace

int fast_selftest(float* h_maxerr, int n_cases, cudaStream_t s) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(h_maxerr && n_cases >= 1, WN_ERR_INVALID, "fast_selftest: bad arguments");
  __nv_bfloat16 *A = nullptr, *B = nullptr, *obf = nullptr;
  float *out = nullptr, *err = nullptr;
  WN_CHECK_CUDA(cudaMalloc(&A, 256 * 128 * 2));
  WN_CHECK_CUDA(cudaMalloc(&B, 256 * 128 * 2));
  WN_CHECK_CUDA(cudaMalloc(&obf, 128 * 64 * 2));
  WN_CHECK_CUDA(cudaMalloc(&out, 128 * 256 * 4));
  WN_CHECK_CUDA(cudaMalloc(&err, NCASE * 4));
  WN_CHECK_CUDA(cudaMemsetAsync(err, 0, NCASE * 4, s));
  fill_kernel<<<64, 256, 0, s>>>(A, 256 * 128, 17u);
  fill_kernel<<<64, 256, 0, s>>>(B, 256 * 128, 91u);
  CUtensorMap tmA, tmB256, tmB64, tmOut, tmA3, tmOut3;
  WN_PROPAGATE(tmap_3d(&tmA3, A, 128, 128, 2, 128, 128 * 128, 128));
  WN_PROPAGATE(tmap_3d(&tmOut3, obf, 64, 64, 2, 64, 64 * 64, 128));
  WN_PROPAGATE(tmap_2d(&tmA, A, 128, 256, 128, 128));
  WN_PROPAGATE(tmap_2d(&tmB256, B, 128, 256, 128, 256));
  WN_PROPAGATE(tmap_2d(&tmB64, B, 128, 256, 128, 64));
  WN_PROPAGATE(tmap_2d(&tmOut, obf, 64, 128, 64, 128));
  const int smem = 98304 + 1024;
  WN_CHECK_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int c = 0; c < NCASE && c < n_cases; ++c) {
    WN_CHECK_CUDA(cudaMemsetAsync(out, 0xFF, 128 * 256 * 4, s));
    WN_CHECK_CUDA(cudaMemsetAsync(obf, 0xFF, 128 * 64 * 2, s));
    selftest_kernel<<<1, 128, smem, s>>>(tmA, tmB256, tmB64, tmOut, tmA3, tmOut3, A, c, out);
    WN_CHECK_LAUNCH();
    ref_kernel<<<128, 256, 0, s>>>(A, B, c, out, obf, err + c);
    WN_CHECK_LAUNCH();
  }
  WN_CHECK_CUDA(cudaStreamSynchronize(s));
  float herr[NCASE];
  WN_CHECK_CUDA(cudaMemcpy(herr, err, NCASE * 4, cudaMemcpyDeviceToHost));
  for (int c = 0; c < n_cases; ++c) h_maxerr[c] = c < NCASE ? herr[c] : -1.f;
  cudaFree(A); cudaFree(B); cudaFree(obf); cudaFree(out); cudaFree(err);
  return WN_OK;
}

}  // namespace wn

extern "C" int wn_selftest_umma(float* h_maxerr, int32_t n_cases, void* stream) {
  return wn::fast_selftest(h_maxerr, n_cases, (cudaStream_t)stream);
}
