// bench_kernels.cu - measurement helpers for bench.py (not part of the hot path): an L2 -> SM read-bandwidth probe that gives
// the generation kernel's L2 traffic a MEASURED denominator (the HBM copy peak in MEASURED_PEAKS.json says nothing about L2).
#include "common.cuh"

namespace wn {
namespace {

// every CTA streams the whole buffer `iters` times with 128-bit loads (the buffer is sized to stay L2-resident), as the
// generation kernel's CTAs all stream the same weight image
__global__ void __launch_bounds__(256) l2_read_kernel(const uint4* __restrict__ buf, int64_t n_vec, int iters, uint32_t* __restrict__ sink) {
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    // rotate the start per CTA and iteration so that the CTAs do not walk the same lines in lock step
    const int64_t rot = ((int64_t)blockIdx.x * 977 + (int64_t)it * 131) * 256 % n_vec;
    for (int64_t i = threadIdx.x; i < n_vec; i += 256) {
      int64_t j = i + rot;
      if (j >= n_vec) j -= n_vec;
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + j));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;      // keeps the loads alive
}

}  // namespace
}  // namespace wn

using namespace wn;

extern "C" int wn_bench_l2_read(const void* d_buf, int64_t bytes, int32_t n_ctas, int32_t iters, void* d_sink, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(d_buf && d_sink && bytes >= 4096 && bytes % 16 == 0 && n_ctas > 0 && iters > 0, WN_ERR_INVALID, "wn_bench_l2_read: bad argument");
  l2_read_kernel<<<(unsigned)n_ctas, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(d_buf), bytes / 16, iters,
                                                                    reinterpret_cast<uint32_t*>(d_sink));
  WN_CHECK_LAUNCH();
  return WN_OK;
}
