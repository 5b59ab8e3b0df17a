// Token hand-off latency between two CTAs of a cluster (timing experiment, not part of the product).
//   mode 0: st.shared + fence.proxy.async + bar(128) + one thread cp.async.bulk shared::cta -> shared::cluster, complete_tx
//   mode 1: st.shared::cluster.v4 per thread, __syncwarp, lane 0 mbarrier.arrive.release.cluster (count = warps)
//   mode 2: st.async.shared::cluster.mbarrier::complete_tx::bytes.v4 per thread (data and signal travel together)
//   mode 3: like 0 but every warp sends its own slice (no CTA barrier)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void wait_par(uint32_t bar, uint32_t par) {
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    if (ok) return;
  }
}
constexpr int T = 128;   // sender threads (4 warps), 16 B each = 2 KB; plus 8 B each = 1 KB
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) pp(long long* out, int iters) {
  __shared__ __align__(128) float4 inbuf[T];
  __shared__ __align__(128) uint2 inbuf2[T];
  __shared__ __align__(128) float4 stage[T];
  __shared__ __align__(128) uint2 stage2[T];
  __shared__ __align__(8) uint64_t bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t b = smem_u32(&bar);
  const uint32_t bytes = MODE == 4 || MODE == 5 ? T * 8 : MODE == 6 ? T * 16 : T * 24;
  if (tid == 0) {
    const int cnt = MODE == 1 ? T / 32 : 1;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(cnt));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (MODE != 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t peer = rank ^ 1;
  const uint32_t r_in = map_to(smem_u32(inbuf), peer), r_in2 = map_to(smem_u32(inbuf2), peer), r_bar = map_to(b, peer);
  float4 v = make_float4(tid, 1, 2, 3);
  uint2 h = make_uint2(tid, 5);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if ((it & 1) == (int)rank) {
      // ---- send
      if (MODE == 0) {
        stage[tid] = v; stage2[tid] = h;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(r_in), "r"(smem_u32(stage)), "r"(T * 16), "r"(r_bar) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(r_in2), "r"(smem_u32(stage2)), "r"(T * 8), "r"(r_bar) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (MODE == 3) {
        stage[tid] = v; stage2[tid] = h;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(r_in + warp * 512), "r"(smem_u32(stage) + warp * 512), "r"(512), "r"(r_bar) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(r_in2 + warp * 256), "r"(smem_u32(stage2) + warp * 256), "r"(256), "r"(r_bar) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (MODE == 4) {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1,%2}, [%3];" ::"r"(r_in2 + tid * 8), "r"(h.x), "r"(h.y), "r"(r_bar) : "memory");
      } else if (MODE == 6) {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(r_in + tid * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(r_bar) : "memory");
      } else if (MODE == 5) {
        stage2[tid] = h;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(r_in2), "r"(smem_u32(stage2)), "r"(T * 8), "r"(r_bar) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (MODE == 1) {
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(r_in + tid * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        asm volatile("st.shared::cluster.v2.u32 [%0], {%1,%2};" ::"r"(r_in2 + tid * 8), "r"(h.x), "r"(h.y) : "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r_bar) : "memory");
      } else {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(r_in + tid * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(r_bar) : "memory");
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1,%2}, [%3];" ::"r"(r_in2 + tid * 8), "r"(h.x), "r"(h.y), "r"(r_bar) : "memory");
      }
    } else {
      // ---- receive: everybody waits, reads its slot
      wait_par(b, (it >> 1) & 1);
      if (MODE != 1 && tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
      float4 x = inbuf[tid]; uint2 y = inbuf2[tid];
      v.x += x.x; h.x += y.x;
      if (MODE == 0 || MODE == 3 || MODE == 5) { if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
    }
  }
  long long t1 = clock64();
  if (tid == 0 && rank == 0) out[0] = t1 - t0;
  if (tid == 0) out[1 + rank] = (long long)v.x + h.x;
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[3];
  const int it = 2000;
#define RUN(M) pp<M><<<2, T>>>(d, it); pp<M><<<2, T>>>(d, it); cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost); \
  printf("mode %d: %.1f cycles per hand-off   (%lld %lld)  %s\n", M, (double)h[0] / it, h[1], h[2], cudaGetErrorString(cudaDeviceSynchronize()));
  RUN(0) RUN(3) RUN(1) RUN(2) RUN(4) RUN(6) RUN(5)
}
