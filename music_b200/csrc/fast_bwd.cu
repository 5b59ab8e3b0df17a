// fast_bwd.cu - bf16 tensor-core backward of the WaveNet stack (mirror of fast_fwd.cu).
//
//  gemm_nt_kernel<NT> : data-gradient GEMMs  OUT = epi(A W)   (head backward, dskip -> dz, dFG -> dx)
//                       warp-specialised (TMA producer / MMA issuer / 4 epilogue warps), smem ring,
//                       double-buffered TMEM accumulators, epilogue fuses ReLU mask / residual add.
//  gemm_tn_kernel<NB> : weight-gradient GEMMs dW = A^T B with the reduction over time; both operands are
//                       the same [time][64 ch] swizzled tiles, consumed MN-major; persistent CTAs keep
//                       their partial sum in TMEM and flush it once with fp32 atomics.
//  block_bwd_kernel   : per residual block: recompute [f|g] (UMMA), dz = dx_{i+1} Wd (UMMA) + dzs,
//                       gate backward in the epilogue -> dFG and z tiles (TMA stores).
//  block_bwd2 / block_bwd3_kernel : persistent versions that also accumulate the block's weight gradients in TMEM
//                       (block_bwd3, three-stage input ring + tiled skip gradient, is the one that runs by default).
#include <stdlib.h>

#include "check_kernels.cuh"
#include "fast_bwd_kernels.cuh"
#include "fast_kernels.cuh"
#include "fast_layout.cuh"
#include "tc05.cuh"

namespace wn {
using namespace tc;

namespace {

constexpr uint32_t TILE = 16384;   // [128 rows][64 bf16]
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ============================================================================================ gemm_nt
template <int NT>
struct NtCfg {
  // NT = 64 (the dx GEMM): the weight chunks of all K segments (up to 4 x 8 KB) are RESIDENT, only A tiles stream through a
  // 3-stage ring; re-fetching them per item was 29 % of the kernel's L2 -> SM traffic (8.3 TB/s, at the L2 rate limit).
  static constexpr bool RES_B = NT == 64;
  static constexpr int STAGES = 3;
  static constexpr uint32_t A_BYTES = TILE, B_BYTES = NT * 128, STAGE = RES_B ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr uint32_t BRES_OFF = STAGES * STAGE, BRES_BYTES = RES_B ? 4 * B_BYTES : 0;
  static constexpr uint32_t OUT_OFF = BRES_OFF + BRES_BYTES, TOTAL = OUT_OFF + (NT / 64) * TILE;
  static constexpr int TMEM_COLS = NT == 256 ? 512 : 128;
  static constexpr int CTAS_PER_SM = NT == 256 ? 1 : 2;      // NT=64: 97 KB smem, 128 TMEM columns per CTA
};

template <int NT, int EPI>      // EPI is a compile-time constant: a run-time p.epi made the epilogue branch per element
__global__ void __launch_bounds__(192, NtCfg<NT>::CTAS_PER_SM)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmOut, GemmNtParams p) {
  using Cfg = NtCfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[Cfg::STAGES], empty[Cfg::STAGES], acc_full[2], acc_empty[2], b_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 1);
    }
    mbar_init(&b_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_launch_dependents();      // programmatic dependent launch: see tc05.cuh
  pdl_wait();
  const uint32_t sbase = smem_u32(sm);
  const int n_items = p.n_batches * p.tiles_per_batch * p.n_ntiles;

  if (warp == 4) {
    if (lane == 0) {
      if (Cfg::RES_B && (int)blockIdx.x < n_items) {      // n_ntiles == 1: every item uses the same weight chunks
        mbar_expect_tx(&b_full, (p.nk[0] + p.nk[1]) * Cfg::B_BYTES);
        int c = 0;
        for (int seg = 0; seg < 2; ++seg)
          for (int k = 0; k < p.nk[seg]; ++k, ++c)
            tma_load_2d(sm + Cfg::BRES_OFF + c * Cfg::B_BYTES, seg == 0 ? &tmB0 : &tmB1, &b_full, p.b_col0[seg] + 64 * k, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nt = item % p.n_ntiles, rt = item / p.n_ntiles;
        const int b = rt / p.tiles_per_batch, row0 = (p.tile0 + rt % p.tiles_per_batch) * 128;
        for (int seg = 0; seg < 2; ++seg) {
          for (int k = 0; k < p.nk[seg]; ++k) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = sm + stage * Cfg::STAGE;
            mbar_expect_tx(&full[stage], Cfg::STAGE);
            tma_load_3d(sa, seg == 0 ? &tmA0 : &tmA1, &full[stage], p.a_col0[seg] + 64 * k, row0 + p.a_row_off[seg], b, p.pol_a[seg]);
            if (!Cfg::RES_B)
              tma_load_2d(sa + Cfg::A_BYTES, seg == 0 ? &tmB0 : &tmB1, &full[stage], p.b_col0[seg] + 64 * k, nt * NT);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      constexpr uint32_t idn = idesc_bf16(128, NT, 0, 0);
      if (Cfg::RES_B && (int)blockIdx.x < n_items) mbar_wait(&b_full, 0);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t dst = tmem + as * NT;
        const int nk = p.nk[0] + p.nk[1];
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = sbase + stage * Cfg::STAGE;
          const uint32_t sb = Cfg::RES_B ? sbase + Cfg::BRES_OFF + kc * Cfg::B_BYTES : sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(dst, desc_kmajor(sa, k), desc_kmajor(sb, k), idn, (kc | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int nt = item % p.n_ntiles, rt = item / p.n_ntiles;
      const int b = rt / p.tiles_per_batch, row0 = (p.tile0 + rt % p.tiles_per_batch) * 128;
      const int row = row0 + tid;
      const bool valid = row >= p.row_lo && row < p.row_hi;
      const uint32_t src = tmem_addr(tmem, warp * 32, as * NT);
      const __nv_bfloat16* auxp =
          p.aux ? p.aux + (int64_t)b * p.aux_bstride + (int64_t)row * p.aux_rstride + p.aux_col0 + nt * NT : nullptr;
      if constexpr (NT == 64) {
        // the whole 128-byte aux row is fetched BEFORE waiting for the accumulator, and the TMA store of the
        // previous item is only checked right before the staging tile is rewritten: nobody waits for a store
        uint4 ax4[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) ax4[q] = make_uint4(0, 0, 0, 0);
        if (EPI != EPI_PLAIN && valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q) ax4[q] = *reinterpret_cast<const uint4*>(auxp + q * 8);
        }
        mbar_wait(&acc_full[as], aph);
        tc_fence_after();
        uint32_t packed[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(src + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a0 = __uint_as_float(v[2 * j]), a1 = __uint_as_float(v[2 * j + 1]);
            const uint32_t axw = reinterpret_cast<const uint32_t*>(&ax4[c * 4 + (j >> 2)])[j & 3];
            const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&axw);
            if (EPI == EPI_MASK) {
              a0 = bf16lo(x2) > 0.f ? a0 : 0.f;
              a1 = bf16hi(x2) > 0.f ? a1 : 0.f;
            } else if (EPI == EPI_ADD) {
              a0 += bf16lo(x2);
              a1 += bf16hi(x2);
            }
            packed[c * 16 + j] = valid ? pack_bf16(a0, a1) : 0u;
          }
        }
        tc_fence_before();
        if (tid == 0) tma_store_wait_read();
        epi_bar_sync();
        if (tid == 0) mbar_arrive(&acc_empty[as]);
        uint8_t* ot = sm + Cfg::OUT_OFF;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sts128(smem_u32(ot) + sw128_chunk(tid, q), packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        fence_proxy_async_smem();
        epi_bar_sync();
        if (tid == 0) {
          tma_store_3d(&tmOut, ot, p.out_col0 + nt * NT, row0, b, p.pol_out);
          tma_store_commit();
        }
      } else if constexpr (EPI == EPI_LOGITS || EPI == EPI_LOGITS_BIAS) {
        // head output: fp32 logits written straight into the (B, Q, W) tensor (lanes = consecutive time steps: coalesced per column)
        mbar_wait(&acc_full[as], aph);
        tc_fence_after();
        const int tw = row - p.lg_pad;
        const bool ok = valid && tw >= 0 && tw < p.lg_W;
        float* out = p.logits + ((int64_t)b * p.lg_Q + nt * NT) * p.lg_W + tw;
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(src + c * 32, v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float a = __uint_as_float(v[j]);
              if (EPI == EPI_LOGITS_BIAS) a += p.bias[nt * NT + c * 32 + j];
              if (nt * NT + c * 32 + j < p.lg_Q) out[(int64_t)(c * 32 + j) * p.lg_W] = a;
            }
          }
        }
        tc_fence_before();
        epi_bar_sync();
        if (tid == 0) mbar_arrive(&acc_empty[as]);
      } else {
        mbar_wait(&acc_full[as], aph);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(src + c * 32, v);
          tmem_ld_wait();
          uint32_t packed[16];
          uint32_t ax[16];
          const float* condp = nullptr;
          if (EPI == EPI_RELU_COND) {
            const int tw = min(max(row - p.lg_pad, 0), p.lg_W - 1);      // pad rows: any frame (their gradients are zero)
            condp = p.cond + ((int64_t)b * p.cond_frames + cond_frame(tw, p.lg_W, p.cond_frames)) * p.n_total + nt * NT + c * 32;
          }
          float4 cv[8];      // (16-byte loads: n_total and the column offsets are multiples of 32 floats)
          if (EPI == EPI_RELU_COND) {
#pragma unroll
            for (int q = 0; q < 8; ++q) cv[q] = __ldg(reinterpret_cast<const float4*>(condp) + q);
          }
          if ((EPI == EPI_MASK || EPI == EPI_ADD) && valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 a4 = *reinterpret_cast<const uint4*>(auxp + c * 32 + q * 8);
              ax[4 * q] = a4.x; ax[4 * q + 1] = a4.y; ax[4 * q + 2] = a4.z; ax[4 * q + 3] = a4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) ax[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a0 = __uint_as_float(v[2 * j]), a1 = __uint_as_float(v[2 * j + 1]);
            const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&ax[j]);
            if (EPI == EPI_MASK) {
              a0 = bf16lo(x2) > 0.f ? a0 : 0.f;
              a1 = bf16hi(x2) > 0.f ? a1 : 0.f;
            } else if (EPI == EPI_ADD) {
              a0 += bf16lo(x2);
              a1 += bf16hi(x2);
            } else if (EPI == EPI_RELU || EPI == EPI_RELU_BIAS || EPI == EPI_RELU_COND) {
              if (EPI == EPI_RELU_BIAS) {
                a0 += p.bias[nt * NT + c * 32 + 2 * j];
                a1 += p.bias[nt * NT + c * 32 + 2 * j + 1];
              }
              if (EPI == EPI_RELU_COND) {
                const float4 c4 = cv[j >> 1];
                a0 += (j & 1) ? c4.z : c4.x;
                a1 += (j & 1) ? c4.w : c4.y;
              }
              a0 = fmaxf(a0, 0.f);
              a1 = fmaxf(a1, 0.f);
            }
            packed[j] = valid ? pack_bf16(a0, a1) : 0u;
          }
          uint8_t* ot = sm + Cfg::OUT_OFF + (c >> 1) * TILE;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 val = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            const int qq = (c & 1) * 4 + q;      // 16-byte chunk of this thread's row inside the 64-column staging tile
            const uint32_t o = p.out_tiled ? (uint32_t)(((((tid >> 5) * 4 + (qq >> 1)) * 32 + (tid & 31)) << 5) + ((qq & 1) << 4)) : sw128_chunk(tid, qq);
            sts128(smem_u32(ot) + o, val.x, val.y, val.z, val.w);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        epi_bar_sync();
        if (tid == 0) {
          mbar_arrive(&acc_empty[as]);
          for (int j = 0; j < NT / 64; ++j) {
            if (p.n_total > 0 && nt * NT + 64 * j >= p.n_total) break;      // partial last column tile
            if (p.out_tiled) {      // tiled skip-gradient layout (see gemm_nt_resb_kernel): the in-range 32-row blocks are one contiguous run
              const int rb0 = row0 >> 5, nb = min(4, p.out_nblk - rb0);
              const int64_t lb = (int64_t)((p.out_col0 + nt * NT) / 64 + j) * p.n_batches + b;
              if (nb > 0) bulk_store(p.out_tiled + ((lb * p.out_nblk + rb0) << 11), sm + Cfg::OUT_OFF + j * TILE, (uint32_t)nb << 12);
            } else {
              tma_store_3d(&tmOut, sm + Cfg::OUT_OFF + j * TILE, p.out_col0 + nt * NT + 64 * j, row0, b);
            }
          }
          tma_store_commit();
          tma_store_wait_read();
        }
        epi_bar_sync();
      }
    }
    if (tid == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ============================================================================================ gemm_nt, resident B
// OUT[rows, 256 nt .. 256 nt + 256) = epi(A[rows, 0..256) W^T) with K = 256: the 128 KB weight block of the CTA's column
// tile stays in shared memory for the whole launch and only the A row tiles stream through (64 KB per 16.8 MFLOP item).
// The streaming version above re-reads the weight block per item (192 KB per item, 28 MB per wave of 148 CTAs) and is
// bound by L2 -> SM bandwidth at a quarter of the tensor rate.  CTA c owns column tile c mod n_ntiles and every
// (number of CTAs with that tile)-th row tile.  Output goes through a ring of three 64-column staging tiles whose TMA
// stores are only checked two tiles later.
struct NtResCfg {
  static constexpr int A_STAGES = 3, OUT_SLOTS = 3;
  static constexpr uint32_t B_OFF = 0, B_BYTES = 4 * 256 * 128;            // 4 k-chunks of [256 n][64 k]
  static constexpr uint32_t A_OFF = B_BYTES, OUT_OFF = A_OFF + A_STAGES * TILE, TOTAL = OUT_OFF + OUT_SLOTS * TILE;
};

template <int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_nt_resb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, GemmNtParams p) {
  using Cfg = NtResCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t b_full, full[Cfg::A_STAGES], empty[Cfg::A_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&b_full, 1);
    for (int i = 0; i < Cfg::A_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_launch_dependents();      // programmatic dependent launch: see tc05.cuh
  pdl_wait();
  const uint32_t sbase = smem_u32(sm);
  const int T = p.n_ntiles;
  const int nt = (int)blockIdx.x % T, rank = (int)blockIdx.x / T;
  const int peers = ((int)gridDim.x - 1 - nt) / T + 1;                    // CTAs that own column tile nt
  const int n_rows = p.n_batches * p.tiles_per_batch;
  const int n_mine = rank < n_rows ? (n_rows - 1 - rank) / peers + 1 : 0;
  const int n_valid = min(256, p.n_total - nt * 256);                      // multiple of 64
  const int n_otiles = n_valid / 64;

  if (warp == 4) {
    if (lane == 0 && n_mine > 0) {
      mbar_expect_tx(&b_full, Cfg::B_BYTES);
      for (int k = 0; k < 4; ++k) tma_load_2d(sm + Cfg::B_OFF + k * 256 * 128, &tmB, &b_full, p.b_col0[0] + 64 * k, nt * 256);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_mine; ++it) {
        const int rt = rank + it * peers;
        const int b = rt / p.tiles_per_batch, row0 = (p.tile0 + rt % p.tiles_per_batch) * 128;
        for (int k = 0; k < 4; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], TILE);
          tma_load_3d(sm + Cfg::A_OFF + stage * TILE, &tmA, &full[stage], p.a_col0[0] + 64 * k, row0 + p.a_row_off[0], b);
          if (++stage == Cfg::A_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && n_mine > 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idn = idesc_bf16(128, n_valid, 0, 0);
      mbar_wait(&b_full, 0);
      for (int it = 0; it < n_mine; ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t dst = tmem + as * 256;
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = sbase + Cfg::A_OFF + stage * TILE, sb = sbase + Cfg::B_OFF + kc * 256 * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(dst, desc_kmajor(sa, k), desc_kmajor(sb, k), idn, (kc | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == Cfg::A_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    int slot = 0;
    for (int it = 0; it < n_mine; ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int rt = rank + it * peers;
      const int b = rt / p.tiles_per_batch, row0 = (p.tile0 + rt % p.tiles_per_batch) * 128;
      const int row = row0 + tid;
      const bool valid = row >= p.row_lo && row < p.row_hi;
      const uint32_t src = tmem_addr(tmem, warp * 32, as * 256);
      const __nv_bfloat16* auxp =
          p.aux ? p.aux + (int64_t)b * p.aux_bstride + (int64_t)row * p.aux_rstride + p.aux_col0 + nt * 256 : nullptr;
      for (int j = 0; j < n_otiles; ++j) {
        uint4 ax4[8];
        if (EPI != EPI_PLAIN && valid) {         // the 128-byte aux row of this tile, requested before the accumulator is read
#pragma unroll
          for (int q = 0; q < 8; ++q) ax4[q] = *reinterpret_cast<const uint4*>(auxp + j * 64 + q * 8);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) ax4[q] = make_uint4(0, 0, 0, 0);
        }
        if (j == 0) {
          mbar_wait(&acc_full[as], aph);
          tc_fence_after();
        }
        uint32_t packed[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(src + j * 64 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float a0 = __uint_as_float(v[2 * e]), a1 = __uint_as_float(v[2 * e + 1]);
            const uint32_t axw = reinterpret_cast<const uint32_t*>(&ax4[c * 4 + (e >> 2)])[e & 3];
            const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&axw);
            if (EPI == EPI_MASK) {
              a0 = bf16lo(x2) > 0.f ? a0 : 0.f;
              a1 = bf16hi(x2) > 0.f ? a1 : 0.f;
            } else if (EPI == EPI_ADD) {
              a0 += bf16lo(x2);
              a1 += bf16hi(x2);
            }
            packed[c * 16 + e] = valid ? pack_bf16(a0, a1) : 0u;
          }
        }
        // staging slot `slot` was stored three tiles ago; thread 0 checked that store before the previous barrier
        uint8_t* ot = sm + Cfg::OUT_OFF + slot * TILE;
#pragma unroll
        for (int q = 0; q < 8; ++q) {     // 16-byte chunk q of this thread's row: swizzled tile, or the tiled global layout's image
          const uint32_t o = p.out_tiled ? (uint32_t)(((((tid >> 5) * 4 + (q >> 1)) * 32 + (tid & 31)) << 5) + ((q & 1) << 4)) : sw128_chunk(tid, q);
          sts128(smem_u32(ot) + o, packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
        fence_proxy_async_smem();
        if (j == n_otiles - 1) tc_fence_before();
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // all but the newest store have left smem
        epi_bar_sync();
        if (tid == 0) {
          if (j == n_otiles - 1) mbar_arrive(&acc_empty[as]);
          if (p.out_tiled) {              // the in-range 32-row blocks of this tile are one contiguous run
            const int rb0 = row0 >> 5, nb = min(4, p.out_nblk - rb0);
            const int64_t lb = (int64_t)((p.out_col0 + nt * 256) / 64 + j) * p.n_batches + b;
            if (nb > 0) bulk_store(p.out_tiled + ((lb * p.out_nblk + rb0) << 11), ot, (uint32_t)nb << 12);
          } else {
            tma_store_3d(&tmOut, ot, p.out_col0 + nt * 256 + 64 * j, row0, b);
          }
          tma_store_commit();
        }
        if (++slot == Cfg::OUT_SLOTS) slot = 0;
      }
    }
    if (tid == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ============================================================================================ gemm_tn
template <int NB>
struct TnCfg {
  static constexpr int STAGES = NB == 4 ? 2 : (NB == 2 ? 3 : 4);
  static constexpr uint32_t A_BYTES = 2 * TILE, B_BYTES = NB * TILE, STAGE = A_BYTES + B_BYTES;
  static constexpr uint32_t TOTAL = STAGES * STAGE;
  static constexpr int TMEM_COLS = NB == 4 ? 256 : (NB == 2 ? 128 : 64);
};

template <int NB>
__global__ void __launch_bounds__(192, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmB1, GemmTnParams p) {
  using Cfg = TnCfg<NB>;
  if (p.y_layers > 0) {      // derive this y-slice's operands (uniform over the CTA)
    const int n_mt = p.n_mtiles > 0 ? p.n_mtiles : 2;
    const int mt = blockIdx.y % n_mt, g = blockIdx.y / n_mt;
    p.a_col0 = 128 * mt;
    p.out0 += (int64_t)128 * mt * p.s_m;
    p.out1 += (int64_t)128 * mt * p.s_m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int li = 4 * g + j;
      p.b_col[j] = 64 * li;                                   // past the last layer: outside the tensor -> zero fill
      p.blk_off[j] = p.y_off0 + (int64_t)min(li, p.y_layers - 1) * p.y_stride;
    }
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[Cfg::STAGES], empty[Cfg::STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_launch_dependents();      // programmatic dependent launch: see tc05.cuh
  pdl_wait();
  const uint32_t sbase = smem_u32(sm);
  const int n_items = p.n_batches * p.tiles_per_batch;
  const bool have_work = (int)blockIdx.x < n_items;
  const bool two_a = p.m_valid > 64;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / p.tiles_per_batch, row0 = (p.tile0 + item % p.tiles_per_batch) * 128;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = sm + stage * Cfg::STAGE;
        uint8_t* sb = sa + Cfg::A_BYTES;
        mbar_expect_tx(&full[stage], (two_a ? 2 : 1) * TILE + NB * TILE);
        tma_load_3d(sa, &tmA, &full[stage], p.a_col0, row0, b);
        if (two_a) tma_load_3d(sa + TILE, &tmA, &full[stage], p.a_col0 + 64, row0, b);
#pragma unroll
        for (int j = 0; j < NB; ++j)
          tma_load_3d(sb + j * TILE, p.b_map[j] == 0 ? &tmB0 : &tmB1, &full[stage], p.b_col[j], row0 + p.b_row_off[j], b);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && have_work) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      constexpr uint32_t idn = idesc_bf16(128, 64 * NB, 1, 1);
      const uint32_t lbo_a = two_a ? TILE : 0u;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * Cfg::STAGE, sb = sa + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem, desc_mnmajor(sa, k, lbo_a), desc_mnmajor(sb, k, TILE), idn, (it | (uint32_t)k) != 0);
        umma_commit(&empty[stage]);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&acc_full);
    }
  } else if (have_work) {
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int m = tid;
    float* base = (m < 64 ? p.out0 : p.out1) + (int64_t)(m & 63) * p.s_m;
    const uint32_t src = tmem_addr(tmem, warp * 32, 0);
#pragma unroll 1
    for (int c = 0; c < 2 * NB; ++c) {
      uint32_t v[32];
      tmem_ld32(src + c * 32, v);
      tmem_ld_wait();
      if (m < p.m_valid) {
        float* ob = base + p.blk_off[c >> 1] + (int64_t)((c & 1) * 32) * p.s_n;
        const int nv = (p.n_valid > 0 ? p.n_valid : 64) - (c & 1) * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nv) atomicAdd(ob + (int64_t)j * p.s_n, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ===================================================================================== causal_wgrad
// dW_causal (R = 64, Q = 256, 2) = scatter of dx0 rows by the input codes, as a tensor-core GEMM with a ONE-HOT A operand that
// is built in shared memory from the indices (never materialised in HBM):
//   tap 1:  dW[r, q, 1] = sum_t [idx[t] == q] dx0[t, r]          tap 0:  dW[r, q, 0] = sum_t [idx[t] == q] dx0[t + 1, r]
// Per 128-row tile: A = OH [K = 128 rows t][M = 256 q] (MN-major, four 64-column blocks), B = [dx0[t+1] | dx0[t]] (MN-major,
// N = 128), two UMMA groups (q < 128, q >= 128) accumulate in TMEM across all tiles of the CTA.  The one-hot tiles are zeroed
// once; afterwards every thread only clears the element it set two tiles ago and sets the new one.  Unlike the shared-memory
// histogram this replaces (39 M shared atomics, 0.2 ms), the cost does not depend on how the codes are distributed.
struct CwCfg {
  static constexpr int STAGES = 2;
  static constexpr uint32_t A_BYTES = 4 * TILE, B_BYTES = 2 * TILE, STAGE = A_BYTES + B_BYTES, TOTAL = STAGES * STAGE;
};
__global__ void __launch_bounds__(192, 1)
causal_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dx0, const int64_t* __restrict__ idx, float* __restrict__ dW, int L,
                    int n_batches, int tiles_per_batch, int R) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[CwCfg::STAGES], oh_full[CwCfg::STAGES], empty[CwCfg::STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < CwCfg::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&oh_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  if (warp < 4) {      // zero both one-hot areas once (128 threads, 16-byte stores)
    for (int st = 0; st < CwCfg::STAGES; ++st)
      for (uint32_t o = tid * 16; o < CwCfg::A_BYTES; o += 128 * 16) *reinterpret_cast<uint4*>(sm + st * CwCfg::STAGE + o) = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_launch_dependents();      // programmatic dependent launch: see tc05.cuh
  pdl_wait();
  const uint32_t sbase = smem_u32(sm);
  const int n_items = n_batches * tiles_per_batch;
  const bool have_work = (int)blockIdx.x < n_items;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / tiles_per_batch, row0 = (item % tiles_per_batch) * 128;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sb = sm + stage * CwCfg::STAGE + CwCfg::A_BYTES;
        mbar_expect_tx(&full[stage], CwCfg::B_BYTES);
        tma_load_3d(sb, &tm_dx0, &full[stage], 0, row0 + 1, b);          // tap 0 pairs idx[t] with dx0[t + 1]
        tma_load_3d(sb + TILE, &tm_dx0, &full[stage], 0, row0, b);       // tap 1 pairs idx[t] with dx0[t]
        if (++stage == CwCfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && have_work) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      constexpr uint32_t idn = idesc_bf16(128, 128, 1, 1);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        mbar_wait(&full[stage], phase);
        mbar_wait(&oh_full[stage], phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * CwCfg::STAGE, sb = sa + CwCfg::A_BYTES;
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + mh * 128, desc_mnmajor(sa + mh * 2 * TILE, k, TILE), desc_mnmajor(sb, k, TILE), idn, (it | (uint32_t)k) != 0);
        umma_commit(&empty[stage]);
        if (++stage == CwCfg::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&acc_full);
    }
  } else if (have_work) {
    // ---- one-hot builders: thread i owns row i of every tile
    int stage = 0;
    uint32_t phase = 0;
    int prev[CwCfg::STAGES];
#pragma unroll
    for (int st = 0; st < CwCfg::STAGES; ++st) prev[st] = -1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / tiles_per_batch, row0 = (item % tiles_per_batch) * 128;
      const int t = row0 + tid;
      const int q = t < L ? (int)idx[(int64_t)b * L + t] : -1;
      mbar_wait(&empty[stage], phase ^ 1);        // the MMAs that read this stage's previous tile have completed
      uint8_t* sa = sm + stage * CwCfg::STAGE;
      // static indexing of prev[] (two stages)
      int& pv = stage == 0 ? prev[0] : prev[1];
      if (pv >= 0) *reinterpret_cast<uint16_t*>(sa + pv) = 0;
      if (q >= 0) {
        pv = (q >> 6) * (int)TILE + (int)sw128_offset(tid, q & 63);
        *reinterpret_cast<uint16_t*>(sa + pv) = 0x3F80;      // bf16 1.0
      } else {
        pv = -1;
      }
      fence_proxy_async_smem();
      epi_bar_sync();
      if (tid == 0) mbar_arrive(&oh_full[stage]);
      if (++stage == CwCfg::STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- flush: TMEM row m of group mh = code q = 128 mh + m; columns [0,64) tap 0, [64,128) tap 1
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const uint32_t src = tmem_addr(tmem, warp * 32, 0);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t v[32];
      tmem_ld32(src + c * 32, v);
      tmem_ld_wait();
      const int mh = c >> 2, tap = (c >> 1) & 1, r0 = (c & 1) * 32;
      float* ob = dW + (int64_t)(128 * mh + tid) * 2 + tap;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(v[j]);
        if (x != 0.f && r0 + j < R) atomicAdd(ob + (int64_t)(r0 + j) * 512, x);      // dW[r][q][tap], Q = 256
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

__device__ __forceinline__ void epi8_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// ========================================================================================= block_bwd3
// The default block backward (WN_BWD3=0 selects block_bwd2; history and numbers: profiles/r1_summary.md, "r1d").  Same work
// and warp roles as block_bwd2 with the per-stage dependency loop shortened - there a two-stage input ring is refilled only
// after the weight-gradient MMAs of the tile that held it, so load issue -> arrival -> recompute -> epilogue -> weight
// gradients runs once per TWO tiles:
//  * x taps in a THREE-stage ring, dx_{i+1} in its own two-stage ring.  That fits the same 216 KB because the skip-path
//    gradient does not pass through shared memory: the dZcat GEMM writes it in a tiled layout (GemmNtParams::out_tiled)
//    in which the 32 rows x 16 channels of one epilogue warp are 1 KB contiguous, the producer prefetches the 16 KB tile
//    into L2 two to three tiles ahead and every epilogue thread reads its 32 bytes with one 256-bit load.  (From the
//    row-major dZcat the same loads - one lane per 3840-byte-pitched row, 32 lines per request - cost 0.3-0.6 ms per step.)
//  * f|g and dz complete on separate barriers and the epilogue does the tanh / sigmoid part, which does not depend on
//    dz, before it waits for dz - a late dx tile costs only the last 48 FMAs of the tile.
// Measured at cfg 2: 1.56-1.60 ms per step (block_bwd2: 1.65-1.72 ms; 1.45 ms without the skip-gradient loads).
struct Bwd3Smem {
  static constexpr uint32_t W0 = 0, W1 = TILE, WDT = 2 * TILE;                 // resident weights (40 KB)
  static constexpr uint32_t NXS = 3, XR = 2 * TILE + 8192, X_STAGE = 2 * TILE; // 3 x {x tap0, x tap1}
  static constexpr uint32_t DXR = XR + NXS * X_STAGE;                          // 2 x dx_{i+1}
  static constexpr uint32_t DF = DXR + 2 * TILE, DG = DF + TILE, Z = DG + TILE;
  static constexpr uint32_t TOTAL = Z + TILE;                                  // 216 KB
};
// 32 bytes per thread in one request (256-bit LDG), not allocated in L1; `policy` = L2 eviction hint or 0
__device__ __forceinline__ void ldg_stream32(const void* ptr, uint64_t policy, uint32_t (&v)[8]) {
  if (policy)
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(ptr), "l"(policy));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(ptr));
}

// 32 bytes per thread in one request (256-bit STG): a full sector per lane; `policy` = L2 eviction hint or 0
__device__ __forceinline__ void stg32(void* ptr, uint64_t policy, const uint32_t* v) {
  if (policy)
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "l"(policy) : "memory");
  else
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

template <bool BIAS, bool DENSE, bool COND>
__global__ void __launch_bounds__(576, 1)
block_bwd3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w0,
                  const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_dx,
                  const __grid_constant__ CUtensorMap tm_wdT, const __grid_constant__ CUtensorMap tm_dfg, BlockBwd2Params pp) {
  const BlockBwdParams& p = pp.b;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t w_full, x_full[3], x_empty[3], dxi_full[2], dxi_empty[2], fg_full[2], fg_empty[2];
  __shared__ __align__(8) uint64_t dz_full, dz_empty, out_full, out_empty, wg_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dxi_full[i], 1);
      mbar_init(&dxi_empty[i], 1);
      mbar_init(&fg_full[i], 1);
      mbar_init(&fg_empty[i], 1);
    }
    mbar_init(&dz_full, 1);
    mbar_init(&dz_empty, 1);
    mbar_init(&out_full, 1);
    mbar_init(&out_empty, 1);
    mbar_init(&wg_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_launch_dependents();      // programmatic dependent launch: see tc05.cuh
  pdl_wait();
  const uint32_t sbase = smem_u32(sm);
  const int n_items = pp.n_batches * p.tiles_per_batch;
  const int n_mine = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  constexpr uint32_t C_DZ = 256, C_WFG = 320, C_WD = 448;      // TMEM columns as in block_bwd2

  if (warp == 16) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0 && n_mine > 0) {
      mbar_expect_tx(&w_full, 2 * TILE + (DENSE ? 8192 : 0));
      tma_load_2d(sm + Bwd3Smem::W0, &tm_w0, &w_full, 0, 0);
      tma_load_2d(sm + Bwd3Smem::W1, &tm_w1, &w_full, 0, 0);
      if (DENSE) tma_load_2d(sm + Bwd3Smem::WDT, &tm_wdT, &w_full, 0, 0);
      int sx = 0, xph = 0;
      for (int it = 0; it < n_mine; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int b = item / p.tiles_per_batch, tau0 = (p.tile0 + item % p.tiles_per_batch) * 128;
        mbar_wait(&x_empty[sx], xph ^ 1);
        uint8_t* sxp = sm + Bwd3Smem::XR + sx * Bwd3Smem::X_STAGE;
        mbar_expect_tx(&x_full[sx], 2 * TILE);
        tma_load_3d(sxp, &tm_x, &x_full[sx], 0, tau0 - p.d, b, p.pol_first);     // the last read of these rows of x_i
        tma_load_3d(sxp + TILE, &tm_x, &x_full[sx], 0, tau0, b);
        // the epilogue reads this tile's skip-path gradient straight from global memory two to three tiles from now
        if (tau0 >= p.tw_al) {
          const int rb0 = (tau0 - p.tw_al) >> 5, nb = min(4, p.dzs_nblk - rb0);
          if (nb > 0) bulk_prefetch_l2(p.dzs + (((int64_t)(p.dzs_lb0 + b) * p.dzs_nblk + rb0) << 11), (uint32_t)nb << 12);
        }
        if (DENSE) {
          const int sd = it & 1;
          mbar_wait(&dxi_empty[sd], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&dxi_full[sd], TILE);
          tma_load_3d(sm + Bwd3Smem::DXR + sd * TILE, &tm_dx, &dxi_full[sd], 0, tau0, b);
        }
        if (++sx == 3) { sx = 0; xph ^= 1; }
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------ MMA issuer (polling)
    if (lane == 0 && n_mine > 0) {
      constexpr uint32_t id_fg = idesc_bf16(128, 128, 0, 0), id_dz = idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t id_wfg = idesc_bf16(128, 128, 1, 1), id_wd = idesc_bf16(128, 64, 1, 1);
      mbar_wait(&w_full, 0);
      int jf = 0, jd = 0, jw = 0;       // next tile for: f|g recompute, dz, weight gradients
      int sf = 0, fph = 0, sw = 0;      // x-ring stage / phase of tile jf, stage of tile jw
      while (jw < n_mine) {
        // (1) dz of tile jd (on the epilogue's critical path): its dx tile has landed, the previous epilogue has drained the accumulator
        if (DENSE && jd < n_mine && mbar_test_wait(&dz_empty, (jd & 1) ^ 1) && mbar_test_wait(&dxi_full[jd & 1], (jd >> 1) & 1)) {
          tc_fence_after();
          const uint32_t sd = sbase + Bwd3Smem::DXR + (jd & 1) * TILE;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + C_DZ, desc_kmajor(sd, k), desc_kmajor(sbase + Bwd3Smem::WDT, k), id_dz, k > 0);
          umma_commit(&dz_full);
          ++jd;
          continue;
        }
        // (2) weight gradients of tile jw once its epilogue has produced dF | dG | z in shared memory
        if (jw < jf && mbar_test_wait(&out_full, jw & 1)) {
          tc_fence_after();
          const uint32_t sxa = sbase + Bwd3Smem::XR + sw * Bwd3Smem::X_STAGE;
#pragma unroll
          for (int k = 0; k < 8; ++k)     // dW_fg[o, (tap, r)] += sum_t dFG[t, o] * x[t - (1 - tap) d, r]
            umma_bf16(tmem + C_WFG, desc_mnmajor(sbase + Bwd3Smem::DF, k, TILE), desc_mnmajor(sxa, k, TILE), id_wfg, (jw | k) != 0);
          umma_commit(&x_empty[sw]);
          if (DENSE) {
            const uint32_t sd = sbase + Bwd3Smem::DXR + (jw & 1) * TILE;
#pragma unroll
            for (int k = 0; k < 8; ++k)   // dW_dense[r, d] += sum_t dx_{i+1}[t, r] * z[t, d]   (rows 64..127 unused)
              umma_bf16(tmem + C_WD, desc_mnmajor(sd, k, 0), desc_mnmajor(sbase + Bwd3Smem::Z, k, TILE), id_wd, (jw | k) != 0);
            umma_commit(&dxi_empty[jw & 1]);
          }
          umma_commit(&out_empty);
          ++jw;
          if (++sw == 3) sw = 0;
          continue;
        }
        // (3) f|g recompute of tile jf into accumulator buffer jf & 1
        if (jf < n_mine && mbar_test_wait(&x_full[sf], fph) && mbar_test_wait(&fg_empty[jf & 1], ((jf >> 1) & 1) ^ 1)) {
          tc_fence_after();
          const uint32_t sxa = sbase + Bwd3Smem::XR + sf * Bwd3Smem::X_STAGE, acc = tmem + (jf & 1) * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc, desc_kmajor(sxa, k), desc_kmajor(sbase + Bwd3Smem::W0, k), id_fg, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc, desc_kmajor(sxa + TILE, k), desc_kmajor(sbase + Bwd3Smem::W1, k), id_fg, true);
          umma_commit(&fg_full[jf & 1]);
          ++jf;
          if (++sf == 3) { sf = 0; fph ^= 1; }
          continue;
        }
      }
      umma_commit(&wg_done);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 0-15
    const int q4 = warp & 3, cg = warp >> 2;          // TMEM lane quarter, 16-column group
    const int row = q4 * 32 + lane;
    const uint32_t lane_addr = tmem_addr(tmem, q4 * 32, 0);
    // (batch row, tile) of this CTA's next work item, stepped without a division per tile
    int b = (int)blockIdx.x / p.tiles_per_batch, tl = (int)blockIdx.x % p.tiles_per_batch;
    for (int it = 0; it < n_mine; ++it) {
      const uint32_t ph = it & 1, ph2 = (it >> 1) & 1;
      const int tau0 = (p.tile0 + tl) * 128;
      const int tau = tau0 + row;
      const bool valid = tau >= p.s_out && tau < p.L;
      // skip-path gradient of this thread's 16 channels (zero before the last W time steps and past the end of the row):
      // one 32-byte request per thread, in flight during the gate math
      uint32_t zs[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if (tau0 >= p.tw_al && tau < p.L)      // tiled layout (GemmNtParams::out_tiled): this warp's 32 rows x 16 channels are 1 KB contiguous
        ldg_stream32(p.dzs + (((((int64_t)(p.dzs_lb0 + b) * p.dzs_nblk + ((tau0 - p.tw_al) >> 5) + q4) * 4 + cg) * 32 + lane) << 4), p.pol_first, zs);
      uint32_t cw[32];      // this row's 16 filter + 16 gate conditioning values (fp32, one 128-byte line): four 256-bit loads in flight during the wait
      if (COND) {
        const int fr = valid ? cond_frame(tau - p.s_out, p.L - p.s_out, p.cond_frames) : 0;
        const uint4* cp16 = p.cond16 + (((((int64_t)b * p.cond_layers + p.cond_layer) * 4 + cg) * 4) * p.cond_frames + fr) * 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t t8[8];
          ldg_stream32(cp16 + (int64_t)q * p.cond_frames * 2, 0, t8);
#pragma unroll
          for (int k = 0; k < 8; ++k) cw[8 * q + k] = t8[k];
        }
      }
      mbar_wait(&fg_full[ph], ph2);
      tc_fence_after();
      uint32_t f[16], g[16];
      tmem_ld16(lane_addr + ph * 128 + cg * 16, f);
      tmem_ld16(lane_addr + ph * 128 + 64 + cg * 16, g);
      tmem_ld_wait();
      // the part of the gate backward that does not need dz: z = t sg, dF = dz * [sg (1 - t^2)], dG = dz * [z (1 - sg)]
      float ca[16], cb[16];
      uint32_t pz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float zo[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float fv = __uint_as_float(f[2 * j + e]), gv = __uint_as_float(g[2 * j + e]);
          if (BIAS) {
            fv += p.bias_fg[cg * 16 + 2 * j + e];
            gv += p.bias_fg[64 + cg * 16 + 2 * j + e];
          }
          if (COND) {
            fv += __uint_as_float(cw[2 * j + e]);
            gv += __uint_as_float(cw[16 + 2 * j + e]);
          }
          const float t = tanh_fast(fv), sg = sigmoid_fast(gv);
          zo[e] = t * sg;
          ca[2 * j + e] = sg * (1.f - t * t);
          cb[2 * j + e] = zo[e] * (1.f - sg);
        }
        pz[j] = valid ? pack_bf16(zo[0], zo[1]) : 0u;
      }
      uint32_t dzv[16];
      if (DENSE) {
        mbar_wait(&dz_full, ph);
        tc_fence_after();
        tmem_ld16(lane_addr + C_DZ + cg * 16, dzv);
        tmem_ld_wait();
      }
      // (opaque to the compiler: without it the bf16 unpack of the loaded words - and with it the wait for the load - is
      //  hoisted to the top of the tile)
      asm volatile("" : "+r"(zs[0]), "+r"(zs[1]), "+r"(zs[2]), "+r"(zs[3]), "+r"(zs[4]), "+r"(zs[5]), "+r"(zs[6]), "+r"(zs[7]));
      uint32_t pf[8], pg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162*>(&zs[j]);
        float dz0 = bf16lo(s2), dz1 = bf16hi(s2);
        if (DENSE) {
          dz0 += __uint_as_float(dzv[2 * j]);
          dz1 += __uint_as_float(dzv[2 * j + 1]);
        }
        pf[j] = valid ? pack_bf16(dz0 * ca[2 * j], dz1 * ca[2 * j + 1]) : 0u;
        pg[j] = valid ? pack_bf16(dz0 * cb[2 * j], dz1 * cb[2 * j + 1]) : 0u;
      }
      // the dF | dG | z tiles may be rewritten once (a) the previous tile's weight-gradient MMAs have read them
      // (out_empty) and (b) its TMA stores have read them (thread 0 checks the bulk group, the barrier publishes it)
      if (it > 0) mbar_wait(&out_empty, ph ^ 1);
      tc_fence_before();
      if (tid == 0) tma_store_wait_read();
      epi8_bar_sync();                     // also: every thread has drained this tile's TMEM accumulators
      if (tid == 0) {
        mbar_arrive(&fg_empty[ph]);
        mbar_arrive(&dz_empty);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t o = sw128_chunk(row, cg * 2 + q);
        sts128(smem_u32(sm + Bwd3Smem::Z) + o, pz[4 * q], pz[4 * q + 1], pz[4 * q + 2], pz[4 * q + 3]);
        sts128(smem_u32(sm + Bwd3Smem::DF) + o, pf[4 * q], pf[4 * q + 1], pf[4 * q + 2], pf[4 * q + 3]);
        sts128(smem_u32(sm + Bwd3Smem::DG) + o, pg[4 * q], pg[4 * q + 1], pg[4 * q + 2], pg[4 * q + 3]);
      }
      fence_proxy_async_smem();
      epi8_bar_sync();
      if (tid == 0) {
        mbar_arrive(&out_full);
        tma_store_3d(&tm_dfg, sm + Bwd3Smem::DF, 0, tau0, b, p.pol_last);          // read by the dx GEMM that follows
        tma_store_3d(&tm_dfg, sm + Bwd3Smem::DG, 64, tau0, b, p.pol_last);
        tma_store_commit();                // checked one phase later (above), nobody waits here
      }
      tl += (int)gridDim.x;
      while (tl >= p.tiles_per_batch) { tl -= p.tiles_per_batch; ++b; }
    }
    if (tid == 0) tma_store_wait_read();
    // ---- flush the weight-gradient accumulators as a per-CTA partial tile [128][192] (see block_bwd2)
    {
      float* prow = pp.partial + ((int64_t)blockIdx.x * 128 + row) * 192;
      if (n_mine > 0) {
        mbar_wait(&wg_done, 0);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_addr + C_WFG + cg * 32, v);         // dW_fg columns [32 cg, 32 cg + 32)
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(prow + cg * 32 + q * 4) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        uint32_t u[16];
        if (DENSE) {
          tmem_ld16(lane_addr + C_WD + cg * 16, u);        // dW_dense columns [16 cg, 16 cg + 16)
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) u[j] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(prow + 128 + cg * 16 + q * 4) = make_uint4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ========================================================================================= block_bwd6 (1/2: the idea)
// block_bwd3 + the data-gradient GEMM of the dilated conv fused in, so that dF | dG (256 B per row and layer, written once and
// read twice by gemm_nt_kernel<64>) never reach HBM.  The obstacle to that fusion is the dilation: row tau of dx_i needs
// W1^T dFG[tau] AND W0^T dFG[tau + d], i.e. rows of another tile.  Here the two products are kept apart until they are consumed:
//     A_i[tau] = dx_{i+1}[tau] + dFG_i[tau] W1_i          Q_i[tau] = dFG_i[tau] W0_i              dx_i[tau] = A_i[tau] + Q_i[tau + d_i]
// and the kernel of layer i - 1 loads the tile A_i[tau0..] and the SHIFTED tile Q_i[tau0 + d_i ..] (TMA, out-of-range rows read as
// zeros) and uses dx_i only through linear maps: dz = A Wd + Q Wd (two accumulating UMMAs), dW_dense = (A + Q)^T z and the residual
// pass-through A_{i-1} = (A + Q) + ... (summed once per tile in fp32 in the epilogue).  No atomics, no zero-filled buffers.
// Per tile:  TMA {x[tau-d], x[tau]} (2 stages), {A', Q'} (2 stages)
//            UMMA  f|g (recompute)  ->  TMEM fg[n & 1];   dz = A' Wd + Q' Wd  ->  TMEM dz
//            epilogue 1: gate backward -> z, dF, dG tiles
//            UMMA  P = [dF|dG] [W0|W1] -> TMEM fg[n & 1] (the drained f|g buffer);  dW_dense += (A' + Q')^T z;  dW_fg += dFG^T [x taps]
//            epilogue 2: A_i = (A' + Q') + P[:, 64:128) over the A' tile, Q_i = P[:, 0:64) over the Q' tile -> TMA stores (store warp)
// The resident forward weights W0 / W1 ([128 o][64 r], K-major B operands of the recompute) are also the MN-major B operand of
// the P product (K = o): no transposed weight copies in shared memory.  Tiles start at the previous layer's first tile so that
// everything layer i - 1 reads has been written in this step (rows below s_out are written as zeros).
// timing experiments (WN_TS=1): clock64 stamps of CTA 0, 16 per tile (tools/ts_bwd.py reads them through wn_debug_ts)
__device__ long long g_ts[16 * 64];
#define WN_TS_MARK(it, k)                                                        \
  do {                                                                           \
    if (p.trace && blockIdx.x == 0 && (it) < 64) g_ts[(it) * 16 + (k)] = clock64(); \
  } while (0)

struct Bwd5Smem {
  static constexpr uint32_t W0 = 0, W1 = TILE, WDT = 2 * TILE;                 // resident weights (40 KB)
  static constexpr uint32_t XR = 2 * TILE + 8192, X_STAGE = 2 * TILE;          // 2 x {x tap0, x tap1}
  static constexpr uint32_t DXR = XR + 2 * X_STAGE, DX_STAGE = 2 * TILE;       // 2 x {A' (-> A_i), Q' (-> Q_i)}
  static constexpr uint32_t DF = DXR + 2 * DX_STAGE, DG = DF + TILE, Z = DG + TILE;
  static constexpr uint32_t TOTAL = Z + TILE;                                  // 216 KB
};

// ========================================================================================= block_bwd6 (2/2: the kernel)
// 19 warps: 0-15 epilogue in TWO GROUPS of 8 that take alternate tiles, 16 TMA producer, 17 MMA issuer, 18 store thread.
// The per-tile resources belong to the tile's parity = its group (TMEM f|g / P buffer, x stage, {A', Q'} stage), so group 1
// runs the gate math of tile n + 1 while group 0 waits for the MMAs of tile n and writes its outputs.  What stays shared is
// sequenced by barriers that complete once per tile, in tile order: the dz accumulator (dz_full / dz_empty) and the
// dF | dG | z staging tiles (out_full / out_empty).  A barrier whose completions are consumed ALTERNATELY by the two groups
// (dz_full, out_empty) is split per tile parity: an mbarrier wait only names the phase parity, and a group that arrives
// early at "completion #1" of a barrier still in phase 0 would pass immediately (the first version of this kernel hung on
// exactly that as soon as a CTA had more than two tiles).  Every thread owns 32 rows x 32 columns as two passes of 16
// (packed results of the first pass are held in registers); dW_dense is accumulated from A' and Q' separately (two MMA
// groups) so that epilogue 2 re-reads both in bf16 and sums them in fp32, exactly as the unfused path adds its bf16 dx.
// Measured (cfg 2, profiles/r2_summary.md): 2.41 ms per step for the 30 launches, against 1.57 (block_bwd3) + 1.06 (dx GEMM)
// of the unfused pair, and 3.3 GB less HBM traffic.  The clock64 timeline (tools/ts_bwd.py) shows what bounds it: with the
// 216 KB of shared memory two {x, A'/Q'} stages are all that fits, each {A', Q'} stage doubles as the output staging area and
// is refilled only after its TMA store has been read, so per parity the loop  load (2-4 k cycles) -> dz -> epilogue 1 (3.2 k)
// -> P (1.5 k) -> epilogue 2 (2.1 k) -> store (1.4 k)  is serial: ~12 k cycles per two tiles.
template <bool DENSE, bool DIRECT>
__global__ void __launch_bounds__(608, 1)
block_bwd6_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w0,
                  const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_a_in,
                  const __grid_constant__ CUtensorMap tm_q_in, const __grid_constant__ CUtensorMap tm_wdT,
                  const __grid_constant__ CUtensorMap tm_a_out, const __grid_constant__ CUtensorMap tm_q_out, BlockBwd2Params pp) {
  const BlockBwdParams& p = pp.b;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t w_full, x_full[2], x_empty[2], dxi_full[2], dxi_empty[2], fg_full[2], fg_empty[2], p_full[2], st_req[2], wd_done[2];
  __shared__ __align__(8) uint64_t dz_full[2], dz_empty, out_full, out_empty[2], wg_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&dxi_full[i], 1);
      mbar_init(&dxi_empty[i], DIRECT ? 9 : 1);      // DIRECT: the group's 8 warps (after their reads) + the dW_dense commit
      mbar_init(&fg_full[i], 1);
      mbar_init(&fg_empty[i], 1);
      mbar_init(&p_full[i], 1);
      mbar_init(&st_req[i], 1);
      mbar_init(&wd_done[i], 1);
      mbar_init(&dz_full[i], 1);
      mbar_init(&out_empty[i], 1);
    }
    mbar_init(&dz_empty, 1);
    mbar_init(&out_full, 1);
    mbar_init(&wg_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t sbase = smem_u32(sm);
  const int n_items = pp.n_batches * p.tiles_per_batch;
  const int n_mine = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  constexpr uint32_t C_DZ = 256, C_WFG = 320, C_WD = 448;      // TMEM columns; f|g / P buffers at 0 and 128

  if (warp == 16) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0 && n_mine > 0) {
      mbar_expect_tx(&w_full, 2 * TILE + (DENSE ? 8192 : 0));
      tma_load_2d(sm + Bwd5Smem::W0, &tm_w0, &w_full, 0, 0);
      tma_load_2d(sm + Bwd5Smem::W1, &tm_w1, &w_full, 0, 0);
      if (DENSE) tma_load_2d(sm + Bwd5Smem::WDT, &tm_wdT, &w_full, 0, 0);
      for (int it = 0; it < n_mine; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int b = item / p.tiles_per_batch, tau0 = (p.tile0 + item % p.tiles_per_batch) * 128;
        const int st = it & 1;
        const uint32_t eph = ((it >> 1) & 1) ^ 1;
        if (it + 2 < n_mine) {      // the tile that will use these stages next: pull its boxes into L2 now (a whole stage period ahead),
          const int item2 = item + 2 * (int)gridDim.x;      // so that the loads issued when the stages free up do not pay DRAM latency
          const int b2 = item2 / p.tiles_per_batch, tau2 = (p.tile0 + item2 % p.tiles_per_batch) * 128;
          tma_prefetch_3d(&tm_x, 0, tau2 - p.d, b2);
          tma_prefetch_3d(&tm_x, 0, tau2, b2);
          if (DENSE && tau2 >= p.own_row0) {
            tma_prefetch_3d(&tm_a_in, 0, tau2, b2);
            tma_prefetch_3d(&tm_q_in, 0, tau2 + p.d_next, b2);
          }
        }
        mbar_wait(&x_empty[st], eph);
        uint8_t* sxp = sm + Bwd5Smem::XR + st * Bwd5Smem::X_STAGE;
        mbar_expect_tx(&x_full[st], 2 * TILE);
        tma_load_3d(sxp, &tm_x, &x_full[st], 0, tau0 - p.d, b, p.pol_first);
        tma_load_3d(sxp + TILE, &tm_x, &x_full[st], 0, tau0, b);
        WN_TS_MARK(it, 0);
        if (tau0 >= p.tw_al) {
          const int rb0 = (tau0 - p.tw_al) >> 5, nb = min(4, p.dzs_nblk - rb0);
          if (nb > 0) bulk_prefetch_l2(p.dzs + (((int64_t)(p.dzs_lb0 + b) * p.dzs_nblk + rb0) << 11), (uint32_t)nb << 12);
        }
        if (DENSE) {
          const int row_a = tau0 >= p.own_row0 ? tau0 : p.L, row_q = tau0 >= p.own_row0 ? tau0 + p.d_next : p.L;
          mbar_wait(&dxi_empty[st], eph);
          uint8_t* sd = sm + Bwd5Smem::DXR + st * Bwd5Smem::DX_STAGE;
          mbar_expect_tx(&dxi_full[st], 2 * TILE);
          tma_load_3d(sd, &tm_a_in, &dxi_full[st], 0, row_a, b, p.pol_first);
          tma_load_3d(sd + TILE, &tm_q_in, &dxi_full[st], 0, row_q, b, p.pol_first);
          WN_TS_MARK(it, 1);
        }
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------ MMA issuer (polling)
    if (lane == 0 && n_mine > 0) {
      constexpr uint32_t id_fg = idesc_bf16(128, 128, 0, 0), id_dz = idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t id_wfg = idesc_bf16(128, 128, 1, 1), id_wd = idesc_bf16(128, 64, 1, 1);
      constexpr uint32_t id_p = idesc_bf16(128, 128, 0, 1);
      mbar_wait(&w_full, 0);
      int jf = 0, jd = 0, jw = 0;
      while (jw < n_mine) {
        if (DENSE && jd < n_mine && mbar_test_wait(&dz_empty, (jd & 1) ^ 1) && mbar_test_wait(&dxi_full[jd & 1], (jd >> 1) & 1)) {
          tc_fence_after();
          const uint32_t sd = sbase + Bwd5Smem::DXR + (jd & 1) * Bwd5Smem::DX_STAGE;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + C_DZ, desc_kmajor(sd, k), desc_kmajor(sbase + Bwd5Smem::WDT, k), id_dz, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + C_DZ, desc_kmajor(sd + TILE, k), desc_kmajor(sbase + Bwd5Smem::WDT, k), id_dz, true);
          umma_commit(&dz_full[jd & 1]);
          WN_TS_MARK(jd, 2);
          ++jd;
          continue;
        }
        if (jw < jf && mbar_test_wait(&out_full, jw & 1)) {
          tc_fence_after();
          const uint32_t sxa = sbase + Bwd5Smem::XR + (jw & 1) * Bwd5Smem::X_STAGE;
          auto job_p = [&]() {
#pragma unroll
            for (int k = 0; k < 8; ++k)     // P[t, (tap, r)] = sum_o dFG[t, o] * W_tap[o, r]
              umma_bf16(tmem + (jw & 1) * 128, desc_kmajor(sbase + (k < 4 ? Bwd5Smem::DF : Bwd5Smem::DG), k & 3),
                        desc_mnmajor(sbase + Bwd5Smem::W0, k, TILE), id_p, k > 0);
            umma_commit(&p_full[jw & 1]);
          };
          auto job_wd = [&]() {
            if (DENSE) {      // dW_dense[r, d] += sum_t (A'[t, r] + Q'[t, r]) * z[t, d]   (rows 64..127 unused)
              const uint32_t sd = sbase + Bwd5Smem::DXR + (jw & 1) * Bwd5Smem::DX_STAGE;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16(tmem + C_WD, desc_mnmajor(sd, k, 0), desc_mnmajor(sbase + Bwd5Smem::Z, k, TILE), id_wd, (jw | k) != 0);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16(tmem + C_WD, desc_mnmajor(sd + TILE, k, 0), desc_mnmajor(sbase + Bwd5Smem::Z, k, TILE), id_wd, true);
            }
            if (DIRECT) {
              if (DENSE) umma_commit(&dxi_empty[jw & 1]);     // dW_dense has read A' and Q' (the stage's other 8 arrivals: the group's warps)
            } else {
              umma_commit(&wd_done[jw & 1]);     // dW_dense has read A' and Q': epilogue 2 may overwrite them in place
            }
          };
          auto job_wfg = [&]() {
#pragma unroll
            for (int k = 0; k < 8; ++k)     // dW_fg[o, (tap, r)] += sum_t dFG[t, o] * x[t - (1 - tap) d, r]
              umma_bf16(tmem + C_WFG, desc_mnmajor(sbase + Bwd5Smem::DF, k, TILE), desc_mnmajor(sxa, k, TILE), id_wfg, (jw | k) != 0);
            umma_commit(&x_empty[jw & 1]);
          };
          if (pp.w_order == 1) {      // stages first: the {x} and {A', Q'} stages are what the tile loop waits for
            job_wfg();
            job_wd();
            job_p();
          } else {                    // P first: epilogue 2 starts earliest
            job_p();
            job_wd();
            job_wfg();
          }
          umma_commit(&out_empty[jw & 1]);
          WN_TS_MARK(jw, 3);
          ++jw;
          continue;
        }
        if (jf < n_mine && mbar_test_wait(&x_full[jf & 1], (jf >> 1) & 1) && mbar_test_wait(&fg_empty[jf & 1], ((jf >> 1) & 1) ^ 1)) {
          tc_fence_after();
          const uint32_t sxa = sbase + Bwd5Smem::XR + (jf & 1) * Bwd5Smem::X_STAGE, acc = tmem + (jf & 1) * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc, desc_kmajor(sxa, k), desc_kmajor(sbase + Bwd5Smem::W0, k), id_fg, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc, desc_kmajor(sxa + TILE, k), desc_kmajor(sbase + Bwd5Smem::W1, k), id_fg, true);
          umma_commit(&fg_full[jf & 1]);
          WN_TS_MARK(jf, 4);
          ++jf;
          continue;
        }
      }
      umma_commit(&wg_done);
    }
  } else if (warp == 18) {
    // ------------------------------------------------------------ store thread: tiles in order, alternating between the groups
    if (!DIRECT && lane == 0 && n_mine > 0) {
      int b = (int)blockIdx.x / p.tiles_per_batch, tl = (int)blockIdx.x % p.tiles_per_batch;
      for (int it = 0; it < n_mine; ++it) {
        const int tau0 = (p.tile0 + tl) * 128;
        uint8_t* sd = sm + Bwd5Smem::DXR + (it & 1) * Bwd5Smem::DX_STAGE;
        mbar_spin_wait(&st_req[it & 1], (it >> 1) & 1);
        tma_store_3d(&tm_a_out, sd, 0, tau0, b, p.pol_last);
        tma_store_3d(&tm_q_out, sd + TILE, 0, tau0, b, p.pol_last);
        tma_store_commit();
        tma_store_wait_read();
        WN_TS_MARK(it, 5);
        mbar_arrive(&dxi_empty[it & 1]);
        tl += (int)gridDim.x;
        while (tl >= p.tiles_per_batch) { tl -= p.tiles_per_batch; ++b; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: group g = warps 8g .. 8g+7, tiles it = g, g+2, ...
    const int g = warp >> 3, q4 = warp & 3, h = (warp >> 2) & 1;      // TMEM lane quarter (= warp % 4), 32-column half
    const int row = q4 * 32 + lane;
    const uint32_t lane_addr = tmem_addr(tmem, q4 * 32, 0);
    const bool leader = (tid & 255) == 0;
    auto group_bar = [&]() {
      if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
    };
    uint8_t* sa = sm + Bwd5Smem::DXR + g * Bwd5Smem::DX_STAGE;      // this group's {A' -> A_i, Q' -> Q_i} stage
    int b = 0, tl = (int)blockIdx.x + g * (int)gridDim.x;
    while (tl >= p.tiles_per_batch) { tl -= p.tiles_per_batch; ++b; }
    for (int it = g; it < n_mine; it += 2) {
      const uint32_t ph2 = (it >> 1) & 1;
      const int tau0 = (p.tile0 + tl) * 128;
      const int tau = tau0 + row;
      const bool valid = tau >= p.s_out && tau < p.L;
      const bool any_invalid = __any_sync(0xffffffffu, !valid);
      const bool has_zs = tau0 >= p.tw_al && tau < p.L;
      const __nv_bfloat16* zsp =
          p.dzs + (((((int64_t)(p.dzs_lb0 + b) * p.dzs_nblk + ((tau0 - p.tw_al) >> 5) + q4) * 4 + h * 2) * 32 + lane) << 4);
      // ---------------- epilogue 1: gate backward, two passes of 16 columns
      // The single dz accumulator is what the two groups share inside a tile period: it is drained FIRST - dz + skip-path
      // gradient of both passes, held as packed bf16 (16 registers) - and handed to the other group's tile before any of the
      // gate math, so that dz(n + 1) is computed under the math of tile n.
      uint32_t dzp[16];
      if (leader) WN_TS_MARK(it, 6);
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        uint32_t zs[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (has_zs) ldg_stream32(zsp + ps * 512, p.pol_first, zs);      // next column group of 16: + 32 rows x 16 channels
        uint32_t dzv[16];
        if (DENSE) {
          if (ps == 0) {
            mbar_wait(&dz_full[g], ph2);
            tc_fence_after();
          }
          tmem_ld16(lane_addr + C_DZ + h * 32 + ps * 16, dzv);
          tmem_ld_wait();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162*>(&zs[j]);
          float dz0 = bf16lo(s2), dz1 = bf16hi(s2);
          if (DENSE) {
            dz0 += __uint_as_float(dzv[2 * j]);
            dz1 += __uint_as_float(dzv[2 * j + 1]);
          }
          dzp[ps * 8 + j] = pack_bf16(dz0, dz1);
        }
      }
      if (DENSE) {
        tc_fence_before();
        group_bar();
        if (leader) mbar_arrive(&dz_empty);
      }
      uint32_t keep[24];                       // packed z | dF | dG of the first pass
      uint32_t pz[8], pf[8], pg[8];
      if (leader) WN_TS_MARK(it, 7);
      mbar_wait(&fg_full[g], ph2);
      tc_fence_after();
      if (leader) WN_TS_MARK(it, 8);
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const int c0 = h * 32 + ps * 16;
        uint32_t f[16], gq[16];
        tmem_ld16(lane_addr + g * 128 + c0, f);
        tmem_ld16(lane_addr + g * 128 + 64 + c0, gq);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162 d2 = *reinterpret_cast<const __nv_bfloat162*>(&dzp[ps * 8 + j]);
          const float dz[2] = {bf16lo(d2), bf16hi(d2)};
          float zo[2], df[2], dg[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float fv = __uint_as_float(f[2 * j + e]), gv = __uint_as_float(gq[2 * j + e]);
            const float t = tanh_fast(fv), sg = sigmoid_fast(gv);
            zo[e] = t * sg;
            df[e] = dz[e] * (sg * (1.f - t * t));
            dg[e] = dz[e] * (zo[e] * (1.f - sg));
          }
          pz[j] = pack_bf16(zo[0], zo[1]);
          pf[j] = pack_bf16(df[0], df[1]);
          pg[j] = pack_bf16(dg[0], dg[1]);
        }
        if (any_invalid && !valid) {      // rows outside the layer's valid range are stored as zeros (boundary tiles only)
#pragma unroll
          for (int j = 0; j < 8; ++j) { pz[j] = 0u; pf[j] = 0u; pg[j] = 0u; }
        }
        if (ps == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { keep[j] = pz[j]; keep[8 + j] = pf[j]; keep[16 + j] = pg[j]; }
        }
      }
      if (!DENSE) {      // (dz_empty still completes once per tile: the issuer's job order does not depend on DENSE)
        tc_fence_before();
        group_bar();
        if (leader) mbar_arrive(&dz_empty);
      }
      // the dF | dG | z tiles are free once the MMAs of tile it - 1 (the other group's) have read them
      if (leader) WN_TS_MARK(it, 9);
      if (it > 0) mbar_wait(&out_empty[g ^ 1], ((it - 1) >> 1) & 1);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t o0 = sw128_chunk(row, h * 4 + q), o1 = sw128_chunk(row, h * 4 + 2 + q);
        sts128(smem_u32(sm + Bwd5Smem::Z) + o0, keep[4 * q], keep[4 * q + 1], keep[4 * q + 2], keep[4 * q + 3]);
        sts128(smem_u32(sm + Bwd5Smem::DF) + o0, keep[8 + 4 * q], keep[9 + 4 * q], keep[10 + 4 * q], keep[11 + 4 * q]);
        sts128(smem_u32(sm + Bwd5Smem::DG) + o0, keep[16 + 4 * q], keep[17 + 4 * q], keep[18 + 4 * q], keep[19 + 4 * q]);
        sts128(smem_u32(sm + Bwd5Smem::Z) + o1, pz[4 * q], pz[4 * q + 1], pz[4 * q + 2], pz[4 * q + 3]);
        sts128(smem_u32(sm + Bwd5Smem::DF) + o1, pf[4 * q], pf[4 * q + 1], pf[4 * q + 2], pf[4 * q + 3]);
        sts128(smem_u32(sm + Bwd5Smem::DG) + o1, pg[4 * q], pg[4 * q + 1], pg[4 * q + 2], pg[4 * q + 3]);
      }
      fence_proxy_async_smem();
      group_bar();                           // the group's 8 warps together wrote the whole [128][64] tile of each of z, dF, dG
      if (leader) {
        mbar_arrive(&out_full);
        WN_TS_MARK(it, 10);
      }
      if (DIRECT) {
        // ---------------- epilogue 2, direct: the residual pass-through A' + Q' is read (fp32 sums held in registers) BEFORE the
        // wait for P, so the {A', Q'} stage goes back to the producer as soon as dW_dense has read it too - the refill of the
        // stage no longer waits for P, epilogue 2 and a TMA store out of the same shared memory; A_i / Q_i go from registers
        // to global memory, one full 32-byte sector per thread and request (each thread owns 64 contiguous bytes of a row)
        float sum[32];
        if (DENSE) {
          mbar_wait(&dxi_full[g], ph2);                            // (TMA data of this stage observed by this thread as well)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t o = sw128_chunk(row, h * 4 + q);
            const uint4 a4 = lds128(smem_u32(sa) + o), q4v = lds128(smem_u32(sa + TILE) + o);
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, qw[4] = {q4v.x, q4v.y, q4v.z, q4v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
              const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(&qw[j]);
              sum[q * 8 + 2 * j] = bf16lo(a2) + bf16lo(q2);
              sum[q * 8 + 2 * j + 1] = bf16hi(a2) + bf16hi(q2);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&dxi_empty[g]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = 0.f;
        }
        mbar_wait(&p_full[g], ph2);
        tc_fence_after();
        if (leader) WN_TS_MARK(it, 11);
        const bool in_range = tau < p.L;
        __nv_bfloat16* arow = pp.a_out_p + ((int64_t)b * p.L + tau) * 64 + h * 32;
        __nv_bfloat16* qrow = pp.q_out_p + ((int64_t)b * p.L + tau) * 64 + h * 32;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int c0 = h * 32 + ps * 16;
          uint32_t p0[16], p1[16], pa[8], pq[8];
          tmem_ld16(lane_addr + g * 128 + c0, p0);
          tmem_ld16(lane_addr + g * 128 + 64 + c0, p1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pa[j] = pack_bf16(sum[ps * 16 + 2 * j] + __uint_as_float(p1[2 * j]), sum[ps * 16 + 2 * j + 1] + __uint_as_float(p1[2 * j + 1]));
            pq[j] = pack_bf16(__uint_as_float(p0[2 * j]), __uint_as_float(p0[2 * j + 1]));
          }
          if (any_invalid && !valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { pa[j] = 0u; pq[j] = 0u; }
          }
          if (in_range) {
            stg32(arow + ps * 16, p.pol_last, pa);
            stg32(qrow + ps * 16, p.pol_last, pq);
          }
        }
        tc_fence_before();
        group_bar();                           // every thread of the group has drained P
        if (leader) {
          mbar_arrive(&fg_empty[g]);
          WN_TS_MARK(it, 12);
        }
      } else {
      // ---------------- epilogue 2: A_i = (A' + Q') + P1, Q_i = P0, in place over A' / Q'
      mbar_wait(&p_full[g], ph2);            // P complete; dW_dense has read A' and Q'
      if (DENSE) mbar_wait(&dxi_full[g], ph2);                  // (TMA data of this stage observed by this thread as well)
      else mbar_wait(&dxi_empty[g], ph2 ^ 1);                   // no producer waits for the staging slots: the store of tile it - 2
      tc_fence_after();
      if (leader) WN_TS_MARK(it, 11);
      uint32_t pa[16], pq[16];               // packed A_i / Q_i of both passes: written once dW_dense has read A' and Q'
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const int c0 = h * 32 + ps * 16;
        uint32_t p0[16], p1[16];
        tmem_ld16(lane_addr + g * 128 + c0, p0);
        tmem_ld16(lane_addr + g * 128 + 64 + c0, p1);
        const uint32_t o0 = sw128_chunk(row, h * 4 + ps * 2), o1 = sw128_chunk(row, h * 4 + ps * 2 + 1);
        uint32_t aw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, qw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (DENSE) {
          const uint4 a0 = lds128(smem_u32(sa) + o0), a1 = lds128(smem_u32(sa) + o1);
          const uint4 q0 = lds128(smem_u32(sa + TILE) + o0), q1 = lds128(smem_u32(sa + TILE) + o1);
          aw[0] = a0.x; aw[1] = a0.y; aw[2] = a0.z; aw[3] = a0.w; aw[4] = a1.x; aw[5] = a1.y; aw[6] = a1.z; aw[7] = a1.w;
          qw[0] = q0.x; qw[1] = q0.y; qw[2] = q0.z; qw[3] = q0.w; qw[4] = q1.x; qw[5] = q1.y; qw[6] = q1.z; qw[7] = q1.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
          const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(&qw[j]);
          const float s0 = bf16lo(a2) + bf16lo(q2), s1 = bf16hi(a2) + bf16hi(q2);
          pa[ps * 8 + j] = valid ? pack_bf16(s0 + __uint_as_float(p1[2 * j]), s1 + __uint_as_float(p1[2 * j + 1])) : 0u;
          pq[ps * 8 + j] = valid ? pack_bf16(__uint_as_float(p0[2 * j]), __uint_as_float(p0[2 * j + 1])) : 0u;
        }
      }
      if (DENSE) mbar_wait(&wd_done[g], ph2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t o = sw128_chunk(row, h * 4 + q);
        sts128(smem_u32(sa) + o, pa[4 * q], pa[4 * q + 1], pa[4 * q + 2], pa[4 * q + 3]);
        sts128(smem_u32(sa + TILE) + o, pq[4 * q], pq[4 * q + 1], pq[4 * q + 2], pq[4 * q + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      group_bar();                           // every thread of the group has drained P and written its part of both output tiles
      if (leader) {
        mbar_arrive(&fg_empty[g]);
        mbar_arrive(&st_req[g]);
        WN_TS_MARK(it, 12);
      }
      }
      tl += 2 * (int)gridDim.x;
      while (tl >= p.tiles_per_batch) { tl -= p.tiles_per_batch; ++b; }
    }
    // ---- flush the weight-gradient accumulators as a per-CTA partial tile [128][192]; all 16 warps
    {
      const int cg = warp >> 2;
      float* prow = pp.partial + ((int64_t)blockIdx.x * 128 + row) * 192;
      if (n_mine > 0) {
        mbar_wait(&wg_done, 0);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_addr + C_WFG + cg * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(prow + cg * 32 + q * 4) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        uint32_t u[16];
        if (DENSE) {
          tmem_ld16(lane_addr + C_WD + cg * 16, u);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) u[j] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(prow + 128 + cg * 16 + q * 4) = make_uint4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// dx_0[tau] = A_0[tau] + Q_0[tau + d_0]: the input gradient of the causal layer's weight-gradient kernel (the only consumer
// of a data gradient that is not a block_bwd6 launch)
__global__ void __launch_bounds__(256) combine_dx0_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Qd,
                                                          __nv_bfloat16* __restrict__ out, int L, int d0) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)L * 8;            // 8 chunks of 8 channels per row
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int tau = (int)(e >> 3), c = (int)(e & 7) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(A + ((int64_t)b * L + tau) * 64 + c);
    uint4 q = make_uint4(0, 0, 0, 0);
    if (tau + d0 < L) q = *reinterpret_cast<const uint4*>(Qd + ((int64_t)b * L + tau + d0) * 64 + c);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, qw[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
      const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(&qw[j]);
      o[j] = pack_bf16(bf16lo(a2) + bf16lo(q2), bf16hi(a2) + bf16hi(q2));
    }
    *reinterpret_cast<uint4*>(out + ((int64_t)b * L + tau) * 64 + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// dW = sum over CTAs of the partial tiles written by block_bwd2 / block_bwd3 (fixed summation order: deterministic)
struct WgradReduceArgs {
  int64_t filt0, gate0, dense0, layer_stride;     // flat-vector offsets of layer 0's filter / gate / dense weights
  int n_layers, R, D;                              // real (unpadded) residual / dilation channel counts
  int n_ctas[64];                                  // partial tiles written per layer
  int layer0;                                      // blockIdx.y = 0 is this layer
};
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial_all, int64_t layer_pitch, WgradReduceArgs a,
                                                           float* __restrict__ G) {
  // blockIdx.y = layer; a block sums 32 consecutive elements of a tile row: 8 threads x float4 per partial tile, 32 groups of partial
  // tiles (group g takes tiles g, g + 32, ...), then the 32 group sums in fixed order - deterministic, and the same order of additions
  // as the first version of this kernel (one float per thread: 2.6 TB/s; this one keeps four 16-byte loads per thread in flight)
  const int layer = a.layer0 + blockIdx.y;
  const float* partial = partial_all + (int64_t)layer * layer_pitch;
  const int n_ctas = a.n_ctas[layer];
  float* g_filt = G + a.filt0 + (int64_t)layer * a.layer_stride;
  float* g_gate = G + a.gate0 + (int64_t)layer * a.layer_stride;
  float* g_dense = (layer + 1 < a.n_layers) ? G + a.dense0 + (int64_t)layer * a.layer_stride : nullptr;
  const int e0 = blockIdx.x * 32;                      // 192 = 6 x 32: a block never straddles two rows
  const int m = e0 / 192, c0 = e0 % 192;
  if (c0 >= 128 && (m >= 64 || g_dense == nullptr)) return;      // the unused quarter of the dense columns: not even read
  const int e = e0 + (threadIdx.x & 7) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int k = threadIdx.x >> 3;
  for (; k + 96 < n_ctas; k += 128) {
    const float4 v0 = *reinterpret_cast<const float4*>(partial + (int64_t)k * 128 * 192 + e);
    const float4 v1 = *reinterpret_cast<const float4*>(partial + (int64_t)(k + 32) * 128 * 192 + e);
    const float4 v2 = *reinterpret_cast<const float4*>(partial + (int64_t)(k + 64) * 128 * 192 + e);
    const float4 v3 = *reinterpret_cast<const float4*>(partial + (int64_t)(k + 96) * 128 * 192 + e);
    s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
    s.x += v1.x; s.y += v1.y; s.z += v1.z; s.w += v1.w;
    s.x += v2.x; s.y += v2.y; s.z += v2.z; s.w += v2.w;
    s.x += v3.x; s.y += v3.y; s.z += v3.z; s.w += v3.w;
  }
  for (; k < n_ctas; k += 32) {
    const float4 v = *reinterpret_cast<const float4*>(partial + (int64_t)k * 128 * 192 + e);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  __shared__ float red[32][33];
  {
    float* r = &red[threadIdx.x >> 3][(threadIdx.x & 7) * 4];
    r[0] = s.x; r[1] = s.y; r[2] = s.z; r[3] = s.w;
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  float t = 0.f;
#pragma unroll
  for (int g = 0; g < 32; ++g) t += red[g][threadIdx.x];      // fixed order: deterministic
  const int c = c0 + threadIdx.x;
  if (c < 128) {          // column = (tap, r): tap = c / 64, r = c % 64 ; row m = output channel (filter 0..63 | gate 64..127)
    float* base = m < 64 ? g_filt : g_gate;
    if ((m & 63) < a.D && (c & 63) < a.R) base[(int64_t)(m & 63) * (2 * a.R) + (c & 63) * 2 + (c >> 6)] = t;      // (D, R, 2)
  } else if (m < a.R && c - 128 < a.D) {
    g_dense[(int64_t)m * a.D + (c - 128)] = t;                                                                    // (R, D, 1)
  }
}

// ============================================================================================ SIMT helpers
// dlogits (B,Q,W) fp32 -> DLG [B][Wp][Q] bf16 in the padded skip row space (pad rows = 0)
__global__ void __launch_bounds__(256) dlogits_transpose_kernel(const float* __restrict__ dl, __nv_bfloat16* __restrict__ out, int Q,
                                                                int W, int Wp, int pad) {
  __shared__ float tile[32][257];
  const int b = blockIdx.y, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 warps
  const int tw = j0 + tx - pad;
  const bool in = tw >= 0 && tw < W;
  const float* src = dl + (int64_t)b * Q * W + tw;
  for (int q0 = ty; q0 < Q; q0 += 64) {      // eight 128-byte row segments per warp in flight (one at a time ran at 3.5 TB/s)
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (in && q0 + 8 * u < Q) ? src[(int64_t)(q0 + 8 * u) * W] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (q0 + 8 * u < Q) tile[tx][q0 + 8 * u] = v[u];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r;
    if (j >= Wp) continue;
    __nv_bfloat16* o = out + ((int64_t)b * Wp + j) * Q;
    for (int q = tx * 2; q < Q; q += 64)
      *reinterpret_cast<__nv_bfloat162*>(o + q) = __floats2bfloat162_rn(tile[r][q], tile[r][q + 1]);
  }
}

// dWc (R=64, Q, 2) += scatter of dx0 rows: shared-memory privatised histogram of 64-vectors
__global__ void __launch_bounds__(1024) causal_scatter_bwd_kernel(const int64_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dx0,
                                                                 float* __restrict__ dW, int L, int Q, int rows_per_cta) {
  extern __shared__ float acc[];       // [2][Q][64]
  const int b = blockIdx.y;
  const int t_begin = 1 + blockIdx.x * rows_per_cta, t_end = min(L, t_begin + rows_per_cta);
  for (int e = threadIdx.x; e < 2 * Q * 64; e += blockDim.x) acc[e] = 0.f;
  __syncthreads();
  const int r = threadIdx.x & 63, sub = threadIdx.x >> 6, nsub = blockDim.x >> 6;     // 16 rows in flight
  for (int tau = t_begin + sub; tau < t_end; tau += nsub) {
    const float g = __bfloat162float(dx0[((int64_t)b * L + tau) * 64 + r]);
    const int q0 = (int)idx[(int64_t)b * L + tau - 1], q1 = (int)idx[(int64_t)b * L + tau];
    atomicAdd(&acc[(0 * Q + q0) * 64 + r], g);
    atomicAdd(&acc[(1 * Q + q1) * 64 + r], g);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * Q * 64; e += blockDim.x) {
    const float v = acc[e];
    if (v != 0.f) {
      const int k = e / (Q * 64), q = (e / 64) % Q, rr = e % 64;
      atomicAdd(dW + ((int64_t)rr * Q + q) * 2 + k, v);
    }
  }
}

__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    dst[e] = __bfloat162float(src[e]);
}

// out[c] += sum over b, rows in [row_lo,row_hi) of src[b][row][c]   (bias gradients); C in {64,128,256}
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, int C, int64_t bstride, int row_lo,
                                                          int row_hi, float* __restrict__ out, float* __restrict__ out_hi, int n_valid,
                                                          int pitch) {
  const int b = blockIdx.y;
  const int col = threadIdx.x % C, rsub = threadIdx.x / C, nsub = 256 / C;
  const int r0 = row_lo + blockIdx.x * 512, r1 = min(row_hi, r0 + 512);
  float acc = 0.f;
  for (int r = r0 + rsub; r < r1; r += nsub) acc += __bfloat162float(src[(int64_t)b * bstride + (int64_t)r * pitch + col]);
  if ((col & 63) >= n_valid) return;                               // channel-padded models: only the real channels have a bias
  if (out_hi && col >= 64) atomicAdd(out_hi + col - 64, acc);      // split output: columns [64,128) go to a second vector
  else atomicAdd(out + col, acc);
}
// every layer's skip bias receives the same gradient (the skip outputs are summed): copy slot 0 to the others
__global__ void replicate_kernel(float* __restrict__ G, const int64_t* __restrict__ offs, int n, int C) {
  const int i = blockIdx.x + 1, c = threadIdx.x;
  if (i < n && c < C) G[offs[i] + c] = G[offs[0] + c];
}

// out[b][frame][col_map(c)] += sum over the rows tl of a (B, rows, pitch) bf16 tensor that belong to `frame` (the transpose of the
// conditioning broadcast, model1.py:227-247).  Rows t0 + tl, tl in [0, len); 128 columns starting at c0 per launch slab.
// grid (frames, segments, B), 64 threads = 64 column pairs: a CTA walks ONE frame's rows - tl = f + j frames when the encoding is
// tiled along time, tl = f (len / frames) + j when every frame is held for len / frames steps - so each thread keeps its two sums
// in registers (no shared-memory atomics: the first version spent 47 us per layer in them), every row is one 256-byte request and
// a segment ends in 2 atomics per thread.  dd > 0: columns are the kernels' padded [filter 64 | gate 64] order and go to the
// autoencoder's raw (2 dd) order, gate first; dd == 0: identity.
__global__ void __launch_bounds__(64) frame_sum_bf16_kernel(const __nv_bfloat16* __restrict__ src, int64_t bstride, int pitch, int c0,
                                                            int t0, int len, int frames, float* __restrict__ out, int out_pitch, int dd) {
  const int f = blockIdx.x, seg = blockIdx.y, nseg = gridDim.y, b = blockIdx.z;
  const bool held = len % frames == 0;
  const int m = held ? len / frames : (len - f + frames - 1) / frames;      // rows of this frame
  const int64_t first = held ? (int64_t)f * m : f, step = held ? 1 : frames;
  const int per = (m + nseg - 1) / nseg, j0 = seg * per, j1 = min(m, j0 + per);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(src + (int64_t)b * bstride + (int64_t)t0 * pitch + c0) + threadIdx.x;
  const int64_t pitch2 = pitch / 2;
  float a0 = 0.f, a1 = 0.f;
  int j = j0;
  for (; j + 4 <= j1; j += 4) {
    uint32_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(base + (first + (int64_t)(j + u) * step) * pitch2);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&v[u]);
      a0 += bf16lo(h);
      a1 += bf16hi(h);
    }
  }
  for (; j < j1; ++j) {
    const uint32_t v = __ldg(base + (first + (int64_t)j * step) * pitch2);
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&v);
    a0 += bf16lo(h);
    a1 += bf16hi(h);
  }
  if (j1 <= j0) return;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int c = 2 * threadIdx.x + e;
    int oc = c0 + c;
    if (dd > 0) {
      const int ch = c & 63;
      if (ch >= dd) continue;
      oc = c < 64 ? dd + ch : ch;      // filter -> second half, gate -> first half
    }
    atomicAdd(out + ((int64_t)b * frames + f) * out_pitch + oc, e ? a1 : a0);
  }
}

// raw conditioning conv output (B * frames, 2 dd), gate first (model1.py:188-192) [+ the conv bias (2 dd) of the conditioned layer]
// -> the kernels' table slot: out[row * out_stride + c], c < 64: filter channel c, c >= 64: gate channel c - 64, zero padded
__global__ void cond_table_kernel(const float* __restrict__ raw, const float* __restrict__ bias, int64_t n_rows, int dd,
                                  float* __restrict__ out, int64_t out_stride) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_rows * 128; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e >> 7;
    const int c = (int)(e & 127), ch = c & 63;
    float v = 0.f;
    if (ch < dd) {
      const int rc = c < 64 ? dd + ch : ch;
      v = raw[row * 2 * dd + rc] + (bias ? bias[rc] : 0.f);
    }
    out[row * out_stride + c] = v;
  }
}
// table rows [128] = [filter 64 | gate 64] per (b, frame, layer)  ->  the block kernels' load order
//   out[b][layer][column group 4][chunk 4][frame][8 floats],   chunk q = floats [8q, 8q + 8) of a thread's {16 filter | 16 gate} values
// A warp of the epilogue holds 32 consecutive rows = (mostly) consecutive frames of ONE (layer, column group): with this order its
// 256-bit load of chunk q touches 32 consecutive 32-byte sectors = 8 lines.  (Frame-major rows - one 128-byte line per thread -
// cost 32 lines per request and ~10 us per launch at the autoencoder's shape.)
__global__ void cond_pack16_kernel(const float* __restrict__ tab, float* __restrict__ out, int64_t n_rows, int frames, int layers) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_rows * 128; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e >> 7;                                  // (b * frames + f) * layers + layer
    const int o = (int)(e & 127), cg = o >> 5, w = o & 31, q = w >> 3, k = w & 7;
    const int layer = (int)(row % layers);
    const int64_t bf = row / layers;
    const int f = (int)(bf % frames);
    const int64_t b = bf / frames;
    const float v = tab[row * 128 + (w < 16 ? cg * 16 + w : 64 + cg * 16 + (w - 16))];
    out[(((((b * layers + layer) * 4 + cg) * 4 + q) * frames) + f) * 8 + k] = v;
  }
}
// out[row][c] = raw[row][c] + bias[c]   (head conditioning + connection_1 bias)
__global__ void add_bias_rows_kernel(const float* __restrict__ raw, const float* __restrict__ bias, int64_t n_rows, int C, float* __restrict__ out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_rows * C; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = raw[e] + (bias ? bias[e % C] : 0.f);
}

template <typename K>
int set_smem(K kernel, int bytes) {
  WN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return WN_OK;
}
// same, once per kernel (for call sites that pick one of several instantiations at run time)
template <typename K>
int set_smem_once(K kernel, int bytes) {
  static std::vector<const void*> done;
  const void* key = reinterpret_cast<const void*>(kernel);
  for (const void* d : done)
    if (d == key) return WN_OK;
  WN_PROPAGATE(set_smem(kernel, bytes));
  done.push_back(key);
  return WN_OK;
}

int launch_colsum_bf16(const void* src, int C, int B, int64_t rows_per_batch, int row_lo, int row_hi, float* out, cudaStream_t s,
                       int n_valid = 64, int pitch = 0) {
  if (row_hi <= row_lo) return WN_OK;
  if (pitch == 0) pitch = C;
  dim3 grid((unsigned)ceil_div(row_hi - row_lo, 512), (unsigned)B);
  WN_PROF("colsum_bias", s);
  colsum_bf16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), C, rows_per_batch * pitch, row_lo, row_hi, out,
                                          nullptr, n_valid, pitch);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
// dFG (128 columns): filter-bias gradient from columns [0,64), gate-bias gradient from [64,128)
int launch_colsum_bf16_split(const void* src, int B, int64_t rows_per_batch, int row_lo, float* out_f, float* out_g, cudaStream_t s,
                             int n_valid) {
  const int row_hi = (int)rows_per_batch;
  if (row_hi <= row_lo) return WN_OK;
  dim3 grid((unsigned)ceil_div(row_hi - row_lo, 512), (unsigned)B);
  WN_PROF("colsum_bias", s);
  colsum_bf16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), 128, rows_per_batch * 128, row_lo, row_hi,
                                          out_f, out_g, n_valid, 128);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace

// environment switches of the backward (timing experiments; read once)
struct BwdEnv {
  bool nt_stream, bwd6, bwd6_tma, l2hint_dx, scatter_simt;
  int bwd6_worder, wgrad_side;      // wgrad_side: 1 on, 0 off, -1 auto
};
static const BwdEnv& bwd_env() {
  static const BwdEnv e = [] {
    auto on = [](const char* name) { const char* v = getenv(name); return v && v[0] == '1'; };
    auto off = [](const char* name) { const char* v = getenv(name); return v && v[0] == '0'; };
    BwdEnv r{};
    r.nt_stream = on("WN_NT_STREAM");
    r.bwd6 = !off("WN_BWD6");             // block_bwd6 (default): the dx GEMM fused into the block backward; WN_BWD6=0: block_bwd3 + dx GEMM
    r.bwd6_tma = on("WN_BWD6_TMA");       // A_i / Q_i staged in shared memory and written by TMA (the first version of block_bwd6)
    r.bwd6_worder = off("WN_BWD6_WORDER") ? 0 : 1;      // 1 (default): dW_fg, dW_dense, P - the stage releases first (2.33 vs 2.39 ms per step)
    // per-layer weight-gradient reductions on a side stream: on for conditioned models (their frame sums run there too); for the
    // plain model ONE reduction after the last block kernel is ~1 % faster (same box: 5.25-5.28 vs 5.30-5.34 ms per step) - the
    // reduction CTAs cannot co-reside with the 608-thread persistent block kernel and only delay the next launch
    r.wgrad_side = on("WN_WGRAD_SIDE") ? 1 : (off("WN_WGRAD_SIDE") ? 0 : -1);
    r.l2hint_dx = on("WN_L2HINT_DX");
    r.scatter_simt = on("WN_SCATTER_SIMT");
    return r;
  }();
  return e;
}

int launch_gemm_nt(int NT, const GemmNtMaps& m, const GemmNtParams& p, cudaStream_t s) {
  const int n_items = p.n_batches * p.tiles_per_batch * p.n_ntiles;
  if (n_items <= 0) return WN_OK;
  const int grid = std::min(n_items, g_sm_count * (NT == 64 ? NtCfg<64>::CTAS_PER_SM : 1));
  WN_PROF(p.tag ? p.tag : "gemm_nt", s);
  auto pick = [&](auto kernel_plain, auto kernel_mask, auto kernel_add) {
    return p.epi == EPI_PLAIN ? kernel_plain : (p.epi == EPI_MASK ? kernel_mask : kernel_add);
  };
  if (NT == 256 && p.epi <= EPI_ADD && p.nk[0] == 4 && p.nk[1] == 0 && p.n_total > 0 && p.n_total % 64 == 0 && !bwd_env().nt_stream) {
    // K = 256: weight block resident in shared memory, one column tile per CTA
    const int smem = NtResCfg::TOTAL + 1024;
    auto k = pick(gemm_nt_resb_kernel<EPI_PLAIN>, gemm_nt_resb_kernel<EPI_MASK>, gemm_nt_resb_kernel<EPI_ADD>);
    WN_PROPAGATE(set_smem_once(k, smem));
    const int g2 = std::min(g_sm_count, p.n_batches * p.tiles_per_batch * p.n_ntiles);
    WN_CHECK_CUDA(launch_pdl(k, dim3((unsigned)g2), dim3(192), smem, s, m.a[0], m.b[0], m.out, p));
  } else if (NT == 256) {
    const int smem = NtCfg<256>::TOTAL + 1024;
    auto k = pick(gemm_nt_kernel<256, EPI_PLAIN>, gemm_nt_kernel<256, EPI_MASK>, gemm_nt_kernel<256, EPI_ADD>);
    if (p.epi == EPI_RELU) k = gemm_nt_kernel<256, EPI_RELU>;
    else if (p.epi == EPI_RELU_BIAS) k = gemm_nt_kernel<256, EPI_RELU_BIAS>;
    else if (p.epi == EPI_LOGITS) k = gemm_nt_kernel<256, EPI_LOGITS>;
    else if (p.epi == EPI_LOGITS_BIAS) k = gemm_nt_kernel<256, EPI_LOGITS_BIAS>;
    else if (p.epi == EPI_RELU_COND) k = gemm_nt_kernel<256, EPI_RELU_COND>;
    WN_PROPAGATE(set_smem_once(k, smem));
    WN_CHECK_CUDA(launch_pdl(k, dim3((unsigned)grid), dim3(192), smem, s, m.a[0], m.a[1], m.b[0], m.b[1], m.out, p));
  } else if (NT == 64) {
    const int smem = NtCfg<64>::TOTAL + 1024;
    auto k = pick(gemm_nt_kernel<64, EPI_PLAIN>, gemm_nt_kernel<64, EPI_MASK>, gemm_nt_kernel<64, EPI_ADD>);
    WN_PROPAGATE(set_smem_once(k, smem));
    WN_CHECK_CUDA(launch_pdl(k, dim3((unsigned)grid), dim3(192), smem, s, m.a[0], m.a[1], m.b[0], m.b[1], m.out, p));
  } else {
    set_error("launch_gemm_nt: NT=%d", NT);
    return WN_ERR_INVALID;
  }
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_gemm_tn(int NB, const GemmTnMaps& m, const GemmTnParams& p, cudaStream_t s) {
  const int n_items = p.n_batches * p.tiles_per_batch;
  if (n_items <= 0) return WN_OK;
  int gx = std::min(n_items, g_sm_count), gy = 1;
  if (p.y_layers > 0) {
    gy = (p.n_mtiles > 0 ? p.n_mtiles : 2) * (int)ceil_div(p.y_layers, 4);
    gx = std::max(1, std::min(n_items, g_sm_count / gy));
  }
  const dim3 grid((unsigned)gx, (unsigned)gy);
  WN_PROF(p.tag ? p.tag : "gemm_tn", s);
  WN_REQUIRE(NB == 4, WN_ERR_INVALID, "launch_gemm_tn: NB=%d", NB);
  const int smem = TnCfg<4>::TOTAL + 1024;
  WN_PROPAGATE(set_smem_once(gemm_tn_kernel<4>, smem));
  WN_CHECK_CUDA(launch_pdl(gemm_tn_kernel<4>, grid, dim3(192), smem, s, m.a, m.b[0], m.b[1], p));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// block backward with the weight gradients accumulated in TMEM (block_bwd3)
int launch_block_bwd2(const BlockBwdMaps& m, const BlockBwd2Params& p, cudaStream_t s) {
  const int n_items = p.n_batches * p.b.tiles_per_batch;
  if (n_items <= 0) return WN_OK;
  const int n_ctas = std::min(n_items, g_sm_count);
  const bool bias = p.b.bias_fg != nullptr, dense = p.b.has_dense != 0;
  const int smem3 = Bwd3Smem::TOTAL + 1024;
  WN_PROF("block_bwd3", s);
  auto k3 = bias ? (dense ? block_bwd3_kernel<true, true, false> : block_bwd3_kernel<true, false, false>)
                 : (dense ? block_bwd3_kernel<false, true, false> : block_bwd3_kernel<false, false, false>);
  if (p.b.cond) {      // conditioned decoder of the autoencoder (its conv biases are folded into the conditioning table)
    WN_REQUIRE(!bias, WN_ERR_INVALID, "conditioned block backward: fold the bias into the conditioning table");
    k3 = dense ? block_bwd3_kernel<false, true, true> : block_bwd3_kernel<false, false, true>;
  }
  WN_PROPAGATE(set_smem_once(k3, smem3));
  WN_CHECK_CUDA(launch_pdl(k3, dim3((unsigned)n_ctas), dim3(576), smem3, s, m.x, m.w0, m.w1, m.dx, m.wdT, m.dfg, p));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// block backward with the dilated conv's data-gradient GEMM fused in (block_bwd6)
int launch_block_bwd6(const BlockBwdMaps& m, const BlockBwd2Params& p, cudaStream_t s) {
  const int n_items = p.n_batches * p.b.tiles_per_batch;
  if (n_items <= 0) return WN_OK;
  const int n_ctas = std::min(n_items, g_sm_count);
  const int smem = Bwd5Smem::TOTAL + 1024;
  WN_REQUIRE(p.b.bias_fg == nullptr, WN_ERR_INVALID, "block_bwd6 serves models without bias");
  const bool direct = !bwd_env().bwd6_tma;
  auto k = p.b.has_dense ? (direct ? block_bwd6_kernel<true, true> : block_bwd6_kernel<true, false>)
                         : (direct ? block_bwd6_kernel<false, true> : block_bwd6_kernel<false, false>);
  WN_PROPAGATE(set_smem_once(k, smem));
  WN_PROF("block_bwd6", s);
  WN_CHECK_CUDA(launch_pdl(k, dim3((unsigned)n_ctas), dim3(608), smem, s, m.x, m.w0, m.w1, m.a_in, m.q_in, m.wdT, m.a_out, m.q_out, p));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_frame_sum_bf16(const void* src, int64_t rows_per_batch, int pitch, int c0, int B, int t0, int len, int frames, float* out,
                          int out_pitch, int dd, cudaStream_t s) {
  if (len <= 0) return WN_OK;
  WN_REQUIRE(pitch % 2 == 0 && c0 % 2 == 0, WN_ERR_INVALID, "frame sum: odd pitch / column offset");
  const int rows_per_frame = (int)ceil_div(len, frames);
  const int nseg = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows_per_frame, 32), ceil_div(16 * g_sm_count, (int64_t)frames * B)));
  WN_PROF("frame_sum", s);
  frame_sum_bf16_kernel<<<dim3((unsigned)frames, (unsigned)nseg, (unsigned)B), 64, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), rows_per_batch * pitch, pitch, c0, t0, len, frames, out, out_pitch, dd);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_cond_table(const float* raw, const float* bias, int64_t n_rows, int dd, float* out, int64_t out_stride, cudaStream_t s) {
  cond_table_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_rows * 128, 256), 1184), 256, 0, s>>>(raw, bias, n_rows, dd, out, out_stride);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_cond_pack16(const float* tab, void* out, int64_t n_rows, int frames, int layers, cudaStream_t s) {
  cond_pack16_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_rows * 128, 256), 1184), 256, 0, s>>>(tab, reinterpret_cast<float*>(out), n_rows, frames,
                                                                                                      layers);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_add_bias_rows(const float* raw, const float* bias, int64_t n_rows, int C, float* out, cudaStream_t s) {
  add_bias_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_rows * C, 256), 1184), 256, 0, s>>>(raw, bias, n_rows, C, out);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// skip sum + head for skip widths other than 256 (the fused skip_head_kernel keeps its [128 x 256] accumulators in TMEM and is
// written for S = 256): three GEMMs of the streaming kernel with fused epilogues, intermediates h0 / h1 in HBM (they are
// saved for the backward anyway):  h0 = relu(Zcat Wskip_cat^T + b) ; h1 = relu(h0 P1^T + b1) ; logits = h1 P2^T + b2   (model.py:134-138)
int head_forward_generic(const Model& m, const BwdMaps& M, const SkipHeadMaps& H, const SkipHeadParams& hp, int B, int L, cudaStream_t s) {
  const int Wpad = skip_wp(m, L), N = m.n_layers, S = m.S;
  const bool bias = hp.bias_skip != nullptr;
  GemmNtMaps gm{};
  GemmNtParams gp{};
  gp.n_batches = B; gp.tile0 = 0; gp.tiles_per_batch = (int)ceil_div(Wpad, 128);
  gp.row_lo = 0; gp.row_hi = Wpad; gp.nk[1] = 0;
  // h0
  gm.a[0] = M.zcat; gm.a[1] = M.zcat; gm.b[0] = H.wsk; gm.b[1] = H.wsk; gm.out = M.h0;
  gp.n_ntiles = S / 256; gp.n_total = S; gp.nk[0] = N;
  gp.epi = bias ? EPI_RELU_BIAS : EPI_RELU; gp.bias = hp.bias_skip; gp.tag = "head_skip_gemm";
  WN_PROPAGATE(launch_gemm_nt(256, gm, gp, s));
  WN_DEBUG_SYNC("head skip gemm", s);
  // h1
  gm.a[0] = M.h0; gm.a[1] = M.h0; gm.b[0] = H.p1; gm.b[1] = H.p1; gm.out = M.h1;
  gp.nk[0] = S / 64; gp.bias = hp.bias_p1; gp.tag = "head_p1_gemm";
  if (m.cond_head) {      // autoencoder decoder: + cond_N[frame] before the ReLU (model1.py:216-219); connection_1's bias is folded in
    gp.epi = EPI_RELU_COND; gp.cond = m.cond_head; gp.cond_frames = m.cond_frames;
    gp.lg_W = hp.W; gp.lg_pad = hp.pad;
  }
  WN_PROPAGATE(launch_gemm_nt(256, gm, gp, s));
  WN_DEBUG_SYNC("head p1 gemm", s);
  // logits
  gm.a[0] = M.h1; gm.a[1] = M.h1; gm.b[0] = H.p2; gm.b[1] = H.p2; gm.out = M.h1;      // (out map unused)
  gp.n_ntiles = 1; gp.n_total = 256; gp.nk[0] = S / 64; gp.cond = nullptr;
  gp.epi = bias ? EPI_LOGITS_BIAS : EPI_LOGITS; gp.bias = hp.bias_p2;
  gp.logits = hp.logits; gp.lg_W = hp.W; gp.lg_Q = hp.Q; gp.lg_pad = hp.pad; gp.tag = "head_p2_gemm";
  WN_PROPAGATE(launch_gemm_nt(256, gm, gp, s));
  WN_DEBUG_SYNC("head p2 gemm", s);
  return WN_OK;
}

}  // namespace wn
namespace wn { int fast_gen_debug_ts(long long* h_buf, int n); }
extern "C" int wn_debug_ts(long long* h_buf, int32_t n) {      // timing experiments: the clock64 stamps of block_bwd6's CTA 0
  if (h_buf && n < 0 && n >= -3 * 16 * 64) return wn::fast_gen_debug_ts(h_buf, -n);      // n < 0: the stamps of gen_pipe_kernel (fast_gen.cu)
  if (!h_buf || n <= 0 || n > 16 * 64) return WN_ERR_INVALID;
  cudaError_t e = cudaMemcpyFromSymbol(h_buf, wn::g_ts, (size_t)n * sizeof(long long));
  return e == cudaSuccess ? WN_OK : WN_ERR_CUDA;
}
namespace wn {

// =============================================================================================== host
int build_bwd_maps(const Model& m, const PackLayout& pl, const WsLayout& wl, int B, int L, const uint8_t* P, uint8_t* Wp,
                   const std::vector<CUtensorMap>& xm, BwdMaps* out) {
  const int N = m.n_layers, Wpad = skip_wp(m, L);
  out->layer.assign(N, BwdLayerMaps{});
  for (int i = 0; i < N; ++i) {
    BwdLayerMaps& l = out->layer[i];
    l.x = xm[i];
    WN_PROPAGATE(tmap_2d(&l.w0, P + pl.wfg0 + (size_t)i * 128 * 64 * 2, 64, 128, 64, 128));
    WN_PROPAGATE(tmap_2d(&l.w1, P + pl.wfg1 + (size_t)i * 128 * 64 * 2, 64, 128, 64, 128));
    WN_PROPAGATE(tmap_2d(&l.wdT, P + pl.wdT + (size_t)i * 64 * 64 * 2, 64, 64, 64, 64));
    WN_PROPAGATE(tmap_2d(&l.wfgT0, P + pl.wfgT0 + (size_t)i * 64 * 128 * 2, 128, 64, 128, 64));
    WN_PROPAGATE(tmap_2d(&l.wfgT1, P + pl.wfgT1 + (size_t)i * 64 * 128 * 2, 128, 64, 128, 64));
  }
  auto skip3 = [&](CUtensorMap* t, size_t off, int cols) {
    return tmap_3d(t, Wp + off, cols, Wpad, B, cols, (uint64_t)Wpad * cols, 128);
  };
  WN_PROPAGATE(skip3(&out->dlg, wl.DLG, 256));
  WN_PROPAGATE(skip3(&out->h0, wl.H0, m.S));
  WN_PROPAGATE(skip3(&out->h1, wl.H1, m.S));
  WN_PROPAGATE(skip3(&out->dh1, wl.DH1, m.S));
  WN_PROPAGATE(skip3(&out->dsk, wl.DSK, m.S));
  WN_PROPAGATE(skip3(&out->zcat, wl.Zcat, 64 * N));
  WN_PROPAGATE(skip3(&out->dzcat, wl.DZcat, 64 * N));
  WN_PROPAGATE(tmap_2d(&out->p1T, P + pl.p1T, m.S, m.S, m.S, 256));                    // [S in][S out]: B of dSK = dH1 P1
  WN_PROPAGATE(tmap_2d(&out->p2T, P + pl.p2T, 256, m.S, 256, 256));                    // [S][Q]:        B of dH1 = dLg P2
  WN_PROPAGATE(tmap_2d(&out->wsTcat, P + pl.wsT, m.S, (uint64_t)64 * N, m.S, 256));    // [64 N][S]:     B of dZcat = dSK Ws
  WN_PROPAGATE(tmap_3d(&out->dxa, Wp + wl.DXa, 64, L, B, 64, (uint64_t)L * 64, 128));
  WN_PROPAGATE(tmap_3d(&out->dxb, Wp + wl.DXb, 64, L, B, 64, (uint64_t)L * 64, 128));
  WN_PROPAGATE(tmap_3d(&out->dfg, Wp + wl.DFG, 128, L, B, 128, (uint64_t)L * 128, 128));
  WN_PROPAGATE(tmap_3d(&out->dfg2, Wp + wl.DFG2, 128, L, B, 128, (uint64_t)L * 128, 128));
  WN_PROPAGATE(tmap_3d(&out->zf, Wp + wl.Zf, 64, L, B, 64, (uint64_t)L * 64, 128));
  // block_bwd6: the Q halves of the split data gradient live in the two halves of the (then unused) dFG buffer
  const size_t half = align_up((size_t)B * L * 64 * 2, 1024);
  WN_PROPAGATE(tmap_3d(&out->dqa, Wp + wl.DFG, 64, L, B, 64, (uint64_t)L * 64, 128));
  WN_PROPAGATE(tmap_3d(&out->dqb, Wp + wl.DFG + half, 64, L, B, 64, (uint64_t)L * 64, 128));
  return WN_OK;
}

// per-device side stream + fork / join events of the backward (one process drives one GPU; a second device gets its own set)
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t dfg_free[2] = {nullptr, nullptr};      // conditioned models: the frame sums over dF|dG buffer k have been taken
};
static int side_stream(SideStream** out) {
  static SideStream tab[64];
  int dev = 0;
  WN_CHECK_CUDA(cudaGetDevice(&dev));
  WN_REQUIRE(dev >= 0 && dev < 64, WN_ERR_UNSUPPORTED, "device index %d", dev);
  SideStream& t = tab[dev];
  if (!t.stream) {
    WN_CHECK_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    WN_CHECK_CUDA(cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming));
    WN_CHECK_CUDA(cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming));
    WN_CHECK_CUDA(cudaEventCreateWithFlags(&t.dfg_free[0], cudaEventDisableTiming));
    WN_CHECK_CUDA(cudaEventCreateWithFlags(&t.dfg_free[1], cudaEventDisableTiming));
  }
  *out = &t;
  return WN_OK;
}

int fast_backward_impl(Model& m, const BwdMaps& M, int B, int L, const float* d_x, const int64_t* d_idx, const void* d_packed,
                       void* d_ws, float* d_dlogits, float* G, cudaStream_t s) {
  const PackLayout pl = pack_layout(m);
  const WsLayout wl = ws_layout(m, B, L);
  const int W = L - m.rf + 1, N = m.n_layers, Wpad = skip_wp(m, L), tw_al = skip_tw_al(m, L), pad = (L - W) - tw_al;
  const uint8_t* P = reinterpret_cast<const uint8_t*>(d_packed);
  uint8_t* Wp = reinterpret_cast<uint8_t*>(d_ws);
  const bool bias = m.use_bias != 0;
  const BwdEnv& env = bwd_env();
  const bool dz_tiled = !env.nt_stream;       // the block backward reads the skip-path gradient from the tiled layout (GemmNtParams::out_tiled)
  WN_REQUIRE(dz_tiled, WN_ERR_UNSUPPORTED, "WN_NT_STREAM=1 (row-major dZcat) has no block-backward kernel any more");
  const bool fuse_dx = env.bwd6 && !bias && m.cond_fg == nullptr;      // (conditioning gradients need dF|dG in memory too)     // block_bwd6 (bias gradients need dF|dG in memory: those models keep block_bwd3 + the dx GEMM)
  WN_CHECK_CUDA(cudaMemsetAsync(G, 0, (size_t)m.n_params * sizeof(float), s));
  if (!fuse_dx) {
    // dx ping-pong buffers: layer i writes tiles >= its own first tile only, so rows below hold the previous step's values of
    // shallower layers (read by the next layer's tile loads) unless cleared.  (block_bwd6 writes every row that is read.)
    WN_CHECK_CUDA(cudaMemsetAsync(Wp + wl.DXa, 0, (size_t)B * L * 64 * 2, s));
    WN_CHECK_CUDA(cudaMemsetAsync(Wp + wl.DXb, 0, (size_t)B * L * 64 * 2, s));
    WN_CHECK_CUDA(cudaMemsetAsync(Wp + wl.DFG, 0, (size_t)B * L * 128 * 2, s));
    if (m.cond_fg_grad) WN_CHECK_CUDA(cudaMemsetAsync(Wp + wl.DFG2, 0, (size_t)B * L * 128 * 2, s));
  }
  {
    dim3 grid((unsigned)ceil_div(Wpad, 32), (unsigned)B);
    WN_PROF("dlogits_transpose", s);
    dlogits_transpose_kernel<<<grid, 256, 0, s>>>(d_dlogits, reinterpret_cast<__nv_bfloat16*>(Wp + wl.DLG), m.Q, W, Wpad, pad);
    WN_CHECK_LAUNCH();
  }
  WN_DEBUG_SYNC("dlogits_transpose", s);
  const int skip_tiles = (int)ceil_div(Wpad, 128);
  // ---- head: dH1 = mask_{h1}(dLg P2), dSK = mask_{h0}(dH1 P1)            (model.py:135-138 backwards)
  {
    GemmNtMaps gm{};
    gm.a[0] = M.dlg; gm.a[1] = M.dlg; gm.b[0] = M.p2T; gm.b[1] = M.p2T; gm.out = M.dh1;
    GemmNtParams gp{};
    const int S = m.S;
    gp.n_batches = B; gp.tile0 = 0; gp.tiles_per_batch = skip_tiles; gp.n_ntiles = S / 256; gp.n_total = S;
    gp.nk[0] = 4; gp.nk[1] = 0;                                  // K = Q = 256
    gp.epi = EPI_MASK; gp.aux = reinterpret_cast<const __nv_bfloat16*>(Wp + wl.H1);
    gp.aux_bstride = (int64_t)Wpad * S; gp.aux_rstride = S; gp.aux_col0 = 0;
    gp.row_lo = 0; gp.row_hi = Wpad; gp.tag = "gemm_nt_dH1";
    WN_PROPAGATE(launch_gemm_nt(256, gm, gp, s));
    WN_DEBUG_SYNC("gemm_nt dH1", s);
    if (m.cond_head_grad)      // d loss / d cond_N[frame] = sum of dH1 (gradient at the head's pre-activation) over the frame's time steps
      for (int c0 = 0; c0 < S; c0 += 128)
        WN_PROPAGATE(launch_frame_sum_bf16(Wp + wl.DH1, Wpad, S, c0, B, pad, W, m.cond_frames, m.cond_head_grad, S, 0, s));
    gm.a[0] = M.dh1; gm.a[1] = M.dh1; gm.b[0] = M.p1T; gm.b[1] = M.p1T; gm.out = M.dsk;
    gp.nk[0] = S / 64;                                           // K = S (S = 512: the streaming kernel)
    gp.aux = reinterpret_cast<const __nv_bfloat16*>(Wp + wl.H0); gp.tag = "gemm_nt_dSK";
    WN_PROPAGATE(launch_gemm_nt(256, gm, gp, s));
    WN_DEBUG_SYNC("gemm_nt dSK", s);
  }
  if (bias) {   // bias gradients = column sums of the matching output gradients (pad rows are zero)
    WN_PROPAGATE(launch_colsum_bf16(Wp + wl.DLG, 256, B, Wpad, 0, Wpad, G + m.post2.b, s));
    for (int c0 = 0; c0 < m.S; c0 += 256) {      // 256 columns per launch of the (B, Wpad, S) tensors
      WN_PROPAGATE(launch_colsum_bf16(Wp + wl.DH1 + (size_t)c0 * 2, 256, B, Wpad, 0, Wpad, G + m.post1.b + c0, s, 64, m.S));
      WN_PROPAGATE(launch_colsum_bf16(Wp + wl.DSK + (size_t)c0 * 2, 256, B, Wpad, 0, Wpad, G + m.layers[0].skip.b + c0, s, 64, m.S));
    }
    if (N > 1) {   // offsets of the skip biases were uploaded behind the pack-job table by fast_pack
      replicate_kernel<<<N - 1, m.S, 0, s>>>(G, reinterpret_cast<const int64_t*>(P + pl.jobs + skip_bias_offs_pos(m)), N, m.S);
      WN_CHECK_LAUNCH();
    }
    WN_DEBUG_SYNC("head bias grads", s);
  }
  // ---- head weight gradients: dP2 = dLg^T relu(h1), dP1 = dH1^T relu(h0); both m-tiles of each in one launch
  for (int which = 0; which < 2; ++which) {
    GemmTnMaps tm{};
    tm.a = which == 0 ? M.dlg : M.dh1;
    tm.b[0] = which == 0 ? M.h1 : M.h0; tm.b[1] = tm.b[0];
    GemmTnParams tp{};
    tp.n_batches = B; tp.tile0 = 0; tp.tiles_per_batch = skip_tiles; tp.m_valid = 128;
    float* base = G + (which == 0 ? m.post2.w : m.post1.w);
    tp.out0 = base; tp.out1 = base + 64 * m.S; tp.s_m = m.S; tp.s_n = 1;      // (Q, S) / (S, S): rows = A columns, row pitch S
    tp.n_mtiles = (which == 0 ? m.Q : m.S) / 128;
    tp.y_layers = m.S / 64; tp.y_off0 = 0; tp.y_stride = 64;                    // 64-column blocks of the S-wide B operand
    tp.tag = "gemm_tn_head";
    WN_PROPAGATE(launch_gemm_tn(4, tm, tp, s));
    WN_DEBUG_SYNC("gemm_tn head", s);
  }
  // ---- dZcat = dSK Wskip_cat (every layer's skip data-gradient at once; dSK is read once)
  {
    GemmNtMaps gm{};
    gm.a[0] = M.dsk; gm.a[1] = M.dsk; gm.b[0] = M.wsTcat; gm.b[1] = M.wsTcat; gm.out = M.dzcat;
    GemmNtParams gp{};
    if (dz_tiled) { gp.out_tiled = reinterpret_cast<__nv_bfloat16*>(Wp + wl.DZcat); gp.out_nblk = (int)ceil_div(Wpad, 32); }
    gp.n_batches = B; gp.tile0 = 0; gp.tiles_per_batch = skip_tiles; gp.n_ntiles = (int)ceil_div(64 * N, 256); gp.n_total = 64 * N;
    gp.nk[0] = m.S / 64; gp.nk[1] = 0; gp.epi = EPI_PLAIN; gp.row_lo = 0; gp.row_hi = Wpad; gp.tag = "gemm_nt_dZcat";
    WN_PROPAGATE(launch_gemm_nt(256, gm, gp, s));
    WN_DEBUG_SYNC("gemm_nt dZcat", s);
  }
  // ---- skip weight gradients for all layers in ONE launch: dWs_i = dSK^T Zcat[:, 64 i : 64 i + 64]
  //      (grid.y = 2 m-tiles x ceil(N/4) layer groups run concurrently, so dSK and Zcat tiles are shared through L2)
  {
    GemmTnMaps tm{};
    tm.a = M.dsk; tm.b[0] = M.zcat; tm.b[1] = M.zcat;
    GemmTnParams tp{};
    tp.n_batches = B; tp.tile0 = 0; tp.tiles_per_batch = skip_tiles; tp.m_valid = 128;
    tp.out0 = G; tp.out1 = G + 64 * m.D; tp.s_m = m.D; tp.s_n = 1; tp.n_valid = m.D;      // skip weight (S, D, 1)
    tp.n_mtiles = m.S / 128;
    tp.y_layers = N; tp.y_off0 = m.layers[0].skip.w;
    tp.y_stride = N > 1 ? m.layers[1].skip.w - m.layers[0].skip.w : 0;
    tp.tag = "gemm_tn_dWs";
    WN_PROPAGATE(launch_gemm_tn(4, tm, tp, s));
    WN_DEBUG_SYNC("gemm_tn dWs", s);
  }
  // ---- residual blocks, last to first
  // side stream for the per-layer reductions of the weight-gradient partial tiles (WN_WGRAD_SIDE=0: one reduction at the end)
  SideStream* side = nullptr;
  if (env.wgrad_side == 1 || (env.wgrad_side == -1 && m.cond_fg_grad != nullptr)) WN_PROPAGATE(side_stream(&side));
  WgradReduceArgs ra{};
  ra.filt0 = m.layers[0].filt.w; ra.gate0 = m.layers[0].gate.w; ra.dense0 = m.layers[0].dense.w;
  ra.layer_stride = N > 1 ? m.layers[1].filt.w - m.layers[0].filt.w : 0;
  ra.n_layers = N; ra.R = m.R; ra.D = m.D;
  const int tiles_total = (int)ceil_div(L, 128);
  // first tile of layer i: its own first valid one; block_bwd6 starts at the previous layer's, so that every row layer i - 1
  // reads has been written in this step (rows below s_out as zeros)
  auto first_tile = [&](int i) { return fuse_dx ? (i > 0 ? m.layers[i - 1].start / 128 : 0) : m.layers[i].start / 128; };
  for (int i = 0; i < N; ++i) ra.n_ctas[i] = std::min(B * (tiles_total - first_tile(i)), g_sm_count);
  const bool split_ok = fuse_dx && side != nullptr && m.split_event != nullptr && m.split_layer >= 0;
  bool split_done = false;
  for (int i = N - 1; i >= 0; --i) {
    const LayerP& l = m.layers[i];
    const int d = l.dilation, s_out = l.start, s_in = s_out - d;
    const bool has_dense = i + 1 < N;
    const CUtensorMap& dx_next = ((i + 1) & 1) ? M.dxb : M.dxa;       // dx_{i+1}  (block_bwd6: A_{i+1})
    const CUtensorMap& dx_cur = (i & 1) ? M.dxb : M.dxa;              // dx_i      (block_bwd6: A_i)
    const size_t dx_next_off = ((i + 1) & 1) ? wl.DXb : wl.DXa;
    const int tile0 = first_tile(i), tpb = tiles_total - tile0;
    const bool fs_side = side != nullptr && m.cond_fg_grad != nullptr;
    const int dfg_alt = fs_side ? (i & 1) : 0;
    const size_t dfg_off = dfg_alt ? wl.DFG2 : wl.DFG;
    if (fuse_dx) {
      BlockBwdMaps bm{};
      bm.x = M.layer[i].x; bm.w0 = M.layer[i].w0; bm.w1 = M.layer[i].w1; bm.wdT = M.layer[i].wdT;
      bm.a_in = dx_next; bm.q_in = ((i + 1) & 1) ? M.dqb : M.dqa;
      bm.a_out = dx_cur; bm.q_out = (i & 1) ? M.dqb : M.dqa;
      BlockBwd2Params b2{};
      BlockBwdParams& bp = b2.b;
      bp.L = L; bp.d = d; bp.s_out = s_out; bp.tile0 = tile0; bp.tiles_per_batch = tpb; bp.has_dense = has_dense;
      bp.tw0 = L - W; bp.tw_al = tw_al; bp.Wp = Wpad;
      bp.dzs = reinterpret_cast<const __nv_bfloat16*>(Wp + wl.DZcat); bp.dzs_pitch = 64 * N; bp.dzs_col = 64 * i;
      bp.dzs_lb0 = i * B; bp.dzs_nblk = (int)ceil_div(Wpad, 32);
      static const bool ts_env = getenv("WN_TS") != nullptr;
      bp.trace = (ts_env && i == N / 2) ? 1 : 0;
      bp.d_next = has_dense ? m.layers[i + 1].dilation : 0;
      bp.own_row0 = (s_out / 128) * 128;
      if (l2_hints_on()) { bp.pol_first = kL2EvictFirst; bp.pol_last = kL2EvictLast; }
      b2.n_batches = B;
      b2.partial = reinterpret_cast<float*>(Wp + wl.WGP) + (int64_t)i * WGP_LAYER_FLOATS;
      b2.w_order = env.bwd6_worder;
      b2.a_out_p = reinterpret_cast<__nv_bfloat16*>(Wp + ((i & 1) ? wl.DXb : wl.DXa));
      b2.q_out_p = reinterpret_cast<__nv_bfloat16*>(Wp + wl.DFG + ((i & 1) ? align_up((size_t)B * L * 64 * 2, 1024) : 0));
      WN_PROPAGATE(launch_block_bwd6(bm, b2, s));
      WN_DEBUG_SYNC("block_bwd6", s);
      if (side) {
        WN_CHECK_CUDA(cudaEventRecord(side->fork, s));
        WN_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        ra.layer0 = i;
        WN_PROF("wgrad_reduce (side stream, overlapped)", side->stream);
        wgrad_reduce_kernel<<<dim3((128 * 192) / 32, 1), 256, 0, side->stream>>>(reinterpret_cast<const float*>(Wp + wl.WGP), WGP_LAYER_FLOATS,
                                                                                 ra, G);
        WN_CHECK_LAUNCH();
        if (split_ok && i == m.split_layer) {      // blocks >= i, the skip weights and the head are final: that bucket may be exchanged
          WN_CHECK_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(m.split_event), side->stream));
          split_done = true;
        }
      }
      continue;
    }
    {
      BlockBwdMaps bm{};
      bm.x = M.layer[i].x; bm.w0 = M.layer[i].w0; bm.w1 = M.layer[i].w1; bm.dx = dx_next; bm.wdT = M.layer[i].wdT;
      bm.dfg = dfg_alt ? M.dfg2 : M.dfg;
      BlockBwd2Params b2{};
      BlockBwdParams& bp = b2.b;
      bp.L = L; bp.d = d; bp.s_out = s_out; bp.tile0 = tile0; bp.tiles_per_batch = tpb; bp.has_dense = has_dense;
      bp.tw0 = L - W; bp.tw_al = tw_al; bp.Wp = Wpad;
      bp.dzs = reinterpret_cast<const __nv_bfloat16*>(Wp + wl.DZcat); bp.dzs_pitch = 64 * N; bp.dzs_col = 64 * i;
      bp.dzs_lb0 = i * B; bp.dzs_nblk = (int)ceil_div(Wpad, 32);
      bp.bias_fg = bias ? reinterpret_cast<const float*>(P + pl.bias_fg) + i * 128 : nullptr;
      if (m.cond_fg) {      // conditioned decoder: recompute with the same table (the conv bias is part of it)
        WN_REQUIRE(m.cond_fg16, WN_ERR_INVALID, "conditioned backward: the bf16 conditioning table is missing (launch_cond_pack16)");
        bp.cond = m.cond_fg; bp.cond16 = reinterpret_cast<const uint4*>(m.cond_fg16); bp.cond_frames = m.cond_frames; bp.cond_layers = N; bp.cond_layer = i;
        bp.bias_fg = nullptr;
      }
      if (l2_hints_on()) { bp.pol_first = kL2EvictFirst; bp.pol_last = kL2EvictLast; }
      b2.n_batches = B;
      b2.partial = reinterpret_cast<float*>(Wp + wl.WGP) + (int64_t)i * WGP_LAYER_FLOATS;
      if (fs_side && i + 2 < N) WN_CHECK_CUDA(cudaStreamWaitEvent(s, side->dfg_free[dfg_alt], 0));      // layer i + 2's frame sums have read this buffer
      WN_PROPAGATE(launch_block_bwd2(bm, b2, s));
      WN_DEBUG_SYNC("block_bwd3", s);
      if (m.cond_fg_grad && !fs_side)   // d loss / d cond_i[frame] in the autoencoder's raw (gate | filter) order: (B, frames, N, 2 D)
        WN_PROPAGATE(launch_frame_sum_bf16(Wp + dfg_off, L, 128, 0, B, s_out, L - s_out, m.cond_frames,
                                           m.cond_fg_grad + (int64_t)i * 2 * m.D, N * 2 * m.D, m.D, s));
      if (side) {   // sum this layer's per-CTA weight-gradient tiles next to the kernels that follow (its CTAs need ~1 KB of smem)
        WN_CHECK_CUDA(cudaEventRecord(side->fork, s));
        WN_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        if (fs_side) {      // the frame sums too: dF|dG alternates between two buffers, so the next layer's kernels do not wait for them
          WN_PROPAGATE(launch_frame_sum_bf16(Wp + dfg_off, L, 128, 0, B, s_out, L - s_out, m.cond_frames,
                                             m.cond_fg_grad + (int64_t)i * 2 * m.D, N * 2 * m.D, m.D, side->stream));
          WN_CHECK_CUDA(cudaEventRecord(side->dfg_free[dfg_alt], side->stream));
        }
        ra.layer0 = i;
        WN_PROF("wgrad_reduce (side stream, overlapped)", side->stream);
        wgrad_reduce_kernel<<<dim3((128 * 192) / 32, 1), 256, 0, side->stream>>>(reinterpret_cast<const float*>(Wp + wl.WGP), WGP_LAYER_FLOATS,
                                                                                 ra, G);
        WN_CHECK_LAUNCH();
      }
    }
    if (bias) {   // dFG columns [0,64) = filter, [64,128) = gate: two separate bias vectors in the flat layout
      WN_PROPAGATE(launch_colsum_bf16_split(Wp + dfg_off, B, L, s_out, G + l.filt.b, G + l.gate.b, s, m.D));
      if (has_dense) WN_PROPAGATE(launch_colsum_bf16(Wp + dx_next_off, 64, B, L, s_out, L, G + l.dense.b, s, m.R));
    }
    {   // dx_i[tau] = dx_{i+1}[tau] + W1^T dFG[tau] + W0^T dFG[tau + d]
      GemmNtMaps gm{};
      gm.a[0] = dfg_alt ? M.dfg2 : M.dfg; gm.a[1] = gm.a[0]; gm.b[0] = M.layer[i].wfgT1; gm.b[1] = M.layer[i].wfgT0; gm.out = dx_cur;
      GemmNtParams gp{};
      const int t0 = s_in / 128;
      gp.n_batches = B; gp.tile0 = t0; gp.tiles_per_batch = tiles_total - t0; gp.n_ntiles = 1;
      gp.nk[0] = 2; gp.nk[1] = 2; gp.a_row_off[0] = 0; gp.a_row_off[1] = d;
      gp.epi = EPI_ADD; gp.aux = reinterpret_cast<const __nv_bfloat16*>(Wp + dx_next_off);
      gp.aux_bstride = (int64_t)L * 64; gp.aux_rstride = 64; gp.aux_col0 = 0;
      gp.row_lo = s_in; gp.row_hi = L; gp.tag = "gemm_nt_dx";
      // (L2 hints on this GEMM's dFG loads / dx stores were measured to cost 0.1 ms per step: left off)
      if (env.l2hint_dx) { gp.pol_a[0] = kL2EvictFirst; gp.pol_out = kL2EvictLast; }
      WN_PROPAGATE(launch_gemm_nt(64, gm, gp, s));
      WN_DEBUG_SYNC("gemm_nt dx", s);
    }
  }
  if (m.split_event && m.split_layer >= 0 && !split_done) {      // configurations without the early record: the event still fires
    if (side) {
      WN_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
      WN_CHECK_CUDA(cudaStreamWaitEvent(s, side->join, 0));
    }
    WN_CHECK_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(m.split_event), s));
  }
  if (side) {                 // join: everything after this point on `s` sees the reduced weight gradients
    WN_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
    WN_CHECK_CUDA(cudaStreamWaitEvent(s, side->join, 0));
  } else {                    // one reduction of every layer's per-CTA weight-gradient tiles (fixed order: deterministic)
    ra.layer0 = 0;
    WN_PROF("wgrad_reduce", s);
    dim3 grid((128 * 192) / 32, (unsigned)N);
    wgrad_reduce_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(Wp + wl.WGP), WGP_LAYER_FLOATS, ra, G);
    WN_CHECK_LAUNCH();
  }
  // ---- causal layer
  if (fuse_dx) {      // dx_0[tau] = A_0[tau] + Q_0[tau + d_0]  ->  the buffer behind M.zf
    dim3 grid((unsigned)std::min<int64_t>(ceil_div((int64_t)L * 8, 256), 2048), (unsigned)B);
    WN_PROF("combine_dx0", s);
    combine_dx0_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(Wp + wl.DXa), reinterpret_cast<const __nv_bfloat16*>(Wp + wl.DFG),
                                            reinterpret_cast<__nv_bfloat16*>(Wp + wl.Zf), L, m.layers[0].dilation);
    WN_CHECK_LAUNCH();
  }
  const __nv_bfloat16* dx0 = reinterpret_cast<const __nv_bfloat16*>(Wp + (fuse_dx ? wl.Zf : wl.DXa));
  if (bias) WN_PROPAGATE(launch_colsum_bf16(dx0, 64, B, L, 1, L, G + m.causal.b, s, m.R));
  if (d_idx && m.Q == 256 && !env.scatter_simt) {
    const int smem = CwCfg::TOTAL + 1024;
    WN_PROPAGATE(set_smem_once(causal_wgrad_kernel, smem));
    const int tiles = (int)ceil_div(L, 128);
    WN_PROF("causal_wgrad", s);
    WN_CHECK_CUDA(launch_pdl(causal_wgrad_kernel, dim3((unsigned)std::min(B * tiles, g_sm_count)), dim3(192), smem, s, fuse_dx ? M.zf : M.dxa, d_idx,
                             G + m.causal.w, L, B, tiles, m.R));
    WN_DEBUG_SYNC("causal_wgrad", s);
  } else if (d_idx) {
    const int smem = 2 * m.Q * 64 * 4;
    WN_PROPAGATE(set_smem_once(causal_scatter_bwd_kernel, smem));
    const int per_batch = std::max(1, (g_sm_count + B - 1) / B);
    const int rows_per_cta = (int)ceil_div(L - 1, per_batch);
    dim3 grid((unsigned)ceil_div(L - 1, rows_per_cta), (unsigned)B);
    {
      WN_PROF("causal_scatter_bwd", s);
      causal_scatter_bwd_kernel<<<grid, 1024, smem, s>>>(d_idx, dx0, G + m.causal.w, L, m.Q, rows_per_cta);
      WN_CHECK_LAUNCH();
    }
    WN_DEBUG_SYNC("causal_scatter", s);
  } else {
    float* dx0f = reinterpret_cast<float*>(Wp + wl.DX0f);
    const int64_t n = (int64_t)B * L * 64;
    bf16_to_f32_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 4096), 256, 0, s>>>(dx0, dx0f, n);
    WN_CHECK_LAUNCH();
    for (int tap = 0; tap < 2; ++tap) {      // dense (B,Q,L) input: fp32 SIMT weight gradient of the causal conv
      WgArgs g;
      g.X.p = d_x; g.X.sb = (int64_t)m.Q * L; g.X.st = 1; g.X.sc = L;
      g.x_lo = 0; g.x_hi = L; g.n_in = m.Q; g.off = tap == 0 ? -1 : 0;
      g.dY.p = dx0f; g.dY.sb = (int64_t)L * 64; g.dY.st = 64; g.dY.sc = 1;
      g.n_out = m.R; g.dW = G + m.causal.w + tap; g.s_out = (int64_t)m.Q * 2; g.s_in = 2;
      g.B = B; g.t0 = 1; g.t1 = L;
      WN_PROPAGATE(launch_wgrad(g, s));
    }
  }
  return WN_OK;
}

}  // namespace wn
