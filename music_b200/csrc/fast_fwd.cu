// fast_fwd.cu - bf16 tensor-core forward of the WaveNet stack (R = D = 64, S = Q = 256).
//
// block_fwd2_kernel : one residual block (wavenet/model.py:118-129), persistent CTAs over [128 time x 64 ch] tiles:
//     TMA: x_i[tau-d], x_i[tau] tiles + W_fg (both taps) + W_dense          -> shared memory (SW128)
//     UMMA #1 (M128 N128 K128): [f|g] = x[tau-d] W0^T + x[tau] W1^T         -> TMEM cols [0,128)
//     epilogue #1: z = sigmoid(g) * tanh(f) -> bf16 -> shared (A operand of #2) and TMA-stored into
//                  the skip operand matrix Zcat[:, 64 i : 64 i + 64] for rows in the last W
//     UMMA #2 (M128 N64 K64): dense = z Wd^T                                -> TMEM cols [128,192)
//     epilogue #2: x_{i+1} = dense + x_i[tau] (residual read back from the shared x tile) -> TMA store.
//                  The residual stream is carried as hi + lo bf16 pairs (hi feeds the MMAs, hi+lo the
//                  residual add) so that 30 layers of bf16 rounding do not accumulate in x.
// skip_head_kernel  : for a [128 rows of B*W] tile, chained GEMMs with TMEM-resident accumulators
//     acc1 = Zcat[128 x 64N] Wskip_cat^T  (= sum_i skip_i, model.py:127-134; K = 64 N_layers)
//     h0 = relu(acc1) -> smem/HBM ; acc2 = h0 P1^T ; h1 = relu(acc2) -> smem/HBM ; acc1 = h1 P2^T
//     logits (fp32) -> (B,Q,W).   Warp-specialised: warps 0-3 epilogue, warp 4 TMA, warp 5 MMA.
#include "fast_kernels.cuh"
#include "tc05.cuh"

namespace wn {
using namespace tc;

// ======================================================================================= block
namespace {
constexpr uint32_t TILE_BYTES = 128 * 128;     // [128 rows][64 bf16]
constexpr uint32_t WD_BYTES = 64 * 128;
}  // namespace

// ====================================================================================== block (persistent)
// 18 warps: 0-15 epilogue (4 per TMEM lane quarter, 16 columns each), 16 TMA producer, 17 MMA issuer.
// TMEM: two accumulator buffers {f|g: 128 cols, dense: 64 cols}; UMMA #1 of tile n+1/n+2 overlaps the epilogues of tile n.
// Shared memory: resident weights, THREE input stages {x tap0, x tap1}, two z tiles (A operand of UMMA #2 and source of
// the Zcat store), one staging tile for x_{i+1} hi and two lo tiles (TMA-loaded, lo' written in place).  The TMA stores of
// a tile are checked one epilogue phase later, just before their staging tiles are rewritten, so nobody waits for them;
// the MMA issuer polls its two job queues (UMMA #2 of tile j2, UMMA #1 of tile j1) and never blocks on one while the
// other is ready.  Round 2 (profiles/r2_summary.md): a store warp, epilogue 2 software-pipelined behind the next tile's epilogue 1 and
// outputs written straight from registers were all built and verified - 1.34 to 1.61 ms per step against 1.32 for this kernel:
// at 7470 warp-instructions per tile (1870 issue slots per scheduler, issue active 45 %) the tile period is set by the epilogue's
// instruction stream, not by any wait those variants remove.  History (clock64 instrumentation, profiles/): the first
// persistent version spent, per 6800-cycle tile, 1250 cycles with all threads waiting for the TMA store to drain and
// 2000 waiting for UMMA #2 queued behind the next tile's UMMA #1 in the in-order tensor pipe.
namespace {

struct Fwd2Smem {
  static constexpr uint32_t W0 = 0, W1 = TILE_BYTES, WD = 2 * TILE_BYTES;
  static constexpr uint32_t IN = 2 * TILE_BYTES + WD_BYTES, IN_STAGE = 2 * TILE_BYTES;     // {x tap0, x tap1}
  static constexpr int STAGES = 3;
  static constexpr uint32_t Z = IN + STAGES * IN_STAGE;                                    // 2 z tiles
  static constexpr uint32_t XO = Z + 2 * TILE_BYTES;                                       // staging of x_{i+1} hi
  static constexpr uint32_t LO = XO + TILE_BYTES;                                          // 2 lo tiles: TMA-loaded, lo' written in place
  static constexpr uint32_t TOTAL = LO + 2 * TILE_BYTES;                                   // 216 KB
};
__device__ __forceinline__ void epi16_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// 32 bytes per thread in one request (256-bit LDG through the read-only path)
__device__ __forceinline__ void ldg_nc32(const void* ptr, uint32_t* v) {
  asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(ptr));
}

// BIAS and TRACE are compile-time: run-time tests inside the unrolled epilogue loops cost a branch per element
template <bool BIAS, bool TRACE, bool COND>
__global__ void __launch_bounds__(576, 1)
block_fwd2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w0,
                  const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_wd,
                  const __grid_constant__ CUtensorMap tm_xo, const __grid_constant__ CUtensorMap tm_loo,
                  const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_lo, BlockFwdParams p,
                  BlockFwdPtrs g, int n_batches) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t w_full, in_full[3], in_empty[3], fg_full[2], dense_full[2], acc_empty[2], z_ready;
  __shared__ __align__(8) uint64_t lo_full[2], lo_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 2);       // UMMA #1 commit + epilogue (residual tile read)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&fg_full[i], 1);
      mbar_init(&dense_full[i], 1);
      mbar_init(&acc_empty[i], 1);
      mbar_init(&lo_full[i], 1);
      mbar_init(&lo_empty[i], 1);
    }
    mbar_init(&z_ready, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sbase = smem_u32(sm);
  const int n_items = n_batches * p.tiles_per_batch;
  const bool dense = p.has_dense != 0;
  const int n_mine = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  pdl_launch_dependents();      // the next launch may set itself up on SMs this grid has left ...
  pdl_wait();                   // ... and this one touches activations only after its predecessor has completed

  if (warp == 16) {
    if (lane == 0 && n_mine > 0) {
      mbar_expect_tx(&w_full, 2 * TILE_BYTES + (dense ? WD_BYTES : 0));
      tma_load_2d(sm + Fwd2Smem::W0, &tm_w0, &w_full, 0, 0);
      tma_load_2d(sm + Fwd2Smem::W1, &tm_w1, &w_full, 0, 0);
      if (dense) tma_load_2d(sm + Fwd2Smem::WD, &tm_wd, &w_full, 0, 0);
      for (int it = 0; it < n_mine; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int st = it % 3;
        const int b = item / p.tiles_per_batch, tau0 = (p.tile0 + item % p.tiles_per_batch) * 128;
        mbar_wait(&in_empty[st], ((it / 3) & 1) ^ 1);
        uint8_t* si = sm + Fwd2Smem::IN + st * Fwd2Smem::IN_STAGE;
        mbar_expect_tx(&in_full[st], 2 * TILE_BYTES);
        // rows of x_i are read as tap 1 by their own tile and, d rows later, as tap 0: that second read is the last one
        tma_load_3d(si, &tm_x, &in_full[st], 0, tau0 - p.d, b, p.pol_first);
        tma_load_3d(si + TILE_BYTES, &tm_x, &in_full[st], 0, tau0, b);
        if (TRACE && p.ts && blockIdx.x == 0) p.ts[1024 + it * 4 + 2] = clock64();
        if (dense && !(TRACE && (p.dbg & 4))) {      // low half of the residual stream; its tile is reused for lo' and released after that store
          mbar_wait(&lo_empty[it & 1], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&lo_full[it & 1], TILE_BYTES);
          tma_load_3d(sm + Fwd2Smem::LO + (it & 1) * TILE_BYTES, &tm_lo, &lo_full[it & 1], 0, tau0, b, p.pol_first);
        }
      }
    }
  } else if (warp == 17) {
    if (lane == 0 && n_mine > 0) {
      constexpr uint32_t id1 = idesc_bf16(128, 128, 0, 0), id2 = idesc_bf16(128, 64, 0, 0);
      mbar_wait(&w_full, 0);
      int j1 = 0, j2 = 0;          // next UMMA #1 / UMMA #2 tile
      while (j1 < n_mine || (dense && j2 < n_mine)) {
        bool progressed = false;
        if (dense && j2 < j1 && mbar_test_wait(&z_ready, j2 & 1)) {      // UMMA #2 first: never queue it behind a UMMA #1
          tc_fence_after();
          const uint32_t acc = tmem + (j2 & 1) * 192 + 128;
          const uint32_t zt = sbase + Fwd2Smem::Z + (j2 & 1) * TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(acc, desc_kmajor(zt, kk), desc_kmajor(sbase + Fwd2Smem::WD, kk), id2, kk > 0);
          umma_commit(&dense_full[j2 & 1]);
          if (TRACE && p.ts && blockIdx.x == 0) p.ts[1024 + j2 * 4 + 1] = clock64();
          ++j2;
          progressed = true;
        }
        if (j1 < n_mine && mbar_test_wait(&in_full[j1 % 3], (j1 / 3) & 1) && mbar_test_wait(&acc_empty[j1 & 1], ((j1 >> 1) & 1) ^ 1)) {
          tc_fence_after();
          const uint32_t si = sbase + Fwd2Smem::IN + (j1 % 3) * Fwd2Smem::IN_STAGE, acc = tmem + (j1 & 1) * 192;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(acc, desc_kmajor(si, kk), desc_kmajor(sbase + Fwd2Smem::W0, kk), id1, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(acc, desc_kmajor(si + TILE_BYTES, kk), desc_kmajor(sbase + Fwd2Smem::W1, kk), id1, true);
          umma_commit(&fg_full[j1 & 1]);
          umma_commit(&in_empty[j1 % 3]);
          if (TRACE && p.ts && blockIdx.x == 0) p.ts[1024 + j1 * 4 + 0] = clock64();
          ++j1;
          progressed = true;
        }
        (void)progressed;
      }
    }
  } else {
    const int q4 = warp & 3, cg = warp >> 2;
    const int row = q4 * 32 + lane;
    // (batch row, tile) of this CTA's next work item, stepped without a division per tile
    int b = (int)blockIdx.x / p.tiles_per_batch, tl = (int)blockIdx.x % p.tiles_per_batch, st = 0;
    for (int it = 0; it < n_mine; ++it) {
      const int ab = it & 1;
      const uint32_t ph2 = (it >> 1) & 1;
      const int tau0 = (p.tile0 + tl) * 128;
      const int tau = tau0 + row;
      const bool valid = (tau >= p.s_out) && (tau < p.L);
      const uint32_t lane_addr = tmem_addr(tmem, q4 * 32, ab * 192);
      const uint8_t* si = sm + Fwd2Smem::IN + st * Fwd2Smem::IN_STAGE;
      uint8_t* zt = sm + Fwd2Smem::Z + ab * TILE_BYTES;
      const bool rec = TRACE && p.ts != nullptr && blockIdx.x == 0 && tid == 0;
      long long* ts = p.ts + (int64_t)it * 8;
      if (rec) ts[0] = clock64();
      uint8_t* lot = sm + Fwd2Smem::LO + ab * TILE_BYTES;
      // ---- epilogue 1: gate -> z tile (smem, A operand of UMMA #2) and Zcat (global).  z tile `ab` was last read by
      //      UMMA #2 of tile it-2, whose completion (dense_full) every thread waited for in that tile's epilogue 2.
      // this row's conditioning values (autoencoder decoder; rows outside the valid range are masked anyway): 16 filter + 16 gate
      // fp32 values as four 256-bit loads issued BEFORE the wait for the accumulator; the table is ordered so that a warp's 32 rows
      // read 32 consecutive sectors per load (cond_pack16_kernel).  (The row-major
      // table read with scalar loads - 64 requests of 32 different lines each - cost ~8 us per tile, with 16-byte loads ~2.  A
      // bf16 table halves the requests again but adds its rounding to the pre-activations: gradient noise 0.14 -> 0.151.)
      uint32_t cw[32];
      if (COND) {
        const int fr = valid ? cond_frame(tau - p.s_out, p.L - p.s_out, p.cond_frames) : 0;
        const uint4* cp16 = p.cond16 + (((((int64_t)b * p.cond_layers + p.cond_layer) * 4 + cg) * 4) * p.cond_frames + fr) * 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) ldg_nc32(cp16 + (int64_t)q * p.cond_frames * 2, cw + 8 * q);
      }
      mbar_wait(&fg_full[ab], ph2);
      tc_fence_after();
      if (rec) ts[1] = clock64();
      uint32_t f[16], gq[16];
      tmem_ld16(lane_addr + cg * 16, f);
      tmem_ld16(lane_addr + 64 + cg * 16, gq);
      tmem_ld_wait();
      uint32_t pz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f0 = __uint_as_float(f[2 * j]), f1 = __uint_as_float(f[2 * j + 1]);
        float g0 = __uint_as_float(gq[2 * j]), g1 = __uint_as_float(gq[2 * j + 1]);
        if (BIAS) {
          f0 += p.bias_fg[cg * 16 + 2 * j];
          f1 += p.bias_fg[cg * 16 + 2 * j + 1];
          g0 += p.bias_fg[64 + cg * 16 + 2 * j];
          g1 += p.bias_fg[64 + cg * 16 + 2 * j + 1];
        }
        if (COND) {
          f0 += __uint_as_float(cw[2 * j]);
          f1 += __uint_as_float(cw[2 * j + 1]);
          g0 += __uint_as_float(cw[16 + 2 * j]);
          g1 += __uint_as_float(cw[16 + 2 * j + 1]);
        }
        float z0 = sigmoid_fast(g0) * tanh_fast(f0), z1 = sigmoid_fast(g1) * tanh_fast(f1);
        if (TRACE && (p.dbg & 8)) { z0 = g0 * f0; z1 = g1 * f1; }
        pz[j] = valid ? pack_bf16(z0, z1) : 0u;
      }
      const uint4 zv0 = make_uint4(pz[0], pz[1], pz[2], pz[3]), zv1 = make_uint4(pz[4], pz[5], pz[6], pz[7]);
      sts128(smem_u32(zt) + sw128_chunk(row, cg * 2), zv0.x, zv0.y, zv0.z, zv0.w);
      sts128(smem_u32(zt) + sw128_chunk(row, cg * 2 + 1), zv1.x, zv1.y, zv1.z, zv1.w);
      if (rec) ts[2] = clock64();
      fence_proxy_async_smem();
      tc_fence_before();
      // all TMA stores committed so far (tiles <= it-1) have been read: the x_{i+1} staging tiles may be rewritten in
      // epilogue 2, and z tile 1-ab in the next tile.  They were issued a whole epilogue phase ago: normally no wait.
      if (tid == 0) {
        tma_store_wait_read();
        if (dense && it > 0) mbar_arrive(&lo_empty[ab ^ 1]);       // the previous tile's lo' store has been read
      }
      epi16_bar_sync();
      if (rec) ts[3] = clock64();
      if (tid == 0 && dense) mbar_arrive(&z_ready);
      // ---- epilogue 2: x_{i+1} = dense + (hi + lo) in fp32, split again into hi + lo
      if (dense) {
        if (!(TRACE && (p.dbg & 4))) mbar_wait(&lo_full[ab], ph2);
        mbar_wait(&dense_full[ab], ph2);
        tc_fence_after();
        if (rec) ts[4] = clock64();
        uint32_t dv[16];
        tmem_ld16(lane_addr + 128 + cg * 16, dv);
        tmem_ld_wait();
        uint32_t ph[8], pl[8];
        const uint4 lv0 = lds128(smem_u32(lot) + sw128_chunk(row, cg * 2));
        const uint4 lv1 = lds128(smem_u32(lot) + sw128_chunk(row, cg * 2 + 1));
        const uint32_t ll[8] = {lv0.x, lv0.y, lv0.z, lv0.w, lv1.x, lv1.y, lv1.z, lv1.w};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 rv = lds128(smem_u32(si) + TILE_BYTES + sw128_chunk(row, cg * 2 + q));
          const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            float x0 = __uint_as_float(dv[2 * j]) + (bf16lo(rr[e]) + bf16lo(ll[j]));
            float x1 = __uint_as_float(dv[2 * j + 1]) + (bf16hi(rr[e]) + bf16hi(ll[j]));
            if (BIAS) {
              x0 += p.bias_d[cg * 16 + 2 * j];
              x1 += p.bias_d[cg * 16 + 2 * j + 1];
            }
            const uint32_t h = pack_bf16(x0, x1);
            ph[j] = h;
            pl[j] = pack_bf16(x0 - bf16lo(h), x1 - bf16hi(h));
          }
        }
        if (__any_sync(0xffffffffu, !valid)) {      // rows below the layer's valid start are stored as zeros (first tiles only)
          if (!valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ph[j] = 0u; pl[j] = 0u; }
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t o = sw128_chunk(row, cg * 2 + q);
          sts128(sbase + Fwd2Smem::XO + o, ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
          sts128(smem_u32(lot) + o, pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);   // in place
        }
        fence_proxy_async_smem();
      }
      if (rec) ts[5] = clock64();
      tc_fence_before();
      epi16_bar_sync();              // every thread has finished reading this tile's TMEM buffer and residual tile
      if (rec) ts[6] = clock64();
      if (tid == 0) {
        mbar_arrive(&acc_empty[ab]);
        mbar_arrive(&in_empty[st]);
        // z is read again only by the skip GEMM at the end of the forward; x_{i+1} hi / lo' by the very next launch
        if (tau0 >= p.tw_al && !(TRACE && (p.dbg & 1))) tma_store_3d(&tm_z, zt, p.zcol, tau0 - p.tw_al, b, p.pol_first);
        if (dense && !(TRACE && (p.dbg & 2))) {
          tma_store_3d(&tm_xo, sm + Fwd2Smem::XO, 0, tau0, b, p.pol_last);
          tma_store_3d(&tm_loo, lot, 0, tau0, b, p.pol_last);
        }
        tma_store_commit();          // not waited for here: checked before the staging tiles are rewritten (above)
      }
      if (rec) ts[7] = clock64();
      st = st == 2 ? 0 : st + 1;
      tl += (int)gridDim.x;
      while (tl >= p.tiles_per_batch) { tl -= p.tiles_per_batch; ++b; }
    }
    if (tid == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace

int launch_block_fwd2(const BlockFwdMaps& m, const BlockFwdParams& p, const BlockFwdPtrs& g, int n_batches, cudaStream_t s) {
  const int smem = Fwd2Smem::TOTAL + 1024;
  const int n_items = n_batches * p.tiles_per_batch;
  if (n_items <= 0) return WN_OK;
  const bool trace = p.ts != nullptr || p.dbg != 0, bias = p.bias_fg != nullptr;
  auto k = trace ? (bias ? block_fwd2_kernel<true, true, false> : block_fwd2_kernel<false, true, false>)
                 : (bias ? block_fwd2_kernel<true, false, false> : block_fwd2_kernel<false, false, false>);
  static const void* configured[6] = {};
  int slot = (trace ? 2 : 0) + (bias ? 1 : 0);
  if (p.cond) {      // conditioned decoder of the autoencoder
    k = bias ? block_fwd2_kernel<true, false, true> : block_fwd2_kernel<false, false, true>;
    slot = 4 + (bias ? 1 : 0);
  }
  if (configured[slot] == nullptr) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[slot] = reinterpret_cast<const void*>(k);
  }
  WN_PROF("block_fwd", s);
  WN_CHECK_CUDA(launch_pdl(k, dim3((unsigned)std::min(n_items, g_sm_count)), dim3(576), smem, s, m.x, m.w0, m.w1, m.wd, m.xo, m.loo, m.z, m.lo, p, g,
                           n_batches));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

// ================================================================================== skip + head
namespace {

constexpr int SH_THREADS = 192;
constexpr int SH_STAGES = 3;
constexpr uint32_t SH_A_BYTES = TILE_BYTES;          // [128][64]
constexpr uint32_t SH_B_BYTES = 256 * 128;           // [256][64]
constexpr uint32_t SH_STAGE_BYTES = SH_A_BYTES + SH_B_BYTES;
constexpr uint32_t SH_H_OFF = SH_STAGES * SH_STAGE_BYTES;
constexpr uint32_t SH_TOTAL = SH_H_OFF + 4 * TILE_BYTES;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <bool BIAS>
__global__ void __launch_bounds__(SH_THREADS, 1)
skip_head_kernel(const __grid_constant__ CUtensorMap tm_zcat, const __grid_constant__ CUtensorMap tm_wsk,
                 const __grid_constant__ CUtensorMap tm_p1, const __grid_constant__ CUtensorMap tm_p2,
                 const __grid_constant__ CUtensorMap tm_h0, const __grid_constant__ CUtensorMap tm_h1, SkipHeadParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[SH_STAGES], empty[SH_STAGES];
  __shared__ __align__(8) uint64_t acc_full[3], h_ready[2], tmem_free;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < SH_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&acc_full[i], 1);
    mbar_init(&h_ready[0], 1);
    mbar_init(&h_ready[1], 1);
    mbar_init(&tmem_free, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sbase = smem_u32(sm);
  const int kc_skip = p.k_skip / 64;
  pdl_launch_dependents();      // programmatic dependent launch: see tc05.cuh
  pdl_wait();

  if (warp == 4) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int row0 = tile * 128;
        for (int kc = 0; kc < kc_skip + 8; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = sm + stage * SH_STAGE_BYTES;
          uint8_t* sb = sa + SH_A_BYTES;
          if (kc < kc_skip) {
            mbar_expect_tx(&full[stage], SH_STAGE_BYTES);
            tma_load_2d(sa, &tm_zcat, &full[stage], kc * 64, row0);
            tma_load_2d(sb, &tm_wsk, &full[stage], kc * 64, 0);
          } else {
            mbar_expect_tx(&full[stage], SH_B_BYTES);
            const int k2 = kc - kc_skip;
            tma_load_2d(sb, k2 < 4 ? &tm_p1 : &tm_p2, &full[stage], (k2 & 3) * 64, 0);
          }
          if (++stage == SH_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      constexpr uint32_t idn = idesc_bf16(128, 256, 0, 0);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t tph = it & 1;
        mbar_wait(&tmem_free, tph ^ 1);            // epilogue of the previous tile has drained acc1
        tc_fence_after();
        for (int kc = 0; kc < kc_skip; ++kc) {     // acc1 = Zcat * Wskip_cat^T
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = sbase + stage * SH_STAGE_BYTES, sb = sa + SH_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_kmajor(sa, k), desc_kmajor(sb, k), idn, (kc | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == SH_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[0]);
        for (int g = 0; g < 2; ++g) {              // acc2 = h0 P1^T ; acc1 = h1 P2^T
          mbar_wait(&h_ready[g], tph);
          tc_fence_after();
          const uint32_t dst = g == 0 ? tmem + 256 : tmem;
          for (int kc = 0; kc < 4; ++kc) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = sbase + SH_H_OFF + kc * TILE_BYTES;
            const uint32_t sb = sbase + stage * SH_STAGE_BYTES + SH_A_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(dst, desc_kmajor(sa, k), desc_kmajor(sb, k), idn, (kc | k) != 0);
            umma_commit(&empty[stage]);
            if (++stage == SH_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&acc_full[1 + g]);
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps 0-3
    uint32_t it = 0;
    const int row = tid;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t tph = it & 1;
      const int row0 = tile * 128;
      const uint32_t lane_addr = tmem_addr(tmem, warp * 32, 0);
      for (int g = 0; g < 2; ++g) {
        // relu(acc) -> bf16 H tile in shared memory (A operand of the next GEMM) and to HBM for backward
        mbar_wait(&acc_full[g], tph);
        tc_fence_after();
        const uint32_t src = lane_addr + (g == 0 ? 0 : 256);
        const float* bias = g == 0 ? p.bias_skip : p.bias_p1;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t v[32];
          tmem_ld32(src + c * 32, v);
          tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a0 = __uint_as_float(v[2 * j]), a1 = __uint_as_float(v[2 * j + 1]);
            if (BIAS) {
              a0 += bias[c * 32 + 2 * j];
              a1 += bias[c * 32 + 2 * j + 1];
            }
            packed[j] = pack_bf16(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
          }
          uint8_t* ht = sm + SH_H_OFF + (c >> 1) * TILE_BYTES;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 val = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            sts128(smem_u32(ht) + sw128_chunk(row, (c & 1) * 4 + q), val.x, val.y, val.z, val.w);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        epi_bar_sync();
        if (tid == 0) {
          mbar_arrive(&h_ready[g]);
          for (int kc = 0; kc < 4; ++kc)
            tma_store_2d(g == 0 ? &tm_h0 : &tm_h1, sm + SH_H_OFF + kc * TILE_BYTES, kc * 64, row0);
          tma_store_commit();
          tma_store_wait_read();       // H tile may be overwritten once the stores have read it ...
        }
        // ... and once the GEMM that consumes it has completed (acc_full[g+1], waited below / next loop)
        if (g == 0) {
          // nothing: epilogue g=1 waits acc_full[1], which implies the MMAs reading h0 are done
        }
        epi_bar_sync();
      }
      // logits
      mbar_wait(&acc_full[2], tph);
      tc_fence_after();
      const int r = row0 + row;
      const int bb = r / p.Wp, tw = r % p.Wp - p.pad;
      const bool ok = r < p.n_rows && tw >= 0;
      float* out = p.logits + ((int64_t)bb * p.Q) * p.W + tw;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + c * 32, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(v[j]);
            if (BIAS) a += p.bias_p2[c * 32 + j];
            out[(int64_t)(c * 32 + j) * p.W] = a;
          }
        }
      }
      tc_fence_before();
      epi_bar_sync();
      if (tid == 0) mbar_arrive(&tmem_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace

int launch_skip_head(const SkipHeadMaps& m, const SkipHeadParams& p, cudaStream_t s) {
  const int smem = SH_TOTAL + 1024;
  const bool bias = p.bias_skip != nullptr;
  auto k = bias ? skip_head_kernel<true> : skip_head_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[bias]) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[bias] = true;
  }
  int grid = std::min(p.n_tiles, g_sm_count);
  WN_PROF("skip_head", s);
  WN_CHECK_CUDA(launch_pdl(k, dim3((unsigned)grid), dim3(SH_THREADS), smem, s, m.zcat, m.wsk, m.p1, m.p2, m.h0, m.h1, p));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace wn
