// fast_kernels.cuh - parameter blocks and launchers of the tcgen05 kernels (fast_fwd.cu, fast_bwd.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace wn {

// ---------------------------------------------------------------- forward: residual block
struct BlockFwdMaps {
  CUtensorMap x;    // x_i      (C=64, L, B) bf16, box {64,128,1}, SW128 : load (both taps)
  CUtensorMap xo;   // x_{i+1}  same geometry                               : store
  CUtensorMap w0;   // W_fg tap 0 [128 rows (f|g)][64]   box {64,128}
  CUtensorMap w1;   // W_fg tap 1
  CUtensorMap wd;   // W_dense [64][64]                  box {64,64}
  CUtensorMap z;    // Zcat (C=64*N, Wp, B), box {64,128,1}                 : store (Wp = L - tw_al)
  CUtensorMap lo;   // low half of x_i (ping-pong buffer)                    : load
  CUtensorMap loo;  // low half of x_{i+1}                                   : store
};
struct BlockFwdParams {
  int L, d, s_out;          // length, dilation, first valid output time index
  int tile0, tiles_per_batch;
  int tw0;                  // L - W: first time index that feeds the skip path
  int tw_al;                // tw0 rounded down to a tile boundary: row 0 of the (padded) skip row space
  int zcol;                 // 64 * layer
  int Wp, zpitch;           // rows per batch and row pitch (elements) of Zcat
  int has_dense;            // 0 for the last layer (its dense output is discarded, model.py:121-124)
  const float* bias_fg;     // [128] or null
  const float* bias_d;      // [64] or null
  const float* cond;        // per-frame conditioning of [f|g] (see BlockBwdParams::cond) or null
  const uint4* cond16;      // copy in load order: [b][layer][column group 4][chunk 4][frame][8 floats] (cond_pack16_kernel)
  int cond_frames, cond_layers, cond_layer;
  unsigned long long pol_first, pol_last;   // L2 eviction hints for tiles read for the last time / read by the next kernel (0: none)
  long long* ts;            // optional timestamp buffer (timing experiments): CTA 0 writes 8 clock64 values per tile
  int dbg;                  // WN_DBG bit mask (timing experiments only): 1 no Zcat store, 2 no x stores, 4 no lo load, 8 no MUFU
};
// persistent, warp-specialised variant (one CTA per SM loops over the (batch, tile) items); outputs are written
// to global memory straight from the epilogue registers
struct BlockFwdPtrs {
  const __nv_bfloat16* lo_in;    // low half of x_i      (B, L, 64)
  __nv_bfloat16* lo_out;         // low half of x_{i+1}
  __nv_bfloat16* x_out;          // x_{i+1} (hi)
  __nv_bfloat16* zcat;           // (B, Wp, zpitch)
};
int launch_block_fwd2(const BlockFwdMaps& m, const BlockFwdParams& p, const BlockFwdPtrs& g, int n_batches, cudaStream_t s);


// ---------------------------------------------------------------- forward: skip GEMM + head
struct SkipHeadMaps {
  CUtensorMap zcat;  // (64*N, B*W) box {64,128}   load
  CUtensorMap wsk;   // Wskip_cat [256][64*N] box {64,256}
  CUtensorMap p1;    // [256][256] box {64,256}
  CUtensorMap p2;    // [256][256] box {64,256}
  CUtensorMap h0;    // (256, B*W) box {64,128}    store
  CUtensorMap h1;
};
struct SkipHeadParams {
  int n_tiles, n_rows;      // rows of the padded skip row space, B * Wp
  int Wp, pad, W, Q, k_skip; // row j of a batch is time step tw = j - pad (pad = tw0 - tw_al leading rows are unused)
  float* logits;            // (B,Q,W) fp32
  const float* bias_skip;   // [256] = sum over layers of skip biases, or null
  const float* bias_p1;
  const float* bias_p2;
};
int launch_skip_head(const SkipHeadMaps& m, const SkipHeadParams& p, cudaStream_t s);

}  // namespace wn
