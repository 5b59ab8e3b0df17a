// fast_bwd_kernels.cuh - parameter blocks of the backward tensor-core kernels (fast_bwd.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace wn {

// ---------------------------------------------------------------------------------------------
// gemm_nt:  OUT[b, row, n] = epi( sum_seg sum_k A_seg[b, row + a_row_off[seg], a_col0[seg] + k] *
//                                               B_seg[n0 + n, b_col0[seg] + k] )            (bf16 out)
// rows are tiled by 128 inside each batch of a 3-D tensor (cols, rows, batches); n0 = NT * n-tile.
// ---------------------------------------------------------------------------------------------
// streaming NT = 256 kernel only: EPI_RELU (out = relu(acc [+ bias]) in bf16) and EPI_LOGITS (fp32 (B,Q,W) output of the head)
// EPI_RELU_COND: out = relu(acc + cond[(b * cond_frames + frame(row)) * n_total + column]) (per-frame conditioning of the head)
enum { EPI_PLAIN = 0, EPI_MASK = 1, EPI_ADD = 2, EPI_RELU = 3, EPI_RELU_BIAS = 4, EPI_LOGITS = 5, EPI_LOGITS_BIAS = 6, EPI_RELU_COND = 7 };
struct GemmNtMaps {
  CUtensorMap a[2];   // 3-D (cols, rows, batches) box {64,128,1}
  CUtensorMap b[2];   // 2-D [N rows][K cols]       box {64, NT}
  CUtensorMap out;    // 3-D box {64,128,1}
};
struct GemmNtParams {
  int n_batches, tile0, tiles_per_batch, n_ntiles;
  int n_total;                  // output columns in total (0: every column tile is full) - resident-B kernel only
  int nk[2], a_row_off[2], a_col0[2], b_col0[2];
  int out_col0;
  int epi;
  const __nv_bfloat16* aux;     // EPI_MASK: out = aux > 0 ? acc : 0 ; EPI_ADD: out = acc + aux
  int64_t aux_bstride, aux_rstride;
  int aux_col0;
  int row_lo, row_hi;           // rows outside are written as zeros
  unsigned long long pol_a[2];  // L2 eviction hint of the A loads of each segment (0: none)
  unsigned long long pol_out;   // ... of the output stores
  // resident-B kernel only: write OUT in the tiled skip-gradient layout instead (block_bwd3 reads it):
  // [column block of 64 * n_batches + b][row / 32][column group of 16][row % 32][16 columns], out_nblk 32-row blocks per batch row
  __nv_bfloat16* out_tiled;
  int out_nblk;
  const float* bias;            // EPI_RELU_BIAS / EPI_LOGITS_BIAS: [n_total] added to the accumulator
  float* logits;                // EPI_LOGITS*: (B, lg_Q, lg_W) fp32; row j of a batch is time step j - lg_pad (rows with j < lg_pad are skipped)
  int lg_W, lg_Q, lg_pad;
  const float* cond;            // EPI_RELU_COND: (B, cond_frames, n_total) fp32; the frame follows the time step row - lg_pad of lg_W
  int cond_frames;
  const char* tag;              // profiler label (host only)
};
int launch_gemm_nt(int NT, const GemmNtMaps& m, const GemmNtParams& p, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// gemm_tn:  OUT[m, j, n] += sum_{b,row} A[b, row, a_col0 + m] * B_j[b, row + b_row_off[j], b_col[j] + n]
// (weight gradients: the reduction runs over time).  m in [0,128) (m_valid 64 or 128), j < NB blocks
// of 64 columns.  fp32 atomics into  out[(m < 64 ? out0 : out1) + (m % 64) * s_m + blk_off[j] + n * s_n].
// ---------------------------------------------------------------------------------------------
struct GemmTnMaps {
  CUtensorMap a;      // 3-D box {64,128,1}
  CUtensorMap b[2];   // 3-D box {64,128,1}
};
struct GemmTnParams {
  int n_batches, tile0, tiles_per_batch;
  int a_col0, m_valid;
  int n_mtiles;                 // "all layers" mode: 128-row m-tiles of A (0: the historical 2)
  int n_valid;                  // columns n < n_valid of every 64-column block are flushed (0: all 64) - channel-padded models
  int b_map[4], b_row_off[4], b_col[4];
  float* out0;
  float* out1;
  int64_t s_m, s_n, blk_off[4];
  // "all layers" mode (skip weight gradients): blockIdx.y = m_tile + 2 * group; group g covers layers 4g..4g+3:
  // a_col0 = 128 m_tile, b_col[j] = 64 (4g + j), blk_off[j] = y_off0 + min(4g + j, y_layers - 1) * y_stride,
  // out0/out1 shifted by 128 m_tile * s_m.  y_layers == 0: disabled.
  int y_layers;
  int64_t y_off0, y_stride;
  const char* tag;              // profiler label (host only)
};
int launch_gemm_tn(int NB, const GemmTnMaps& m, const GemmTnParams& p, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// block_bwd: recompute [f|g] of block i, dz = dx_{i+1} Wd + dzs, gate backward -> dFG, z  (one tile / CTA)
// ---------------------------------------------------------------------------------------------
struct BlockBwdMaps {
  CUtensorMap a_in, q_in, a_out, q_out;   // block_bwd6: A / Q of layer i + 1 (loads) and of layer i (stores), each (64, L, B)
  CUtensorMap x;      // x_i (64, L, B)
  CUtensorMap w0, w1; // W_fg taps [128][64]
  CUtensorMap dx;     // dx_{i+1} (64, L, B)
  CUtensorMap wdT;    // [64 d][64 r]
  CUtensorMap dfg;    // (128, L, B) store
  CUtensorMap zf;     // (64, L, B) store
  CUtensorMap dzs;    // dZcat (64 N, Wp, B): this layer's 64 columns of the skip-path gradient (load)
};
struct BlockBwdParams {
  int L, d, s_out, tile0, tiles_per_batch, has_dense;
  int tw0, tw_al, Wp;
  const __nv_bfloat16* dzs;     // [B*Wp][dzs_pitch], this layer's 64 columns start at dzs_col
  int dzs_pitch, dzs_col;
  int dzs_lb0, dzs_nblk;        // block_bwd3 (tiled dZcat): layer * B, 32-row blocks per batch row
  int d_next, own_row0;         // block_bwd6: dilation of layer i + 1 (row shift of its Q tile); first row of this layer's own first tile
  const float* bias_fg;
  unsigned long long pol_first, pol_last;   // L2 eviction hints (0: none)
  // additive per-frame conditioning of the [f|g] pre-activations (wavenet_autoencoder `_conditon`, model1.py:227-247): row tau of
  // batch b adds cond[((b * cond_frames + frame) * cond_layers + layer) * 128 + column], frame = the reference's rule on the local
  // index tau - s_out and length L - s_out.  null: none.
  const float* cond;
  const uint4* cond16;      // copy in load order: [b][layer][column group 4][chunk 4][frame][8 floats] (cond_pack16_kernel)
  int cond_frames, cond_layers, cond_layer;
  int trace;                    // WN_TS=1 (timing experiments): CTA 0 of block_bwd6 writes clock64 stamps per tile (wn_debug_ts)
};

// ---------------------------------------------------------------------------------------------
// block_bwd2: persistent, warp-specialised version that also accumulates the block's weight gradients
// (dW_filter, dW_gate for both taps and dW_dense) in TMEM while the operands are in shared memory.
// ---------------------------------------------------------------------------------------------
struct BlockBwd2Params {
  BlockBwdParams b;
  int n_batches;
  float* g_filt;      // dW filter (D,R,2) in the flat gradient vector
  float* g_gate;      // dW gate
  float* g_dense;     // dW dense (R,D,1) or null
  float* partial;     // [n_ctas][128][192] fp32 per-CTA weight-gradient tiles (reduced by wgrad_reduce_kernel)
  __nv_bfloat16* a_out_p;     // block_bwd6 (direct stores): A_i and Q_i (B, L, 64), the arrays behind BlockBwdMaps::a_out / q_out
  __nv_bfloat16* q_out_p;
  int w_order;                // block_bwd6: 0 = P, dW_dense, dW_fg; 1 = dW_fg, dW_dense, P (stage releases first)
};
int launch_block_bwd2(const BlockBwdMaps& m, const BlockBwd2Params& p, cudaStream_t s);

}  // namespace wn
