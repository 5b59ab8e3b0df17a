// ae_fast.cuh - launchers of the autoencoder's encoder / conditioning kernels used by mode 1 ("bf16") of ae.cu (ae_fast.cu).
//
// The encoder of wavenet_autoencoder (wavenet_autoencoder/model1.py:137-156) is 11 % of the model's FLOPs at the shipped parameters
// and, at 32 channels, a memory- and launch-bound chain of 40 small layers: each layer is ONE kernel whose two contractions run on
// the tensor cores as register-chained mma.sync tiles (no shared memory in the data path: the accumulator layout of the first
// product is the A-fragment layout of the second), the residual stream stays fp32.  All 120 weight gradients of the stack come from
// ONE launch (operands transposed by ldmatrix.trans), bottleneck + AvgPool collapse into "pool first, then a 32 -> BW product per
// frame" (both are linear), and the N + 1 conditioning convs are one grouped GEMM forward and one backward.
#pragma once
#include "common.cuh"

namespace wn {

constexpr int kEncC = 32;      // encoder residual = dilation channels served by these kernels

// one encoder layer (model1.py:146-150): T = W_dil * relu(x_i) (taps t - d, t), x_{i+1} = x_i + W_dense * relu(T); rows [s_out, L)
int launch_enc_fwd_layer(const float* x, float* T, float* x_out, const float* w_dil, const float* w_dense, int B, int L, int d, int s_out,
                         cudaStream_t s);
// its data gradient: dT = [T > 0] (W_dense^T gX_{i+1}),  gX_i = gX_{i+1} + [x_i > 0] (W_1^T dT[t] + W_0^T dT[t + d]); rows [s_out - d, L)
int launch_enc_bwd_layer(const float* gx_next, const float* T, const float* x, float* gx, float* dT, const float* w_dil,
                         const float* w_dense, int B, int L, int d, int s_out, cudaStream_t s);
// weight gradients of every layer in one launch.  Layer i reads slot i of x / T / dT and slot i + 1 of gx (slot strides in floats).
struct EncWgradArgs {
  const float *x, *T, *dT, *gx;
  int64_t x_stride, t_stride, dt_stride, gx_stride;
  int N, B, L;
  int dil[64], s_out[64];
  int64_t w_dil[64], w_dense[64];      // offsets of the two weights of layer i in the flat gradient vector
};
int launch_enc_wgrad(const EncWgradArgs& a, float* G, cudaStream_t s);
// xbar[b, f, :] = mean of x_N over frame f (rows tw + f pool ...), enc[b, f, :] = W_b xbar (+ bias)      (model1.py:152-155)
int launch_enc_pool_bottleneck(const float* xN, const float* w_b, const float* bias, float* xbar, float* enc, int B, int L, int tw, int pool,
                               int frames, int BW, cudaStream_t s);
// backward: gxN[b, t, :] = W_b^T genc[b, frame(t), :] / pool inside the pooled range, 0 elsewhere (all L rows written);
// dW_b += genc^T xbar, db += column sums of genc
int launch_enc_pool_bottleneck_bwd(const float* genc, const float* xbar, const float* w_b, float* gxN, float* v_scratch, float* dW, float* db,
                                   int B, int L, int tw, int pool, int frames, int BW, cudaStream_t s);

// grouped small GEMMs of the conditioning convs (model1.py:178-179, 216-217) on the (M = B frames, K = BW) encoding
struct CondBlock {
  int64_t w_off, b_off;      // rows [ncols][K] of the conv weight and its bias, offsets into the conditioning vector
  float* out;                // forward: table columns of this block (row stride out_stride); backward: unused
  const float* cg;           // backward: gradient columns of this block (row stride cg_stride)
  int out_stride, cg_stride, ncols;
};
constexpr int kCondBlocksMax = 96;
struct CondBlocks {
  CondBlock blk[kCondBlocksMax];
  int n;
};
// out_j[m, c] = b_j[c] + sum_k enc[m, k] W_j[c, k]
int launch_cond_tables(const float* enc, const float* cond_params, const CondBlocks& blocks, int M, int K, cudaStream_t s);
// genc[m, k] = sum_j sum_c cg_j[m, c] W_j[c, k]
int launch_cond_bwd(const float* cond_params, const CondBlocks& blocks, float* genc, int M, int K, cudaStream_t s);

}  // namespace wn
