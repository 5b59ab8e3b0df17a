// check_kernels.cuh - launchers of the fp32 SIMT "check mode" kernels (check_kernels.cu).
// They are deliberately simple: one generic tap-GEMM, one generic weight-gradient reduction and a
// few element-wise kernels.  The check mode is the 1e-4-parity path named by the north star and the
// on-GPU cross-check for the tcgen05 kernels.
#pragma once
#include "common.cuh"

namespace wn {

struct TensorView {          // element strides of a (batch, time, channel) view
  const float* p = nullptr;
  int64_t sb = 0, st = 0, sc = 0;
  int shift = 0;             // index along time = tau + shift
};

struct PwArgs {
  TensorView X;              // input
  int x_lo = 0, x_hi = 0;    // rows tau' outside [x_lo, x_hi) read as 0 (tau' = tau + tap offset)
  int n_in = 0, n_taps = 1, off[2] = {0, 0}, x_relu = 0;
  const float* Wt = nullptr; // [tap][n_in][n_out]
  const float* bias = nullptr;
  TensorView Res;            // optional residual (same tau)
  TensorView Mask;           // optional: out *= (Mask > 0)
  TensorView Y;              // output (p is written)
  int n_out = 0, accumulate = 0;
  int B = 0, t0 = 0, t1 = 0; // output rows tau in [t0, t1)
};
int launch_pw_gemm(const PwArgs& a, cudaStream_t s);

struct WgArgs {
  TensorView X;
  int x_lo = 0, x_hi = 0, n_in = 0, x_relu = 0, off = 0;
  TensorView dY;
  int n_out = 0;
  float* dW = nullptr;       // dW[o*s_out + i*s_in] += sum_rows X[row+off, i] * dY[row, o]
  int64_t s_out = 0, s_in = 0;
  int B = 0, t0 = 0, t1 = 0;
};
int launch_wgrad(const WgArgs& a, cudaStream_t s);
int launch_colsum(const TensorView& dY, int n_out, int B, int t0, int t1, float* db, cudaStream_t s);

// FG is (B,L,2D): filter pre-activation in [0,D), gate pre-activation in [D,2D)
int launch_gate_fwd(const float* FG, float* Z, int B, int L, int D, int t0, int t1, cudaStream_t s);
int launch_gate_bwd(const float* FG, const float* dZ, float* dFG, int B, int L, int D, int t0, int t1, cudaStream_t s);
// x0[b,tau,:] = Wt[0][idx[b,tau-1]][:] + Wt[1][idx[b,tau]][:] (+bias), tau in [1,L)
int launch_causal_idx_fwd(const int64_t* idx, const float* Wt, const float* bias, float* X0, int B, int L, int R, int Q, cudaStream_t s);
// dW (R,Q,2) += scatter of dX0 rows
int launch_causal_idx_bwd(const int64_t* idx, const float* dX0, float* dW, int B, int L, int R, int Q, cudaStream_t s);
// Conv1d (out,in,k) -> Wt[k][in][out] and Wtt[k][out][in]
int launch_pack_f32(const float* W, float* Wt, float* Wtt, int out, int in, int k, cudaStream_t s);

}  // namespace wn
