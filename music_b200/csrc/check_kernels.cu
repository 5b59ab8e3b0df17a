// check_kernels.cu - fp32 SIMT kernels of the check mode (see check_kernels.cuh).
#include "check_kernels.cuh"

namespace wn {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float ld_view(const TensorView& v, int b, int tau, int c) {
  return v.p[(int64_t)b * v.sb + (int64_t)(tau + v.shift) * v.st + (int64_t)c * v.sc];
}

// Y[b,tau,o] = (acc? Y) + bias[o] + Res[b,tau,o] + sum_tap sum_i act(X[b,tau+off,i]) * Wt[tap][i][o]
__global__ void __launch_bounds__(256) pw_gemm_kernel(PwArgs a) {
  __shared__ __align__(16) float Xs[TK][TM + 4];
  __shared__ __align__(16) float Ws[TK][TN + 4];
  const int b = blockIdx.z;
  const int row0 = a.t0 + blockIdx.x * TM;
  const int col0 = blockIdx.y * TN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool chan_contig = (a.X.sc == 1);
  // 16-byte path: channel-contiguous input whose rows / batches / base are 4-float aligned, channel counts multiples of 4
  const bool vec4 = chan_contig && (a.n_in & 3) == 0 && (a.n_out & 3) == 0 && (a.X.st & 3) == 0 && (a.X.sb & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.X.p) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.Wt) & 15) == 0;
  for (int tap = 0; tap < a.n_taps; ++tap) {
    const int off = a.off[tap];
    const float* Wtap = a.Wt + (int64_t)tap * a.n_in * a.n_out;
    for (int k0 = 0; k0 < a.n_in; k0 += TK) {
      if (vec4) {          // channels-last input and 4-aligned channel counts: one 16-byte load per thread and operand
        const int kk = (tid & 3) * 4, rr = tid >> 2;               // 64 rows x 16 channels
        const int tau = row0 + rr, k = k0 + kk, tsrc = tau + off;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tau < a.t1 && k < a.n_in && tsrc >= a.x_lo && tsrc < a.x_hi) {
          v = *reinterpret_cast<const float4*>(a.X.p + (int64_t)b * a.X.sb + (int64_t)(tsrc + a.X.shift) * a.X.st + k);
          if (a.x_relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        }
        Xs[kk][rr] = v.x; Xs[kk + 1][rr] = v.y; Xs[kk + 2][rr] = v.z; Xs[kk + 3][rr] = v.w;
        const int cc = (tid & 15) * 4, kw = tid >> 4;               // 16 k-rows x 64 output channels
        const int kg = k0 + kw, c = col0 + cc;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kg < a.n_in && c < a.n_out) w = *reinterpret_cast<const float4*>(Wtap + (int64_t)kg * a.n_out + c);
        *reinterpret_cast<float4*>(&Ws[kw][cc]) = w;
      } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int e = tid + j * 256;
        int kk, rr;
        if (chan_contig) { kk = e % TK; rr = e / TK; } else { rr = e % TM; kk = e / TM; }
        int tau = row0 + rr, k = k0 + kk;
        float v = 0.f;
        int tsrc = tau + off;
        if (tau < a.t1 && k < a.n_in && tsrc >= a.x_lo && tsrc < a.x_hi) {
          v = ld_view(a.X, b, tsrc, k);
          if (a.x_relu) v = fmaxf(v, 0.f);
        }
        Xs[kk][rr] = v;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int e = tid + j * 256;
        int cc = e % TN, kk = e / TN;
        int k = k0 + kk, c = col0 + cc;
        Ws[kk][cc] = (k < a.n_in && c < a.n_out) ? Wtap[(int64_t)k * a.n_out + c] : 0.f;
      }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float4 av = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int tau = row0 + ty * 4 + i;
    if (tau >= a.t1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = col0 + tx * 4 + j;
      if (c >= a.n_out) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[c];
      if (a.Res.p) v += ld_view(a.Res, b, tau, c);
      if (a.Mask.p) v = (ld_view(a.Mask, b, tau, c) > 0.f) ? v : 0.f;
      float* yp = const_cast<float*>(a.Y.p) + (int64_t)b * a.Y.sb + (int64_t)(tau + a.Y.shift) * a.Y.st + (int64_t)c * a.Y.sc;
      if (a.accumulate) v += *yp;
      *yp = v;
    }
  }
}

// dW[o,i] += sum_{b,tau in chunk} act(X[b,tau+off,i]) * dY[b,tau,o]
__global__ void __launch_bounds__(256) wgrad_kernel(WgArgs a, int rows_per_block, int chunks_per_batch) {
  __shared__ __align__(16) float Xs[TK][TM + 4];   // [row][i]
  __shared__ __align__(16) float Ys[TK][TN + 4];   // [row][o]
  const int b = blockIdx.z / chunks_per_batch;
  const int chunk = blockIdx.z % chunks_per_batch;
  const int i0 = blockIdx.x * TM, o0 = blockIdx.y * TN;
  const int r_begin = a.t0 + chunk * rows_per_block;
  const int r_end = min(a.t1, r_begin + rows_per_block);
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool x_chan_contig = (a.X.sc == 1);
  const bool y_chan_contig = (a.dY.sc == 1);
  for (int r0 = r_begin; r0 < r_end; r0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * 256;
      int rr, ii;
      if (x_chan_contig) { ii = e % TM; rr = e / TM; } else { rr = e % TK; ii = e / TK; }
      int tau = r0 + rr, i = i0 + ii, tsrc = tau + a.off;
      float v = 0.f;
      if (tau < r_end && i < a.n_in && tsrc >= a.x_lo && tsrc < a.x_hi) {
        v = ld_view(a.X, b, tsrc, i);
        if (a.x_relu) v = fmaxf(v, 0.f);
      }
      Xs[rr][ii] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * 256;
      int rr, oo;
      if (y_chan_contig) { oo = e % TN; rr = e / TN; } else { rr = e % TK; oo = e / TK; }
      int tau = r0 + rr, o = o0 + oo;
      Ys[rr][oo] = (tau < r_end && o < a.n_out) ? ld_view(a.dY, b, tau, o) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < TK; ++rr) {
      float4 av = *reinterpret_cast<const float4*>(&Xs[rr][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Ys[rr][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ii = i0 + ty * 4 + i;
    if (ii >= a.n_in) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = o0 + tx * 4 + j;
      if (o >= a.n_out) continue;
      atomicAdd(a.dW + (int64_t)o * a.s_out + (int64_t)ii * a.s_in, acc[i][j]);
    }
  }
}

__global__ void colsum_kernel(TensorView dY, int n_out, int t0, int t1, int rows_per_block, int chunks_per_batch, float* db) {
  const int b = blockIdx.y / chunks_per_batch, chunk = blockIdx.y % chunks_per_batch;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int r_begin = t0 + chunk * rows_per_block, r_end = min(t1, r_begin + rows_per_block);
  float s = 0.f;
  for (int tau = r_begin; tau < r_end; ++tau) s += ld_view(dY, b, tau, o);
  atomicAdd(db + o, s);
}

__global__ void gate_fwd_kernel(const float* __restrict__ FG, float* __restrict__ Z, int L, int D, int t0, int t1) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)(t1 - t0) * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int tau = t0 + (int)(e / D), d = (int)(e % D);
    const float* fg = FG + ((int64_t)b * L + tau) * 2 * D;
    float f = fg[d], g = fg[D + d];
    Z[((int64_t)b * L + tau) * D + d] = (1.f / (1.f + expf(-g))) * tanhf(f);
  }
}

__global__ void gate_bwd_kernel(const float* __restrict__ FG, const float* __restrict__ dZ, float* __restrict__ dFG,
                                int L, int D, int t0, int t1) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)(t1 - t0) * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int tau = t0 + (int)(e / D), d = (int)(e % D);
    const int64_t row = (int64_t)b * L + tau;
    float f = FG[row * 2 * D + d], g = FG[row * 2 * D + D + d];
    float t = tanhf(f), s = 1.f / (1.f + expf(-g));
    float dz = dZ[row * D + d];
    dFG[row * 2 * D + d] = dz * s * (1.f - t * t);
    dFG[row * 2 * D + D + d] = dz * t * s * (1.f - s);
  }
}

__global__ void causal_idx_fwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ Wt,
                                      const float* __restrict__ bias, float* __restrict__ X0, int L, int R, int Q) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)(L - 1) * R;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int tau = 1 + (int)(e / R), r = (int)(e % R);
    int q0 = (int)idx[(int64_t)b * L + tau - 1], q1 = (int)idx[(int64_t)b * L + tau];
    float v = Wt[(int64_t)q0 * R + r] + Wt[((int64_t)Q + q1) * R + r];
    if (bias) v += bias[r];
    X0[((int64_t)b * L + tau) * R + r] = v;
  }
}

__global__ void causal_idx_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dX0,
                                      float* __restrict__ dW, int L, int R, int Q) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)(L - 1) * R;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int tau = 1 + (int)(e / R), r = (int)(e % R);
    int q0 = (int)idx[(int64_t)b * L + tau - 1], q1 = (int)idx[(int64_t)b * L + tau];
    float g = dX0[((int64_t)b * L + tau) * R + r];
    atomicAdd(dW + ((int64_t)r * Q + q0) * 2 + 0, g);
    atomicAdd(dW + ((int64_t)r * Q + q1) * 2 + 1, g);
  }
}

__global__ void pack_f32_kernel(const float* __restrict__ W, float* __restrict__ Wt, float* __restrict__ Wtt,
                                int out, int in, int k) {
  const int64_t n = (int64_t)out * in * k;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int kk = (int)(e % k);
    int i = (int)((e / k) % in);
    int o = (int)(e / ((int64_t)k * in));
    float v = W[e];
    Wt[((int64_t)kk * in + i) * out + o] = v;
    Wtt[((int64_t)kk * out + o) * in + i] = v;
  }
}

inline int ew_blocks(int64_t n) { return (int)std::min<int64_t>(ceil_div(n, 256), 148 * 16); }

}  // namespace

int launch_pw_gemm(const PwArgs& a, cudaStream_t s) {
  if (a.t1 <= a.t0 || a.B <= 0) return WN_OK;
  dim3 grid((unsigned)ceil_div(a.t1 - a.t0, TM), (unsigned)ceil_div(a.n_out, TN), (unsigned)a.B);
  WN_PROF("check_pw_gemm", s);
  pw_gemm_kernel<<<grid, 256, 0, s>>>(a);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_wgrad(const WgArgs& a, cudaStream_t s) {
  if (a.t1 <= a.t0 || a.B <= 0) return WN_OK;
  const int rows = a.t1 - a.t0;
  // split the time axis only as far as needed to fill the GPU (~4 CTAs per SM): every chunk ends in one fp32 atomic per
  // weight, and with 512-channel convs a fixed 1024-row chunk meant 16 M atomics per call
  const int tiles = (int)(ceil_div(a.n_in, TM) * ceil_div(a.n_out, TN)) * a.B;
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(4 * g_sm_count, tiles), ceil_div(rows, 256)));
  int rows_per_block = (int)ceil_div(rows, chunks);
  rows_per_block = (rows_per_block + 15) / 16 * 16;
  chunks = (int)ceil_div(rows, rows_per_block);
  dim3 grid((unsigned)ceil_div(a.n_in, TM), (unsigned)ceil_div(a.n_out, TN), (unsigned)(a.B * chunks));
  WN_PROF("check_wgrad", s);
  wgrad_kernel<<<grid, 256, 0, s>>>(a, rows_per_block, chunks);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_colsum(const TensorView& dY, int n_out, int B, int t0, int t1, float* db, cudaStream_t s) {
  if (t1 <= t0 || B <= 0) return WN_OK;
  int rows_per_block = 512;
  int chunks = (int)ceil_div(t1 - t0, rows_per_block);
  dim3 grid((unsigned)ceil_div(n_out, 64), (unsigned)(B * chunks));
  colsum_kernel<<<grid, 64, 0, s>>>(dY, n_out, t0, t1, rows_per_block, chunks, db);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_gate_fwd(const float* FG, float* Z, int B, int L, int D, int t0, int t1, cudaStream_t s) {
  if (t1 <= t0) return WN_OK;
  dim3 grid((unsigned)ew_blocks((int64_t)(t1 - t0) * D), (unsigned)B);
  gate_fwd_kernel<<<grid, 256, 0, s>>>(FG, Z, L, D, t0, t1);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_gate_bwd(const float* FG, const float* dZ, float* dFG, int B, int L, int D, int t0, int t1, cudaStream_t s) {
  if (t1 <= t0) return WN_OK;
  dim3 grid((unsigned)ew_blocks((int64_t)(t1 - t0) * D), (unsigned)B);
  gate_bwd_kernel<<<grid, 256, 0, s>>>(FG, dZ, dFG, L, D, t0, t1);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_causal_idx_fwd(const int64_t* idx, const float* Wt, const float* bias, float* X0, int B, int L, int R, int Q, cudaStream_t s) {
  dim3 grid((unsigned)ew_blocks((int64_t)(L - 1) * R), (unsigned)B);
  causal_idx_fwd_kernel<<<grid, 256, 0, s>>>(idx, Wt, bias, X0, L, R, Q);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_causal_idx_bwd(const int64_t* idx, const float* dX0, float* dW, int B, int L, int R, int Q, cudaStream_t s) {
  dim3 grid((unsigned)ew_blocks((int64_t)(L - 1) * R), (unsigned)B);
  causal_idx_bwd_kernel<<<grid, 256, 0, s>>>(idx, dX0, dW, L, R, Q);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int launch_pack_f32(const float* W, float* Wt, float* Wtt, int out, int in, int k, cudaStream_t s) {
  pack_f32_kernel<<<ew_blocks((int64_t)out * in * k), 256, 0, s>>>(W, Wt, Wtt, out, in, k);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace wn
