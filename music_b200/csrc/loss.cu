// loss.cu - softmax over the reference's (-1,Q) view and the (double-softmax) cross entropy.
//   wavenet/model.py:142-144  total.view(-1, Q) -> nn.Softmax()        (rows = flat Q-chunks)
//   wavenet/train.py:146,179  nn.CrossEntropyLoss()(probabilities, target.view(-1))
// One warp per row; HBM-bound (reads the logits once, writes probabilities and/or dlogits once).
#include "common.cuh"

namespace wn {
namespace {

constexpr int MAXQ_PER_LANE = 32;   // Q <= 1024

struct RowMap {
  int Q, W, rows_mode;
  // element k of row r in the (B,Q,W) logits buffer
  __device__ __forceinline__ int64_t addr(int64_t r, int k) const {
    if (rows_mode == WN_ROWS_REFERENCE) return r * Q + k;
    int64_t b = r / W, t = r % W;
    return (b * Q + k) * (int64_t)W + t;
  }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// p = softmax(row) into registers e[]; returns nothing (e[j] holds p for k = lane + 32 j)
__device__ __forceinline__ void row_softmax(const float* __restrict__ logits, const RowMap& m, int64_t r, int lane,
                                            float (&e)[MAXQ_PER_LANE], int nper) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < MAXQ_PER_LANE; ++j) {
    if (j < nper) {
      int k = lane + 32 * j;
      e[j] = (k < m.Q) ? logits[m.addr(r, k)] : -INFINITY;
      mx = fmaxf(mx, e[j]);
    }
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAXQ_PER_LANE; ++j) {
    if (j < nper) {
      e[j] = expf(e[j] - mx);
      s += e[j];
    }
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int j = 0; j < MAXQ_PER_LANE; ++j)
    if (j < nper) e[j] *= inv;
}

__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ logits, RowMap m, int64_t n_rows,
                                                          float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int nper = (m.Q + 31) / 32;
  float e[MAXQ_PER_LANE];
  row_softmax(logits, m, r, lane, e, nper);
#pragma unroll
  for (int j = 0; j < MAXQ_PER_LANE; ++j)
    if (j < nper) {
      int k = lane + 32 * j;
      if (k < m.Q) probs[r * m.Q + k] = e[j];
    }
}

// dlogits = p * (dp - <dp,p>) per row
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ dprobs,
                                                          RowMap m, int64_t n_rows, float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int nper = (m.Q + 31) / 32;
  float p[MAXQ_PER_LANE], g[MAXQ_PER_LANE];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < MAXQ_PER_LANE; ++j)
    if (j < nper) {
      int k = lane + 32 * j;
      p[j] = (k < m.Q) ? probs[r * m.Q + k] : 0.f;
      g[j] = (k < m.Q) ? dprobs[r * m.Q + k] : 0.f;
      dot += p[j] * g[j];
    }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < MAXQ_PER_LANE; ++j)
    if (j < nper) {
      int k = lane + 32 * j;
      if (k < m.Q) dlogits[m.addr(r, k)] = p[j] * (g[j] - dot);
    }
}

// fused probabilities -> loss -> dlogits
__global__ void __launch_bounds__(256) loss_fwd_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                                                           RowMap m, int64_t n_rows, float grad_scale,
                                                           float* __restrict__ row_loss, float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int nper = (m.Q + 31) / 32;
  float p[MAXQ_PER_LANE];
  row_softmax(logits, m, r, lane, p, nper);
  const int y = (int)target[r];
  const float inv_n = grad_scale / (float)n_rows;
  // p[y]
  float py = 0.f;
#pragma unroll
  for (int j = 0; j < MAXQ_PER_LANE; ++j)
    if (j < nper && lane + 32 * j == y) py = p[j];
  py = warp_sum(py);
  if (m.rows_mode == WN_ROWS_REFERENCE) {
    // second softmax over the probabilities: loss = log sum_k exp(p_k) - p_y
    float s2 = 0.f;
    float q[MAXQ_PER_LANE];
#pragma unroll
    for (int j = 0; j < MAXQ_PER_LANE; ++j)
      if (j < nper) {
        int k = lane + 32 * j;
        q[j] = (k < m.Q) ? expf(p[j]) : 0.f;
        s2 += q[j];
      }
    s2 = warp_sum(s2);
    if (lane == 0) row_loss[r] = logf(s2) - py;
    if (dlogits) {
      const float inv2 = 1.f / s2;
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < MAXQ_PER_LANE; ++j)
        if (j < nper) {
          int k = lane + 32 * j;
          q[j] = (q[j] * inv2 - ((k == y) ? 1.f : 0.f)) * inv_n;   // g = dL/dp
          dot += q[j] * p[j];
        }
      dot = warp_sum(dot);
#pragma unroll
      for (int j = 0; j < MAXQ_PER_LANE; ++j)
        if (j < nper) {
          int k = lane + 32 * j;
          if (k < m.Q) dlogits[m.addr(r, k)] = p[j] * (q[j] - dot);
        }
    }
  } else {
    if (lane == 0) row_loss[r] = -logf(py);
    if (dlogits) {
#pragma unroll
      for (int j = 0; j < MAXQ_PER_LANE; ++j)
        if (j < nper) {
          int k = lane + 32 * j;
          if (k < m.Q) dlogits[m.addr(r, k)] = (p[j] - ((k == y) ? 1.f : 0.f)) * inv_n;
        }
    }
  }
}

// Q == 256, reference rows (contiguous): one warp per row, 8 contiguous floats per lane (two 16-byte loads),
// small register footprint so that >= 32 warps per SM are in flight: HBM-bound (1 KB read + 1 KB written per row).
__global__ void __launch_bounds__(256, 4) loss_fwd_bwd_q256_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                                                                   int64_t n_rows, float grad_scale, float* __restrict__ row_loss,
                                                                   float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const float4* src = reinterpret_cast<const float4*>(logits + r * 256 + lane * 8);
  const float4 a = src[0], b4 = src[1];
  float p[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
  float mx = p[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) mx = fmaxf(mx, p[j]);
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p[j] = expf(p[j] - mx);
    s += p[j];
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
  const int y = (int)target[r];
  float py = 0.f, s2 = 0.f, q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p[j] *= inv;
    if (lane * 8 + j == y) py = p[j];
    q[j] = expf(p[j]);
    s2 += q[j];
  }
  py = warp_sum(py);
  s2 = warp_sum(s2);
  if (lane == 0) row_loss[r] = logf(s2) - py;
  if (dlogits) {
    const float inv2 = 1.f / s2, inv_n = grad_scale / (float)n_rows;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      q[j] = (q[j] * inv2 - ((lane * 8 + j == y) ? 1.f : 0.f)) * inv_n;
      dot += q[j] * p[j];
    }
    dot = warp_sum(dot);
    float4* dst = reinterpret_cast<float4*>(dlogits + r * 256 + lane * 8);
    dst[0] = make_float4(p[0] * (q[0] - dot), p[1] * (q[1] - dot), p[2] * (q[2] - dot), p[3] * (q[3] - dot));
    dst[1] = make_float4(p[4] * (q[4] - dot), p[5] * (q[5] - dot), p[6] * (q[6] - dot), p[7] * (q[7] - dot));
  }
}

// deterministic mean of n floats in two fixed-order stages: MEAN_BLOCKS partial sums (double), then one block over them
constexpr int MEAN_BLOCKS = 128;
__global__ void __launch_bounds__(1024) partial_sum_kernel(const float* __restrict__ v, int64_t n, double* __restrict__ partial) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += (double)v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sh[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(MEAN_BLOCKS) mean_of_partials_kernel(const double* __restrict__ partial, int64_t n, float* __restrict__ out) {
  __shared__ double sh[MEAN_BLOCKS / 32];
  double s = partial[threadIdx.x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < MEAN_BLOCKS / 32; ++k) t += sh[k];
    out[0] = (float)(t / (double)n);
  }
}
// single-block variant (kept for small inputs)
__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = (float)(s / (double)n);
  }
}

}  // namespace
}  // namespace wn

using namespace wn;

extern "C" int wn_softmax_fwd(const float* d_logits, int32_t B, int32_t Q, int32_t W, int32_t rows, float* d_probs, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(Q > 0 && Q <= 1024 && B > 0 && W > 0, WN_ERR_INVALID, "wn_softmax_fwd: bad shape B=%d Q=%d W=%d", B, Q, W);
  RowMap m{Q, W, rows};
  int64_t n_rows = (int64_t)B * W;
  softmax_fwd_kernel<<<(unsigned)ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(d_logits, m, n_rows, d_probs);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_softmax_bwd(const float* d_probs, const float* d_dprobs, int32_t B, int32_t Q, int32_t W, int32_t rows,
                              float* d_dlogits, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(Q > 0 && Q <= 1024 && B > 0 && W > 0, WN_ERR_INVALID, "wn_softmax_bwd: bad shape");
  RowMap m{Q, W, rows};
  int64_t n_rows = (int64_t)B * W;
  softmax_bwd_kernel<<<(unsigned)ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(d_probs, d_dprobs, m, n_rows, d_dlogits);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_loss_scratch_bytes(int32_t B, int32_t W, size_t* bytes) {
  WN_REQUIRE(bytes && B > 0 && W > 0, WN_ERR_INVALID, "wn_loss_scratch_bytes: bad args");
  *bytes = align_up((size_t)B * W * sizeof(float), 256) + MEAN_BLOCKS * sizeof(double);     // row losses + partial sums
  return WN_OK;
}

extern "C" int wn_loss_fwd_bwd(const float* d_logits, const int64_t* d_target, int32_t B, int32_t Q, int32_t W, int32_t rows,
                               float grad_scale, float* d_loss, float* d_dlogits, void* d_scratch, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(Q > 0 && Q <= 1024 && B > 0 && W > 0 && d_scratch && d_loss, WN_ERR_INVALID, "wn_loss_fwd_bwd: bad args");
  RowMap m{Q, W, rows};
  int64_t n_rows = (int64_t)B * W;
  cudaStream_t s = (cudaStream_t)stream;
  WN_PROF("loss_fwd_bwd", s);
  if (Q == 256 && rows == WN_ROWS_REFERENCE)
    loss_fwd_bwd_q256_kernel<<<(unsigned)ceil_div(n_rows, 8), 256, 0, s>>>(d_logits, d_target, n_rows, grad_scale,
                                                                           (float*)d_scratch, d_dlogits);
  else
    loss_fwd_bwd_kernel<<<(unsigned)ceil_div(n_rows, 8), 256, 0, s>>>(d_logits, d_target, m, n_rows, grad_scale,
                                                                      (float*)d_scratch, d_dlogits);
  WN_CHECK_LAUNCH();
  if (n_rows < 64 * 1024) {
    mean_kernel<<<1, 1024, 0, s>>>((const float*)d_scratch, n_rows, d_loss);
    WN_CHECK_LAUNCH();
  } else {
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(d_scratch) + align_up((size_t)B * W * sizeof(float), 256));
    partial_sum_kernel<<<MEAN_BLOCKS, 1024, 0, s>>>((const float*)d_scratch, n_rows, partial);
    WN_CHECK_LAUNCH();
    mean_of_partials_kernel<<<1, MEAN_BLOCKS, 0, s>>>(partial, n_rows, d_loss);
    WN_CHECK_LAUNCH();
  }
  return WN_OK;
}
