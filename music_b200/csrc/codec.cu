// codec.cu - mu-law companding, wavenet/audio_func.py:5-22 (encode) and :24-39 (decode).
// Bit-exactness with the reference's torch-CPU fp32 arithmetic is obtained by construction:
// encode is monotone in x, so code(x) = #{k >= 1 : T[k] <= clamp(x,-1,1)} where T[k] is the smallest
// float the REFERENCE maps to a code >= k (bisected on the host over the float order by
// music_b200/wavenet/audio_func.py); decode has only Q possible outputs, tabulated the same way.
// HBM-bound element-wise kernels: 4 B in / 8 B out per sample (encode), 8 B in / 4 B out (decode).
#include "common.cuh"

namespace wn {
namespace {

__global__ void __launch_bounds__(256) mulaw_encode_kernel(const float* __restrict__ x, int64_t n, int q,
                                                           const float* __restrict__ thr, int64_t* __restrict__ out) {
  extern __shared__ float sthr[];
  for (int i = threadIdx.x; i < q; i += blockDim.x) sthr[i] = thr[i];
  __syncthreads();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float v = fminf(fmaxf(x[e], -1.0f), 1.0f);
    // largest k in [0,q) with (k == 0 || sthr[k] <= v)
    int lo = 0, hi = q;        // invariant: cond(lo) true, cond(hi) false
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (sthr[mid] <= v) lo = mid; else hi = mid;
    }
    out[e] = lo;
  }
}

__global__ void __launch_bounds__(256) mulaw_decode_kernel(const int64_t* __restrict__ codes, int64_t n, int q,
                                                           const float* __restrict__ values, float* __restrict__ out) {
  extern __shared__ float sval[];
  for (int i = threadIdx.x; i < q; i += blockDim.x) sval[i] = values[i];
  __syncthreads();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = codes[e];
    c = c < 0 ? 0 : (c >= q ? q - 1 : c);
    out[e] = sval[c];
  }
}

// one_hot_encode of wavenet/faster_audio_data.py:62-83 on the device.  The reference builds a (T, Q) one-hot and RESHAPES it
// to (Q, T) (:77-81), so the 1 of sample t sits at flat offset t * Q + code[t] of the (Q, T) plane; the true one-hot
// (transpose) has it at code[t] * T + t.  The plane is cleared by a memset; this kernel writes the T ones.
__global__ void __launch_bounds__(256) onehot_scatter_kernel(const int64_t* __restrict__ codes, int T, int q, int transpose,
                                                             float* __restrict__ out) {
  const int b = blockIdx.y;
  float* plane = out + (int64_t)b * q * T;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    int64_t c = codes[(int64_t)b * T + t];
    c = c < 0 ? 0 : (c >= q ? q - 1 : c);
    plane[transpose ? c * T + t : (int64_t)t * q + c] = 1.0f;
  }
}

}  // namespace
}  // namespace wn

using namespace wn;

extern "C" int wn_onehot_encode(const int64_t* d_codes, int32_t B, int32_t T, int32_t q, int32_t transpose, float* d_out, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(d_codes && d_out && B > 0 && T > 0 && q >= 2, WN_ERR_INVALID, "wn_onehot_encode: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  WN_CHECK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)B * q * T * sizeof(float), s));
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(T, 256), 64), (unsigned)B);
  onehot_scatter_kernel<<<grid, 256, 0, s>>>(d_codes, T, q, transpose, d_out);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_mulaw_encode(const float* d_audio, int64_t n, int32_t q, const float* d_thresholds, int64_t* d_codes, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(q >= 2 && q <= 4096 && n >= 0, WN_ERR_INVALID, "wn_mulaw_encode: bad q=%d n=%lld", q, (long long)n);
  if (n == 0) return WN_OK;
  int blocks = (int)std::min<int64_t>(ceil_div(n, 256 * 4), (int64_t)g_sm_count * 8);
  mulaw_encode_kernel<<<blocks, 256, q * sizeof(float), (cudaStream_t)stream>>>(d_audio, n, q, d_thresholds, d_codes);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_mulaw_decode(const int64_t* d_codes, int64_t n, int32_t q, const float* d_values, float* d_audio, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(q >= 2 && q <= 4096 && n >= 0, WN_ERR_INVALID, "wn_mulaw_decode: bad q=%d n=%lld", q, (long long)n);
  if (n == 0) return WN_OK;
  int blocks = (int)std::min<int64_t>(ceil_div(n, 256 * 4), (int64_t)g_sm_count * 8);
  mulaw_decode_kernel<<<blocks, 256, q * sizeof(float), (cudaStream_t)stream>>>(d_codes, n, q, d_values, d_audio);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
