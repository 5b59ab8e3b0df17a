// optim.cu - fused optimizer steps over the flat fp32 parameter / gradient vectors.
// Semantics = torch.optim defaults as built by get_optimizer (wavenet/train.py:28-42):
//   Adam(lr): betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad
//   SGD(lr, momentum): dampening 0, no nesterov;  RMSprop(lr, momentum): alpha 0.99, eps 1e-8, not centered
// HBM-bound: Adam moves 16 B read + 12 B written per parameter.
#include "common.cuh"

namespace wn {
namespace {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float mi = m[i] + (gi - m[i]) * (1.f - b1);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;             // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p[i] = p[i] - (lr / bc1) * (mi / denom);              // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

// the same step with the step count read from device memory (CUDA-graph replay: a captured launch cannot carry a new host scalar).
// *d_step is incremented by step_inc_kernel right before; the bias corrections are evaluated in double like the host version.
__global__ void step_inc_kernel(int32_t* d_step) { *d_step += 1; }
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                       const int32_t* __restrict__ d_step) {
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double step = (double)*d_step;
    s_bc[0] = (float)(1.0 - pow((double)b1, step));
    s_bc[1] = (float)sqrt(1.0 - pow((double)b2, step));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float mi = m[i] + (gi - m[i]) * (1.f - b1);
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                                  int64_t n, float lr, float mom, int first) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (mom != 0.f) {
      float b = first ? gi : buf[i] * mom + gi;
      buf[i] = b;
      gi = b;
    }
    p[i] -= lr * gi;
  }
}

__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq,
                                                      float* __restrict__ buf, int64_t n, float lr, float alpha, float eps, float mom) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float s = sq[i] * alpha + (1.f - alpha) * gi * gi;
    sq[i] = s;
    float avg = sqrtf(s) + eps;
    if (mom > 0.f) {
      float b = buf[i] * mom + gi / avg;
      buf[i] = b;
      p[i] -= lr * b;
    } else {
      p[i] -= lr * gi / avg;
    }
  }
}

inline int blocks_for(int64_t n) { return (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)g_sm_count * 8); }

}  // namespace
}  // namespace wn

using namespace wn;

extern "C" int wn_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                            float beta2, float eps, int32_t step, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(n > 0 && step >= 1, WN_ERR_INVALID, "wn_adam_step: bad n/step");
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  WN_PROF("adam", (cudaStream_t)stream);
  adam_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(d_params, d_grads, d_m, d_v, n, lr, beta1, beta2, eps, (float)bc1,
                                                               (float)sqrt(bc2));
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_adam_step_dev(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                                float beta2, float eps, int32_t* d_step, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(n > 0 && d_step, WN_ERR_INVALID, "wn_adam_step_dev: bad n / null step counter");
  step_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_step);
  WN_CHECK_LAUNCH();
  WN_PROF("adam", (cudaStream_t)stream);
  adam_dev_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(d_params, d_grads, d_m, d_v, n, lr, beta1, beta2, eps, d_step);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_sgd_step(float* d_params, const float* d_grads, float* d_momentum_buf, int64_t n, float lr, float momentum,
                           int32_t first_step, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(n > 0, WN_ERR_INVALID, "wn_sgd_step: bad n");
  sgd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(d_params, d_grads, d_momentum_buf, n, lr, momentum, first_step);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_rmsprop_step(float* d_params, const float* d_grads, float* d_square_avg, float* d_momentum_buf, int64_t n,
                               float lr, float alpha, float eps, float momentum, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(n > 0, WN_ERR_INVALID, "wn_rmsprop_step: bad n");
  rmsprop_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(d_params, d_grads, d_square_avg, d_momentum_buf, n, lr, alpha,
                                                                  eps, momentum);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
