// common.cuh - shared host/device helpers for libwavenet_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/wavenet_b200.h"

namespace wn {

void set_error(const char* fmt, ...);

#define WN_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      wn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return WN_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

// every kernel launch of the library goes through this: counts launches (wn_launch_count) and, when the
// profiler is on (wn_profile_enable), brackets the launch with CUDA events on its stream.
void prof_begin(const char* name, cudaStream_t s);
void prof_end(cudaStream_t s);
extern unsigned long long g_launches;
#define WN_CHECK_LAUNCH()                \
  do {                                   \
    ++wn::g_launches;                    \
    WN_CHECK_CUDA(cudaGetLastError());   \
  } while (0)
struct ProfScope {
  cudaStream_t s;
  ProfScope(const char* name, cudaStream_t st) : s(st) { prof_begin(name, st); }
  ~ProfScope() { prof_end(s); }
};
#define WN_PROF(name, stream) wn::ProfScope _prof_scope_(name, stream)

#define WN_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      wn::set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

#define WN_PROPAGATE(expr)      \
  do {                          \
    int _s = (expr);            \
    if (_s != WN_OK) return _s; \
  } while (0)

// WN_DEBUG_SYNC=1 in the environment: synchronise after every tensor-core kernel and name the one that failed
int debug_sync(const char* what, cudaStream_t s);
#define WN_DEBUG_SYNC(what, stream) WN_PROPAGATE(wn::debug_sync(what, stream))

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- parameter table: offsets into the flat fp32 vector (reference state_dict order) ----
struct ConvP {
  int64_t w = -1;   // offset of weight (out,in,k)
  int64_t b = -1;   // offset of bias (out) or -1
  int out = 0, in = 0, k = 0;
};

struct LayerP {
  ConvP filt, gate, dense, skip;
  int dilation = 1;
  int start = 0;   // first valid absolute time index of this layer's OUTPUT (s_{i+1})
};

struct Model {
  int n_layers = 0, R = 0, D = 0, S = 0, Q = 0, use_bias = 0, fw = 2;
  std::vector<int> dil;
  ConvP causal, post1, post2;
  std::vector<LayerP> layers;
  int64_t n_params = 0;
  int rf = 0;
  bool fast_ok = false;   // shapes supported by the tcgen05 kernels
  void* tmaps = nullptr;  // FastPlan cache (fast_plan.cu)
  // optional additive per-frame conditioning in the bf16 path (the decoder of wavenet_autoencoder, model1.py:158-247); the tables
  // are in the kernels' padded layout and must outlive the calls.  Conv biases of the conditioned layers are folded into them.
  const float* cond_fg = nullptr;     // (B, frames, n_layers, 128) fp32, [filter 64 | gate 64]: added to the [f|g] pre-activations
  const void* cond_fg16 = nullptr;    // the same table in the block kernels' per-thread order (launch_cond_pack16): per (b, frame, layer)
                                      // [b][layer][column group 4][chunk 4][frame][8 floats] (cond_pack16_kernel): a warp's loads are sector-contiguous
  const float* cond_head = nullptr;   // (B, frames, S) fp32: added to post_process_1's output before its ReLU
  float* cond_fg_grad = nullptr;      // backward: sums of d[f|g] / d(head pre-activation) over the rows of each frame, same shapes
  float* cond_head_grad = nullptr;    //           (must be zero-filled by the caller)
  int cond_frames = 0;
  // overlapped gradient exchange (wn_backward_set_split): record split_event once the gradients of blocks >= split_layer are final
  int split_layer = -1;
  void* split_event = nullptr;
};

// fp32 packed image: per conv Wt[k][in][out], Wtt[k][out][in], bias copy (offsets in floats)
struct Pack32 {
  int64_t wt, wtt, b;
};
std::vector<const ConvP*> all_convs(const Model& m);
Pack32 pack32_of(const Model& m, const ConvP* target, int64_t* total = nullptr);

// L2 eviction-priority hints on the TMA transfers of the bf16 path (WN_L2HINT=0 disables them; timing experiments)
inline bool l2_hints_on() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("WN_L2HINT");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}
// Launch with programmatic stream serialization (see pdl_wait() in tc05.cuh).  Only for kernels that call pdl_wait() before
// their first access to memory written by earlier kernels.  WN_PDL=0 turns the attribute off (timing experiments).
inline bool pdl_on() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("WN_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_on() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr unsigned long long kL2EvictFirst = 0x12F0000000000000ull, kL2EvictLast = 0x14F0000000000000ull;

extern int g_device;
extern int g_sm_count;
extern bool g_inited;

}  // namespace wn

struct wn_model {
  wn::Model m;
  wn_gen_cond cond{};       // optional conditioning descriptor (wn_set_conditioning); d_fg == nullptr: none
};
