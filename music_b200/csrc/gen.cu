// gen.cu - incremental generation, fast_generate.predict_next (wavenet/fast_generate.py:13-141).
//
// State per stream (device, caller-owned): {int64 t; int64 last_note; float ring[sum_k d_k * R]}.
// The reference keeps, per block, a (1,R,d_k) queue that it shift-copies every step (:99-104); here
// each block has a RING of d_k vectors: the vector pushed at push-count c lives in slot c mod d_k,
// so the oldest vector (the dilated tap, column 0 of the reference queue) is slot t mod d_k and is
// the slot the new vector overwrites.  Algorithmic traffic per stream-step: one R-vector read and one
// written per block instead of the whole queue.
//
// This file holds the fp32 check-mode kernel (one CTA per stream, weights streamed from L2 in the
// transposed fp32 image) plus prime / import / export.  The bf16 persistent kernel is fast_gen.cu.
#include "check_kernels.cuh"
#include "common.cuh"
#include "fast.cuh"

namespace wn {

constexpr int GEN_MAXL = 96;

struct GenParams {
  int n_layers, R, D, S, Q, has_bias;
  int dil[GEN_MAXL];
  int ring_off[GEN_MAXL];       // in vectors (multiply by R)
  int64_t state_stride;          // bytes per stream
  // fp32 packed image offsets (floats)
  int64_t causal_wt, causal_b;
  int64_t layer0, layer_sz;      // layer block base / size
  int64_t f_wt, f_b, g_wt, g_b, d_wt, d_b, s_wt, s_b;   // within a layer block
  int64_t p1_wt, p1_b, p2_wt, p2_b;
  // optional conditioning (wn_set_conditioning): tables, frame rule parameters, first valid index of every block's output
  const float* cond_fg;
  const float* cond_head;
  int cond_frames, cond_total, cond_gate_first, rf;
  int s_out[GEN_MAXL];
};
__host__ __device__ inline int cond_frame_of(int t_local, int len, int frames) {
  return (len % frames == 0) ? t_local / (len / frames) : t_local % frames;      // model1.py:233-246
}

static int64_t gen_state_stride(const Model& m) {
  int64_t sum = 0;
  for (int d : m.dil) sum += d;
  return (int64_t)align_up(16 + (size_t)sum * m.R * sizeof(float), 16);
}

static int fill_params(const Model& m, GenParams* gp) {
  WN_REQUIRE(m.n_layers <= GEN_MAXL, WN_ERR_UNSUPPORTED, "generation supports at most %d layers", GEN_MAXL);
  GenParams& p = *gp;
  p.n_layers = m.n_layers; p.R = m.R; p.D = m.D; p.S = m.S; p.Q = m.Q; p.has_bias = m.use_bias;
  int acc = 0;
  for (int i = 0; i < m.n_layers; ++i) {
    p.dil[i] = m.dil[i];
    p.ring_off[i] = acc;
    acc += m.dil[i];
  }
  p.state_stride = gen_state_stride(m);
  Pack32 c = pack32_of(m, &m.causal);
  p.causal_wt = c.wt; p.causal_b = c.b;
  Pack32 f0 = pack32_of(m, &m.layers[0].filt), g0 = pack32_of(m, &m.layers[0].gate);
  Pack32 d0 = pack32_of(m, &m.layers[0].dense), s0 = pack32_of(m, &m.layers[0].skip);
  p.layer0 = f0.wt;
  Pack32 p1 = pack32_of(m, &m.post1), p2 = pack32_of(m, &m.post2);
  p.layer_sz = m.n_layers > 1 ? pack32_of(m, &m.layers[1].filt).wt - f0.wt : p1.wt - f0.wt;
  p.f_wt = 0; p.f_b = f0.b - f0.wt;
  p.g_wt = g0.wt - f0.wt; p.g_b = g0.b - f0.wt;
  p.d_wt = d0.wt - f0.wt; p.d_b = d0.b - f0.wt;
  p.s_wt = s0.wt - f0.wt; p.s_b = s0.b - f0.wt;
  p.p1_wt = p1.wt; p.p1_b = p1.b; p.p2_wt = p2.wt; p.p2_b = p2.b;
  p.cond_fg = nullptr; p.cond_head = nullptr; p.cond_frames = 0; p.cond_total = 0; p.cond_gate_first = 0;
  p.rf = m.rf;
  for (int i = 0; i < m.n_layers; ++i) p.s_out[i] = m.layers[i].start;
  return WN_OK;
}
static void attach_cond(GenParams* p, const wn_gen_cond& c) {
  if (c.d_fg == nullptr) return;
  p->cond_fg = c.d_fg; p->cond_head = c.d_head; p->cond_frames = c.frames; p->cond_total = c.total_len;
  p->cond_gate_first = c.gate_first;
}

namespace {

// softmax over sl[0..Q) by one thread, then greedy (topk(1), fast_generate.py:139-140) or inverse CDF
__device__ int pick_from_logits(const float* sl, int Q, const float* u) {
  float mx = sl[0];
  for (int k = 1; k < Q; ++k) mx = fmaxf(mx, sl[k]);
  float sum = 0.f;
  for (int k = 0; k < Q; ++k) sum += expf(sl[k] - mx);
  const float inv = 1.f / sum;
  if (u == nullptr) {
    int best = 0;
    float bp = -1.f;
    for (int k = 0; k < Q; ++k) {
      float p = expf(sl[k] - mx) * inv;
      if (p > bp) { bp = p; best = k; }
    }
    return best;
  }
  float total = 0.f;
  for (int k = 0; k < Q; ++k) total += expf(sl[k] - mx) * inv;
  const float thr = (*u) * total;
  float c = 0.f;
  for (int k = 0; k < Q; ++k) {
    c += expf(sl[k] - mx) * inv;
    if (c > thr) return k;
  }
  return Q - 1;
}

__global__ void __launch_bounds__(256) gen_steps_f32_kernel(GenParams p, const float* __restrict__ P, char* __restrict__ state,
                                                            int n_streams, int n_steps, int push,
                                                            const int64_t* __restrict__ first_note,
                                                            const float* __restrict__ uniforms, int64_t* __restrict__ out,
                                                            float* __restrict__ logits_out) {
  extern __shared__ float sm[];
  const int R = p.R, D = p.D, S = p.S, Q = p.Q;
  float* x = sm;              // R   block input
  float* old = x + R;         // R   dilated tap
  float* fg = old + R;        // 2D
  float* z = fg + 2 * D;      // D
  float* y = z + D;           // R   block output
  float* sk = y + R;          // S
  float* h1 = sk + S;         // S
  float* lg = h1 + S;         // Q
  __shared__ int s_note;
  const int st = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  char* sp = state + (int64_t)st * p.state_stride;
  int64_t* hdr = reinterpret_cast<int64_t*>(sp);
  float* rings = reinterpret_cast<float*>(sp + 16);
  int64_t t = hdr[0];
  int last = (int)hdr[1];
  int note = (int)first_note[st];
  for (int step = 0; step < n_steps; ++step) {
    // causal layer: W[:, last, 0] + W[:, note, 1]  (fast_generate.py:111-116)
    for (int r = tid; r < R; r += nt) {
      float v = P[p.causal_wt + (int64_t)last * R + r] + P[p.causal_wt + ((int64_t)Q + note) * R + r];
      if (p.has_bias) v += P[p.causal_b + r];
      x[r] = v;
    }
    for (int s = tid; s < S; s += nt) sk[s] = 0.f;
    last = note;
    __syncthreads();
    for (int i = 0; i < p.n_layers; ++i) {
      const int d = p.dil[i];
      float* ring = rings + (int64_t)p.ring_off[i] * R;
      const int slot = (int)(t % d);
      const float* L = P + p.layer0 + (int64_t)i * p.layer_sz;
      const float* cfg = nullptr;        // this block's conditioning vector for the current time step (2D values)
      if (p.cond_fg) {
        const int tau = p.rf + (int)t;   // absolute index of the sample being consumed
        const int f = cond_frame_of(tau - p.s_out[i], p.cond_total - p.s_out[i], p.cond_frames);
        cfg = p.cond_fg + (((int64_t)st * p.cond_frames + f) * p.n_layers + i) * 2 * D;
      }
      for (int r = tid; r < R; r += nt) old[r] = ring[(int64_t)slot * R + r];
      __syncthreads();
      // filter / gate: W0 * old + W1 * x   (:118-121 via one_layer_forward :71-95)
      for (int o = tid; o < 2 * D; o += nt) {
        const bool is_g = o >= D;
        const int oo = is_g ? o - D : o;
        const float* Wt = L + (is_g ? p.g_wt : p.f_wt);      // [tap][R][D]
        float acc = p.has_bias ? L[(is_g ? p.g_b : p.f_b) + oo] : 0.f;
        if (cfg) acc += cfg[p.cond_gate_first ? (is_g ? oo : D + oo) : o];
        for (int c = 0; c < R; ++c) acc = fmaf(Wt[(int64_t)c * D + oo], old[c], acc);
        const float* Wt1 = Wt + (int64_t)R * D;
        for (int c = 0; c < R; ++c) acc = fmaf(Wt1[(int64_t)c * D + oo], x[c], acc);
        fg[o] = acc;
      }
      __syncthreads();
      for (int o = tid; o < D; o += nt) z[o] = (1.f / (1.f + expf(-fg[D + o]))) * tanhf(fg[o]);
      __syncthreads();
      // dense + residual, skip
      for (int o = tid; o < R + S; o += nt) {
        if (o < R) {
          const float* Wt = L + p.d_wt;                      // [D][R]
          float acc = p.has_bias ? L[p.d_b + o] : 0.f;
          for (int c = 0; c < D; ++c) acc = fmaf(Wt[(int64_t)c * R + o], z[c], acc);
          y[o] = acc + x[o];
        } else {
          const int s = o - R;
          const float* Wt = L + p.s_wt;                      // [D][S]
          float acc = p.has_bias ? L[p.s_b + s] : 0.f;
          for (int c = 0; c < D; ++c) acc = fmaf(Wt[(int64_t)c * S + s], z[c], acc);
          sk[s] += acc;
        }
      }
      __syncthreads();
      // queue push (:128-129): reference pushes the block OUTPUT
      for (int r = tid; r < R; r += nt) {
        ring[(int64_t)slot * R + r] = (push == WN_PUSH_OUTPUT) ? y[r] : x[r];
        x[r] = y[r];
      }
      __syncthreads();
    }
    // head (:130-134)
    for (int s = tid; s < S; s += nt) {
      const float* Wt = P + p.p1_wt;                          // [S][S]
      float acc = p.has_bias ? P[p.p1_b + s] : 0.f;
      if (p.cond_head) {
        const int tau = p.rf + (int)t;
        const int f = cond_frame_of(tau - (p.rf - 1), p.cond_total - (p.rf - 1), p.cond_frames);
        acc += p.cond_head[((int64_t)st * p.cond_frames + f) * S + s];
      }
      for (int c = 0; c < S; ++c) acc = fmaf(Wt[(int64_t)c * S + s], fmaxf(sk[c], 0.f), acc);
      h1[s] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int q = tid; q < Q; q += nt) {
      const float* Wt = P + p.p2_wt;                          // [S][Q]
      float acc = p.has_bias ? P[p.p2_b + q] : 0.f;
      for (int c = 0; c < S; ++c) acc = fmaf(Wt[(int64_t)c * Q + q], h1[c], acc);
      lg[q] = acc;
      if (logits_out) logits_out[((int64_t)step * n_streams + st) * Q + q] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      const float* u = uniforms ? uniforms + (int64_t)step * n_streams + st : nullptr;
      int k = pick_from_logits(lg, Q, u);
      s_note = k;
      out[(int64_t)step * n_streams + st] = k;
    }
    __syncthreads();
    note = s_note;
    t += 1;
  }
  if (tid == 0) {
    hdr[0] = t;
    hdr[1] = last;
  }
}

__global__ void gen_pick_kernel(const float* __restrict__ logits, int Q, int n_streams, const float* __restrict__ uniforms,
                                int64_t* __restrict__ out) {
  int st = blockIdx.x * blockDim.x + threadIdx.x;
  if (st >= n_streams) return;
  out[st] = pick_from_logits(logits + (int64_t)st * Q, Q, uniforms ? uniforms + st : nullptr);
}

// rings <- last d_k rows of block inputs X_k (fp32 workspace, (B,L,R)), header <- {0, last prime code}
__global__ void gen_fill_f32_kernel(GenParams p, char* __restrict__ state, const float* __restrict__ X, int64_t x_stride, int L,
                                    const int64_t* __restrict__ prime_idx) {
  const int st = blockIdx.y, i = blockIdx.x;
  char* sp = state + (int64_t)st * p.state_stride;
  float* ring = reinterpret_cast<float*>(sp + 16) + (int64_t)p.ring_off[i] * p.R;
  const int d = p.dil[i];
  const float* Xi = X + x_stride * i + ((int64_t)st * L + (L - d)) * p.R;
  for (int e = threadIdx.x; e < d * p.R; e += blockDim.x) ring[e] = Xi[e];
  if (i == 0 && threadIdx.x == 0) {
    int64_t* hdr = reinterpret_cast<int64_t*>(sp);
    hdr[0] = 0;
    hdr[1] = prime_idx[(int64_t)st * L + L - 1];
  }
}

// queues (n_streams, sum d, R) oldest-first  <->  rings
__global__ void gen_xfer_kernel(GenParams p, char* __restrict__ state, float* __restrict__ queues, int64_t* __restrict__ last_note,
                                int sum_d, int to_state) {
  const int st = blockIdx.y, i = blockIdx.x;
  char* sp = state + (int64_t)st * p.state_stride;
  int64_t* hdr = reinterpret_cast<int64_t*>(sp);
  float* ring = reinterpret_cast<float*>(sp + 16) + (int64_t)p.ring_off[i] * p.R;
  float* q = queues + ((int64_t)st * sum_d + p.ring_off[i]) * p.R;
  const int d = p.dil[i];
  const int64_t t = to_state ? 0 : hdr[0];
  for (int e = threadIdx.x; e < d * p.R; e += blockDim.x) {
    int j = e / p.R, r = e % p.R;                 // j-th oldest
    int slot = (int)((t + j) % d);
    if (to_state) ring[(int64_t)slot * p.R + r] = q[e];
    else q[e] = ring[(int64_t)slot * p.R + r];
  }
  if (i == 0 && threadIdx.x == 0) {
    if (to_state) { hdr[0] = 0; hdr[1] = last_note[st]; }
    else last_note[st] = hdr[1];
  }
}

}  // namespace
}  // namespace wn

using namespace wn;

extern "C" int wn_gen_state_bytes(const wn_model* h, int32_t mode, int32_t n_streams, size_t* bytes) {
  WN_REQUIRE(h && bytes && n_streams > 0, WN_ERR_INVALID, "wn_gen_state_bytes: bad argument");
  (void)mode;
  *bytes = (size_t)gen_state_stride(h->m) * n_streams;
  return WN_OK;
}

extern "C" int wn_gen_prime(wn_model* h, int32_t mode, int32_t n_streams, const int64_t* d_prime_idx, const void* d_packed,
                            void* d_state, void* d_workspace, size_t workspace_bytes, const float* d_uniforms, int64_t* d_out,
                            float* d_logits, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(h && d_prime_idx && d_packed && d_state && d_workspace && d_out && d_logits, WN_ERR_INVALID, "wn_gen_prime: null argument");
  WN_REQUIRE(mode == WN_MODE_FP32, WN_ERR_UNSUPPORTED, "wn_gen_prime: mode %d not available (prime runs in fp32)", mode);
  const Model& m = h->m;
  cudaStream_t s = (cudaStream_t)stream;
  const int L = m.rf;                                     // assert note.size(2) == rf (fast_generate.py:30)
  size_t need = 0;
  WN_PROPAGATE(wn_workspace_bytes(h, mode, n_streams, L, &need));
  WN_REQUIRE(workspace_bytes >= need, WN_ERR_INVALID, "wn_gen_prime: workspace %zu < %zu", workspace_bytes, need);
  WN_PROPAGATE(wn_forward(h, mode, n_streams, L, nullptr, d_prime_idx, d_packed, d_workspace, d_logits, stream));
  GenParams gp;
  WN_PROPAGATE(fill_params(m, &gp));
  // X_k live at the start of the fp32 workspace with a per-layer stride (api.cu ws32_layout)
  int64_t x_stride = (int64_t)align_up((size_t)n_streams * L * m.R * sizeof(float), 256) / 4;
  dim3 grid((unsigned)m.n_layers, (unsigned)n_streams);
  gen_fill_f32_kernel<<<grid, 256, 0, s>>>(gp, (char*)d_state, (const float*)d_workspace, x_stride, L, d_prime_idx);
  WN_CHECK_LAUNCH();
  gen_pick_kernel<<<(unsigned)ceil_div(n_streams, 64), 64, 0, s>>>(d_logits, m.Q, n_streams, d_uniforms, d_out);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_gen_steps(wn_model* h, int32_t mode, int32_t n_streams, int32_t n_steps, int32_t push,
                            const int64_t* d_first_note, const void* d_packed, void* d_state, const float* d_uniforms,
                            int64_t* d_out, float* d_logits, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(h && d_first_note && d_packed && d_state && d_out, WN_ERR_INVALID, "wn_gen_steps: null argument");
  WN_REQUIRE(push == WN_PUSH_OUTPUT || push == WN_PUSH_INPUT, WN_ERR_INVALID, "wn_gen_steps: bad push mode %d", push);
  WN_REQUIRE(n_streams > 0 && n_steps >= 0, WN_ERR_INVALID, "wn_gen_steps: bad sizes");
  if (n_steps == 0) return WN_OK;
  const Model& m = h->m;
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == WN_MODE_BF16) {
    return fast_gen_steps(h->m, n_streams, n_steps, push, d_first_note, d_packed, d_state, d_uniforms, d_out, d_logits, s, &h->cond);
  }
  WN_REQUIRE(mode == WN_MODE_FP32, WN_ERR_INVALID, "wn_gen_steps: unknown mode %d", mode);
  GenParams gp;
  WN_PROPAGATE(fill_params(m, &gp));
  attach_cond(&gp, h->cond);
  size_t smem = (size_t)(3 * m.R + 3 * m.D + 2 * m.S + m.Q) * sizeof(float);
  gen_steps_f32_kernel<<<n_streams, 256, smem, s>>>(gp, (const float*)d_packed, (char*)d_state, n_streams, n_steps, push,
                                                    d_first_note, d_uniforms, d_out, d_logits);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_gen_export(const wn_model* h, int32_t mode, int32_t n_streams, const void* d_state, float* d_queues,
                             int64_t* d_last_note, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(h && d_state && d_queues && d_last_note, WN_ERR_INVALID, "wn_gen_export: null argument");
  (void)mode;
  GenParams gp;
  WN_PROPAGATE(fill_params(h->m, &gp));
  int sum_d = 0;
  for (int d : h->m.dil) sum_d += d;
  dim3 grid((unsigned)h->m.n_layers, (unsigned)n_streams);
  gen_xfer_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gp, (char*)d_state, d_queues, d_last_note, sum_d, 0);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

extern "C" int wn_gen_import(const wn_model* h, int32_t mode, int32_t n_streams, void* d_state, const float* d_queues,
                             const int64_t* d_last_note, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(h && d_state && d_queues && d_last_note, WN_ERR_INVALID, "wn_gen_import: null argument");
  (void)mode;
  GenParams gp;
  WN_PROPAGATE(fill_params(h->m, &gp));
  int sum_d = 0;
  for (int d : h->m.dil) sum_d += d;
  dim3 grid((unsigned)h->m.n_layers, (unsigned)n_streams);
  gen_xfer_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gp, (char*)d_state, const_cast<float*>(d_queues),
                                                          const_cast<int64_t*>(d_last_note), sum_d, 1);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
