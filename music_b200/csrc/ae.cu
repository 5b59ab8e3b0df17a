// ae.cu - wavenet_autoencoder forward and backward (wavenet_autoencoder/model1.py:137-268), fp32 check mode.
//
//   _encode  (:137-156): en_causal(k=2) -> N x [relu -> dilated conv(k=2) -> relu -> 1x1 -> + residual] -> bottleneck 1x1
//                        -> AvgPool1d(pool)                         => encoding (B, BW, frames), frames = floor(W / pool)
//   _decode  (:158-225): de_causal(k=2) -> N x [filter_gate conv (2D out) + cond_i(encoding) -> gate = FIRST half,
//                        filter = second half -> z = tanh(xf) sigmoid(xg) -> dense + residual ; skip on the last W]
//                        -> relu(sum skips) -> connection_1 -> + cond_N(encoding) -> relu -> connection_2
//   _conditon(:227-247): when len % frames == 0 every frame is held for len/frames steps (broadcast), otherwise the
//                        whole encoding is TILED along time: index = t mod frames.
// The reference creates the N+1 conditioning convs (bias=True) afresh, at random, on every call (:178,216); here they
// are an explicit second parameter vector so that results are reproducible.
//
// Activations are (B, L, C) fp32 in absolute time coordinates like the WaveNet check mode; everything is composed from
// the generic tap-GEMM of check_kernels.cu plus three element-wise kernels below.
#include "ae_fast.cuh"
#include "check_kernels.cuh"
#include "common.cuh"
#include "fast.cuh"
#include "fast_layout.cuh"

struct wn_ae {
  int N = 0, Q = 0, Re = 0, De = 0, BW = 0, pool = 0, Rd = 0, Dd = 0, Sd = 0, use_bias = 0, rf = 0;
  std::vector<int> dil, start;                 // start[i] = first valid time index of layer i's output
  // offsets into the flat parameter vector (reference state_dict order, model1.py:55-58)
  std::vector<wn::ConvP> en_dil, en_dense, de_fg, de_dense, de_skip;
  wn::ConvP en_causal, bottleneck, de_causal, conn1, conn2;
  std::vector<wn::ConvP> cond;                 // offsets into the conditioning vector (always biased)
  int64_t n_params = 0, n_cond = 0;
  // mode 1: the conditioned decoder (89 % of the FLOPs) runs through the bf16 tcgen05 WaveNet kernels.  `dec` is a WaveNet plan
  // whose parameter offsets point INTO the autoencoder's flat vector (filter = second half of filter_gate, gate = first half,
  // post_process_1/2 = connection_1/2), so packing reads and the backward writes the autoencoder's own parameters / gradients.
  int mode = 0;
  bool enc_fast = false;      // mode 1 and 32-channel encoder stacks: the kernels of ae_fast.cu (else the fp32 SIMT encoder)
  wn::Model dec;
  ~wn_ae() { wn::fast_release(dec); }
};

namespace wn {
namespace {

// encoding[b, f, c] = mean_{j<pool} Hb[b, t0 + f*pool + j, c]            (nn.AvgPool1d(pool), model1.py:154-155)
__global__ void avgpool_kernel(const float* __restrict__ Hb, float* __restrict__ enc, int L, int C, int t0, int pool, int frames) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)frames * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(e / C), c = (int)(e % C);
    float s = 0.f;
    for (int j = 0; j < pool; ++j) s += Hb[((int64_t)b * L + t0 + f * pool + j) * C + c];
    enc[((int64_t)b * frames + f) * C + c] = s / (float)pool;
  }
}

__device__ __forceinline__ int cond_frame(int t_local, int len, int frames) {
  return (len % frames == 0) ? t_local / (len / frames) : t_local % frames;      // model1.py:233-246
}

// Y (B,L,2D) pre-activation + cond (B,frames,2D) -> Z (B,L,D) = tanh(Y[D:2D]) * sigmoid(Y[0:D])        (model1.py:183-192)
__global__ void ae_gate_kernel(const float* __restrict__ Y, const float* __restrict__ cond, float* __restrict__ Z, int L, int D,
                               int t0, int t1, int frames) {
  const int b = blockIdx.y, len = t1 - t0;
  const int64_t n = (int64_t)len * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int tl = (int)(e / D), d = (int)(e % D);
    const int f = cond_frame(tl, len, frames);
    const float* y = Y + ((int64_t)b * L + t0 + tl) * 2 * D;
    const float* c = cond + ((int64_t)b * frames + f) * 2 * D;
    const float xg = y[d] + c[d], xf = y[D + d] + c[D + d];
    Z[((int64_t)b * L + t0 + tl) * D + d] = tanhf(xf) * (1.f / (1.f + expf(-xg)));
  }
}

// H (B,W,S) += cond (B,frames,S) with the same frame rule (model1.py:218)
__global__ void ae_cond_add_kernel(float* __restrict__ H, const float* __restrict__ cond, int W, int S, int frames) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)W * S;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int tl = (int)(e / S), s = (int)(e % S);
    H[((int64_t)b * W + tl) * S + s] += cond[((int64_t)b * frames + cond_frame(tl, W, frames)) * S + s];
  }
}

// backward of ae_gate_kernel: dYc[b,t,0:D] = d(gate pre-activation), dYc[b,t,D:2D] = d(filter pre-activation)
__global__ void ae_gate_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ cond, const float* __restrict__ dZ,
                                   float* __restrict__ dYc, int L, int D, int t0, int t1, int frames) {
  const int b = blockIdx.y, len = t1 - t0;
  const int64_t n = (int64_t)len * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int tl = (int)(e / D), d = (int)(e % D);
    const int f = cond_frame(tl, len, frames);
    const int64_t row = (int64_t)b * L + t0 + tl;
    const float* y = Y + row * 2 * D;
    const float* c = cond + ((int64_t)b * frames + f) * 2 * D;
    const float xg = y[d] + c[d], xf = y[D + d] + c[D + d];
    const float t = tanhf(xf), sg = 1.f / (1.f + expf(-xg)), dz = dZ[row * D + d];
    dYc[row * 2 * D + d] = dz * t * sg * (1.f - sg);
    dYc[row * 2 * D + D + d] = dz * sg * (1.f - t * t);
  }
}

// backward of `_conditon` (model1.py:227-247): dcond[b,f,c] = sum over the local rows tl in [0,len) that read frame f.
// G (B, *, C) rows are addressed as t0 + tl with row pitch `pitch` per batch.
__global__ void ae_cond_reduce_kernel(const float* __restrict__ G, float* __restrict__ dcond, int64_t pitch, int C, int t0,
                                      int len, int frames) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)frames * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(e / C), c = (int)(e % C);
    float s = 0.f;
    if (len % frames == 0) {
      const int m = len / frames;
      for (int j = 0; j < m; ++j) s += G[((int64_t)b * pitch + t0 + f * m + j) * C + c];
    } else {
      for (int tl = f; tl < len; tl += frames) s += G[((int64_t)b * pitch + t0 + tl) * C + c];
    }
    dcond[((int64_t)b * frames + f) * C + c] = s;
  }
}

// backward of AvgPool1d: rows t0 + f*pool + j (f < frames) get dENC[b,f,:] / pool, rows past the last full window get 0
__global__ void avgpool_bwd_kernel(const float* __restrict__ denc, float* __restrict__ dHb, int L, int C, int t0, int pool, int frames) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)(L - t0) * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int tl = (int)(e / C), c = (int)(e % C);
    const int f = tl / pool;
    dHb[((int64_t)b * L + t0 + tl) * C + c] = f < frames ? denc[((int64_t)b * frames + f) * C + c] / (float)pool : 0.f;
  }
}

__global__ void onehot_rows_kernel(const int64_t* __restrict__ idx, float* __restrict__ X, int64_t n_rows, int Q) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_rows * Q; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / Q;
    X[e] = ((int)(e % Q) == (int)idx[r]) ? 1.f : 0.f;
  }
}

inline int ew_blocks(int64_t n) { return (int)std::min<int64_t>(ceil_div(n, 256), 148 * 16); }

struct AeWs {
  float *Xin, *S0, *S1, *T, *HB, *ENC, *COND, *Y, *Z, *SK, *H1, *WT;
  // training only: per-layer saved activations and the gradient buffers of wn_ae_backward
  float *EX, *ET, *DX, *DY;             // encoder x_i (N+1), encoder conv output (N), decoder x_i (N+1), decoder pre-activation (N)
  int64_t ex_stride, et_stride, dx_stride, dy_stride;
  float *gX, *gT, *gZ, *gY, *gSK, *gH1, *gENC, *gCOND, *gHB;
  size_t bytes;
};
int64_t conv_elems(const ConvP& c) { return (int64_t)c.out * c.in * c.k; }

AeWs ae_ws(const wn_ae& a, int B, int L, bool need_onehot, void* base, bool train = false) {
  const int W = L - a.rf + 1, frames = std::max(1, W / a.pool);
  AeWs w{};
  size_t off = 0;
  auto take = [&](int64_t n) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up((size_t)std::max<int64_t>(n, 1) * sizeof(float), 256);
    return p;
  };
  const int Cmax = std::max(std::max(a.Re, a.De), std::max(a.Rd, 2 * a.Dd));
  w.Xin = take(need_onehot ? (int64_t)B * L * a.Q : 0);
  w.S0 = take((int64_t)B * L * Cmax);
  w.S1 = take((int64_t)B * L * Cmax);
  w.T = take((int64_t)B * L * Cmax);
  w.HB = take((int64_t)B * L * a.BW);
  w.ENC = take((int64_t)B * frames * a.BW);
  w.COND = take((int64_t)B * frames * std::max(2 * a.Dd, a.Sd));
  w.Y = take((int64_t)B * L * 2 * a.Dd);
  w.Z = take((int64_t)B * L * a.Dd);
  w.SK = take((int64_t)B * W * a.Sd);
  w.H1 = take((int64_t)B * W * a.Sd);
  int64_t wmax = 0;
  for (auto* v : {&a.en_dil, &a.en_dense, &a.de_fg, &a.de_dense, &a.de_skip, &a.cond})
    for (auto& c : *v) wmax = std::max(wmax, conv_elems(c));
  for (const ConvP* c : {&a.en_causal, &a.bottleneck, &a.de_causal, &a.conn1, &a.conn2}) wmax = std::max(wmax, conv_elems(*c));
  w.WT = take(2 * wmax);      // Wt + Wtt of the conv being applied (repacked per call; check mode)
  if (train) {
    w.ex_stride = (int64_t)B * L * a.Re; w.et_stride = (int64_t)B * L * a.De;
    w.dx_stride = (int64_t)B * L * a.Rd; w.dy_stride = (int64_t)B * L * 2 * a.Dd;
    w.EX = take(w.ex_stride * (a.N + 1));
    w.ET = take(w.et_stride * a.N);
    w.DX = take(w.dx_stride * (a.N + 1));
    w.DY = take(w.dy_stride * a.N);
    w.gX = take((int64_t)B * L * std::max(a.Re, a.Rd));
    w.gT = take((int64_t)B * L * a.De);
    w.gZ = take((int64_t)B * L * a.Dd);
    w.gY = take((int64_t)B * L * 2 * a.Dd);
    w.gSK = take((int64_t)B * W * a.Sd);
    w.gH1 = take((int64_t)B * W * a.Sd);
    w.gENC = take((int64_t)B * frames * a.BW);
    w.gCOND = take((int64_t)B * frames * std::max(2 * a.Dd, a.Sd));
    w.gHB = take((int64_t)B * L * a.BW);
  }
  w.bytes = off;
  return w;
}

// mode 1: what follows the fp32 workspace - packed bf16 weight image, bf16 activation workspace of the decoder, conditioning
// tables in the kernels' layout and their gradients (frame sums) in the autoencoder's raw layout
struct AeFast {
  size_t packed, ws, ctab_fg, ctab_fg16, ctab_head, cgrad_fg, cgrad_head, total;
  size_t xbar, vbar, gx_all, dt_all;      // fast encoder: pooled x_N, its gradient, per-layer gX (N + 1 slots) and dT (N slots)
};
AeFast ae_fast_layout(wn_ae& a, int B, int L, size_t base, bool train = true) {
  AeFast f{};
  const int W = L - a.rf + 1, frames = std::max(1, W / a.pool);
  size_t off = align_up(base, 1024);
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 1024);
    return o;
  };
  size_t pb = 0, wb = 0;
  fast_packed_bytes(a.dec, &pb);
  fast_workspace_bytes(a.dec, B, L, &wb);
  f.packed = take(pb);
  f.ws = take(wb);
  f.ctab_fg = take((size_t)B * frames * a.N * 128 * 4);
  f.ctab_fg16 = take((size_t)B * frames * a.N * 128 * 4);
  f.ctab_head = take((size_t)B * frames * a.Sd * 4);
  f.cgrad_fg = take((size_t)B * frames * a.N * 2 * a.Dd * 4);
  f.cgrad_head = take((size_t)B * frames * a.Sd * 4);
  if (a.enc_fast) {
    f.xbar = take((size_t)B * frames * a.Re * 4);
    f.vbar = take((size_t)B * frames * a.Re * 4);
    if (train) {
      f.gx_all = take((size_t)B * L * a.Re * 4 * (a.N + 1));
      f.dt_all = take((size_t)B * L * a.De * 4 * a.N);
    }
  }
  f.total = off;
  return f;
}

TensorView tv(const float* p, int64_t sb, int64_t st, int64_t sc, int shift = 0) {
  TensorView v;
  v.p = p; v.sb = sb; v.st = st; v.sc = sc; v.shift = shift;
  return v;
}

// Y = conv(X) for one Conv1d taken from a flat vector: repack (out,in,k) -> [k][in][out], then the generic tap-GEMM
int apply_conv(const float* params, const ConvP& c, float* wt_scratch, PwArgs a, int dilation, cudaStream_t s) {
  WN_PROPAGATE(launch_pack_f32(params + c.w, wt_scratch, wt_scratch + conv_elems(c), c.out, c.in, c.k, s));
  a.Wt = wt_scratch;
  a.bias = c.b >= 0 ? params + c.b : nullptr;
  a.n_in = c.in; a.n_out = c.out;
  a.n_taps = c.k;
  if (c.k == 2) { a.off[0] = -dilation; a.off[1] = 0; } else { a.off[0] = 0; }
  return launch_pw_gemm(a, s);
}

}  // namespace
}  // namespace wn

using namespace wn;

extern "C" int wn_ae_create(const wn_ae_config* cfg, wn_ae** out) {
  WN_REQUIRE(cfg && out, WN_ERR_INVALID, "wn_ae_create: null argument");
  WN_REQUIRE(cfg->filter_width == 2, WN_ERR_UNSUPPORTED, "wn_ae_create: filter_width %d (only 2 is supported)", cfg->filter_width);
  WN_REQUIRE(cfg->n_layers > 0 && cfg->dilations && cfg->en_pool_kernel_size > 0, WN_ERR_INVALID, "wn_ae_create: bad config");
  wn_ae* a = new wn_ae();
  a->N = cfg->n_layers; a->Q = cfg->quantization_channel;
  a->Re = cfg->en_residual_channel; a->De = cfg->en_dilation_channel; a->BW = cfg->en_bottleneck_width;
  a->pool = cfg->en_pool_kernel_size; a->Rd = cfg->de_residual_channel; a->Dd = cfg->de_dilation_channel;
  a->Sd = cfg->de_skip_channel; a->use_bias = cfg->use_bias ? 1 : 0;
  int64_t off = 0;
  auto add = [&](ConvP& c, int o, int i, int k, bool bias, int64_t& cursor) {
    c.out = o; c.in = i; c.k = k;
    c.w = cursor; cursor += (int64_t)o * i * k;
    if (bias) { c.b = cursor; cursor += o; }
  };
  const bool ub = a->use_bias != 0;
  a->en_dil.resize(a->N); a->en_dense.resize(a->N); a->de_fg.resize(a->N); a->de_dense.resize(a->N); a->de_skip.resize(a->N);
  int st = 1, sum = 0;
  for (int i = 0; i < a->N; ++i) {
    a->dil.push_back(cfg->dilations[i]);
    st += cfg->dilations[i]; sum += cfg->dilations[i];
    a->start.push_back(st);
  }
  // registration order of the reference constructor: _init_encoding, _init_decoding, _init_causal_layer, _init_connection
  for (int i = 0; i < a->N; ++i) add(a->en_dil[i], a->De, a->Re, 2, ub, off);
  for (int i = 0; i < a->N; ++i) add(a->en_dense[i], a->Re, a->De, 1, ub, off);
  for (int i = 0; i < a->N; ++i) {
    add(a->de_fg[i], 2 * a->Dd, a->Rd, 2, ub, off);
    add(a->de_dense[i], a->Rd, a->Dd, 1, ub, off);
    add(a->de_skip[i], a->Sd, a->Dd, 1, ub, off);
  }
  add(a->en_causal, a->Re, a->Q, 2, ub, off);
  add(a->bottleneck, a->BW, a->Re, 1, ub, off);
  add(a->de_causal, a->Rd, a->Q, 2, ub, off);
  add(a->conn1, a->Sd, a->Sd, 1, ub, off);
  add(a->conn2, a->Q, a->Sd, 1, ub, off);
  a->n_params = off;
  int64_t coff = 0;
  a->cond.resize(a->N + 1);
  for (int i = 0; i < a->N; ++i) add(a->cond[i], 2 * a->Dd, a->BW, 1, true, coff);
  add(a->cond[a->N], a->Sd, a->BW, 1, true, coff);
  a->n_cond = coff;
  a->rf = sum + 2;
  a->mode = cfg->mode;
  if (a->mode != 0) {
    WN_REQUIRE(a->mode == 1, WN_ERR_INVALID, "wn_ae_create: unknown mode %d", a->mode);
    Model& d = a->dec;
    d.n_layers = a->N; d.R = a->Rd; d.D = a->Dd; d.S = a->Sd; d.Q = a->Q; d.use_bias = a->use_bias; d.fw = 2;
    d.dil = a->dil; d.rf = a->rf; d.n_params = a->n_params;
    d.layers.resize(a->N);
    for (int i = 0; i < a->N; ++i) {
      LayerP& l = d.layers[i];
      const ConvP& fg = a->de_fg[i];
      l.dilation = a->dil[i]; l.start = a->start[i];
      l.gate = fg; l.gate.out = a->Dd;                                           // first half of filter_gate's outputs (model1.py:188)
      l.filt = fg; l.filt.out = a->Dd; l.filt.w = fg.w + (int64_t)a->Dd * a->Rd * 2;      // second half (:190)
      if (fg.b >= 0) l.filt.b = fg.b + a->Dd;
      l.dense = a->de_dense[i];
      l.skip = a->de_skip[i];
    }
    d.causal = a->de_causal; d.post1 = a->conn1; d.post2 = a->conn2;
    d.fast_ok = fast_supported(d);
    if (!d.fast_ok || a->use_bias) {
      set_error("wn_ae_create: mode 1 (bf16 tensor-core decoder) needs decoder residual, dilation <= 64, skip 256 or 512, quantization 256 "
                "channels and use_bias = 0 (got %d / %d / %d / %d, bias %d); use mode 0 (fp32)", a->Rd, a->Dd, a->Sd, a->Q, a->use_bias);
      delete a;
      return WN_ERR_UNSUPPORTED;
    }
    a->enc_fast = a->Re == kEncC && a->De == kEncC && a->BW % 64 == 0 && a->N <= 64 && 2 * a->Dd <= 128 &&
                  2 * a->N + a->Sd / 64 <= kCondBlocksMax;
    if (getenv("WN_AE_ENC_SIMT")) a->enc_fast = false;      // timing experiments: the fp32 SIMT encoder under the bf16 decoder
  }
  *out = a;
  return WN_OK;
}
extern "C" int wn_ae_destroy(wn_ae* a) { delete a; return WN_OK; }
extern "C" int64_t wn_ae_param_count(const wn_ae* a) { return a ? a->n_params : -1; }
extern "C" int64_t wn_ae_cond_param_count(const wn_ae* a) { return a ? a->n_cond : -1; }
extern "C" int32_t wn_ae_receptive_field(const wn_ae* a) { return a ? a->rf : -1; }

extern "C" int wn_ae_workspace_bytes(const wn_ae* a, int32_t B, int32_t L, size_t* bytes) {
  WN_REQUIRE(a && bytes && B > 0, WN_ERR_INVALID, "wn_ae_workspace_bytes: bad argument");
  WN_REQUIRE(L - a->rf + 1 > 0, WN_ERR_SHAPE, "wave sample not long enough");
  *bytes = ae_ws(*a, B, L, true, nullptr).bytes;
  if (a->mode == 1) *bytes = ae_fast_layout(*const_cast<wn_ae*>(a), B, L, *bytes, false).total;
  return WN_OK;
}

static int ae_forward_impl(wn_ae* a, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx, const float* d_params,
                           const float* d_cond, void* d_workspace, float* d_logits, float* d_encoding, void* stream, bool train) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(a && d_params && d_cond && d_workspace && d_logits, WN_ERR_INVALID, "wn_ae_forward: null argument");
  WN_REQUIRE((d_x != nullptr) != (d_idx != nullptr), WN_ERR_INVALID, "wn_ae_forward: exactly one of d_x / d_idx must be given");
  const int W = L - a->rf + 1;
  WN_REQUIRE(W > 0, WN_ERR_SHAPE, "wave sample not long enough");
  const int frames = W / a->pool;
  WN_REQUIRE(frames >= 1, WN_ERR_SHAPE, "output width %d shorter than the pooling window %d", W, a->pool);
  cudaStream_t s = (cudaStream_t)stream;
  AeWs w = ae_ws(*a, B, L, true, d_workspace, train);
  const int N = a->N, Q = a->Q, tw = L - W;
  // in training every layer writes its own slot (x_i, conv outputs) instead of the ping-pong buffers
  auto enc_x = [&](int i) { return train ? w.EX + w.ex_stride * i : ((i & 1) ? w.S1 : w.S0); };
  auto enc_t = [&](int i) { return train ? w.ET + w.et_stride * i : w.T; };
  auto dec_x = [&](int i) { return train ? w.DX + w.dx_stride * i : ((i & 1) ? w.S1 : w.S0); };
  auto dec_y = [&](int i) { return train ? w.DY + w.dy_stride * i : w.Y; };
  TensorView Xin;
  const bool enc_fast = a->mode == 1 && a->enc_fast;
  if (d_x) {
    Xin = tv(d_x, (int64_t)Q * L, 1, L);                                  // (B,Q,L) as the reference takes it
  } else if (!enc_fast) {
    onehot_rows_kernel<<<ew_blocks((int64_t)B * L * Q), 256, 0, s>>>(d_idx, w.Xin, (int64_t)B * L, Q);
    WN_CHECK_LAUNCH();
    Xin = tv(w.Xin, (int64_t)L * Q, Q, 1);
  }
  auto view = [&](const float* p, int C) { return tv(p, (int64_t)L * C, C, 1); };
  float* cur = nullptr;
  float* nxt = nullptr;
  int s_in = 1;
  if (enc_fast) {
    // ---------------------------------------------------------------- encoder on the tensor cores (ae_fast.cu; model1.py:137-156)
    const AeFast fl = ae_fast_layout(*a, B, L, w.bytes, train);
    uint8_t* base = reinterpret_cast<uint8_t*>(d_workspace);
    if (d_idx) {      // causal conv of a one-hot input = two table rows per time step
      WN_PROPAGATE(launch_pack_f32(d_params + a->en_causal.w, w.WT, w.WT + conv_elems(a->en_causal), a->en_causal.out, a->en_causal.in, 2, s));
      WN_PROPAGATE(launch_causal_idx_fwd(d_idx, w.WT, nullptr, enc_x(0), B, L, a->Re, Q, s));
    } else {
      PwArgs p; p.X = Xin; p.x_lo = 0; p.x_hi = L; p.Y = view(enc_x(0), a->Re); p.B = B; p.t0 = 1; p.t1 = L;
      WN_PROPAGATE(apply_conv(d_params, a->en_causal, w.WT, p, 1, s));
    }
    for (int i = 0; i < N; ++i)
      WN_PROPAGATE(launch_enc_fwd_layer(enc_x(i), enc_t(i), enc_x(i + 1), d_params + a->en_dil[i].w, d_params + a->en_dense[i].w, B, L,
                                        a->dil[i], a->start[i], s));
    // AvgPool1d and the 1x1 bottleneck are both linear: pool the 32 channels first, then one 32 -> BW product per frame
    WN_PROPAGATE(launch_enc_pool_bottleneck(enc_x(N), d_params + a->bottleneck.w, nullptr, reinterpret_cast<float*>(base + fl.xbar), w.ENC, B, L,
                                            tw, a->pool, frames, a->BW, s));
    if (d_encoding)
      WN_CHECK_CUDA(cudaMemcpyAsync(d_encoding, w.ENC, (size_t)B * frames * a->BW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
  // ------------------------------------------------------------------ encoder (model1.py:137-156)
  {
    PwArgs p; p.X = Xin; p.x_lo = 0; p.x_hi = L; p.Y = view(enc_x(0), a->Re); p.B = B; p.t0 = 1; p.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->en_causal, w.WT, p, 1, s));
  }
  cur = enc_x(0);
  for (int i = 0; i < N; ++i) {
    const int d = a->dil[i], s_out = s_in + d;
    nxt = enc_x(i + 1);
    float* T = enc_t(i);
    PwArgs p1; p1.X = view(cur, a->Re); p1.x_lo = s_in; p1.x_hi = L; p1.x_relu = 1; p1.Y = view(T, a->De);
    p1.B = B; p1.t0 = s_out; p1.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->en_dil[i], w.WT, p1, d, s));
    PwArgs p2; p2.X = view(T, a->De); p2.x_lo = s_out; p2.x_hi = L; p2.x_relu = 1; p2.Res = view(cur, a->Re);
    p2.Y = view(nxt, a->Re); p2.B = B; p2.t0 = s_out; p2.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->en_dense[i], w.WT, p2, 1, s));
    cur = nxt;
    s_in = s_out;
  }
  {
    PwArgs p; p.X = view(cur, a->Re); p.x_lo = tw; p.x_hi = L; p.Y = view(w.HB, a->BW); p.B = B; p.t0 = tw; p.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->bottleneck, w.WT, p, 1, s));
    dim3 grid((unsigned)ew_blocks((int64_t)frames * a->BW), (unsigned)B);
    avgpool_kernel<<<grid, 256, 0, s>>>(w.HB, w.ENC, L, a->BW, tw, a->pool, frames);
    WN_CHECK_LAUNCH();
    if (d_encoding)        // channels-last copy (B, frames, BW); `_encode` returns its transpose (B, BW, frames)
      WN_CHECK_CUDA(cudaMemcpyAsync(d_encoding, w.ENC, (size_t)B * frames * a->BW * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  }
  // ------------------------------------------------------------------ decoder (model1.py:158-225)
  if (a->mode == 1) {
    // bf16 tensor-core decoder: the N + 1 conditioning convs are applied to the (B, frames, BW) encoding in fp32 (tiny), their
    // outputs become per-frame tables that the WaveNet block kernels add to the [f|g] pre-activations (and the head GEMM to
    // connection_1's output); everything else is the conditioned WaveNet forward of fast_host.cu on the shared parameters.
    uint8_t* base = reinterpret_cast<uint8_t*>(d_workspace);
    const AeFast fl = ae_fast_layout(*a, B, L, w.bytes, train);
    float* ctab_fg = reinterpret_cast<float*>(base + fl.ctab_fg);
    float* ctab_head = reinterpret_cast<float*>(base + fl.ctab_head);
    TensorView ENCf = tv(w.ENC, (int64_t)frames * a->BW, a->BW, 1);
    if (a->enc_fast) {
      // all N + 1 conditioning convs as ONE grouped GEMM that writes the kernels' table layout directly: per layer the gate rows
      // [0, Dd) of the conv go to table columns [64, 64 + Dd), the filter rows [Dd, 2 Dd) to [0, Dd) (padding columns stay zero)
      CondBlocks cb{};
      for (int i = 0; i < N; ++i)
        for (int half = 0; half < 2; ++half) {      // 0: gate, 1: filter
          CondBlock& k = cb.blk[cb.n++];
          k.w_off = a->cond[i].w + (int64_t)half * a->Dd * a->BW; k.b_off = a->cond[i].b + half * a->Dd;
          k.out = ctab_fg + (int64_t)i * 128 + (half == 0 ? 64 : 0); k.out_stride = N * 128; k.ncols = a->Dd;
        }
      for (int c0 = 0; c0 < a->Sd; c0 += 64) {
        CondBlock& k = cb.blk[cb.n++];
        k.w_off = a->cond[N].w + (int64_t)c0 * a->BW; k.b_off = a->cond[N].b + c0;
        k.out = ctab_head + c0; k.out_stride = a->Sd; k.ncols = std::min(64, a->Sd - c0);
      }
      WN_PROPAGATE(launch_cond_tables(w.ENC, d_cond, cb, B * frames, a->BW, s));
    } else {
    for (int i = 0; i < N; ++i) {
      PwArgs pc; pc.X = ENCf; pc.x_lo = 0; pc.x_hi = frames; pc.Y = tv(w.COND, (int64_t)frames * 2 * a->Dd, 2 * a->Dd, 1);
      pc.B = B; pc.t0 = 0; pc.t1 = frames;
      WN_PROPAGATE(apply_conv(d_cond, a->cond[i], w.WT, pc, 1, s));                       // (:178-179)
      WN_PROPAGATE(launch_cond_table(w.COND, nullptr, (int64_t)B * frames, a->Dd, ctab_fg + (int64_t)i * 128, (int64_t)N * 128, s));
    }
    {
      PwArgs pc; pc.X = ENCf; pc.x_lo = 0; pc.x_hi = frames; pc.Y = tv(ctab_head, (int64_t)frames * a->Sd, a->Sd, 1);
      pc.B = B; pc.t0 = 0; pc.t1 = frames;
      WN_PROPAGATE(apply_conv(d_cond, a->cond[N], w.WT, pc, 1, s));                       // (:216-217)
    }
    }
    WN_PROPAGATE(launch_cond_pack16(ctab_fg, base + fl.ctab_fg16, (int64_t)B * frames * N, frames, N, s));
    Model& d = a->dec;
    d.cond_fg = ctab_fg; d.cond_fg16 = base + fl.ctab_fg16; d.cond_head = ctab_head; d.cond_frames = frames;
    d.cond_fg_grad = nullptr; d.cond_head_grad = nullptr;
    WN_PROPAGATE(fast_pack(d, d_params, base + fl.packed, s));
    return fast_forward(d, B, L, d_x, d_idx, base + fl.packed, base + fl.ws, d_logits, s);
  }
  {
    PwArgs p; p.X = Xin; p.x_lo = 0; p.x_hi = L; p.Y = view(dec_x(0), a->Rd); p.B = B; p.t0 = 1; p.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->de_causal, w.WT, p, 1, s));
  }
  cur = dec_x(0); s_in = 1;
  TensorView ENCv = tv(w.ENC, (int64_t)frames * a->BW, a->BW, 1);
  for (int i = 0; i < N; ++i) {
    const int d = a->dil[i], s_out = s_in + d;
    PwArgs pc; pc.X = ENCv; pc.x_lo = 0; pc.x_hi = frames; pc.Y = tv(w.COND, (int64_t)frames * 2 * a->Dd, 2 * a->Dd, 1);
    pc.B = B; pc.t0 = 0; pc.t1 = frames;
    WN_PROPAGATE(apply_conv(d_cond, a->cond[i], w.WT, pc, 1, s));                       // fresh conv on the encoding (:178-179)
    nxt = dec_x(i + 1);
    float* Yi = dec_y(i);
    PwArgs p1; p1.X = view(cur, a->Rd); p1.x_lo = s_in; p1.x_hi = L; p1.Y = view(Yi, 2 * a->Dd); p1.B = B; p1.t0 = s_out; p1.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->de_fg[i], w.WT, p1, d, s));                    // filter_gate conv (:175)
    {
      dim3 grid((unsigned)ew_blocks((int64_t)(L - s_out) * a->Dd), (unsigned)B);
      ae_gate_kernel<<<grid, 256, 0, s>>>(Yi, w.COND, w.Z, L, a->Dd, s_out, L, frames);   // _conditon + gate (:183-192)
      WN_CHECK_LAUNCH();
    }
    PwArgs p2; p2.X = view(w.Z, a->Dd); p2.x_lo = s_out; p2.x_hi = L; p2.Res = view(cur, a->Rd); p2.Y = view(nxt, a->Rd);
    p2.B = B; p2.t0 = s_out; p2.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->de_dense[i], w.WT, p2, 1, s));                 // dense + residual (:194-202)
    PwArgs p3; p3.X = view(w.Z, a->Dd); p3.x_lo = tw; p3.x_hi = L; p3.Y = tv(w.SK, (int64_t)W * a->Sd, a->Sd, 1, -tw);
    p3.accumulate = i > 0; p3.B = B; p3.t0 = tw; p3.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->de_skip[i], w.WT, p3, 1, s));                  // skip on the last W (:204-208)
    cur = nxt;
    s_in = s_out;
  }
  TensorView SKv = tv(w.SK, (int64_t)W * a->Sd, a->Sd, 1, -tw), H1v = tv(w.H1, (int64_t)W * a->Sd, a->Sd, 1, -tw);
  {
    PwArgs p; p.X = SKv; p.x_lo = tw; p.x_hi = L; p.x_relu = 1; p.Y = H1v; p.B = B; p.t0 = tw; p.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->conn1, w.WT, p, 1, s));                        // relu -> connection_1 (:210-213)
    PwArgs pc; pc.X = ENCv; pc.x_lo = 0; pc.x_hi = frames; pc.Y = tv(w.COND, (int64_t)frames * a->Sd, a->Sd, 1);
    pc.B = B; pc.t0 = 0; pc.t1 = frames;
    WN_PROPAGATE(apply_conv(d_cond, a->cond[N], w.WT, pc, 1, s));                       // second fresh conv (:216-217)
    dim3 grid((unsigned)ew_blocks((int64_t)W * a->Sd), (unsigned)B);
    ae_cond_add_kernel<<<grid, 256, 0, s>>>(w.H1, w.COND, W, a->Sd, frames);            // _conditon (:218)
    WN_CHECK_LAUNCH();
    PwArgs p2; p2.X = H1v; p2.x_lo = tw; p2.x_hi = L; p2.x_relu = 1; p2.Y = tv(d_logits, (int64_t)Q * W, 1, W, -tw);
    p2.B = B; p2.t0 = tw; p2.t1 = L;
    WN_PROPAGATE(apply_conv(d_params, a->conn2, w.WT, p2, 1, s));                       // relu -> connection_2 (:219-221)
  }
  return WN_OK;
}

extern "C" int wn_ae_forward(wn_ae* a, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx, const float* d_params,
                             const float* d_cond, void* d_workspace, float* d_logits, float* d_encoding, void* stream) {
  return ae_forward_impl(a, B, L, d_x, d_idx, d_params, d_cond, d_workspace, d_logits, d_encoding, stream, false);
}

// Conditioning tables for incremental generation with the decoder (wn_set_conditioning, include/wavenet_b200.h):
//   d_fg[n, f, i, :]   = cond_i(encoding)[n, :, f]   (2Dd channels in the conv's own order: gate first, model1.py:188-192)
//   d_head[n, f, :]    = cond_N(encoding)[n, :, f]
extern "C" int wn_ae_cond_tables(wn_ae* a, int32_t n_streams, int32_t frames, const float* d_encoding, const float* d_cond,
                                 void* d_workspace, float* d_fg, float* d_head, void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(a && d_encoding && d_cond && d_workspace && d_fg && d_head && n_streams > 0 && frames > 0, WN_ERR_INVALID,
             "wn_ae_cond_tables: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  AeWs w = ae_ws(*a, n_streams, a->rf, true, d_workspace);       // only the weight-repack scratch is used
  const int N = a->N, C2 = 2 * a->Dd;
  TensorView ENCv = tv(d_encoding, (int64_t)frames * a->BW, a->BW, 1);
  for (int i = 0; i < N; ++i) {
    PwArgs pc; pc.X = ENCv; pc.x_lo = 0; pc.x_hi = frames;
    pc.Y = tv(d_fg + (int64_t)i * C2, (int64_t)frames * N * C2, (int64_t)N * C2, 1);
    pc.B = n_streams; pc.t0 = 0; pc.t1 = frames;
    WN_PROPAGATE(apply_conv(d_cond, a->cond[i], w.WT, pc, 1, s));
  }
  PwArgs ph; ph.X = ENCv; ph.x_lo = 0; ph.x_hi = frames; ph.Y = tv(d_head, (int64_t)frames * a->Sd, a->Sd, 1);
  ph.B = n_streams; ph.t0 = 0; ph.t1 = frames;
  WN_PROPAGATE(apply_conv(d_cond, a->cond[N], w.WT, ph, 1, s));
  return WN_OK;
}

extern "C" int wn_ae_train_workspace_bytes(const wn_ae* a, int32_t B, int32_t L, size_t* bytes) {
  WN_REQUIRE(a && bytes && B > 0, WN_ERR_INVALID, "wn_ae_train_workspace_bytes: bad argument");
  WN_REQUIRE(L - a->rf + 1 > 0, WN_ERR_SHAPE, "wave sample not long enough");
  *bytes = ae_ws(*a, B, L, true, nullptr, true).bytes;
  if (a->mode == 1) *bytes = ae_fast_layout(*const_cast<wn_ae*>(a), B, L, *bytes, true).total;
  return WN_OK;
}

extern "C" int wn_ae_forward_train(wn_ae* a, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx, const float* d_params,
                                   const float* d_cond, void* d_workspace, float* d_logits, float* d_encoding, void* stream) {
  return ae_forward_impl(a, B, L, d_x, d_idx, d_params, d_cond, d_workspace, d_logits, d_encoding, stream, true);
}

// Backward of the whole autoencoder (what autograd does for model1.py:137-268 given d loss / d connection_2 output).
// The workspace must be the one wn_ae_forward_train filled for the same inputs.  d_grads has the layout of d_params,
// d_cond_grads (optional) the layout of d_cond.
extern "C" int wn_ae_backward(wn_ae* a, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx, const float* d_params,
                              const float* d_cond, void* d_workspace, const float* d_dlogits, float* d_grads, float* d_cond_grads,
                              void* stream) {
  WN_REQUIRE(g_inited, WN_ERR_UNSUPPORTED, "wn_init() has not succeeded: no sm_100 device, no fallback");
  WN_REQUIRE(a && d_params && d_cond && d_workspace && d_dlogits && d_grads, WN_ERR_INVALID, "wn_ae_backward: null argument");
  WN_REQUIRE((d_x != nullptr) != (d_idx != nullptr), WN_ERR_INVALID, "wn_ae_backward: exactly one of d_x / d_idx must be given");
  const int W = L - a->rf + 1;
  WN_REQUIRE(W > 0, WN_ERR_SHAPE, "wave sample not long enough");
  const int frames = W / a->pool;
  WN_REQUIRE(frames >= 1, WN_ERR_SHAPE, "output width %d shorter than the pooling window %d", W, a->pool);
  cudaStream_t s = (cudaStream_t)stream;
  AeWs w = ae_ws(*a, B, L, true, d_workspace, true);
  const int N = a->N, Q = a->Q, tw = L - W, Sd = a->Sd, Dd = a->Dd, Rd = a->Rd, Re = a->Re, De = a->De, BW = a->BW;
  float* G = d_grads;
  float* Gc = d_cond_grads;
  WN_CHECK_CUDA(cudaMemsetAsync(G, 0, (size_t)a->n_params * sizeof(float), s));
  if (Gc) WN_CHECK_CUDA(cudaMemsetAsync(Gc, 0, (size_t)a->n_cond * sizeof(float), s));
  auto view = [&](const float* p, int C) { return tv(p, (int64_t)L * C, C, 1); };
  auto wview = [&](const float* p) { return tv(p, (int64_t)W * Sd, Sd, 1, -tw); };
  TensorView Xin = d_x ? tv(d_x, (int64_t)Q * L, 1, L) : tv(w.Xin, (int64_t)L * Q, Q, 1);     // the one-hot copy is still in the workspace
  TensorView ENCv = tv(w.ENC, (int64_t)frames * BW, BW, 1), gENCv = tv(w.gENC, (int64_t)frames * BW, BW, 1);
  // weight gradient of a Conv1d (out,in,k): one reduction per tap; bias = column sums
  auto conv_wgrad = [&](float* Gbase, const ConvP& c, TensorView X, int x_lo, int x_hi, int x_relu, TensorView dY, int t0, int t1,
                        int dilation, int nb) -> int {
    for (int tap = 0; tap < c.k; ++tap) {
      WgArgs g;
      g.X = X; g.x_lo = x_lo; g.x_hi = x_hi; g.n_in = c.in; g.x_relu = x_relu; g.off = (c.k == 2 && tap == 0) ? -dilation : 0;
      g.dY = dY; g.n_out = c.out; g.dW = Gbase + c.w + tap; g.s_out = (int64_t)c.in * c.k; g.s_in = c.k;
      g.B = nb; g.t0 = t0; g.t1 = t1;
      WN_PROPAGATE(launch_wgrad(g, s));
    }
    if (c.b >= 0) WN_PROPAGATE(launch_colsum(dY, c.out, nb, t0, t1, Gbase + c.b, s));
    return WN_OK;
  };
  // Y (+)= W^T dY : transposed weights [k][out][in] are the second half of the repack scratch
  auto conv_dgrad = [&](const float* pbase, const ConvP& c, PwArgs p, int dilation) -> int {
    WN_PROPAGATE(launch_pack_f32(pbase + c.w, w.WT, w.WT + conv_elems(c), c.out, c.in, c.k, s));
    p.Wt = w.WT + conv_elems(c);
    p.n_in = c.out; p.n_out = c.in; p.n_taps = c.k;
    if (c.k == 2) { p.off[0] = dilation; p.off[1] = 0; } else { p.off[0] = 0; }
    return launch_pw_gemm(p, s);
  };
  // conditioning conv i applied to the encoding -> w.COND (needed again by the gate backward)
  auto cond_fwd = [&](int i, int C) -> int {
    PwArgs pc; pc.X = ENCv; pc.x_lo = 0; pc.x_hi = frames; pc.Y = tv(w.COND, (int64_t)frames * C, C, 1); pc.B = B; pc.t0 = 0; pc.t1 = frames;
    return apply_conv(d_cond, a->cond[i], w.WT, pc, 1, s);
  };
  // gradient arriving at conditioning conv i's output (already reduced over time into w.gCOND): its weights, and the encoding
  bool enc_started = false;
  auto cond_bwd = [&](int i, int C) -> int {
    TensorView gC = tv(w.gCOND, (int64_t)frames * C, C, 1);
    if (Gc) WN_PROPAGATE(conv_wgrad(Gc, a->cond[i], ENCv, 0, frames, 0, gC, 0, frames, 1, B));
    PwArgs p; p.X = gC; p.x_lo = 0; p.x_hi = frames; p.Y = gENCv; p.accumulate = enc_started ? 1 : 0; p.B = B; p.t0 = 0; p.t1 = frames;
    enc_started = true;
    return conv_dgrad(d_cond, a->cond[i], p, 1);
  };

  if (a->mode == 1) {
    // decoder backward on the tensor cores: writes every decoder gradient straight into G (shared offsets; it zero-fills G
    // first) and leaves, per conditioning conv, the sum of the gradient at its output over the rows of each frame
    uint8_t* base = reinterpret_cast<uint8_t*>(d_workspace);
    const AeFast fl = ae_fast_layout(*a, B, L, w.bytes);
    float* cg_fg = reinterpret_cast<float*>(base + fl.cgrad_fg);
    float* cg_head = reinterpret_cast<float*>(base + fl.cgrad_head);
    WN_CHECK_CUDA(cudaMemsetAsync(cg_fg, 0, (size_t)B * frames * N * 2 * Dd * sizeof(float), s));
    WN_CHECK_CUDA(cudaMemsetAsync(cg_head, 0, (size_t)B * frames * Sd * sizeof(float), s));
    Model& d = a->dec;
    d.cond_fg = reinterpret_cast<const float*>(base + fl.ctab_fg); d.cond_fg16 = base + fl.ctab_fg16;
    d.cond_head = reinterpret_cast<const float*>(base + fl.ctab_head);
    d.cond_frames = frames; d.cond_fg_grad = cg_fg; d.cond_head_grad = cg_head;
    WN_PROPAGATE(fast_backward(d, B, L, d_x, d_idx, base + fl.packed, base + fl.ws, const_cast<float*>(d_dlogits), G, s));
    // conditioning convs: their weights (optional) and the encoding, head first as in the fp32 path
    auto cond_bwd_from = [&](int i, const float* gsrc, int64_t g_bstride, int64_t g_rstride, int C, bool dgrad) -> int {
      TensorView gC = tv(gsrc, g_bstride, g_rstride, 1);
      if (Gc) WN_PROPAGATE(conv_wgrad(Gc, a->cond[i], ENCv, 0, frames, 0, gC, 0, frames, 1, B));
      if (!dgrad) return WN_OK;
      PwArgs p; p.X = gC; p.x_lo = 0; p.x_hi = frames; p.Y = gENCv; p.accumulate = enc_started ? 1 : 0; p.B = B; p.t0 = 0; p.t1 = frames;
      enc_started = true;
      (void)C;
      return conv_dgrad(d_cond, a->cond[i], p, 1);
    };
    const bool grouped = a->enc_fast;      // d loss / d encoding of all N + 1 convs in one grouped GEMM (ae_fast.cu)
    if (Gc || !grouped) {
      WN_PROPAGATE(cond_bwd_from(N, cg_head, (int64_t)frames * Sd, Sd, Sd, !grouped));
      for (int i = N - 1; i >= 0; --i)
        WN_PROPAGATE(cond_bwd_from(i, cg_fg + (int64_t)i * 2 * Dd, (int64_t)frames * N * 2 * Dd, (int64_t)N * 2 * Dd, 2 * Dd, !grouped));
    }
    if (grouped) {
      CondBlocks cb{};
      for (int i = 0; i < N; ++i) {
        CondBlock& k = cb.blk[cb.n++];
        k.w_off = a->cond[i].w; k.cg = cg_fg + (int64_t)i * 2 * Dd; k.cg_stride = N * 2 * Dd; k.ncols = 2 * Dd;
      }
      {
        CondBlock& k = cb.blk[cb.n++];
        k.w_off = a->cond[N].w; k.cg = cg_head; k.cg_stride = Sd; k.ncols = Sd;
      }
      WN_PROPAGATE(launch_cond_bwd(d_cond, cb, w.gENC, B * frames, BW, s));
    }
  } else {
  // ------------------------------------------------------------------ head (model1.py:210-221)
  TensorView dLg = tv(d_dlogits, (int64_t)Q * W, 1, W, -tw);
  TensorView SKv = wview(w.SK), H1v = wview(w.H1), gH1v = wview(w.gH1), gSKv = wview(w.gSK);
  WN_PROPAGATE(conv_wgrad(G, a->conn2, H1v, tw, L, 1, dLg, tw, L, 1, B));
  {
    PwArgs p; p.X = dLg; p.x_lo = tw; p.x_hi = L; p.Mask = H1v; p.Y = gH1v; p.B = B; p.t0 = tw; p.t1 = L;
    WN_PROPAGATE(conv_dgrad(d_params, a->conn2, p, 1));
  }
  {
    dim3 grid((unsigned)ew_blocks((int64_t)frames * Sd), (unsigned)B);
    ae_cond_reduce_kernel<<<grid, 256, 0, s>>>(w.gH1, w.gCOND, W, Sd, 0, W, frames);
    WN_CHECK_LAUNCH();
    WN_PROPAGATE(cond_bwd(N, Sd));
  }
  WN_PROPAGATE(conv_wgrad(G, a->conn1, SKv, tw, L, 1, gH1v, tw, L, 1, B));
  {
    PwArgs p; p.X = gH1v; p.x_lo = tw; p.x_hi = L; p.Mask = SKv; p.Y = gSKv; p.B = B; p.t0 = tw; p.t1 = L;
    WN_PROPAGATE(conv_dgrad(d_params, a->conn1, p, 1));
  }
  // ------------------------------------------------------------------ decoder blocks in reverse (model1.py:171-208)
  WN_CHECK_CUDA(cudaMemsetAsync(w.gX, 0, (size_t)B * L * std::max(Re, Rd) * sizeof(float), s));
  TensorView gXv = view(w.gX, Rd), gZv = view(w.gZ, Dd), gYv = view(w.gY, 2 * Dd), Zv = view(w.Z, Dd);
  for (int i = N - 1; i >= 0; --i) {
    const int d = a->dil[i], s_out = a->start[i], s_in = s_out - d;
    const float* Xi = w.DX + w.dx_stride * i;
    const float* Yi = w.DY + w.dy_stride * i;
    WN_PROPAGATE(cond_fwd(i, 2 * Dd));
    {
      dim3 grid((unsigned)ew_blocks((int64_t)(L - s_out) * Dd), (unsigned)B);
      ae_gate_kernel<<<grid, 256, 0, s>>>(Yi, w.COND, w.Z, L, Dd, s_out, L, frames);          // recompute z
      WN_CHECK_LAUNCH();
    }
    // dZ = dense^T dX_{i+1} + skip^T dSkip
    {
      PwArgs p; p.X = gXv; p.x_lo = s_out; p.x_hi = L; p.Y = gZv; p.B = B; p.t0 = s_out; p.t1 = L;
      WN_PROPAGATE(conv_dgrad(d_params, a->de_dense[i], p, 1));
      PwArgs q; q.X = gSKv; q.x_lo = tw; q.x_hi = L; q.Y = gZv; q.accumulate = 1; q.B = B; q.t0 = tw; q.t1 = L;
      WN_PROPAGATE(conv_dgrad(d_params, a->de_skip[i], q, 1));
    }
    WN_PROPAGATE(conv_wgrad(G, a->de_dense[i], Zv, s_out, L, 0, gXv, s_out, L, 1, B));
    WN_PROPAGATE(conv_wgrad(G, a->de_skip[i], Zv, tw, L, 0, gSKv, tw, L, 1, B));
    {
      dim3 grid((unsigned)ew_blocks((int64_t)(L - s_out) * Dd), (unsigned)B);
      ae_gate_bwd_kernel<<<grid, 256, 0, s>>>(Yi, w.COND, w.gZ, w.gY, L, Dd, s_out, L, frames);
      WN_CHECK_LAUNCH();
      dim3 grid2((unsigned)ew_blocks((int64_t)frames * 2 * Dd), (unsigned)B);
      ae_cond_reduce_kernel<<<grid2, 256, 0, s>>>(w.gY, w.gCOND, L, 2 * Dd, s_out, L - s_out, frames);
      WN_CHECK_LAUNCH();
      WN_PROPAGATE(cond_bwd(i, 2 * Dd));
    }
    WN_PROPAGATE(conv_wgrad(G, a->de_fg[i], view(Xi, Rd), s_in, L, 0, gYv, s_out, L, d, B));
    {   // dX_i[tau] += W1^T dY[tau] + W0^T dY[tau + d]   (the residual pass-through is already in gX)
      PwArgs p; p.X = gYv; p.x_lo = s_out; p.x_hi = L; p.Y = gXv; p.accumulate = 1; p.B = B; p.t0 = s_in; p.t1 = L;
      WN_PROPAGATE(conv_dgrad(d_params, a->de_fg[i], p, d));
    }
  }
  WN_PROPAGATE(conv_wgrad(G, a->de_causal, Xin, 0, L, 0, gXv, 1, L, 1, B));
  }      // mode 0
  if (a->mode == 1 && a->enc_fast) {
    // ---------------------------------------------------------------- encoder backward on the tensor cores (ae_fast.cu)
    uint8_t* base = reinterpret_cast<uint8_t*>(d_workspace);
    const AeFast fl = ae_fast_layout(*a, B, L, w.bytes, true);
    float* gx_all = reinterpret_cast<float*>(base + fl.gx_all);
    float* dt_all = reinterpret_cast<float*>(base + fl.dt_all);
    const int64_t slot = (int64_t)B * L * Re;
    WN_PROPAGATE(launch_enc_pool_bottleneck_bwd(w.gENC, reinterpret_cast<const float*>(base + fl.xbar), d_params + a->bottleneck.w,
                                                gx_all + slot * N, reinterpret_cast<float*>(base + fl.vbar), G + a->bottleneck.w, nullptr, B,
                                                L, tw, a->pool, frames, BW, s));
    for (int i = N - 1; i >= 0; --i)
      WN_PROPAGATE(launch_enc_bwd_layer(gx_all + slot * (i + 1), w.ET + w.et_stride * i, w.EX + w.ex_stride * i, gx_all + slot * i,
                                        dt_all + slot * i, d_params + a->en_dil[i].w, d_params + a->en_dense[i].w, B, L, a->dil[i],
                                        a->start[i], s));
    EncWgradArgs wa{};
    wa.x = w.EX; wa.T = w.ET; wa.dT = dt_all; wa.gx = gx_all;
    wa.x_stride = w.ex_stride; wa.t_stride = w.et_stride; wa.dt_stride = slot; wa.gx_stride = slot;
    wa.N = N; wa.B = B; wa.L = L;
    for (int i = 0; i < N; ++i) {
      wa.dil[i] = a->dil[i]; wa.s_out[i] = a->start[i];
      wa.w_dil[i] = a->en_dil[i].w; wa.w_dense[i] = a->en_dense[i].w;
    }
    WN_PROPAGATE(launch_enc_wgrad(wa, G, s));
    if (d_idx) {
      WN_PROPAGATE(launch_causal_idx_bwd(d_idx, gx_all, G + a->en_causal.w, B, L, Re, Q, s));
    } else {
      WN_PROPAGATE(conv_wgrad(G, a->en_causal, Xin, 0, L, 0, view(gx_all, Re), 1, L, 1, B));
    }
    return WN_OK;
  }
  // ------------------------------------------------------------------ encoder (model1.py:137-156)
  {
    dim3 grid((unsigned)ew_blocks((int64_t)W * BW), (unsigned)B);
    avgpool_bwd_kernel<<<grid, 256, 0, s>>>(w.gENC, w.gHB, L, BW, tw, a->pool, frames);
    WN_CHECK_LAUNCH();
  }
  TensorView gHBv = view(w.gHB, BW);
  WN_PROPAGATE(conv_wgrad(G, a->bottleneck, view(w.EX + w.ex_stride * N, Re), tw, L, 0, gHBv, tw, L, 1, B));
  WN_CHECK_CUDA(cudaMemsetAsync(w.gX, 0, (size_t)B * L * std::max(Re, Rd) * sizeof(float), s));
  TensorView gXe = view(w.gX, Re), gTv = view(w.gT, De);
  {
    PwArgs p; p.X = gHBv; p.x_lo = tw; p.x_hi = L; p.Y = gXe; p.B = B; p.t0 = tw; p.t1 = L;
    WN_PROPAGATE(conv_dgrad(d_params, a->bottleneck, p, 1));
  }
  for (int i = N - 1; i >= 0; --i) {
    const int d = a->dil[i], s_out = a->start[i], s_in = s_out - d;
    TensorView Xi = view(w.EX + w.ex_stride * i, Re), Ti = view(w.ET + w.et_stride * i, De);
    WN_PROPAGATE(conv_wgrad(G, a->en_dense[i], Ti, s_out, L, 1, gXe, s_out, L, 1, B));
    {   // dT = (dense^T dX_{i+1}) where T > 0
      PwArgs p; p.X = gXe; p.x_lo = s_out; p.x_hi = L; p.Mask = Ti; p.Y = gTv; p.B = B; p.t0 = s_out; p.t1 = L;
      WN_PROPAGATE(conv_dgrad(d_params, a->en_dense[i], p, 1));
    }
    WN_PROPAGATE(conv_wgrad(G, a->en_dil[i], Xi, s_in, L, 1, gTv, s_out, L, d, B));
    {   // dX_i[tau] += [x_i > 0] (W1^T dT[tau] + W0^T dT[tau + d])
      PwArgs p; p.X = gTv; p.x_lo = s_out; p.x_hi = L; p.Mask = Xi; p.Y = gXe; p.accumulate = 1; p.B = B; p.t0 = s_in; p.t1 = L;
      WN_PROPAGATE(conv_dgrad(d_params, a->en_dil[i], p, d));
    }
  }
  WN_PROPAGATE(conv_wgrad(G, a->en_causal, Xin, 0, L, 0, gXe, 1, L, 1, B));
  return WN_OK;
}
