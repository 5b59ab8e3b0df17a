// fast_gen.cu - bf16-weight incremental generation kernel (fast_generate.predict_next step branch,
// wavenet/fast_generate.py:66-141), R = D = 64, S = Q = 256.
//
// One CTA advances SPC = 2 independent streams through n_steps samples.  Per step and stream the work is
// 31 small GEMVs (1.27 M MAC); the weights (2.4 MB bf16, [in][out] layout of the packed image) do not fit
// in shared memory, so they are streamed from L2 with 16-byte loads and every load is used for both
// streams.  State is the same per-stream ring-buffer block as the fp32 kernel (gen.cu): the dilated tap
// is slot t mod d, and the new vector overwrites it.  HBM/L2-bound by construction: per step a CTA moves
// 2.4 MB of weights + 2 x 15 KB of ring vectors; the skip sums stay in registers across the 30 blocks.
#include "fast.cuh"
#include "fast_layout.cuh"

namespace wn {
namespace {

constexpr int SPC = 2;            // streams per CTA
constexpr int GEN_MAXL = 40;      // static shared memory budget (prefetched taps: 2 x 40 x 64 floats)

struct FastGenParams {
  int n_layers, n_streams, n_steps, push, has_bias;
  int dil[GEN_MAXL];
  int ring_off[GEN_MAXL];
  int64_t state_stride;
  const float* wc_t;               // [2][256][64] fp32
  const float* bias_c;             // 64
  const float* bias_fg;            // [N][128]
  const float* bias_d;             // [N][64]
  const float* bias_skip;          // 256 (sum over layers)
  const float* bias_p1;
  const float* bias_p2;
  const __nv_bfloat16* wfgT0;      // [N][64 r][128 o]
  const __nv_bfloat16* wfgT1;
  const __nv_bfloat16* wdT;        // [N][64 d][64 r]
  const __nv_bfloat16* wsT;        // [N][64 d][256 s]
  const __nv_bfloat16* p1T;        // [256 in][256 out]
  const __nv_bfloat16* p2T;
};

__device__ __forceinline__ void fma8(float (&acc)[8], const uint4& w, float x) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    acc[2 * j] = fmaf(f.x, x, acc[2 * j]);
    acc[2 * j + 1] = fmaf(f.y, x, acc[2 * j + 1]);
  }
}

__global__ void __launch_bounds__(256, 1)
gen_steps_bf16_kernel(FastGenParams p, char* __restrict__ state, const int64_t* __restrict__ first_note,
                      const float* __restrict__ uniforms, int64_t* __restrict__ out, float* __restrict__ logits_out) {
  __shared__ __align__(16) float s_x[SPC][64];
  __shared__ __align__(16) float s_old[SPC][GEN_MAXL][64];
  __shared__ __align__(16) float s_z[SPC][64];
  __shared__ __align__(16) float s_y[SPC][64];
  __shared__ __align__(16) float s_part[16][SPC][128];     // reduction scratch (also [8][SPC][256] and [32][SPC][64])
  __shared__ __align__(16) float s_h[SPC][256];
  __shared__ __align__(16) float s_lg[SPC][256];
  __shared__ int s_note[SPC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int st0 = blockIdx.x * SPC;
  int n_act = min(SPC, p.n_streams - st0);
  char* sp[SPC];
  float* rings[SPC];
  int64_t t[SPC];
  int last[SPC], note[SPC];
#pragma unroll
  for (int s = 0; s < SPC; ++s) {
    const int st = min(st0 + s, p.n_streams - 1);
    sp[s] = state + (int64_t)st * p.state_stride;
    rings[s] = reinterpret_cast<float*>(sp[s] + 16);
    t[s] = reinterpret_cast<const int64_t*>(sp[s])[0];
    last[s] = (int)reinterpret_cast<const int64_t*>(sp[s])[1];
    note[s] = (int)first_note[st];
  }
  float* part8 = &s_part[0][0][0];                          // viewed as [8][SPC][256]
  float* part32 = &s_part[0][0][0];                         // viewed as [32][SPC][64]

  // per-thread weight slices of one block: [f|g] 8 rows x 8 outputs, dense 2 x 8, skip 8 x 8 (16-byte loads from L2)
  struct LayerW {
    uint4 fg[8], d[2], s[8];
  };
  const int fg_n0 = (tid & 15) * 8, fg_kg = tid >> 4, fg_k0 = fg_kg * 8;
  const int d_n0 = (tid & 7) * 8, d_kg = tid >> 3, d_k0 = d_kg * 2;
  const int s_n0 = (tid & 31) * 8, s_k0 = (tid >> 5) * 8;
  auto load_layer = [&](int i, LayerW& w) {
    const __nv_bfloat16* wf = (fg_k0 < 64 ? p.wfgT0 + ((int64_t)i * 64 + fg_k0) * 128 : p.wfgT1 + ((int64_t)i * 64 + (fg_k0 - 64)) * 128) + fg_n0;
#pragma unroll
    for (int k = 0; k < 8; ++k) w.fg[k] = *reinterpret_cast<const uint4*>(wf + (int64_t)k * 128);
    const __nv_bfloat16* wd = p.wdT + ((int64_t)i * 64 + d_k0) * 64 + d_n0;
    w.d[0] = *reinterpret_cast<const uint4*>(wd);
    w.d[1] = *reinterpret_cast<const uint4*>(wd + 64);
    const __nv_bfloat16* ws = p.wsT + ((int64_t)i * 64 + s_k0) * 256 + s_n0;
#pragma unroll
    for (int k = 0; k < 8; ++k) w.s[k] = *reinterpret_cast<const uint4*>(ws + (int64_t)k * 256);
  };

  LayerW wa, wb;
  load_layer(0, wa);
  for (int step = 0; step < p.n_steps; ++step) {
    // ---- prefetch every block's dilated tap (slot t mod d) and run the causal layer (a gather)
    for (int e = tid; e < SPC * p.n_layers * 16; e += 256) {
      const int s = e / (p.n_layers * 16), r = e % (p.n_layers * 16), i = r >> 4, c4 = (r & 15) * 4;
      const int slot = (int)(t[s] % p.dil[i]);
      *reinterpret_cast<float4*>(&s_old[s][i][c4]) =
          *reinterpret_cast<const float4*>(rings[s] + ((int64_t)p.ring_off[i] + slot) * 64 + c4);
    }
    if (tid < SPC * 64) {
      const int s = tid >> 6, r = tid & 63;
      float v = p.wc_t[(int64_t)last[s] * 64 + r] + p.wc_t[((int64_t)256 + note[s]) * 64 + r];
      if (p.has_bias) v += p.bias_c[r];
      s_x[s][r] = v;
    }
#pragma unroll
    for (int s = 0; s < SPC; ++s) last[s] = note[s];
    float sk[SPC][8];
#pragma unroll
    for (int s = 0; s < SPC; ++s)
#pragma unroll
      for (int j = 0; j < 8; ++j) sk[s][j] = 0.f;
    __syncthreads();

    // one residual block with its weights already in registers (wavenet/fast_generate.py:118-129)
    auto run_layer = [&](int i, const LayerW& w) {
      // ---- [f|g] = W0 old + W1 x : thread = (8 outputs, 8 inputs)
#pragma unroll
      for (int s = 0; s < SPC; ++s) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const float* in = fg_k0 < 64 ? &s_old[s][i][fg_k0] : &s_x[s][fg_k0 - 64];
#pragma unroll
        for (int k = 0; k < 8; ++k) fma8(acc, w.fg[k], in[k]);
        *reinterpret_cast<float4*>(&s_part[fg_kg][s][fg_n0]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(&s_part[fg_kg][s][fg_n0 + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      __syncthreads();
      if (tid < SPC * 64) {        // reduce the 16 partial sums, gate
        const int s = tid >> 6, c = tid & 63;
        float f = p.has_bias ? p.bias_fg[i * 128 + c] : 0.f, g = p.has_bias ? p.bias_fg[i * 128 + 64 + c] : 0.f;
#pragma unroll
        for (int kg = 0; kg < 16; ++kg) {
          f += s_part[kg][s][c];
          g += s_part[kg][s][64 + c];
        }
        s_z[s][c] = (1.f / (1.f + __expf(-g))) * tanhf(f);
      }
      __syncthreads();
      // ---- dense (64 -> 64) partials and skip (64 -> 256) accumulation in registers
#pragma unroll
      for (int s = 0; s < SPC; ++s) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        fma8(acc, w.d[0], s_z[s][d_k0]);
        fma8(acc, w.d[1], s_z[s][d_k0 + 1]);
        float* dst = part32 + ((int64_t)d_kg * SPC + s) * 64 + d_n0;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int s = 0; s < SPC; ++s) fma8(sk[s], w.s[k], s_z[s][s_k0 + k]);
      __syncthreads();
      if (tid < SPC * 64) {        // dense reduce + residual, queue push, next block input
        const int s = tid >> 6, r = tid & 63;
        float y = p.has_bias ? p.bias_d[i * 64 + r] : 0.f;
#pragma unroll
        for (int kg = 0; kg < 32; ++kg) y += part32[((int64_t)kg * SPC + s) * 64 + r];
        const float xin = s_x[s][r];
        y += xin;
        if (s < n_act) {
          const int slot = (int)(t[s] % p.dil[i]);
          rings[s][((int64_t)p.ring_off[i] + slot) * 64 + r] = (p.push == WN_PUSH_OUTPUT) ? y : xin;   // fast_generate.py:128-129
        }
        s_x[s][r] = y;
      }
      __syncthreads();
    };
    // weights of block i+1 are requested from L2 before block i is computed (ping-pong register sets)
    for (int i = 0; i < p.n_layers; i += 2) {
      if (i + 1 < p.n_layers) load_layer(i + 1, wb);
      run_layer(i, wa);
      if (i + 1 < p.n_layers) {
        if (i + 2 < p.n_layers) load_layer(i + 2, wa);
        run_layer(i + 1, wb);
      }
    }
    // ---- head weights (two 256 x 256 GEMVs, 8 groups of 32 inputs) are requested a phase ahead as well
    const int h_n0 = (tid & 31) * 8, h_k0 = (tid >> 5) * 32;
    uint4 hwa[16], hwb[16];
    auto load_head = [&](const __nv_bfloat16* Wm, int half, uint4 (&w)[16]) {
#pragma unroll
      for (int k = 0; k < 16; ++k) w[k] = *reinterpret_cast<const uint4*>(Wm + (int64_t)(h_k0 + half * 16 + k) * 256 + h_n0);
    };
    load_head(p.p1T, 0, hwa);
    // ---- skip sum: reduce the 8 input groups, bias, relu
#pragma unroll
    for (int s = 0; s < SPC; ++s) {
      float* dst = part8 + ((int64_t)(tid >> 5) * SPC + s) * 256 + (tid & 31) * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(sk[s][0], sk[s][1], sk[s][2], sk[s][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(sk[s][4], sk[s][5], sk[s][6], sk[s][7]);
    }
    __syncthreads();
    for (int e = tid; e < SPC * 256; e += 256) {
      const int s = e >> 8, c = e & 255;
      float v = p.has_bias ? p.bias_skip[c] : 0.f;
#pragma unroll
      for (int kg = 0; kg < 8; ++kg) v += part8[((int64_t)kg * SPC + s) * 256 + c];
      s_h[s][c] = fmaxf(v, 0.f);
    }
    __syncthreads();
    // ---- head: two 256 x 256 GEMVs (8 groups of 32 inputs)
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const int n0 = h_n0, k0 = h_k0;
      float acc[SPC][8];
#pragma unroll
      for (int s = 0; s < SPC; ++s)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[s][j] = 0.f;
      load_head(which == 0 ? p.p1T : p.p2T, 1, hwb);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int s = 0; s < SPC; ++s) fma8(acc[s], hwa[k], s_h[s][k0 + k]);
      }
      if (which == 0) load_head(p.p2T, 0, hwa);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int s = 0; s < SPC; ++s) fma8(acc[s], hwb[k], s_h[s][k0 + 16 + k]);
      }
#pragma unroll
      for (int s = 0; s < SPC; ++s) {
        float* dst = part8 + ((int64_t)(tid >> 5) * SPC + s) * 256 + n0;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7]);
      }
      if (which == 1 && step + 1 < p.n_steps) load_layer(0, wa);
      __syncthreads();
      for (int e = tid; e < SPC * 256; e += 256) {
        const int s = e >> 8, c = e & 255;
        const float* bias = which == 0 ? p.bias_p1 : p.bias_p2;
        float v = p.has_bias ? bias[c] : 0.f;
#pragma unroll
        for (int kg = 0; kg < 8; ++kg) v += part8[((int64_t)kg * SPC + s) * 256 + c];
        if (which == 0) s_h[s][c] = fmaxf(v, 0.f);
        else {
          s_lg[s][c] = v;
          if (logits_out && s < n_act) logits_out[((int64_t)step * p.n_streams + st0 + s) * 256 + c] = v;
        }
      }
      __syncthreads();
    }
    // ---- pick: greedy topk(1) over the softmax (fast_generate.py:138-140) or inverse CDF (extension); one warp per stream
    if (warp < SPC) {
      const int s = warp;
      float v[8];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = s_lg[s][lane * 8 + j];
        mx = fmaxf(mx, v[j]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      int pick;
      if (uniforms == nullptr) {
        int best = 1 << 30;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] == mx) best = min(best, lane * 8 + j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        pick = best;
      } else {
        // same arithmetic as the fp32 kernel / oracle: p = exp(l - max) / sum, sequential fp32 running sum
        pick = 255;
        if (lane == 0) {
          float sum = 0.f;
          for (int k = 0; k < 256; ++k) sum += expf(s_lg[s][k] - mx);
          const float inv = 1.f / sum;
          float total = 0.f;
          for (int k = 0; k < 256; ++k) total += expf(s_lg[s][k] - mx) * inv;
          const float thr = uniforms[(int64_t)step * p.n_streams + min(st0 + s, p.n_streams - 1)] * total;
          float c = 0.f;
          for (int k = 0; k < 256; ++k) {
            c += expf(s_lg[s][k] - mx) * inv;
            if (c > thr) { pick = k; break; }
          }
        }
        pick = __shfl_sync(0xffffffffu, pick, 0);
      }
      if (lane == 0) {
        s_note[s] = pick;
        if (s < n_act) out[(int64_t)step * p.n_streams + st0 + s] = pick;
      }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < SPC; ++s) {
      note[s] = s_note[s];
      t[s] += 1;
    }
  }
  if (tid < n_act) {
    reinterpret_cast<int64_t*>(sp[tid])[0] = t[tid];
    reinterpret_cast<int64_t*>(sp[tid])[1] = last[tid];
  }
}

}  // namespace

int fast_gen_steps(Model& m, int n_streams, int n_steps, int push, const int64_t* d_first_note, const void* d_packed, void* d_state,
                   const float* d_uniforms, int64_t* d_out, float* d_logits, cudaStream_t s) {
  WN_REQUIRE(fast_supported(m) && m.n_layers <= GEN_MAXL, WN_ERR_UNSUPPORTED,
             "bf16 generation is specialised for 64/64/256/256 channels and <= %d layers; use mode fp32", GEN_MAXL);
  const PackLayout pl = pack_layout(m);
  const uint8_t* P = reinterpret_cast<const uint8_t*>(d_packed);
  FastGenParams p{};
  p.n_layers = m.n_layers; p.n_streams = n_streams; p.n_steps = n_steps; p.push = push; p.has_bias = m.use_bias;
  int acc = 0;
  for (int i = 0; i < m.n_layers; ++i) {
    p.dil[i] = m.dil[i];
    p.ring_off[i] = acc;
    acc += m.dil[i];
  }
  p.state_stride = (int64_t)align_up(16 + (size_t)acc * m.R * sizeof(float), 16);
  p.wc_t = reinterpret_cast<const float*>(P + pl.wc_t);
  p.bias_c = reinterpret_cast<const float*>(P + pl.bias_c);
  p.bias_fg = reinterpret_cast<const float*>(P + pl.bias_fg);
  p.bias_d = reinterpret_cast<const float*>(P + pl.bias_d);
  p.bias_skip = reinterpret_cast<const float*>(P + pl.bias_skip);
  p.bias_p1 = reinterpret_cast<const float*>(P + pl.bias_p1);
  p.bias_p2 = reinterpret_cast<const float*>(P + pl.bias_p2);
  p.wfgT0 = reinterpret_cast<const __nv_bfloat16*>(P + pl.wfgT0);
  p.wfgT1 = reinterpret_cast<const __nv_bfloat16*>(P + pl.wfgT1);
  p.wdT = reinterpret_cast<const __nv_bfloat16*>(P + pl.wdT);
  p.wsT = reinterpret_cast<const __nv_bfloat16*>(P + pl.wsT);
  p.p1T = reinterpret_cast<const __nv_bfloat16*>(P + pl.p1T);
  p.p2T = reinterpret_cast<const __nv_bfloat16*>(P + pl.p2T);
  WN_PROF("gen_steps_bf16", s);
  gen_steps_bf16_kernel<<<(unsigned)ceil_div(n_streams, SPC), 256, 0, s>>>(p, reinterpret_cast<char*>(d_state), d_first_note,
                                                                          d_uniforms, d_out, d_logits);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace wn
