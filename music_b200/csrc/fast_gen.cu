// fast_gen.cu - half-precision incremental generation kernel (fast_generate.predict_next step branch,
// wavenet/fast_generate.py:66-141), R = D = 64, S = Q = 256.
//
// One CTA advances G = 8 independent streams through n_steps samples; the 8 streams are the N dimension of
// mma.sync.m16n8k16 tiles whose A operand is the weight matrix ([out][in]) and whose B operand is the activation
// matrix ([in][stream]).  Weights and MMA operands are fp16 (11-bit significand: tighter than the bf16 of the
// training kernels), accumulation, the residual stream and the ring-buffer state are fp32.  The weights (2.47 MB)
// do not fit in one SM's shared memory, so they are streamed from L2 once per step and CTA, pre-arranged in
// A-FRAGMENT order (gen_frag_pack_kernel): every warp load is one coalesced 512-byte line that lands directly in
// the registers the MMA reads, requested 1.5 blocks ahead of its use.  Measured on B200 (tools/mma_bench.cu):
// mma.sync.m16n8k16 issues at 0.5 per clock per SM with 21 cycles latency, which is why operands are single fp16
// values rather than bf16 hi+lo pairs (half the MMAs) and why every accumulator chain is split in two.
// State is the same per-stream ring-buffer block as the fp32 kernel (gen.cu): the dilated tap is slot t mod d,
// and the new vector overwrites it; the taps of the NEXT step are fetched with cp.async while the head runs.
// Bounds (measured, profiles/): one SM pulls the 72 KB of a block's fragments from L2 in ~1150 cycles (64 B/clk), so a step
// cannot be shorter than ~20 us while every CTA streams all weights; staging them in shared memory instead (TMA bulk ring, tried)
// is no faster because the LDS traffic of the fragments then costs 576 cycles per block on the critical path.  The next step is a
// weights-stationary pipeline: fragments resident in the REGISTERS of 17 CTAs (2 blocks each + 2 head CTAs), streams hopping
// from CTA to CTA.  Latency-bound by construction (31 dependent stages per sample, 2 block-wide barriers per stage).
#include <cuda_fp16.h>
#include <cuda_pipeline.h>

#include <type_traits>

#include "fast.cuh"
#include "fast_layout.cuh"
#include "tc05.cuh"

namespace wn {
using namespace tc;
namespace {

constexpr int G = 8;              // streams per CTA
constexpr int GEN_MAXL = 40;
constexpr int XS = 72;            // padded fp32 row of 64
constexpr int XH = 72;            // padded fp16 row of 64: 36 words -> the 32 lanes of a fragment load hit 32 banks
constexpr int HH = 264;           // padded fp16 row of 256 (132 words)
constexpr int HS = 264;           // padded fp32 row of 256
// fragment image (uint4 per lane and tile): per block [f|g: 8 m-tiles x 8 k-tiles][dense: 4 x 4][skip: 16 x 4], then the head
constexpr int FRAG_FG = 0, FRAG_D = 8 * 8 * 32, FRAG_S = FRAG_D + 4 * 4 * 32, FRAG_LAYER = FRAG_S + 16 * 4 * 32;
constexpr int FRAG_HEAD = 16 * 16 * 32;

struct FastGenParams {
  int n_layers, n_streams, n_steps, push, has_bias;
  int gpc;                         // gen_pipe_kernel: groups of 8 streams per cluster (1..NG)
  int trace;                       // WN_TS=1: clock64 stamps (GEN_TS)
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
  const uint4* frag;               // A fragments: [N][FRAG_LAYER] then P1 [FRAG_HEAD], P2 [FRAG_HEAD]
  // optional conditioning (wn_set_conditioning; the autoencoder's decoder): per-frame additive terms of every block's [f|g]
  // pre-activation, [stream][frame][layer][128], and of post_process_1's, [stream][frame][256]; the frame of a time step follows
  // model1.py:233-246 (gen.cu cond_frame_of).  gen_pipe_kernel only.
  const float* cond_fg;
  const float* cond_head;
  int cond_frames, cond_total, cond_gate_first, rf;
  int s_out[GEN_MAXL];             // first valid output index of every block (the frame rule's origin)
};
__device__ __forceinline__ int gen_cond_frame(int t_local, int len, int frames) {
  return (len % frames == 0) ? t_local / (len / frames) : t_local % frames;      // model1.py:233-246
}

// ---------------------------------------------------------------- fragment image
struct FragPackArgs {
  int n_layers;
  int64_t filt[GEN_MAXL], gate[GEN_MAXL], dense[GEN_MAXL], skip[GEN_MAXL], post1, post2;     // float offsets of the Conv1d weights
  const float* params;
  uint4* frag;
};
// element (row, k) of A-fragment register e (0..7) of lane l:  row = l/4 + 8*((e>>1)&1),  k = 2*(l%4) + (e&1) + 8*(e>>2)
__global__ void __launch_bounds__(256) gen_frag_pack_kernel(FragPackArgs a) {
  const int64_t total = (int64_t)a.n_layers * FRAG_LAYER + 2 * FRAG_HEAD;
  const float* P = a.params;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int lane = (int)(e & 31);
    const int64_t tile = e >> 5;
    const int64_t layer_tiles = FRAG_LAYER / 32;
    __half v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = (lane >> 2) + 8 * ((r >> 1) & 1), kk = 2 * (lane & 3) + (r & 1) + 8 * (r >> 2);
      float w;
      if (tile < (int64_t)a.n_layers * layer_tiles) {
        const int i = (int)(tile / layer_tiles), t = (int)(tile % layer_tiles);
        if (t < 64) {               // f|g: m-tile j holds filter channels 8j..8j+7 (rows 0-7) and gate channels 8j..8j+7 (rows 8-15);
                                    // k < 64: tap 0 (applied to the queue head), k >= 64: tap 1 (applied to the new vector)
          const int mt = t >> 3, kt = t & 7, k = kt * 16 + kk;
          const int64_t base = row < 8 ? a.filt[i] : a.gate[i];
          const int o = 8 * mt + (row & 7), in = k & 63, tap = k >> 6;
          w = P[base + ((int64_t)o * 64 + in) * 2 + tap];
        } else if (t < 64 + 16) {   // dense (R out, D in, 1)
          const int u = t - 64, mt = u >> 2, kt = u & 3;
          w = P[a.dense[i] + (int64_t)(mt * 16 + row) * 64 + kt * 16 + kk];
        } else {                    // skip (S out, D in, 1)
          const int u = t - 80, mt = u >> 2, kt = u & 3;
          w = P[a.skip[i] + (int64_t)(mt * 16 + row) * 64 + kt * 16 + kk];
        }
      } else {
        const int64_t u = tile - (int64_t)a.n_layers * layer_tiles;
        const int uu = (int)(u & 255), mt = uu >> 4, kt = uu & 15;
        w = P[(u < 256 ? a.post1 : a.post2) + (int64_t)(mt * 16 + row) * 256 + kt * 16 + kk];
      }
      v[r] = __float2half_rn(w);
    }
    a.frag[e] = *reinterpret_cast<const uint4*>(v);
  }
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gate_z(float f, float g) {       // tanh(f) * sigmoid(g), sigmoid(g) = 0.5 tanh(g/2) + 0.5
  return tanh_approx(f) * fmaf(0.5f, tanh_approx(0.5f * g), 0.5f);
}

struct GenSmem {
  float x[G][XS];                  // residual stream (block input / output), fp32
  __half xh[G][XH];                // the same, fp16 (MMA operand)
  __half zh[G][XH];                // gated activations
  __half hh[2][G][HH];             // head activations (ping-pong)
  float lg[G][HS];                 // logits
  int note[G], last[G];
  int slot[2][GEN_MAXL][G];        // ring slot t mod d of every block and stream, for the current / the next step
  // dynamic tail: float old[N][G][XS], the dilated taps of this step (cp.async landing zone, read as MMA operands)
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(256, 1)
gen_steps_bf16_kernel(FastGenParams p, char* __restrict__ state, const int64_t* __restrict__ first_note,
                      const float* __restrict__ uniforms, int64_t* __restrict__ out, float* __restrict__ logits_out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  GenSmem& S = *reinterpret_cast<GenSmem*>(smem_raw);
  const int N = p.n_layers;
  float* const old = reinterpret_cast<float*>(smem_raw + ((sizeof(GenSmem) + 15) & ~size_t(15)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n8 = lane >> 2, q = lane & 3;          // fragment coordinates: B column / C row = n8, C columns = 2q, 2q+1
  const int st0 = blockIdx.x * G;
  const int n_act = min(G, p.n_streams - st0);
  auto stream_ptr = [&](int s) { return state + (int64_t)min(st0 + s, p.n_streams - 1) * p.state_stride; };
  if (tid < G) {
    S.note[tid] = (int)first_note[min(st0 + tid, p.n_streams - 1)];
    S.last[tid] = (int)reinterpret_cast<const int64_t*>(stream_ptr(tid))[1];
  }
  for (int e = tid; e < N * G; e += 256) {          // the only 64-bit divisions: afterwards the slots advance by one per step
    const int i = e / G, s = e % G;
    S.slot[0][i][s] = (int)(reinterpret_cast<const int64_t*>(stream_ptr(s))[0] % p.dil[i]);
  }
  __syncthreads();
  // rings of the two streams whose results this thread holds in its accumulator fragments (columns 2q, 2q+1)
  float* const ring0 = reinterpret_cast<float*>(stream_ptr(2 * q) + 16);
  float* const ring1 = reinterpret_cast<float*>(stream_ptr(2 * q + 1) + 16);

  // ---- A fragments in registers, two parity sets (block i uses set i & 1).  fg: the warp's f|g m-tile (8 k-tiles).  w2: warps 0-3
  // hold their dense m-tile in [0,4) and skip m-tile `warp` in [4,8); warps 4-7 hold skip m-tiles 4 + 3 (warp - 4) + {0,1,2} in
  // [0,12).  (Warps 0-3 also run the residual epilogue, so they get one skip tile instead of three.)
  uint4 fgA[8], fgB[8], w2A[12], w2B[12];
  const uint4* const fbase = p.frag + lane;
  auto load_fg = [&](int i, uint4 (&w)[8]) {
    const uint4* base = fbase + (int64_t)i * FRAG_LAYER + FRAG_FG + warp * 8 * 32;
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = base[k * 32];
  };
  auto load_w2 = [&](int i, uint4 (&w)[12]) {
    const uint4* base = fbase + (int64_t)i * FRAG_LAYER;
    if (warp < 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = base[FRAG_D + (warp * 4 + k) * 32];
#pragma unroll
      for (int k = 0; k < 4; ++k) w[4 + k] = base[FRAG_S + (warp * 4 + k) * 32];
    } else {
#pragma unroll
      for (int k = 0; k < 12; ++k) w[k] = base[FRAG_S + ((4 + 3 * (warp - 4)) * 4 + k) * 32];
    }
  };
  // ---- asynchronous fetch of every block's dilated tap (slot t mod d of the per-stream queues) of all streams.  advance = 0: the
  // slots of buffer `cur` as they are; advance = 1: the slots of the NEXT step, which are also written to buffer cur ^ 1.
  auto prefetch_taps = [&](int cur, int advance) {
    for (int e = tid; e < G * N * 16; e += 256) {
      const int c4 = (e & 15) * 4, s = (e >> 4) % G, i = (e >> 4) / G;
      int slot = S.slot[cur][i][s];
      if (advance) {
        slot = slot + 1 == p.dil[i] ? 0 : slot + 1;
        if (c4 == 0) S.slot[cur ^ 1][i][s] = slot;
      }
      const float* src = reinterpret_cast<const float*>(stream_ptr(s) + 16) + ((int64_t)p.ring_off[i] + slot) * 64 + c4;
      __pipeline_memcpy_async(old + (size_t)(e >> 4) * XS + c4, src, 16);
    }
    __pipeline_commit();
  };
  load_fg(0, fgA);
  load_w2(0, w2A);
  if (N > 1) load_fg(1, fgB);
  prefetch_taps(0, 0);

  const uint4* const headA = fbase + (int64_t)N * FRAG_LAYER;

  for (int step = 0; step < p.n_steps; ++step) {
    const int cur = step & 1;
    // ---- causal layer: a gather of two embedding rows (fast_generate.py:111-116)
    for (int e = tid; e < G * 64; e += 256) {
      const int s = e >> 6, r = e & 63;
      float v = p.wc_t[(int64_t)S.last[s] * 64 + r] + p.wc_t[((int64_t)256 + S.note[s]) * 64 + r];
      if (p.has_bias) v += p.bias_c[r];
      S.x[s][r] = v;
      S.xh[s][r] = __float2half_rn(v);
    }
    float sk[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) sk[j][r] = 0.f;
    __pipeline_wait_prior(0);                        // this step's taps
    __syncthreads();
    if (tid < G) S.last[tid] = S.note[tid];          // next step's tap 0 of the causal layer (read again only after later barriers)

    // one residual block (fast_generate.py:118-129).  While block i runs, the dense + skip fragments of block i+1 and (after its
    // f|g MMAs) the f|g fragments of block i+2 are requested from L2.
    auto run_layer = [&](int i, uint4 (&fg)[8], uint4 (&w2)[12], uint4 (&w2_next)[12]) {
      if (i + 1 < N) load_w2(i + 1, w2_next);
      // ---- [f|g] = W0 old + W1 x : warp w owns filter and gate channels 8w..8w+7 for all 8 streams
      const int ch = 8 * warp + n8;
      float c0[4], c1[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const float bf = p.has_bias ? p.bias_fg[i * 128 + ch] : 0.f, bg = p.has_bias ? p.bias_fg[i * 128 + 64 + ch] : 0.f;
        c0[0] = c0[1] = bf;
        c0[2] = c0[3] = bg;
      }
      const float* orow = old + ((size_t)i * G + n8) * XS + 2 * q;
      const __half* xrow = &S.xh[n8][2 * q];
      uint32_t bo[4][2], bx[4][2];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const float2 v0 = *reinterpret_cast<const float2*>(orow + kt * 16), v1 = *reinterpret_cast<const float2*>(orow + kt * 16 + 8);
        bo[kt][0] = pack_h2(v0.x, v0.y);
        bo[kt][1] = pack_h2(v1.x, v1.y);
        bx[kt][0] = *reinterpret_cast<const uint32_t*>(xrow + kt * 16);
        bx[kt][1] = *reinterpret_cast<const uint32_t*>(xrow + kt * 16 + 8);
      }
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        mma_f16(c0, fg[kt], bo[kt][0], bo[kt][1]);
        mma_f16(c1, fg[4 + kt], bx[kt][0], bx[kt][1]);
      }
      if (i + 2 < N) load_fg(i + 2, fg);
      S.zh[2 * q][ch] = __float2half_rn(gate_z(c0[0] + c1[0], c0[2] + c1[2]));
      S.zh[2 * q + 1][ch] = __float2half_rn(gate_z(c0[1] + c1[1], c0[3] + c1[3]));
      __syncthreads();
      // ---- dense (64 -> 64) + residual on warps 0-3 first (it is the critical path), skip (64 -> 256) accumulates in registers
      uint32_t bz[4][2];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const __half* zr = &S.zh[n8][kt * 16 + 2 * q];
        bz[kt][0] = *reinterpret_cast<const uint32_t*>(zr);
        bz[kt][1] = *reinterpret_cast<const uint32_t*>(zr + 8);
      }
      if (warp < 4) {
        const int chd = 16 * warp + n8;
        const int s0 = 2 * q, s1 = 2 * q + 1;
        float dn[4];
        dn[0] = S.x[s0][chd]; dn[1] = S.x[s1][chd]; dn[2] = S.x[s0][chd + 8]; dn[3] = S.x[s1][chd + 8];     // residual
        const float xin[4] = {dn[0], dn[1], dn[2], dn[3]};
        if (p.has_bias) {
          const float b0 = p.bias_d[i * 64 + chd], b1 = p.bias_d[i * 64 + chd + 8];
          dn[0] += b0; dn[1] += b0; dn[2] += b1; dn[3] += b1;
        }
        const int o0 = (p.ring_off[i] + S.slot[cur][i][s0]) * 64 + chd, o1 = (p.ring_off[i] + S.slot[cur][i][s1]) * 64 + chd;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) mma_f16(dn, w2[kt], bz[kt][0], bz[kt][1]);
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) mma_f16(sk[0], w2[4 + kt], bz[kt][0], bz[kt][1]);
        S.x[s0][chd] = dn[0]; S.x[s1][chd] = dn[1]; S.x[s0][chd + 8] = dn[2]; S.x[s1][chd + 8] = dn[3];
        S.xh[s0][chd] = __float2half_rn(dn[0]); S.xh[s1][chd] = __float2half_rn(dn[1]);
        S.xh[s0][chd + 8] = __float2half_rn(dn[2]); S.xh[s1][chd + 8] = __float2half_rn(dn[3]);
        const bool out_push = p.push == WN_PUSH_OUTPUT;                                     // fast_generate.py:128-129
        if (s0 < n_act) { ring0[o0] = out_push ? dn[0] : xin[0]; ring0[o0 + 8] = out_push ? dn[2] : xin[2]; }
        if (s1 < n_act) { ring1[o1] = out_push ? dn[1] : xin[1]; ring1[o1 + 8] = out_push ? dn[3] : xin[3]; }
      } else {
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
          mma_f16(sk[0], w2[kt], bz[kt][0], bz[kt][1]);
          mma_f16(sk[1], w2[4 + kt], bz[kt][0], bz[kt][1]);
          mma_f16(sk[2], w2[8 + kt], bz[kt][0], bz[kt][1]);
        }
      }
      __syncthreads();
    };
    for (int i = 0; i < N; i += 2) {
      run_layer(i, fgA, w2A, w2B);
      if (i + 1 < N) run_layer(i + 1, fgB, w2B, w2A);
    }
    // the taps of the next step (slot (t+1) mod d; for d = 1 that is the vector pushed just now, visible after the barrier)
    if (step + 1 < p.n_steps) prefetch_taps(cur, 1);

    // ---- head: relu(sum skips) -> P1 -> relu -> P2 ; warp w owns rows 32w..32w+31 of P1 / P2 outputs
    uint4 ha[8], hb[8];
    auto head_load = [&](int which, int part, uint4 (&dst)[8]) {               // both m-tiles, k-tiles [4 part, 4 part + 4)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[j * 4 + k] = headA[(int64_t)which * FRAG_HEAD + ((2 * warp + j) * 16 + part * 4 + k) * 32];
    };
    head_load(0, 0, ha);
    head_load(0, 1, hb);
    {   // skip sums -> relu -> fp16, rows of the skip m-tiles this warp accumulated
      const int nt = warp < 4 ? 1 : 3, mt0 = warp < 4 ? warp : 4 + 3 * (warp - 4);
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (j < nt) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = 16 * (mt0 + j) + n8 + 8 * (r >> 1), s = 2 * q + (r & 1);
            S.hh[0][s][row] = __float2half_rn(fmaxf(sk[j][r] + (p.has_bias ? p.bias_skip[row] : 0.f), 0.f));
          }
        }
    }
    __syncthreads();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float c[2][2][4];               // [m-tile][k parity][fragment]
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1);
          const float* bias = which == 0 ? p.bias_p1 : p.bias_p2;
          c[j][0][r] = p.has_bias ? bias[row] : 0.f;
          c[j][1][r] = 0.f;
        }
      const __half* hr = &S.hh[which][n8][2 * q];
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        uint4 (&cur_w)[8] = (part & 1) ? hb : ha;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int kt = part * 4 + k;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(hr + kt * 16), b1 = *reinterpret_cast<const uint32_t*>(hr + kt * 16 + 8);
          mma_f16(c[0][k & 1], cur_w[k], b0, b1);
          mma_f16(c[1][k & 1], cur_w[4 + k], b0, b1);
        }
        // refill the set just consumed: two parts ahead, rolling over into the second matrix
        if (part + 2 < 4) head_load(which, part + 2, cur_w);
        else if (which == 0) head_load(1, part - 2, cur_w);
      }
      if (which == 1 && step + 1 < p.n_steps) {          // the first blocks of the next step
        load_fg(0, fgA);
        load_w2(0, w2A);
        if (N > 1) load_fg(1, fgB);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1), s = 2 * q + (r & 1);
          const float v = c[j][0][r] + c[j][1][r];
          if (which == 0) {
            S.hh[1][s][row] = __float2half_rn(fmaxf(v, 0.f));
          } else {
            S.lg[s][row] = v;
            if (logits_out && s < n_act) logits_out[((int64_t)step * p.n_streams + st0 + s) * 256 + row] = v;
          }
        }
      __syncthreads();
    }
    // ---- pick: greedy topk(1) over the softmax (fast_generate.py:138-140) or inverse CDF (extension); one warp per stream
    {
      const int s = warp;
      const float* lg = S.lg[s];
      float v[8];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = lg[lane * 8 + j];
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
          for (int k = 0; k < 256; ++k) sum += expf(lg[k] - mx);
          const float inv = 1.f / sum;
          float total_p = 0.f;
          for (int k = 0; k < 256; ++k) total_p += expf(lg[k] - mx) * inv;
          const float thr = uniforms[(int64_t)step * p.n_streams + min(st0 + s, p.n_streams - 1)] * total_p;
          float cdf = 0.f;
          for (int k = 0; k < 256; ++k) {
            cdf += expf(lg[k] - mx) * inv;
            if (cdf > thr) { pick = k; break; }
          }
        }
        pick = __shfl_sync(0xffffffffu, pick, 0);
      }
      if (lane == 0) {
        S.note[s] = pick;
        if (s < n_act) out[(int64_t)step * p.n_streams + st0 + s] = pick;
      }
    }
    __syncthreads();
  }
  __pipeline_wait_prior(0);
  if (tid < n_act) {
    reinterpret_cast<int64_t*>(stream_ptr(tid))[0] += p.n_steps;
    reinterpret_cast<int64_t*>(stream_ptr(tid))[1] = S.last[tid];
  }
}


// ================================================================================================ weights-stationary pipeline
// gen_pipe_kernel: a thread-block cluster holds the whole model for the whole launch and serves groups of G = 8 streams (the N
// dimension of m16n8k16).  Block CTAs keep the A fragments of their residual blocks resident, the last CTA ("head") keeps
// post_process_1 and half of post_process_2 in registers and the rest in shared memory: nothing is re-read from L2 per step but
// the ring state.  Two geometries (template BPC, chosen by fast_gen_steps):
//   BPC = 4  CTA 0 and the last block CTA have 2 blocks with every fragment in registers; the CTAs in between have 4 blocks with
//            W1 (current-sample tap of [f|g]) and the dense convolution in registers - what the token waits for - and W0 (old tap)
//            and the skip convolution, which run ahead of / behind the token, in shared memory.  10 CTAs for 30 blocks, ONE group
//            per cluster, as many clusters as fit on the GPU (up to 16 = 128 streams).
//   BPC = 2  every CTA 2 blocks, everything in registers, 16 CTAs for 30 blocks, up to NG = 8 groups per cluster (576 streams).
// A group's state is a token that hops CTA -> CTA through distributed shared memory.  Its critical path per block is
//   LDS x fragments -> 4 independent MMAs (W1 x, on top of the precomputed W0 old + bias + conditioning) -> gate -> STS z, barrier ->
//   (4 dense warps) LDS z fragments -> 4 independent MMAs -> + residual (kept in registers from block to block) -> STS x, barrier
// and after a CTA's last block the 4 dense warps send the token straight from their registers with st.async (data and mbarrier
// complete_tx travel together): 1 KB of fp16 B fragments (the accumulator fragments transposed by movmatrix: exactly what the next
// CTA's MMAs load) on one mbarrier, then 2 KB of fp32 accumulator fragments (the residual) on a second one that the consumer probes
// a phase early.  Everything else is off that path: W0.old of a group's next token is computed as soon as its taps (cp.async,
// requested one ring period ahead) have landed; skip MMAs, queue pushes and the running skip sums (an 8 KB bulk copy behind the
// token) come after the token has left; barriers are re-armed in the tail.  The last block CTA does its skip MMAs first and sends
// relu(sum + bias) as post_process_1's fp16 B fragments; the head finds the greedy pick in registers (shuffles, then one exchange
// through 512 B of shared memory) and st.asyncs it to CTA 0, which gathers the causal layer's two embedding rows from its
// shared-memory copy of the table.  Every group owns one slot per CTA and carries exactly one token around the ring, so there is no
// back-pressure and no barrier between groups.  The CTA's role (first / in between / short in between / last) is a compile-time
// parameter of the loop body; waits use CTA-scope acquires (a cluster-scope acquire compiles to CCTL.IVALL per token).
// Measured history and the micro-benchmarks behind these choices: profiles/r2_summary.md (32.1 -> 9.8 us per step).
constexpr int NG = 8;
// timing experiments (WN_TS=1): clock64 stamps of CTA 1 of cluster 0, group 0, 16 per step (wn_debug_ts with n < 0 reads them)
__device__ long long g_gen_ts[3 * 16 * 64];      // [0,1024): clock64 stamps of CTA 1; [1024,2048): globaltimer at token arrival, per rank; [2048,3072): clock64 stamps of the head
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define HEAD_TS(k)                                                                                  \
  do {                                                                                              \
    if (TRACE && cid == 0 && g == 0 && tid == 0 && step < 64) g_gen_ts[2048 + step * 16 + (k)] = clock64(); \
  } while (0)
#define GEN_TS(k)                                                                                   \
  do {                                                                                              \
    if (TRACE && rank == p.trace && cid == 0 && g == 0 && tid == 0 && step < 64) g_gen_ts[step * 16 + (k)] = clock64(); \
  } while (0)
namespace pipe {
constexpr uint32_t XR_BYTES = 128 * 16, XF_BYTES = 4 * 32 * 8;      // one group's x: fp32 residual (the dense warps' accumulator fragments),
                                                                      // fp16 B fragments [k-tile 4][lane 32] of every warp's MMAs
constexpr uint32_t X_BYTES = G * XS * 4, SK_BYTES = G * HS * 4;      // a group's taps of one block (fp32 rows), its skip sums
constexpr uint32_t XH_BYTES = G * XH * 2;                            // 8 streams x 64 channels of fp16, rows padded to 72 (conflict-free fragment loads)
constexpr uint32_t HF_BYTES = 16 * 32 * 8;                           // a group's relu(skip sums), fp16 B fragments [k-tile 16][lane 32]
constexpr uint32_t WS_BLOCK = 12 * 256 * 16;                         // one block's off-critical-path A fragments in shared memory: per thread
                                                                      // 4 of W0 (the old tap) and 8 of the skip convolution, [fragment 12][thread 256]
// Shared-memory map for NGv group slots per CTA and up to NBv blocks per CTA (CTA 0 always has two).  The slots other CTAs write into
// (notes, x, skip sums, the head's fragments) sit at the same offsets in every role.
template <int NGv, int NBv>
struct PipeL {
  static constexpr uint32_t OFF_BAR = 0;                                   // xhfull[NGv], xrfull[NGv], skfull[NGv], notefull[NGv]
  static constexpr uint32_t OFF_NOTE = 256;                                // int note[NGv][G], last[NGv][G]
  static constexpr uint32_t OFF_SLOT = OFF_NOTE + 3 * NGv * G * 4;         // (+ int tb[NGv][G]: step counters at launch) int slot[2][NGv][NBv][G]
  static constexpr uint32_t OFF_ZF = OFF_SLOT + 2 * NGv * NBv * G * 4;         // __half zh[NBv][G][XH]: gated activations of the CTA's blocks
  static constexpr uint32_t OFF_XL = OFF_ZF + NBv * XH_BYTES;              // __half xl[G][XH]: a block's output (CTA 0: also the embedding)
  static constexpr uint32_t OFF_SKIN = (OFF_XL + XH_BYTES + 127) & ~127u;  // float skin[NGv][G][HS]  (same offset in every role but CTA 0)
  static constexpr uint32_t OFF_XIN = OFF_SKIN + NGv * SK_BYTES;           // block CTAs: float4 xr[NGv][128]
  static constexpr uint32_t OFF_XHIN = OFF_XIN + NGv * XR_BYTES;           //             uint2 xf[NGv][4][32]
  static constexpr uint32_t OFF_TAPS = OFF_XHIN + NGv * XF_BYTES;          //             float taps[NGv][NBv][G][XS]
  static constexpr uint32_t OFF_WS = OFF_TAPS + NGv * NBv * X_BYTES;       //             uint4 ws[NBv][12][256] (NBv > 2 only)
  static constexpr uint32_t OFF_SKIN2 = OFF_WS;                            // the last block CTA (two blocks, no ws): float skin2[NGv][G][HS]
  static constexpr uint32_t BLOCK_BYTES = OFF_WS + (NBv > 2 ? NBv * WS_BLOCK : 0);
  static constexpr uint32_t OFF_WC = OFF_SKIN;                             // CTA 0 (nothing arrives but notes): float wc[2][256][64]
  static constexpr uint32_t OFF_TAPS0 = OFF_WC + 2 * 256 * 64 * 4;         //        its taps [NGv][2][G][XS]
  static constexpr uint32_t OFF_STG0 = OFF_TAPS0 + NGv * 2 * X_BYTES;      //        two staging sets of outgoing skip sums
  static constexpr uint32_t STG_BYTES = SK_BYTES;
  static constexpr uint32_t CTA0_BYTES = OFF_STG0 + 2 * STG_BYTES;
  static constexpr uint32_t OFF_HF = OFF_SKIN;                             // head: uint2 hf[NGv][16][32] (what the last block CTA sends)
  static constexpr uint32_t OFF_P2 = OFF_HF + NGv * HF_BYTES;              //       uint4 p2[FRAG_HEAD]
  static constexpr uint32_t OFF_HH = OFF_P2 + FRAG_HEAD * 16;              //       relu(post_process_1) fragments, greedy candidates
  static constexpr uint32_t OFF_LG = OFF_HH + 2 * G * HH * 2;              //       float lg[G][HS]
  static constexpr uint32_t OFF_CH = OFF_LG + G * HS * 4;                  //       float condh[8][256]: a step's post_process_1 conditioning
  static constexpr uint32_t HEAD_BYTES = OFF_CH + 8 * 256 * 4;
  static constexpr uint32_t TOTAL =
      HEAD_BYTES > CTA0_BYTES ? (HEAD_BYTES > BLOCK_BYTES ? HEAD_BYTES : BLOCK_BYTES) : (CTA0_BYTES > BLOCK_BYTES ? CTA0_BYTES : BLOCK_BYTES);
  static_assert(TOTAL <= 227 * 1024, "pipeline generation kernel: shared memory");
  static_assert(4 * NGv * 8 + 8 <= 256 || NBv == 2, "barriers");
  static_assert(NBv == 2 || NGv == 1, "the second skip slot is for one group");
  static_assert(SK_BYTES % 16 == 0 && OFF_SKIN % 16 == 0 && OFF_STG0 % 16 == 0 && OFF_XIN % 16 == 0 && OFF_XHIN % 16 == 0 && OFF_WS % 16 == 0, "16-byte units");
};
constexpr int pipe_groups(int bpc) { return bpc == 2 ? 8 : 1; }      // group slots per cluster

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// local shared memory -> another CTA's shared memory; the bytes complete on THAT CTA's mbarrier
__device__ __forceinline__ void bulk_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(src_cta),
               "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// one thread's 16 / 8 bytes -> another CTA's shared memory; the bytes complete on THAT CTA's mbarrier (data and signal travel together:
// measured 310 cycles per hand-off of 1 KB against 530 for st.shared + fence.proxy.async + barrier + cp.async.bulk of the 3 KB token)
__device__ __forceinline__ void st_async_f4(uint32_t dst_cluster, const float (&v)[4], uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(dst_cluster), "f"(v[0]),
               "f"(v[1]), "f"(v[2]), "f"(v[3]), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_async_u1(uint32_t dst_cluster, uint32_t a, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst_cluster), "r"(a), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_async_u2(uint32_t dst_cluster, uint32_t a, uint32_t b, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1,%2}, [%3];" ::"r"(dst_cluster), "r"(a), "r"(b),
               "r"(bar_cluster) : "memory");
}
// 8x8 fp16 transpose inside a warp: an accumulator fragment (row = channel, two streams per thread) becomes a B fragment (row = stream's
// column, two channels per thread)
__device__ __forceinline__ uint32_t movm_t(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// wait for completion `parity` of a local mbarrier whose bytes come from another CTA of the cluster (st.async / bulk copies into THIS
// CTA's shared memory: the transaction count orders them before the phase flips, so the default CTA-scope acquire is enough - a
// cluster-scope acquire compiles to CCTL.IVALL, an L1 invalidation on every token).  Bounded: a token that never arrives (a bug) traps
// after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void wait_token(uint64_t* bar, uint32_t parity) {
#ifdef WN_GEN_WAIT_LANE0
  if ((threadIdx.x & 31) != 0) { __syncwarp(); return; }
#endif
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
#ifdef WN_GEN_WAIT_LANE0
  __syncwarp();
#endif
}
// float -> unsigned key with the same order (-0 < +0; NaNs sort to the ends)
__device__ __forceinline__ uint32_t ordered_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint32_t try_wait_once(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
template <int N_>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
}  // namespace pipe

// HAS_BIAS / PUSH_OUT are compile-time: every block of the ring is a chain of ~400 dependent-latency instructions per warp, and the
// run-time tests (and the predicated bias loads behind them) were a measurable part of it
// BPC = blocks per CTA behind CTA 0 (which always has two).  2: every fragment in registers, 8 group slots per cluster.  4: only what the
// token waits for stays in registers (W1 of [f|g], the dense convolution); W0 and the skip convolution, which run behind / ahead of
// the token, are read from shared memory - 8 block CTAs instead of 15, seven hops less per step, one group per cluster.
template <bool HAS_BIAS, bool PUSH_OUT, bool TRACE, int BPC, bool COND = false>
__global__ void __launch_bounds__(256, 1)
gen_pipe_kernel(FastGenParams p, char* __restrict__ state, const int64_t* __restrict__ first_note, const float* __restrict__ uniforms,
                int64_t* __restrict__ out, float* __restrict__ logits_out) {
  using namespace pipe;
  constexpr int NGV = pipe_groups(BPC);
  using L = PipeL<NGV, BPC>;
  constexpr uint32_t OFF_BAR = L::OFF_BAR, OFF_NOTE = L::OFF_NOTE, OFF_SLOT = L::OFF_SLOT, OFF_ZF = L::OFF_ZF, OFF_XL = L::OFF_XL,
                     OFF_SKIN = L::OFF_SKIN, OFF_XIN = L::OFF_XIN, OFF_XHIN = L::OFF_XHIN, OFF_TAPS = L::OFF_TAPS, OFF_WS = L::OFF_WS,
                     OFF_WC = L::OFF_WC, OFF_TAPS0 = L::OFF_TAPS0, OFF_STG0 = L::OFF_STG0, STG_BYTES = L::STG_BYTES, OFF_HF = L::OFF_HF,
                     OFF_P2 = L::OFF_P2, OFF_HH = L::OFF_HH, OFF_LG = L::OFF_LG, OFF_SKIN2 = L::OFF_SKIN2, OFF_CH = L::OFF_CH;
  extern __shared__ __align__(128) uint8_t sm[];
  const int N = p.n_layers;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n8 = lane >> 2, q = lane & 3;
  const int CS = (int)cluster_size(), rank = (int)cluster_rank(), cid = (int)blockIdx.x / CS;
  const int groups_total = (p.n_streams + G - 1) / G;
  const int g0 = cid * p.gpc, ng = min(p.gpc, groups_total - g0);
  const bool is_head = rank == CS - 1;
  uint64_t* xhfull = reinterpret_cast<uint64_t*>(sm + OFF_BAR);    // a group's x has arrived, fp16 fragments (what the first MMAs need): tx bytes
  uint64_t* xrfull = xhfull + NGV;                                  // its fp32 residual (needed one exchange phase later)
  uint64_t* skfull = xhfull + 2 * NGV;                              // its running skip sums (a bulk copy, behind the token)
  uint64_t* notefull = xhfull + 3 * NGV;                            // CTA 0: the head's picks of a group
  int (*note)[G] = reinterpret_cast<int (*)[G]>(sm + OFF_NOTE);
  int (*last)[G] = reinterpret_cast<int (*)[G]>(sm + OFF_NOTE + NGV * G * 4);
  int (*tb)[G] = reinterpret_cast<int (*)[G]>(sm + OFF_NOTE + 2 * NGV * G * 4);      // every stream's step counter at launch (conditioning)
  int (*slot)[NGV][BPC][G] = reinterpret_cast<int (*)[NGV][BPC][G]>(sm + OFF_SLOT);      // [step parity]: ring slot t mod d of every block and stream
  float (*skin)[G][HS] = reinterpret_cast<float (*)[G][HS]>(sm + OFF_SKIN);
  // BPC > 2, blocks that do not divide: the CTA before the last has at most two blocks ("small", register-resident).  The chain of
  // running skip sums then bypasses it - the CTA before it sends to the last block CTA directly, and the small CTA sends its own
  // contribution there too (second slot) - because a CTA with shared-memory skip weights emits its sums too late for a short hop.
  const int n_last_b = N > 2 ? min(2, N - 2) : 0, n_mid_b = max(0, N - 2 - n_last_b);
  const bool has_small = BPC > 2 && n_mid_b > 0 && (n_mid_b % BPC == 1 || n_mid_b % BPC == 2);
  uint64_t* skfull2 = xhfull + 4 * NGV;                             // the last block CTA: the small CTA's own skip sums (has_small)
  float (*skin2)[HS] = reinterpret_cast<float (*)[HS]>(sm + OFF_SKIN2);
  const uint32_t sm_base = smem_u32(sm);
  if (tid == 0) {
    for (int g = 0; g < NGV; ++g) {
      mbar_init(&xhfull[g], 1);       // the consumer's own arming arrival; the data arrive as tx bytes
      mbar_init(&xrfull[g], 1);
      mbar_init(&skfull[g], 1);
      if (BPC > 2) mbar_init(&skfull2[g], 1);
      mbar_init(&notefull[g], 1);     // CTA 0: the 8 picks of the head arrive as 32 tx bytes
    }
    fence_barrier_init();
    if (rank == 0)
      for (int g = 0; g < NGV; ++g) mbar_expect_tx(&notefull[g], G * 4);
    if (rank > 0)
      for (int g = 0; g < NGV; ++g) {      // armed for step 0
        if (!is_head) {
          mbar_expect_tx(&xhfull[g], XF_BYTES);
          mbar_expect_tx(&xrfull[g], XR_BYTES);
        }
        if (!(has_small && rank == CS - 3)) mbar_expect_tx(&skfull[g], is_head ? HF_BYTES : SK_BYTES);      // (nothing is sent to the small CTA)
        if (has_small && rank == CS - 2) mbar_expect_tx(&skfull2[g], SK_BYTES);
      }
  }
  auto stream_of = [&](int g, int s) { return min((g0 + g) * G + s, p.n_streams - 1); };
  auto sptr = [&](int g, int s) { return state + (int64_t)stream_of(g, s) * p.state_stride; };
  auto n_act_of = [&](int g) { return min(G, p.n_streams - (g0 + g) * G); };
  if (COND)
    for (int e = tid; e < max(ng, 0) * G; e += 256) tb[e / G][e % G] = (int)reinterpret_cast<const int64_t*>(sptr(e / G, e % G))[0];
  cluster_sync_all();
  if (ng <= 0) {      // (cannot happen with the launch geometry of fast_gen_steps; all CTAs of the cluster agree)
    cluster_sync_all();
    return;
  }

  if (!is_head) {
    // ============================================================ block CTA: blocks l0, l0 + 1
    // The CTA's role (first: gathers the embedding, nothing arrives but notes; to_head: its skip sums go to the head) is a compile-time
    // constant of the body: the role tests sat on the token's critical path as taken branches.
    auto block_cta = [&](auto first_c, auto to_head_c, auto small_c) {
    constexpr bool first = decltype(first_c)::value, to_head = decltype(to_head_c)::value, small = decltype(small_c)::value;
    // CTA 0 and the last block CTA have two blocks with every fragment in registers (CTA 0's shared memory holds the embedding table;
    // the last CTA's skip MMAs are what the head waits for).  The CTAs in between have up to BPC.
    // A CTA in between with at most two blocks (the one before the last, when the blocks do not divide) keeps them in registers too:
    // its skip sums then reach the last CTA before that one needs them.
    constexpr int NB = (first || to_head || small) ? 2 : BPC;           // this role's blocks (at most)
    constexpr bool SMEMW = !first && !to_head && !small && BPC > 2;      // W0 and skip fragments in shared memory
    const int n_last = first ? 0 : min(2, N - 2);              // (a model of one or two blocks has CTA 0 only)
    const int l0 = first ? 0 : to_head ? N - n_last : 2 + BPC * (rank - 1);
    const int nl = first ? min(2, N) : to_head ? n_last : min(BPC, N - n_last - l0);
    // exchanges inside the CTA go through [stream][channel] rows of fp16: the producers scatter halves, the consumers read B fragments
    // as 32-bit words (the transposing movmatrix costs 25 cycles of latency: only the token that leaves the CTA pays it)
    __half (*zh)[G][XH] = reinterpret_cast<__half (*)[G][XH]>(sm + OFF_ZF);      // [block][stream][channel]
    __half (*xl)[XH] = reinterpret_cast<__half (*)[XH]>(sm + OFF_XL);            // [stream][channel]
    float* taps = reinterpret_cast<float*>(sm + (first ? OFF_TAPS0 : OFF_TAPS));      // [NGV][NB][G][XS]
    float* wc = reinterpret_cast<float*>(sm + OFF_WC);                                 // CTA 0
    uint4* const wsm = reinterpret_cast<uint4*>(sm + OFF_WS) + tid;                    // SMEMW: [block][fragment 12][thread]
    // ---- resident A fragments (same per-warp ownership as gen_steps_bf16_kernel): fgw = [f|g] rows 8w..8w+7 (k-tiles 0..3 of the old
    // tap W0, 4..7 of the current sample W1); w2w = dense m-tile w (warps 0-3), then skip m-tiles 2w, 2w+1
    uint4 fgw[NB][8], w2w[NB][12];
    {
      const uint4* const fbase = p.frag + lane;
#pragma unroll
      for (int li = 0; li < NB; ++li) {
        const int i = min(l0 + li, N - 1);
        const uint4* base = fbase + (int64_t)i * FRAG_LAYER;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 v = base[FRAG_FG + (warp * 8 + k) * 32];
          if (SMEMW) wsm[(li * 12 + k) * 256] = v;
          else fgw[li][k] = v;
        }
#pragma unroll
        for (int k = 4; k < 8; ++k) fgw[li][k] = base[FRAG_FG + (warp * 8 + k) * 32];
#pragma unroll
        for (int k = 0; k < 4; ++k) w2w[li][k] = warp < 4 ? base[FRAG_D + (warp * 4 + k) * 32] : make_uint4(0u, 0u, 0u, 0u);      // dense: warps 0-3
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint4 v = base[FRAG_S + (2 * warp * 4 + k) * 32];      // skip: m-tiles 2w, 2w+1
          if (SMEMW) wsm[(li * 12 + 4 + k) * 256] = v;
          else w2w[li][4 + k] = v;
        }
      }
    }
    // (every thread reads back only the shared-memory fragments it wrote itself)
    auto w0_frag = [&](int li, int k) -> uint4 { return SMEMW ? wsm[(li * 12 + k) * 256] : fgw[li][k]; };
    auto sk_frag = [&](int li, int k) -> uint4 { return SMEMW ? wsm[(li * 12 + 4 + k) * 256] : w2w[li][4 + k]; };
    for (int e = tid; e < ng * nl * G; e += 256) {
      const int s = e % G, li = (e / G) % nl, g = e / (G * nl);
      slot[0][g][li][s] = (int)(reinterpret_cast<const int64_t*>(sptr(g, s))[0] % p.dil[l0 + li]);
    }
    if (first) {
      for (int e = tid; e < ng * G; e += 256) {
        const int g = e / G, s = e % G;
        note[g][s] = (int)first_note[stream_of(g, s)];
        last[g][s] = (int)reinterpret_cast<const int64_t*>(sptr(g, s))[1];
      }
      for (int e = tid; e < 2 * 256 * 64 / 4; e += 256) reinterpret_cast<float4*>(wc)[e] = reinterpret_cast<const float4*>(p.wc_t)[e];
    }
    __syncthreads();
    // taps of (group g, its next use): slot as stored for step parity par (advance = 0) or slot + 1 (advance = 1, stored for the other
    // parity: the 16 lanes that read a slot and the one that advances it never touch the same word)
    auto prefetch_taps = [&](int g, int advance, int par) {
      for (int e = tid; e < nl * G * 16; e += 256) {
        const int c4 = (e & 15) * 4, s = (e >> 4) % G, li = (e >> 4) / G;
        int sl = slot[par][g][li][s];
        if (advance) {
          sl = sl + 1 == p.dil[l0 + li] ? 0 : sl + 1;
          if (c4 == 0) slot[par ^ 1][g][li][s] = sl;
        }
        const float* src = reinterpret_cast<const float*>(sptr(g, s) + 16) + ((int64_t)p.ring_off[l0 + li] + sl) * 64 + c4;
        __pipeline_memcpy_async(taps + ((size_t)(g * NB + li) * G + s) * XS + c4, src, 16);
      }
      __pipeline_commit();
    };
    for (int g = 0; g < ng; ++g) prefetch_taps(g, 0, 0);
    for (int g = ng; g < NG; ++g) __pipeline_commit();      // always NG commit groups per ring period: the wait below is a constant
    // remote addresses of the next CTA's slots (same offsets there)
    const uint32_t r_xr = map_to(sm_base + OFF_XIN, rank + 1), r_xf = map_to(sm_base + OFF_XHIN, rank + 1);
    const int sk_dst = rank + ((has_small && rank == CS - 4) ? 2 : 1);      // where this CTA's running skip sums go
    const uint32_t r_skin = small ? map_to(sm_base + OFF_SKIN2, rank + 1) : map_to(sm_base + OFF_SKIN, sk_dst);
    const uint32_t r_xhfull = map_to(smem_u32(xhfull), rank + 1), r_xrfull = map_to(smem_u32(xrfull), rank + 1);
    const uint32_t r_skfull = small ? map_to(smem_u32(skfull2), rank + 1) : map_to(smem_u32(skfull), to_head ? rank + 1 : sk_dst);
    const uint32_t r_hf = map_to(sm_base + OFF_HF, rank + 1);      // (the last block CTA: the head's fragment slots)
    // skip MMAs of this CTA's blocks: warp w owns skip channels 32w..32w+31 (m-tiles 2w, 2w+1)
    auto skip_block = [&](float (&sk)[2][4], int li) {
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const __half* zr = &zh[li][n8][kt * 16 + 2 * q];
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(zr), b1 = *reinterpret_cast<const uint32_t*>(zr + 8);
        mma_f16(sk[0], sk_frag(li, kt), b0, b1);
        mma_f16(sk[1], sk_frag(li, 4 + kt), b0, b1);
      }
    };
    auto skip_mmas = [&](float (&sk)[2][4]) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) sk[j][r] = 0.f;
#pragma unroll
      for (int li = 0; li < NB; ++li)
        if (li < nl) skip_block(sk, li);
    };

    // W0 . old of both blocks of group gg (the taps are known one ring period ahead): computed while the group's token is still
    // on its way, so that the token's critical path is four independent MMAs per block instead of four chains of two
    float pre[NB][4];
    auto compute_pre = [&](int gg, int step_of) {      // step_of: the step the token of group gg will belong to
#pragma unroll
      for (int li = 0; li < NB; ++li) {
        if (li < nl) {
          const int ch = 8 * warp + n8;
          const float bf = HAS_BIAS ? p.bias_fg[(l0 + li) * 128 + ch] : 0.f, bg = HAS_BIAS ? p.bias_fg[(l0 + li) * 128 + 64 + ch] : 0.f;
          float c[4] = {bf, bf, bg, bg}, e[4] = {0.f, 0.f, 0.f, 0.f};
          if (COND) {      // this block's conditioning vector for the time step being consumed (gen.cu, same indexing)
            const int i = l0 + li, so = p.s_out[i];
            const int fi = p.cond_gate_first ? 64 + ch : ch, gi = p.cond_gate_first ? ch : 64 + ch;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              const int s_ = 2 * q + b;
              const int tau = p.rf + tb[gg][s_] + step_of;
              const int f = gen_cond_frame(tau - so, p.cond_total - so, p.cond_frames);
              const float* cv = p.cond_fg + (((int64_t)stream_of(gg, s_) * p.cond_frames + f) * N + i) * 128;
              c[b] += cv[fi];
              c[2 + b] += cv[gi];
            }
          }
          const float* orow = taps + ((size_t)(gg * NB + li) * G + n8) * XS + 2 * q;
          uint32_t bo[4][2];
#pragma unroll
          for (int kt = 0; kt < 4; ++kt) {
            const float2 v0 = *reinterpret_cast<const float2*>(orow + kt * 16), v1 = *reinterpret_cast<const float2*>(orow + kt * 16 + 8);
            bo[kt][0] = pack_h2(v0.x, v0.y);
            bo[kt][1] = pack_h2(v1.x, v1.y);
          }
          mma_f16(c, w0_frag(li, 0), bo[0][0], bo[0][1]);
          mma_f16(e, w0_frag(li, 1), bo[1][0], bo[1][1]);
          mma_f16(c, w0_frag(li, 2), bo[2][0], bo[2][1]);
          mma_f16(e, w0_frag(li, 3), bo[3][0], bo[3][1]);
#pragma unroll
          for (int r = 0; r < 4; ++r) pre[li][r] = c[r] + e[r];
        }
      }
    };
    cp_wait<NG - 1>();      // group 0's taps
    __syncthreads();
    compute_pre(0, 0);

    for (int step = 0; step < p.n_steps; ++step) {
      for (int g = 0; g < ng; ++g) {
        const int n_act = n_act_of(g);
        const int chd = 16 * (warp & 3) + n8, s0 = 2 * q, s1 = 2 * q + 1;      // dense warps: accumulator fragment = channels chd, chd + 8
        float (*sko)[HS];       // where this CTA's outgoing skip sums are staged
        float dn[4] = {0.f, 0.f, 0.f, 0.f};      // dense warps: the residual stream, fp32, in accumulator-fragment layout
        if (first) {
          // CTA 0: the causal layer is two embedding rows (fast_generate.py:111-116), gathered by the dense warps straight into their
          // fragments.  Outgoing skip sums alternate between two staging sets; a set is rewritten two groups later.
          sko = reinterpret_cast<float (*)[HS]>(sm + OFF_STG0 + ((step * ng + g) & 1) * STG_BYTES);
          if (tid == 0) bulk_wait_read_1();      // the bulk copy of the iteration that used this staging set has been read
          if (step > 0) {
            wait_token(&notefull[g], (step - 1) & 1);
          }
          if (warp < 4) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int s_ = 2 * q + (r & 1), c_ = chd + 8 * (r >> 1);
              dn[r] = wc[last[g][s_] * 64 + c_] + wc[(256 + note[g][s_]) * 64 + c_] + (HAS_BIAS ? p.bias_c[c_] : 0.f);
            }
            xl[s0][chd] = __float2half_rn(dn[0]); xl[s1][chd] = __float2half_rn(dn[1]);
            xl[s0][chd + 8] = __float2half_rn(dn[2]); xl[s1][chd + 8] = __float2half_rn(dn[3]);
          }
          __syncthreads();      // (also tid 0's wait on the staging set before anybody writes it)
          if (tid < G) last[g][tid] = note[g][tid];
        } else {
          GEN_TS(0);
          wait_token(&xhfull[g], step & 1);      // every thread waits for itself: no barrier between the token and its first use
          GEN_TS(1);
          if (TRACE && g == 0 && tid == 0 && step < 64 && cid == 0) g_gen_ts[1024 + step * 16 + rank] = (long long)global_ns();
          sko = skin[g];
        }
        float pv[NB][4];      // what the blocks push into their queues (dense warps)
        float sk[2][4];       // (the last block CTA: skip sums, accumulated block by block)
#pragma unroll
        for (int li = 0; li < NB; ++li) {
          if (li < nl) {
            const int i = l0 + li;
            const bool last_block = li == nl - 1;
            // ---- [f|g] = (W0 old, precomputed) + W1 x : warp w owns filter and gate channels 8w..8w+7 for all 8 streams
            uint2 bx[4];      // this block's input as B fragments: the token's own layout, or rows of the local exchange buffer
            if (li == 0 && !first) {
#pragma unroll
              for (int kt = 0; kt < 4; ++kt) bx[kt] = reinterpret_cast<const uint2*>(sm + OFF_XHIN)[(g * 4 + kt) * 32 + lane];
            } else {
              const __half* xrow = &xl[n8][2 * q];
#pragma unroll
              for (int kt = 0; kt < 4; ++kt)
                bx[kt] = make_uint2(*reinterpret_cast<const uint32_t*>(xrow + kt * 16), *reinterpret_cast<const uint32_t*>(xrow + kt * 16 + 8));
            }
            float c0[4] = {pre[li][0], pre[li][1], pre[li][2], pre[li][3]}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f},
                  c3[4] = {0.f, 0.f, 0.f, 0.f};
            mma_f16(c0, fgw[li][4], bx[0].x, bx[0].y);
            mma_f16(c1, fgw[li][5], bx[1].x, bx[1].y);
            mma_f16(c2, fgw[li][6], bx[2].x, bx[2].y);
            mma_f16(c3, fgw[li][7], bx[3].x, bx[3].y);
            uint32_t res_ok = 1;
            if (li == 0 && !first && warp < 4) res_ok = try_wait_once(&xrfull[g], step & 1);      // (the answer is read behind the barrier below)
            uint32_t sk_ok = 1;
            if (to_head) {
              // the last block CTA: the previous block's skip MMAs run in the shadow of this block's gate; the upstream sums are probed
              if (li == 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                  for (int r = 0; r < 4; ++r) sk[j][r] = 0.f;
              } else {
                skip_block(sk, li - 1);
              }
              if (last_block && !first) sk_ok = try_wait_once(&skfull[g], step & 1);
            }
            GEN_TS(2 + 4 * (li ? 1 : 0));
            {
              const float z0 = gate_z((c0[0] + c1[0]) + (c2[0] + c3[0]), (c0[2] + c1[2]) + (c2[2] + c3[2]));
              const float z1 = gate_z((c0[1] + c1[1]) + (c2[1] + c3[1]), (c0[3] + c1[3]) + (c2[3] + c3[3]));
              zh[li][2 * q][8 * warp + n8] = __float2half_rn(z0);      // warp w: channels 8w..8w+7 of streams 2q, 2q+1
              zh[li][2 * q + 1][8 * warp + n8] = __float2half_rn(z1);
            }
            __syncthreads();
            GEN_TS(3 + 4 * (li ? 1 : 0));
            if (last_block && to_head) {
              // the last block CTA: what the head waits for is the skip sum, so it goes first - relu and the fp16 conversion applied
              // here, sent as the B fragments of post_process_1's MMAs straight from the accumulators
              skip_block(sk, li);
              GEN_TS(10);
              if (!sk_ok) wait_token(&skfull[g], step & 1);
              if (has_small) wait_token(&skfull2[g], step & 1);
              GEN_TS(11);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                float v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1), s_ = 2 * q + (r & 1);
                  v[r] = fmaxf(sk[j][r] + (first ? 0.f : skin[g][s_][row]) + (has_small ? skin2[s_][row] : 0.f) + (HAS_BIAS ? p.bias_skip[row] : 0.f), 0.f);
                }
                st_async_u2(r_hf + (uint32_t)((g * 16 + 2 * warp + j) * 32 + lane) * 8, movm_t(pack_h2(v[0], v[1])), movm_t(pack_h2(v[2], v[3])),
                            r_skfull + g * 8);
              }
              GEN_TS(12);
            }
            if (warp < 4) {
              if (li == 0 && !first) {      // the fp32 residual travels behind the fragments, on its own barrier (probed one phase ago)
                if (!res_ok) wait_token(&xrfull[g], step & 1);
                const uint4 rv = lds128(sm_base + OFF_XIN + (uint32_t)(g * 128 + tid) * 16);
                dn[0] = __uint_as_float(rv.x); dn[1] = __uint_as_float(rv.y); dn[2] = __uint_as_float(rv.z); dn[3] = __uint_as_float(rv.w);
              }
              uint32_t bz[4][2];
#pragma unroll
              for (int kt = 0; kt < 4; ++kt) {
                const __half* zr = &zh[li][n8][kt * 16 + 2 * q];
                bz[kt][0] = *reinterpret_cast<const uint32_t*>(zr);
                bz[kt][1] = *reinterpret_cast<const uint32_t*>(zr + 8);
              }
              float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f}, d3[4] = {0.f, 0.f, 0.f, 0.f};
              if (HAS_BIAS) {
                const float b0 = p.bias_d[i * 64 + chd], b1 = p.bias_d[i * 64 + chd + 8];
                d0[0] = b0; d0[1] = b0; d0[2] = b1; d0[3] = b1;
              }
              mma_f16(d0, w2w[li][0], bz[0][0], bz[0][1]);
              mma_f16(d1, w2w[li][1], bz[1][0], bz[1][1]);
              mma_f16(d2, w2w[li][2], bz[2][0], bz[2][1]);
              mma_f16(d3, w2w[li][3], bz[3][0], bz[3][1]);
              const float xi[4] = {dn[0], dn[1], dn[2], dn[3]};
#pragma unroll
              for (int r = 0; r < 4; ++r) dn[r] = ((d0[r] + d1[r]) + (d2[r] + d3[r])) + dn[r];
              GEN_TS(4 + 4 * (li ? 1 : 0));
#pragma unroll
              for (int r = 0; r < 4; ++r) pv[li][r] = PUSH_OUT ? dn[r] : xi[r];      // fast_generate.py:128-129
              if (!last_block) {
                xl[s0][chd] = __float2half_rn(dn[0]); xl[s1][chd] = __float2half_rn(dn[1]);      // the next block's input
                xl[s0][chd + 8] = __float2half_rn(dn[2]); xl[s1][chd + 8] = __float2half_rn(dn[3]);
              } else if (!to_head) {
                // the token leaves straight from the dense warps' registers: first the fragments the next CTA's MMAs wait for, then the
                // residual.  The skip MMAs, the queue pushes and the skip sums below are off the ring's critical path.
                st_async_u2(r_xf + (uint32_t)(g * 128 + tid) * 8, movm_t(pack_h2(dn[0], dn[1])), movm_t(pack_h2(dn[2], dn[3])), r_xhfull + g * 8);
                st_async_f4(r_xr + (uint32_t)(g * 128 + tid) * 16, dn, r_xrfull + g * 8);
              }
            }
            if (!last_block) {
              __syncthreads();      // block li + 1 reads what the dense warps wrote
            }
            GEN_TS(5 + 4 * (li ? 1 : 0));
          }
        }
        // ================= behind the token: skip MMAs of both blocks, running skip sums, queue pushes, next taps, next W0 . old
        if (!to_head) {
          skip_mmas(sk);
          GEN_TS(10);
          if (!first && !small) wait_token(&skfull[g], step & 1);
          GEN_TS(11);
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1), s = 2 * q + (r & 1);
              sko[s][row] = sk[j][r] + ((first || small) ? 0.f : skin[g][s][row]);      // (in place behind CTA 0)
            }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the bulk copy below reads these generic-proxy writes
        }
        __syncthreads();
        GEN_TS(12);
        if (tid == 0) {
          if (!to_head) {      // the running skip sums -> the next CTA's slot of this group
            bulk_to_peer(r_skin + g * SK_BYTES, smem_u32(sko), SK_BYTES, r_skfull + g * 8);
            bulk_commit();
          }
          // this group's barriers, armed for its next token (here, not behind the waits: nothing of it on the ring's critical path)
          if (first && step > 0) mbar_expect_tx(&notefull[g], G * 4);      // (the last step's picks are awaited before the kernel ends)
          if (!first && step + 1 < p.n_steps) {
            mbar_expect_tx(&xhfull[g], XF_BYTES);
            mbar_expect_tx(&xrfull[g], XR_BYTES);
            if (!small) mbar_expect_tx(&skfull[g], SK_BYTES);
            if (to_head && has_small) mbar_expect_tx(&skfull2[g], SK_BYTES);
          }
        }
        if (warp < 4) {
          float* const ring0 = reinterpret_cast<float*>(sptr(g, s0) + 16);
          float* const ring1 = reinterpret_cast<float*>(sptr(g, s1) + 16);
#pragma unroll
          for (int li = 0; li < NB; ++li)
            if (li < nl) {
              const int i = l0 + li;
              const int o0 = (p.ring_off[i] + slot[step & 1][g][li][s0]) * 64 + chd, o1 = (p.ring_off[i] + slot[step & 1][g][li][s1]) * 64 + chd;
              if (s0 < n_act) { ring0[o0] = pv[li][0]; ring0[o0 + 8] = pv[li][2]; }
              if (s1 < n_act) { ring1[o1] = pv[li][1]; ring1[o1 + 8] = pv[li][3]; }
            }
        }
        __syncthreads();      // the queue pushes above happen before the tap requests below (a queue of dilation 1 is read back at once)
        if (step + 1 < p.n_steps) prefetch_taps(g, 1, step & 1);
        else __pipeline_commit();               // (keeps the commit-group accounting uniform)
        if (g == ng - 1)
          for (int k = ng; k < NG; ++k) __pipeline_commit();
        GEN_TS(13);
        if (g + 1 < ng || step + 1 < p.n_steps) {
          cp_wait<NG - 1>();      // the next group's taps were requested NG commit groups ago
          __syncthreads();
          compute_pre(g + 1 < ng ? g + 1 : 0, g + 1 < ng ? step : step + 1);
        }
        GEN_TS(14);
      }
    }
    cp_wait<0>();
    if (tid == 0) bulk_wait_read_0();
    if (first) {
      for (int g = 0; g < ng; ++g) wait_token(&notefull[g], (p.n_steps - 1) & 1);      // nothing is in flight towards this CTA at exit
      __syncthreads();
      for (int e = tid; e < ng * G; e += 256) {
        const int g = e / G, s = e % G;
        if (s < n_act_of(g)) {
          reinterpret_cast<int64_t*>(sptr(g, s))[0] += p.n_steps;
          reinterpret_cast<int64_t*>(sptr(g, s))[1] = last[g][s];
        }
      }
    }
    };
    if (rank == 0) {
      if (CS == 2) block_cta(std::true_type{}, std::true_type{}, std::false_type{});
      else block_cta(std::true_type{}, std::false_type{}, std::false_type{});
    } else if (rank == CS - 2) {
      block_cta(std::false_type{}, std::true_type{}, std::false_type{});
    } else if (BPC > 2 && N - std::min(2, N - 2) - (2 + BPC * (rank - 1)) <= 2) {
      block_cta(std::false_type{}, std::false_type{}, std::true_type{});
    } else {
      block_cta(std::false_type{}, std::false_type{}, std::false_type{});
    }
  } else {
    // ============================================================ head CTA: relu(sum skips) -> P1 -> relu -> P2 -> pick
    uint4* p2s = reinterpret_cast<uint4*>(sm + OFF_P2);
    float (*lg)[HS] = reinterpret_cast<float (*)[HS]>(sm + OFF_LG);
    const uint4* const headA = p.frag + (int64_t)N * FRAG_LAYER;
    uint4 p1w[2][16];      // post_process_1: this warp's two m-tiles, resident
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 16; ++k) p1w[j][k] = headA[((2 * warp + j) * 16 + k) * 32 + lane];
    uint4 p2w[2][8];       // post_process_2: k-tiles 0..7 of the two m-tiles resident too; k-tiles 8..15 come from shared memory
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) p2w[j][k] = headA[FRAG_HEAD + ((2 * warp + j) * 16 + k) * 32 + lane];
    for (int e = tid; e < FRAG_HEAD; e += 256) p2s[e] = headA[FRAG_HEAD + e];
    __syncthreads();
    const uint32_t r_note = map_to(sm_base + OFF_NOTE, 0), r_notefull = map_to(smem_u32(notefull), 0);
    const uint2* const hf = reinterpret_cast<const uint2*>(sm + OFF_HF);      // [group][k-tile 16][lane]: relu(skip sums), B fragments
    uint2* const h1f = reinterpret_cast<uint2*>(sm + OFF_HH);                  // [k-tile 16][lane]: relu(post_process_1), B fragments
    float2* const cand = reinterpret_cast<float2*>(sm + OFF_HH + 16 * 32 * 8);  // [warp 8][stream 8]: greedy candidates (value, row)
    float* const condh = reinterpret_cast<float*>(sm + OFF_CH);                  // [8][256 threads]: a step's conditioning, each thread's own
    for (int step = 0; step < p.n_steps; ++step) {
      for (int g = 0; g < ng; ++g) {
        const int n_act = n_act_of(g);
        if (COND) {      // this step's conditioning of post_process_1, fetched while the token is on its way; parked in shared memory
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1), s_ = 2 * q + (r & 1);
              const int tau = p.rf + tb[g][s_] + step;
              const int f = gen_cond_frame(tau - (p.rf - 1), p.cond_total - (p.rf - 1), p.cond_frames);
              condh[(j * 4 + r) * 256 + tid] = p.cond_head[((int64_t)stream_of(g, s_) * p.cond_frames + f) * 256 + row];
            }
        }
        HEAD_TS(0);
        wait_token(&skfull[g], step & 1);
        HEAD_TS(1);
        if (TRACE && g == 0 && tid == 0 && step < 64 && cid == 0) g_gen_ts[1024 + step * 16 + rank] = (long long)global_ns();
        float c[2][2][4];      // [m-tile][chain][fragment]: two chains of eight MMAs per m-tile (the pipe's issue rate bounds a phase)
#pragma unroll
        for (int which = 0; which < 2; ++which) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1);
              const float* bias = which == 0 ? p.bias_p1 : p.bias_p2;
              c[j][0][r] = HAS_BIAS ? bias[row] : 0.f;
              c[j][1][r] = 0.f;
            }
          const uint2* bsrc = (which == 0 ? hf + g * 16 * 32 : h1f) + lane;
#pragma unroll
          for (int kt = 0; kt < 16; ++kt) {
            const uint2 b = bsrc[kt * 32];
            if (which == 0) {
              mma_f16(c[0][kt & 1], p1w[0][kt], b.x, b.y);
              mma_f16(c[1][kt & 1], p1w[1][kt], b.x, b.y);
            } else if (kt < 8) {
              mma_f16(c[0][kt & 1], p2w[0][kt], b.x, b.y);
              mma_f16(c[1][kt & 1], p2w[1][kt], b.x, b.y);
            } else {
              const uint4 a0 = p2s[((2 * warp) * 16 + kt) * 32 + lane], a1 = p2s[((2 * warp + 1) * 16 + kt) * 32 + lane];
              mma_f16(c[0][kt & 1], a0, b.x, b.y);
              mma_f16(c[1][kt & 1], a1, b.x, b.y);
            }
          }
          HEAD_TS(3 + 2 * which);
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) c[j][0][r] = c[j][0][r] + c[j][1][r];
          if (which == 0 && COND) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int r = 0; r < 4; ++r) c[j][0][r] += condh[(j * 4 + r) * 256 + tid];      // (the thread's own values: no barrier)
          }
          if (which == 0) {
            // relu, fp16, and the accumulator fragment transposed into the B fragment of k-tile 2w + j of post_process_2
#pragma unroll
            for (int j = 0; j < 2; ++j)
              h1f[(2 * warp + j) * 32 + lane] = make_uint2(movm_t(pack_h2(fmaxf(c[j][0][0], 0.f), fmaxf(c[j][0][1], 0.f))),
                                                           movm_t(pack_h2(fmaxf(c[j][0][2], 0.f), fmaxf(c[j][0][3], 0.f))));
            __syncthreads();
            HEAD_TS(4);
          }
        }
        // ---- logits of rows 32w..32w+31 are in c[.][0][.]: thread (n8, q) has rows 16(2w+j) + n8 + 8h of streams 2q, 2q+1
        if (logits_out) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1), s = 2 * q + (r & 1);
              if (s < n_act) logits_out[((int64_t)step * p.n_streams + (g0 + g) * G + s) * 256 + row] = c[j][0][r];
            }
        }
        if (uniforms == nullptr) {
          // greedy topk(1) over the softmax (fast_generate.py:138-140) = the first maximum of the logits; found in registers:
          // per thread over its four rows, over the 8 row lanes by shuffles, over the 8 warps through 512 bytes of shared memory
          float bv[2];
          int bi[2];
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            bv[b] = c[0][0][b];
            bi[b] = 32 * warp + n8;
#pragma unroll
            for (int k = 1; k < 4; ++k) {      // rows in increasing order: (j, h) = (0,1), (1,0), (1,1)
              const float v = c[k >> 1][0][2 * (k & 1) + b];
              if (v > bv[b]) { bv[b] = v; bi[b] = 32 * warp + 16 * (k >> 1) + 8 * (k & 1) + n8; }
            }
          }
          // over the 8 row lanes by shuffles, floats compared as ordered unsigned keys (measured: a partial-mask redux.sync costs 350
          // cycles, and eight full-mask ones with neutral elements for the other streams ~90 each: CREDUX does not pipeline)
          uint32_t key[2] = {ordered_key(bv[0]), ordered_key(bv[1])};
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              const uint32_t ok = __shfl_xor_sync(0xffffffffu, key[b], o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi[b], o);
              if (ok > key[b] || (ok == key[b] && oi < bi[b])) { key[b] = ok; bi[b] = oi; }
            }
          }
          if (n8 == 0) {
            cand[warp * G + 2 * q] = make_float2(__uint_as_float(key[0]), __int_as_float(bi[0]));
            cand[warp * G + 2 * q + 1] = make_float2(__uint_as_float(key[1]), __int_as_float(bi[1]));
          }
          __syncthreads();
          HEAD_TS(6);
          {
            const int s = warp;      // one warp per stream; lanes 0..7 hold the 8 warps' candidates (rows increase with the warp)
            const float2 cv = cand[(lane & 7) * G + s];
            const uint32_t key = __float_as_uint(cv.x);
            const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
            const int idx = (int)__reduce_min_sync(0xffffffffu, key == kmax ? __float_as_uint(cv.y) : 0xffffu);
            if (lane == 0) {
              st_async_u1(r_note + (uint32_t)(g * G + s) * 4, (uint32_t)idx, r_notefull + g * 8);      // CTA 0's note[g][s]
              if (s < n_act) out[(int64_t)step * p.n_streams + (g0 + g) * G + s] = idx;
            }
          }
        } else {
          // inverse CDF of the softmax; one warp per stream, the logits through shared memory
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) lg[2 * q + (r & 1)][16 * (2 * warp + j) + n8 + 8 * (r >> 1)] = c[j][0][r];
          __syncthreads();
          HEAD_TS(6);
          const int s = warp;
          const float* lgs = lg[s];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 8; ++j) mx = fmaxf(mx, lgs[j * 32 + lane]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          if (lane == 0) {
            int pick = 255;
            float sum = 0.f;
            for (int k = 0; k < 256; ++k) sum += expf(lgs[k] - mx);
            const float inv = 1.f / sum;
            float total_p = 0.f;
            for (int k = 0; k < 256; ++k) total_p += expf(lgs[k] - mx) * inv;
            const float thr = uniforms[(int64_t)step * p.n_streams + stream_of(g, s)] * total_p;
            float cdf = 0.f;
            for (int k = 0; k < 256; ++k) {
              cdf += expf(lgs[k] - mx) * inv;
              if (cdf > thr) { pick = k; break; }
            }
            st_async_u1(r_note + (uint32_t)(g * G + s) * 4, (uint32_t)pick, r_notefull + g * 8);
            if (s < n_act) out[(int64_t)step * p.n_streams + (g0 + g) * G + s] = pick;
          }
        }
        HEAD_TS(7);
        if (tid == 0 && step + 1 < p.n_steps) mbar_expect_tx(&skfull[g], HF_BYTES);      // armed for the group's next sums
        // (the next group's / step's writes of h1f, cand and lg come behind its barriers; every warp has read them by then)
      }
    }
  }
  cluster_sync_all();      // nobody exits while a peer may still write into its shared memory
}

size_t gen_smem_bytes(int n_layers) {
  return ((sizeof(GenSmem) + 15) & ~size_t(15)) + (size_t)n_layers * G * XS * sizeof(float);
}

}  // namespace

int fast_gen_debug_ts(long long* h_buf, int n) {
  return cudaMemcpyFromSymbol(h_buf, g_gen_ts, (size_t)n * sizeof(long long)) == cudaSuccess ? WN_OK : WN_ERR_CUDA;
}

size_t fast_gen_frag_bytes(const Model& m) { return ((size_t)m.n_layers * FRAG_LAYER + 2 * FRAG_HEAD) * sizeof(uint4); }

// Builds the fp16 A-fragment image behind the regular packed image (called at the end of the weight pack, same stream).
int fast_gen_pack(const Model& m, const float* d_params, uint8_t* P, cudaStream_t s) {
  if (m.n_layers > GEN_MAXL) return WN_OK;      // generation refuses such models (fast_gen_steps)
  const PackLayout pl = pack_layout(m);
  FragPackArgs a{};
  a.n_layers = m.n_layers;
  for (int i = 0; i < m.n_layers; ++i) {
    a.filt[i] = m.layers[i].filt.w; a.gate[i] = m.layers[i].gate.w;
    a.dense[i] = m.layers[i].dense.w; a.skip[i] = m.layers[i].skip.w;
  }
  a.post1 = m.post1.w; a.post2 = m.post2.w;
  a.params = d_params;
  a.frag = reinterpret_cast<uint4*>(P + pl.gen_frag);
  gen_frag_pack_kernel<<<148, 256, 0, s>>>(a);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

int fast_gen_steps(Model& m, int n_streams, int n_steps, int push, const int64_t* d_first_note, const void* d_packed, void* d_state,
                   const float* d_uniforms, int64_t* d_out, float* d_logits, cudaStream_t s, const wn_gen_cond* cond) {
  WN_REQUIRE(fast_gen_supported(m) && m.n_layers <= GEN_MAXL, WN_ERR_UNSUPPORTED,
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
  p.frag = reinterpret_cast<const uint4*>(P + pl.gen_frag);
  const bool conditioned = cond && cond->d_fg;
  p.cond_fg = conditioned ? cond->d_fg : nullptr;
  p.cond_head = conditioned ? cond->d_head : nullptr;
  p.cond_frames = conditioned ? cond->frames : 0;
  p.cond_total = conditioned ? cond->total_len : 0;
  p.cond_gate_first = conditioned ? cond->gate_first : 0;
  p.rf = m.rf;
  for (int i = 0; i < m.n_layers; ++i) p.s_out[i] = m.layers[i].start;
  // (the WN_GEN_* switches are read on every call: the geometry tests flip them inside one process)
  const char* const pipe_env = getenv("WN_GEN_PIPE");
  const bool pipe_off = pipe_env && pipe_env[0] == '0';
  if (!pipe_off && m.n_layers <= 30) {
    // Weights-stationary cluster pipeline.  The step is the ring's latency (one group's token through every CTA and back), so the
    // groups are spread over as many clusters as fit on the GPU at once: a cluster serving one group never makes a token wait for
    // another group's tail work.  Two geometries: 4 blocks per CTA (9 CTAs for 30 layers, one group per cluster, up to 16 clusters)
    // when every group gets its own cluster, else 2 blocks per CTA (16 CTAs, up to 8 groups per cluster, up to 9 clusters).  With more
    // groups than that the clusters would run in waves, and the one-CTA-per-8-streams kernel below (same step time for any number
    // of streams up to 8 x 148) has the higher throughput.
    p.trace = [] { const char* e = getenv("WN_TS_RANK"); return e ? atoi(e) : 1; }();      // (the CTA whose timeline WN_TS=1 records)
    const int groups = (int)ceil_div(n_streams, G);
    const bool out_push = push == WN_PUSH_OUTPUT;
    static const bool ts_env = getenv("WN_TS") != nullptr;
    const bool pipe_force = pipe_env && pipe_env[0] == '1';
    const int gpc_env = [] { const char* e = getenv("WN_GEN_GPC"); return e ? atoi(e) : 0; }();      // groups per cluster, at least
    const int bpc_env = [] { const char* e = getenv("WN_GEN_BPC"); return e ? atoi(e) : 0; }();      // 2 or 4: that geometry only
    using Kern = void (*)(FastGenParams, char*, const int64_t*, const float*, int64_t*, float*);
    const int ki = (ts_env ? 4 : 0) + (m.use_bias ? 2 : 0) + (out_push ? 1 : 0);
    static const Kern kerns[2][8] = {
        {gen_pipe_kernel<false, false, false, 2>, gen_pipe_kernel<false, true, false, 2>, gen_pipe_kernel<true, false, false, 2>,
         gen_pipe_kernel<true, true, false, 2>, gen_pipe_kernel<false, false, true, 2>, gen_pipe_kernel<false, true, true, 2>,
         gen_pipe_kernel<true, false, true, 2>, gen_pipe_kernel<true, true, true, 2>},
        {gen_pipe_kernel<false, false, false, 4>, gen_pipe_kernel<false, true, false, 4>, gen_pipe_kernel<true, false, false, 4>,
         gen_pipe_kernel<true, true, false, 4>, gen_pipe_kernel<false, false, true, 4>, gen_pipe_kernel<false, true, true, 4>,
         gen_pipe_kernel<true, false, true, 4>, gen_pipe_kernel<true, true, true, 4>}};
    static const Kern kerns_cond[2][4] = {
        {gen_pipe_kernel<false, false, false, 2, true>, gen_pipe_kernel<false, true, false, 2, true>, gen_pipe_kernel<true, false, false, 2, true>,
         gen_pipe_kernel<true, true, false, 2, true>},
        {gen_pipe_kernel<false, false, false, 4, true>, gen_pipe_kernel<false, true, false, 4, true>, gen_pipe_kernel<true, false, false, 4, true>,
         gen_pipe_kernel<true, true, false, 4, true>}};
    static const size_t smem_of[2] = {pipe::PipeL<pipe::pipe_groups(2), 2>::TOTAL, pipe::PipeL<pipe::pipe_groups(4), 4>::TOTAL};
    static bool pipe_once[2][12] = {};
    static int max_clusters[2][32] = {};
    for (int v = 1; v >= 0; --v) {      // 4 blocks per CTA first
      const int bpc = v ? 4 : 2;
      if (bpc_env && bpc_env != bpc) continue;
      const Kern kp = conditioned ? kerns_cond[v][ki & 3] : kerns[v][ki];
      const int oi = conditioned ? 8 + (ki & 3) : ki;
      // CTA 0 (two blocks), the CTAs in between (bpc blocks), the last block CTA (two blocks), the head
      const int n_last = m.n_layers > 2 ? std::min(2, m.n_layers - 2) : 0, n_mid = std::max(0, m.n_layers - 2 - n_last);
      const int cs = 1 + (int)ceil_div(n_mid, bpc) + (n_last ? 1 : 0) + 1;
      if (!pipe_once[v][oi]) {
        WN_CHECK_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of[v]));
        WN_CHECK_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        pipe_once[v][oi] = true;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)cs); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem_of[v]; cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (max_clusters[v][cs] == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kp, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = -1; }
        max_clusters[v][cs] = n > 0 ? n : -1;
      }
      const int slots = pipe::pipe_groups(bpc);
      int gpc = max_clusters[v][cs] > 0 ? (int)ceil_div(groups, max_clusters[v][cs]) : slots + 1;
      if (gpc_env >= 1 && gpc_env <= slots && gpc_env >= gpc) gpc = gpc_env;
      if (pipe_force && v == 0 && gpc > slots) gpc = slots;
      if (gpc > slots) continue;
      p.gpc = gpc;
      cfg.gridDim = dim3((unsigned)(cs * (int)ceil_div(groups, gpc)));
      WN_PROF(v ? "gen_pipe4" : "gen_pipe", s);
      WN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kp, p, reinterpret_cast<char*>(d_state), d_first_note, d_uniforms, d_out, d_logits));
      WN_CHECK_LAUNCH();
      return WN_OK;
    }
  }
  WN_REQUIRE(!conditioned, WN_ERR_UNSUPPORTED,
             "conditioned half-precision generation runs on the cluster pipeline only (at most 30 blocks, %d streams on this GPU)", 9 * NG * G);
  const size_t smem = gen_smem_bytes(m.n_layers);
  static bool once = false;
  if (!once) {
    WN_CHECK_CUDA(cudaFuncSetAttribute(gen_steps_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gen_smem_bytes(GEN_MAXL)));
    once = true;
  }
  WN_PROF("gen_steps_bf16", s);
  gen_steps_bf16_kernel<<<(unsigned)ceil_div(n_streams, G), 256, smem, s>>>(p, reinterpret_cast<char*>(d_state), d_first_note,
                                                                           d_uniforms, d_out, d_logits);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace wn
