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
  int lpc;                         // gen_pipe_kernel: blocks per CTA (2; 1 in timing experiments)
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
};

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
// gen_pipe_kernel: ONE thread-block cluster of ceil(N / 2) + 1 CTAs (16 for the 30-layer model) serves up to NG = 8 groups of
// G = 8 streams.  CTA r < CS - 1 keeps the A fragments of blocks 2r and 2r + 1 (144 KB of fp16) in its REGISTERS for the whole
// launch; the last CTA keeps post_process_1 in registers and post_process_2 in shared memory.  Nothing is re-read from L2 per
// step any more.  A group's state hops from CTA to CTA through distributed shared memory as ONE token: after the last barrier
// of its second block a CTA issues three bulk copies (cp.async.bulk shared::cta -> shared::cluster: x fp32, x fp16, the running
// skip sums) into the next CTA's slot of that group; they complete on that slot's mbarrier there (complete_tx), which the
// consumer has armed with the byte count.  (The first version wrote the slot with per-thread st.shared::cluster: 36 cycles per
// warp store, 1150 + 1400 cycles per hop.)  The head picks the sample and sends it to CTA 0, which gathers the causal layer's two
// embedding rows from its shared-memory copy of the table.  Every group owns one slot per CTA and carries exactly one token
// around the ring, so there is no back-pressure and no barrier between groups: while group g is in CTA r, groups g - 1, g - 2,
// ... are in CTAs r + 1, r + 2, ...  Ring-buffer state stays in global memory (L2); the taps of a group's next step are requested
// with cp.async right after its token has left.
// Measured (tools/ts_gen.py): a dependent mma.sync.m16n8k16 chain costs ~85 cycles per link on B200, so every accumulation is
// split into chains of two; a block is then two exchange phases of ~350-400 cycles.  The step is the ring latency
// (30 blocks + 15 hops + head), the same for 8 and for 64 streams per cluster; independent clusters serve more streams.
constexpr int NG = 8;
// timing experiments (WN_TS=1): clock64 stamps of CTA 1 of cluster 0, group 0, 16 per step (wn_debug_ts with n < 0 reads them)
__device__ long long g_gen_ts[16 * 64];
#define GEN_TS(k)                                                                                   \
  do {                                                                                              \
    if (TRACE && rank == 1 && cid == 0 && g == 0 && tid == 0 && step < 64) g_gen_ts[step * 16 + (k)] = clock64(); \
  } while (0)
namespace pipe {
constexpr uint32_t X_BYTES = G * XS * 4, XH_BYTES = G * XH * 2, SK_BYTES = G * HS * 4;      // one group's x (fp32), x (fp16), skip sums
constexpr uint32_t OFF_BAR = 0;                                   // xfull[NG], skfull[NG], notefull[NG]
constexpr uint32_t OFF_NOTE = 256;                                // int note[NG][G], last[NG][G]
constexpr uint32_t OFF_SLOT = OFF_NOTE + 2 * NG * G * 4;          // int slot[NG][2][G]
constexpr uint32_t OFF_ZH = OFF_SLOT + NG * 2 * G * 4;            // __half zh[G][XH]
constexpr uint32_t OFF_SKIN = (OFF_ZH + XH_BYTES + 127) & ~127u;  // float skin[NG][G][HS]  (same offset in every role but CTA 0)
constexpr uint32_t OFF_XIN = OFF_SKIN + NG * SK_BYTES;            // block CTAs: float xin[NG][G][XS]   (same offset in all of them)
constexpr uint32_t OFF_XHIN = OFF_XIN + NG * X_BYTES;             //             __half xhin[NG][G][XH]
constexpr uint32_t OFF_TAPS = OFF_XHIN + NG * XH_BYTES;           //             float taps[NG][2][G][XS]
constexpr uint32_t BLOCK_BYTES = OFF_TAPS + NG * 2 * X_BYTES;
constexpr uint32_t OFF_WC = OFF_SKIN;                             // CTA 0 (nothing arrives but notes): float wc[2][256][64]
constexpr uint32_t OFF_TAPS0 = OFF_WC + 2 * 256 * 64 * 4;         //        its taps
constexpr uint32_t OFF_STG0 = OFF_TAPS0 + NG * 2 * X_BYTES;       //        two staging sets {x fp32, x fp16, skip sums} (outgoing tokens)
constexpr uint32_t STG_BYTES = X_BYTES + XH_BYTES + SK_BYTES;
constexpr uint32_t CTA0_BYTES = OFF_STG0 + 2 * STG_BYTES;
constexpr uint32_t OFF_P2 = OFF_SKIN + NG * SK_BYTES;             // head: uint4 p2[FRAG_HEAD]
constexpr uint32_t OFF_HH = OFF_P2 + FRAG_HEAD * 16;              //       __half hh[2][G][HH]
constexpr uint32_t OFF_LG = OFF_HH + 2 * G * HH * 2;              //       float lg[G][HS]
constexpr uint32_t HEAD_BYTES = OFF_LG + G * HS * 4;
constexpr uint32_t TOTAL = HEAD_BYTES > CTA0_BYTES ? (HEAD_BYTES > BLOCK_BYTES ? HEAD_BYTES : BLOCK_BYTES) : (CTA0_BYTES > BLOCK_BYTES ? CTA0_BYTES : BLOCK_BYTES);
static_assert(TOTAL <= 227 * 1024, "pipeline generation kernel: shared memory");
static_assert(X_BYTES % 16 == 0 && XH_BYTES % 16 == 0 && SK_BYTES % 16 == 0 && OFF_XIN % 16 == 0 && OFF_XHIN % 16 == 0 && OFF_STG0 % 16 == 0,
              "bulk copies move 16-byte units");

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
__device__ __forceinline__ void st_remote_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void arrive_remote(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// local shared memory -> another CTA's shared memory; the bytes complete on THAT CTA's mbarrier
__device__ __forceinline__ void bulk_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(src_cta),
               "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_2() { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// wait for completion `parity` of a local mbarrier whose arrivals / bytes come from another CTA of the cluster.  Bounded: a token that
// never arrives (a bug) traps after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void wait_token(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
template <int N_>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ void cp_wait_dyn(int n) {      // all but the newest n committed groups are complete
  switch (n) {
    case 0: cp_wait<0>(); break;
    case 1: cp_wait<1>(); break;
    case 2: cp_wait<2>(); break;
    case 3: cp_wait<3>(); break;
    case 4: cp_wait<4>(); break;
    case 5: cp_wait<5>(); break;
    case 6: cp_wait<6>(); break;
    default: cp_wait<7>(); break;
  }
}
}  // namespace pipe

// HAS_BIAS / PUSH_OUT are compile-time: every block of the ring is a chain of ~400 dependent-latency instructions per warp, and the
// run-time tests (and the predicated bias loads behind them) were a measurable part of it
template <bool HAS_BIAS, bool PUSH_OUT, bool TRACE>
__global__ void __launch_bounds__(256, 1)
gen_pipe_kernel(FastGenParams p, char* __restrict__ state, const int64_t* __restrict__ first_note, const float* __restrict__ uniforms,
                int64_t* __restrict__ out, float* __restrict__ logits_out) {
  using namespace pipe;
  extern __shared__ __align__(128) uint8_t sm[];
  const int N = p.n_layers;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n8 = lane >> 2, q = lane & 3;
  const int CS = (int)cluster_size(), rank = (int)cluster_rank(), cid = (int)blockIdx.x / CS;
  const int groups_total = (p.n_streams + G - 1) / G;
  const int g0 = cid * NG, ng = min(NG, groups_total - g0);
  const bool is_head = rank == CS - 1;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(sm + OFF_BAR);     // a group's x (fp32 + fp16) has arrived: tx bytes
  uint64_t* skfull = xfull + NG;                                    // its running skip sums have arrived (a second, larger copy: waited for later)
  uint64_t* notefull = xfull + 2 * NG;                              // CTA 0: the head's picks of a group
  int (*note)[G] = reinterpret_cast<int (*)[G]>(sm + OFF_NOTE);
  int (*last)[G] = reinterpret_cast<int (*)[G]>(sm + OFF_NOTE + NG * G * 4);
  int (*slot)[2][G] = reinterpret_cast<int (*)[2][G]>(sm + OFF_SLOT);
  float (*skin)[G][HS] = reinterpret_cast<float (*)[G][HS]>(sm + OFF_SKIN);
  const uint32_t sm_base = smem_u32(sm);
  if (tid == 0) {
    for (int g = 0; g < NG; ++g) {
      mbar_init(&xfull[g], 1);        // the consumer's own arming arrival; the data arrive as tx bytes
      mbar_init(&skfull[g], 1);
      mbar_init(&notefull[g], 8);     // the 8 pick warps of the head
    }
    fence_barrier_init();
    if (rank > 0)
      for (int g = 0; g < NG; ++g) {      // armed for step 0
        if (!is_head) mbar_expect_tx(&xfull[g], X_BYTES + XH_BYTES);
        mbar_expect_tx(&skfull[g], SK_BYTES);
      }
  }
  auto stream_of = [&](int g, int s) { return min((g0 + g) * G + s, p.n_streams - 1); };
  auto sptr = [&](int g, int s) { return state + (int64_t)stream_of(g, s) * p.state_stride; };
  auto n_act_of = [&](int g) { return min(G, p.n_streams - (g0 + g) * G); };
  cluster_sync_all();
  if (ng <= 0) {      // (cannot happen with the launch geometry of fast_gen_steps; all CTAs of the cluster agree)
    cluster_sync_all();
    return;
  }

  if (!is_head) {
    // ============================================================ block CTA: blocks l0, l0 + 1
    const int l0 = p.lpc * rank, nl = min(p.lpc, N - l0);
    const bool first = rank == 0, to_head = rank == CS - 2;
    __half (*zh)[XH] = reinterpret_cast<__half (*)[XH]>(sm + OFF_ZH);
    float* taps = reinterpret_cast<float*>(sm + (first ? OFF_TAPS0 : OFF_TAPS));      // [NG][2][G][XS]
    float* wc = reinterpret_cast<float*>(sm + OFF_WC);                                 // CTA 0
    // ---- resident A fragments (same per-warp ownership as gen_steps_bf16_kernel)
    uint4 fgw[2][8], w2w[2][12];
    {
      const uint4* const fbase = p.frag + lane;
#pragma unroll
      for (int li = 0; li < 2; ++li) {
        const int i = min(l0 + li, N - 1);
        const uint4* base = fbase + (int64_t)i * FRAG_LAYER;
#pragma unroll
        for (int k = 0; k < 8; ++k) fgw[li][k] = base[FRAG_FG + (warp * 8 + k) * 32];
        if (warp < 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) w2w[li][k] = base[FRAG_D + (warp * 4 + k) * 32];
#pragma unroll
          for (int k = 0; k < 4; ++k) w2w[li][4 + k] = base[FRAG_S + (warp * 4 + k) * 32];
#pragma unroll
          for (int k = 8; k < 12; ++k) w2w[li][k] = make_uint4(0u, 0u, 0u, 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 12; ++k) w2w[li][k] = base[FRAG_S + ((4 + 3 * (warp - 4)) * 4 + k) * 32];
        }
      }
    }
    for (int e = tid; e < ng * nl * G; e += 256) {
      const int s = e % G, li = (e / G) % nl, g = e / (G * nl);
      slot[g][li][s] = (int)(reinterpret_cast<const int64_t*>(sptr(g, s))[0] % p.dil[l0 + li]);
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
    // taps of (group g, its next use): slot as stored (advance = 0) or slot + 1 (advance = 1, also stored)
    auto prefetch_taps = [&](int g, int advance) {
      for (int e = tid; e < nl * G * 16; e += 256) {
        const int c4 = (e & 15) * 4, s = (e >> 4) % G, li = (e >> 4) / G;
        int sl = slot[g][li][s];
        if (advance) {
          sl = sl + 1 == p.dil[l0 + li] ? 0 : sl + 1;
          if (c4 == 0) slot[g][li][s] = sl;
        }
        const float* src = reinterpret_cast<const float*>(sptr(g, s) + 16) + ((int64_t)p.ring_off[l0 + li] + sl) * 64 + c4;
        __pipeline_memcpy_async(taps + ((size_t)(g * 2 + li) * G + s) * XS + c4, src, 16);
      }
      __pipeline_commit();
    };
    for (int g = 0; g < ng; ++g) prefetch_taps(g, 0);
    for (int g = ng; g < NG; ++g) __pipeline_commit();      // always NG commit groups per ring period: the wait below is a constant
    // remote addresses of the next CTA's slots (same offsets there)
    const uint32_t r_xin = map_to(sm_base + OFF_XIN, rank + 1), r_xhin = map_to(sm_base + OFF_XHIN, rank + 1);
    const uint32_t r_skin = map_to(sm_base + OFF_SKIN, rank + 1);
    const uint32_t r_xfull = map_to(smem_u32(xfull), rank + 1), r_skfull = map_to(smem_u32(skfull), rank + 1);

    for (int step = 0; step < p.n_steps; ++step) {
      for (int g = 0; g < ng; ++g) {
        const int n_act = n_act_of(g);
        float (*xs)[XS];
        __half (*xhs)[XH];
        float (*sko)[HS];       // where this CTA's outgoing skip sums are staged
        if (first) {
          // CTA 0: the gathered input and the outgoing token live in one of two staging sets; a set is rewritten two groups later,
          // when its bulk copies have long been read
          uint8_t* stg = sm + OFF_STG0 + ((step * ng + g) & 1) * STG_BYTES;
          xs = reinterpret_cast<float (*)[XS]>(stg);
          xhs = reinterpret_cast<__half (*)[XH]>(stg + X_BYTES);
          sko = reinterpret_cast<float (*)[HS]>(stg + X_BYTES + XH_BYTES);
          if (tid == 0) bulk_wait_read_2();      // the two bulk groups of the iteration that used this staging set have been read
          if (step > 0) wait_token(&notefull[g], (step - 1) & 1);
          __syncthreads();      // (tid 0's wait on the staging set before anybody writes it)
          for (int e = tid; e < G * 64; e += 256) {      // causal layer: two embedding rows (fast_generate.py:111-116)
            const int s = e >> 6, r = e & 63;
            float v = wc[last[g][s] * 64 + r] + wc[(256 + note[g][s]) * 64 + r];
            if (HAS_BIAS) v += p.bias_c[r];
            xs[s][r] = v;
            xhs[s][r] = __float2half_rn(v);
          }
        } else {
          GEN_TS(0);
          wait_token(&xfull[g], step & 1);
          GEN_TS(1);
          if (tid == 0 && step + 1 < p.n_steps) mbar_expect_tx(&xfull[g], X_BYTES + XH_BYTES);      // armed for the group's next token
          xs = reinterpret_cast<float (*)[G][XS]>(sm + OFF_XIN)[g];
          xhs = reinterpret_cast<__half (*)[G][XH]>(sm + OFF_XHIN)[g];
          sko = skin[g];
        }
        cp_wait<NG - 1>();                     // this group's taps (requested one ring period = NG commit groups ago)
        __syncthreads();
        GEN_TS(2);
        if (first && tid < G) last[g][tid] = note[g][tid];
        float sk[3][4];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int r = 0; r < 4; ++r) sk[j][r] = 0.f;
        float* const ring0 = reinterpret_cast<float*>(sptr(g, 2 * q) + 16);
        float* const ring1 = reinterpret_cast<float*>(sptr(g, 2 * q + 1) + 16);
#pragma unroll
        for (int li = 0; li < 2; ++li) {
          if (li < nl) {
            const int i = l0 + li;
            const bool last_block = li == nl - 1;
            // ---- [f|g] = W0 old + W1 x : warp w owns filter and gate channels 8w..8w+7 for all 8 streams.  Four chains of two MMAs
            const int ch = 8 * warp + n8;
            float c0[4], c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f}, c3[4] = {0.f, 0.f, 0.f, 0.f};
            {
              const float bf = HAS_BIAS ? p.bias_fg[i * 128 + ch] : 0.f, bg = HAS_BIAS ? p.bias_fg[i * 128 + 64 + ch] : 0.f;
              c0[0] = c0[1] = bf;
              c0[2] = c0[3] = bg;
            }
            const float* orow = taps + ((size_t)(g * 2 + li) * G + n8) * XS + 2 * q;
            const __half* xrow = &xhs[n8][2 * q];
            uint32_t bo[4][2], bx[4][2];
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
              const float2 v0 = *reinterpret_cast<const float2*>(orow + kt * 16), v1 = *reinterpret_cast<const float2*>(orow + kt * 16 + 8);
              bo[kt][0] = pack_h2(v0.x, v0.y);
              bo[kt][1] = pack_h2(v1.x, v1.y);
              bx[kt][0] = *reinterpret_cast<const uint32_t*>(xrow + kt * 16);
              bx[kt][1] = *reinterpret_cast<const uint32_t*>(xrow + kt * 16 + 8);
            }
            mma_f16(c1, fgw[li][4], bx[0][0], bx[0][1]);
            mma_f16(c3, fgw[li][6], bx[2][0], bx[2][1]);
            mma_f16(c0, fgw[li][0], bo[0][0], bo[0][1]);
            mma_f16(c2, fgw[li][2], bo[2][0], bo[2][1]);
            mma_f16(c1, fgw[li][5], bx[1][0], bx[1][1]);
            mma_f16(c3, fgw[li][7], bx[3][0], bx[3][1]);
            mma_f16(c0, fgw[li][1], bo[1][0], bo[1][1]);
            mma_f16(c2, fgw[li][3], bo[3][0], bo[3][1]);
            GEN_TS(3 + 5 * li);
            zh[2 * q][ch] = __float2half_rn(gate_z((c0[0] + c1[0]) + (c2[0] + c3[0]), (c0[2] + c1[2]) + (c2[2] + c3[2])));
            zh[2 * q + 1][ch] = __float2half_rn(gate_z((c0[1] + c1[1]) + (c2[1] + c3[1]), (c0[3] + c1[3]) + (c2[3] + c3[3])));
            __syncthreads();
            GEN_TS(4 + 5 * li);
            uint32_t bz[4][2];
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
              const __half* zr = &zh[n8][kt * 16 + 2 * q];
              bz[kt][0] = *reinterpret_cast<const uint32_t*>(zr);
              bz[kt][1] = *reinterpret_cast<const uint32_t*>(zr + 8);
            }
            if (warp < 4) {
              const int chd = 16 * warp + n8;
              const int s0 = 2 * q, s1 = 2 * q + 1;
              float dn[4], dm[4] = {0.f, 0.f, 0.f, 0.f};
              dn[0] = xs[s0][chd]; dn[1] = xs[s1][chd]; dn[2] = xs[s0][chd + 8]; dn[3] = xs[s1][chd + 8];     // residual
              const float xi[4] = {dn[0], dn[1], dn[2], dn[3]};
              if (HAS_BIAS) {
                const float b0 = p.bias_d[i * 64 + chd], b1 = p.bias_d[i * 64 + chd + 8];
                dn[0] += b0; dn[1] += b0; dn[2] += b1; dn[3] += b1;
              }
              const int o0 = (p.ring_off[i] + slot[g][li][s0]) * 64 + chd, o1 = (p.ring_off[i] + slot[g][li][s1]) * 64 + chd;
              mma_f16(dn, w2w[li][0], bz[0][0], bz[0][1]);
              mma_f16(dm, w2w[li][2], bz[2][0], bz[2][1]);
              mma_f16(dn, w2w[li][1], bz[1][0], bz[1][1]);
              mma_f16(dm, w2w[li][3], bz[3][0], bz[3][1]);
#pragma unroll
              for (int r = 0; r < 4; ++r) dn[r] += dm[r];
              GEN_TS(5 + 5 * li);
              if (!(last_block && to_head)) {      // in place: the next block's input, or the token's payload
                xs[s0][chd] = dn[0]; xs[s1][chd] = dn[1]; xs[s0][chd + 8] = dn[2]; xs[s1][chd + 8] = dn[3];
                xhs[s0][chd] = __float2half_rn(dn[0]); xhs[s1][chd] = __float2half_rn(dn[1]);
                xhs[s0][chd + 8] = __float2half_rn(dn[2]); xhs[s1][chd + 8] = __float2half_rn(dn[3]);
                if (last_block) {
                  // the token leaves as soon as the four dense warps have written it: the skip MMAs, the ring pushes and the skip
                  // sums below are off the ring's critical path
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                  asm volatile("bar.sync 1, 128;" ::: "memory");
                  if (tid == 0) {
                    bulk_to_peer(r_xin + g * X_BYTES, smem_u32(xs), X_BYTES, r_xfull + g * 8);
                    bulk_to_peer(r_xhin + g * XH_BYTES, smem_u32(xhs), XH_BYTES, r_xfull + g * 8);
                    bulk_commit();
                  }
                }
              } else if (tid == 0) {
                bulk_commit();      // (two bulk groups per iteration in every CTA: see bulk_wait_read_2)
              }
#pragma unroll
              for (int kt = 0; kt < 4; ++kt) mma_f16(sk[0], w2w[li][4 + kt], bz[kt][0], bz[kt][1]);
              constexpr bool out_push = PUSH_OUT;                                     // fast_generate.py:128-129
              if (s0 < n_act) { ring0[o0] = out_push ? dn[0] : xi[0]; ring0[o0 + 8] = out_push ? dn[2] : xi[2]; }
              if (s1 < n_act) { ring1[o1] = out_push ? dn[1] : xi[1]; ring1[o1 + 8] = out_push ? dn[3] : xi[3]; }
              GEN_TS(6 + 5 * li);
            } else {
#pragma unroll
              for (int kt = 0; kt < 4; ++kt) {
                mma_f16(sk[0], w2w[li][kt], bz[kt][0], bz[kt][1]);
                mma_f16(sk[1], w2w[li][4 + kt], bz[kt][0], bz[kt][1]);
                mma_f16(sk[2], w2w[li][8 + kt], bz[kt][0], bz[kt][1]);
              }
            }
            if (last_block) {
              // running skip sums: what arrived behind the token + this CTA's two blocks, in place
              if (!first) {
                wait_token(&skfull[g], step & 1);
                if (tid == 0 && step + 1 < p.n_steps) mbar_expect_tx(&skfull[g], SK_BYTES);
              }
              const int nt = warp < 4 ? 1 : 3, mt0 = warp < 4 ? warp : 4 + 3 * (warp - 4);
#pragma unroll
              for (int j = 0; j < 3; ++j)
                if (j < nt) {
#pragma unroll
                  for (int r = 0; r < 4; ++r) {
                    const int row = 16 * (mt0 + j) + n8 + 8 * (r >> 1), s = 2 * q + (r & 1);
                    sko[s][row] = sk[j][r] + (first ? 0.f : skin[g][s][row]);
                  }
                }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the bulk copies below read these generic-proxy writes
            }
            __syncthreads();
            GEN_TS(7 + 5 * li);
          }
        }
        if (tid == 0) {      // behind the token: the running skip sums -> the next CTA's slot of this group
          bulk_to_peer(r_skin + g * SK_BYTES, smem_u32(sko), SK_BYTES, r_skfull + g * 8);
          bulk_commit();
        }
        GEN_TS(13);
        if (step + 1 < p.n_steps) prefetch_taps(g, 1);
        else __pipeline_commit();               // (keeps the group accounting of cp_wait_dyn uniform)
        GEN_TS(14);
      }
      for (int g = ng; g < NG; ++g) __pipeline_commit();
    }
    cp_wait<0>();
    if (tid == 0) bulk_wait_read_0();
    if (first) {
      __syncthreads();
      for (int e = tid; e < ng * G; e += 256) {
        const int g = e / G, s = e % G;
        if (s < n_act_of(g)) {
          reinterpret_cast<int64_t*>(sptr(g, s))[0] += p.n_steps;
          reinterpret_cast<int64_t*>(sptr(g, s))[1] = last[g][s];
        }
      }
    }
  } else {
    // ============================================================ head CTA: relu(sum skips) -> P1 -> relu -> P2 -> pick
    uint4* p2s = reinterpret_cast<uint4*>(sm + OFF_P2);
    __half (*hh)[G][HH] = reinterpret_cast<__half (*)[G][HH]>(sm + OFF_HH);
    float (*lg)[HS] = reinterpret_cast<float (*)[HS]>(sm + OFF_LG);
    const uint4* const headA = p.frag + (int64_t)N * FRAG_LAYER;
    uint4 p1w[2][16];      // post_process_1: this warp's two m-tiles, resident
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 16; ++k) p1w[j][k] = headA[((2 * warp + j) * 16 + k) * 32 + lane];
    for (int e = tid; e < FRAG_HEAD; e += 256) p2s[e] = headA[FRAG_HEAD + e];
    __syncthreads();
    const uint32_t r_note = map_to(sm_base + OFF_NOTE, 0), r_notefull = map_to(smem_u32(notefull), 0);
    for (int step = 0; step < p.n_steps; ++step) {
      for (int g = 0; g < ng; ++g) {
        const int n_act = n_act_of(g);
        wait_token(&skfull[g], step & 1);
        if (tid == 0 && step + 1 < p.n_steps) mbar_expect_tx(&skfull[g], SK_BYTES);
        for (int e = tid; e < G * 256; e += 256) {
          const int s = e >> 8, row = e & 255;
          hh[0][s][row] = __float2half_rn(fmaxf(skin[g][s][row] + (HAS_BIAS ? p.bias_skip[row] : 0.f), 0.f));
        }
        __syncthreads();
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          float c[2][4][4];      // [m-tile][chain][fragment]: four chains of four MMAs per m-tile
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1);
              const float* bias = which == 0 ? p.bias_p1 : p.bias_p2;
              c[j][0][r] = HAS_BIAS ? bias[row] : 0.f;
              c[j][1][r] = c[j][2][r] = c[j][3][r] = 0.f;
            }
          const __half* hr = &hh[which][n8][2 * q];
#pragma unroll
          for (int kt = 0; kt < 16; ++kt) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(hr + kt * 16), b1 = *reinterpret_cast<const uint32_t*>(hr + kt * 16 + 8);
            if (which == 0) {
              mma_f16(c[0][kt & 3], p1w[0][kt], b0, b1);
              mma_f16(c[1][kt & 3], p1w[1][kt], b0, b1);
            } else {
              const uint4 a0 = p2s[((2 * warp) * 16 + kt) * 32 + lane], a1 = p2s[((2 * warp + 1) * 16 + kt) * 32 + lane];
              mma_f16(c[0][kt & 3], a0, b0, b1);
              mma_f16(c[1][kt & 3], a1, b0, b1);
            }
          }
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int row = 16 * (2 * warp + j) + n8 + 8 * (r >> 1), s = 2 * q + (r & 1);
              const float v = (c[j][0][r] + c[j][1][r]) + (c[j][2][r] + c[j][3][r]);
              if (which == 0) {
                hh[1][s][row] = __float2half_rn(fmaxf(v, 0.f));
              } else {
                lg[s][row] = v;
                if (logits_out && s < n_act) logits_out[((int64_t)step * p.n_streams + (g0 + g) * G + s) * 256 + row] = v;
              }
            }
          __syncthreads();
        }
        // ---- pick: greedy topk(1) over the softmax (fast_generate.py:138-140) or inverse CDF; one warp per stream
        {
          const int s = warp;
          const float* lgs = lg[s];
          float v[8];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = lgs[lane * 8 + j];
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
            pick = 255;
            if (lane == 0) {
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
            }
            pick = __shfl_sync(0xffffffffu, pick, 0);
          }
          if (lane == 0) {
            if (s < n_act) out[(int64_t)step * p.n_streams + (g0 + g) * G + s] = pick;
            st_remote_u32(r_note + (uint32_t)(g * G + s) * 4, (uint32_t)pick);      // CTA 0's note[g][s]
            arrive_remote(r_notefull + g * 8);
          }
        }
        // (no barrier here: the next group's first barrier comes after every warp has finished this pick)
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
                   const float* d_uniforms, int64_t* d_out, float* d_logits, cudaStream_t s) {
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
  static const bool pipe_off = [] { const char* e = getenv("WN_GEN_PIPE"); return e && e[0] == '0'; }();
  if (!pipe_off && m.n_layers <= 30) {
    // weights-stationary cluster pipeline: ceil(N / 2) block CTAs + the head CTA per cluster, 8 groups of 8 streams per cluster
    static const int lpc_env = [] { const char* e = getenv("WN_GEN_LPC"); return e && e[0] == '1' ? 1 : 2; }();
    p.lpc = (lpc_env == 1 && m.n_layers <= 15) ? 1 : 2;
    p.trace = 0;
    const int cs = (m.n_layers + p.lpc - 1) / p.lpc + 1;
    const int groups = (int)ceil_div(n_streams, G), n_clusters = (int)ceil_div(groups, NG);
    const bool out_push = push == WN_PUSH_OUTPUT;
    static const bool ts_env = getenv("WN_TS") != nullptr;
    auto kp = ts_env ? (m.use_bias ? (out_push ? gen_pipe_kernel<true, true, true> : gen_pipe_kernel<true, false, true>)
                                   : (out_push ? gen_pipe_kernel<false, true, true> : gen_pipe_kernel<false, false, true>))
                     : (m.use_bias ? (out_push ? gen_pipe_kernel<true, true, false> : gen_pipe_kernel<true, false, false>)
                                   : (out_push ? gen_pipe_kernel<false, true, false> : gen_pipe_kernel<false, false, false>));
    static bool pipe_once[4] = {false, false, false, false};
    const int ki = (m.use_bias ? 2 : 0) + (out_push ? 1 : 0);
    if (!pipe_once[ki]) {
      WN_CHECK_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe::TOTAL));
      WN_CHECK_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      pipe_once[ki] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * n_clusters)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = pipe::TOTAL; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // The pipeline is the lower-latency kernel (26 vs 32 us per step) as long as every cluster is resident at once; with more stream
    // groups than that its clusters would run in waves, and the one-CTA-per-8-streams kernel below (same step time for any number of
    // streams up to 8 x 148) has the higher throughput.
    static int max_clusters[32] = {};
    if (max_clusters[cs] == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kp, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = -1; }
      max_clusters[cs] = n > 0 ? n : -1;
    }
    static const bool pipe_force = [] { const char* e = getenv("WN_GEN_PIPE"); return e && e[0] == '1'; }();
    if (pipe_force || n_clusters <= max_clusters[cs]) {
      WN_PROF("gen_pipe", s);
      WN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kp, p, reinterpret_cast<char*>(d_state), d_first_note, d_uniforms, d_out, d_logits));
      WN_CHECK_LAUNCH();
      return WN_OK;
    }
  }
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
