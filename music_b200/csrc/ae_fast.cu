// ae_fast.cu - encoder and conditioning kernels of the autoencoder's "bf16" mode (see ae_fast.cuh for the design).
//   _encode   wavenet_autoencoder/model1.py:137-156        _conditon / fresh 1x1 convs   :178-179, :216-217, :227-247
// Tensor-core tiles are mma.sync.m16n8k16 (bf16 operands, fp32 accumulation): at 32 channels a layer is 6 KFLOP per time step
// against 384 bytes of fp32 activations - the kernels are bound by HBM and launch latency, not by the tensor pipe, and the
// register-chained form needs no shared-memory round trip between the two products of a layer.
#include "ae_fast.cuh"
#include "tc05.cuh"

namespace wn {
using tc::pdl_launch_dependents;
using tc::pdl_wait;
namespace {

constexpr int C = kEncC;

__device__ __forceinline__ uint32_t bf2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t bf2_relu(float2 v) { return bf2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)); }
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// Fragment conventions (PTX m16n8k16, g = lane / 4, tg = lane % 4):
//   A (16 x 16, row):  a0 = (row g,     k 2tg, 2tg+1)   a1 = (row g + 8, k 2tg..)   a2 = (row g, k 2tg + 8..)   a3 = (row g + 8, k 2tg + 8..)
//   B (16 x 8,  col):  b0 = (k 2tg, 2tg+1; n g)         b1 = (k 2tg + 8..; n g)
//   C (16 x 8):        c0, c1 = (row g, n 2tg, 2tg+1)   c2, c3 = (row g + 8, n 2tg, 2tg+1)
// A 16 x 32 activation tile held in C layout - float2 v[nt][h] = (row g + 8h, channels 8nt + 2tg, +1) - IS the A operand of the next
// product over those 32 channels: k-step ks uses v[2ks][0], v[2ks][1], v[2ks+1][0], v[2ks+1][1].
struct Tile32 {
  float2 v[4][2];
};
__device__ __forceinline__ void a_frag(const Tile32& t, int ks, bool relu, uint32_t (&a)[4]) {
  if (relu) {
    a[0] = bf2_relu(t.v[2 * ks][0]); a[1] = bf2_relu(t.v[2 * ks][1]);
    a[2] = bf2_relu(t.v[2 * ks + 1][0]); a[3] = bf2_relu(t.v[2 * ks + 1][1]);
  } else {
    a[0] = bf2(t.v[2 * ks][0].x, t.v[2 * ks][0].y); a[1] = bf2(t.v[2 * ks][1].x, t.v[2 * ks][1].y);
    a[2] = bf2(t.v[2 * ks + 1][0].x, t.v[2 * ks + 1][0].y); a[3] = bf2(t.v[2 * ks + 1][1].x, t.v[2 * ks + 1][1].y);
  }
}
// rows r0 = t0 + g and r0 + 8 of a (L, 32) fp32 array; rows outside [lo, hi) read as zero
__device__ __forceinline__ void load_tile(Tile32& t, const float* base, int r0, int lo, int hi, int tg) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = r0 + 8 * h;
    const bool ok = r >= lo && r < hi;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) t.v[nt][h] = ok ? ld2(base + (int64_t)r * C + 8 * nt + 2 * tg) : make_float2(0.f, 0.f);
  }
}
__device__ __forceinline__ void zero_acc(float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// ------------------------------------------------------------------------------------------------ encoder layer, forward
// weights in the Conv1d layouts of the flat parameter vector: w_dil (De, Re, 2), w_dense (Re, De, 1)
constexpr int ENC_WARPS = 8;
// The 48 B-fragment registers of a thread are gathered once per CTA into shared memory ([slot][lane], conflict-free) - read per
// thread straight from the Conv1d layouts they cost 96 scattered loads per warp, more than a 16-row tile's worth of work.
__global__ void __launch_bounds__(32 * ENC_WARPS) enc_fwd_layer_kernel(const float* __restrict__ x, float* __restrict__ T, float* __restrict__ xo,
                                                                       const float* __restrict__ w_dil, const float* __restrict__ w_dense,
                                                                       int L, int d, int s_out, int n_chunks) {
  __shared__ uint32_t wfrag[48][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < 48 * 32; e += 32 * ENC_WARPS) {
    const int slot = e >> 5, ln = e & 31, gg = ln >> 2, tt = ln & 3;
    uint32_t v;
    if (slot < 32) {      // dilated conv: K = 64 = (tap 0: x[t - d] channels, tap 1: x[t] channels), N = De;  B[k][n] = w_dil[n][k % 32][k / 32]
      const int ks = slot >> 3, nt = (slot >> 1) & 3, j = slot & 1;
      const int n = 8 * nt + gg, k = 16 * (ks & 1) + 8 * j + 2 * tt, tap = ks >> 1;
      v = bf2(w_dil[(n * C + k) * 2 + tap], w_dil[(n * C + k + 1) * 2 + tap]);
    } else {              // dense: B[k = de][n = r] = w_dense[r][de]
      const int sl = slot - 32, ks = sl >> 3, nt = (sl >> 1) & 3, j = sl & 1;
      const int n = 8 * nt + gg, k = 16 * ks + 8 * j + 2 * tt;
      v = bf2(w_dense[n * C + k], w_dense[n * C + k + 1]);
    }
    wfrag[slot][ln] = v;
  }
  pdl_launch_dependents();      // (programmatic dependent launch: the weight gather above overlaps the previous layer's tail ...
  pdl_wait();                   //  ... activations are touched only after the previous kernel has completed)
  __syncthreads();
  uint32_t wT[4][4][2], wD[2][4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) wT[ks][nt][j] = wfrag[(ks * 4 + nt) * 2 + j][lane];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) wD[ks][nt][j] = wfrag[32 + (ks * 4 + nt) * 2 + j][lane];
  const float* xb = x + (int64_t)b * L * C;
  float* Tb = T + (int64_t)b * L * C;
  float* xob = xo + (int64_t)b * L * C;
  for (int ch = blockIdx.x * ENC_WARPS + warp; ch < n_chunks; ch += gridDim.x * ENC_WARPS) {
    const int r0 = s_out + ch * 16 + g;
    Tile32 cur, prv;
    load_tile(cur, xb, r0, 0, L, tg);
    load_tile(prv, xb - (int64_t)d * C, r0, d, L + d, tg);      // rows r - d (always >= s_out - d >= 0 for valid rows)
    float acc[4][4];
    zero_acc(acc);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      a_frag(ks < 2 ? prv : cur, ks & 1, true, a);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(acc[nt], a, wT[ks][nt]);
    }
    Tile32 t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      t.v[nt][0] = make_float2(acc[nt][0], acc[nt][1]);
      t.v[nt][1] = make_float2(acc[nt][2], acc[nt][3]);
    }
    float o[4][4];
    zero_acc(o);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
      a_frag(t, ks, true, a);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(o[nt], a, wD[ks][nt]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + 8 * h;
      if (r < L) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int64_t e = (int64_t)r * C + 8 * nt + 2 * tg;
          st2(Tb + e, t.v[nt][h].x, t.v[nt][h].y);
          st2(xob + e, cur.v[nt][h].x + o[nt][2 * h], cur.v[nt][h].y + o[nt][2 * h + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ encoder layer, data gradient
__global__ void __launch_bounds__(32 * ENC_WARPS) enc_bwd_layer_kernel(const float* __restrict__ gxn, const float* __restrict__ T,
                                                                       const float* __restrict__ x, float* __restrict__ gx,
                                                                       float* __restrict__ dT, const float* __restrict__ w_dil,
                                                                       const float* __restrict__ w_dense, int L, int d, int s_out,
                                                                       int n_chunks) {
  __shared__ uint32_t wfrag[48][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
  const int b = blockIdx.y, s_in = s_out - d;
  // dT_pre[t][de] = sum_r gX[t][r] w_dense[r][de]:  B[k = r][n = de] = w_dense[k][n]
  // Y[t][r] = sum_de dT[t + d][de] w_dil[de][r][0] + dT[t][de] w_dil[de][r][1]:  B[k][n = r] = w_dil[k % 32][n][k / 32]
  for (int e = threadIdx.x; e < 48 * 32; e += 32 * ENC_WARPS) {
    const int slot = e >> 5, ln = e & 31, gg = ln >> 2, tt = ln & 3;
    uint32_t v;
    if (slot < 16) {
      const int ks = slot >> 3, nt = (slot >> 1) & 3, j = slot & 1;
      const int n = 8 * nt + gg, k = 16 * ks + 8 * j + 2 * tt;
      v = bf2(w_dense[k * C + n], w_dense[(k + 1) * C + n]);
    } else {
      const int sl = slot - 16, ks = sl >> 3, nt = (sl >> 1) & 3, j = sl & 1;
      const int n = 8 * nt + gg, k = 16 * (ks & 1) + 8 * j + 2 * tt, tap = ks >> 1;
      v = bf2(w_dil[(k * C + n) * 2 + tap], w_dil[((k + 1) * C + n) * 2 + tap]);
    }
    wfrag[slot][ln] = v;
  }
  pdl_launch_dependents();
  pdl_wait();
  __syncthreads();
  uint32_t wA[2][4][2], wB[4][4][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) wA[ks][nt][j] = wfrag[(ks * 4 + nt) * 2 + j][lane];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) wB[ks][nt][j] = wfrag[16 + (ks * 4 + nt) * 2 + j][lane];
  const int64_t boff = (int64_t)b * L * C;
  const float *gb = gxn + boff, *Tb = T + boff, *xb = x + boff;
  float *gxb = gx + boff, *dTb = dT + boff;
  for (int ch = blockIdx.x * ENC_WARPS + warp; ch < n_chunks; ch += gridDim.x * ENC_WARPS) {
    const int r0 = s_in + ch * 16 + g;
    Tile32 dts[2];      // dT[t + d] (tap 0) and dT[t] (tap 1), masked; rows outside the layer's output range are zero
    Tile32 gcur;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const int shift = which == 0 ? d : 0;
      Tile32 gt, tt;
      load_tile(gt, gb + (int64_t)shift * C, r0, s_out - shift, L - shift, tg);
      load_tile(tt, Tb + (int64_t)shift * C, r0, s_out - shift, L - shift, tg);
      float acc[4][4];
      zero_acc(acc);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
        a_frag(gt, ks, false, a);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816(acc[nt], a, wA[ks][nt]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          dts[which].v[nt][h] = make_float2(tt.v[nt][h].x > 0.f ? acc[nt][2 * h] : 0.f, tt.v[nt][h].y > 0.f ? acc[nt][2 * h + 1] : 0.f);
      if (which == 1) {
        gcur = gt;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = r0 + 8 * h;
          if (r >= s_out && r < L) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) st2(dTb + (int64_t)r * C + 8 * nt + 2 * tg, dts[1].v[nt][h].x, dts[1].v[nt][h].y);
          }
        }
      }
    }
    float y[4][4];
    zero_acc(y);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      a_frag(dts[ks >> 1], ks & 1, false, a);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(y[nt], a, wB[ks][nt]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + 8 * h;
      if (r < L) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int64_t e = (int64_t)r * C + 8 * nt + 2 * tg;
          const float2 xv = ld2(xb + e);
          st2(gxb + e, gcur.v[nt][h].x + (xv.x > 0.f ? y[nt][2 * h] : 0.f), gcur.v[nt][h].y + (xv.y > 0.f ? y[nt][2 * h + 1] : 0.f));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ weight gradients, all layers
// grid (row chunks, layers, B).  Per 16-row step a warp stages five [16][32] bf16 tiles (row pitch 80 bytes: conflict-free
// ldmatrix) and takes both MMA operands from them with ldmatrix.trans - the reduction index of these products is TIME.
//   dW_dense[r][de] += sum_t gX_{i+1}[t][r] relu(T[t][de])       dW_dil[de][r][1] += sum_t dT[t][de] relu(x[t][r])
//                                                                dW_dil[de][r][0] += sum_t dT[t][de] relu(x[t - d][r])
constexpr int WG_ROWS = 2048, WG_PITCH = 40;      // rows per CTA; tile row pitch in bf16
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__global__ void __launch_bounds__(128) enc_wgrad_kernel(EncWgradArgs a, float* __restrict__ G) {
  __shared__ __align__(16) __nv_bfloat16 tiles[4][5][16 * WG_PITCH];
  __shared__ float red[3][C * C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
  const int ly = blockIdx.y, b = blockIdx.z;
  const int d = a.dil[ly], s_out = a.s_out[ly], L = a.L;
  const int row_lo = s_out + blockIdx.x * WG_ROWS;
  for (int e = threadIdx.x; e < 3 * C * C; e += 128) (&red[0][0])[e] = 0.f;
  __syncthreads();
  if (row_lo >= L) return;
  const int row_hi = min(L, row_lo + WG_ROWS);
  const int64_t boff = (int64_t)b * L * C;
  const float* xs = a.x + a.x_stride * ly + boff;
  const float* Ts = a.T + a.t_stride * ly + boff;
  const float* dTs = a.dT + a.dt_stride * ly + boff;
  const float* gs = a.gx + a.gx_stride * (ly + 1) + boff;
  float acc[3][2][4][4];
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) zero_acc(acc[p][mt]);
  __nv_bfloat16* my = &tiles[warp][0][0];
  const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
  const int srow = lane >> 1, shalf = lane & 1;      // staging: lane -> (tile row, 16-channel half)
  // ldmatrix.x4.trans lane addresses.  A (m-tile mt): matrices {k 0-7, m +0}, {k 0-7, m +8}, {k 8-15, m +0}, {k 8-15, m +8};
  // B (n-tile pair np): {k 0-7, n +0}, {k 8-15, n +0}, {k 0-7, n +8}, {k 8-15, n +8}
  const uint32_t a_lane = (uint32_t)(((lane & 7) + 8 * ((lane >> 4) & 1)) * WG_PITCH + 8 * ((lane >> 3) & 1)) * 2;
  const uint32_t b_lane = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * WG_PITCH + 8 * ((lane >> 4) & 1)) * 2;
  constexpr uint32_t TILE_B = 16 * WG_PITCH * 2;
  for (int t0 = row_lo + warp * 16; t0 < row_hi; t0 += 64) {
    const int r = t0 + srow;
    const bool ok = r < row_hi;
    auto stage = [&](int slot, const float* src, int rr, bool valid, bool relu) {
      float4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = valid ? *reinterpret_cast<const float4*>(src + (int64_t)rr * C + shalf * 16 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t w[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (relu) {
          w[2 * q] = bf2(fmaxf(v[q].x, 0.f), fmaxf(v[q].y, 0.f));
          w[2 * q + 1] = bf2(fmaxf(v[q].z, 0.f), fmaxf(v[q].w, 0.f));
        } else {
          w[2 * q] = bf2(v[q].x, v[q].y);
          w[2 * q + 1] = bf2(v[q].z, v[q].w);
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(my + slot * 16 * WG_PITCH + srow * WG_PITCH + shalf * 16);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    };
    stage(0, dTs, r, ok, false);
    stage(1, xs, r, ok, true);
    stage(2, xs, r - d, ok, true);
    stage(3, gs, r, ok, false);
    stage(4, Ts, r, ok, true);
    __syncwarp();
    uint32_t adT[2][4], agx[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      ldsm4t(my_s + 0 * TILE_B + a_lane + mt * 32, adT[mt]);
      ldsm4t(my_s + 3 * TILE_B + a_lane + mt * 32, agx[mt]);
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bx[4], bxp[4], bt[4];
      ldsm4t(my_s + 1 * TILE_B + b_lane + np * 32, bx);
      ldsm4t(my_s + 2 * TILE_B + b_lane + np * 32, bxp);
      ldsm4t(my_s + 4 * TILE_B + b_lane + np * 32, bt);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t b1[2] = {bx[2 * q], bx[2 * q + 1]}, b0[2] = {bxp[2 * q], bxp[2 * q + 1]}, bd[2] = {bt[2 * q], bt[2 * q + 1]};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[0][mt][2 * np + q], adT[mt], b0);
          mma16816(acc[1][mt][2 * np + q], adT[mt], b1);
          mma16816(acc[2][mt][2 * np + q], agx[mt], bd);
        }
      }
    }
    __syncwarp();
  }
  // CTA reduction in shared memory, then one atomic per element and CTA
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int m = 16 * mt + g + 8 * (j >> 1), n = 8 * nt + 2 * tg + (j & 1);
          atomicAdd(&red[p][m * C + n], acc[p][mt][nt][j]);
        }
  __syncthreads();
  for (int e = threadIdx.x; e < C * C; e += 128) {
    const int m = e / C, n = e % C;
    atomicAdd(G + a.w_dil[ly] + (int64_t)(m * C + n) * 2 + 0, red[0][e]);      // [de = m][r = n][tap 0]
    atomicAdd(G + a.w_dil[ly] + (int64_t)(m * C + n) * 2 + 1, red[1][e]);
    atomicAdd(G + a.w_dense[ly] + e, red[2][e]);                               // [r = m][de = n]
  }
}

// ------------------------------------------------------------------------------------------------ bottleneck + AvgPool
// grid (frames, B), 256 threads: xbar = mean over the frame's rows, enc = W_b xbar (+ bias)
__global__ void __launch_bounds__(256) enc_pool_bottleneck_kernel(const float* __restrict__ xN, const float* __restrict__ w_b,
                                                                  const float* __restrict__ bias, float* __restrict__ xbar,
                                                                  float* __restrict__ enc, int L, int tw, int pool, int frames, int BW) {
  __shared__ float part[8][C];
  __shared__ float xm[C];
  const int f = blockIdx.x, b = blockIdx.y, c = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const float* src = xN + ((int64_t)b * L + tw + (int64_t)f * pool) * C;
  float s = 0.f;
  for (int j = sub; j < pool; j += 8) s += src[(int64_t)j * C + c];
  part[sub][c] = s;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    t /= (float)pool;
    xm[threadIdx.x] = t;
    xbar[((int64_t)b * frames + f) * C + threadIdx.x] = t;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < BW; o += 256) {
    float t = bias ? bias[o] : 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) t += w_b[o * C + k] * xm[k];
    enc[((int64_t)b * frames + f) * BW + o] = t;
  }
}
// v[b, f, c] = (1 / pool) sum_o W_b[o][c] genc[b, f, o]
__global__ void __launch_bounds__(256) enc_bottleneck_dgrad_kernel(const float* __restrict__ genc, const float* __restrict__ w_b,
                                                                   float* __restrict__ v, int pool, int frames, int BW) {
  __shared__ float part[8][C];
  const int f = blockIdx.x, b = blockIdx.y, c = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const float* ge = genc + ((int64_t)b * frames + f) * BW;
  float s = 0.f;
  for (int o = sub; o < BW; o += 8) s += w_b[o * C + c] * ge[o];
  part[sub][c] = s;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    v[((int64_t)b * frames + f) * C + threadIdx.x] = t / (float)pool;
  }
}
// gxN[b, t, :] = v[b, (t - tw) / pool, :] inside [tw, tw + frames pool), zero elsewhere
__global__ void __launch_bounds__(256) enc_gx_broadcast_kernel(const float* __restrict__ v, float* __restrict__ gxN, int L, int tw, int pool,
                                                               int frames) {
  const int b = blockIdx.y;
  const int64_t n4 = (int64_t)L * (C / 4);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(e / (C / 4)), q = (int)(e % (C / 4));
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= tw && t < tw + frames * pool) val = *reinterpret_cast<const float4*>(v + ((int64_t)b * frames + (t - tw) / pool) * C + q * 4);
    *reinterpret_cast<float4*>(gxN + ((int64_t)b * L + t) * C + q * 4) = val;
  }
}
// dW_b[o][c] += sum_{b, f} genc[b, f, o] xbar[b, f, c];  db[o] += sum genc[b, f, o].   grid (BW / 8), 256 threads = 8 o x 32 c
__global__ void __launch_bounds__(256) enc_bottleneck_wgrad_kernel(const float* __restrict__ genc, const float* __restrict__ xbar,
                                                                   float* __restrict__ dW, float* __restrict__ db, int M, int BW) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), c = threadIdx.x & 31;
  if (o >= BW) return;
  float s = 0.f, sb = 0.f;
  for (int m = 0; m < M; ++m) {
    const float ge = genc[(int64_t)m * BW + o];
    s += ge * xbar[(int64_t)m * C + c];
    sb += ge;
  }
  dW[o * C + c] += s;
  if (db && c == 0) db[o] += sb;
}

// ------------------------------------------------------------------------------------------------ conditioning convs (grouped)
// CTA: 16 rows x 64 columns, 256 threads = (row, 4 columns cq + 16 j); K in steps of 64.  Operand tiles are [*][64 + 4] floats
// (pitch 68: 16-byte aligned rows, conflict-free 128-bit reads across consecutive rows).
constexpr int CP = 68;
__device__ __forceinline__ void dot_tile(const float* As, const float* Bs, int row, int cq, float (&acc)[4]) {
#pragma unroll 4
  for (int k = 0; k < 64; k += 4) {
    const float4 av = *reinterpret_cast<const float4*>(As + row * CP + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bv = *reinterpret_cast<const float4*>(Bs + (cq + 16 * j) * CP + k);
      acc[j] += av.x * bv.x + av.y * bv.y + av.z * bv.z + av.w * bv.w;
    }
  }
}
__global__ void __launch_bounds__(256) cond_tables_kernel(const float* __restrict__ enc, const float* __restrict__ cp,
                                                          const __grid_constant__ CondBlocks blocks, int M, int K) {
  __shared__ __align__(16) float As[16 * CP], Bs[64 * CP];
  const CondBlock& bk = blocks.blk[blockIdx.y];
  const int row = threadIdx.x >> 4, cq = threadIdx.x & 15, m0 = blockIdx.x * 16;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 64) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int r = e >> 6, k = e & 63;
      As[r * CP + k] = (m0 + r < M) ? enc[(int64_t)(m0 + r) * K + k0 + k] : 0.f;
    }
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
      const int c = e >> 6, k = e & 63;
      Bs[c * CP + k] = c < bk.ncols ? cp[bk.w_off + (int64_t)c * K + k0 + k] : 0.f;
    }
    __syncthreads();
    dot_tile(As, Bs, row, cq, acc);
    __syncthreads();
  }
  if (m0 + row < M) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cq + 16 * j;
      if (c < bk.ncols) bk.out[(int64_t)(m0 + row) * bk.out_stride + c] = acc[j] + cp[bk.b_off + c];
    }
  }
}
// genc[m][k] = sum over blocks j, columns c:  cg_j[m][c] W_j[c][k].   grid (ceil(M / 16), K / 64); per block the reduction
// index c is processed in steps of 64: As = cg tile [16][64 c], Bs = W^T tile [64 k][64 c]
__global__ void __launch_bounds__(256) cond_bwd_kernel(const float* __restrict__ cp, const __grid_constant__ CondBlocks blocks,
                                                       float* __restrict__ genc, int M, int K) {
  __shared__ __align__(16) float As[16 * CP], Bs[64 * CP];
  const int row = threadIdx.x >> 4, cq = threadIdx.x & 15, m0 = blockIdx.x * 16, k0 = blockIdx.y * 64;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int jb = blockIdx.z; jb < blocks.n; jb += gridDim.z) {      // (the blocks are spread over grid.z: partial sums meet in atomics)
    const CondBlock& bk = blocks.blk[jb];
    for (int c0 = 0; c0 < bk.ncols; c0 += 64) {
      for (int e = threadIdx.x; e < 16 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        As[r * CP + c] = (m0 + r < M && c0 + c < bk.ncols) ? bk.cg[(int64_t)(m0 + r) * bk.cg_stride + c0 + c] : 0.f;
      }
      for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        const int c = e >> 6, k = e & 63;      // coalesced along k in global memory, transposed into [k][c]
        Bs[k * CP + c] = (c0 + c < bk.ncols) ? cp[bk.w_off + (int64_t)(c0 + c) * K + k0 + k] : 0.f;
      }
      __syncthreads();
      dot_tile(As, Bs, row, cq, acc);
      __syncthreads();
    }
  }
  if (m0 + row < M) {
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(genc + (int64_t)(m0 + row) * K + k0 + cq + 16 * j, acc[j]);
  }
}

// two 8-warp CTAs per SM; every warp walks its 16-row chunks with the weight fragments resident in registers
inline int chunk_grid(int n_chunks, int B) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_chunks, ENC_WARPS), std::max(1, 148 * 2 / B))); }

}  // namespace

int launch_enc_fwd_layer(const float* x, float* T, float* x_out, const float* w_dil, const float* w_dense, int B, int L, int d, int s_out,
                         cudaStream_t s) {
  const int n_chunks = (int)ceil_div(L - s_out, 16);
  if (n_chunks <= 0) return WN_OK;
  WN_PROF("enc_fwd_layer", s);
  WN_CHECK_CUDA(launch_pdl(enc_fwd_layer_kernel, dim3((unsigned)chunk_grid(n_chunks, B), (unsigned)B), dim3(32 * ENC_WARPS), 0, s, x, T, x_out, w_dil,
                           w_dense, L, d, s_out, n_chunks));
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_enc_bwd_layer(const float* gx_next, const float* T, const float* x, float* gx, float* dT, const float* w_dil,
                         const float* w_dense, int B, int L, int d, int s_out, cudaStream_t s) {
  const int n_chunks = (int)ceil_div(L - (s_out - d), 16);
  if (n_chunks <= 0) return WN_OK;
  WN_PROF("enc_bwd_layer", s);
  WN_CHECK_CUDA(launch_pdl(enc_bwd_layer_kernel, dim3((unsigned)chunk_grid(n_chunks, B), (unsigned)B), dim3(32 * ENC_WARPS), 0, s, gx_next, T, x, gx, dT,
                           w_dil, w_dense, L, d, s_out, n_chunks));
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_enc_wgrad(const EncWgradArgs& a, float* G, cudaStream_t s) {
  WN_REQUIRE(a.N <= 64, WN_ERR_UNSUPPORTED, "encoder weight-gradient kernel: %d layers (max 64)", a.N);
  int rows = 0;
  for (int i = 0; i < a.N; ++i) rows = std::max(rows, a.L - a.s_out[i]);
  if (rows <= 0) return WN_OK;
  WN_PROF("enc_wgrad", s);
  enc_wgrad_kernel<<<dim3((unsigned)ceil_div(rows, WG_ROWS), (unsigned)a.N, (unsigned)a.B), 128, 0, s>>>(a, G);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_enc_pool_bottleneck(const float* xN, const float* w_b, const float* bias, float* xbar, float* enc, int B, int L, int tw, int pool,
                               int frames, int BW, cudaStream_t s) {
  WN_PROF("enc_pool_bottleneck", s);
  enc_pool_bottleneck_kernel<<<dim3((unsigned)frames, (unsigned)B), 256, 0, s>>>(xN, w_b, bias, xbar, enc, L, tw, pool, frames, BW);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_enc_pool_bottleneck_bwd(const float* genc, const float* xbar, const float* w_b, float* gxN, float* v_scratch, float* dW, float* db,
                                   int B, int L, int tw, int pool, int frames, int BW, cudaStream_t s) {
  WN_PROF("enc_pool_bottleneck_bwd", s);
  enc_bottleneck_dgrad_kernel<<<dim3((unsigned)frames, (unsigned)B), 256, 0, s>>>(genc, w_b, v_scratch, pool, frames, BW);
  WN_CHECK_LAUNCH();
  enc_gx_broadcast_kernel<<<dim3((unsigned)std::min<int64_t>(ceil_div((int64_t)L * (C / 4), 256), 148 * 8), (unsigned)B), 256, 0, s>>>(
      v_scratch, gxN, L, tw, pool, frames);
  WN_CHECK_LAUNCH();
  enc_bottleneck_wgrad_kernel<<<(unsigned)ceil_div(BW, 8), 256, 0, s>>>(genc, xbar, dW, db, B * frames, BW);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_cond_tables(const float* enc, const float* cond_params, const CondBlocks& blocks, int M, int K, cudaStream_t s) {
  WN_REQUIRE(K % 64 == 0 && blocks.n > 0 && blocks.n <= kCondBlocksMax, WN_ERR_INVALID, "launch_cond_tables: K %d, %d blocks", K, blocks.n);
  WN_PROF("cond_tables", s);
  cond_tables_kernel<<<dim3((unsigned)ceil_div(M, 16), (unsigned)blocks.n), 256, 0, s>>>(enc, cond_params, blocks, M, K);
  WN_CHECK_LAUNCH();
  return WN_OK;
}
int launch_cond_bwd(const float* cond_params, const CondBlocks& blocks, float* genc, int M, int K, cudaStream_t s) {
  WN_REQUIRE(K % 64 == 0 && blocks.n > 0 && blocks.n <= kCondBlocksMax, WN_ERR_INVALID, "launch_cond_bwd: K %d, %d blocks", K, blocks.n);
  WN_CHECK_CUDA(cudaMemsetAsync(genc, 0, (size_t)M * K * sizeof(float), s));
  WN_PROF("cond_bwd", s);
  const int nz = std::min(blocks.n, 12);
  cond_bwd_kernel<<<dim3((unsigned)ceil_div(M, 16), (unsigned)(K / 64), (unsigned)nz), 256, 0, s>>>(cond_params, blocks, genc, M, K);
  WN_CHECK_LAUNCH();
  return WN_OK;
}

}  // namespace wn
