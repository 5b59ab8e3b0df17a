// fast_layout.cuh - byte layouts of the bf16 packed-weight image and of the training workspace, plus the
// backward-side descriptor bundle (shared between fast_host.cu and fast_bwd.cu).
#pragma once
#include <cuda.h>

#include <vector>

#include "common.cuh"

namespace wn {

enum { PJ_BF16 = 0, PJ_BF16_T = 1, PJ_F32_T = 2, PJ_COPY = 3 };
struct PackJob {
  int32_t kind, out, in, k, tap, pitch, row0, col0;
  int64_t src;   // float offset into the flat parameter vector
  int64_t dst;   // byte offset into the packed image
};

struct PackLayout {     // byte offsets
  size_t jobs, wc_t, bias_c, bias_fg, bias_d, bias_skip, bias_p1, bias_p2;
  size_t wfg0, wfg1, wd, wscat, p1, p2;          // K-major [out][in] bf16 (forward B operands)
  size_t wfgT0, wfgT1, wdT, wsT, p1T, p2T;       // [in][out] bf16 (data-gradient B operands)
  size_t gen_frag;                               // mma.sync A-fragment image of all weights (generation kernel, fast_gen.cu)
  size_t total;
};
PackLayout pack_layout(const Model& m);
// the flat-vector offsets of the N skip biases (int64) live behind the job table, at pl.jobs + this
inline size_t skip_bias_offs_pos(const Model& m) { return sizeof(PackJob) * (size_t)(16 * m.n_layers + 32); }

struct WsLayout {       // byte offsets
  size_t X, x_stride, XLO, Zcat, H0, H1, X0f;     // XLO: 2 ping-pong buffers of x_stride bytes
  size_t DLG, DH1, DSK, DZcat, DXa, DXb, DFG, DFG2, Zf, DX0f, WGP;   // WGP: per-CTA weight-gradient partial tiles
  size_t total;
};
WsLayout ws_layout(const Model& m, int B, int L);
constexpr int64_t WGP_LAYER_FLOATS = (int64_t)160 * 128 * 192;     // per-layer region of per-CTA weight-gradient tiles
// padded skip row space: rows per batch Wp = L - tw_al, tw_al = floor((L-W)/128)*128
inline int skip_tw_al(const Model& m, int L) { return ((m.rf - 1) / 128) * 128; }
inline int skip_wp(const Model& m, int L) { return L - skip_tw_al(m, L); }

int tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems, uint32_t box_rows);
int tmap_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batches, uint64_t pitch_elems,
            uint64_t batch_pitch_elems, uint32_t box_rows);

struct BwdLayerMaps {
  CUtensorMap x;              // x_i (64, L, B)
  CUtensorMap w0, w1;         // W_fg taps [128][64]        (recompute)
  CUtensorMap wdT;            // [64 d][64 r]   box {64,64}
  CUtensorMap wfgT0, wfgT1;   // [64 r][128 o]  box {64,64}
};
struct BwdMaps {
  std::vector<BwdLayerMaps> layer;
  CUtensorMap dlg, h0, h1, dh1, dsk;     // (256, Wp, B) box {64,128,1}   (padded skip row space)
  CUtensorMap zcat, dzcat;               // (64 N, Wp, B)
  CUtensorMap p1T, p2T;                  // [256][256] box {64,256}
  CUtensorMap wsTcat;                    // [64 N rows (layer, d)][256 s] box {64,256}
  CUtensorMap dxa, dxb, dfg, dfg2, zf;   // (64|128, L, B); dfg2: second dF|dG buffer of conditioned models
  CUtensorMap dqa, dqb;                  // block_bwd6: Q_i ping-pong buffers (64, L, B) inside the dFG region
};
int build_bwd_maps(const Model& m, const PackLayout& pl, const WsLayout& wl, int B, int L, const uint8_t* P, uint8_t* Wp,
                   const std::vector<CUtensorMap>& xm, BwdMaps* out);
// conditioning helpers of the autoencoder's bf16 decoder (fast_bwd.cu)
int launch_frame_sum_bf16(const void* src, int64_t rows_per_batch, int pitch, int c0, int B, int t0, int len, int frames, float* out,
                          int out_pitch, int dd, cudaStream_t s);
int launch_cond_table(const float* raw, const float* bias, int64_t n_rows, int dd, float* out, int64_t out_stride, cudaStream_t s);
// fp32 table rows of 128 ([filter 64 | gate 64]) -> rows in the block kernels' per-thread order (Model::cond_fg16)
int launch_cond_pack16(const float* tab, void* out, int64_t n_rows, int frames, int layers, cudaStream_t s);
int launch_add_bias_rows(const float* raw, const float* bias, int64_t n_rows, int C, float* out, cudaStream_t s);
struct SkipHeadMaps;
struct SkipHeadParams;
int head_forward_generic(const Model& m, const BwdMaps& M, const SkipHeadMaps& H, const SkipHeadParams& hp, int B, int L, cudaStream_t s);
int fast_backward_impl(Model& m, const BwdMaps& maps, int B, int L, const float* d_x, const int64_t* d_idx, const void* d_packed,
                       void* d_ws, float* d_dlogits, float* d_grads, cudaStream_t s);

}  // namespace wn
