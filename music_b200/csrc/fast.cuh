// fast.cuh - interface of the bf16 tensor-core path (tcgen05 / TMEM / TMA), fast_*.cu.
#pragma once
#include "common.cuh"

namespace wn {

int fast_init();                         // resolve cuTensorMapEncodeTiled, set smem attributes
bool fast_supported(const Model& m);     // shapes the tcgen05 training kernels serve: R, D <= 64 (zero-padded to 64), S = Q = 256
bool fast_gen_supported(const Model& m); // shapes the half-precision generation kernel serves: exactly 64/64/256/256
void fast_release(Model& m);
int fast_packed_bytes(const Model& m, size_t* bytes);
int fast_pack(Model& m, const float* d_params, void* d_packed, cudaStream_t s);
int fast_workspace_bytes(const Model& m, int B, int L, size_t* bytes);
int fast_forward(Model& m, int B, int L, const float* d_x, const int64_t* d_idx, const void* d_packed, void* d_ws,
                 float* d_logits, cudaStream_t s);
int fast_backward(Model& m, int B, int L, const float* d_x, const int64_t* d_idx, const void* d_packed, void* d_ws,
                  float* d_dlogits, float* d_grads, cudaStream_t s);

int fast_gen_steps(Model& m, int n_streams, int n_steps, int push, const int64_t* d_first_note, const void* d_packed,
                   void* d_state, const float* d_uniforms, int64_t* d_out, float* d_logits, cudaStream_t s,
                   const wn_gen_cond* cond = nullptr);      // cond: optional conditioning tables (wn_set_conditioning)
size_t fast_gen_frag_bytes(const Model& m);                         // A-fragment image appended to the packed weights
int fast_gen_pack(const Model& m, const float* d_params, uint8_t* P, cudaStream_t s);   // builds it (fp16) from the fp32 weights
int fast_gen_debug_ts(long long* h_buf, int n);                     // WN_TS=1: clock64 stamps of gen_pipe_kernel (timing experiments)
int fast_selftest(float* h_maxerr, int n_cases, cudaStream_t s);

}  // namespace wn
