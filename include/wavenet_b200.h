/* wavenet_b200.h - C ABI of libwavenet_b200.so
 *
 * B200 (sm_100a) implementation of the WaveNet dilated-causal-convolution hot path of
 * deep-art-project/Music.  The reference is pure Python on PyTorch and has no FFI of its own;
 * the boundary it exposes is its Python surface.  Each entry point below cites the reference
 * interface (file:line, relative to the reference root) whose arithmetic it replaces; the
 * Python layer in music_b200/ re-creates that surface on top of these calls.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types in any signature.
 *  - every function returns 0 on success, a negative wn_status otherwise;
 *    wn_last_error() returns a thread-local message for the last failure.
 *  - every pointer named d_* is a DEVICE pointer owned by the caller; the library never
 *    allocates device memory: sizes come from the *_bytes queries.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it and
 *    the library never synchronises.
 *  - there is no CPU fallback: without an sm_100 device wn_init() fails and nothing else runs.
 *
 * Layouts
 *  - parameters: ONE flat fp32 vector in the reference state_dict order
 *    (wavenet/model.py:46-84): causal_layer.weight (R,Q,2) [.bias (R)], then per layer
 *    filter (D,R,2), gate (D,R,2), dense (R,D,1), skip (S,D,1) [each followed by its bias],
 *    post_process_1 (S,S,1), post_process_2 (Q,S,1).  Gradients use the same flat layout.
 *  - dense input  : (B,Q,L) fp32 contiguous, as `wavenet.forward` takes it (model.py:86-98).
 *  - index input  : (B,L) int64 mu-law codes == a true one-hot (B,Q,L) input.
 *  - logits       : (B,Q,W) fp32 contiguous, W = L - rf + 1: the output of post_process_2
 *    (model.py:138), i.e. the tensor `forward` views as (-1,Q) before its softmax (:142-144).
 */
#ifndef WAVENET_B200_H_
#define WAVENET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  WN_OK = 0,
  WN_ERR_INVALID = -1,      /* bad argument */
  WN_ERR_CUDA = -2,         /* CUDA runtime / driver error */
  WN_ERR_UNSUPPORTED = -3,  /* valid request this build cannot serve (e.g. not sm_100) */
  WN_ERR_SHAPE = -4         /* "wave sample not long enough" (model.py:100-101) and friends */
} wn_status;

/* arithmetic mode of the conv stack */
enum { WN_MODE_FP32 = 0,    /* fp32 SIMT check mode (1e-4 parity) */
       WN_MODE_BF16 = 1 };  /* tcgen05/TMEM bf16 tensor-core path, fp32 accumulate */

/* softmax row order + loss definition */
enum { WN_ROWS_REFERENCE = 0,   /* rows = flat Q-chunks of the (B,Q,W) buffer (model.py:142-143) and
                                   CrossEntropyLoss applied to probabilities (train.py:146,179) */
       WN_ROWS_CORRECTED = 1 }; /* rows = time steps, single softmax (extension) */

/* which vector a block pushes into its generation queue */
enum { WN_PUSH_OUTPUT = 0,      /* block output: what fast_generate.py:128-129 does */
       WN_PUSH_INPUT = 1 };     /* block input: consistent with the full forward (extension) */

typedef struct {
  int32_t n_layers;
  const int32_t* dilations;        /* host array, n_layers entries */
  int32_t residual_channels;       /* R */
  int32_t dilation_channels;       /* D */
  int32_t skip_channels;           /* S */
  int32_t quantization_channels;   /* Q */
  int32_t use_bias;
  int32_t filter_width;            /* must be 2 (the only width the reference ships or tests) */
} wn_config;

typedef struct wn_model wn_model;  /* host-side plan: shapes, offsets, cached TMA descriptors */

/* ---- lifecycle ---------------------------------------------------------------------- */
int wn_version(void);
const char* wn_last_error(void);
/* Selects `device`, checks compute capability 10.x, resolves the driver entry points used for
 * TMA descriptors.  Must be called once per process before any launch. */
int wn_init(int device);

/* ---- mu-law codec : wavenet/audio_func.py:5-22 (encode), :24-39 (decode) -------------- */
/* bit-exact with the reference's torch-CPU fp32 arithmetic via threshold / value tables
 * (d_thresholds: Q floats, [0] unused; d_values: Q floats) evaluated once on the host. */
int wn_mulaw_encode(const float* d_audio, int64_t n, int32_t q, const float* d_thresholds,
                    int64_t* d_codes, void* stream);
int wn_mulaw_decode(const int64_t* d_codes, int64_t n, int32_t q, const float* d_values,
                    float* d_audio, void* stream);

/* ---- loader one-hot : one_hot_encode, wavenet/faster_audio_data.py:62-83, on the device ------------------------------
 * d_codes (B,T) int64 -> d_out (B,Q,T) fp32.  transpose = 0 reproduces the reference bit for bit: it builds a (T,Q) one-hot
 * and RESHAPES it to (Q,T) (:77-81), which is not a one-hot; transpose = 1 gives the true one-hot.  Only the codes cross PCIe. */
int wn_onehot_encode(const int64_t* d_codes, int32_t B, int32_t T, int32_t q, int32_t transpose, float* d_out, void* stream);

/* ---- model plan : wavenet.__init__ / calc_receptive_field, wavenet/model.py:8-84 ------ */
int wn_model_create(const wn_config* cfg, wn_model** out);
int wn_model_destroy(wn_model* m);
int64_t wn_model_param_count(const wn_model* m);
/* Offset (in floats) of residual block `layer`'s first parameter in the flat vector: [offset, param_count) holds blocks
 * layer .. N-1 and the two post-processing convs (layer == N: only those). */
int64_t wn_model_layer_offset(const wn_model* m, int32_t layer);
/* Overlapped data-parallel gradient exchange (replaces what nn.DataParallel's reducer does, wavenet/train.py:121): when set, the
 * bf16 wn_backward records `cuda_event` (a cudaEvent_t) as soon as the gradients in [wn_model_layer_offset(layer), param_count)
 * are final - the blocks below `layer` are still running - so the caller can start reducing that bucket on another stream.
 * layer < 0 or a NULL event: off.  Models with bias and the fp32 mode ignore it (the event is then recorded at the end). */
int wn_backward_set_split(wn_model* m, int32_t layer, void* cuda_event);
int32_t wn_model_receptive_field(const wn_model* m);          /* model.py:43-44 */
/* 1 if this build serves the model's shape with: what = 0 the bf16 tcgen05 training path (residual, dilation <= 64 channels,
 * zero-padded to 64; skip = quantization = 256), what = 1 the half-precision generation kernel (exactly 64/64/256/256);
 * 0 otherwise (such shapes run in WN_MODE_FP32).  The Python layer's mode="auto" asks this. */
int32_t wn_model_supports(const wn_model* m, int32_t what);

/* kernel-side weight images derived from the flat fp32 parameters (refresh after every
 * optimizer step).  fp32 mode: transposed fp32 copies; bf16 mode: bf16 K-major matrices that
 * TMA loads straight into the UMMA shared-memory layout. */
int wn_packed_bytes(const wn_model* m, int32_t mode, size_t* bytes);
int wn_pack_weights(wn_model* m, int32_t mode, const float* d_params, void* d_packed, void* stream);

/* ---- training forward / backward : wavenet.forward, wavenet/model.py:86-138 ----------- */
int wn_workspace_bytes(const wn_model* m, int32_t mode, int32_t B, int32_t L, size_t* bytes);
/* Exactly one of d_x (dense (B,Q,L) fp32) / d_idx ((B,L) int64) is non-NULL.
 * Writes d_logits (B,Q,W) fp32 and keeps what backward needs inside d_workspace (which must be
 * zero-filled once when allocated and must not be touched between forward and backward). */
int wn_forward(wn_model* m, int32_t mode, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx,
               const void* d_packed, void* d_workspace, float* d_logits, void* stream);
/* Gradient of a scalar loss w.r.t. the flat parameters given d_dlogits (B,Q,W) fp32.
 * d_grads (param_count floats) is OVERWRITTEN.  Same d_x/d_idx/d_workspace as the forward.
 * d_dlogits may be clobbered. */
int wn_backward(wn_model* m, int32_t mode, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx,
                const void* d_packed, void* d_workspace, float* d_dlogits, float* d_grads, void* stream);

/* ---- softmax + loss : model.py:142-144 (nn.Softmax on the (-1,Q) view),
 *                       train.py:146,177-179 (CrossEntropyLoss on the probabilities) ------ */
int wn_softmax_fwd(const float* d_logits, int32_t B, int32_t Q, int32_t W, int32_t rows,
                   float* d_probs /* (B*W,Q) */, void* stream);
int wn_softmax_bwd(const float* d_probs, const float* d_dprobs, int32_t B, int32_t Q, int32_t W,
                   int32_t rows, float* d_dlogits /* (B,Q,W) */, void* stream);
/* Fused: probabilities, loss (mean over the B*W rows, written to d_loss[0]) and, if
 * d_dlogits != NULL, grad_scale * dloss/dlogits.  d_scratch: wn_loss_scratch_bytes(). */
int wn_loss_scratch_bytes(int32_t B, int32_t W, size_t* bytes);
int wn_loss_fwd_bwd(const float* d_logits, const int64_t* d_target /* (B,W) */, int32_t B, int32_t Q,
                    int32_t W, int32_t rows, float grad_scale, float* d_loss, float* d_dlogits,
                    void* d_scratch, void* stream);

/* ---- optimizers : get_optimizer, wavenet/train.py:28-42 (torch.optim defaults) --------- */
int wn_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n, float lr,
                 float beta1, float beta2, float eps, int32_t step, void* stream);
/* wn_adam_step with the step count kept in device memory: *d_step is incremented, then used for the bias corrections.  For steps replayed
 * from a CUDA graph (a captured launch cannot carry a new host scalar); same arithmetic as wn_adam_step. */
int wn_adam_step_dev(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n, float lr,
                     float beta1, float beta2, float eps, int32_t* d_step, void* stream);
int wn_sgd_step(float* d_params, const float* d_grads, float* d_momentum_buf, int64_t n, float lr,
                float momentum, int32_t first_step, void* stream);
int wn_rmsprop_step(float* d_params, const float* d_grads, float* d_square_avg, float* d_momentum_buf,
                    int64_t n, float lr, float alpha, float eps, float momentum, void* stream);

/* ---- incremental generation : fast_generate.predict_next, wavenet/fast_generate.py:13-141 -- */
/* Per-stream state: ring buffers replacing the reference's shift-copied queues. */
int wn_gen_state_bytes(const wn_model* m, int32_t mode, int32_t n_streams, size_t* bytes);
/* Prime branch (:29-65): d_prime_idx (n_streams, rf) int64 codes of the priming one-hot piece.
 * Fills the queues and writes the first pick per stream to d_out[n_streams]. */
int wn_gen_prime(wn_model* m, int32_t mode, int32_t n_streams, const int64_t* d_prime_idx,
                 const void* d_packed, void* d_state, void* d_workspace, size_t workspace_bytes,
                 const float* d_uniforms /* n_streams or NULL=greedy */, int64_t* d_out,
                 float* d_logits /* (n_streams,Q) or NULL */, void* stream);
/* Step branch (:66-141), n_steps times per stream, feeding each pick back as the next one-hot
 * note (generate(), :166-172).  d_first_note (n_streams) = the note fed to the first step.
 * d_out is (n_steps, n_streams); d_uniforms (n_steps, n_streams) or NULL for the reference's
 * greedy topk(1); d_logits (n_steps, n_streams, Q) or NULL. */
int wn_gen_steps(wn_model* m, int32_t mode, int32_t n_streams, int32_t n_steps, int32_t push,
                 const int64_t* d_first_note, const void* d_packed, void* d_state,
                 const float* d_uniforms, int64_t* d_out, float* d_logits, void* stream);
/* Queue interchange with the reference's OrderedDict layout ('causal_layer': (1,Q,1) as a code,
 * 'block_k': (1,R,d_k) oldest first): d_queues is (n_streams, sum(d_k), R) fp32 time-major,
 * d_last_note (n_streams) int64. */
int wn_gen_export(const wn_model* m, int32_t mode, int32_t n_streams, const void* d_state,
                  float* d_queues, int64_t* d_last_note, void* stream);
int wn_gen_import(const wn_model* m, int32_t mode, int32_t n_streams, void* d_state,
                  const float* d_queues, const int64_t* d_last_note, void* stream);

/* ---- conditioned generation (extension; SURVEY.md 8f rank 3: incremental generation for the autoencoder's decoder) ----
 * The decoder of wavenet_autoencoder (model1.py:158-225) is a WaveNet stack whose block pre-activations and head receive an
 * additive, per-frame conditioning vector (`_conditon`, :227-247).  With a descriptor installed, wn_forward (fp32),
 * wn_gen_prime and wn_gen_steps (fp32, and the half-precision cluster pipeline of the 64/64/256/256 shape) add  d_fg[stream, frame_i(t), block i, :]  to the [f|g] pre-activations of block i and
 * d_head[stream, frame(t), :]  to post_process_1's output, where frame follows the reference rule on the TOTAL sequence:
 * len % frames == 0 ? t_local / (len / frames) : t_local % frames, with len = total_len - (first valid index of that tensor).
 * gate_first != 0 means channels [0,D) of d_fg belong to the gate and [D,2D) to the filter (the autoencoder's split, :188-192).
 * Pass NULL to remove the descriptor.  The tables are read at call time; they must outlive the calls. */
typedef struct {
  const float* d_fg;      /* (n_streams, frames, n_layers, 2D) fp32 */
  const float* d_head;    /* (n_streams, frames, S) fp32 */
  int32_t frames, total_len, gate_first;
} wn_gen_cond;
int wn_set_conditioning(wn_model* m, const wn_gen_cond* cond);
/* Builds those tables from an autoencoder: d_encoding (n_streams, frames, BW) channels-last, d_cond = its conditioning
 * convs (layout of wn_ae_forward).  d_workspace: wn_ae_workspace_bytes(a, n_streams, L >= rf) bytes. */
typedef struct wn_ae wn_ae;
int wn_ae_cond_tables(wn_ae* a, int32_t n_streams, int32_t frames, const float* d_encoding, const float* d_cond,
                      void* d_workspace, float* d_fg, float* d_head, void* stream);

/* ---- autoencoder : wavenet_autoencoder.forward, wavenet_autoencoder/model1.py:137-268 (fp32 check mode) ----
 * Parameters: one flat fp32 vector in the reference state_dict order (model1.py:55-58: en_dilation_layer_stack.*,
 * en_dense_layer_stack.*, de_dilation_layer_stack.{3i,3i+1,3i+2}, en_causal_layer, bottleneck_layer, de_causal_layer,
 * connection_1, connection_2).  The N+1 conditioning convs the reference creates at random on every call
 * (:178-179, :216-217; always biased) are a second flat vector d_cond: N x [(2Dd,BW,1) weight, (2Dd) bias] then
 * [(Sd,BW,1), (Sd)].  d_logits is (B,Q,W) = the output of connection_2 (:221); d_encoding, if non-NULL, receives the
 * pooled encoding channels-last (B, frames, BW), frames = W / pool. */
typedef struct {
  int32_t n_layers;
  const int32_t* dilations;
  int32_t quantization_channel;
  int32_t en_residual_channel, en_dilation_channel, en_bottleneck_width, en_pool_kernel_size;
  int32_t de_residual_channel, de_dilation_channel, de_skip_channel;
  int32_t use_bias;
  int32_t filter_width;            /* must be 2 */
  int32_t mode;                    /* 0: fp32 check mode everywhere; 1: the conditioned decoder (model1.py:158-247, 89 % of the FLOPs)
                                      on the bf16 tcgen05 WaveNet kernels, encoder in fp32.  Mode 1 needs decoder residual, dilation <= 64,
                                      skip 256 or 512, quantization 256 channels, use_bias = 0 (the shipped model_params.json qualifies). */
} wn_ae_config;
int wn_ae_create(const wn_ae_config* cfg, wn_ae** out);
int wn_ae_destroy(wn_ae* a);
int64_t wn_ae_param_count(const wn_ae* a);
int64_t wn_ae_cond_param_count(const wn_ae* a);
int32_t wn_ae_receptive_field(const wn_ae* a);
int wn_ae_workspace_bytes(const wn_ae* a, int32_t B, int32_t L, size_t* bytes);
int wn_ae_forward(wn_ae* a, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx, const float* d_params,
                  const float* d_cond, void* d_workspace, float* d_logits, float* d_encoding, void* stream);
/* Training (what autograd does for wavenet_autoencoder/train.py's loss.backward() through model1.py:137-268):
 * wn_ae_forward_train is wn_ae_forward that keeps every layer's input and pre-activation in the (larger) training
 * workspace; wn_ae_backward consumes that workspace and d loss / d logits (B,Q,W) and writes d_grads (layout of
 * d_params) and, if non-NULL, d_cond_grads (layout of d_cond; the reference discards these convs, so this is an
 * extension for trainable conditioning). */
int wn_ae_train_workspace_bytes(const wn_ae* a, int32_t B, int32_t L, size_t* bytes);
int wn_ae_forward_train(wn_ae* a, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx, const float* d_params,
                        const float* d_cond, void* d_workspace, float* d_logits, float* d_encoding, void* stream);
int wn_ae_backward(wn_ae* a, int32_t B, int32_t L, const float* d_x, const int64_t* d_idx, const float* d_params,
                   const float* d_cond, void* d_workspace, const float* d_dlogits, float* d_grads, float* d_cond_grads,
                   void* stream);

/* ---- instrumentation (bench.py): kernel-launch counter and per-kernel CUDA-event profiler ---------- */
uint64_t wn_launch_count(void);                 /* kernels launched by this library so far */
int wn_profile_enable(int32_t on);              /* bracket every launch with CUDA events on its stream */
int wn_profile_report(char* buf, size_t cap);   /* sync; "name count total_ms" lines, longest first; clears */

/* timing experiments (WN_TS=1 in the environment): 16 clock64 stamps per tile of CTA 0 of the fused block backward;
 * n < 0: -n stamps of the pipelined generation kernel (16 per step of CTA 1, group 0) */
int wn_debug_ts(long long* h_buf, int32_t n);

/* L2 -> SM read-bandwidth probe (bench.py): n_ctas CTAs each stream the L2-resident buffer `iters` times with 128-bit loads,
 * like the generation kernel's CTAs stream the shared weight image.  bytes * n_ctas * iters / time = achieved L2 read rate. */
int wn_bench_l2_read(const void* d_buf, int64_t bytes, int32_t n_ctas, int32_t iters, void* d_sink, void* stream);

/* ---- self tests of the tcgen05 / TMA building blocks (used by tests/, not by the product) -- */
/* Runs D = A*B^T tiles through TMA -> UMMA -> TMEM -> registers for the operand layouts the
 * kernels rely on; writes max |err| vs. an in-kernel fp32 SIMT product to h_maxerr[case]. */
int wn_selftest_umma(float* h_maxerr, int32_t n_cases, void* stream);
/* Test hook for the element-wise gradient parity test at the benchmarked shape: gives the fp32 check-mode workspace of a
 * forward on the same input (d_ws_fp32) the ReLU masks of the bf16 forward (d_ws_bf16) - model.py:135,137 - by setting the
 * sign of the stored fp32 pre-activations; magnitudes are untouched.  wn_backward(fp32) then differentiates through the
 * bf16 run's masks, so that what remains between the two gradients is kernel arithmetic, not mask flips. */
int wn_test_impose_relu_masks(const wn_model* m, int32_t B, int32_t L, const void* d_ws_bf16, void* d_ws_fp32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVENET_B200_H_ */
