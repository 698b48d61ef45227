/*
 * matgcn.h - C ABI of the B200 (sm_100a) Multi-ATGCN recurrent graph-convolution library.
 *
 * This is the drop-in boundary for the one hot path this repository accelerates: the
 * AGCRN-style cell of the reference's
 *   libcity/model/traffic_flow_prediction/MultiATGCN.py   ("MA.py" below)
 * The reference has no native code and no FFI; its "operator API" for
 * this path is the set of nn.Module.forward methods cited on each entry point, which autograd
 * differentiates.  Each forward entry point therefore has a hand-written backward twin.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller; the library never allocates or frees
 *    caller-visible memory; scratch comes in through explicit workspace arguments whose sizes
 *    the *_bytes() queries return;
 *  - all tensors are float32, dense, row-major with the shapes written next to each argument;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *  - return value 0 = ok, negative = argument/launch error (text via matgcn_last_error());
 *  - no global mutable state except the thread-local error string; one call at a time per stream.
 *
 * Device layouts ("node-major"): activations are [T, N, B, C] (time, node, batch, channel) so that
 * one time step is an [N, B*C] matrix whose rows the support propagation contracts over.
 * Base matrices are [Kp, N, ldm] with ldm >= N (row pitch; columns >= N are never read).
 * K = Kp + 1 counts the implicit identity support T_0 = I.
 */
#ifndef MATGCN_H_
#define MATGCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MATGCN_ABI_VERSION 4

/* flags of the contraction-heavy entry points */
#define MATGCN_FLAG_EXACT 0 /* fp32 FFMA kernels: 1e-4 parity with the reference */
#define MATGCN_FLAG_TF32 1  /* fast mode: contractions on tcgen05 tensor cores as TF32 (fp32 storage and accumulation) */
#define MATGCN_FLAG_BF16 2  /* with TF32: the support-propagation contractions read bf16 twins of their operands (kind::f16 MMAs) */

/* ABI version of the loaded library (compare with MATGCN_ABI_VERSION). */
int matgcn_abi_version(void);

/* Last error text of the calling thread ("" if none). */
const char* matgcn_last_error(void);

/* Number of kernels this library has launched in this process (monotonic counter). */
unsigned long long matgcn_launch_count(void);

/* Number of those launches that went to the tcgen05/TMA tensor-core kernel. */
unsigned long long matgcn_tc_launch_count(void);

/* -------------------------------------------------------------------------------------------
 * Adaptive adjacency  A = softmax(relu(L * Rt^T), dim=1)
 * replaces MA.py:80-83 (AGCN.forward): bidirection passes L = Rt = node_emb [N, D];
 * unidirection passes L = node_vec1 [N, D], Rt = node_vec2^T [N, D].
 *   A   [N, ldm] out (columns >= N are zero-filled)
 * ----------------------------------------------------------------------------------------- */
int matgcn_adaptive_adj_fwd(const float* L, const float* Rt, int N, int D, float* A, int ldm, void* stream);

/* Backward of the above.  dA [N, ldm]; dL, dRt [N, D] are overwritten.
 * scratch: N*N floats (holds the pre-softmax gradient). For bidirection the caller adds dL + dRt. */
int matgcn_adaptive_adj_bwd(const float* L, const float* Rt, const float* A, const float* dA, int N, int D,
                            int ldm, float* dL, float* dRt, float* scratch, void* stream);

/* -------------------------------------------------------------------------------------------
 * Per-node weight generation from the embedding-indexed pools
 *   W[n,k,i,o] = c[k] * sum_d E[n,d] * pool[d,k,i,o]        b[n,o] = sum_d E[n,d] * bias_pool[d,o]
 * replaces MA.py:104-105 (and folds the view weights softmax(weights_g) of MA.py:102-103 into W:
 * scaling support k by c[k] equals scaling the weights that multiply its output).
 *   E [N,D]  pool [D,K,I,O]  bias_pool [D,O]  c [K]   ->   W [N,K,I,O]  b [N,O]
 * ----------------------------------------------------------------------------------------- */
int matgcn_nodeweights_fwd(const float* E, const float* pool, const float* bias_pool, const float* c,
                           int N, int D, int K, int I, int O, float* W, float* b, void* stream);

/* Backward: given dW [N,K,I,O], db [N,O] writes dE [N,D], dpool [D,K,I,O], dbias_pool [D,O], dc [K]. */
int matgcn_nodeweights_bwd(const float* E, const float* pool, const float* bias_pool, const float* c,
                           const float* dW, const float* db, int N, int D, int K, int I, int O,
                           float* dE, float* dpool, float* dbias_pool, float* dc, void* stream);

/* The same two operators with the arithmetic of the big products selectable (flags as for the layer entry points):
 * MATGCN_FLAG_TF32 runs E x pool, dW x pool^T and E^T x dW on the tensor-core engine. */
int matgcn_nodeweights_fwd_ex(const float* E, const float* pool, const float* bias_pool, const float* c,
                              int N, int D, int K, int I, int O, float* W, float* b, int flags, void* stream);
int matgcn_nodeweights_bwd_ex(const float* E, const float* pool, const float* bias_pool, const float* c,
                              const float* dW, const float* db, int N, int D, int K, int I, int O,
                              float* dE, float* dpool, float* dbias_pool, float* dc, int flags, void* stream);

/* -------------------------------------------------------------------------------------------
 * Support propagation as a standalone operator (the dominant contraction of the path):
 *   P[k, n, col] = sum_m M[k, n, m] * X[m, col]
 * replaces the einsum 'knm,bmc->bknc' of MA.py:106 in node-major layout for the non-identity
 * supports.   M [Kp, N, ldm]   X [N, cols] (cols = B*C)   ->   P [Kp, N, cols]
 * ----------------------------------------------------------------------------------------- */
int matgcn_propagate_fwd(const float* M, int Kp, int N, int ldm, const float* X, int cols, float* P, int flags,
                         void* stream);

/* Persistent recurrence kernels (bf16 mode, rnn_units = 64): each layer's 24-step recurrence (MA.py:200-211) runs as ONE
 * cooperative launch whose phases are separated by grid barriers (csrc/rec_fwd.cuh) instead of four launches per time
 * step.  on = 0 selects one launch per phase.  Returns the previous setting.  Default on (or MATGCN_REC=0 in the environment). */
int matgcn_set_recurrent_kernel(int on);

/* Time-batched residual-cell weight / bias gradients of the layer backward (MA.py:142-150): in the fast modes with rnn_units = 64
 * one pass over the pre-activation gradients (csrc/dr_pass.cuh) replaces four split-K contractions and a column sum.
 * on = 0 keeps the separate launches.  Returns the previous setting.  Default on (or MATGCN_DR_PASS=0 in the environment). */
int matgcn_set_dr_pass(int on);

/* Device timing of the persistent recurrence kernels (measurement only): while on, each of their launches is bracketed by CUDA
 * events on its stream; the read call waits for them and returns the summed milliseconds and launch counts since it was turned on. */
int matgcn_rec_timing(int on);
int matgcn_rec_timing_read(double* fwd_ms, int* fwd_launches, double* bwd_ms, int* bwd_launches);

/* Fused tail of the forward step (candidate contraction + residual GRU cell + mix in one launch; tensor-core engine,
 * rnn_units = 64).  on = 0 selects the three separate contractions.  Returns the previous setting.  Default on
 * (or MATGCN_FUSED_TAIL=0 in the environment). */
int matgcn_set_fused_tail(int on);

/* The same contraction with bf16 twins of M and X (device arrays of __nv_bfloat16), float32 output: the kernel the
 * layer entry points launch for the propagation when MATGCN_FLAG_BF16 is set. */
int matgcn_propagate_fwd_bf16(const void* M16, int Kp, int N, int ldm, const void* X16, int cols, float* P, void* stream);
/* The same launch as a bf16-mode step issues it: only the bf16 twin P16 [Kp, N, cols] of the result is stored. */
int matgcn_propagate_fwd_bf16_twin(const void* M16, int Kp, int N, int ldm, const void* X16, int cols, void* P16, void* stream);

/* Diagnostics: device buffer (int64, >= 8 per tile of CTA 0) that later tensor-core launches fill with clock64()
 * stamps [producer start, mma wait, mma start, mma committed, epilogue wait, epilogue start, epilogue end]; NULL disables. */
int matgcn_debug_set_timeline(long long* buf);
/* Diagnostics: record the timeline only for the tensor-core launch that follows `launches` others. */
int matgcn_debug_set_timeline_skip(int launches);
/* Diagnostics: bit 0 = tensor-core epilogue skips its global stores, bit 1 = skips the smem transpose (results invalid). */
int matgcn_debug_set_mode(int mode);

/* Diagnostics: plain C[M,N] = A*B through the selected engine, for unit tests of the GEMM kernels.
 * a_kc: A element (m,k) at m*lda+k (else k*lda+m); b_kc: B element (k,n) at n*ldb+k (else k*ldb+n).
 * splits > 1 exercises the split-K / atomic epilogue. */
int matgcn_gemm_debug(int a_kc, int b_kc, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                      float* C, int ldc, int splits, int flags, void* stream);

/* Diagnostics: the same product with bf16 operands (device arrays of __nv_bfloat16) on the bf16 tensor-core engine. */
int matgcn_gemm_debug_bf16(int a_kc, int b_kc, int M, int N, int K, const void* A16, int lda, const void* B16, int ldb,
                           float* C, int ldc, int splits, void* stream);

/* -------------------------------------------------------------------------------------------
 * One encoder layer over the whole input window.
 * replaces, for one layer i, the body of ATGRUEncoder.forward MA.py:200-211:
 *   for t: s = ATGRUCell(x_t, s)            MA.py:120-128 with AGCN MA.py:106-108 for gate/update
 *          r = GRUCell(x_t, s)              MA.py:142-150 (residual, shared nn.Linear)
 *          s = mix[t]*s + (1-mix[t])*r      MA.py:208   (mix = sigmoid(weights_gru[i]))
 * with supports and per-node weights hoisted out of the loop.
 *
 * dims: T steps, N nodes, B batch, Cin input channels, H hidden, K supports incl. identity,
 *       I = Cin + H.
 *   x      [T, N, B, Cin] with time stride x_tstride (floats); each step block contiguous
 *   h0     [N, B, H] or NULL (= zeros)
 *   M      [K-1, N, ldm]  base matrices (unscaled T_k, k >= 1)
 *   Wg,bg  [N, K, I, 2H], [N, 2H]   gate weights (z first, r second: MA.py:124)
 *   Wu,bu  [N, K, I, H],  [N, H]    candidate weights
 *   Rgw,Rgb [2H, I], [2H]; Ruw,Rub [H, I], [H]   residual GRU nn.Linear weights (out x in)
 *   mix    [T]  already passed through sigmoid
 *   ws     forward workspace of matgcn_encoder_layer_fwd_ws_bytes(); it holds every saved
 *          activation and must stay untouched until the matching backward call has run.
 * The layer output y[t] = hidden state after step t lives INSIDE the workspace at float offset
 * matgcn_encoder_layer_y_offset() with time stride matgcn_encoder_layer_y_tstride().
 * ----------------------------------------------------------------------------------------- */
size_t matgcn_encoder_layer_fwd_ws_bytes(int T, int N, int B, int Cin, int H, int K);
size_t matgcn_encoder_layer_bwd_ws_bytes(int T, int N, int B, int Cin, int H, int K, int n_adp);
size_t matgcn_encoder_layer_y_offset(int T, int N, int B, int Cin, int H, int K);
size_t matgcn_encoder_layer_y_tstride(int T, int N, int B, int Cin, int H, int K);

/* With MATGCN_FLAG_TF32 | MATGCN_FLAG_BF16 (and H % 8 == 0, ldm % 8 == 0, B >= 8) the propagated slots k >= 1 of PX (wide
 * layers), PH and PZ exist only as their bf16 twins: their fp32 areas in the workspace are left unwritten.  Slot 0 of each
 * (x_t, h_t, z*h) and every other saved activation stay float32 in all modes. */
/* Float offsets of the named workspace slots, for tests/diagnostics.  names: "PX","GX","RX","PH",
 * "PZ","Z","R","HC","H1","Z2","R2","HC2","ZH2".  Returns (size_t)-1 for an unknown name. */
size_t matgcn_encoder_layer_slot_offset(const char* name, int T, int N, int B, int Cin, int H, int K);

int matgcn_encoder_layer_fwd(int T, int N, int B, int Cin, int H, int K, int ldm,
                             const float* x, long long x_tstride, const float* h0, const float* M,
                             const float* Wg, const float* bg, const float* Wu, const float* bu,
                             const float* Rgw, const float* Rgb, const float* Ruw, const float* Rub,
                             const float* mix, float* ws, int flags, void* stream);

/* Backward of one encoder layer (reverse-time BPTT + time-batched parameter gradients).
 *   dy [T, N, B, H] with time stride dy_tstride: gradient w.r.t. every step's output
 *   n_adp: the first n_adp base matrices (the adaptive-adjacency slices) receive a gradient
 * Outputs (all overwritten):
 *   dx [T, N, B, Cin] contiguous; dh0 [N, B, H] or NULL; dM [K-1, N, ldm] (slices >= n_adp zeroed);
 *   dWg, dbg, dWu, dbu, dRgw, dRgb, dRuw, dRub as their forward shapes; dmix [T]
 * ws is the forward workspace (its GX/RX slots are overwritten by the pre-activation gradients),
 * bws a scratch of matgcn_encoder_layer_bwd_ws_bytes(). */
int matgcn_encoder_layer_bwd(int T, int N, int B, int Cin, int H, int K, int ldm, int n_adp,
                             const float* dy, long long dy_tstride, const float* M,
                             const float* Wg, const float* Wu, const float* Rgw, const float* Ruw,
                             const float* mix, float* ws, float* bws,
                             float* dx, float* dh0, float* dM,
                             float* dWg, float* dbg, float* dWu, float* dbu,
                             float* dRgw, float* dRgb, float* dRuw, float* dRub, float* dmix,
                             int flags, void* stream);

/* Layer chaining (fast bf16 mode, Cin == H): an inner ATGRUEncoder layer (MA.py:200-211: layer l+1 consumes layer l's output
 * sequence) reads its input where the previous layer left it instead of copying and re-propagating it.  x must be the previous
 * layer's output view (ws_prev + matgcn_encoder_layer_y_offset, time stride matgcn_encoder_layer_y_tstride) and x16_chain the
 * address of the previous layer's bf16 state twin one step in: (char*)ws_prev + 4*matgcn_encoder_layer_slot_offset("PH16", prev dims)
 * + 2*K*N*B*H.  The previous layer's PH16[t+1, 1..K) = M h_t IS this layer's M x_t; only the last step is propagated, into the
 * spare slots PH16_prev[T, 1..K).  matgcn_encoder_layer_chain_ok tells whether the shape / mode qualifies (the _chained entry
 * points fail otherwise); the previous layer's workspace must stay alive and unchanged until this layer's backward has run. */
int matgcn_encoder_layer_chain_ok(int T, int N, int B, int Cin, int H, int K, int ldm, int flags);
int matgcn_encoder_layer_fwd_chained(int T, int N, int B, int Cin, int H, int K, int ldm, const float* x, long long x_tstride,
                                     void* x16_chain, const float* h0, const float* M, const float* Wg, const float* bg,
                                     const float* Wu, const float* bu, const float* Rgw, const float* Rgb, const float* Ruw,
                                     const float* Rub, const float* mix, float* ws, int flags, void* stream);
int matgcn_encoder_layer_bwd_chained(int T, int N, int B, int Cin, int H, int K, int ldm, int n_adp, const float* dy,
                                     long long dy_tstride, const float* M, const float* Wg, const float* Wu, const float* Rgw,
                                     const float* Ruw, const float* mix, float* ws, float* bws, float* dx, float* dh0, float* dM,
                                     float* dWg, float* dbg, float* dWu, float* dbu, float* dRgw, float* dRgb, float* dRuw,
                                     float* dRub, float* dmix, int flags, const float* x_chain, const void* x16_chain, void* stream);

/* -------------------------------------------------------------------------------------------
 * gcn_off ablation layer: a plain GRU whose nn.Linear weights are shared by all nodes,
 * replaces GRUCell.forward MA.py:142-150 used as the main cell (MA.py:187-192, 204) over the whole window.
 *   x [T, N, B, Cin] (time stride x_tstride), h0 [N, B, H] or NULL, Gw/Gb [2H, I]/[2H] (gate), Uw/Ub [H, I]/[H]
 * The output y[t] lives in the workspace at float offset matgcn_dense_gru_layer_y_offset(), time stride N*B*H.
 * Backward overwrites dx [T,N,B,Cin], dh0 (or NULL), dGw, dGb, dUw, dUb.
 * ----------------------------------------------------------------------------------------- */
size_t matgcn_dense_gru_layer_fwd_ws_bytes(int T, int N, int B, int Cin, int H);
size_t matgcn_dense_gru_layer_bwd_ws_bytes(int T, int N, int B, int Cin, int H);
size_t matgcn_dense_gru_layer_y_offset(int T, int N, int B, int Cin, int H);
int matgcn_dense_gru_layer_fwd(int T, int N, int B, int Cin, int H, const float* x, long long x_tstride, const float* h0,
                               const float* Gw, const float* Gb, const float* Uw, const float* Ub, float* ws, int flags,
                               void* stream);
int matgcn_dense_gru_layer_bwd(int T, int N, int B, int Cin, int H, const float* dy, long long dy_tstride, const float* x,
                               long long x_tstride, const float* Gw, const float* Uw, float* ws, float* bws, float* dx,
                               float* dh0, float* dGw, float* dGb, float* dUw, float* dUb, int flags, void* stream);

/* -------------------------------------------------------------------------------------------
 * Callers either side of the path.
 *
 * f1 - optimiser half of TrafficStateExecutor._train_epoch (libcity/executor/traffic_state_executor.py:413-422):
 *      torch.nn.utils.clip_grad_norm_(parameters, max_norm) (executor:420-421) followed by
 *      torch.optim.Adam(lr, eps, betas, weight_decay).step() (executor:146-147), fused over ONE flat fp32 bucket
 *      (parameters, gradients and both Adam moments contiguous, n elements each, 16-byte aligned).
 *   matgcn_grad_sumsq:    *sumsq = sum_i grad[i]^2   (device double; overwritten)
 *   matgcn_adam_clip_step: g = grad * grad_scale * min(1, max_norm / (|grad_scale| * sqrt(*sumsq) + 1e-6));
 *                          then torch.optim.Adam's update (no amsgrad) with bias corrections of `step` (>= 1).
 *      grad_scale folds the 1/world of the data-parallel mean; max_norm <= 0 disables clipping (sumsq may be NULL);
 *      write_grad != 0 stores the scaled/clipped gradient back (what clip_grad_norm_ leaves in .grad);
 *      norm_out (device float, may be NULL) receives the total gradient norm clip_grad_norm_ returns.
 * ----------------------------------------------------------------------------------------- */
int matgcn_grad_sumsq(const float* grad, long long n, double* sumsq, void* stream);
int matgcn_adam_clip_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, const double* sumsq,
                          float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                          long long step, int write_grad, float* norm_out, void* stream);
/* The same step with the two values that change from step to step resident in device memory, so that ONE captured CUDA graph of
 * the whole loop body (executor:413-422) can be replayed for every step (train.GraphedTrainStep):
 *   matgcn_adam_clip_step_dev: lr_dev (device float) is read by the kernel (an lr scheduler, executor:155-197, writes it between
 *                          replays), step_dev (device int64, >= 1 when the kernel runs) yields the bias corrections;
 *   matgcn_step_tick:      first node of the captured step: *step_dev += 1 (torch.optim.Adam increments before the update) and
 *                          *seed_dev moves to the next dropout key (either pointer may be NULL). */
int matgcn_adam_clip_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, const double* sumsq,
                              float max_norm, float grad_scale, const float* lr_dev, float beta1, float beta2, float eps,
                              float weight_decay, const long long* step_dev, int write_grad, float* norm_out, void* stream);
int matgcn_step_tick(unsigned long long* seed_dev, long long* step_dev, void* stream);

/* f2 - batch assembly: replaces MTHDataset._get_sample_indices / _generate_input_data
 *      (libcity/data/dataset/dataset_subclass/mth_dataset.py:31-60, 62-158) plus the per-batch collate and upload
 *      (libcity/data/utils.py:68-72, libcity/data/batch.py:43-57) with a gather from a series resident in HBM.
 *   series       [T_total, N, F]   the (already scaled) time series
 *   seg_offsets  [n_seg] (device int)  time-slice distance of each input segment's start before the label start, in
 *                the order the reference concatenates them: closeness oldest..newest, period oldest..newest, trend
 *   label_starts [B] (device int64)  first predicted time slice of each sample
 *   X            [B, n_seg*in_window, N, F] out;  y [B, out_window, N, F] out
 *   bad_flag     device int, set to 1 if any sample is not a valid one under the reference's rules (its chunks are
 *                then left unwritten); the caller zeroes it.
 * ----------------------------------------------------------------------------------------- */
int matgcn_assemble_windows(const float* series, long long T_total, int N, int F, const int* seg_offsets, int n_seg,
                            int in_window, int out_window, const long long* label_starts, int B, float* X, float* y,
                            int* bad_flag, void* stream);

/* f3 - dropout + output head: replaces MA.py:416-417 (F.dropout(p=0.1, training) on the encoder output, then
 *      end_conv = Conv2d(T -> T_out*C, kernel (1, H)), MA.py:340-344: the time steps are the channels), on the
 *      node-major encoder output as it sits in the layer workspace:
 *        out[r, o] = bias[o] + sum_t sum_h drop(y[t, r, h]) * w[o, t, h]        r = (node, batch) row, rows = N*B
 *   y     [Tc, rows, H] with time stride y_tstride (floats); H must be 64
 *   w     [O, Tc, H] (= end_conv.weight[:, :, 0, :]), bias [O], out [rows, O]
 *   p_drop in [0,1): 0 = eval mode.  The mask is counter-based (Philox4x32-10 keyed by `seed`, 16 bits per element)
 *   and never stored: the backward regenerates it from the same seed.  The realised drop probability is
 *   round(p*65536)/65536 and kept values are scaled by matgcn_head_dropout_scale(p) = 1/(1 - that), so E[drop(x)] = x.
 *   Backward overwrites dy [Tc, rows, H] (contiguous), dw [O, Tc, H], dbias [O].
 *   matgcn_head_dropout_mask writes the multipliers (0 or scale) of elements 0..n-1 (tests / diagnostics).
 * All arithmetic is fp32 FFMA (no TF32), so the 1e-4 parity bound of the head holds in every mode. */
float matgcn_head_dropout_scale(float p_drop);
int matgcn_head_fwd(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, const float* bias, int O,
                    float p_drop, unsigned long long seed, float* out, void* stream);
int matgcn_head_bwd(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, int O, float p_drop,
                    unsigned long long seed, const float* dout, float* dy, float* dw, float* dbias, void* stream);
int matgcn_head_dropout_mask(long long n, float p_drop, unsigned long long seed, float* mult, void* stream);
/* Same two operators with the mask keyed by `seed ^ *seed_dev` (device uint64 advanced by matgcn_step_tick): the form a captured
 * train step uses, so that a replay draws a new mask (MA.py:416: F.dropout draws a new mask every call). */
int matgcn_head_fwd_dev(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, const float* bias, int O,
                        float p_drop, unsigned long long seed, const unsigned long long* seed_dev, float* out, void* stream);
int matgcn_head_bwd_dev(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, int O, float p_drop,
                        unsigned long long seed, const unsigned long long* seed_dev, const float* dout, float* dy, float* dw,
                        float* dbias, void* stream);

/* f3, second half - calculate_loss (MultiATGCN.py:422-427): StandardScaler.inverse_transform (libcity/utils/normalization.py:62-76)
 *      of forecast and target, then masked_mae_torch(pred, true, 0) (libcity/model/loss.py:17-29), as one streaming pass.
 *      pred and y are 4-d float32 device tensors [B, T_out, N, C] described by `sizes` and their element strides (the forecast
 *      is a permuted view of the head's output, the target a channel slice of batch['y']); acc is a 3-double device scratch
 *      kept for the backward; loss receives the scalar.  bwd: dpred (contiguous [B, T_out, N, C]) = grad_loss * dloss/dpred. */
int matgcn_masked_mae_fwd(const float* pred, const float* y, const long long* sizes, const long long* pred_strides,
                          const long long* y_strides, float mean, float std, float min_s, double* acc, float* loss, void* stream);
int matgcn_masked_mae_bwd(const float* pred, const float* y, const long long* sizes, const long long* pred_strides,
                          const long long* y_strides, float mean, float std, float min_s, const double* acc, const float* grad_loss,
                          float* dpred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MATGCN_H_ */
