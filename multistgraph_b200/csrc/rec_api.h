// Host interface of the persistent recurrence kernels (their own translation unit, rec.cu: the tensor-core unit matgcn.cu
// takes minutes to build).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace matgcn {

struct RecFwdArgs {
    int T, N, B, Cin, K, ldm;
    const __nv_bfloat16* M16; __nv_bfloat16* PH16; __nv_bfloat16* PZ16; const __nv_bfloat16* WG16; const __nv_bfloat16* WU16;
    const float* GX; const float* RX; float* PH; float* PZ;
    float* Z; float* R; float* HC; float* H1; float* Z2; float* R2; float* HC2; float* ZH2;
    const float* RgH; const float* RuH; const float* mix;
    unsigned int* gbar;
};

// cudaErrorNotSupported: the shape does not meet the kernel's requirements (the caller runs one launch per phase instead)
cudaError_t launch_rec_fwd(const RecFwdArgs& a, cudaStream_t st);

// Reverse-time recurrence of one layer (rec_bwd.cuh): data gradients and pre-activation gradients of all T steps.
struct RecBwdArgs {
    int T, N, B, Cin, K, ldm, n_adp;
    const float* dy; long long dy_tstride;
    const __nv_bfloat16* M16; const __nv_bfloat16* WG16; const __nv_bfloat16* WU16;
    // saved by the forward pass: PH (slot 0 of step t = h_{t-1}), Z, R, HC, H1, Z2, R2, HC2  [T, N*B, H]
    const float* PH; const float* Z; const float* R; const float* HC; const float* H1; const float* Z2; const float* R2; const float* HC2;
    const float* RgH; const float* RuH; const float* mix;
    float* DG; float* DR;            // [T, N*B, 3H] out (in place of GX / RX); DG may be null: only the bf16 twin is written
    __nv_bfloat16* DG16;             // [T, N*B, 3H] out: bf16 twin of DG
    float* DPT0;                     // [N*B*H] scratch: slot 0 of the per-node products
    __nv_bfloat16* DPT16;            // [K, N*B*H] scratch: slots 1.. of the per-node products (bf16)
    float* DHD; float* DHD2; float* DZC;   // [N*B*H] scratch (DHD2 and DZC zero-filled by the caller)
    __nv_bfloat16* DPZA; __nv_bfloat16* DPHA;   // [T, n_adp, N*B*H] out: per-step copies of the adaptive slices (bf16)
    float* DHC;                      // [N*B*H] out: gradient w.r.t. the initial state
    float* dmix;                     // [T] out (zero-filled by the caller)
    unsigned int* gbar;
};
cudaError_t launch_rec_bwd(const RecBwdArgs& a, cudaStream_t st);

// One pass over DR for all its reductions over (t, n, b): residual-cell weight and bias gradients (dr_pass.cuh).  X (the layer
// input, [T][NB, 64] with x_tstride floats between steps) may be null: the input columns are then left to the caller; dRgb /
// dRub may be null (no bias sums).  All outputs are ACCUMULATED into (atomics): the caller zero-fills them.
struct DrPassArgs {
    int T, NB, Cin, H;
    const float* DR; const float* H1; const float* ZH2; const float* X; long long x_tstride;
    float* dRgw; float* dRuw; float* dRgb; float* dRub;
};
cudaError_t launch_dr_pass(const DrPassArgs& a, cudaStream_t st);

// Optional device timing of the two persistent kernels (bench.py's roofline): while enabled, every launch is bracketed by CUDA
// events on its stream; rec_timing_read() waits for them and returns the summed durations and launch counts since enabling.
void rec_timing_enable(bool on);
void rec_timing_read(double* fwd_ms, int* fwd_n, double* bwd_ms, int* bwd_n);

}  // namespace matgcn
