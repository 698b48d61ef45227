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

}  // namespace matgcn
