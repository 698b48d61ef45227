// Throughput of the legacy warp-level MMA paths on sm_100a (per SM), to decide what the small in-epilogue products
// should use: mma.sync TF32 m16n8k8, BF16 m16n8k16, and plain FFMA for scale.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, long long* cycles) {
    float c[8][4];
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9f + i + j;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 11, b1 = 13;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = fmaf(c[i][j], 1.0001f, 0.5f);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double macs_per_instr) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    rate_kernel<MODE><<<148, 256>>>(out, iters, cyc);
    rate_kernel<MODE><<<148, 256>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double instr_per_sm = 8.0 * iters * 8;  // 8 warps x iters x 8 independent instructions (warp-level)
    printf("%-28s %8lld cycles  -> %.2f cycles per warp-instruction per SM, %.1f MAC/cycle/SM\n", name, h[0],
           h[0] / instr_per_sm, instr_per_sm * macs_per_instr / h[0]);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("mma.sync m16n8k8 tf32", 16 * 8 * 8);
    run<1>("mma.sync m16n8k16 bf16", 16 * 8 * 16);
    run<2>("mma.sync m16n8k16 f16", 16 * 8 * 16);
    run<3>("FFMA x4 (per 'instruction')", 4 * 32);
    return 0;
}
