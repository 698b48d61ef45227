// Probe of the tcgen05.ld.16x256b register layout: fill TMEM lanes 0..31 x 32 columns with lane*100 + col through
// tcgen05.st.32x32b, read back with 16x256b.x4, print what every thread got.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o ldtm_probe ldtm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(float* out) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    __syncthreads();
    const uint32_t tm = slot;
    const int lane = threadIdx.x;
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint((float)(lane * 100 + c));
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tm),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31]));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tm));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[lane * 16 + i] = __uint_as_float(r[i]);
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}
int main() {
    float* d;
    cudaMalloc(&d, 32 * 16 * 4);
    k<<<1, 32>>>(d);
    float h[32 * 16];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    int bad = 0;
    for (int lane = 0; lane < 32; ++lane) {
        for (int i = 0; i < 16; ++i) {
            const int j = i / 4, w = i % 4;
            const int row = lane / 4 + (w >= 2 ? 8 : 0), col = 8 * j + 2 * (lane % 4) + (w & 1);
            if (h[lane * 16 + i] != (float)(row * 100 + col)) ++bad;
        }
    }
    printf("mismatches against the assumed layout (reg 4j+w: row lane/4 + 8*(w/2), col 8j + 2*(lane%%4) + w%%2): %d\n", bad);
    for (int lane = 0; lane < 6; ++lane) {
        printf("lane %d:", lane);
        for (int i = 0; i < 16; ++i) printf(" %g", h[lane * 16 + i]);
        printf("\n");
    }
    return 0;
}
