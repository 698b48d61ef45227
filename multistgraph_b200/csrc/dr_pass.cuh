// One pass over the residual-cell pre-activation gradients DR of a layer for everything that reduces them over (t, n, b):
//   dRgw[:, Cin:] += DR[:, 0:2H]^T H1      dRuw[:, Cin:] += DR[:, 2H:]^T ZH2        (MA.py:142-150, hidden columns)
//   dRgw[:, 0:Cin] += DR[:, 0:2H]^T x      dRuw[:, 0:Cin] += DR[:, 2H:]^T x         (input columns, Cin = 64 only)
//   dRgb += colsum DR[:, 0:2H]             dRub += colsum DR[:, 2H:]
// The per-phase path ran these as four split-K launches plus a column-sum kernel, i.e. DR (475 MB per layer at the Baltimore
// shape) was streamed five times; here every byte is read once.
//
// The reduction index (rows of DR) is the K dimension of the MMAs, so DR, H1, ZH2 and x are all MN-major TF32 operands and
// arrive by TMA exactly as they sit in memory ({32 columns, 32 rows} boxes, 32-byte-atom 128B swizzle).  Per 32-row block:
//   D_g [128 x 144] += DR[:, 0:128]^T  [H1 | x | 1]          D_u [64 x 160] += DR[:, 128:192]^T  [x | 1 | ZH2]
// - the B operands are two overlapping windows of one row of 32-column atoms  H1 H1 x x 1 ZH2 ZH2  in the stage, and the
// "1" atom (written once, never overwritten by TMA) turns the bias column sums into one more accumulator column.
// Every CTA reduces a contiguous range of row blocks into TMEM and adds its partial sums to the outputs with fp32 atomics.
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (warp 2 owns the TMEM allocation).
#pragma once
#include "gemm_tc.cuh"
#include "rec_api.h"
#include "rec_fwd.cuh"

namespace matgcn {

constexpr int DRP_STAGES = 4;
constexpr int DRP_ATOM = 4096;                       // one {32 cols, 32 rows} fp32 box
constexpr int DRP_OFF_B = 6 * DRP_ATOM;              // DR atoms 0..5, then  H1 H1 x x 1 ZH2 ZH2
constexpr int DRP_STAGE_BYTES = 13 * DRP_ATOM;       // 52 KB
constexpr int DRP_OFF_BAR = DRP_STAGES * DRP_STAGE_BYTES;
constexpr int DRP_SMEM_TOTAL = DRP_OFF_BAR + 128 + 1024;
constexpr int DRP_THREADS = 192;
constexpr int DRP_NG = 144;                          // D_g columns: H1 64 | x 64 | ones 16 (of the atom's 32)
constexpr int DRP_NU = 160;                          // D_u columns: x 64 | ones 32 | ZH2 64
constexpr int DRP_TMEM_DU = 256;

struct DrpMaps { CUtensorMap DR, H1, ZH2, X; };
struct DrpP {
    int T, NB, Cin, I;
    int nkb_t;            // 32-row blocks per time step
    int total_kb;
    int has_x, bias;
    float* dRgw; float* dRuw; float* dRgb; float* dRub;
};

__device__ __forceinline__ void drp_tmem_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

__global__ void __launch_bounds__(DRP_THREADS, 1) dr_pass_kernel(const __grid_constant__ DrpMaps maps, const DrpP p) {
    extern __shared__ uint8_t drp_smem_raw[];
    const uint32_t base = (smem_u32(drp_smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + DRP_OFF_BAR;
    const uint32_t full0 = bars, empty0 = bars + 8 * DRP_STAGES, acc_full = bars + 16 * DRP_STAGES, tmem_slot = acc_full + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb0 = (int)((long long)p.total_kb * blockIdx.x / gridDim.x);
    const int kb1 = (int)((long long)p.total_kb * (blockIdx.x + 1) / gridDim.x);

    // constant atoms: ones (bias sums); zeros in place of x when the layer has no 64-wide input
    for (int i = threadIdx.x; i < DRP_STAGES * (DRP_ATOM / 16); i += DRP_THREADS) {
        const int s = i / (DRP_ATOM / 16), o = (i % (DRP_ATOM / 16)) * 16;
        const uint32_t st = base + s * DRP_STAGE_BYTES + DRP_OFF_B;
        rf_sts4(st + 4 * DRP_ATOM + o, make_float4(1.f, 1.f, 1.f, 1.f));
        if (!p.has_x) {
            rf_sts4(st + 2 * DRP_ATOM + o, make_float4(0.f, 0.f, 0.f, 0.f));
            rf_sts4(st + 3 * DRP_ATOM + o, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < DRP_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0 && kb1 > kb0) {
            const uint32_t bytes = (p.has_x ? 12 : 10) * DRP_ATOM;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const int t = kb / p.nkb_t, row = (kb - t * p.nkb_t) * 32;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t st = base + stage * DRP_STAGE_BYTES, fb = full0 + 8 * stage;
                mbar_expect_tx(fb, bytes);
#pragma unroll
                for (int j = 0; j < 6; ++j) tma_load_5d(st + j * DRP_ATOM, &maps.DR, fb, 32 * j, row, t, 0, 0);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    tma_load_5d(st + DRP_OFF_B + j * DRP_ATOM, &maps.H1, fb, 32 * j, row, t, 0, 0);
                    tma_load_5d(st + DRP_OFF_B + (5 + j) * DRP_ATOM, &maps.ZH2, fb, 32 * j, row, t, 0, 0);
                    if (p.has_x) tma_load_5d(st + DRP_OFF_B + (2 + j) * DRP_ATOM, &maps.X, fb, 32 * j, row, t, 0, 0);
                }
                if (++stage == DRP_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && kb1 > kb0) {
            // D = f32, A = B = tf32, both MN-major, N >> 3, M >> 4
            const uint32_t id_common = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16);
            const uint32_t idg = id_common | ((uint32_t)(DRP_NG >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idu = id_common | ((uint32_t)(DRP_NU >> 3) << 17) | ((64u >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t st = base + stage * DRP_STAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t acc = (kb > kb0 || kk > 0) ? 1u : 0u;
                    const uint64_t ag = umma_desc(st + kk * 1024, DRP_ATOM, 512, 1);
                    const uint64_t au = umma_desc(st + 4 * DRP_ATOM + kk * 1024, DRP_ATOM, 512, 1);
                    const uint64_t bg = umma_desc(st + DRP_OFF_B + kk * 1024, DRP_ATOM, 512, 1);
                    const uint64_t bu = umma_desc(st + DRP_OFF_B + 2 * DRP_ATOM + kk * 1024, DRP_ATOM, 512, 1);
                    umma_tf32(tmem, ag, bg, idg, acc);
                    umma_tf32(tmem + DRP_TMEM_DU, au, bu, idu, acc);
                }
                umma_commit(empty0 + 8 * stage);
                if (++stage == DRP_STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else if (kb1 > kb0) {
        const int q = warp & 3;      // TMEM lane quadrant this warp may read
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
        float v[16];
        // D_g: row m = 32q + lane; columns  H1 (0..63) | x (64..127) | 1 (128)
        {
            const int m = 32 * q + lane;
            float* g = p.dRgw + (long long)m * p.I;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                drp_tmem_ld16(lane_addr + 16 * c, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) atomicAdd(g + p.Cin + 16 * c + i, v[i]);
            }
            if (p.has_x) {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    drp_tmem_ld16(lane_addr + 64 + 16 * c, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(g + 16 * c + i, v[i]);
                }
            }
            if (p.bias) {
                drp_tmem_ld16(lane_addr + 128, v);
                atomicAdd(p.dRgb + m, v[0]);
            }
        }
        // D_u (M = 64: 16 rows per quadrant on lanes 0..15): columns  x (0..63) | 1 (64..95) | ZH2 (96..159)
        {
            const int r = 16 * q + (lane & 15);
            const bool ok = lane < 16;
            float* u = p.dRuw + (long long)r * p.I;
            if (p.has_x) {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    drp_tmem_ld16(lane_addr + DRP_TMEM_DU + 16 * c, v);
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) atomicAdd(u + 16 * c + i, v[i]);
                    }
                }
            }
            if (p.bias) {
                drp_tmem_ld16(lane_addr + DRP_TMEM_DU + 64, v);
                if (ok) atomicAdd(p.dRub + r, v[0]);
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                drp_tmem_ld16(lane_addr + DRP_TMEM_DU + 96 + 16 * c, v);
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(u + p.Cin + 16 * c + i, v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// fp32 tensor map {cols, rows, T} with {32, 32, 1} boxes in the MN-major TF32 operand layout
inline bool drp_make_map(CUtensorMap* map, const float* base, int cols, int ld, int rows, long long tstride, int T) {
    TmapEncodeFn enc = tmap_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3) || (tstride & 3) || (cols & 31)) return false;
    cuuint64_t d[5] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)T, 1, 1};
    cuuint64_t s[4] = {(cuuint64_t)ld * 4, (cuuint64_t)tstride * 4, (cuuint64_t)tstride * 4 * T, (cuuint64_t)tstride * 4 * T};
    cuuint32_t b[5] = {32, 32, 1, 1, 1};
    cuuint32_t e[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 4; ++i)
        if (s[i] == 0 || s[i] >= (1ULL << 40)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_dr_pass(const DrPassArgs& a, cudaStream_t st) {
    if (a.H != 64 || a.T <= 0 || a.NB <= 0) return cudaErrorNotSupported;
    const bool has_x = a.X != nullptr;
    if (has_x && a.Cin != 64) return cudaErrorNotSupported;
    DrpMaps maps;
    const long long U = (long long)a.NB * 64;
    if (!drp_make_map(&maps.DR, a.DR, 192, 192, a.NB, 3 * U, a.T) || !drp_make_map(&maps.H1, a.H1, 64, 64, a.NB, U, a.T) ||
        !drp_make_map(&maps.ZH2, a.ZH2, 64, 64, a.NB, U, a.T))
        return cudaErrorNotSupported;
    if (has_x) {
        if (!drp_make_map(&maps.X, a.X, 64, 64, a.NB, a.x_tstride, a.T)) return cudaErrorNotSupported;
    } else {
        maps.X = maps.H1;
    }
    DrpP p;
    p.T = a.T; p.NB = a.NB; p.Cin = a.Cin; p.I = a.Cin + 64;
    p.nkb_t = (a.NB + 31) / 32;
    const long long total = (long long)p.nkb_t * a.T;
    if (total > 2147483647LL) return cudaErrorNotSupported;
    p.total_kb = (int)total;
    p.has_x = has_x ? 1 : 0;
    p.bias = a.dRgb ? 1 : 0;
    p.dRgw = a.dRgw; p.dRuw = a.dRuw; p.dRgb = a.dRgb; p.dRub = a.dRub;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64) return cudaErrorNotSupported;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dr_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DRP_SMEM_TOTAL);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int sms = sm_count();
    const int grid = p.total_kb < sms ? p.total_kb : sms;
    dr_pass_kernel<<<grid, DRP_THREADS, DRP_SMEM_TOTAL, st>>>(maps, p);
    return cudaGetLastError();
}

}  // namespace matgcn
