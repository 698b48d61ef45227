// Batched fp32 SIMT GEMM with strided k-batches, split-K and a fused epilogue functor.
//
// This is the "exact mode" engine (fp32 FFMA, fp32 accumulate) behind every contraction of
// the Multi-ATGCN path; the bf16 tcgen05 engine (gemm_tc.cuh) replaces the propagation
// GEMMs in fast mode.  One template covers
//   * support propagation         C[(k,n), col]  = sum_m  M[(k,n), m] * X[m, col]
//   * its transpose (backward)    dX[m, col]     = sum_(k,n) M[(k,n), m] * dP[(k,n), col]
//   * node-batched contractions   out_n[b, o]    = sum_k sum_i P_k[n, b, i] * W[n, k, i, o]
//   * time-batched weight grads   dW[n,k][i, o]  = sum_t sum_b P_k,t[n, b, i] * D_t[n, b, o]
// through three knobs: operand layouts, a two-level batch index z = (z1, z2) and "k-batches"
// (the reduction runs over KB strided slabs of inner length K).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace matgcn {

// kernels launched by this library in this process (reported by bench.py as gpu_launches)
inline std::atomic<unsigned long long> g_launches{0};
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// values an epilogue reads from global memory for one output element (see the epilogue section of matgcn.cu)
struct EpiIn { float a, b, c, d, e, f; };
// the same for four consecutive output columns (vectorised tensor-core epilogue)
struct EpiIn4 { float4 a, b, c, d, e, f; };
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
// four floats -> four bf16 (round to nearest even), one 8-byte store; p must be 8-byte aligned
__device__ __forceinline__ void st4_bf16(__nv_bfloat16* p, const float4& v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// a range to pull into L2: n chunks of `chunk` bytes (a multiple of 128), `stride` bytes apart
struct PfRange { const char* base; long long stride; int chunk, n; };

struct GemmP {
    const float* A;
    const float* B;
    // optional bf16 twins of the operands (same logical layout and strides): used by the bf16 tensor-core engine
    const __nv_bfloat16* A16;
    const __nv_bfloat16* B16;
    int keepB;    // tensor-core engine: load B with the L2 evict-last policy (an operand re-read by every step of a recurrence)
    PfRange pf[8];        // tensor-core engine: L2 warm-up for the next launches (see TcP::pf); npf ranges in use
    int npf;
    int need16;   // the fp32 operands are NOT valid (their producer skipped the fp32 store): only the bf16 engine may run this
    int M, N, K;  // K: inner reduction length of one k-batch
    int KB;       // number of k-batches
    int lda, ldb;
    long long sA1, sA2, sAk;  // A offsets: z1, z2, k-batch
    long long sB1, sB2, sBk;
    int Z2;      // z = z1 * Z2 + z2
    int splits;  // split-K factor (epilogue must be atomic when > 1)
};

template <int BM_, int BN_, int BK_, int TM_, int TN_>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
    static constexpr int TX = BN / TN, TY = BM / TM;
    static constexpr int NT = TX * TY;
    static constexpr int RM = TM / 4, RN = TN / 4;  // float4 groups per thread
    static_assert(TM % 4 == 0 && TN % 4 == 0, "microtile is built from float4 groups");
    static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile loads must divide evenly");
};
using CfgBig = TileCfg<128, 128, 16, 8, 8>;    // 256 threads
using CfgMid = TileCfg<64, 64, 16, 4, 4>;      // 256 threads
using CfgSkinnyM = TileCfg<32, 128, 16, 4, 4>; // 256 threads, few output rows
using CfgSkinnyN = TileCfg<128, 32, 16, 4, 4>; // 256 threads, few output columns

// A_KC: element (m,k) at m*lda + k (K contiguous); otherwise at k*lda + m (M contiguous).
// B_KC: element (k,n) at n*ldb + k (K contiguous); otherwise at k*ldb + n (N contiguous).
template <class Cfg, bool A_KC, bool B_KC, class Epi>
__global__ void __launch_bounds__(Cfg::NT) gemm_kernel(const GemmP p, const Epi epi, const int tiles_n, const int tiles) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, TM = Cfg::TM, TN = Cfg::TN, NT = Cfg::NT;
    constexpr int LA = BM * BK / NT, LB = BN * BK / NT;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int z = blockIdx.x / tiles;
    const int tile = blockIdx.x - z * tiles;
    const int z1 = z / p.Z2, z2 = z - z1 * p.Z2;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const float* __restrict__ Az = p.A + z1 * p.sA1 + z2 * p.sA2;
    const float* __restrict__ Bz = p.B + z1 * p.sB1 + z2 * p.sB2;

    const int kt_per_kb = (p.K + BK - 1) / BK;
    const int kt_total = p.KB * kt_per_kb;
    const int kt_per_split = (kt_total + p.splits - 1) / p.splits;
    const int kt0 = blockIdx.y * kt_per_split;
    const int kt1 = min(kt_total, kt0 + kt_per_split);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int tx = tid % Cfg::TX, ty = tid / Cfg::TX;
    float ra[LA], rb[LB];

    auto fetch = [&](int kt) {
        const int kb = kt / kt_per_kb;
        const int k0 = (kt - kb * kt_per_kb) * BK;
        const float* __restrict__ Ab = Az + kb * p.sAk;
        const float* __restrict__ Bb = Bz + kb * p.sBk;
#pragma unroll
        for (int l = 0; l < LA; ++l) {
            const int i = tid + l * NT;
            int m, k;
            if (A_KC) { k = i % BK; m = i / BK; } else { m = i % BM; k = i / BM; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < p.M && gk < p.K) v = A_KC ? __ldg(Ab + (long long)gm * p.lda + gk) : __ldg(Ab + (long long)gk * p.lda + gm);
            ra[l] = v;
        }
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const int i = tid + l * NT;
            int n, k;
            if (B_KC) { k = i % BK; n = i / BK; } else { n = i % BN; k = i / BN; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < p.N && gk < p.K) v = B_KC ? __ldg(Bb + (long long)gn * p.ldb + gk) : __ldg(Bb + (long long)gk * p.ldb + gn);
            rb[l] = v;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int l = 0; l < LA; ++l) {
            const int i = tid + l * NT;
            int m, k;
            if (A_KC) { k = i % BK; m = i / BK; } else { m = i % BM; k = i / BM; }
            As[buf][k][m] = ra[l];
        }
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const int i = tid + l * NT;
            int n, k;
            if (B_KC) { k = i % BK; n = i / BK; } else { n = i % BN; k = i / BN; }
            Bs[buf][k][n] = rb[l];
        }
    };

    if (kt0 < kt1) {
        fetch(kt0);
        stash(0);
    }
    __syncthreads();
    int buf = 0;
    for (int kt = kt0; kt < kt1; ++kt) {
        const bool more = (kt + 1 < kt1);
        if (more) fetch(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int r = 0; r < Cfg::RM; ++r) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][r * (BM / Cfg::RM) + ty * 4]);
                a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < Cfg::RN; ++r) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][r * (BN / Cfg::RN) + tx * 4]);
                b[r * 4 + 0] = v.x; b[r * 4 + 1] = v.y; b[r * 4 + 2] = v.z; b[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    if (kt0 >= kt1 && p.splits > 1) return;  // an empty split contributes nothing
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + (i / 4) * (BM / Cfg::RM) + ty * 4 + (i % 4);
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + (j / 4) * (BN / Cfg::RN) + tx * 4 + (j % 4);
            if (col < p.N) epi(z1, z2, row, col, acc[i][j]);
        }
    }
}

template <class Cfg, bool A_KC, bool B_KC, class Epi>
inline cudaError_t launch_gemm(const GemmP& p, const Epi& epi, int Z, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0 || Z <= 0) return cudaSuccess;
    const int tiles_m = (p.M + Cfg::BM - 1) / Cfg::BM;
    const int tiles_n = (p.N + Cfg::BN - 1) / Cfg::BN;
    const long long gx = (long long)tiles_m * tiles_n * Z;
    if (gx > 2147483647LL) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)gx, (unsigned)(p.splits > 0 ? p.splits : 1), 1);
    GemmP q = p;
    if (q.splits < 1) q.splits = 1;
    if (q.Z2 < 1) q.Z2 = 1;
    gemm_kernel<Cfg, A_KC, B_KC, Epi><<<grid, Cfg::NT, 0, st>>>(q, epi, tiles_n, tiles_m * tiles_n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace matgcn
