// Input-side gradients of a layer with a tiny channel count (layer 0: Cin = 2) in the fast modes, as warp-level TF32
// mma.sync products instead of the shuffle-reduction FFMA kernels (xside_bwd_dg_small / xside_bwd_dr_small, which stay the
// exact-mode path): one pass over G = DG (per node) or G = DR (flat rows), 16 rows x 192 columns per warp tile,
//
//   (1)  DX[16 rows, kc]   = G_tile [16, 192] x W3^T [192, kc]        input-side data gradient (stored, or added for DR)
//   (2)  dW3[kc, 192]     += [X | 1]^T [kc+1, 16 rows] x G_tile        weight gradient, and the bias gradient through the ones column
//
// kc = (support k, channel i) for DG (K*Cin <= 14), kc = channel i for DR.  The tile goes global -> shared memory with
// cp.async (double-buffered per warp, no registers), both products read their fragments from it conflict-free (pitch 196,
// and rows paired (2 tig, 2 tig + 1) along the k index of product (2)).  The per-warp dW3 accumulators (24 n-tiles) live in
// registers over all tiles of the warp and are summed once per block through the (then idle) tile buffers.
#pragma once
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace matgcn {

constexpr int XM_H3 = 192;                 // 3 * rnn_units (H = 64)
constexpr int XM_LD = 196;                 // shared-memory pitch of the G tile and of W3 (floats)
constexpr int XM_WARPS = 8;
constexpr int XM_TILE = 16 * XM_LD;        // floats per tile buffer
constexpr int XM_SMEM_FLOATS = 16 * XM_LD + XM_WARPS * 2 * XM_TILE;

struct XsMmaArgs {
    const float* G;      // NODE: DG [T, N, B, 192]   else DR [T*N*B, 192]
    const float* PX;     // [T, K, N, B, Cin]
    float* DPX;          // [T, K, N, B, Cin]
    const float* Wa;     // NODE: Wg [N, K, I, 2H]    else Rgw [2H, I]
    const float* Wb;     // NODE: Wu [N, K, I, H]     else Ruw [H, I]
    float* dWa;          // same shapes as Wa / Wb (only the input rows / columns are written)
    float* dWb;
    float* dba;          // NODE: dbg [N, 2H], dbu [N, H]   else dRgb [2H], dRub [H]
    float* dbb;
    int T, N, B, Cin, H, K;
};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool NODE>
__global__ void __launch_bounds__(XM_WARPS * 32, 1) xside_bwd_mma_kernel(const XsMmaArgs a) {
    extern __shared__ __align__(16) float xm_smem[];
    float* w3 = xm_smem;                                            // [16 kc][XM_LD]: W3[kc][o], rows >= KC zero
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    float* tilebuf = xm_smem + 16 * XM_LD + warp * 2 * XM_TILE;    // this warp's two tile buffers
    const int Cin = a.Cin, H = a.H, K = a.K, I = Cin + H;
    const int KC = NODE ? K * Cin : Cin;                            // real rows of W3; row KC of dW3 is the bias gradient
    const long long NB = (long long)a.N * a.B, UX = NB * Cin;
    const int n = blockIdx.x;                                       // NODE: the node of this block
    for (int j = threadIdx.x; j < 16 * XM_H3; j += blockDim.x) {
        const int kc = j / XM_H3, o = j - kc * XM_H3;
        float v = 0.f;
        if (kc < KC) {
            if (NODE) {
                const int k = kc / Cin, i = kc - k * Cin;
                v = o < 2 * H ? a.Wa[(((long long)n * K + k) * I + i) * 2 * H + o] : a.Wb[(((long long)n * K + k) * I + i) * H + o - 2 * H];
            } else {
                v = o < 2 * H ? a.Wa[(long long)o * I + kc] : a.Wb[(long long)(o - 2 * H) * I + kc];
            }
        }
        w3[kc * XM_LD + o] = v;
    }
    __syncthreads();
    // rows of this block / warp: NODE -> r = (t, b) of node n, T*B rows; else flat rows (t, n, b)
    const long long rows_total = NODE ? (long long)a.T * a.B : (long long)a.T * NB;
    const long long ntiles = (rows_total + 15) / 16;
    const long long tile0 = NODE ? warp : (long long)blockIdx.x * XM_WARPS + warp;
    const long long tstep = NODE ? XM_WARPS : (long long)gridDim.x * XM_WARPS;
    // Row r of this block = (time step t, position q inside the step): q = batch index b (NODE, period B) or flat (n, b)
    // (period N*B).  One 64-bit division per TILE (t0, q0 of its first row); the rows of the tile step from there.
    const long long period = NODE ? (long long)a.B : NB;
    long long cur_t0 = 0, cur_q0 = 0, nxt_t0 = 0, nxt_q0 = 0;
    auto split = [&](long long t0, long long q0, int row, long long& t, long long& q) {   // (t, q) of row r0 + row
        t = t0; q = q0 + row;
        while (q >= period) { q -= period; ++t; }
    };
    auto grow = [&](long long t, long long q) -> const float* {   // global address of that row of G
        return NODE ? a.G + ((t * a.N + n) * a.B + q) * XM_H3 : a.G + (t * NB + q) * XM_H3;
    };
    auto xoff = [&](long long t, long long q, int kc) -> long long {   // offset of X[row][kc] inside PX / DPX
        if (NODE) { const int k = kc / Cin, i = kc - k * Cin; return (t * K + k) * UX + ((long long)n * a.B + q) * Cin + i; }
        return t * K * UX + q * Cin + kc;
    };
    auto issue_tile = [&](long long tile, float* buf, long long& t0, long long& q0) {   // 16 rows x 48 float4, coalesced; rows past the end zeroed
        const long long r0 = tile * 16;
        t0 = r0 / period; q0 = r0 - t0 * period;
        int prow = -1;
        const float* src = nullptr;
#pragma unroll
        for (int it = 0; it < 24; ++it) {
            const int idx = lane + 32 * it, row = idx / 48, c4 = idx - row * 48;
            if (row != prow) { long long t, q; split(t0, q0, row, t, q); src = grow(t, q); prow = row; }
            float* dst = buf + row * XM_LD + 4 * c4;
            if (r0 + row < rows_total) cp_async16(dst, src + 4 * c4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };
    float acc[24][4];
#pragma unroll
    for (int nt = 0; nt < 24; ++nt)
#pragma unroll
        for (int x = 0; x < 4; ++x) acc[nt][x] = 0.f;
    int cur = 0;
    if (tile0 < ntiles) issue_tile(tile0, tilebuf, nxt_t0, nxt_q0);
    for (long long tile = tile0; tile < ntiles; tile += tstep) {
        const long long r0 = tile * 16;
        float* buf = tilebuf + cur * XM_TILE;
        cur_t0 = nxt_t0; cur_q0 = nxt_q0;
        if (tile + tstep < ntiles) { issue_tile(tile + tstep, tilebuf + (cur ^ 1) * XM_TILE, nxt_t0, nxt_q0); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        // ---- (1) DX = G_tile x W3^T : M = 16 rows, N = kc (n-tile 0: kc 0..7, n-tile 1: kc 8..15), K = 192 ----
        float dx0[4] = {0.f, 0.f, 0.f, 0.f}, dx1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int ks = 0; ks < 24; ++ks) {
            uint32_t af[4];
            af[0] = __float_as_uint(buf[g * XM_LD + 8 * ks + tig]); af[1] = __float_as_uint(buf[(g + 8) * XM_LD + 8 * ks + tig]);
            af[2] = __float_as_uint(buf[g * XM_LD + 8 * ks + tig + 4]); af[3] = __float_as_uint(buf[(g + 8) * XM_LD + 8 * ks + tig + 4]);
            mma_tf32_16x8x8(dx0, af, __float_as_uint(w3[g * XM_LD + 8 * ks + tig]), __float_as_uint(w3[g * XM_LD + 8 * ks + tig + 4]));
            if (NODE) mma_tf32_16x8x8(dx1, af, __float_as_uint(w3[(g + 8) * XM_LD + 8 * ks + tig]), __float_as_uint(w3[(g + 8) * XM_LD + 8 * ks + tig + 4]));
        }
        // DX fragment: rows g / g+8, kc = 8*nt + 2*tig, +1.  Cin = 2: kc pair = (support k = 4*nt + tig, channels 0 and 1)
        {
            const long long ra = r0 + g, rb = ra + 8;
            long long ta, qa, tb, qb;
            split(cur_t0, cur_q0, g, ta, qa);
            split(cur_t0, cur_q0, g + 8, tb, qb);
            if (NODE) {
                if (tig < K) {
                    if (ra < rows_total) st2(a.DPX + xoff(ta, qa, 2 * tig), dx0[0], dx0[1]);
                    if (rb < rows_total) st2(a.DPX + xoff(tb, qb, 2 * tig), dx0[2], dx0[3]);
                }
                if (4 + tig < K) {
                    if (ra < rows_total) st2(a.DPX + xoff(ta, qa, 8 + 2 * tig), dx1[0], dx1[1]);
                    if (rb < rows_total) st2(a.DPX + xoff(tb, qb, 8 + 2 * tig), dx1[2], dx1[3]);
                }
            } else if (tig == 0) {   // the residual path's share of dx is ADDED into DPX[t, 0]
                if (ra < rows_total) { float* d = a.DPX + xoff(ta, qa, 0); const float2 o = ld2(d); st2(d, o.x + dx0[0], o.y + dx0[1]); }
                if (rb < rows_total) { float* d = a.DPX + xoff(tb, qb, 0); const float2 o = ld2(d); st2(d, o.x + dx0[2], o.y + dx0[3]); }
            }
        }
        // ---- (2) dW3 += [X | 1]^T x G_tile : M = kc (16), N = 192 (24 n-tiles), K = 16 rows (2 k-steps, rows paired 2tig / 2tig+1) ----
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const long long rlo = r0 + 8 * ks + 2 * tig, rhi = rlo + 1;
            uint32_t af[4];
            // A[m = kc][k = row]: a0 = (g, row lo), a1 = (g+8, row lo), a2 = (g, row hi), a3 = (g+8, row hi)
            float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
            long long tl, ql, th, qh;
            split(cur_t0, cur_q0, 8 * ks + 2 * tig, tl, ql);
            split(cur_t0, cur_q0, 8 * ks + 2 * tig + 1, th, qh);
            if (rlo < rows_total) {
                if (g < KC) x0 = a.PX[xoff(tl, ql, g)]; else if (g == KC) x0 = 1.f;
                if (g + 8 < KC) x1 = a.PX[xoff(tl, ql, g + 8)]; else if (g + 8 == KC) x1 = 1.f;
            }
            if (rhi < rows_total) {
                if (g < KC) x2 = a.PX[xoff(th, qh, g)]; else if (g == KC) x2 = 1.f;
                if (g + 8 < KC) x3 = a.PX[xoff(th, qh, g + 8)]; else if (g + 8 == KC) x3 = 1.f;
            }
            af[0] = __float_as_uint(x0); af[1] = __float_as_uint(x1); af[2] = __float_as_uint(x2); af[3] = __float_as_uint(x3);
            const float* blo = buf + (8 * ks + 2 * tig) * XM_LD + g;     // B[k = row][n = o]: b0 = (row lo, o = 8nt+g), b1 = (row hi, ..)
#pragma unroll
            for (int nt = 0; nt < 24; ++nt)
                mma_tf32_16x8x8(acc[nt], af, __float_as_uint(blo[8 * nt]), __float_as_uint(blo[XM_LD + 8 * nt]));
        }
        __syncwarp();   // every lane is done with buf before the next iteration's cp.async overwrites the other buffer's twin
        cur ^= 1;
    }
    // ---- block reduction of dW3 (rows 0..KC; row KC = bias gradient): every warp parks its fragment accumulators in its own
    // (now idle) tile buffer, then the block sums the eight partials per element - no atomics, no zero fill ----
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 24; ++nt) {
        const int o = 8 * nt + 2 * tig;
        *reinterpret_cast<float2*>(tilebuf + g * XM_LD + o) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(tilebuf + (g + 8) * XM_LD + o) = make_float2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
    const float* part = xm_smem + 16 * XM_LD;   // warp w's partial: part + w * 2 * XM_TILE, [16][XM_LD]
    for (int j = threadIdx.x; j < XM_H3 * (KC + 1); j += blockDim.x) {
        const int kc = j / XM_H3, o = j - kc * XM_H3;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < XM_WARPS; ++w) v += part[w * 2 * XM_TILE + kc * XM_LD + o];
        if (NODE) {
            if (kc == KC) {
                if (o < 2 * H) a.dba[(long long)n * 2 * H + o] = v; else a.dbb[(long long)n * H + o - 2 * H] = v;
            } else {
                const int k = kc / Cin, i = kc - k * Cin;
                if (o < 2 * H) a.dWa[(((long long)n * K + k) * I + i) * 2 * H + o] = v;
                else a.dWb[(((long long)n * K + k) * I + i) * H + o - 2 * H] = v;
            }
        } else {
            if (kc == KC) atomicAdd(o < 2 * H ? a.dba + o : a.dbb + (o - 2 * H), v);
            else atomicAdd(o < 2 * H ? a.dWa + (long long)o * I + kc : a.dWb + (long long)(o - 2 * H) * I + kc, v);
        }
    }
}

inline bool xside_mma_ok(int Cin, int H, int K, const void* G, const void* PX, const void* DPX) {
    return Cin == 2 && H == 64 && K * Cin <= 14 && aligned16(G) && !(reinterpret_cast<uintptr_t>(PX) & 7) && !(reinterpret_cast<uintptr_t>(DPX) & 7);
}

template <bool NODE>
inline cudaError_t launch_xside_bwd_mma(const XsMmaArgs& a, cudaStream_t st) {
    static bool configured[64] = {};   // per device
    const size_t smem = sizeof(float) * XM_SMEM_FLOATS;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(xside_bwd_mma_kernel<NODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const unsigned grid = NODE ? (unsigned)a.N : (unsigned)sm_count();
    xside_bwd_mma_kernel<NODE><<<grid, XM_WARPS * 32, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace matgcn
