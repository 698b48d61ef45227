// Persistent forward recurrence of one encoder layer (bf16 mode, rnn_units = 64): all T time steps of
//   PH[t,1..] = M h_{t-1}  ->  gate (per node)  ->  PZ[t,1..] = M (z h)  ->  candidate + residual GRU cell + mix -> h_t
// (MA.py:120-128, 142-150, 200-211) run as ONE cooperative launch instead of 4 launches per step.
//
// Why: one launch of the per-phase path does <= 3 tiles per SM and pays the launch floor, a cold instruction cache and a
// full pipeline ramp / drain every time (~25 us per phase for ~5 us of work at the Baltimore shape).  Here the four
// phases of a step are separated by a grid-wide barrier (one L2 atomic + an acquire spin, ~1 us) and the tile ->
// CTA assignment is static, so that
//   * every generic-proxy read of data produced inside the kernel (h_{t-1}, r) is a read of what the SAME CTA wrote;
//   * the only data that crosses CTAs are the bf16 operand twins, written by epilogue warps (generic proxy) and read by
//     TMA (async proxy) after the barrier;
//   * the per-node weight blocks of a CTA's nodes are the same every step (L2-resident, evict-last).
//
// Roles (384 threads, one CTA per SM): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator,
// warp 3 = L2 prefetch of the next phase's epilogue inputs, warps 4-11 = epilogue.  4-stage ring of 32 KB (A 16 KB | B 16 KB),
// two 128-column fp32 accumulators in TMEM.
#pragma once
#include "epilogues.cuh"
#include "gemm_tc.cuh"
#include "rec_api.h"

namespace matgcn {

constexpr int RF_STAGES = 4;
constexpr int RF_A_BYTES = 16384;
constexpr int RF_STAGE_BYTES = 32768;
constexpr int RF_EPI_BYTES = TC_EPI_WARPS * 32 * TC_EPI_LD * 4;
constexpr int RF_BAR_BYTES = 256;
constexpr int RF_RES_BYTES = (3 * 64) * TC_RES_LD * 4;
constexpr int RF_SMEM_TOTAL = RF_STAGES * RF_STAGE_BYTES + RF_EPI_BYTES + RF_BAR_BYTES + RF_RES_BYTES + 1024;

struct RecMaps {
    CUtensorMap M;    // base matrices (A of the propagation): {N, Kp*N}
    CUtensorMap Hs;   // h_{t-1} as the B operand of the propagation: PH16 {B*64, N, slot}
    CUtensorMap Zs;   // z*h as the B operand of the propagation: PZ16 {B*64, N, slot}
    CUtensorMap PHa;  // propagated state as the A operand of the gate contraction: PH16 {64, B, N, slot}
    CUtensorMap PZa;  // same for the candidate contraction: PZ16 {64, B, N, slot}
    CUtensorMap WG;   // per-node gate weights (B operand): WG16 {128, I, K, N}
    CUtensorMap WU;   // per-node candidate weights: WU16 {64, I, K, N}
};

struct RecFwdP {
    int T, N, B, K, Cin;
    int m64;            // B <= 64: M = 64 MMAs for the per-node contractions
    int node_tiles_m;   // row tiles per node (ceil(B / 128); 1 when m64)
    int prop_tiles_m, prop_tiles_n, prop_kt;
    long long U;        // N * B * 64
    const float* GX; const float* RX;      // [T, N*B, 3H]
    float* PH; float* PZ;                  // fp32 slot arrays (only slot 0 of every step is written)
    float* Z; float* R; float* HC; float* H1; float* Z2; float* R2; float* HC2; float* ZH2;  // [T, N*B, H]
    const float* RgH; const float* RuH; const float* mix;
    __nv_bfloat16* PH16; __nv_bfloat16* PZ16;
    unsigned int* gbar;  // zeroed grid-barrier counter
    long long* dbg;      // optional timeline of CTA 0 (tools/rec_timeline.py)
};

__device__ __forceinline__ unsigned int rf_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rf_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void rf_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t rf_pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// 32-byte global accesses (one full sector per lane)
__device__ __forceinline__ void rf_st8(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void rf_st8f(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void rf_ld8f(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

__global__ void __launch_bounds__(TC_THREADS, 1) rec_fwd_kernel(const __grid_constant__ RecMaps maps, const RecFwdP p) {
    constexpr int H = 64;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* stage_base = smem;
    float* epi_buf = reinterpret_cast<float*>(smem + RF_STAGES * RF_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RF_STAGES * RF_STAGE_BYTES + RF_EPI_BYTES);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], phase_bar; then the TMEM base slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RF_STAGES + 5);
    volatile uint32_t* phase_cnt = tmem_slot + 1;   // number of grid barriers this CTA has passed (polled by the prefetch warp)
    float* res_w = reinterpret_cast<float*>(smem + RF_STAGES * RF_STAGE_BYTES + RF_EPI_BYTES + RF_BAR_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full0 = bar0, empty0 = bar0 + 8u * RF_STAGES, tfull0 = bar0 + 8u * (2 * RF_STAGES);
    const uint32_t tempty0 = tfull0 + 16u, phase_bar = tfull0 + 32u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RF_STAGES; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8u * a, 1);
            mbar_init(tempty0 + 8u * a, TC_EPI_WARPS);
        }
        mbar_init(phase_bar, 1);
        *phase_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int T = p.T, K = p.K;
    const int G = gridDim.x;
    const int prop_tiles = p.prop_tiles_m * p.prop_tiles_n;
    const int node_tiles = p.N * p.node_tiles_m;
    const uint32_t a_rows_bytes = p.m64 ? 8192u : 16384u;   // A box of the per-node contractions: 64 or 128 rows of 128 bytes

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, nbar = 0;
            const uint64_t pol = l2_policy_evict_last();
            for (int t = 0; t < T; ++t) {
                for (int ph = 0; ph < 4; ++ph) {
                    if (t | ph) {
                        mbar_wait(phase_bar, nbar & 1u);  // the previous phase is complete on every CTA
                        ++nbar;
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    if (ph == 0 || ph == 2) {
                        const CUtensorMap* tb = ph == 0 ? &maps.Hs : &maps.Zs;
                        const int slot = t * K;
                        for (int tile = blockIdx.x; tile < prop_tiles; tile += G) {
                            const int tn = tile / p.prop_tiles_m, tm = tile - tn * p.prop_tiles_m;
                            for (int kt = 0; kt < p.prop_kt; ++kt) {
                                mbar_wait(empty0 + 8u * stage, phase ^ 1);
                                const uint32_t sa = smem_u32(stage_base + stage * RF_STAGE_BYTES), sb = sa + RF_A_BYTES;
                                const uint32_t fb = full0 + 8u * stage;
                                mbar_expect_tx(fb, RF_STAGE_BYTES);
                                tma_load_5d_hint(sa, &maps.M, fb, kt * 64, tm * 128, 0, 0, 0, pol);
                                tma_load_5d(sb, tb, fb, tn * 128, kt * 64, slot, 0, 0);
                                tma_load_5d(sb + 8192, tb, fb, tn * 128 + 64, kt * 64, slot, 0, 0);
                                if (++stage == RF_STAGES) { stage = 0; phase ^= 1; }
                            }
                        }
                    } else {
                        const bool gate = ph == 1;
                        const CUtensorMap* ta = gate ? &maps.PHa : &maps.PZa;
                        const CUtensorMap* tw = gate ? &maps.WG : &maps.WU;
                        const uint32_t tx = a_rows_bytes + (gate ? 16384u : 8192u);
                        for (int tile = blockIdx.x; tile < node_tiles; tile += G) {
                            const int n = tile / p.node_tiles_m, m0 = (tile - n * p.node_tiles_m) * 128;
                            for (int k = 0; k < K; ++k) {
                                mbar_wait(empty0 + 8u * stage, phase ^ 1);
                                const uint32_t sa = smem_u32(stage_base + stage * RF_STAGE_BYTES), sb = sa + RF_A_BYTES;
                                const uint32_t fb = full0 + 8u * stage;
                                mbar_expect_tx(fb, tx);
                                tma_load_5d(sa, ta, fb, 0, m0, n, t * K + k, 0);
                                tma_load_5d_hint(sb, tw, fb, 0, p.Cin, k, n, 0, pol);
                                if (gate) tma_load_5d_hint(sb + 8192, tw, fb, 64, p.Cin, k, n, 0, pol);
                                if (++stage == RF_STAGES) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // instruction descriptors: D = f32, A = B = bf16, A K-major, B MN-major, N >> 3, M >> 4
            const uint32_t id_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16);
            const uint32_t id_prop = id_base | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t mnode = p.m64 ? 64u : 128u;
            const uint32_t id_gate = id_base | ((uint32_t)(128 >> 3) << 17) | ((mnode >> 4) << 24);
            const uint32_t id_cand = id_base | ((uint32_t)(64 >> 3) << 17) | ((mnode >> 4) << 24);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int t = 0; t < T; ++t) {
                for (int ph = 0; ph < 4; ++ph) {
                    const bool prop = ph == 0 || ph == 2;
                    const int ntiles = prop ? prop_tiles : node_tiles;
                    const int nk = prop ? p.prop_kt : K;
                    const uint32_t idesc = prop ? id_prop : (ph == 1 ? id_gate : id_cand);
                    for (int tile = blockIdx.x; tile < ntiles; tile += G) {
                        mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1);
                        tc_fence_after();
                        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 128);
                        for (int kt = 0; kt < nk; ++kt) {
                            mbar_wait(full0 + 8u * stage, phase);
                            tc_fence_after();
                            const uint32_t sa = smem_u32(stage_base + stage * RF_STAGE_BYTES), sb = sa + RF_A_BYTES;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint64_t da = umma_desc(sa + kk * 32, 16, 1024, 2);
                                const uint64_t db = umma_desc(sb + kk * 2048, 8192, 1024, 2);
                                umma_bf16(tmem_d, da, db, idesc, (kt > 0 || kk > 0) ? 1u : 0u);
                            }
                            umma_commit(empty0 + 8u * stage);
                            if (++stage == RF_STAGES) { stage = 0; phase ^= 1; }
                        }
                        umma_commit(tfull0 + 8u * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ================================ L2 prefetch of epilogue inputs ================================
        // While a propagation phase runs (L2 -> SM bound, HBM idle) pull the pre-activation rows the NEXT per-node phase of
        // this CTA reads: GX[t] gate columns before the gate phase, GX[t] candidate columns and RX[t] before the tail.
        for (int t = 0; t < T; ++t) {
            const float* GXt = p.GX + (long long)t * 3 * p.U;
            const float* RXt = p.RX + (long long)t * 3 * p.U;
            while (*phase_cnt < (uint32_t)(4 * t)) __nanosleep(256);   // step t has begun
            for (int tile = blockIdx.x; tile < node_tiles; tile += G) {
                const int n = tile / p.node_tiles_m, m0 = (tile - n * p.node_tiles_m) * 128;
                const int rows = min(p.m64 ? 64 : 128, p.B - m0);
                const long long g0 = (long long)n * p.B + m0;
                // rows of 3H floats = 6 lines of 128 bytes each: GX whole row (gate: 4 lines, candidate: 2), RX whole row
                for (int idx = lane; idx < rows * 6; idx += 32) {
                    const long long off = (g0 + idx / 6) * 3 * H + (idx % 6) * 32;
                    pf_l2(GXt + off);
                    pf_l2(RXt + off);
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        const int q = warp & 3;             // TMEM lane quadrant
        const int half = (warp - 4) >> 2;   // which of the two warps of that quadrant
        float* buf = epi_buf + (warp - 4) * (32 * TC_EPI_LD);
        float* buf_other = epi_buf + ((warp - 4) ^ 4) * (32 * TC_EPI_LD);
        {
            // residual-cell weights Rg_h [128][64] and Ru_h [64][64] -> padded tiles in shared memory, once
            const int et = threadIdx.x - 128;
            for (int idx = et; idx < 3 * 64 * 16; idx += TC_EPI_WARPS * 32) {
                const int n = idx >> 4, k4 = (idx & 15) * 4;
                const float4 v = n < 128 ? ld4(p.RgH + n * 64 + k4) : ld4(p.RuH + (n - 128) * 64 + k4);
                *reinterpret_cast<float4*>(res_w + n * TC_RES_LD + k4) = v;
            }
            asm volatile("bar.sync 5, 256;" ::: "memory");
        }
        TcP tp;   // the fields the fused-tail epilogue reads
        tp.M = p.B;
        tp.m64 = p.m64;
        int acc = 0;
        uint32_t acc_phase = 0, nbar = 0;
        const int rows_q = p.m64 ? 16 : 32;
        const int ldc = p.B * H;
        const long long prop_rows = (long long)(K - 1) * p.N;
        for (int t = 0; t < T; ++t) {
            const long long tU = (long long)t * p.U;
            const float* GXt = p.GX + 3 * tU;
            const float* RXt = p.RX + 3 * tU;
            float* PHt = p.PH + (long long)t * K * p.U;     // h_{t-1} (fp32)
            float* PZt = p.PZ + (long long)t * K * p.U;     // z*h (fp32)
            __nv_bfloat16* PH16t = p.PH16 + (long long)t * K * p.U;
            __nv_bfloat16* PZ16t = p.PZ16 + (long long)t * K * p.U;
            for (int ph = 0; ph < 4; ++ph) {
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[(t * 4 + ph) * 4 + 0] = clock64();
                if (ph == 0 || ph == 2) {
                    // ---- propagation: bf16 twin of the accumulator -> slots 1.. of PH16 / PZ16 ----
                    __nv_bfloat16* dst = (ph == 0 ? PH16t : PZ16t) + p.U;
                    for (int tile = blockIdx.x; tile < prop_tiles; tile += G) {
                        const int tn = tile / p.prop_tiles_m, tm = tile - tn * p.prop_tiles_m;
                        const long long row = (long long)tm * 128 + q * 32 + lane;
                        mbar_wait(tfull0 + 8u * acc, acc_phase);
                        tc_fence_after();
                        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 128) + ((uint32_t)(q * 32) << 16);
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            const int c = half + 2 * cc;
                            const int col = tn * 128 + c * 32;
                            uint32_t r[32];
                            rf_tmem_ld32(tmem_acc + (uint32_t)(c * 32), r);
                            rf_tmem_wait_ld();
                            if (row < prop_rows && col < ldc) {
                                __nv_bfloat16* o = dst + row * ldc + col;
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    uint32_t w[8];
#pragma unroll
                                    for (int u = 0; u < 8; ++u)
                                        w[u] = rf_pack_bf16(__uint_as_float(r[16 * j + 2 * u]), __uint_as_float(r[16 * j + 2 * u + 1]));
                                    rf_st8(o + 16 * j, w);
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                } else if (ph == 1) {
                    // ---- gate: z = sigma(acc + GX[:, 0:H]), r = sigma(acc + GX[:, H:2H]); z*h -> slot 0 of PZ / PZ16 ----
                    float* Zt = p.Z + tU;
                    float* Rt = p.R + tU;
                    for (int tile = blockIdx.x; tile < node_tiles; tile += G) {
                        const int n = tile / p.node_tiles_m, m0 = (tile - n * p.node_tiles_m) * 128;
                        const int row = m0 + q * rows_q + lane;
                        const bool ok = lane < rows_q && row < p.B;
                        const long long g = (long long)n * p.B + row;
                        const int cz = half * 32;       // this warp's columns of the z half (and, + H, of the r half)
                        float gz[32], hz[32];
                        if (ok) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                rf_ld8f(GXt + g * 3 * H + cz + 8 * j, gz + 8 * j);
                                rf_ld8f(PHt + g * H + cz + 8 * j, hz + 8 * j);
                            }
                        }
                        mbar_wait(tfull0 + 8u * acc, acc_phase);
                        tc_fence_after();
                        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 128) + ((uint32_t)(q * 32) << 16);
                        {
                            uint32_t r[32];
                            rf_tmem_ld32(tmem_acc + (uint32_t)cz, r);
                            rf_tmem_wait_ld();
                            float gr[32];
                            if (ok) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) rf_ld8f(GXt + g * 3 * H + H + cz + 8 * j, gr + 8 * j);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float z[8], zh[8];
                                    uint32_t w[4];
#pragma unroll
                                    for (int u = 0; u < 8; ++u) {
                                        z[u] = sigmoid_fast(__uint_as_float(r[8 * j + u]) + gz[8 * j + u]);
                                        zh[u] = z[u] * hz[8 * j + u];
                                    }
#pragma unroll
                                    for (int u = 0; u < 4; ++u) w[u] = rf_pack_bf16(zh[2 * u], zh[2 * u + 1]);
                                    rf_st8f(Zt + g * H + cz + 8 * j, z);
                                    rf_st8f(PZt + g * H + cz + 8 * j, zh);
                                    *reinterpret_cast<uint4*>(PZ16t + g * H + cz + 8 * j) = make_uint4(w[0], w[1], w[2], w[3]);
                                }
                            }
                            rf_tmem_ld32(tmem_acc + (uint32_t)(H + cz), r);
                            rf_tmem_wait_ld();
                            if (ok) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float rr[8];
#pragma unroll
                                    for (int u = 0; u < 8; ++u) rr[u] = sigmoid_fast(__uint_as_float(r[8 * j + u]) + gr[8 * j + u]);
                                    rf_st8f(Rt + g * H + cz + 8 * j, rr);
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                } else {
                    // ---- candidate + residual GRU cell + mix (gemm_tc.cuh: tc_epilogue_tile_candres) ----
                    EpiCandRes ef{GXt, PHt, p.R + tU, p.HC + tU, p.H1 + tU, p.B, H, 1, RXt, p.Z2 + tU, p.R2 + tU, p.ZH2 + tU, p.HC2 + tU,
                                  PHt + (long long)K * p.U, p.mix + t, PH16t + (long long)K * p.U, p.RgH, p.RuH};
                    for (int tile = blockIdx.x; tile < node_tiles; tile += G) {
                        const int n = tile / p.node_tiles_m, m0 = (tile - n * p.node_tiles_m) * 128;
                        tc_epilogue_tile_candres(ef, tp, tmem_base + (uint32_t)(acc * 128), tfull0 + 8u * acc, acc_phase, q, half, buf, buf_other,
                                                 res_w, res_w + 128 * TC_RES_LD, lane, n, m0);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                }
                // ---- end of phase: publish this CTA's writes, wait for every CTA ----
                if (t == T - 1 && ph == 3) break;
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[(t * 4 + ph) * 4 + 1] = clock64();
                asm volatile("bar.sync 6, 256;" ::: "memory");
                if (threadIdx.x == 128) {
                    if (p.dbg && blockIdx.x == 0) p.dbg[(t * 4 + ph) * 4 + 2] = clock64();
                    __threadfence();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    atomicAdd(p.gbar, 1u);
                    const unsigned int target = (nbar + 1u) * (unsigned int)G;
                    long long t0 = 0;
                    for (uint32_t it = 0; rf_ld_acquire(p.gbar) < target; ++it) {
                        if (it == 1024) t0 = clock64();
                        if (it > 1024 && (it & 255) == 0 && clock64() - t0 > 4000000000LL) __trap();
                    }
                    __threadfence();
                    if (p.dbg && blockIdx.x == 0) p.dbg[(t * 4 + ph) * 4 + 3] = clock64();
                    *phase_cnt = nbar + 1u;
                    mbar_arrive(phase_bar);
                }
                ++nbar;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// bf16 tensor map of rank <= 5 (padded to 5 with unit dimensions), 128-byte swizzle; strides in bytes for dims 1..rank-1
inline bool rf_make_map(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides,
                        const unsigned int* box) {
    TmapEncodeFn enc = tmap_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    cuuint64_t d[5] = {1, 1, 1, 1, 1};
    cuuint64_t s[4] = {0, 0, 0, 0};
    cuuint32_t b[5] = {1, 1, 1, 1, 1};
    cuuint32_t e[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i < rank - 1; ++i) s[i] = strides[i];
    for (int i = rank - 1; i < 4; ++i) s[i] = (i == 0 ? d[0] * 2 : s[i - 1] * d[i]);   // packed continuation for the unit dimensions
    for (int i = 0; i < 4; ++i)
        if (s[i] == 0 || (s[i] & 15) || s[i] >= (1ULL << 40)) return false;
    for (int i = 0; i < 5; ++i)
        if (d[i] == 0 || d[i] > 0xffffffffULL || b[i] == 0 || b[i] > 256) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


// cudaErrorNotSupported: the shape does not meet the kernel's requirements (the caller runs one launch per phase instead)
cudaError_t launch_rec_fwd(const RecFwdArgs& a, cudaStream_t st) {
    constexpr int H = 64;
    const int Kp = a.K - 1, I = a.Cin + H;
    if (a.B < 8 || (a.ldm & 7) || a.K < 2 || a.N < 1 || a.T < 1) return cudaErrorNotSupported;
    const unsigned long long U = (unsigned long long)a.N * a.B * H;
    const unsigned long long slots_h = (unsigned long long)a.T * a.K + 1, slots_z = (unsigned long long)a.T * a.K;
    RecFwdP p;
    memset(&p, 0, sizeof(p));
    p.T = a.T; p.N = a.N; p.B = a.B; p.K = a.K; p.Cin = a.Cin;
    p.m64 = a.B <= 64 ? 1 : 0;
    p.node_tiles_m = p.m64 ? 1 : (a.B + 127) / 128;
    p.prop_tiles_m = (Kp * a.N + 127) / 128;
    p.prop_tiles_n = (a.B * H + 127) / 128;
    p.prop_kt = (a.N + 63) / 64;
    p.U = (long long)U;
    p.GX = a.GX; p.RX = a.RX; p.PH = a.PH; p.PZ = a.PZ;
    p.Z = a.Z; p.R = a.R; p.HC = a.HC; p.H1 = a.H1; p.Z2 = a.Z2; p.R2 = a.R2; p.HC2 = a.HC2; p.ZH2 = a.ZH2;
    p.RgH = a.RgH; p.RuH = a.RuH; p.mix = a.mix;
    p.PH16 = a.PH16; p.PZ16 = a.PZ16;
    p.gbar = a.gbar;
    p.dbg = tc_debug_buffer();
    const float* al[] = {a.GX, a.RX, a.PH, a.PZ, a.Z, a.R, a.HC, a.H1, a.Z2, a.R2, a.HC2, a.ZH2, a.RgH, a.RuH};
    for (const float* q : al)
        if (reinterpret_cast<uintptr_t>(q) & 31) return cudaErrorNotSupported;   // 32-byte accesses
    if ((reinterpret_cast<uintptr_t>(a.PH16) & 31) || (reinterpret_cast<uintptr_t>(a.PZ16) & 31)) return cudaErrorNotSupported;

    RecMaps maps;
    const unsigned int arows = p.m64 ? 64u : 128u;
    {
        const unsigned long long d[2] = {(unsigned long long)a.N, (unsigned long long)Kp * a.N}, s[1] = {(unsigned long long)a.ldm * 2};
        const unsigned int b[2] = {64, 128};
        if (!rf_make_map(&maps.M, a.M16, 2, d, s, b)) return cudaErrorNotSupported;
    }
    {
        const unsigned long long s[2] = {(unsigned long long)a.B * H * 2, U * 2};
        const unsigned int b[3] = {64, 64, 1};
        const unsigned long long dh[3] = {(unsigned long long)a.B * H, (unsigned long long)a.N, slots_h};
        const unsigned long long dz[3] = {(unsigned long long)a.B * H, (unsigned long long)a.N, slots_z};
        if (!rf_make_map(&maps.Hs, a.PH16, 3, dh, s, b) || !rf_make_map(&maps.Zs, a.PZ16, 3, dz, s, b)) return cudaErrorNotSupported;
    }
    {
        const unsigned long long s[3] = {(unsigned long long)H * 2, (unsigned long long)a.B * H * 2, U * 2};
        const unsigned int b[4] = {64, arows, 1, 1};
        const unsigned long long dh[4] = {64, (unsigned long long)a.B, (unsigned long long)a.N, slots_h};
        const unsigned long long dz[4] = {64, (unsigned long long)a.B, (unsigned long long)a.N, slots_z};
        if (!rf_make_map(&maps.PHa, a.PH16, 4, dh, s, b) || !rf_make_map(&maps.PZa, a.PZ16, 4, dz, s, b)) return cudaErrorNotSupported;
    }
    {
        const unsigned int b[4] = {64, 64, 1, 1};
        const unsigned long long dg[4] = {128, (unsigned long long)I, (unsigned long long)a.K, (unsigned long long)a.N};
        const unsigned long long sg[3] = {256, (unsigned long long)I * 256, (unsigned long long)a.K * I * 256};
        const unsigned long long du[4] = {64, (unsigned long long)I, (unsigned long long)a.K, (unsigned long long)a.N};
        const unsigned long long su[3] = {128, (unsigned long long)I * 128, (unsigned long long)a.K * I * 128};
        if (!rf_make_map(&maps.WG, a.WG16, 4, dg, sg, b) || !rf_make_map(&maps.WU, a.WU16, 4, du, su, b)) return cudaErrorNotSupported;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64) return cudaErrorNotSupported;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(rec_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_SMEM_TOTAL);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int most = max(p.prop_tiles_m * p.prop_tiles_n, a.N * p.node_tiles_m);
    const int grid = most < sms ? most : sms;
    cudaError_t e = cudaMemsetAsync(a.gbar, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&maps, (void*)&p};
    e = cudaLaunchCooperativeKernel((void*)rec_fwd_kernel, dim3((unsigned)grid), dim3(TC_THREADS), args, RF_SMEM_TOTAL, st);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace matgcn
