// Persistent forward recurrence of one encoder layer (bf16 mode, rnn_units = 64, at most 64 samples per GPU): all T steps of
//   PH[t,1..] = M h_{t-1}  ->  gate (per node)  ->  PZ[t,1..] = M (z h)  ->  candidate + residual GRU cell + mix -> h_t
// (MA.py:120-128, 142-150, 200-211) run as ONE cooperative launch instead of 4 launches per step.
//
// Why: one launch of the per-phase path does <= 3 tiles per SM and pays the launch floor, a cold instruction cache and a
// full pipeline ramp / drain every time.  Here the four phases of a step are separated by a grid-wide barrier (one L2 atomic
// + an acquire spin) and the tile -> CTA assignment is static, so that
//   * every generic-proxy read of data produced inside the kernel (h_{t-1}, r) is a read of what the SAME CTA wrote;
//   * the only data that crosses CTAs are the bf16 operand twins, written by epilogue warps (generic proxy) and read by
//     TMA (async proxy) after the barrier;
//   * the per-node weight blocks of a CTA's nodes are the same every step (evict-last in L2).
//
// Roles (384 threads, one CTA per SM): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator,
// warp 3 = L2 prefetch of the next per-node phase's epilogue inputs, warps 4-11 = epilogue.
// Shared memory: 4-stage operand ring of 32 KB (A 16 KB | B 16 KB), the residual-cell weights as resident TF32 operand
// tiles (48 KB), two 16 KB operand tiles the epilogue writes (h1, z2*h1), barriers.
// Tensor memory (512 columns): two 128-column accumulators for the streamed contractions, 128 + 64 columns for the
// residual-cell products.
//
// Epilogues read the accumulator of a 64-row tile (M = 64 MMA: 16 rows per TMEM lane quadrant) with tcgen05.ld.16x256b,
// whose register layout (thread -> rows lane/4 and lane/4 + 8, column pairs 8j + 2*(lane%4); tools/ubench/ldtm_probe.cu)
// keeps all 32 lanes busy and gives 32-byte sectors per row quad; everything a tile's epilogue reads from global memory is
// requested before the accumulator is waited for.
// The tail (candidate -> residual GRU cell -> mix) keeps its two small products on the tensor cores: the epilogue warps
// write h1 (then z2*h1) as a 128B-swizzled K-major TF32 operand tile, the MMA warp multiplies it with the resident weights
// into TMEM and the same warps pick the result up again - three accumulator round trips per tile, no mma.sync.
#pragma once
#include "epilogues.cuh"
#include "gemm_tc.cuh"
#include "rec_api.h"

#include <utility>
#include <vector>

namespace matgcn {

constexpr int RF_STAGES = 4;
constexpr int RF_A_BYTES = 16384;
constexpr int RF_STAGE_BYTES = 32768;
constexpr int RF_WG_BYTES = 2 * 128 * 128;   // Rg_h [128 outputs][64 inputs] fp32: two 32-wide K slabs of 128 rows x 128 bytes
constexpr int RF_WU_BYTES = 2 * 64 * 128;    // Ru_h [64][64]
constexpr int RF_S_BYTES = 2 * 64 * 128;     // h1 / z2*h1 operand tile [64 rows][64] fp32
constexpr int RF_BAR_BYTES = 256;
constexpr int RF_OFF_WG = RF_STAGES * RF_STAGE_BYTES;
constexpr int RF_OFF_WU = RF_OFF_WG + RF_WG_BYTES;
constexpr int RF_OFF_S1 = RF_OFF_WU + RF_WU_BYTES;
constexpr int RF_OFF_S2 = RF_OFF_S1 + RF_S_BYTES;
constexpr int RF_OFF_BAR = RF_OFF_S2 + RF_S_BYTES;
constexpr int RF_SMEM_TOTAL = RF_OFF_BAR + RF_BAR_BYTES + 1024;
constexpr int RF_TMEM_D2 = 256, RF_TMEM_D3 = 384;

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
    int prop_tiles_m, prop_tiles_n, prop_kt;
    int tn_fast;         // dense-phase tile order: 1 = column tiles fastest (see rf_tile_decode)
    long long U;        // N * B * 64
    const float* GX; const float* RX;      // [T, N*B, 3H]
    float* PH; float* PZ;                  // fp32 slot arrays (only slot 0 of every step is written)
    float* Z; float* R; float* HC; float* H1; float* Z2; float* R2; float* HC2; float* ZH2;  // [T, N*B, H]
    const float* RgH; const float* RuH; const float* mix;
    __nv_bfloat16* PH16; __nv_bfloat16* PZ16;
    unsigned int* gbar;  // zeroed grid-barrier counter
    long long* dbg;      // optional timeline of CTA 0 (tools/rec_timeline.py)
    int prefetch;        // warp 3, during the propagation phases (L2 -> SM bound, HBM idle): bit 0 pulls GX[t] / RX[t] into L2, bit 1 the
                         // hidden-row weight blocks of this CTA's nodes that the per-node phase after it streams (MATGCN_REC_PF)
    const __nv_bfloat16* WG16; const __nv_bfloat16* WU16;   // raw pointers of the weight twins (for the prefetch)
    int stream_hint;     // bit 0: read-once inputs / write-once outputs carry the L2 evict-first hint, bit 1: so do the TMA reads of the
                         // propagated state rows (MATGCN_REC_HINT, default 3)
};

__device__ __forceinline__ unsigned int rf_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// non-blocking test of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
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
// 16 TMEM lanes x 32 columns: register 4j + w of a thread = (row lane/4 + 8*(w/2), column 8j + 2*(lane%4) + w%2)
__device__ __forceinline__ void rf_tmem_ld16x4(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
// Exchange with the neighbouring lane (lane ^ 1) so that every thread of a 16x256b.x4 fragment holds four float4 instead of
// eight column pairs: f[2w + m] = row lane/4 + 8w, columns 16m + 4*pc .. + 3 with pc = ((lane & 1) << 1) | ((lane >> 1) & 1).
// The four lanes of a row then cover 64 contiguous bytes per access: half as many L1 wavefronts as the pair layout.
__device__ __forceinline__ void rf_quad4(const float (&v)[16], float4 (&f)[4], bool odd) {
#pragma unroll
    for (int w = 0; w < 2; ++w) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int je = 8 * m + 2 * w, jo = 8 * m + 4 + 2 * w;   // registers of the column groups 2m and 2m + 1
            const float s0 = odd ? v[je] : v[jo], s1 = odd ? v[je + 1] : v[jo + 1];
            const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
            f[2 * w + m] = odd ? make_float4(r0, r1, v[jo], v[jo + 1]) : make_float4(v[je], v[je + 1], r0, r1);
        }
    }
}
__device__ __forceinline__ void rf_sts4(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void rf_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t rf_pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// 32-byte global store (one full sector per lane)
__device__ __forceinline__ void rf_st8(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ float2 rf_ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void rf_st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
// L2 eviction hints for the data that is touched once per launch (pre-activation inputs read, saved activations written): marked
// evict-first so that what IS reused every step - the per-node weight blocks (evict-last), the state twins - stays in the 126 MB L2
__device__ __forceinline__ uint64_t rf_policy_stream(int on) {
    uint64_t pol;
    if (on) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld4h(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st4h(float* p, const float4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// Dense-phase tile -> (row tile tm, column tile tn).  Row tiles fastest (the default at the BASELINE shapes: everything fits L2) makes
// the CTAs of a wave share one column block of the state and stream DIFFERENT row bands of the base-matrix stack; once that stack
// exceeds L2 (N = 8192: 537 MB of bf16) it would be re-read from HBM once per column tile.  Column tiles fastest lets the CTAs that
// run together share a row band and keeps the (small) state slab in L2: the stack is read once per phase.
__device__ __forceinline__ void rf_tile_decode(int tile, int tiles_m, int tiles_n, int tn_fast, int& tm, int& tn) {
    if (tn_fast) { tm = tile / tiles_n; tn = tile - tm * tiles_n; }
    else { tn = tile / tiles_m; tm = tile - tn * tiles_m; }
}
inline int rec_tn_fast(int Kp, int N) {
    const char* e = getenv("MATGCN_REC_TN_FAST");
    if (e) return atoi(e) != 0;
    return (long long)Kp * N * N * 2 > (48LL << 20);   // bf16 base-matrix stack larger than ~half of what L2 can keep
}
__device__ __forceinline__ void rf_proxy_fence_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// fine-grained timeline of CTA 0 at the middle time step: slot s of tile i (first four tiles of the CTA) of phase ph
#define RF_STAMP(ph, i, s)                                                                                                   \
    do {                                                                                                                     \
        if (p.dbg && blockIdx.x == 0 && t == (T >> 1) && (i) < 4) p.dbg[T * 16 + (((ph) * 4 + (i)) << 4) + (s)] = clock64(); \
    } while (0)
#define RF_STAMP_E(ph, i, s)                        \
    do {                                            \
        if (threadIdx.x == 128) RF_STAMP(ph, i, s); \
    } while (0)

__global__ void __launch_bounds__(TC_THREADS, 1) rec_fwd_kernel(const __grid_constant__ RecMaps maps, const RecFwdP p) {
    constexpr int H = 64;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RF_OFF_BAR);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], phase_bar, s1_full, s2_full, r2_full, r3_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RF_STAGES + 9);
    volatile uint32_t* phase_cnt = tmem_slot + 1;   // number of grid barriers this CTA has passed (polled by the prefetch warp)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full0 = bar0, empty0 = bar0 + 8u * RF_STAGES, tfull0 = bar0 + 8u * (2 * RF_STAGES);
    const uint32_t tempty0 = tfull0 + 16u, phase_bar = tfull0 + 32u;
    const uint32_t s1_full = phase_bar + 8u, s2_full = phase_bar + 16u, r2_full = phase_bar + 24u, r3_full = phase_bar + 32u;
    const uint32_t wg_s = smem_u32(smem + RF_OFF_WG), wu_s = smem_u32(smem + RF_OFF_WU);
    const uint32_t s1_s = smem_u32(smem + RF_OFF_S1), s2_s = smem_u32(smem + RF_OFF_S2);

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
        mbar_init(s1_full, TC_EPI_WARPS);
        mbar_init(s2_full, TC_EPI_WARPS);
        mbar_init(r2_full, 1);
        mbar_init(r3_full, 1);
        *phase_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp >= 4) {
        // residual-cell weights Rg_h [128][64], Ru_h [64][64] (fp32, row = output) -> K-major 128B-swizzled operand tiles:
        // element (o, i) -> slab i/32, row o, 16-byte chunk ((i%32)/4) ^ (o%8)
        const int et = threadIdx.x - 128;
        for (int idx = et; idx < 3 * 64 * 16; idx += TC_EPI_WARPS * 32) {
            const int o = idx >> 4, i4 = (idx & 15) * 4;
            const bool g = o < 128;
            const int oo = g ? o : o - 128;
            const float4 v = g ? ld4(p.RgH + oo * 64 + i4) : ld4(p.RuH + oo * 64 + i4);
            uint8_t* dst = smem + (g ? RF_OFF_WG + (i4 >> 5) * (128 * 128) : RF_OFF_WU + (i4 >> 5) * (64 * 128)) + oo * 128 +
                           ((((i4 & 31) >> 2) ^ (oo & 7)) << 4);
            *reinterpret_cast<float4*>(dst) = v;
        }
        rf_proxy_fence_smem();   // generic-proxy writes -> tensor-core (async proxy) reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int T = p.T, K = p.K;
    const int G = gridDim.x;
    const int prop_tiles = p.prop_tiles_m * p.prop_tiles_n;
    const int node_tiles = p.N;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, nbar = 0;
            const uint64_t pol = l2_policy_evict_last();
            const uint64_t pol_once = rf_policy_stream(p.stream_hint & 2);   // propagated state rows: dead (for this launch) once read here
            for (int t = 0; t < T; ++t) {
                for (int ph = 0; ph < 4; ++ph) {
                    const bool prop = ph == 0 || ph == 2, gate = ph == 1;
                    const int ntiles = prop ? prop_tiles : node_tiles;
                    const int nk = prop ? p.prop_kt : K;
                    const uint32_t tx = prop ? (uint32_t)RF_STAGE_BYTES : 8192u + (gate ? 16384u : 8192u);
                    const CUtensorMap* ta = gate ? &maps.PHa : &maps.PZa;
                    const CUtensorMap* tw = gate ? &maps.WG : &maps.WU;
                    const CUtensorMap* tb = ph == 0 ? &maps.Hs : &maps.Zs;
                    const int slot = t * K;
                    // the operand of a k-block that does NOT depend on the previous phase (base matrices / per-node weights)
                    // (row / column tile of a dense tile: decoded once per tile, not once per k-block - this thread's latency is on the
                    // critical path of the operand stream)
                    int c_tile = -1, c_tm = 0, c_tn = 0;
                    auto decode = [&](int tile) {
                        if (tile != c_tile) {
                            rf_tile_decode(tile, p.prop_tiles_m, p.prop_tiles_n, p.tn_fast, c_tm, c_tn);
                            c_tile = tile;
                        }
                    };
                    auto issue_const = [&](int tile, int kt, uint32_t sa, uint32_t fb) {
                        if (prop) {
                            decode(tile);
                            tma_load_5d_hint(sa, &maps.M, fb, kt * 64, c_tm * 128, 0, 0, 0, pol);
                        } else {
                            tma_load_5d_hint(sa + RF_A_BYTES, tw, fb, 0, p.Cin, kt, tile, 0, pol);
                            if (gate) tma_load_5d_hint(sa + RF_A_BYTES + 8192, tw, fb, 64, p.Cin, kt, tile, 0, pol);
                        }
                    };
                    // ... and the one that does (the state written by the other CTAs in the previous phase)
                    auto issue_state = [&](int tile, int kt, uint32_t sa, uint32_t fb) {
                        if (prop) {
                            decode(tile);
                            tma_load_5d(sa + RF_A_BYTES, tb, fb, c_tn * 128, kt * 64, slot, 0, 0);
                            tma_load_5d(sa + RF_A_BYTES + 8192, tb, fb, c_tn * 128 + 64, kt * 64, slot, 0, 0);
                        } else {
                            if (kt > 0) tma_load_5d_hint(sa, ta, fb, 0, 0, tile, slot + kt, 0, pol_once);
                            else tma_load_5d(sa, ta, fb, 0, 0, tile, slot + kt, 0);   // (slot 0 = the state itself: the dense phase's operand)
                        }
                    };
                    // before the grid barrier: arm the first stages of this CTA's first tile and request their constant operands
                    int pre = 0;
                    if ((t | ph) && (int)blockIdx.x < ntiles) {
                        int s2 = stage;
                        uint32_t p2 = phase;
                        for (; pre < nk && pre < RF_STAGES; ++pre) {
                            mbar_wait(empty0 + 8u * s2, p2 ^ 1);
                            const uint32_t fb = full0 + 8u * s2;
                            mbar_expect_tx(fb, tx);
                            issue_const(blockIdx.x, pre, smem_u32(stage_base + s2 * RF_STAGE_BYTES), fb);
                            if (++s2 == RF_STAGES) { s2 = 0; p2 ^= 1; }
                        }
                    }
                    if (t | ph) {
                        mbar_wait(phase_bar, nbar & 1u);  // the previous phase is complete on every CTA
                        ++nbar;
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    for (int tile = blockIdx.x; tile < ntiles; tile += G) {
                        RF_STAMP(ph, tile / G, 0);
                        for (int kt = 0; kt < nk; ++kt) {
                            const uint32_t sa = smem_u32(stage_base + stage * RF_STAGE_BYTES);
                            const uint32_t fb = full0 + 8u * stage;
                            if (pre > 0) {
                                --pre;   // armed above, constant operand already in flight
                            } else {
                                mbar_wait(empty0 + 8u * stage, phase ^ 1);
                                mbar_expect_tx(fb, tx);
                                issue_const(tile, kt, sa, fb);
                            }
                            issue_state(tile, kt, sa, fb);
                            if (++stage == RF_STAGES) { stage = 0; phase ^= 1; }
                        }
                        RF_STAMP(ph, tile / G, 1);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // instruction descriptors: D = f32; bf16 operands, A K-major, B MN-major (streamed contractions);
            // tf32 operands, both K-major (residual-cell products); N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t id_bf = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16);
            const uint32_t id_prop = id_bf | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t id_gate = id_bf | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            const uint32_t id_cand = id_bf | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            const uint32_t id_tf = (1u << 4) | (2u << 7) | (2u << 10);
            const uint32_t id_rg = id_tf | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            const uint32_t id_ru = id_tf | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, res_par = 0;
            int t = 0, ph = 0;
            auto issue_tile = [&](int nk, uint32_t idesc, int ti) {   // one streamed contraction into the next accumulator buffer
                mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1);
                tc_fence_after();
                RF_STAMP(ph, ti, 2);
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 128);
                for (int kt = 0; kt < nk; ++kt) {
                    mbar_wait(full0 + 8u * stage, phase);
                    tc_fence_after();
                    if (kt == 0) RF_STAMP(ph, ti, 3);
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
                RF_STAMP(ph, ti, 4);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            };
            for (t = 0; t < T; ++t) {
                for (ph = 0; ph < 3; ++ph) {
                    const bool prop = ph == 0 || ph == 2;
                    const int ntiles = prop ? prop_tiles : node_tiles;
                    for (int tile = blockIdx.x; tile < ntiles; tile += G) issue_tile(prop ? p.prop_kt : K, prop ? id_prop : id_gate, tile / G);
                }
                // tail: the candidate contractions of this CTA's tiles and the residual-cell products the epilogue warps ask for
                // are interleaved by readiness (non-blocking barrier tests): a residual product is eight MMAs on the critical
                // path of a tile's epilogue and must not queue behind the operand stream of the next tile
                ph = 3;
                int mt = blockIdx.x, rt = blockIdx.x, mk = 0, rstage = 0;
                bool open = false;
                long long spin0 = 0;
                for (uint32_t it = 0; rt < node_tiles; ++it) {
                    bool progress = false;
                    if (rstage == 0 ? mbar_test(s1_full, res_par) : mbar_test(s2_full, res_par)) {
                        tc_fence_after();
                        if (rstage == 0) {   // h1 operand tile written: [z2 | r2] pre-activations
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma_tf32(tmem_base + RF_TMEM_D2, umma_desc(s1_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2),
                                          umma_desc(wg_s + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, 2), id_rg, kk > 0 ? 1u : 0u);
                            umma_commit(r2_full);
                            rstage = 1;
                        } else {             // z2*h1 operand tile written: candidate pre-activation of the residual cell
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma_tf32(tmem_base + RF_TMEM_D3, umma_desc(s2_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2),
                                          umma_desc(wu_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2), id_ru, kk > 0 ? 1u : 0u);
                            umma_commit(r3_full);
                            rstage = 0;
                            res_par ^= 1;
                            rt += G;
                        }
                        progress = true;
                    }
                    if (mt < node_tiles) {
                        if (!open && mbar_test(tempty0 + 8u * acc, acc_phase ^ 1)) {
                            tc_fence_after();
                            RF_STAMP(ph, mt / G, 2);
                            open = true;
                            mk = 0;
                        }
                        if (open && mbar_test(full0 + 8u * stage, phase)) {
                            tc_fence_after();
                            if (mk == 0) RF_STAMP(ph, mt / G, 3);
                            const uint32_t sa = smem_u32(stage_base + stage * RF_STAGE_BYTES), sb = sa + RF_A_BYTES;
                            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 128);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16(tmem_d, umma_desc(sa + kk * 32, 16, 1024, 2), umma_desc(sb + kk * 2048, 8192, 1024, 2), id_cand,
                                          (mk > 0 || kk > 0) ? 1u : 0u);
                            umma_commit(empty0 + 8u * stage);
                            if (++stage == RF_STAGES) { stage = 0; phase ^= 1; }
                            if (++mk == K) {
                                umma_commit(tfull0 + 8u * acc);
                                RF_STAMP(ph, mt / G, 4);
                                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                                open = false;
                                mt += G;
                            }
                            progress = true;
                        }
                    }
                    if (progress) {
                        it = 0;
                    } else {
                        if (it == 256) spin0 = clock64();
                        if (it > 256 && (it & 1023) == 0 && clock64() - spin0 > 4000000000LL) __trap();
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ================================ L2 prefetch of epilogue inputs ================================
        // While a propagation phase runs (L2 -> SM bound, HBM idle) pull the pre-activation rows the per-node phase after it
        // reads on this CTA: GX[t] during M*h, RX[t] during M*(z*h).
        const int I = p.Cin + H;
        for (int t = 0; t < T && p.prefetch; ++t) {
            for (int part = 0; part < 2; ++part) {
                const float* X = (part == 0 ? p.GX : p.RX) + (long long)t * 3 * p.U;
                while (*phase_cnt < (uint32_t)(4 * t + 2 * part)) __nanosleep(256);
                if (p.prefetch & 2) {
                    // per node and support one contiguous block: gate 64 rows x 128 bf16 = 16 KB, candidate 64 x 64 = 8 KB
                    const int ow = part == 0 ? 2 * H : H;
                    const __nv_bfloat16* W = part == 0 ? p.WG16 : p.WU16;
                    const int units = ((node_tiles - (int)blockIdx.x + G - 1) / G) * K;
                    for (int u = lane; u < units; u += 32) {
                        const int n = blockIdx.x + (u / K) * G, k = u - (u / K) * K;
                        const __nv_bfloat16* src = W + (((long long)n * K + k) * I + p.Cin) * ow;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(H * ow * 2) : "memory");
                    }
                }
                if (p.prefetch & 1) {
                    for (int n = blockIdx.x; n < node_tiles; n += G) {
                        const float* base = X + (long long)n * p.B * 3 * H;
                        for (int idx = lane; idx < p.B * 6; idx += 32) pf_l2(base + idx * 32);   // rows of 3H floats = 6 lines
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        const int q = warp & 3;             // TMEM lane quadrant
        const int half = (warp - 4) >> 2;   // which of the two warps of that quadrant
        int acc = 0;
        uint32_t acc_phase = 0, nbar = 0, res_par = 0;
        const int ldc = p.B * H;
        const long long prop_rows = (long long)(K - 1) * p.N;
        // coordinates of this thread inside a 64-row tile after rf_quad4: rows b0 and b0 + 8, columns ch + 16m .. + 3 (m = 0, 1)
        const int b0 = q * 16 + (lane >> 2);
        const bool odd = lane & 1;
        const int pc = ((lane & 1) << 1) | ((lane >> 1) & 1);
        const bool ok[2] = {b0 < p.B, b0 + 8 < p.B};
        const int ch = half * 32 + 4 * pc;
        // row offsets (relative to row b0 of the node) used for LOADS: rows past the batch read row 0 of the node instead
        const long long rd[2] = {ok[0] ? 0 : -(long long)b0, ok[1] ? 8 : -(long long)b0};
        // this thread's 16-byte slots in the operand tiles: slab `half`, rows b0 / b0 + 8 (same swizzle phase: 8 rows apart),
        // chunk (4m + pc) ^ (b0 % 8)
        const uint32_t s_off = (uint32_t)(half * 8192 + b0 * 128);
        const uint32_t s_x = (uint32_t)(b0 & 7);
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        const uint64_t pol = rf_policy_stream(p.stream_hint & 1);
        for (int t = 0; t < T; ++t) {
            for (int ph = 0; ph < 4; ++ph) {
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[(t * 4 + ph) * 4 + 0] = clock64();
                if (ph == 0 || ph == 2) {
                    // ---- propagation (128-row tiles, lane = row): bf16 twin of the accumulator -> slots 1.. of PH16 / PZ16 ----
                    __nv_bfloat16* dst = (ph == 0 ? p.PH16 : p.PZ16) + ((long long)t * K + 1) * p.U;
                    for (int tile = blockIdx.x; tile < prop_tiles; tile += G) {
                        int tm, tn;
                            rf_tile_decode(tile, p.prop_tiles_m, p.prop_tiles_n, p.tn_fast, tm, tn);
                        const long long row = (long long)tm * 128 + q * 32 + lane;
                        RF_STAMP_E(ph, tile / G, 5);
                        mbar_wait(tfull0 + 8u * acc, acc_phase);
                        tc_fence_after();
                        RF_STAMP_E(ph, tile / G, 6);
                        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 128) + tlane;
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
                        RF_STAMP_E(ph, tile / G, 7);
                        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                } else if (ph == 1) {
                    // ---- gate: z = sigma(acc + GX[:, 0:H]), r = sigma(acc + GX[:, H:2H]); z*h -> slot 0 of PZ / PZ16 ----
                    // per-step bases (derived below an empty asm so that nothing of it is hoisted above the phase: the per-step
                    // pointers of the other phases then do not stay live here)
                    long long tl = t;
                    asm volatile("" : "+l"(tl));
                    const long long tU = tl * p.U;
                    const float* GXt = p.GX + 3 * tU;
                    const float* PHt = p.PH + tl * K * p.U;
                    __nv_bfloat16* PZ16t = p.PZ16 + tl * K * p.U;
                    float* Zt = p.Z + tU;
                    float* Rt = p.R + tU;
                    float4 gz[4], hz[4], gr[4];   // [2w + m]: this thread's pre-activation inputs and h
                    // (rows past the batch are redirected to row 0 of the node: every load is unconditional, which keeps these
                    // arrays in registers; their results are never stored)
                    auto load_inputs = [&](int n) {
                        const long long g0 = (long long)n * p.B + b0;
                        const long long o0 = g0 * H + ch, x0 = g0 * 3 * H + ch;
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const long long o = o0 + rd[w] * H + 16 * m2, x = x0 + rd[w] * 3 * H + 16 * m2;
                                gz[2 * w + m2] = ld4h(GXt + x, pol);
                                hz[2 * w + m2] = ld4(PHt + o);
                                gr[2 * w + m2] = ld4h(GXt + x + H, pol);
                            }
                        }
                    };
                    load_inputs(min((int)blockIdx.x, node_tiles - 1));   // (unconditional: keeps the arrays in registers)
                    for (int n = blockIdx.x; n < node_tiles; n += G) {
                        const long long o0 = ((long long)n * p.B + b0) * H + ch;
                        RF_STAMP_E(ph, n / G, 5);
                        mbar_wait(tfull0 + 8u * acc, acc_phase);
                        tc_fence_after();
                        RF_STAMP_E(ph, n / G, 6);
                        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 128) + tlane;
                        float az[16], ar[16];
                        rf_tmem_ld16x4(tmem_acc + (uint32_t)(half * 32), az);
                        rf_tmem_ld16x4(tmem_acc + (uint32_t)(H + half * 32), ar);
                        rf_tmem_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);   // the accumulator is in registers: release it
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                        float4 fz[4], fr[4];
                        rf_quad4(az, fz, odd);
                        rf_quad4(ar, fr, odd);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            fz[e] = sigmoid4(fz[e] + gz[e], 1);
                            fr[e] = sigmoid4(fr[e] + gr[e], 1);
                            hz[e] = fz[e] * hz[e];
                        }
                        // the next tile's inputs are requested BEFORE this tile's stores enter the memory pipeline
                        float4 zh[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) zh[e] = hz[e];
                        load_inputs(min(n + G, node_tiles - 1));
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                if (ok[w]) {
                                    const int e = 2 * w + m2;
                                    const long long o = o0 + w * 8 * H + 16 * m2;
                                    st4h(Zt + o, fz[e], pol);
                                    st4_bf16(PZ16t + o, zh[e]);   // (every consumer of z*h reads the bf16 twin: no fp32 copy)
                                    st4(Rt + o, fr[e]);
                                }
                            }
                        }
                        RF_STAMP_E(ph, n / G, 7);
                    }
                } else {
                    // ---- tail: candidate, then the residual GRU cell (two TF32 products on the tensor cores) and the mix ----
                    const float m = __ldg(p.mix + t);
                    long long tl = t;
                    asm volatile("" : "+l"(tl));
                    const long long tU = tl * p.U;
                    const float* GXt = p.GX + 3 * tU;
                    const float* RXt = p.RX + 3 * tU;
                    const float* PHt = p.PH + tl * K * p.U;
                    const float* Rt = p.R + tU;
                    float* Yt = p.PH + (tl + 1) * K * p.U;
                    __nv_bfloat16* Y16t = p.PH16 + (tl + 1) * K * p.U;
                    float4 gc[4], rr[4], hh[4], xz[4], xr[4], xu[4];
                    auto load_stage1 = [&](int n) {   // candidate pre-activation input, r and h_{t-1} of a tile
                        const long long g0 = (long long)n * p.B + b0;
                        const long long o0 = g0 * H + ch, x0 = g0 * 3 * H + ch;
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const long long o = o0 + rd[w] * H + 16 * m2, x = x0 + rd[w] * 3 * H + 16 * m2;
                                gc[2 * w + m2] = ld4h(GXt + x + 2 * H, pol);
                                rr[2 * w + m2] = ld4(Rt + o);
                                hh[2 * w + m2] = ld4(PHt + o);
                            }
                        }
                    };
                    load_stage1(min((int)blockIdx.x, node_tiles - 1));   // (unconditional: keeps the arrays in registers)
                    for (int n = blockIdx.x; n < node_tiles; n += G) {
                        const long long g0 = (long long)n * p.B + b0;
                        const long long o0 = g0 * H + ch, x0 = g0 * 3 * H + ch;
                        RF_STAMP_E(ph, n / G, 5);
                        mbar_wait(tfull0 + 8u * acc, acc_phase);
                        tc_fence_after();
                        RF_STAMP_E(ph, n / G, 6);
                        float a[16], c2[16];
                        float4 fa[4], fc[4], h1[4], r2[4];
                        rf_tmem_ld16x4(tmem_base + (uint32_t)(acc * 128) + tlane + (uint32_t)(half * 32), a);
                        rf_tmem_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                        // (the residual cell's pre-activation inputs are requested one stage ahead of their use)
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const long long x = x0 + rd[w] * 3 * H + 16 * m2;
                                xz[2 * w + m2] = ld4h(RXt + x, pol);
                                xr[2 * w + m2] = ld4h(RXt + x + H, pol);
                            }
                        }
                        // stage 1: hc = tanh(acc + GX[:, 2H:]); h1 = r*h + (1-r)*hc -> HC, H1, operand tile S1
                        rf_quad4(a, fa, odd);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const int e = 2 * w + m2;
                                const float4 hc = tanh4(fa[e] + gc[e], 1);
                                h1[e] = rr[e] * hh[e] + one_minus(rr[e]) * hc;
                                if (ok[w]) {
                                    const long long o = o0 + w * 8 * H + 16 * m2;
                                    st4h(p.HC + tU + o, hc, pol);
                                    st4h(p.H1 + tU + o, h1[e], pol);
                                }
                                rf_sts4(s1_s + s_off + (uint32_t)(w * 1024) + ((((uint32_t)(4 * m2 + pc)) ^ s_x) << 4), h1[e]);
                            }
                        }
                        rf_proxy_fence_smem();
                        __syncwarp();
                        RF_STAMP_E(ph, n / G, 8);
                        if (lane == 0) mbar_arrive(s1_full);
                        load_stage1(min(n + G, node_tiles - 1));   // (gc / rr / hh are dead from here on: the next tile's go in flight)
                        // stage 2: [z2 | r2] = sigma(h1 Rg_h^T + RX[:, 0:2H]); z2*h1 -> Z2, R2, ZH2, operand tile S2
                        mbar_wait(r2_full, res_par);
                        tc_fence_after();
                        RF_STAMP_E(ph, n / G, 9);
                        rf_tmem_ld16x4(tmem_base + RF_TMEM_D2 + tlane + (uint32_t)(half * 32), a);
                        rf_tmem_ld16x4(tmem_base + RF_TMEM_D2 + tlane + (uint32_t)(H + half * 32), c2);
                        rf_tmem_wait_ld();
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) xu[2 * w + m2] = ld4h(RXt + x0 + rd[w] * 3 * H + 16 * m2 + 2 * H, pol);
                        }
                        rf_quad4(a, fa, odd);
                        rf_quad4(c2, fc, odd);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const int e = 2 * w + m2;
                                const float4 z2 = sigmoid4(fa[e] + xz[e], 1);
                                r2[e] = sigmoid4(fc[e] + xr[e], 1);
                                const float4 zh2 = z2 * h1[e];
                                if (ok[w]) {
                                    const long long o = o0 + w * 8 * H + 16 * m2;
                                    st4h(p.Z2 + tU + o, z2, pol);
                                    st4h(p.R2 + tU + o, r2[e], pol);
                                    st4h(p.ZH2 + tU + o, zh2, pol);
                                }
                                rf_sts4(s2_s + s_off + (uint32_t)(w * 1024) + ((((uint32_t)(4 * m2 + pc)) ^ s_x) << 4), zh2);
                            }
                        }
                        tc_fence_before();
                        rf_proxy_fence_smem();
                        __syncwarp();
                        RF_STAMP_E(ph, n / G, 10);
                        if (lane == 0) mbar_arrive(s2_full);
                        // stage 3: hc2 = tanh(z2*h1 Ru_h^T + RX[:, 2H:]); residual GRU output, mix -> HC2, h_t (fp32 + bf16 twin)
                        mbar_wait(r3_full, res_par);
                        tc_fence_after();
                        RF_STAMP_E(ph, n / G, 11);
                        rf_tmem_ld16x4(tmem_base + RF_TMEM_D3 + tlane + (uint32_t)(half * 32), a);
                        rf_tmem_wait_ld();
                        tc_fence_before();
                        rf_quad4(a, fa, odd);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                if (ok[w]) {
                                    const int e = 2 * w + m2;
                                    const long long o = o0 + w * 8 * H + 16 * m2;
                                    const float4 hc2 = tanh4(fa[e] + xu[e], 1);
                                    const float4 res = r2[e] * h1[e] + one_minus(r2[e]) * hc2;
                                    const float4 y = m * h1[e] + (1.f - m) * res;
                                    st4h(p.HC2 + tU + o, hc2, pol);
                                    st4(Yt + o, y);
                                    st4_bf16(Y16t + o, y);
                                }
                            }
                        }
                        res_par ^= 1;
                        RF_STAMP_E(ph, n / G, 7);
                    }
                }
                // ---- end of phase: publish this CTA's writes, wait for every CTA ----
                if (t == T - 1 && ph == 3) break;
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[(t * 4 + ph) * 4 + 1] = clock64();
                asm volatile("bar.sync 6, 256;" ::: "memory");
                if (threadIdx.x == 128) {
                    if (p.dbg && blockIdx.x == 0) p.dbg[(t * 4 + ph) * 4 + 2] = clock64();
                    // release: the writes of every epilogue thread (ordered before this point by the CTA barrier) become visible at
                    // gpu scope before the arrival is; the other CTAs' TMA reads them through the async proxy
                    asm volatile("fence.proxy.async;" ::: "memory");
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.gbar) : "memory");
                    const unsigned int target = (nbar + 1u) * (unsigned int)G;
                    long long t0 = 0;
                    for (uint32_t it = 0; rf_ld_acquire(p.gbar) < target; ++it) {
                        if (it == 1024) t0 = clock64();
                        if (it > 1024 && (it & 255) == 0 && clock64() - t0 > 4000000000LL) __trap();
                    }
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
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

struct RecTiming {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[2];   // [0] forward, [1] backward
};
inline RecTiming& rec_timing() {
    static RecTiming t;
    return t;
}
void rec_timing_enable(bool on) {
    RecTiming& t = rec_timing();
    for (auto& v : t.ev) {
        for (auto& e : v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        v.clear();
    }
    t.on = on;
}
void rec_timing_read(double* fwd_ms, int* fwd_n, double* bwd_ms, int* bwd_n) {
    RecTiming& t = rec_timing();
    double ms[2] = {0.0, 0.0};
    for (int i = 0; i < 2; ++i)
        for (auto& e : t.ev[i]) {
            cudaEventSynchronize(e.second);
            float f = 0.f;
            if (cudaEventElapsedTime(&f, e.first, e.second) == cudaSuccess) ms[i] += f;
        }
    *fwd_ms = ms[0]; *fwd_n = (int)t.ev[0].size();
    *bwd_ms = ms[1]; *bwd_n = (int)t.ev[1].size();
}
// cooperative launch, bracketed by events when the timing hook is on
inline cudaError_t rec_launch(const void* kern, int grid, void** args, int smem, cudaStream_t st, int which) {
    RecTiming& t = rec_timing();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (t.on) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(TC_THREADS), args, smem, st);
    if (t.on) {
        cudaEventRecord(e1, st);
        t.ev[which].emplace_back(e0, e1);
    }
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_rec_fwd(const RecFwdArgs& a, cudaStream_t st) {
    constexpr int H = 64;
    const int Kp = a.K - 1, I = a.Cin + H;
    if (a.B < 8 || a.B > 64 || (a.ldm & 7) || a.K < 2 || a.N < 1 || a.T < 1) return cudaErrorNotSupported;
    const unsigned long long U = (unsigned long long)a.N * a.B * H;
    const unsigned long long slots_h = (unsigned long long)a.T * a.K + 1, slots_z = (unsigned long long)a.T * a.K;
    RecFwdP p;
    memset(&p, 0, sizeof(p));
    p.T = a.T; p.N = a.N; p.B = a.B; p.K = a.K; p.Cin = a.Cin;
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
    {
        const char* e = getenv("MATGCN_REC_PF");
        p.prefetch = e ? atoi(e) & 3 : 0;   // bit 0 (GX / RX) measured within noise: the fill traffic costs the propagation what the epilogues gain
    }
    p.WG16 = a.WG16; p.WU16 = a.WU16;
    p.tn_fast = rec_tn_fast(Kp, a.N);
    {
        const char* e = getenv("MATGCN_REC_HINT");
        p.stream_hint = e ? (atoi(e) & 3) : 3;   // measured: forward launch -1.3 %, reverse launch -3.3 % (profiles/r2k_ab_l2_hints.txt)
    }
    const float* al[] = {a.GX, a.RX, a.PH, a.PZ, a.Z, a.R, a.HC, a.H1, a.Z2, a.R2, a.HC2, a.ZH2, a.RgH, a.RuH};
    for (const float* q : al)
        if (reinterpret_cast<uintptr_t>(q) & 31) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(a.PH16) & 31) || (reinterpret_cast<uintptr_t>(a.PZ16) & 31)) return cudaErrorNotSupported;

    RecMaps maps;
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
        const unsigned int b[4] = {64, 64, 1, 1};
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
    const int most = max(p.prop_tiles_m * p.prop_tiles_n, a.N);
    const int grid = most < sms ? most : sms;
    cudaError_t e = cudaMemsetAsync(a.gbar, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&maps, (void*)&p};
    return rec_launch((const void*)rec_fwd_kernel, grid, args, RF_SMEM_TOTAL, st, 0);
}

}  // namespace matgcn
