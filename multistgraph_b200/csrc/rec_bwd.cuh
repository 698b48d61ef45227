// Persistent reverse-time recurrence of one encoder layer (bf16 mode, rnn_units = 64, at most 64 samples per GPU): the
// backward of MA.py:120-128, 142-150, 200-211 for all T steps as ONE cooperative launch - the counterpart of rec_fwd.cuh.
// Per step (t = T-1 .. 0), four phases separated by grid barriers:
//
//   A  per node   first half: head of the reverse step for each of the CTA's nodes: dy = dY_t + carry, chain rule of the sigma-mix
//                 and of the residual GRU cell (its two small products dzh2 = da3 Ru_h and dh1 += [daz2 | dar2] Rg_h as TF32
//                 tcgen05 MMAs on operand tiles the epilogue warps write), main-cell gate algebra -> DR[t], DG[t][:, H:3H], DHD.
//                 second half (after a CTA-local barrier): the per-node products DPT[k] = gu Wu[n,k]^T as an ordinary pipelined
//                 GEMM - gu rows back through TMA from DG16[t], the K weight blocks side by side as one [64 x 64K] operand
//   B  dense      DZ = sum_k M_k^T DPT[k]                 (128 x 128 tiles, K = (K-1) N: TMA-fed tcgen05, plain fp32 store)
//   C  per node   first half: dzh = DZ + DPT[0]; DHD += dzh z; gz = dzh h z (1-z) -> DG[t][:, 0:H]; second half:
//                 DPT[k] = [gz | gr] Wg[n,k]^T (same pipelined form); DHD2 = DHD + DPT[0]
//   D  dense      DC = sum_k M_k^T DPT[k]                 (the carry of step t-1 is DC + DHD2, formed by its phase A)
//
// The elementwise parts of the two dense phases of the per-phase path (EpiB4 / EpiB6) are moved into the per-node phases
// that follow them: there every access is a row of a node's [B, H] block (coalesced), the dense phases keep a plain store,
// and DHD / DPT[0] / DHD2 are produced and consumed by the same CTA (static node -> CTA assignment).  Data that crosses CTAs:
// the bf16 DPT twins (generic-proxy writes -> TMA reads after the barrier) and DZ / DC (read with ld.global.cg: the
// addresses are reused every step, a stale L1 line must not be hit).
#pragma once
#include "rec_fwd.cuh"

namespace matgcn {

constexpr int RB2_STAGES = 3;
constexpr int RB2_NSTAGES = 2;       // per-node phases: whole-tile weight buffers over the same memory (see the producer)
constexpr int RB2_WBUF = 49152;      // two stages [K x 8 KB weight blocks of one K slab | 8 KB operand columns at +40 KB]
constexpr int RB2_OFF_WUT = RB2_STAGES * RF_STAGE_BYTES;   // Ru_h^T [64 inputs][64 outputs] fp32 K-major: 2 slabs of 64 rows x 128 B
constexpr int RB2_OFF_WGT = RB2_OFF_WUT + 16384;           // Rg_h^T [64 inputs][128 outputs]: 4 slabs
constexpr int RB2_OFF_A1 = RB2_OFF_WGT + 32768;            // da3 [64 rows][64] fp32: 2 slabs
constexpr int RB2_OFF_A2 = RB2_OFF_A1 + 16384;             // [daz2 | dar2] [64 rows][128] fp32: 4 slabs
constexpr int RB2_OFF_BAR = RB2_OFF_A2 + 32768;
constexpr int RB2_SMEM_TOTAL = RB2_OFF_BAR + 256 + 1024;
constexpr int RB2_TMEM_D1 = 320, RB2_TMEM_D2 = 384;        // D3[k] at columns 64 k (k < 5); dense accumulators at 0 and 128

struct RecBwdMaps {
    CUtensorMap MT;   // base matrices as the M-contiguous A operand of the transposed propagation: {N, Kp*N}, box 64 x 64
    CUtensorMap DP;   // DPT16 slots 1.. as its B operand: {B*64, Kp*N}, box 64 x 64
    CUtensorMap WG;   // per-node gate weights, K-major B operand of DPT = [gz | gr] Wg^T: WG16 {128, I, K, N}
    CUtensorMap WU;   // per-node candidate weights: WU16 {64, I, K, N}
    CUtensorMap DG;   // bf16 pre-activation gradients as the A operand of the per-node products: DG16 {192, B, N, T}, box 64 x 64
};

struct RecBwdP {
    int T, N, B, K, Cin, n_adp;
    int prop_tiles_m, prop_tiles_n, prop_kt;
    int tn_fast;         // dense-phase tile order: 1 = column tiles fastest (see rf_tile_decode)
    long long U, dy_tstride;
    const float* dy;
    const float* PH; const float* Z; const float* R; const float* HC; const float* H1; const float* Z2; const float* R2; const float* HC2;
    const float* RgH; const float* RuH; const float* mix;
    float* DG; float* DR; __nv_bfloat16* DG16;
    float* DPT0; __nv_bfloat16* DPT16;
    float* DHD; float* DHD2; float* DZC;
    __nv_bfloat16* DPZA; __nv_bfloat16* DPHA;
    float* DHC; float* dmix;
    unsigned int* gbar;
    long long* dbg;      // optional phase timeline of CTA 0 (tools/rec_timeline.py)
    int prefetch;        // warp 3, during the dense phases: bit 0 pulls the next step's saved activations into L2 (phase B), bit 1 the
                         // hidden-row weight blocks of this CTA's nodes for the per-node products that follow (MATGCN_REC_BWD_PF)
    const __nv_bfloat16* WG16; const __nv_bfloat16* WU16;
    int stream_hint;     // 1: saved activations (read once) and DG / DR (written once) carry the L2 evict-first hint (MATGCN_REC_BWD_HINT)
};

__device__ __forceinline__ float4 rb2_ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void rf_st8f(float* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void rb2_sts_bf16x4(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(rf_pack_bf16(v.x, v.y)), "r"(rf_pack_bf16(v.z, v.w)) : "memory");
}

#define RB2_STAMP(i, s)                                                                                                     \
    do {                                                                                                                    \
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128 && t == (T >> 1) && (i) < 4) p.dbg[T * 16 + ((i) << 4) + (s)] = clock64(); \
    } while (0)

__global__ void __launch_bounds__(TC_THREADS, 1) rec_bwd_kernel(const __grid_constant__ RecBwdMaps maps, const RecBwdP p) {
    constexpr int H = 64;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RB2_OFF_BAR);
    // bars: full[S], empty[S], tmem_full[2], tmem_empty[2], phase_bar, a1_full, a2_full, d3_empty, d1_full, d2_full, d3_full,
    //       nfull[2], nempty[2]: the per-node phases use the same ring memory as whole-tile weight buffers - the K weight blocks of
    //       a node are laid side by side as ONE K-major operand with 64 K rows, so that the K products of a tile are one
    //       [64 x 64K] MMA per k-step instead of K small ones (an MMA of this size costs ~150 cycles whatever its N);
    //       the two uses of the memory are never active at the same time
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RB2_STAGES + 12 + 2 * RB2_NSTAGES);
    volatile uint32_t* phase_cnt = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full0 = bar0, empty0 = bar0 + 8u * RB2_STAGES, tfull0 = bar0 + 8u * (2 * RB2_STAGES);
    const uint32_t tempty0 = tfull0 + 16u, phase_bar = tfull0 + 32u;
    const uint32_t a1_full = phase_bar + 8u, a2_full = phase_bar + 16u, d3_empty = phase_bar + 24u;
    const uint32_t d1_full = phase_bar + 32u, d2_full = phase_bar + 40u, d3_full = phase_bar + 48u;
    const uint32_t nfull0 = phase_bar + 56u, nempty0 = nfull0 + 8u * RB2_NSTAGES, sub_bar = nempty0 + 8u * RB2_NSTAGES;
    const uint32_t wut_s = smem_u32(smem + RB2_OFF_WUT), wgt_s = smem_u32(smem + RB2_OFF_WGT);
    const uint32_t a1_s = smem_u32(smem + RB2_OFF_A1), a2_s = smem_u32(smem + RB2_OFF_A2);

    if (threadIdx.x == 0) {
        for (int s = 0; s < RB2_STAGES; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8u * a, 1);
            mbar_init(tempty0 + 8u * a, TC_EPI_WARPS);
        }
        mbar_init(phase_bar, 1);
        mbar_init(a1_full, TC_EPI_WARPS);
        mbar_init(a2_full, TC_EPI_WARPS);
        mbar_init(d3_empty, TC_EPI_WARPS);
        mbar_init(sub_bar, TC_EPI_WARPS);
        mbar_init(d1_full, 1);
        mbar_init(d2_full, 1);
        mbar_init(d3_full, 1);
        for (int s = 0; s < RB2_NSTAGES; ++s) {
            mbar_init(nfull0 + 8u * s, 1);
            mbar_init(nempty0 + 8u * s, 1);
        }
        *phase_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp >= 4) {
        // residual-cell weights, TRANSPOSED (the reverse products contract over the output index o):
        //   Ru_h [64 o][64 i], Rg_h [128 o][64 i]  ->  K-major operand tiles with row = i, K = o:
        //   element (i, o) -> slab o/32, row i, 16-byte chunk ((o%32)/4) ^ (i%8), word o%4
        const int et = threadIdx.x - 128;
        for (int idx = et; idx < 3 * 64 * 16; idx += TC_EPI_WARPS * 32) {
            const int o = idx >> 4, i4 = (idx & 15) * 4;
            const bool g = o >= 64;
            const int oo = g ? o - 64 : o;
            const float4 v = g ? ld4(p.RgH + oo * 64 + i4) : ld4(p.RuH + oo * 64 + i4);
            uint8_t* base = smem + (g ? RB2_OFF_WGT : RB2_OFF_WUT) + (oo >> 5) * 8192 + (oo & 3) * 4;
            const int ch = (oo & 31) >> 2;
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i4 + u;
                *reinterpret_cast<float*>(base + i * 128 + ((ch ^ (i & 7)) << 4)) = vv[u];
            }
        }
        rf_proxy_fence_smem();
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
            uint32_t phase = 0, nbar = 0, wuse0 = 0, wuse1 = 0, subpar = 0;
            const uint64_t pol = l2_policy_evict_last();
            for (int t = T - 1; t >= 0; --t) {
                for (int ph = 0; ph < 4; ++ph) {
                    const bool first = (t == T - 1) && ph == 0;
                    if (ph == 1 || ph == 3) {
                        // dense phases: base matrices (constant: requested before the grid barrier) and the DPT twins
                        int pre = 0;
                        // the ring memory is shared with the per-node weight buffers: both must have been consumed
                        mbar_wait(nempty0, (wuse0 & 1u) ^ 1u);
                        mbar_wait(nempty0 + 8u, (wuse1 & 1u) ^ 1u);
                        if ((int)blockIdx.x < prop_tiles) {
                            int tm, tn;
                            rf_tile_decode(blockIdx.x, p.prop_tiles_m, p.prop_tiles_n, p.tn_fast, tm, tn);
                            int s2 = stage;
                            uint32_t p2 = phase;
                            for (; pre < p.prop_kt && pre < RB2_STAGES; ++pre) {
                                mbar_wait(empty0 + 8u * s2, p2 ^ 1);
                                const uint32_t fb = full0 + 8u * s2, sa = smem_u32(stage_base + s2 * RF_STAGE_BYTES);
                                mbar_expect_tx(fb, RF_STAGE_BYTES);
                                tma_load_5d_hint(sa, &maps.MT, fb, tm * 128, pre * 64, 0, 0, 0, pol);
                                tma_load_5d_hint(sa + 8192, &maps.MT, fb, tm * 128 + 64, pre * 64, 0, 0, 0, pol);
                                if (++s2 == RB2_STAGES) { s2 = 0; p2 ^= 1; }
                            }
                        }
                        // (this thread skips the barriers that precede the per-node phases: the counter makes sure the parity test below
                        // is not answered by an older completion)
                        while (*phase_cnt < nbar + 1u) __nanosleep(32);
                        mbar_wait(phase_bar, nbar & 1u);
                        ++nbar;
                        asm volatile("fence.proxy.async;" ::: "memory");
                        for (int tile = blockIdx.x; tile < prop_tiles; tile += G) {
                            int tm, tn;
                            rf_tile_decode(tile, p.prop_tiles_m, p.prop_tiles_n, p.tn_fast, tm, tn);
                            for (int kt = 0; kt < p.prop_kt; ++kt) {
                                const uint32_t sa = smem_u32(stage_base + stage * RF_STAGE_BYTES);
                                const uint32_t fb = full0 + 8u * stage;
                                if (pre > 0) {
                                    --pre;
                                } else {
                                    mbar_wait(empty0 + 8u * stage, phase ^ 1);
                                    mbar_expect_tx(fb, RF_STAGE_BYTES);
                                    tma_load_5d_hint(sa, &maps.MT, fb, tm * 128, kt * 64, 0, 0, 0, pol);
                                    tma_load_5d_hint(sa + 8192, &maps.MT, fb, tm * 128 + 64, kt * 64, 0, 0, 0, pol);
                                }
                                tma_load_5d(sa + RF_A_BYTES, &maps.DP, fb, tn * 128, kt * 64, 0, 0, 0);
                                tma_load_5d(sa + RF_A_BYTES + 8192, &maps.DP, fb, tn * 128 + 64, kt * 64, 0, 0, 0);
                                if (++stage == RB2_STAGES) { stage = 0; phase ^= 1; }
                            }
                        }
                    } else {
                        // per-node phases stream only the weights, which never depend on the previous phase: they are requested as soon
                        // as the dense ring (same memory) has been consumed, without waiting for the grid barrier
                        {
                            int s2 = stage;
                            uint32_t p2 = phase;
                            for (int i = 0; i < RB2_STAGES; ++i) {
                                mbar_wait(empty0 + 8u * s2, p2 ^ 1);
                                if (++s2 == RB2_STAGES) { s2 = 0; p2 ^= 1; }
                            }
                        }
                        if (!first) ++nbar;   // (the grid barrier before this phase is not waited for here)
                        // second half of a per-node phase: work items (tile, K slab) - one per tile in phase A (K = 64: gu), two in
                        // phase C (K = 128: gz | gr) - each with its own 48 KB stage: the K weight blocks of that slab side by side
                        // (constant: requested right away for the first two items) and the 64 operand columns of DG16[t] (written by
                        // this CTA's epilogue warps in the first half: requested after the sub-phase barrier)
                        const int halves = ph == 0 ? 1 : 2;
                        const uint32_t tx = (uint32_t)K * 8192u + 8192u;
                        auto issue_w = [&](int n, int h, uint32_t dst, uint32_t fb) {
                            for (int k = 0; k < K; ++k) {
                                if (ph == 0) tma_load_5d_hint(dst + k * 8192, &maps.WU, fb, 0, p.Cin, k, n, 0, pol);
                                else tma_load_5d_hint(dst + k * 8192, &maps.WG, fb, 64 * h, p.Cin, k, n, 0, pol);
                            }
                        };
                        int early = 0;
                        {
                            uint32_t u0 = wuse0, u1 = wuse1;
                            for (int n = blockIdx.x; n < node_tiles && early < 2; n += G) {
                                for (int h = 0; h < halves && early < 2; ++h, ++early) {
                                    const int bb = early & 1;
                                    uint32_t& use = bb ? u1 : u0;
                                    mbar_wait(nempty0 + 8u * bb, (use & 1u) ^ 1u);
                                    mbar_expect_tx(nfull0 + 8u * bb, tx);
                                    issue_w(n, h, smem_u32(stage_base + bb * RB2_WBUF), nfull0 + 8u * bb);
                                    ++use;
                                }
                            }
                        }
                        mbar_wait(sub_bar, subpar);   // DG16[t] rows of this CTA's nodes written (generic proxy, fenced at gpu scope) ...
                        subpar ^= 1u;
                        asm volatile("fence.proxy.async;" ::: "memory");   // ... and ordered before the TMA reads below
                        int item = 0;
                        for (int n = blockIdx.x; n < node_tiles; n += G) {
                            for (int h = 0; h < halves; ++h, ++item) {
                                const int bb = item & 1;
                                uint32_t& use = bb ? wuse1 : wuse0;
                                const uint32_t fb = nfull0 + 8u * bb;
                                const uint32_t dst = smem_u32(stage_base + bb * RB2_WBUF);
                                if (item >= early) {
                                    mbar_wait(nempty0 + 8u * bb, (use & 1u) ^ 1u);
                                    mbar_expect_tx(fb, tx);
                                    issue_w(n, h, dst, fb);
                                }
                                tma_load_5d(dst + 40960, &maps.DG, fb, ph == 0 ? 128 : 64 * h, 0, n, t, 0);
                                ++use;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // D = f32.  dense phases: bf16, A and B both MN-major (bits 15, 16), 128 x 128.  per-node products: bf16, both K-major,
            // 64 x 64.  residual-cell products: tf32, both K-major, 64 x 64.
            const uint32_t id_prop = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t id_node0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 4) << 24);   // (N is set per instruction)
            const uint32_t id_res = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, par12 = 0, par3 = 0, wuse0 = 0, wuse1 = 0;
            for (int t = T - 1; t >= 0; --t) {
                for (int ph = 0; ph < 4; ++ph) {
                    if (ph == 1 || ph == 3) {
                        for (int tile = blockIdx.x; tile < prop_tiles; tile += G) {
                            mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1);
                            tc_fence_after();
                            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 128);
                            for (int kt = 0; kt < p.prop_kt; ++kt) {
                                mbar_wait(full0 + 8u * stage, phase);
                                tc_fence_after();
                                const uint32_t sa = smem_u32(stage_base + stage * RF_STAGE_BYTES), sb = sa + RF_A_BYTES;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_bf16(tmem_d, umma_desc(sa + kk * 2048, 8192, 1024, 2), umma_desc(sb + kk * 2048, 8192, 1024, 2), id_prop,
                                              (kt > 0 || kk > 0) ? 1u : 0u);
                                umma_commit(empty0 + 8u * stage);
                                if (++stage == RB2_STAGES) { stage = 0; phase ^= 1; }
                            }
                            umma_commit(tfull0 + 8u * acc);
                            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                        }
                    } else {
                        if (ph == 0) {
                            for (int n = blockIdx.x; n < node_tiles; n += G) {
                                mbar_wait(a1_full, par12);   // da3 tile written: dzh2 = da3 Ru_h
                                tc_fence_after();
#pragma unroll
                                for (int kk = 0; kk < 8; ++kk)
                                    umma_tf32(tmem_base + RB2_TMEM_D1, umma_desc(a1_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2),
                                              umma_desc(wut_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2), id_res, kk > 0 ? 1u : 0u);
                                umma_commit(d1_full);
                                // dh1 += [daz2 | dar2] Rg_h: the dar2 half (written together with da3) runs while the epilogue warps turn
                                // dzh2 into daz2; only the daz2 half is on the critical path
#pragma unroll
                                for (int kk = 8; kk < 16; ++kk)
                                    umma_tf32(tmem_base + RB2_TMEM_D2, umma_desc(a2_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2),
                                              umma_desc(wgt_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2), id_res, kk > 8 ? 1u : 0u);
                                mbar_wait(a2_full, par12);   // daz2 written
                                tc_fence_after();
#pragma unroll
                                for (int kk = 0; kk < 8; ++kk)
                                    umma_tf32(tmem_base + RB2_TMEM_D2, umma_desc(a2_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2),
                                              umma_desc(wgt_s + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024, 2), id_res, 1u);
                                umma_commit(d2_full);
                                par12 ^= 1;
                            }
                        }
                        // second half of the phase: D3 [64 x 64K] (+)= operand columns (TMA) x [W_0^T | .. | W_{K-1}^T] per work item
                        // (tile, K slab), an ordinary pipelined GEMM - per k-step one MMA of up to 256 columns (+ the rest)
                        const int halves = ph == 0 ? 1 : 2;
                        int item = 0;
                        for (int n = blockIdx.x; n < node_tiles; n += G) {
                            for (int h = 0; h < halves; ++h, ++item) {
                                const int bb = item & 1;
                                uint32_t& use = bb ? wuse1 : wuse0;
                                mbar_wait(nfull0 + 8u * bb, use & 1u);                  // weights and operand columns landed
                                if (h == 0) mbar_wait(d3_empty, par3 ^ 1u);             // the previous tile's products are in registers
                                tc_fence_after();
                                const uint32_t wb = smem_u32(stage_base + bb * RB2_WBUF);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    const uint64_t da = umma_desc(wb + 40960 + kk * 32, 16, 1024, 2);
                                    const uint32_t acc_in = (h > 0 || kk > 0) ? 1u : 0u;
                                    for (int k0 = 0; k0 < K; k0 += 4) {
                                        const uint32_t nn = (uint32_t)min(4, K - k0) * 64u;
                                        umma_bf16(tmem_base + (uint32_t)(64 * k0), da, umma_desc(wb + k0 * 8192 + kk * 32, 16, 1024, 2),
                                                  id_node0 | ((nn >> 3) << 17), acc_in);
                                    }
                                }
                                umma_commit(nempty0 + 8u * bb);
                                ++use;
                            }
                            umma_commit(d3_full);
                            par3 ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ================================ L2 prefetch ================================
        // the saved activations the head of the NEXT reverse step (t - 1) reads on this CTA, requested while the dense phase B
        // of step t runs
        const int I = p.Cin + H;
        auto pf_weights = [&](const __nv_bfloat16* W, int ow) {   // per node and support one contiguous block of 64 rows x ow bf16
            const int units = ((node_tiles - (int)blockIdx.x + G - 1) / G) * K;
            for (int u = lane; u < units; u += 32) {
                const int n = blockIdx.x + (u / K) * G, k = u - (u / K) * K;
                const __nv_bfloat16* src = W + (((long long)n * K + k) * I + p.Cin) * ow;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(H * ow * 2) : "memory");
            }
        };
        uint32_t nb = 1;   // barriers passed when phase B of the first step begins
        for (int t = T - 1; t >= 0 && p.prefetch; --t, nb += 4) {
            while (*phase_cnt < nb) __nanosleep(256);
            if (p.prefetch & 2) pf_weights(p.WG16, 2 * H);   // phase C: [gz | gr] Wg^T
            if ((p.prefetch & 1) && t >= 1) {
                const long long tU = (long long)(t - 1) * p.U;
                const float* arr[9] = {p.H1 + tU, p.R2 + tU, p.HC2 + tU, p.Z2 + tU, p.R + tU, p.HC + tU, p.Z + tU,
                                       p.PH + (long long)(t - 1) * K * p.U, p.dy + (long long)(t - 1) * p.dy_tstride};
                for (int n = blockIdx.x; n < node_tiles; n += G) {
                    const long long base = (long long)n * p.B * H;
                    for (int a = 0; a < 9; ++a)
                        for (int idx = lane; idx < p.B * 2; idx += 32) pf_l2(arr[a] + base + idx * 32);
                }
            }
            if ((p.prefetch & 2) && t >= 1) {
                while (*phase_cnt < nb + 2) __nanosleep(256);   // phase D of this step: the next step's phase A multiplies with Wu^T
                pf_weights(p.WU16, H);
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        const int q = warp & 3, half = (warp - 4) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0, nbar = 0, par12 = 0, par3 = 0;
        const int ldc = p.B * H;
        // per-node tiles (64 rows): the thread's coordinates after rf_quad4 (see rec_fwd.cuh)
        const int b0 = q * 16 + (lane >> 2);
        const bool odd = lane & 1;
        const int pc = ((lane & 1) << 1) | ((lane >> 1) & 1);
        const bool ok[2] = {b0 < p.B, b0 + 8 < p.B};
        const int ch = half * 32 + 4 * pc;
        const long long rd[2] = {ok[0] ? 0 : -(long long)b0, ok[1] ? 8 : -(long long)b0};
        const uint32_t s_off = (uint32_t)(half * 8192 + b0 * 128);   // tf32 tiles: slab `half` (+ 2 for the dar2 columns)
        const uint32_t s_x = (uint32_t)(b0 & 7);
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        const uint64_t pol = rf_policy_stream(p.stream_hint);
        for (int t = T - 1; t >= 0; --t) {
            for (int ph = 0; ph < 4; ++ph) {
                long long tl = t;
                asm volatile("" : "+l"(tl));   // (nothing below is hoisted above the phase: see rec_fwd.cuh)
                const long long tU = tl * p.U;
                // this phase reads what OTHER CTAs wrote in the previous one (DZ / DC): every epilogue thread waits for the grid
                // barrier, not only the TMA producer
                if (nbar > 0) mbar_wait(phase_bar, (nbar - 1u) & 1u);
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[((T - 1 - t) * 4 + ph) * 4 + 0] = clock64();
                if (ph == 1 || ph == 3) {
                    // ---- dense phases: plain fp32 store of the 128 x 128 tile (lane = node row) ----
                    for (int tile = blockIdx.x; tile < prop_tiles; tile += G) {
                        int tm, tn;
                            rf_tile_decode(tile, p.prop_tiles_m, p.prop_tiles_n, p.tn_fast, tm, tn);
                        const long long row = (long long)tm * 128 + q * 32 + lane;
                        mbar_wait(tfull0 + 8u * acc, acc_phase);
                        tc_fence_after();
                        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 128) + tlane;
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            const int c = half + 2 * cc;
                            const int col = tn * 128 + c * 32;
                            uint32_t r[32];
                            rf_tmem_ld32(tmem_acc + (uint32_t)(c * 32), r);
                            rf_tmem_wait_ld();
                            if (row < p.N && col < ldc) {
                                float* o = p.DZC + row * ldc + col;
#pragma unroll
                                for (int j = 0; j < 4; ++j) rf_st8f(o + 8 * j, r + 8 * j);
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                } else if (ph == 0) {
                    // ---- phase A: head of the reverse step, then DPT[k] = gu Wu[n,k]^T ----
                    const float* dYt = p.dy + tl * p.dy_tstride;
                    const float* Hp = p.PH + tl * K * p.U;
                    const float gmix = __ldg(p.mix + t);
                    float dmix_part = 0.f;
                    float4 dyv[4], cv[4], dv[4], h1[4], r2v[4], hc2[4];
                    auto load_s0 = [&](int n) {
                        const long long o0 = ((long long)n * p.B + b0) * H + ch;
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const long long o = o0 + rd[w] * H + 16 * m2;
                                const int e = 2 * w + m2;
                                dyv[e] = ld4h(dYt + o, pol);
                                cv[e] = rb2_ldcg4(p.DZC + o);
                                dv[e] = ld4(p.DHD2 + o);
                                h1[e] = ld4h(p.H1 + tU + o, pol);
                                r2v[e] = ld4h(p.R2 + tU + o, pol);
                                hc2[e] = ld4h(p.HC2 + tU + o, pol);
                            }
                        }
                    };
                    load_s0(min((int)blockIdx.x, node_tiles - 1));   // (unconditional: keeps the arrays in registers)
                    for (int n = blockIdx.x; n < node_tiles; n += G) {
                        const long long g0 = (long long)n * p.B + b0;
                        const long long o0 = g0 * H + ch, x0 = g0 * 3 * H + ch;
                        float4 z2[4], dh1[4];
                        RB2_STAMP(n / G, 0);
                        // S0: dy, sigma-mix and the elementwise half of the residual cell's chain rule
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const int e = 2 * w + m2;
                                const float4 dy = dyv[e] + cv[e] + dv[e];
                                const float4 res = r2v[e] * h1[e] + one_minus(r2v[e]) * hc2[e];
                                const float4 pr = dy * (h1[e] - res);
                                const float4 dres = (1.f - gmix) * dy;
                                const float4 da3 = dres * one_minus(r2v[e]) * one_minus(hc2[e] * hc2[e]);
                                const float4 dar = dres * (h1[e] - hc2[e]) * r2v[e] * one_minus(r2v[e]);
                                dh1[e] = gmix * dy + dres * r2v[e];
                                if (ok[w]) {
                                    dmix_part += (pr.x + pr.y) + (pr.z + pr.w);
                                    const long long x = x0 + w * 8 * 3 * H + 16 * m2;
                                    st4h(p.DR + 3 * tU + x + 2 * H, da3, pol);
                                    st4h(p.DR + 3 * tU + x + H, dar, pol);
                                }
                                const uint32_t so = s_off + (uint32_t)(w * 1024) + ((((uint32_t)(4 * m2 + pc)) ^ s_x) << 4);
                                rf_sts4(a1_s + so, da3);
                                rf_sts4(a2_s + 16384 + so, dar);
                            }
                        }
                        rf_proxy_fence_smem();
                        __syncwarp();
                        RB2_STAMP(n / G, 1);
                        if (lane == 0) mbar_arrive(a1_full);
                        // (inputs of S1 and S2, requested ahead of their use)
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) z2[2 * w + m2] = ld4h(p.Z2 + tU + o0 + rd[w] * H + 16 * m2, pol);
                        }
                        float4 rr[4], hc[4], hp[4];
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const long long o = o0 + rd[w] * H + 16 * m2;
                                rr[2 * w + m2] = ld4h(p.R + tU + o, pol);
                                hc[2 * w + m2] = ld4h(p.HC + tU + o, pol);
                                hp[2 * w + m2] = ld4(Hp + o);
                            }
                        }
                        // S1: dzh2 -> daz2, dh1
                        float a[16];
                        float4 fa[4];
                        mbar_wait(d1_full, par12);
                        tc_fence_after();
                        RB2_STAMP(n / G, 2);
                        rf_tmem_ld16x4(tmem_base + RB2_TMEM_D1 + tlane + (uint32_t)(half * 32), a);
                        rf_tmem_wait_ld();
                        rf_quad4(a, fa, odd);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const int e = 2 * w + m2;
                                const float4 daz = fa[e] * h1[e] * z2[e] * one_minus(z2[e]);
                                dh1[e] = dh1[e] + fa[e] * z2[e];
                                if (ok[w]) st4h(p.DR + 3 * tU + x0 + w * 8 * 3 * H + 16 * m2, daz, pol);
                                rf_sts4(a2_s + s_off + (uint32_t)(w * 1024) + ((((uint32_t)(4 * m2 + pc)) ^ s_x) << 4), daz);
                            }
                        }
                        tc_fence_before();
                        rf_proxy_fence_smem();
                        __syncwarp();
                        RB2_STAMP(n / G, 3);
                        if (lane == 0) mbar_arrive(a2_full);
                        // S2: dh1 complete -> DHD, candidate / r-gate pre-activation gradients, bf16 operand tile of DPT = gu Wu^T
                        mbar_wait(d2_full, par12);
                        tc_fence_after();
                        RB2_STAMP(n / G, 4);
                        rf_tmem_ld16x4(tmem_base + RB2_TMEM_D2 + tlane + (uint32_t)(half * 32), a);
                        rf_tmem_wait_ld();
                        rf_quad4(a, fa, odd);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const int e = 2 * w + m2;
                                const float4 d = dh1[e] + fa[e];
                                const float4 gu = d * one_minus(rr[e]) * one_minus(hc[e] * hc[e]);
                                const float4 gr = d * (hp[e] - hc[e]) * rr[e] * one_minus(rr[e]);
                                if (ok[w]) {
                                    const long long o = o0 + w * 8 * H + 16 * m2, x = x0 + w * 8 * 3 * H + 16 * m2;
                                    st4(p.DHD + o, d * rr[e]);
                                    if (p.DG) {   // (null: every consumer of DG reads the bf16 twin)
                                        st4h(p.DG + 3 * tU + x + 2 * H, gu, pol);
                                        st4h(p.DG + 3 * tU + x + H, gr, pol);
                                    }
                                    st4_bf16(p.DG16 + 3 * tU + x + 2 * H, gu);
                                    st4_bf16(p.DG16 + 3 * tU + x + H, gr);
                                }
                            }
                        }
                        tc_fence_before();
                        RB2_STAMP(n / G, 5);
                        load_s0(min(n + G, node_tiles - 1));   // (dyv .. hc2 are dead: the next tile's S0 inputs go in flight)
                        par12 ^= 1;
                    }
                    // the bf16 gu rows of this CTA's nodes are in DG16[t]: hand them to the TMA producer (second half of the phase)
                    __threadfence();   // the rows must have reached L2, where TMA reads them, before the producer is told
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sub_bar);
                    // DPT[k] = gu Wu[n,k]^T for all k: products -> registers (accumulator released at once) -> DPT[0] (fp32),
                    // DPT[k >= 1] (bf16 twins only), adaptive slices per step
                    for (int n = blockIdx.x; n < node_tiles; n += G) {
                        const long long o0 = ((long long)n * p.B + b0) * H + ch;
                        float fr[5][16];
                        mbar_wait(d3_full, par3);
                        tc_fence_after();
                        RB2_STAMP(n / G, 6);
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            rf_tmem_ld16x4(tmem_base + (uint32_t)(64 * k) + tlane + (uint32_t)(half * 32), fr[k]);   // (k >= K: unused columns)
                        rf_tmem_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(d3_empty);
                        par3 ^= 1;
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (k < K) {
                                float4 fa[4];
                                rf_quad4(fr[k], fa, odd);
#pragma unroll
                                for (int w = 0; w < 2; ++w) {
#pragma unroll
                                    for (int m2 = 0; m2 < 2; ++m2) {
                                        if (ok[w]) {
                                            const long long o = o0 + w * 8 * H + 16 * m2;
                                            if (k == 0) {
                                                st4(p.DPT0 + o, fa[2 * w + m2]);
                                            } else {
                                                st4_bf16(p.DPT16 + (long long)k * p.U + o, fa[2 * w + m2]);
                                                if (k <= p.n_adp) st4_bf16(p.DPZA + (tl * p.n_adp + (k - 1)) * p.U + o, fa[2 * w + m2]);
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        RB2_STAMP(n / G, 7);
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) dmix_part += __shfl_xor_sync(0xffffffffu, dmix_part, off);
                    if (lane == 0 && dmix_part != 0.f) atomicAdd(p.dmix + t, dmix_part);
                } else {
                    // ---- phase C: z-gate algebra on dzh = DZ + DPT[0], then DPT[k] = [gz | gr] Wg[n,k]^T, DHD2 = DHD + dzh z + DPT[0] ----
                    const float* Hp = p.PH + tl * K * p.U;
                    float4 dz[4], d0[4], zz[4], hp[4], dd[4];
                    auto load_c = [&](int n) {
                        const long long o0 = ((long long)n * p.B + b0) * H + ch;
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const long long o = o0 + rd[w] * H + 16 * m2;
                                const int e = 2 * w + m2;
                                dz[e] = rb2_ldcg4(p.DZC + o);
                                d0[e] = ld4(p.DPT0 + o);
                                zz[e] = ld4h(p.Z + tU + o, pol);
                                hp[e] = ld4(Hp + o);
                                dd[e] = ld4(p.DHD + o);
                            }
                        }
                    };
                    load_c(min((int)blockIdx.x, node_tiles - 1));   // (unconditional: keeps the arrays in registers)
                    for (int n = blockIdx.x; n < node_tiles; n += G) {
                        const long long g0 = (long long)n * p.B + b0;
                        const long long o0 = g0 * H + ch, x0 = g0 * 3 * H + ch;
                        float4 dk[4];
                        RB2_STAMP(n / G, 8);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) {
                                const int e = 2 * w + m2;
                                const float4 dzh = dz[e] + d0[e];
                                dk[e] = dd[e] + dzh * zz[e];
                                const float4 gz = dzh * hp[e] * zz[e] * one_minus(zz[e]);
                                if (ok[w]) {
                                    const long long x = x0 + w * 8 * 3 * H + 16 * m2;
                                    if (p.DG) st4h(p.DG + 3 * tU + x, gz, pol);
                                    st4_bf16(p.DG16 + 3 * tU + x, gz);
                                }
                                if (ok[w]) st4(p.DHD + o0 + w * 8 * H + 16 * m2, dk[e]);   // DHD += dzh z (read back after the products)
                            }
                        }
                        RB2_STAMP(n / G, 9);
                        load_c(min(n + G, node_tiles - 1));   // (the next tile's inputs go in flight)
                    }
                    __threadfence();   // the rows must have reached L2, where TMA reads them, before the producer is told
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sub_bar);
                    // DPT[k] = [gz | gr] Wg[n,k]^T: DHD2 = DHD + DPT[0]; DPT[k >= 1] as bf16 twins, adaptive slices per step
                    for (int n = blockIdx.x; n < node_tiles; n += G) {
                        const long long o0 = ((long long)n * p.B + b0) * H + ch;
                        float4 dk2[4];
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
#pragma unroll
                            for (int m2 = 0; m2 < 2; ++m2) dk2[2 * w + m2] = ld4(p.DHD + o0 + rd[w] * H + 16 * m2);
                        }
                        float fr[5][16];
                        mbar_wait(d3_full, par3);
                        tc_fence_after();
                        RB2_STAMP(n / G, 10);
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            rf_tmem_ld16x4(tmem_base + (uint32_t)(64 * k) + tlane + (uint32_t)(half * 32), fr[k]);   // (k >= K: unused columns)
                        rf_tmem_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(d3_empty);
                        par3 ^= 1;
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (k < K) {
                                float4 fa[4];
                                rf_quad4(fr[k], fa, odd);
#pragma unroll
                                for (int w = 0; w < 2; ++w) {
#pragma unroll
                                    for (int m2 = 0; m2 < 2; ++m2) {
                                        if (ok[w]) {
                                            const long long o = o0 + w * 8 * H + 16 * m2;
                                            if (k == 0) {
                                                st4(p.DHD2 + o, dk2[2 * w + m2] + fa[2 * w + m2]);
                                            } else {
                                                st4_bf16(p.DPT16 + (long long)k * p.U + o, fa[2 * w + m2]);
                                                if (k <= p.n_adp) st4_bf16(p.DPHA + (tl * p.n_adp + (k - 1)) * p.U + o, fa[2 * w + m2]);
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        RB2_STAMP(n / G, 11);
                    }
                }
                // ---- end of phase: publish this CTA's writes, wait for every CTA ----
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[((T - 1 - t) * 4 + ph) * 4 + 1] = clock64();
                asm volatile("bar.sync 6, 256;" ::: "memory");
                if (threadIdx.x == 128) {
                    if (p.dbg && blockIdx.x == 0) p.dbg[((T - 1 - t) * 4 + ph) * 4 + 2] = clock64();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.gbar) : "memory");
                    const unsigned int target = (nbar + 1u) * (unsigned int)G;
                    long long t0 = 0;
                    for (uint32_t it = 0; rf_ld_acquire(p.gbar) < target; ++it) {
                        if (it == 1024) t0 = clock64();
                        if (it > 1024 && (it & 255) == 0 && clock64() - t0 > 4000000000LL) __trap();
                    }
                    if (p.dbg && blockIdx.x == 0) p.dbg[((T - 1 - t) * 4 + ph) * 4 + 3] = clock64();
                    *phase_cnt = nbar + 1u;
                    mbar_arrive(phase_bar);
                }
                ++nbar;
            }
        }
        // gradient w.r.t. the initial state: the carry the (non-existent) step -1 would read
        asm volatile("bar.sync 6, 256;" ::: "memory");   // (thread 128 has passed the last grid barrier)
        for (int n = blockIdx.x; n < node_tiles; n += G) {
            const long long o0 = ((long long)n * p.B + b0) * H + ch;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
#pragma unroll
                for (int m2 = 0; m2 < 2; ++m2) {
                    if (ok[w]) {
                        const long long o = o0 + w * 8 * H + 16 * m2;
                        st4(p.DHC + o, rb2_ldcg4(p.DZC + o) + ld4(p.DHD2 + o));
                    }
                }
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

cudaError_t launch_rec_bwd(const RecBwdArgs& a, cudaStream_t st) {
    constexpr int H = 64;
    const int Kp = a.K - 1, I = a.Cin + H;
    if (a.B < 8 || a.B > 64 || (a.ldm & 7) || a.K < 2 || a.K > 5 || a.N < 1 || a.T < 1 || a.n_adp < 0 || a.n_adp > Kp) return cudaErrorNotSupported;
    const unsigned long long U = (unsigned long long)a.N * a.B * H;
    RecBwdP p;
    memset(&p, 0, sizeof(p));
    p.T = a.T; p.N = a.N; p.B = a.B; p.K = a.K; p.Cin = a.Cin; p.n_adp = a.n_adp;
    p.prop_tiles_m = (a.N + 127) / 128;
    p.prop_tiles_n = (a.B * H + 127) / 128;
    p.prop_kt = (Kp * a.N + 63) / 64;
    p.U = (long long)U; p.dy_tstride = a.dy_tstride;
    p.dy = a.dy;
    p.PH = a.PH; p.Z = a.Z; p.R = a.R; p.HC = a.HC; p.H1 = a.H1; p.Z2 = a.Z2; p.R2 = a.R2; p.HC2 = a.HC2;
    p.RgH = a.RgH; p.RuH = a.RuH; p.mix = a.mix;
    p.DG = a.DG; p.DR = a.DR; p.DG16 = a.DG16;
    p.DPT0 = a.DPT0; p.DPT16 = a.DPT16;
    p.DHD = a.DHD; p.DHD2 = a.DHD2; p.DZC = a.DZC;
    p.DPZA = a.DPZA; p.DPHA = a.DPHA;
    p.DHC = a.DHC; p.dmix = a.dmix;
    p.gbar = a.gbar;
    p.WG16 = a.WG16; p.WU16 = a.WU16;
    p.tn_fast = rec_tn_fast(a.K - 1, a.N);
    {
        const char* e = getenv("MATGCN_REC_BWD_HINT");
        p.stream_hint = e ? (atoi(e) & 1) : 1;   // measured: forward launch -1.3 %, reverse launch -3.3 % (profiles/r2k_ab_l2_hints.txt)
    }
    p.dbg = tc_debug_buffer();
    {
        const char* e = getenv("MATGCN_REC_BWD_PF");
        p.prefetch = e ? atoi(e) & 3 : 0;   // bit 0 off by default: the fill traffic costs the dense phase more than the head gains
    }
    const void* al[] = {a.dy, a.PH, a.Z, a.R, a.HC, a.H1, a.Z2, a.R2, a.HC2, a.RgH, a.RuH, a.DR, a.DG16, a.DPT0, a.DPT16, a.DHD,
                        a.DHD2, a.DZC, a.DHC};
    for (const void* q : al)
        if (!q || (reinterpret_cast<uintptr_t>(q) & 31)) return cudaErrorNotSupported;
    if (reinterpret_cast<uintptr_t>(a.DG) & 31) return cudaErrorNotSupported;   // (DG may be null: no fp32 copy of the pre-activation gradients)
    if ((a.dy_tstride & 7) || (a.n_adp > 0 && (!a.DPZA || !a.DPHA || (reinterpret_cast<uintptr_t>(a.DPZA) & 15) || (reinterpret_cast<uintptr_t>(a.DPHA) & 15))))
        return cudaErrorNotSupported;

    RecBwdMaps maps;
    {
        const unsigned long long d[2] = {(unsigned long long)a.N, (unsigned long long)Kp * a.N}, s[1] = {(unsigned long long)a.ldm * 2};
        const unsigned int b[2] = {64, 64};
        if (!rf_make_map(&maps.MT, a.M16, 2, d, s, b)) return cudaErrorNotSupported;
    }
    {
        const unsigned long long d[2] = {(unsigned long long)a.B * H, (unsigned long long)Kp * a.N}, s[1] = {(unsigned long long)a.B * H * 2};
        const unsigned int b[2] = {64, 64};
        if (!rf_make_map(&maps.DP, a.DPT16 + U, 2, d, s, b)) return cudaErrorNotSupported;
    }
    {
        const unsigned int b[4] = {64, 64, 1, 1};
        const unsigned long long dg[4] = {128, (unsigned long long)I, (unsigned long long)a.K, (unsigned long long)a.N};
        const unsigned long long sg[3] = {256, (unsigned long long)I * 256, (unsigned long long)a.K * I * 256};
        const unsigned long long du[4] = {64, (unsigned long long)I, (unsigned long long)a.K, (unsigned long long)a.N};
        const unsigned long long su[3] = {128, (unsigned long long)I * 128, (unsigned long long)a.K * I * 128};
        if (!rf_make_map(&maps.WG, a.WG16, 4, dg, sg, b) || !rf_make_map(&maps.WU, a.WU16, 4, du, su, b)) return cudaErrorNotSupported;
        const unsigned long long dd[4] = {192, (unsigned long long)a.B, (unsigned long long)a.N, (unsigned long long)a.T};
        const unsigned long long sd[3] = {384, (unsigned long long)a.B * 384, U * 3 * 2};
        if (!rf_make_map(&maps.DG, a.DG16, 4, dd, sd, b)) return cudaErrorNotSupported;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64) return cudaErrorNotSupported;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(rec_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB2_SMEM_TOTAL);
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
    return rec_launch((const void*)rec_bwd_kernel, grid, args, RB2_SMEM_TOTAL, st, 1);
}

}  // namespace matgcn
