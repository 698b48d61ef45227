// EXPERIMENTAL (off by default; matgcn_set_persistent(1) or MATGCN_MULTI=1 enables it).  Results are identical to
// the one-launch-per-phase path and it is covered by tests, but as measured in round 1 it is slower (a step of the
// Baltimore shape: 40 ms vs 28 ms): the union of nine epilogue flavours does not fit the 168 registers/thread of a
// 384-thread block without spilling in the epilogue loops, and a phase boundary (CTA barrier + grid barrier) costs
// 5-8 us - no less than a PDL kernel boundary.  Kept as the starting point for round 2.
//
// Persistent multi-phase tensor-core kernel: a whole recurrence (all time steps of an encoder layer, forward or
// reverse) runs as ONE cooperative launch.  Each phase is one contraction of the step (same TMA / tcgen05 / TMEM
// pipeline and the same epilogue functors as gemm_tc.cuh) or the elementwise head of the reverse step; phases are
// separated by a grid-wide barrier instead of a kernel boundary, so the hidden state and every intermediate of
// the step stay in L2 between phases and the ~10 us launch + prologue + drain cost of a kernel is paid once per
// layer instead of 13 times per time step.
//
// Memory model across a phase boundary (data written with st.global by epilogue warps of any CTA, read in the next
// phase by TMA (async proxy, from L2) and by epilogue loads (through L1)):
//   writers : st.global ... __threadfence() -> bar.sync among the epilogue warps -> one thread: atomicAdd(gbar)
//   readers : that thread spins with ld.acquire.gpu on gbar -> __threadfence() -> arrives on a CTA mbarrier;
//             the TMA producer waits on it and issues fence.proxy.async; every epilogue thread waits on it and
//             executes __threadfence() (gpu-scope fence: also drops stale L1 lines of buffers reused every step).
#pragma once
#include <vector>

#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace matgcn {

enum EpiKind : int { EK_STORE = 0, EK_ATOMIC, EK_GATE, EK_CAND, EK_RESCAND, EK_B1, EK_B2, EK_B4, EK_B6, EK_PLAIN };
enum PhaseKind : int { PK_GEMM = 0, PK_HEAD = 1 };

// elementwise head of the reverse step (B0 of DESIGN.md section 3)
struct HeadArgs {
    const float* dY; const float* carry; const float* H1; const float* R2; const float* HC2; const float* mix_t;
    long long n;
    int H;
    float* DH1; float* DRES; float* DR; float* dmix_t;
};

template <class E> struct EpiKindOf;
template <> struct EpiKindOf<EpiStore> { static constexpr int v = EK_STORE; };
template <> struct EpiKindOf<EpiPlain> { static constexpr int v = EK_PLAIN; };
template <> struct EpiKindOf<EpiAtomic> { static constexpr int v = EK_ATOMIC; };
template <> struct EpiKindOf<EpiGate> { static constexpr int v = EK_GATE; };
template <> struct EpiKindOf<EpiCand> { static constexpr int v = EK_CAND; };
template <> struct EpiKindOf<EpiResCand> { static constexpr int v = EK_RESCAND; };
template <> struct EpiKindOf<EpiB1> { static constexpr int v = EK_B1; };
template <> struct EpiKindOf<EpiB2> { static constexpr int v = EK_B2; };
template <> struct EpiKindOf<EpiB4> { static constexpr int v = EK_B4; };
template <> struct EpiKindOf<EpiB6> { static constexpr int v = EK_B6; };

constexpr int MP_BLOB = 160;
constexpr int MP_SLOTS = 8;  // distinct contractions per time step
// Tensor maps live in kernel-parameter space (descriptors fetched from global memory throttle TMA); one pair per
// contraction of the step, with the time step as an extra tensor dimension.
struct MpMaps {
    CUtensorMap a[MP_SLOTS], b[MP_SLOTS];
};
struct alignas(128) MPhase {
    TcP p;
    int kind, a_kc, b_kc, bn, epi_kind, slot;
    const float* copy_src;  // optional side copy performed by the epilogue warps at the start of the phase
    float* copy_dst;
    long long copy_n;       // floats, multiple of 4
    alignas(16) unsigned char blob[MP_BLOB];  // the epilogue functor (or HeadArgs)
};

constexpr int MP_BN = 128;  // widest tile; TMEM = 2 x 128 columns
constexpr int MP_STAGES = 5;
struct MpSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;
    static constexpr int B_BYTES = MP_BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_BYTES = TC_EPI_WARPS * 32 * TC_EPI_LD * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = MP_STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The functor is copied field by field into registers (a byte-wise copy through its address would pin it to
// local memory and turn every field access of the epilogue into a local load).
template <class Epi>
__device__ __forceinline__ Epi load_functor(const MPhase* ph) {
    return *reinterpret_cast<const Epi*>(ph->blob);
}

struct MpEpiState {
    int acc;
    uint32_t acc_phase;
};

// tile index -> coordinates (shared by all roles)
__device__ __forceinline__ void mp_decode(const TcP& p, int bn, int tile, int& z1, int& z2, int& m0, int& n0, int& kt0, int& kt1) {
    const int tiles_mn = p.tiles_m * p.tiles_n;
    const int kt_per_kb = (p.K + TC_BK - 1) / TC_BK;
    const int kt_total = p.KB * kt_per_kb;
    const int kt_per_split = (kt_total + p.splits - 1) / p.splits;
    const int rest = (int)fdiv((uint32_t)tile, p.d_mn);
    const int mn = tile - rest * tiles_mn;
    const int z = (int)fdiv((uint32_t)rest, p.d_splits);
    const int split = rest - z * p.splits;
    z1 = (int)fdiv((uint32_t)z, p.d_z2);
    z2 = z - z1 * p.Z2;
    const int tm = (int)fdiv((uint32_t)mn, p.d_tn);
    m0 = tm * TC_BM;
    n0 = (mn - tm * p.tiles_n) * bn;
    kt0 = split * kt_per_split;
    kt1 = min(kt_total, kt0 + kt_per_split);
}

// (not inlined: each epilogue flavour gets its own register allocation instead of the union of all nine)
template <class Epi>
__device__ __noinline__ void mp_epilogue_phase(const MPhase* ph, int bn, uint32_t tmem_base, uint32_t tfull0,
                                               uint32_t tempty0, MpEpiState& st_io, int q, int half, float* buf, int lane) {
    // by-value copies: everything the hot loops read lives in registers, not behind a pointer that every
    // global store would force the compiler to re-read
    const Epi epi = load_functor<Epi>(ph);
    const TcP p = ph->p;
    MpEpiState st = st_io;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int z1, z2, m0, n0, kt0, kt1;
        mp_decode(p, bn, tile, z1, z2, m0, n0, kt0, kt1);
        if (kt0 >= kt1) continue;
        mbar_wait(tfull0 + 8u * st.acc, st.acc_phase);
        tc_fence_after();
        tc_epilogue_tile(epi, p, tmem_base + (uint32_t)(st.acc * MP_BN), bn, q, half, buf, lane, z1, z2, m0, n0);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8u * st.acc);
        if (++st.acc == 2) { st.acc = 0; st.acc_phase ^= 1; }
    }
    st_io = st;
}

__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_multi_kernel(const __grid_constant__ MpMaps maps, const MPhase* __restrict__ phases, const int nph,
                                                                           unsigned int* gbar, long long* dbg) {
    using S = MpSmem;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* stage_base = smem;
    float* epi_buf = reinterpret_cast<float*>(smem + MP_STAGES * S::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MP_STAGES * S::STAGE_BYTES + S::EPI_BYTES);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], phase_bar; then the TMEM base slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MP_STAGES + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full0 = bar0, empty0 = bar0 + 8u * MP_STAGES, tfull0 = bar0 + 8u * (2 * MP_STAGES);
    const uint32_t tempty0 = tfull0 + 16u, phase_bar = tfull0 + 32u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MP_STAGES; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8u * a, 1);
            mbar_init(tempty0 + 8u * a, TC_EPI_WARPS);
        }
        mbar_init(phase_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * MP_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int ph = 0; ph < nph; ++ph) {
                const MPhase* P = phases + ph;
                if (ph > 0) {
                    mbar_wait(phase_bar, (uint32_t)((ph - 1) & 1));  // previous phase complete on every CTA
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                if (dbg && blockIdx.x == 0) dbg[1024 + ph * 16 + 0] = clock64();
                if (P->kind != PK_GEMM) continue;
                const TcP p = P->p;
                const int bn = P->bn, a_kc = P->a_kc, b_kc = P->b_kc;
                const int kt_per_kb = (p.K + TC_BK - 1) / TC_BK;
                const uint32_t tx_bytes = (uint32_t)(S::A_BYTES + bn * TC_BK * 4);
                const CUtensorMap* tmA = &maps.a[P->slot];
                const CUtensorMap* tmB = &maps.b[P->slot];
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    int z1, z2, m0, n0, kt0, kt1;
                    mp_decode(p, bn, tile, z1, z2, m0, n0, kt0, kt1);
                    for (int kt = kt0; kt < kt1; ++kt) {
                        const int kb = kt / kt_per_kb;
                        const int k0 = (kt - kb * kt_per_kb) * TC_BK;
                        mbar_wait(empty0 + 8u * stage, phase ^ 1);
                        const uint32_t sa = smem_u32(stage_base + stage * S::STAGE_BYTES);
                        const uint32_t sb = sa + S::A_BYTES;
                        const uint32_t fb = full0 + 8u * stage;
                        mbar_expect_tx(fb, tx_bytes);
                        int ca[3] = {kb * p.cAk, z2 * p.cA2, z1 * p.cA1};
                        int cb[3] = {kb * p.cBk, z2 * p.cB2, z1 * p.cB1};
                        if (p.tA_dim) ca[p.tA_dim - 2] = p.t;
                        if (p.tB_dim) cb[p.tB_dim - 2] = p.t;
                        if (a_kc) {
                            tma_load_5d(sa, tmA, fb, k0, m0, ca[0], ca[1], ca[2]);
                        } else {
                            for (int j = 0; j < TC_BM / 32; ++j) tma_load_5d(sa + j * 4096, tmA, fb, m0 + 32 * j, k0, ca[0], ca[1], ca[2]);
                        }
                        if (b_kc) {
                            tma_load_5d(sb, tmB, fb, k0, n0, cb[0], cb[1], cb[2]);
                        } else {
                            for (int j = 0; j < bn / 32; ++j) tma_load_5d(sb + j * 4096, tmB, fb, n0 + 32 * j, k0, cb[0], cb[1], cb[2]);
                        }
                        if (++stage == MP_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (dbg && blockIdx.x == 0 && tile < 3 * (int)gridDim.x) dbg[1024 + ph * 16 + 1 + tile / gridDim.x] = clock64();
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int ph = 0; ph < nph; ++ph) {
                const MPhase* P = phases + ph;
                if (P->kind != PK_GEMM) continue;
                const TcP p = P->p;
                const int bn = P->bn, a_kc = P->a_kc, b_kc = P->b_kc;
                const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_kc ? 0u : 1u) << 15) | ((b_kc ? 0u : 1u) << 16) |
                                       ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((p.m64 ? 64 : TC_BM) >> 4) << 24);
                for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    int z1, z2, m0, n0, kt0, kt1;
                    mp_decode(p, bn, tile, z1, z2, m0, n0, kt0, kt1);
                    if (kt0 >= kt1) continue;
                    mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1);
                    tc_fence_after();
                    if (dbg && blockIdx.x == 0 && tile < 3 * (int)gridDim.x) dbg[1024 + ph * 16 + 4 + 2 * (tile / gridDim.x)] = clock64();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * MP_BN);
                    for (int kt = kt0; kt < kt1; ++kt) {
                        mbar_wait(full0 + 8u * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + stage * S::STAGE_BYTES);
                        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                        for (int kk = 0; kk < TC_BK / 8; ++kk) {
                            const uint64_t da = a_kc ? umma_desc(sa + kk * 32, 16, 1024, 2) : umma_desc(sa + kk * 1024, 4096, 512, 1);
                            const uint64_t db = b_kc ? umma_desc(sb + kk * 32, 16, 1024, 2) : umma_desc(sb + kk * 1024, 4096, 512, 1);
                            umma_tf32(tmem_d, da, db, idesc, (kt > kt0 || kk > 0) ? 1u : 0u);
                        }
                        umma_commit(empty0 + 8u * stage);
                        if (++stage == MP_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(tfull0 + 8u * acc);
                    if (dbg && blockIdx.x == 0 && tile < 3 * (int)gridDim.x) dbg[1024 + ph * 16 + 5 + 2 * (tile / gridDim.x)] = clock64();
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue / elementwise warps ================================
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;                         // 0..255
        constexpr int ET = TC_EPI_WARPS * 32;
        float* buf = epi_buf + (warp - 4) * (32 * TC_EPI_LD);
        MpEpiState st{0, 0u};
        for (int ph = 0; ph < nph; ++ph) {
            const MPhase* P = phases + ph;
            if (ph > 0) {
                // the releasing thread's gpu-scope fence (below) has already invalidated this SM's L1
                mbar_wait(phase_bar, (uint32_t)((ph - 1) & 1));
            }
            if (dbg && blockIdx.x == 0 && et == 0) dbg[ph * 4 + 0] = clock64();
            if (P->copy_n > 0) {
                const float4* src = reinterpret_cast<const float4*>(P->copy_src);
                float4* dst = reinterpret_cast<float4*>(P->copy_dst);
                const long long n4 = P->copy_n >> 2;
                for (long long i = (long long)blockIdx.x * ET + et; i < n4; i += (long long)gridDim.x * ET) dst[i] = src[i];
            }
            if (P->kind == PK_GEMM) {
                const int bn = P->bn;
                switch (P->epi_kind) {
                    case EK_STORE: mp_epilogue_phase<EpiStore>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_PLAIN: mp_epilogue_phase<EpiPlain>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_ATOMIC: mp_epilogue_phase<EpiAtomic>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_GATE: mp_epilogue_phase<EpiGate>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_CAND: mp_epilogue_phase<EpiCand>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_RESCAND: mp_epilogue_phase<EpiResCand>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_B1: mp_epilogue_phase<EpiB1>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_B2: mp_epilogue_phase<EpiB2>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_B4: mp_epilogue_phase<EpiB4>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    case EK_B6: mp_epilogue_phase<EpiB6>(P, bn, tmem_base, tfull0, tempty0, st, q, half, buf, lane); break;
                    default: __trap();
                }
            } else {
                // reverse-step head: dy = dY[t] + carry ; residual-mix backward up to da3 ; d(mix) partial sums
                const HeadArgs a = load_functor<HeadArgs>(P);
                const float g = __ldg(a.mix_t);
                float part = 0.f;
                // float4 per thread (H % 4 == 0 is guaranteed by the builder), two quads in flight
                const long long n4 = a.n >> 2;
                const long long stride = (long long)gridDim.x * ET;
                for (long long i0 = (long long)blockIdx.x * ET + et; i0 < n4; i0 += 2 * stride) {
                    float4 dyv[2], cv[2], h1v[2], r2v[2], hcv[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const long long i = i0 + u * stride;
                        if (i < n4) {
                            dyv[u] = ld4(a.dY + 4 * i); cv[u] = ld4(a.carry + 4 * i); h1v[u] = ld4(a.H1 + 4 * i);
                            r2v[u] = ld4(a.R2 + 4 * i); hcv[u] = ld4(a.HC2 + 4 * i);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const long long i = i0 + u * stride;
                        if (i >= n4) continue;
                        const float4 dy = dyv[u] + cv[u];
                        const float4 res = r2v[u] * h1v[u] + one_minus(r2v[u]) * hcv[u];
                        const float4 pr = dy * (h1v[u] - res);
                        part += pr.x + pr.y + pr.z + pr.w;
                        const float4 dres = (1.f - g) * dy;
                        st4(a.DRES + 4 * i, dres);
                        st4(a.DH1 + 4 * i, g * dy + dres * r2v[u]);
                        const long long e = 4 * i, row = e / a.H;
                        const int c = (int)(e - row * a.H);
                        st4(a.DR + row * 3 * a.H + 2 * a.H + c, dres * one_minus(r2v[u]) * one_minus(hcv[u] * hcv[u]));
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (lane == 0) atomicAdd(a.dmix_t, part);
            }
            // ---- end of phase: publish this CTA's writes, then wait for every CTA ----
            // (the cooperative-groups grid.sync pattern: CTA barrier, then ONE thread fences, signals, spins, fences)
            if (dbg && blockIdx.x == 0 && et == 0) dbg[ph * 4 + 1] = clock64();
            asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
            if (et == 0) {
                if (dbg && blockIdx.x == 0) dbg[ph * 4 + 2] = clock64();
                if (ph + 1 < nph) {
                    __threadfence();
                    atomicAdd(gbar, 1u);
                    const unsigned int target = (unsigned int)(ph + 1) * gridDim.x;
                    long long t0 = 0;
                    for (uint32_t it = 0; ld_acquire_gpu(gbar) < target; ++it) {
                        if (it == 1024) t0 = clock64();
                        if (it > 1024 && (it & 255) == 0 && clock64() - t0 > 4000000000LL) __trap();
                    }
                    __threadfence();
                }
                if (dbg && blockIdx.x == 0) dbg[ph * 4 + 3] = clock64();
                mbar_arrive(phase_bar);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * MP_BN));
    }
}

// ---------------------------------------------------------------------------------------------
// host side: phase list builder
// ---------------------------------------------------------------------------------------------
struct MultiBuilder {
    std::vector<MPhase> ph;
    bool ok = true;
    struct Slot {
        bool seen = false, have_stride = false;
        GemmP p;
        int t = 0, Z = 0, bn = 0;
        bool a_kc = false, b_kc = false;
        long long sAt = 0, sBt = 0;
    };
    Slot slots[MP_SLOTS];
    int t_max = 0;

    static bool same_shape(const GemmP& a, const GemmP& b) {
        return a.M == b.M && a.N == b.N && a.K == b.K && a.KB == b.KB && a.lda == b.lda && a.ldb == b.ldb && a.sA1 == b.sA1 &&
               a.sA2 == b.sA2 && a.sAk == b.sAk && a.sB1 == b.sB1 && a.sB2 == b.sB2 && a.sBk == b.sBk && a.Z2 == b.Z2 &&
               a.splits == b.splits;
    }

    // slot: which contraction of the step this is (same shapes for every t, operand bases moving by a constant stride)
    template <bool A_KC, bool B_KC, class Epi>
    void add_gemm(int slot, int t, const GemmP& p, const Epi& epi, int Z) {
        static_assert(sizeof(Epi) <= MP_BLOB, "epilogue functor does not fit the phase blob");
        if (!ok) return;
        if (slot < 0 || slot >= MP_SLOTS || t < 0) { ok = false; return; }
        Slot& s = slots[slot];
        if (!s.seen) {
            s.seen = true; s.p = p; s.t = t; s.Z = Z; s.bn = p.N <= 64 ? 64 : 128; s.a_kc = A_KC; s.b_kc = B_KC;
        } else {
            if (!same_shape(p, s.p) || Z != s.Z || s.a_kc != A_KC || s.b_kc != B_KC) { ok = false; return; }
            const long long dt = t - s.t;
            if (dt != 0) {
                const long long dA = p.A - s.p.A, dB = p.B - s.p.B;
                if (dA % dt || dB % dt) { ok = false; return; }
                if (!s.have_stride) { s.sAt = dA / dt; s.sBt = dB / dt; s.have_stride = true; }
                else if (dA != dt * s.sAt || dB != dt * s.sBt) { ok = false; return; }
            } else if (p.A != s.p.A || p.B != s.p.B) { ok = false; return; }
        }
        if (t > t_max) t_max = t;
        MPhase m;
        memset(&m, 0, sizeof(m));
        const int bn = s.bn;
        const int Z2 = p.Z2 > 0 ? p.Z2 : 1;
        const int splits = p.splits > 0 ? p.splits : 1;
        TcP& tp = m.p;
        tp.M = p.M; tp.N = p.N; tp.K = p.K; tp.KB = p.KB; tp.Z2 = Z2; tp.splits = splits;
        tp.tiles_m = (p.M + TC_BM - 1) / TC_BM;
        tp.tiles_n = (p.N + bn - 1) / bn;
        const long long total = (long long)tp.tiles_m * tp.tiles_n * splits * Z;
        if (total > 2147483647LL || p.K < 8) { ok = false; return; }
        tp.total_tiles = (int)total;
        tp.d_mn = make_fastdiv((uint32_t)(tp.tiles_m * tp.tiles_n));
        tp.d_splits = make_fastdiv((uint32_t)splits);
        tp.d_z2 = make_fastdiv((uint32_t)Z2);
        tp.d_tn = make_fastdiv((uint32_t)tp.tiles_n);
        tp.vec = ((p.N & 3) == 0 && epi.vec_ok()) ? 1 : 0;
        tp.m64 = p.M <= 64 ? 1 : 0;
        tp.t = t;
        m.kind = PK_GEMM; m.a_kc = A_KC ? 1 : 0; m.b_kc = B_KC ? 1 : 0; m.bn = bn; m.epi_kind = EpiKindOf<Epi>::v; m.slot = slot;
        memcpy(m.blob, &epi, sizeof(Epi));
        ph.push_back(m);
    }
    void add_head(const HeadArgs& a) {
        if (!ok) return;
        if ((a.H & 3) || (a.n & 3) || !aligned16(a.dY) || !aligned16(a.carry) || !aligned16(a.H1) || !aligned16(a.R2) ||
            !aligned16(a.HC2) || !aligned16(a.DH1) || !aligned16(a.DRES) || !aligned16(a.DR)) {
            ok = false;
            return;
        }
        MPhase m;
        memset(&m, 0, sizeof(m));
        m.kind = PK_HEAD;
        memcpy(m.blob, &a, sizeof(a));
        ph.push_back(m);
    }
    // side copy executed at the start of the most recently added phase
    void copy_on_last(const float* src, float* dst, long long n) {
        if (!ok || ph.empty() || n <= 0) return;
        if ((n & 3) || !aligned16(src) || !aligned16(dst)) { ok = false; return; }
        ph.back().copy_src = src; ph.back().copy_dst = dst; ph.back().copy_n = n;
    }
    // dev_phases: device buffer of ph.size()*sizeof(MPhase) bytes (128-byte aligned); gbar: one device counter
    cudaError_t launch(void* dev_phases, unsigned int* gbar, cudaStream_t st) {
        if (!ok || ph.empty()) return cudaErrorNotSupported;
        if (reinterpret_cast<uintptr_t>(dev_phases) & 127) return cudaErrorNotSupported;
        MpMaps maps;
        memset(&maps, 0, sizeof(maps));
        int cA[MP_SLOTS][3], cB[MP_SLOTS][3], tdA[MP_SLOTS], tdB[MP_SLOTS];
        const int nT = t_max + 1;
        for (int i = 0; i < MP_SLOTS; ++i) {
            Slot& s = slots[i];
            if (!s.seen) continue;
            if (s.sAt < 0 || s.sBt < 0) return cudaErrorNotSupported;
            const GemmP& p = s.p;
            const int Z2 = p.Z2 > 0 ? p.Z2 : 1;
            const int Z1 = (s.Z + Z2 - 1) / Z2;
            const float* A0 = p.A - (long long)s.t * s.sAt;   // operand base at t = 0
            const float* B0 = p.B - (long long)s.t * s.sBt;
            if (!make_operand_map(&maps.a[i], A0, s.a_kc, p.M, p.K, p.lda, p.sAk, p.sA2, p.sA1, p.KB, Z2, Z1, TC_BM, &cA[i][0], &cA[i][1],
                                  &cA[i][2], s.sAt, nT, &tdA[i]) ||
                !make_operand_map(&maps.b[i], B0, s.b_kc, p.N, p.K, p.ldb, p.sBk, p.sB2, p.sB1, p.KB, Z2, Z1, s.bn, &cB[i][0], &cB[i][1],
                                  &cB[i][2], s.sBt, nT, &tdB[i]))
                return cudaErrorNotSupported;
        }
        for (MPhase& m : ph) {
            if (m.kind != PK_GEMM) continue;
            const int i = m.slot;
            m.p.cAk = cA[i][0]; m.p.cA2 = cA[i][1]; m.p.cA1 = cA[i][2];
            m.p.cBk = cB[i][0]; m.p.cB2 = cB[i][1]; m.p.cB1 = cB[i][2];
            m.p.tA_dim = tdA[i]; m.p.tB_dim = tdB[i];
        }
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(gemm_tc_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MpSmem::TOTAL);
            if (e != cudaSuccess) return e;
            configured = true;
        }
        cudaError_t e = cudaMemcpyAsync(dev_phases, ph.data(), ph.size() * sizeof(MPhase), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(gbar, 0, sizeof(unsigned int), st);
        if (e != cudaSuccess) return e;
        const MPhase* dp = reinterpret_cast<const MPhase*>(dev_phases);
        int n = (int)ph.size();
        long long* dbg = tc_debug_buffer();
        void* args[] = {(void*)&maps, (void*)&dp, (void*)&n, (void*)&gbar, (void*)&dbg};
        e = cudaLaunchCooperativeKernel((void*)gemm_tc_multi_kernel, dim3((unsigned)sm_count()), dim3(TC_THREADS), args, MpSmem::TOTAL, st);
        count_launch();
        return e != cudaSuccess ? e : cudaGetLastError();
    }
};

}  // namespace matgcn
