// TMA-fed tcgen05 (5th-gen tensor core) GEMM for sm_100a with TMEM accumulators.
//
// "Fast mode" engine: same problem description (GemmP), same epilogue functors and the same
// semantics as the fp32 SIMT kernel in gemm_simt.cuh, but the products run on the tensor cores
// as kind::tf32 MMAs (fp32 operands straight from HBM via TMA, truncated to TF32 by the tensor
// core, fp32 accumulation in tensor memory).  No operand copies or layout changes are needed,
// so every contraction of the path can be switched between the two engines per call.
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   : TMA producer  - cp.async.bulk.tensor.5d into a STAGES-deep 128B-swizzled smem ring
//   warp 1   : MMA issuer    - one lane issues tcgen05.mma (M=128, N=BN, K=8), commits to mbarriers
//   warp 2   : TMEM allocator (2 x BN fp32 columns: double-buffered accumulator)
//   warps 4-7: epilogue      - tcgen05.ld 32 rows x 32 columns per warp, transpose through padded
//                              smem so that the epilogue functor sees consecutive columns on
//                              consecutive lanes (coalesced global access)
// Tiles: 128 x BN output, 32-element (128 B) K slabs; batching (z1, z2) and k-batches are extra
// tensor-map dimensions, so one launch covers e.g. all nodes of a node-batched contraction.
// Out-of-range rows/columns/K are zero-filled by TMA and masked in the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_simt.cuh"

namespace matgcn {

constexpr int TC_EPI_LD = 36;  // epilogue staging row pitch (floats)
constexpr int TC_BM = 128;
constexpr int TC_BK = 32;  // fp32 elements per K slab = one 128-byte swizzle row
constexpr int TC_EPI_WARPS = 8;  // two warps per TMEM lane quadrant, each taking every other 32-column chunk
constexpr int TC_THREADS = (4 + TC_EPI_WARPS) * 32;

// Unsigned division by a launch-time constant (Granlund-Montgomery round-up method): the tile decode
// runs once per tile per role, and hardware integer division would cost ~200 cycles apiece there.
struct FastDiv {
    uint32_t d, mul, sh1, sh2;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1ULL << l) < d) ++l;  // ceil(log2 d)
    f.mul = (uint32_t)(((1ULL << 32) * ((1ULL << l) - d)) / d + 1);
    f.sh1 = l < 1 ? l : 1;
    f.sh2 = l == 0 ? 0 : l - 1;
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
    const uint32_t t = __umulhi(f.mul, n);
    return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

struct TcP {
    int M, N, K, KB, Z2, splits;
    int tiles_m, tiles_n;
    int total_tiles;        // Z * splits * tiles_m * tiles_n
    FastDiv d_mn, d_splits, d_z2, d_tn;
    int cA1, cA2, cAk;      // 1 if the operand really has that (strided) dimension, else coordinate 0
    int cB1, cB2, cBk;
    int tA_dim, tB_dim, t;  // persistent kernel: which coordinate (2..4, 0 = none) carries the time step, and its value
    int m64;                // 1: issue M=64 MMAs (tile rows 0..63 only; 16 accumulator rows per TMEM lane quadrant)
    int keepB;              // 1: B operand loads carry the L2 evict-last policy
    int vec;                // 1: N % 4 == 0 and the epilogue's pointers/pitches allow 16-byte accesses
    int dbg_mode;           // diagnostics only: 1 = skip epilogue stores
    long long* dbg;         // optional per-tile timeline of CTA 0 (tools/tc_timeline.py); null in production
    // L2 warm-up for the NEXT launches of the stream (weights a following per-node contraction streams from HBM, saved
    // activations the next elementwise-heavy kernel reads): this launch is L2-bound and leaves the HBM pipe idle.
    PfRange pf[8];
    int npf;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (it == 64) t0 = clock64();
        if (it > 64 && (it & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// same with an L2 cache policy (createpolicy) attached to the load
__device__ __forceinline__ void tma_load_5d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                 int c3, int c4, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor (SM100 UMMA): start address, leading/stride byte offsets (all >> 4),
// descriptor version 1, 128-byte swizzle.
// layout_type: 2 = SWIZZLE_128B (16-byte atoms; K-major tiles), 1 = SWIZZLE_128B_BASE32B (32-byte atoms; the only
// layout the tensor core accepts for MN-major 32-bit operands).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}

// Operand element traits: TF32 reads fp32 operands (32 per 128-byte slab, K = 8 per MMA), BF16 reads bf16 twins
// (64 per slab, K = 16 per MMA).  MN-major tiles: rows are K indices holding MN_BLOCK contiguous elements;
// 32-bit operands need the 32-byte-atom swizzle (4-row K groups), 16-bit ones the plain 128-byte swizzle (8-row groups).
template <bool BF16>
struct TcElem {
    static constexpr int ESIZE = BF16 ? 2 : 4;
    static constexpr int BK = 128 / ESIZE;            // elements per 128-byte K slab
    static constexpr int UMMA_K = 32 / ESIZE;         // K per tcgen05.mma
    static constexpr int MN_BLOCK = 128 / ESIZE;      // MN elements per swizzle row of an MN-major tile
    static constexpr int MN_BOX_BYTES = BK * 128;     // one MN-major TMA box: BK rows of 128 bytes
    static constexpr int MN_KSTEP_BYTES = UMMA_K * 128;
    static constexpr int MN_SBO = BF16 ? 1024 : 512;  // K-group stride
    static constexpr int MN_LAYOUT = BF16 ? 2 : 1;    // SWIZZLE_128B : SWIZZLE_128B_BASE32B
    static constexpr uint32_t FMT = BF16 ? 1u : 2u;   // instruction-descriptor operand format
};

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int TC_RES_LD = 68;  // row pitch (floats) of the residual-cell weight tiles in shared memory: conflict-free B fragments
template <int BN, bool FUSED = false>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * 128;  // 16 KB: 128 rows of one 128-byte slab
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = FUSED ? 5 : ((BN <= 64) ? 6 : (BN <= 128 ? 5 : 3));
    static constexpr int EPI_LD = TC_EPI_LD;  // staging row pitch in floats: 16-byte aligned rows, conflict-free float4 access
    static constexpr int EPI_BYTES = TC_EPI_WARPS * 32 * EPI_LD * 4;  // per epilogue warp: [32][36] floats
    static constexpr int BAR_BYTES = 256;
    static constexpr int EXTRA_BYTES = FUSED ? (3 * 64) * TC_RES_LD * 4 : 0;  // fused tail: Rg_h [128][68] + Ru_h [64][68]
    static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + EXTRA_BYTES + 1024;  // + alignment slack
};

// Epilogue of one output tile, run by one epilogue warp: TMEM -> registers -> staging tile -> epilogue functor.
//   tmem_acc: TMEM address of the accumulator (lane 0, first column); bn: tile width; q: lane quadrant of this warp;
//   half: which of the two warps of the quadrant (takes every other 32-column chunk); buf: [32][EPI_LD] staging tile.
template <class Epi>
__device__ __forceinline__ void tc_epilogue_tile(const Epi& epi, const TcP& p, uint32_t tmem_acc, int bn, int q, int half,
                                                 float* buf, int lane, int z1, int z2, int m0, int n0) {
    constexpr int LD = TC_EPI_LD;
    const int pM = p.M, pN = p.N, pvec = p.vec, pm64 = p.m64, pdbg = p.dbg_mode;  // hoisted: p may sit in memory
    // M=128: accumulator row r lives in TMEM lane r.  M=64: row r lives in lane 32*(r/16) + r%16, i.e. every
    // lane quadrant holds 16 rows, so all epilogue warps stay busy on the half-height tiles of this path.
    const int rows_per_q = pm64 ? 16 : 32;
    const int row_base = m0 + q * rows_per_q;
    const int row_lim = (pdbg & 4) ? (m0 + 128) : min(pM, row_base + rows_per_q);  // bit 2: raw lane dump
    if (row_base < row_lim) {
#pragma unroll 1
        for (int c = half; c < bn / 32; c += TC_EPI_WARPS / 4) {
            const int col_base = n0 + c * 32;
            if (col_base >= pN) break;
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // lane = accumulator row: park the 32 columns of that row in the staging tile
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(buf + lane * LD + 4 * j) =
                    make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3]));
            __syncwarp();
            if (pvec) {
                // lane -> (row = 4*it + lane/8, columns 4*(lane%8) .. +3): one 16-byte access per lane, 4 rows of
                // 128 contiguous bytes per warp instruction; loads of a batch are issued before any is consumed
                const int rq = lane >> 3, cq = lane & 7;
                const int col = col_base + 4 * cq;
                const bool col_ok = col < pN;
                constexpr int NBATCH = Epi::kBatch;
#pragma unroll 1
                for (int it0 = 0; it0 < 8; it0 += NBATCH) {
                    EpiIn4 in[NBATCH];
                    float4 v[NBATCH];
#pragma unroll
                    for (int u = 0; u < NBATCH; ++u) {
                        const int rl = 4 * (it0 + u) + rq;
                        v[u] = *reinterpret_cast<const float4*>(buf + rl * LD + 4 * cq);
                        if (col_ok && row_base + rl < row_lim) in[u] = epi.load4(z1, z2, row_base + rl, col);
                    }
#pragma unroll
                    for (int u = 0; u < NBATCH; ++u) {
                        const int rl = 4 * (it0 + u) + rq;
                        if (col_ok && row_base + rl < row_lim && !(pdbg & 1)) epi.store4(z1, z2, row_base + rl, col, v[u], in[u]);
                    }
                }
            } else {
                const int col = col_base + lane;
                const bool col_ok = col < pN;
#pragma unroll 1
                for (int rr0 = 0; rr0 < 32; rr0 += 8) {
                    EpiIn in[8];
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int row = row_base + rr0 + u;
                        v[u] = buf[(rr0 + u) * LD + lane];
                        if (col_ok && row < row_lim) in[u] = epi.load(z1, z2, row, col);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int row = row_base + rr0 + u;
                        if (col_ok && row < row_lim) epi.store(z1, z2, row, col, v[u], in[u]);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// Compile-time row index for the software-pipelined epilogue: every access to the register ring below uses a
// constant index, so the ring is never indexed dynamically (which would put it in local memory).
template <int V> struct IntC { static constexpr int value = V; };
template <int S, int N, class F>
__device__ __forceinline__ void tc_static_for(F& f) {
    if constexpr (S < N) {
        f(IntC<S>{});
        tc_static_for<S + 1, N>(f);
    }
}

// Software-pipelined form of the vectorised epilogue (p.vec): the whole tile loop of one epilogue warp.
// The epilogue functor's global reads do not depend on the accumulator and cost ~1000+ cycles under load, so they run
// R = Epi::kPipe accesses ahead of their use through a ring of registers - across tile boundaries: while the last R
// accesses of a tile are computed, the first R of the CTA's next tile are already in flight (for the half-height
// tiles of the node-batched contractions, M64, that is the whole next tile).
// Positions are cursors (Epi::Cur) computed once per 32-column chunk and bumped by four rows per access - no
// per-element index arithmetic.  A 32x32 chunk = 8 accesses per lane (lane -> row 4*it + lane/8, columns 4*(lane%8)..+3);
// half-height tiles keep 16 rows per TMEM lane quadrant, i.e. 4 accesses per chunk.
// SPLIT (half-height 64-wide tiles: the node-batched contractions of the reverse step and of the time-batched gradients, whose
// epilogue is a latency chain of ~2k cycles for 16 rows x 32 columns per warp): the two warps of a TMEM lane quadrant take
// ALTERNATE TILES instead of alternate column chunks, each with the whole 64-column row block, and the accumulator ring has
// four buffers - two tiles are in their epilogue at any time.
template <int BN, bool M64, bool SPLIT, class Epi, class Dec>
__device__ __forceinline__ void tc_epilogue_loop_pipe(const Epi& epi, const TcP& p, uint32_t tmem_base, const Dec& decode, uint32_t tfull0,
                                                      uint32_t tempty0, int q, int half, float* buf, int lane) {
    constexpr int LD = TC_EPI_LD;
    constexpr int NCH = SPLIT ? BN / 32 : (BN / 32 + 1) / 2;  // 32-column chunks per warp (two warps share a TMEM lane quadrant)
    constexpr int CSTEP = SPLIT ? 32 : 64;                    // column distance between consecutive chunks of this warp
    constexpr int NACC = SPLIT ? 4 : 2;                       // accumulator buffers in TMEM
    const int hoff = SPLIT ? 0 : half * 32;                   // first column of this warp inside the tile
    constexpr int ITS = M64 ? 4 : 8;        // accesses per chunk
    constexpr int ACC = NCH * ITS;          // accesses per lane and tile
    constexpr int R = Epi::kPipe < ACC ? Epi::kPipe : ACC;  // ring size = prefetch distance in accesses
    constexpr int ROWS_Q = M64 ? 16 : 32;
    const int pM = p.M, pN = p.N, pdbg = p.dbg_mode;
    const int rq = lane >> 3, cq = lane & 7;
    const bool do_loads = !(pdbg & 32);  // debug mode bit 5: no epilogue global reads (A/B measurements)
    const int total = p.total_tiles, stride = gridDim.x;

    struct Tile { int z1, z2, m0, n0; };
    auto next_nonempty = [&](int t, Tile& tl) {  // first tile >= t of this CTA with work in its split; total if none
        for (; t < total; t += stride) {
            int kt0, kt1;
            decode(t, tl.z1, tl.z2, tl.m0, tl.n0, kt0, kt1);
            if (kt0 < kt1) return t;
        }
        return total;
    };
    Tile cur, nxt;
    int tile = next_nonempty(blockIdx.x, cur);
    if (SPLIT && half == 1 && tile < total) tile = next_nonempty(tile + stride, cur);  // this warp takes the odd tiles of the CTA
    if (tile >= total) return;

    typename Epi::Cur lc, sc;
    EpiIn4 ring[R];
    // The accesses of a tile run in NR rounds of R (one ring revolution each): only the R accesses of a round are
    // unrolled (slot j = compile time), the round index is a run-time loop - this keeps the epilogue code small, which
    // matters because every launch of these short kernels starts with a cold instruction cache.
    constexpr int NR = ACC / R;
    static_assert(ACC % R == 0, "ring size must divide the accesses of a tile");
    // read access (round rd, slot j) of tile tl into ring[j]
    auto issue = [&](const Tile& tl, int rd, auto J_) {
        constexpr int j = decltype(J_)::value;
        const int a = rd * R + j;
        const int ci = a / ITS, it = a % ITS;
        const int row_base = tl.m0 + q * ROWS_Q;
        const int col = tl.n0 + hoff + 4 * cq + CSTEP * ci;
        if (it == 0) lc = epi.begin4(tl.z1, tl.z2, row_base + rq, col);
        if (col < pN && row_base + 4 * it + rq < min(pM, row_base + ROWS_Q) && do_loads) ring[j] = epi.load4(lc);
        epi.advance4(lc, 4);
    };
    {
        auto pro = [&](auto J_) { issue(cur, 0, J_); };
        tc_static_for<0, R>(pro);
    }
    int acc = SPLIT ? half : 0;
    uint32_t acc_phase = 0;
    while (true) {
        int ntile = next_nonempty(tile + stride, nxt);
        if (SPLIT && ntile < total) ntile = next_nonempty(ntile + stride, nxt);
        const bool has_next = ntile < total;
        if (p.dbg && blockIdx.x == 0 && (threadIdx.x == 128 || (SPLIT && threadIdx.x == 256))) p.dbg[(tile / stride) * 8 + 4] = clock64();
        mbar_wait(tfull0 + 8u * acc, acc_phase);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN);
        const int row_base = cur.m0 + q * ROWS_Q;
        const int row_lim = min(pM, row_base + ROWS_Q);
#pragma unroll 1
        for (int rd = 0; rd < NR; ++rd) {
            auto step = [&](auto J_) {
                constexpr int j = decltype(J_)::value;
                const int a = rd * R + j;
                const int ci = a / ITS, it = a % ITS;
                const int col = cur.n0 + hoff + 4 * cq + CSTEP * ci;
                const bool chunk_ok = cur.n0 + hoff + CSTEP * ci < pN && row_base < row_lim;  // warp-uniform
                if (it == 0 && chunk_ok) {
                    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(hoff + CSTEP * ci);
                    uint32_t r[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    __syncwarp();  // every lane is done reading the previous chunk (or tile) from the staging tile
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        *reinterpret_cast<float4*>(buf + lane * LD + 4 * jj) =
                            make_float4(__uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1]), __uint_as_float(r[4 * jj + 2]),
                                        __uint_as_float(r[4 * jj + 3]));
                    __syncwarp();
                }
                if (it == 0) sc = epi.begin4(cur.z1, cur.z2, row_base + rq, col);
                // arithmetic + writes of this access, then its ring slot is refilled with the access R further on
                if (chunk_ok) {
                    const int rl = 4 * it + rq;
                    const float4 v = *reinterpret_cast<const float4*>(buf + rl * LD + 4 * cq);
                    if (col < pN && row_base + rl < row_lim && !(pdbg & 1)) epi.store4(sc, v, ring[j]);
                }
                epi.advance4(sc, 4);
                if (rd + 1 < NR) issue(cur, rd + 1, J_);
                else if (has_next) issue(nxt, 0, J_);
            };
            tc_static_for<0, R>(step);
        }
        tc_fence_before();
        __syncwarp();
        if (p.dbg && blockIdx.x == 0 && (threadIdx.x == 128 || (SPLIT && threadIdx.x == 256))) p.dbg[(tile / stride) * 8 + 6] = clock64();
        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
        acc += SPLIT ? 2 : 1;
        if (acc >= NACC) { acc -= NACC; acc_phase ^= 1; }
        if (!has_next) break;
        tile = ntile;
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused tail of the forward step (EpiCandRes; BN = H = 64, one 32-column chunk per epilogue warp).
// Phase 1 is the candidate epilogue in the coalesced layout; h1 goes back into the staging tile, where the two warps
// of a TMEM lane quadrant (columns 0..31 / 32..63 of the same rows) find the full 64-wide rows.  Phases 2/3 are the
// residual gate and candidate products as mma.sync m16n8k8 TF32 (A fragments from the staging tiles, B fragments from
// the weight tiles in shared memory) with their elementwise epilogues in the accumulator-fragment layout.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

template <class Epi>
__device__ __forceinline__ void tc_epilogue_tile_candres(const Epi& e, const TcP& p, uint32_t tmem_acc, uint32_t tfull, uint32_t tfull_parity,
                                                         int q, int half, float* buf, float* buf_other, const float* wg_s,
                                                         const float* wu_s, int lane, int z1, int m0) {
    constexpr int LD = TC_EPI_LD;
    constexpr int H = 64;
    constexpr int WL = TC_RES_LD;
    const int rows_per_q = p.m64 ? 16 : 32;
    const int row_base = m0 + q * rows_per_q;           // row inside the node's [B, H] block
    const int row_lim = min(p.M, row_base + rows_per_q);
    const bool any_row = row_base < row_lim;            // same for both warps of the quadrant
    const long long g0 = (long long)z1 * e.rows_per_z;  // first (node, batch) row of this node

    // ---------------- phase 1: hc = tanh(acc + GX[2H:3H]); h1 = r*h + (1-r)*hc ----------------
    const int rq = lane >> 3, cq = lane & 7;
    const int col = half * 32 + 4 * cq;
    long long i = (g0 + row_base + rq) * H + col;
    long long j = (g0 + row_base + rq) * 3 * H + 2 * H + col;
    float4 gx[4], rr[4], hh[4];
    if (any_row) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (row_base + 4 * u + rq < row_lim) {
                gx[u] = ld4(e.GX + j + (long long)u * 12 * H);
                rr[u] = ld4(e.R + i + (long long)u * 4 * H);
                hh[u] = ld4(e.Hprev + i + (long long)u * 4 * H);
            }
    }
    mbar_wait(tfull, tfull_parity);
    tc_fence_after();
    if (!any_row) return;
    {
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
            *reinterpret_cast<float4*>(buf + lane * LD + 4 * jj) =
                make_float4(__uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1]), __uint_as_float(r[4 * jj + 2]),
                            __uint_as_float(r[4 * jj + 3]));
        __syncwarp();
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        if (16 * b >= rows_per_q) break;  // half-height tiles keep 16 rows per quadrant
        if (b == 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (row_base + 16 + 4 * u + rq < row_lim) {
                    gx[u] = ld4(e.GX + j + (long long)(4 + u) * 12 * H);
                    rr[u] = ld4(e.R + i + (long long)(4 + u) * 4 * H);
                    hh[u] = ld4(e.Hprev + i + (long long)(4 + u) * 4 * H);
                }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rl = 16 * b + 4 * u + rq;
            if (row_base + rl < row_lim) {
                float* sp = buf + rl * LD + 4 * cq;
                const float4 acc = *reinterpret_cast<const float4*>(sp);
                const float4 hc = tanh4(acc + gx[u], 1);
                const float4 h1 = rr[u] * hh[u] + one_minus(rr[u]) * hc;
                st4(e.HC + i + (long long)(4 * b + u) * 4 * H, hc);
                st4(e.H1 + i + (long long)(4 * b + u) * 4 * H, h1);
                *reinterpret_cast<float4*>(sp) = h1;
            }
        }
    }
    pair_sync(q);  // both 32-column halves of h1 are in the two staging tiles

    // ---------------- phases 2/3: residual GRU cell on h1, mix ----------------
    const int g = lane >> 2, tig = lane & 3;
    const float* k_lo = half == 0 ? buf : buf_other;   // columns (= reduction index) 0..31 of the rows
    const float* k_hi = half == 0 ? buf_other : buf;   // columns 32..63
    const float m = __ldg(e.mix_t);
    for (int mb = 0; mb * 16 < rows_per_q; ++mb) {
        const int ra = 16 * mb + g, rb = ra + 8;       // staging rows of this thread's fragments
        const bool va = row_base + ra < row_lim, vb = row_base + rb < row_lim;
        const long long ga = g0 + row_base + ra, gb = ga + 8;
        const int c0 = 32 * half + 2 * tig;            // first of this thread's output columns (+ 8*nt)
        // pre-activation inputs of the residual gate (issued before the products)
        float2 xz[4][2], xr[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            if (va) { xz[nt][0] = ld2(e.RX + ga * 3 * H + c0 + 8 * nt); xr[nt][0] = ld2(e.RX + ga * 3 * H + H + c0 + 8 * nt); }
            if (vb) { xz[nt][1] = ld2(e.RX + gb * 3 * H + c0 + 8 * nt); xr[nt][1] = ld2(e.RX + gb * 3 * H + H + c0 + 8 * nt); }
        }
        float cz[4][4], cr[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int x = 0; x < 4; ++x) { cz[nt][x] = 0.f; cr[nt][x] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const float* src = (ks < 4 ? k_lo : k_hi) + (ks & 3) * 8 + tig;
            uint32_t a[4];
            a[0] = __float_as_uint(src[ra * LD]); a[1] = __float_as_uint(src[rb * LD]);
            a[2] = __float_as_uint(src[ra * LD + 4]); a[3] = __float_as_uint(src[rb * LD + 4]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float* wz = wg_s + (32 * half + 8 * nt + g) * WL + 8 * ks + tig;
                mma_tf32_16x8x8(cz[nt], a, __float_as_uint(wz[0]), __float_as_uint(wz[4]));
                const float* wr = wz + 64 * WL;
                mma_tf32_16x8x8(cr[nt], a, __float_as_uint(wr[0]), __float_as_uint(wr[4]));
            }
        }
        float h1v[4][4], r2v[4][4], zhv[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float2 ha = *reinterpret_cast<const float2*>(buf + ra * LD + 8 * nt + 2 * tig);
            const float2 hb = *reinterpret_cast<const float2*>(buf + rb * LD + 8 * nt + 2 * tig);
            h1v[nt][0] = ha.x; h1v[nt][1] = ha.y; h1v[nt][2] = hb.x; h1v[nt][3] = hb.y;
            float z2[4];
            z2[0] = cz[nt][0] + xz[nt][0].x; z2[1] = cz[nt][1] + xz[nt][0].y; z2[2] = cz[nt][2] + xz[nt][1].x; z2[3] = cz[nt][3] + xz[nt][1].y;
            r2v[nt][0] = cr[nt][0] + xr[nt][0].x; r2v[nt][1] = cr[nt][1] + xr[nt][0].y;
            r2v[nt][2] = cr[nt][2] + xr[nt][1].x; r2v[nt][3] = cr[nt][3] + xr[nt][1].y;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                z2[x] = sigmoid_fast(z2[x]);
                r2v[nt][x] = sigmoid_fast(r2v[nt][x]);
                zhv[nt][x] = z2[x] * h1v[nt][x];
            }
            const int c = c0 + 8 * nt;
            if (va) { st2(e.Z2 + ga * H + c, z2[0], z2[1]); st2(e.R2 + ga * H + c, r2v[nt][0], r2v[nt][1]); st2(e.ZH2 + ga * H + c, zhv[nt][0], zhv[nt][1]); }
            if (vb) { st2(e.Z2 + gb * H + c, z2[2], z2[3]); st2(e.R2 + gb * H + c, r2v[nt][2], r2v[nt][3]); st2(e.ZH2 + gb * H + c, zhv[nt][2], zhv[nt][3]); }
        }
        // candidate pre-activation inputs (in flight during the exchange below)
        float2 xu[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            if (va) xu[nt][0] = ld2(e.RX + ga * 3 * H + 2 * H + c0 + 8 * nt);
            if (vb) xu[nt][1] = ld2(e.RX + gb * 3 * H + 2 * H + c0 + 8 * nt);
        }
        pair_sync(q);  // both warps are done reading h1 from the staging tiles
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<float2*>(buf + ra * LD + 8 * nt + 2 * tig) = make_float2(zhv[nt][0], zhv[nt][1]);
            *reinterpret_cast<float2*>(buf + rb * LD + 8 * nt + 2 * tig) = make_float2(zhv[nt][2], zhv[nt][3]);
        }
        pair_sync(q);  // z2*h1 is in place
        float cf[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int x = 0; x < 4; ++x) cf[nt][x] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const float* src = (ks < 4 ? k_lo : k_hi) + (ks & 3) * 8 + tig;
            uint32_t a[4];
            a[0] = __float_as_uint(src[ra * LD]); a[1] = __float_as_uint(src[rb * LD]);
            a[2] = __float_as_uint(src[ra * LD + 4]); a[3] = __float_as_uint(src[rb * LD + 4]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float* wu = wu_s + (32 * half + 8 * nt + g) * WL + 8 * ks + tig;
                mma_tf32_16x8x8(cf[nt], a, __float_as_uint(wu[0]), __float_as_uint(wu[4]));
            }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float hc2[4], y[4];
            hc2[0] = cf[nt][0] + xu[nt][0].x; hc2[1] = cf[nt][1] + xu[nt][0].y; hc2[2] = cf[nt][2] + xu[nt][1].x; hc2[3] = cf[nt][3] + xu[nt][1].y;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                hc2[x] = tanh_fast(hc2[x]);
                const float res = r2v[nt][x] * h1v[nt][x] + (1.f - r2v[nt][x]) * hc2[x];
                y[x] = m * h1v[nt][x] + (1.f - m) * res;
            }
            const int c = c0 + 8 * nt;
            if (va) {
                st2(e.HC2 + ga * H + c, hc2[0], hc2[1]);
                st2(e.Y + ga * H + c, y[0], y[1]);
                if (e.Y16) *reinterpret_cast<__nv_bfloat162*>(e.Y16 + ga * H + c) = __floats2bfloat162_rn(y[0], y[1]);
            }
            if (vb) {
                st2(e.HC2 + gb * H + c, hc2[2], hc2[3]);
                st2(e.Y + gb * H + c, y[2], y[3]);
                if (e.Y16) *reinterpret_cast<__nv_bfloat162*>(e.Y16 + gb * H + c) = __floats2bfloat162_rn(y[2], y[3]);
            }
        }
    }
    pair_sync(q);  // the partner is done reading this warp's staging tile: the next tile may overwrite it
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// VEC: the vectorised, software-pipelined epilogue (N % 4 == 0 and 16-byte aligned epilogue operands); the scalar
// epilogue is a separate instantiation because ptxas spills when both live in one kernel.
// M64 (with VEC): half-height tiles, known at compile time so that the epilogue pipeline enumerates only their rows.
template <int BN, bool A_KC, bool B_KC, bool BF16, class Epi, bool VEC, bool M64 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcP p, const Epi epi) {
    constexpr bool FUSED = IsFusedRes<Epi>::value;
    using S = TcSmem<BN, FUSED>;
    using E = TcElem<BF16>;
    constexpr int STAGES = S::STAGES;
    constexpr int BK = E::BK;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 128B-swizzled tiles need 1024-byte alignment; keep the pointer derived from the __shared__ symbol so that
    // the epilogue's staging accesses compile to LDS/STS instead of generic loads/stores
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* stage_base = smem;
    float* epi_buf = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::EPI_BYTES);
    // bars: [0,STAGES) full, [STAGES,2*STAGES) empty, then tmem_full[2], tmem_empty[2]; then the TMEM base slot
    constexpr bool SPLIT = VEC && M64 && BN == 64 && !FUSED;   // see tc_epilogue_loop_pipe
    constexpr int NACC = SPLIT ? 4 : 2;                        // accumulator buffers of BN TMEM columns each
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NACC);
    float* res_w = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::EPI_BYTES + S::BAR_BYTES);  // FUSED only

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + NACC + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < NACC; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), SPLIT ? TC_EPI_WARPS / 2 : TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(NACC * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the tail of the
    // previous kernel in the stream; from here on its results are needed.  No-ops for an ordinary launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int kt_per_kb = (p.K + BK - 1) / BK;
    const int kt_total = p.KB * kt_per_kb;
    const int kt_per_split = (kt_total + p.splits - 1) / p.splits;
    const int tiles_mn = p.tiles_m * p.tiles_n;

    // tile -> (z1, z2, split, m0, n0, kt0, kt1)
    auto decode = [&](int tile, int& z1, int& z2, int& m0, int& n0, int& kt0, int& kt1) {
        const int rest = (int)fdiv((uint32_t)tile, p.d_mn);
        const int mn = tile - rest * tiles_mn;
        const int z = (int)fdiv((uint32_t)rest, p.d_splits);
        const int split = rest - z * p.splits;
        z1 = (int)fdiv((uint32_t)z, p.d_z2);
        z2 = z - z1 * p.Z2;
        const int tm = (int)fdiv((uint32_t)mn, p.d_tn);
        m0 = tm * TC_BM;
        n0 = (mn - tm * p.tiles_n) * BN;
        kt0 = split * kt_per_split;
        kt1 = min(kt_total, kt0 + kt_per_split);
    };

    // Register re-partition between the warpgroups: the producer / MMA / allocator warps (warpgroup 0) need few
    // registers, the two epilogue warpgroups hold double-buffered epilogue inputs (128*56 + 256*224 = 384*168).
    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const bool keepB = p.keepB != 0;
            const uint64_t polB = l2_policy_evict_last();
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int z1, z2, m0, n0, kt0, kt1;
                decode(tile, z1, z2, m0, n0, kt0, kt1);
                if (p.dbg && blockIdx.x == 0) p.dbg[(tile / gridDim.x) * 8 + 0] = clock64();
                // (k-batch, offset inside it) advance incrementally: this single thread's instruction latency is on the critical
                // path of the short launches, a division per k-block is not free
                int kb = kt0 / kt_per_kb;
                int k0 = (kt0 - kb * kt_per_kb) * BK;
                const int k_end = kt_per_kb * BK;
                for (int kt = kt0; kt < kt1; ++kt, k0 += BK) {
                    if (k0 == k_end) { k0 = 0; ++kb; }
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_u32(stage_base + stage * S::STAGE_BYTES);
                    const uint32_t sb = sa + S::A_BYTES;
                    mbar_expect_tx(full_bar(stage), S::STAGE_BYTES);
                    if (A_KC) {
                        tma_load_5d(sa, &tmA, full_bar(stage), k0, m0, kb * p.cAk, z2 * p.cA2, z1 * p.cA1);
                    } else {
#pragma unroll
                        for (int j = 0; j < TC_BM / E::MN_BLOCK; ++j)
                            tma_load_5d(sa + j * E::MN_BOX_BYTES, &tmA, full_bar(stage), m0 + E::MN_BLOCK * j, k0, kb * p.cAk, z2 * p.cA2,
                                        z1 * p.cA1);
                    }
                    if (B_KC) {
                        if (keepB) tma_load_5d_hint(sb, &tmB, full_bar(stage), k0, n0, kb * p.cBk, z2 * p.cB2, z1 * p.cB1, polB);
                        else tma_load_5d(sb, &tmB, full_bar(stage), k0, n0, kb * p.cBk, z2 * p.cB2, z1 * p.cB1);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / E::MN_BLOCK; ++j) {
                            if (keepB)
                                tma_load_5d_hint(sb + j * E::MN_BOX_BYTES, &tmB, full_bar(stage), n0 + E::MN_BLOCK * j, k0, kb * p.cBk,
                                                 z2 * p.cB2, z1 * p.cB1, polB);
                            else
                                tma_load_5d(sb + j * E::MN_BOX_BYTES, &tmB, full_bar(stage), n0 + E::MN_BLOCK * j, k0, kb * p.cBk, z2 * p.cB2,
                                            z1 * p.cB1);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (E::FMT << 7) | (E::FMT << 10) | ((A_KC ? 0u : 1u) << 15) | ((B_KC ? 0u : 1u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((p.m64 ? 64 : TC_BM) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int z1, z2, m0, n0, kt0, kt1;
                decode(tile, z1, z2, m0, n0, kt0, kt1);
                if (kt0 >= kt1) continue;  // empty split: nothing to accumulate, epilogue skips it too
                if (p.dbg && blockIdx.x == 0) p.dbg[(tile / gridDim.x) * 8 + 1] = clock64();
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                if (p.dbg && blockIdx.x == 0) p.dbg[(tile / gridDim.x) * 8 + 2] = clock64();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kt = kt0; kt < kt1; ++kt) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_base + stage * S::STAGE_BYTES);
                    const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BK / E::UMMA_K; ++kk) {
                        // K-major : 128B swizzle, 8-row groups 1024 B apart (SBO); one MMA step advances 32 B inside the row
                        // MN-major: rows are K indices holding MN_BLOCK contiguous elements; K groups MN_SBO apart,
                        //           MN blocks one TMA box (MN_BOX_BYTES) apart (LBO); one MMA step advances UMMA_K rows
                        const uint64_t da = A_KC ? umma_desc(sa + kk * 32, 16, 1024, 2)
                                                 : umma_desc(sa + kk * E::MN_KSTEP_BYTES, E::MN_BOX_BYTES, E::MN_SBO, E::MN_LAYOUT);
                        const uint64_t db = B_KC ? umma_desc(sb + kk * 32, 16, 1024, 2)
                                                 : umma_desc(sb + kk * E::MN_KSTEP_BYTES, E::MN_BOX_BYTES, E::MN_SBO, E::MN_LAYOUT);
                        if (BF16) umma_bf16(tmem_d, da, db, idesc, (kt > kt0 || kk > 0) ? 1u : 0u);
                        else umma_tf32(tmem_d, da, db, idesc, (kt > kt0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));  // accumulator complete
                if (p.dbg && blockIdx.x == 0) p.dbg[(tile / gridDim.x) * 8 + 3] = clock64();
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ================================ L2 warm-up for the next launch ================================
        for (int r = 0; r < p.npf; ++r) {
            const PfRange pr = p.pf[r];
            const int lines = pr.chunk >> 7;                          // 128-byte lines per chunk
            const long long total = (long long)pr.n * lines;
            for (long long i = (long long)blockIdx.x * 32 + lane; i < total; i += (long long)gridDim.x * 32) {
                const long long ch = i / lines;
                const int ln = (int)(i - ch * lines);
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pr.base + ch * pr.stride + ((long long)ln << 7)));
            }
        }
        // ================================ epilogue-input prefetch ================================
        // The epilogue's global reads (saved activations, pre-activations: mostly HBM-resident) queue behind the
        // operand stream when they are issued on demand; this warp touches them tile by tile right away, so that
        // the epilogue warps find them in L2.
        if ((VEC || FUSED) && !(p.dbg_mode & 16)) {  // debug mode bit 4: no prefetch (A/B measurements)
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int z1, z2, m0, n0, kt0, kt1;
                decode(tile, z1, z2, m0, n0, kt0, kt1);
                if (kt0 >= kt1) continue;
                const int rows = min(p.m64 ? 64 : TC_BM, p.M - m0);
                constexpr int CH = BN / 32;
                for (int idx = lane; idx < rows * CH; idx += 32) {
                    const int row = m0 + idx / CH, col = n0 + (idx % CH) * 32;
                    if (col < p.N) epi.prefetch4(epi.begin4(z1, z2, row, col));
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        const int q = warp & 3;             // TMEM lane quadrant this warp may read: lanes [32q, 32q+32)
        const int half = (warp - 4) >> 2;   // which of the two warps of that quadrant (column-chunk parity)
        float* buf = epi_buf + (warp - 4) * (32 * S::EPI_LD);  // staging tile [32][EPI_LD] of this warp
        if constexpr (FUSED) {
            // residual-cell weights Rg_h [128][64] and Ru_h [64][64] -> padded tiles in shared memory, once per CTA
            const int et = threadIdx.x - 128;
            for (int idx = et; idx < 3 * 64 * 16; idx += TC_EPI_WARPS * 32) {
                const int n = idx >> 4, k4 = (idx & 15) * 4;
                const float4 v = n < 128 ? ld4(epi.RgH + n * 64 + k4) : ld4(epi.RuH + (n - 128) * 64 + k4);
                *reinterpret_cast<float4*>(res_w + n * TC_RES_LD + k4) = v;
            }
            asm volatile("bar.sync 5, 256;" ::: "memory");
        }
        if constexpr (VEC && !FUSED) {
            tc_epilogue_loop_pipe<BN, M64, SPLIT>(epi, p, tmem_base, decode, tfull_bar(0), tempty_bar(0), q, half, buf, lane);
        } else {
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int z1, z2, m0, n0, kt0, kt1;
                decode(tile, z1, z2, m0, n0, kt0, kt1);
                if (kt0 >= kt1) continue;
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[(tile / gridDim.x) * 8 + 4] = clock64();
                if constexpr (FUSED) {
                    tc_epilogue_tile_candres(epi, p, tmem_base + (uint32_t)(acc * BN), tfull_bar(acc), acc_phase, q, half, buf,
                                             epi_buf + ((warp - 4) ^ 4) * (32 * S::EPI_LD), res_w, res_w + 128 * TC_RES_LD, lane, z1, m0);
                } else {
                    mbar_wait(tfull_bar(acc), acc_phase);
                    tc_fence_after();
                    tc_epilogue_tile(epi, p, tmem_base + (uint32_t)(acc * BN), BN, q, half, buf, lane, z1, z2, m0, n0);
                }
                tc_fence_before();
                __syncwarp();
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[(tile / gridDim.x) * 8 + 6] = clock64();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NACC * BN));
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmapEncodeFn tmap_encoder() {
    static TmapEncodeFn fn = []() -> TmapEncodeFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        if (q != cudaDriverEntryPointSuccess) return nullptr;
        return (TmapEncodeFn)p;
    }();
    return fn;
}

// Debug hook: when set (matgcn_debug_set_timeline), CTA 0 of every tensor-core launch writes clock64() stamps.
inline long long*& tc_debug_buffer() {
    static long long* p = nullptr;
    return p;
}

// record the timeline only for the n-th tensor-core launch after it was armed
inline int& tc_debug_countdown() {
    static int n = 0;
    return n;
}
inline int& tc_debug_mode() {
    static int m = 0;
    return m;
}

// Programmatic dependent launch of the tensor-core kernels (MATGCN_PDL=0 disables it).
inline bool tc_use_pdl() {
    static bool on = []() {
        const char* e = getenv("MATGCN_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

inline int sm_count() {   // of the current device (cached per device: one process may drive several)
    static int n[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (n[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[dev] = v > 0 ? v : 148;
    }
    return n[dev];
}

// Operand as a rank-5 tensor map {inner, outer, KB, Z2, Z1}.  `kc`: K is the contiguous (inner) dimension.
// Optional time dimension (persistent recurrence kernel): when nT > 1 and st != 0 one of the unused outer
// dimensions becomes the time step (size nT, stride st floats) and *tdim tells which coordinate carries t.
inline bool make_operand_map(CUtensorMap* map, const void* base, bool kc, int mn, int K, int ld, long long sk, long long s2,
                             long long s1, int KB, int Z2, int Z1, int box_mn_rows, int* ck, int* c2, int* c1,
                             long long st = 0, int nT = 1, int* tdim = nullptr, bool bf16 = false) {
    TmapEncodeFn enc = tmap_encoder();
    if (!enc) return false;
    const int es = bf16 ? 2 : 4;            // element size
    const int am = 16 / es - 1;             // strides (in elements) must keep 16-byte alignment
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & am) || (sk & am) || (s2 & am) || (s1 & am) || (st & am) || st < 0) return false;
    *ck = (KB > 1 && sk != 0) ? 1 : 0;
    *c2 = (Z2 > 1 && s2 != 0) ? 1 : 0;
    *c1 = (Z1 > 1 && s1 != 0) ? 1 : 0;
    if (KB > 1 && sk == 0) return false;  // a reduction over identical slabs never occurs on this path
    const cuuint64_t inner = kc ? (cuuint64_t)K : (cuuint64_t)mn;
    const cuuint64_t outer = kc ? (cuuint64_t)mn : (cuuint64_t)K;
    const cuuint64_t row_bytes = (cuuint64_t)ld * es;
    cuuint64_t dims[5] = {inner, outer, (cuuint64_t)(*ck ? KB : 1), (cuuint64_t)(*c2 ? Z2 : 1), (cuuint64_t)(*c1 ? Z1 : 1)};
    const cuuint64_t dummy = row_bytes * outer;
    cuuint64_t strides[4] = {row_bytes, *ck ? (cuuint64_t)sk * es : dummy, *c2 ? (cuuint64_t)s2 * es : dummy,
                             *c1 ? (cuuint64_t)s1 * es : dummy};
    if (tdim) *tdim = 0;
    if (nT > 1 && st != 0) {
        int d = !*ck ? 2 : (!*c2 ? 3 : (!*c1 ? 4 : -1));
        if (d < 0 || !tdim) return false;
        dims[d] = (cuuint64_t)nT;
        strides[d - 1] = (cuuint64_t)st * es;
        *tdim = d;
    }
    for (int i = 0; i < 4; ++i)
        if (strides[i] == 0 || (strides[i] & 15) || strides[i] >= (1ULL << 40)) return false;
    const cuuint32_t slab = (cuuint32_t)(128 / es);  // elements per 128-byte row
    cuuint32_t box[5] = {slab, (cuuint32_t)(kc ? box_mn_rows : slab), 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (kc || bf16) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Returns cudaErrorNotSupported when the problem does not meet the TMA alignment rules (caller falls back
// to the SIMT engine); any other error is a real launch failure.
template <int BN, bool A_KC, bool B_KC, class Epi, bool BF16 = false>
inline cudaError_t launch_gemm_tc(const GemmP& p, const Epi& epi, int Z, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0 || Z <= 0) return cudaSuccess;
    const int Z2 = p.Z2 > 0 ? p.Z2 : 1;
    const int Z1 = (Z + Z2 - 1) / Z2;
    const int splits = p.splits > 0 ? p.splits : 1;
    TcP t;
    t.M = p.M; t.N = p.N; t.K = p.K; t.KB = p.KB; t.Z2 = Z2; t.splits = splits;
    t.tiles_m = (p.M + TC_BM - 1) / TC_BM;
    t.tiles_n = (p.N + BN - 1) / BN;
    const long long total = (long long)t.tiles_m * t.tiles_n * splits * Z;
    if (total > 2147483647LL) return cudaErrorNotSupported;
    t.total_tiles = (int)total;
    t.d_mn = make_fastdiv((uint32_t)(t.tiles_m * t.tiles_n));
    t.d_splits = make_fastdiv((uint32_t)splits);
    t.d_z2 = make_fastdiv((uint32_t)Z2);
    t.d_tn = make_fastdiv((uint32_t)t.tiles_n);
    t.tA_dim = t.tB_dim = t.t = 0;
    t.keepB = p.keepB;
    t.npf = p.npf < 0 ? 0 : (p.npf > 8 ? 8 : p.npf);
    for (int r = 0; r < t.npf; ++r) t.pf[r] = p.pf[r];
    t.vec = ((p.N & 3) == 0 && epi.vec_ok()) ? 1 : 0;
    // M <= 64 (every node-batched contraction at batch 64): M=64 MMAs, whose accumulator spreads 16 rows over each
    // TMEM lane quadrant (layout verified on hardware by tools/m64_probe.py), halve the MMA work and keep all
    // epilogue warps busy
    t.m64 = (p.M <= 64 || (tc_debug_mode() & 8)) ? 1 : 0;
    t.dbg = (tc_debug_buffer() && tc_debug_countdown()-- == 0) ? tc_debug_buffer() : nullptr;
    t.dbg_mode = tc_debug_mode();
    CUtensorMap ma, mb;
    const void* Aop = BF16 ? (const void*)p.A16 : (const void*)p.A;
    const void* Bop = BF16 ? (const void*)p.B16 : (const void*)p.B;
    if (!Aop || !Bop) return cudaErrorNotSupported;
    if (!make_operand_map(&ma, Aop, A_KC, p.M, p.K, p.lda, p.sAk, p.sA2, p.sA1, p.KB, Z2, Z1, TC_BM, &t.cAk, &t.cA2, &t.cA1, 0, 1,
                          nullptr, BF16))
        return cudaErrorNotSupported;
    if (!make_operand_map(&mb, Bop, B_KC, p.N, p.K, p.ldb, p.sBk, p.sB2, p.sB1, p.KB, Z2, Z1, BN, &t.cBk, &t.cB2, &t.cB1, 0, 1,
                          nullptr, BF16))
        return cudaErrorNotSupported;
    constexpr bool FUSED = IsFusedRes<Epi>::value;
    using S = TcSmem<BN, FUSED>;
    void (*kern)(const CUtensorMap, const CUtensorMap, const TcP, const Epi);
    int variant = 0;
    if constexpr (FUSED) {
        if (!t.vec || BN != 64 || p.N != 64 || t.tiles_n != 1) return cudaErrorNotSupported;  // caller launches the three unfused contractions
        kern = gemm_tc_kernel<BN, A_KC, B_KC, BF16, Epi, true>;
    } else if (!t.vec) {
        if constexpr (EpiVecOnly<Epi>::value) return cudaErrorNotSupported;
        else kern = gemm_tc_kernel<BN, A_KC, B_KC, BF16, Epi, false>;
    } else if (t.m64) {
        kern = gemm_tc_kernel<BN, A_KC, B_KC, BF16, Epi, true, true>;
        variant = 2;
    } else {
        kern = gemm_tc_kernel<BN, A_KC, B_KC, BF16, Epi, true, false>;
        variant = 1;
    }
    static bool configured[64][3] = {};  // per device (the attribute is per device), template instantiation and epilogue flavour
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev][variant]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev][variant] = true;
    }
    const int grid = total < sm_count() ? (int)total : sm_count();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tc_use_pdl() ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ma, mb, t, epi);
    count_launch();
    return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace matgcn
