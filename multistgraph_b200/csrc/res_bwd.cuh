// Fused head of the reverse step (fast modes, rnn_units = 64): the elementwise chain rule of the sigma-mix and of the
// residual GRU cell together with its two small products, for 16 (node, batch) rows per warp:
//
//   B0 (was bwd_head_kernel)   dy = dY_t + carry ; dmix_t += sum dy*(h1 - res) ; dres = (1-g) dy ; dh1 = g dy + dres r2
//                               DR[2H:3H] = dres (1-r2)(1-hc2^2) ; DR[H:2H] = dres (h1-hc2) r2 (1-r2)
//   B1 (was a tcgen05 launch)   dzh2 = DR[2H:3H] x Ru_h ; dh1 += dzh2 z2 ; DR[0:H] = dzh2 h1 z2 (1-z2)
//   B2 (was a tcgen05 launch)   dh1 += DR[0:2H] x Rg_h ; DHD = dh1 r ; DG[2H:3H] = dh1 (1-r)(1-hc^2) ; DG[H:2H] = dh1 (h-hc) r (1-r)
//
// The products are [16, 64] x [64, 64] and [16, 128] x [128, 64] per warp: far too small for a TMA/tcgen05 pipeline
// (three launches of ~20 us each were almost all fixed cost), so they run as warp-level mma.sync m16n8k8 TF32 with
// the two weight tiles resident in shared memory.  dh1, dres and dzh2 never touch global memory.
//
// Everything lives in the accumulator-fragment layout, with both the k and the n index of the products permuted so
// that (i) a thread owns FOUR consecutive columns of rows g and g+8 in every 16-column block p (16-byte global
// accesses, 64 contiguous bytes per row per instruction) and (ii) the values a thread computes elementwise are already
// the A fragments of the next product (no shared-memory round trip, no shuffles):
//   n-tile nt = 2p+s, hardware column j      <->  actual column 16p + 4(j/2) + 2s + (j%2)
//   k-step ks = 2p+q, hardware k = tig|tig+4 <->  actual k      16p + 4tig + 2q + (0|1)
// The weights sit in shared memory as interleaved k-pairs, [k/2][n][2], so a B fragment is one 8-byte LDS; each row is
// rotated by rot(tig) in {0,4,16,20} floats, which makes those loads conflict-free under the permuted n.
// The second product is accumulated as soon as its A columns exist (block p of DR[H:2H] during the head, block p of
// DR[0:H] during the epilogue of the first product), so at most two 16x64 accumulators are live.
#pragma once
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace matgcn {

struct ResBwdArgs {
    const float* dY;     // [rows, H] upstream gradient of the layer output at step t
    const float* carry;  // [rows, H] gradient carried from step t+1
    const float* H1; const float* R2; const float* HC2; const float* Z2;   // saved residual-cell activations of step t
    const float* Hprev; const float* R; const float* HC;                    // saved main-cell activations of step t
    const float* mix_t;  // sigma(weights_gru[l, t])
    const float* RuH;    // [H, H]  dense Ruw[:, Cin:]
    const float* RgH;    // [2H, H] dense Rgw[:, Cin:]
    float* DR;           // [rows, 3H] out: residual-cell pre-activation gradients
    float* DG;           // [rows, 3H] out: columns H..3H of the main-cell pre-activation gradients
    __nv_bfloat16* DG16; // may be null: bf16 twin of DG
    float* DHD;          // [rows, H] out: dh1 * r (direct path to h_{t-1})
    float* dmix_t;       // += sum dy * (h1 - res)
    long long rows;
};

constexpr int RB_H = 64;
constexpr int RB_WARPS = 12;
constexpr int RB_ROW = 128;                          // floats per k-pair row: 64 n x 2
constexpr int RB_SMEM_FLOATS = (32 + 64) * RB_ROW;   // Ru_h: 32 k-pairs, Rg_h: 64 k-pairs

__device__ __forceinline__ int rb_rot(int kpair) { const int t = (kpair >> 1) & 3; return 4 * (t & 1) + 16 * (t >> 1); }

// one k-step (8 permuted k's) of a [16, *] x [*, 64] product: A fragment af, B from k-pair row `wrow`
__device__ __forceinline__ void rb_kstep(float (&acc)[8][4], const uint32_t (&af)[4], const float* wrow, int cg) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const float2 w = *reinterpret_cast<const float2*>(wrow + ((32 * (nt >> 1) + 4 * (nt & 1) + cg) & (RB_ROW - 1)));
        mma_tf32_16x8x8(acc[nt], af, __float_as_uint(w.x), __float_as_uint(w.y));
    }
}
__device__ __forceinline__ void rb_afrag(uint32_t (&af)[4], const float4& va, const float4& vb, int q) {
    af[0] = __float_as_uint(q ? va.z : va.x); af[1] = __float_as_uint(q ? vb.z : vb.x);
    af[2] = __float_as_uint(q ? va.w : va.y); af[3] = __float_as_uint(q ? vb.w : vb.y);
}
// block p of an accumulator as the two float4 (rows g, g+8) the thread owns
__device__ __forceinline__ float4 rb_blk_a(const float (&acc)[8][4], int p) { return make_float4(acc[2 * p][0], acc[2 * p][1], acc[2 * p + 1][0], acc[2 * p + 1][1]); }
__device__ __forceinline__ float4 rb_blk_b(const float (&acc)[8][4], int p) { return make_float4(acc[2 * p][2], acc[2 * p][3], acc[2 * p + 1][2], acc[2 * p + 1][3]); }
__device__ __forceinline__ void rb_blk_add(float (&acc)[8][4], int p, const float4& va, const float4& vb) {
    acc[2 * p][0] += va.x; acc[2 * p][1] += va.y; acc[2 * p + 1][0] += va.z; acc[2 * p + 1][1] += va.w;
    acc[2 * p][2] += vb.x; acc[2 * p][3] += vb.y; acc[2 * p + 1][2] += vb.z; acc[2 * p + 1][3] += vb.w;
}

__global__ void __launch_bounds__(RB_WARPS * 32, 1) res_bwd_fused_kernel(const ResBwdArgs a) {
    extern __shared__ __align__(16) float rb_smem[];
    constexpr int H = RB_H;
    float* ru_s = rb_smem;                 // [32 k-pairs][128]
    float* rg_s = ru_s + 32 * RB_ROW;      // [64 k-pairs][128]: pairs 0..31 multiply DR[0:H], pairs 32..63 DR[H:2H]
    for (int idx = threadIdx.x; idx < 192 * 16; idx += blockDim.x) {
        const int o = idx >> 4, j4 = (idx & 15) * 4;
        const float4 v = o < 64 ? ld4(a.RuH + o * H + j4) : ld4(a.RgH + (o - 64) * H + j4);
        const int kk = o < 64 ? o : o - 64, kpair = kk >> 1, rot = rb_rot(kpair);
        float* row = (o < 64 ? ru_s : rg_s) + kpair * RB_ROW + (kk & 1);
        row[(2 * j4 + rot) & (RB_ROW - 1)] = v.x;
        row[(2 * j4 + 2 + rot) & (RB_ROW - 1)] = v.y;
        row[(2 * j4 + 4 + rot) & (RB_ROW - 1)] = v.z;
        row[(2 * j4 + 6 + rot) & (RB_ROW - 1)] = v.w;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float gmix = __ldg(a.mix_t);
    const int g = lane >> 2, tig = lane & 3;
    const int cg = 8 * (g >> 1) + 2 * (g & 1) + 4 * (tig & 1) + 16 * (tig >> 1);  // B-fragment column term + rot(tig)
    const long long ntiles = (a.rows + 15) / 16;
    float dmix_part = 0.f;
    for (long long tile = (long long)blockIdx.x * RB_WARPS + warp; tile < ntiles; tile += (long long)gridDim.x * RB_WARPS) {
        const long long ra = tile * 16 + g, rb = ra + 8;
        const bool va = ra < a.rows, vb = rb < a.rows;
        float acc1[8][4], acc2[8][4];   // dzh2 ; dh1
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int x = 0; x < 4; ++x) { acc1[nt][x] = 0.f; acc2[nt][x] = 0.f; }
        // ---- head (B0), block by block; its outputs feed product 1 (da3) and the DR[H:2H] half of product 2 (dar) ----
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int c = 16 * p + 4 * tig;
            float4 dya = f4(0.f), h1a = f4(0.f), r2a = f4(0.f), hc2a = f4(0.f), dyb = f4(0.f), h1b = f4(0.f), r2b = f4(0.f), hc2b = f4(0.f);
            if (va) { const long long o = ra * H + c; dya = ld4(a.dY + o) + ld4(a.carry + o); h1a = ld4(a.H1 + o); r2a = ld4(a.R2 + o); hc2a = ld4(a.HC2 + o); }
            if (vb) { const long long o = rb * H + c; dyb = ld4(a.dY + o) + ld4(a.carry + o); h1b = ld4(a.H1 + o); r2b = ld4(a.R2 + o); hc2b = ld4(a.HC2 + o); }
            float4 da3a, dara, da3b, darb;
            {
                const float4 res = r2a * h1a + one_minus(r2a) * hc2a, pr = dya * (h1a - res), dres = (1.f - gmix) * dya;
                dmix_part += (pr.x + pr.y) + (pr.z + pr.w);
                da3a = dres * one_minus(r2a) * one_minus(hc2a * hc2a);
                dara = dres * (h1a - hc2a) * r2a * one_minus(r2a);
                const float4 dh1 = gmix * dya + dres * r2a;
                acc2[2 * p][0] += dh1.x; acc2[2 * p][1] += dh1.y; acc2[2 * p + 1][0] += dh1.z; acc2[2 * p + 1][1] += dh1.w;
                if (va) { st4(a.DR + ra * 3 * H + 2 * H + c, da3a); st4(a.DR + ra * 3 * H + H + c, dara); }
            }
            {
                const float4 res = r2b * h1b + one_minus(r2b) * hc2b, pr = dyb * (h1b - res), dres = (1.f - gmix) * dyb;
                dmix_part += (pr.x + pr.y) + (pr.z + pr.w);
                da3b = dres * one_minus(r2b) * one_minus(hc2b * hc2b);
                darb = dres * (h1b - hc2b) * r2b * one_minus(r2b);
                const float4 dh1 = gmix * dyb + dres * r2b;
                acc2[2 * p][2] += dh1.x; acc2[2 * p][3] += dh1.y; acc2[2 * p + 1][2] += dh1.z; acc2[2 * p + 1][3] += dh1.w;
                if (vb) { st4(a.DR + rb * 3 * H + 2 * H + c, da3b); st4(a.DR + rb * 3 * H + H + c, darb); }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int kpair = 8 * p + 2 * tig + q;
                uint32_t af[4];
                rb_afrag(af, da3a, da3b, q);
                rb_kstep(acc1, af, ru_s + kpair * RB_ROW, cg);
                rb_afrag(af, dara, darb, q);
                rb_kstep(acc2, af, rg_s + (32 + kpair) * RB_ROW, cg);
            }
        }
        // ---- epilogue of product 1 (B1): DR[0:H], dh1 += dzh2 z2, and the DR[0:H] half of product 2 ----
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int c = 16 * p + 4 * tig;
            float4 z2a = f4(0.f), h1a = f4(0.f), z2b = f4(0.f), h1b = f4(0.f);
            if (va) { z2a = ld4(a.Z2 + ra * H + c); h1a = ld4(a.H1 + ra * H + c); }
            if (vb) { z2b = ld4(a.Z2 + rb * H + c); h1b = ld4(a.H1 + rb * H + c); }
            const float4 dza = rb_blk_a(acc1, p), dzb = rb_blk_b(acc1, p);
            const float4 daza = dza * h1a * z2a * one_minus(z2a), dazb = dzb * h1b * z2b * one_minus(z2b);
            rb_blk_add(acc2, p, dza * z2a, dzb * z2b);
            if (va) st4(a.DR + ra * 3 * H + c, daza);
            if (vb) st4(a.DR + rb * 3 * H + c, dazb);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t af[4];
                rb_afrag(af, daza, dazb, q);
                rb_kstep(acc2, af, rg_s + (8 * p + 2 * tig + q) * RB_ROW, cg);
            }
        }
        // ---- epilogue of product 2 (B2): DHD and the r / candidate pre-activation gradients of the main cell ----
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int c = 16 * p + 4 * tig;
            if (va) {
                const float4 r = ld4(a.R + ra * H + c), hc = ld4(a.HC + ra * H + c), hp = ld4(a.Hprev + ra * H + c);
                const float4 d = rb_blk_a(acc2, p);
                const float4 gu = d * one_minus(r) * one_minus(hc * hc), gr = d * (hp - hc) * r * one_minus(r);
                st4(a.DHD + ra * H + c, d * r);
                st4(a.DG + ra * 3 * H + 2 * H + c, gu);
                st4(a.DG + ra * 3 * H + H + c, gr);
                if (a.DG16) { st4_bf16(a.DG16 + ra * 3 * H + 2 * H + c, gu); st4_bf16(a.DG16 + ra * 3 * H + H + c, gr); }
            }
            if (vb) {
                const float4 r = ld4(a.R + rb * H + c), hc = ld4(a.HC + rb * H + c), hp = ld4(a.Hprev + rb * H + c);
                const float4 d = rb_blk_b(acc2, p);
                const float4 gu = d * one_minus(r) * one_minus(hc * hc), gr = d * (hp - hc) * r * one_minus(r);
                st4(a.DHD + rb * H + c, d * r);
                st4(a.DG + rb * 3 * H + 2 * H + c, gu);
                st4(a.DG + rb * 3 * H + H + c, gr);
                if (a.DG16) { st4_bf16(a.DG16 + rb * 3 * H + 2 * H + c, gu); st4_bf16(a.DG16 + rb * 3 * H + H + c, gr); }
            }
        }
    }
    // dmix_t: one atomic per warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dmix_part += __shfl_xor_sync(0xffffffffu, dmix_part, off);
    if (lane == 0 && dmix_part != 0.f) atomicAdd(a.dmix_t, dmix_part);
}

inline bool res_bwd_fused_ok(const ResBwdArgs& a, int H) {
    return H == RB_H && aligned16(a.dY) && aligned16(a.carry) && aligned16(a.H1) && aligned16(a.R2) && aligned16(a.HC2) &&
           aligned16(a.Z2) && aligned16(a.Hprev) && aligned16(a.R) && aligned16(a.HC) && aligned16(a.RuH) && aligned16(a.RgH) &&
           aligned16(a.DR) && aligned16(a.DG) && aligned16(a.DHD) && (!a.DG16 || aligned16(a.DG16));
}

inline cudaError_t launch_res_bwd_fused(const ResBwdArgs& a, cudaStream_t st) {
    static bool configured[64] = {};   // per device
    const size_t smem = sizeof(float) * RB_SMEM_FLOATS;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(res_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const long long ntiles = (a.rows + 15) / 16;
    long long blocks = (ntiles + RB_WARPS - 1) / RB_WARPS;
    if (blocks > sm_count()) blocks = sm_count();
    if (blocks < 1) blocks = 1;
    res_bwd_fused_kernel<<<(unsigned)blocks, RB_WARPS * 32, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace matgcn
