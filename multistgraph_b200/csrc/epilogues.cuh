// Epilogue functors shared by the SIMT engine (gemm_simt.cuh) and the tensor-core engines (gemm_tc*.cuh):
// all the elementwise GRU algebra of the path (forward gating, reverse-step chain rule) lives here.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include "gemm_simt.cuh"

namespace matgcn {

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
// fast-mode activations (SFU approximations, relative error ~2^-11: the same order as the TF32 products they follow)
__device__ __forceinline__ float tanh_fast(float v) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// sigma(v) = (1 + tanh(v/2)) / 2: one SFU operation instead of exp + reciprocal (the SFU pipe bounds the epilogues)
__device__ __forceinline__ float sigmoid_fast(float v) { return fmaf(tanh_fast(0.5f * v), 0.5f, 0.5f); }

// ------------------------------------------------------------------------------------------
// epilogues
//
// Every epilogue is split in two phases so that the tensor-core kernel (few epilogue warps, no
// thread-level parallelism to hide latency) can issue the global loads of a whole batch of output
// elements before consuming any of them:
//   EpiIn in = epi.load(z1, z2, row, col);      // every global read the element needs
//   epi.store(z1, z2, row, col, acc, in);       // arithmetic + global writes
// operator() = store(load()) is what the SIMT kernel calls.  load4/store4 are the same for four
// consecutive columns (col % 4 == 0) with 16-byte accesses; vec_ok() tells the host whether every
// pointer/pitch involved allows that.
// ------------------------------------------------------------------------------------------
#define EPI_CALL_OPERATOR                                                                              \
    __device__ __forceinline__ void operator()(int z1, int z2, int row, int col, float acc) const {    \
        store(z1, z2, row, col, acc, load(z1, z2, row, col));                                          \
    }
// L2 prefetch of the 128-byte line holding *p (tensor-core engine: a spare warp runs one tile ahead of the epilogue
// warps and pulls their global inputs into L2, see gemm_tc.cuh)
__device__ __forceinline__ void pf_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 operator+(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(const float4& a, const float4& b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float a, const float4& b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ float4 one_minus(const float4& a) { return make_float4(1.f - a.x, 1.f - a.y, 1.f - a.z, 1.f - a.w); }
__device__ __forceinline__ float4 sigmoid4(const float4& a, int fast) {
    return fast ? make_float4(sigmoid_fast(a.x), sigmoid_fast(a.y), sigmoid_fast(a.z), sigmoid_fast(a.w))
                : make_float4(sigmoidf_(a.x), sigmoidf_(a.y), sigmoidf_(a.z), sigmoidf_(a.w));
}
__device__ __forceinline__ float4 tanh4(const float4& a, int fast) {
    return fast ? make_float4(tanh_fast(a.x), tanh_fast(a.y), tanh_fast(a.z), tanh_fast(a.w))
                : make_float4(tanhf(a.x), tanhf(a.y), tanhf(a.z), tanhf(a.w));
}

// C[z1*s1 + z2*s2 + row*ldc + col] (=|+=) acc * scale[(col / scale_div)] + bias[z1*bias_s1 + col] + add[...]
struct EpiStore {
    static constexpr int kBatch = 4;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 8;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    float* C;
    long long s1, s2;
    int ldc;
    const float* bias;   // may be null
    long long bias_s1;
    const float* scale;  // may be null: per column-group scale (view weights)
    int scale_div;
    int accumulate;      // 1: C += value
    const float* add;    // may be null: extra addend with C's indexing (offsets add_s1/add_s2, ld add_ld)
    long long add_s1, add_s2;
    int add_ld;
    __nv_bfloat16* C16;  // may be null: bf16 twin of C (same indexing), operand of a later bf16 contraction
    float* D2;           // may be null: second copy of the blocks z2 in [d2_lo, d2_hi), at D2 + z1*s1 + (z2-d2_lo)*s2 + row*ldc + col
    int d2_lo, d2_hi;
    bool vec_ok() const {
        return aligned16(C) && (!D2 || aligned16(D2)) && (!C16 || aligned16(C16)) && !(s1 & 3) && !(s2 & 3) && !(ldc & 3) && (!bias || (aligned16(bias) && !(bias_s1 & 3))) &&
               (!scale || !(scale_div & 3)) && (!add || (aligned16(add) && !(add_s1 & 3) && !(add_s2 & 3) && !(add_ld & 3)));
    }
    __device__ __forceinline__ EpiIn load(int z1, int z2, int row, int col) const {
        EpiIn in;
        in.a = scale ? __ldg(scale + col / scale_div) : 1.f;
        in.b = bias ? __ldg(bias + z1 * bias_s1 + col) : 0.f;
        in.c = add ? add[z1 * add_s1 + z2 * add_s2 + (long long)row * add_ld + col] : 0.f;
        in.d = accumulate ? C[z1 * s1 + z2 * s2 + (long long)row * ldc + col] : 0.f;
        return in;
    }
    __device__ __forceinline__ void store(int z1, int z2, int row, int col, float acc, const EpiIn& in) const {
        float v = acc;
        if (scale) v *= in.a;
        if (bias) v += in.b;
        if (add) v += in.c;
        if (accumulate) v += in.d;
        C[z1 * s1 + z2 * s2 + (long long)row * ldc + col] = v;
        if (C16) C16[z1 * s1 + z2 * s2 + (long long)row * ldc + col] = __float2bfloat16_rn(v);
        if (D2 && z2 >= d2_lo && z2 < d2_hi) D2[z1 * s1 + (z2 - d2_lo) * s2 + (long long)row * ldc + col] = v;
    }
    __device__ __forceinline__ EpiIn4 load4(int z1, int z2, int row, int col) const {
        EpiIn4 in;
        in.a = f4(scale ? __ldg(scale + col / scale_div) : 1.f);
        in.b = bias ? ld4(bias + z1 * bias_s1 + col) : f4(0.f);
        in.c = add ? ld4(add + z1 * add_s1 + z2 * add_s2 + (long long)row * add_ld + col) : f4(0.f);
        in.d = accumulate ? ld4(C + z1 * s1 + z2 * s2 + (long long)row * ldc + col) : f4(0.f);
        return in;
    }
    __device__ __forceinline__ void store4(int z1, int z2, int row, int col, const float4& acc, const EpiIn4& in) const {
        float4 v = acc;
        if (scale) v = v * in.a;
        if (bias) v = v + in.b;
        if (add) v = v + in.c;
        if (accumulate) v = v + in.d;
        st4(C + z1 * s1 + z2 * s2 + (long long)row * ldc + col, v);
        if (C16) st4_bf16(C16 + z1 * s1 + z2 * s2 + (long long)row * ldc + col, v);
        if (D2 && z2 >= d2_lo && z2 < d2_hi) st4(D2 + z1 * s1 + (z2 - d2_lo) * s2 + (long long)row * ldc + col, v);
    }
    // cursor form (tensor-core epilogue only: always the fast activations, which keeps the unrolled epilogue code small -
    // these short kernels start with a cold instruction cache): the position is computed once per 32-column chunk and
    // bumped row by row
    struct Cur { long long off, aoff, off2; float4 bias; float scale; };
    __device__ __forceinline__ Cur begin4(int z1, int z2, int row, int col) const {
        Cur c;
        c.off = z1 * s1 + z2 * s2 + (long long)row * ldc + col;
        c.off2 = (D2 && z2 >= d2_lo && z2 < d2_hi) ? z1 * s1 + (z2 - d2_lo) * s2 + (long long)row * ldc + col : -1;
        c.aoff = add ? z1 * add_s1 + z2 * add_s2 + (long long)row * add_ld + col : 0;
        c.bias = bias ? ld4(bias + z1 * bias_s1 + col) : f4(0.f);
        c.scale = scale ? __ldg(scale + col / scale_div) : 1.f;
        return c;
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const {
        c.off += (long long)rows * ldc; c.aoff += (long long)rows * add_ld;
        if (c.off2 >= 0) c.off2 += (long long)rows * ldc;
    }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.c = add ? ld4(add + c.aoff) : f4(0.f);
        in.d = accumulate ? ld4(C + c.off) : f4(0.f);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const {
        if (add) pf_l2(add + c.aoff);
        if (accumulate) pf_l2(C + c.off);
    }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        float4 v = c.scale * acc + c.bias;
        if (add) v = v + in.c;
        if (accumulate) v = v + in.d;
        st4(C + c.off, v);
        if (C16) st4_bf16(C16 + c.off, v);
        if (c.off2 >= 0) st4(D2 + c.off2, v);
    }
    EPI_CALL_OPERATOR
};
inline EpiStore epi_store(float* C, long long s1, long long s2, int ldc) {
    EpiStore e;
    memset(&e, 0, sizeof(e));
    e.C = C; e.s1 = s1; e.s2 = s2; e.ldc = ldc; e.scale_div = 1;
    return e;
}

// EpiStore with its feature set fixed at compile time (F = OR of the ES_* bits in use) for the tensor-core engine's vectorised
// epilogue: the run-time tests of the generic functor (is there a bias / scale / addend / twin / second copy, reloaded from the
// constant bank around every 16-byte access) made that epilogue ~130 instructions per access and the time-batched launches
// with large outputs instruction-bound (profiles/r2g_rx_epilogue_ncu.txt); with the unused paths compiled out it is ~25.
enum { ES_BIAS = 1, ES_ADD = 2, ES_ACC = 4, ES_C16 = 8, ES_SCALE = 16 };
template <int F>
struct EpiStoreT {
    static constexpr int kBatch = 4;
    static constexpr int kPipe = 8;
    static constexpr bool kVecOnly = true;   // no scalar-epilogue kernel is built for it (the caller falls back to EpiStore)
    float* C;
    long long s1, s2;
    int ldc;
    const float* bias; long long bias_s1;
    const float* scale; int scale_div;
    const float* add; long long add_s1, add_s2; int add_ld;
    __nv_bfloat16* C16;
    EpiStoreT() = default;
    explicit EpiStoreT(const EpiStore& e)
        : C(e.C), s1(e.s1), s2(e.s2), ldc(e.ldc), bias(e.bias), bias_s1(e.bias_s1), scale(e.scale), scale_div(e.scale_div), add(e.add),
          add_s1(e.add_s1), add_s2(e.add_s2), add_ld(e.add_ld), C16(e.C16) {}
    static bool matches(const EpiStore& e) {
        return !e.D2 && !!e.bias == !!(F & ES_BIAS) && !!e.add == !!(F & ES_ADD) && !!e.accumulate == !!(F & ES_ACC) &&
               !!e.C16 == !!(F & ES_C16) && !!e.scale == !!(F & ES_SCALE);
    }
    bool vec_ok() const {
        return aligned16(C) && (!(F & ES_C16) || aligned16(C16)) && !(s1 & 3) && !(s2 & 3) && !(ldc & 3) &&
               (!(F & ES_BIAS) || (aligned16(bias) && !(bias_s1 & 3))) && (!(F & ES_SCALE) || !(scale_div & 3)) &&
               (!(F & ES_ADD) || (aligned16(add) && !(add_s1 & 3) && !(add_s2 & 3) && !(add_ld & 3)));
    }
    __device__ __forceinline__ EpiIn load(int z1, int z2, int row, int col) const {
        EpiIn in;
        in.a = (F & ES_SCALE) ? __ldg(scale + col / scale_div) : 1.f;
        in.b = (F & ES_BIAS) ? __ldg(bias + z1 * bias_s1 + col) : 0.f;
        in.c = (F & ES_ADD) ? add[z1 * add_s1 + z2 * add_s2 + (long long)row * add_ld + col] : 0.f;
        in.d = (F & ES_ACC) ? C[z1 * s1 + z2 * s2 + (long long)row * ldc + col] : 0.f;
        return in;
    }
    __device__ __forceinline__ void store(int z1, int z2, int row, int col, float acc, const EpiIn& in) const {
        const float v = acc * in.a + in.b + in.c + in.d;
        C[z1 * s1 + z2 * s2 + (long long)row * ldc + col] = v;
        if (F & ES_C16) C16[z1 * s1 + z2 * s2 + (long long)row * ldc + col] = __float2bfloat16_rn(v);
    }
    __device__ __forceinline__ EpiIn4 load4(int z1, int z2, int row, int col) const {
        EpiIn4 in;
        in.a = f4((F & ES_SCALE) ? __ldg(scale + col / scale_div) : 1.f);
        in.b = (F & ES_BIAS) ? ld4(bias + z1 * bias_s1 + col) : f4(0.f);
        in.c = (F & ES_ADD) ? ld4(add + z1 * add_s1 + z2 * add_s2 + (long long)row * add_ld + col) : f4(0.f);
        in.d = (F & ES_ACC) ? ld4(C + z1 * s1 + z2 * s2 + (long long)row * ldc + col) : f4(0.f);
        return in;
    }
    __device__ __forceinline__ void store4(int z1, int z2, int row, int col, const float4& acc, const EpiIn4& in) const {
        const float4 v = in.a * acc + in.b + in.c + in.d;
        st4(C + z1 * s1 + z2 * s2 + (long long)row * ldc + col, v);
        if (F & ES_C16) st4_bf16(C16 + z1 * s1 + z2 * s2 + (long long)row * ldc + col, v);
    }
    struct Cur { long long off, aoff; float4 bias; float scale; };
    __device__ __forceinline__ Cur begin4(int z1, int z2, int row, int col) const {
        Cur c;
        c.off = z1 * s1 + z2 * s2 + (long long)row * ldc + col;
        c.aoff = (F & ES_ADD) ? z1 * add_s1 + z2 * add_s2 + (long long)row * add_ld + col : 0;
        c.bias = (F & ES_BIAS) ? ld4(bias + z1 * bias_s1 + col) : f4(0.f);
        c.scale = (F & ES_SCALE) ? __ldg(scale + col / scale_div) : 1.f;
        return c;
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const {
        c.off += (long long)rows * ldc;
        if (F & ES_ADD) c.aoff += (long long)rows * add_ld;
    }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        if (F & ES_ADD) in.c = ld4(add + c.aoff);
        if (F & ES_ACC) in.d = ld4(C + c.off);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const {
        if (F & ES_ADD) pf_l2(add + c.aoff);
        if (F & ES_ACC) pf_l2(C + c.off);
    }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        float4 v = acc;
        if (F & ES_SCALE) v = c.scale * v;
        if (F & ES_BIAS) v = v + c.bias;
        if (F & ES_ADD) v = v + in.c;
        if (F & ES_ACC) v = v + in.d;
        st4(C + c.off, v);
        if (F & ES_C16) st4_bf16(C16 + c.off, v);
    }
    EPI_CALL_OPERATOR
};
template <class E, class = void> struct EpiVecOnly { static constexpr bool value = false; };
template <class E> struct EpiVecOnly<E, decltype((void)E::kVecOnly)> { static constexpr bool value = E::kVecOnly; };

// Lean form of EpiStore for the contractions of the recurrence (support propagation, the per-node products of the
// reverse step): C = acc, optionally with a bf16 twin and a second copy of the blocks z2 in [d2_lo, d2_hi).
// (EpiStore's optional bias / scale / addend / accumulate paths cost code even when unused, and every launch of these
// short kernels starts with a cold instruction cache.)
struct EpiPlain {
    static constexpr int kBatch = 8;
    static constexpr int kPipe = 8;   // no global reads; 8 = that many accesses unrolled per round (independent LDS -> STG chains)
    float* C;
    long long s1, s2;
    int ldc;
    __nv_bfloat16* C16;  // may be null
    float* D2;           // may be null: see EpiStore
    int d2_lo, d2_hi;
    __nv_bfloat16* D2h;  // may be null: the same second copy as bf16 (same block range and indexing as D2)
    int c_z2_hi;         // the fp32 copy C is written only for blocks z2 < c_z2_hi (bf16 mode: slots that are consumed as
                         // bf16 twins only skip their fp32 store; 0 = never, INT_MAX = always)
    bool vec_ok() const { return aligned16(C) && (!C16 || aligned16(C16)) && (!D2 || aligned16(D2)) && (!D2h || aligned16(D2h)) && !(s1 & 3) && !(s2 & 3) && !(ldc & 3); }
    __device__ __forceinline__ EpiIn load(int, int, int, int) const { return EpiIn{}; }
    __device__ __forceinline__ void store(int z1, int z2, int row, int col, float acc, const EpiIn&) const {
        const long long o = z1 * s1 + z2 * s2 + (long long)row * ldc + col;
        if (z2 < c_z2_hi) C[o] = acc;
        if (C16) C16[o] = __float2bfloat16_rn(acc);
        if (D2 && z2 >= d2_lo && z2 < d2_hi) D2[o - (long long)d2_lo * s2] = acc;
        if (D2h && z2 >= d2_lo && z2 < d2_hi) D2h[o - (long long)d2_lo * s2] = __float2bfloat16_rn(acc);
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int, int) const { return EpiIn4{}; }
    __device__ __forceinline__ void store4(int z1, int z2, int row, int col, const float4& acc, const EpiIn4&) const {
        const long long o = z1 * s1 + z2 * s2 + (long long)row * ldc + col;
        if (z2 < c_z2_hi) st4(C + o, acc);
        if (C16) st4_bf16(C16 + o, acc);
        if (D2 && z2 >= d2_lo && z2 < d2_hi) st4(D2 + o - (long long)d2_lo * s2, acc);
        if (D2h && z2 >= d2_lo && z2 < d2_hi) st4_bf16(D2h + o - (long long)d2_lo * s2, acc);
    }
    struct Cur { long long off; int dup; int main; };
    __device__ __forceinline__ Cur begin4(int z1, int z2, int row, int col) const {
        return Cur{z1 * s1 + z2 * s2 + (long long)row * ldc + col, ((D2 || D2h) && z2 >= d2_lo && z2 < d2_hi) ? 1 : 0, z2 < c_z2_hi ? 1 : 0};
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.off += (long long)rows * ldc; }
    __device__ __forceinline__ EpiIn4 load4(const Cur&) const { return EpiIn4{}; }
    __device__ __forceinline__ void prefetch4(const Cur&) const {}
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4&) const {
        if (c.main) st4(C + c.off, acc);
        if (C16) st4_bf16(C16 + c.off, acc);
        if (c.dup) {
            if (D2) st4(D2 + c.off - (long long)d2_lo * s2, acc);
            if (D2h) st4_bf16(D2h + c.off - (long long)d2_lo * s2, acc);
        }
    }
    EPI_CALL_OPERATOR
};
inline EpiPlain epi_plain(float* C, long long s1, long long s2, int ldc) {
    EpiPlain e;
    memset(&e, 0, sizeof(e));
    e.C = C; e.s1 = s1; e.s2 = s2; e.ldc = ldc; e.c_z2_hi = 0x7fffffff;
    return e;
}

// Plain store whose columns are BLOCKS of `blk` columns living `sblk` floats apart (block j = col / blk):
//   C[j*sblk + z1*s1 + z2*s2 + row*ld + col % blk] = acc,   bf16 twin for the blocks j >= c16_lo.
// One launch with N = (number of blocks) * blk then replaces one launch per block when the blocks share the A operand
// (the per-support input gradients DPX[t, k, n] = DG[t, n] WX[n, k]^T: DG is streamed once instead of K times).
// blk % 32 == 0: a 32-column chunk of the tensor-core epilogue never straddles two blocks.
struct EpiBlocks {
    static constexpr int kBatch = 8;
    static constexpr int kPipe = 8;
    float* C;
    long long s1, s2, sblk;
    int ld, blk;
    __nv_bfloat16* C16;  // may be null
    int c16_lo;          // column blocks >= c16_lo get the bf16 twin ...
    int c32_hi;          // ... and only blocks < c32_hi the fp32 store (blocks whose consumers all read the twin skip it)
    bool vec_ok() const { return aligned16(C) && (!C16 || aligned16(C16)) && !(s1 & 3) && !(s2 & 3) && !(sblk & 3) && !(ld & 3) && !(blk & 31); }
    __device__ __forceinline__ long long off(int z1, int z2, int row, int col, int& j) const {
        j = col / blk;
        return j * sblk + z1 * s1 + z2 * s2 + (long long)row * ld + (col - j * blk);
    }
    __device__ __forceinline__ EpiIn load(int, int, int, int) const { return EpiIn{}; }
    __device__ __forceinline__ void store(int z1, int z2, int row, int col, float acc, const EpiIn&) const {
        int j;
        const long long o = off(z1, z2, row, col, j);
        if (j < c32_hi) C[o] = acc;
        if (C16 && j >= c16_lo) C16[o] = __float2bfloat16_rn(acc);
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int, int) const { return EpiIn4{}; }
    __device__ __forceinline__ void store4(int z1, int z2, int row, int col, const float4& acc, const EpiIn4&) const {
        int j;
        const long long o = off(z1, z2, row, col, j);
        if (j < c32_hi) st4(C + o, acc);
        if (C16 && j >= c16_lo) st4_bf16(C16 + o, acc);
    }
    struct Cur { long long off; int twin; };
    __device__ __forceinline__ Cur begin4(int z1, int z2, int row, int col) const {
        int j;
        const long long o = off(z1, z2, row, col, j);
        return Cur{o, ((C16 && j >= c16_lo) ? 1 : 0) | (j < c32_hi ? 2 : 0)};
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.off += (long long)rows * ld; }
    __device__ __forceinline__ EpiIn4 load4(const Cur&) const { return EpiIn4{}; }
    __device__ __forceinline__ void prefetch4(const Cur&) const {}
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4&) const {
        if (c.twin & 2) st4(C + c.off, acc);
        if (c.twin & 1) st4_bf16(C16 + c.off, acc);
    }
    EPI_CALL_OPERATOR
};

struct EpiAtomic {  // split-K partial sums into a zeroed C
    static constexpr int kBatch = 8;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 1;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    float* C;
    long long s1, s2;
    int ldc;
    bool vec_ok() const { return true; }
    __device__ __forceinline__ EpiIn load(int, int, int, int) const { return EpiIn{}; }
    __device__ __forceinline__ void store(int z1, int z2, int row, int col, float acc, const EpiIn&) const {
        atomicAdd(C + z1 * s1 + z2 * s2 + (long long)row * ldc + col, acc);
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int, int) const { return EpiIn4{}; }
    __device__ __forceinline__ void store4(int z1, int z2, int row, int col, const float4& acc, const EpiIn4&) const {
        float* d = C + z1 * s1 + z2 * s2 + (long long)row * ldc + col;
        atomicAdd(d, acc.x); atomicAdd(d + 1, acc.y); atomicAdd(d + 2, acc.z); atomicAdd(d + 3, acc.w);
    }
    struct Cur { long long off; };
    __device__ __forceinline__ Cur begin4(int z1, int z2, int row, int col) const { return Cur{z1 * s1 + z2 * s2 + (long long)row * ldc + col}; }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.off += (long long)rows * ldc; }
    __device__ __forceinline__ EpiIn4 load4(const Cur&) const { return EpiIn4{}; }
    __device__ __forceinline__ void prefetch4(const Cur&) const {}
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4&) const {
        float* d = C + c.off;
        atomicAdd(d, acc.x); atomicAdd(d + 1, acc.y); atomicAdd(d + 2, acc.z); atomicAdd(d + 3, acc.w);
    }
    EPI_CALL_OPERATOR
};

// Forward step epilogues.  g = z1*rows_per_z + row indexes (node, batch) pairs; activations
// are [N*B, H] blocks, pre-activation inputs [N*B, 3H] blocks.
struct EpiGate {  // sigma(acc + GX[:, 0:2H]) -> z (cols < H): Z, ZH = z*h ; r (cols >= H): R
    static constexpr int kBatch = 8;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 8;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    const float* GX; const float* Hprev; float* Z; float* R; float* ZH;
    int rows_per_z, H, fast;
    __nv_bfloat16* ZH16;  // may be null: bf16 twin of ZH
    bool vec_ok() const { return !(H & 3) && aligned16(GX) && aligned16(Hprev) && aligned16(Z) && aligned16(R) && aligned16(ZH); }
    __device__ __forceinline__ EpiIn load(int z1, int, int row, int col) const {
        const long long g = (long long)z1 * rows_per_z + row;
        EpiIn in;
        in.a = GX[g * 3 * H + col];
        in.b = col < H ? Hprev[g * H + col] : 0.f;
        return in;
    }
    __device__ __forceinline__ void store(int z1, int, int row, int col, float acc, const EpiIn& in) const {
        const long long g = (long long)z1 * rows_per_z + row;
        const float s = fast ? sigmoid_fast(acc + in.a) : sigmoidf_(acc + in.a);
        if (col < H) {
            Z[g * H + col] = s;
            ZH[g * H + col] = s * in.b;
            if (ZH16) ZH16[g * H + col] = __float2bfloat16_rn(s * in.b);
        } else {
            R[g * H + col - H] = s;
        }
    }
    __device__ __forceinline__ EpiIn4 load4(int z1, int, int row, int col) const {
        const long long g = (long long)z1 * rows_per_z + row;
        EpiIn4 in;
        in.a = ld4(GX + g * 3 * H + col);
        in.b = col < H ? ld4(Hprev + g * H + col) : f4(0.f);
        return in;
    }
    __device__ __forceinline__ void store4(int z1, int, int row, int col, const float4& acc, const EpiIn4& in) const {
        const long long g = (long long)z1 * rows_per_z + row;
        const float4 s = sigmoid4(acc + in.a, fast);
        if (col < H) {
            st4(Z + g * H + col, s);
            st4(ZH + g * H + col, s * in.b);
            if (ZH16) st4_bf16(ZH16 + g * H + col, s * in.b);
        } else {
            st4(R + g * H + col - H, s);
        }
    }
    // i: offset into the [rows, H] blocks (column already reduced by H for the r half), j: offset into GX [rows, 3H]
    struct Cur { long long i, j; int is_z; };
    __device__ __forceinline__ Cur begin4(int z1, int, int row, int col) const {
        const long long g = (long long)z1 * rows_per_z + row;
        Cur c;
        c.is_z = col < H;
        c.i = g * H + (c.is_z ? col : col - H);
        c.j = g * 3 * H + col;
        return c;
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.i += (long long)rows * H; c.j += (long long)rows * 3 * H; }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.a = ld4(GX + c.j);
        in.b = c.is_z ? ld4(Hprev + c.i) : f4(0.f);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const {
        pf_l2(GX + c.j);
        if (c.is_z) pf_l2(Hprev + c.i);
    }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        const float4 s = sigmoid4(acc + in.a, 1);
        if (c.is_z) {
            st4(Z + c.i, s);
            st4(ZH + c.i, s * in.b);
            if (ZH16) st4_bf16(ZH16 + c.i, s * in.b);
        } else {
            st4(R + c.i, s);
        }
    }
    EPI_CALL_OPERATOR
};
struct EpiCand {  // hc = tanh(acc + GX[:, 2H:3H]); h1 = r*h + (1-r)*hc
    static constexpr int kBatch = 4;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 8;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    const float* GX; const float* Hprev; const float* R; float* HC; float* H1;
    int rows_per_z, H, fast;
    bool vec_ok() const { return !(H & 3) && aligned16(GX) && aligned16(Hprev) && aligned16(R) && aligned16(HC) && aligned16(H1); }
    __device__ __forceinline__ EpiIn load(int z1, int, int row, int col) const {
        const long long g = (long long)z1 * rows_per_z + row;
        EpiIn in;
        in.a = GX[g * 3 * H + 2 * H + col];
        in.b = R[g * H + col];
        in.c = Hprev[g * H + col];
        return in;
    }
    __device__ __forceinline__ void store(int z1, int, int row, int col, float acc, const EpiIn& in) const {
        const long long g = (long long)z1 * rows_per_z + row;
        const float hc = fast ? tanh_fast(acc + in.a) : tanhf(acc + in.a);
        HC[g * H + col] = hc;
        H1[g * H + col] = in.b * in.c + (1.f - in.b) * hc;
    }
    __device__ __forceinline__ EpiIn4 load4(int z1, int, int row, int col) const {
        const long long g = (long long)z1 * rows_per_z + row;
        EpiIn4 in;
        in.a = ld4(GX + g * 3 * H + 2 * H + col);
        in.b = ld4(R + g * H + col);
        in.c = ld4(Hprev + g * H + col);
        return in;
    }
    __device__ __forceinline__ void store4(int z1, int, int row, int col, const float4& acc, const EpiIn4& in) const {
        const long long g = (long long)z1 * rows_per_z + row;
        const float4 hc = tanh4(acc + in.a, fast);
        st4(HC + g * H + col, hc);
        st4(H1 + g * H + col, in.b * in.c + one_minus(in.b) * hc);
    }
    struct Cur { long long i, j; };
    __device__ __forceinline__ Cur begin4(int z1, int, int row, int col) const {
        const long long g = (long long)z1 * rows_per_z + row;
        return Cur{g * H + col, g * 3 * H + 2 * H + col};
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.i += (long long)rows * H; c.j += (long long)rows * 3 * H; }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.a = ld4(GX + c.j); in.b = ld4(R + c.i); in.c = ld4(Hprev + c.i);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const { pf_l2(GX + c.j); pf_l2(R + c.i); pf_l2(Hprev + c.i); }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        const float4 hc = tanh4(acc + in.a, 1);
        st4(HC + c.i, hc);
        st4(H1 + c.i, in.b * in.c + one_minus(in.b) * hc);
    }
    EPI_CALL_OPERATOR
};
struct EpiResCand {  // residual candidate + mix: y = g*h1 + (1-g)*(r2*h1 + (1-r2)*hc2)
    static constexpr int kBatch = 4;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 8;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    const float* RX; const float* H1; const float* R2; float* HC2; float* Y; const float* mix_t;
    int H, fast;
    __nv_bfloat16* Y16;  // may be null: bf16 twin of Y
    bool vec_ok() const { return !(H & 3) && aligned16(RX) && aligned16(H1) && aligned16(R2) && aligned16(HC2) && aligned16(Y); }
    __device__ __forceinline__ EpiIn load(int, int, int row, int col) const {
        const long long g = row;
        EpiIn in;
        in.a = RX[g * 3 * H + 2 * H + col];
        in.b = R2[g * H + col];
        in.c = H1[g * H + col];
        in.d = __ldg(mix_t);
        return in;
    }
    __device__ __forceinline__ void store(int, int, int row, int col, float acc, const EpiIn& in) const {
        const long long g = row;
        const float hc2 = fast ? tanh_fast(acc + in.a) : tanhf(acc + in.a);
        const float r2 = in.b, h1 = in.c, m = in.d;
        const float res = r2 * h1 + (1.f - r2) * hc2;
        HC2[g * H + col] = hc2;
        Y[g * H + col] = m * h1 + (1.f - m) * res;
        if (Y16) Y16[g * H + col] = __float2bfloat16_rn(m * h1 + (1.f - m) * res);
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int row, int col) const {
        const long long g = row;
        EpiIn4 in;
        in.a = ld4(RX + g * 3 * H + 2 * H + col);
        in.b = ld4(R2 + g * H + col);
        in.c = ld4(H1 + g * H + col);
        in.d = f4(__ldg(mix_t));
        return in;
    }
    __device__ __forceinline__ void store4(int, int, int row, int col, const float4& acc, const EpiIn4& in) const {
        const long long g = row;
        const float4 hc2 = tanh4(acc + in.a, fast);
        const float4 res = in.b * in.c + one_minus(in.b) * hc2;
        st4(HC2 + g * H + col, hc2);
        const float4 y = in.d * in.c + one_minus(in.d) * res;
        st4(Y + g * H + col, y);
        if (Y16) st4_bf16(Y16 + g * H + col, y);
    }
    struct Cur { long long i, j; float m; };
    __device__ __forceinline__ Cur begin4(int, int, int row, int col) const {
        return Cur{(long long)row * H + col, (long long)row * 3 * H + 2 * H + col, __ldg(mix_t)};
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.i += (long long)rows * H; c.j += (long long)rows * 3 * H; }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.a = ld4(RX + c.j); in.b = ld4(R2 + c.i); in.c = ld4(H1 + c.i);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const { pf_l2(RX + c.j); pf_l2(R2 + c.i); pf_l2(H1 + c.i); }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        const float4 hc2 = tanh4(acc + in.a, 1);
        const float4 res = in.b * in.c + one_minus(in.b) * hc2;
        st4(HC2 + c.i, hc2);
        const float4 y = c.m * in.c + (1.f - c.m) * res;
        st4(Y + c.i, y);
        if (Y16) st4_bf16(Y16 + c.i, y);
    }
    EPI_CALL_OPERATOR
};

// Fused tail of the forward step (tensor-core engine only, H = 64): the candidate contraction's epilogue
// (EpiCand) followed, for the same (node, batch) rows, by the whole residual GRU cell and the mix
// (EpiGate on RX / EpiResCand) - the two small [rows, 64] x [64, 192] products run as warp-level mma.sync
// inside the epilogue warps (gemm_tc.cuh: tc_epilogue_tile_candres), so one launch replaces three.
struct EpiCandRes {
    static constexpr bool kFusedRes = true;
    // candidate part (EpiCand)
    const float* GX; const float* Hprev; const float* R; float* HC; float* H1;
    int rows_per_z, H, fast;
    // residual cell + mix (EpiGate on RX, EpiResCand)
    const float* RX; float* Z2; float* R2; float* ZH2; float* HC2; float* Y; const float* mix_t;
    __nv_bfloat16* Y16;   // may be null: bf16 twin of Y
    const float* RgH;     // [2H, H] dense: Rgw[:, Cin:]
    const float* RuH;     // [H, H]  dense: Ruw[:, Cin:]
    bool vec_ok() const {
        return H == 64 && aligned16(GX) && aligned16(Hprev) && aligned16(R) && aligned16(HC) && aligned16(H1) && aligned16(RX) &&
               aligned16(Z2) && aligned16(R2) && aligned16(ZH2) && aligned16(HC2) && aligned16(Y) && aligned16(RgH) && aligned16(RuH) &&
               (!Y16 || aligned16(Y16));
    }
    // only the prefetch warp uses cursors here (col < 64: one 128-byte line per stream and 32-column chunk)
    struct Cur { long long i, j; };
    __device__ __forceinline__ Cur begin4(int z1, int, int row, int col) const {
        const long long g = (long long)z1 * rows_per_z + row;
        return Cur{g * H + col, g * 3 * H + col};
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const {
        pf_l2(GX + c.j + 2 * H); pf_l2(R + c.i); pf_l2(Hprev + c.i);
        pf_l2(RX + c.j); pf_l2(RX + c.j + H); pf_l2(RX + c.j + 2 * H);
    }
};
template <class E, class = void> struct IsFusedRes { static constexpr bool value = false; };
template <class E> struct IsFusedRes<E, decltype((void)E::kFusedRes)> { static constexpr bool value = true; };

// Backward step epilogues (notation of DESIGN.md section 3 / tests/host_mirror.py).
struct EpiB1 {  // acc = dzh2 ; DH1 += dzh2*z2 ; DR[0:H] = dzh2*h1*z2(1-z2) ; DR[H:2H] = dres*(h1-hc2)*r2(1-r2)
    static constexpr int kBatch = 2;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 4;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    float* DH1; float* DR; const float* DRES; const float* H1; const float* Z2; const float* R2; const float* HC2;
    int H;
    bool vec_ok() const {
        return !(H & 3) && aligned16(DH1) && aligned16(DR) && aligned16(DRES) && aligned16(H1) && aligned16(Z2) && aligned16(R2) && aligned16(HC2);
    }
    __device__ __forceinline__ EpiIn load(int, int, int row, int col) const {
        const long long i = (long long)row * H + col;
        EpiIn in;
        in.a = Z2[i]; in.b = R2[i]; in.c = H1[i]; in.d = DH1[i]; in.e = DRES[i]; in.f = HC2[i];
        return in;
    }
    __device__ __forceinline__ void store(int, int, int row, int col, float acc, const EpiIn& in) const {
        const long long i = (long long)row * H + col;
        const float z2 = in.a, r2 = in.b, h1 = in.c;
        DH1[i] = in.d + acc * z2;
        DR[(long long)row * 3 * H + col] = acc * h1 * z2 * (1.f - z2);
        DR[(long long)row * 3 * H + H + col] = in.e * (h1 - in.f) * r2 * (1.f - r2);
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int row, int col) const {
        const long long i = (long long)row * H + col;
        EpiIn4 in;
        in.a = ld4(Z2 + i); in.b = ld4(R2 + i); in.c = ld4(H1 + i); in.d = ld4(DH1 + i); in.e = ld4(DRES + i); in.f = ld4(HC2 + i);
        return in;
    }
    __device__ __forceinline__ void store4(int, int, int row, int col, const float4& acc, const EpiIn4& in) const {
        const long long i = (long long)row * H + col;
        st4(DH1 + i, in.d + acc * in.a);
        st4(DR + (long long)row * 3 * H + col, acc * in.c * in.a * one_minus(in.a));
        st4(DR + (long long)row * 3 * H + H + col, in.e * (in.c - in.f) * in.b * one_minus(in.b));
    }
    struct Cur { long long i, j; };
    __device__ __forceinline__ Cur begin4(int, int, int row, int col) const { return Cur{(long long)row * H + col, (long long)row * 3 * H + col}; }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.i += (long long)rows * H; c.j += (long long)rows * 3 * H; }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.a = ld4(Z2 + c.i); in.b = ld4(R2 + c.i); in.c = ld4(H1 + c.i); in.d = ld4(DH1 + c.i); in.e = ld4(DRES + c.i); in.f = ld4(HC2 + c.i);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const {
        pf_l2(Z2 + c.i); pf_l2(R2 + c.i); pf_l2(H1 + c.i); pf_l2(DH1 + c.i); pf_l2(DRES + c.i); pf_l2(HC2 + c.i);
    }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        st4(DH1 + c.i, in.d + acc * in.a);
        st4(DR + c.j, acc * in.c * in.a * one_minus(in.a));
        st4(DR + c.j + H, in.e * (in.c - in.f) * in.b * one_minus(in.b));
    }
    EPI_CALL_OPERATOR
};
struct EpiB2 {  // dh1 = DH1 + acc ; cell backward elementwise
    static constexpr int kBatch = 4;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 4;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    const float* DH1; const float* Hprev; const float* R; const float* HC; float* DHD; float* DG;
    int H;
    __nv_bfloat16* DG16;  // may be null: bf16 twin of the current step's DG block (same [rows, 3H] layout)
    bool vec_ok() const { return !(H & 3) && aligned16(DH1) && aligned16(Hprev) && aligned16(R) && aligned16(HC) && aligned16(DHD) && aligned16(DG); }
    __device__ __forceinline__ EpiIn load(int, int, int row, int col) const {
        const long long i = (long long)row * H + col;
        EpiIn in;
        in.a = DH1[i]; in.b = R[i]; in.c = HC[i]; in.d = Hprev[i];
        return in;
    }
    __device__ __forceinline__ void store(int, int, int row, int col, float acc, const EpiIn& in) const {
        const long long i = (long long)row * H + col;
        const float dh1 = in.a + acc;
        const float r = in.b, hc = in.c, h = in.d;
        DHD[i] = dh1 * r;
        const float gu = dh1 * (1.f - r) * (1.f - hc * hc), gr = dh1 * (h - hc) * r * (1.f - r);
        DG[(long long)row * 3 * H + 2 * H + col] = gu;
        DG[(long long)row * 3 * H + H + col] = gr;
        if (DG16) {
            DG16[(long long)row * 3 * H + 2 * H + col] = __float2bfloat16_rn(gu);
            DG16[(long long)row * 3 * H + H + col] = __float2bfloat16_rn(gr);
        }
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int row, int col) const {
        const long long i = (long long)row * H + col;
        EpiIn4 in;
        in.a = ld4(DH1 + i); in.b = ld4(R + i); in.c = ld4(HC + i); in.d = ld4(Hprev + i);
        return in;
    }
    __device__ __forceinline__ void store4(int, int, int row, int col, const float4& acc, const EpiIn4& in) const {
        const long long i = (long long)row * H + col;
        const float4 dh1 = in.a + acc;
        st4(DHD + i, dh1 * in.b);
        const float4 gu = dh1 * one_minus(in.b) * one_minus(in.c * in.c), gr = dh1 * (in.d - in.c) * in.b * one_minus(in.b);
        st4(DG + (long long)row * 3 * H + 2 * H + col, gu);
        st4(DG + (long long)row * 3 * H + H + col, gr);
        if (DG16) {
            st4_bf16(DG16 + (long long)row * 3 * H + 2 * H + col, gu);
            st4_bf16(DG16 + (long long)row * 3 * H + H + col, gr);
        }
    }
    struct Cur { long long i, j; };
    __device__ __forceinline__ Cur begin4(int, int, int row, int col) const { return Cur{(long long)row * H + col, (long long)row * 3 * H + col}; }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.i += (long long)rows * H; c.j += (long long)rows * 3 * H; }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.a = ld4(DH1 + c.i); in.b = ld4(R + c.i); in.c = ld4(HC + c.i); in.d = ld4(Hprev + c.i);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const { pf_l2(DH1 + c.i); pf_l2(R + c.i); pf_l2(HC + c.i); pf_l2(Hprev + c.i); }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        const float4 dh1 = in.a + acc;
        st4(DHD + c.i, dh1 * in.b);
        const float4 gu = dh1 * one_minus(in.b) * one_minus(in.c * in.c), gr = dh1 * (in.d - in.c) * in.b * one_minus(in.b);
        st4(DG + c.j + 2 * H, gu);
        st4(DG + c.j + H, gr);
        if (DG16) {
            st4_bf16(DG16 + c.j + 2 * H, gu);
            st4_bf16(DG16 + c.j + H, gr);
        }
    }
    EPI_CALL_OPERATOR
};
struct EpiB4 {  // dzh = acc + DP0 ; DHD += dzh*z ; DG[0:H] = dzh*h*z(1-z)      (row = node m, col = (b,c))
    static constexpr int kBatch = 4;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 4;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    const float* DP0; const float* Hprev; const float* Z; float* DHD; float* DG;
    int H, BH;  // BH = B*H columns per node
    __nv_bfloat16* DG16;  // may be null: bf16 twin of the current step's DG block
    bool vec_ok() const { return !(H & 3) && !(BH & 3) && aligned16(DP0) && aligned16(Hprev) && aligned16(Z) && aligned16(DHD) && aligned16(DG); }
    __device__ __forceinline__ EpiIn load(int, int, int row, int col) const {
        const long long i = (long long)row * BH + col;
        EpiIn in;
        in.a = DP0[i]; in.b = Z[i]; in.c = DHD[i]; in.d = Hprev[i];
        return in;
    }
    __device__ __forceinline__ void store(int, int, int row, int col, float acc, const EpiIn& in) const {
        const long long i = (long long)row * BH + col;
        const float dzh = acc + in.a;
        const float z = in.b;
        DHD[i] = in.c + dzh * z;
        const long long g = i / H;
        const int c = (int)(i - g * H);
        const float gz = dzh * in.d * z * (1.f - z);
        DG[g * 3 * H + c] = gz;
        if (DG16) DG16[g * 3 * H + c] = __float2bfloat16_rn(gz);
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int row, int col) const {
        const long long i = (long long)row * BH + col;
        EpiIn4 in;
        in.a = ld4(DP0 + i); in.b = ld4(Z + i); in.c = ld4(DHD + i); in.d = ld4(Hprev + i);
        return in;
    }
    __device__ __forceinline__ void store4(int, int, int row, int col, const float4& acc, const EpiIn4& in) const {
        const long long i = (long long)row * BH + col;
        const float4 dzh = acc + in.a;
        st4(DHD + i, in.c + dzh * in.b);
        const long long g = i / H;
        const int c = (int)(i - g * H);
        const float4 gz = dzh * in.d * in.b * one_minus(in.b);
        st4(DG + g * 3 * H + c, gz);
        if (DG16) st4_bf16(DG16 + g * 3 * H + c, gz);
    }
    struct Cur { long long i, j; };  // j: position of the same element in the [rows, 3H] gradient block
    __device__ __forceinline__ Cur begin4(int, int, int row, int col) const {
        const long long i = (long long)row * BH + col;
        const long long g = i / H;
        return Cur{i, g * 3 * H + (i - g * H)};
    }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.i += (long long)rows * BH; c.j += (long long)rows * 3 * BH; }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.a = ld4(DP0 + c.i); in.b = ld4(Z + c.i); in.c = ld4(DHD + c.i); in.d = ld4(Hprev + c.i);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const { pf_l2(DP0 + c.i); pf_l2(Z + c.i); pf_l2(DHD + c.i); pf_l2(Hprev + c.i); }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const {
        const float4 dzh = acc + in.a;
        st4(DHD + c.i, in.c + dzh * in.b);
        const float4 gz = dzh * in.d * in.b * one_minus(in.b);
        st4(DG + c.j, gz);
        if (DG16) st4_bf16(DG16 + c.j, gz);
    }
    EPI_CALL_OPERATOR
};

struct EpiB6 {  // carry = acc + DP0 + DHD
    static constexpr int kBatch = 8;  // float4 rows whose global loads the tensor-core epilogue keeps in flight per lane
    static constexpr int kPipe = 8;   // tensor-core epilogue: how many accesses (float4 rows per lane) its global reads run ahead
    const float* DP0; const float* DHD; float* OUT;
    int BH;
    bool vec_ok() const { return !(BH & 3) && aligned16(DP0) && aligned16(DHD) && aligned16(OUT); }
    __device__ __forceinline__ EpiIn load(int, int, int row, int col) const {
        const long long i = (long long)row * BH + col;
        EpiIn in;
        in.a = DP0[i]; in.b = DHD[i];
        return in;
    }
    __device__ __forceinline__ void store(int, int, int row, int col, float acc, const EpiIn& in) const {
        const long long i = (long long)row * BH + col;
        OUT[i] = acc + in.a + in.b;
    }
    __device__ __forceinline__ EpiIn4 load4(int, int, int row, int col) const {
        const long long i = (long long)row * BH + col;
        EpiIn4 in;
        in.a = ld4(DP0 + i); in.b = ld4(DHD + i);
        return in;
    }
    __device__ __forceinline__ void store4(int, int, int row, int col, const float4& acc, const EpiIn4& in) const {
        const long long i = (long long)row * BH + col;
        st4(OUT + i, acc + in.a + in.b);
    }
    struct Cur { long long i; };
    __device__ __forceinline__ Cur begin4(int, int, int row, int col) const { return Cur{(long long)row * BH + col}; }
    __device__ __forceinline__ void advance4(Cur& c, int rows) const { c.i += (long long)rows * BH; }
    __device__ __forceinline__ EpiIn4 load4(const Cur& c) const {
        EpiIn4 in;
        in.a = ld4(DP0 + c.i); in.b = ld4(DHD + c.i);
        return in;
    }
    __device__ __forceinline__ void prefetch4(const Cur& c) const { pf_l2(DP0 + c.i); pf_l2(DHD + c.i); }
    __device__ __forceinline__ void store4(const Cur& c, const float4& acc, const EpiIn4& in) const { st4(OUT + c.i, acc + in.a + in.b); }
    EPI_CALL_OPERATOR
};


}  // namespace matgcn
