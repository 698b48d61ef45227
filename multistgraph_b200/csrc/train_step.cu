// Kernels for the callers either side of the Multi-ATGCN path (f3 = dropout + output
// head, MA.py:416-417, documented at its kernels below):
//
//  f1  the optimiser half of TrafficStateExecutor._train_epoch (traffic_state_executor.py:413-422):
//      clip_grad_norm_(parameters, max_norm) followed by torch.optim.Adam.step() (executor:146-147), on ONE flat
//      fp32 bucket (parameters, gradients and both moments are contiguous; multistgraph_b200/dp.py lays them out):
//      a sum-of-squares reduction and one streaming update kernel instead of ~60 per-parameter launches.
//  f2  batch assembly: MTHDataset._get_sample_indices / _generate_input_data (mth_dataset.py:31-60, 62-158) cut every
//      sample out of one [T_total, N, F] series as (closeness, period, trend) segments + a target window, and
//      data/utils.py:68-72 + batch.py:43-57 then deep-copy / stack / upload them per batch.  Here the series stays
//      resident in HBM and one kernel gathers the batch: a batched contiguous copy, purely HBM-bound.
//
// Both are byte/float streaming work: coalesced 16-byte (or 8-byte) accesses, grid sized from the SM count.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/matgcn.h"

namespace matgcn {
inline std::atomic<unsigned long long> g_launches{0};  // the same inline variable as in gemm_simt.cuh (one definition, C++17)
}
extern "C" void matgcn_internal_set_error(const char* where, const char* what);

namespace {

int fail(const char* where, const char* what) {
    matgcn_internal_set_error(where, what);
    return -1;
}
#define TS_REQUIRE(cond, msg)                       \
    do {                                            \
        if (!(cond)) return fail(__func__, msg);    \
    } while (0)
#define TS_CK(call)                                                                    \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) return fail(__func__, cudaGetErrorString(e_));          \
    } while (0)

int sm_count_ts() {
    static int n = []() {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    return n;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------
// f1: sum of squares of the flat gradient bucket (fp32 partial sums per thread, fp64 across the block and grid: the
// result is order-independent to ~1e-16 relative, so the clip coefficient is reproducible run to run)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = g4[i];
        s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1); s2 = fmaf(v.z, v.z, s2); s3 = fmaf(v.w, v.w, s3);
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s0 = fmaf(g[i], g[i], s0);
    double s = (double)s0 + (double)s1 + (double)s2 + (double)s3;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = part[threadIdx.x];
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) s += __shfl_xor_sync(0xffu, s, off);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}

struct AdamP {
    float lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale;
    float bc1, bc2_sqrt;  // 1 - beta1^t, sqrt(1 - beta2^t)
    int write_grad;
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float coef, const AdamP& a) {
    g *= coef;                                        // all-reduce mean (grad_scale) and clip_grad_norm_ in one factor
    float gd = g;
    if (a.weight_decay != 0.f) gd = fmaf(a.weight_decay, p, gd);       // torch Adam: grad = grad + wd * param
    m = m + (gd - m) * (1.f - a.beta1);               // exp_avg.lerp_(grad, 1 - beta1)
    v = a.beta2 * v + (1.f - a.beta2) * gd * gd;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - (a.lr / a.bc1) * (m / denom);             // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, const double* __restrict__ sumsq,
                                                        AdamP a, float* __restrict__ norm_out) {
    // clip coefficient: torch.nn.utils.clip_grad_norm_ -> clamp(max_norm / (total_norm + 1e-6), max=1)
    float coef = a.grad_scale;
    if (sumsq) {
        const float total = (float)(sqrt(*sumsq) * (double)fabsf(a.grad_scale));
        if (a.max_norm > 0.f) coef *= fminf(a.max_norm / (total + 1e-6f), 1.f);
        if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
    }
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float4* p4 = reinterpret_cast<float4*>(p); float4* g4 = reinterpret_cast<float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, coef, a); adam_one(pp.y, gg.y, mm.y, vv.y, coef, a);
        adam_one(pp.z, gg.z, mm.z, vv.z, coef, a); adam_one(pp.w, gg.w, mm.w, vv.w, coef, a);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
        if (a.write_grad) g4[i] = gg;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        adam_one(pp, gg, mm, vv, coef, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (a.write_grad) g[i] = gg;
    }
}

// ------------------------------------------------------------------------------------------
// f2: window gather.  One "chunk" = one segment of one sample: `len` consecutive time slices = len * row floats that
// are contiguous both in the series and in the destination.  grid.y walks the chunks, grid.x tiles a chunk.
// VEC = floats per access (4, 2 or 1) chosen by the host from the alignment of row, base pointers and strides.
// ------------------------------------------------------------------------------------------
template <int VEC>
struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int VEC>
__global__ void __launch_bounds__(256) assemble_windows_kernel(const float* __restrict__ series, long long T_total, long long row,
                                                               const int* __restrict__ seg_offsets, int n_seg, int in_window,
                                                               int out_window, const long long* __restrict__ label_starts, int B,
                                                               float* __restrict__ X, float* __restrict__ y, int* __restrict__ bad) {
    using V = typename VecT<VEC>::type;
    const int chunks = B * (n_seg + 1);
    for (int ch = blockIdx.y; ch < chunks; ch += gridDim.y) {
        const int b = ch / (n_seg + 1), s = ch - b * (n_seg + 1);
        const long long t0 = label_starts[b];
        long long src_t, len;
        float* dst;
        if (s < n_seg) {
            src_t = t0 - (long long)seg_offsets[s]; len = in_window;
            dst = X + ((long long)b * n_seg + s) * in_window * row;
        } else {
            src_t = t0; len = out_window;
            dst = y + (long long)b * out_window * row;
        }
        if (src_t < 0 || src_t + len > T_total || t0 + in_window > T_total) {   // mth_dataset.py:45-46, 55-58, 78-79: not a valid sample
            if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(bad, 1);
            continue;
        }
        const V* sv = reinterpret_cast<const V*>(series + src_t * row);
        V* dv = reinterpret_cast<V*>(dst);
        const long long nv = len * row / VEC;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) dv[i] = sv[i];
    }
}


// ------------------------------------------------------------------------------------------
// f3: dropout + output head.  MA.py:416-417: F.dropout(p, training) on the encoder output, then
// Conv2d(T -> T_out*C, kernel (1, H)): time steps are the channels, i.e. out[r, o] = bias[o] + sum_t sum_h drop(y[t,r,h]) w[o,t,h]
// for every (node, batch) row r.  fp32 FFMA (the 1e-4 parity bound holds in every mode); HBM-bound on the 4*T*rows*H
// bytes of y, which are read once in the forward and once in the backward; the mask is never stored.
//
// Dropout mask: counter-based (Philox4x32-10 keyed by the seed): element e = (t*rows + r)*H + h belongs to group e/8,
// whose 128 random bits give eight 16-bit lanes; lane < thr drops.  thr = round(p*65536); the realised drop
// probability thr/65536 (0.100006 for p = 0.1) is what the kept values are rescaled by, so E[drop(x)] = x exactly.
// ------------------------------------------------------------------------------------------
struct DropP {
    unsigned long long seed;
    uint32_t thr;   // 0: no dropout
    float scale;    // 1 / (1 - thr/65536)
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
    uint32_t x0 = c0, x1 = c1, x2 = 0u, x3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        x0 = hi1 ^ x1 ^ k0; x1 = lo1; x2 = hi0 ^ x3 ^ k1; x3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
}
// keep-multipliers (0 or scale) of the four elements e..e+3 (e % 4 == 0)
__device__ __forceinline__ float4 drop_mult4(const DropP& d, unsigned long long e) {
    if (d.thr == 0u) return make_float4(1.f, 1.f, 1.f, 1.f);
    const unsigned long long grp = e >> 3;
    const uint4 rnd = philox4x32_10((uint32_t)grp, (uint32_t)(grp >> 32), (uint32_t)d.seed, (uint32_t)(d.seed >> 32));
    const uint32_t a = (e & 4ull) ? rnd.z : rnd.x, b = (e & 4ull) ? rnd.w : rnd.y;
    return make_float4((a & 0xFFFFu) >= d.thr ? d.scale : 0.f, (a >> 16) >= d.thr ? d.scale : 0.f,
                       (b & 0xFFFFu) >= d.thr ? d.scale : 0.f, (b >> 16) >= d.thr ? d.scale : 0.f);
}
__device__ __forceinline__ float4 mul4(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}

constexpr int HD_H = 64;        // rnn_units the head kernels are written for
constexpr int HD_RT = 64;       // rows per tile
constexpr int HD_LDY = 68;      // smem pitch of the y tile / w tile (floats): conflict-free 16-byte row accesses
constexpr int HD_OMAX = 32;     // output channels per backward pass
constexpr int HD_OFWD = 24;     // output channels per forward pass (6 per thread)

// forward: one block per 64-row tile, loop over t; thread (row = tid/4, og = tid%4) owns OPT outputs o = og + 4*j
template <int OPT>
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ y, long long y_tstride, int Tc, long long rows,
                                                       const float* __restrict__ w, const float* __restrict__ bias, int O, int o0,
                                                       DropP dp, float* __restrict__ out) {
    __shared__ __align__(16) float ys[2][HD_RT * HD_LDY];
    __shared__ __align__(16) float wsm[2][4 * OPT * HD_LDY];
    const int tid = threadIdx.x, row = tid >> 2, og = tid & 3;
    const long long r0 = (long long)blockIdx.x * HD_RT;
    float acc[OPT];
#pragma unroll
    for (int j = 0; j < OPT; ++j) acc[j] = 0.f;
    auto load_tile = [&](int t, int buf) {
        // y tile: 64 rows x 16 float4; thread -> float4 pairs (one 8-element mask group)
        for (int idx = tid; idx < HD_RT * 8; idx += 256) {
            const int r = idx >> 3, h8 = (idx & 7) * 8;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (r0 + r < rows) {
                const float* src = y + (long long)t * y_tstride + (r0 + r) * HD_H + h8;
                a = *reinterpret_cast<const float4*>(src); b = *reinterpret_cast<const float4*>(src + 4);
                const unsigned long long e = ((unsigned long long)t * rows + (r0 + r)) * HD_H + h8;
                a = mul4(a, drop_mult4(dp, e)); b = mul4(b, drop_mult4(dp, e + 4));
            }
            *reinterpret_cast<float4*>(&ys[buf][r * HD_LDY + h8]) = a;
            *reinterpret_cast<float4*>(&ys[buf][r * HD_LDY + h8 + 4]) = b;
        }
        for (int idx = tid; idx < 4 * OPT * 16; idx += 256) {
            const int oo = idx >> 4, h4 = (idx & 15) * 4, o = o0 + oo;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o < O && oo < HD_OFWD) v = *reinterpret_cast<const float4*>(w + ((long long)o * Tc + t) * HD_H + h4);
            *reinterpret_cast<float4*>(&wsm[buf][oo * HD_LDY + h4]) = v;
        }
    };
    load_tile(0, 0);
    __syncthreads();
    for (int t = 0; t < Tc; ++t) {
        const int buf = t & 1;
        if (t + 1 < Tc) load_tile(t + 1, buf ^ 1);
        const float* yr = &ys[buf][row * HD_LDY];
#pragma unroll 4
        for (int h4 = 0; h4 < 16; ++h4) {
            const float4 yv = *reinterpret_cast<const float4*>(yr + 4 * h4);
#pragma unroll
            for (int j = 0; j < OPT; ++j) acc[j] = dot4(yv, *reinterpret_cast<const float4*>(&wsm[buf][(og + 4 * j) * HD_LDY + 4 * h4]), acc[j]);
        }
        __syncthreads();
    }
    if (r0 + row < rows) {
#pragma unroll
        for (int j = 0; j < OPT; ++j) {
            const int o = o0 + og + 4 * j;
            if (o < O && og + 4 * j < HD_OFWD) out[(r0 + row) * O + o] = acc[j] + bias[o];
        }
    }
}

// backward: block (chunk, t) walks its row tiles: dy tile = mask * (dout tile x w_t), dw_t += dout^T x drop(y) (registers,
// one atomic per element per block at the end), dbias by the t == 0 blocks.
//   dy mapping: thread (rg = tid/16, hg = tid%16) -> rows 4rg..4rg+3, columns 4hg..4hg+3
//   dw mapping: thread (os = tid/16, hg = tid%16) -> outputs os, os+16 (of this 32-wide pass), columns 4hg..4hg+3
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ y, long long y_tstride, int Tc, long long rows,
                                                       const float* __restrict__ w, int O, int o0, DropP dp,
                                                       const float* __restrict__ dout, float* __restrict__ dy, int dy_accumulate,
                                                       float* __restrict__ dw, float* __restrict__ dbias) {
    __shared__ __align__(16) float ys[HD_RT * HD_LDY];          // dropped y tile [row][h]
    __shared__ __align__(16) float wsm[HD_OMAX * HD_LDY];       // w_t [o][h]
    __shared__ __align__(16) float ds[HD_OMAX * HD_LDY];        // dout tile transposed [o][row]
    const int tid = threadIdx.x, t = blockIdx.y;
    const int rg = tid >> 4, hg = tid & 15, os = tid >> 4;
    const int ow = min(O - o0, HD_OMAX);
    for (int idx = tid; idx < HD_OMAX * 16; idx += 256) {
        const int oo = idx >> 4, h4 = (idx & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oo < ow) v = *reinterpret_cast<const float4*>(w + ((long long)(o0 + oo) * Tc + t) * HD_H + h4);
        *reinterpret_cast<float4*>(&wsm[oo * HD_LDY + h4]) = v;
    }
    float4 dwa = make_float4(0.f, 0.f, 0.f, 0.f), dwb = dwa;
    float db = 0.f;
    const long long ntiles = (rows + HD_RT - 1) / HD_RT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * HD_RT;
        __syncthreads();   // previous tile's readers are done (and wsm is visible on the first pass)
        for (int idx = tid; idx < HD_RT * 8; idx += 256) {
            const int r = idx >> 3, h8 = (idx & 7) * 8;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (r0 + r < rows) {
                const float* src = y + (long long)t * y_tstride + (r0 + r) * HD_H + h8;
                a = *reinterpret_cast<const float4*>(src); b = *reinterpret_cast<const float4*>(src + 4);
                const unsigned long long e = ((unsigned long long)t * rows + (r0 + r)) * HD_H + h8;
                a = mul4(a, drop_mult4(dp, e)); b = mul4(b, drop_mult4(dp, e + 4));
            }
            *reinterpret_cast<float4*>(&ys[r * HD_LDY + h8]) = a;
            *reinterpret_cast<float4*>(&ys[r * HD_LDY + h8 + 4]) = b;
        }
        for (int idx = tid; idx < HD_RT * HD_OMAX; idx += 256) {
            const int r = idx / HD_OMAX, oo = idx % HD_OMAX;
            ds[oo * HD_LDY + r] = (oo < ow && r0 + r < rows) ? dout[(r0 + r) * O + o0 + oo] : 0.f;
        }
        __syncthreads();
        // ---- dy tile ----
        float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0, d2 = d0, d3 = d0;
        for (int oo = 0; oo < ow; ++oo) {
            const float4 dv = *reinterpret_cast<const float4*>(&ds[oo * HD_LDY + 4 * rg]);
            const float4 wv = *reinterpret_cast<const float4*>(&wsm[oo * HD_LDY + 4 * hg]);
            d0.x = fmaf(dv.x, wv.x, d0.x); d0.y = fmaf(dv.x, wv.y, d0.y); d0.z = fmaf(dv.x, wv.z, d0.z); d0.w = fmaf(dv.x, wv.w, d0.w);
            d1.x = fmaf(dv.y, wv.x, d1.x); d1.y = fmaf(dv.y, wv.y, d1.y); d1.z = fmaf(dv.y, wv.z, d1.z); d1.w = fmaf(dv.y, wv.w, d1.w);
            d2.x = fmaf(dv.z, wv.x, d2.x); d2.y = fmaf(dv.z, wv.y, d2.y); d2.z = fmaf(dv.z, wv.z, d2.z); d2.w = fmaf(dv.z, wv.w, d2.w);
            d3.x = fmaf(dv.w, wv.x, d3.x); d3.y = fmaf(dv.w, wv.y, d3.y); d3.z = fmaf(dv.w, wv.z, d3.z); d3.w = fmaf(dv.w, wv.w, d3.w);
        }
        {
            const float4* dd[4] = {&d0, &d1, &d2, &d3};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long r = r0 + 4 * rg + i;
                if (r < rows) {
                    const unsigned long long e = ((unsigned long long)t * rows + r) * HD_H + 4 * hg;
                    float4 v = mul4(*dd[i], drop_mult4(dp, e));
                    float4* dst = reinterpret_cast<float4*>(dy + ((long long)t * rows + r) * HD_H + 4 * hg);
                    if (dy_accumulate) { const float4 old = *dst; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
                    *dst = v;
                }
            }
        }
        // ---- dw_t partial: outputs os and os+16, columns 4hg.. ----
#pragma unroll 4
        for (int r4 = 0; r4 < HD_RT / 4; ++r4) {
            const float4 da = *reinterpret_cast<const float4*>(&ds[os * HD_LDY + 4 * r4]);
            const float4 dbv = *reinterpret_cast<const float4*>(&ds[(os + 16) * HD_LDY + 4 * r4]);
            const float4 y0 = *reinterpret_cast<const float4*>(&ys[(4 * r4 + 0) * HD_LDY + 4 * hg]);
            const float4 y1 = *reinterpret_cast<const float4*>(&ys[(4 * r4 + 1) * HD_LDY + 4 * hg]);
            const float4 y2 = *reinterpret_cast<const float4*>(&ys[(4 * r4 + 2) * HD_LDY + 4 * hg]);
            const float4 y3 = *reinterpret_cast<const float4*>(&ys[(4 * r4 + 3) * HD_LDY + 4 * hg]);
            dwa.x = fmaf(da.x, y0.x, fmaf(da.y, y1.x, fmaf(da.z, y2.x, fmaf(da.w, y3.x, dwa.x))));
            dwa.y = fmaf(da.x, y0.y, fmaf(da.y, y1.y, fmaf(da.z, y2.y, fmaf(da.w, y3.y, dwa.y))));
            dwa.z = fmaf(da.x, y0.z, fmaf(da.y, y1.z, fmaf(da.z, y2.z, fmaf(da.w, y3.z, dwa.z))));
            dwa.w = fmaf(da.x, y0.w, fmaf(da.y, y1.w, fmaf(da.z, y2.w, fmaf(da.w, y3.w, dwa.w))));
            dwb.x = fmaf(dbv.x, y0.x, fmaf(dbv.y, y1.x, fmaf(dbv.z, y2.x, fmaf(dbv.w, y3.x, dwb.x))));
            dwb.y = fmaf(dbv.x, y0.y, fmaf(dbv.y, y1.y, fmaf(dbv.z, y2.y, fmaf(dbv.w, y3.y, dwb.y))));
            dwb.z = fmaf(dbv.x, y0.z, fmaf(dbv.y, y1.z, fmaf(dbv.z, y2.z, fmaf(dbv.w, y3.z, dwb.z))));
            dwb.w = fmaf(dbv.x, y0.w, fmaf(dbv.y, y1.w, fmaf(dbv.z, y2.w, fmaf(dbv.w, y3.w, dwb.w))));
            if (t == 0 && hg == 0) db += (da.x + da.y) + (da.z + da.w);
            if (t == 0 && hg == 1) db += (dbv.x + dbv.y) + (dbv.z + dbv.w);
        }
    }
    if (os < ow) {
        float* dst = dw + ((long long)(o0 + os) * Tc + t) * HD_H + 4 * hg;
        atomicAdd(dst, dwa.x); atomicAdd(dst + 1, dwa.y); atomicAdd(dst + 2, dwa.z); atomicAdd(dst + 3, dwa.w);
    }
    if (os + 16 < ow) {
        float* dst = dw + ((long long)(o0 + os + 16) * Tc + t) * HD_H + 4 * hg;
        atomicAdd(dst, dwb.x); atomicAdd(dst + 1, dwb.y); atomicAdd(dst + 2, dwb.z); atomicAdd(dst + 3, dwb.w);
    }
    if (t == 0 && dbias) {
        if (hg == 0 && os < ow) atomicAdd(dbias + o0 + os, db);
        if (hg == 1 && os + 16 < ow) atomicAdd(dbias + o0 + os + 16, db);
    }
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(long long n4, DropP dp, float* __restrict__ mult) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(mult)[i] = drop_mult4(dp, (unsigned long long)i * 4);
}

DropP make_drop(float p, unsigned long long seed) {
    DropP d;
    d.seed = seed;
    d.thr = p > 0.f ? (uint32_t)lrintf(p * 65536.f) : 0u;
    if (d.thr > 65535u) d.thr = 65535u;
    d.scale = 65536.f / (float)(65536u - d.thr);
    return d;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int matgcn_grad_sumsq(const float* grad, long long n, double* sumsq, void* stream) {
    TS_REQUIRE(grad && sumsq && n >= 0, "null pointer or negative length");
    TS_REQUIRE(aligned16(grad), "grad must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TS_CK(cudaMemsetAsync(sumsq, 0, sizeof(double), st));
    if (n == 0) return 0;
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = 4LL * sm_count_ts();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    grad_sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(grad, n, sumsq);
    matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_adam_clip_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                     const double* sumsq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                                     float eps, float weight_decay, long long step, int write_grad, float* norm_out, void* stream) {
    TS_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0, "null pointer or negative length");
    TS_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), "buffers must be 16-byte aligned");
    TS_REQUIRE(step >= 1, "step counts from 1 (torch.optim.Adam increments before the update)");
    TS_REQUIRE(max_norm <= 0.f || sumsq, "clipping needs the sum of squares from matgcn_grad_sumsq");
    if (n == 0) return 0;
    AdamP a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    a.grad_scale = grad_scale;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.write_grad = write_grad;
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = 8LL * sm_count_ts();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_clip_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, sumsq, a, norm_out);
    matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_assemble_windows(const float* series, long long T_total, int N, int F, const int* seg_offsets, int n_seg,
                                       int in_window, int out_window, const long long* label_starts, int B, float* X, float* y,
                                       int* bad_flag, void* stream) {
    TS_REQUIRE(T_total > 0 && N > 0 && F > 0 && n_seg > 0 && in_window > 0 && out_window > 0 && B >= 0, "bad dimensions");
    if (B == 0) return 0;   // an empty batch has no buffers to speak of
    TS_REQUIRE(series && seg_offsets && label_starts && X && y && bad_flag, "null pointer");
    const long long row = (long long)N * F;
    const uintptr_t bases = reinterpret_cast<uintptr_t>(series) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(y);
    const int vec = (!(row & 3) && !(bases & 15)) ? 4 : (!(row & 1) && !(bases & 7)) ? 2 : 1;
    const long long per_chunk = (long long)in_window * row / vec;
    long long gx = (per_chunk + 255) / 256;
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    long long gy = (long long)B * (n_seg + 1);
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec == 4)
        assemble_windows_kernel<4><<<grid, 256, 0, st>>>(series, T_total, row, seg_offsets, n_seg, in_window, out_window, label_starts, B, X, y, bad_flag);
    else if (vec == 2)
        assemble_windows_kernel<2><<<grid, 256, 0, st>>>(series, T_total, row, seg_offsets, n_seg, in_window, out_window, label_starts, B, X, y, bad_flag);
    else
        assemble_windows_kernel<1><<<grid, 256, 0, st>>>(series, T_total, row, seg_offsets, n_seg, in_window, out_window, label_starts, B, X, y, bad_flag);
    matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" float matgcn_head_dropout_scale(float p_drop) { return make_drop(p_drop, 0).scale; }

extern "C" int matgcn_head_fwd(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, const float* bias,
                               int O, float p_drop, unsigned long long seed, float* out, void* stream) {
    TS_REQUIRE(y && w && bias && out, "null pointer");
    TS_REQUIRE(Tc > 0 && rows >= 0 && O > 0 && p_drop >= 0.f && p_drop < 1.f, "bad dimensions or dropout probability");
    TS_REQUIRE(H == HD_H, "the head kernels are written for rnn_units = 64");
    TS_REQUIRE(aligned16(y) && aligned16(w) && !(y_tstride & 3), "y and w must be 16-byte aligned");
    if (rows == 0) return 0;
    const DropP dp = make_drop(p_drop, seed);
    const unsigned grid = (unsigned)((rows + HD_RT - 1) / HD_RT);
    cudaStream_t st = (cudaStream_t)stream;
    for (int o0 = 0; o0 < O; o0 += HD_OFWD) {
        const int ow = O - o0 < HD_OFWD ? O - o0 : HD_OFWD;
        if (ow <= 4) head_fwd_kernel<1><<<grid, 256, 0, st>>>(y, y_tstride, Tc, rows, w, bias, O, o0, dp, out);
        else if (ow <= 12) head_fwd_kernel<3><<<grid, 256, 0, st>>>(y, y_tstride, Tc, rows, w, bias, O, o0, dp, out);
        else head_fwd_kernel<6><<<grid, 256, 0, st>>>(y, y_tstride, Tc, rows, w, bias, O, o0, dp, out);
        matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_head_bwd(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, int O, float p_drop,
                               unsigned long long seed, const float* dout, float* dy, float* dw, float* dbias, void* stream) {
    TS_REQUIRE(y && w && dout && dy && dw && dbias, "null pointer");
    TS_REQUIRE(Tc > 0 && rows >= 0 && O > 0 && p_drop >= 0.f && p_drop < 1.f, "bad dimensions or dropout probability");
    TS_REQUIRE(H == HD_H, "the head kernels are written for rnn_units = 64");
    TS_REQUIRE(aligned16(y) && aligned16(w) && aligned16(dy) && !(y_tstride & 3), "y, w and dy must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TS_CK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)O * Tc * H, st));
    TS_CK(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)O, st));
    if (rows == 0) return 0;
    const DropP dp = make_drop(p_drop, seed);
    const long long ntiles = (rows + HD_RT - 1) / HD_RT;
    long long chunks = (2LL * sm_count_ts() + Tc - 1) / Tc;      // ~2 blocks per SM in total
    if (chunks > ntiles) chunks = ntiles;
    if (chunks < 1) chunks = 1;
    dim3 grid((unsigned)chunks, (unsigned)Tc);
    for (int o0 = 0; o0 < O; o0 += HD_OMAX) {
        head_bwd_kernel<<<grid, 256, 0, st>>>(y, y_tstride, Tc, rows, w, O, o0, dp, dout, dy, o0 > 0 ? 1 : 0, dw, dbias);
        matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_head_dropout_mask(long long n, float p_drop, unsigned long long seed, float* mult, void* stream) {
    TS_REQUIRE(mult && n >= 0 && !(n & 3) && aligned16(mult), "mask length must be a multiple of 4 and the buffer 16-byte aligned");
    if (n == 0) return 0;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 8LL * sm_count_ts()) blocks = 8LL * sm_count_ts();
    dropout_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n / 4, make_drop(p_drop, seed), mult);
    matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    TS_CK(cudaGetLastError());
    return 0;
}
