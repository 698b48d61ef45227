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
    const long long* step_dev;   // optional device-resident step count / learning rate (a captured step replays with new values):
    const float* lr_dev;         // the kernel derives the bias corrections from *step_dev itself
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
    if (a.step_dev) {   // same double-precision expressions as the host form (matgcn_adam_clip_step)
        const double st = (double)__ldg(a.step_dev);
        a.bc1 = (float)(1.0 - pow((double)a.beta1, st));
        a.bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, st));
    }
    if (a.lr_dev) a.lr = __ldg(a.lr_dev);
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
    const unsigned long long* seed_dev;   // optional device-resident key, XORed into `seed` by the kernel (CUDA-graph replays)
};
// the key a kernel uses: the host half XOR the device half (advanced by matgcn_step_tick between replays of a captured step)
__device__ __forceinline__ void drop_resolve(DropP& d) {
    if (d.seed_dev) d.seed ^= __ldg(d.seed_dev);
}

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
constexpr int HD_RT = 32;       // rows per tile: 128 threads, thread (rg = tid/16, hg = tid%16) owns rows 4rg..4rg+3, columns 4hg..4hg+3
constexpr int HD_THREADS = 128;
constexpr int HD_LD = 68;       // smem pitch (floats): conflict-free 16-byte row accesses
constexpr int HD_OMAX = 24;     // output channels per pass (one accumulator set per channel and thread)

// the thread's 4x4 block of y (rows 4rg+i, columns 4hg..) -> registers; rows beyond the end read as zero
__device__ __forceinline__ void head_load_y(const float* __restrict__ y, long long y_tstride, int t, long long rows, long long r0,
                                            int rg, int hg, float4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = r0 + 4 * rg + i;
        v[i] = r < rows ? *reinterpret_cast<const float4*>(y + (long long)t * y_tstride + r * HD_H + 4 * hg) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// keep-multipliers of the same 4x4 block.  The two lanes of a pair (hg even / odd) share every 8-element mask group:
// each computes the Philox block of two of the four rows and they swap halves.
__device__ __forceinline__ void head_mults(const DropP& dp, int t, long long rows, long long r0, int rg, int hg, float4 (&m)[4]) {
    if (dp.thr == 0u) {
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        return;
    }
    const bool odd = hg & 1;
    uint32_t oa[2], ob[2], ra[2], rb[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const long long r = r0 + 4 * rg + (odd ? 2 : 0) + j;
        const unsigned long long grp = (((unsigned long long)t * rows + r) * HD_H + 4 * hg) >> 3;
        const uint4 rnd = philox4x32_10((uint32_t)grp, (uint32_t)(grp >> 32), (uint32_t)dp.seed, (uint32_t)(dp.seed >> 32));
        oa[j] = odd ? rnd.z : rnd.x; ob[j] = odd ? rnd.w : rnd.y;
        ra[j] = __shfl_xor_sync(0xffffffffu, odd ? rnd.x : rnd.z, 1);
        rb[j] = __shfl_xor_sync(0xffffffffu, odd ? rnd.y : rnd.w, 1);
    }
    const uint32_t a[4] = {odd ? ra[0] : oa[0], odd ? ra[1] : oa[1], odd ? oa[0] : ra[0], odd ? oa[1] : ra[1]};
    const uint32_t b[4] = {odd ? rb[0] : ob[0], odd ? rb[1] : ob[1], odd ? ob[0] : rb[0], odd ? ob[1] : rb[1]};
#pragma unroll
    for (int i = 0; i < 4; ++i)
        m[i] = make_float4((a[i] & 0xFFFFu) >= dp.thr ? dp.scale : 0.f, (a[i] >> 16) >= dp.thr ? dp.scale : 0.f,
                           (b[i] & 0xFFFFu) >= dp.thr ? dp.scale : 0.f, (b[i] >> 16) >= dp.thr ? dp.scale : 0.f);
}

// forward: one block per 32-row tile, loop over t with the accumulators (OW outputs x 4 rows, partial over the thread's 4
// columns) in registers; one butterfly over the 16 column groups at the end.  y goes global -> registers, w_t through smem.
template <int OW>
__global__ void __launch_bounds__(HD_THREADS) head_fwd_kernel(const float* __restrict__ y, long long y_tstride, int Tc, long long rows,
                                                              const float* __restrict__ w, const float* __restrict__ bias, int O, int o0,
                                                              DropP dp, float* __restrict__ out) {
    __shared__ __align__(16) float wsm[2][OW * HD_LD];
    drop_resolve(dp);
    const int tid = threadIdx.x, rg = tid >> 4, hg = tid & 15;
    const long long r0 = (long long)blockIdx.x * HD_RT;
    const int ow = min(O - o0, OW);
    float acc[OW][4];
#pragma unroll
    for (int oo = 0; oo < OW; ++oo)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[oo][i] = 0.f;
    // w_t tile: global -> registers (issued a whole step ahead) -> shared memory after the step's arithmetic
    constexpr int WPT = (OW * 16 + HD_THREADS - 1) / HD_THREADS;
    float4 wreg[WPT];
    auto fetch_w = [&](int t) {
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const int idx = tid + k * HD_THREADS, oo = idx >> 4, h4 = (idx & 15) * 4;
            wreg[k] = (idx < OW * 16 && oo < ow) ? *reinterpret_cast<const float4*>(w + ((long long)(o0 + oo) * Tc + t) * HD_H + h4)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto stash_w = [&](int buf) {
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const int idx = tid + k * HD_THREADS, oo = idx >> 4, h4 = (idx & 15) * 4;
            if (idx < OW * 16) *reinterpret_cast<float4*>(&wsm[buf][oo * HD_LD + h4]) = wreg[k];
        }
    };
    float4 yv[4], yn[4];
    head_load_y(y, y_tstride, 0, rows, r0, rg, hg, yv);
    fetch_w(0);
    stash_w(0);
    __syncthreads();
    for (int t = 0; t < Tc; ++t) {
        const int buf = t & 1;
        if (t + 1 < Tc) { head_load_y(y, y_tstride, t + 1, rows, r0, rg, hg, yn); fetch_w(t + 1); }
        float4 m[4];
        head_mults(dp, t, rows, r0, rg, hg, m);
#pragma unroll
        for (int i = 0; i < 4; ++i) yv[i] = mul4(yv[i], m[i]);
#pragma unroll
        for (int oo = 0; oo < OW; ++oo) {
            const float4 wv = *reinterpret_cast<const float4*>(&wsm[buf][oo * HD_LD + 4 * hg]);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[oo][i] = dot4(yv[i], wv, acc[oo][i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) yv[i] = yn[i];
        if (t + 1 < Tc) stash_w(buf ^ 1);   // (buffer buf^1 was last read in step t-1, which every thread has left)
        __syncthreads();
    }
    // sum over the 16 column groups (lanes hg of a half-warp), then lane hg writes outputs hg and hg+16
#pragma unroll
    for (int oo = 0; oo < OW; ++oo)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = acc[oo][i];
            v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8);
            acc[oo][i] = v;
        }
#pragma unroll
    for (int oo = 0; oo < OW; ++oo) {
        if ((oo & 15) == hg && oo < ow) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long r = r0 + 4 * rg + i;
                if (r < rows) out[r * O + o0 + oo] = acc[oo][i] + bias[o0 + oo];
            }
        }
    }
}

// backward: block (chunk, t) walks its row tiles.  Per tile and thread: dy block = mask * (dout rows x w_t) and the partial
// dw_t[o][4hg..] += sum over its 4 rows of dout[r][o] * drop(y)[r][4hg..] - both from the SAME two shared-memory reads per
// output channel (32 FMA per 2 LDS.128).  The dw partials stay in registers over all tiles of the block and are reduced over
// the row groups once at the end (shuffle + shared memory), then one atomic per element and block.
template <int OW>
__global__ void __launch_bounds__(HD_THREADS) head_bwd_kernel(const float* __restrict__ y, long long y_tstride, int Tc, long long rows,
                                                              const float* __restrict__ w, int O, int o0, DropP dp,
                                                              const float* __restrict__ dout, float* __restrict__ dy, int dy_accumulate,
                                                              float* __restrict__ dw) {
    __shared__ __align__(16) float wsm[OW * HD_LD];            // w_t [o][h]
    __shared__ __align__(16) float ds[OW * HD_LD];             // dout tile transposed [o][row] (32 rows used)
    __shared__ __align__(16) float red[4 * 8 * HD_H];          // end-of-block reduction: [warp][8 outputs][h]
    drop_resolve(dp);
    const int tid = threadIdx.x, t = blockIdx.y, rg = tid >> 4, hg = tid & 15, warp = tid >> 5, lane = tid & 31;
    const int ow = min(O - o0, OW);
    for (int idx = tid; idx < OW * 16; idx += HD_THREADS) {
        const int oo = idx >> 4, h4 = (idx & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oo < ow) v = *reinterpret_cast<const float4*>(w + ((long long)(o0 + oo) * Tc + t) * HD_H + h4);
        *reinterpret_cast<float4*>(&wsm[oo * HD_LD + h4]) = v;
    }
    float4 dwacc[OW];
#pragma unroll
    for (int oo = 0; oo < OW; ++oo) dwacc[oo] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long ntiles = (rows + HD_RT - 1) / HD_RT;
    float4 yv[4];
    constexpr int DPT_ = HD_RT * OW / HD_THREADS;   // dout elements per thread and tile
    static_assert(HD_RT * OW % HD_THREADS == 0, "tile must divide evenly");
    float dreg[DPT_];
    auto fetch_d = [&](long long r0) {
#pragma unroll
        for (int k = 0; k < DPT_; ++k) {
            const int idx = tid + k * HD_THREADS, r = idx / OW, oo = idx - r * OW;
            dreg[k] = (oo < ow && r0 + r < rows) ? dout[(r0 + r) * O + o0 + oo] : 0.f;
        }
    };
    if ((long long)blockIdx.x < ntiles) {
        head_load_y(y, y_tstride, t, rows, (long long)blockIdx.x * HD_RT, rg, hg, yv);
        fetch_d((long long)blockIdx.x * HD_RT);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * HD_RT;
        __syncthreads();   // the previous tile's readers of ds are done (and wsm is visible on the first pass)
#pragma unroll
        for (int k = 0; k < DPT_; ++k) {
            const int idx = tid + k * HD_THREADS, r = idx / OW, oo = idx - r * OW;
            ds[oo * HD_LD + r] = dreg[k];
        }
        float4 m[4], yn[4];
        head_mults(dp, t, rows, r0, rg, hg, m);
        if (tile + gridDim.x < ntiles) {
            head_load_y(y, y_tstride, t, rows, (tile + gridDim.x) * HD_RT, rg, hg, yn);
            fetch_d((tile + gridDim.x) * HD_RT);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) yv[i] = mul4(yv[i], m[i]);
        __syncthreads();
        float4 d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int oo = 0; oo < OW; ++oo) {
            const float4 dv = *reinterpret_cast<const float4*>(&ds[oo * HD_LD + 4 * rg]);
            const float4 wv = *reinterpret_cast<const float4*>(&wsm[oo * HD_LD + 4 * hg]);
            const float dvv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                d[i].x = fmaf(dvv[i], wv.x, d[i].x); d[i].y = fmaf(dvv[i], wv.y, d[i].y);
                d[i].z = fmaf(dvv[i], wv.z, d[i].z); d[i].w = fmaf(dvv[i], wv.w, d[i].w);
                dwacc[oo].x = fmaf(dvv[i], yv[i].x, dwacc[oo].x); dwacc[oo].y = fmaf(dvv[i], yv[i].y, dwacc[oo].y);
                dwacc[oo].z = fmaf(dvv[i], yv[i].z, dwacc[oo].z); dwacc[oo].w = fmaf(dvv[i], yv[i].w, dwacc[oo].w);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = r0 + 4 * rg + i;
            if (r < rows) {
                float4 v = mul4(d[i], m[i]);
                float4* dst = reinterpret_cast<float4*>(dy + ((long long)t * rows + r) * HD_H + 4 * hg);
                if (dy_accumulate) { const float4 old = *dst; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
                *dst = v;
            }
            yv[i] = yn[i];
        }
    }
    // dw_t: sum the two row groups of a warp by shuffle, the four warps through shared memory, 8 output channels at a time
#pragma unroll
    for (int oo = 0; oo < OW; ++oo) {
        dwacc[oo].x += __shfl_xor_sync(0xffffffffu, dwacc[oo].x, 16); dwacc[oo].y += __shfl_xor_sync(0xffffffffu, dwacc[oo].y, 16);
        dwacc[oo].z += __shfl_xor_sync(0xffffffffu, dwacc[oo].z, 16); dwacc[oo].w += __shfl_xor_sync(0xffffffffu, dwacc[oo].w, 16);
    }
#pragma unroll
    for (int c = 0; c < (OW + 7) / 8; ++c) {
        __syncthreads();
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (8 * c + j < OW) *reinterpret_cast<float4*>(&red[(warp * 8 + j) * HD_H + 4 * hg]) = dwacc[8 * c + j];
        }
        __syncthreads();
        for (int idx = tid; idx < 8 * HD_H; idx += HD_THREADS) {
            const int j = idx >> 6, h = idx & 63, oo = 8 * c + j;
            if (oo < ow) {
                const float v = (red[(0 * 8 + j) * HD_H + h] + red[(1 * 8 + j) * HD_H + h]) + (red[(2 * 8 + j) * HD_H + h] + red[(3 * 8 + j) * HD_H + h]);
                atomicAdd(dw + ((long long)(o0 + oo) * Tc + t) * HD_H + h, v);
            }
        }
    }
}

// dbias[o] = sum_r dout[r, o]
__global__ void __launch_bounds__(256) head_dbias_kernel(const float* __restrict__ dout, long long rows, int O, float* __restrict__ dbias) {
    const long long n = rows * O;
    // thread -> fixed column class: element index i = base + k*stride with stride a multiple of O keeps i % O constant
    const long long stride = (long long)gridDim.x * blockDim.x / O * O;
    const long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (stride == 0 || base >= stride) return;
    float s = 0.f;
    for (long long i = base; i < n; i += stride) s += dout[i];
    atomicAdd(dbias + base % O, s);
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(long long n4, DropP dp, float* __restrict__ mult) {
    drop_resolve(dp);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(mult)[i] = drop_mult4(dp, (unsigned long long)i * 4);
}

__global__ void step_tick_kernel(unsigned long long* seed_dev, long long* step_dev) {
    if (threadIdx.x != 0) return;
    if (step_dev) *step_dev += 1;
    if (seed_dev) *seed_dev += 0x9E3779B97F4A7C15ull;   // Weyl sequence: 2^64 distinct keys; Philox decorrelates neighbouring keys
}

DropP make_drop(float p, unsigned long long seed, const unsigned long long* seed_dev = nullptr) {
    DropP d;
    d.seed = seed;
    d.seed_dev = seed_dev;
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

static int adam_clip_step_impl(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                     const double* sumsq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                                     float eps, float weight_decay, long long step, const long long* step_dev, const float* lr_dev,
                                     int write_grad, float* norm_out, void* stream) {
    TS_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0, "null pointer or negative length");
    TS_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), "buffers must be 16-byte aligned");
    TS_REQUIRE(step_dev || step >= 1, "step counts from 1 (torch.optim.Adam increments before the update)");
    TS_REQUIRE(max_norm <= 0.f || sumsq, "clipping needs the sum of squares from matgcn_grad_sumsq");
    if (n == 0) return 0;
    AdamP a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    a.grad_scale = grad_scale;
    a.bc1 = step_dev ? 1.f : (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = step_dev ? 1.f : (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.write_grad = write_grad;
    a.step_dev = step_dev;
    a.lr_dev = lr_dev;
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = 8LL * sm_count_ts();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_clip_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, sumsq, a, norm_out);
    matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_adam_clip_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                     const double* sumsq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                                     float eps, float weight_decay, long long step, int write_grad, float* norm_out, void* stream) {
    return adam_clip_step_impl(param, grad, exp_avg, exp_avg_sq, n, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, weight_decay,
                               step, nullptr, nullptr, write_grad, norm_out, stream);
}

extern "C" int matgcn_adam_clip_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                         const double* sumsq, float max_norm, float grad_scale, const float* lr_dev, float beta1,
                                         float beta2, float eps, float weight_decay, const long long* step_dev, int write_grad,
                                         float* norm_out, void* stream) {
    TS_REQUIRE(lr_dev && step_dev, "device-resident learning rate and step count required");
    return adam_clip_step_impl(param, grad, exp_avg, exp_avg_sq, n, sumsq, max_norm, grad_scale, 0.f, beta1, beta2, eps, weight_decay,
                               0, step_dev, lr_dev, write_grad, norm_out, stream);
}

// Start of a captured train step: the Adam step count advances by one and the dropout key moves to the next value of a
// Weyl sequence, both in device memory, so that every replay of the same CUDA graph is a NEW step.
extern "C" int matgcn_step_tick(unsigned long long* seed_dev, long long* step_dev, void* stream) {
    TS_REQUIRE(seed_dev || step_dev, "null pointer");
    step_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(seed_dev, step_dev);
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

static int head_fwd_impl(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, const float* bias,
                               int O, float p_drop, unsigned long long seed, const unsigned long long* seed_dev, float* out, void* stream) {
    TS_REQUIRE(y && w && bias && out, "null pointer");
    TS_REQUIRE(Tc > 0 && rows >= 0 && O > 0 && p_drop >= 0.f && p_drop < 1.f, "bad dimensions or dropout probability");
    TS_REQUIRE(H == HD_H, "the head kernels are written for rnn_units = 64");
    TS_REQUIRE(aligned16(y) && aligned16(w) && !(y_tstride & 3), "y and w must be 16-byte aligned");
    if (rows == 0) return 0;
    const DropP dp = make_drop(p_drop, seed, seed_dev);
    const unsigned grid = (unsigned)((rows + HD_RT - 1) / HD_RT);
    cudaStream_t st = (cudaStream_t)stream;
    for (int o0 = 0; o0 < O; o0 += HD_OMAX) {
        const int ow = O - o0 < HD_OMAX ? O - o0 : HD_OMAX;
        if (ow <= 4) head_fwd_kernel<4><<<grid, HD_THREADS, 0, st>>>(y, y_tstride, Tc, rows, w, bias, O, o0, dp, out);
        else if (ow <= 12) head_fwd_kernel<12><<<grid, HD_THREADS, 0, st>>>(y, y_tstride, Tc, rows, w, bias, O, o0, dp, out);
        else head_fwd_kernel<24><<<grid, HD_THREADS, 0, st>>>(y, y_tstride, Tc, rows, w, bias, O, o0, dp, out);
        matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_head_fwd(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, const float* bias,
                               int O, float p_drop, unsigned long long seed, float* out, void* stream) {
    return head_fwd_impl(y, y_tstride, Tc, rows, H, w, bias, O, p_drop, seed, nullptr, out, stream);
}
extern "C" int matgcn_head_fwd_dev(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, const float* bias,
                                   int O, float p_drop, unsigned long long seed, const unsigned long long* seed_dev, float* out,
                                   void* stream) {
    TS_REQUIRE(seed_dev, "null device seed");
    return head_fwd_impl(y, y_tstride, Tc, rows, H, w, bias, O, p_drop, seed, seed_dev, out, stream);
}

static int head_bwd_impl(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, int O, float p_drop,
                               unsigned long long seed, const unsigned long long* seed_dev, const float* dout, float* dy, float* dw,
                               float* dbias, void* stream) {
    TS_REQUIRE(y && w && dout && dy && dw && dbias, "null pointer");
    TS_REQUIRE(Tc > 0 && rows >= 0 && O > 0 && p_drop >= 0.f && p_drop < 1.f, "bad dimensions or dropout probability");
    TS_REQUIRE(H == HD_H, "the head kernels are written for rnn_units = 64");
    TS_REQUIRE(aligned16(y) && aligned16(w) && aligned16(dy) && !(y_tstride & 3), "y, w and dy must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TS_CK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)O * Tc * H, st));
    TS_CK(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)O, st));
    if (rows == 0) return 0;
    const DropP dp = make_drop(p_drop, seed, seed_dev);
    const long long ntiles = (rows + HD_RT - 1) / HD_RT;
    long long chunks = (3LL * sm_count_ts() + Tc - 1) / Tc;      // ~3 resident blocks per SM in total
    if (chunks > ntiles) chunks = ntiles;
    if (chunks < 1) chunks = 1;
    dim3 grid((unsigned)chunks, (unsigned)Tc);
    for (int o0 = 0; o0 < O; o0 += HD_OMAX) {
        const int ow = O - o0 < HD_OMAX ? O - o0 : HD_OMAX;
        const int accum = o0 > 0 ? 1 : 0;
        if (ow <= 4) head_bwd_kernel<4><<<grid, HD_THREADS, 0, st>>>(y, y_tstride, Tc, rows, w, O, o0, dp, dout, dy, accum, dw);
        else if (ow <= 12) head_bwd_kernel<12><<<grid, HD_THREADS, 0, st>>>(y, y_tstride, Tc, rows, w, O, o0, dp, dout, dy, accum, dw);
        else head_bwd_kernel<24><<<grid, HD_THREADS, 0, st>>>(y, y_tstride, Tc, rows, w, O, o0, dp, dout, dy, accum, dw);
        matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    {
        long long blocks = (rows * O + 256 * 64 - 1) / (256 * 64);
        if (blocks > 2LL * sm_count_ts()) blocks = 2LL * sm_count_ts();
        if (blocks * 256 < O) blocks = (O + 255) / 256;
        head_dbias_kernel<<<(unsigned)blocks, 256, 0, st>>>(dout, rows, O, dbias);
        matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_head_bwd(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, int O, float p_drop,
                               unsigned long long seed, const float* dout, float* dy, float* dw, float* dbias, void* stream) {
    return head_bwd_impl(y, y_tstride, Tc, rows, H, w, O, p_drop, seed, nullptr, dout, dy, dw, dbias, stream);
}
extern "C" int matgcn_head_bwd_dev(const float* y, long long y_tstride, int Tc, long long rows, int H, const float* w, int O, float p_drop,
                                   unsigned long long seed, const unsigned long long* seed_dev, const float* dout, float* dy, float* dw,
                                   float* dbias, void* stream) {
    TS_REQUIRE(seed_dev, "null device seed");
    return head_bwd_impl(y, y_tstride, Tc, rows, H, w, O, p_drop, seed, seed_dev, dout, dy, dw, dbias, stream);
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

// ------------------------------------------------------------------------------------------------------------------------
//  f3 (second half)  calculate_loss (MA.py:422-427): StandardScaler.inverse_transform of forecast and target
//  (normalization.py:62-76) followed by masked_mae_torch(pred, true, null_val = 0) (loss.py:17-29):
//      labels[|labels| < min_s] = 0 ; mask = labels != 0 ; mask /= mean(mask) ; loss = mean(|pred - labels| * mask), NaN -> 0
//  = sum(|pred - labels| over the unmasked elements) / count(unmasked).  One streaming pass over the two [B, T_out, N, C]
//  tensors (given with their element strides: the forecast is a permuted view of the head's output, the target a channel
//  slice of the batch) accumulates both sums in fp64; the last block to finish divides them.  The backward is one more pass:
//      d pred = g * std * sign(pred - labels) * [unmasked] / count.
// ------------------------------------------------------------------------------------------------------------------------
namespace {
struct Loss4 {
    int d1, d2, d3;                 // sizes of dimensions 1..3 (dimension 0 follows from n)
    long long ps[4], ys[4];         // element strides of the forecast and of the target
};
__device__ __forceinline__ void loss_offsets(const Loss4& g, long long i, long long& po, long long& yo) {
    const long long i3 = i % g.d3, r2 = i / g.d3;
    const long long i2 = r2 % g.d2, r1 = r2 / g.d2;
    const long long i1 = r1 % g.d1, i0 = r1 / g.d1;
    po = i0 * g.ps[0] + i1 * g.ps[1] + i2 * g.ps[2] + i3 * g.ps[3];
    yo = i0 * g.ys[0] + i1 * g.ys[1] + i2 * g.ys[2] + i3 * g.ys[3];
}
// acc: [0] sum of |pred - label| over unmasked elements, [1] their count, [2] (as unsigned) blocks done
__global__ void masked_mae_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ y, Loss4 g, long long n, float mean, float sd,
                                      float min_s, double* __restrict__ acc, float* __restrict__ loss) {
    double s = 0.0, c = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long po, yo;
        loss_offsets(g, i, po, yo);
        float l = y[yo] * sd + mean;
        const float p = pred[po] * sd + mean;
        if (fabsf(l) < min_s) l = 0.f;
        if (l != 0.f) {   // (a NaN label compares unequal to 0: it counts, and its NaN difference becomes 0 below, as in loss.py:26-28)
            c += 1.0;
            const float d = fabsf(p - l);
            if (d == d) s += (double)d;
        }
    }
    __shared__ double ss[32], sc[32];
    __shared__ bool last;
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bs = 0.0, bc = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { bs += ss[w]; bc += sc[w]; }
        atomicAdd(acc, bs);
        atomicAdd(acc + 1, bc);
        __threadfence();
        const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(acc + 2), 1u);
        last = done + 1 == gridDim.x;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const double ts = *reinterpret_cast<volatile double*>(acc), tc = *reinterpret_cast<volatile double*>(acc + 1);
        *loss = tc > 0.0 ? (float)(ts / tc) : 0.f;   // no unmasked element: mask / mean(mask) is NaN -> 0 (loss.py:25)
    }
}
__global__ void masked_mae_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ y, Loss4 g, long long n, float mean, float sd,
                                      float min_s, const double* __restrict__ acc, const float* __restrict__ gout, float* __restrict__ dpred) {
    const double cnt = acc[1];
    const float k = cnt > 0.0 ? (float)((double)(*gout) * (double)sd / cnt) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long po, yo;
        loss_offsets(g, i, po, yo);
        float l = y[yo] * sd + mean;
        const float p = pred[po] * sd + mean;
        if (fabsf(l) < min_s) l = 0.f;
        const float d = p - l;
        dpred[i] = (l != 0.f && d == d) ? (d > 0.f ? k : (d < 0.f ? -k : 0.f)) : 0.f;
    }
}
}  // namespace

static bool loss_geometry(const long long* sizes, const long long* pstr, const long long* ystr, Loss4& g, long long& n) {
    for (int i = 0; i < 4; ++i)
        if (sizes[i] <= 0 || sizes[i] > 2147483647LL) return false;
    g.d1 = (int)sizes[1]; g.d2 = (int)sizes[2]; g.d3 = (int)sizes[3];
    for (int i = 0; i < 4; ++i) { g.ps[i] = pstr[i]; g.ys[i] = ystr[i]; }
    n = sizes[0] * sizes[1] * sizes[2] * sizes[3];
    return true;
}

extern "C" int matgcn_masked_mae_fwd(const float* pred, const float* y, const long long* sizes, const long long* pred_strides,
                                     const long long* y_strides, float mean, float std, float min_s, double* acc, float* loss,
                                     void* stream) {
    TS_REQUIRE(pred && y && sizes && pred_strides && y_strides && acc && loss, "null pointer");
    Loss4 g;
    long long n;
    TS_REQUIRE(loss_geometry(sizes, pred_strides, y_strides, g, n), "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    TS_CK(cudaMemsetAsync(acc, 0, 3 * sizeof(double), st));
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 4LL * sm_count_ts()) blocks = 4LL * sm_count_ts();
    if (blocks < 1) blocks = 1;
    masked_mae_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(pred, y, g, n, mean, std, min_s, acc, loss);
    matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    TS_CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_masked_mae_bwd(const float* pred, const float* y, const long long* sizes, const long long* pred_strides,
                                     const long long* y_strides, float mean, float std, float min_s, const double* acc,
                                     const float* grad_loss, float* dpred, void* stream) {
    TS_REQUIRE(pred && y && sizes && pred_strides && y_strides && acc && grad_loss && dpred, "null pointer");
    Loss4 g;
    long long n;
    TS_REQUIRE(loss_geometry(sizes, pred_strides, y_strides, g, n), "bad sizes");
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 4LL * sm_count_ts()) blocks = 4LL * sm_count_ts();
    if (blocks < 1) blocks = 1;
    masked_mae_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, y, g, n, mean, std, min_s, acc, grad_loss, dpred);
    matgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
    TS_CK(cudaGetLastError());
    return 0;
}
