// C-ABI entry points of include/matgcn.h: fp32 ("exact mode") path.
//
// Each entry point is a host function that enqueues a fixed sequence of kernels on the
// caller's stream.  The reference lines each one replaces are cited in include/matgcn.h;
// the decomposition (hoisted supports / weights, time-batched input half, reverse-time BPTT,
// time-batched parameter gradients) is described in DESIGN.md section 3.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/matgcn.h"
#include "gemm_simt.cuh"
#include "epilogues.cuh"
#include "gemm_tc.cuh"
#include "rec_api.h"
#include "res_bwd.cuh"
#include "xside_mma.cuh"

using namespace matgcn;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(const char* where, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, what);
    return -1;
}
// error sink of the library's other translation units (train_step.cu); not part of the public header
extern "C" void matgcn_internal_set_error(const char* where, const char* what) { snprintf(g_err, sizeof(g_err), "%s: %s", where, what); }
#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) return fail(__func__, cudaGetErrorString(e_));          \
    } while (0)
#define REQUIRE(cond, msg)                          \
    do {                                            \
        if (!(cond)) return fail(__func__, msg);    \
    } while (0)

// Engine dispatch: fast mode (flags & MATGCN_FLAG_TF32) sends a contraction to the tcgen05/TMA kernel when
// its operands meet the TMA alignment rules, otherwise (and always in exact mode) to the fp32 SIMT kernel.
static std::atomic<unsigned long long> g_tc_launches{0};
// L2 warm-up of the next launch's weight blocks (TcP::pf_*); MATGCN_L2_WARM=0 switches it off for A/B measurements
static int l2_warm_level();
static void add_pf(GemmP& p, const void* base, long long stride_bytes, long long chunk_bytes, int n) {
    if (!base || p.npf >= 8 || chunk_bytes < 128 || chunk_bytes > 0x7fffff00LL || n <= 0) return;
    // budget: what one launch pulls in must sit in L2 next to its own working set (126 MB L2; large shapes simply skip ranges)
    long long have = chunk_bytes * n;
    for (int r = 0; r < p.npf; ++r) have += (long long)p.pf[r].chunk * p.pf[r].n;
    if (have > (64LL << 20)) return;
    p.pf[p.npf++] = PfRange{reinterpret_cast<const char*>(base), stride_bytes, (int)(chunk_bytes & ~127LL), n};
}
static bool l2_warm_enabled() {
    return l2_warm_level() > 0;
}
// Warm-up pays where (a) a step's working set - the bf16 weight twins it streams plus ~20 [N,B,H] fp32 blocks of saved
// activations, pre-activations and gradients - clearly exceeds L2, so that the next launch would otherwise start on HBM
// latency (measured at N=403, B=64: -2.4 % step time with both layers warmed; at N=237, where the set is about L2-sized,
// +2 %), and (b) the largest block to pull in, the gate weights' hidden rows, fits the per-launch budget (N=883: +0.6 % when
// only parts fit).  MATGCN_L2_WARM=3 forces it on, 0 switches it off, 1 restricts it to the weight blocks.
// Fused reduction pass over DR (dr_pass.cuh).  MATGCN_DR_PASS=0 or matgcn_set_dr_pass(0) keeps the separate split-K contractions.
static int& dr_pass_flag() {
    static int f = []() { const char* e = getenv("MATGCN_DR_PASS"); return (e && e[0] == '0') ? 0 : 1; }();
    return f;
}
static bool dr_pass_enabled() { return dr_pass_flag() != 0; }
static bool dg32_forced() {
    static const bool on = []() { const char* e = getenv("MATGCN_DG32"); return e && e[0] == '1'; }();
    return on;
}
extern "C" int matgcn_set_dr_pass(int on) {
    const int prev = dr_pass_flag();
    dr_pass_flag() = on ? 1 : 0;
    return prev;
}
static bool l2_warm_layer(int N, int K, int Cin, int H, int B) {
    if (!l2_warm_enabled()) return false;
    if (l2_warm_level() >= 3) return true;
    static const long long l2 = []() {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev);
        return v > 0 ? (long long)v : (126LL << 20);
    }();
    const long long twins = (long long)N * K * (Cin + H) * 3 * H * 2;      // Wg16 + Wu16 bytes
    const long long acts = 20LL * N * B * H * 4;
    const long long gate_rows = (long long)N * K * H * 2 * H * 2;          // Wg16[n, k, Cin:, :] bytes
    return twins + acts > l2 + l2 / 4 && gate_rows <= (64LL << 20);
}
// level 1: weight blocks only; level 2 (default): also the saved activations / pre-activations the next launches read
static int l2_warm_level() {
    static const int lvl = []() { const char* e = getenv("MATGCN_L2_WARM"); return (e && e[0] >= '0' && e[0] <= '9') ? e[0] - '0' : 2; }();
    return lvl;
}
template <class Cfg, bool A_KC, bool B_KC, class Epi>
static cudaError_t gemm_any(bool tc, const GemmP& p, const Epi& epi, int Z, cudaStream_t st) {
    if (tc && p.K >= 8 && p.A16 && p.B16) {
        // bf16 twins of both operands are available: bf16 MMAs (half the operand traffic, twice the MMA rate)
        cudaError_t e = (p.N <= 64) ? launch_gemm_tc<64, A_KC, B_KC, Epi, true>(p, epi, Z, st)
                                    : launch_gemm_tc<128, A_KC, B_KC, Epi, true>(p, epi, Z, st);
        if (e == cudaSuccess) g_tc_launches.fetch_add(1, std::memory_order_relaxed);
        if (e != cudaErrorNotSupported) return e;
    }
    if (p.need16) return cudaErrorNotSupported;  // no fallback may read the fp32 operands: fail loudly
    if (tc && p.K >= 8) {
        cudaError_t e = (p.N <= 64) ? launch_gemm_tc<64, A_KC, B_KC, Epi>(p, epi, Z, st)
                                    : launch_gemm_tc<128, A_KC, B_KC, Epi>(p, epi, Z, st);
        if (e == cudaSuccess) g_tc_launches.fetch_add(1, std::memory_order_relaxed);
        if (e != cudaErrorNotSupported) return e;
    }
    return launch_gemm<Cfg, A_KC, B_KC, Epi>(p, epi, Z, st);
}
// Tensor-core launch with EpiStore's feature set fixed at compile time (EpiStoreT<F>, epilogues.cuh) for the time-batched
// launches whose run time is their epilogue.  cudaErrorNotSupported: the functor uses other features than F, or the shape does
// not qualify - the caller takes the generic route.  MATGCN_LEAN_EPI=0 switches it off (A/B measurements).
static bool lean_epi_enabled() {
    static const bool on = []() { const char* e = getenv("MATGCN_LEAN_EPI"); return !(e && e[0] == '0'); }();
    return on;
}
template <int F, int BN, bool A_KC, bool B_KC, bool BF16>
static cudaError_t launch_store_lean(const GemmP& p, const EpiStore& e, int Z, cudaStream_t st) {
    if (!lean_epi_enabled() || !EpiStoreT<F>::matches(e) || p.K < 8) return cudaErrorNotSupported;
    if (BF16 ? !(p.A16 && p.B16) : !(p.A && p.B)) return cudaErrorNotSupported;
    const cudaError_t le = launch_gemm_tc<BN, A_KC, B_KC, EpiStoreT<F>, BF16>(p, EpiStoreT<F>(e), Z, st);
    if (le == cudaSuccess) g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return le;
}
extern "C" unsigned long long matgcn_tc_launch_count(void) { return g_tc_launches.load(); }

// Optional per-launch tracing (MATGCN_TRACE=1): a CUDA event after every enqueued operation of the encoder
// entry points; on exit the entry point synchronises and prints the time spent between consecutive events,
// aggregated by source line, to stderr.  Diagnostics only - it serialises host and device.
#include <map>
#include <string>
#include <vector>
struct Tracer {
    bool on;
    cudaStream_t st;
    std::vector<std::pair<int, cudaEvent_t>> ev;
    explicit Tracer(cudaStream_t s) : st(s) {
        static const bool enabled = []() { const char* e = getenv("MATGCN_TRACE"); return e && e[0] == '1'; }();
        on = enabled;
        if (on) mark(0);
    }
    void mark(int line) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back({line, e});
    }
    void report(const char* what) {
        if (!on || ev.size() < 2) return;
        cudaEventSynchronize(ev.back().second);
        std::map<int, std::pair<int, float>> agg;
        float total = 0.f;
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
            agg[ev[i].first].first++;
            agg[ev[i].first].second += ms;
            total += ms;
        }
        fprintf(stderr, "[matgcn trace] %s: total %.3f ms\n", what, total);
        for (auto& kv : agg)
            fprintf(stderr, "[matgcn trace]   line %4d: %4d x  %8.3f ms total  %7.1f us avg\n", kv.first, kv.second.first,
                    kv.second.second, 1e3f * kv.second.second / kv.second.first);
        for (auto& e : ev) cudaEventDestroy(e.second);
        ev.clear();
    }
};
#define TR() tr.mark(__LINE__)

extern "C" int matgcn_abi_version(void) { return MATGCN_ABI_VERSION; }
extern "C" const char* matgcn_last_error(void) { return g_err; }
extern "C" unsigned long long matgcn_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = sh[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
    return r;
}

// one CTA per row n: A[n, :] = softmax_m(relu(L[n] . Rt[m]))
__global__ void adaptive_adj_fwd_kernel(const float* __restrict__ L, const float* __restrict__ Rt, int N, int D,
                                        float* __restrict__ A, int ldm) {
    extern __shared__ float sm[];
    float* row = sm;          // N
    float* lrow = sm + N;     // D
    __shared__ float red[32];
    const int n = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) lrow[d] = L[(long long)n * D + d];
    __syncthreads();
    float mx = 0.f;  // relu output is >= 0
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        const float* r = Rt + (long long)m * D;
        float s = 0.f;
        for (int d = 0; d < D; ++d) s = fmaf(lrow[d], __ldg(r + d), s);
        s = fmaxf(s, 0.f);
        row[m] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        const float e = expf(row[m] - mx);
        row[m] = e;
        sum += e;
    }
    sum = block_reduce(sum, red, false);
    const float inv = 1.f / sum;
    for (int m = threadIdx.x; m < ldm; m += blockDim.x) A[(long long)n * ldm + m] = m < N ? row[m] * inv : 0.f;
}

// one CTA per row n: dpre[n, m] = [L[n].Rt[m] > 0] * A[n,m] * (dA[n,m] - sum_j dA[n,j] A[n,j])
__global__ void adaptive_adj_bwd_kernel(const float* __restrict__ L, const float* __restrict__ Rt,
                                        const float* __restrict__ A, const float* __restrict__ dA, int N, int D,
                                        int ldm, float* __restrict__ dpre) {
    extern __shared__ float sm[];
    float* lrow = sm;  // D
    __shared__ float red[32];
    const int n = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) lrow[d] = L[(long long)n * D + d];
    float dot = 0.f;
    for (int m = threadIdx.x; m < N; m += blockDim.x) dot += A[(long long)n * ldm + m] * dA[(long long)n * ldm + m];
    dot = block_reduce(dot, red, false);  // also orders the lrow writes before the reads below
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        const float* r = Rt + (long long)m * D;
        float s = 0.f;
        for (int d = 0; d < D; ++d) s = fmaf(lrow[d], __ldg(r + d), s);
        const float a = A[(long long)n * ldm + m];
        dpre[(long long)n * N + m] = s > 0.f ? a * (dA[(long long)n * ldm + m] - dot) : 0.f;
    }
}

// out[i] = a[i] * c[(i / div) % K]   (pool scaled by its view weight)
__global__ void scale_groups_kernel(const float* __restrict__ a, const float* __restrict__ c, long long n, int div,
                                    int K, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * __ldg(c + (i / div) % K);
}

// dc[k] = sum_{d,i,o} dpool[d,k,i,o] * pool[d,k,i,o] / c[k]     grid (K, chunks)
__global__ void view_weight_grad_kernel(const float* __restrict__ dpool, const float* __restrict__ pool,
                                        const float* __restrict__ c, int D, int K, int IO, float* __restrict__ dc) {
    __shared__ float red[32];
    const int k = blockIdx.x;
    const long long total = (long long)D * IO;
    float s = 0.f;
    for (long long j = (long long)blockIdx.y * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.y * blockDim.x) {
        const long long d = j / IO, io = j - d * IO;
        const long long idx = (d * K + k) * IO + io;
        s += dpool[idx] * pool[idx];
    }
    s = block_reduce(s, red, false);
    if (threadIdx.x == 0) atomicAdd(dc + k, s / c[k]);
}

// Column sums of a [T][Z][rows][ld] array over (t, rows) for the first ncol columns; columns < split go to
// out1[z*ld1 + col], the rest to out2[z*ld2 + col - split].   grid (Z, chunks), partial sums by atomics.
// Vectorised column sums for the bias gradients: src[t*st + z*sz + r*ncol + c] summed over (t, r) for every block z
// (rows are dense: ld == ncol, ncol % 4 == 0, 16-byte aligned).  grid (Z, S): block (z, y) takes the 64-row chunks
// y, y+S, ... of the (t, chunk) space; thread -> (row lane, float4 column); four independent 16-byte loads in flight.
__global__ void __launch_bounds__(256) colsum4_kernel(const float* __restrict__ src, int T, long long st, long long sz, int rows,
                                                      int ncol, int split, float* __restrict__ out1, int ld1,
                                                      float* __restrict__ out2, int ld2) {
    extern __shared__ float4 cs_red[];  // [lanes][ncol / 4]
    const int nc4 = ncol >> 2;
    const int lanes = blockDim.x / nc4;
    const int c4 = threadIdx.x % nc4, lane = threadIdx.x / nc4;
    const int z = blockIdx.x;
    const int chunks_per_t = (rows + 63) >> 6;
    const int nchunks = T * chunks_per_t;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes) {
        for (int ch = blockIdx.y; ch < nchunks; ch += gridDim.y) {
            const int t = ch / chunks_per_t, r0 = (ch - t * chunks_per_t) << 6;
            const int r1 = min(rows, r0 + 64);
            const float* base = src + (long long)t * st + (long long)z * sz + 4 * c4;
            for (int r = r0 + lane; r < r1; r += 4 * lanes) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ru = r + u * lanes;
                    v[u] = ru < r1 ? *reinterpret_cast<const float4*>(base + (long long)ru * ncol) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
            }
        }
        cs_red[lane * nc4 + c4] = s;
    }
    __syncthreads();
    if (lane == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 o = cs_red[l * nc4 + c4];
            s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        }
        const float vals[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int col = 4 * c4 + u;
            if (col < split) atomicAdd(out1 + (long long)z * ld1 + col, vals[u]);
            else atomicAdd(out2 + (long long)z * ld2 + col - split, vals[u]);
        }
    }
}
// The same over a bf16 array (the bf16 twin of DG when the persistent reverse kernel wrote no fp32 copy): eight columns per thread.
__global__ void __launch_bounds__(256) colsum8_bf16_kernel(const __nv_bfloat16* __restrict__ src, int T, long long st, long long sz, int rows,
                                                           int ncol, int split, float* __restrict__ out1, int ld1,
                                                           float* __restrict__ out2, int ld2) {
    extern __shared__ float4 cs_red[];  // [lanes][ncol / 8][2]
    const int nc8 = ncol >> 3;
    const int lanes = blockDim.x / nc8;
    const int c8 = threadIdx.x % nc8, lane = threadIdx.x / nc8;
    const int z = blockIdx.x;
    const int chunks_per_t = (rows + 63) >> 6;
    const int nchunks = T * chunks_per_t;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (lane < lanes) {
        for (int ch = blockIdx.y; ch < nchunks; ch += gridDim.y) {
            const int t = ch / chunks_per_t, r0 = (ch - t * chunks_per_t) << 6;
            const int r1 = min(rows, r0 + 64);
            const __nv_bfloat16* base = src + (long long)t * st + (long long)z * sz + 8 * c8;
            for (int r = r0 + lane; r < r1; r += 4 * lanes) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ru = r + u * lanes;
                    v[u] = ru < r1 ? *reinterpret_cast<const uint4*>(base + (long long)ru * ncol) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {   // bf16 -> fp32: the 16 bits are the high half of the float
                        s[2 * j] += __uint_as_float(w4[j] << 16);
                        s[2 * j + 1] += __uint_as_float(w4[j] & 0xffff0000u);
                    }
                }
            }
        }
        cs_red[(lane * nc8 + c8) * 2] = make_float4(s[0], s[1], s[2], s[3]);
        cs_red[(lane * nc8 + c8) * 2 + 1] = make_float4(s[4], s[5], s[6], s[7]);
    }
    __syncthreads();
    if (lane == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 a = cs_red[(l * nc8 + c8) * 2], b = cs_red[(l * nc8 + c8) * 2 + 1];
            s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w; s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int col = 8 * c8 + u;
            if (col < split) atomicAdd(out1 + (long long)z * ld1 + col, s[u]);
            else atomicAdd(out2 + (long long)z * ld2 + col - split, s[u]);
        }
    }
}
static bool colsum4_ok(const float* src, long long st, long long sz, int ncol) {
    return aligned16(src) && !(st & 3) && !(sz & 3) && !(ncol & 3) && ncol / 4 <= 256;
}
__global__ void colsum_kernel(const float* __restrict__ src, int T, long long st, long long sz, int rows, int ld,
                              int ncol, int split, float* __restrict__ out1, int ld1, float* __restrict__ out2, int ld2) {
    const int z = blockIdx.x;
    const int lanes = blockDim.x / ncol;  // row lanes per block
    const int col = threadIdx.x % ncol, lane = threadIdx.x / ncol;
    if (lane >= lanes) return;
    const long long total = (long long)T * rows;
    float s = 0.f;
    long long j = (long long)blockIdx.y * lanes + lane;
    const long long step = (long long)gridDim.y * lanes;
    for (; j < total; j += 8 * step) {  // eight independent loads in flight per thread
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long ju = j + u * step;
            if (ju < total) {
                const long long t = ju / rows, r = ju - t * rows;
                v[u] = src[t * st + z * sz + r * ld + col];
            } else {
                v[u] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    if (col < split) atomicAdd(out1 + (long long)z * ld1 + col, s);
    else atomicAdd(out2 + (long long)z * ld2 + col - split, s);
}

// B0: head of the reverse step.  dy = dY[t] + carry ; residual-mix backward up to da3.
// 16-byte form of the kernel below (H % 4 == 0, n % 4 == 0, aligned pointers): one float4 per thread and stream
__global__ void __launch_bounds__(256) bwd_head4_kernel(const float* __restrict__ dY, const float* __restrict__ carry,
                                                        const float* __restrict__ H1, const float* __restrict__ R2,
                                                        const float* __restrict__ HC2, const float* __restrict__ mix_t, long long n4,
                                                        int H, float* __restrict__ DH1, float* __restrict__ DRES,
                                                        float* __restrict__ DR, float* __restrict__ dmix_t) {
    __shared__ float red[32];
    const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (i4 < n4) {
        const long long i = 4 * i4;
        const float g = __ldg(mix_t);
        const float4 a = ld4(dY + i), b = ld4(carry + i), h1 = ld4(H1 + i), r2 = ld4(R2 + i), hc2 = ld4(HC2 + i);
        const float4 dy = a + b;
        const float4 res = r2 * h1 + one_minus(r2) * hc2;
        const float4 pr = dy * (h1 - res);
        part = (pr.x + pr.y) + (pr.z + pr.w);
        const float4 dres = (1.f - g) * dy;
        st4(DRES + i, dres);
        st4(DH1 + i, g * dy + dres * r2);
        const long long row = i / H;
        const int c = (int)(i - row * H);
        st4(DR + row * 3 * H + 2 * H + c, dres * one_minus(r2) * one_minus(hc2 * hc2));
    }
    part = block_reduce(part, red, false);
    if (threadIdx.x == 0) atomicAdd(dmix_t, part);
}
__global__ void bwd_head_kernel(const float* __restrict__ dY, const float* __restrict__ carry,
                                const float* __restrict__ H1, const float* __restrict__ R2,
                                const float* __restrict__ HC2, const float* __restrict__ mix_t, long long n, int H,
                                float* __restrict__ DH1, float* __restrict__ DRES, float* __restrict__ DR,
                                float* __restrict__ dmix_t) {
    __shared__ float red[32];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (i < n) {
        const float g = __ldg(mix_t);
        const float dy = dY[i] + carry[i];
        const float h1 = H1[i], r2 = R2[i], hc2 = HC2[i];
        const float res = r2 * h1 + (1.f - r2) * hc2;
        part = dy * (h1 - res);
        const float dres = (1.f - g) * dy;
        DRES[i] = dres;
        DH1[i] = g * dy + dres * r2;
        const long long row = i / H;
        const int c = (int)(i - row * H);
        DR[row * 3 * H + 2 * H + c] = dres * (1.f - r2) * (1.f - hc2 * hc2);
    }
    part = block_reduce(part, red, false);
    if (threadIdx.x == 0) atomicAdd(dmix_t, part);
}


// ------------------------------------------------------------------------------------------
// Input-side kernels for a tiny channel count (layer 0: Cin = 2).  With K*Cin <= 10 the contractions over the
// input rows are a handful of FMAs per output element, so GEMM tiles would be >90% padding; these kernels
// stream the big [T,N,B,3H] arrays exactly once instead.
// ------------------------------------------------------------------------------------------
constexpr int XS_KC = 10;  // max K*Cin
constexpr int XS_J = 6;    // max ceil(3H/32)
static bool xside_small_ok(int Cin, int H, int K) { return Cin <= 4 && K * Cin <= XS_KC && 3 * H <= 32 * XS_J; }

// GX[t,n,b,o] = bias3[n,o] + sum_{k,i} PX[t,k,n,b,i] * W3[n,k,i,o] ;  RX[t,n,b,o] = rbias3[o] + sum_i x[t,n,b,i] * Rx3[o,i]
// (W3 = [Wg | Wu] input rows, Rx3 = [Rgw ; Ruw][:, 0:Cin]).   grid (N, T), thread = output column o.
__global__ void xside_fwd_small_kernel(const float* __restrict__ PX, const float* __restrict__ Wg, const float* __restrict__ bg,
                                       const float* __restrict__ Wu, const float* __restrict__ bu,
                                       const float* __restrict__ Rgw, const float* __restrict__ Rgb,
                                       const float* __restrict__ Ruw, const float* __restrict__ Rub, int T, int N, int B, int Cin,
                                       int H, int K, float* __restrict__ GX, float* __restrict__ RX) {
    extern __shared__ float xs[];  // [K][B][Cin]
    const int n = blockIdx.x, t = blockIdx.y, o = threadIdx.x;
    const int I = Cin + H, KC = K * Cin;
    const long long UX = (long long)N * B * Cin;
    for (int j = threadIdx.x; j < K * B * Cin; j += blockDim.x) {
        const int k = j / (B * Cin), r = j - k * (B * Cin);
        xs[j] = PX[((long long)t * K + k) * UX + (long long)n * B * Cin + r];
    }
    float w[XS_KC], rw[4], bias = 0.f, rb = 0.f;
    int xo[XS_KC];  // smem offset of x[k][b=0][i] for each (k,i) pair: no index arithmetic in the inner loop
#pragma unroll
    for (int kc = 0; kc < XS_KC; ++kc) {
        const int k = kc / Cin, i = kc - k * Cin;
        xo[kc] = k * B * Cin + i;
        w[kc] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) rw[i] = 0.f;
    if (o < 3 * H) {
#pragma unroll
        for (int kc = 0; kc < XS_KC; ++kc)
            if (kc < KC) {
                const int k = kc / Cin, i = kc - k * Cin;
                w[kc] = o < 2 * H ? Wg[(((long long)n * K + k) * I + i) * 2 * H + o] : Wu[(((long long)n * K + k) * I + i) * H + o - 2 * H];
            } else {
                xo[kc] = 0;  // weight 0: any valid slot
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < Cin) rw[i] = o < 2 * H ? Rgw[(long long)o * I + i] : Ruw[(long long)(o - 2 * H) * I + i];
        bias = o < 2 * H ? bg[(long long)n * 2 * H + o] : bu[(long long)n * H + o - 2 * H];
        rb = o < 2 * H ? Rgb[o] : Rub[o - 2 * H];
    }
    __syncthreads();
    if (o >= 3 * H) return;
    float* gx = GX + (((long long)t * N + n) * B) * 3 * H + o;
    float* rx = RX + (((long long)t * N + n) * B) * 3 * H + o;
    const int cin_m1 = Cin - 1;
    for (int b = 0; b < B; ++b) {
        const float* xb = xs + b * Cin;
        float g = bias, r = rb;
#pragma unroll
        for (int kc = 0; kc < XS_KC; ++kc) g = fmaf(w[kc], xb[xo[kc]], g);
#pragma unroll
        for (int i = 0; i < 4; ++i) r = fmaf(rw[i], xb[min(i, cin_m1)], r);
        gx[(long long)b * 3 * H] = g;
        rx[(long long)b * 3 * H] = r;
    }
}

// The same with four output columns per thread (3H % 4 == 0, 16-byte aligned operands): the scalar form above issues ~35
// instructions per output element and is bound by instruction issue (80 % of the issue slots, 40 % of the DRAM bandwidth:
// profiles/r2h_xside_ncu.txt); here the 14 shared-memory reads of a row serve four columns and the stores are 16 bytes.
// grid (N, T), thread = (column quad q, row group rg of 4): rows b = rg, rg + 4, ...
constexpr int XS_RG = 4;
__global__ void __launch_bounds__(256) xside_fwd_small4_kernel(const float* __restrict__ PX, const float* __restrict__ Wg,
                                                               const float* __restrict__ bg, const float* __restrict__ Wu,
                                                               const float* __restrict__ bu, const float* __restrict__ Rgw,
                                                               const float* __restrict__ Rgb, const float* __restrict__ Ruw,
                                                               const float* __restrict__ Rub, int T, int N, int B, int Cin, int H, int K,
                                                               float* __restrict__ GX, float* __restrict__ RX) {
    extern __shared__ float xs[];  // [K][B][Cin]
    const int n = blockIdx.x, t = blockIdx.y;
    const int Q = 3 * H / 4;
    const int q = threadIdx.x % Q, rg = threadIdx.x / Q, o = 4 * q;
    const int I = Cin + H, KC = K * Cin;
    const long long UX = (long long)N * B * Cin;
    for (int j = threadIdx.x; j < K * B * Cin; j += blockDim.x) {
        const int k = j / (B * Cin), r = j - k * (B * Cin);
        xs[j] = PX[((long long)t * K + k) * UX + (long long)n * B * Cin + r];
    }
    const bool gate = o < 2 * H;
    float4 w[XS_KC], rw[4];
    int xo[XS_KC];
#pragma unroll
    for (int kc = 0; kc < XS_KC; ++kc) {
        const int k = kc / Cin, i = kc - k * Cin;
        if (kc < KC) {
            xo[kc] = k * B * Cin + i;
            w[kc] = gate ? ld4(Wg + (((long long)n * K + k) * I + i) * 2 * H + o) : ld4(Wu + (((long long)n * K + k) * I + i) * H + o - 2 * H);
        } else {
            xo[kc] = 0;  // weight 0: any valid slot
            w[kc] = f4(0.f);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rw[i] = f4(0.f);
        if (i < Cin) {
            const float* r0 = gate ? Rgw + (long long)o * I + i : Ruw + (long long)(o - 2 * H) * I + i;
            rw[i] = make_float4(r0[0], r0[I], r0[2 * I], r0[3 * I]);
        }
    }
    const float4 bias = gate ? ld4(bg + (long long)n * 2 * H + o) : ld4(bu + (long long)n * H + o - 2 * H);
    const float4 rb = gate ? ld4(Rgb + o) : ld4(Rub + o - 2 * H);
    __syncthreads();
    float* gx = GX + (((long long)t * N + n) * B) * 3 * H + o;
    float* rx = RX + (((long long)t * N + n) * B) * 3 * H + o;
    const int cin_m1 = Cin - 1;
    for (int b = rg; b < B; b += XS_RG) {
        const float* xb = xs + b * Cin;
        float4 g = bias, r = rb;
#pragma unroll
        for (int kc = 0; kc < XS_KC; ++kc) {
            const float x = xb[xo[kc]];
            g.x = fmaf(w[kc].x, x, g.x); g.y = fmaf(w[kc].y, x, g.y); g.z = fmaf(w[kc].z, x, g.z); g.w = fmaf(w[kc].w, x, g.w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x = xb[min(i, cin_m1)];
            r.x = fmaf(rw[i].x, x, r.x); r.y = fmaf(rw[i].y, x, r.y); r.z = fmaf(rw[i].z, x, r.z); r.w = fmaf(rw[i].w, x, r.w);
        }
        st4(gx + (long long)b * 3 * H, g);
        st4(rx + (long long)b * 3 * H, r);
    }
}

// One pass over DG[:, n]: weight gradient of the input rows, bias gradient, and the input-side data gradient
//   dW3[n,k,i,o] = sum_{t,b} PX[t,k,n,b,i] * DG[t,n,b,o]     db3[n,o] = sum_{t,b} DG[t,n,b,o]
//   DPX[t,k,n,b,i] = sum_o DG[t,n,b,o] * W3[n,k,i,o]
// grid (N), 8 warps; a warp takes every 8th (t,b) row, lane l owns columns l, l+32, ...
__global__ void __launch_bounds__(256) xside_bwd_dg_small_kernel(
    const float* __restrict__ PX, const float* __restrict__ DG, const float* __restrict__ Wg, const float* __restrict__ Wu, int T,
    int N, int B, int Cin, int H, int K, float* __restrict__ dWg, float* __restrict__ dWu, float* __restrict__ dbg,
    float* __restrict__ dbu, float* __restrict__ DPX) {
    __shared__ float w3[XS_KC][32 * XS_J];
    __shared__ float red[32 * XS_J][XS_KC + 1];
    const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int I = Cin + H, KC = K * Cin, H3 = 3 * H;
    const long long UX = (long long)N * B * Cin;
    for (int j = threadIdx.x; j < XS_KC * 32 * XS_J; j += blockDim.x) {
        const int kc = j / (32 * XS_J), o = j - kc * (32 * XS_J);
        float v = 0.f;
        if (kc < KC && o < H3) {
            const int k = kc / Cin, i = kc - k * Cin;
            v = o < 2 * H ? Wg[(((long long)n * K + k) * I + i) * 2 * H + o] : Wu[(((long long)n * K + k) * I + i) * H + o - 2 * H];
        }
        w3[kc][o] = v;
    }
    for (int j = threadIdx.x; j < 32 * XS_J * (XS_KC + 1); j += blockDim.x) (&red[0][0])[j] = 0.f;
    __syncthreads();
    float acc[XS_J][XS_KC], bsum[XS_J];
#pragma unroll
    for (int j = 0; j < XS_J; ++j) {
        bsum[j] = 0.f;
#pragma unroll
        for (int kc = 0; kc < XS_KC; ++kc) acc[j][kc] = 0.f;
    }
    const int rows = T * B;
    for (int r = warp; r < rows; r += 8) {
        const int t = r / B, b = r - t * B;
        const float* dgp = DG + (((long long)t * N + n) * B + b) * H3;
        float dg[XS_J];
#pragma unroll
        for (int j = 0; j < XS_J; ++j) dg[j] = (lane + 32 * j < H3) ? dgp[lane + 32 * j] : 0.f;
        float xv = 0.f;
        if (lane < KC) {
            const int k = lane / Cin, i = lane - k * Cin;
            xv = PX[((long long)t * K + k) * UX + ((long long)n * B + b) * Cin + i];
        }
        float part[XS_KC];
#pragma unroll
        for (int kc = 0; kc < XS_KC; ++kc) {
            const float x = __shfl_sync(0xffffffffu, xv, kc);
            float pp = 0.f;
#pragma unroll
            for (int j = 0; j < XS_J; ++j) {
                acc[j][kc] = fmaf(x, dg[j], acc[j][kc]);
                pp = fmaf(dg[j], w3[kc][lane + 32 * j], pp);
            }
            part[kc] = pp;
        }
#pragma unroll
        for (int j = 0; j < XS_J; ++j) bsum[j] += dg[j];
#pragma unroll
        for (int kc = 0; kc < XS_KC; ++kc) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part[kc] += __shfl_xor_sync(0xffffffffu, part[kc], off);
        }
        float mine = 0.f;
#pragma unroll
        for (int kc = 0; kc < XS_KC; ++kc)
            if (lane == kc) mine = part[kc];
        if (lane < KC) {
            const int k = lane / Cin, i = lane - k * Cin;
            DPX[((long long)t * K + k) * UX + ((long long)n * B + b) * Cin + i] = mine;
        }
    }
    // cross-warp reduction through shared memory, one warp at a time
    for (int wsel = 0; wsel < 8; ++wsel) {
        if (warp == wsel) {
#pragma unroll
            for (int j = 0; j < XS_J; ++j) {
#pragma unroll
                for (int kc = 0; kc < XS_KC; ++kc) red[lane + 32 * j][kc] += acc[j][kc];
                red[lane + 32 * j][XS_KC] += bsum[j];
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < H3 * (KC + 1); j += blockDim.x) {
        const int kc = j / H3, o = j - kc * H3;
        if (kc == KC) {
            const float v = red[o][XS_KC];
            if (o < 2 * H) dbg[(long long)n * 2 * H + o] = v;
            else dbu[(long long)n * H + o - 2 * H] = v;
        } else {
            const int k = kc / Cin, i = kc - k * Cin;
            const float v = red[o][kc];
            if (o < 2 * H) dWg[(((long long)n * K + k) * I + i) * 2 * H + o] = v;
            else dWu[(((long long)n * K + k) * I + i) * H + o - 2 * H] = v;
        }
    }
}

// One pass over DR (flat rows (t,n,b)): residual-GRU input-column weight gradient, bias gradient, and the
// residual path's share of the input gradient, added into DPX[t,0]:
//   dRx3[o,i] += sum_rows DR[row,o] * x[row,i]    drb3[o] += sum_rows DR[row,o]    DPX[t,0,n,b,i] += sum_o DR[row,o] * Rx3[o,i]
// grid (chunks), 8 warps, warp per row; outputs are pre-zeroed and receive one atomic per CTA and element.
__global__ void __launch_bounds__(256) xside_bwd_dr_small_kernel(
    const float* __restrict__ PX, const float* __restrict__ DR, const float* __restrict__ Rgw, const float* __restrict__ Ruw,
    int T, int N, int B, int Cin, int H, int K, float* __restrict__ dRgw, float* __restrict__ dRuw, float* __restrict__ dRgb,
    float* __restrict__ dRub, float* __restrict__ DPX) {
    __shared__ float rx3[4][32 * XS_J];
    __shared__ float red[32 * XS_J][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int I = Cin + H, H3 = 3 * H;
    const long long UX = (long long)N * B * Cin, NB = (long long)N * B;
    for (int j = threadIdx.x; j < 4 * 32 * XS_J; j += blockDim.x) {
        const int i = j / (32 * XS_J), o = j - i * (32 * XS_J);
        float v = 0.f;
        if (i < Cin && o < H3) v = o < 2 * H ? Rgw[(long long)o * I + i] : Ruw[(long long)(o - 2 * H) * I + i];
        rx3[i][o] = v;
    }
    for (int j = threadIdx.x; j < 32 * XS_J * 5; j += blockDim.x) (&red[0][0])[j] = 0.f;
    __syncthreads();
    float acc[XS_J][4], bsum[XS_J];
#pragma unroll
    for (int j = 0; j < XS_J; ++j) {
        bsum[j] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    }
    const long long rows = (long long)T * NB;
    float drn[XS_J], xvn = 0.f;
    auto fetch = [&](long long r) {   // loads of the warp's next row, issued before the arithmetic of the current one
        const long long t = r / NB, nb = r - t * NB;
        const float* drp = DR + r * H3;
#pragma unroll
        for (int j = 0; j < XS_J; ++j) drn[j] = (lane + 32 * j < H3) ? drp[lane + 32 * j] : 0.f;
        xvn = lane < Cin ? PX[t * K * UX + nb * Cin + lane] : 0.f;
    };
    if ((long long)blockIdx.x * 8 + warp < rows) fetch((long long)blockIdx.x * 8 + warp);
    for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
        const long long t = r / NB, nb = r - t * NB;
        float dr[XS_J];
#pragma unroll
        for (int j = 0; j < XS_J; ++j) dr[j] = drn[j];
        float* x0 = DPX + t * K * UX + nb * Cin;          // DPX[t,0,n,b,:]
        const float xv = xvn;
        if (r + (long long)gridDim.x * 8 < rows) fetch(r + (long long)gridDim.x * 8);
        float part[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x = __shfl_sync(0xffffffffu, xv, i);
            float pp = 0.f;
#pragma unroll
            for (int j = 0; j < XS_J; ++j) {
                acc[j][i] = fmaf(x, dr[j], acc[j][i]);
                pp = fmaf(dr[j], rx3[i][lane + 32 * j], pp);
            }
            part[i] = pp;
        }
#pragma unroll
        for (int j = 0; j < XS_J; ++j) bsum[j] += dr[j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part[i] += __shfl_xor_sync(0xffffffffu, part[i], off);
        }
        float mine = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (lane == i) mine = part[i];
        if (lane < Cin) x0[lane] += mine;
    }
    for (int wsel = 0; wsel < 8; ++wsel) {
        if (warp == wsel) {
#pragma unroll
            for (int j = 0; j < XS_J; ++j) {
#pragma unroll
                for (int i = 0; i < 4; ++i) red[lane + 32 * j][i] += acc[j][i];
                red[lane + 32 * j][4] += bsum[j];
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < H3 * (Cin + 1); j += blockDim.x) {
        const int i = j / H3, o = j - i * H3;
        if (i == Cin) {
            atomicAdd(o < 2 * H ? dRgb + o : dRub + (o - 2 * H), red[o][4]);
        } else {
            atomicAdd(o < 2 * H ? dRgw + (long long)o * I + i : dRuw + (long long)(o - 2 * H) * I + i, red[o][i]);
        }
    }
}


// fp32 -> bf16 twins (round to nearest even).  2-D form: `rows` blocks of `n` floats, source/destination pitches in elements.
__global__ void to_bf16_kernel(const float* __restrict__ src, long long spitch, __nv_bfloat16* __restrict__ dst, long long dpitch,
                               long long n, int rows) {
    const long long total = n * rows;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n, c = i - r * n;
        dst[r * dpitch + c] = __float2bfloat16_rn(src[r * spitch + c]);
    }
}
// contiguous form: 16-byte loads, 8-byte stores (n % 4 == 0, aligned pointers)
__global__ void to_bf16_vec_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        st4_bf16(dst + 4 * i, ld4(src + 4 * i));
}
static cudaError_t to_bf16(const float* src, long long spitch, __nv_bfloat16* dst, long long dpitch, long long n, int rows, cudaStream_t st) {
    if (rows == 1 && !(n & 3) && aligned16(src) && !(reinterpret_cast<uintptr_t>(dst) & 7)) {
        const long long n4 = n >> 2;
        long long blocks = (n4 + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        to_bf16_vec_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n4);
        count_launch();
        return cudaGetLastError();
    }
    const long long total = n * rows;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    to_bf16_kernel<<<blocks, 256, 0, st>>>(src, spitch, dst, dpitch, n, rows);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// 3xTF32 (exact mode on the tensor cores, dense propagations only): x = hi + lo with hi = x truncated to TF32 (10 explicit
// mantissa bits; the tensor core then reads it exactly) and lo = x - hi (exact in fp32, |lo| < 2^-10 |x|, read by the tensor core
// to 10 bits: relative error 2^-20 of x).  a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (the dropped lo*lo term is 2^-20 relative) as
// THREE k-batches of one TF32 contraction accumulated in fp32 in TMEM: operand slabs [hi, hi, lo] (pattern 0) against [hi, lo, hi]
// (pattern 1).  src: n contiguous floats (n % 4 == 0, 16-byte aligned); dst: three slabs `slab` floats apart.
// ------------------------------------------------------------------------------------------
// (rows x n floats, source row pitch `spitch`; the three destination slabs are dense [rows, n] blocks `slab` floats apart)
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, long long n4, float* __restrict__ dst, long long slab,
                                                     int pattern, int rows, long long spitch) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n4 * rows; j += (long long)gridDim.x * blockDim.x) {
        const long long r = j / n4, i = j - r * n4;
        const float4 x = ld4(src + r * spitch + 4 * i);
        float* __restrict__ d = dst + r * (n4 * 4);
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); lo.x = x.x - hi.x;
        hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); lo.y = x.y - hi.y;
        hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); lo.z = x.z - hi.z;
        hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); lo.w = x.w - hi.w;
        st4(d + 4 * i, hi);
        st4(d + slab + 4 * i, pattern == 0 ? hi : lo);
        st4(d + 2 * slab + 4 * i, pattern == 0 ? lo : hi);
    }
}
static cudaError_t split3(const float* src, long long n, float* dst, long long slab, int pattern, cudaStream_t st, int rows = 1,
                          long long spitch = 0) {
    const long long n4 = n >> 2;
    long long blocks = (n4 * rows + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    split3_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, n4, dst, slab, pattern, rows, spitch);
    count_launch();
    return cudaGetLastError();
}
// MATGCN_EXACT_TC=0 keeps every contraction of the exact mode on the fp32 FFMA kernels (A/B comparisons, tests)
static bool exact_tc_enabled() {
    const char* e = getenv("MATGCN_EXACT_TC");
    return !(e && e[0] == '0');
}
constexpr int EXACT_TC_MAX_N = 2048;   // the split copies of the base matrices are 3 (K-1) N ldm floats: kept for N up to here
// One dense propagation of the exact mode as a 3-k-batch TF32 contraction on the tensor-core engine.  A3: the [hi, hi, lo] slabs of
// the base matrices (slab stride sa), B: the fp32 operand (nb contiguous floats) whose [hi, lo, hi] slabs go to B3.
// cudaErrorNotSupported: shape / alignment does not qualify - the caller runs the FFMA kernel.
template <bool A_KC, class Epi>
static cudaError_t prop_3xtf32(GemmP p, const Epi& epi, const float* A3, long long sa, float* B3, long long nb, cudaStream_t st) {
    if (!A3 || !B3 || (nb & 3) || !aligned16(p.B) || !aligned16(B3) || !aligned16(A3) || (p.lda & 3) || (p.ldb & 3) || p.KB != 1)
        return cudaErrorNotSupported;
    cudaError_t e = split3(p.B, nb, B3, nb, 1, st);
    if (e != cudaSuccess) return e;
    p.A = A3; p.sAk = sa; p.B = B3; p.sBk = nb; p.KB = 3;
    p.A16 = nullptr; p.B16 = nullptr; p.need16 = 0; p.npf = 0;
    e = launch_gemm_tc<128, A_KC, false, Epi>(p, epi, 1, st);
    if (e == cudaSuccess) g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return e;
}

// A time-batched split-K contraction of the exact mode (dM: both operands K-major, one dense [M, K] / [N, K] block per time step,
// KB = T k-batches, atomic epilogue) as 3xTF32: chunks of tc time steps, operand copies [term][tt][block] so that the k-batch
// index (term, tt) keeps ONE stride; the chunks accumulate through the atomic epilogue.  scratch: >= 6 * tc * max(nA, nB) floats.
template <class Epi>
static cudaError_t tb_3xtf32(const GemmP& p, const Epi& epi, float* scratch, long long scratch_floats, cudaStream_t st) {
    const long long nA = (long long)p.M * p.lda, nB = (long long)p.N * p.ldb;
    if (!scratch || p.lda != p.K || p.ldb != p.K || (nA & 3) || (nB & 3) || (p.sAk & 3) || (p.sBk & 3) || !aligned16(p.A) || !aligned16(p.B) ||
        !aligned16(scratch) || p.KB < 1)
        return cudaErrorNotSupported;
    long long tc = scratch_floats / (3 * (nA + nB));
    if (tc > p.KB) tc = p.KB;
    if (tc > 8) tc = 8;
    if (tc < 1) return cudaErrorNotSupported;
    float* A3 = scratch;
    float* B3 = scratch + 3 * tc * nA;
    for (int t0 = 0; t0 < p.KB; t0 += (int)tc) {
        const int c = (int)((p.KB - t0) < tc ? (p.KB - t0) : tc);
        cudaError_t e = split3(p.A + (long long)t0 * p.sAk, nA, A3, (long long)c * nA, 0, st, c, p.sAk);
        if (e != cudaSuccess) return e;
        e = split3(p.B + (long long)t0 * p.sBk, nB, B3, (long long)c * nB, 1, st, c, p.sBk);
        if (e != cudaSuccess) return e;
        GemmP q = p;
        q.A = A3; q.sAk = nA; q.B = B3; q.sBk = nB; q.KB = 3 * c; q.splits = 3 * c;
        q.A16 = nullptr; q.B16 = nullptr; q.need16 = 0; q.npf = 0;
        e = launch_gemm_tc<128, true, true, Epi>(q, epi, 1, st);
        if (e != cudaSuccess) return e;   // (NotSupported can only come from the first chunk: nothing has been accumulated yet)
        g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return cudaSuccess;
}

// bf16 [N*K, Cin, 3H] concatenation of the input rows of the gate and candidate weights: WX[nk, i, 0:2H] = Wg[nk, i, :],
// WX[nk, i, 2H:3H] = Wu[nk, i, :] (i < Cin) - lets the time-batched input gradient run as ONE contraction over 3H per support.
// (H % 4 == 0: four columns per thread, 16-byte loads and 8-byte stores)
__global__ void pack_wx16_kernel(const float* __restrict__ Wg, const float* __restrict__ Wu, int NK, int Cin, int I, int H,
                                 __nv_bfloat16* __restrict__ WX) {
    const int q = 3 * H / 4;   // column quads per row
    const long long total = (long long)NK * Cin * q;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % q) * 4;
        const long long ri = idx / q;
        const int i = (int)(ri % Cin);
        const long long nk = ri / Cin;
        const float4 v = c < 2 * H ? ld4(Wg + (nk * I + i) * 2 * H + c) : ld4(Wu + (nk * I + i) * H + c - 2 * H);
        st4_bf16(WX + ri * 3 * H + c, v);
    }
}
// rows of n floats: fp32 copy and bf16 twin in one pass (n % 4 == 0, 16-byte aligned rows)
__global__ void copy_twin_kernel(const float* __restrict__ src, long long spitch, float* __restrict__ dst, __nv_bfloat16* __restrict__ dst16,
                                 long long dpitch, long long n4, int rows) {
    const long long total = n4 * rows;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n4, c = (i - r * n4) * 4;
        const float4 v = ld4(src + r * spitch + c);
        st4(dst + r * dpitch + c, v);
        st4_bf16(dst16 + r * dpitch + c, v);
    }
}

// ------------------------------------------------------------------------------------------
// adaptive adjacency
// ------------------------------------------------------------------------------------------
extern "C" int matgcn_adaptive_adj_fwd(const float* L, const float* Rt, int N, int D, float* A, int ldm, void* stream) {
    REQUIRE(L && Rt && A, "null pointer");
    REQUIRE(N > 0 && D > 0 && ldm >= N, "bad dims");
    const size_t smem = (size_t)(N + D) * sizeof(float);
    REQUIRE(smem <= 200 * 1024, "N too large for the single-row softmax kernel");
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(adaptive_adj_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    adaptive_adj_fwd_kernel<<<N, 256, smem, st>>>(L, Rt, N, D, A, ldm);
    count_launch();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int matgcn_adaptive_adj_bwd(const float* L, const float* Rt, const float* A, const float* dA, int N, int D,
                                       int ldm, float* dL, float* dRt, float* scratch, void* stream) {
    REQUIRE(L && Rt && A && dA && dL && dRt && scratch, "null pointer");
    REQUIRE(N > 0 && D > 0 && ldm >= N, "bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    adaptive_adj_bwd_kernel<<<N, 256, (size_t)D * sizeof(float), st>>>(L, Rt, A, dA, N, D, ldm, scratch);
    count_launch();
    CK(cudaGetLastError());
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = 1;
    // dL = dpre [N,N] * Rt [N,D]
    p.A = scratch; p.lda = N; p.B = Rt; p.ldb = D; p.M = N; p.N = D; p.K = N;
    CK((launch_gemm<CfgSkinnyN, true, false>(p, epi_store(dL, 0, 0, D), 1, st)));
    // dRt = dpre^T * L
    p.A = scratch; p.lda = N; p.B = L; p.ldb = D;
    CK((launch_gemm<CfgSkinnyN, false, false>(p, epi_store(dRt, 0, 0, D), 1, st)));
    return 0;
}

// ------------------------------------------------------------------------------------------
// per-node weights
// ------------------------------------------------------------------------------------------
extern "C" int matgcn_nodeweights_fwd_ex(const float* E, const float* pool, const float* bias_pool, const float* c,
                                         int N, int D, int K, int I, int O, float* W, float* b, int flags, void* stream);
extern "C" int matgcn_nodeweights_fwd(const float* E, const float* pool, const float* bias_pool, const float* c,
                                      int N, int D, int K, int I, int O, float* W, float* b, void* stream) {
    return matgcn_nodeweights_fwd_ex(E, pool, bias_pool, c, N, D, K, I, O, W, b, MATGCN_FLAG_EXACT, stream);
}
// flags & MATGCN_FLAG_TF32: the big product (E x pool, a stream of N*K*I*O outputs) runs on the tensor-core engine
extern "C" int matgcn_nodeweights_fwd_ex(const float* E, const float* pool, const float* bias_pool, const float* c,
                                         int N, int D, int K, int I, int O, float* W, float* b, int flags, void* stream) {
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    REQUIRE(E && pool && bias_pool && c && W && b, "null pointer");
    REQUIRE(N > 0 && D > 0 && K > 0 && I > 0 && O > 0, "bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    const long long KIO = (long long)K * I * O;
    REQUIRE(KIO < 2147483647LL, "K*I*O overflows int");
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = 1;
    p.A = E; p.lda = D; p.B = pool; p.ldb = (int)KIO; p.M = N; p.N = (int)KIO; p.K = D;
    EpiStore e = epi_store(W, 0, 0, (int)KIO);
    e.scale = c; e.scale_div = I * O;
    {
        const cudaError_t le = (tc && D >= 8) ? launch_store_lean<ES_SCALE, 128, true, false, false>(p, e, 1, st) : cudaErrorNotSupported;
        if (le == cudaErrorNotSupported) CK((gemm_any<CfgBig, true, false>(tc, p, e, 1, st)));
        else CK(le);
    }
    p.B = bias_pool; p.ldb = O; p.N = O;
    CK((launch_gemm<CfgMid, true, false>(p, epi_store(b, 0, 0, O), 1, st)));
    return 0;
}

extern "C" int matgcn_nodeweights_bwd_ex(const float* E, const float* pool, const float* bias_pool, const float* c,
                                         const float* dW, const float* db, int N, int D, int K, int I, int O,
                                         float* dE, float* dpool, float* dbias_pool, float* dc, int flags, void* stream);
extern "C" int matgcn_nodeweights_bwd(const float* E, const float* pool, const float* bias_pool, const float* c,
                                      const float* dW, const float* db, int N, int D, int K, int I, int O,
                                      float* dE, float* dpool, float* dbias_pool, float* dc, void* stream) {
    return matgcn_nodeweights_bwd_ex(E, pool, bias_pool, c, dW, db, N, D, K, I, O, dE, dpool, dbias_pool, dc, MATGCN_FLAG_EXACT, stream);
}
extern "C" int matgcn_nodeweights_bwd_ex(const float* E, const float* pool, const float* bias_pool, const float* c,
                                         const float* dW, const float* db, int N, int D, int K, int I, int O,
                                         float* dE, float* dpool, float* dbias_pool, float* dc, int flags, void* stream) {
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    REQUIRE(E && pool && bias_pool && c && dW && db && dE && dpool && dbias_pool && dc, "null pointer");
    REQUIRE(N > 0 && D > 0 && K > 0 && I > 0 && O > 0, "bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    const long long KIO = (long long)K * I * O;
    REQUIRE(KIO < 2147483647LL, "K*I*O overflows int");
    const long long npool = (long long)D * KIO;
    CK(cudaMemsetAsync(dE, 0, sizeof(float) * (size_t)N * D, st));
    CK(cudaMemsetAsync(dc, 0, sizeof(float) * (size_t)K, st));
    // dpool is first used as scratch for c[k]*pool, consumed by the dE contraction below
    scale_groups_kernel<<<(unsigned)((npool + 255) / 256), 256, 0, st>>>(pool, c, npool, I * O, K, dpool);
    count_launch();
    CK(cudaGetLastError());
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1;
    // dE[n,d] = sum_col dW[n,col] * (c*pool)[d,col]   (+ db * bias_pool^T), split-K with atomics
    p.A = dW; p.lda = (int)KIO; p.B = dpool; p.ldb = (int)KIO; p.M = N; p.N = D; p.K = (int)KIO;
    p.splits = (int)((KIO + 2047) / 2048);
    if (p.splits > 128) p.splits = 128;
    EpiAtomic ea{dE, 0, 0, D};
    if (tc) {
        // both big contractions stream dW once each on the tensor-core engine (split-K for the N x D one)
        GemmP q = p;
        q.splits = (int)((KIO + 4095) / 4096);
        if (q.splits > 64) q.splits = 64;
        CK((gemm_any<CfgSkinnyN, true, true>(true, q, ea, 1, st)));
    } else {
        CK((launch_gemm<CfgSkinnyN, true, true>(p, ea, 1, st)));
    }
    p.A = db; p.lda = O; p.B = bias_pool; p.ldb = O; p.K = O; p.splits = 1;
    CK((launch_gemm<CfgSkinnyN, true, true>(p, ea, 1, st)));
    // G[d,col] = sum_n E[n,d] dW[n,col] ; dpool = c[k] * G
    p.A = E; p.lda = D; p.B = dW; p.ldb = (int)KIO; p.M = D; p.N = (int)KIO; p.K = N; p.splits = 1;
    EpiStore eg = epi_store(dpool, 0, 0, (int)KIO);
    eg.scale = c; eg.scale_div = I * O;
    {
        const cudaError_t le = (tc && N >= 8) ? launch_store_lean<ES_SCALE, 128, false, false, false>(p, eg, 1, st) : cudaErrorNotSupported;
        if (le == cudaErrorNotSupported) CK((gemm_any<CfgSkinnyM, false, false>(tc, p, eg, 1, st)));
        else CK(le);
    }
    // dc[k] = sum G * pool = sum dpool * pool / c[k]
    {
        dim3 grid(K, 64);
        view_weight_grad_kernel<<<grid, 256, 0, st>>>(dpool, pool, c, D, K, I * O, dc);
        count_launch();
        CK(cudaGetLastError());
    }
    // dbias_pool = E^T db
    p.A = E; p.lda = D; p.B = db; p.ldb = O; p.M = D; p.N = O; p.K = N;
    CK((launch_gemm<CfgSkinnyM, false, false>(p, epi_store(dbias_pool, 0, 0, O), 1, st)));
    return 0;
}

// Persistent recurrence kernels (rec_fwd.cuh: one cooperative launch per layer instead of four launches per time step; bf16 mode,
// rnn_units = 64).  MATGCN_REC=0 or matgcn_set_recurrent_kernel(0) selects one launch per phase (A/B comparisons, tests).
static int& rec_flag() {
    static int f = []() { const char* e = getenv("MATGCN_REC"); return (e && e[0] == '0') ? 0 : 1; }();
    return f;
}
extern "C" int matgcn_set_recurrent_kernel(int on) {
    const int prev = rec_flag();
    rec_flag() = on ? 1 : 0;
    return prev;
}
extern "C" int matgcn_rec_timing(int on) {
    rec_timing_enable(on != 0);
    return 0;
}
extern "C" int matgcn_rec_timing_read(double* fwd_ms, int* fwd_launches, double* bwd_ms, int* bwd_launches) {
    REQUIRE(fwd_ms && fwd_launches && bwd_ms && bwd_launches, "null pointer");
    rec_timing_read(fwd_ms, fwd_launches, bwd_ms, bwd_launches);
    return 0;
}
// Fused tail of the forward step (candidate + residual cell + mix in one launch); MATGCN_FUSED_TAIL=0 or
// matgcn_set_fused_tail(0) selects the three separate contractions (A/B comparisons, tests).
static int& fused_tail_flag() {
    static int f = []() { const char* e = getenv("MATGCN_FUSED_TAIL"); return (e && e[0] == '0') ? 0 : 1; }();
    return f;
}
static bool fused_tail_enabled() { return fused_tail_flag() != 0; }
extern "C" int matgcn_set_fused_tail(int on) {
    const int prev = fused_tail_flag();
    fused_tail_flag() = on ? 1 : 0;
    return prev;
}
// One contraction of a recurrence step on the one-launch-per-phase path (the modes and shapes the persistent kernels of rec.cu
// do not cover).
#define STEP_GEMM(SLOT, Cfg, AKC, BKC, P, EPI, Z)                            \
    do {                                                                     \
        CK((gemm_any<Cfg, AKC, BKC>(tc, P, EPI, Z, st)));                    \
        TR();                                                                \
    } while (0)
// Support propagation of one step as GemmP: dst[1..K) = M * src   (slot stride U, cols = B*C)
static GemmP prop_params(const float* M, int ldm, int N, int Kp, const float* slot0, int cols) {
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = 1;
    p.A = M; p.lda = ldm; p.M = Kp * N; p.K = N;
    p.B = slot0; p.ldb = cols; p.N = cols;
    return p;
}

// ------------------------------------------------------------------------------------------
// encoder layer: workspace layout
// ------------------------------------------------------------------------------------------
struct LayerWs {
    size_t PX, GX, RX, PH, PZ, Z, R, HC, H1, Z2, R2, HC2, ZH2, RGH, RUH, R3X, RB3, BX3, MPH, M16, PH16, PZ16, PX16, WG16, WU16, WX16, M3, X3, total;
    size_t U, UX;  // floats of one [N,B,H] / [N,B,Cin] block
};
static size_t align64(size_t v) { return (v + 63) / 64 * 64; }
static LayerWs layer_ws(int T, int N, int B, int Cin, int H, int K) {
    LayerWs w;
    w.U = (size_t)N * B * H;
    w.UX = (size_t)N * B * Cin;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o = align64(o + n); return r; };
    w.PX = take((size_t)T * K * w.UX);
    w.GX = take((size_t)T * 3 * w.U);
    w.RX = take((size_t)T * 3 * w.U);
    w.PH = take(((size_t)T * K + K) * w.U);   // (+ slots [T, 1..K): written by a chained next layer in the fp32-slot modes)
    w.PZ = take((size_t)T * K * w.U);
    w.Z = take((size_t)T * w.U);
    w.R = take((size_t)T * w.U);
    w.HC = take((size_t)T * w.U);
    w.H1 = take((size_t)T * w.U);
    w.Z2 = take((size_t)T * w.U);
    w.R2 = take((size_t)T * w.U);
    w.HC2 = take((size_t)T * w.U);
    w.ZH2 = take((size_t)T * w.U);
    w.RGH = take((size_t)2 * H * H);  // Rgw[:, Cin:] and Ruw[:, Cin:] repacked densely (16-byte aligned rows for TMA)
    w.RUH = take((size_t)H * H);
    w.R3X = take((size_t)3 * H * Cin);  // [Rgw[:, 0:Cin]; Ruw[:, 0:Cin]] packed [3H, Cin]: the residual cell's input side as ONE operand
    w.RB3 = take((size_t)3 * H);        // [Rgb; Rub]
    w.BX3 = take((size_t)N * 3 * H);    // [bg | bu] per node
    w.MPH = take(256);  // grid-barrier counter of the persistent kernel
    // bf16 twins (sizes in floats = elements / 2): base matrices, PH / PZ / PX (same layouts as the fp32 arrays: slot 0 =
    // the state itself, slots 1.. = its propagated copies) and the per-node weights
    w.M16 = take(((size_t)(K - 1) * N * 8 * ((N + 7) / 8 + 1)) / 2 + 64);
    w.PH16 = take((((size_t)T * K + K) * w.U) / 2 + 64);   // (+ slots [T, 1..K): written by a chained next layer, see encoder_layer_fwd_impl)
    w.PZ16 = take(((size_t)T * K * w.U) / 2 + 64);
    w.PX16 = take(((size_t)T * K * w.UX) / 2 + 64);
    w.WG16 = take(((size_t)N * K * (Cin + H) * 2 * H) / 2 + 64);
    w.WU16 = take(((size_t)N * K * (Cin + H) * H) / 2 + 64);
    w.WX16 = take(((size_t)N * K * Cin * 3 * H) / 2 + 64);  // bf16 [N, K, Cin, 3H]: gate | candidate input-row weights (forward GX, backward DPX)
    // exact mode, dense propagations as 3xTF32 (prop_3xtf32): [hi, hi, lo] slabs of the base matrices (kept for the backward) and the
    // [hi, lo, hi] slabs of one step's state
    const bool x3 = N <= EXACT_TC_MAX_N;
    w.M3 = take(x3 ? (size_t)3 * (K - 1) * N * 8 * ((N + 7) / 8 + 1) : 0);
    w.X3 = take(x3 ? (size_t)3 * w.U : 0);
    w.total = o;
    return w;
}
struct LayerBws {
    size_t DPX, DPT, DPHA, DPZA, DH1, DHD, DHC, DRES, MPH, DPT16, DPX16, DG16, DPT3, TB3, TB3_floats, total;
};
static LayerBws layer_bws(int T, int N, int B, int Cin, int H, int K, int n_adp) {
    LayerBws w;
    const size_t U = (size_t)N * B * H, UX = (size_t)N * B * Cin;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o = align64(o + n); return r; };
    w.DPX = take((size_t)T * K * UX);
    w.DPT = take((size_t)K * U);
    w.DPHA = take((size_t)T * (n_adp > 0 ? n_adp : 0) * U);
    w.DPZA = take((size_t)T * (n_adp > 0 ? n_adp : 0) * U);
    w.DH1 = take(U);
    w.DHD = take(U);
    w.DHC = take(U);
    w.DRES = take(U);
    w.MPH = take(256);
    w.DPT16 = take(((size_t)K * U) / 2 + 64);
    w.DPX16 = take(((size_t)T * K * UX) / 2 + 64);
    w.DG16 = take(((size_t)T * 3 * U) / 2 + 64);  // bf16 twin of the pre-activation gradients DG [T, N*B, 3H]
    w.DPT3 = take(N <= EXACT_TC_MAX_N ? (size_t)3 * (K - 1) * U : 0);   // exact mode: [hi, lo, hi] slabs of DPT[1..K) (prop_3xtf32)
    // exact mode: operand copies of the dM contractions (tb_3xtf32): up to eight time steps per chunk, at most 256 MB
    w.TB3_floats = 0;
    if (N <= EXACT_TC_MAX_N && n_adp > 0) {
        const size_t um = U > UX ? U : UX;
        size_t tcn = ((size_t)64 << 20) / (6 * um);
        if (tcn > 8) tcn = 8;
        if (tcn > (size_t)T) tcn = (size_t)T;
        w.TB3_floats = tcn * 6 * um;
    }
    w.TB3 = take(w.TB3_floats);
    w.total = o;
    return w;
}

extern "C" size_t matgcn_encoder_layer_fwd_ws_bytes(int T, int N, int B, int Cin, int H, int K) {
    return layer_ws(T, N, B, Cin, H, K).total * sizeof(float);
}
extern "C" size_t matgcn_encoder_layer_bwd_ws_bytes(int T, int N, int B, int Cin, int H, int K, int n_adp) {
    return layer_bws(T, N, B, Cin, H, K, n_adp).total * sizeof(float);
}
extern "C" size_t matgcn_encoder_layer_y_offset(int T, int N, int B, int Cin, int H, int K) {
    const LayerWs w = layer_ws(T, N, B, Cin, H, K);
    return w.PH + (size_t)K * w.U;  // PH[t+1, 0]
}
extern "C" size_t matgcn_encoder_layer_y_tstride(int T, int N, int B, int Cin, int H, int K) {
    (void)T; (void)Cin;
    return (size_t)K * N * B * H;
}
extern "C" size_t matgcn_encoder_layer_slot_offset(const char* name, int T, int N, int B, int Cin, int H, int K) {
    const LayerWs w = layer_ws(T, N, B, Cin, H, K);
    struct { const char* n; size_t v; } tab[] = {{"PX", w.PX}, {"GX", w.GX}, {"RX", w.RX}, {"PH", w.PH}, {"PZ", w.PZ},
                                                 {"Z", w.Z}, {"R", w.R}, {"HC", w.HC}, {"H1", w.H1}, {"Z2", w.Z2},
                                                 {"R2", w.R2}, {"HC2", w.HC2}, {"ZH2", w.ZH2}, {"PH16", w.PH16}};
    for (auto& e : tab)
        if (strcmp(e.n, name) == 0) return e.v;
    return (size_t)-1;
}

static int check_layer_dims(int T, int N, int B, int Cin, int H, int K, int ldm) {
    if (T <= 0 || N <= 0 || B <= 0 || Cin <= 0 || H <= 0 || K < 2) return fail("encoder_layer", "bad dims (need K >= 2)");
    if (ldm < N) return fail("encoder_layer", "ldm < N");
    if ((long long)N * B * 3 * H >= 2147483647LL) return fail("encoder_layer", "N*B*3H overflows int (shard the batch)");
    if ((long long)(K - 1) * N * (long long)ldm >= 2147483647LL) return fail("encoder_layer", "(K-1)*N*ldm overflows int");
    return 0;
}

// Support propagation for a block of Z time steps: dst[z][1..K) = M * src[z][0]   (slot stride U)
static cudaError_t propagate(bool tc, const float* M, int ldm, int N, int Kp, const float* slot0, long long zstride, int cols,
                             float* slot1, int Z, cudaStream_t st) {
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = 1;
    p.A = M; p.lda = ldm; p.M = Kp * N; p.K = N;
    p.B = slot0; p.ldb = cols; p.N = cols; p.sB1 = zstride;
    return gemm_any<CfgBig, true, false>(tc, p, epi_plain(slot1, zstride, 0, cols), Z, st);  // the epilogue the layer entry points use
}

extern "C" int matgcn_propagate_fwd(const float* M, int Kp, int N, int ldm, const float* X, int cols, float* P,
                                    int flags, void* stream) {
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    REQUIRE(M && X && P, "null pointer");
    REQUIRE(Kp > 0 && N > 0 && cols > 0 && ldm >= N, "bad dims");
    CK(propagate(tc, M, ldm, N, Kp, X, 0, cols, P, 1, (cudaStream_t)stream));
    return 0;
}

// Diagnostics: device buffer (>= 8 * tiles-per-CTA int64) that CTA 0 of later tensor-core launches fills with
// clock64() stamps per tile: [producer start, mma wait-begin, mma start, mma committed, epi wait-begin, epi start, epi end].
extern "C" int matgcn_debug_set_mode(int mode) {
    tc_debug_mode() = mode;
    return 0;
}
extern "C" int matgcn_debug_set_timeline_skip(int launches) {
    tc_debug_countdown() = launches;
    return 0;
}
extern "C" int matgcn_debug_set_timeline(long long* buf) {
    tc_debug_buffer() = buf;
    tc_debug_countdown() = 0;
    return 0;
}

// bf16 form of matgcn_propagate_fwd: M16 [Kp, N, ldm] and X16 [N, cols] are bf16 twins, P is float32.
extern "C" int matgcn_propagate_fwd_bf16(const void* M16, int Kp, int N, int ldm, const void* X16, int cols, float* P, void* stream) {
    REQUIRE(M16 && X16 && P, "null pointer");
    REQUIRE(Kp > 0 && N > 0 && cols > 0 && ldm >= N, "bad dims");
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = 1;
    p.A16 = (const __nv_bfloat16*)M16; p.lda = ldm; p.M = Kp * N; p.K = N;
    p.B16 = (const __nv_bfloat16*)X16; p.ldb = cols; p.N = cols;
    cudaError_t e = launch_gemm_tc<128, true, false, EpiPlain, true>(p, epi_plain(P, 0, 0, cols), 1, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(__func__, "operands do not meet the TMA alignment rules");
    CK(e);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// The propagation launch exactly as a bf16-mode step issues it: bf16 operands, and only the bf16 twin of the result is stored
// (P16 [Kp, N, cols]; the fp32 copy of the propagated slots is skipped - EpiPlain::c_z2_hi = 0).
extern "C" int matgcn_propagate_fwd_bf16_twin(const void* M16, int Kp, int N, int ldm, const void* X16, int cols, void* P16, void* stream) {
    REQUIRE(M16 && X16 && P16, "null pointer");
    REQUIRE(Kp > 0 && N > 0 && cols > 0 && ldm >= N, "bad dims");
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = 1;
    p.A16 = (const __nv_bfloat16*)M16; p.lda = ldm; p.M = Kp * N; p.K = N;
    p.B16 = (const __nv_bfloat16*)X16; p.ldb = cols; p.N = cols;
    EpiPlain e = epi_plain(reinterpret_cast<float*>(P16), 0, 0, cols);   // C is never dereferenced
    e.C16 = (__nv_bfloat16*)P16;
    e.c_z2_hi = 0;
    cudaError_t err = launch_gemm_tc<128, true, false, EpiPlain, true>(p, e, 1, (cudaStream_t)stream);
    if (err == cudaErrorNotSupported) return fail(__func__, "operands do not meet the TMA alignment rules");
    CK(err);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// Plain C = A*B through either engine, for unit tests of the GEMM kernels (all operand layouts).
extern "C" int matgcn_gemm_debug(int a_kc, int b_kc, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                                 float* C, int ldc, int splits, int flags, void* stream) {
    REQUIRE(A && B && C, "null pointer");
    REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, "bad dims");
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    cudaStream_t st = (cudaStream_t)stream;
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = splits;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.M = M; p.N = N; p.K = K;
    if (splits > 1) {
        CK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
        EpiAtomic e{C, 0, 0, ldc};
        if (a_kc && !b_kc) CK((gemm_any<CfgBig, true, false>(tc, p, e, 1, st)));
        else if (!a_kc && !b_kc) CK((gemm_any<CfgBig, false, false>(tc, p, e, 1, st)));
        else if (a_kc && b_kc) CK((gemm_any<CfgBig, true, true>(tc, p, e, 1, st)));
        else return fail(__func__, "layout combination (A M-contiguous, B K-contiguous) is not used on this path");
    } else {
        EpiStore e = epi_store(C, 0, 0, ldc);
        if (a_kc && !b_kc) CK((gemm_any<CfgBig, true, false>(tc, p, e, 1, st)));
        else if (!a_kc && !b_kc) CK((gemm_any<CfgBig, false, false>(tc, p, e, 1, st)));
        else if (a_kc && b_kc) CK((gemm_any<CfgBig, true, true>(tc, p, e, 1, st)));
        else return fail(__func__, "layout combination (A M-contiguous, B K-contiguous) is not used on this path");
    }
    return 0;
}

// Same with bf16 operands (A16/B16 are bf16 device arrays in the layouts above; C is float32).
extern "C" int matgcn_gemm_debug_bf16(int a_kc, int b_kc, int M, int N, int K, const void* A16, int lda, const void* B16,
                                      int ldb, float* C, int ldc, int splits, void* stream) {
    REQUIRE(A16 && B16 && C, "null pointer");
    REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, "bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    GemmP p;
    memset(&p, 0, sizeof(p));
    p.KB = 1; p.Z2 = 1; p.splits = splits;
    p.A16 = (const __nv_bfloat16*)A16; p.B16 = (const __nv_bfloat16*)B16;
    p.lda = lda; p.ldb = ldb; p.M = M; p.N = N; p.K = K;
    cudaError_t e = cudaErrorNotSupported;
    if (splits > 1) {
        CK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
        EpiAtomic ep{C, 0, 0, ldc};
        if (a_kc && !b_kc) e = launch_gemm_tc<128, true, false, EpiAtomic, true>(p, ep, 1, st);
        else if (!a_kc && !b_kc) e = launch_gemm_tc<128, false, false, EpiAtomic, true>(p, ep, 1, st);
        else if (a_kc && b_kc) e = launch_gemm_tc<128, true, true, EpiAtomic, true>(p, ep, 1, st);
    } else {
        EpiStore ep = epi_store(C, 0, 0, ldc);
        if (a_kc && !b_kc) e = launch_gemm_tc<128, true, false, EpiStore, true>(p, ep, 1, st);
        else if (!a_kc && !b_kc) e = launch_gemm_tc<128, false, false, EpiStore, true>(p, ep, 1, st);
        else if (a_kc && b_kc) e = launch_gemm_tc<128, true, true, EpiStore, true>(p, ep, 1, st);
    }
    if (e == cudaErrorNotSupported) return fail(__func__, "operands do not meet the TMA alignment rules (or unused layout)");
    CK(e);
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// ------------------------------------------------------------------------------------------
// encoder layer forward
// ------------------------------------------------------------------------------------------
// Layer chaining (Cin == H): the input of an inner layer is the previous layer's output, and its propagated copies
// M x_t are what the previous layer's recurrence already computed for its OWN next step (PH16[t+1, 1..K) = M h_t, same bf16
// operands).  With x16_chain = &PH16_prev[1, 0] the layer reads x (fp32: slot 0 of the previous PH, same time stride) and its
// bf16 slots where they sit: no copy, and only the last time step is propagated (into PH16_prev[T, 1..K), which exists for this).
// The exact and tf32 modes keep the propagated slots in fp32 (PH_prev[t+1, 1..K), same layout as PX) and chain onto those.
static bool chain_ok(int T, int N, int B, int Cin, int H, int K, int ldm, int flags) {
    const char* e = getenv("MATGCN_CHAIN");
    if (e && e[0] == '0') return false;
    if (T < 1 || Cin != H || xside_small_ok(Cin, H, K)) return false;
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    const size_t m16_cap = ((size_t)(K - 1) * N * 8 * ((N + 7) / 8 + 1));
    const bool bf = tc && (flags & MATGCN_FLAG_BF16) != 0 && (size_t)(K - 1) * N * ldm <= m16_cap;
    if (!bf) return true;   // fp32 slots (exact and tf32 modes): PH_prev[t+1, 1..K) = M h_t is this layer's PX[t, 1..K) as it stands
    const bool skip32 = !(H & 7) && !(ldm & 7) && B >= 8;
    return skip32 && !(Cin & 7);
}
extern "C" int matgcn_encoder_layer_chain_ok(int T, int N, int B, int Cin, int H, int K, int ldm, int flags) {
    return chain_ok(T, N, B, Cin, H, K, ldm, flags) ? 1 : 0;
}

static int encoder_layer_fwd_impl(int T, int N, int B, int Cin, int H, int K, int ldm,
                                        const float* x, long long x_tstride, const float* h0, const float* M,
                                        const float* Wg, const float* bg, const float* Wu, const float* bu,
                                        const float* Rgw, const float* Rgb, const float* Ruw, const float* Rub,
                                        const float* mix, float* ws, int flags, void* x16_chain, void* stream) {
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    REQUIRE(x && M && Wg && bg && Wu && bu && Rgw && Rgb && Ruw && Rub && mix && ws, "null pointer");
    if (check_layer_dims(T, N, B, Cin, H, K, ldm)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    Tracer tr(st);
    const LayerWs w = layer_ws(T, N, B, Cin, H, K);
    const int Kp = K - 1, I = Cin + H;
    const long long U = (long long)w.U, UX = (long long)w.UX;
    float* PX = ws + w.PX; float* GX = ws + w.GX; float* RX = ws + w.RX; float* PH = ws + w.PH; float* PZ = ws + w.PZ;

    // bf16 propagation (flags & MATGCN_FLAG_BF16): bf16 twins of the base matrices and of every propagated tensor
    const size_t m16_cap = ((size_t)Kp * N * 8 * ((N + 7) / 8 + 1));
    const bool bf = tc && (flags & MATGCN_FLAG_BF16) != 0 && (size_t)Kp * N * ldm <= m16_cap;
    __nv_bfloat16* M16 = reinterpret_cast<__nv_bfloat16*>(ws + w.M16);
    __nv_bfloat16* PH16 = reinterpret_cast<__nv_bfloat16*>(ws + w.PH16);
    __nv_bfloat16* PZ16 = reinterpret_cast<__nv_bfloat16*>(ws + w.PZ16);
    __nv_bfloat16* PX16 = reinterpret_cast<__nv_bfloat16*>(ws + w.PX16);
    __nv_bfloat16* WG16 = reinterpret_cast<__nv_bfloat16*>(ws + w.WG16);
    __nv_bfloat16* WU16 = reinterpret_cast<__nv_bfloat16*>(ws + w.WU16);
    const bool chained = x16_chain != nullptr;
    if (chained) {
        REQUIRE(chain_ok(T, N, B, Cin, H, K, ldm, flags) && x_tstride == K * UX && aligned16(x) && aligned16(x16_chain),
                "chained input does not qualify (matgcn_encoder_layer_chain_ok)");
        PX = const_cast<float*>(x);   // read only from here on: slot 0 of the previous layer's PH, time stride K*U
        PX16 = reinterpret_cast<__nv_bfloat16*>(x16_chain);
    }
    // x -> slot 0 of PX[t] (and of its bf16 twin: one pass over x when the rows allow 16-byte accesses)
    const bool twin_copy = bf && !(UX & 3) && !(x_tstride & 3) && aligned16(x) && aligned16(PX) && aligned16(PX16);
    if (chained) {
        // nothing to copy
    } else if (twin_copy) {
        copy_twin_kernel<<<148 * 16, 256, 0, st>>>(x, x_tstride, PX, PX16, K * UX, UX >> 2, T);
        count_launch();
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpy2DAsync(PX, sizeof(float) * K * UX, x, sizeof(float) * x_tstride, sizeof(float) * UX, T,
                             cudaMemcpyDeviceToDevice, st));
    }
    if (bf) {
        CK(to_bf16(M, 0, M16, 0, (long long)Kp * N * ldm, 1, st));
        if (!twin_copy && !chained) CK(to_bf16(x, x_tstride, PX16, K * UX, UX, T, st));
        // per-node weights: 2-byte twins are what the step contractions stream (and, marked evict-last, what stays in L2
        // across the 24 steps: 49 MB per layer at the Baltimore size instead of 99 MB of fp32 from HBM every step)
        CK(to_bf16(Wg, 0, WG16, 0, (long long)N * K * I * 2 * H, 1, st));
        CK(to_bf16(Wu, 0, WU16, 0, (long long)N * K * I * H, 1, st));
    }
    // bf16 mode: the propagated slots (k >= 1) of PX / PH / PZ exist only as bf16 twins - their fp32 stores are skipped, and the
    // contractions that consume them are marked need16 (no fallback engine may touch the unwritten fp32 slots).
    // Needs shapes for which the bf16 tensor-core launches are always eligible (16-byte pitches) and the per-phase launch path.
    const bool skip32 = bf && !(H & 7) && !(ldm & 7) && B >= 8;  // (B = reduction length of the weight gradients)
    const bool skip32x = skip32 && !xside_small_ok(Cin, H, K) && !(Cin & 7);
    const bool warm = l2_warm_layer(N, K, Cin, H, B);
    // exact mode: the dense propagations of the recurrence run as 3xTF32 on the tensor-core engine (prop_3xtf32)
    const long long m3_slab = (long long)Kp * N * ldm;
    const bool x3 = !tc && exact_tc_enabled() && N <= EXACT_TC_MAX_N && !(ldm & 3) && !((B * H) & 3) && aligned16(M) &&
                    (size_t)m3_slab <= (size_t)Kp * N * 8 * ((N + 7) / 8 + 1);
    float* M3 = x3 ? ws + w.M3 : nullptr;
    float* X3 = x3 ? ws + w.X3 : nullptr;
    if (x3) CK(split3(M, m3_slab, M3, m3_slab, 0, st));
    // PX[t, 1..K) = M * x_t  (all t at once; a chained layer finds t < T-1 in place and propagates the last step only)
    {
        const long long t0 = chained ? (long long)(T - 1) * K * UX : 0;
        GemmP pp = prop_params(M, ldm, N, Kp, PX + t0, B * Cin);
        pp.sB1 = K * UX;
        EpiPlain e = epi_plain(PX + t0 + UX, K * UX, 0, B * Cin);
        if (bf) { pp.A16 = M16; pp.B16 = PX16 + t0; e.C16 = PX16 + t0 + UX; }
        if (skip32x) e.c_z2_hi = 0;   // every consumer of PX[t, k >= 1] reads the bf16 twin
        if (chained && skip32x) pp.need16 = 1;   // (bf16 mode: the fp32 slots k >= 1 do not exist behind a chained input)
        CK((gemm_any<CfgBig, true, false>(tc, pp, e, chained ? 1 : T, st)));
    }
    TR();

    GemmP p;
    if (xside_small_ok(Cin, H, K)) {
        // tiny channel count (layer 0): one streaming kernel writes GX and RX
        dim3 grid(N, T);
        const int threads = (3 * H + 31) / 32 * 32;
        if (!(H & 3) && 3 * H / 4 * XS_RG <= 256 && aligned16(Wg) && aligned16(Wu) && aligned16(bg) && aligned16(bu) && aligned16(Rgb) &&
            aligned16(Rub) && aligned16(GX) && aligned16(RX))
            xside_fwd_small4_kernel<<<grid, 3 * H / 4 * XS_RG, sizeof(float) * (size_t)K * B * Cin, st>>>(PX, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub,
                                                                                                  T, N, B, Cin, H, K, GX, RX);
        else
            xside_fwd_small_kernel<<<grid, threads, sizeof(float) * (size_t)K * B * Cin, st>>>(PX, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, T, N,
                                                                                          B, Cin, H, K, GX, RX);
        count_launch();
        TR();
        CK(cudaGetLastError());
    } else {
    // the residual cell's input-side weights and biases packed as ONE [3H, Cin] operand (also read by the layer backward)
    float* R3X = ws + w.R3X; float* RB3 = ws + w.RB3;
    CK(cudaMemcpy2DAsync(R3X, sizeof(float) * Cin, Rgw, sizeof(float) * I, sizeof(float) * Cin, 2 * H, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpy2DAsync(R3X + (size_t)2 * H * Cin, sizeof(float) * Cin, Ruw, sizeof(float) * I, sizeof(float) * Cin, H,
                         cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(RB3, Rgb, sizeof(float) * 2 * H, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(RB3 + 2 * H, Rub, sizeof(float) * H, cudaMemcpyDeviceToDevice, st));
    // GX[t, n, :, 0:2H] = bg[n] + sum_k PX[t,k,n] * Wg[n,k,0:Cin,:]   z = (n, t), k-batches = k
    memset(&p, 0, sizeof(p));
    p.splits = 1; p.Z2 = T; p.KB = K;
    p.A = PX; p.lda = Cin; p.sA1 = (long long)B * Cin; p.sA2 = K * UX; p.sAk = UX;
    p.M = B; p.K = Cin;
    bool gx_done = false;
    if (bf && !(Cin & 7) && !(H & 3)) {
        // bf16 mode: gate and candidate columns in ONE contraction against the packed input-row weights WX16 [N, K, Cin, 3H]
        // (PX16 is streamed once; the layer backward reuses the pack for the input gradients)
        __nv_bfloat16* WX16 = reinterpret_cast<__nv_bfloat16*>(ws + w.WX16);
        float* BX3 = ws + w.BX3;
        pack_wx16_kernel<<<148 * 8, 256, 0, st>>>(Wg, Wu, N * K, Cin, I, H, WX16);
        count_launch();
        CK(cudaGetLastError());
        CK(cudaMemcpy2DAsync(BX3, sizeof(float) * 3 * H, bg, sizeof(float) * 2 * H, sizeof(float) * 2 * H, N, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpy2DAsync(BX3 + 2 * H, sizeof(float) * 3 * H, bu, sizeof(float) * H, sizeof(float) * H, N, cudaMemcpyDeviceToDevice, st));
        GemmP q = p;
        q.A16 = PX16; q.B16 = WX16; q.ldb = 3 * H; q.N = 3 * H; q.sB1 = (long long)K * Cin * 3 * H; q.sB2 = 0; q.sBk = (long long)Cin * 3 * H;
        EpiStore e = epi_store(GX, (long long)B * 3 * H, 3 * U, 3 * H);
        e.bias = BX3; e.bias_s1 = 3 * H;
        cudaError_t ge = launch_store_lean<ES_BIAS, 128, true, false, true>(q, e, N * T, st);
        if (ge == cudaErrorNotSupported) {
            ge = launch_gemm_tc<128, true, false, EpiStore, true>(q, e, N * T, st);
            if (ge == cudaSuccess) g_tc_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (ge == cudaSuccess) {
            TR();
            gx_done = true;
        } else if (ge != cudaErrorNotSupported) {
            CK(ge);
        }
    }
    if (!gx_done) {
        p.B = Wg; p.ldb = 2 * H; p.N = 2 * H; p.sB1 = (long long)K * I * 2 * H; p.sB2 = 0; p.sBk = (long long)I * 2 * H;
        if (bf) { p.A16 = PX16; p.B16 = WG16; }
        p.need16 = skip32x ? 1 : 0;
        EpiStore e = epi_store(GX, (long long)B * 3 * H, 3 * U, 3 * H);
        e.bias = bg; e.bias_s1 = 2 * H;
        CK((gemm_any<CfgMid, true, false>(tc, p, e, N * T, st)));
        TR();
        p.B = Wu; p.ldb = H; p.N = H; p.sB1 = (long long)K * I * H; p.sBk = (long long)I * H;
        if (bf) p.B16 = WU16;
        e = epi_store(GX + 2 * H, (long long)B * 3 * H, 3 * U, 3 * H);
        e.bias = bu; e.bias_s1 = H;
        CK((gemm_any<CfgMid, true, false>(tc, p, e, N * T, st)));
        TR();
    }
    // RX[t] = x_t * [Rgw[:, 0:Cin]; Ruw[:, 0:Cin]]^T + [Rgb; Rub]          flat rows (n,b), z = t
    memset(&p, 0, sizeof(p));
    p.splits = 1; p.Z2 = 1; p.KB = 1;
    p.A = PX; p.lda = Cin; p.sA1 = K * UX; p.M = N * B; p.K = Cin;
    bool rx_done = false;
    if (tc && Cin >= 8) {
        // fast modes: all 3H columns in one contraction against the packed operand (x is streamed once)
        GemmP q = p;
        q.B = R3X; q.ldb = Cin; q.N = 3 * H;
        EpiStore e = epi_store(RX, 3 * U, 0, 3 * H);
        e.bias = RB3; e.bias_s1 = 0;
        cudaError_t re = launch_store_lean<ES_BIAS, 128, true, true, false>(q, e, T, st);
        if (re == cudaErrorNotSupported) {
            re = launch_gemm_tc<128, true, true, EpiStore>(q, e, T, st);
            if (re == cudaSuccess) g_tc_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (re == cudaSuccess) {
            TR();
            rx_done = true;
        } else if (re != cudaErrorNotSupported) {
            CK(re);
        }
    }
    if (!rx_done) {
        p.B = Rgw; p.ldb = I; p.N = 2 * H;
        EpiStore e = epi_store(RX, 3 * U, 0, 3 * H);
        e.bias = Rgb; e.bias_s1 = 0;
        CK((gemm_any<CfgMid, true, true>(tc, p, e, T, st)));
        TR();
        p.B = Ruw; p.ldb = I; p.N = H;
        e = epi_store(RX + 2 * H, 3 * U, 0, 3 * H);
        e.bias = Rub; e.bias_s1 = 0;
        CK((gemm_any<CfgMid, true, true>(tc, p, e, T, st)));
        TR();
    }
    }
    // dense, aligned copies of the hidden-state columns of the residual GRU weights
    float* RgH = ws + w.RGH; float* RuH = ws + w.RUH;
    CK(cudaMemcpy2DAsync(RgH, sizeof(float) * H, Rgw + Cin, sizeof(float) * I, sizeof(float) * H, 2 * H, cudaMemcpyDeviceToDevice, st));
    TR();
    CK(cudaMemcpy2DAsync(RuH, sizeof(float) * H, Ruw + Cin, sizeof(float) * I, sizeof(float) * H, H, cudaMemcpyDeviceToDevice, st));
    TR();
    // initial state
    if (h0) CK(cudaMemcpyAsync(PH, h0, sizeof(float) * U, cudaMemcpyDeviceToDevice, st));
    else CK(cudaMemsetAsync(PH, 0, sizeof(float) * U, st));
    if (bf) {
        if (h0) CK(to_bf16(h0, 0, PH16, 0, U, 1, st));
        else CK(cudaMemsetAsync(PH16, 0, sizeof(__nv_bfloat16) * U, st));
    }

    constexpr bool use_multi = false;   // (the cooperative multi-phase kernel of round 1 is gone: rec_fwd.cuh replaces it)
    if (skip32 && H == 64 && rec_flag() && fused_tail_enabled()) {
        // the whole recurrence as one persistent cooperative launch (rec_fwd.cuh)
        RecFwdArgs ra{T, N, B, Cin, K, ldm, M16, PH16, PZ16, WG16, WU16, GX, RX, PH, PZ,
                      ws + w.Z, ws + w.R, ws + w.HC, ws + w.H1, ws + w.Z2, ws + w.R2, ws + w.HC2, ws + w.ZH2,
                      RgH, RuH, mix, reinterpret_cast<unsigned int*>(ws + w.MPH)};
        const cudaError_t re = launch_rec_fwd(ra, st);
        if (re == cudaSuccess) {
            g_tc_launches.fetch_add(1, std::memory_order_relaxed);
            TR();
            tr.report("encoder_layer_fwd");
            return 0;
        }
        if (re != cudaErrorNotSupported) return fail(__func__, cudaGetErrorString(re));
    }
    {
        for (int t = 0; t < T; ++t) {
            float* PHt = PH + (long long)t * K * U;
            float* PZt = PZ + (long long)t * K * U;
            float* Zt = ws + w.Z + t * U; float* Rt_ = ws + w.R + t * U; float* HCt = ws + w.HC + t * U;
            float* H1t = ws + w.H1 + t * U; float* Z2t = ws + w.Z2 + t * U; float* R2t = ws + w.R2 + t * U;
            float* HC2t = ws + w.HC2 + t * U; float* ZH2t = ws + w.ZH2 + t * U;
            const float* GXt = GX + (long long)t * 3 * U; const float* RXt = RX + (long long)t * 3 * U;
            // (a) PH[t,1..] = M * h
            p = prop_params(M, ldm, N, Kp, PHt, B * H);
            __nv_bfloat16* PH16t = PH16 + (long long)t * K * U;
            __nv_bfloat16* PZ16t = PZ16 + (long long)t * K * U;
            {
                EpiPlain e = epi_plain(PHt + U, 0, 0, B * H);
                if (bf) { p.A16 = M16; p.B16 = PH16t; e.C16 = PH16t + U; }
                if (skip32) e.c_z2_hi = 0;
                if (bf && warm) {   // the gate contraction that follows streams Wg16[n, k, Cin:, :] from HBM
                    add_pf(p, WG16 + (long long)Cin * 2 * H, (long long)I * 2 * H * 2, (long long)H * 2 * H * 2, N * K);
                    if (l2_warm_level() >= 2) add_pf(p, GXt, 0, (long long)3 * U * 4, 1);   // its epilogue reads GX[t] (and the tail after it)
                }
                cudaError_t xe = x3 ? prop_3xtf32<true>(p, e, M3, m3_slab, X3, U, st) : cudaErrorNotSupported;
                if (xe == cudaSuccess) { TR(); }
                else if (xe != cudaErrorNotSupported) return fail(__func__, cudaGetErrorString(xe));
                else STEP_GEMM(0, CfgBig, true, false, p, e, 1);
            }
            // (b) gate: per node [B, K*H] x [K*H, 2H]
            memset(&p, 0, sizeof(p));
            p.splits = 1; p.Z2 = 1; p.KB = K;
            p.A = PHt; p.lda = H; p.sA1 = (long long)B * H; p.sAk = U; p.M = B; p.K = H;
            p.B = Wg + (long long)Cin * 2 * H; p.ldb = 2 * H; p.N = 2 * H; p.sB1 = (long long)K * I * 2 * H; p.sBk = (long long)I * 2 * H;
            if (bf) { p.A16 = PH16t; p.B16 = WG16 + (long long)Cin * 2 * H; p.keepB = 1; }
            p.need16 = skip32 ? 1 : 0;
            STEP_GEMM(1, CfgMid, true, false, p, (EpiGate{GXt, PHt, Zt, Rt_, PZt, B, H, tc ? 1 : 0, bf ? PZ16t : nullptr}), N);
            // (c) PZ[t,1..] = M * (z*h)
            {
                GemmP pp = prop_params(M, ldm, N, Kp, PZt, B * H);
                EpiPlain e = epi_plain(PZt + U, 0, 0, B * H);
                if (bf) { pp.A16 = M16; pp.B16 = PZ16t; e.C16 = PZ16t + U; }
                if (skip32) e.c_z2_hi = 0;
                if (bf && warm) {   // the candidate contraction that follows streams Wu16[n, k, Cin:, :]
                    add_pf(pp, WU16 + (long long)Cin * H, (long long)I * H * 2, (long long)H * H * 2, N * K);
                    if (l2_warm_level() >= 2) add_pf(pp, RXt, 0, (long long)3 * U * 4, 1);  // the fused tail reads RX[t]
                }
                cudaError_t xe = x3 ? prop_3xtf32<true>(pp, e, M3, m3_slab, X3, U, st) : cudaErrorNotSupported;
                if (xe == cudaSuccess) { TR(); }
                else if (xe != cudaErrorNotSupported) return fail(__func__, cudaGetErrorString(xe));
                else STEP_GEMM(2, CfgBig, true, false, pp, e, 1);
            }
            // (d) candidate
            p.A = PZt;
            p.B = Wu + (long long)Cin * H; p.ldb = H; p.N = H; p.sB1 = (long long)K * I * H; p.sBk = (long long)I * H;
            if (bf) { p.A16 = PZ16t; p.B16 = WU16 + (long long)Cin * H; p.keepB = 1; }
            // (d)+(e)+(f) in one launch when the fused tail applies (tensor-core engine, H = 64): see EpiCandRes
            if (tc && !use_multi && H == 64 && fused_tail_enabled()) {
                EpiCandRes ef{GXt, PHt, Rt_, HCt, H1t, B, H, 1, RXt, Z2t, R2t, ZH2t, HC2t, PHt + (long long)K * U, mix + t,
                              bf ? PH16t + (long long)K * U : nullptr, RgH, RuH};
                cudaError_t fe = cudaErrorNotSupported;
                if (bf) fe = launch_gemm_tc<64, true, false, EpiCandRes, true>(p, ef, N, st);
                if (fe == cudaErrorNotSupported && !p.need16) fe = launch_gemm_tc<64, true, false, EpiCandRes>(p, ef, N, st);
                if (fe == cudaErrorNotSupported && p.need16) return fail(__func__, "bf16 engine unavailable for a contraction whose fp32 operand was skipped");
                if (fe == cudaSuccess) {
                    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
                    TR();
                    continue;
                }
                if (fe != cudaErrorNotSupported) return fail(__func__, cudaGetErrorString(fe));
            }
            STEP_GEMM(3, CfgMid, true, false, p, (EpiCand{GXt, PHt, Rt_, HCt, H1t, B, H, tc ? 1 : 0}), N);
            // (e) residual gate: [N*B, H] x Rgw[:, Cin:]^T
            memset(&p, 0, sizeof(p));
            p.splits = 1; p.Z2 = 1; p.KB = 1;
            p.A = H1t; p.lda = H; p.M = N * B; p.K = H;
            p.B = RgH; p.ldb = H; p.N = 2 * H;
            STEP_GEMM(4, CfgMid, true, true, p, (EpiGate{RXt, H1t, Z2t, R2t, ZH2t, 0, H, tc ? 1 : 0}), 1);
            // (f) residual candidate + mix -> PH[t+1, 0]
            p.A = ZH2t;
            p.B = RuH; p.ldb = H; p.N = H;
            STEP_GEMM(5, CfgMid, true, true, p,
                      (EpiResCand{RXt, H1t, R2t, HC2t, PHt + (long long)K * U, mix + t, H, tc ? 1 : 0, bf ? PH16t + (long long)K * U : nullptr}), 1);
        }
    }
    tr.report("encoder_layer_fwd");
    return 0;
}
extern "C" int matgcn_encoder_layer_fwd(int T, int N, int B, int Cin, int H, int K, int ldm,
                                        const float* x, long long x_tstride, const float* h0, const float* M,
                                        const float* Wg, const float* bg, const float* Wu, const float* bu,
                                        const float* Rgw, const float* Rgb, const float* Ruw, const float* Rub,
                                        const float* mix, float* ws, int flags, void* stream) {
    return encoder_layer_fwd_impl(T, N, B, Cin, H, K, ldm, x, x_tstride, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix, ws, flags,
                                  nullptr, stream);
}
extern "C" int matgcn_encoder_layer_fwd_chained(int T, int N, int B, int Cin, int H, int K, int ldm,
                                                const float* x, long long x_tstride, void* x16_chain, const float* h0, const float* M,
                                                const float* Wg, const float* bg, const float* Wu, const float* bu,
                                                const float* Rgw, const float* Rgb, const float* Ruw, const float* Rub,
                                                const float* mix, float* ws, int flags, void* stream) {
    REQUIRE(x16_chain, "null chain pointer");
    return encoder_layer_fwd_impl(T, N, B, Cin, H, K, ldm, x, x_tstride, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix, ws, flags,
                                  x16_chain, stream);
}

// ------------------------------------------------------------------------------------------
// encoder layer backward
// ------------------------------------------------------------------------------------------
static int encoder_layer_bwd_impl(int T, int N, int B, int Cin, int H, int K, int ldm, int n_adp,
                                        const float* dy, long long dy_tstride, const float* M,
                                        const float* Wg, const float* Wu, const float* Rgw, const float* Ruw,
                                        const float* mix, float* ws, float* bws,
                                        float* dx, float* dh0, float* dM,
                                        float* dWg, float* dbg, float* dWu, float* dbu,
                                        float* dRgw, float* dRgb, float* dRuw, float* dRub, float* dmix,
                                        int flags, const float* x_chain, const void* x16_chain, void* stream) {
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    REQUIRE(dy && M && Wg && Wu && Rgw && Ruw && mix && ws && bws, "null pointer");
    REQUIRE(dx && dM && dWg && dbg && dWu && dbu && dRgw && dRgb && dRuw && dRub && dmix, "null output pointer");
    if (check_layer_dims(T, N, B, Cin, H, K, ldm)) return -1;
    REQUIRE(n_adp >= 0 && n_adp <= K - 1, "n_adp out of range");
    cudaStream_t st = (cudaStream_t)stream;
    Tracer tr(st);
    const LayerWs w = layer_ws(T, N, B, Cin, H, K);
    const LayerBws bw = layer_bws(T, N, B, Cin, H, K, n_adp);
    const int Kp = K - 1, I = Cin + H;
    const long long U = (long long)w.U, UX = (long long)w.UX;
    float* PX = ws + w.PX; float* DG = ws + w.GX; float* DR = ws + w.RX; float* PH = ws + w.PH; float* PZ = ws + w.PZ;
    if (x16_chain) {   // the forward pass read its input where the previous layer left it (encoder_layer_fwd_impl)
        REQUIRE(x_chain && chain_ok(T, N, B, Cin, H, K, ldm, flags), "chained input does not qualify");
        PX = const_cast<float*>(x_chain);   // read only
    }
    float* DPX = bws + bw.DPX; float* DPT = bws + bw.DPT; float* DPHA = bws + bw.DPHA; float* DPZA = bws + bw.DPZA;
    float* DH1 = bws + bw.DH1; float* DHD = bws + bw.DHD; float* DHC = bws + bw.DHC; float* DRES = bws + bw.DRES;
    const int NB = N * B;

    const size_t m16_cap = ((size_t)Kp * N * 8 * ((N + 7) / 8 + 1));
    const bool bf = tc && (flags & MATGCN_FLAG_BF16) != 0 && (size_t)Kp * N * ldm <= m16_cap;
    const __nv_bfloat16* M16 = reinterpret_cast<const __nv_bfloat16*>(ws + w.M16);  // written by the forward pass
    __nv_bfloat16* DPT16 = reinterpret_cast<__nv_bfloat16*>(bws + bw.DPT16);
    __nv_bfloat16* DPX16 = reinterpret_cast<__nv_bfloat16*>(bws + bw.DPX16);
    __nv_bfloat16* DG16T = bf ? reinterpret_cast<__nv_bfloat16*>(bws + bw.DG16) : nullptr;
    const __nv_bfloat16* WG16 = reinterpret_cast<const __nv_bfloat16*>(ws + w.WG16);  // written by the forward pass
    const __nv_bfloat16* WU16 = reinterpret_cast<const __nv_bfloat16*>(ws + w.WU16);
    CK(cudaMemsetAsync(DHC, 0, sizeof(float) * U, st));
    TR();
    CK(cudaMemsetAsync(dmix, 0, sizeof(float) * T, st));
    TR();
    GemmP p;
    constexpr bool use_multi = false;
    // same rule as the forward pass: in bf16 mode the propagated slots k >= 1 (PH / PZ / PX there, DPT here) exist only as bf16 twins
    const bool skip32 = bf && !use_multi && !(H & 7) && !(ldm & 7) && B >= 8;
    const bool warm = l2_warm_layer(N, K, Cin, H, B);
    // exact mode: the dense transposed propagations B4 / B6 as 3xTF32 against the [hi, hi, lo] slabs the forward pass left in ws
    const long long m3_slab = (long long)Kp * N * ldm;
    const bool x3 = !tc && exact_tc_enabled() && N <= EXACT_TC_MAX_N && !(ldm & 3) && !((B * H) & 3) && aligned16(M) &&
                    (size_t)m3_slab <= (size_t)Kp * N * 8 * ((N + 7) / 8 + 1);
    const float* M3 = x3 ? ws + w.M3 : nullptr;
    float* DPT3 = x3 ? bws + bw.DPT3 : nullptr;
    bool rec_done = false, dg16_only = false;
    if (skip32 && H == 64 && rec_flag() && fused_tail_enabled()) {
        // the whole reverse-time recurrence as one persistent cooperative launch (rec_bwd.cuh)
        CK(cudaMemsetAsync(DH1, 0, sizeof(float) * U, st));    // DHD2: DHD + DPT[0] of the step after (none yet)
        CK(cudaMemsetAsync(DRES, 0, sizeof(float) * U, st));   // DZ / DC: the dense phases' output (no carry yet)
        // inner layers in bf16 mode: every consumer of the main cell's pre-activation gradients reads the bf16 twin, and the
        // reverse kernel is bound in part by its store traffic - no fp32 copy then (MATGCN_DG32=1 keeps it)
        dg16_only = !xside_small_ok(Cin, H, K) && !(Cin & 7) && !(H & 7) && 3 * H / 8 <= 256 && !dg32_forced();
        RecBwdArgs ra{T, N, B, Cin, K, ldm, n_adp, dy, dy_tstride, M16, WG16, WU16,
                      PH, ws + w.Z, ws + w.R, ws + w.HC, ws + w.H1, ws + w.Z2, ws + w.R2, ws + w.HC2,
                      ws + w.RGH, ws + w.RUH, mix, dg16_only ? nullptr : DG, DR, DG16T, DPT, DPT16, DHD, DH1, DRES,
                      reinterpret_cast<__nv_bfloat16*>(DPZA), reinterpret_cast<__nv_bfloat16*>(DPHA), DHC, dmix,
                      reinterpret_cast<unsigned int*>(bws + bw.MPH)};
        const cudaError_t re = launch_rec_bwd(ra, st);
        if (re == cudaSuccess) {
            g_tc_launches.fetch_add(1, std::memory_order_relaxed);
            TR();
            rec_done = true;
        } else if (re != cudaErrorNotSupported) {
            return fail(__func__, cudaGetErrorString(re));
        }
        if (!rec_done) dg16_only = false;
    }
    if (!rec_done) {
        for (int t = T - 1; t >= 0; --t) {
            const float* PHt = PH + (long long)t * K * U;
            const float* Zt = ws + w.Z + t * U; const float* Rt_ = ws + w.R + t * U; const float* HCt = ws + w.HC + t * U;
            const float* H1t = ws + w.H1 + t * U; const float* Z2t = ws + w.Z2 + t * U; const float* R2t = ws + w.R2 + t * U;
            const float* HC2t = ws + w.HC2 + t * U;
            float* DGt = DG + (long long)t * 3 * U; float* DRt = DR + (long long)t * 3 * U;
            __nv_bfloat16* DG16 = bf ? DG16T + (long long)t * 3 * U : nullptr;
            // B0 + B1 + B2 in one launch when the fused form applies (fast modes, H = 64): see res_bwd.cuh
            bool rb_fused = false;
            if (tc && !use_multi && fused_tail_enabled()) {
                ResBwdArgs ra{dy + (long long)t * dy_tstride, DHC, H1t, R2t, HC2t, Z2t, PHt, Rt_, HCt, mix + t, ws + w.RUH, ws + w.RGH,
                              DRt, DGt, DG16, DHD, dmix + t, (long long)NB};
                if (res_bwd_fused_ok(ra, H)) {
                    CK(launch_res_bwd_fused(ra, st));
                    TR();
                    rb_fused = true;
                }
            }
            if (!rb_fused) {
                // B0
                {
                    const float* dyt = dy + (long long)t * dy_tstride;
                    if (!(H & 3) && !(U & 3) && aligned16(dyt) && aligned16(DHC) && aligned16(H1t) && aligned16(R2t) && aligned16(HC2t) &&
                        aligned16(DH1) && aligned16(DRES) && aligned16(DRt))
                        bwd_head4_kernel<<<(unsigned)((U / 4 + 255) / 256), 256, 0, st>>>(dyt, DHC, H1t, R2t, HC2t, mix + t, U / 4, H, DH1,
                                                                                         DRES, DRt, dmix + t);
                    else
                        bwd_head_kernel<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(dyt, DHC, H1t, R2t, HC2t, mix + t, U, H, DH1, DRES,
                                                                                    DRt, dmix + t);
                    count_launch();
                    TR();
                    CK(cudaGetLastError());
                }
                // B1: dzh2 = da3 [NB,H] * Ruw[:, Cin:]  (B element (k=o, n=j) at o*H + j of the dense copy)
                memset(&p, 0, sizeof(p));
                p.splits = 1; p.Z2 = 1; p.KB = 1;
                p.A = DRt + 2 * H; p.lda = 3 * H; p.M = NB; p.K = H;
                p.B = ws + w.RUH; p.ldb = H; p.N = H;
                STEP_GEMM(0, CfgMid, true, false, p, (EpiB1{DH1, DRt, DRES, H1t, Z2t, R2t, HC2t, H}), 1);
                // B2: dh1 += da2 [NB,2H] * Rgw[:, Cin:]
                p.A = DRt; p.K = 2 * H;
                p.B = ws + w.RGH;
                STEP_GEMM(1, CfgMid, true, false, p, (EpiB2{DH1, PHt, Rt_, HCt, DHD, DGt, H, DG16}), 1);
            }
            // B3: DPT[k][n] = dau[n] [B,H] * Wu[n,k,Cin:,:]^T      z = (n, k)
            memset(&p, 0, sizeof(p));
            p.splits = 1; p.Z2 = K; p.KB = 1;
            p.A = DGt + 2 * H; p.lda = 3 * H; p.sA1 = (long long)B * 3 * H; p.sA2 = 0; p.M = B; p.K = H;
            p.B = Wu + (long long)Cin * H; p.ldb = H; p.N = H; p.sB1 = (long long)K * I * H; p.sB2 = (long long)I * H;
            if (bf) { p.A16 = DG16 + 2 * H; p.B16 = WU16 + (long long)Cin * H; p.keepB = 1; }
            {
                EpiPlain e = epi_plain(DPT, (long long)B * H, U, H);
                if (bf) e.C16 = DPT16;
                if (skip32) e.c_z2_hi = 1;   // DPT[k >= 1] is consumed as its bf16 twin only (B4); DPT[0] stays fp32 (EpiB4 reads it)
                // the adaptive slices are also kept per step (operands of dM): second destination instead of a copy
                if (n_adp && !use_multi) {
                    // (bf16 mode: kept as bf16 in the same storage - the dM contraction reads it as a bf16 operand)
                    if (skip32) e.D2h = reinterpret_cast<__nv_bfloat16*>(DPZA) + (long long)t * n_adp * U;
                    else e.D2 = DPZA + (long long)t * n_adp * U;
                    e.d2_lo = 1; e.d2_hi = 1 + n_adp;
                }
                STEP_GEMM(2, CfgMid, true, true, p, e, N * K);
            }
            // B4: dzh = DPT[0] + sum_{k>=1} M_k^T DPT[k]
            memset(&p, 0, sizeof(p));
            p.splits = 1; p.Z2 = 1; p.KB = 1;
            p.A = M; p.lda = ldm; p.M = N; p.K = Kp * N;
            p.B = DPT + U; p.ldb = B * H; p.N = B * H;
            if (bf) { p.A16 = M16; p.B16 = DPT16 + U; }
            p.need16 = skip32 ? 1 : 0;
            if (bf && warm) {   // B5 streams Wg16[n, k, Cin:, :]
                add_pf(p, WG16 + (long long)Cin * 2 * H, (long long)I * 2 * H * 2, (long long)H * 2 * H * 2, N * K);
            }
            {
                const EpiB4 e4{DPT, PHt, Zt, DHD, DGt, H, B * H, DG16};
                cudaError_t xe = x3 ? prop_3xtf32<false>(p, e4, M3, m3_slab, DPT3, (long long)Kp * U, st) : cudaErrorNotSupported;
                if (xe == cudaSuccess) { TR(); }
                else if (xe != cudaErrorNotSupported) return fail(__func__, cudaGetErrorString(xe));
                else STEP_GEMM(3, CfgBig, false, false, p, e4, 1);
            }
            // B5: DPT[k][n] = dag[n] [B,2H] * Wg[n,k,Cin:,:]^T
            memset(&p, 0, sizeof(p));
            p.splits = 1; p.Z2 = K; p.KB = 1;
            p.A = DGt; p.lda = 3 * H; p.sA1 = (long long)B * 3 * H; p.M = B; p.K = 2 * H;
            p.B = Wg + (long long)Cin * 2 * H; p.ldb = 2 * H; p.N = H; p.sB1 = (long long)K * I * 2 * H; p.sB2 = (long long)I * 2 * H;
            if (bf) { p.A16 = DG16; p.B16 = WG16 + (long long)Cin * 2 * H; p.keepB = 1; }
            {
                EpiPlain e = epi_plain(DPT, (long long)B * H, U, H);
                if (bf) e.C16 = DPT16;
                if (skip32) e.c_z2_hi = 1;
                if (n_adp && !use_multi) {
                    if (skip32) e.D2h = reinterpret_cast<__nv_bfloat16*>(DPHA) + (long long)t * n_adp * U;
                    else e.D2 = DPHA + (long long)t * n_adp * U;
                    e.d2_lo = 1; e.d2_hi = 1 + n_adp;
                }
                STEP_GEMM(4, CfgMid, true, true, p, e, N * K);
            }
            // B6: carry = DHD + DPT[0] + sum M_k^T DPT[k]
            memset(&p, 0, sizeof(p));
            p.splits = 1; p.Z2 = 1; p.KB = 1;
            p.A = M; p.lda = ldm; p.M = N; p.K = Kp * N;
            p.B = DPT + U; p.ldb = B * H; p.N = B * H;
            if (bf) { p.A16 = M16; p.B16 = DPT16 + U; }
            p.need16 = skip32 ? 1 : 0;
            if (bf && warm && t > 0) {   // B3 of the next reverse step streams Wu16[n, k, Cin:, :]
                add_pf(p, WU16 + (long long)Cin * H, (long long)I * H * 2, (long long)H * H * 2, N * K);
                if (l2_warm_level() >= 2) {   // ... and its fused head reads the saved activations of step t-1
                    const long long ub = (long long)U * 4;
                    add_pf(p, ws + w.H1 + (t - 1) * U, 0, ub, 1); add_pf(p, ws + w.R2 + (t - 1) * U, 0, ub, 1);
                    add_pf(p, ws + w.HC2 + (t - 1) * U, 0, ub, 1); add_pf(p, ws + w.Z2 + (t - 1) * U, 0, ub, 1);
                    add_pf(p, PH + (long long)(t - 1) * K * U, 0, ub, 1); add_pf(p, ws + w.R + (t - 1) * U, 0, ub, 1);
                    add_pf(p, ws + w.HC + (t - 1) * U, 0, ub, 1);
                }
            }
            {
                const EpiB6 e6{DPT, DHD, DHC, B * H};
                cudaError_t xe = x3 ? prop_3xtf32<false>(p, e6, M3, m3_slab, DPT3, (long long)Kp * U, st) : cudaErrorNotSupported;
                if (xe == cudaSuccess) { TR(); }
                else if (xe != cudaErrorNotSupported) return fail(__func__, cudaGetErrorString(xe));
                else STEP_GEMM(5, CfgBig, false, false, p, e6, 1);
            }
        }
    }
    if (dh0) CK(cudaMemcpyAsync(dh0, DHC, sizeof(float) * U, cudaMemcpyDeviceToDevice, st));

    // ---- time-batched parameter gradients ------------------------------------------------
    // dWg[n,k,Cin:,:] = sum_{t,b} PH[t,k,n]^T DG[t,n][:, 0:2H]        z = (n,k), k-batches = t
    memset(&p, 0, sizeof(p));
    p.splits = 1; p.Z2 = K; p.KB = T;
    p.lda = H; p.sA1 = (long long)B * H; p.sA2 = U; p.sAk = K * U; p.M = H; p.K = B;
    p.ldb = 3 * H; p.sB1 = (long long)B * 3 * H; p.sB2 = 0; p.sBk = 3 * U;
    p.A = PH; p.B = DG; p.N = 2 * H;
    const __nv_bfloat16* PH16 = reinterpret_cast<const __nv_bfloat16*>(ws + w.PH16);  // bf16 twins written by the forward pass
    const __nv_bfloat16* PZ16 = reinterpret_cast<const __nv_bfloat16*>(ws + w.PZ16);
    const __nv_bfloat16* PX16 = x16_chain ? reinterpret_cast<const __nv_bfloat16*>(x16_chain) : reinterpret_cast<const __nv_bfloat16*>(ws + w.PX16);
    if (bf) { p.A16 = PH16; p.B16 = DG16T; }   // these contractions stream the saved state: half the bytes with the twins
    p.need16 = skip32 ? 1 : 0;                 // (and in that case the forward never wrote the fp32 PH / PZ slots k >= 1)
    CK((gemm_any<CfgMid, false, false>(tc, p, epi_store(dWg + (long long)Cin * 2 * H, (long long)K * I * 2 * H, (long long)I * 2 * H, 2 * H), N * K, st)));
    TR();
    p.A = PZ; p.B = DG + 2 * H; p.N = H;
    if (bf) { p.A16 = PZ16; p.B16 = DG16T + 2 * H; }
    CK((gemm_any<CfgMid, false, false>(tc, p, epi_store(dWu + (long long)Cin * H, (long long)K * I * H, (long long)I * H, H), N * K, st)));
    TR();
    p.A16 = nullptr; p.B16 = nullptr; p.need16 = 0;
    // residual-GRU weight gradients accumulate by atomics: clear them first
    CK(cudaMemsetAsync(dRgw, 0, sizeof(float) * (size_t)2 * H * I, st));
    TR();
    CK(cudaMemsetAsync(dRuw, 0, sizeof(float) * (size_t)H * I, st));
    TR();
    CK(cudaMemsetAsync(dRgb, 0, sizeof(float) * 2 * H, st));
    TR();
    CK(cudaMemsetAsync(dRub, 0, sizeof(float) * H, st));
    TR();
    const bool small_x = xside_small_ok(Cin, H, K);
    // fast modes, H = 64: ONE pass over DR for the residual-cell weight / bias gradients (dr_pass.cuh) instead of four split-K
    // contractions and a column-sum kernel (MATGCN_DR_PASS=0 keeps those)
    bool dr_hidden_done = false, dr_x_done = false, dr_bias_done = false;
    if (tc && H == 64 && dr_pass_enabled()) {
        const bool with_x = !small_x && Cin == 64;
        DrPassArgs da{T, (int)NB, Cin, H, DR, ws + w.H1, ws + w.ZH2, with_x ? PX : nullptr, K * UX,
                      dRgw, dRuw, small_x ? nullptr : dRgb, small_x ? nullptr : dRub};
        const cudaError_t de = launch_dr_pass(da, st);
        if (de == cudaSuccess) {
            count_launch();
            g_tc_launches.fetch_add(1, std::memory_order_relaxed);
            TR();
            dr_hidden_done = true; dr_x_done = with_x; dr_bias_done = !small_x;
        } else if (de != cudaErrorNotSupported) {
            CK(de);
        }
    }
    bool dpx16_only = false;   // the fused DPX launch keeps slots k >= 1 as bf16 only: their consumers must read the twins
    const int cs_threads = 3 * H <= 256 ? 256 : 3 * H;
    REQUIRE(3 * H <= 1024, "hidden size too large for the column-sum kernels");
    if (small_x) {
        // tiny channel count (layer 0): one pass over DG (input-row weight gradient, bias gradient, DPX) and one pass
        // over DR (residual input-column gradients, residual bias gradient, residual share of dx into DPX[t,0])
        if (tc && xside_mma_ok(Cin, H, K, DG, PX, DPX) && aligned16(DR)) {
            // fast modes: the same two passes as warp-level TF32 mma.sync products (xside_mma.cuh)
            XsMmaArgs xa{DG, PX, DPX, Wg, Wu, dWg, dWu, dbg, dbu, T, N, B, Cin, H, K};
            CK(launch_xside_bwd_mma<true>(xa, st));
            TR();
            XsMmaArgs xr{DR, PX, DPX, Rgw, Ruw, dRgw, dRuw, dRgb, dRub, T, N, B, Cin, H, K};
            CK(launch_xside_bwd_mma<false>(xr, st));
            TR();
        } else {
            xside_bwd_dg_small_kernel<<<N, 256, 0, st>>>(PX, DG, Wg, Wu, T, N, B, Cin, H, K, dWg, dWu, dbg, dbu, DPX);
            count_launch();
            TR();
            xside_bwd_dr_small_kernel<<<592, 256, 0, st>>>(PX, DR, Rgw, Ruw, T, N, B, Cin, H, K, dRgw, dRuw, dRgb, dRub, DPX);
            count_launch();
            TR();
            CK(cudaGetLastError());
        }
    } else {
        // input rows 0:Cin from PX, and the bias gradients (column sums of DG over (t, b))
        p.lda = Cin; p.sA1 = (long long)B * Cin; p.sA2 = UX; p.sAk = K * UX; p.M = Cin;
        p.A = PX; p.B = DG; p.N = 2 * H;
        if (bf) { p.A16 = PX16; p.B16 = DG16T; }
        p.need16 = (skip32 && !(Cin & 7)) ? 1 : 0;
        CK((gemm_any<CfgMid, false, false>(tc, p, epi_store(dWg, (long long)K * I * 2 * H, (long long)I * 2 * H, 2 * H), N * K, st)));
        TR();
        p.B = DG + 2 * H; p.N = H;
        if (bf) p.B16 = DG16T + 2 * H;
        CK((gemm_any<CfgMid, false, false>(tc, p, epi_store(dWu, (long long)K * I * H, (long long)I * H, H), N * K, st)));
        TR();
        p.A16 = nullptr; p.B16 = nullptr; p.need16 = 0;
        CK(cudaMemsetAsync(dbg, 0, sizeof(float) * (size_t)N * 2 * H, st));
        TR();
        CK(cudaMemsetAsync(dbu, 0, sizeof(float) * (size_t)N * H, st));
        TR();
        const size_t cs_smem = sizeof(float4) * (256 / (3 * H / 4 > 0 ? 3 * H / 4 : 1)) * (3 * H / 4);
        dim3 g1(N, 8);
        if (dg16_only)
            colsum8_bf16_kernel<<<g1, 256, cs_smem * 2, st>>>(DG16T, T, 3 * U, (long long)B * 3 * H, B, 3 * H, 2 * H, dbg, 2 * H, dbu, H);
        else if (colsum4_ok(DG, 3 * U, (long long)B * 3 * H, 3 * H))
            colsum4_kernel<<<g1, 256, cs_smem, st>>>(DG, T, 3 * U, (long long)B * 3 * H, B, 3 * H, 2 * H, dbg, 2 * H, dbu, H);
        else
            colsum_kernel<<<g1, cs_threads, 0, st>>>(DG, T, 3 * U, (long long)B * 3 * H, B, 3 * H, 3 * H, 2 * H, dbg, 2 * H, dbu, H);
        count_launch();
        TR();
        dim3 g2(1, 1184);
        if (!dr_bias_done) {
            if (colsum4_ok(DR, 3 * U, 0, 3 * H))
                colsum4_kernel<<<g2, 256, cs_smem, st>>>(DR, T, 3 * U, 0, NB, 3 * H, 2 * H, dRgb, 2 * H, dRub, H);
            else
                colsum_kernel<<<g2, cs_threads, 0, st>>>(DR, T, 3 * U, 0, NB, 3 * H, 3 * H, 2 * H, dRgb, 2 * H, dRub, H);
            count_launch();
            TR();
        }
        CK(cudaGetLastError());
        // DPX[t,k,n] = DG[t,n][:,0:2H] * Wg[n,k,0:Cin,:]^T + DG[t,n][:,2H:] * Wu[n,k,0:Cin,:]^T     per k: z = (t, n)
        bool dpx_done = false;
        if (bf && !(Cin & 7) && !(H & 3)) {
            // bf16 mode: one contraction over all 3H pre-activation gradients per support, against the packed input-row weights
            // (the packed input-row weights WX16 [N, K, Cin, 3H] were written by the forward pass)
            const __nv_bfloat16* WX16 = reinterpret_cast<const __nv_bfloat16*>(ws + w.WX16);
            dpx_done = false;
            if (!(Cin & 31)) {
                // all K supports in ONE launch: the output columns are K blocks of Cin (EpiBlocks), DG is streamed once
                memset(&p, 0, sizeof(p));
                p.splits = 1; p.Z2 = N; p.KB = 1;
                p.A = DG; p.lda = 3 * H; p.sA1 = 3 * U; p.sA2 = (long long)B * 3 * H; p.M = B; p.K = 3 * H;
                p.A16 = DG16T;
                p.B16 = WX16; p.ldb = 3 * H; p.N = K * Cin; p.sB1 = 0; p.sB2 = (long long)K * Cin * 3 * H;
                EpiBlocks e{DPX, K * UX, (long long)B * Cin, UX, Cin, Cin, DPX16, 1, (ldm & 7) ? K : 1};  // (slots k >= 1: bf16 only, see dpx16_only)
                const cudaError_t de = launch_gemm_tc<128, true, true, EpiBlocks, true>(p, e, T * N, st);
                if (de == cudaSuccess) {
                    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
                    TR();
                    dpx_done = true;
                } else if (de != cudaErrorNotSupported) {
                    CK(de);
                }
            }
            const bool fused = dpx_done;
            dpx16_only = fused && !(ldm & 7);
            dpx_done = true;
            for (int k = 0; k < K && dpx_done && !fused; ++k) {
                memset(&p, 0, sizeof(p));
                p.splits = 1; p.Z2 = N; p.KB = 1;
                p.A = DG; p.lda = 3 * H; p.sA1 = 3 * U; p.sA2 = (long long)B * 3 * H; p.M = B; p.K = 3 * H;
                p.A16 = DG16T;
                p.B16 = WX16 + (long long)k * Cin * 3 * H; p.ldb = 3 * H; p.N = Cin; p.sB1 = 0; p.sB2 = (long long)K * Cin * 3 * H;
                EpiStore e = epi_store(DPX + (long long)k * UX, K * UX, (long long)B * Cin, Cin);
                if (k >= 1) e.C16 = DPX16 + (long long)k * UX;
                const cudaError_t de = Cin <= 64 ? launch_gemm_tc<64, true, true, EpiStore, true>(p, e, T * N, st)
                                                 : launch_gemm_tc<128, true, true, EpiStore, true>(p, e, T * N, st);
                if (de == cudaErrorNotSupported && k == 0) { dpx_done = false; break; }  // fall back to the two-pass form below
                CK(de);
                g_tc_launches.fetch_add(1, std::memory_order_relaxed);
                TR();
            }
        }
        REQUIRE(dpx_done || !dg16_only, "the input-gradient contraction needs the fp32 pre-activation gradients, which were not written");
        for (int k = 0; k < K && !dpx_done; ++k) {
            memset(&p, 0, sizeof(p));
            p.splits = 1; p.Z2 = N; p.KB = 1;
            p.A = DG; p.lda = 3 * H; p.sA1 = 3 * U; p.sA2 = (long long)B * 3 * H; p.M = B; p.K = 2 * H;
            p.B = Wg + (long long)k * I * 2 * H; p.ldb = 2 * H; p.N = Cin; p.sB1 = 0; p.sB2 = (long long)K * I * 2 * H;
            EpiStore e = epi_store(DPX + (long long)k * UX, K * UX, (long long)B * Cin, Cin);
            if (bf && k >= 1) e.C16 = DPX16 + (long long)k * UX;  // (the accumulating second launch writes the final twin)
            if (bf) { p.A16 = DG16T; p.B16 = WG16 + (long long)k * I * 2 * H; }
            CK((gemm_any<CfgMid, true, true>(tc, p, e, T * N, st)));
            TR();
            p.A = DG + 2 * H; p.K = H;
            p.B = Wu + (long long)k * I * H; p.ldb = H; p.sB2 = (long long)K * I * H;
            if (bf) { p.A16 = DG16T + 2 * H; p.B16 = WU16 + (long long)k * I * H; }
            e.accumulate = 1;
            CK((gemm_any<CfgMid, true, true>(tc, p, e, T * N, st)));
            TR();
        }
    }
    // dx[t] = DPX[t,0] + sum_{k>=1} M_k^T DPX[t,k] + DR[t][:,0:2H]*Rgw[:,0:Cin] + DR[t][:,2H:]*Ruw[:,0:Cin]
    memset(&p, 0, sizeof(p));
    p.splits = 1; p.Z2 = 1; p.KB = 1;
    p.A = M; p.lda = ldm; p.M = N; p.K = Kp * N;
    p.B = DPX + UX; p.ldb = B * Cin; p.N = B * Cin; p.sB1 = K * UX;
    if (bf && !small_x) { p.A16 = M16; p.B16 = DPX16 + UX; }
    p.need16 = dpx16_only ? 1 : 0;
    {
        EpiStore e = epi_store(dx, UX, 0, B * Cin);
        e.add = DPX; e.add_s1 = K * UX; e.add_ld = B * Cin;
        const cudaError_t le = tc ? launch_store_lean<ES_ADD, 128, false, false, true>(p, e, T, st) : cudaErrorNotSupported;
        if (le == cudaErrorNotSupported) CK((gemm_any<CfgBig, false, false>(tc, p, e, T, st)));
        else CK(le);
        TR();
    }
    memset(&p, 0, sizeof(p));
    p.splits = 1; p.Z2 = 1; p.KB = 1;
    p.A = DR; p.lda = 3 * H; p.sA1 = 3 * U; p.M = NB; p.K = 2 * H;
    p.B = Rgw; p.ldb = I; p.N = Cin;
    if (!small_x) {
        // one pass over DR (475 MB per layer at the Baltimore shape) instead of two: [Rgw[:, 0:Cin]; Ruw[:, 0:Cin]] as one
        // [3H, Cin] operand
        const float* R3 = ws + w.R3X;   // (packed by the forward pass)
        p.K = 3 * H; p.B = R3; p.ldb = Cin;
        EpiStore e = epi_store(dx, UX, 0, Cin);
        e.accumulate = 1;
        const cudaError_t le = (tc && Cin <= 64) ? launch_store_lean<ES_ACC, 64, true, false, false>(p, e, T, st) : cudaErrorNotSupported;
        if (le == cudaErrorNotSupported) CK((gemm_any<CfgMid, true, false>(tc, p, e, T, st)));
        else CK(le);
        TR();
    }
    // dM[a] = sum_t DPHA[t,a] PH[t,0]^T + DPZA[t,a] PZ[t,0]^T + DPX[t,a+1] PX[t,0]^T     (split-K, atomics)
    CK(cudaMemsetAsync(dM, 0, sizeof(float) * (size_t)Kp * N * ldm, st));
    TR();
    for (int a = 0; a < n_adp; ++a) {
        EpiAtomic ea{dM + (long long)a * N * ldm, 0, 0, ldm};
        memset(&p, 0, sizeof(p));
        p.Z2 = 1; p.KB = T; p.M = N; p.N = N;
        p.K = B * H; p.lda = B * H; p.ldb = B * H;
        p.splits = T;
        p.A = DPHA + (long long)a * U; p.sAk = (long long)n_adp * U; p.B = PH; p.sBk = K * U;
        if (skip32) {   // the per-step adaptive slices were kept as bf16 (D2h above); h_t has its bf16 twin in slot 0 of PH16
            p.A16 = reinterpret_cast<const __nv_bfloat16*>(DPHA) + (long long)a * U; p.B16 = PH16; p.need16 = 1;
        }
        // (exact mode: 3xTF32 on the tensor-core engine when the shape qualifies, else the FFMA kernel)
        auto dm_term = [&](const GemmP& q) -> cudaError_t {
            cudaError_t xe = (x3 && bw.TB3_floats) ? tb_3xtf32(q, ea, bws + bw.TB3, (long long)bw.TB3_floats, st) : cudaErrorNotSupported;
            if (xe == cudaErrorNotSupported) xe = gemm_any<CfgBig, true, true>(tc, q, ea, 1, st);
            return xe;
        };
        CK(dm_term(p));
        TR();
        p.A = DPZA + (long long)a * U; p.B = PZ;
        if (skip32) { p.A16 = reinterpret_cast<const __nv_bfloat16*>(DPZA) + (long long)a * U; p.B16 = PZ16; }
        CK(dm_term(p));
        TR();
        p.K = B * Cin; p.lda = B * Cin; p.ldb = B * Cin;
        p.A = DPX + (long long)(a + 1) * UX; p.sAk = K * UX; p.B = PX; p.sBk = K * UX;
        p.A16 = nullptr; p.B16 = nullptr; p.need16 = 0;
        if (bf && !small_x && !(Cin & 7)) { p.A16 = DPX16 + (long long)(a + 1) * UX; p.B16 = PX16; }
        p.need16 = dpx16_only ? 1 : 0;
        CK(dm_term(p));
        TR();
    }
    // residual GRU weights: dRgw[:, Cin:] = sum DR[:,0:2H]^T H1 ; dRgw[:, 0:Cin] = sum DR[:,0:2H]^T x ; same for Ruw
    {
        memset(&p, 0, sizeof(p));
        p.Z2 = 1; p.KB = T; p.K = NB; p.lda = 3 * H; p.sAk = 3 * U;
        int splits = (int)(((long long)T * NB + 4095) / 4096);
        if (splits > 592) splits = 592;
        if (splits < 1) splits = 1;
        p.splits = splits;
        if (!dr_hidden_done) {
            // gate, hidden columns
            p.A = DR; p.M = 2 * H; p.B = ws + w.H1; p.ldb = H; p.sBk = U; p.N = H;
            CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dRgw + Cin, 0, 0, I}, 1, st)));
            TR();
            // candidate, hidden columns
            p.A = DR + 2 * H; p.M = H; p.B = ws + w.ZH2;
            CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dRuw + Cin, 0, 0, I}, 1, st)));
            TR();
        }
        // input columns
        p.A = DR + 2 * H; p.M = H;
        if (!small_x && !dr_x_done) {
            p.B = PX; p.ldb = Cin; p.sBk = K * UX; p.N = Cin;
            CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dRuw, 0, 0, I}, 1, st)));
            TR();
            p.A = DR; p.M = 2 * H;
            CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dRgw, 0, 0, I}, 1, st)));
            TR();
        }
    }
    tr.report("encoder_layer_bwd");
    return 0;
}
extern "C" int matgcn_encoder_layer_bwd(int T, int N, int B, int Cin, int H, int K, int ldm, int n_adp,
                                        const float* dy, long long dy_tstride, const float* M,
                                        const float* Wg, const float* Wu, const float* Rgw, const float* Ruw,
                                        const float* mix, float* ws, float* bws,
                                        float* dx, float* dh0, float* dM,
                                        float* dWg, float* dbg, float* dWu, float* dbu,
                                        float* dRgw, float* dRgb, float* dRuw, float* dRub, float* dmix,
                                        int flags, void* stream) {
    return encoder_layer_bwd_impl(T, N, B, Cin, H, K, ldm, n_adp, dy, dy_tstride, M, Wg, Wu, Rgw, Ruw, mix, ws, bws, dx, dh0, dM, dWg, dbg,
                                  dWu, dbu, dRgw, dRgb, dRuw, dRub, dmix, flags, nullptr, nullptr, stream);
}
extern "C" int matgcn_encoder_layer_bwd_chained(int T, int N, int B, int Cin, int H, int K, int ldm, int n_adp,
                                                const float* dy, long long dy_tstride, const float* M,
                                                const float* Wg, const float* Wu, const float* Rgw, const float* Ruw,
                                                const float* mix, float* ws, float* bws,
                                                float* dx, float* dh0, float* dM,
                                                float* dWg, float* dbg, float* dWu, float* dbu,
                                                float* dRgw, float* dRgb, float* dRuw, float* dRub, float* dmix,
                                                int flags, const float* x_chain, const void* x16_chain, void* stream) {
    REQUIRE(x_chain && x16_chain, "null chain pointer");
    return encoder_layer_bwd_impl(T, N, B, Cin, H, K, ldm, n_adp, dy, dy_tstride, M, Wg, Wu, Rgw, Ruw, mix, ws, bws, dx, dh0, dM, dWg, dbg,
                                  dWu, dbu, dRgw, dRgb, dRuw, dRub, dmix, flags, x_chain, x16_chain, stream);
}

// ------------------------------------------------------------------------------------------
// gcn_off ablation: the layer is a plain GRU whose two nn.Linear are shared by all nodes (MA.py:134-153 used as
// the main cell, MA.py:187-192).  It is the residual-GRU half of the full layer with the hidden state itself
// as input, so it reuses the same epilogues (with a constant mix of 0: y = GRU(x, h)).
//   workspace: RX [T,NB,3H] (-> DR in backward) | HS [T+1,NB,H] (HS[0] = h0, HS[t+1] = h_t) | Z2,R2,HC2,ZH2 [T,NB,H]
//              | GH [2H,H] | UH [H,H] | zero, scratch
// ------------------------------------------------------------------------------------------
struct DenseWs {
    size_t RX, HS, Z2, R2, HC2, ZH2, GH, UH, ZERO, total;
};
static DenseWs dense_ws(int T, int N, int B, int H) {
    DenseWs w;
    const size_t U = (size_t)N * B * H;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o = align64(o + n); return r; };
    w.RX = take((size_t)T * 3 * U);
    w.HS = take((size_t)(T + 1) * U);
    w.Z2 = take((size_t)T * U);
    w.R2 = take((size_t)T * U);
    w.HC2 = take((size_t)T * U);
    w.ZH2 = take((size_t)T * U);
    w.GH = take((size_t)2 * H * H);
    w.UH = take((size_t)H * H);
    w.ZERO = take(64);
    w.total = o;
    return w;
}
extern "C" size_t matgcn_dense_gru_layer_fwd_ws_bytes(int T, int N, int B, int Cin, int H) {
    (void)Cin;
    return dense_ws(T, N, B, H).total * sizeof(float);
}
extern "C" size_t matgcn_dense_gru_layer_bwd_ws_bytes(int T, int N, int B, int Cin, int H) {
    (void)T; (void)Cin;
    return (align64((size_t)N * B * H) * 3 + 64) * sizeof(float);
}
extern "C" size_t matgcn_dense_gru_layer_y_offset(int T, int N, int B, int Cin, int H) {
    (void)Cin;
    return dense_ws(T, N, B, H).HS + (size_t)N * B * H;
}

extern "C" int matgcn_dense_gru_layer_fwd(int T, int N, int B, int Cin, int H, const float* x, long long x_tstride,
                                          const float* h0, const float* Gw, const float* Gb, const float* Uw, const float* Ub,
                                          float* ws, int flags, void* stream) {
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    REQUIRE(x && Gw && Gb && Uw && Ub && ws, "null pointer");
    REQUIRE(T > 0 && N > 0 && B > 0 && Cin > 0 && H > 0, "bad dims");
    REQUIRE((long long)N * B * 3 * H < 2147483647LL, "N*B*3H overflows int (shard the batch)");
    cudaStream_t st = (cudaStream_t)stream;
    const DenseWs w = dense_ws(T, N, B, H);
    const int I = Cin + H, NB = N * B;
    const long long U = (long long)NB * H;
    float* RX = ws + w.RX; float* HS = ws + w.HS; float* GH = ws + w.GH; float* UH = ws + w.UH;
    CK(cudaMemsetAsync(ws + w.ZERO, 0, sizeof(float) * 64, st));
    CK(cudaMemcpy2DAsync(GH, sizeof(float) * H, Gw + Cin, sizeof(float) * I, sizeof(float) * H, 2 * H, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpy2DAsync(UH, sizeof(float) * H, Uw + Cin, sizeof(float) * I, sizeof(float) * H, H, cudaMemcpyDeviceToDevice, st));
    if (h0) CK(cudaMemcpyAsync(HS, h0, sizeof(float) * U, cudaMemcpyDeviceToDevice, st));
    else CK(cudaMemsetAsync(HS, 0, sizeof(float) * U, st));
    GemmP p;
    // input half for all t: RX[t] = x_t * [Gw[:, :Cin]; Uw[:, :Cin]]^T + [Gb; Ub]
    memset(&p, 0, sizeof(p));
    p.splits = 1; p.Z2 = 1; p.KB = 1;
    p.A = x; p.lda = Cin; p.sA1 = x_tstride; p.M = NB; p.K = Cin;
    {
        p.B = Gw; p.ldb = I; p.N = 2 * H;
        EpiStore e = epi_store(RX, 3 * U, 0, 3 * H);
        e.bias = Gb;
        CK((gemm_any<CfgMid, true, true>(tc, p, e, T, st)));
        p.B = Uw; p.N = H;
        e = epi_store(RX + 2 * H, 3 * U, 0, 3 * H);
        e.bias = Ub;
        CK((gemm_any<CfgMid, true, true>(tc, p, e, T, st)));
    }
    for (int t = 0; t < T; ++t) {
        const float* Ht = HS + t * U;
        const float* RXt = RX + (long long)t * 3 * U;
        float* Z2t = ws + w.Z2 + t * U; float* R2t = ws + w.R2 + t * U; float* HC2t = ws + w.HC2 + t * U; float* ZH2t = ws + w.ZH2 + t * U;
        memset(&p, 0, sizeof(p));
        p.splits = 1; p.Z2 = 1; p.KB = 1;
        p.A = Ht; p.lda = H; p.M = NB; p.K = H;
        p.B = GH; p.ldb = H; p.N = 2 * H;
        CK((gemm_any<CfgMid, true, true>(tc, p, EpiGate{RXt, Ht, Z2t, R2t, ZH2t, 0, H, tc ? 1 : 0}, 1, st)));
        p.A = ZH2t; p.B = UH; p.N = H;
        CK((gemm_any<CfgMid, true, true>(tc, p, EpiResCand{RXt, Ht, R2t, HC2t, HS + (t + 1) * U, ws + w.ZERO, H, tc ? 1 : 0}, 1, st)));
    }
    return 0;
}

extern "C" int matgcn_dense_gru_layer_bwd(int T, int N, int B, int Cin, int H, const float* dy, long long dy_tstride,
                                          const float* x, long long x_tstride, const float* Gw, const float* Uw, float* ws,
                                          float* bws, float* dx, float* dh0, float* dGw, float* dGb, float* dUw, float* dUb,
                                          int flags, void* stream) {
    const bool tc = (flags & MATGCN_FLAG_TF32) != 0;
    REQUIRE(dy && x && Gw && Uw && ws && bws && dx && dGw && dGb && dUw && dUb, "null pointer");
    REQUIRE(T > 0 && N > 0 && B > 0 && Cin > 0 && H > 0, "bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    const DenseWs w = dense_ws(T, N, B, H);
    const int I = Cin + H, NB = N * B;
    const long long U = (long long)NB * H, UX = (long long)NB * Cin;
    float* DR = ws + w.RX; const float* HS = ws + w.HS;
    const size_t Ua = align64((size_t)U);
    float* DH1 = bws; float* DRES = bws + Ua; float* DHC = bws + 2 * Ua; float* scratch = bws + 3 * Ua;
    CK(cudaMemsetAsync(DHC, 0, sizeof(float) * U, st));
    CK(cudaMemsetAsync(scratch, 0, sizeof(float) * 64, st));
    GemmP p;
    for (int t = T - 1; t >= 0; --t) {
        const float* Ht = HS + t * U;
        const float* Z2t = ws + w.Z2 + t * U; const float* R2t = ws + w.R2 + t * U; const float* HC2t = ws + w.HC2 + t * U;
        float* DRt = DR + (long long)t * 3 * U;
        bwd_head_kernel<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(dy + (long long)t * dy_tstride, DHC, Ht, R2t, HC2t, ws + w.ZERO, U, H,
                                                                    DH1, DRES, DRt, scratch);
        count_launch();
        CK(cudaGetLastError());
        memset(&p, 0, sizeof(p));
        p.splits = 1; p.Z2 = 1; p.KB = 1;
        p.A = DRt + 2 * H; p.lda = 3 * H; p.M = NB; p.K = H;
        p.B = ws + w.UH; p.ldb = H; p.N = H;
        CK((gemm_any<CfgMid, true, false>(tc, p, EpiB1{DH1, DRt, DRES, Ht, Z2t, R2t, HC2t, H}, 1, st)));
        p.A = DRt; p.K = 2 * H; p.B = ws + w.GH;
        {
            EpiStore e = epi_store(DHC, 0, 0, H);   // carry = dh1 (direct) + da2 * Gw[:, Cin:]
            e.add = DH1; e.add_ld = H;
            CK((gemm_any<CfgMid, true, false>(tc, p, e, 1, st)));
        }
    }
    if (dh0) CK(cudaMemcpyAsync(dh0, DHC, sizeof(float) * U, cudaMemcpyDeviceToDevice, st));
    // parameter gradients over all steps
    CK(cudaMemsetAsync(dGw, 0, sizeof(float) * (size_t)2 * H * I, st));
    CK(cudaMemsetAsync(dUw, 0, sizeof(float) * (size_t)H * I, st));
    CK(cudaMemsetAsync(dGb, 0, sizeof(float) * 2 * H, st));
    CK(cudaMemsetAsync(dUb, 0, sizeof(float) * H, st));
    {
        REQUIRE(3 * H <= 1024, "hidden size too large for the column-sum kernels");
        const int cs_threads = 3 * H <= 256 ? 256 : 3 * H;
        dim3 g2(1, 1184);
        colsum_kernel<<<g2, cs_threads, 0, st>>>(DR, T, 3 * U, 0, NB, 3 * H, 3 * H, 2 * H, dGb, 2 * H, dUb, H);
        count_launch();
        CK(cudaGetLastError());
        memset(&p, 0, sizeof(p));
        p.Z2 = 1; p.KB = T; p.K = NB; p.lda = 3 * H; p.sAk = 3 * U;
        int splits = (int)(((long long)T * NB + 4095) / 4096);
        if (splits > 592) splits = 592;
        if (splits < 1) splits = 1;
        p.splits = splits;
        p.A = DR; p.M = 2 * H; p.B = HS; p.ldb = H; p.sBk = U; p.N = H;
        CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dGw + Cin, 0, 0, I}, 1, st)));
        p.A = DR + 2 * H; p.M = H; p.B = ws + w.ZH2;
        CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dUw + Cin, 0, 0, I}, 1, st)));
        p.B = x; p.ldb = Cin; p.sBk = x_tstride; p.N = Cin;
        CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dUw, 0, 0, I}, 1, st)));
        p.A = DR; p.M = 2 * H;
        CK((gemm_any<CfgMid, false, false>(tc, p, EpiAtomic{dGw, 0, 0, I}, 1, st)));
    }
    // dx[t] = DR[t][:, :2H] * Gw[:, :Cin] + DR[t][:, 2H:] * Uw[:, :Cin]
    memset(&p, 0, sizeof(p));
    p.splits = 1; p.Z2 = 1; p.KB = 1;
    p.A = DR; p.lda = 3 * H; p.sA1 = 3 * U; p.M = NB; p.K = 2 * H;
    p.B = Gw; p.ldb = I; p.N = Cin;
    {
        EpiStore e = epi_store(dx, UX, 0, Cin);
        CK((gemm_any<CfgMid, true, false>(tc, p, e, T, st)));
        e.accumulate = 1;
        p.A = DR + 2 * H; p.K = H; p.B = Uw;
        CK((gemm_any<CfgMid, true, false>(tc, p, e, T, st)));
    }
    return 0;
}
