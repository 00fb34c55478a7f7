// Error plumbing, the host-side graph build (collation, H1), type-matched program features (H2),
// Gumbel-softmax straight-through (H7) and the generic segment primitives.
#include <algorithm>
#include <vector>

#include "bg_common.cuh"

#include <stdlib.h>

namespace bg {

// Programmatic dependent launch: default ON for single-stream use (BG_PDL=0 turns it off): on the batch-32 training step the
// 2-5 us idle gaps between dependent kernels (profiles/tools/timeline.py: 7.5 -> 0.9 ms per two steps) are worth 47.5 -> 53.9 steps/s.
// It does NOT mix with the multi-stream step (step.Lanes): launches carrying the attribute into streams that fork/join through
// events cost 26.3 vs 15.5 ms per step there, so the overlapped step switches it off for its own launches through bg_set_pdl(0) and restores it (step.lanes_pdl_scope).
static int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) g_pdl = !(getenv("BG_PDL") && atoi(getenv("BG_PDL")) == 0);
    return g_pdl != 0;
}

// Device-resident addend for every in-kernel Philox offset (NULL = none).  A CUDA graph freezes kernel arguments, so a
// captured pass would replay the same dropout masks / Gumbel noise; with a base pointer set during capture the replays read
// the addend from device memory, which the owner of the graph bumps between replays (graphs.py).
static const uint64_t* g_rng_base = nullptr;
const uint64_t* rng_base() { return g_rng_base; }

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return BG_ECUDA;
    }
    return BG_OK;
}

// ---------------------------------------------------------------------------------------------
// H2: type table.  One warp per (type, feature): lane-strided partial sums in a fixed order, then a
// fixed butterfly.  M is small (tens of program nodes per building).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) type_table_kernel(const float* __restrict__ lx, const int64_t* __restrict__ ltype,
                                                              int64_t M, int F, int K, float* __restrict__ table) {
    pdl_prologue();
    const int w = (blockIdx.x * kThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= K * F) return;
    const int t = w / F, f = w % F;
    float sum = 0.f, cnt = 0.f;
    for (int64_t r = lane; r < M; r += 32) {
        if (ltype[r] == (int64_t)t) {
            sum += lx[r * F + f];
            cnt += 1.f;
        }
    }
    sum = group_sum<32>(sum);
    cnt = group_sum<32>(cnt);
    if (lane == 0) table[t * F + f] = cnt > 0.f ? sum / cnt : 0.f;
}

// Backward of the type gather: per-CTA partial [K][C] over a row chunk, last CTA folds.
__global__ void __launch_bounds__(kThreads) type_scatter_kernel(const float* __restrict__ g, int64_t ld, const int32_t* __restrict__ type,
                                                                int64_t N, int C, int K, int G, float* out,
                                                                unsigned int* counter, float* partials) {
    pdl_prologue();
    extern __shared__ float acc[];  // [K*C]
    const int64_t chunk = ceil_div(N, G);
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(N, r0 + chunk);
    // thread c owns column c (C <= kThreads): rows are visited in order => fixed summation order
    for (int i = threadIdx.x; i < K * C; i += kThreads) acc[i] = 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads)
        for (int64_t r = r0; r < r1; ++r) acc[__ldg(type + r) * C + c] += __ldg(g + r * ld + c);
    __syncthreads();
    for (int i = threadIdx.x; i < K * C; i += kThreads) partials[(int64_t)blockIdx.x * K * C + i] = acc[i];
    __shared__ float red[kThreads];
    __syncthreads();
    if (!hier_fold(partials, partials + (int64_t)G * K * C, K * C, counter, red, acc)) return;
    for (int i = threadIdx.x; i < K * C; i += kThreads) out[i] = acc[i];
}

// The same for K <= 8 types (the models' 7) and C a multiple of 32: per-type sums in REGISTERS, grid = row chunks x 32-column
// groups.  The kernel above walks its rows one at a time through a shared-memory read-modify-write (a dependent chain with the
// global load inside) and ends in a last-CTA fold of ~120 partial vectors of K*C = 896 floats by ONE CTA: 37-49 us at N = 15 k
// on the generator's backward chain (profiles/r02b_summary.md).  Here warp `slot` of a CTA owns every 8th row of the chunk and
// lane l one column: 8 accumulators per thread, eight independent row loads in flight; the 8 slots fold through shared
// memory, the <= 64 row chunks of a column group through one ticket round by the group's last CTA (224 floats x <= 64 partials,
// 32 loads in flight per thread).  Fixed order everywhere: bitwise reproducible.
constexpr int kTsSlots = kThreads / 32;
static inline int ts_row_chunks(int64_t N) {
    int64_t rc = ceil_div(N, 256);
    return (int)(rc < 1 ? 1 : rc > 64 ? 64 : rc);
}
__global__ void __launch_bounds__(kThreads) type_scatter8_kernel(const float* __restrict__ g, int64_t ld, const int32_t* __restrict__ type,
                                                                 int64_t N, int C, int K, float* out, unsigned int* counters,
                                                                 float* partials) {
    pdl_prologue();
    __shared__ float sacc[kTsSlots][8 * 32];
    const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
    const int RC = gridDim.x, rc = blockIdx.x, cg = blockIdx.y, c = cg * 32 + lane;
    const int64_t chunk = ceil_div(N, RC);
    const int64_t r0 = (int64_t)rc * chunk, r1 = min(N, r0 + chunk);
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0.f;
    int64_t r = r0 + slot;
    for (; r + 7 * kTsSlots < r1; r += 8 * kTsSlots) {
        int t[8];
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            t[q] = __ldg(type + r + q * kTsSlots);
            v[q] = __ldg(g + (r + q * kTsSlots) * ld + c);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += t[q] == k ? v[q] : 0.f;
    }
    for (; r < r1; r += kTsSlots) {
        const int t = __ldg(type + r);
        const float v = __ldg(g + r * ld + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += t == k ? v : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sacc[slot][k * 32 + lane] = a[k];
    __syncthreads();
    const int L = K * 32;
    float* mine = partials + ((int64_t)cg * RC + rc) * L;
    if (threadIdx.x < L) {
        float t = 0.f;
#pragma unroll
        for (int sl = 0; sl < kTsSlots; ++sl) t += sacc[sl][threadIdx.x];
        mine[threadIdx.x] = t;
    }
    if (!ticket_last(counters + 1 + cg, RC)) return;
    if (threadIdx.x < L) {
        const float* col = partials + (int64_t)cg * RC * L + threadIdx.x;
        float t = 0.f;
        for (int g0 = 0; g0 < RC; g0 += 32) {
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] = g0 + u < RC ? col[(int64_t)(g0 + u) * L] : 0.f;
#pragma unroll
            for (int u = 0; u < 32; ++u) t += v[u];
        }
        out[(threadIdx.x >> 5) * C + c] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// H7: Gumbel-softmax + straight-through, one thread per row (K <= 16)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gumbel_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                                              uint64_t seed, uint64_t offset, const uint64_t* __restrict__ base,
                                                              int64_t N, int K, float* __restrict__ soft, float* __restrict__ hard,
                                                              int32_t* __restrict__ amax) {
    pdl_prologue();
    if (base) offset += *base;
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= N) return;
    float v[16], mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
        float g;
        if (noise) g = noise[r * K + k];
        else {  // Gumbel(0,1) = -log(Exp(1)), Exp(1) = -log(U): what F.gumbel_softmax draws, from Philox
            if ((k & 3) == 0) {
                const uint4 rnd = philox4x32((uint64_t)r * 4 + (k >> 2), offset, seed);
                v[12] = u01(rnd.x); v[13] = u01(rnd.y); v[14] = u01(rnd.z); v[15] = u01(rnd.w);
            }
            g = -logf(-logf(v[12 + (k & 3)]) + 1e-20f);
        }
        v[k] = (logits[r * K + k] + g) / 1.0f;
        mx = fmaxf(mx, v[k]);
    }
    float sum = 0.f;
    for (int k = 0; k < K; ++k) {
        v[k] = expf(v[k] - mx);
        sum += v[k];
    }
    int best = 0;
    float bv = -INFINITY;
    for (int k = 0; k < K; ++k) {
        v[k] = v[k] / sum;
        if (v[k] > bv) {  // first maximum wins, as torch.argmax does
            bv = v[k];
            best = k;
        }
    }
    for (int k = 0; k < K; ++k) {
        soft[r * K + k] = v[k];
        const float one = (k == best) ? 1.f : 0.f;
        hard[r * K + k] = (one - v[k]) + v[k];  // label_hard - label_soft.detach() + label_soft
    }
    if (amax) amax[r] = best;
}

__global__ void __launch_bounds__(kThreads) gumbel_bwd_kernel(const float* __restrict__ g_hard, const float* __restrict__ g_soft,
                                                              const float* __restrict__ soft, int64_t N, int K,
                                                              float* __restrict__ g_logits) {
    pdl_prologue();
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= N) return;
    float g[16], dot = 0.f;
    for (int k = 0; k < K; ++k) {
        g[k] = (g_hard ? g_hard[r * K + k] : 0.f) + (g_soft ? g_soft[r * K + k] : 0.f);
        dot = fmaf(g[k], soft[r * K + k], dot);
    }
    for (int k = 0; k < K; ++k) g_logits[r * K + k] = soft[r * K + k] * (g[k] - dot);
}

// ---------------------------------------------------------------------------------------------
// generic segment primitives (one warp per segment)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) segment_softmax_kernel(const float* __restrict__ v, const int32_t* __restrict__ ptr,
                                                                   int64_t S, float* __restrict__ out) {
    pdl_prologue();
    const int64_t sgm = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (sgm >= S) return;
    const int b = ptr[sgm], e = ptr[sgm + 1];
    float mx = -INFINITY;
    for (int i = b + lane; i < e; i += 32) mx = fmaxf(mx, v[i]);
    mx = group_max<32>(mx);
    float sum = 0.f;
    for (int i = b + lane; i < e; i += 32) sum += expf(v[i] - mx);
    sum = group_sum<32>(sum) + 1e-16f;
    for (int i = b + lane; i < e; i += 32) out[i] = expf(v[i] - mx) / sum;
}

__global__ void __launch_bounds__(kThreads) segment_pool_kernel(const float* __restrict__ x, const int32_t* __restrict__ ptr,
                                                                int64_t S, int C, int mode, float* __restrict__ out) {
    pdl_prologue();
    // one CTA per segment; thread c owns column c and walks the rows in order
    const int64_t sgm = blockIdx.x;
    if (sgm >= S) return;
    const int b = ptr[sgm], e = ptr[sgm + 1];
    for (int c = threadIdx.x; c < C; c += kThreads) {
        float a = mode == 1 ? -INFINITY : 0.f;
        for (int r = b; r < e; ++r) {
            const float t = x[(int64_t)r * C + c];
            a = mode == 1 ? fmaxf(a, t) : a + t;
        }
        if (mode == 0) a = a / (float)max(1, e - b);
        if (mode == 1 && e == b) a = 0.f;
        out[sgm * C + c] = a;
    }
}


// ---------------------------------------------------------------------------------------------
// H11 / H12 loss glue of one critic update (reference trainer.py:298-301, 314, 323): the interpolate between the real one-hot
// labels and the generated soft labels, and  loss = mean(D(fake)) - mean(D(real)) + lambda * mean((||grad_i||_2 - 1)^2)
// with its backward.  torch runs this as ~30 elementwise / reduction launches per critic update, all of them on the
// critical path between the first and the second-order backward of the gradient-penalty pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gp_mix_kernel(const float* __restrict__ e, const int64_t* __restrict__ onehot_i64,
                                                          const float* __restrict__ onehot_f32, const float* __restrict__ soft,
                                                          int64_t N, int K, float* __restrict__ mixed) {
    pdl_prologue();
    const int64_t total = N * K;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const float ev = __ldg(e + i / K);
        const float oh = onehot_i64 ? (float)__ldg(onehot_i64 + i) : __ldg(onehot_f32 + i);
        mixed[i] = ev * oh + (1.f - ev) * __ldg(soft + i);  // e * real + (1 - e) * fake, in torch's evaluation order
    }
}

constexpr int kLossTerms = 3;  // sum D(fake), sum D(real), sum (||grad_i|| - 1)^2
__global__ void __launch_bounds__(kThreads) critic_loss_fwd_kernel(const float* __restrict__ d_fake, const float* __restrict__ d_real,
                                                                   const float* __restrict__ grad, int64_t N, int K, float lambda,
                                                                   float* __restrict__ coef, float* partials, unsigned int* counters,
                                                                   float* __restrict__ out /* loss, mean fake, mean real, gp */) {
    pdl_prologue();
    __shared__ float red[kThreads];
    __shared__ float wsum[kLossTerms][kThreads / 32];
    __shared__ float result[kLossTerms];
    const int G = gridDim.x;
    const int64_t chunk = ceil_div(N, (int64_t)G);
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(N, r0 + chunk);
    float acc[kLossTerms] = {0.f, 0.f, 0.f};
    for (int64_t i = r0 + threadIdx.x; i < r1; i += kThreads) {
        acc[0] += __ldg(d_fake + i);
        acc[1] += __ldg(d_real + i);
        float ss = 0.f;
        for (int k = 0; k < K; ++k) {
            const float g = __ldg(grad + i * K + k);
            ss = fmaf(g, g, ss);
        }
        const float nrm = sqrtf(ss), dev = nrm - 1.f;
        acc[2] = fmaf(dev, dev, acc[2]);
        // d/d grad_i of lambda * mean((||grad_i|| - 1)^2) = coef_i * grad_i; torch's norm backward is 0 at a zero row
        coef[i] = nrm > 0.f ? lambda * 2.f * dev / ((float)N * nrm) : 0.f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < kLossTerms; ++t) {
        float v = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wsum[t][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLossTerms) {
        float v = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) v += wsum[threadIdx.x][w];
        partials[(int64_t)blockIdx.x * kLossTerms + threadIdx.x] = v;
    }
    if (!hier_fold(partials, partials + (int64_t)G * kLossTerms, kLossTerms, counters, red, result)) return;
    if (threadIdx.x == 0) {
        const float mf = result[0] / (float)N, mr = result[1] / (float)N, gp = lambda * (result[2] / (float)N);
        out[0] = mf - mr + gp;
        out[1] = mf;
        out[2] = mr;
        out[3] = gp;
    }
}

// upstream gradient g (a device scalar): d loss / d D(fake)_i = g / N, d loss / d D(real)_i = -g / N, d loss / d grad_i = g coef_i grad_i
__global__ void __launch_bounds__(kThreads) critic_loss_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ coef,
                                                                   const float* __restrict__ grad, int64_t N, int K,
                                                                   float* __restrict__ g_fake, float* __restrict__ g_real,
                                                                   float* __restrict__ g_grad) {
    pdl_prologue();
    const float g = __ldg(g_loss);
    const float inv = g / (float)N;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        if (g_fake) g_fake[i] = inv;
        if (g_real) g_real[i] = -inv;
        if (g_grad) {
            const float c = g * __ldg(coef + i);
            for (int k = 0; k < K; ++k) g_grad[i * K + k] = c * __ldg(grad + i * K + k);
        }
    }
}
}  // namespace bg

using namespace bg;

extern "C" int bg_version(void) { return 101; }
extern "C" int bg_set_pdl(int32_t on) {
    const int prev = bg::pdl_enabled() ? 1 : 0;
    bg::g_pdl = on ? 1 : 0;
    return prev;
}
extern "C" int bg_set_rng_base(const uint64_t* base) {
    bg::g_rng_base = base;
    return BG_OK;
}
extern "C" const char* bg_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// H1: host CSR/CSC build.  Stable counting sort by destination keeps, per row, the COO order of the
// input (the reference's CPU scatter order); the self loop goes last.  No CUDA calls.
// ---------------------------------------------------------------------------------------------
extern "C" int bg_csr_build_host(const int64_t* coo, int64_t E_in, int64_t N, int32_t* rowptr, int32_t* col,
                                 int32_t* cscptr, int32_t* cscrow, int32_t* perm, int64_t* E_out, int32_t* max_deg) {
    BG_REQUIRE(rowptr && col && cscptr && cscrow && perm && E_out && max_deg, BG_EINVAL, "bg_csr_build_host: null output");
    BG_REQUIRE(N > 0 && E_in >= 0 && (coo || E_in == 0), BG_EINVAL, "bg_csr_build_host: bad sizes N=%lld E=%lld",
               (long long)N, (long long)E_in);
    BG_REQUIRE(N < (int64_t)1 << 31 && E_in + N < (int64_t)1 << 31, BG_ERANGE, "bg_csr_build_host: graph exceeds int32 indexing");
    const int64_t* src = coo;
    const int64_t* dst = coo + E_in;
    std::vector<int32_t> indeg(N + 1, 0), outdeg(N + 1, 0);
    for (int64_t e = 0; e < E_in; ++e) {
        const int64_t s = src[e], t = dst[e];
        if (s < 0 || s >= N || t < 0 || t >= N) {
            set_error("bg_csr_build_host: edge %lld = (%lld -> %lld) out of range [0,%lld)", (long long)e, (long long)s,
                      (long long)t, (long long)N);
            return BG_ERANGE;
        }
        if (s == t) continue;  // remove_self_loops
        ++indeg[t];
        ++outdeg[s];
    }
    rowptr[0] = 0;
    cscptr[0] = 0;
    int32_t md = 0;
    for (int64_t i = 0; i < N; ++i) {
        rowptr[i + 1] = rowptr[i] + indeg[i] + 1;  // +1: add_self_loops
        cscptr[i + 1] = cscptr[i] + outdeg[i] + 1;
        md = std::max(md, indeg[i] + 1);
    }
    std::vector<int32_t> rfill(rowptr, rowptr + N), cfill(cscptr, cscptr + N);
    // CSR fill in COO order
    for (int64_t e = 0; e < E_in; ++e) {
        const int64_t s = src[e], t = dst[e];
        if (s == t) continue;
        col[rfill[t]++] = (int32_t)s;
    }
    for (int64_t i = 0; i < N; ++i) col[rfill[i]++] = (int32_t)i;  // self loop last
    // CSC: walk the CSR in order so every source's out-edges are listed by ascending destination
    for (int64_t i = 0; i < N; ++i)
        for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int32_t s = col[e];
            const int32_t k = cfill[s]++;
            cscrow[k] = (int32_t)i;
            perm[k] = e;
        }
    *E_out = rowptr[N];
    *max_deg = md;
    return BG_OK;
}

extern "C" int bg_type_table(const float* local_x, const int64_t* local_type, int64_t M, int32_t F, int32_t K, float* table,
                             void* stream) {
    BG_REQUIRE(local_x && local_type && table, BG_EINVAL, "bg_type_table: null pointer");
    BG_REQUIRE(M >= 0 && F > 0 && K > 0, BG_EINVAL, "bg_type_table: bad sizes");
    const int warps = K * F;
    launch_k(type_table_kernel, (unsigned)ceil_div(warps * 32, kThreads), kThreads, 0, as_stream(stream), local_x, local_type, M, F, K, table);
    return check_launch("bg_type_table");
}

static inline int scatter_splits(int64_t N) {
    int64_t g = ceil_div(N, 128);
    if (g > kSMs) g = kSMs;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" size_t bg_type_scatter_sum_ws(int64_t N, int32_t C, int32_t K) {
    return (size_t)kCounterBytes + (size_t)(scatter_splits(N) + scatter_splits(N) / kFoldGroup + 2) * (size_t)K * (size_t)C * sizeof(float);
}

extern "C" int bg_type_scatter_sum(const float* g, int64_t ld, const int32_t* type, int64_t N, int32_t C, int32_t K, float* out,
                                   float* workspace, size_t ws_bytes, void* stream) {
    BG_REQUIRE(g && type && out && workspace, BG_EINVAL, "bg_type_scatter_sum: null pointer");
    BG_REQUIRE(ws_bytes >= bg_type_scatter_sum_ws(N, C, K), BG_EINVAL, "bg_type_scatter_sum: workspace too small");
    BG_REQUIRE((size_t)K * C * sizeof(float) <= 48 * 1024, BG_EUNSUPPORTED, "bg_type_scatter_sum: K*C too large");
    const int G = scatter_splits(N);
    if (K <= 8 && (C & 31) == 0) {
        const int RC = ts_row_chunks(N);  // partials: RC x K x C floats <= the (scatter_splits(N) + 2) x K x C the caller provides
        launch_k(type_scatter8_kernel, dim3((unsigned)RC, (unsigned)(C / 32)), kThreads, 0, as_stream(stream), g, ld, type, N, C, K, out,
                 reinterpret_cast<unsigned int*>(workspace), workspace + kCounterBytes / sizeof(float));
    }
    else
        launch_k(type_scatter_kernel, G, kThreads, (size_t)K * C * sizeof(float), as_stream(stream),
            g, ld, type, N, C, K, G, out, reinterpret_cast<unsigned int*>(workspace), workspace + kCounterBytes / sizeof(float));
    return check_launch("bg_type_scatter_sum");
}

extern "C" int bg_gumbel_st_fwd(const float* logits, const float* noise, uint64_t seed, uint64_t offset, int64_t N, int32_t K,
                                float* soft, float* hard, int32_t* argmax, void* stream) {
    BG_REQUIRE(logits && soft && hard, BG_EINVAL, "bg_gumbel_st_fwd: null pointer");
    BG_REQUIRE(K >= 1 && K <= 12, BG_EUNSUPPORTED, "bg_gumbel_st_fwd: K=%d not in [1,12]", K);
    launch_k(gumbel_fwd_kernel, (unsigned)ceil_div(N, kThreads), kThreads, 0, as_stream(stream), logits, noise, seed, offset, rng_base(), N, K, soft,
                                                                                          hard, argmax);
    return check_launch("bg_gumbel_st_fwd");
}

extern "C" int bg_gumbel_st_bwd(const float* g_hard, const float* g_soft, const float* soft, int64_t N, int32_t K,
                                float* g_logits, void* stream) {
    BG_REQUIRE(soft && g_logits && (g_hard || g_soft), BG_EINVAL, "bg_gumbel_st_bwd: null pointer");
    BG_REQUIRE(K >= 1 && K <= 16, BG_EUNSUPPORTED, "bg_gumbel_st_bwd: K=%d not in [1,16]", K);
    launch_k(gumbel_bwd_kernel, (unsigned)ceil_div(N, kThreads), kThreads, 0, as_stream(stream), g_hard, g_soft, soft, N, K, g_logits);
    return check_launch("bg_gumbel_st_bwd");
}

extern "C" int bg_segment_softmax(const float* v, const int32_t* seg_ptr, int64_t S, float* out, void* stream) {
    BG_REQUIRE(v && seg_ptr && out, BG_EINVAL, "bg_segment_softmax: null pointer");
    if (S <= 0) return BG_OK;
    launch_k(segment_softmax_kernel, (unsigned)ceil_div(S * 32, kThreads), kThreads, 0, as_stream(stream), v, seg_ptr, S, out);
    return check_launch("bg_segment_softmax");
}

// Per-segment confusion matrices: cm[s, t, p] = #{rows of segment s with target t and prediction p}, prediction =
// argmax of a score row (first maximum, like torch.argmax).  One CTA per segment, integer shared-memory counters
// (integer addition is associative: deterministic).  Replaces the B+4 sklearn calls per step of trainer.py:387-443.
__global__ void __launch_bounds__(kThreads) segment_confusion_kernel(const float* __restrict__ score,
                                                                     const int64_t* __restrict__ target,
                                                                     const int32_t* __restrict__ seg_ptr, int K,
                                                                     int32_t* __restrict__ cm) {
    pdl_prologue();
    __shared__ int cnt[16 * 16];
    for (int i = threadIdx.x; i < K * K; i += kThreads) cnt[i] = 0;
    __syncthreads();
    const int s = blockIdx.x;
    for (int r = seg_ptr[s] + threadIdx.x; r < seg_ptr[s + 1]; r += kThreads) {
        const float* row = score + (int64_t)r * K;
        int best = 0;
        float bv = __ldg(row);
        for (int k = 1; k < K; ++k) {
            const float v = __ldg(row + k);
            if (v > bv) {
                bv = v;
                best = k;
            }
        }
        const int t = (int)target[r];
        if (t >= 0 && t < K) atomicAdd(&cnt[t * K + best], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += kThreads) cm[(int64_t)s * K * K + i] = cnt[i];
}

extern "C" int bg_segment_confusion(const float* score, const int64_t* target, const int32_t* seg_ptr, int64_t S, int32_t K,
                                    int32_t* cm, void* stream) {
    BG_REQUIRE(score && target && seg_ptr && cm, BG_EINVAL, "bg_segment_confusion: null pointer");
    BG_REQUIRE(K >= 1 && K <= 16, BG_EUNSUPPORTED, "bg_segment_confusion: K=%d out of range [1,16]", (int)K);
    if (S <= 0) return BG_OK;
    launch_k(segment_confusion_kernel, (unsigned)S, kThreads, 0, as_stream(stream), score, target, seg_ptr, K, cm);
    return check_launch("bg_segment_confusion");
}

extern "C" int bg_segment_pool(const float* x, const int32_t* seg_ptr, int64_t S, int32_t C, int32_t mode, float* out,
                               void* stream) {
    BG_REQUIRE(x && seg_ptr && out, BG_EINVAL, "bg_segment_pool: null pointer");
    BG_REQUIRE(mode >= 0 && mode <= 2, BG_EINVAL, "bg_segment_pool: mode must be 0 (mean), 1 (max) or 2 (sum)");
    if (S <= 0) return BG_OK;
    launch_k(segment_pool_kernel, (unsigned)S, kThreads, 0, as_stream(stream), x, seg_ptr, S, C, mode, out);
    return check_launch("bg_segment_pool");
}

extern "C" int bg_gp_mix(const float* e, const void* onehot, int32_t onehot_is_i64, const float* soft, int64_t N, int32_t K, float* mixed,
                         void* stream) {
    BG_REQUIRE(e && onehot && soft && mixed, BG_EINVAL, "bg_gp_mix: null pointer");
    if (N <= 0) return BG_OK;
    const int64_t grid = std::min<int64_t>(bg::ceil_div(N * K, (int64_t)bg::kThreads), 8 * bg::kSMs);
    bg::launch_k(bg::gp_mix_kernel, (int)grid, bg::kThreads, 0, bg::as_stream(stream), e,
                 onehot_is_i64 ? static_cast<const int64_t*>(onehot) : nullptr, onehot_is_i64 ? nullptr : static_cast<const float*>(onehot),
                 soft, N, (int)K, mixed);
    return bg::check_launch("bg_gp_mix");
}

extern "C" size_t bg_critic_loss_ws(int64_t N) { return bg::kCounterBytes + (size_t)(64 + 8) * bg::kLossTerms * sizeof(float) + 256; }

extern "C" int bg_critic_loss_fwd(const float* d_fake, const float* d_real, const float* grad, int64_t N, int32_t K, float lambda,
                                  float* coef, float* workspace, size_t ws_bytes, float* out4, void* stream) {
    BG_REQUIRE(d_fake && d_real && grad && coef && workspace && out4, BG_EINVAL, "bg_critic_loss_fwd: null pointer");
    BG_REQUIRE(N >= 1 && K >= 1, BG_EINVAL, "bg_critic_loss_fwd: N=%lld K=%d", (long long)N, K);
    BG_REQUIRE(ws_bytes >= bg_critic_loss_ws(N), BG_EINVAL, "bg_critic_loss_fwd: workspace too small");
    unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
    float* partials = workspace + bg::kCounterBytes / sizeof(float);
    const int G = (int)std::min<int64_t>(64, bg::ceil_div(N, (int64_t)bg::kThreads));  // function of N only: reproducible
    bg::launch_k(bg::critic_loss_fwd_kernel, G, bg::kThreads, 0, bg::as_stream(stream), d_fake, d_real, grad, N, (int)K, lambda, coef, partials,
                 counters, out4);
    return bg::check_launch("bg_critic_loss_fwd");
}

extern "C" int bg_critic_loss_bwd(const float* g_loss, const float* coef, const float* grad, int64_t N, int32_t K, float* g_fake,
                                  float* g_real, float* g_grad, void* stream) {
    BG_REQUIRE(g_loss && coef && grad, BG_EINVAL, "bg_critic_loss_bwd: null pointer");
    if (N <= 0) return BG_OK;
    const int64_t grid = std::min<int64_t>(bg::ceil_div(N, (int64_t)bg::kThreads), 4 * bg::kSMs);
    bg::launch_k(bg::critic_loss_bwd_kernel, (int)grid, bg::kThreads, 0, bg::as_stream(stream), g_loss, coef, grad, N, (int)K, g_fake, g_real, g_grad);
    return bg::check_launch("bg_critic_loss_bwd");
}
