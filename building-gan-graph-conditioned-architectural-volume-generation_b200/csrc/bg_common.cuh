// Shared device/host helpers for libbgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <vector>

#include "bg_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbgb200 targets sm_100a (B200) only"
#endif

namespace bg {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSMs = 148;  // B200
constexpr int kCounterBytes = 4096;  // head of every reduction workspace: self-resetting ticket counters

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// ---- deferred weight-gradient folds (bg_dense.cu): partial sums are queued and folded once per pass
struct FoldEntry {
    const float* partial;  // [S][Cout*K]
    float* dW;
    float* dbias;          // last column of K routed here when non-null
    int64_t ld_dw;
    int S, Cout, K, accumulate, block_base;
};
struct FoldBatch {
    static constexpr int kMax = 56;
    int n;
    FoldEntry e[kMax];
};
struct WgradQueue {
    static constexpr int kPhases = 4;
    std::vector<FoldEntry> phase[kPhases];
    std::vector<BgWgrad> pending;  // deferred mode: problems recorded by wgrad_launch, launched in batches of kWgMax
    bool defer = false;
    // deferred mode with overlap: batches run on a library-owned side stream, forked from / joined back into the caller's
    // stream with events (capturable), so the weight gradients overlap the latency-bound dgrad chain of the same pass
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool forked = false;
    float* buf = nullptr;
    size_t cap = 0, used = 0;  // floats
};
// bg_gat_bwd / bg_gat_bwd_gn with the second-order sweep's cotangent at h added in the source pass's epilogue (bg_gat.cu)
int gat_bwd_inj(const BgGraph* g, const float* gout, const float* h, const float* s, const float* d, const float* m, const float* z,
                const float* a_src, const float* a_dst, float* P, float* DU, float* gh_tot, float* gsd, int32_t C, float slope,
                const float* inj_h, void* stream);
int gat_bwd_gn_inj(const BgGraph* g, const float* gx1, const float* o, const float* x1, const float* gn_w, const float* gn_alpha,
                   const float* gn_stats, const float* gn_bstats, float keep_scale, const float* inj_o, const float* h, const float* s,
                   const float* d, const float* m, const float* z, const float* a_src, const float* a_dst, float* P, float* DU,
                   float* go, float* gh_tot, float* gsd, int32_t C, float slope, const float* inj_h, void* stream);
int wgrad_launch(const BgWgrad* probs, int nprob, WgradQueue& q, cudaStream_t st);
int wgrad_flush(WgradQueue& q, cudaStream_t st);

// GraphNorm backward moments of the block BELOW, fused into the epilogue of the backward-input product that produces that
// block's gx1 (bg_dense.cu: dense_fwd_moments): the tile a CTA just computed is exactly the operand gn_bwd_moments_kernel
// would read back, so the column sums (sum gy, sum gy*(o - alpha mu)) come for two extra loads per element and the separate
// 10-us launch (165 per training step, 80 of them on the step's critical chain) disappears.  Same outputs as
// bg_graphnorm_bwd_moments: bstats[2C] = (G0, G1), dparams[3C] = (dw, dbeta, dalpha) (+= when accumulate).
struct GnMomFuse {
    const float *o, *x1, *alpha, *stats, *w;
    float keep_scale;
    float* dparams;
    int accumulate;
    float* bstats;
    unsigned int* counters;  // reduction workspace head (self-resetting tickets)
    float* partials;         // >= (G + G / kFoldGroup + 2) * 2 * BN floats, G = ceil(N / 64)
};
int dense_fwd_moments(const BgDense* a, const GnMomFuse* f, cudaStream_t st);  // f == nullptr: plain bg_dense_fwd

// row-per-thread kernel for the small layers of the latency-bound regime (bg_rowdense.cu): BG_OK when launched, 1 when
// the shape is not eligible, <0 on error.  seg_off = the BG_MAX_SEG + 1 column offsets of the segmented input.
int rowdense_try(const BgDense* a, const int* seg_off, int K, const GnMomFuse* mom, cudaStream_t st);

// TMA-gather variant of the aggregation forward (bg_gat_tma.cu); same return convention
int gat_tma_enabled();
int gat_fwd_tma_try(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias, float* out, float* m,
                    float* z, int C, float slope, cudaStream_t st);

// warp-MMA 3xTF32 kernel for small single-segment layers (bg_dense_mma.cu); same return convention
int dense_mma_try(const BgDense* a, int K, const GnMomFuse* mom, cudaStream_t st);

// tcgen05 3xTF32 dense path (bg_dense_tc.cu): BG_OK when launched, 1 when the shape is not eligible, <0 on error
int dense_tc_try(const BgDense* a, int K, cudaStream_t st);

// ---- launches: every kernel of the library goes through launch_k().  With programmatic dependent launch (PDL; default ON,
// BG_PDL=0 / bg_set_pdl(0) switch it off, see pdl_enabled() in bg_misc.cu) kernel N+1 is scheduled while kernel N drains: each kernel starts with pdl_prologue() =
// `griddepcontrol.launch_dependents` (the next kernel may be made resident once all CTAs of this one have started) followed by
// `griddepcontrol.wait` (block until the previous kernel has completed and its writes are visible) BEFORE its first global
// access, so stream order is preserved exactly.  Measured on the batch-32 training step (profiles/tools/timeline.py): 4400 kernels
// per two steps left 7.5 ms of 2-5 us idle gaps between dependent launches.
bool pdl_enabled();
const uint64_t* rng_base();  // bg_set_rng_base
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define BG_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            bg::set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Number of row-chunks a column reduction over N rows is split into.  Fixed function of N only,
// so the reduction tree (and therefore the rounding) is reproducible run to run.
static inline int reduce_splits(int64_t N, int rows_per_cta_iter) {
    int64_t want = ceil_div(N, (int64_t)rows_per_cta_iter * 4);
    if (want < 1) want = 1;
    // latency-bound sizes: at most one CTA per SM, so the cross-CTA fold is a single ticket round (<= kSingleFold partials)
    const int64_t cap = N <= 65536 ? kSMs : 2 * kSMs;
    if (want > cap) want = cap;
    return (int)want;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void store(float* p) const { p[0] = v[0]; }
};
template <>
struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float* p) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

// 256-bit global load / store (sm_100: LDG.E.256 / STG.E.256 - what nvcc emits for a 32-byte aligned aggregate; p must be
// 32-byte aligned and, for load(), point to data that is read-only for the kernel's lifetime)
struct alignas(32) Float8 {
    float v[8];
};
template <>
struct Vec<8> {
    float v[8];
    __device__ __forceinline__ void load(const float* __restrict__ p) {
        const Float8 t = *reinterpret_cast<const Float8*>(p);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = t.v[q];
    }
    __device__ __forceinline__ void store(float* p) const {
        Float8 t;
#pragma unroll
        for (int q = 0; q < 8; ++q) t.v[q] = v[q];
        *reinterpret_cast<Float8*>(p) = t;
    }
};

// Butterfly sum over a group of LANES consecutive lanes (LANES power of two <= 32): every lane of
// the group ends with the same, order-fixed result.
template <int LANES>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
template <int LANES>
__device__ __forceinline__ float group_max(float x) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// One group of LANES = C/VEC consecutive lanes owns one node row (VEC = min(4, C) channels per lane).
template <int C>
struct RowMap {
    static constexpr int VEC = C >= 4 ? 4 : C;
    static constexpr int LANES = C / VEC;
    static constexpr int RPW = 32 / LANES;       // rows per warp
    static constexpr int RPC = RPW * kWarps;     // rows per CTA
};

template <int LANES>
__device__ __forceinline__ unsigned group_mask(int lane) {
    if (LANES >= 32) return 0xffffffffu;
    return ((1u << (LANES & 31)) - 1u) << ((lane / LANES) * LANES);
}
template <int LANES>
__device__ __forceinline__ float gsum(float x, unsigned mask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
    return x;
}
template <int LANES>
__device__ __forceinline__ float gmax(float x, unsigned mask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(mask, x, o));
    return x;
}

// Philox4x32-10 counter-based generator (Salmon et al., SC'11): 4 uniform 32-bit words per (counter, key).
// Used for the fused dropout masks and Gumbel noise: element i of call `offset` reads word i%4 of
// philox(counter = (i/4, offset), key = seed) - stateless, reproducible, CUDA-graph safe.
__device__ __forceinline__ uint4 philox4x32(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t seed) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// 32 random bits -> uniform in (0, 1]
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// Fire-and-forget L2 prefetch of the 128-byte line holding p: unlike a load it holds no register, so the bytes in
// flight are not bounded by the register file.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float lrelu(float u, float slope) { return u > 0.f ? u : u * slope; }
__device__ __forceinline__ float lrelu_grad(float u, float slope) { return u > 0.f ? 1.f : slope; }

// "Last CTA finalises" ticket: every CTA publishes its partial, fences, takes a ticket; the CTA
// that draws the last ticket returns true (and resets the counter so the buffer is reusable
// without a memset).  The finaliser reads the partials in a FIXED order => deterministic.
__device__ __forceinline__ bool last_cta_ticket(unsigned int* counter, unsigned int total) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == total - 1);
        if (is_last) *counter = 0u;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// "Last CTA" ticket on an arbitrary counter (self-resetting).
__device__ __forceinline__ bool ticket_last(unsigned int* counter, unsigned int total) {
    __shared__ bool is_last_t;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(counter, 1u);
        is_last_t = (t == total - 1);
        if (is_last_t) *counter = 0u;
    }
    __syncthreads();
    const bool r = is_last_t;
    if (r) __threadfence();
    __syncthreads();
    return r;
}

// Fixed-order fold of `G` partial vectors of length L (partials[g*L + i]) by the last CTA.
// Uses S = kThreads / Lp slices (Lp = L rounded up to a power of two, capped at kThreads).  Every thread sums its slice
// g = sl, sl + S, ... in that order; the loads are issued in register batches of kFoldBatch (all of a batch in flight before
// the first add): at G ~ 150 partials of 128 floats a thread has ~80 loads to make, and eight at a time (what a plain
// unrolled loop gives) made this fold 10 dependent L2 round trips = most of a 10-us moments launch.
constexpr int kFoldBatch = 32;
__device__ __forceinline__ float fold_slice(const float* __restrict__ col, int G, int sl, int S, int64_t L) {
    float t = 0.f;
    for (int g0 = sl; g0 < G; g0 += S * kFoldBatch) {
        float v[kFoldBatch];
#pragma unroll
        for (int u = 0; u < kFoldBatch; ++u) {
            const int g = g0 + u * S;
            v[u] = g < G ? col[(int64_t)g * L] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kFoldBatch; ++u) t += v[u];  // adding 0 for g >= G leaves the bits unchanged
    }
    return t;
}
__device__ __forceinline__ void fold_partials(const float* partials, int G, int L, float* red /*>= kThreads*/,
                                              float* result /*shared, >= L*/) {
    for (int base = 0; base < L; base += kThreads) {
        const int Lt = min(L - base, kThreads);
        int Lp = 1;
        while (Lp < Lt) Lp <<= 1;
        const int S = kThreads / Lp;
        const int i = threadIdx.x % Lp, sl = threadIdx.x / Lp;
        red[threadIdx.x] = i < Lt ? fold_slice(partials + base + i, G, sl, S, L) : 0.f;
        __syncthreads();
        if (threadIdx.x < Lt) {
            float a = 0.f;
            for (int q = 0; q < S; ++q) a += red[q * Lp + threadIdx.x];
            result[base + threadIdx.x] = a;
        }
        __syncthreads();
    }
}


// Two-level deterministic fold of G per-CTA partial vectors (length L): the last CTA of every group of
// kFoldGroup folds its group, the last group folds the group results.  Both folds are <= ~20 deep, so the
// critical path stays short even with ~300 CTAs.  counters[0] = global ticket, counters[1+g] = group g.
// Returns true in exactly one CTA, whose `result` (shared, >= L floats) then holds the column sums.
// Up to kSingleFold CTAs (the small, L2-resident graphs of the training step: N ~ 15 k rows -> G ~ 120) ONE ticket round is
// enough: the last CTA sums all G partials itself (<= 40 loads per thread, 8 in flight) instead of paying a second
// fence -> atomic -> reload round trip (~2 us of a ~10 us launch).  The choice depends on G only, so results stay reproducible.
constexpr int kFoldGroup = 16;
constexpr int kSingleFold = 160;
__device__ __forceinline__ bool hier_fold(float* partials, float* gpartials, int L, unsigned int* counters, float* red,
                                          float* result) {
    const int G = gridDim.x;
    if (G <= kSingleFold) {
        if (!ticket_last(counters, G)) return false;
        fold_partials(partials, G, L, red, result);
        return true;
    }
    const int group = blockIdx.x / kFoldGroup, ngroups = (G + kFoldGroup - 1) / kFoldGroup;
    const int gsize = min(kFoldGroup, G - group * kFoldGroup);
    if (!ticket_last(counters + 1 + group, gsize)) return false;
    fold_partials(partials + (int64_t)group * kFoldGroup * L, gsize, L, red, result);
    if (ngroups == 1) return true;
    for (int i = threadIdx.x; i < L; i += kThreads) gpartials[(int64_t)group * L + i] = result[i];
    if (!ticket_last(counters, ngroups)) return false;
    fold_partials(gpartials, ngroups, L, red, result);
    return true;
}

// hier_fold for CTAs of any size: kThreads virtual folding threads (same slices, same order => same bits as fold_partials),
// spread over however many threads the CTA has.
__device__ __forceinline__ void fold_partials_any(const float* partials, int G, int L, float* red /*>= kThreads*/,
                                                  float* result /*shared, >= L*/) {
    for (int base = 0; base < L; base += kThreads) {
        const int Lt = min(L - base, kThreads);
        int Lp = 1;
        while (Lp < Lt) Lp <<= 1;
        const int S = kThreads / Lp;
        for (int vt = threadIdx.x; vt < kThreads; vt += blockDim.x) {  // kThreads virtual folding threads
            const int i = vt % Lp, sl = vt / Lp;
            red[vt] = i < Lt ? fold_slice(partials + base + i, G, sl, S, L) : 0.f;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < Lt; t += blockDim.x) {
            float a = 0.f;
            for (int q = 0; q < S; ++q) a += red[q * Lp + t];
            result[base + t] = a;
        }
        __syncthreads();
    }
}
__device__ __forceinline__ bool hier_fold_any(float* partials, float* gpartials, int L, unsigned int* counters, float* red,
                                              float* result) {
    const int G = gridDim.x;
    if (G <= kSingleFold) {
        if (!ticket_last(counters, G)) return false;
        fold_partials_any(partials, G, L, red, result);
        return true;
    }
    const int group = blockIdx.x / kFoldGroup, ngroups = (G + kFoldGroup - 1) / kFoldGroup;
    const int gsize = min(kFoldGroup, G - group * kFoldGroup);
    if (!ticket_last(counters + 1 + group, gsize)) return false;
    fold_partials_any(partials + (int64_t)group * kFoldGroup * L, gsize, L, red, result);
    if (ngroups == 1) return true;
    for (int i = threadIdx.x; i < L; i += blockDim.x) gpartials[(int64_t)group * L + i] = result[i];
    if (!ticket_last(counters, ngroups)) return false;
    fold_partials_any(gpartials, ngroups, L, red, result);
    return true;
}

}  // namespace bg
