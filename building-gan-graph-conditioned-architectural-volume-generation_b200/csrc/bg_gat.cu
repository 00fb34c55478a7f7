// GATConv attention aggregation on the destination-sorted CSR: forward, first-order backward
// and second-order backward (WGAN-GP).  Replaces the index_select / scatter_reduce_(amax) /
// scatter_add_ chain PyG's GATConv runs for the reference (models.py:72,82,192,202).
//
// Mapping: one group of LANES = C/VEC consecutive lanes owns one node row (VEC = min(4, C)
// channels per lane, 128-bit loads for C >= 4); a warp therefore carries 32/LANES rows.  All
// per-row reductions are butterflies inside the group with a FIXED edge order (CSR order = the
// reference's COO order, self loop last) - no atomics, bitwise reproducible.
// Roofline: HBM.  Algorithmic bytes per launch are stated in DESIGN.md section "Kernels".
#include <stdlib.h>

#include <type_traits>

#include "bg_common.cuh"

namespace bg {

// ------------------------------------------------------------------------------------------
// Row-block mapping shared by the forward and first-order backward kernels.
//
// A CTA owns a CONTIGUOUS chunk of node rows and sweeps it front to back (on large graphs the grid is SMs x resident
// CTAs, so every SM walks a few contiguous windows): voxel grids are numbered floor-major, so a row's x/y neighbours
// and the row itself are touched by the same CTA within a short window and stay L1-resident; only the first touch of
// a row and the floor +-1 neighbours go to L2 / HBM.
//
// The per-edge scalar work (col index, s_j, leaky-relu, exp, softmax weight) is done ONCE per edge: edge slot t of a
// row (t < CAP = 8) lives in lane t % LANES of the row's lane group (register k = t / LANES) and is broadcast with one
// SHFL when the gather loop reaches it.  (r01b, before: every lane recomputed exp()/div per edge, 464 instructions per
// warp, issue-bound at 61% SM throughput.)  The gather runs as two blocks of four independent 128-bit loads; its trip
// count is warp-uniform (padding slots carry weight 0 and point at the row itself - an L1 hit), so all 32 lanes stay
// convergent and every shuffle uses the full mask.  Warps that hold a row with more than CAP edges take the generic
// per-row path.
//
// Graphs larger than L2 use the software-pipelined variants: a warp's dependent chain rowptr -> col -> s[j] -> h[j]
// is spread over four consecutive sweep iterations,
//   A (it+3): rowptr     B (it+2): col, d     C (it+1): s[j], L2 prefetch of far gather rows     D (it): math + gather,
// so every load was issued one iteration before it is consumed, and the DRAM misses of the gather (first touch of the
// row itself - prefetched as a sequential stream a few iterations ahead - and the floor +-1 neighbours) are taken by
// register-free prefetches.  Stages that refer to an iteration < 0 redo iteration 0 with degree 0.
// ------------------------------------------------------------------------------------------
constexpr int kGatMaxThreads = 1024;
template <int C>
constexpr bool kPipeOK = C >= 8;  // narrower rows hold 8 edge slots per lane: the pipeline registers would spill
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNearRows = 512;  // neighbours closer than this are L1/L2-resident through the sweep itself

// Lane mapping of the aggregation kernels: a group of LANES lanes owns one row; a lane holds HALVES 128-bit pieces of
// it (channels [4*sub, 4*sub+4) of each half-row), so rows of 64 / 128 floats take 8 / 16 lanes and a warp carries
// 4 / 2 of them: the per-row scalar work and the slot shuffles (which share the LSU data pipe with the gathers - the
// saturated unit in r01e, 89%) are amortised over twice as many rows, while every 128-bit load instruction still
// covers whole 128-byte lines.
template <int C>
struct GatMap {
    static constexpr int HALVES = C >= 64 ? 2 : 1;
    static constexpr int VEC = C >= 4 ? 4 : C;
#ifdef BG_GAT_V8
    // experiment: a lane's two 128-bit pieces are CONTIGUOUS (channels [8 sub, 8 sub + 8)), so that a gathered row is ONE 256-bit
    // load per lane (LDG.E.256) instead of two 128-bit loads half a row apart
    static constexpr bool V8 = HALVES == 2;
#else
    static constexpr bool V8 = false;
#endif
    static constexpr int LOFF = V8 ? 2 * VEC : VEC;          // channel offset of lane `sub`'s first piece: sub * LOFF
    static constexpr int HOFF = V8 ? VEC : C / 2;            // channel offset between a lane's pieces
    static constexpr int NV = VEC * HALVES;                  // floats per lane
    static constexpr int LANES = C / NV;
    static constexpr int RPW = 32 / LANES;                   // rows per warp
    static constexpr int EPL = LANES >= 8 ? 1 : 8 / LANES;   // edge slots per lane
    static constexpr int CAP = 8;                            // edge slots per row
    static constexpr int LINES = (C * 4 + 127) / 128;        // 128-byte lines per feature row
};
template <int C>
__device__ __forceinline__ void row_load(float (&v)[GatMap<C>::NV], const float* __restrict__ xrow, int sub) {
    using M = GatMap<C>;
#pragma unroll
    for (int hh = 0; hh < M::HALVES; ++hh) {
        Vec<M::VEC> t;
        t.load(xrow + hh * M::HOFF + sub * M::LOFF);
#pragma unroll
        for (int q = 0; q < M::VEC; ++q) v[hh * M::VEC + q] = t.v[q];
    }
}
template <int C>
__device__ __forceinline__ void row_store(const float (&v)[GatMap<C>::NV], float* __restrict__ xrow, int sub) {
    using M = GatMap<C>;
#pragma unroll
    for (int hh = 0; hh < M::HALVES; ++hh) {
        Vec<M::VEC> t;
#pragma unroll
        for (int q = 0; q < M::VEC; ++q) t.v[q] = v[hh * M::VEC + q];
        t.store(xrow + hh * M::HOFF + sub * M::LOFF);
    }
}
// channel index of element q of a lane's row piece
template <int C>
__device__ __forceinline__ int chan(int sub, int q) {
    using M = GatMap<C>;
    return (q / M::VEC) * M::HOFF + sub * M::LOFF + q % M::VEC;
}

__device__ __forceinline__ float rcp_fast(float x) {  // x >= 1 here (softmax denominators): MUFU.RCP, <= 1 ulp
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// value of edge slot t (compile-time constant after unrolling), broadcast inside the row's lane group
template <int LANES, int EPL, typename T>
__device__ __forceinline__ T slot_get(const T (&a)[EPL], int t) {
    if constexpr (LANES == 1) {
        return a[t];
    } else {
        return __shfl_sync(kFull, a[t / LANES], t % LANES, LANES);
    }
}

// a lane's HALVES pieces of the row at `xlane` (= row base + the lane's channel offset): two 128-bit loads, or one 256-bit load
template <int C>
__device__ __forceinline__ void load_pieces(Vec<GatMap<C>::VEC> (&x)[GatMap<C>::HALVES], const float* __restrict__ xlane) {
    using M = GatMap<C>;
    if constexpr (M::V8) {
        Vec<8> t;
        t.load(xlane);
#pragma unroll
        for (int q = 0; q < 4; ++q) x[0].v[q] = t.v[q], x[1].v[q] = t.v[4 + q];
    } else {
#pragma unroll
        for (int hh = 0; hh < M::HALVES; ++hh) x[hh].load(xlane + hh * M::HOFF);
    }
}
// acc += sum_{q < 4} w[t0+q] * X[idx[t0+q], :]   (4 independent 128-bit gathers in flight)
template <int C, int EPL>
__device__ __forceinline__ void gather_fma4(const float* __restrict__ xb, const int (&idx)[EPL], const float (&w)[EPL],
                                            int t0, float (&acc)[GatMap<C>::NV]) {
    constexpr int VEC = GatMap<C>::VEC, LANES = GatMap<C>::LANES, HALVES = GatMap<C>::HALVES;
    constexpr int NB = 4 / HALVES;  // 4 independent 128-bit loads in flight per lane
#pragma unroll
    for (int b = 0; b < 4; b += NB) {
        int jj[NB];
        float pp[NB];
        Vec<VEC> xv[NB][HALVES];
#pragma unroll
        for (int q = 0; q < NB; ++q) jj[q] = slot_get<LANES, EPL>(idx, t0 + b + q);
#pragma unroll
        for (int q = 0; q < NB; ++q) load_pieces<C>(xv[q], xb + (int64_t)jj[q] * C);
#pragma unroll
        for (int q = 0; q < NB; ++q) pp[q] = slot_get<LANES, EPL>(w, t0 + b + q);
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int hh = 0; hh < HALVES; ++hh)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[hh * VEC + v] = fmaf(pp[q], xv[q][hh].v[v], acc[hh * VEC + v]);
    }
}

// Sweep schedule: the rows are cut into chunks of `chunk_rows` (a multiple of the CTA's rows per iteration) that are
// dealt round-robin to the CTAs, and a CTA sweeps each of its chunks front to back.  Inside a chunk the sweep is
// contiguous (x / y neighbours hit L1); across the grid all CTAs advance through the same moving window of
// gridDim.x * chunk_rows rows, which is sized to stay L2-resident, so the floor +-1 neighbours (thousands of rows away)
// are served by L2 whatever the floor size - every feature row comes from HBM once.  (A plain 1/gridDim.x split only
// does that when the chunk length happens to divide the floor stride: measured 0.45 vs 0.59 of peak, sweep4.)
// Small graphs use ONE chunk per CTA (ipc_shift = 30).
struct Sweep {
    int chunk_rows, ipc_shift, stride, lane_off, N, niter;
    __device__ __forceinline__ Sweep(int N_, int chunk_rows_, int ipc_shift_, int rpw, int grow) {
        N = N_, chunk_rows = chunk_rows_, ipc_shift = ipc_shift_;
        const int warp = threadIdx.x >> 5;
        stride = (blockDim.x >> 5) * rpw;
        lane_off = warp * rpw + grow;
        const int nchunks = (N + chunk_rows - 1) / chunk_rows;
        const int nmy = ((int)blockIdx.x < nchunks) ? (nchunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        niter = 0;
        if (nmy > 0) {
            const int r0l = ((int)blockIdx.x + (nmy - 1) * (int)gridDim.x) * chunk_rows;
            const int left = min(N, r0l + chunk_rows) - (r0l + warp * rpw);
            niter = ((nmy - 1) << ipc_shift) + (left > 0 ? (left + stride - 1) / stride : 0);
        }
    }
    // this lane's row in sweep iteration max(it, 0) (may be >= N in the last chunk)
    __device__ __forceinline__ int raw(int it) const {
        const int itc = it > 0 ? it : 0;
        return ((int)blockIdx.x + (itc >> ipc_shift) * (int)gridDim.x) * chunk_rows +
               (itc & ((1 << ipc_shift) - 1)) * stride + lane_off;
    }
    __device__ __forceinline__ int at(int it) const {
        const int r = raw(it);
        return r < N ? r : N - 1;
    }
};

// L2 prefetch of the feature row `j` by the slot's owner lane when the row is far from the sweep position
template <int C>
__device__ __forceinline__ void prefetch_far_row(const float* __restrict__ x, int j, int row, bool ok) {
    const int dist = j - row;
    if (ok && (dist >= kNearRows || dist <= -kNearRows)) {
#pragma unroll
        for (int l = 0; l < GatMap<C>::LINES; ++l) prefetch_l2(x + (int64_t)j * C + l * 32);
    }
}
// sequential prefetch stream of the rows a warp will own `ahead` iterations later (first touch of the row itself)
template <int C>
__device__ __forceinline__ void prefetch_stream(const float* __restrict__ x, int row_ahead, int r1, int sub) {
    using M = GatMap<C>;
    if constexpr (M::V8) {  // contiguous pieces: lanes whose 32 bytes start a 128-byte line prefetch that line
        if (row_ahead < r1 && ((sub * M::LOFF * 4) & 127) == 0) prefetch_l2(x + (int64_t)row_ahead * C + sub * M::LOFF);
    } else if (row_ahead < r1 && ((sub * M::VEC * 4) & 127) == 0) {
#pragma unroll
        for (int hh = 0; hh < M::HALVES; ++hh) prefetch_l2(x + (int64_t)row_ahead * C + hh * (C / 2) + sub * M::VEC);
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int C>
__device__ __noinline__ void gat_fwd_row_generic(int row, int sub, unsigned gm, const int32_t* __restrict__ rowptr,
                                                 const int32_t* __restrict__ col, const float* __restrict__ h,
                                                 const float* __restrict__ s, const float* __restrict__ d,
                                                 const float* __restrict__ bias, float* __restrict__ out,
                                                 float* __restrict__ m_out, float* __restrict__ z_out, float slope) {
    using M = GatMap<C>;
    constexpr int NV = M::NV, LANES = M::LANES;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float di = __ldg(d + row);
    float mx = -INFINITY;
    for (int e = beg + sub; e < end; e += LANES) mx = fmaxf(mx, lrelu(__ldg(s + __ldg(col + e)) + di, slope));
    mx = gmax<LANES>(mx, gm);
    float zs = 0.f;
    for (int e = beg + sub; e < end; e += LANES) zs += expf(lrelu(__ldg(s + __ldg(col + e)) + di, slope) - mx);
    zs = gsum<LANES>(zs, gm) + 1e-16f;
    const float inv = rcp_fast(zs);
    float acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        float hv[NV];
        row_load<C>(hv, h + (int64_t)j * C, sub);
        const float p = expf(lrelu(__ldg(s + j) + di, slope) - mx) * inv;
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = fmaf(p, hv[v], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += bias ? __ldg(bias + chan<C>(sub, v)) : 0.f;
    row_store<C>(acc, out + (int64_t)row * C, sub);
    if (sub == 0) {
        m_out[row] = mx;
        z_out[row] = zs;
    }
}

// fast path: softmax over the register-resident slots (u = leaky-relu'd logits, -inf in padding slots) + gather
template <int C>
__device__ __forceinline__ void gat_fwd_row_fast(int row, bool valid, int maxdeg, int sub,
                                                 const int (&j)[GatMap<C>::EPL], const float (&u)[GatMap<C>::EPL],
                                                 const float* __restrict__ hb, const float* __restrict__ bias,
                                                 float* __restrict__ out, float* __restrict__ m_out,
                                                 float* __restrict__ z_out, float (&st1)[GatMap<C>::NV],
                                                 float (&st2)[GatMap<C>::NV], bool stats /* compile-time at the call sites */,
                                                 const float* __restrict__ kshift) {
    using M = GatMap<C>;
    constexpr int NV = M::NV, LANES = M::LANES, EPL = M::EPL;
    float p[EPL];
    float mx = u[0];
#pragma unroll
    for (int k = 1; k < EPL; ++k) mx = fmaxf(mx, u[k]);
    mx = group_max<LANES>(mx);
    float zs = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        p[k] = expf(u[k] - mx);  // exp(-inf) == 0 in the padding slots
        zs += p[k];
    }
    zs = group_sum<LANES>(zs) + 1e-16f;
    const float inv = rcp_fast(zs);
#pragma unroll
    for (int k = 0; k < EPL; ++k) p[k] *= inv;
    float acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
    gather_fma4<C, EPL>(hb, j, p, 0, acc);
    if (maxdeg > 4) gather_fma4<C, EPL>(hb, j, p, 4, acc);
    if (valid) {
        if (stats) {  // GraphNorm moments of (out - bias - k), k = row 0's aggregate: one shift common to every CTA, so the
                      // partial sums simply add, and a sample of the column keeps S2/n - (S1/n)^2 free of cancellation
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float a = acc[v] - kshift[chan<C>(sub, v)];
                st1[v] += a;
                st2[v] = fmaf(a, a, st2[v]);
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += bias ? __ldg(bias + chan<C>(sub, v)) : 0.f;
        row_store<C>(acc, out + (int64_t)row * C, sub);
        if (sub == 0) {
            m_out[row] = mx;
            z_out[row] = zs;
        }
    }
}

// GraphNorm statistics fused into the aggregation (mu, then var of o - alpha*mu over ALL rows): per-lane moment
// sums -> warp (lanes holding the same channels) -> CTA -> per-CTA partial -> two-level last-CTA fold (fixed order,
// bitwise reproducible) -> stats[3C] = (mean, rstd, var), exactly what gn_stats_kernel writes.
struct GnFuse {
    float* partials;         // [gridDim.x][2C]; NULL = no statistics
    float* gpartials;        // [ceil(gridDim.x / 16)][2C]
    unsigned int* counters;  // self-resetting tickets
    const float* alpha;      // GraphNorm mean_scale
    float* stats;            // out: [3C]
    float eps;
    int64_t n_rows;
};

template <int C>
__device__ __forceinline__ void gat_fwd_row_stats_from_out(int row, int sub, const float* __restrict__ out,
                                                           const float* __restrict__ bias, float (&st1)[GatMap<C>::NV],
                                                           float (&st2)[GatMap<C>::NV], const float* __restrict__ kshift) {
    float o[GatMap<C>::NV];
    row_load<C>(o, out + (int64_t)row * C, sub);  // written by this lane a moment ago (generic high-degree path)
#pragma unroll
    for (int v = 0; v < GatMap<C>::NV; ++v) {
        const float a = o[v] - (bias ? __ldg(bias + chan<C>(sub, v)) : 0.f) - kshift[chan<C>(sub, v)];
        st1[v] += a;
        st2[v] = fmaf(a, a, st2[v]);
    }
}

// Common shift of the fused statistics: k = h[0, :], one coalesced load per CTA (same bits everywhere).  Any value inside the
// range of the column does: the aggregate out_i - bias = sum_e p_e h_j is a convex combination of rows of h, so a sample row
// of h sits within the column's spread and S2/n - (S1/n)^2 stays free of cancellation.  (Round 1 aggregated row 0 itself here:
// a chain of four dependent loads - rowptr -> col -> s[j] -> h[j] - in front of every CTA, ~3 us of a 12 us launch.)
template <int C>
__device__ __forceinline__ void gat_fwd_row0_shift(const int32_t* __restrict__, const int32_t* __restrict__,
                                                   const float* __restrict__ h, const float* __restrict__,
                                                   const float* __restrict__, float, float* kshift) {
    if (threadIdx.x < C) kshift[threadIdx.x] = __ldg(h + threadIdx.x);
    __syncthreads();
}

template <int C>
__device__ __forceinline__ void gat_fwd_finish_stats(const GnFuse& gn, const float* __restrict__ bias, int sub,
                                                     float (&st1)[GatMap<C>::NV], float (&st2)[GatMap<C>::NV],
                                                     const float* __restrict__ kshift) {
    using M = GatMap<C>;
    constexpr int NV = M::NV, LANES = M::LANES;
    __shared__ float red[(kGatMaxThreads / 32) * 2 * C > kThreads ? (kGatMaxThreads / 32) * 2 * C : kThreads];
    __shared__ float sums[2 * C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int o = 16; o >= LANES; o >>= 1) {
            st1[v] += __shfl_xor_sync(kFull, st1[v], o);
            st2[v] += __shfl_xor_sync(kFull, st2[v], o);
        }
    if (lane < LANES) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            red[warp * 2 * C + chan<C>(sub, v)] = st1[v];
            red[warp * 2 * C + C + chan<C>(sub, v)] = st2[v];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < nwarp; ++w) t += red[w * 2 * C + i];
        gn.partials[(int64_t)blockIdx.x * 2 * C + i] = t;
    }
    __syncthreads();
    if (!hier_fold_any(gn.partials, gn.gpartials, 2 * C, gn.counters, red, sums)) return;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float n = (float)gn.n_rows;
        const float md = sums[c] / n;
        const float mean = (bias ? __ldg(bias + c) : 0.f) + kshift[c] + md;
        const float M2 = fmaxf(sums[C + c] - sums[c] * md, 0.f);
        const float shift = mean * (1.f - __ldg(gn.alpha + c));  // mean of (o - alpha*mu)
        const float var = (M2 + n * shift * shift) / n;
        gn.stats[c] = mean;
        gn.stats[C + c] = 1.f / sqrtf(var + gn.eps);
        gn.stats[2 * C + c] = var;
    }
}

// STATS variants carry 2*NV more live registers per lane: they are built for 768-thread CTAs (85 registers) instead of
// spilling inside the gather loop at the 64-register cap of a 1024-thread CTA.
constexpr int kGatStatsThreads = 768;
template <int C, bool PIPE, bool STATS>
__global__ void __launch_bounds__(STATS ? kGatStatsThreads : kGatMaxThreads, 1) gat_fwd_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ h,
    const float* __restrict__ s, const float* __restrict__ d, const float* __restrict__ bias,
    float* __restrict__ out, float* __restrict__ m_out, float* __restrict__ z_out, int N, float slope,
    int chunk_rows, int ipc_shift, int ahead, const GnFuse gn) {
    pdl_prologue();
    using M = GatMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES, EPL = M::EPL, CAP = M::CAP, RPW = M::RPW, NV = M::NV;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES, grow = lane / LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const Sweep sw(N, chunk_rows, ipc_shift, RPW, grow);
    const float* hb = h + sub * M::LOFF;
    constexpr bool stats = STATS;  // a compile-time switch: the 2*NV accumulators must not cost the plain kernel registers
    __shared__ float kshift[STATS ? C : 1];
    if constexpr (STATS) gat_fwd_row0_shift<C>(rowptr, col, h, s, d, slope, kshift);
    float st1[NV], st2[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) st1[v] = st2[v] = 0.f;
    if constexpr (!PIPE) {
        for (int it = 0; it < sw.niter; ++it) {
            const int row = sw.raw(it);
            const bool valid = row < N;
            const int rr = valid ? row : N - 1;
            const int beg = __ldg(rowptr + rr);
            const int deg = __ldg(rowptr + rr + 1) - beg;
            const int maxdeg = __reduce_max_sync(kFull, deg);
            if (maxdeg > CAP) {
                if (valid) {
                    gat_fwd_row_generic<C>(row, sub, gm, rowptr, col, h, s, d, bias, out, m_out, z_out, slope);
                    if (stats) gat_fwd_row_stats_from_out<C>(row, sub, out, bias, st1, st2, kshift);
                }
                continue;
            }
            const float di = __ldg(d + rr);
            int j[EPL];
            float u[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const bool ok = sub + k * LANES < deg;
                j[k] = ok ? __ldg(col + beg + sub + k * LANES) : rr;
                u[k] = ok ? lrelu(__ldg(s + j[k]) + di, slope) : -INFINITY;
            }
            gat_fwd_row_fast<C>(row, valid, maxdeg, sub, j, u, hb, bias, out, m_out, z_out, st1, st2, stats, kshift);
        }
    } else {
        if (sw.niter == 0 && !stats) return;
        int beg2 = 0, deg2 = 0, deg1 = 0, deg0 = 0;
        int j1[EPL], j0[EPL];
        float s0[EPL], d1 = 0.f, d0 = 0.f;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            j1[k] = j0[k] = sw.at(0);
            s0[k] = 0.f;
        }
        for (int it = -3; it < sw.niter; ++it) {
            // A: rowptr of iteration it+3
            const int r3 = sw.at(it + 3);
            const int beg3 = __ldg(rowptr + r3);
            const int deg3 = __ldg(rowptr + r3 + 1) - beg3;
            // B: col / d of iteration it+2
            const int r2 = sw.at(it + 2);
            int j2[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) j2[k] = sub + k * LANES < deg2 ? __ldg(col + beg2 + sub + k * LANES) : r2;
            const float d2 = __ldg(d + r2);
            // C: s[j] of iteration it+1, prefetch of its far gather rows and of the row stream `ahead` iterations on
            const int rc = sw.at(it + 1);
            float s1[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const bool ok = sub + k * LANES < deg1;
                s1[k] = __ldg(s + j1[k]);
                prefetch_far_row<C>(h, j1[k], rc, ok);
            }
            prefetch_stream<C>(h, sw.raw(it + 1 + ahead), N, sub);
            // D: softmax + aggregation of iteration it
            if (it >= 0) {
                const int row = sw.raw(it);
                const bool valid = row < N;
                const int maxdeg = __reduce_max_sync(kFull, deg0);
                if (maxdeg > CAP) {
                    if (valid) {
                        gat_fwd_row_generic<C>(row, sub, gm, rowptr, col, h, s, d, bias, out, m_out, z_out, slope);
                        if (stats) gat_fwd_row_stats_from_out<C>(row, sub, out, bias, st1, st2, kshift);
                    }
                } else {
                    float u[EPL];
#pragma unroll
                    for (int k = 0; k < EPL; ++k) u[k] = sub + k * LANES < deg0 ? lrelu(s0[k] + d0, slope) : -INFINITY;
                    gat_fwd_row_fast<C>(row, valid, maxdeg, sub, j0, u, hb, bias, out, m_out, z_out, st1, st2, stats, kshift);
                }
            }
            // rotate the pipeline registers
            deg0 = deg1, d0 = d1;
#pragma unroll
            for (int k = 0; k < EPL; ++k) j0[k] = j1[k], s0[k] = s1[k], j1[k] = j2[k];
            deg1 = deg2, d1 = d2;
            beg2 = beg3, deg2 = deg3;
        }
    }
    if (stats) gat_fwd_finish_stats<C>(gn, bias, sub, st1, st2, kshift);
}

// ------------------------------------------------------------------------------------------
// backward, pass A: per destination row (CSR).  Writes P[e] (softmax weight), DU[e] (d loss /
// d pre-activation logit u_e = s_j + d_i) and gsd[2i+1] = d loss / d d_i.
// ------------------------------------------------------------------------------------------
// GraphNorm backward "apply" fused into the destination pass (SURVEY H5b/H5c backward): the gradient at the conv output,
//   go = w r gy - w r^3 G1 (o - alpha mu) - alpha M[..]   with gy = gx1 * keep_scale * [x1 > 0]   (+ an injected cotangent),
// is computed by the lane group that owns the row right where the aggregation backward needs it (its g_i), and written
// once for the source pass / the bias gradient - instead of a separate elementwise launch that writes go and this kernel
// reading it back.  The per-channel constants are the ones gn_bwd_apply_kernel uses, evaluated by every CTA into shared
// memory (same expressions => same bits as the unfused pair).
struct GnBwdFuse {
    const float *gx1, *o, *x1, *inj;  // inj: optional cotangent added to go (second-order sweep)
    const float *w, *alpha, *stats, *bstats;
    float* go_out;
    float keep_scale;
};
template <int C>
__device__ __forceinline__ void gn_bwd_consts(const GnBwdFuse& f, float* kc) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mu = f.stats[c], r = f.stats[C + c], a = f.alpha[c], wc = f.w[c];
        const float G0 = f.bstats[c], G1 = f.bstats[C + c];
        kc[c] = wc * r;
        kc[C + c] = wc * r * r * r * G1;
        kc[2 * C + c] = a * (wc * r * G0 - wc * r * r * r * G1 * mu * (1.f - a));
        kc[3 * C + c] = a * mu;
    }
    __syncthreads();
}
// g_i of row rr: loaded (FUSE = false) or computed from the GraphNorm backward and stored to go_out (FUSE = true)
template <int C, bool FUSE>
__device__ __forceinline__ void load_gi(float (&gi)[GatMap<C>::NV], const float* __restrict__ gout, const GnBwdFuse& f,
                                        const float* kc, int rr, int sub, bool valid) {
    constexpr int NV = GatMap<C>::NV;
    if constexpr (!FUSE) {
        row_load<C>(gi, gout + (int64_t)rr * C, sub);
    } else {
        float g[NV], ov[NV], xv[NV];
        row_load<C>(g, f.gx1 + (int64_t)rr * C, sub);
        row_load<C>(ov, f.o + (int64_t)rr * C, sub);
        row_load<C>(xv, f.x1 + (int64_t)rr * C, sub);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = chan<C>(sub, v);
            const float gy = xv[v] > 0.f ? g[v] * f.keep_scale : 0.f;
            gi[v] = gy * kc[c] - (ov[v] - kc[3 * C + c]) * kc[C + c] - kc[2 * C + c];
        }
        if (f.inj) {
            float iv[NV];
            row_load<C>(iv, f.inj + (int64_t)rr * C, sub);
#pragma unroll
            for (int v = 0; v < NV; ++v) gi[v] += iv[v];
        }
        if (valid) row_store<C>(gi, f.go_out + (int64_t)rr * C, sub);
    }
}

template <int C>
__device__ __noinline__ void gat_bwd_dst_row_generic(int row, int sub, unsigned gm, const int32_t* __restrict__ rowptr,
                                                     const int32_t* __restrict__ col, const float* __restrict__ gout,
                                                     const float* __restrict__ h, const float* __restrict__ s,
                                                     const float* __restrict__ d, const float* __restrict__ m_in,
                                                     const float* __restrict__ z_in, float* __restrict__ P,
                                                     float* __restrict__ DU, float* __restrict__ gsd, float slope) {
    using M = GatMap<C>;
    constexpr int NV = M::NV, LANES = M::LANES;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float di = __ldg(d + row), mi = __ldg(m_in + row), inv = rcp_fast(__ldg(z_in + row));
    float gi[NV];
    row_load<C>(gi, gout + (int64_t)row * C, sub);
    float r = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        float hv[NV];
        row_load<C>(hv, h + (int64_t)j * C, sub);
        float c = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) c = fmaf(gi[v], hv[v], c);
        c = gsum<LANES>(c, gm);
        const float p = expf(lrelu(__ldg(s + j) + di, slope) - mi) * inv;
        r = fmaf(p, c, r);
        if (sub == 0) {
            P[e] = p;
            DU[e] = c;
        }
    }
    __syncwarp(gm);
    float gd = 0.f;
    for (int e = beg + sub; e < end; e += LANES) {
        const float u = __ldg(s + __ldg(col + e)) + di;
        const float du = lrelu_grad(u, slope) * P[e] * (DU[e] - r);
        DU[e] = du;
        gd += du;
    }
    gd = gsum<LANES>(gd, gm);
    if (sub == 0) gsd[2 * (int64_t)row + 1] = gd;
}

// c[slot] = g_i . h_j for 4 slots: gather, partial dots, butterfly, the slot's owner lane keeps the result
template <int C, int EPL>
__device__ __forceinline__ void gather_dot4(const float* __restrict__ hb, const int (&j)[EPL],
                                            const float (&gi)[GatMap<C>::NV], int t0, int sub, float (&c)[EPL]) {
    constexpr int VEC = GatMap<C>::VEC, LANES = GatMap<C>::LANES, HALVES = GatMap<C>::HALVES;
    constexpr int NB = 4 / HALVES;
#pragma unroll
    for (int b = 0; b < 4; b += NB) {
        int jj[NB];
        Vec<VEC> hv[NB][HALVES];
        float cc[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) jj[q] = slot_get<LANES, EPL>(j, t0 + b + q);
#pragma unroll
        for (int q = 0; q < NB; ++q) load_pieces<C>(hv[q], hb + (int64_t)jj[q] * C);
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            cc[q] = 0.f;
#pragma unroll
            for (int hh = 0; hh < HALVES; ++hh)
#pragma unroll
                for (int v = 0; v < VEC; ++v) cc[q] = fmaf(gi[hh * VEC + v], hv[q][hh].v[v], cc[q]);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            cc[q] = group_sum<LANES>(cc[q]);
            const int t = t0 + b + q;
            if (sub == t % LANES) c[t / LANES] = cc[q];
        }
    }
}

template <int C, bool FUSE>
__device__ __forceinline__ void gat_bwd_dst_row_fast(int row, bool valid, int beg, int deg, int maxdeg, int sub,
                                                     const int (&j)[GatMap<C>::EPL], const float (&sj)[GatMap<C>::EPL],
                                                     float di, float mi, float zi, const float* __restrict__ gout,
                                                     const float* __restrict__ hb, float* __restrict__ P,
                                                     float* __restrict__ DU, float* __restrict__ gsd, float slope,
                                                     int rr, const GnBwdFuse& fuse, const float* kc) {
    using M = GatMap<C>;
    constexpr int NV = M::NV, LANES = M::LANES, EPL = M::EPL;
    float gi[NV];
    load_gi<C, FUSE>(gi, gout, fuse, kc, rr, sub, valid);
    const float inv = rcp_fast(zi);
    float p[EPL], lg[EPL], c[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const bool ok = sub + k * LANES < deg;
        const float u = sj[k] + di;
        p[k] = ok ? expf(lrelu(u, slope) - mi) * inv : 0.f;
        lg[k] = lrelu_grad(u, slope);
        c[k] = 0.f;
    }
    gather_dot4<C, EPL>(hb, j, gi, 0, sub, c);
    if (maxdeg > 4) gather_dot4<C, EPL>(hb, j, gi, 4, sub, c);
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) r = fmaf(p[k], c[k], r);
    r = group_sum<LANES>(r);
    float gd = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const float du = lg[k] * p[k] * (c[k] - r);
        if (valid && sub + k * LANES < deg) {
            P[beg + sub + k * LANES] = p[k];
            DU[beg + sub + k * LANES] = du;
        }
        gd += du;
    }
    gd = group_sum<LANES>(gd);
    if (valid && sub == 0) gsd[2 * (int64_t)row + 1] = gd;
}

template <int C, bool PIPE, bool FUSE>
__global__ void __launch_bounds__(FUSE ? kGatStatsThreads : kGatMaxThreads, 1) gat_bwd_dst_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ gout_in,
    const float* __restrict__ h, const float* __restrict__ s, const float* __restrict__ d,
    const float* __restrict__ m_in, const float* __restrict__ z_in, float* __restrict__ P,
    float* __restrict__ DU, float* __restrict__ gsd, int N, float slope, int chunk_rows, int ipc_shift, int ahead,
    const GnBwdFuse fuse) {
    pdl_prologue();
    using M = GatMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES, EPL = M::EPL, CAP = M::CAP, RPW = M::RPW;
    __shared__ float kc[FUSE ? 4 * C : 1];
    if constexpr (FUSE) gn_bwd_consts<C>(fuse, kc);
    const float* gout = FUSE ? fuse.go_out : gout_in;  // the generic row path reads the row back from where it was stored
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES, grow = lane / LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const Sweep sw(N, chunk_rows, ipc_shift, RPW, grow);
    const float* hb = h + sub * M::LOFF;
    if constexpr (!PIPE) {
        for (int it = 0; it < sw.niter; ++it) {
            const int row = sw.raw(it);
            const bool valid = row < N;
            const int rr = valid ? row : N - 1;
            const int beg = __ldg(rowptr + rr);
            const int deg = __ldg(rowptr + rr + 1) - beg;
            const int maxdeg = __reduce_max_sync(kFull, deg);
            if (maxdeg > CAP) {
                if (valid) {
                    if constexpr (FUSE) {
                        float gi[M::NV];
                        load_gi<C, true>(gi, gout, fuse, kc, row, sub, true);
                    }
                    gat_bwd_dst_row_generic<C>(row, sub, gm, rowptr, col, gout, h, s, d, m_in, z_in, P, DU, gsd, slope);
                }
                continue;
            }
            int j[EPL];
            float sj[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                j[k] = sub + k * LANES < deg ? __ldg(col + beg + sub + k * LANES) : rr;
                sj[k] = __ldg(s + j[k]);
            }
            gat_bwd_dst_row_fast<C, FUSE>(row, valid, beg, deg, maxdeg, sub, j, sj, __ldg(d + rr), __ldg(m_in + rr),
                                          __ldg(z_in + rr), gout, hb, P, DU, gsd, slope, rr, fuse, kc);
        }
    } else {
        if (sw.niter == 0) return;
        int beg2 = 0, deg2 = 0, beg1 = 0, deg1 = 0, beg0 = 0, deg0 = 0;
        int j1[EPL], j0[EPL];
        float s0[EPL], d1 = 0.f, d0 = 0.f, m0 = 0.f, z0 = 1.f;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            j1[k] = j0[k] = sw.at(0);
            s0[k] = 0.f;
        }
        for (int it = -3; it < sw.niter; ++it) {
            const int r3 = sw.at(it + 3);
            const int beg3 = __ldg(rowptr + r3);
            const int deg3 = __ldg(rowptr + r3 + 1) - beg3;
            const int r2 = sw.at(it + 2);
            int j2[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) j2[k] = sub + k * LANES < deg2 ? __ldg(col + beg2 + sub + k * LANES) : r2;
            const float d2 = __ldg(d + r2);
            const int rc = sw.at(it + 1);
            float s1[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                s1[k] = __ldg(s + j1[k]);
                prefetch_far_row<C>(h, j1[k], rc, sub + k * LANES < deg1);
            }
            const float m1 = __ldg(m_in + rc), z1 = __ldg(z_in + rc);
            prefetch_stream<C>(h, sw.raw(it + 1 + ahead), N, sub);
            if constexpr (FUSE) {
                prefetch_stream<C>(fuse.gx1, sw.raw(it + 1 + ahead), N, sub);
                prefetch_stream<C>(fuse.o, sw.raw(it + 1 + ahead), N, sub);
                prefetch_stream<C>(fuse.x1, sw.raw(it + 1 + ahead), N, sub);
            } else {
                prefetch_stream<C>(gout, sw.raw(it + 1 + ahead), N, sub);
            }
            if (it >= 0) {
                const int row = sw.raw(it);
                const bool valid = row < N;
                const int maxdeg = __reduce_max_sync(kFull, deg0);
                if (maxdeg > CAP) {
                    if (valid) {
                        if constexpr (FUSE) {
                            float gi[M::NV];
                            load_gi<C, true>(gi, gout, fuse, kc, row, sub, true);
                        }
                        gat_bwd_dst_row_generic<C>(row, sub, gm, rowptr, col, gout, h, s, d, m_in, z_in, P, DU, gsd, slope);
                    }
                } else {
                    gat_bwd_dst_row_fast<C, FUSE>(row, valid, beg0, deg0, maxdeg, sub, j0, s0, d0, m0, z0, gout, hb, P, DU, gsd,
                                                  slope, valid ? row : N - 1, fuse, kc);
                }
            }
            beg0 = beg1, deg0 = deg1, d0 = d1, m0 = m1, z0 = z1;
#pragma unroll
            for (int k = 0; k < EPL; ++k) j0[k] = j1[k], s0[k] = s1[k], j1[k] = j2[k];
            beg1 = beg2, deg1 = deg2, d1 = d2;
            beg2 = beg3, deg2 = deg3;
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward, pass B: per source row (CSC).  out_tot[j] = sum_{e in out(j)} P[e] * G[i(e)]
//   + (sum_e DU[e]) * a_src + gsd[2j+1] * a_dst ;  gsd[2j] = sum_e DU[e].
// Shared by the first- and second-order backward (different P/DU/G).
// ------------------------------------------------------------------------------------------
template <int C>
__device__ __noinline__ void gat_bwd_src_row_generic(int row, int sub, const int32_t* __restrict__ cscptr,
                                                     const int32_t* __restrict__ cscrow, const int32_t* __restrict__ perm,
                                                     const float* __restrict__ P, const float* __restrict__ DU,
                                                     const float* __restrict__ G, const float* __restrict__ a_src,
                                                     const float* __restrict__ a_dst, float* __restrict__ out_tot,
                                                     float* __restrict__ gsd, const float* __restrict__ add) {
    using M = GatMap<C>;
    constexpr int NV = M::NV;
    const int beg = __ldg(cscptr + row), end = __ldg(cscptr + row + 1);
    float acc[NV], inj[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = inj[v] = 0.f;
    if (add) row_load<C>(inj, add + (int64_t)row * C, sub);
    float gs = 0.f;
    for (int k = beg; k < end; ++k) {
        const int i = __ldg(cscrow + k), e = __ldg(perm + k);
        float gv[NV];
        row_load<C>(gv, G + (int64_t)i * C, sub);
        const float p = P[e];
        gs += DU[e];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = fmaf(p, gv[v], acc[v]);
    }
    const float gd = gsd[2 * (int64_t)row + 1];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += gs * __ldg(a_src + chan<C>(sub, v)) + gd * __ldg(a_dst + chan<C>(sub, v));
    if (add) {
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += inj[v];
    }
    row_store<C>(acc, out_tot + (int64_t)row * C, sub);
    if (sub == 0) gsd[2 * (int64_t)row] = gs;
}

template <int C>
__device__ __forceinline__ void gat_bwd_src_row_fast(int row, bool valid, int maxdeg, int sub,
                                                     const int (&i)[GatMap<C>::EPL], const float (&p)[GatMap<C>::EPL],
                                                     const float (&du)[GatMap<C>::EPL], float gd,
                                                     const float* __restrict__ gb, const float* __restrict__ a_src,
                                                     const float* __restrict__ a_dst, float* __restrict__ out_tot,
                                                     float* __restrict__ gsd, const float* __restrict__ add) {
    using M = GatMap<C>;
    constexpr int NV = M::NV, LANES = M::LANES, EPL = M::EPL;
    float gs = 0.f;
#pragma unroll
    for (int k = 0; k < EPL; ++k) gs += du[k];
    gs = group_sum<LANES>(gs);
    float acc[NV], inj[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = inj[v] = 0.f;
    // injected cotangent at this row (second-order sweep): issued with the gather, added last - the same single rounding as
    // the separate axpy it replaces (out + 1.0 * add)
    if (add && valid) row_load<C>(inj, add + (int64_t)row * C, sub);
    gather_fma4<C, EPL>(gb, i, p, 0, acc);
    if (maxdeg > 4) gather_fma4<C, EPL>(gb, i, p, 4, acc);
    if (valid) {
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] += gs * __ldg(a_src + chan<C>(sub, v)) + gd * __ldg(a_dst + chan<C>(sub, v));
        if (add) {
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] += inj[v];
        }
        row_store<C>(acc, out_tot + (int64_t)row * C, sub);
        if (sub == 0) gsd[2 * (int64_t)row] = gs;
    }
}

template <int C, bool PIPE>
__global__ void __launch_bounds__(kGatMaxThreads, 1) gat_bwd_src_kernel(
    const int32_t* __restrict__ cscptr, const int32_t* __restrict__ cscrow, const int32_t* __restrict__ perm,
    const float* __restrict__ P, const float* __restrict__ DU, const float* __restrict__ G,
    const float* __restrict__ a_src, const float* __restrict__ a_dst, float* __restrict__ out_tot,
    float* __restrict__ gsd, const float* __restrict__ add, int N, int chunk_rows, int ipc_shift, int ahead) {
    pdl_prologue();
    using M = GatMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES, EPL = M::EPL, CAP = M::CAP, RPW = M::RPW;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES, grow = lane / LANES;
    const Sweep sw(N, chunk_rows, ipc_shift, RPW, grow);
    const float* gb = G + sub * M::LOFF;
    if constexpr (!PIPE) {
        for (int it = 0; it < sw.niter; ++it) {
            const int row = sw.raw(it);
            const bool valid = row < N;
            const int rr = valid ? row : N - 1;
            const int beg = __ldg(cscptr + rr);
            const int deg = __ldg(cscptr + rr + 1) - beg;
            const int maxdeg = __reduce_max_sync(kFull, deg);
            if (maxdeg > CAP) {
                if (valid) gat_bwd_src_row_generic<C>(row, sub, cscptr, cscrow, perm, P, DU, G, a_src, a_dst, out_tot, gsd, add);
                continue;
            }
            int i[EPL];
            float p[EPL], du[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const bool ok = sub + k * LANES < deg;
                i[k] = ok ? __ldg(cscrow + beg + sub + k * LANES) : rr;
                const int e = ok ? __ldg(perm + beg + sub + k * LANES) : 0;
                p[k] = ok ? P[e] : 0.f;
                du[k] = ok ? DU[e] : 0.f;
            }
            gat_bwd_src_row_fast<C>(row, valid, maxdeg, sub, i, p, du, gsd[2 * (int64_t)rr + 1], gb, a_src, a_dst, out_tot, gsd, add);
        }
    } else {
        if (sw.niter == 0) return;
        int beg2 = 0, deg2 = 0, deg1 = 0, deg0 = 0;
        int i1[EPL], e1[EPL], i0[EPL];
        float p0[EPL], du0[EPL], gd0 = 0.f;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            i1[k] = i0[k] = sw.at(0);
            e1[k] = 0;
            p0[k] = du0[k] = 0.f;
        }
        for (int it = -3; it < sw.niter; ++it) {
            const int r3 = sw.at(it + 3);
            const int beg3 = __ldg(cscptr + r3);
            const int deg3 = __ldg(cscptr + r3 + 1) - beg3;
            const int r2 = sw.at(it + 2);
            int i2[EPL], e2[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const bool ok = sub + k * LANES < deg2;
                i2[k] = ok ? __ldg(cscrow + beg2 + sub + k * LANES) : r2;
                e2[k] = ok ? __ldg(perm + beg2 + sub + k * LANES) : 0;
            }
            const int rc = sw.at(it + 1);
            float p1[EPL], du1[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const bool ok = sub + k * LANES < deg1;
                p1[k] = ok ? P[e1[k]] : 0.f;
                du1[k] = ok ? DU[e1[k]] : 0.f;
                prefetch_far_row<C>(G, i1[k], rc, ok);
            }
            const float gd1 = gsd[2 * (int64_t)rc + 1];
            prefetch_stream<C>(G, sw.raw(it + 1 + ahead), N, sub);
            if (it >= 0) {
                const int row = sw.raw(it);
                const bool valid = row < N;
                const int maxdeg = __reduce_max_sync(kFull, deg0);
                if (maxdeg > CAP) {
                    if (valid) gat_bwd_src_row_generic<C>(row, sub, cscptr, cscrow, perm, P, DU, G, a_src, a_dst, out_tot, gsd, add);
                } else {
                    gat_bwd_src_row_fast<C>(row, valid, maxdeg, sub, i0, p0, du0, gd0, gb, a_src, a_dst, out_tot, gsd, add);
                }
            }
            deg0 = deg1, gd0 = gd1;
#pragma unroll
            for (int k = 0; k < EPL; ++k) i0[k] = i1[k], p0[k] = p1[k], du0[k] = du1[k], i1[k] = i2[k], e1[k] = e2[k];
            deg1 = deg2;
            beg2 = beg3, deg2 = deg3;
        }
    }
}

// ------------------------------------------------------------------------------------------
// second-order backward, destination pass.  Notation (re-derived and
// checked against fp64 autograd in tests/test_oracle_pyg.py): edge e = (j -> i),
//   c = g_i.h_j   a = Ht_j.g_i   t = (St_j + Dt_i) phi'(u)   r = sum p c   T = sum p t
//   w = t - T     pi = a + c w - t r     Pi = sum p pi
//   gt_i = sum p (Ht_j + w h_j)          du2 = phi'(u) p (pi - Pi)       dt_i = sum du2
// scratch A0..A3 are E floats each; on exit A1 = p*w (for the source pass), A2 = du2.
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gat_bwd2_dst_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ Ht,
    const float* __restrict__ St, const float* __restrict__ Dt, const float* __restrict__ gout,
    const float* __restrict__ h, const float* __restrict__ s, const float* __restrict__ d,
    const float* __restrict__ m_in, const float* __restrict__ z_in, float* __restrict__ A0,
    float* __restrict__ A1, float* __restrict__ A2, float* __restrict__ A3, float* __restrict__ gt,
    float* __restrict__ sdt, int64_t N, float slope) {
    pdl_prologue();
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float di = __ldg(d + row), mi = __ldg(m_in + row), zi = __ldg(z_in + row), Dti = __ldg(Dt + row);
    Vec<VEC> gi;
    gi.load(gout + row * C + sub * VEC);
    const float* hb = h + sub * VEC;
    const float* Hb = Ht + sub * VEC;
    if (end - beg <= 8) {
        // Rows with at most 8 in-edges (every row of a 6-neighbour grid + self loop): everything of the row's edges lives in
        // registers.  The loop version below walks the edges one at a time through three passes (col -> gathers -> shuffles ->
        // scratch store, re-read, re-gather): ~7 dependent L2 round trips per pass, 14-24 us per launch on the critic update's
        // chain.  Here the 8 column indices, then the 8 + 8 scalars and 16 row pieces are each in flight together, p / c / a / t
        // stay in registers (no A0 / A3 scratch round trip) and the rows are not gathered a second time.  Same sums in the same
        // edge order (Pi is summed by every lane over all edges instead of lane-strided).
        const int deg = end - beg;
        int jj[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) jj[q] = q < deg ? __ldg(col + beg + q) : (int)row;
        float sj[8], stj[8];
        Vec<VEC> hv[8], Hv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            sj[q] = __ldg(s + jj[q]);
            stj[q] = __ldg(St + jj[q]);
            hv[q].load(hb + (int64_t)jj[q] * C);
            Hv[q].load(Hb + (int64_t)jj[q] * C);
        }
        float pq[8], cq[8], aq[8], tq[8], gq[8];
        float r = 0.f, T = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float c = 0.f, a = 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                c = fmaf(gi.v[v], hv[q].v[v], c);
                a = fmaf(gi.v[v], Hv[q].v[v], a);
            }
            cq[q] = gsum<LANES>(c, gm);
            aq[q] = gsum<LANES>(a, gm);
            const float u = sj[q] + di;
            gq[q] = lrelu_grad(u, slope);
            pq[q] = q < deg ? expf(lrelu(u, slope) - mi) / zi : 0.f;
            tq[q] = (stj[q] + Dti) * gq[q];
            r = fmaf(pq[q], cq[q], r);   // padding slots: p = 0 leaves r, T, Pi and the sums below unchanged
            T = fmaf(pq[q], tq[q], T);
        }
        float Pi = 0.f, piq[8], pwq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float w = tq[q] - T;
            piq[q] = aq[q] + cq[q] * w - tq[q] * r;
            Pi = fmaf(pq[q], piq[q], Pi);
            pwq[q] = pq[q] * w;
        }
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        float dt = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = fmaf(pq[q], Hv[q].v[v], fmaf(pwq[q], hv[q].v[v], acc[v]));
            const float du2 = gq[q] * pq[q] * (piq[q] - Pi);
            dt += du2;
            if (sub == 0 && q < deg) {
                A1[beg + q] = pwq[q];
                A2[beg + q] = du2;
            }
        }
        Vec<VEC> o;
#pragma unroll
        for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
        o.store(gt + row * C + sub * VEC);
        if (sub == 0) sdt[2 * row + 1] = dt;
        return;
    }
    float r = 0.f, T = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        Vec<VEC> hv, Hv;
        hv.load(hb + (int64_t)j * C);
        Hv.load(Hb + (int64_t)j * C);
        float c = 0.f, a = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            c = fmaf(gi.v[v], hv.v[v], c);
            a = fmaf(gi.v[v], Hv.v[v], a);
        }
        c = gsum<LANES>(c, gm);
        a = gsum<LANES>(a, gm);
        const float u = __ldg(s + j) + di;
        const float p = expf(lrelu(u, slope) - mi) / zi;
        const float t = (__ldg(St + j) + Dti) * lrelu_grad(u, slope);
        r = fmaf(p, c, r);
        T = fmaf(p, t, T);
        if (sub == 0) {
            A0[e] = p;
            A1[e] = c;
            A2[e] = a;
            A3[e] = t;
        }
    }
    __syncwarp(gm);
    float Pi = 0.f;
    for (int e = beg + sub; e < end; e += LANES) {
        const float p = A0[e], c = A1[e], a = A2[e], t = A3[e];
        const float w = t - T;
        const float pi = a + c * w - t * r;
        Pi = fmaf(p, pi, Pi);
        A1[e] = p * w;
        A3[e] = pi;
    }
    Pi = gsum<LANES>(Pi, gm);
    __syncwarp(gm);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    float dt = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        Vec<VEC> hv, Hv;
        hv.load(hb + (int64_t)j * C);
        Hv.load(Hb + (int64_t)j * C);
        const float p = A0[e], pw = A1[e];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(p, Hv.v[v], fmaf(pw, hv.v[v], acc[v]));
        const float u = __ldg(s + j) + di;
        const float du2 = lrelu_grad(u, slope) * p * (A3[e] - Pi);
        dt += du2;
        if (sub == 0) A2[e] = du2;
    }
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
    o.store(gt + row * C + sub * VEC);
    if (sub == 0) sdt[2 * row + 1] = dt;
}

// Launch geometry.  Small graphs (working set L2-resident, latency-bound): 256-thread CTAs, one sweep iteration each
// where possible.  Large graphs: SMs x `cps` CTAs of `threads` threads, each sweeping one contiguous row chunk.
// bg_tune() overrides (used by the kernel sweep in bench.py --workload c4).
static int g_tune[8] = {1024, 1, 128, 1, 0, 64, 0, 0};  // see bg_tune() in include/bg_b200.h
struct GatCfg {
    unsigned grid, threads;
    int chunk_rows, ipc_shift, ahead;
    bool pipe;
};
// `fat_ctas`: the forward kernel with fused statistics ends in a cross-CTA fold of one partial vector per CTA; in the small
// regime it therefore runs max_threads-wide CTAs (768 threads: ~150 CTAs at N ~ 15 k, C = 64 instead of ~470 of 256 threads), which
// keeps the fold to ONE ticket round over <= kSingleFold partials (measured: 13.7 -> ~7 us per launch on the chain).
template <int C>
static GatCfg gat_cfg(int64_t N, int max_threads = kGatMaxThreads, bool fat_ctas = false) {
    const bool small = !g_tune[4] && N * C * 4 < (int64_t)(24 << 20);
    const int threads = small ? (fat_ctas ? max_threads : 256) : (g_tune[0] < max_threads ? g_tune[0] : max_threads);
    const int64_t rows_iter = (int64_t)(threads / 32) * GatMap<C>::RPW;
    const int64_t iters = ceil_div(N, rows_iter);
    if (small) {  // one contiguous chunk per CTA, one sweep iteration each while the grid fits 8 CTAs per SM
        const int64_t cap = (int64_t)kSMs * 8;
        int64_t grid = iters < cap ? iters : cap;
        const int64_t rpc = ceil_div(iters, grid) * rows_iter;
        grid = ceil_div(N, rpc);
        return GatCfg{(unsigned)grid, (unsigned)threads, (int)rpc, 30, 0, false};
    }
    // chunks of ~g_tune[5] KiB of feature rows (power-of-two iterations per chunk), dealt round-robin
    int shift = 0;
    while (shift < 12 && (rows_iter << (shift + 1)) * C * 4 <= (int64_t)g_tune[5] * 1024) ++shift;
    const int64_t chunk = rows_iter << shift;
    const int64_t nchunks = ceil_div(N, chunk), cap = g_tune[6] > 0 ? g_tune[6] : (int64_t)kSMs * g_tune[1];
    const int64_t grid = nchunks < cap ? nchunks : cap;
    return GatCfg{(unsigned)grid, (unsigned)threads, (int)chunk, shift, (int)ceil_div(g_tune[2], rows_iter), g_tune[3] != 0};
}
#define BG_GAT_LAUNCH(KERNEL, ...)                                                      \
    do {                                                                                \
        if (kPipeOK<C> && c.pipe) /* narrower rows hold 8 edge slots per lane: the pipeline registers would spill */ \
            launch_k(KERNEL<C, kPipeOK<C>>, c.grid, c.threads, 0, st, __VA_ARGS__, c.chunk_rows, c.ipc_shift, c.ahead);  \
        else                                                                            \
            launch_k(KERNEL<C, false>, c.grid, c.threads, 0, st, __VA_ARGS__, c.chunk_rows, c.ipc_shift, c.ahead); \
    } while (0)

template <int C>
static int launch_fwd(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias,
                      float* out, float* m, float* z, float slope, cudaStream_t st, const float* gn_alpha = nullptr,
                      float gn_eps = 0.f, float* gn_stats = nullptr, float* ws = nullptr) {
    const GatCfg c = gat_cfg<C>(g->N, gn_stats ? kGatStatsThreads : kGatMaxThreads, gn_stats != nullptr);
    GnFuse gn{};
    if (gn_stats) {  // workspace: [counters 4 KiB][partials grid x 2C][group partials]
        gn.counters = reinterpret_cast<unsigned int*>(ws);
        gn.partials = ws + kCounterBytes / sizeof(float);
        gn.gpartials = gn.partials + (size_t)c.grid * 2 * C;
        gn.alpha = gn_alpha, gn.stats = gn_stats, gn.eps = gn_eps, gn.n_rows = g->N;
    }
#define BG_FWD(PIPE_, STATS_)                                                                                              \
    launch_k(gat_fwd_kernel<C, PIPE_, STATS_>, c.grid, c.threads, 0, st, g->rowptr, g->col, h, s, d, bias, out, m, z, (int)g->N, slope, \
                                                                   c.chunk_rows, c.ipc_shift, c.ahead, gn)
    if (kPipeOK<C> && c.pipe) {
        if (gn_stats) BG_FWD(kPipeOK<C>, true); else BG_FWD(kPipeOK<C>, false);
    } else {
        if (gn_stats) BG_FWD(false, true); else BG_FWD(false, false);
    }
#undef BG_FWD
    return check_launch("bg_gat_fwd");
}
template <int C>
static int launch_bwd(const BgGraph* g, const float* gout, const float* h, const float* s, const float* d,
                      const float* m, const float* z, const float* a_src, const float* a_dst, float* P,
                      float* DU, float* gh_tot, float* gsd, float slope, cudaStream_t st, const float* inj_h = nullptr) {
    const GatCfg c = gat_cfg<C>(g->N);
    const GnBwdFuse nofuse{};
    if (kPipeOK<C> && c.pipe)
        launch_k(gat_bwd_dst_kernel<C, kPipeOK<C>, false>, c.grid, c.threads, 0, st, g->rowptr, g->col, gout, h, s, d, m, z, P, DU, gsd,
                                                                            (int)g->N, slope, c.chunk_rows, c.ipc_shift, c.ahead, nofuse);
    else
        launch_k(gat_bwd_dst_kernel<C, false, false>, c.grid, c.threads, 0, st, g->rowptr, g->col, gout, h, s, d, m, z, P, DU, gsd,
                                                                         (int)g->N, slope, c.chunk_rows, c.ipc_shift, c.ahead, nofuse);
    BG_GAT_LAUNCH(gat_bwd_src_kernel, g->cscptr, g->cscrow, g->perm, P, DU, gout, a_src, a_dst, gh_tot, gsd, inj_h, (int)g->N);
    return check_launch("bg_gat_bwd");
}
// destination pass with the GraphNorm backward fused in (go is produced here), then the source pass on that go
template <int C>
static int launch_bwd_gn(const BgGraph* g, const GnBwdFuse& f, const float* h, const float* s, const float* d, const float* m,
                         const float* z, const float* a_src, const float* a_dst, float* P, float* DU, float* gh_tot, float* gsd,
                         float slope, cudaStream_t st, const float* inj_h = nullptr) {
    const GatCfg cd = gat_cfg<C>(g->N, kGatStatsThreads);
    if (kPipeOK<C> && cd.pipe)
        launch_k(gat_bwd_dst_kernel<C, kPipeOK<C>, true>, cd.grid, cd.threads, 0, st, g->rowptr, g->col, nullptr, h, s, d, m, z, P, DU, gsd,
                                                                             (int)g->N, slope, cd.chunk_rows, cd.ipc_shift, cd.ahead, f);
    else
        launch_k(gat_bwd_dst_kernel<C, false, true>, cd.grid, cd.threads, 0, st, g->rowptr, g->col, nullptr, h, s, d, m, z, P, DU, gsd,
                                                                          (int)g->N, slope, cd.chunk_rows, cd.ipc_shift, cd.ahead, f);
    const GatCfg c = gat_cfg<C>(g->N);
    BG_GAT_LAUNCH(gat_bwd_src_kernel, g->cscptr, g->cscrow, g->perm, P, DU, f.go_out, a_src, a_dst, gh_tot, gsd, inj_h, (int)g->N);
    return check_launch("bg_gat_bwd_gn");
}
template <int C>
static int launch_bwd2(const BgGraph* g, const float* Ht, const float* St, const float* Dt, const float* gout,
                       const float* h, const float* s, const float* d, const float* m, const float* z,
                       const float* a_src, const float* a_dst, float* scratch, float* gt, float* ht_tot,
                       float* sdt, float slope, cudaStream_t st) {
    const int64_t grid = ceil_div(g->N, RowMap<C>::RPC);
    float *A0 = scratch, *A1 = scratch + g->E, *A2 = scratch + 2 * g->E, *A3 = scratch + 3 * g->E;
    launch_k(gat_bwd2_dst_kernel<C>, (unsigned)grid, kThreads, 0, st, g->rowptr, g->col, Ht, St, Dt, gout, h, s, d, m, z,
                                                               A0, A1, A2, A3, gt, sdt, g->N, slope);
    const GatCfg c = gat_cfg<C>(g->N);
    BG_GAT_LAUNCH(gat_bwd_src_kernel, g->cscptr, g->cscrow, g->perm, A1, A2, gout, a_src, a_dst, ht_tot, sdt, (const float*)nullptr, (int)g->N);
    return check_launch("bg_gat_bwd2");
}

#define BG_DISPATCH_C(C, CALL)                                                         \
    switch (C) {                                                                       \
        case 1: return CALL(1);                                                        \
        case 2: return CALL(2);                                                        \
        case 4: return CALL(4);                                                        \
        case 8: return CALL(8);                                                        \
        case 16: return CALL(16);                                                      \
        case 32: return CALL(32);                                                      \
        case 64: return CALL(64);                                                      \
        case 128: return CALL(128);                                                    \
        default:                                                                       \
            bg::set_error("unsupported channel width C=%d (supported: 1,2,4,...,128)", (int)(C)); \
            return BG_EUNSUPPORTED;                                                    \
    }

static int check_graph(const BgGraph* g) {
    BG_REQUIRE(g && g->rowptr && g->col && g->cscptr && g->cscrow && g->perm, BG_EINVAL, "BgGraph has null arrays");
    BG_REQUIRE(g->N < ((int64_t)1 << 31) - ((int64_t)1 << 27) && g->E < (int64_t)1 << 31, BG_EINVAL, "BgGraph: N and E must fit int32");
    BG_REQUIRE(g->N > 0 && g->E >= g->N, BG_EINVAL, "BgGraph: need N>0 and E>=N (self loops), got N=%lld E=%lld",
               (long long)g->N, (long long)g->E);
    return BG_OK;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_tune(int32_t key, int32_t value) {
    BG_REQUIRE(key >= 0 && key < 8, BG_EINVAL, "bg_tune: unknown key %d", (int)key);
    BG_REQUIRE(key != 0 || (value >= 32 && value <= kGatMaxThreads && value % 32 == 0), BG_EINVAL,
               "bg_tune: threads must be a multiple of 32 in [32, 1024]");
    BG_REQUIRE(key != 1 || value >= 1, BG_EINVAL, "bg_tune: CTAs per SM must be >= 1");
    g_tune[key] = value;
    return BG_OK;
}

extern "C" int bg_gat_fwd(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias,
                          float* out, float* m, float* z, int32_t C, float slope, void* stream) {
    if (int rc = check_graph(g)) return rc;
    BG_REQUIRE(h && s && d && out && m && z, BG_EINVAL, "bg_gat_fwd: null pointer");
    if (gat_tma_enabled() && g->N * C * 4 >= (int64_t)(24 << 20)) {  // HBM-sized graphs: TMA-gather kernel where eligible
        const int rc = gat_fwd_tma_try(g, h, s, d, bias, out, m, z, C, slope, as_stream(stream));
        if (rc <= 0) return rc;
    }
#define CALL(CC) launch_fwd<CC>(g, h, s, d, bias, out, m, z, slope, as_stream(stream))
    BG_DISPATCH_C(C, CALL)
#undef CALL
}

// Upper bound of the grid any launch geometry uses: 8 CTAs per SM (small graphs) or g_tune CTAs per SM.
extern "C" size_t bg_gat_fwd_gn_ws(int64_t N, int32_t C) {
    (void)N;
    const size_t gmax = (size_t)kSMs * (g_tune[1] > 8 ? g_tune[1] : 8) + 16;
    return (size_t)kCounterBytes + (gmax + gmax / kFoldGroup + 2) * 2 * (size_t)C * sizeof(float);
}

extern "C" int bg_gat_fwd_gn(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias, float* out,
                             float* m, float* z, int32_t C, float slope, const float* gn_alpha, float gn_eps, float* gn_stats,
                             float* workspace, size_t ws_bytes, void* stream) {
    if (int rc = check_graph(g)) return rc;
    BG_REQUIRE(h && s && d && out && m && z && gn_alpha && gn_stats && workspace, BG_EINVAL, "bg_gat_fwd_gn: null pointer");
    BG_REQUIRE(ws_bytes >= bg_gat_fwd_gn_ws(g->N, C), BG_EINVAL, "bg_gat_fwd_gn: workspace too small");
    BG_REQUIRE(g_tune[6] == 0 || g_tune[6] <= kSMs * 8, BG_EINVAL, "bg_gat_fwd_gn: grid cap too large");
#define CALL(CC) launch_fwd<CC>(g, h, s, d, bias, out, m, z, slope, as_stream(stream), gn_alpha, gn_eps, gn_stats, workspace)
    BG_DISPATCH_C(C, CALL)
#undef CALL
}

// bg_gat_bwd / bg_gat_bwd_gn with a cotangent injected at h (second-order sweep, bg_passes.cu): gh_tot = (source pass) + inj_h,
// added in the source pass's epilogue instead of by a separate axpy launch on the critic update's critical chain
namespace bg {
int gat_bwd_inj(const BgGraph* g, const float* gout, const float* h, const float* s, const float* d, const float* m, const float* z,
                const float* a_src, const float* a_dst, float* P, float* DU, float* gh_tot, float* gsd, int32_t C, float slope,
                const float* inj_h, void* stream) {
    if (int rc = check_graph(g)) return rc;
    BG_REQUIRE(gout && h && s && d && m && z && a_src && a_dst && P && DU && gh_tot && gsd, BG_EINVAL,
               "bg_gat_bwd: null pointer");
#define CALL(CC) launch_bwd<CC>(g, gout, h, s, d, m, z, a_src, a_dst, P, DU, gh_tot, gsd, slope, as_stream(stream), inj_h)
    BG_DISPATCH_C(C, CALL)
#undef CALL
}
}  // namespace bg

extern "C" int bg_gat_bwd(const BgGraph* g, const float* gout, const float* h, const float* s, const float* d,
                          const float* m, const float* z, const float* a_src, const float* a_dst, float* P,
                          float* DU, float* gh_tot, float* gsd, int32_t C, float slope, void* stream) {
    return bg::gat_bwd_inj(g, gout, h, s, d, m, z, a_src, a_dst, P, DU, gh_tot, gsd, C, slope, nullptr, stream);
}

namespace bg {
int gat_bwd_gn_inj(const BgGraph* g, const float* gx1, const float* o, const float* x1, const float* gn_w, const float* gn_alpha,
                   const float* gn_stats, const float* gn_bstats, float keep_scale, const float* inj_o, const float* h, const float* s,
                   const float* d, const float* m, const float* z, const float* a_src, const float* a_dst, float* P, float* DU,
                   float* go, float* gh_tot, float* gsd, int32_t C, float slope, const float* inj_h, void* stream) {
    if (int rc = check_graph(g)) return rc;
    BG_REQUIRE(gx1 && o && x1 && gn_w && gn_alpha && gn_stats && gn_bstats && h && s && d && m && z && a_src && a_dst && P && DU &&
                   go && gh_tot && gsd,
               BG_EINVAL, "bg_gat_bwd_gn: null pointer");
    const GnBwdFuse f{gx1, o, x1, inj_o, gn_w, gn_alpha, gn_stats, gn_bstats, go, keep_scale};
#define CALL(CC) launch_bwd_gn<CC>(g, f, h, s, d, m, z, a_src, a_dst, P, DU, gh_tot, gsd, slope, as_stream(stream), inj_h)
    BG_DISPATCH_C(C, CALL)
#undef CALL
}
}  // namespace bg

extern "C" int bg_gat_bwd_gn(const BgGraph* g, const float* gx1, const float* o, const float* x1, const float* gn_w,
                             const float* gn_alpha, const float* gn_stats, const float* gn_bstats, float keep_scale,
                             const float* inj_o, const float* h, const float* s, const float* d, const float* m, const float* z,
                             const float* a_src, const float* a_dst, float* P, float* DU, float* go, float* gh_tot, float* gsd,
                             int32_t C, float slope, void* stream) {
    return bg::gat_bwd_gn_inj(g, gx1, o, x1, gn_w, gn_alpha, gn_stats, gn_bstats, keep_scale, inj_o, h, s, d, m, z, a_src, a_dst, P, DU,
                              go, gh_tot, gsd, C, slope, nullptr, stream);
}

extern "C" int bg_gat_bwd2(const BgGraph* g, const float* Ht, const float* St, const float* Dt, const float* gout,
                           const float* h, const float* s, const float* d, const float* m, const float* z,
                           const float* a_src, const float* a_dst, float* scratch, float* gt, float* ht_tot,
                           float* sdt, int32_t C, float slope, void* stream) {
    if (int rc = check_graph(g)) return rc;
    BG_REQUIRE(Ht && St && Dt && gout && h && s && d && m && z && a_src && a_dst && scratch && gt && ht_tot && sdt,
               BG_EINVAL, "bg_gat_bwd2: null pointer");
#define CALL(CC) \
    launch_bwd2<CC>(g, Ht, St, Dt, gout, h, s, d, m, z, a_src, a_dst, scratch, gt, ht_tot, sdt, slope, as_stream(stream))
    BG_DISPATCH_C(C, CALL)
#undef CALL
}
