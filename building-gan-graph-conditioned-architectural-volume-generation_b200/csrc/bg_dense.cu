// Node-wise dense layers: y = act(LayerNorm(X W^T + b)) over a SEGMENTED input (the reference's
// torch.cat at models.py:135-141,146,239 is never materialised), optional fused attention dots
// s = y.a_src, d = y.a_dst (the `lin` of a GATConv), the split-N deterministic weight gradient
// and the LayerNorm/activation backward.  fp32 FFMA: this is the rel-1e-5 parity mode.
#include <algorithm>
#include <type_traits>
#include <mutex>
#include <vector>

#include <stdlib.h>

#include "bg_common.cuh"

namespace bg {

constexpr int BM = 64;  // rows per CTA
constexpr int BK = 16;  // k-slab

struct SegView {
    int nseg;
    int off[BG_MAX_SEG + 1];
    BgSeg seg[BG_MAX_SEG];
};

__device__ __forceinline__ float seg_fetch(const SegView& sv, int64_t row, int k) {
#pragma unroll
    for (int q = 0; q < BG_MAX_SEG; ++q) {
        if (q < sv.nseg && k < sv.off[q + 1]) {
            const BgSeg& sg = sv.seg[q];
            if (sg.ptr == nullptr) return 1.f;
            const int64_t r = sg.gather ? (int64_t)__ldg(sg.gather + row) : row;
            return __ldg(sg.ptr + r * sg.ld + (k - sv.off[q]));
        }
    }
    return 0.f;
}

static int fill_segview(SegView& sv, int nseg, const BgSeg* seg, int* K);

// Column k of the segmented input resolved once (the column a thread fetches is loop-invariant): afterwards a
// fetch is one (or, with a gather index, two dependent) loads without any segment search.
struct SegCol {
    const float* base;      // segment pointer + column offset; nullptr with ones => constant 1, without => constant 0
    const int32_t* gather;
    int ld;
    bool ones;
};
__device__ __forceinline__ SegCol seg_resolve(const SegView& sv, int k, int K) {
    SegCol c{nullptr, nullptr, 0, false};
    if (k >= K) return c;
#pragma unroll
    for (int q = 0; q < BG_MAX_SEG; ++q) {
        if (q < sv.nseg && k >= sv.off[q] && k < sv.off[q + 1]) {
            const BgSeg& sg = sv.seg[q];
            c.ones = (sg.ptr == nullptr);
            c.base = sg.ptr ? sg.ptr + (k - sv.off[q]) : nullptr;
            c.gather = sg.gather;
            c.ld = sg.ld;
        }
    }
    return c;
}
__device__ __forceinline__ float seg_load(const SegCol& c, int64_t row) {
    if (c.base == nullptr) return c.ones ? 1.f : 0.f;
    const int64_t r = c.gather ? (int64_t)__ldg(c.gather + row) : row;
    return __ldg(c.base + r * c.ld);
}

struct DenseParams {
    int64_t N;
    SegView x;
    int K;
    const float* W;
    int64_t w_so, w_sk;
    int Cout;
    const float *bias, *gamma, *beta, *att_src, *att_dst;
    int act;
    float* out;
    int64_t ld_out;
    float *xhat, *rstd, *s, *d;
    const float* gate;
    int64_t ld_gate;
    float gate_slope;
    GnMomFuse mom;  // only read by the MOM instantiations
};

template <int BN, int TM, int TN, bool MOM = false>
__global__ void __launch_bounds__(kThreads) dense_fwd_kernel(const DenseParams p) {
    pdl_prologue();
    constexpr int TX = BN / TN, TY = BM / TM;
    static_assert(TX * TY == kThreads, "tile/thread mismatch");
    __shared__ float Xs[BK][BM + 1];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;  // column tile (only without LayerNorm / attention dots)
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // Software pipeline over the K slabs: the global loads of slab k+1 are issued (into registers) before slab k is
    // multiplied, so their L2 latency overlaps the FMAs instead of being paid once per slab - at N ~ 15 k rows these
    // launches are latency-bound and K = 64 means four dependent round trips otherwise.
    constexpr int XPER = BM * BK / kThreads, WPER = (BN * BK + kThreads - 1) / kThreads;
    float xv[XPER], wv[WPER];
    auto fetch = [&](int k0) {
        const SegCol xcol = seg_resolve(p.x, k0 + tid % BK, p.K);  // this thread's column of the slab (256 % BK == 0)
#pragma unroll
        for (int q = 0; q < XPER; ++q) {
            const int idx = tid + q * kThreads, r = idx / BK;
            const int64_t grow = row0 + r;
            xv[q] = (grow < p.N) ? seg_load(xcol, grow) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < WPER; ++q) {
            const int idx = tid + q * kThreads;
            int c, kk;
            if (p.w_sk == 1) { c = idx / BK; kk = idx % BK; } else { kk = idx / BN; c = idx % BN; }
            const int gc = col0 + c, gk = k0 + kk;
            wv[q] = (idx < BN * BK && gc < p.Cout && gk < p.K) ? __ldg(p.W + gc * p.w_so + gk * p.w_sk) : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
        for (int q = 0; q < XPER; ++q) {
            const int idx = tid + q * kThreads;
            Xs[idx % BK][idx / BK] = xv[q];
        }
#pragma unroll
        for (int q = 0; q < WPER; ++q) {
            const int idx = tid + q * kThreads;
            if (idx < BN * BK) {
                int c, kk;
                if (p.w_sk == 1) { c = idx / BK; kk = idx % BK; } else { kk = idx / BN; c = idx % BN; }
                Ws[kk][c] = wv[q];
            }
        }
        __syncthreads();
        if (k0 + BK < p.K) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float xr[TM], wr[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) xr[i] = Xs[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) wr[j] = Ws[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, LayerNorm, activation, attention dots
    float ms0[MOM ? TN : 1], ms1[MOM ? TN : 1], am[MOM ? TN : 1];  // MOM: column sums of gy and gy * (o - alpha mu)
    if constexpr (MOM) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int c = tx * TN + j;
            ms0[j] = ms1[j] = 0.f;
            am[j] = c < p.Cout ? __ldg(p.mom.alpha + c) * __ldg(p.mom.stats + c) : 0.f;
        }
    }
    float bj[TN], gj[TN], tj[TN], asj[TN], adj[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int c = col0 + tx * TN + j;
        const bool ok = c < p.Cout;
        bj[j] = (ok && p.bias) ? __ldg(p.bias + c) : 0.f;
        gj[j] = (ok && p.gamma) ? __ldg(p.gamma + c) : 1.f;
        tj[j] = (ok && p.beta) ? __ldg(p.beta + c) : 0.f;
        asj[j] = (ok && p.att_src) ? __ldg(p.att_src + c) : 0.f;
        adj[j] = (ok && p.att_dst) ? __ldg(p.att_dst + c) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t grow = row0 + ty * TM + i;
        float y[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) y[j] = acc[i][j] + bj[j];
        if (p.gamma) {  // LayerNorm over the BN (== Cout) columns of this row, eps = 1e-5
            float sm = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) sm += y[j];
            sm = group_sum<TX>(sm);
            const float mean = sm / (float)BN;
            float vs = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const float dlt = y[j] - mean;
                vs = fmaf(dlt, dlt, vs);
            }
            vs = group_sum<TX>(vs);
            const float rs = 1.f / sqrtf(vs / (float)BN + 1e-5f);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const float xh = (y[j] - mean) * rs;
                if (p.xhat && grow < p.N) p.xhat[grow * p.Cout + tx * TN + j] = xh;
                y[j] = fmaf(xh, gj[j], tj[j]);
            }
            if (p.rstd && tx == 0 && grow < p.N) p.rstd[grow] = rs;
        }
        if (p.act == BG_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < TN; ++j) y[j] = y[j] > 0.f ? y[j] : 0.f;
        } else if (p.act == BG_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < TN; ++j) y[j] = y[j] > 0.f ? y[j] : 0.2f * y[j];
        }
        if (p.att_src) {
            float ss = 0.f, dd = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                ss = fmaf(y[j], asj[j], ss);
                dd = fmaf(y[j], adj[j], dd);
            }
            ss = group_sum<TX>(ss);
            dd = group_sum<TX>(dd);
            if (tx == 0 && grow < p.N) {
                p.s[grow] = ss;
                p.d[grow] = dd;
            }
        }
        if (p.gate && grow < p.N) {  // fused activation backward of the layer below: zero / scale where its output was off
            const float* grw = p.gate + grow * p.ld_gate + col0 + tx * TN;
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (col0 + tx * TN + j < p.Cout) y[j] *= (__ldg(grw + j) > 0.f ? 1.f : p.gate_slope);
        }
        if constexpr (MOM) {
            if (grow < p.N) {  // y = this row's gx1 entries (after the fused gate): GraphNorm backward moments of the block below
                const float* orw = p.mom.o + grow * p.Cout + tx * TN;
                const float* xrw = p.mom.x1 + grow * p.Cout + tx * TN;
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (tx * TN + j < p.Cout) {
                        const float gy = __ldg(xrw + j) > 0.f ? y[j] * p.mom.keep_scale : 0.f;
                        ms0[j] += gy;
                        ms1[j] = fmaf(gy, __ldg(orw + j) - am[j], ms1[j]);
                    }
            }
        }
        if (grow < p.N) {
            float* orow = p.out + grow * p.ld_out + col0 + tx * TN;
            if (TN % 4 == 0 && (p.ld_out & 3) == 0 && col0 + tx * TN + TN <= p.Cout &&
                ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < TN; j += 4)
                    *reinterpret_cast<float4*>(orow + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (col0 + tx * TN + j < p.Cout) orow[j] = y[j];
            }
        }
    }
    if constexpr (MOM) {
        // CTA column sums in a fixed order: rows of a warp by shuffles (lanes with equal tx), warps through shared memory,
        // CTAs through the last-CTA fold; the last CTA finishes exactly like gn_bwd_moments_kernel.
        constexpr int NW = kThreads / 32;
        __shared__ float wpart[NW][2 * BN];
        __shared__ float fred[kThreads];
        __shared__ float fsum[2 * BN];
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
#pragma unroll
            for (int o = TX; o < 32; o <<= 1) {
                ms0[j] += __shfl_xor_sync(0xffffffffu, ms0[j], o);
                ms1[j] += __shfl_xor_sync(0xffffffffu, ms1[j], o);
            }
        }
        if (lane < TX) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                wpart[warp][lane * TN + j] = ms0[j];
                wpart[warp][BN + lane * TN + j] = ms1[j];
            }
        }
        __syncthreads();
        float* partials = p.mom.partials;
        for (int i = tid; i < 2 * BN; i += kThreads) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += wpart[w][i];
            partials[(int64_t)blockIdx.x * 2 * BN + i] = t;
        }
        if (!hier_fold(partials, partials + (int64_t)gridDim.x * 2 * BN, 2 * BN, p.mom.counters, fred, fsum)) return;
        if (tid < p.Cout) {
            const int c = tid, C = p.Cout;
            const float n = (float)p.N;
            const float G0 = fsum[c] / n, G1 = fsum[BN + c] / n;
            const float mu = p.mom.stats[c], r = p.mom.stats[C + c], a = p.mom.alpha[c], wc = p.mom.w[c];
            const float mean_ohat = wc * r * G0 - wc * r * r * r * G1 * mu * (1.f - a);  // M[d loss/d ohat]
            p.mom.bstats[c] = G0;
            p.mom.bstats[C + c] = G1;
            const float dw = n * r * G1, db = n * G0, da = -mu * n * mean_ohat;
            float* dp = p.mom.dparams;
            if (p.mom.accumulate) {
                dp[c] += dw;
                dp[C + c] += db;
                dp[2 * C + c] += da;
            } else {
                dp[c] = dw;
                dp[C + c] = db;
                dp[2 * C + c] = da;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// weight gradients: dW[o,k] = sum_n gz[n,o] X[n,k].  Up to BG_MAX_WGRAD independent problems per launch
// (a conv block's lin.weight, [att_src;att_dst] and bias reductions share one launch).  grid =
// (64x64 output tiles of all problems, S row splits).  Every CTA reduces its row range for its tile and
// writes a compact partial [S][Cout*K]; a fold kernel sums the S partials of every queued problem in split
// order (deterministic, no float atomics).  The whole-pass executors queue the folds and run ONE fold
// launch per pass (bg_passes.cu); the stand-alone entry point folds immediately.
// ------------------------------------------------------------------------------------------
struct WgProblem {
    const float* gz;
    int64_t ld_gz, N, rows_per_split;  // every problem has its own row count (the 7-row type table next to N voxels)
    int Cout, K;
    SegView x;
    float* partial;  // [S][Cout*K]
    int tiles_k, tile_base, ntiles;
};
constexpr int kWgMax = 16;  // problems per launch (keeps the parameter block under 4 KiB)
constexpr int kWgKick = 8;  // deferred mode: a batch leaves for the side stream as soon as this many problems are queued
constexpr double kWgKickWork = 1.2e8;  // ... or as soon as the queued multiply-adds reach this (N x Cout x K; one 128 x 128 layer at N = 8 k)
struct WgBatch {
    int nprob, S;
    WgProblem p[kWgMax];
};

__global__ void __launch_bounds__(kThreads) wgrad_multi_kernel(const WgBatch b) {
    pdl_prologue();
    constexpr int T = 64, RB = 32;
    __shared__ __align__(16) float smem[2][RB][T + 4];
    float(*Gs)[T + 4] = smem[0];
    float(*Xs)[T + 4] = smem[1];
    int pi = 0;
#pragma unroll 1
    for (int q = 1; q < b.nprob; ++q)
        if ((int)blockIdx.x >= b.p[q].tile_base) pi = q;
    const WgProblem& p = b.p[pi];
    const int lt = blockIdx.x - p.tile_base;
    const int k0 = (lt % p.tiles_k) * T, o0 = (lt / p.tiles_k) * T;
    const int tid = threadIdx.x;
    const int64_t rbeg = (int64_t)blockIdx.y * p.rows_per_split, rend = min(p.N, rbeg + p.rows_per_split);
    const int no = min(T, p.Cout - o0), nk = min(T, p.K - k0);  // valid extent of this tile
    // Thread mapping.  A full 64x64 tile is a 16x16 grid of 4x4 micro-tiles = 256 threads.  Most tiles of this model
    // are much smaller (C <= 64 in the discriminator, [2,C] attention and [C,1] bias gradients): their to x tk micro-
    // tiles occupy only `tiles` threads, so the CTA is split into R = 256 / tiles row groups that each take every
    // R-th row of a slab and are summed in group order at the end (fixed order => deterministic).  r01f before this:
    // 68 us per batched launch with 7/8 of the warps idle at the barriers.
    const int to = (no + 3) >> 2, tk = (nk + 3) >> 2, tiles = to * tk;
    int R = 1;
    while (R < RB && 2 * R * tiles <= kThreads) R <<= 1;
    const int grp = tid / tiles, slot = tid - grp * tiles;
    const bool active = grp < R;
    const int ty = slot / tk, tx = slot - ty * tk;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // every thread always fetches column tid % 64 of both operands (256 % 64 == 0): resolve it once
    const int mycol = tid % T;
    const SegCol xcol = seg_resolve(p.x, k0 + mycol, mycol < nk ? p.K : 0);
    const float* gcol = mycol < no ? p.gz + o0 + mycol : nullptr;
    // software pipeline: the global loads of slab i+1 are issued (into registers) before slab i is multiplied, so the
    // load latency overlaps the FMAs / the barrier instead of adding to every iteration
    constexpr int PER = RB * T / kThreads, RSTEP = kThreads / T;  // a thread's q-th element of a slab: row (tid / T) + q * RSTEP
    float gv[PER], xv[PER];
    // Address arithmetic hoisted out of the slab loop (ncu r01i: FFMA was 17% of the instructions of this kernel, IMAD / ISETP /
    // LEA / branches of the per-element index math and of seg_load 50%): per thread one base pointer per operand and constant
    // strides; full slabs (all but possibly the last of a split) carry no row predicates; a gather index or a partial slab
    // takes the general path.
    const int rl = tid / T;
    const float* gp = gcol ? gcol + (rbeg + rl) * p.ld_gz : nullptr;
    const int64_t gstep = (int64_t)RSTEP * p.ld_gz;
    const bool xdirect = xcol.base != nullptr && xcol.gather == nullptr;
    const float* xp = xdirect ? xcol.base + (rbeg + rl) * (int64_t)xcol.ld : nullptr;
    const int64_t xstep = (int64_t)RSTEP * xcol.ld;
    const float xconst = (xcol.base == nullptr && xcol.ones) ? 1.f : 0.f;
    auto fetch = [&](int64_t r0) {
        if (r0 + RB <= rend) {
            if (gp) {
                const float* g0 = gp + (r0 - rbeg) * p.ld_gz;
#pragma unroll
                for (int q = 0; q < PER; ++q) gv[q] = __ldg(g0 + q * gstep);
            } else {
#pragma unroll
                for (int q = 0; q < PER; ++q) gv[q] = 0.f;
            }
            if (xdirect) {
                const float* x0 = xp + (r0 - rbeg) * (int64_t)xcol.ld;
#pragma unroll
                for (int q = 0; q < PER; ++q) xv[q] = __ldg(x0 + q * xstep);
            } else if (xcol.base == nullptr) {
#pragma unroll
                for (int q = 0; q < PER; ++q) xv[q] = xconst;
            } else {
#pragma unroll
                for (int q = 0; q < PER; ++q) xv[q] = seg_load(xcol, r0 + rl + q * RSTEP);
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int64_t r = r0 + (tid + q * kThreads) / T;
            gv[q] = (r < rend && gcol) ? __ldg(gcol + r * p.ld_gz) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int64_t r = r0 + (tid + q * kThreads) / T;
            xv[q] = (r < rend) ? seg_load(xcol, r) : 0.f;
        }
    };
    if (rbeg < rend) fetch(rbeg);
    for (int64_t r0 = rbeg; r0 < rend; r0 += RB) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            Gs[rl + q * RSTEP][mycol] = gv[q];
            Xs[rl + q * RSTEP][mycol] = xv[q];
        }
        __syncthreads();
        if (r0 + RB < rend) fetch(r0 + RB);
        if (active) {
            // rows grp, grp + R, ... of the slab, fully unrolled per R (same order as a plain loop): constant shared-memory
            // offsets and one 128-bit load per operand instead of a dynamic row loop with eight scalar loads
            auto rows = [&](auto rc) {
                constexpr int RR = decltype(rc)::value;
#pragma unroll
                for (int i = 0; i < RB / RR; ++i) {
                    const float4 g4 = *reinterpret_cast<const float4*>(&Gs[grp + i * RR][ty * 4]);
                    const float4 x4 = *reinterpret_cast<const float4*>(&Xs[grp + i * RR][tx * 4]);
                    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(g[a], x[c], acc[a][c]);
                }
            };
            switch (R) {
                case 1: rows(std::integral_constant<int, 1>{}); break;
                case 2: rows(std::integral_constant<int, 2>{}); break;
                case 4: rows(std::integral_constant<int, 4>{}); break;
                case 8: rows(std::integral_constant<int, 8>{}); break;
                case 16: rows(std::integral_constant<int, 16>{}); break;
                default: rows(std::integral_constant<int, 32>{}); break;
            }
        }
        __syncthreads();
    }
    if (R > 1) {  // fold the row groups through shared memory (Gs and Xs are free now: 2*32*68 >= 256*16 floats)
        float* red = &smem[0][0][0];
        static_assert(sizeof(smem) >= kThreads * 16 * sizeof(float), "fold buffer");
        if (active) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) red[(i * 4 + j) * kThreads + tid] = acc[i][j];
        }
        __syncthreads();
        if (grp == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float t = 0.f;
                    for (int q = 0; q < R; ++q) t += red[(i * 4 + j) * kThreads + q * tiles + slot];
                    acc[i][j] = t;
                }
        }
    }
    if (grp != 0) return;
    float* mine = p.partial + (int64_t)blockIdx.y * p.Cout * p.K;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + ty * 4 + i, k = k0 + tx * 4 + j;
            if (o < p.Cout && k < p.K) mine[(int64_t)o * p.K + k] = acc[i][j];
        }
}

// The same partial sums on the warp-level tensor cores (3xTF32 mma.sync.m16n8k8, fp32-accurate): D[o, k] = sum_rows gz[row, o] X[row, k]
// is an M = Cout, N = K, "K" = rows product.  Same tiles, same splits, same slab loader (coalesced, segment / gather aware, software
// pipelined) and the same partial layout as wgrad_multi_kernel - only the inner product differs: the slab's operands are read
// from shared memory as MMA fragments (A[m][kk] = Gs[kk][m], B[kk][n] = Xs[kk][n]; row stride 72 floats: conflict-free), split
// into TF32 hi / lo on the fly, hi*hi into a main accumulator, the two correction products into a second one.  Warp w of the 8
// owns the 16 x 32 sub-tile (m-tile w % MT', k-half ...) of the 64 x 64 tile; tiles with fewer than 8 sub-tiles split the slab's
// four 8-row k-steps over R row groups of warps, folded in group order at the end (deterministic).  Why: the FFMA kernel
// was the second-largest consumer of SM time in the training step (5.5 of 22.7 ms of kernel time per step, round 2), bound
// by instruction issue; one MMA replaces 32 warp-FFMAs.
__global__ void __launch_bounds__(kThreads) wgrad_mma_kernel(const WgBatch b) {
    pdl_prologue();
    constexpr int T = 64, RB = 32, LD = T + 8;
    __shared__ __align__(16) float smem[2][RB][LD];
    float(*Gs)[LD] = smem[0];
    float(*Xs)[LD] = smem[1];
    int pi = 0;
#pragma unroll 1
    for (int q = 1; q < b.nprob; ++q)
        if ((int)blockIdx.x >= b.p[q].tile_base) pi = q;
    const WgProblem& p = b.p[pi];
    const int lt = blockIdx.x - p.tile_base;
    const int k0 = (lt % p.tiles_k) * T, o0 = (lt / p.tiles_k) * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int64_t rbeg = (int64_t)blockIdx.y * p.rows_per_split, rend = min(p.N, rbeg + p.rows_per_split);
    const int no = min(T, p.Cout - o0), nk = min(T, p.K - k0);  // valid extent of this tile
    // warp -> (sub-tile, row group): MT m-tiles of 16 outputs x NG groups of 32 k columns, units rounded up to a power of two
    const int MT = (no + 15) >> 4, NG = (nk + 31) >> 5, units = MT * NG;
    int U = 1;
    while (U < units) U <<= 1;            // 1, 2, 4, 8
    const int R = 8 / U;                   // row groups: k-steps rg, rg + R, ... of every slab (R in {1, 2, 4, 8}; 8 > 4 k-steps: some idle)
    const int unit = warp % U, rg = warp / U;
    const bool active = unit < units;
    const int m0 = (unit % MT) * 16, n0 = (unit / MT) * 32;
    float cm[4][4], cc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) cm[j][i] = cc[j][i] = 0.f;
    // ---- slab loader: identical to wgrad_multi_kernel
    const int mycol = tid % T;
    const SegCol xcol = seg_resolve(p.x, k0 + mycol, mycol < nk ? p.K : 0);
    const float* gcol = mycol < no ? p.gz + o0 + mycol : nullptr;
    constexpr int PER = RB * T / kThreads, RSTEP = kThreads / T;
    float gv[PER], xv[PER];
    const int rl = tid / T;
    const float* gp = gcol ? gcol + (rbeg + rl) * p.ld_gz : nullptr;
    const int64_t gstep = (int64_t)RSTEP * p.ld_gz;
    const bool xdirect = xcol.base != nullptr && xcol.gather == nullptr;
    const float* xp = xdirect ? xcol.base + (rbeg + rl) * (int64_t)xcol.ld : nullptr;
    const int64_t xstep = (int64_t)RSTEP * xcol.ld;
    const float xconst = (xcol.base == nullptr && xcol.ones) ? 1.f : 0.f;
    auto fetch = [&](int64_t r0) {
        if (r0 + RB <= rend) {
            if (gp) {
                const float* g0 = gp + (r0 - rbeg) * p.ld_gz;
#pragma unroll
                for (int q = 0; q < PER; ++q) gv[q] = __ldg(g0 + q * gstep);
            } else {
#pragma unroll
                for (int q = 0; q < PER; ++q) gv[q] = 0.f;
            }
            if (xdirect) {
                const float* x0 = xp + (r0 - rbeg) * (int64_t)xcol.ld;
#pragma unroll
                for (int q = 0; q < PER; ++q) xv[q] = __ldg(x0 + q * xstep);
            } else if (xcol.base == nullptr) {
#pragma unroll
                for (int q = 0; q < PER; ++q) xv[q] = xconst;
            } else {
#pragma unroll
                for (int q = 0; q < PER; ++q) xv[q] = seg_load(xcol, r0 + rl + q * RSTEP);
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int64_t r = r0 + (tid + q * kThreads) / T;
            gv[q] = (r < rend && gcol) ? __ldg(gcol + r * p.ld_gz) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int64_t r = r0 + (tid + q * kThreads) / T;
            xv[q] = (r < rend) ? seg_load(xcol, r) : 0.f;
        }
    };
    auto split = [](float v, uint32_t& hi, uint32_t& lo) {
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(v - __uint_as_float(hi)));
    };
    auto mma = [](float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    };
    if (rbeg < rend) fetch(rbeg);
    for (int64_t r0 = rbeg; r0 < rend; r0 += RB) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            Gs[rl + q * RSTEP][mycol] = gv[q];
            Xs[rl + q * RSTEP][mycol] = xv[q];
        }
        __syncthreads();
        if (r0 + RB < rend) fetch(r0 + RB);
        if (active) {
            for (int ks = rg; ks < RB / 8; ks += R) {  // the slab's 8-row k-steps of this row group, ascending
                const int ra = 8 * ks + t, rb = ra + 4;
                uint32_t ah[4], al[4];
                split(Gs[ra][m0 + g], ah[0], al[0]);
                split(Gs[ra][m0 + g + 8], ah[1], al[1]);
                split(Gs[rb][m0 + g], ah[2], al[2]);
                split(Gs[rb][m0 + g + 8], ah[3], al[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t b0h, b0l, b1h, b1l;
                    split(Xs[ra][n0 + 8 * j + g], b0h, b0l);
                    split(Xs[rb][n0 + 8 * j + g], b1h, b1l);
                    mma(cm[j], ah, b0h, b1h);
                    mma(cc[j], al, b0h, b1h);
                    mma(cc[j], ah, b0l, b1l);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) cm[j][i] += cc[j][i];
    if (R > 1) {  // fold the row groups through shared memory in group order (Gs / Xs are free now)
        float* red = &smem[0][0][0];
        static_assert(sizeof(smem) >= kThreads * 16 * sizeof(float), "fold buffer");
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) red[(j * 4 + i) * kThreads + tid] = active ? cm[j][i] : 0.f;
        __syncthreads();
        if (rg == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float s = 0.f;
                    for (int q = 0; q < R; ++q) s += red[(j * 4 + i) * kThreads + (q * U + unit) * 32 + lane];
                    cm[j][i] = s;
                }
        }
    }
    if (rg != 0 || !active) return;
    float* mine = p.partial + (int64_t)blockIdx.y * p.Cout * p.K;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // C fragment: (row g | g + 8, column 2t | 2t + 1) of n-tile j
            const int o = o0 + m0 + g + ((i & 2) ? 8 : 0), k = k0 + n0 + 8 * j + 2 * t + (i & 1);
            if (o < p.Cout && k < p.K) mine[(int64_t)o * p.K + k] = cm[j][i];
        }
}

__global__ void __launch_bounds__(kThreads) wgrad_fold_kernel(const FoldBatch fb) {
    pdl_prologue();
    int e = 0;
#pragma unroll 1
    for (int q = 1; q < fb.n; ++q)
        if ((int)blockIdx.x >= fb.e[q].block_base) e = q;
    const FoldEntry& f = fb.e[e];
    const int count = f.Cout * f.K;
    const int i = ((int)blockIdx.x - f.block_base) * kThreads + threadIdx.x;
    if (i >= count) return;
    float t = 0.f;
    const float* src = f.partial + i;
    int sp = 0;
    for (; sp + 8 <= f.S; sp += 8) {  // 8 independent loads in flight, summed in split order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (int64_t)(sp + u) * count);
#pragma unroll
        for (int u = 0; u < 8; ++u) t += v[u];
    }
    for (; sp < f.S; ++sp) t += __ldg(src + (int64_t)sp * count);
    const int o = i / f.K, k = i % f.K;
    float* dst = (f.dbias && k == f.K - 1) ? f.dbias + o : f.dW + (int64_t)o * f.ld_dw + k;
    *dst = f.accumulate ? *dst + t : t;
}

static inline int wgrad_splits(int64_t N, int total_tiles, int64_t max_ck = 0) {
    // small outputs (a handful of tiles) are pure latency: use many short row ranges (the folds of a whole pass run as
    // one parallel launch, so more partials are cheap); large outputs keep the partial traffic bounded
    int64_t ns = ceil_div(4 * kSMs, total_tiles);
    const int64_t maxs = ceil_div(N, 64);
    // a handful of SMALL outputs (the last batch of a pass: its launch sits between the end of the dgrad chain and the fold):
    // 128 short row ranges; wide outputs (a 128 x 128 layer alone is 4 tiles too) keep the partial-sum traffic at 64
    const int64_t cap = (total_tiles <= 2 || (total_tiles <= 4 && max_ck <= 64 * 65)) ? 128 : 64;
    if (ns > maxs) ns = maxs;
    if (ns > cap) ns = cap;
    if (ns < 1) ns = 1;
    return (int)ns;
}
static int wgrad_tiles(int Cout, int K) { return (int)(ceil_div(K, 64) * ceil_div(Cout, 64)); }

// One partial-sum launch for up to kWgMax problems (any mix of row counts); queues their folds.
static int wgrad_launch_batch(const BgWgrad* probs, int nprob, WgradQueue& q, cudaStream_t st) {
    WgBatch b;
    b.nprob = nprob;
    int tiles = 0;
    int64_t nmax = 0;
    for (int i = 0; i < nprob; ++i) {
        const BgWgrad& a = probs[i];
        BG_REQUIRE(a.gz && (a.dW || a.dbias), BG_EINVAL, "wgrad: problem %d has null pointers", i);
        WgProblem& p = b.p[i];
        p.gz = a.gz; p.ld_gz = a.ld_gz; p.Cout = a.Cout; p.N = a.N;
        if (int rc = fill_segview(p.x, a.nseg, a.seg, &p.K)) return rc;
        p.tiles_k = (int)ceil_div(p.K, 64);
        p.tile_base = tiles;
        p.ntiles = wgrad_tiles(p.Cout, p.K);
        tiles += p.ntiles;
        if (a.N > nmax) nmax = a.N;
    }
    int64_t max_ck = 0;
    for (int i = 0; i < nprob; ++i) max_ck = std::max<int64_t>(max_ck, (int64_t)b.p[i].Cout * b.p[i].K);
    b.S = wgrad_splits(nmax, tiles, max_ck);
    {   // the partial sums of every batch of a pass live side by side until the fold: shorten the splits rather than fail
        size_t per_split = 0;
        for (int i = 0; i < nprob; ++i) per_split += ((size_t)b.p[i].Cout * b.p[i].K + 63) / 64 * 64 + 64;
        const size_t room = q.cap > q.used ? q.cap - q.used : 0;
        if ((size_t)b.S * per_split > room && room / per_split >= 1) b.S = (int)(room / per_split);
    }
    for (int i = 0; i < nprob; ++i) {
        WgProblem& p = b.p[i];
        p.rows_per_split = ceil_div(ceil_div(p.N, b.S), 32) * 32;
        const size_t need = (size_t)b.S * p.Cout * p.K;
        BG_REQUIRE(q.used + need <= q.cap, BG_EINVAL, "wgrad: partial-sum workspace too small (%zu + %zu > %zu floats)", q.used,
                   need, q.cap);
        p.partial = q.buf + q.used;
        q.used += (need + 63) / 64 * 64;
        FoldEntry f{p.partial, probs[i].dW, probs[i].dbias, probs[i].ld_dw, b.S, p.Cout, p.K, probs[i].accumulate, 0};
        // a contribution whose destination overlaps an already queued one goes to a later fold launch (the folds of
        // one launch run concurrently: no two may read-modify-write the same addresses)
        auto overlaps = [](const FoldEntry& a, const FoldEntry& b) {
            auto hit = [](const float* p0, size_t n0, const float* p1, size_t n1) {
                return p0 && p1 && p0 < p1 + n1 && p1 < p0 + n0;
            };
            const size_t aw = a.dW ? (size_t)(a.Cout - 1) * a.ld_dw + a.K : 0, bw = b.dW ? (size_t)(b.Cout - 1) * b.ld_dw + b.K : 0;
            return hit(a.dW, aw, b.dW, bw) || hit(a.dW, aw, b.dbias, b.Cout) || hit(a.dbias, a.Cout, b.dW, bw) ||
                   hit(a.dbias, a.Cout, b.dbias, b.Cout);
        };
        int phase = 0;
        for (int ph = 0; ph < WgradQueue::kPhases; ++ph)
            for (const FoldEntry& g : q.phase[ph])
                if (ph >= phase && overlaps(f, g)) phase = ph + 1;
        BG_REQUIRE(phase < WgradQueue::kPhases, BG_EUNSUPPORTED, "wgrad: too many contributions to one destination");
        if (phase > 0) f.accumulate = 1;
        q.phase[phase].push_back(f);
    }
    dim3 grid((unsigned)tiles, (unsigned)b.S);
    static const bool skip = getenv("BG_DEBUG_SKIP_WGRAD") && atoi(getenv("BG_DEBUG_SKIP_WGRAD"));  // timing experiments only: WRONG gradients
    static const bool use_mma = !(getenv("BG_WGRAD_MMA") && atoi(getenv("BG_WGRAD_MMA")) == 0);  // default: tensor-core partial sums
    if (!skip) {
        if (use_mma) launch_k(wgrad_mma_kernel, grid, kThreads, 0, st, b);
        else launch_k(wgrad_multi_kernel, grid, kThreads, 0, st, b);
    }
    return check_launch("wgrad");
}

// Immediate mode: launch now.  Deferred mode (q.defer, used by the whole-pass executors, whose operands stay alive
// until the end of the pass): only record the problems; wgrad_flush() then runs ALL weight gradients of the pass as
// one launch per kWgMax problems - ~25 tiny latency-bound launches per backward pass become 2.
// Side stream + fork/join events of the overlap mode: one set per (device, caller stream) - passes running concurrently on
// different caller streams (step.py's discriminator lanes) must not share a side stream, or each one's join would wait for the
// other's weight gradients.  Created on first use, never destroyed (the only library-owned CUDA objects; at most kMaxSide sets,
// further caller streams stay on their own stream; BG_WGRAD_OVERLAP=0 keeps everything on the caller's stream).
struct SideStream {
    int dev = -1;
    cudaStream_t caller = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream* side_stream(cudaStream_t caller) {
    static const bool enabled = !(getenv("BG_WGRAD_OVERLAP") && atoi(getenv("BG_WGRAD_OVERLAP")) == 0);
    if (!enabled) return nullptr;
    constexpr int kMaxSide = 64;
    static SideStream pool[kMaxSide];
    static int used = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < used; ++i)
        if (pool[i].dev == dev && pool[i].caller == caller) return &pool[i];
    if (used == kMaxSide) return nullptr;
    SideStream& s = pool[used];
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming);
    s.dev = dev;
    s.caller = caller;
    ++used;
    return &s;
}

// Launch the first `n` pending problems: on the side stream (after everything enqueued so far on `st`) when available
static int wgrad_kick(WgradQueue& q, size_t n, cudaStream_t st) {
    cudaStream_t where = st;
    if (!q.side) {
        if (SideStream* s = side_stream(st)) { q.side = s->st; q.ev_fork = s->fork; q.ev_join = s->join; }
    }
    if (q.side) {
        if (cudaEventRecord(q.ev_fork, st) != cudaSuccess || cudaStreamWaitEvent(q.side, q.ev_fork, 0) != cudaSuccess) {
            set_error("wgrad: fork to the side stream failed: %s", cudaGetErrorString(cudaGetLastError()));
            return BG_ECUDA;
        }
        where = q.side;
        q.forked = true;
    }
    const int rc = wgrad_launch_batch(q.pending.data(), (int)n, q, where);
    q.pending.erase(q.pending.begin(), q.pending.begin() + n);
    return rc;
}

int wgrad_launch(const BgWgrad* probs, int nprob, WgradQueue& q, cudaStream_t st) {
    BG_REQUIRE(nprob >= 1, BG_EINVAL, "wgrad: nprob %d out of range", nprob);
    if (q.defer) {
        q.pending.insert(q.pending.end(), probs, probs + nprob);
        while (q.pending.size() >= (size_t)kWgKick)
            if (int rc = wgrad_kick(q, kWgKick, st)) return rc;
        // a LARGE problem does not wait for seven more: the generator's 128-wide layers are ~25 us of side-stream work each,
        // and batched by count the last five of a backward pass left together AFTER the dgrad chain had ended (141 us batch,
        // ~100 us of it exposed at the end of the step: profiles/r02b_summary.md); kicked as they arrive only the last one is
        double work = 0.0;
        for (const BgWgrad& a : q.pending) {
            int64_t k = 0;
            for (int i = 0; i < a.nseg; ++i) k += a.seg[i].width;
            work += (double)a.N * (double)a.Cout * (double)k;
        }
        if (work >= kWgKickWork && !q.pending.empty())
            if (int rc = wgrad_kick(q, q.pending.size(), st)) return rc;
        return BG_OK;
    }
    for (int i = 0; i < nprob; i += kWgMax)
        if (int rc = wgrad_launch_batch(probs + i, nprob - i < kWgMax ? nprob - i : kWgMax, q, st)) return rc;
    return BG_OK;
}

// Fold every queued problem: one launch per phase (and per kMaxFold entries).
int wgrad_flush(WgradQueue& q, cudaStream_t st) {
    while (!q.pending.empty())
        if (int rc = wgrad_kick(q, q.pending.size() < (size_t)kWgMax ? q.pending.size() : (size_t)kWgMax, st)) return rc;
    if (q.forked) {  // join: the folds (and everything the caller enqueues next) wait for the side-stream batches
        if (cudaEventRecord(q.ev_join, q.side) != cudaSuccess || cudaStreamWaitEvent(st, q.ev_join, 0) != cudaSuccess) {
            set_error("wgrad: join from the side stream failed: %s", cudaGetErrorString(cudaGetLastError()));
            return BG_ECUDA;
        }
        q.forked = false;
    }
    for (int ph = 0; ph < WgradQueue::kPhases; ++ph) {
        std::vector<FoldEntry>& v = q.phase[ph];
        for (size_t base = 0; base < v.size(); base += FoldBatch::kMax) {
            FoldBatch fb;
            fb.n = (int)std::min<size_t>(FoldBatch::kMax, v.size() - base);
            int blocks = 0;
            for (int i = 0; i < fb.n; ++i) {
                fb.e[i] = v[base + i];
                fb.e[i].block_base = blocks;
                blocks += (int)ceil_div((int64_t)fb.e[i].Cout * fb.e[i].K, kThreads);
            }
            launch_k(wgrad_fold_kernel, blocks, kThreads, 0, st, fb);
        }
        v.clear();
    }
    q.used = 0;
    return check_launch("wgrad fold");
}

// ------------------------------------------------------------------------------------------
// LayerNorm + activation backward (row-wise) with deterministic column sums for dgamma/dbeta
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) ln_act_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                                              const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, int64_t N, int G, int act,
                                                              float* __restrict__ gz, float* dgamma, float* dbeta,
                                                              int accumulate, unsigned int* counter, float* partials) {
    pdl_prologue();
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES, RPC = M::RPC;
    __shared__ float red[kThreads * 2 * VEC];
    __shared__ float sums[2 * C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES, slot = warp * M::RPW + lane / LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const int64_t chunk = ceil_div(N, G);
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(N, r0 + chunk);
    float gam[VEC], dg[VEC], db[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        gam[v] = __ldg(gamma + sub * VEC + v);
        dg[v] = 0.f;
        db[v] = 0.f;
    }
    // four rows of a lane group in flight per trip (a warp walks ~13 rows of a 128-wide layer at N = 15 k: one row at a time
    // was 13 dependent load -> shuffle-reduce -> store chains, 18-25 us per launch on the generator's backward chain); the
    // per-thread dgamma / dbeta sums still add the rows in ascending order: same bits as the one-row loop
    constexpr int U = 4;
    for (int64_t r = r0 + slot; r < r1; r += U * RPC) {
        Vec<VEC> g[U], o[U], xh[U];
        float rs[U];
        bool ok[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const int64_t rq = r + (int64_t)q * RPC;
            ok[q] = rq < r1;
            if (ok[q]) {
                const int64_t off = rq * C + sub * VEC;
                g[q].load(gout + off);
                o[q].load(out + off);
                xh[q].load(xhat + off);
                rs[q] = __ldg(rstd + rq);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) g[q].v[v] = o[q].v[v] = xh[q].v[v] = 0.f;
                rs[q] = 0.f;
            }
        }
        float gx[U][VEC], s1[U], s2[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            s1[q] = s2[q] = 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float gy = g[q].v[v];
                if (act == BG_ACT_LRELU) gy = o[q].v[v] > 0.f ? gy : 0.2f * gy;
                else if (act == BG_ACT_RELU) gy = o[q].v[v] > 0.f ? gy : 0.f;
                if (ok[q]) {
                    dg[v] = fmaf(gy, xh[q].v[v], dg[v]);
                    db[v] += gy;
                }
                gx[q][v] = gy * gam[v];
                s1[q] += gx[q][v];
                s2[q] = fmaf(gx[q][v], xh[q].v[v], s2[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            s1[q] = gsum<LANES>(s1[q], gm) / (float)C;
            s2[q] = gsum<LANES>(s2[q], gm) / (float)C;
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (ok[q]) {
                Vec<VEC> z;
#pragma unroll
                for (int v = 0; v < VEC; ++v) z.v[v] = rs[q] * (gx[q][v] - s1[q] - xh[q].v[v] * s2[q]);
                z.store(gz + (r + (int64_t)q * RPC) * C + sub * VEC);
            }
        }
    }
    // column sums over the CTA's row slots: fixed order
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        red[(slot * LANES + sub) * 2 * VEC + v] = dg[v];
        red[(slot * LANES + sub) * 2 * VEC + VEC + v] = db[v];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kThreads) {
        const int which = i / C, c = i % C, sb = c / VEC, v = c % VEC;
        float t = 0.f;
        for (int sl = 0; sl < RPC; ++sl) t += red[(sl * LANES + sb) * 2 * VEC + which * VEC + v];
        partials[(int64_t)blockIdx.x * 2 * C + i] = t;
    }
    __syncthreads();
    if (!hier_fold(partials, partials + (int64_t)G * 2 * C, 2 * C, counter, red, sums)) return;
    for (int c = threadIdx.x; c < C; c += kThreads) {
        if (accumulate) {
            dgamma[c] += sums[c];
            dbeta[c] += sums[C + c];
        } else {
            dgamma[c] = sums[c];
            dbeta[c] = sums[C + c];
        }
    }
}

__global__ void __launch_bounds__(kThreads) act_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                                           int64_t total, int act, float* __restrict__ gz) {
    pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const float o = out[i], g = gout[i];
        gz[i] = act == BG_ACT_RELU ? (o > 0.f ? g : 0.f) : act == BG_ACT_LRELU ? (o > 0.f ? g : 0.2f * g) : g;
    }
}

// ------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, int64_t n) {
    pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        y[i] = fmaf(a, x[i], y[i]);
}
__global__ void __launch_bounds__(kThreads) fill_kernel(float* __restrict__ y, float v, int64_t n) {
    pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) y[i] = v;
}

// Adam over one flat parameter buffer (H15: torch.optim.Adam semantics, amsgrad/maximize off): one launch per optimiser
// step instead of torch's foreach sequence over ~190 parameter tensors.  Same update order as torch's _multi_tensor_adam:
// m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).  The step count comes
// either from the host (step_size / bc2_sqrt precomputed in double, like torch) or - CUDA-graph replay - from a device
// counter that the caller increments before the launch.
__global__ void __launch_bounds__(kThreads) adam_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                             float4* __restrict__ v, int64_t n4, double lr, double b1, double b2d, float b2,
                                                             float w1, float w2, float eps, float wd, float step_size, float bc2_sqrt, const int64_t* __restrict__ step_dev) {
    pdl_prologue();
    if (step_dev) {
        const double t = (double)*step_dev;
        step_size = (float)(lr / (1.0 - pow(b1, t)));
        bc2_sqrt = (float)sqrt(1.0 - pow(b2d, t));
    }
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = wd != 0.f ? fmaf(wd, P[k], G[k]) : G[k];
            const float diff = gr - M[k];  // torch lerp: a + w d for |w| < 0.5, else b - d (1 - w)
            M[k] = w1 < 0.5f ? M[k] + w1 * diff : gr - diff * (1.f - w1);
            V[k] = V[k] * b2 + w2 * gr * gr;
            const float denom = sqrtf(V[k]) / bc2_sqrt + eps;
            P[k] = P[k] - step_size * (M[k] / denom);
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

static inline int flat_grid(int64_t total) {
    int64_t g = ceil_div(total, (int64_t)kThreads * 4);
    if (g < 1) g = 1;
    if (g > 8 * kSMs) g = 8 * kSMs;
    return (int)g;
}

static int fill_segview(SegView& sv, int nseg, const BgSeg* seg, int* K) {
    BG_REQUIRE(nseg >= 1 && nseg <= BG_MAX_SEG, BG_EINVAL, "segment count %d out of range [1,%d]", nseg, BG_MAX_SEG);
    sv.nseg = nseg;
    sv.off[0] = 0;
    for (int q = 0; q < BG_MAX_SEG; ++q) {
        if (q < nseg) {
            BG_REQUIRE(seg[q].width > 0, BG_EINVAL, "segment %d has width %d", q, seg[q].width);
            sv.seg[q] = seg[q];
            sv.off[q + 1] = sv.off[q] + seg[q].width;
        } else {
            sv.seg[q] = BgSeg{nullptr, nullptr, 0, 0};
            sv.off[q + 1] = sv.off[q];
        }
    }
    *K = sv.off[nseg];
    return BG_OK;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_dense_fwd(const BgDense* a, void* stream) { return dense_fwd_moments(a, nullptr, as_stream(stream)); }

int bg::dense_fwd_moments(const BgDense* a, const GnMomFuse* mom, cudaStream_t stream) {
    BG_REQUIRE(a && a->W && a->out, BG_EINVAL, "bg_dense_fwd: null pointer");
    BG_REQUIRE(a->N > 0 && a->Cout > 0, BG_EINVAL, "bg_dense_fwd: N and Cout must be positive");
    DenseParams p;
    p.N = a->N;
    if (int rc = fill_segview(p.x, a->nseg, a->seg, &p.K)) return rc;
    p.W = a->W; p.w_so = a->w_so; p.w_sk = a->w_sk; p.Cout = a->Cout;
    p.bias = a->bias; p.gamma = a->ln_gamma; p.beta = a->ln_beta;
    p.att_src = a->att_src; p.att_dst = a->att_dst; p.act = a->act;
    p.out = a->out; p.ld_out = a->ld_out; p.xhat = a->xhat; p.rstd = a->rstd; p.s = a->s; p.d = a->d;
    p.gate = a->gate; p.ld_gate = a->ld_gate; p.gate_slope = a->gate_slope;
    const bool rowwise = a->ln_gamma || a->att_src;
    BG_REQUIRE(!a->ln_gamma || a->ln_beta, BG_EINVAL, "bg_dense_fwd: LayerNorm needs gamma and beta");
    BG_REQUIRE(!a->att_src || (a->att_dst && a->s && a->d), BG_EINVAL, "bg_dense_fwd: attention dots need att_dst, s, d");
    BG_REQUIRE(!rowwise || a->Cout <= 128, BG_EUNSUPPORTED, "bg_dense_fwd: row-wise epilogue needs Cout<=128 (got %d)", a->Cout);
    {  // 128/64-wide layers with plain row-major weights go to the tensor cores (tcgen05, 3xTF32 split: fp32-accurate)
        const int rc = (a->gate || mom) ? 1 : dense_tc_try(a, p.K, stream);
        if (rc <= 0) return rc;
    }
    if (mom) {
        BG_REQUIRE(!rowwise && a->Cout <= 128 && a->ld_out == a->Cout, BG_EINVAL, "dense_fwd_moments: plain product with Cout <= 128 only");
        BG_REQUIRE(mom->o && mom->x1 && mom->alpha && mom->stats && mom->w && mom->dparams && mom->bstats && mom->counters && mom->partials,
                   BG_EINVAL, "dense_fwd_moments: null pointer");
    }
    {  // small layers in the latency-bound regime: warp-MMA 3xTF32 (bg_dense_mma.cu), very narrow ones row-per-thread (bg_rowdense.cu)
        int rc = dense_mma_try(a, p.K, mom, stream);
        if (rc <= 0) return rc;
        rc = rowdense_try(a, p.x.off, p.K, mom, stream);
        if (rc <= 0) return rc;
    }
    int bn = 8;
    while (bn < a->Cout && bn < 128) bn <<= 1;
    BG_REQUIRE(!a->ln_gamma || bn == a->Cout, BG_EUNSUPPORTED,
               "bg_dense_fwd: LayerNorm width must be one of 8,16,32,64,128 (got %d)", a->Cout);
    dim3 grid((unsigned)ceil_div(a->N, BM), (unsigned)ceil_div(a->Cout, bn));
    cudaStream_t st = stream;
    if (mom) {
        BG_REQUIRE(!rowwise && a->Cout <= 128 && a->ld_out == a->Cout, BG_EINVAL, "dense_fwd_moments: plain product with Cout <= 128 only");
        BG_REQUIRE(mom->o && mom->x1 && mom->alpha && mom->stats && mom->w && mom->dparams && mom->bstats && mom->counters && mom->partials,
                   BG_EINVAL, "dense_fwd_moments: null pointer");
        p.mom = *mom;
        switch (bn) {
            case 8: launch_k(dense_fwd_kernel<8, 2, 1, true>, grid, kThreads, 0, st, p); break;
            case 16: launch_k(dense_fwd_kernel<16, 2, 2, true>, grid, kThreads, 0, st, p); break;
            case 32: launch_k(dense_fwd_kernel<32, 2, 4, true>, grid, kThreads, 0, st, p); break;
            case 64: launch_k(dense_fwd_kernel<64, 4, 4, true>, grid, kThreads, 0, st, p); break;
            default: launch_k(dense_fwd_kernel<128, 4, 8, true>, grid, kThreads, 0, st, p); break;
        }
        return check_launch("dense_fwd_moments");
    }
    p.mom = GnMomFuse{};
    switch (bn) {
        case 8: launch_k(dense_fwd_kernel<8, 2, 1>, grid, kThreads, 0, st, p); break;
        case 16: launch_k(dense_fwd_kernel<16, 2, 2>, grid, kThreads, 0, st, p); break;
        case 32: launch_k(dense_fwd_kernel<32, 2, 4>, grid, kThreads, 0, st, p); break;
        case 64: launch_k(dense_fwd_kernel<64, 4, 4>, grid, kThreads, 0, st, p); break;
        default: launch_k(dense_fwd_kernel<128, 4, 8>, grid, kThreads, 0, st, p); break;
    }
    return check_launch("bg_dense_fwd");
}

extern "C" size_t bg_wgrad_multi_ws(int64_t N, int32_t nprob, const int32_t* Cout, const int32_t* K) {
    int tiles = 0;
    size_t elems = 0;
    for (int i = 0; i < nprob; ++i) {
        tiles += wgrad_tiles(Cout[i], K[i]);
        elems += ((size_t)Cout[i] * K[i] + 63) / 64 * 64;
    }
    return (size_t)kCounterBytes + elems * (size_t)wgrad_splits(N, tiles) * sizeof(float) + 4096;
}

extern "C" int bg_wgrad_multi(const BgWgrad* probs, int32_t nprob, float* workspace, size_t ws_bytes, void* stream) {
    BG_REQUIRE(probs && workspace, BG_EINVAL, "bg_wgrad_multi: null pointer");
    BG_REQUIRE(nprob >= 1 && nprob <= BG_MAX_WGRAD, BG_EINVAL, "bg_wgrad_multi: nprob %d out of range [1,%d]", (int)nprob, BG_MAX_WGRAD);
    BG_REQUIRE(ws_bytes > (size_t)kCounterBytes, BG_EINVAL, "bg_wgrad_multi: workspace too small");
    WgradQueue q;
    q.buf = workspace + kCounterBytes / sizeof(float);
    q.cap = (ws_bytes - kCounterBytes) / sizeof(float);
    q.used = 0;
    if (int rc = wgrad_launch(probs, nprob, q, as_stream(stream))) return rc;
    return wgrad_flush(q, as_stream(stream));
}

extern "C" size_t bg_dense_wgrad_ws(int64_t N, int32_t Cout, int32_t K) { return bg_wgrad_multi_ws(N, 1, &Cout, &K); }

extern "C" int bg_dense_wgrad(const BgWgrad* a, void* stream) {
    BG_REQUIRE(a && a->workspace, BG_EINVAL, "bg_dense_wgrad: null pointer");
    return bg_wgrad_multi(a, 1, a->workspace, a->ws_bytes, stream);
}

extern "C" size_t bg_ln_act_bwd_ws(int64_t N, int32_t C) {
    (void)N;
    return (size_t)kCounterBytes + (size_t)(2 * kSMs + 2 * kSMs / kFoldGroup + 2) * 2 * (size_t)C * sizeof(float);
}

extern "C" int bg_ln_act_bwd(const float* gout, const float* out, const float* xhat, const float* rstd, const float* gamma,
                             int64_t N, int32_t C, int32_t act, float* gz, float* dgamma, float* dbeta, int32_t accumulate,
                             float* workspace, size_t ws_bytes, void* stream) {
    BG_REQUIRE(gout && out && gz, BG_EINVAL, "bg_ln_act_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
    if (!xhat) {
        const int64_t total = N * C;
        launch_k(act_bwd_kernel, flat_grid(total), kThreads, 0, st, gout, out, total, act, gz);
        return check_launch("bg_ln_act_bwd(act)");
    }
    BG_REQUIRE(rstd && gamma && dgamma && dbeta && workspace, BG_EINVAL, "bg_ln_act_bwd: LayerNorm path needs rstd/gamma/dgamma/dbeta/workspace");
    BG_REQUIRE(ws_bytes >= bg_ln_act_bwd_ws(N, C), BG_EINVAL, "bg_ln_act_bwd: workspace too small");
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    float* partials = workspace + kCounterBytes / sizeof(float);
#define CALL(CC)                                                                                                   \
    {                                                                                                              \
        const int G = reduce_splits(N, RowMap<CC>::RPC);                                                           \
        launch_k(ln_act_bwd_kernel<CC>, G, kThreads, 0, st, gout, out, xhat, rstd, gamma, N, G, act, gz, dgamma, dbeta, \
                                                      accumulate, counter, partials);                              \
    }
    switch (C) {
        case 8: CALL(8); break;
        case 16: CALL(16); break;
        case 32: CALL(32); break;
        case 64: CALL(64); break;
        case 128: CALL(128); break;
        default:
            bg::set_error("bg_ln_act_bwd: unsupported LayerNorm width %d (8,16,32,64,128)", C);
            return BG_EUNSUPPORTED;
    }
#undef CALL
    return check_launch("bg_ln_act_bwd");
}

extern "C" int bg_axpy(float* y, const float* x, float a, int64_t n, void* stream) {
    BG_REQUIRE(y && x, BG_EINVAL, "bg_axpy: null pointer");
    if (n <= 0) return BG_OK;
    launch_k(axpy_kernel, flat_grid(n), kThreads, 0, as_stream(stream), y, x, a, n);
    return check_launch("bg_axpy");
}

extern "C" int bg_fill(float* y, float v, int64_t n, void* stream) {
    BG_REQUIRE(y, BG_EINVAL, "bg_fill: null pointer");
    if (n <= 0) return BG_OK;
    launch_k(fill_kernel, flat_grid(n), kThreads, 0, as_stream(stream), y, v, n);
    return check_launch("bg_fill");
}

extern "C" int bg_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                            double weight_decay, int64_t step, const int64_t* step_dev, void* stream) {
    BG_REQUIRE(p && g && m && v, BG_EINVAL, "bg_adam_flat: null pointer");
    BG_REQUIRE(n >= 0 && n % 4 == 0, BG_EINVAL, "bg_adam_flat: n=%lld must be a multiple of 4 (flat buckets are padded)", (long long)n);
    BG_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, BG_EINVAL, "bg_adam_flat: buffers must be 16-byte aligned");
    BG_REQUIRE(step_dev || step >= 1, BG_EINVAL, "bg_adam_flat: step=%lld (counts from 1)", (long long)step);
    if (n == 0) return BG_OK;
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (!step_dev) {
        step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
        bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    }
    const int64_t grid = std::min<int64_t>(ceil_div(n / 4, (int64_t)kThreads), 8 * kSMs);
    launch_k(adam_flat_kernel, (int)grid, kThreads, 0, as_stream(stream), reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
             reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n / 4, lr, beta1, beta2, (float)beta2,
             (float)(1.0 - beta1), (float)(1.0 - beta2),  // torch forms 1 - beta in double, then narrows
             (float)eps, (float)weight_decay, step_size, bc2_sqrt, step_dev);
    return check_launch("bg_adam_flat");
}
