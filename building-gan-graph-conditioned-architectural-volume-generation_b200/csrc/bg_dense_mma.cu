// Small dense layers of the latency-bound regime on the warp-level tensor-core path: 3xTF32 mma.sync.m16n8k8, fp32-accurate.
//
// Scope: the layers of the training step that are too small for the tcgen05 kernel to amortise its fixed costs (TMEM
// allocation, mbarrier pipeline, swizzled shared-memory operand tiles) and too instruction-heavy for FFMA: plain single-
// segment inputs, K in {8,16,...,128}, Cout in {8,16,32,64}, N <= 131072 - the conv `lin`s of the GNN blocks, the
// discriminator's 64-wide MLP layers and every backward-input product of those (reference models.py:72,82,177-225).
//
// Why.  ncu of the tiled FFMA kernel on the 64->64 layer at N = 15145 (profiles/r02_ncu_dense_small.csv): 4.9 M warp
// instructions for 1.9 M useful FFMAs, issue slots 49 % busy, 13.6 us per launch on a chain where an elementwise pass over the
// same rows costs 3 us; a row-per-thread FFMA variant was worse (20 us: one warp per scheduler, cold straight-line code,
// shared-memory broadcast bandwidth).  The work is instruction-issue bound, so the fix is fewer instructions: one
// m16n8k8 MMA replaces 32 warp-FFMAs.
//
// Mapping.  One warp owns 16 rows; a CTA is 8 warps = 128 rows.
//   A (16 x 8 per k-step) comes straight from global memory into registers: lane (g = lane/4, t = lane%4) loads the float2
//     X[row g | g+8][8 ks + 2t .. 2t+1] - a quad reads 32 contiguous bytes of a row, every 32-byte sector exactly once, all K/8
//     loads of a lane in flight at once (a single L2 round trip).  The two values are used as the fragment's columns t and
//     t+4, i.e. the k index inside a k-step is permuted (logical t <-> physical 2t, logical t+4 <-> physical 2t+1);
//   B (8 x 8 per k-step and n-tile) is staged once per CTA in shared memory IN FRAGMENT ORDER with the same permutation
//     and already split into (hi, lo) TF32 halves: one conflict-free LDS.64 per fragment register;
//   3xTF32: a = a_hi + a_lo (cvt.rna: |a_lo| <= 2^-12 |a|, and rounding a_lo itself to TF32 loses <= 2^-24 |a| - fp32-level),
//     acc_main += a_hi b_hi, acc_corr += a_lo b_hi + a_hi b_lo.  The tensor core accumulates with truncation, one rounding per
//     k-step, biased toward zero: the hi*hi products alternate between TWO main accumulators (even / odd k-steps: <= 8
//     roundings each at K = 128, on half the magnitude) and the correction terms get their own (they are 2^-12 of the result,
//     their roundings vanish); out = (main0 + main1) + corr in fp32 round-to-nearest.
//   Epilogue on the C fragment (rows g, g+8; columns 8 nt + 2t, +1): bias, LayerNorm (row sums over the quad: two
//     shuffles), activation, attention dots, the fused activation-backward gate, saved xhat / rstd, float2 stores (a quad
//     writes one 32-byte sector per row and n-tile), and the optional GraphNorm-backward column moments of the block below
//     (GnMomFuse: shuffles over g, shared memory over the warps, last-CTA fold across CTAs).
#include <stdlib.h>

#include "bg_common.cuh"

namespace bg {

constexpr int MM_WARPS = 8;
constexpr int MM_T = MM_WARPS * 32;
constexpr int MM_ROWS = MM_WARPS * 16;  // rows per CTA
constexpr int64_t MM_MAX_N = 131072;
constexpr int MM_MAX_K = 128;

struct MmParams {
    int64_t N;
    const float* X;
    int64_t ld_x;
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
    GnMomFuse mom;
};

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

template <int COUT, int KSM, bool MOM>  // KSM = compile-time bound on K / 8 (register arrays of the A operand)
__global__ void __launch_bounds__(MM_T) dense_mma_kernel(const MmParams p) {
    pdl_prologue();
    constexpr int NT = COUT / 8;
    extern __shared__ __align__(16) float2 mm_wf[];  // [K/8][NT][2][32] (hi, lo) B fragments
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int K = p.K, KS = K >> 3;
    // column half of a 128-wide plain product (grid.y = 2: the generator's 128 -> 128 backward-input products; LayerNorm /
    // attention dots / moments need the whole row in one warp and never come with grid.y > 1)
    const int col0 = blockIdx.y * COUT;
    const int64_t row_a = (int64_t)blockIdx.x * MM_ROWS + warp * 16 + g, row_b = row_a + 8;
    const bool live_a = row_a < p.N, live_b = row_b < p.N;

    // ---- A: all K/8 float2 pairs of this lane's two rows, issued before anything else (one L2 round trip)
    float2 xa[KSM], xb[KSM];
    {
        const float* pa = p.X + row_a * p.ld_x + 2 * t;
        const float* pb = p.X + row_b * p.ld_x + 2 * t;
#pragma unroll
        for (int ks = 0; ks < KSM; ++ks) {
            xa[ks] = make_float2(0.f, 0.f);
            xb[ks] = make_float2(0.f, 0.f);
            if (ks < KS) {
                if (live_a) xa[ks] = __ldg(reinterpret_cast<const float2*>(pa + 8 * ks));
                if (live_b) xb[ks] = __ldg(reinterpret_cast<const float2*>(pb + 8 * ks));
            }
        }
    }
    // ---- B: fragment-ordered (hi, lo) weights.  Element e = ((ks * NT + nt) * 2 + r) * 32 + l holds
    // Wop[n = 8 nt + l/4][k = 8 ks + 2 (l%4) + r], Wop[n][k] = W[n * w_so + k * w_sk]; K * COUT elements, <= 32 per thread, all loads of a
    // thread in flight before the first shared store.
    {
        const int total = K * COUT;
        for (int base = 0; base < total; base += 16 * MM_T) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int e = base + u * MM_T + tid;
                const int l = e & 31, r = (e >> 5) & 1, q = e >> 6;  // q = ks * NT + nt
                const int nt = q % NT, ks = q / NT;
                const int n = col0 + 8 * nt + (l >> 2), k = 8 * ks + 2 * (l & 3) + r;
                v[u] = (e < total && n < p.Cout) ? __ldg(p.W + (int64_t)n * p.w_so + (int64_t)k * p.w_sk) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int e = base + u * MM_T + tid;
                if (e < total) {
                    const float hi = __uint_as_float(to_tf32(v[u]));
                    mm_wf[e] = make_float2(hi, __uint_as_float(to_tf32(v[u] - hi)));
                }
            }
        }
    }
    __syncthreads();

    float cm[NT][4], cn[NT][4], cc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) cm[nt][i] = cn[nt][i] = cc[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSM; ++ks) {
        if (ks < KS) {
            // fragment registers: a0 = (row g, col t), a1 = (row g+8, col t), a2 = (row g, col t+4), a3 = (row g+8, col t+4)
            const float av[4] = {xa[ks].x, xb[ks].x, xa[ks].y, xb[ks].y};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ah[i] = to_tf32(av[i]);
                al[i] = to_tf32(av[i] - __uint_as_float(ah[i]));
            }
            const float2* wf = mm_wf + (size_t)ks * NT * 64 + lane;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 b0 = wf[nt * 64], b1 = wf[nt * 64 + 32];
                const uint32_t b0h = __float_as_uint(b0.x), b0l = __float_as_uint(b0.y);
                const uint32_t b1h = __float_as_uint(b1.x), b1l = __float_as_uint(b1.y);
                if (ks & 1) mma_tf32(cn[nt], ah, b0h, b1h);
                else mma_tf32(cm[nt], ah, b0h, b1h);
                mma_tf32(cc[nt], al, b0h, b1h);
                mma_tf32(cc[nt], ah, b0l, b1l);
            }
        }
    }

    // ---- epilogue on the C fragment: y[nt][0..1] = row_a cols 8nt + 2t, +1 ; y[nt][2..3] = row_b same cols
    float y[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int c = col0 + 8 * nt + 2 * t;
        float b0 = 0.f, b1 = 0.f;
        if (p.bias) {
            const float2 bv = __ldg(reinterpret_cast<const float2*>(p.bias + c));
            b0 = bv.x;
            b1 = bv.y;
        }
        y[nt][0] = (cm[nt][0] + cn[nt][0]) + cc[nt][0] + b0;
        y[nt][1] = (cm[nt][1] + cn[nt][1]) + cc[nt][1] + b1;
        y[nt][2] = (cm[nt][2] + cn[nt][2]) + cc[nt][2] + b0;
        y[nt][3] = (cm[nt][3] + cn[nt][3]) + cc[nt][3] + b1;
    }
    if (p.gamma) {  // LayerNorm over the COUT columns of each row (eps = 1e-5): row sums over the quad
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            sa += y[nt][0] + y[nt][1];
            sb += y[nt][2] + y[nt][3];
        }
        const float ma = quad_sum(sa) / (float)COUT, mb = quad_sum(sb) / (float)COUT;
        float va = 0.f, vb = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            va = fmaf(y[nt][0] - ma, y[nt][0] - ma, va);
            va = fmaf(y[nt][1] - ma, y[nt][1] - ma, va);
            vb = fmaf(y[nt][2] - mb, y[nt][2] - mb, vb);
            vb = fmaf(y[nt][3] - mb, y[nt][3] - mb, vb);
        }
        const float ra = 1.f / sqrtf(quad_sum(va) / (float)COUT + 1e-5f), rb = 1.f / sqrtf(quad_sum(vb) / (float)COUT + 1e-5f);
        if (p.rstd && t == 0) {
            if (live_a) p.rstd[row_a] = ra;
            if (live_b) p.rstd[row_b] = rb;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = 8 * nt + 2 * t;
            const float2 gv = __ldg(reinterpret_cast<const float2*>(p.gamma + c)), bv = __ldg(reinterpret_cast<const float2*>(p.beta + c));
            const float h0 = (y[nt][0] - ma) * ra, h1 = (y[nt][1] - ma) * ra, h2 = (y[nt][2] - mb) * rb, h3 = (y[nt][3] - mb) * rb;
            if (p.xhat) {
                if (live_a) *reinterpret_cast<float2*>(p.xhat + row_a * COUT + c) = make_float2(h0, h1);
                if (live_b) *reinterpret_cast<float2*>(p.xhat + row_b * COUT + c) = make_float2(h2, h3);
            }
            y[nt][0] = fmaf(h0, gv.x, bv.x);
            y[nt][1] = fmaf(h1, gv.y, bv.y);
            y[nt][2] = fmaf(h2, gv.x, bv.x);
            y[nt][3] = fmaf(h3, gv.y, bv.y);
        }
    }
    if (p.act != BG_ACT_NONE) {
        const float slope = p.act == BG_ACT_RELU ? 0.f : 0.2f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) y[nt][i] = y[nt][i] > 0.f ? y[nt][i] : slope * y[nt][i];
    }
    if (p.att_src) {
        float s_a = 0.f, d_a = 0.f, s_b = 0.f, d_b = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = 8 * nt + 2 * t;
            const float2 as = __ldg(reinterpret_cast<const float2*>(p.att_src + c)), ad = __ldg(reinterpret_cast<const float2*>(p.att_dst + c));
            s_a = fmaf(y[nt][0], as.x, fmaf(y[nt][1], as.y, s_a));
            d_a = fmaf(y[nt][0], ad.x, fmaf(y[nt][1], ad.y, d_a));
            s_b = fmaf(y[nt][2], as.x, fmaf(y[nt][3], as.y, s_b));
            d_b = fmaf(y[nt][2], ad.x, fmaf(y[nt][3], ad.y, d_b));
        }
        s_a = quad_sum(s_a), d_a = quad_sum(d_a), s_b = quad_sum(s_b), d_b = quad_sum(d_b);
        if (t == 0) {
            if (live_a) { p.s[row_a] = s_a; p.d[row_a] = d_a; }
            if (live_b) { p.s[row_b] = s_b; p.d[row_b] = d_b; }
        }
    }
    if (p.gate) {  // fused activation backward of the layer below
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = col0 + 8 * nt + 2 * t;
            if (live_a) {
                const float2 gv = __ldg(reinterpret_cast<const float2*>(p.gate + row_a * p.ld_gate + c));
                y[nt][0] *= gv.x > 0.f ? 1.f : p.gate_slope;
                y[nt][1] *= gv.y > 0.f ? 1.f : p.gate_slope;
            }
            if (live_b) {
                const float2 gv = __ldg(reinterpret_cast<const float2*>(p.gate + row_b * p.ld_gate + c));
                y[nt][2] *= gv.x > 0.f ? 1.f : p.gate_slope;
                y[nt][3] *= gv.y > 0.f ? 1.f : p.gate_slope;
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int c = col0 + 8 * nt + 2 * t;
        if (live_a) *reinterpret_cast<float2*>(p.out + row_a * p.ld_out + c) = make_float2(y[nt][0], y[nt][1]);
        if (live_b) *reinterpret_cast<float2*>(p.out + row_b * p.ld_out + c) = make_float2(y[nt][2], y[nt][3]);
    }

    if constexpr (MOM) {
        // y = gx1 of rows a / b: column sums of gy = gx1 * keep_scale * [x1 > 0] and gy * (o - alpha mu): over g by shuffles
        // (fixed order), over the warps through shared memory (warp order), over the CTAs by the last-CTA fold; the last CTA
        // finishes exactly like gn_bwd_moments_kernel.
        __shared__ float wpart[MM_WARPS][2 * COUT];
        __shared__ float fred[kThreads];
        __shared__ float fsum[2 * COUT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = 8 * nt + 2 * t;
            const float2 am = make_float2(__ldg(p.mom.alpha + c) * __ldg(p.mom.stats + c), __ldg(p.mom.alpha + c + 1) * __ldg(p.mom.stats + c + 1));
            float m0[2] = {0.f, 0.f}, m1[2] = {0.f, 0.f};
            if (live_a) {
                const float2 xv = __ldg(reinterpret_cast<const float2*>(p.mom.x1 + row_a * COUT + c));
                const float2 ov = __ldg(reinterpret_cast<const float2*>(p.mom.o + row_a * COUT + c));
                const float g0 = xv.x > 0.f ? y[nt][0] * p.mom.keep_scale : 0.f, g1 = xv.y > 0.f ? y[nt][1] * p.mom.keep_scale : 0.f;
                m0[0] += g0; m0[1] += g1;
                m1[0] = fmaf(g0, ov.x - am.x, m1[0]); m1[1] = fmaf(g1, ov.y - am.y, m1[1]);
            }
            if (live_b) {
                const float2 xv = __ldg(reinterpret_cast<const float2*>(p.mom.x1 + row_b * COUT + c));
                const float2 ov = __ldg(reinterpret_cast<const float2*>(p.mom.o + row_b * COUT + c));
                const float g0 = xv.x > 0.f ? y[nt][2] * p.mom.keep_scale : 0.f, g1 = xv.y > 0.f ? y[nt][3] * p.mom.keep_scale : 0.f;
                m0[0] += g0; m0[1] += g1;
                m1[0] = fmaf(g0, ov.x - am.x, m1[0]); m1[1] = fmaf(g1, ov.y - am.y, m1[1]);
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                m0[0] += __shfl_xor_sync(0xffffffffu, m0[0], o); m0[1] += __shfl_xor_sync(0xffffffffu, m0[1], o);
                m1[0] += __shfl_xor_sync(0xffffffffu, m1[0], o); m1[1] += __shfl_xor_sync(0xffffffffu, m1[1], o);
            }
            if (g == 0) {
                wpart[warp][c] = m0[0]; wpart[warp][c + 1] = m0[1];
                wpart[warp][COUT + c] = m1[0]; wpart[warp][COUT + c + 1] = m1[1];
            }
        }
        __syncthreads();
        float* partials = p.mom.partials;
        for (int i = tid; i < 2 * COUT; i += MM_T) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < MM_WARPS; ++w) s += wpart[w][i];
            partials[(int64_t)blockIdx.x * 2 * COUT + i] = s;
        }
        if (!hier_fold(partials, partials + (int64_t)gridDim.x * 2 * COUT, 2 * COUT, p.mom.counters, fred, fsum)) return;
        if (tid < COUT) {
            const int c = tid, C = COUT;
            const float n = (float)p.N;
            const float G0 = fsum[c] / n, G1 = fsum[COUT + c] / n;
            const float mu = p.mom.stats[c], r = p.mom.stats[C + c], a = p.mom.alpha[c], wc = p.mom.w[c];
            const float mean_ohat = wc * r * G0 - wc * r * r * r * G1 * mu * (1.f - a);  // M[d loss/d ohat]
            p.mom.bstats[c] = G0;
            p.mom.bstats[C + c] = G1;
            const float dw = n * r * G1, db = n * G0, da = -mu * n * mean_ohat;
            float* dp = p.mom.dparams;
            if (p.mom.accumulate) {
                dp[c] += dw; dp[C + c] += db; dp[2 * C + c] += da;
            } else {
                dp[c] = dw; dp[C + c] = db; dp[2 * C + c] = da;
            }
        }
    }
}

static const bool g_dense_mma128 = !(getenv("BG_DENSE_MMA128") && atoi(getenv("BG_DENSE_MMA128")) == 0);  // A/B switch of the column-half mode
static int g_dense_mma = -1;  // BG_DENSE_MMA=0: off (A/B switch)
static int dense_mma_on() {
    if (g_dense_mma < 0) g_dense_mma = getenv("BG_DENSE_MMA") ? atoi(getenv("BG_DENSE_MMA")) : 1;
    return g_dense_mma;
}

template <int COUT, int KSM, bool MOM>
static void mm_launch2(const MmParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    static bool once = (cudaFuncSetAttribute(dense_mma_kernel<COUT, KSM, MOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), true);
    (void)once;
    launch_k(dense_mma_kernel<COUT, KSM, MOM>, grid, MM_T, smem, st, p);
}
template <int COUT>
static void mm_launch(const MmParams& p, bool mom, dim3 grid, size_t smem, cudaStream_t st) {
    const int ks = p.K >> 3;
    if (mom) {
        if (ks <= 4) mm_launch2<COUT, 4, true>(p, grid, smem, st);
        else if (ks <= 8) mm_launch2<COUT, 8, true>(p, grid, smem, st);
        else mm_launch2<COUT, 16, true>(p, grid, smem, st);
    } else {
        if (ks <= 4) mm_launch2<COUT, 4, false>(p, grid, smem, st);
        else if (ks <= 8) mm_launch2<COUT, 8, false>(p, grid, smem, st);
        else mm_launch2<COUT, 16, false>(p, grid, smem, st);
    }
}

// BG_OK when launched, 1 when the shape is not eligible, < 0 on error.
int dense_mma_try(const BgDense* a, int K, const GnMomFuse* mom, cudaStream_t st) {
    if (!dense_mma_on()) return 1;
    const int C = a->Cout;
    // 128 columns: two column halves per row tile (grid.y = 2), plain epilogue only (bias / activation / gate)
    const bool halves = C == 128 && !a->ln_gamma && !a->att_src && !a->xhat && !mom && g_dense_mma128;
    if (a->N > MM_MAX_N || K > MM_MAX_K || (K & 7) || !(C == 8 || C == 16 || C == 32 || C == 64 || halves)) return 1;
    if (a->nseg != 1 || !a->seg[0].ptr || a->seg[0].gather || a->seg[0].width != K) return 1;
    const auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7) == 0; };
    if ((a->seg[0].ld & 1) || !al8(a->seg[0].ptr) || (a->ld_out & 1) || !al8(a->out)) return 1;
    if (a->gate && ((a->ld_gate & 1) || !al8(a->gate))) return 1;
    if ((a->bias && !al8(a->bias)) || (a->ln_gamma && (!al8(a->ln_gamma) || !al8(a->ln_beta))) ||
        (a->att_src && (!al8(a->att_src) || !al8(a->att_dst))) || (a->xhat && !al8(a->xhat)))
        return 1;
    if (mom && (!al8(mom->o) || !al8(mom->x1))) return 1;
    MmParams p;
    p.N = a->N; p.X = a->seg[0].ptr; p.ld_x = a->seg[0].ld; p.K = K;
    p.W = a->W; p.w_so = a->w_so; p.w_sk = a->w_sk; p.Cout = C;
    p.bias = a->bias; p.gamma = a->ln_gamma; p.beta = a->ln_beta;
    p.att_src = a->att_src; p.att_dst = a->att_dst; p.act = a->act;
    p.out = a->out; p.ld_out = a->ld_out; p.xhat = a->xhat; p.rstd = a->rstd; p.s = a->s; p.d = a->d;
    p.gate = a->gate; p.ld_gate = a->ld_gate; p.gate_slope = a->gate_slope;
    p.mom = mom ? *mom : GnMomFuse{};
    const int Ct = halves ? 64 : C;  // columns per CTA
    const size_t smem = (size_t)K * Ct * sizeof(float2);
    const dim3 grid((unsigned)ceil_div(a->N, MM_ROWS), halves ? 2u : 1u);
    switch (Ct) {
        case 8: mm_launch<8>(p, mom != nullptr, grid, smem, st); break;
        case 16: mm_launch<16>(p, mom != nullptr, grid, smem, st); break;
        case 32: mm_launch<32>(p, mom != nullptr, grid, smem, st); break;
        default: mm_launch<64>(p, mom != nullptr, grid, smem, st); break;
    }
    return check_launch(mom ? "dense_fwd_moments(mma)" : "bg_dense_fwd(mma)");
}

}  // namespace bg

// 1 = warp-MMA 3xTF32 kernel for the small dense layers (default), 0 = off.  Returns the previous setting.
extern "C" int bg_set_dense_mma(int32_t on) {
    const int prev = bg::dense_mma_on();
    bg::g_dense_mma = on ? 1 : 0;
    return prev;
}
