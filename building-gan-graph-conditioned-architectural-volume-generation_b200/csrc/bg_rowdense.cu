// Row-per-thread dense layer for the LATENCY-BOUND regime of the training step (N ~ 15 k rows, K <= 128, Cout <= 64:
// the conv `lin`s of the narrow GNN blocks, the discriminator's MLPs, every backward-input product of those layers -
// ~600 of the ~700 dense launches of one step, reference models.py:72,82,177-225).
//
// Why a second kernel.  At this size the whole problem lives in L2 and a launch is a dependency-chain link, not a
// throughput problem: measured on the chain (profiles/tools/chain_latency.py, round 2) the tiled 64x64 FFMA kernel of
// bg_dense.cu costs 12-16 us per 64-wide layer (four K-slab round trips through shared memory with two barriers each,
// scalar shared loads at one per two FMAs, a cross-lane LayerNorm / attention-dot epilogue), against ~3 us for an
// elementwise pass over the same rows.  Here ONE thread owns ONE row:
//   * the CTA stages its 128 input rows into shared memory with ONE wave of coalesced loads (every thread has K scalar
//     loads in flight at once: a single L2 round trip for the whole tile, segments / row gathers resolved once per thread
//     because a thread always fetches the same column), and the whole weight matrix W[K, Cout] next to it;
//   * the K loop then reads one x value (conflict-free: odd row stride) and Cout/4 broadcast float4 weights per k and
//     issues Cout independent FMAs: no barrier inside the loop, accumulation order = k ascending (deterministic);
//   * bias, LayerNorm, activation, attention dots, the fused activation-backward gate and the saved xhat / rstd are
//     thread-local arithmetic (no shuffles);
//   * the optional GraphNorm-backward column moments of the block below (GnMomFuse, same contract as dense_fwd_moments)
//     are summed over the CTA's rows through shared memory in row order and folded across CTAs by the last CTA.
// Falls back (returns 1) for anything else; bg_dense.cu keeps the tiled kernel for large N and wide layers.
#include <stdlib.h>

#include "bg_common.cuh"

namespace bg {

constexpr int RD_T = 128;            // threads per CTA = rows per CTA
constexpr int64_t RD_MAX_N = 131072;  // latency regime only (larger problems are bandwidth-bound: tiled kernel)

struct RdParams {
    int64_t N;
    int nseg;
    int off[BG_MAX_SEG + 1];
    BgSeg seg[BG_MAX_SEG];
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

template <int COUT, bool MOM>
__global__ void __launch_bounds__(RD_T) rowdense_kernel(const RdParams p) {
    pdl_prologue();
    extern __shared__ __align__(16) float rd_smem[];
    constexpr int WSTR = COUT + 4;  // weight row stride: float4-aligned, spreads the staging stores over 8 banks
    const int K = p.K, KS = K | 1;  // odd x row stride: thread t reads Xs[t * KS + k] conflict-free
    float* Ws = rd_smem;                                  // [K][WSTR]
    float* Xs = rd_smem + (size_t)K * WSTR;               // [RD_T][KS]   (MOM: reused as [RD_T][2 COUT + 1])
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * RD_T;

    // ---- stage W: Ws[k][c] = W[c * w_so + k * w_sk] (zero beyond Cout).  Loads are issued in register batches (all of a
    // batch in flight before the first shared store) so the staging costs one or two L2 round trips, not one per element.
    {
        const bool rowmajor = p.w_sk == 1;
        const int inner = rowmajor ? K : p.Cout;        // contiguous run length in W
        const int outer = rowmajor ? p.Cout : K;        // number of runs
        const int64_t run_ld = rowmajor ? p.w_so : p.w_sk;
        const bool vec = (inner & 3) == 0 && (run_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0;
        if (vec) {
            const int i4 = inner >> 2, total = outer * i4;
            for (int base = 0; base < total; base += 8 * RD_T) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = base + u * RD_T + tid;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < total) v[u] = __ldg(reinterpret_cast<const float4*>(p.W + (int64_t)(idx / i4) * run_ld) + (idx % i4));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = base + u * RD_T + tid;
                    if (idx < total) {
                        const int o = idx / i4, i = (idx % i4) * 4;
                        if (rowmajor) {  // o = output column c, i = k
                            Ws[(i + 0) * WSTR + o] = v[u].x;
                            Ws[(i + 1) * WSTR + o] = v[u].y;
                            Ws[(i + 2) * WSTR + o] = v[u].z;
                            Ws[(i + 3) * WSTR + o] = v[u].w;
                        } else {         // o = k, i = output column c
                            *reinterpret_cast<float4*>(Ws + o * WSTR + i) = v[u];
                        }
                    }
                }
            }
        } else {
            const int total = outer * inner;
            for (int base = 0; base < total; base += 16 * RD_T) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = base + u * RD_T + tid;
                    v[u] = idx < total ? __ldg(p.W + (int64_t)(idx / inner) * run_ld + (idx % inner)) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = base + u * RD_T + tid;
                    if (idx < total) {
                        const int o = idx / inner, i = idx % inner;
                        Ws[rowmajor ? i * WSTR + o : o * WSTR + i] = v[u];
                    }
                }
            }
        }
        // zero the padded output columns [Cout, COUT)
        if (p.Cout < COUT)
            for (int idx = tid; idx < K * (COUT - p.Cout); idx += RD_T) {
                const int k = idx / (COUT - p.Cout), c = p.Cout + idx % (COUT - p.Cout);
                Ws[k * WSTR + c] = 0.f;
            }
    }
    // ---- stage X
    {
        const BgSeg& s0 = p.seg[0];
        const bool flat = p.nseg == 1 && s0.ptr && !s0.gather && (K & 3) == 0 && (s0.ld & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(s0.ptr) & 15) == 0;
        if (flat) {  // one plain segment: the tile is RD_T rows of K/4 float4, up to 16 per thread in flight at once
            const int k4 = K >> 2, total = RD_T * k4;
            for (int base = 0; base < total; base += 16 * RD_T) {
                float4 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = base + u * RD_T + tid;
                    const int r = idx / k4;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < total && row0 + r < p.N) v[u] = __ldg(reinterpret_cast<const float4*>(s0.ptr + (row0 + r) * s0.ld) + (idx % k4));
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = base + u * RD_T + tid;
                    if (idx < total) {
                        float* dst = Xs + (idx / k4) * KS + (idx % k4) * 4;
                        dst[0] = v[u].x; dst[1] = v[u].y; dst[2] = v[u].z; dst[3] = v[u].w;
                    }
                }
            }
        } else {  // general: thread -> fixed column kc = tid % KP (segment / gather resolved once), rows r0 + i * RPI
            int KP = 1;
            while (KP < K) KP <<= 1;
            const int kc = tid % KP, RPI = RD_T / KP, r0 = tid / KP;
            const float* base = nullptr;
            const int32_t* gather = nullptr;
            int ld = 0;
            bool ones = false;
            if (kc < K) {
#pragma unroll
                for (int q = 0; q < BG_MAX_SEG; ++q)
                    if (q < p.nseg && kc >= p.off[q] && kc < p.off[q + 1]) {
                        ones = p.seg[q].ptr == nullptr;
                        base = p.seg[q].ptr ? p.seg[q].ptr + (kc - p.off[q]) : nullptr;
                        gather = p.seg[q].gather;
                        ld = p.seg[q].ld;
                    }
                for (int i0 = 0; i0 < KP; i0 += 16) {
                    int64_t gi[16];
                    float v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int64_t gr = row0 + r0 + (int64_t)(i0 + u) * RPI;
                        gi[u] = (i0 + u < KP && gr < p.N) ? (gather ? (int64_t)__ldg(gather + gr) : gr) : -1;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) v[u] = gi[u] < 0 ? 0.f : (base ? __ldg(base + gi[u] * ld) : (ones ? 1.f : 0.f));
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (i0 + u < KP) Xs[(r0 + (i0 + u) * RPI) * KS + kc] = v[u];
                }
            }
        }
    }
    __syncthreads();

    // ---- K loop: Cout independent FMA chains per thread, k ascending
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
    const float* xrow = Xs + tid * KS;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float xv = xrow[k];
        const float4* w4 = reinterpret_cast<const float4*>(Ws + k * WSTR);
#pragma unroll
        for (int j4 = 0; j4 < COUT / 4; ++j4) {
            const float4 w = w4[j4];
            acc[4 * j4 + 0] = fmaf(xv, w.x, acc[4 * j4 + 0]);
            acc[4 * j4 + 1] = fmaf(xv, w.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(xv, w.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(xv, w.w, acc[4 * j4 + 3]);
        }
    }

    // ---- epilogue, thread-local
    const int64_t grow = row0 + tid;
    const bool live = grow < p.N;
    const int Cout = p.Cout;
    if (p.bias) {
#pragma unroll
        for (int j = 0; j < COUT; ++j)
            if (j < Cout) acc[j] += __ldg(p.bias + j);
    }
    if (p.gamma) {  // LayerNorm over the row (Cout == COUT, checked on the host), eps = 1e-5
        float sm = 0.f;
#pragma unroll
        for (int j = 0; j < COUT; ++j) sm += acc[j];
        const float mean = sm / (float)COUT;
        float vs = 0.f;
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            const float dlt = acc[j] - mean;
            vs = fmaf(dlt, dlt, vs);
        }
        const float rs = 1.f / sqrtf(vs / (float)COUT + 1e-5f);
        if (p.rstd && live) p.rstd[grow] = rs;
#pragma unroll
        for (int j4 = 0; j4 < COUT / 4; ++j4) {
            float xh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = 4 * j4 + u;
                xh[u] = (acc[j] - mean) * rs;
                acc[j] = fmaf(xh[u], __ldg(p.gamma + j), __ldg(p.beta + j));
            }
            if (p.xhat && live) *reinterpret_cast<float4*>(p.xhat + grow * COUT + 4 * j4) = make_float4(xh[0], xh[1], xh[2], xh[3]);
        }
    }
    if (p.act == BG_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = acc[j] > 0.f ? acc[j] : 0.f;
    } else if (p.act == BG_ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = acc[j] > 0.f ? acc[j] : 0.2f * acc[j];
    }
    if (p.att_src) {
        float ss = 0.f, dd = 0.f;
#pragma unroll
        for (int j = 0; j < COUT; ++j)
            if (j < Cout) {
                ss = fmaf(acc[j], __ldg(p.att_src + j), ss);
                dd = fmaf(acc[j], __ldg(p.att_dst + j), dd);
            }
        if (live) {
            p.s[grow] = ss;
            p.d[grow] = dd;
        }
    }
    if (p.gate && live) {  // fused activation backward of the layer below
        const float* grw = p.gate + grow * p.ld_gate;
        if ((p.ld_gate & 3) == 0 && (Cout & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gate) & 15) == 0) {
#pragma unroll
            for (int j4 = 0; j4 < COUT / 4; ++j4)
                if (4 * j4 < Cout) {
                    const float4 gv = __ldg(reinterpret_cast<const float4*>(grw) + j4);
                    acc[4 * j4 + 0] *= gv.x > 0.f ? 1.f : p.gate_slope;
                    acc[4 * j4 + 1] *= gv.y > 0.f ? 1.f : p.gate_slope;
                    acc[4 * j4 + 2] *= gv.z > 0.f ? 1.f : p.gate_slope;
                    acc[4 * j4 + 3] *= gv.w > 0.f ? 1.f : p.gate_slope;
                }
        } else {
#pragma unroll
            for (int j = 0; j < COUT; ++j)
                if (j < Cout) acc[j] *= (__ldg(grw + j) > 0.f ? 1.f : p.gate_slope);
        }
    }
    if (live) {
        float* orow = p.out + grow * p.ld_out;
        if ((p.ld_out & 3) == 0 && (Cout & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
#pragma unroll
            for (int j4 = 0; j4 < COUT / 4; ++j4)
                if (4 * j4 < Cout)
                    *reinterpret_cast<float4*>(orow + 4 * j4) = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < COUT; ++j)
                if (j < Cout) orow[j] = acc[j];
        }
    }

    if constexpr (MOM) {
        // acc = this row's gx1 (after the gate): column sums of gy = gx1 * keep_scale * [x1 > 0] and gy * (o - alpha mu) over
        // the CTA's rows (row order), then the cross-CTA fold + the final arithmetic of gn_bwd_moments_kernel by the last CTA.
        constexpr int MS = 2 * COUT + 1;
        __syncthreads();  // every thread is done reading its x row: the tile is reused
        float* Ms = Xs;
        {
            const float* orw = p.mom.o + grow * Cout;
            const float* xrw = p.mom.x1 + grow * Cout;
#pragma unroll
            for (int j = 0; j < COUT; ++j) {
                float g0 = 0.f, g1 = 0.f;
                if (live && j < Cout) {
                    const float gy = __ldg(xrw + j) > 0.f ? acc[j] * p.mom.keep_scale : 0.f;
                    g0 = gy;
                    g1 = gy * (__ldg(orw + j) - __ldg(p.mom.alpha + j) * __ldg(p.mom.stats + j));
                }
                Ms[tid * MS + j] = g0;
                Ms[tid * MS + COUT + j] = g1;
            }
        }
        __syncthreads();
        __shared__ float fred[kThreads];
        __shared__ float fsum[2 * COUT];
        float* partials = p.mom.partials;
        for (int i = tid; i < 2 * COUT; i += RD_T) {
            float t = 0.f;
#pragma unroll 8
            for (int r = 0; r < RD_T; ++r) t += Ms[r * MS + i];
            partials[(int64_t)blockIdx.x * 2 * COUT + i] = t;
        }
        if (!hier_fold_any(partials, partials + (int64_t)gridDim.x * 2 * COUT, 2 * COUT, p.mom.counters, fred, fsum)) return;
        for (int c = tid; c < Cout; c += RD_T) {
            const int C = Cout;
            const float n = (float)p.N;
            const float G0 = fsum[c] / n, G1 = fsum[COUT + c] / n;
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

static int g_rowdense = -1;  // BG_ROWDENSE=0 keeps every dense layer on the tiled kernel (A/B switch)

template <int COUT>
static void rd_launch(const RdParams& p, bool mom, unsigned grid, size_t smem, cudaStream_t st) {
    if (mom) {
        static bool once = (cudaFuncSetAttribute(rowdense_kernel<COUT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024), true);
        (void)once;
        launch_k(rowdense_kernel<COUT, true>, grid, RD_T, smem, st, p);
    } else {
        static bool once = (cudaFuncSetAttribute(rowdense_kernel<COUT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024), true);
        (void)once;
        launch_k(rowdense_kernel<COUT, false>, grid, RD_T, smem, st, p);
    }
}

// BG_OK when launched, 1 when the shape is not eligible (caller continues with the tiled kernel), < 0 on error.
int rowdense_try(const BgDense* a, const int* seg_off, int K, const GnMomFuse* mom, cudaStream_t st) {
    if (g_rowdense < 0) g_rowdense = getenv("BG_ROWDENSE") ? atoi(getenv("BG_ROWDENSE")) : 1;
    if (!g_rowdense) return 1;
    // measured against the tiled kernel on the dependency chain (profiles/r02_summary.md): ahead only for very narrow layers
    // (the 1/2/4-channel bottleneck blocks, the critic's 8 -> 1 score layer); BG_ROWDENSE=2 lifts the limit (K <= 128, Cout <= 64)
    if (a->N > RD_MAX_N || K < 1 || (g_rowdense == 2 ? (K > 128 || a->Cout > 64) : (K > 16 || a->Cout > 8))) return 1;
    int cp = 4;
    while (cp < a->Cout) cp <<= 1;
    if (a->ln_gamma && cp != a->Cout) return 1;
    if (a->xhat && (reinterpret_cast<uintptr_t>(a->xhat) & 15)) return 1;
    RdParams p;
    p.N = a->N;
    p.nseg = a->nseg;
    for (int q = 0; q < BG_MAX_SEG; ++q) p.seg[q] = q < a->nseg ? a->seg[q] : BgSeg{nullptr, nullptr, 0, 0};
    for (int q = 0; q <= BG_MAX_SEG; ++q) p.off[q] = seg_off[q];
    p.K = K;
    p.W = a->W; p.w_so = a->w_so; p.w_sk = a->w_sk; p.Cout = a->Cout;
    p.bias = a->bias; p.gamma = a->ln_gamma; p.beta = a->ln_beta;
    p.att_src = a->att_src; p.att_dst = a->att_dst; p.act = a->act;
    p.out = a->out; p.ld_out = a->ld_out; p.xhat = a->xhat; p.rstd = a->rstd; p.s = a->s; p.d = a->d;
    p.gate = a->gate; p.ld_gate = a->ld_gate; p.gate_slope = a->gate_slope;
    p.mom = mom ? *mom : GnMomFuse{};
    const int KS = K | 1;
    size_t xs = (size_t)RD_T * KS;
    if (mom && (size_t)RD_T * (2 * cp + 1) > xs) xs = (size_t)RD_T * (2 * cp + 1);
    const size_t smem = ((size_t)K * (cp + 4) + xs) * sizeof(float);
    if (smem > 160 * 1024) return 1;
    const unsigned grid = (unsigned)ceil_div(a->N, RD_T);
    switch (cp) {
        case 4: rd_launch<4>(p, mom != nullptr, grid, smem, st); break;
        case 8: rd_launch<8>(p, mom != nullptr, grid, smem, st); break;
        case 16: rd_launch<16>(p, mom != nullptr, grid, smem, st); break;
        case 32: rd_launch<32>(p, mom != nullptr, grid, smem, st); break;
        default: rd_launch<64>(p, mom != nullptr, grid, smem, st); break;
    }
    return check_launch(mom ? "dense_fwd_moments(row)" : "bg_dense_fwd(row)");
}

}  // namespace bg

// 1 = row-per-thread kernel for the small dense layers (default), 0 = tiled kernel everywhere.  Returns the previous setting.
extern "C" int bg_set_rowdense(int32_t on) {
    if (bg::g_rowdense < 0) bg::g_rowdense = getenv("BG_ROWDENSE") ? atoi(getenv("BG_ROWDENSE")) : 1;
    const int prev = bg::g_rowdense;
    bg::g_rowdense = on < 0 ? 0 : (on > 2 ? 2 : on);
    return prev;
}
