// Aggregation kernels of the non-default conv types the reference can select with GENERATOR_CONV_TYPE /
// DISCRIMINATOR_CONV_TYPE (models.py:22-31,166-175; config.py:89,93): GCNConv, GraphConv (weighted neighbour sums:
// bg_gcn_norm + bg_spmm) and GATv2Conv (bg_gatv2_fwd / _bwd / _bwd2).  Same conventions as bg_gat.cu: destination-sorted
// CSR with the self loop last in every row, its transpose (CSC + perm) for the out-edge sums, one lane group per node
// row, fixed edge order, no atomics.  These types are not exercised by any shipped configuration of the reference, so
// the kernels are the plain one-row-per-group form (no chunked sweep / software pipeline).
//
// GATv2 math (torch_geometric GATv2Conv, heads=1, share_weights=False; restated in oracle/pyg.py::GATv2Conv), edge e = j -> i:
//   v_e = xl_j + xr_i        logit_e = att . lrelu(v_e)        p = softmax_i(logit)        out_i = sum_e p_e xl_j + bias
// first order (g = d loss / d out), with w_e = att * lrelu'(v_e) (a vector), c_e = g_i . xl_j, r_i = sum_e p_e c_e,
// delta_e = p_e (c_e - r_i):
//   gxl_j = sum_{e in out(j)} [p_e g_i + delta_e w_e]     gxr_i = sum_{e in in(i)} delta_e w_e
//   gatt  = sum_e delta_e lrelu(v_e)      (returned per destination row, the column sum is one bg_dense_wgrad)
// second order (WGAN-GP): cotangents Hl on gxl, Hr on gxr; A_e = Hl_j . g_i, B_e = (Hl_j + Hr_i) . w_e,
// Bbar_i = sum p B, T_e = B_e - Bbar_i, pi_e = A_e + c_e B_e - c_e Bbar_i - r_i B_e, Pibar_i = sum p pi,
// dl2_e = p_e (pi_e - Pibar_i)   (lrelu is piecewise linear: its second derivative vanishes a.e.):
//   cot(g_i)  = sum_{in(i)} [p_e Hl_j + p_e T_e xl_j]          cot(xr_i) = sum_{in(i)} dl2_e w_e
//   cot(xl_j) = sum_{out(j)} [p_e T_e g_i + dl2_e w_e]         (the first-order source kernel with other scalars)
//   cot(att)  = sum_e [dl2_e lrelu(v_e) + delta_e (Hl_j + Hr_i) * lrelu'(v_e)]
// derived for this file and checked against fp64 autograd of the oracle in tests/test_conv_types_gpu.py.
#include "bg_common.cuh"

namespace bg {

// ------------------------------------------------------------------------------------------
// GCN symmetric normalisation, CSR order: w[e] = deg(i)^-1/2 * deg(j)^-1/2, deg = in-degree incl. the self loop
// (gcn_norm with add_remaining_self_loops, fill 1).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gcn_norm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                            float* __restrict__ w, int64_t N) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float di = 1.0f / sqrtf((float)(end - beg));
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        const float dj = 1.0f / sqrtf((float)(__ldg(rowptr + j + 1) - __ldg(rowptr + j)));
        w[e] = dj * di;
    }
}

// ------------------------------------------------------------------------------------------
// weighted neighbour sum: out[r] = sum over the row's entries of w * x[other end]  (+ bias)
//   TRANSPOSE = false: CSR (in-edges of destination r, weight w[e]);  true: CSC (out-edges of source r, weight w[perm[k]])
//   self_loops = 0 skips the self-loop entry of every row (GraphConv adds none)
// ------------------------------------------------------------------------------------------
template <int C, bool TRANSPOSE>
__global__ void __launch_bounds__(kThreads) spmm_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                        const int32_t* __restrict__ perm, const float* __restrict__ w,
                                                        const float* __restrict__ x, const float* __restrict__ bias,
                                                        float* __restrict__ out, int64_t N, int self_loops) {
    pdl_prologue();
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(ptr + row), end = __ldg(ptr + row + 1);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const float* xb = x + sub * VEC;
    for (int k = beg; k < end; ++k) {
        const int o = __ldg(idx + k);
        if (!self_loops && o == (int)row) continue;
        const float we = w ? __ldg(w + (TRANSPOSE ? __ldg(perm + k) : k)) : 1.f;
        Vec<VEC> xv;
        xv.load(xb + (int64_t)o * C);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(we, xv.v[v], acc[v]);
    }
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = acc[v] + (bias ? __ldg(bias + sub * VEC + v) : 0.f);
    o.store(out + row * C + sub * VEC);
}

// ------------------------------------------------------------------------------------------
// GATv2 forward (per destination row)
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gatv2_fwd_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ xl,
    const float* __restrict__ xr, const float* __restrict__ att, const float* __restrict__ bias,
    float* __restrict__ out, float* __restrict__ logit, float* __restrict__ m_out, float* __restrict__ z_out, int64_t N,
    float slope) {
    pdl_prologue();
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    Vec<VEC> ri, a;
    ri.load(xr + row * C + sub * VEC);
    a.load(att + sub * VEC);
    const float* lb = xl + sub * VEC;
    float mx = -INFINITY;
    for (int e = beg; e < end; ++e) {
        Vec<VEC> lv;
        lv.load(lb + (int64_t)__ldg(col + e) * C);
        float t = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) t = fmaf(a.v[v], lrelu(lv.v[v] + ri.v[v], slope), t);
        t = gsum<LANES>(t, gm);
        mx = fmaxf(mx, t);
        if (sub == 0) logit[e] = t;
    }
    __syncwarp(gm);
    float zs = 0.f;
    for (int e = beg + sub; e < end; e += LANES) zs += expf(logit[e] - mx);
    zs = gsum<LANES>(zs, gm) + 1e-16f;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int e = beg; e < end; ++e) {
        Vec<VEC> lv;
        lv.load(lb + (int64_t)__ldg(col + e) * C);
        const float p = expf(logit[e] - mx) / zs;
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(p, lv.v[v], acc[v]);
    }
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = acc[v] + (bias ? __ldg(bias + sub * VEC + v) : 0.f);
    o.store(out + row * C + sub * VEC);
    if (sub == 0) {
        m_out[row] = mx;
        z_out[row] = zs;
    }
}

// ------------------------------------------------------------------------------------------
// GATv2 backward, destination pass: P[e] = p_e, DL[e] = delta_e, gxr, per-row attention-vector gradient
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gatv2_bwd_dst_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ gout,
    const float* __restrict__ xl, const float* __restrict__ xr, const float* __restrict__ att,
    const float* __restrict__ logit, const float* __restrict__ m_in, const float* __restrict__ z_in, float* __restrict__ P,
    float* __restrict__ DL, float* __restrict__ gxr, float* __restrict__ garow, int64_t N, float slope) {
    pdl_prologue();
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float mi = __ldg(m_in + row), zi = __ldg(z_in + row);
    Vec<VEC> gi, ri, a;
    gi.load(gout + row * C + sub * VEC);
    ri.load(xr + row * C + sub * VEC);
    a.load(att + sub * VEC);
    const float* lb = xl + sub * VEC;
    float r = 0.f;
    for (int e = beg; e < end; ++e) {
        Vec<VEC> lv;
        lv.load(lb + (int64_t)__ldg(col + e) * C);
        float c = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) c = fmaf(gi.v[v], lv.v[v], c);
        c = gsum<LANES>(c, gm);
        const float p = expf(__ldg(logit + e) - mi) / zi;
        r = fmaf(p, c, r);
        if (sub == 0) {
            P[e] = p;
            DL[e] = c;
        }
    }
    __syncwarp(gm);
    float gr[VEC], ga[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) gr[v] = ga[v] = 0.f;
    for (int e = beg; e < end; ++e) {
        Vec<VEC> lv;
        lv.load(lb + (int64_t)__ldg(col + e) * C);
        const float dl = P[e] * (DL[e] - r);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float u = lv.v[v] + ri.v[v];
            gr[v] = fmaf(dl * lrelu_grad(u, slope), a.v[v], gr[v]);
            ga[v] = fmaf(dl, lrelu(u, slope), ga[v]);
        }
    }
    __syncwarp(gm);
    for (int e = beg + sub; e < end; e += LANES) DL[e] = P[e] * (DL[e] - r);
    Vec<VEC> o1, o2;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        o1.v[v] = gr[v];
        o2.v[v] = ga[v];
    }
    o1.store(gxr + row * C + sub * VEC);
    o2.store(garow + row * C + sub * VEC);
}

// ------------------------------------------------------------------------------------------
// GATv2 backward, source pass (shared by first and second order):
//   out[j] = sum_{k in out(j)} [ PA[e] * G[i] + PB[e] * att * lrelu'(xl_j + xr_i) ],  e = perm[k], i = cscrow[k]
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gatv2_bwd_src_kernel(
    const int32_t* __restrict__ cscptr, const int32_t* __restrict__ cscrow, const int32_t* __restrict__ perm,
    const float* __restrict__ PA, const float* __restrict__ PB, const float* __restrict__ G, const float* __restrict__ xl,
    const float* __restrict__ xr, const float* __restrict__ att, float* __restrict__ out, int64_t N, float slope) {
    pdl_prologue();
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(cscptr + row), end = __ldg(cscptr + row + 1);
    Vec<VEC> lj, a;
    lj.load(xl + row * C + sub * VEC);
    a.load(att + sub * VEC);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int k = beg; k < end; ++k) {
        const int i = __ldg(cscrow + k), e = __ldg(perm + k);
        Vec<VEC> gv, rv;
        gv.load(G + (int64_t)i * C + sub * VEC);
        rv.load(xr + (int64_t)i * C + sub * VEC);
        const float pa = PA[e], pb = PB[e];
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            acc[v] = fmaf(pa, gv.v[v], fmaf(pb * lrelu_grad(lj.v[v] + rv.v[v], slope), a.v[v], acc[v]));
    }
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
    o.store(out + row * C + sub * VEC);
}

// ------------------------------------------------------------------------------------------
// GATv2 second-order backward, destination pass.  Scratch S0..S5 are E floats each; on exit S4 = p*T, S5 = dl2
// (inputs of the source pass).
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gatv2_bwd2_dst_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ Hl,
    const float* __restrict__ Hr, const float* __restrict__ gout, const float* __restrict__ xl, const float* __restrict__ xr,
    const float* __restrict__ att, const float* __restrict__ logit, const float* __restrict__ m_in,
    const float* __restrict__ z_in, float* __restrict__ S0, float* __restrict__ S1, float* __restrict__ S2,
    float* __restrict__ S3, float* __restrict__ S4, float* __restrict__ S5, float* __restrict__ gt, float* __restrict__ cxr,
    float* __restrict__ carow, int64_t N, float slope) {
    pdl_prologue();
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float mi = __ldg(m_in + row), zi = __ldg(z_in + row);
    Vec<VEC> gi, ri, hri, a;
    gi.load(gout + row * C + sub * VEC);
    ri.load(xr + row * C + sub * VEC);
    hri.load(Hr + row * C + sub * VEC);
    a.load(att + sub * VEC);
    float r = 0.f, Bbar = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        Vec<VEC> lv, hv;
        lv.load(xl + (int64_t)j * C + sub * VEC);
        hv.load(Hl + (int64_t)j * C + sub * VEC);
        float c = 0.f, A = 0.f, B = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            c = fmaf(gi.v[v], lv.v[v], c);
            A = fmaf(hv.v[v], gi.v[v], A);
            B = fmaf((hv.v[v] + hri.v[v]) * lrelu_grad(lv.v[v] + ri.v[v], slope), a.v[v], B);
        }
        c = gsum<LANES>(c, gm);
        A = gsum<LANES>(A, gm);
        B = gsum<LANES>(B, gm);
        const float p = expf(__ldg(logit + e) - mi) / zi;
        r = fmaf(p, c, r);
        Bbar = fmaf(p, B, Bbar);
        if (sub == 0) {
            S0[e] = p;
            S1[e] = c;
            S2[e] = A;
            S3[e] = B;
        }
    }
    __syncwarp(gm);
    float Pibar = 0.f;
    for (int e = beg + sub; e < end; e += LANES) {
        const float p = S0[e], c = S1[e], A = S2[e], B = S3[e];
        const float pi = A + c * B - c * Bbar - r * B;
        Pibar = fmaf(p, pi, Pibar);
        S2[e] = pi;
    }
    Pibar = gsum<LANES>(Pibar, gm);
    __syncwarp(gm);
    float g2[VEC], cr[VEC], ca[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) g2[v] = cr[v] = ca[v] = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        Vec<VEC> lv, hv;
        lv.load(xl + (int64_t)j * C + sub * VEC);
        hv.load(Hl + (int64_t)j * C + sub * VEC);
        const float p = S0[e], c = S1[e], B = S3[e];
        const float pT = p * (B - Bbar), dl = p * (c - r), dl2 = p * (S2[e] - Pibar);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float u = lv.v[v] + ri.v[v], mk = lrelu_grad(u, slope);
            g2[v] = fmaf(p, hv.v[v], fmaf(pT, lv.v[v], g2[v]));
            cr[v] = fmaf(dl2 * mk, a.v[v], cr[v]);
            ca[v] = fmaf(dl2, lrelu(u, slope), fmaf(dl * mk, hv.v[v] + hri.v[v], ca[v]));
        }
        if (sub == 0) {
            S4[e] = pT;
            S5[e] = dl2;
        }
    }
    Vec<VEC> o1, o2, o3;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        o1.v[v] = g2[v];
        o2.v[v] = cr[v];
        o3.v[v] = ca[v];
    }
    o1.store(gt + row * C + sub * VEC);
    o2.store(cxr + row * C + sub * VEC);
    o3.store(carow + row * C + sub * VEC);
}

#define BG_CONV_DISPATCH_C(C, CALL)                                                    \
    switch (C) {                                                                       \
        case 1: CALL(1); break;                                                        \
        case 2: CALL(2); break;                                                        \
        case 4: CALL(4); break;                                                        \
        case 8: CALL(8); break;                                                        \
        case 16: CALL(16); break;                                                      \
        case 32: CALL(32); break;                                                      \
        case 64: CALL(64); break;                                                      \
        case 128: CALL(128); break;                                                    \
        default:                                                                       \
            bg::set_error("unsupported channel width C=%d (supported: 1,2,4,...,128)", (int)(C)); \
            return BG_EUNSUPPORTED;                                                    \
    }

static int check_graph_conv(const BgGraph* g) {
    BG_REQUIRE(g && g->rowptr && g->col && g->cscptr && g->cscrow && g->perm, BG_EINVAL, "BgGraph has null arrays");
    BG_REQUIRE(g->N > 0 && g->E >= g->N, BG_EINVAL, "BgGraph: need N>0 and E>=N (self loops), got N=%lld E=%lld", (long long)g->N,
               (long long)g->E);
    return BG_OK;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_gcn_norm(const BgGraph* g, float* w, void* stream) {
    if (int rc = check_graph_conv(g)) return rc;
    BG_REQUIRE(w, BG_EINVAL, "bg_gcn_norm: null pointer");
    launch_k(gcn_norm_kernel, (unsigned)ceil_div(g->N, kThreads), kThreads, 0, as_stream(stream), g->rowptr, g->col, w, g->N);
    return check_launch("bg_gcn_norm");
}

extern "C" int bg_spmm(const BgGraph* g, const float* w, const float* x, const float* bias, float* out, int32_t C,
                       int32_t transpose, int32_t self_loops, void* stream) {
    if (int rc = check_graph_conv(g)) return rc;
    BG_REQUIRE(x && out, BG_EINVAL, "bg_spmm: null pointer");
    cudaStream_t st = as_stream(stream);
#define CALL(CC)                                                                                                        \
    do {                                                                                                                \
        const unsigned grid = (unsigned)ceil_div(g->N, RowMap<CC>::RPC);                                                \
        if (transpose)                                                                                                  \
            launch_k(spmm_kernel<CC, true>, grid, kThreads, 0, st, g->cscptr, g->cscrow, g->perm, w, x, bias, out, g->N, self_loops); \
        else                                                                                                            \
            launch_k(spmm_kernel<CC, false>, grid, kThreads, 0, st, g->rowptr, g->col, nullptr, w, x, bias, out, g->N, self_loops);   \
    } while (0)
    BG_CONV_DISPATCH_C(C, CALL)
#undef CALL
    return check_launch("bg_spmm");
}

extern "C" int bg_gatv2_fwd(const BgGraph* g, const float* xl, const float* xr, const float* att, const float* bias, float* out,
                            float* logit, float* m, float* z, int32_t C, float slope, void* stream) {
    if (int rc = check_graph_conv(g)) return rc;
    BG_REQUIRE(xl && xr && att && out && logit && m && z, BG_EINVAL, "bg_gatv2_fwd: null pointer");
    cudaStream_t st = as_stream(stream);
#define CALL(CC)                                                                                               \
    launch_k(gatv2_fwd_kernel<CC>, (unsigned)ceil_div(g->N, RowMap<CC>::RPC), kThreads, 0, st, g->rowptr, g->col, xl, xr, att, bias, \
                                                                                          out, logit, m, z, g->N, slope)
    BG_CONV_DISPATCH_C(C, CALL)
#undef CALL
    return check_launch("bg_gatv2_fwd");
}

extern "C" int bg_gatv2_bwd(const BgGraph* g, const float* gout, const float* xl, const float* xr, const float* att,
                            const float* logit, const float* m, const float* z, float* P, float* DL, float* gxl, float* gxr,
                            float* garow, int32_t C, float slope, void* stream) {
    if (int rc = check_graph_conv(g)) return rc;
    BG_REQUIRE(gout && xl && xr && att && logit && m && z && P && DL && gxl && gxr && garow, BG_EINVAL,
               "bg_gatv2_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
#define CALL(CC)                                                                                                          \
    do {                                                                                                                  \
        const unsigned grid = (unsigned)ceil_div(g->N, RowMap<CC>::RPC);                                                  \
        launch_k(gatv2_bwd_dst_kernel<CC>, grid, kThreads, 0, st, g->rowptr, g->col, gout, xl, xr, att, logit, m, z, P, DL, gxr, \
                                                            garow, g->N, slope);                                          \
        launch_k(gatv2_bwd_src_kernel<CC>, grid, kThreads, 0, st, g->cscptr, g->cscrow, g->perm, P, DL, gout, xl, xr, att, gxl,  \
                                                            g->N, slope);                                                 \
    } while (0)
    BG_CONV_DISPATCH_C(C, CALL)
#undef CALL
    return check_launch("bg_gatv2_bwd");
}

extern "C" int bg_gatv2_bwd2(const BgGraph* g, const float* Hl, const float* Hr, const float* gout, const float* xl,
                             const float* xr, const float* att, const float* logit, const float* m, const float* z,
                             float* scratch, float* gt, float* cxl, float* cxr, float* carow, int32_t C, float slope,
                             void* stream) {
    if (int rc = check_graph_conv(g)) return rc;
    BG_REQUIRE(Hl && Hr && gout && xl && xr && att && logit && m && z && scratch && gt && cxl && cxr && carow, BG_EINVAL,
               "bg_gatv2_bwd2: null pointer");
    cudaStream_t st = as_stream(stream);
    float* S[6];
    for (int q = 0; q < 6; ++q) S[q] = scratch + (size_t)q * g->E;
#define CALL(CC)                                                                                                             \
    do {                                                                                                                     \
        const unsigned grid = (unsigned)ceil_div(g->N, RowMap<CC>::RPC);                                                     \
        launch_k(gatv2_bwd2_dst_kernel<CC>, grid, kThreads, 0, st, g->rowptr, g->col, Hl, Hr, gout, xl, xr, att, logit, m, z, S[0], \
                                                             S[1], S[2], S[3], S[4], S[5], gt, cxr, carow, g->N, slope);     \
        launch_k(gatv2_bwd_src_kernel<CC>, grid, kThreads, 0, st, g->cscptr, g->cscrow, g->perm, S[4], S[5], gout, xl, xr, att, cxl, \
                                                            g->N, slope);                                                    \
    } while (0)
    BG_CONV_DISPATCH_C(C, CALL)
#undef CALL
    return check_launch("bg_gatv2_bwd2");
}
