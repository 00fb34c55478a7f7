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

#include "bg_common.cuh"

namespace bg {

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gat_fwd_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ h,
    const float* __restrict__ s, const float* __restrict__ d, const float* __restrict__ bias,
    float* __restrict__ out, float* __restrict__ m_out, float* __restrict__ z_out, int64_t N, float slope, int pf_rows) {
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;  // whole groups leave together; group masks keep the rest legal
    // Large graphs are latency-bound on the dependent chain rowptr -> col -> s -> h (every array cold in L2): pull the
    // lines a block ~2 waves ahead will need into L2 now.  Prefetches hold no registers, so they decouple the bytes in
    // flight from the register file; the block that later owns those rows finds its whole chain L2-resident.
    const int64_t prow = row + pf_rows;
    int pbeg = -1;
    if (pf_rows > 0 && prow < N) {
        if (sub == 0) {
            pbeg = __ldg(rowptr + prow);  // consumed only at the end of the kernel (no stall here)
            prefetch_l2(d + prow);
            prefetch_l2(s + prow);
        }
        if (((sub * VEC * 4) & 127) == 0) prefetch_l2(h + prow * C + sub * VEC);
    }
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float di = __ldg(d + row);

    // softmax statistics, edges spread over the group's lanes
    float mx = -INFINITY;
    for (int e = beg + sub; e < end; e += LANES) mx = fmaxf(mx, lrelu(__ldg(s + __ldg(col + e)) + di, slope));
    mx = gmax<LANES>(mx, gm);
    float zs = 0.f;
    for (int e = beg + sub; e < end; e += LANES) zs += expf(lrelu(__ldg(s + __ldg(col + e)) + di, slope) - mx);
    zs = gsum<LANES>(zs, gm) + 1e-16f;

    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const float* hb = h + sub * VEC;
    int e = beg;
    for (; e + 4 <= end; e += 4) {  // 4 independent gathers in flight
        int j[4];
        float p[4];
        Vec<VEC> hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) j[k] = __ldg(col + e + k);
#pragma unroll
        for (int k = 0; k < 4; ++k) hv[k].load(hb + (int64_t)j[k] * C);
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = expf(lrelu(__ldg(s + j[k]) + di, slope) - mx) / zs;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = fmaf(p[k], hv[k].v[v], acc[v]);
    }
    for (; e < end; ++e) {
        const int j = __ldg(col + e);
        Vec<VEC> hv;
        hv.load(hb + (int64_t)j * C);
        const float p = expf(lrelu(__ldg(s + j) + di, slope) - mx) / zs;
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(p, hv.v[v], acc[v]);
    }
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = acc[v] + (bias ? __ldg(bias + sub * VEC + v) : 0.f);
    o.store(out + row * C + sub * VEC);
    if (sub == 0) {
        m_out[row] = mx;
        z_out[row] = zs;
        if (pbeg >= 0) {
            prefetch_l2(col + pbeg);
            prefetch_l2(col + pbeg + 8);
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward, pass A: per destination row (CSR).  Writes P[e] (softmax weight), DU[e] (d loss /
// d pre-activation logit u_e = s_j + d_i) and gsd[2i+1] = d loss / d d_i.
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gat_bwd_dst_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ gout,
    const float* __restrict__ h, const float* __restrict__ s, const float* __restrict__ d,
    const float* __restrict__ m_in, const float* __restrict__ z_in, float* __restrict__ P,
    float* __restrict__ DU, float* __restrict__ gsd, int64_t N, float slope) {
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const unsigned gm = group_mask<LANES>(lane);
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float di = __ldg(d + row), mi = __ldg(m_in + row), zi = __ldg(z_in + row);
    Vec<VEC> gi;
    gi.load(gout + row * C + sub * VEC);
    const float* hb = h + sub * VEC;
    float r = 0.f;
    for (int e = beg; e < end; ++e) {
        const int j = __ldg(col + e);
        Vec<VEC> hv;
        hv.load(hb + (int64_t)j * C);
        float c = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) c = fmaf(gi.v[v], hv.v[v], c);
        c = gsum<LANES>(c, gm);
        const float p = expf(lrelu(__ldg(s + j) + di, slope) - mi) / zi;
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
    if (sub == 0) gsd[2 * row + 1] = gd;
}

// ------------------------------------------------------------------------------------------
// backward, pass B: per source row (CSC).  out_tot[j] = sum_{e in out(j)} P[e] * G[i(e)]
//   + (sum_e DU[e]) * a_src + gsd[2j+1] * a_dst ;  gsd[2j] = sum_e DU[e].
// Shared by the first- and second-order backward (different P/DU/G).
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) gat_bwd_src_kernel(
    const int32_t* __restrict__ cscptr, const int32_t* __restrict__ cscrow, const int32_t* __restrict__ perm,
    const float* __restrict__ P, const float* __restrict__ DU, const float* __restrict__ G,
    const float* __restrict__ a_src, const float* __restrict__ a_dst, float* __restrict__ out_tot,
    float* __restrict__ gsd, int64_t N) {
    using M = RowMap<C>;
    constexpr int VEC = M::VEC, LANES = M::LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES;
    const int64_t row = (int64_t)blockIdx.x * M::RPC + warp * M::RPW + lane / LANES;
    if (row >= N) return;
    const int beg = __ldg(cscptr + row), end = __ldg(cscptr + row + 1);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    float gs = 0.f;
    const float* gb = G + sub * VEC;
    int k = beg;
    for (; k + 4 <= end; k += 4) {
        int i[4], e[4];
        float p[4], du[4];
        Vec<VEC> gv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            i[q] = __ldg(cscrow + k + q);
            e[q] = __ldg(perm + k + q);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) gv[q].load(gb + (int64_t)i[q] * C);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            p[q] = P[e[q]];
            du[q] = DU[e[q]];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            gs += du[q];
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = fmaf(p[q], gv[q].v[v], acc[v]);
        }
    }
    for (; k < end; ++k) {
        const int i = __ldg(cscrow + k), e = __ldg(perm + k);
        Vec<VEC> gv;
        gv.load(gb + (int64_t)i * C);
        const float p = P[e];
        gs += DU[e];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(p, gv.v[v], acc[v]);
    }
    const float gd = gsd[2 * row + 1];
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v)
        o.v[v] = acc[v] + gs * __ldg(a_src + sub * VEC + v) + gd * __ldg(a_dst + sub * VEC + v);
    o.store(out_tot + row * C + sub * VEC);
    if (sub == 0) gsd[2 * row] = gs;
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

// Prefetch distance in rows: ~2 waves of resident CTAs ahead; 0 (off) for graphs whose working set is L2-resident anyway.
template <int C>
static int prefetch_rows(int64_t N) {
    static const int env = getenv("BG_GAT_PF") ? atoi(getenv("BG_GAT_PF")) : -1;
    if (env >= 0) return env * RowMap<C>::RPC;
    if (N * C * 4 < (int64_t)(24 << 20)) return 0;
    return 2 * kSMs * 5 * RowMap<C>::RPC;
}

template <int C>
static int launch_fwd(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias,
                      float* out, float* m, float* z, float slope, cudaStream_t st) {
    const int64_t grid = ceil_div(g->N, RowMap<C>::RPC);
    gat_fwd_kernel<C><<<(unsigned)grid, kThreads, 0, st>>>(g->rowptr, g->col, h, s, d, bias, out, m, z, g->N, slope,
                                                          prefetch_rows<C>(g->N));
    return check_launch("bg_gat_fwd");
}
template <int C>
static int launch_bwd(const BgGraph* g, const float* gout, const float* h, const float* s, const float* d,
                      const float* m, const float* z, const float* a_src, const float* a_dst, float* P,
                      float* DU, float* gh_tot, float* gsd, float slope, cudaStream_t st) {
    const int64_t grid = ceil_div(g->N, RowMap<C>::RPC);
    gat_bwd_dst_kernel<C><<<(unsigned)grid, kThreads, 0, st>>>(g->rowptr, g->col, gout, h, s, d, m, z, P, DU, gsd,
                                                              g->N, slope);
    gat_bwd_src_kernel<C><<<(unsigned)grid, kThreads, 0, st>>>(g->cscptr, g->cscrow, g->perm, P, DU, gout, a_src,
                                                              a_dst, gh_tot, gsd, g->N);
    return check_launch("bg_gat_bwd");
}
template <int C>
static int launch_bwd2(const BgGraph* g, const float* Ht, const float* St, const float* Dt, const float* gout,
                       const float* h, const float* s, const float* d, const float* m, const float* z,
                       const float* a_src, const float* a_dst, float* scratch, float* gt, float* ht_tot,
                       float* sdt, float slope, cudaStream_t st) {
    const int64_t grid = ceil_div(g->N, RowMap<C>::RPC);
    float *A0 = scratch, *A1 = scratch + g->E, *A2 = scratch + 2 * g->E, *A3 = scratch + 3 * g->E;
    gat_bwd2_dst_kernel<C><<<(unsigned)grid, kThreads, 0, st>>>(g->rowptr, g->col, Ht, St, Dt, gout, h, s, d, m, z,
                                                               A0, A1, A2, A3, gt, sdt, g->N, slope);
    gat_bwd_src_kernel<C><<<(unsigned)grid, kThreads, 0, st>>>(g->cscptr, g->cscrow, g->perm, A1, A2, gout, a_src,
                                                              a_dst, ht_tot, sdt, g->N);
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
    BG_REQUIRE(g->N > 0 && g->E >= g->N, BG_EINVAL, "BgGraph: need N>0 and E>=N (self loops), got N=%lld E=%lld",
               (long long)g->N, (long long)g->E);
    return BG_OK;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_gat_fwd(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias,
                          float* out, float* m, float* z, int32_t C, float slope, void* stream) {
    if (int rc = check_graph(g)) return rc;
    BG_REQUIRE(h && s && d && out && m && z, BG_EINVAL, "bg_gat_fwd: null pointer");
#define CALL(CC) launch_fwd<CC>(g, h, s, d, bias, out, m, z, slope, as_stream(stream))
    BG_DISPATCH_C(C, CALL)
#undef CALL
}

extern "C" int bg_gat_bwd(const BgGraph* g, const float* gout, const float* h, const float* s, const float* d,
                          const float* m, const float* z, const float* a_src, const float* a_dst, float* P,
                          float* DU, float* gh_tot, float* gsd, int32_t C, float slope, void* stream) {
    if (int rc = check_graph(g)) return rc;
    BG_REQUIRE(gout && h && s && d && m && z && a_src && a_dst && P && DU && gh_tot && gsd, BG_EINVAL,
               "bg_gat_bwd: null pointer");
#define CALL(CC) launch_bwd<CC>(g, gout, h, s, d, m, z, a_src, a_dst, P, DU, gh_tot, gsd, slope, as_stream(stream))
    BG_DISPATCH_C(C, CALL)
#undef CALL
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
