// GraphNorm (one segment over all N rows - the reference calls it without `batch`,
// models.py:73,83,193,203) fused with ReLU and the dropout mask: forward, backward, and the
// second-order backward needed by WGAN-GP.  Every pass is "a few [C]-sized column moments over
// N rows, then one elementwise apply" (SURVEY appendix D.2).
//
// Column moments: the N rows are split into a FIXED number of row chunks (function of N only);
// each CTA reduces its chunk with a fixed in-CTA tree and publishes a partial; the CTA that draws
// the last ticket folds the partials in chunk order.  No float atomics => bitwise reproducible.
// The forward statistics use Chan's parallel (mean, M2) combination, so the variance does not
// suffer the E[x^2]-E[x]^2 cancellation.
#include "bg_common.cuh"

namespace bg {

template <int C>
struct ColMap {
    static constexpr int VEC = C >= 4 ? 4 : C;
    static constexpr int TPR = C / VEC;            // threads per row
    static constexpr int RPI = kThreads / TPR;     // rows per CTA iteration
};

// In-CTA reduction of K*VEC per-thread values over the row slots: threads that hold the same
// column group (same tid % TPR) are summed in a fixed order.  Result for column group cv lands in
// red[k*C + cv*VEC + v] (shared).  red must hold kWarps*K*C floats.
template <int C, int K>
__device__ __forceinline__ void cta_colsum(float (&val)[K * ColMap<C>::VEC], float* red) {
    using M = ColMap<C>;
    constexpr int VEC = M::VEC, TPR = M::TPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (TPR < 32) {
#pragma unroll
        for (int i = 0; i < K * VEC; ++i)
#pragma unroll
            for (int o = 16; o >= TPR; o >>= 1) val[i] += __shfl_xor_sync(0xffffffffu, val[i], o);
    }
    __syncthreads();
    if (lane < TPR) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v) red[warp * K * C + k * C + lane * VEC + v] = val[k * VEC + v];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * C; i += kThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w * K * C + i];
        red[i] = t;  // slot of warp 0 is overwritten only after every warp's value was read by this thread
    }
    __syncthreads();
}

// Column of flat element i in an [N, C] row-major tensor.  Every width the models use is a power of two: a mask instead of
// a 64-bit modulo (ncu r01h: gn_apply_kernel spent ~210 instructions per 128-bit element group, issue-bound at 48 % of the copy
// bandwidth at N = 1e6 - four runtime `% C` per group were a third of that).
__device__ __forceinline__ int col_of(int64_t i, int C) {
    return (C & (C - 1)) == 0 ? (int)(i & (int64_t)(C - 1)) : (int)(i % C);
}

// y = w*(o - alpha*mu)*rstd + beta ; x1 = relu(y) * keep / keep_prob      (flat elementwise)
// keep: explicit uint8 mask, or (keep == NULL && keep_prob < 1) a Philox Bernoulli(keep_prob) mask.
__global__ void __launch_bounds__(kThreads) gn_apply_kernel(const float* __restrict__ o, const float* __restrict__ w,
                                                            const float* __restrict__ beta, const float* __restrict__ alpha,
                                                            const float* __restrict__ stats, const uint8_t* __restrict__ keep,
                                                            float keep_prob, uint64_t seed, uint64_t offset,
                                                            const uint64_t* __restrict__ base, int64_t total, int C,
                                                            float* __restrict__ x1) {
    pdl_prologue();
    if (base) offset += *base;
    __shared__ float sc[128], sh[128];  // y = o*sc + sh
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const float mu = stats[c], r = stats[C + c];
        sc[c] = w[c] * r;
        sh[c] = beta[c] - w[c] * r * alpha[c] * mu;
    }
    __syncthreads();
    const bool fused = (keep == nullptr) && keep_prob < 1.f;
    const float keep_scale = (keep != nullptr || fused) ? 1.f / keep_prob : 1.f;
    const int64_t n4 = total / 4;
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(o) + i);
        const int c0 = col_of(i * 4, C);
        float xv[4] = {x.x, x.y, x.z, x.w}, y[4];
        bool kp[4] = {true, true, true, true};
        if (keep) {
            const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(keep) + i);
#pragma unroll
            for (int k = 0; k < 4; ++k) kp[k] = ((m >> (8 * k)) & 0xffu) != 0;
        } else if (fused) {
            const uint4 rnd = philox4x32((uint64_t)i, offset, seed);
            const uint32_t rv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) kp[k] = u01(rv[k]) <= keep_prob;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (C & 3) == 0 ? c0 + k : col_of(c0 + k, C);  // C % 4 == 0: the group never straddles a row
            float t = fmaf(xv[k], sc[c], sh[c]);
            t = t > 0.f ? t : 0.f;
            y[k] = kp[k] ? t * keep_scale : 0.f;
        }
        reinterpret_cast<float4*>(x1)[i] = make_float4(y[0], y[1], y[2], y[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (total & 3)) {
        const int64_t i = n4 * 4 + threadIdx.x;
        const int c = col_of(i, C);
        float t = fmaf(o[i], sc[c], sh[c]);
        t = t > 0.f ? t : 0.f;
        bool kp = true;
        if (keep) kp = keep[i] != 0;
        else if (fused) {
            const uint4 rnd = philox4x32((uint64_t)n4, offset, seed);
            const uint32_t rv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
            kp = u01(rv[threadIdx.x]) <= keep_prob;
        }
        x1[i] = kp ? t * keep_scale : 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// generic K-moment column reduction with a per-element functor, + last-CTA fold into sums[K*C]
// ------------------------------------------------------------------------------------------
struct BwdMoments {  // K=2: sum gy, sum gy*ohat          gy = gx1 * keep_scale*[x1>0]
    const float *gx1, *o, *x1, *alpha, *stats;
    float keep_scale;
};
struct Bwd2Moments {  // K=3: sum Xt, sum Xt*gy, sum Xt*ohat
    const float *Xt, *gx1, *o, *x1, *alpha, *stats;
    float keep_scale;
};

template <int C, int K, class F>
__device__ __forceinline__ void moments_body(const F& f, int64_t N, int G, unsigned int* counters, float* partials,
                                             float* sums_out /*shared result, K*C*/, bool& is_last) {
    using M = ColMap<C>;
    constexpr int VEC = M::VEC, TPR = M::TPR, RPI = M::RPI;
    __shared__ float red[kWarps * K * C > kThreads ? kWarps * K * C : kThreads];
    const int cv = threadIdx.x % TPR, rs = threadIdx.x / TPR;
    const int64_t chunk = ceil_div(N, G);
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(N, r0 + chunk);
    float acc[K * VEC];
#pragma unroll
    for (int i = 0; i < K * VEC; ++i) acc[i] = 0.f;
    // 4 independent row loads in flight per thread (a single dependent load per trip left the kernel latency-bound at
    // ~50% of HBM bandwidth on graphs larger than L2); the summation order per thread is unchanged
    int64_t r = r0 + rs;
    for (; r + 3 * RPI < r1; r += 4 * RPI) f.template accumulate4<C>(r, RPI, cv, acc);
    for (; r < r1; r += RPI) f.template accumulate<C>(r, cv, acc);
    cta_colsum<C, K>(acc, red);
    for (int i = threadIdx.x; i < K * C; i += kThreads) partials[(int64_t)blockIdx.x * K * C + i] = red[i];
    is_last = hier_fold(partials, partials + (int64_t)G * K * C, K * C, counters, red, sums_out);
}

// ------------------------------------------------------------------------------------------
// forward statistics: mu, rstd, var in ONE pass: sums of d = x - shift and d^2 with shift = row 0 of the
// column (a sample of the distribution, so |mean(d)| is a few sigma at most and the variance
// S2/n - (S1/n)^2 does not suffer the E[x^2]-E[x]^2 cancellation of unshifted sums).
// ------------------------------------------------------------------------------------------
struct StatsAcc {
    const float* o;
    template <int CC>
    __device__ __forceinline__ void consume(const Vec<ColMap<CC>::VEC>& x, const Vec<ColMap<CC>::VEC>& sh, float* acc) const {
        constexpr int VEC = ColMap<CC>::VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float d = x.v[v] - sh.v[v];
            acc[v] += d;
            acc[VEC + v] = fmaf(d, d, acc[VEC + v]);
        }
    }
    template <int CC>
    __device__ __forceinline__ void accumulate(int64_t r, int cv, float* acc) const {
        constexpr int VEC = ColMap<CC>::VEC;
        Vec<VEC> x, sh;
        x.load(o + r * CC + cv * VEC);
        sh.load(o + cv * VEC);
        consume<CC>(x, sh, acc);
    }
    template <int CC>
    __device__ __forceinline__ void accumulate4(int64_t r, int64_t step, int cv, float* acc) const {
        constexpr int VEC = ColMap<CC>::VEC;
        Vec<VEC> x[4], sh;
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q].load(o + (r + q * step) * CC + cv * VEC);
        sh.load(o + cv * VEC);
#pragma unroll
        for (int q = 0; q < 4; ++q) consume<CC>(x[q], sh, acc);
    }
};

template <int C>
__global__ void __launch_bounds__(kThreads) gn_stats_kernel(const float* __restrict__ o, const float* __restrict__ alpha,
                                                            int64_t N, int G, float eps, float* __restrict__ stats,
                                                            unsigned int* counters, float* partials) {
    pdl_prologue();
    __shared__ float sums[2 * C];
    bool is_last;
    StatsAcc f{o};
    moments_body<C, 2>(f, N, G, counters, partials, sums, is_last);
    if (!is_last) return;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        const float n = (float)N;
        const float md = sums[c] / n;
        const float mean = o[c] + md;
        const float M2 = fmaxf(sums[C + c] - sums[c] * md, 0.f);
        const float shift = mean * (1.f - __ldg(alpha + c));  // mean of (o - alpha*mu)
        const float var = (M2 + n * shift * shift) / n;
        stats[c] = mean;
        stats[C + c] = 1.f / sqrtf(var + eps);
        stats[2 * C + c] = var;
    }
}

template <int C>
struct BwdAcc {
    BwdMoments p;
    template <int CC>
    __device__ __forceinline__ void consume(const Vec<ColMap<CC>::VEC>& g, const Vec<ColMap<CC>::VEC>& ov,
                                            const Vec<ColMap<CC>::VEC>& xv, int cv, float* acc) const {
        constexpr int VEC = ColMap<CC>::VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int c = cv * VEC + v;
            const float gy = xv.v[v] > 0.f ? g.v[v] * p.keep_scale : 0.f;
            const float oh = ov.v[v] - __ldg(p.alpha + c) * __ldg(p.stats + c);
            acc[v] += gy;
            acc[VEC + v] = fmaf(gy, oh, acc[VEC + v]);
        }
    }
    template <int CC>
    __device__ __forceinline__ void accumulate(int64_t r, int cv, float* acc) const {
        constexpr int VEC = ColMap<CC>::VEC;
        Vec<VEC> g, ov, xv;
        const int64_t off = r * CC + cv * VEC;
        g.load(p.gx1 + off);
        ov.load(p.o + off);
        xv.load(p.x1 + off);
        consume<CC>(g, ov, xv, cv, acc);
    }
    template <int CC>
    __device__ __forceinline__ void accumulate4(int64_t r, int64_t step, int cv, float* acc) const {
        constexpr int VEC = ColMap<CC>::VEC;
        Vec<VEC> g[4], ov[4], xv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t off = (r + q * step) * CC + cv * VEC;
            g[q].load(p.gx1 + off);
            ov[q].load(p.o + off);
            xv[q].load(p.x1 + off);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) consume<CC>(g[q], ov[q], xv[q], cv, acc);
    }
};

// backward moments + parameter gradients.  bstats[2C] = (G0, G1, means); dparams[3C] = (dw, dbeta, dalpha)
template <int C>
__global__ void __launch_bounds__(kThreads) gn_bwd_moments_kernel(BwdMoments p, const float* __restrict__ w, int64_t N, int G,
                                                                  float* dparams, int accumulate, float* bstats,
                                                                  unsigned int* counter, float* partials) {
    pdl_prologue();
    __shared__ float sums[2 * C];
    bool is_last;
    BwdAcc<C> f{p};
    moments_body<C, 2>(f, N, G, counter, partials, sums, is_last);
    if (!is_last) return;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        const float n = (float)N;
        const float G0 = sums[c] / n, G1 = sums[C + c] / n;
        const float mu = p.stats[c], r = p.stats[C + c], a = p.alpha[c], wc = w[c];
        const float mean_ohat = wc * r * G0 - wc * r * r * r * G1 * mu * (1.f - a);  // M[d loss/d ohat]
        bstats[c] = G0;
        bstats[C + c] = G1;
        const float dw = n * r * G1, db = n * G0, da = -mu * n * mean_ohat;
        if (accumulate) {
            dparams[c] += dw;
            dparams[C + c] += db;
            dparams[2 * C + c] += da;
        } else {
            dparams[c] = dw;
            dparams[C + c] = db;
            dparams[2 * C + c] = da;
        }
    }
}

// go = w r gy - w r^3 G1 ohat - alpha * M[ohat_hat]
__global__ void __launch_bounds__(kThreads) gn_bwd_apply_kernel(BwdMoments p, const float* __restrict__ w,
                                                                const float* __restrict__ bstats, int64_t total, int C,
                                                                float* __restrict__ go) {
    pdl_prologue();
    __shared__ float k1[128], k2[128], k3[128], sh[128];  // go = gy*k1 - ohat*k2 - k3 ; ohat = o - sh
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const float mu = p.stats[c], r = p.stats[C + c], a = p.alpha[c], wc = w[c];
        const float G0 = bstats[c], G1 = bstats[C + c];
        k1[c] = wc * r;
        k2[c] = wc * r * r * r * G1;
        k3[c] = a * (wc * r * G0 - wc * r * r * r * G1 * mu * (1.f - a));
        sh[c] = a * mu;
    }
    __syncthreads();
    const int64_t n4 = total / 4;  // 128-bit loads/stores (C >= 4: a float4 never straddles a row)
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    auto one4 = [&](int64_t i, const float4& xa, const float4& ga, const float4& oa) {
        const int c0 = col_of(i * 4, C);
        const float xv[4] = {xa.x, xa.y, xa.z, xa.w}, gv[4] = {ga.x, ga.y, ga.z, ga.w}, ov[4] = {oa.x, oa.y, oa.z, oa.w};
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (C & 3) == 0 ? c0 + k : col_of(c0 + k, C);
            const float gy = xv[k] > 0.f ? gv[k] * p.keep_scale : 0.f;
            r[k] = gy * k1[c] - (ov[k] - sh[c]) * k2[c] - k3[c];
        }
        reinterpret_cast<float4*>(go)[i] = make_float4(r[0], r[1], r[2], r[3]);
    };
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // 12 independent 128-bit loads in flight per thread
        float4 xa[4], ga[4], oa[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            xa[q] = __ldg(reinterpret_cast<const float4*>(p.x1) + i + q * stride);
            ga[q] = __ldg(reinterpret_cast<const float4*>(p.gx1) + i + q * stride);
            oa[q] = __ldg(reinterpret_cast<const float4*>(p.o) + i + q * stride);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) one4(i + q * stride, xa[q], ga[q], oa[q]);
    }
    for (; i < n4; i += stride)
        one4(i, __ldg(reinterpret_cast<const float4*>(p.x1) + i), __ldg(reinterpret_cast<const float4*>(p.gx1) + i),
             __ldg(reinterpret_cast<const float4*>(p.o) + i));
    if (blockIdx.x == 0 && threadIdx.x < (total & 3)) {
        const int64_t t = n4 * 4 + threadIdx.x;
        const int c = col_of(t, C);
        const float gy = p.x1[t] > 0.f ? p.gx1[t] * p.keep_scale : 0.f;
        go[t] = gy * k1[c] - (p.o[t] - sh[c]) * k2[c] - k3[c];
    }
}

template <int C>
struct Bwd2Acc {
    Bwd2Moments p;
    template <int CC>
    __device__ __forceinline__ void accumulate(int64_t r, int cv, float* acc) const {
        constexpr int VEC = ColMap<CC>::VEC;
        Vec<VEC> xt, g, ov, xv;
        const int64_t off = r * CC + cv * VEC;
        xt.load(p.Xt + off);
        g.load(p.gx1 + off);
        ov.load(p.o + off);
        xv.load(p.x1 + off);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int c = cv * VEC + v;
            const float gy = xv.v[v] > 0.f ? g.v[v] * p.keep_scale : 0.f;
            const float oh = ov.v[v] - __ldg(p.alpha + c) * __ldg(p.stats + c);
            acc[v] += xt.v[v];
            acc[VEC + v] = fmaf(xt.v[v], gy, acc[VEC + v]);
            acc[2 * VEC + v] = fmaf(xt.v[v], oh, acc[2 * VEC + v]);
        }
    }
    template <int CC>
    __device__ __forceinline__ void accumulate4(int64_t r, int64_t step, int cv, float* acc) const {
#pragma unroll
        for (int q = 0; q < 4; ++q) accumulate<CC>(r + q * step, cv, acc);
    }
};

// second-order moments.  b2[4C] = (A0, P, Q, MPhi) ; dparams2[3C] = cotangents on (w, beta=0, alpha)
template <int C>
__global__ void __launch_bounds__(kThreads) gn_bwd2_moments_kernel(Bwd2Moments p, const float* __restrict__ w,
                                                                   const float* __restrict__ bstats, int64_t N, int G,
                                                                   float* dparams2, int accumulate, float* b2,
                                                                   unsigned int* counter, float* partials) {
    pdl_prologue();
    __shared__ float sums[3 * C];
    bool is_last;
    Bwd2Acc<C> f{p};
    moments_body<C, 3>(f, N, G, counter, partials, sums, is_last);
    if (!is_last) return;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        const float n = (float)N;
        const float A0 = sums[c] / n, A1 = sums[C + c] / n, A2 = sums[2 * C + c] / n;
        const float mu = p.stats[c], r = p.stats[C + c], a = p.alpha[c], wc = w[c];
        const float G0 = bstats[c], G1 = bstats[C + c];
        const float mo = mu * (1.f - a);  // M[ohat]
        const float r3 = r * r * r;
        const float Pm = A1 - a * A0 * G0;
        const float Qm = A2 - a * A0 * mo;
        const float coef = Pm - 3.f * r * r * G1 * Qm;
        const float MPhi = -wc * r3 * (Qm * G0 + G1 * A0 * (1.f - a) + coef * mo);
        const float mean_ohat = wc * r * G0 - wc * r3 * G1 * mo;
        b2[c] = A0;
        b2[C + c] = Pm;
        b2[2 * C + c] = Qm;
        b2[3 * C + c] = MPhi;
        const float wt = n * (r * Pm - r3 * G1 * Qm);
        const float at = -n * A0 * mean_ohat - mu * n * MPhi;
        if (accumulate) {
            dparams2[c] += wt;
            dparams2[2 * C + c] += at;
        } else {
            dparams2[c] = wt;
            dparams2[C + c] = 0.f;
            dparams2[2 * C + c] = at;
        }
    }
}

// gx1t = w (r Ot - r^3 Q ohat) * mask ;  ot = Phi - alpha*MPhi,
// Phi = -w r^3 (Q gy + G1 Ot + (P - 3 r^2 G1 Q) ohat),  Ot = Xt - alpha*A0
__global__ void __launch_bounds__(kThreads) gn_bwd2_apply_kernel(Bwd2Moments p, const float* __restrict__ w,
                                                                 const float* __restrict__ bstats, const float* __restrict__ b2,
                                                                 int64_t total, int C, float* __restrict__ gx1t,
                                                                 float* __restrict__ ot) {
    pdl_prologue();
    __shared__ float wr[128], wr3[128], Qs[128], G1s[128], cf[128], sh[128], aA0[128], aMPhi[128];
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const float mu = p.stats[c], r = p.stats[C + c], a = p.alpha[c], wc = w[c];
        const float G1 = bstats[C + c];
        const float A0 = b2[c], Pm = b2[C + c], Qm = b2[2 * C + c], MPhi = b2[3 * C + c];
        wr[c] = wc * r;
        wr3[c] = wc * r * r * r;
        Qs[c] = Qm;
        G1s[c] = G1;
        cf[c] = Pm - 3.f * r * r * G1 * Qm;
        sh[c] = a * mu;
        aA0[c] = a * A0;
        aMPhi[c] = a * MPhi;
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int c = col_of(i, C);
        const bool on = p.x1[i] > 0.f;
        const float gy = on ? p.gx1[i] * p.keep_scale : 0.f;
        const float oh = p.o[i] - sh[c];
        const float Ot = p.Xt[i] - aA0[c];
        const float gyt = wr[c] * Ot - wr3[c] * Qs[c] * oh;
        gx1t[i] = on ? gyt * p.keep_scale : 0.f;
        const float Phi = -wr3[c] * (Qs[c] * gy + G1s[c] * Ot + cf[c] * oh);
        ot[i] = Phi - aMPhi[c];
    }
}

static inline int flat_grid(int64_t total) {
    int64_t g = ceil_div(total, (int64_t)kThreads * 4);
    if (g < 1) g = 1;
    if (g > 8 * kSMs) g = 8 * kSMs;
    return (int)g;
}

template <int C>
static int gn_splits(int64_t N) { return reduce_splits(N, ColMap<C>::RPI); }

}  // namespace bg

using namespace bg;

extern "C" size_t bg_graphnorm_ws(int64_t N, int32_t C) {
    (void)N;
    return (size_t)kCounterBytes + (size_t)(2 * kSMs + 2 * kSMs / kFoldGroup + 2) * 3 * (size_t)C * sizeof(float) + 4 * (size_t)C * sizeof(float);
}

#define BG_GN_DISPATCH(C, CALL)                                                                    \
    switch (C) {                                                                                   \
        case 1: CALL(1); break;                                                                    \
        case 2: CALL(2); break;                                                                    \
        case 4: CALL(4); break;                                                                    \
        case 8: CALL(8); break;                                                                    \
        case 16: CALL(16); break;                                                                  \
        case 32: CALL(32); break;                                                                  \
        case 64: CALL(64); break;                                                                  \
        case 128: CALL(128); break;                                                                \
        default:                                                                                   \
            bg::set_error("unsupported channel width C=%d (supported: 1,2,4,...,128)", (int)(C)); \
            return BG_EUNSUPPORTED;                                                                \
    }

extern "C" int bg_graphnorm_fwd(const float* o, const float* w, const float* beta, const float* alpha,
                                const uint8_t* keep, float keep_prob, uint64_t seed, uint64_t offset, int64_t N, int32_t C,
                                float eps, float* x1, float* stats, float* workspace, size_t ws_bytes, void* stream) {
    BG_REQUIRE(o && w && beta && alpha && x1 && stats && workspace, BG_EINVAL, "bg_graphnorm_fwd: null pointer");
    BG_REQUIRE(N > 0, BG_EINVAL, "bg_graphnorm_fwd: N must be > 0");
    BG_REQUIRE(keep_prob > 0.f && keep_prob <= 1.f, BG_EINVAL, "bg_graphnorm_fwd: keep_prob must be in (0,1]");
    BG_REQUIRE(ws_bytes >= bg_graphnorm_ws(N, C), BG_EINVAL, "bg_graphnorm_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    float* partials = workspace + kCounterBytes / sizeof(float);
#define CALL(CC)                                                                                            \
    {                                                                                                       \
        const int G = gn_splits<CC>(N);                                                                     \
        launch_k(gn_stats_kernel<CC>, G, kThreads, 0, st, o, alpha, N, G, eps, stats, counter, partials);         \
    }
    BG_GN_DISPATCH(C, CALL)
#undef CALL
    const int64_t total = N * C;
    launch_k(gn_apply_kernel, flat_grid(total), kThreads, 0, st, o, w, beta, alpha, stats, keep, keep_prob, seed, offset, rng_base(), total, C, x1);
    return check_launch("bg_graphnorm_fwd");
}

// The apply half alone: statistics come from bg_gat_fwd_gn (fused into the aggregation) or any other producer.
extern "C" int bg_graphnorm_apply(const float* o, const float* w, const float* beta, const float* alpha, const float* stats,
                                  const uint8_t* keep, float keep_prob, uint64_t seed, uint64_t offset, int64_t N, int32_t C,
                                  float* x1, void* stream) {
    BG_REQUIRE(o && w && beta && alpha && stats && x1, BG_EINVAL, "bg_graphnorm_apply: null pointer");
    BG_REQUIRE(N > 0 && C >= 1 && C <= 128, BG_EINVAL, "bg_graphnorm_apply: bad shape");
    BG_REQUIRE(keep_prob > 0.f && keep_prob <= 1.f, BG_EINVAL, "bg_graphnorm_apply: keep_prob must be in (0,1]");
    const int64_t total = N * C;
    launch_k(gn_apply_kernel, flat_grid(total), kThreads, 0, as_stream(stream), o, w, beta, alpha, stats, keep, keep_prob, seed, offset, rng_base(),
                                                                          total, C, x1);
    return check_launch("bg_graphnorm_apply");
}

extern "C" int bg_graphnorm_bwd(const float* gx1, const float* o, const float* x1, const float* w, const float* alpha,
                                const float* stats, float keep_scale, int64_t N, int32_t C, float* go, float* dparams,
                                int32_t accumulate, float* bstats, float* workspace, size_t ws_bytes, void* stream) {
    BG_REQUIRE(gx1 && o && x1 && w && alpha && stats && go && dparams && bstats && workspace, BG_EINVAL,
               "bg_graphnorm_bwd: null pointer");
    BG_REQUIRE(ws_bytes >= bg_graphnorm_ws(N, C), BG_EINVAL, "bg_graphnorm_bwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    float* partials = workspace + kCounterBytes / sizeof(float);
    BwdMoments p{gx1, o, x1, alpha, stats, keep_scale};
#define CALL(CC)                                                                                                     \
    {                                                                                                                \
        const int G = gn_splits<CC>(N);                                                                              \
        launch_k(gn_bwd_moments_kernel<CC>, G, kThreads, 0, st, p, w, N, G, dparams, accumulate, bstats, counter, partials); \
    }
    BG_GN_DISPATCH(C, CALL)
#undef CALL
    const int64_t total = N * C;
    launch_k(gn_bwd_apply_kernel, flat_grid(total), kThreads, 0, st, p, w, bstats, total, C, go);
    return check_launch("bg_graphnorm_bwd");
}

// The moments half of bg_graphnorm_bwd alone (parameter gradients + bstats): the elementwise half is fused into the
// aggregation backward (bg_gat_bwd_gn).
extern "C" int bg_graphnorm_bwd_moments(const float* gx1, const float* o, const float* x1, const float* w, const float* alpha,
                                        const float* stats, float keep_scale, int64_t N, int32_t C, float* dparams,
                                        int32_t accumulate, float* bstats, float* workspace, size_t ws_bytes, void* stream) {
    BG_REQUIRE(gx1 && o && x1 && w && alpha && stats && dparams && bstats && workspace, BG_EINVAL,
               "bg_graphnorm_bwd_moments: null pointer");
    BG_REQUIRE(ws_bytes >= bg_graphnorm_ws(N, C), BG_EINVAL, "bg_graphnorm_bwd_moments: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    float* partials = workspace + kCounterBytes / sizeof(float);
    BwdMoments p{gx1, o, x1, alpha, stats, keep_scale};
#define CALL(CC)                                                                                                     \
    {                                                                                                                \
        const int G = gn_splits<CC>(N);                                                                              \
        launch_k(gn_bwd_moments_kernel<CC>, G, kThreads, 0, st, p, w, N, G, dparams, accumulate, bstats, counter, partials); \
    }
    BG_GN_DISPATCH(C, CALL)
#undef CALL
    return check_launch("bg_graphnorm_bwd_moments");
}

extern "C" int bg_graphnorm_bwd2(const float* Xt, const float* gx1, const float* o, const float* x1, const float* w,
                                 const float* alpha, const float* stats, const float* bstats, float keep_scale, int64_t N,
                                 int32_t C, float* gx1t, float* ot, float* dparams2, int32_t accumulate, float* workspace,
                                 size_t ws_bytes, void* stream) {
    BG_REQUIRE(Xt && gx1 && o && x1 && w && alpha && stats && bstats && gx1t && ot && dparams2 && workspace, BG_EINVAL,
               "bg_graphnorm_bwd2: null pointer");
    BG_REQUIRE(ws_bytes >= bg_graphnorm_ws(N, C), BG_EINVAL, "bg_graphnorm_bwd2: workspace too small");
    cudaStream_t st = as_stream(stream);
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    float* b2 = workspace + kCounterBytes / sizeof(float);  // 4*C floats
    float* partials = b2 + 4 * (size_t)C;
    Bwd2Moments p{Xt, gx1, o, x1, alpha, stats, keep_scale};
#define CALL(CC)                                                                                                        \
    {                                                                                                                   \
        const int G = gn_splits<CC>(N);                                                                                 \
        launch_k(gn_bwd2_moments_kernel<CC>, G, kThreads, 0, st, p, w, bstats, N, G, dparams2, accumulate, b2, counter, partials); \
    }
    BG_GN_DISPATCH(C, CALL)
#undef CALL
    const int64_t total = N * C;
    launch_k(gn_bwd2_apply_kernel, flat_grid(total), kThreads, 0, st, p, w, bstats, b2, total, C, gx1t, ot);
    return check_launch("bg_graphnorm_bwd2");
}
