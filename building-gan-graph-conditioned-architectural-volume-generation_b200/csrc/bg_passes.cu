// Whole-pass executors: one C call = one forward / backward / second-order-backward pass of the
// generator or the discriminator (reference models.py:119-155 and :229-245 plus the derivative structure
// trainer.py:291-385 needs).  A pass is a fixed sequence of kernel launches on the caller's stream: no
// allocation (activations live in a caller-owned workspace carved by a deterministic bump layout that
// the backward pass re-derives), no host sync, no data-dependent control flow => graph-capturable, and
// the host cost per launch is the bare cudaLaunchKernel.
//
// Parameter order = torch's named_parameters() order of the reference modules:
//   Linear+LayerNorm layer: W, b, gamma, beta        bare Linear: W, b
//   conv block: att_src, att_dst, bias, lin.weight, gn.weight, gn.bias, gn.mean_scale
#include <stdlib.h>
#include <vector>

#include "bg_common.cuh"

namespace bg {

constexpr int kMaxLayers = 8;
constexpr int kMaxConvs = 16;
constexpr size_t kWgradOffsetBytes = 4u << 20;  // reduction workspace: [0,4 MiB) tickets + small partials, rest = wgrad partials
constexpr float kKeepProb = 0.8f;  // nn.Dropout(0.2), hard-coded in the reference (models.py:75,85,195,205)

struct Arena {
    char* base;
    size_t off;
    size_t cap = ~(size_t)0;  // bytes; an allocation past it sets `overflow` and returns the arena base (in bounds)
    bool overflow = false;
    float* f(size_t n) {
        const size_t a = align_up(off, 256);
        if (a + n * sizeof(float) > cap) {
            overflow = true;
            return reinterpret_cast<float*>(base);
        }
        off = a + n * sizeof(float);
        return base ? reinterpret_cast<float*>(base + a) : nullptr;
    }
};

struct DenseL {  // one Linear (+LayerNorm) (+activation)
    int pW, pb, pg, pbeta;  // parameter indices (-1: absent)
    int cin, cout, act;
    float *out, *xhat, *rstd;  // saved by forward
    float* b_gz;               // saved by the first backward for the second-order sweep
};
struct ConvL {
    int p_as, p_ad, p_bias, p_W, p_gw, p_gb, p_ga;
    int cin, cout;
    float *h, *s, *d, *o, *m, *z, *x1, *stats;
    float *b_gx1, *b_go, *b_gh, *b_gsd, *b_bstats;
    // set by the block ABOVE when its backward-input product already computed this block's GraphNorm backward moments
    // (dense_fwd_moments): the buffers conv_backward must use instead of launching bg_graphnorm_bwd_moments
    float* pre_bst = nullptr;
    float* pre_gnpar = nullptr;
};

static int conv_widths(int hidden, int repeat, int* w) {  // hourglass: halve `repeat` times then double back
    int n = 0;
    for (int k = 0; k <= repeat; ++k) w[n++] = hidden >> k;
    for (int k = repeat - 1; k >= 0; --k) w[n++] = hidden >> k;
    return n;
}

struct GenNet {
    int n_menc, n_mlp, n_conv, n_dec, nparams;
    DenseL menc[kMaxLayers], mlp[kMaxLayers], dec[kMaxLayers];
    ConvL conv[kMaxConvs];
    float* soft;
};
struct DiscNet {
    int n_conv, nparams;
    DenseL pre[2], dec[4];
    ConvL conv[kMaxConvs];
};

static void add_ln_layer(DenseL& l, int& pi, int cin, int cout) {
    l = DenseL{pi, pi + 1, pi + 2, pi + 3, cin, cout, BG_ACT_LRELU, nullptr, nullptr, nullptr, nullptr};
    pi += 4;
}
static void add_plain_layer(DenseL& l, int& pi, int cin, int cout, int act) {
    l = DenseL{pi, pi + 1, -1, -1, cin, cout, act, nullptr, nullptr, nullptr, nullptr};
    pi += 2;
}
static void add_conv(ConvL& c, int& pi, int cin, int cout) {
    c = ConvL{};
    c.p_as = pi; c.p_ad = pi + 1; c.p_bias = pi + 2; c.p_W = pi + 3; c.p_gw = pi + 4; c.p_gb = pi + 5; c.p_ga = pi + 6;
    c.cin = cin; c.cout = cout;
    pi += 7;
}

static int build_gen(const BgModelDesc& md, GenNet& g) {
    BG_REQUIRE(md.le_layers >= 1 && md.le_layers <= kMaxLayers && md.g_mlp_layers >= 1 && md.g_mlp_layers <= kMaxLayers,
               BG_EUNSUPPORTED, "generator: MLP depth out of range");
    BG_REQUIRE(2 * md.g_repeat <= kMaxConvs && md.g_repeat >= 1, BG_EUNSUPPORTED, "generator: encoder repeat out of range");
    int pi = 0;
    g.n_menc = md.le_layers;
    for (int i = 0; i < g.n_menc; ++i) add_ln_layer(g.menc[i], pi, i == 0 ? md.local_dim : md.le_dim, md.le_dim);
    g.n_mlp = md.g_mlp_layers;
    for (int i = 0; i < g.n_mlp; ++i)
        add_ln_layer(g.mlp[i], pi, i == 0 ? md.le_dim + md.voxel_dim + md.z_dim : md.g_hidden, md.g_hidden);
    int w[2 * kMaxConvs];
    const int nw = conv_widths(md.g_hidden, md.g_repeat, w);
    g.n_conv = nw - 1;
    for (int k = 0; k < g.n_conv; ++k) add_conv(g.conv[k], pi, w[k], w[k + 1]);
    const int gh = md.g_hidden;
    g.n_dec = 5;
    add_ln_layer(g.dec[0], pi, md.le_dim + md.voxel_dim + md.z_dim + w[nw - 1] + gh, gh);
    add_ln_layer(g.dec[1], pi, gh, gh / 2);
    add_ln_layer(g.dec[2], pi, gh / 2, gh / 4);
    add_ln_layer(g.dec[3], pi, gh / 4, gh / 8);
    add_plain_layer(g.dec[4], pi, gh / 8, md.num_classes, BG_ACT_NONE);
    g.nparams = pi;
    return BG_OK;
}

static int build_disc(const BgModelDesc& md, DiscNet& d) {
    BG_REQUIRE(2 * md.d_repeat <= kMaxConvs && md.d_repeat >= 1, BG_EUNSUPPORTED, "discriminator: encoder repeat out of range");
    int pi = 0;
    const int dh = md.d_hidden;
    add_plain_layer(d.pre[0], pi, md.local_dim + md.voxel_dim + md.num_classes, dh, BG_ACT_RELU);
    add_plain_layer(d.pre[1], pi, dh, dh, BG_ACT_RELU);
    int w[2 * kMaxConvs];
    const int nw = conv_widths(dh, md.d_repeat, w);
    d.n_conv = nw - 1;
    for (int k = 0; k < d.n_conv; ++k) add_conv(d.conv[k], pi, w[k], w[k + 1]);
    add_plain_layer(d.dec[0], pi, dh, dh / 2, BG_ACT_RELU);
    add_plain_layer(d.dec[1], pi, dh / 2, dh / 4, BG_ACT_RELU);
    add_plain_layer(d.dec[2], pi, dh / 4, dh / 8, BG_ACT_RELU);
    add_plain_layer(d.dec[3], pi, dh / 8, 1, BG_ACT_NONE);
    d.nparams = pi;
    return BG_OK;
}

// ---- deterministic workspace layouts (forward-saved tensors); the backward re-derives the same pointers
static void layout_dense(DenseL& l, int64_t rows, Arena& A, bool with_ln_saves) {
    l.out = A.f((size_t)rows * l.cout);
    if (l.pg >= 0 && with_ln_saves) {
        l.xhat = A.f((size_t)rows * l.cout);
        l.rstd = A.f((size_t)rows);
    }
}
static void layout_conv(ConvL& c, int64_t N, Arena& A) {
    c.h = A.f((size_t)N * c.cout);
    c.s = A.f((size_t)N);
    c.d = A.f((size_t)N);
    c.o = A.f((size_t)N * c.cout);
    c.m = A.f((size_t)N);
    c.z = A.f((size_t)N);
    c.x1 = A.f((size_t)N * c.cout);
    c.stats = A.f((size_t)3 * c.cout);
}
static void layout_gen(const BgModelDesc& md, GenNet& g, int64_t N, Arena& A) {
    for (int i = 0; i < g.n_menc; ++i) layout_dense(g.menc[i], md.num_classes, A, true);
    for (int i = 0; i < g.n_mlp; ++i) layout_dense(g.mlp[i], N, A, true);
    for (int k = 0; k < g.n_conv; ++k) layout_conv(g.conv[k], N, A);
    for (int i = 0; i < g.n_dec; ++i) layout_dense(g.dec[i], N, A, true);
    g.soft = nullptr;  // label_soft is an output tensor of the pass
}
static void layout_disc(DiscNet& d, int64_t N, Arena& A) {
    for (int i = 0; i < 2; ++i) layout_dense(d.pre[i], N, A, false);
    for (int k = 0; k < d.n_conv; ++k) layout_conv(d.conv[k], N, A);
    for (int i = 0; i < 4; ++i) layout_dense(d.dec[i], N, A, false);
}
// tensors the discriminator's first backward keeps for the second-order sweep
static void layout_disc_bwd_saved(DiscNet& d, int64_t N, Arena& A) {
    for (int i = 0; i < 2; ++i) d.pre[i].b_gz = A.f((size_t)N * d.pre[i].cout);
    for (int i = 0; i < 4; ++i) d.dec[i].b_gz = A.f((size_t)N * d.dec[i].cout);
    for (int k = 0; k < d.n_conv; ++k) {
        ConvL& c = d.conv[k];
        c.b_gx1 = A.f((size_t)N * c.cout);
        c.b_go = A.f((size_t)N * c.cout);
        c.b_gh = A.f((size_t)N * c.cout);
        c.b_gsd = A.f((size_t)N * 2);
        c.b_bstats = A.f((size_t)2 * c.cout);
    }
}

// ---- shared per-call context
struct Ctx {
    const float* const* P;   // parameters
    float* G;                // flat grad bucket (may be null)
    const int64_t* goff;     // offsets of each parameter's gradient in G
    const BgGraph* graph;
    int64_t N;
    float* red;              // reduction workspace (counters + partials)
    size_t red_bytes;
    void* st;
    int accumulate;
    WgradQueue* q;           // deferred weight gradients of this pass (launched and folded once, at the end)
    Arena* X;                // per-layer scratch (per-edge arrays); everything a weight gradient reads lives in the pass arena
    float* g(int pidx) const { return (G && pidx >= 0) ? G + goff[pidx] : nullptr; }
};

static int init_queue(WgradQueue& q, float* red, size_t red_bytes) {
    BG_REQUIRE(red_bytes >= 2 * kWgradOffsetBytes, BG_EINVAL, "reduction workspace too small (%zu bytes, need >= %zu)", red_bytes,
               2 * kWgradOffsetBytes);
    q.buf = red + kWgradOffsetBytes / sizeof(float);
    q.cap = (red_bytes - kWgradOffsetBytes) / sizeof(float);
    q.used = 0;
    q.defer = true;  // operands of every weight gradient stay alive until wgrad_flush() at the end of the pass
    return BG_OK;
}

static inline BgSeg seg(const float* p, int width, int ld, const int32_t* gather = nullptr) { return BgSeg{p, gather, width, ld}; }

// The type-matched encoder runs on the K table rows: K <= 8 => the whole chain is one single-CTA launch per direction
// (bg_smallmlp.cu) instead of a launch (backward: three) per layer.  BG_SMALL_MLP=0 keeps the per-layer launches (A/B switch).
static bool small_chain_ok(const DenseL* l, int n, int rows) {
    static const bool on = !(getenv("BG_SMALL_MLP") && atoi(getenv("BG_SMALL_MLP")) == 0);
    if (!on || n < 1 || n > BG_SMALL_MAX_LAYERS || rows < 1 || rows > BG_SMALL_MAX_ROWS) return false;
    for (int i = 0; i < n; ++i)
        if (l[i].cin > 128 || l[i].cout > 128 || (i > 0 && l[i].cin != l[i - 1].cout)) return false;
    return true;
}
static void small_chain_fill(const Ctx& c, const DenseL* l, int n, BgSmallLayer* out, bool grads) {
    for (int i = 0; i < n; ++i) {
        BgSmallLayer& a = out[i];
        a = BgSmallLayer{};
        a.W = c.P[l[i].pW];
        a.bias = l[i].pb >= 0 ? c.P[l[i].pb] : nullptr;
        if (l[i].pg >= 0) { a.gamma = c.P[l[i].pg]; a.beta = c.P[l[i].pbeta]; a.xhat = l[i].xhat; a.rstd = l[i].rstd; }
        a.cin = l[i].cin; a.cout = l[i].cout; a.act = l[i].act;
        a.out = l[i].out;
        if (grads && c.G) {
            a.dW = c.g(l[i].pW);
            a.dbias = c.g(l[i].pb);
            a.dgamma = c.g(l[i].pg);
            a.dbeta = c.g(l[i].pbeta);
        }
    }
}

static int dense_fwd_call(const Ctx& c, const DenseL& l, int64_t rows, const BgSeg* segs, int nseg, bool save_ln) {
    BgDense a{};
    a.N = rows; a.nseg = nseg;
    for (int i = 0; i < nseg; ++i) a.seg[i] = segs[i];
    a.W = c.P[l.pW]; a.w_so = l.cin; a.w_sk = 1; a.Cout = l.cout;
    a.bias = c.P[l.pb];
    if (l.pg >= 0) { a.ln_gamma = c.P[l.pg]; a.ln_beta = c.P[l.pbeta]; }
    a.act = l.act;
    a.out = l.out; a.ld_out = l.cout;
    if (l.pg >= 0 && save_ln) { a.xhat = l.xhat; a.rstd = l.rstd; }
    return bg_dense_fwd(&a, c.st);
}

// out[rows, hi-lo] = X[rows, Kx] @ W[:, lo:hi]   (W is [Kx, ldw] row-major: the backward-input product)
// `gate` (optional, [rows, hi-lo]): the ReLU backward of the layer below fused into the epilogue (out *= gate > 0)
static int matmul_nn(const Ctx& c, const float* X, int64_t rows, int Kx, const float* W, int ldw, int lo, int hi, float* out,
                     const float* gate = nullptr) {
    BgDense a{};
    a.N = rows; a.nseg = 1; a.seg[0] = seg(X, Kx, Kx);
    a.W = W + lo; a.w_so = 1; a.w_sk = ldw; a.Cout = hi - lo;
    a.act = BG_ACT_NONE; a.out = out; a.ld_out = hi - lo;
    a.gate = gate; a.ld_gate = hi - lo; a.gate_slope = 0.f;
    return bg_dense_fwd(&a, c.st);
}
// out[rows, Cout] = X[rows, hi-lo] @ W[:, lo:hi]^T   (W is [Cout, ldw] row-major)
static int matmul_nt(const Ctx& c, const float* X, int64_t rows, const float* W, int Cout, int ldw, int lo, int hi, float* out,
                     const float* a_src = nullptr, const float* a_dst = nullptr, float* s = nullptr, float* d = nullptr,
                     const float* gate = nullptr) {
    BgDense a{};
    a.N = rows; a.nseg = 1; a.seg[0] = seg(X, hi - lo, hi - lo);
    a.W = W + lo; a.w_so = ldw; a.w_sk = 1; a.Cout = Cout;
    a.act = BG_ACT_NONE; a.out = out; a.ld_out = Cout;
    a.att_src = a_src; a.att_dst = a_dst; a.s = s; a.d = d;
    a.gate = gate; a.ld_gate = Cout; a.gate_slope = 0.f;
    return bg_dense_fwd(&a, c.st);
}

static BgWgrad wg(int64_t N, const float* gz, int ld_gz, int Cout, const BgSeg* segs, int nseg, float* dW, int ld_dw,
                  float* dbias, int accumulate) {
    BgWgrad w{};
    w.N = N; w.gz = gz; w.ld_gz = ld_gz; w.Cout = Cout; w.nseg = nseg;
    for (int i = 0; i < nseg; ++i) w.seg[i] = segs[i];
    w.dW = dW; w.ld_dw = ld_dw; w.dbias = dbias; w.accumulate = accumulate;
    return w;
}

#define BG_TRY(expr)              \
    do {                          \
        if (int _rc = (expr)) return _rc; \
    } while (0)

// ---- one conv block --------------------------------------------------------------------------------
static int conv_forward(const Ctx& c, const ConvL& L, const float* x, const uint8_t* keep, float keep_prob, uint64_t seed,
                        uint64_t offset) {
    BG_TRY(matmul_nt(c, x, c.N, c.P[L.p_W], L.cout, L.cin, 0, L.cin, L.h, c.P[L.p_as], c.P[L.p_ad], L.s, L.d));
    // aggregation with the GraphNorm statistics fused into its epilogue, then ONE elementwise pass
    BG_TRY(bg_gat_fwd_gn(c.graph, L.h, L.s, L.d, c.P[L.p_bias], L.o, L.m, L.z, L.cout, 0.2f, c.P[L.p_ga], 1e-5f, L.stats, c.red,
                         c.red_bytes < kWgradOffsetBytes ? c.red_bytes : kWgradOffsetBytes, c.st));
    return bg_graphnorm_apply(L.o, c.P[L.p_gw], c.P[L.p_gb], c.P[L.p_ga], L.stats, keep, keep_prob, seed, offset, c.N, L.cout,
                              L.x1, c.st);
}

// First-order backward of one block.  gx1 may be null (then `inj_o` IS the gradient at o).  Temporaries come
// from T; when `keep` the intermediates are written to the block's b_* buffers instead (second-order sweep).
// Graphs up to this many rows fuse the GraphNorm backward moments of the block below into the backward-input product
// (latency-bound regime: one launch less per block).  Larger graphs keep the separate, bandwidth-tuned reduction.
constexpr int64_t kFuseMomentsMaxN = 65536;

static int conv_backward(const Ctx& c, ConvL& L, const float* x_in, const float* gx1, float keep_scale, const float* inj_o,
                         const float* inj_h, bool keep, Arena& T, float* gx_out, const float* gate_x = nullptr,
                         ConvL* below = nullptr) {
    const int C = L.cout;
    float* go = keep ? L.b_go : T.f((size_t)c.N * C);
    float* gh = keep ? L.b_gh : T.f((size_t)c.N * C);
    float* gsd = keep ? L.b_gsd : T.f((size_t)c.N * 2);
    const bool moments_done = gx1 && L.pre_bst != nullptr;
    float* bst = moments_done ? L.pre_bst : (keep ? L.b_bstats : T.f((size_t)2 * C));
    c.X->off = 0;
    float* Pe = c.X->f((size_t)c.graph->E);
    float* DU = c.X->f((size_t)c.graph->E);
    float* gnpar = moments_done ? L.pre_gnpar : (c.G ? c.g(L.p_gw) : T.f((size_t)3 * C));
    L.pre_bst = L.pre_gnpar = nullptr;
    if (gx1) {
        // GraphNorm backward: column moments (+ parameter gradients) as one launch - unless the product that made gx1 already
        // delivered them; the elementwise half (and the injected cotangent at o) rides in the prologue of the aggregation
        // backward's destination pass
        if (!moments_done)
            BG_TRY(bg_graphnorm_bwd_moments(gx1, L.o, L.x1, c.P[L.p_gw], c.P[L.p_ga], L.stats, keep_scale, c.N, C, gnpar,
                                            c.G ? c.accumulate : 0, bst, c.red, c.red_bytes, c.st));
        BG_TRY(gat_bwd_gn_inj(c.graph, gx1, L.o, L.x1, c.P[L.p_gw], c.P[L.p_ga], L.stats, bst, keep_scale, inj_o, L.h, L.s, L.d, L.m,
                              L.z, c.P[L.p_as], c.P[L.p_ad], Pe, DU, go, gh, gsd, C, 0.2f, inj_h, c.st));
    } else {
        go = const_cast<float*>(inj_o);
        BG_TRY(gat_bwd_inj(c.graph, go, L.h, L.s, L.d, L.m, L.z, c.P[L.p_as], c.P[L.p_ad], Pe, DU, gh, gsd, C, 0.2f, inj_h, c.st));
    }
    BgSeg ones = seg(nullptr, 1, 0), hseg = seg(L.h, C, C), xseg = seg(x_in, L.cin, L.cin);
    if (c.G && inj_h) {  // bias, [att_src; att_dst] see gh BEFORE the injection at h (those cotangents are h-path only)
        BgWgrad pr[2] = {wg(c.N, go, C, C, &ones, 1, c.g(L.p_bias), 1, nullptr, c.accumulate),
                         wg(c.N, gsd, 2, 2, &hseg, 1, c.g(L.p_as), C, nullptr, c.accumulate)};
        BG_TRY(wgrad_launch(pr, 2, *c.q, as_stream(c.st)));
    }
    // the injection at h (gh += inj_h) rides in the source pass's epilogue (gat_bwd_*_inj above)
    if (c.G) {
        BgWgrad pr[3] = {wg(c.N, gh, C, C, &xseg, 1, c.g(L.p_W), L.cin, nullptr, c.accumulate),
                         wg(c.N, go, C, C, &ones, 1, c.g(L.p_bias), 1, nullptr, c.accumulate),
                         wg(c.N, gsd, 2, 2, &hseg, 1, c.g(L.p_as), C, nullptr, c.accumulate)};
        BG_TRY(wgrad_launch(pr, inj_h ? 1 : 3, *c.q, as_stream(c.st)));
    }
    if (gx_out && below && !gate_x && c.N <= kFuseMomentsMaxN && below->cout == L.cin) {
        // gx_out IS the gradient at the output of the block below: its GraphNorm backward moments ride in this product's epilogue
        const int Cb = below->cout;
        below->pre_bst = keep ? below->b_bstats : T.f((size_t)2 * Cb);
        below->pre_gnpar = c.G ? c.g(below->p_gw) : T.f((size_t)3 * Cb);
        GnMomFuse f{};
        f.o = below->o; f.x1 = below->x1; f.alpha = c.P[below->p_ga]; f.stats = below->stats; f.w = c.P[below->p_gw];
        f.keep_scale = keep_scale;
        f.dparams = below->pre_gnpar; f.accumulate = c.G ? c.accumulate : 0; f.bstats = below->pre_bst;
        f.counters = reinterpret_cast<unsigned int*>(c.red);
        f.partials = c.red + kCounterBytes / sizeof(float);
        BgDense a{};
        a.N = c.N; a.nseg = 1; a.seg[0] = seg(gh, C, C);
        a.W = c.P[L.p_W]; a.w_so = 1; a.w_sk = L.cin; a.Cout = L.cin;
        a.act = BG_ACT_NONE; a.out = gx_out; a.ld_out = L.cin;
        BG_TRY(dense_fwd_moments(&a, &f, as_stream(c.st)));
    } else if (gx_out) {
        BG_TRY(matmul_nn(c, gh, c.N, C, c.P[L.p_W], L.cin, 0, L.cin, gx_out, gate_x));
    }
    // second-order sweep needs gx1: the callers let the producer write it straight into the saved slot (no copy on the chain)
    if (keep && gx1 && gx1 != L.b_gx1)
        BG_TRY(cudaMemcpyAsync(L.b_gx1, gx1, (size_t)c.N * C * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(c.st)) == cudaSuccess
                   ? BG_OK
                   : BG_ECUDA);
    return BG_OK;
}

// Second-order step through (lin-bwd, gat-bwd, gn-bwd).  Xt = cotangent on gx; writes the cotangent on gx1 to
// `gx1t`, the injections to (ot, ht); direct parameter cotangents are accumulated into the grad bucket.
static int conv_backward2(const Ctx& c, const ConvL& L, const float* Xt, float keep_scale, Arena& T, float* gx1t, float* ot,
                          float* ht) {
    const int C = L.cout;
    float* Ht = T.f((size_t)c.N * C);
    float* St = T.f((size_t)c.N);
    float* Dt = T.f((size_t)c.N);
    float* gt = T.f((size_t)c.N * C);
    float* sdt = T.f((size_t)c.N * 2);
    c.X->off = 0;
    float* scratch = c.X->f((size_t)4 * c.graph->E);
    BG_TRY(matmul_nt(c, Xt, c.N, c.P[L.p_W], C, L.cin, 0, L.cin, Ht, c.P[L.p_as], c.P[L.p_ad], St, Dt));
    BG_TRY(bg_gat_bwd2(c.graph, Ht, St, Dt, L.b_go, L.h, L.s, L.d, L.m, L.z, c.P[L.p_as], c.P[L.p_ad], scratch, gt, ht, sdt, C,
                       0.2f, c.st));
    {
        BgSeg xt = seg(Xt, L.cin, L.cin), hts = seg(Ht, C, C), hseg = seg(L.h, C, C);
        BgWgrad pr[3] = {wg(c.N, L.b_gh, C, C, &xt, 1, c.g(L.p_W), L.cin, nullptr, 1),
                         wg(c.N, L.b_gsd, 2, 2, &hts, 1, c.g(L.p_as), C, nullptr, 1),
                         wg(c.N, sdt, 2, 2, &hseg, 1, c.g(L.p_as), C, nullptr, 1)};
        BG_TRY(wgrad_launch(pr, 3, *c.q, as_stream(c.st)));  // the two [att_src; att_dst] contributions fold in separate phases
    }
    return bg_graphnorm_bwd2(gt, L.b_gx1, L.o, L.x1, c.P[L.p_gw], c.P[L.p_ga], L.stats, L.b_bstats, keep_scale, c.N, C, gx1t, ot,
                             c.g(L.p_gw), 1, c.red, c.red_bytes, c.st);
}

// Backward of one dense layer: gz (pre-activation gradient), parameter gradients, and the input gradient for
// the requested column windows of the layer input.
// `gout_is_gz`: the incoming gradient already went through this layer's activation backward (the producer's epilogue
// was gated with this layer's output).  `gate_below`: this layer's input is the ReLU output of a plain layer - gate the
// first backward-input product with it, so that layer needs no activation-backward launch of its own.
static int dense_backward(const Ctx& c, const DenseL& l, int64_t rows, const BgSeg* segs, int nseg, const float* gout, float* gz_buf,
                          const int (*win)[2], float* const* gin, int nwin, const float** gz_out, bool gout_is_gz = false,
                          const float* gate_below = nullptr) {
    const float* gz = gout;
    if (gout_is_gz) {
        BG_REQUIRE(l.pg < 0, BG_EINVAL, "dense_backward: a fused activation backward needs a layer without LayerNorm");
    } else if (l.pg >= 0) {
        float* dgam = c.G ? c.g(l.pg) : nullptr;
        float* dbet = c.G ? c.g(l.pbeta) : nullptr;
        BG_REQUIRE(c.G, BG_EINVAL, "LayerNorm backward needs a grad bucket");
        BG_TRY(bg_ln_act_bwd(gout, l.out, l.xhat, l.rstd, c.P[l.pg], rows, l.cout, l.act, gz_buf, dgam, dbet, c.accumulate, c.red,
                             c.red_bytes, c.st));
        gz = gz_buf;
    } else if (l.act != BG_ACT_NONE) {
        BG_TRY(bg_ln_act_bwd(gout, l.out, nullptr, nullptr, nullptr, rows, l.cout, l.act, gz_buf, nullptr, nullptr, 0, nullptr, 0, c.st));
        gz = gz_buf;
    }
    if (c.G) {
        BgSeg all[BG_MAX_SEG];
        for (int i = 0; i < nseg; ++i) all[i] = segs[i];
        all[nseg] = seg(nullptr, 1, 0);
        BgWgrad pr = wg(rows, gz, l.cout, l.cout, all, nseg + 1, c.g(l.pW), l.cin, c.g(l.pb), c.accumulate);
        BG_TRY(wgrad_launch(&pr, 1, *c.q, as_stream(c.st)));
    }
    for (int i = 0; i < nwin; ++i)
        BG_TRY(matmul_nn(c, gz, rows, l.cout, c.P[l.pW], l.cin, win[i][0], win[i][1], gin[i], i == 0 ? gate_below : nullptr));
    if (gz_out) *gz_out = gz;
    return BG_OK;
}

}  // namespace bg

using namespace bg;

// =====================================================================================================
// generator
// =====================================================================================================
extern "C" int32_t bg_gen_num_params(const BgModelDesc* md) {
    GenNet g;
    if (!md || build_gen(*md, g)) return -1;
    return g.nparams;
}
extern "C" int32_t bg_disc_num_params(const BgModelDesc* md) {
    DiscNet d;
    if (!md || build_disc(*md, d)) return -1;
    return d.nparams;
}

extern "C" size_t bg_gen_fwd_ws(const BgModelDesc* md, int64_t N, int64_t E) {
    GenNet g;
    if (!md || build_gen(*md, g)) return 0;
    Arena A{nullptr, 0};
    layout_gen(*md, g, N, A);
    (void)E;
    return align_up(A.off, 256);
}

extern "C" int bg_gen_forward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                              const float* z, const float* noise, const uint8_t* const* keeps, int32_t training, uint64_t seed,
                              uint64_t offset, void* ws, size_t ws_bytes, float* red, size_t red_bytes, float* logits, float* hard,
                              float* soft, void* stream) {
    BG_REQUIRE(md && params && graph && in && z && ws && red && logits && hard && soft, BG_EINVAL, "bg_gen_forward: null pointer");
    GenNet g;
    BG_TRY(build_gen(*md, g));
    const int64_t N = graph->N;
    BG_REQUIRE(ws_bytes >= bg_gen_fwd_ws(md, N, graph->E), BG_EINVAL, "bg_gen_forward: workspace too small");
    Arena A{static_cast<char*>(ws), 0};
    layout_gen(*md, g, N, A);
    Ctx c{params, nullptr, nullptr, graph, N, red, red_bytes, stream, 0, nullptr};
    const int K = md->num_classes;
    // type-matched encoder on the K table rows (row-wise ops commute with the per-voxel gather)
    const float* e = in->table;
    int ew = md->local_dim;
    if (small_chain_ok(g.menc, g.n_menc, K)) {
        BgSmallLayer sl[BG_SMALL_MAX_LAYERS];
        small_chain_fill(c, g.menc, g.n_menc, sl, false);
        BG_TRY(bg_small_mlp_fwd(sl, g.n_menc, e, K, stream));
        e = g.menc[g.n_menc - 1].out;
        ew = g.menc[g.n_menc - 1].cout;
    } else {
        for (int i = 0; i < g.n_menc; ++i) {
            BgSeg s = seg(e, ew, ew);
            BG_TRY(dense_fwd_call(c, g.menc[i], K, &s, 1, true));
            e = g.menc[i].out;
            ew = g.menc[i].cout;
        }
    }
    const BgSeg enc = seg(e, md->le_dim, md->le_dim, in->type32);
    const BgSeg vx = seg(in->vx, md->voxel_dim, md->voxel_dim), zz = seg(z, md->z_dim, md->z_dim);
    {
        BgSeg s3[3] = {enc, vx, zz};
        BG_TRY(dense_fwd_call(c, g.mlp[0], N, s3, 3, true));
        for (int i = 1; i < g.n_mlp; ++i) {
            BgSeg s = seg(g.mlp[i - 1].out, md->g_hidden, md->g_hidden);
            BG_TRY(dense_fwd_call(c, g.mlp[i], N, &s, 1, true));
        }
    }
    const float* x = g.mlp[g.n_mlp - 1].out;
    const float* h = x;
    for (int k = 0; k < g.n_conv; ++k) {
        const uint8_t* keep = (training && keeps) ? keeps[k] : nullptr;
        BG_TRY(conv_forward(c, g.conv[k], h, keep, training ? kKeepProb : 1.f, seed, offset + k));
        h = g.conv[k].x1;
    }
    {
        const int eo = g.conv[g.n_conv - 1].cout;
        BgSeg s5[5] = {seg(h, eo, eo), seg(x, md->g_hidden, md->g_hidden), enc, vx, zz};
        BG_TRY(dense_fwd_call(c, g.dec[0], N, s5, 5, true));
        for (int i = 1; i < g.n_dec; ++i) {
            BgSeg s = seg(g.dec[i - 1].out, g.dec[i].cin, g.dec[i].cin);
            DenseL l = g.dec[i];
            if (i == g.n_dec - 1) l.out = logits;
            BG_TRY(dense_fwd_call(c, l, N, &s, 1, true));
        }
    }
    return bg_gumbel_st_fwd(logits, noise, seed, offset + 1000, N, K, soft, hard, nullptr, stream);
}

// Scratch layout of the backward passes: [0, scratch_bytes) = per-layer scratch (per-edge arrays, rewound by every
// block), rest = pass arena.  Nothing in the pass arena is reused before the end of the pass: every gradient a weight
// gradient reads stays alive until the single deferred wgrad launch (wgrad_flush).
static size_t scratch_bytes(int64_t E) { return align_up((size_t)4 * (size_t)E * sizeof(float), 256) + 4096; }
struct Tally {  // mirrors Arena::f for size queries
    size_t off = 0;
    void f(size_t n) { off = align_up(off, 256) + n * sizeof(float); }
};
static void tally_conv_bwd(Tally& t, const ConvL& L, int64_t N, bool keep) {
    if (!keep) { t.f((size_t)N * L.cout); t.f((size_t)N * L.cout); t.f((size_t)N * 2); t.f((size_t)2 * L.cout); }
    t.f((size_t)3 * L.cout);
    t.f((size_t)N * L.cin);
}

extern "C" size_t bg_gen_bwd_ws(const BgModelDesc* md, int64_t N, int64_t E) {
    GenNet g;
    if (!md || build_gen(*md, g)) return 0;
    const size_t K = md->num_classes, le = md->le_dim, gh = md->g_hidden;
    Tally t;
    t.f(N * gh); t.f(N * le); t.f(N * le); t.f(K * le); t.f(K * le); t.f(N * K);
    for (int i = g.n_dec - 1; i >= 0; --i) { t.f((size_t)N * g.dec[i].cout); t.f((size_t)N * (i ? g.dec[i].cin : g.conv[g.n_conv - 1].cout)); }
    for (int k = 0; k < g.n_conv; ++k) tally_conv_bwd(t, g.conv[k], N, false);
    for (int i = 0; i < g.n_mlp; ++i) { t.f(N * gh); t.f(N * gh); }
    for (int i = 0; i < g.n_menc + 2; ++i) { t.f(K * le); t.f(K * le); }
    return scratch_bytes(E) + align_up(t.off, 256) + 65536;
}

extern "C" int bg_gen_backward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                               const float* z, const void* ws_fwd, const float* logits, const float* soft, const float* g_logits,
                               const float* g_hard, const float* g_soft, int32_t training, float* grad_flat, const int64_t* grad_off,
                               int32_t accumulate, void* tmp, size_t tmp_bytes, float* red, size_t red_bytes, void* stream) {
    BG_REQUIRE(md && params && graph && in && z && ws_fwd && soft && grad_flat && grad_off && tmp && red, BG_EINVAL,
               "bg_gen_backward: null pointer");
    BG_REQUIRE(g_logits || g_hard || g_soft, BG_EINVAL, "bg_gen_backward: no incoming gradient");
    GenNet net;
    BG_TRY(build_gen(*md, net));
    const int64_t N = graph->N;
    Arena A{static_cast<char*>(const_cast<void*>(ws_fwd)), 0};
    layout_gen(*md, net, N, A);
    net.dec[net.n_dec - 1].out = const_cast<float*>(logits);
    BG_REQUIRE(tmp_bytes >= bg_gen_bwd_ws(md, N, graph->E), BG_EINVAL, "bg_gen_backward: scratch too small");
    WgradQueue queue;
    BG_TRY(init_queue(queue, red, red_bytes));
    Ctx c{params, grad_flat, grad_off, graph, N, red, red_bytes, stream, accumulate ? 1 : 0, &queue};
    const int K = md->num_classes, le = md->le_dim, gh = md->g_hidden;
    Arena X{static_cast<char*>(tmp), 0, scratch_bytes(graph->E)};
    Arena T{static_cast<char*>(tmp) + scratch_bytes(graph->E), 0, tmp_bytes - scratch_bytes(graph->E)};
    c.X = &X;
    float* g_skip = T.f((size_t)N * gh);
    float* g_e1 = T.f((size_t)N * le);
    float* g_e2 = T.f((size_t)N * le);
    float* ge = T.f((size_t)K * le);
    float* ge2 = T.f((size_t)K * le);
    float* gl = T.f((size_t)N * K);
    // gumbel straight-through
    const float* g = g_logits;
    if (g_hard || g_soft) {
        BG_TRY(bg_gumbel_st_bwd(g_hard, g_soft, soft, N, K, gl, stream));
        if (g_logits) BG_TRY(bg_axpy(gl, g_logits, 1.f, N * K, stream));
        g = gl;
    }
    const float* e_last = net.n_menc ? net.menc[net.n_menc - 1].out : in->table;
    const BgSeg enc = seg(e_last, le, le, in->type32);
    const BgSeg vx = seg(in->vx, md->voxel_dim, md->voxel_dim), zz = seg(z, md->z_dim, md->z_dim);
    const float* x = net.mlp[net.n_mlp - 1].out;
    // decoder, last to second layer (every layer gets fresh buffers: see scratch_bytes())
    for (int i = net.n_dec - 1; i >= 1; --i) {
        BgSeg s = seg(net.dec[i - 1].out, net.dec[i].cin, net.dec[i].cin);
        int win[1][2] = {{0, net.dec[i].cin}};
        float* gzb = T.f((size_t)N * net.dec[i].cout);
        float* gin[1] = {T.f((size_t)N * net.dec[i].cin)};
        BG_TRY(dense_backward(c, net.dec[i], N, &s, 1, g, gzb, win, gin, 1, nullptr));
        g = gin[0];
    }
    {
        const int eo = net.conv[net.n_conv - 1].cout;
        BgSeg s5[5] = {seg(net.conv[net.n_conv - 1].x1, eo, eo), seg(x, gh, gh), enc, vx, zz};
        int win[3][2] = {{0, eo}, {eo, eo + gh}, {eo + gh, eo + gh + le}};
        float* gzb = T.f((size_t)N * net.dec[0].cout);
        float* gin[3] = {T.f((size_t)N * eo), g_skip, g_e1};
        BG_TRY(dense_backward(c, net.dec[0], N, s5, 5, g, gzb, win, gin, 3, nullptr));
        g = gin[0];
    }
    for (int k = net.n_conv - 1; k >= 0; --k) {
        const float* x_in = k == 0 ? x : net.conv[k - 1].x1;
        float* gx = T.f((size_t)N * net.conv[k].cin);
        BG_TRY(conv_backward(c, net.conv[k], x_in, g, training ? 1.f / kKeepProb : 1.f, nullptr, nullptr, false, T, gx, nullptr,
                             k > 0 ? &net.conv[k - 1] : nullptr));
        g = gx;
    }
    BG_TRY(bg_axpy(const_cast<float*>(g), g_skip, 1.f, N * gh, stream));
    for (int i = net.n_mlp - 1; i >= 1; --i) {
        BgSeg s = seg(net.mlp[i - 1].out, gh, gh);
        int win[1][2] = {{0, gh}};
        float* gzb = T.f((size_t)N * gh);
        float* gin[1] = {T.f((size_t)N * gh)};
        BG_TRY(dense_backward(c, net.mlp[i], N, &s, 1, g, gzb, win, gin, 1, nullptr));
        g = gin[0];
    }
    {
        BgSeg s3[3] = {enc, vx, zz};
        int win[1][2] = {{0, le}};
        float* gin[1] = {g_e2};
        BG_TRY(dense_backward(c, net.mlp[0], N, s3, 3, g, T.f((size_t)N * gh), win, gin, 1, nullptr));
    }
    BG_TRY(bg_type_scatter_sum(g_e1, le, in->type32, N, le, K, ge, red, red_bytes, stream));
    BG_TRY(bg_type_scatter_sum(g_e2, le, in->type32, N, le, K, ge2, red, red_bytes, stream));
    BG_TRY(bg_axpy(ge, ge2, 1.f, (int64_t)K * le, stream));
    const float* gcur = ge;
    if (small_chain_ok(net.menc, net.n_menc, K)) {
        BgSmallLayer sl[BG_SMALL_MAX_LAYERS];
        small_chain_fill(c, net.menc, net.n_menc, sl, true);
        BG_TRY(bg_small_mlp_bwd(sl, net.n_menc, in->table, K, ge, nullptr, c.accumulate, stream));
    } else
    for (int i = net.n_menc - 1; i >= 0; --i) {
        const float* xin = i == 0 ? in->table : net.menc[i - 1].out;
        BgSeg s = seg(xin, net.menc[i].cin, net.menc[i].cin);
        int win[1][2] = {{0, net.menc[i].cin}};
        float* small_gz = T.f((size_t)K * le);
        float* gin[1] = {T.f((size_t)K * le)};
        BG_TRY(dense_backward(c, net.menc[i], K, &s, 1, gcur, small_gz, win, gin, i > 0 ? 1 : 0, nullptr));
        gcur = gin[0];
    }
    BG_REQUIRE(!T.overflow && !X.overflow, BG_EINVAL, "bg_gen_backward: scratch too small");
    return wgrad_flush(queue, as_stream(stream));
}

// =====================================================================================================
// discriminator
// =====================================================================================================
extern "C" size_t bg_disc_fwd_ws(const BgModelDesc* md, int64_t N, int64_t E) {
    DiscNet d;
    if (!md || build_disc(*md, d)) return 0;
    Arena A{nullptr, 0};
    layout_disc(d, N, A);
    (void)E;
    return align_up(A.off, 256);
}

extern "C" int bg_disc_forward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                               const float* label, const uint8_t* const* keeps, int32_t training, uint64_t seed, uint64_t offset,
                               void* ws, size_t ws_bytes, float* red, size_t red_bytes, float* score, void* stream) {
    BG_REQUIRE(md && params && graph && in && label && ws && red && score, BG_EINVAL, "bg_disc_forward: null pointer");
    DiscNet d;
    BG_TRY(build_disc(*md, d));
    const int64_t N = graph->N;
    BG_REQUIRE(ws_bytes >= bg_disc_fwd_ws(md, N, graph->E), BG_EINVAL, "bg_disc_forward: workspace too small");
    Arena A{static_cast<char*>(ws), 0};
    layout_disc(d, N, A);
    Ctx c{params, nullptr, nullptr, graph, N, red, red_bytes, stream, 0, nullptr};
    BgSeg s3[3] = {seg(in->table, md->local_dim, md->local_dim, in->type32), seg(in->vx, md->voxel_dim, md->voxel_dim),
                   seg(label, md->num_classes, md->num_classes)};
    BG_TRY(dense_fwd_call(c, d.pre[0], N, s3, 3, false));
    BgSeg s1 = seg(d.pre[0].out, md->d_hidden, md->d_hidden);
    BG_TRY(dense_fwd_call(c, d.pre[1], N, &s1, 1, false));
    const float* h = d.pre[1].out;
    for (int k = 0; k < d.n_conv; ++k) {
        const uint8_t* keep = (training && keeps) ? keeps[k] : nullptr;
        BG_TRY(conv_forward(c, d.conv[k], h, keep, training ? kKeepProb : 1.f, seed, offset + k));
        h = d.conv[k].x1;
    }
    for (int i = 0; i < 4; ++i) {
        BgSeg s = seg(h, d.dec[i].cin, d.dec[i].cin);
        DenseL l = d.dec[i];
        if (i == 3) l.out = score;
        BG_TRY(dense_fwd_call(c, l, N, &s, 1, false));
        h = l.out;
    }
    return BG_OK;
}

extern "C" size_t bg_disc_bwd_saved_ws(const BgModelDesc* md, int64_t N, int64_t E) {
    DiscNet d;
    if (!md || build_disc(*md, d)) return 0;
    Arena A{nullptr, 0};
    layout_disc_bwd_saved(d, N, A);
    (void)E;
    return align_up(A.off, 256);
}
static size_t disc_bwd_floats_bytes(const BgModelDesc* md, const DiscNet& d, int64_t N, bool keep) {
    const size_t dh = md->d_hidden;
    Tally t;
    for (int i = 3; i >= 0; --i) { t.f((size_t)N * d.dec[i].cout); t.f((size_t)N * d.dec[i].cin); }
    for (int k = 0; k < d.n_conv; ++k) tally_conv_bwd(t, d.conv[k], N, keep);
    t.f(N * dh); t.f(N * dh); t.f(N * dh);
    return align_up(t.off, 256);
}
// scratch of bg_disc_backward; bg_disc_backward2 takes twice this
extern "C" size_t bg_disc_tmp_ws(const BgModelDesc* md, int64_t N, int64_t E) {
    DiscNet d;
    if (!md || build_disc(*md, d)) return 0;
    const size_t dh = md->d_hidden;
    const size_t bwd = disc_bwd_floats_bytes(md, d, N, false);
    Tally t;  // second-order sweep, then the injected first-order sweep
    t.f(N * dh); t.f(N * dh); t.f(N * dh); t.f(N * dh);
    for (int k = 0; k < d.n_conv; ++k) {
        const size_t C = d.conv[k].cout;
        t.f(N * C); t.f(N * C);                                              // injections
        t.f(N * C); t.f(N); t.f(N); t.f(N * C); t.f(N * 2); t.f(N * C);   // Ht, St, Dt, gt, sdt, cotangent out
    }
    for (int i = 0; i < 4; ++i) { t.f((size_t)N * d.dec[i].cout); t.f((size_t)N * d.dec[i].cout); }
    const size_t bwd2 = align_up(t.off, 256) + bwd;
    const size_t need = bwd > (bwd2 + 1) / 2 ? bwd : (bwd2 + 1) / 2;
    return scratch_bytes(E) + need + 65536;
}

// First-order backward.  g_score may be null together with inject != null (second-order sweep's forward-graph
// part).  grad_flat may be null (no parameter gradients).  saved != null keeps the intermediates for bwd2.
static int disc_backward_impl(const BgModelDesc* md, DiscNet& d, Ctx& c, const BgBatchIn* in, const float* label, const float* score,
                              const float* g_score, float training_scale, const float* const* inj_o, const float* const* inj_h,
                              bool keep, Arena& T, float* g_label) {
    const int64_t N = c.N;
    const int dh = md->d_hidden;
    const float* g = g_score;
    bool g_is_gz = false;  // g already carries the activation backward of the layer it enters (fused gate)
    auto plain_relu = [](const DenseL& l) { return l.pg < 0 && l.act == BG_ACT_RELU; };
    if (g) {
        for (int i = 3; i >= 0; --i) {
            const float* xin = i == 0 ? d.conv[d.n_conv - 1].x1 : d.dec[i - 1].out;
            BgSeg s = seg(xin, d.dec[i].cin, d.dec[i].cin);
            DenseL l = d.dec[i];
            if (i == 3) l.out = const_cast<float*>(score);
            int win[1][2] = {{0, l.cin}};
            const bool gate = i > 0 && plain_relu(d.dec[i - 1]);  // dec[i-1]'s ReLU backward rides on this dgrad
            // i == 0: the product's output is gx1 of the top block - written into its saved slot when the sweep is kept
            float* gin[1] = {gate && keep ? d.dec[i - 1].b_gz : (i == 0 && keep ? d.conv[d.n_conv - 1].b_gx1 : T.f((size_t)N * l.cin))};
            float* gzbuf = g_is_gz ? nullptr : (keep ? l.b_gz : T.f((size_t)N * l.cout));
            const float* gz_used = nullptr;
            BG_TRY(dense_backward(c, l, N, &s, 1, g, gzbuf, win, gin, 1, &gz_used, g_is_gz, gate ? d.dec[i - 1].out : nullptr));
            if (keep && !g_is_gz && gz_used != gzbuf)  // bare Linear: gz == gout, keep a copy for the second-order sweep
                BG_TRY(cudaMemcpyAsync(gzbuf, gz_used, (size_t)N * l.cout * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(c.st)) ==
                               cudaSuccess
                           ? BG_OK
                           : BG_ECUDA);
            g = gin[0];
            g_is_gz = gate;
        }
    }
    for (int k = d.n_conv - 1; k >= 0; --k) {
        const float* x_in = k == 0 ? d.pre[1].out : d.conv[k - 1].x1;
        const bool gate = k == 0 && plain_relu(d.pre[1]);
        float* gx = gate && keep ? d.pre[1].b_gz : (k > 0 && keep ? d.conv[k - 1].b_gx1 : T.f((size_t)N * d.conv[k].cin));
        BG_TRY(conv_backward(c, d.conv[k], x_in, g, training_scale, inj_o ? inj_o[k] : nullptr, inj_h ? inj_h[k] : nullptr, keep, T,
                             gx, gate ? d.pre[1].out : nullptr, k > 0 ? &d.conv[k - 1] : nullptr));
        g = gx;
        g_is_gz = gate;
    }
    {
        BgSeg s = seg(d.pre[0].out, dh, dh);
        int win[1][2] = {{0, dh}};
        const bool gate = plain_relu(d.pre[0]);
        float* gin[1] = {gate && keep ? d.pre[0].b_gz : T.f((size_t)N * dh)};
        float* gzbuf = g_is_gz ? nullptr : (keep ? d.pre[1].b_gz : T.f((size_t)N * dh));
        BG_TRY(dense_backward(c, d.pre[1], N, &s, 1, g, gzbuf, win, gin, 1, nullptr, g_is_gz, gate ? d.pre[0].out : nullptr));
        g = gin[0];
        g_is_gz = gate;
    }
    {
        const int lo = md->local_dim + md->voxel_dim;
        BgSeg s3[3] = {seg(in->table, md->local_dim, md->local_dim, in->type32), seg(in->vx, md->voxel_dim, md->voxel_dim),
                       seg(label, md->num_classes, md->num_classes)};
        int win[1][2] = {{lo, lo + md->num_classes}};
        float* gin[1] = {g_label};
        float* gzbuf = g_is_gz ? nullptr : (keep ? d.pre[0].b_gz : T.f((size_t)N * dh));
        BG_TRY(dense_backward(c, d.pre[0], N, s3, 3, g, gzbuf, win, gin, g_label ? 1 : 0, nullptr, g_is_gz));
    }
    return BG_OK;
}

extern "C" int bg_disc_backward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                                const float* label, const void* ws_fwd, const float* score, const float* g_score, int32_t training,
                                float* grad_flat, const int64_t* grad_off, int32_t accumulate, void* saved, size_t saved_bytes, void* tmp,
                                size_t tmp_bytes, float* red, size_t red_bytes, float* g_label, void* stream) {
    BG_REQUIRE(md && params && graph && in && label && ws_fwd && score && g_score && tmp && red, BG_EINVAL,
               "bg_disc_backward: null pointer");
    DiscNet d;
    BG_TRY(build_disc(*md, d));
    const int64_t N = graph->N;
    Arena A{static_cast<char*>(const_cast<void*>(ws_fwd)), 0};
    layout_disc(d, N, A);
    if (saved) {
        BG_REQUIRE(saved_bytes >= bg_disc_bwd_saved_ws(md, N, graph->E), BG_EINVAL, "bg_disc_backward: saved buffer too small");
        Arena S{static_cast<char*>(saved), 0};
        layout_disc_bwd_saved(d, N, S);
    }
    BG_REQUIRE(tmp_bytes >= bg_disc_tmp_ws(md, N, graph->E), BG_EINVAL, "bg_disc_backward: scratch too small");
    WgradQueue queue;
    BG_TRY(init_queue(queue, red, red_bytes));
    Ctx c{params, grad_flat, grad_off, graph, N, red, red_bytes, stream, accumulate ? 1 : 0, &queue, nullptr};
    Arena X{static_cast<char*>(tmp), 0, scratch_bytes(graph->E)};
    Arena T{static_cast<char*>(tmp) + scratch_bytes(graph->E), 0, tmp_bytes - scratch_bytes(graph->E)};
    c.X = &X;
    BG_TRY(disc_backward_impl(md, d, c, in, label, score, g_score, training ? 1.f / kKeepProb : 1.f, nullptr, nullptr,
                              saved != nullptr, T, g_label));
    BG_REQUIRE(!T.overflow && !X.overflow, BG_EINVAL, "bg_disc_backward: scratch too small");
    return wgrad_flush(queue, as_stream(stream));
}

// Second-order sweep (WGAN-GP).  Lt = cotangent on g_label.  grad_flat (zero-initialised by the caller) receives the
// parameter cotangents; gt_score (optional) the cotangent on g_score.
extern "C" int bg_disc_backward2(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                                 const float* label, const void* ws_fwd, const float* score, const void* saved, const float* Lt,
                                 int32_t training, float* grad_flat, const int64_t* grad_off, void* tmp, size_t tmp_bytes, float* red,
                                 size_t red_bytes, float* gt_score, void* stream) {
    BG_REQUIRE(md && params && graph && in && label && ws_fwd && score && saved && Lt && grad_flat && grad_off && tmp && red, BG_EINVAL,
               "bg_disc_backward2: null pointer");
    DiscNet d;
    BG_TRY(build_disc(*md, d));
    const int64_t N = graph->N;
    Arena A{static_cast<char*>(const_cast<void*>(ws_fwd)), 0};
    layout_disc(d, N, A);
    Arena S{static_cast<char*>(const_cast<void*>(saved)), 0};
    layout_disc_bwd_saved(d, N, S);
    BG_REQUIRE(tmp_bytes >= 2 * bg_disc_tmp_ws(md, N, graph->E), BG_EINVAL, "bg_disc_backward2: scratch too small");
    WgradQueue queue;
    BG_TRY(init_queue(queue, red, red_bytes));
    Ctx c{params, grad_flat, grad_off, graph, N, red, red_bytes, stream, 1, &queue, nullptr};
    const float keep_scale = training ? 1.f / kKeepProb : 1.f;
    const int dh = md->d_hidden, K = md->num_classes, lo = md->local_dim + md->voxel_dim;
    Arena X{static_cast<char*>(tmp), 0, scratch_bytes(graph->E)};
    Arena T{static_cast<char*>(tmp) + scratch_bytes(graph->E), 0, tmp_bytes - scratch_bytes(graph->E)};
    c.X = &X;
    float* ta = T.f((size_t)N * dh);
    float* tb = T.f((size_t)N * dh);
    float* inj_o[kMaxConvs];
    float* inj_h[kMaxConvs];
    for (int k = 0; k < d.n_conv; ++k) {
        inj_o[k] = T.f((size_t)N * d.conv[k].cout);
        inj_h[k] = T.f((size_t)N * d.conv[k].cout);
    }
    // pre[0] backward was: gz0 = g_a * [x_a > 0] ; g_label = gz0 @ W0[:, lo:lo+K]
    {
        BgSeg lt = seg(Lt, K, K);
        BgWgrad pr = wg(N, d.pre[0].b_gz, dh, dh, &lt, 1, c.g(d.pre[0].pW) + lo, d.pre[0].cin, nullptr, 1);
        BG_TRY(wgrad_launch(&pr, 1, *c.q, as_stream(stream)));
        // cot(g_a) = cot(gz0) * [x_a > 0], the mask fused into the product's epilogue
        BG_TRY(matmul_nt(c, Lt, N, params[d.pre[0].pW], dh, d.pre[0].cin, lo, lo + K, tb, nullptr, nullptr, nullptr, nullptr,
                         d.pre[0].out));
        (void)ta;
    }
    // pre[1] backward was: gz1 = g_b * [x_b > 0] ; g_a = gz1 @ W1          (tb = cot(g_a))
    {
        BgSeg ts = seg(tb, dh, dh);
        BgWgrad pr = wg(N, d.pre[1].b_gz, dh, dh, &ts, 1, c.g(d.pre[1].pW), dh, nullptr, 1);
        BG_TRY(wgrad_launch(&pr, 1, *c.q, as_stream(stream)));
        float* tc = T.f((size_t)N * dh);  // the previous tb is a queued weight-gradient operand: leave it alone
        BG_TRY(matmul_nt(c, tb, N, params[d.pre[1].pW], dh, dh, 0, dh, tc, nullptr, nullptr, nullptr, nullptr,
                         d.pre[1].out));                                                          // cot(g_b) = cot(gz1) * mask
        tb = tc;
    }
    float* t = tb;       // cotangent flowing up the backward chain (fresh buffer per layer)
    for (int k = 0; k < d.n_conv; ++k) {
        float* out = T.f((size_t)N * d.conv[k].cout);
        BG_TRY(conv_backward2(c, d.conv[k], t, keep_scale, T, out, inj_o[k], inj_h[k]));
        t = out;
    }
    for (int i = 0; i < 4; ++i) {
        // backward was: gz = g_y * act'(y) ; g_in = gz @ W        (t = cot(g_in))
        const DenseL& l = d.dec[i];
        BgSeg ts = seg(t, l.cin, l.cin);
        BgWgrad pr = wg(N, l.b_gz, l.cout, l.cout, &ts, 1, c.g(l.pW), l.cin, nullptr, 1);
        BG_TRY(wgrad_launch(&pr, 1, *c.q, as_stream(stream)));
        if (i == 3 && !gt_score) break;
        float* cz = T.f((size_t)N * l.cout);
        if (l.act == BG_ACT_RELU) {  // cot(g_y) = cot(gz) * [y > 0], fused
            BG_TRY(matmul_nt(c, t, N, params[l.pW], l.cout, l.cin, 0, l.cin, cz, nullptr, nullptr, nullptr, nullptr, l.out));
            t = cz;
        } else {
            BG_TRY(matmul_nt(c, t, N, params[l.pW], l.cout, l.cin, 0, l.cin, cz));                 // cot(gz)
            if (l.act != BG_ACT_NONE) {
                t = T.f((size_t)N * l.cout);
                BG_TRY(bg_ln_act_bwd(cz, l.out, nullptr, nullptr, nullptr, N, l.cout, l.act, t, nullptr, nullptr, 0, nullptr, 0, stream));
            } else {
                t = cz;
            }
        }
    }
    if (gt_score)
        BG_TRY(cudaMemcpyAsync(gt_score, t, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)) == cudaSuccess
                   ? BG_OK
                   : BG_ECUDA);
    // forward-graph sweep with the injected cotangents (nothing flows in from the top: the score itself is
    // not part of the second-order loss)
    BG_TRY(disc_backward_impl(md, d, c, in, label, score, nullptr, keep_scale, inj_o, inj_h, false, T, nullptr));
    BG_REQUIRE(!T.overflow && !X.overflow, BG_EINVAL, "bg_disc_backward2: scratch too small");
    return wgrad_flush(queue, as_stream(stream));
}

// ---- debug / test hooks: byte offsets of the saved post-activation tensors inside the forward workspace, in
// layer order (generator: menc outs, mlp outs, conv x1, dec outs; discriminator: pre outs, conv x1, dec outs).
extern "C" int32_t bg_gen_ws_offsets(const BgModelDesc* md, int64_t N, int64_t* out, int32_t cap) {
    GenNet g;
    if (!md || !out || build_gen(*md, g)) return -1;
    char* const base = reinterpret_cast<char*>(4096);
    Arena A{base, 0};
    layout_gen(*md, g, N, A);
    int n = 0;
    auto put = [&](const float* p) { if (n < cap) out[n] = reinterpret_cast<const char*>(p) - base; ++n; };
    for (int i = 0; i < g.n_menc; ++i) put(g.menc[i].out);
    for (int i = 0; i < g.n_mlp; ++i) put(g.mlp[i].out);
    for (int k = 0; k < g.n_conv; ++k) put(g.conv[k].x1);
    for (int i = 0; i < g.n_dec; ++i) put(g.dec[i].out);
    return n;
}
extern "C" int32_t bg_disc_ws_offsets(const BgModelDesc* md, int64_t N, int64_t* out, int32_t cap) {
    DiscNet d;
    if (!md || !out || build_disc(*md, d)) return -1;
    char* const base = reinterpret_cast<char*>(4096);
    Arena A{base, 0};
    layout_disc(d, N, A);
    int n = 0;
    auto put = [&](const float* p) { if (n < cap) out[n] = reinterpret_cast<const char*>(p) - base; ++n; };
    for (int i = 0; i < 2; ++i) put(d.pre[i].out);
    for (int k = 0; k < d.n_conv; ++k) put(d.conv[k].x1);
    for (int i = 0; i < 4; ++i) put(d.dec[i].out);
    return n;
}
