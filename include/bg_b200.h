/*
 * bg_b200.h - C ABI of libbgb200.so: the B200 (sm_100a) kernels behind the Building-GAN
 * voxel-graph message-passing hot path.
 *
 * Every entry point replaces work the reference does through torch / torch_geometric ops
 * (the reference has no native code of its own); the interface each one stands in for is
 * cited as reference file:line next to it.  Conventions:
 *
 *   - pure C, POD only; device pointers are raw `float*` / `int32_t*`, `stream` is a
 *     `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - the library never allocates or frees device memory and never synchronises: inputs, outputs, saved
 *     tensors and workspaces are caller-owned; every launch is ordered on the caller's stream, so a
 *     sequence of calls is CUDA-graph capturable.  The only library-owned CUDA objects are one
 *     non-blocking side stream + two events per device, created on first use by the whole-pass
 *     backward entry points: the batched weight-gradient launches of a pass run there, forked from
 *     and joined back into the caller's stream with events before the pass returns control of its
 *     buffers (BG_WGRAD_OVERLAP=0 keeps everything on the caller's stream).  Process-global state:
 *     that pool, the launch-geometry knobs of bg_tune() and the bg_set_dense_tc() switch;
 *   - return 0 on success, a negative BG_E* code otherwise; `bg_last_error()` returns a
 *     thread-local human-readable message.  There is NO CPU fallback.
 *   - all activations are row-major fp32 `[N, C]`; supported GNN channel widths are
 *     C in {1,2,4,8,16,32,64,128} (the reference's hourglass, models.py:68-88,187-208).
 */
#ifndef BG_B200_H
#define BG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_OK 0
#define BG_EINVAL (-1)   /* bad argument                                  */
#define BG_EUNSUPPORTED (-2) /* width / mode not compiled                 */
#define BG_ECUDA (-3)    /* launch failure reported by cudaGetLastError() */
#define BG_ERANGE (-4)   /* index out of range in host-side graph build   */

/* Borrowed-pointer view of one collated voxel batch (built once per batch, H1).
 * rowptr/col: in-edges of each destination node (source ids), input self loops removed and one
 * self loop appended LAST (GATConv add_self_loops, reference models.py:72 -> PyG GATConv);
 * per-row order = COO order = the reference's CPU scatter order.
 * cscptr/cscrow/perm: out-edges of each source node, and for each the index of the same edge in
 * the CSR arrays.  graph_ptr: node range of each building (Batch.ptr, data.py:160-161). */
typedef struct BgGraph {
    const int32_t* rowptr;    /* [N+1] */
    const int32_t* col;       /* [E]   */
    const int32_t* cscptr;    /* [N+1] */
    const int32_t* cscrow;    /* [E]   */
    const int32_t* perm;      /* [E]   */
    const int32_t* graph_ptr; /* [B+1] */
    int64_t N;
    int64_t E; /* edges incl. the N self loops */
    int32_t B;
    int32_t max_deg;
} BgGraph;

int bg_version(void);
const char* bg_last_error(void);
/* Launch-geometry knobs of the aggregation kernels on graphs larger than L2 (no reference counterpart; used by the
 * kernel sweep in scratch/gat_sweep.py and by the tests).  key: 0 = threads per CTA (default 1024), 1 = CTAs per SM (1),
 * 2 = distance in rows of the sequential L2 prefetch stream (128), 3 = use the software-pipelined kernels (1),
 * 4 = use the large-graph geometry for every graph (test hook, 0), 5 = KiB of feature rows per sweep chunk (64),
 * 6 = grid size cap (test hook, 0 = SMs x key 1). */
int bg_tune(int32_t key, int32_t value);

/* ---- H1: collation (reference data.py:156-163 Batch.from_data_list; PyG GATConv's
 * remove_self_loops + add_self_loops that the reference re-runs inside every conv call).
 * Host function, no CUDA context touched (fork-safe for DataLoader workers, data.py:177-184).
 * coo = int64 [2,E_in] (row 0 sources, row 1 targets).  col/cscrow/perm need E_in+N slots.
 * Writes *E_out (= kept edges + N) and *max_deg. */
int bg_csr_build_host(const int64_t* coo, int64_t E_in, int64_t N, int32_t* rowptr, int32_t* col,
                      int32_t* cscptr, int32_t* cscrow, int32_t* perm, int64_t* E_out, int32_t* max_deg);

/* ---- H2: type-matched program features (reference models.py:122-129, 230-237).
 * table[t,:] = mean of local_x rows with local_type==t (0 if none), K types, F features.
 * The per-voxel gather table[voxel.type] is fused into the consumers (BgSeg.gather). */
int bg_type_table(const float* local_x, const int64_t* local_type, int64_t M, int32_t F, int32_t K,
                  float* table, void* stream);
/* Backward of the gather: out[t, 0:C] = sum over rows n with type[n]==t of g[n*ld + 0:C]. */
int bg_type_scatter_sum(const float* g, int64_t ld, const int32_t* type, int64_t N, int32_t C, int32_t K,
                        float* out, float* workspace, size_t ws_bytes, void* stream);
size_t bg_type_scatter_sum_ws(int64_t N, int32_t C, int32_t K);

/* ---- H3/H4/H6/H8/H10 and the `lin` of every conv: node-wise dense layers
 * (reference nn.Linear / nn.LayerNorm / nn.LeakyReLU / nn.ReLU at models.py:33-66,92-113,
 * 177-185,212-225 and torch.cat at models.py:135-141,146,239 - the concatenation is never
 * materialised: the input is a list of column segments). */
typedef struct BgSeg {
    const float* ptr;      /* NULL => a column of ones (used to fold the bias into wgrad) */
    const int32_t* gather; /* optional row index: row n reads ptr[gather[n]*ld ...]       */
    int32_t width;
    int32_t ld; /* row stride in floats */
} BgSeg;
#define BG_MAX_SEG 6
#define BG_MAX_WGRAD 4
#define BG_ACT_NONE 0
#define BG_ACT_RELU 1
#define BG_ACT_LRELU 2 /* LeakyReLU(0.2) */

typedef struct BgDense {
    int64_t N;
    int32_t nseg;
    BgSeg seg[BG_MAX_SEG]; /* K = sum of widths */
    const float* W;        /* element (o,k) at W[o*w_so + k*w_sk]: [Cout,K] row-major => (K,1);
                              a transposed use (backward-input) => (1, Cout_fwd)               */
    int64_t w_so, w_sk;
    int32_t Cout;
    const float* bias;                   /* [Cout] or NULL */
    const float* ln_gamma;               /* [Cout] or NULL => no LayerNorm (eps 1e-5) */
    const float* ln_beta;
    int32_t act;
    const float* att_src;                /* [Cout] or NULL: also emit s = y.att_src, d = y.att_dst */
    const float* att_dst;
    float* out; int64_t ld_out;          /* [N,Cout] */
    float* xhat;                         /* optional [N,Cout] LayerNorm normalised value (saved for backward) */
    float* rstd;                         /* optional [N] */
    float* s; float* d;                  /* optional [N] each */
    const float* gate;                   /* optional [N,Cout] (leading dimension ld_gate): out *= gate > 0 ? 1 : gate_slope -
                                            the activation backward of the layer BELOW fused into this backward-input
                                            product (gate = that layer's saved output) */
    int64_t ld_gate;
    float gate_slope;
} BgDense;
int bg_dense_fwd(const BgDense* a, void* stream);
/* 128-wide (K >= 128) and 64-wide (K >= 256) layers with plain row-major weights run on the tensor cores: tcgen05.mma with
 * accumulators in TMEM (csrc/bg_dense_tc.cu).  mode 1 (default): kind::tf32 with the 3xTF32 split (hi*hi + lo*hi + hi*lo) into
 * four TMEM accumulators combined in fp32 round-to-nearest - fp32-accurate, per-layer error <= 1e-5 of max|y| (measured
 * 2.4e-7 .. 3.4e-7 for K = 128 .. 524, FFMA kernel: 3.4e-7 .. 7.6e-7).  mode 2 (BG_DENSE_TC=bf16): kind::f16 with bf16
 * operands, fp32 accumulation - the stated reduced-precision mode, per-layer error <= 1e-2 (measured 2.2e-3 .. 2.5e-3).
 * mode 0 (BG_DENSE_TC=0): FP32 FFMA kernels everywhere (strict parity mode).  Returns the previous mode. */
int bg_set_dense_tc(int32_t mode);
/* Small layers (K <= 128, Cout <= 64) of graphs up to 131072 rows - the latency-bound regime of the training step - run on
 * a row-per-thread FFMA kernel (csrc/bg_rowdense.cu) - by default only the very narrow ones (K <= 16, Cout <= 8: the 1/2/4-channel
 * bottleneck blocks, the critic's 8 -> 1 score layer), 2 lifts that limit; 0 (or BG_ROWDENSE=0) keeps them on the tiled kernel.  Same results up
 * to fp32 summation order inside LayerNorm / the attention dots.  Returns the previous setting. */
int bg_set_rowdense(int32_t on);
/* Single-segment layers with K in {8,16,...,128}, Cout in {8,16,32,64} of graphs up to 131072 rows run on a warp-level
 * tensor-core kernel (mma.sync m16n8k8, 3xTF32 split with separate main / correction accumulators: fp32-accurate,
 * csrc/bg_dense_mma.cu): these launches are instruction-issue bound, one MMA replaces 32 warp-FFMAs.  0 (or BG_DENSE_MMA=0)
 * switches it off.  Returns the previous setting. */
int bg_set_dense_mma(int32_t on);

/* A chain of n <= BG_SMALL_MAX_LAYERS Linear (+LayerNorm eps 1e-5) (+activation) layers on rows <= BG_SMALL_MAX_ROWS rows as ONE
 * single-CTA launch per direction (csrc/bg_smallmlp.cu): the generator's matched_features_encoder on the K = 7 rows of the
 * type-matched program table (reference models.py:36-47,131-133), where a launch per layer is pure latency.  Widths 1..128,
 * layer i+1 reads layer i's `out`.  Forward saves out / xhat / rstd like bg_dense_fwd.  Backward takes gout[rows, cout_last],
 * writes (accumulate != 0: adds into) dW / dbias / dgamma / dbeta of every layer whose pointers are non-null and, when gin is
 * non-null, the input gradient gin[rows, cin_0].  Deterministic (fixed summation order). */
#define BG_SMALL_MAX_LAYERS 8
#define BG_SMALL_MAX_ROWS 8
typedef struct BgSmallLayer {
    const float* W;       /* [cout, cin] row-major */
    const float* bias;    /* [cout] or NULL */
    const float* gamma;   /* LayerNorm weight [cout] or NULL => no LayerNorm */
    const float* beta;
    int32_t cin, cout, act;
    float* out;           /* [rows, cout] */
    float* xhat;          /* [rows, cout], LayerNorm layers (forward: optional; backward: required) */
    float* rstd;          /* [rows] */
    float* dW;            /* backward outputs, each optional */
    float* dbias;
    float* dgamma;
    float* dbeta;
} BgSmallLayer;
int bg_small_mlp_fwd(const BgSmallLayer* layers, int32_t n, const float* x, int32_t rows, void* stream);
int bg_small_mlp_bwd(const BgSmallLayer* layers, int32_t n, const float* x, int32_t rows, const float* gout, float* gin,
                     int32_t accumulate, void* stream);

/* Weight gradient: dW[o,k] = sum_n gz[n,o] * X[n,k] over the segment list X (a ones segment
 * yields the bias gradient as an extra column); deterministic split-N reduction.
 * dW is written with leading dimension ld_dw; accumulate!=0 adds into dW (and dbias).
 * dbias != NULL routes the LAST column of X (put the ones segment there) to dbias[o] and dW
 * receives the first K-1 columns. */
typedef struct BgWgrad {
    int64_t N;
    const float* gz; int64_t ld_gz; int32_t Cout;
    int32_t nseg;
    BgSeg seg[BG_MAX_SEG];
    float* dW; int64_t ld_dw;
    float* dbias;
    int32_t accumulate;
    float* workspace; size_t ws_bytes;
} BgWgrad;
size_t bg_dense_wgrad_ws(int64_t N, int32_t Cout, int32_t K);
int bg_dense_wgrad(const BgWgrad* a, void* stream);
/* Up to BG_MAX_WGRAD weight-gradient problems over the same N rows in ONE launch (the workspace / ws_bytes
 * fields of the individual problems are ignored; all must share N and accumulate).  The first 4096 bytes of
 * every reduction workspace hold self-resetting ticket counters: zero them once, never again. */
size_t bg_wgrad_multi_ws(int64_t N, int32_t nprob, const int32_t* Cout, const int32_t* K);
int bg_wgrad_multi(const BgWgrad* probs, int32_t nprob, float* workspace, size_t ws_bytes, void* stream);

/* Backward of LayerNorm + LeakyReLU(0.2) (or of a bare activation when xhat==NULL):
 * gz = d loss / d (x W^T + b) from gout, the layer output `out`, xhat, rstd, gamma.
 * Also column sums dgamma = sum gy*xhat, dbeta = sum gy (deterministic). */
int bg_ln_act_bwd(const float* gout, const float* out, const float* xhat, const float* rstd,
                  const float* gamma, int64_t N, int32_t C, int32_t act, float* gz,
                  float* dgamma, float* dbeta, int32_t accumulate, float* workspace, size_t ws_bytes, void* stream);
size_t bg_ln_act_bwd_ws(int64_t N, int32_t C);

/* ---- H14: the non-default conv types of GENERATOR_CONV_TYPE / DISCRIMINATOR_CONV_TYPE (reference models.py:22-31,
 * 166-175; config.py:89,93).  GCNConv: h = x W^T (bg_dense_fwd), out = bg_spmm(w = bg_gcn_norm, h) + bias.
 * GraphConv: out = lin_rel(bg_spmm(no weights, no self loops, x)) + lin_root(x).  Both aggregations are linear, so
 * the backward is the transposed product (transpose = 1) and the second-order backward the forward product again. */
/* w[E'] (CSR order) = deg(dst)^-1/2 deg(src)^-1/2, deg = in-degree incl. the self loop (PyG gcn_norm, fill 1). */
int bg_gcn_norm(const BgGraph* g, float* w, void* stream);
/* out[r,:] = sum over row r of w[edge] * x[other end,:] (+ bias).  transpose 0: rows = destinations (CSR);
 * 1: rows = sources (CSC, weights looked up through perm).  w NULL = ones.  self_loops 0 skips the self loop. */
int bg_spmm(const BgGraph* g, const float* w, const float* x, const float* bias, float* out, int32_t C,
            int32_t transpose, int32_t self_loops, void* stream);
/* GATv2Conv (heads=1, share_weights=False): xl = lin_l(x), xr = lin_r(x) come from bg_dense_fwd;
 * out_i = sum_e softmax_i(att . LeakyReLU(xl_j + xr_i)) xl_j + bias.  Saves logit[E'] (CSR order), m[N], z[N]. */
int bg_gatv2_fwd(const BgGraph* g, const float* xl, const float* xr, const float* att, const float* bias, float* out,
                 float* logit, float* m, float* z, int32_t C, float slope, void* stream);
/* First-order backward: gxl[N,C], gxr[N,C], garow[N,C] (column sum = d loss / d att), scratch P[E'], DL[E']. */
int bg_gatv2_bwd(const BgGraph* g, const float* gout, const float* xl, const float* xr, const float* att,
                 const float* logit, const float* m, const float* z, float* P, float* DL, float* gxl, float* gxr,
                 float* garow, int32_t C, float slope, void* stream);
/* Second-order backward (WGAN-GP): cotangents Hl, Hr on (gxl, gxr) -> gt (on gout), cxl, cxr (on xl, xr),
 * carow[N,C] (column sum = cotangent on att).  scratch: 6*E' floats. */
int bg_gatv2_bwd2(const BgGraph* g, const float* Hl, const float* Hr, const float* gout, const float* xl,
                  const float* xr, const float* att, const float* logit, const float* m, const float* z,
                  float* scratch, float* gt, float* cxl, float* cxr, float* carow, int32_t C, float slope,
                  void* stream);

/* ---- H5a/H9: GATConv attention aggregation (PyG 2.6.1 GATConv.edge_update + message +
 * aggregate; reference call sites models.py:72,82,192,202).  h = x W^T, s = h.a_src,
 * d = h.a_dst come from bg_dense_fwd.  out_i = sum_e softmax_i(LeakyReLU(s_j+d_i)) h_j + bias.
 * Saves the per-row softmax max `m` and denominator `z` (incl. PyG's +1e-16). */
int bg_gat_fwd(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias,
               float* out, float* m, float* z, int32_t C, float slope, void* stream);
/* bg_gat_fwd with the neighbour rows gathered by the TMA: cp.async.bulk.tensor.2d tile::gather4 into an mbarrier-pipelined
 * shared-memory ring, consumer warps doing softmax + weighted sum out of shared memory (csrc/bg_gat_tma.cu).  Same
 * arguments and results.  Eligible: C in {64, 128}, max in-degree (incl. the self loop) <= 8, 16-byte aligned h / out;
 * otherwise BG_EUNSUPPORTED.  bg_gat_fwd itself takes this path for HBM-sized graphs when bg_set_gat_tma(1) / BG_GAT_TMA=1. */
int bg_gat_fwd_tma(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias,
                   float* out, float* m, float* z, int32_t C, float slope, void* stream);
int bg_set_gat_tma(int32_t on);
/* The same aggregation with the statistics of the GraphNorm that follows it (models.py:72-73) fused into its epilogue:
 * gn_stats[3C] = (mean, rstd, var of out - mean_scale*mean) over all N rows, as bg_graphnorm_fwd computes them, so that
 * the normalisation is one elementwise pass (bg_graphnorm_apply) and `out` is not re-read for its moments.  The moment
 * sums are shifted by a sample of the column (row 0's aggregate, recomputed by every CTA) and folded in a fixed order
 * (deterministic).  workspace: bg_gat_fwd_gn_ws bytes,
 * first 4096 bytes zero on first use. */
size_t bg_gat_fwd_gn_ws(int64_t N, int32_t C);
int bg_gat_fwd_gn(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias, float* out,
                  float* m, float* z, int32_t C, float slope, const float* gn_alpha, float gn_eps, float* gn_stats,
                  float* workspace, size_t ws_bytes, void* stream);
/* First-order backward.  Inputs gout[N,C] (= d loss/d out), h, s, d, m, z, a_src, a_dst.
 * Outputs: gh_tot[N,C] = d loss/d h including the s- and d-paths, gsd[N,2] = (d loss/d s,
 * d loss/d d) for the attention-vector gradients, and the per-edge scratch P[E], DU[E]
 * (softmax weights and d loss/d logit, CSR order) that bwd2 reuses. */
int bg_gat_bwd(const BgGraph* g, const float* gout, const float* h, const float* s, const float* d,
               const float* m, const float* z, const float* a_src, const float* a_dst,
               float* P, float* DU, float* gh_tot, float* gsd, int32_t C, float slope, void* stream);
/* The same backward with the elementwise half of the preceding GraphNorm backward fused into the destination pass:
 * go[N,C] = d loss / d (conv output) is computed from (gx1, o, x1, the GraphNorm parameters, stats, and the bstats written
 * by bg_graphnorm_bwd_moments) (+ inj_o, an optional injected cotangent) by the lane group that needs it, and written once
 * for the source pass and the bias gradient.  Equals bg_graphnorm_bwd (+ bg_axpy) followed by bg_gat_bwd. */
int bg_gat_bwd_gn(const BgGraph* g, const float* gx1, const float* o, const float* x1, const float* gn_w,
                  const float* gn_alpha, const float* gn_stats, const float* gn_bstats, float keep_scale,
                  const float* inj_o, const float* h, const float* s, const float* d, const float* m, const float* z,
                  const float* a_src, const float* a_dst, float* P, float* DU, float* go, float* gh_tot, float* gsd,
                  int32_t C, float slope, void* stream);
int bg_graphnorm_bwd_moments(const float* gx1, const float* o, const float* x1, const float* w, const float* alpha,
                             const float* stats, float keep_scale, int64_t N, int32_t C, float* dparams,
                             int32_t accumulate, float* bstats, float* workspace, size_t ws_bytes, void* stream);
/* Second-order backward (WGAN-GP, reference trainer.py:306-312 create_graph=True): given the
 * cotangents Ht[N,C], St[N], Dt[N] on (gh, gs, gd) of bg_gat_bwd, returns
 * gt[N,C] (cotangent on gout), ht_tot[N,C] (cotangent on h incl. s/d paths) and sdt[N,2]
 * (cotangents on s, d).  scratch: 4*E floats. */
int bg_gat_bwd2(const BgGraph* g, const float* Ht, const float* St, const float* Dt, const float* gout,
                const float* h, const float* s, const float* d, const float* m, const float* z,
                const float* a_src, const float* a_dst, float* scratch, float* gt, float* ht_tot,
                float* sdt, int32_t C, float slope, void* stream);

/* ---- H5b/H5c/H9: GraphNorm (called with batch=None => ONE segment over all nodes, reference
 * models.py:73,83,193,203) fused with ReLU(inplace) and Dropout(0.2) (models.py:74-75).
 * y = w*(o - alpha*mu)*rstd + beta ; x1 = relu(y) * keep / keep_prob.
 * stats[3*C] = (mu, rstd, var) is written by fwd and read by bwd/bwd2.  Dropout mask, three modes:
 * keep != NULL: explicit uint8 [N,C] Bernoulli mask drawn by the caller (torch RNG, reference draw
 * order); keep == NULL && keep_prob < 1: Philox4x32-10 mask from (seed, offset) generated in the
 * kernel; keep == NULL && keep_prob == 1: eval (no dropout).  The backward kernels take
 * keep_scale = 1/keep_prob (1 in eval) and recover the mask from x1 > 0. */
int bg_graphnorm_apply(const float* o, const float* w, const float* beta, const float* alpha, const float* stats,
                       const uint8_t* keep, float keep_prob, uint64_t seed, uint64_t offset, int64_t N, int32_t C,
                       float* x1, void* stream); /* the elementwise half alone (statistics from bg_gat_fwd_gn) */
int bg_graphnorm_fwd(const float* o, const float* w, const float* beta, const float* alpha,
                     const uint8_t* keep, float keep_prob, uint64_t seed, uint64_t offset, int64_t N, int32_t C,
                     float eps, float* x1, float* stats, float* workspace, size_t ws_bytes, void* stream);
/* gx1 -> go, plus parameter gradients dparams[3*C] = (dw, dbeta, dalpha) and bstats[2*C] =
 * (G0, G1) saved for bwd2. */
int bg_graphnorm_bwd(const float* gx1, const float* o, const float* x1, const float* w, const float* alpha,
                     const float* stats, float keep_scale, int64_t N, int32_t C, float* go, float* dparams,
                     int32_t accumulate, float* bstats, float* workspace, size_t ws_bytes, void* stream);
/* Xt = cotangent on go.  Outputs: gx1t (cotangent on gx1), ot (cotangent on o), dparams2[3*C]
 * (cotangents on w, beta(=0), alpha; accumulated when accumulate!=0). */
int bg_graphnorm_bwd2(const float* Xt, const float* gx1, const float* o, const float* x1, const float* w,
                      const float* alpha, const float* stats, const float* bstats, float keep_scale,
                      int64_t N, int32_t C, float* gx1t, float* ot, float* dparams2, int32_t accumulate,
                      float* workspace, size_t ws_bytes, void* stream);
size_t bg_graphnorm_ws(int64_t N, int32_t C);

/* ---- H7: Gumbel-softmax (tau=1) + straight-through one-hot (reference models.py:150-153).
 * noise = Gumbel(0,1) samples drawn by the caller, or NULL: drawn in the kernel from Philox(seed, offset).
 * hard = (onehot(argmax soft) - soft) + soft. */
int bg_gumbel_st_fwd(const float* logits, const float* noise, uint64_t seed, uint64_t offset, int64_t N,
                     int32_t K, float* soft, float* hard, int32_t* argmax, void* stream);
int bg_gumbel_st_bwd(const float* g_hard, const float* g_soft, const float* soft, int64_t N, int32_t K,
                     float* g_logits, void* stream);

/* ---- generic segment primitives named by north_star (same machinery as the edge softmax and
 * the GraphNorm statistics, exposed with an arbitrary segment pointer). */
int bg_segment_softmax(const float* v, const int32_t* seg_ptr, int64_t S, float* out, void* stream);
/* N1 (trainer.py:387-443 `_compute_metrics`: one sklearn call per building + 4 per batch, each with D2H copies):
 * cm[S,K,K] (int32), cm[s,t,p] = rows of segment s with target class t and predicted class p = argmax(score row). */
int bg_segment_confusion(const float* score, const int64_t* target, const int32_t* seg_ptr, int64_t S, int32_t K,
                         int32_t* cm, void* stream);
/* mode 0 = mean, 1 = max, 2 = sum; x[N,C] -> out[S,C] */
int bg_segment_pool(const float* x, const int32_t* seg_ptr, int64_t S, int32_t C, int32_t mode,
                    float* out, void* stream);

/* ---- whole-pass executors: ONE call = one forward / backward / second-order-backward pass of the reference's
 * VoxelGNNGenerator (models.py:119-155) or VoxelGNNDiscriminator (models.py:229-245) as a fixed sequence of
 * launches on `stream`.  `params` = device pointers in torch's named_parameters() order of the reference
 * modules; `grad_off[i]` = offset (floats) of parameter i's gradient inside `grad_flat` (GraphNorm's
 * weight/bias/mean_scale and GATConv's att_src/att_dst must be back to back).  `ws` holds the activations the
 * backward needs (size from *_fwd_ws, caller-owned, must stay alive until the backward ran); `red` is a
 * reduction workspace whose first 4096 bytes are zero on first use (>= 8 MiB); `tmp` is scratch.
 * Dropout: training!=0 with keeps==NULL draws Philox masks from (seed, offset + block index); keeps[k] != NULL
 * supplies explicit uint8 [N,C_k] masks (reference RNG order); training==0 disables dropout.
 * Gumbel noise: `noise` [N,K] explicit, or NULL => Philox(seed, offset + 1000). */
typedef struct BgModelDesc {
    int32_t local_dim, voxel_dim, num_classes, z_dim;
    int32_t le_dim, le_layers;              /* LOCAL_ENCODER_HIDDEN_DIM, LOCAL_GRAPH_ENCODER_REPEAT + 1 */
    int32_t g_hidden, g_mlp_layers, g_repeat; /* GENERATOR_HIDDEN_DIM, GENERATOR_MLP_ENCODER_REPEAT + 1, GENERATOR_ENCODER_REPEAT */
    int32_t d_hidden, d_repeat;
} BgModelDesc;
typedef struct BgBatchIn {
    const float* table;    /* [num_classes, local_dim] type table (bg_type_table) */
    const int32_t* type32; /* [N] voxel type */
    const float* vx;       /* [N, voxel_dim] voxel features */
} BgBatchIn;
int32_t bg_gen_num_params(const BgModelDesc* md);
int32_t bg_disc_num_params(const BgModelDesc* md);
size_t bg_gen_fwd_ws(const BgModelDesc* md, int64_t N, int64_t E);
size_t bg_gen_bwd_ws(const BgModelDesc* md, int64_t N, int64_t E);
int bg_gen_forward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                   const float* z, const float* noise, const uint8_t* const* keeps, int32_t training, uint64_t seed,
                   uint64_t offset, void* ws, size_t ws_bytes, float* red, size_t red_bytes, float* logits, float* hard,
                   float* soft, void* stream);
int bg_gen_backward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                    const float* z, const void* ws_fwd, const float* logits, const float* soft, const float* g_logits,
                    const float* g_hard, const float* g_soft, int32_t training, float* grad_flat, const int64_t* grad_off,
                    int32_t accumulate, void* tmp, size_t tmp_bytes, float* red, size_t red_bytes, void* stream);
size_t bg_disc_fwd_ws(const BgModelDesc* md, int64_t N, int64_t E);
size_t bg_disc_bwd_saved_ws(const BgModelDesc* md, int64_t N, int64_t E);
size_t bg_disc_tmp_ws(const BgModelDesc* md, int64_t N, int64_t E);
int bg_disc_forward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                    const float* label, const uint8_t* const* keeps, int32_t training, uint64_t seed, uint64_t offset,
                    void* ws, size_t ws_bytes, float* red, size_t red_bytes, float* score, void* stream);
/* grad_flat may be NULL (input gradient only); accumulate != 0 adds the parameter gradients into grad_flat instead of
 * overwriting it; saved != NULL keeps the intermediates bg_disc_backward2 needs. */
int bg_disc_backward(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                     const float* label, const void* ws_fwd, const float* score, const float* g_score, int32_t training,
                     float* grad_flat, const int64_t* grad_off, int32_t accumulate, void* saved, size_t saved_bytes, void* tmp,
                     size_t tmp_bytes, float* red, size_t red_bytes, float* g_label, void* stream);
/* WGAN-GP (reference trainer.py:306-312): Lt = cotangent on g_label; grad_flat must be zero on entry and receives
 * the parameter cotangents (second-order sweep + forward-graph sweep with injections); tmp >= 2*bg_disc_tmp_ws. */
int bg_disc_backward2(const BgModelDesc* md, const float* const* params, const BgGraph* graph, const BgBatchIn* in,
                      const float* label, const void* ws_fwd, const float* score, const void* saved, const float* Lt,
                      int32_t training, float* grad_flat, const int64_t* grad_off, void* tmp, size_t tmp_bytes, float* red,
                      size_t red_bytes, float* gt_score, void* stream);

/* Test hooks: byte offsets of the saved post-activation tensors inside a forward workspace, layer order
 * (generator: menc, mlp, conv x1, dec; discriminator: pre, conv x1, dec).  Returns the count. */
int32_t bg_gen_ws_offsets(const BgModelDesc* md, int64_t N, int64_t* out, int32_t cap);
int32_t bg_disc_ws_offsets(const BgModelDesc* md, int64_t N, int64_t* out, int32_t cap);

/* ---- H11 / H12: the loss glue of one critic update (reference trainer.py:291-332).
 * bg_gp_mix:          mixed = e * onehot + (1 - e) * soft   (trainer.py:298-301; onehot int64 [N,K] as the reference holds it, or fp32)
 * bg_critic_loss_fwd: out4 = {loss, mean D(fake), mean D(real), gp},  loss = mean(d_fake) - mean(d_real) + lambda * mean((||grad_i||_2 - 1)^2)
 *                     (trainer.py:314, 323); coef[N] is saved for the backward.  Deterministic fold, no atomics on floats.
 * bg_critic_loss_bwd: g_fake[i] = g/N, g_real[i] = -g/N, g_grad[i,:] = g * coef[i] * grad[i,:]  (any output may be NULL); g = *g_loss. */
int bg_gp_mix(const float* e, const void* onehot, int32_t onehot_is_i64, const float* soft, int64_t N, int32_t K, float* mixed,
              void* stream);
size_t bg_critic_loss_ws(int64_t N);
int bg_critic_loss_fwd(const float* d_fake, const float* d_real, const float* grad, int64_t N, int32_t K, float lambda, float* coef,
                       float* workspace, size_t ws_bytes, float* out4, void* stream);
int bg_critic_loss_bwd(const float* g_loss, const float* coef, const float* grad, int64_t N, int32_t K, float* g_fake, float* g_real,
                       float* g_grad, void* stream);

/* ---- process-wide knobs (the only library state besides the per-(device, stream) weight-gradient side streams).
 * bg_set_pdl: programmatic dependent launch on/off at run time (initial value: env BG_PDL, default on); returns the previous
 * setting.  bg_set_rng_base: device pointer to a uint64 that every in-kernel Philox offset adds (NULL = none, the default):
 * set it while capturing a CUDA graph so that replays draw fresh dropout masks / Gumbel noise (the graph owner bumps the
 * value between replays), reset it to NULL afterwards. */
int bg_set_pdl(int32_t on);
int bg_set_rng_base(const uint64_t* base);

/* ---- small utilities used by the host-side executor */
int bg_axpy(float* y, const float* x, float a, int64_t n, void* stream);          /* y += a*x */
int bg_fill(float* y, float v, int64_t n, void* stream);

/* ---- H15: torch.optim.Adam.step (reference train.py:36-37, trainer.py:481,495; amsgrad / maximize off) over ONE flat
 * parameter buffer: p, g, m (exp_avg), v (exp_avg_sq) are index-aligned flat fp32 buffers of n elements (n % 4 == 0,
 * 16-byte aligned).  step counts from 1 (the value torch's state["step"] has AFTER its increment); if step_dev is not
 * NULL the count is read from that device counter instead (CUDA-graph replay: the caller increments it in-graph). */
int bg_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                 double weight_decay, int64_t step, const int64_t* step_dev, void* stream);

/* Data-parallel gradient exchange fused with the optimiser step (replaces the flat-bucket ncclAllReduce +
 * scale + bg_adam_flat sequence, i.e. what DistributedDataParallel + torch.optim.Adam.step would run around the reference's
 * train.py:36-37 / trainer.py:481,495).  One launch: waits until every rank's gradient bucket is complete, reads all `world`
 * buckets out of the peers' memory over NVLink / NVSwitch (one-shot all-reduce, summed in rank order => bit-identical on every
 * rank), averages, and applies Adam to the LOCAL flat p / m / v (same arithmetic as bg_adam_flat); with p = m = v = NULL it only
 * writes the averaged gradient to `gavg` (optional otherwise).  grad[r] / flags[r]: rank r's gradient bucket (n floats, 16-byte
 * aligned) and its flag array (2 * world uint32, zero before the first call), both in memory every rank can address (symmetric
 * memory / CUDA IPC; the host side uses torch.distributed._symmetric_memory).  epoch, ticket: LOCAL device words, zero before
 * the first call; the kernel advances them itself, so the launch is CUDA-graph capturable.  Every rank must make the same
 * sequence of calls (like a collective).  No NCCL types: this replaces the `bg_allreduce_flat(ncclComm_t, ...)` entry the
 * survey sketched - the host keeps torch.distributed / NCCL for everything else (barriers, the max-over-ranks timing). */
#define BG_MAX_PEERS 8
typedef struct {
    const float* grad[BG_MAX_PEERS];
    uint32_t* flags[BG_MAX_PEERS];
    int32_t rank, world;
} BgPeers;
int bg_p2p_allreduce_adam(const BgPeers* peers, uint32_t* epoch, uint32_t* ticket, float* p, float* m, float* v, float* gavg,
                          int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                          const int64_t* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BG_B200_H */
