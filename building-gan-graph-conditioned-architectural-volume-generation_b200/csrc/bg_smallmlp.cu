// A chain of Linear (+LayerNorm) (+activation) layers on a HANDFUL of rows, forward and backward, as ONE launch each.
//
// Scope: the generator's `matched_features_encoder` (reference models.py:36-47,131-133) evaluated on the K = 7 rows of the
// type-matched program table (row-wise layers commute with the per-voxel gather): five 128-wide
// Linear + LayerNorm + LeakyReLU layers on 7 rows.  Through the tiled dense kernels that was 5 launches of ~20 us forward
// (a 64-row tile with 7 live rows, K in slabs of 16) and 5 x (LayerNorm backward + backward-input product) + 5 weight-gradient
// problems backward: ~100 us + ~125 us of single-CTA latency on the generator's critical path per training step
// (profiles/r02b_summary.md).  The whole chain is ~80 k multiply-adds per layer: one CTA does all layers back to back with the
// activations in shared memory - one launch forward, one launch backward (parameter gradients included).
//
// Forward, per layer: W[cout, cin] is copied to shared memory as it lies (coalesced); warp w owns output columns w, w+16, ...;
// lane l holds x[r][l + 32 j] of all rows in registers, multiplies by W[c][l + 32 j] (conflict-free: consecutive lanes read
// consecutive words), and the 8 row sums are reduced over the warp by a transposing butterfly (9 shuffles per column); then
// warp r normalises row r (two-pass mean / variance, eps = 1e-5, like bg_dense_fwd), applies the activation and saves out /
// xhat / rstd exactly where the separate launches would have.
// Backward, per layer (top to bottom): activation + LayerNorm backward per row (warp r), then dgamma / dbeta / dbias per column,
// dW[c][k] = sum_r gz[r][c] x[r][k] (8 multiply-adds per element, coalesced read-modify-write into the gradient bucket), and the
// backward-input product gin[r][k] = sum_c gz[r][c] W[c][k] with W read coalesced from global memory (four column quarters
// summed through shared memory in a fixed order).  Everything is summed in a fixed order: bitwise reproducible.
#include "bg_common.cuh"

namespace bg {

constexpr int SM_T = 512;       // threads of the single CTA
constexpr int SM_W = SM_T / 32;
constexpr int SM_R = BG_SMALL_MAX_ROWS;
constexpr int SM_C = 128;       // widest layer

struct SmallChain {
    int n, rows, accumulate;
    const float* x;      // [rows, layers[0].cin]
    const float* gout;   // backward: [rows, layers[n-1].cout]
    float* gin;          // backward: [rows, layers[0].cin] or null
    BgSmallLayer l[BG_SMALL_MAX_LAYERS];
};

__device__ __forceinline__ float act_fwd(float v, int act) {
    return act == BG_ACT_RELU ? (v > 0.f ? v : 0.f) : act == BG_ACT_LRELU ? (v > 0.f ? v : 0.2f * v) : v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// W[cout * cin] -> 32 registers per thread (element i = tid + 512 j, or float4 i = tid + 512 j when the buffer allows 128-bit
// loads): every load of a layer's weights is in flight at once - a single CTA streaming 64 KB through a loop of dependent
// round trips was the whole cost of the first version (6 us per layer).
struct WRegs {
    float v[32];
};
__device__ __forceinline__ bool w_vec4(const float* W, int total) { return (total & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0; }
__device__ __forceinline__ void w_fetch(WRegs& r, const float* __restrict__ W, int total, int tid) {
    if (w_vec4(W, total)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = tid + SM_T * j;
            const float4 q = i < total / 4 ? __ldg(reinterpret_cast<const float4*>(W) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            r.v[4 * j] = q.x; r.v[4 * j + 1] = q.y; r.v[4 * j + 2] = q.z; r.v[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int i = tid + SM_T * j;
            r.v[j] = i < total ? __ldg(W + i) : 0.f;
        }
    }
}
__device__ __forceinline__ void w_store(const WRegs& r, float* sm, const float* W, int total, int tid) {
    if (w_vec4(W, total)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = tid + SM_T * j;
            if (i < total / 4) reinterpret_cast<float4*>(sm)[i] = make_float4(r.v[4 * j], r.v[4 * j + 1], r.v[4 * j + 2], r.v[4 * j + 3]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int i = tid + SM_T * j;
            if (i < total) sm[i] = r.v[j];
        }
    }
}

__global__ void __launch_bounds__(SM_T, 1) small_mlp_fwd_kernel(const SmallChain p) {
    pdl_prologue();
    extern __shared__ __align__(16) float sm_w[];  // [cout * cin] of the current layer
    __shared__ float xs[SM_R][SM_C], ys[SM_R][SM_C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WRegs wr;
    w_fetch(wr, p.l[0].W, p.l[0].cin * p.l[0].cout, tid);
    {
        const int cin0 = p.l[0].cin;
        for (int i = tid; i < SM_R * SM_C; i += SM_T) {
            const int r = i / SM_C, k = i % SM_C;
            xs[r][k] = (r < p.rows && k < cin0) ? __ldg(p.x + (int64_t)r * cin0 + k) : 0.f;
        }
    }
    for (int li = 0; li < p.n; ++li) {
        const BgSmallLayer& L = p.l[li];
        const int cin = L.cin, cout = L.cout;
        w_store(wr, sm_w, L.W, cin * cout, tid);
        if (li + 1 < p.n) w_fetch(wr, p.l[li + 1].W, p.l[li + 1].cin * p.l[li + 1].cout, tid);  // in flight during this layer's work
        // per-column constants of this layer's epilogue, fetched before the barrier
        float bias_c[SM_C / SM_W];
#pragma unroll
        for (int q = 0; q < SM_C / SM_W; ++q) {
            const int c = warp + SM_W * q;
            bias_c[q] = (L.bias && c < cout) ? __ldg(L.bias + c) : 0.f;
        }
        float gam[4], bet[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = lane + 32 * j;
            gam[j] = (L.gamma && c < cout) ? __ldg(L.gamma + c) : 1.f;
            bet[j] = (L.gamma && c < cout) ? __ldg(L.beta + c) : 0.f;
        }
        __syncthreads();  // W staged, xs of this layer complete
        float xr[SM_R][4];
#pragma unroll
        for (int r = 0; r < SM_R; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) xr[r][j] = (lane + 32 * j < cin) ? xs[r][lane + 32 * j] : 0.f;
#pragma unroll
        for (int q = 0; q < SM_C / SM_W; ++q) {
            const int c = warp + SM_W * q;
            if (c >= cout) break;
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = (lane + 32 * j < cin) ? sm_w[c * cin + lane + 32 * j] : 0.f;
            float acc[SM_R];
#pragma unroll
            for (int r = 0; r < SM_R; ++r) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) a = fmaf(wv[j], xr[r][j], a);
                acc[r] = a;
            }
            // transposing butterfly: 8 row sums over 32 lanes in 9 shuffles; row index ends up in lane bits 4,3,2
            float b[4], c2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = (lane & 16) ? acc[i] : acc[i + 4], keep = (lane & 16) ? acc[i + 4] : acc[i];
                b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float send = (lane & 8) ? b[i] : b[i + 2], keep = (lane & 8) ? b[i + 2] : b[i];
                c2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            const float send = (lane & 4) ? c2[0] : c2[1], keep = (lane & 4) ? c2[1] : c2[0];
            float d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            if ((lane & 3) == 0) {
                const int r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                ys[r][c] = d + bias_c[q];
            }
        }
        __syncthreads();  // ys complete; every warp is done with xs and sm_w
        if (warp < SM_R) {
            const int r = warp;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (lane + 32 * j < cout) ? ys[r][lane + 32 * j] : 0.f;
            if (L.gamma) {
                const float mean = warp_sum(v[0] + v[1] + v[2] + v[3]) / (float)cout;
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (lane + 32 * j < cout) q = fmaf(v[j] - mean, v[j] - mean, q);
                const float rs = 1.f / sqrtf(warp_sum(q) / (float)cout + 1e-5f);
                if (L.rstd && lane == 0 && r < p.rows) L.rstd[r] = rs;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = lane + 32 * j;
                    if (c < cout) {
                        const float xh = (v[j] - mean) * rs;
                        if (L.xhat && r < p.rows) L.xhat[(int64_t)r * cout + c] = xh;
                        v[j] = fmaf(xh, gam[j], bet[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                if (c < cout) {
                    const float y = act_fwd(v[j], L.act);
                    if (r < p.rows) L.out[(int64_t)r * cout + c] = y;
                    xs[r][c] = r < p.rows ? y : 0.f;
                }
            }
        }
        // the next iteration's w_store overwrites sm_w: every warp left the product loop before the barrier above; its
        // barrier orders the xs writes of this epilogue before the next product
    }
}

__global__ void __launch_bounds__(SM_T, 1) small_mlp_bwd_kernel(const SmallChain p) {
    pdl_prologue();
    __shared__ float gs[SM_R][SM_C];       // gradient at the current layer's output
    __shared__ float gy[SM_R][SM_C];       // after the activation backward
    __shared__ float gz[SM_R][SM_C];       // pre-activation / pre-LayerNorm gradient
    __shared__ float xin[SM_R][SM_C];      // the layer's input rows
    __shared__ float xh[SM_R][SM_C];       // saved normalised values
    __shared__ float part[4][SM_R][SM_C];  // column-quarter partials of the backward-input product
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const int ct = p.l[p.n - 1].cout;
        for (int i = tid; i < SM_R * SM_C; i += SM_T) {
            const int r = i / SM_C, c = i % SM_C;
            gs[r][c] = (r < p.rows && c < ct) ? __ldg(p.gout + (int64_t)r * ct + c) : 0.f;
        }
    }
    for (int li = p.n - 1; li >= 0; --li) {
        const BgSmallLayer& L = p.l[li];
        const int cin = L.cin, cout = L.cout, total = cin * cout;
        const float* xsrc = li == 0 ? p.x : p.l[li - 1].out;
        const bool need_gin = li > 0 || p.gin != nullptr;
        // ---- every global read of the layer is issued here, before the first barrier: one round trip per layer
        float xv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = tid + SM_T * j, r = i / SM_C, k = i % SM_C;
            xv[j] = (r < p.rows && k < cin) ? __ldg(xsrc + (int64_t)r * cin + k) : 0.f;
        }
        // this thread's quarter-column of W for the backward-input product: W[c0 + u][k], u < 32
        const int kq = tid & (SM_C - 1), cq = tid >> 7;
        const int per = (cout + 3) / 4, c0 = cq * per, c1 = min(cout, c0 + per);
        float wq[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) wq[u] = (need_gin && kq < cin && c0 + u < c1) ? __ldg(L.W + (int64_t)(c0 + u) * cin + kq) : 0.f;
        // existing parameter gradients (accumulate mode): element i = tid + 512 j
        float dwv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int i = tid + SM_T * j;
            dwv[j] = (p.accumulate && L.dW && i < total) ? L.dW[i] : 0.f;
        }
        // row r = warp: saved output, normalised value, LayerNorm constants
        float o4[4], h4[4], gm4[4], rs = 0.f;
        {
            const int r = warp;
            const bool live = r < SM_R && r < p.rows;
            if (live && L.gamma) rs = __ldg(L.rstd + r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                const bool ok = live && c < cout;
                o4[j] = ok ? __ldg(L.out + (int64_t)r * cout + c) : 0.f;
                h4[j] = (ok && L.gamma) ? __ldg(L.xhat + (int64_t)r * cout + c) : 0.f;
                gm4[j] = (ok && L.gamma) ? __ldg(L.gamma + c) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = tid + SM_T * j;
            xin[i / SM_C][i % SM_C] = xv[j];
        }
        __syncthreads();  // gs of this layer complete (written by the previous iteration), xin staged
        if (warp < SM_R) {  // row r: activation backward, LayerNorm backward
            const int r = warp;
            const bool live = r < p.rows;
            float g[4], gx[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                g[j] = gx[j] = 0.f;
                if (c < cout && live) {
                    float t = gs[r][c];
                    if (L.act == BG_ACT_LRELU) t = o4[j] > 0.f ? t : 0.2f * t;
                    else if (L.act == BG_ACT_RELU) t = o4[j] > 0.f ? t : 0.f;
                    g[j] = t;
                    if (L.gamma) {
                        gx[j] = t * gm4[j];
                        s1 += gx[j];
                        s2 = fmaf(gx[j], h4[j], s2);
                    }
                }
            }
            if (L.gamma) {
                s1 = warp_sum(s1) / (float)cout;
                s2 = warp_sum(s2) / (float)cout;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                gy[r][c] = g[j];
                xh[r][c] = h4[j];
                gz[r][c] = (c < cout && live) ? (L.gamma ? rs * (gx[j] - s1 - h4[j] * s2) : g[j]) : 0.f;
            }
        }
        __syncthreads();  // gy, xh, gz complete
        // per-column parameter gradients (rows summed in order)
        for (int c = tid; c < cout; c += SM_T) {
            float dg = 0.f, db = 0.f, dbias = 0.f;
#pragma unroll
            for (int r = 0; r < SM_R; ++r) {
                dg = fmaf(gy[r][c], xh[r][c], dg);
                db += gy[r][c];
                dbias += gz[r][c];
            }
            if (L.gamma && L.dgamma) {
                L.dgamma[c] = p.accumulate ? L.dgamma[c] + dg : dg;
                L.dbeta[c] = p.accumulate ? L.dbeta[c] + db : db;
            }
            if (L.dbias) L.dbias[c] = p.accumulate ? L.dbias[c] + dbias : dbias;
        }
        if (L.dW) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int i = tid + SM_T * j;
                if (i < total) {
                    const int c = i / cin, k = i - c * cin;
                    float t = 0.f;
#pragma unroll
                    for (int r = 0; r < SM_R; ++r) t = fmaf(gz[r][c], xin[r][k], t);
                    L.dW[i] = dwv[j] + t;  // dwv = 0 unless accumulating
                }
            }
        }
        if (need_gin) {  // gin[r][k] = sum_c gz[r][c] W[c][k]: thread = (column quarter, k)
            float acc[SM_R];
#pragma unroll
            for (int r = 0; r < SM_R; ++r) acc[r] = 0.f;
            if (kq < cin) {
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    if (c0 + u < c1) {
#pragma unroll
                        for (int r = 0; r < SM_R; ++r) acc[r] = fmaf(gz[r][c0 + u], wq[u], acc[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < SM_R; ++r) part[cq][r][kq] = acc[r];
        }
        __syncthreads();  // part complete; gs free to be overwritten
        if (need_gin) {
            for (int i = tid; i < SM_R * SM_C; i += SM_T) {
                const int r = i / SM_C, k = i % SM_C;
                const float t = (k < cin && r < p.rows) ? ((part[0][r][k] + part[1][r][k]) + part[2][r][k]) + part[3][r][k] : 0.f;
                gs[r][k] = t;
                if (li == 0 && p.gin && k < cin && r < p.rows) p.gin[(int64_t)r * cin + k] = t;
            }
        }
        // the next iteration's first barrier orders these writes (and the xin restaging) before any read
    }
}

static int check_chain(const BgSmallLayer* layers, int n, const float* x, int rows, const char* who) {
    BG_REQUIRE(layers && x, BG_EINVAL, "%s: null pointer", who);
    BG_REQUIRE(n >= 1 && n <= BG_SMALL_MAX_LAYERS, BG_EINVAL, "%s: %d layers (1..%d)", who, n, BG_SMALL_MAX_LAYERS);
    BG_REQUIRE(rows >= 1 && rows <= BG_SMALL_MAX_ROWS, BG_EUNSUPPORTED, "%s: %d rows (1..%d)", who, rows, BG_SMALL_MAX_ROWS);
    for (int i = 0; i < n; ++i) {
        const BgSmallLayer& L = layers[i];
        BG_REQUIRE(L.W && L.out, BG_EINVAL, "%s: layer %d has null pointers", who, i);
        BG_REQUIRE(L.cin >= 1 && L.cin <= SM_C && L.cout >= 1 && L.cout <= SM_C, BG_EUNSUPPORTED,
                   "%s: layer %d is %d -> %d (widths 1..%d)", who, i, L.cin, L.cout, SM_C);
        BG_REQUIRE(i == 0 || L.cin == layers[i - 1].cout, BG_EINVAL, "%s: layer %d input width %d != layer %d output width %d", who, i,
                   L.cin, i - 1, layers[i - 1].cout);
        BG_REQUIRE(!L.gamma || L.beta, BG_EINVAL, "%s: layer %d has a LayerNorm weight without bias", who, i);
    }
    return BG_OK;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_small_mlp_fwd(const BgSmallLayer* layers, int32_t n, const float* x, int32_t rows, void* stream) {
    if (int rc = check_chain(layers, n, x, rows, "bg_small_mlp_fwd")) return rc;
    SmallChain p{};
    p.n = n; p.rows = rows; p.x = x;
    for (int i = 0; i < n; ++i) p.l[i] = layers[i];
    constexpr size_t smem = (size_t)SM_C * SM_C * sizeof(float);
    static bool once = (cudaFuncSetAttribute(small_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    launch_k(small_mlp_fwd_kernel, 1, SM_T, smem, as_stream(stream), p);
    return check_launch("bg_small_mlp_fwd");
}

extern "C" int bg_small_mlp_bwd(const BgSmallLayer* layers, int32_t n, const float* x, int32_t rows, const float* gout, float* gin,
                                int32_t accumulate, void* stream) {
    if (int rc = check_chain(layers, n, x, rows, "bg_small_mlp_bwd")) return rc;
    BG_REQUIRE(gout, BG_EINVAL, "bg_small_mlp_bwd: null pointer");
    for (int i = 0; i < n; ++i)
        BG_REQUIRE(!layers[i].gamma || (layers[i].xhat && layers[i].rstd), BG_EINVAL,
                   "bg_small_mlp_bwd: layer %d has a LayerNorm but no saved xhat / rstd", i);
    SmallChain p{};
    p.n = n; p.rows = rows; p.accumulate = accumulate ? 1 : 0; p.x = x; p.gout = gout; p.gin = gin;
    for (int i = 0; i < n; ++i) p.l[i] = layers[i];
    launch_k(small_mlp_bwd_kernel, 1, SM_T, 0, as_stream(stream), p);
    return check_launch("bg_small_mlp_bwd");
}
