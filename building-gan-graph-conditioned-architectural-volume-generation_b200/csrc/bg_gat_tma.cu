// GAT aggregation forward with the neighbour rows gathered by the TENSOR MEMORY ACCELERATOR:
// cp.async.bulk.tensor.2d ... tile::gather4 (four rows of h per instruction, chosen by four row indices) into an
// mbarrier-pipelined shared-memory ring, consumer warps doing the edge softmax and the weighted sum out of shared memory.
//
// Why.  ncu of the register-path kernel (bg_gat.cu, profiles/r01g_ncu_full_gat_c64_N1e6.csv): at N = 1e6, C = 64 the gather
// goes through the LSU / L1 data pipe (78-84 % busy), every gathered row costs a 128-bit load instruction per lane plus the
// registers to hold it (64 registers x 1024 threads => 50 % of the warp slots), and the kernel sits at 0.63 of the measured
// copy bandwidth, latency-bound on 6.8 long-scoreboard stalls per issue.  The TMA takes the address generation, the
// request tracking and the landing registers out of the SM's hands: one thread issues "rows (j0, j1, j2, j3), columns 0..C"
// and the bytes arrive in shared memory.
//
// Structure (one CTA = 3 producer warps + 4 consumer warps, persistent over row blocks, round-robin so that all CTAs
// advance through the row range together and the floor +-1 neighbours stay L2-resident; two CTAs per SM at C = 64):
//   row block = R = 16 destination rows, 8 edge slots per row (in-degree <= 8 incl. the self loop: 6-neighbour voxel grids
//   have <= 7; graphs with a larger row are sent to the register-path kernel by the host);
//   producer lane (r = lane / 2, q = lane % 2) reads rowptr / col / s / d of its row, writes the 4 logits of slots
//   4q .. 4q+3 (LeakyReLU(s_j + d_i), -inf for padding) to shared memory and issues ONE gather4 for the 4 source rows
//   (padding slots re-read the destination row itself: always in bounds, weight 0).  The producer's own dependent chain
//   rowptr -> col -> s[j] is software-pipelined with two of its iterations between dependent hops (rowptr of its block k+6,
//   col / d of k+4, s[j] of k+2, issue of k), so a load has two iterations to land before it is used; the producer warps
//   take the CTA's blocks in turn;
//   stage = R x 8 x C floats (32 KiB at C = 64) + R x 8 logits; kStages stages; full[stage] counts the TMA bytes
//   (expect_tx) plus the producer's arrival, empty[stage] the consumer warps' arrivals;
//   consumer warps: a group of 8 lanes owns one destination row (4 rows per warp, 16 per pass of the 4 consumer warps): lane
//   `sub` turns the logit of slot `sub` into its softmax weight (max, exp, sum + 1e-16: PyG's softmax, shuffles inside the
//   group), every lane accumulates its C/8 columns over the slots IN CSR ORDER (sources ascending, self loop last: the
//   reference's CPU summation order), adds the bias and stores out / m / z.  (A first version with one warp per row spent
//   186 warp instructions per destination row and was issue-bound at 0.31 of the copy bandwidth.)
// Results equal the register-path kernel's up to the rounding of expf / the reciprocal (tests: rel 1e-5 against fp64).
#include <cuda.h>
#include <stdlib.h>

#include "bg_common.cuh"

namespace bg {

constexpr int kTmaRows = 16;       // destination rows per stage
constexpr int kTmaSlots = 8;       // edge slots per row
constexpr int kTmaStages = 3;
constexpr int kTmaConsumers = 4;   // consumer warps (4 destination rows per warp pass: 16 rows per stage in one pass)
constexpr int kTmaProducers = 3;   // producer warps; producer p issues iterations p, p + P, ... = ALWAYS stage p (kTmaStages == kTmaProducers):
                                   // a stage has one owner, so its empty-barrier phases are consumed strictly in order (no parity aliasing)
static_assert(kTmaStages % kTmaProducers == 0, "every stage needs exactly one producer warp");
constexpr int kTmaThreads = 32 * (kTmaConsumers + kTmaProducers);

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_addr(dst)), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_addr(bar))
        : "memory");
}

template <int C>
__global__ void __launch_bounds__(kTmaThreads, C <= 64 ? 2 : 1) gat_fwd_tma_kernel(const __grid_constant__ CUtensorMap hmap,
                                                                     const int32_t* __restrict__ rowptr,
                                                                     const int32_t* __restrict__ col, const float* __restrict__ s,
                                                                     const float* __restrict__ d, const float* __restrict__ bias,
                                                                     float* __restrict__ out, float* __restrict__ m_out,
                                                                     float* __restrict__ z_out, int N, int E, float slope, int nblocks) {
    pdl_prologue();
    constexpr int ROW_FLOATS = kTmaSlots * C;                     // one destination row's gathered slots
    constexpr int STAGE_FLOATS = kTmaRows * ROW_FLOATS;
    extern __shared__ uint8_t tma_smem_raw[];
    uint8_t* tma_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tma_smem_raw) + 127) & ~uintptr_t(127));  // TMA: 128-byte aligned
    float* hs = reinterpret_cast<float*>(tma_smem);                                            // [stages][R][8][C]
    float* lg = hs + (size_t)kTmaStages * STAGE_FLOATS;                                        // [stages][R][8] logits
    uint64_t* full = reinterpret_cast<uint64_t*>(lg + kTmaStages * kTmaRows * kTmaSlots);      // [stages]
    uint64_t* empty = full + kTmaStages;                                                        // [stages]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int st = 0; st < kTmaStages; ++st) {
            mbar_init_(full + st, 1);
            mbar_init_(empty + st, kTmaConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kTmaConsumers) {
        // ---------------- producer warps: iterations it = pw, pw + P, pw + 2P, ... of this CTA's block sequence
        const int pw = warp - kTmaConsumers;
        const int r = lane >> 1, q = lane & 1;
        const int niter = blockIdx.x < nblocks ? (nblocks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;  // iterations of this CTA
        auto row_of = [&](int it) {  // destination row of this lane in iteration `it` (clamped: loads stay in bounds)
            const int itc = it < niter ? it : niter - 1;
            const int row = ((int)blockIdx.x + itc * (int)gridDim.x) * kTmaRows + r;
            return row < N ? row : N - 1;
        };
        // Software pipeline of the index chain, G producer-iterations between dependent hops: slot 3G holds the iteration whose
        // rowptr loads are being issued, 2G the one whose col / d loads are issued (rowptr arrived G iterations ago), G the one
        // whose s[j] loads are issued, 0 the one being handed to the TMA.  A load has G iterations to land before it is used.
        constexpr int G = 2, DEPTH = 3 * G + 1;
        int begp[DEPTH], degp[DEPTH], rrp[DEPTH], jp[DEPTH][4];
        float sp[DEPTH][4], dp[DEPTH];
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) {
            begp[i] = degp[i] = rrp[i] = 0;
            dp[i] = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) jp[i][e] = 0, sp[i][e] = 0.f;
        }
        if (niter > 0) {
            for (int k = -3 * G;; ++k) {
                const int it = pw + k * kTmaProducers;
                if (it >= niter) break;
                // A (slot 3G): rowptr of iteration it + 3G P
                rrp[3 * G] = row_of(it + 3 * G * kTmaProducers);
                begp[3 * G] = __ldg(rowptr + rrp[3 * G]);
                degp[3 * G] = __ldg(rowptr + rrp[3 * G] + 1) - begp[3 * G];
                // B (slot 2G): col / d of iteration it + 2G P
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    jp[2 * G][e] = 4 * q + e < degp[2 * G] ? __ldg(col + begp[2 * G] + 4 * q + e) : rrp[2 * G];
                dp[2 * G] = __ldg(d + rrp[2 * G]);
                // C (slot G): s[j] of iteration it + G P
#pragma unroll
                for (int e = 0; e < 4; ++e) sp[G][e] = __ldg(s + jp[G][e]);
                // D (slot 0): issue iteration it
                if (k >= 0) {
                    const int st = it % kTmaStages;
                    if (it >= kTmaStages) mbar_wait_(empty + st, ((it / kTmaStages) - 1) & 1);
                    float u[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) u[e] = 4 * q + e < degp[0] ? lrelu(sp[0][e] + dp[0], slope) : -INFINITY;
                    *reinterpret_cast<float4*>(lg + ((size_t)st * kTmaRows + r) * kTmaSlots + 4 * q) = make_float4(u[0], u[1], u[2], u[3]);
                    __syncwarp();
                    if (lane == 0) mbar_expect_tx(full + st, (uint32_t)(STAGE_FLOATS * sizeof(float)));
                    __syncwarp();
                    tma_gather4(hs + (size_t)st * STAGE_FLOATS + (size_t)r * ROW_FLOATS + (size_t)q * 4 * C, &hmap, 0, jp[0][0], jp[0][1],
                                jp[0][2], jp[0][3], full + st);
                }
                // shift the pipeline by one slot
#pragma unroll
                for (int i = 0; i < DEPTH - 1; ++i) {
                    begp[i] = begp[i + 1], degp[i] = degp[i + 1], rrp[i] = rrp[i + 1], dp[i] = dp[i + 1];
#pragma unroll
                    for (int e = 0; e < 4; ++e) jp[i][e] = jp[i + 1][e], sp[i][e] = sp[i + 1][e];
                }
            }
        }
    } else {
        // ---------------- consumer warps: a group of 8 lanes owns one destination row (4 rows per warp pass), lane `sub` holds
        // the logit / weight of edge slot `sub` and the columns {32 p + 4 sub .. +3 : p < C/32} of the row (128-bit pieces: the 8
        // lanes of a row read 128 contiguous bytes per piece, the 4 rows of the warp 4 wavefronts - conflict-free)
        constexpr int NP = C / 32;  // 128-bit pieces per lane
        const int sub = lane & 7, grp = lane >> 3;
        const unsigned gmask = 0xffu << (grp * 8);
        float4 bv[NP];
#pragma unroll
        for (int pc = 0; pc < NP; ++pc)
            bv[pc] = bias ? __ldg(reinterpret_cast<const float4*>(bias + 32 * pc + 4 * sub)) : make_float4(0.f, 0.f, 0.f, 0.f);
        int it = 0;
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
            const int st = it % kTmaStages;
            mbar_wait_(full + st, (it / kTmaStages) & 1);
            for (int r = warp * 4 + grp; r < kTmaRows; r += kTmaConsumers * 4) {
                const int row = blk * kTmaRows + r;
                const float l = lg[((size_t)st * kTmaRows + r) * kTmaSlots + sub];
                float mx = l;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, o));
                float pq = expf(l - mx);  // exp(-inf) = 0 for padding slots
                float zs = pq;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) zs += __shfl_xor_sync(gmask, zs, o);
                zs += 1e-16f;
                pq = pq / zs;
                float4 acc[NP];
#pragma unroll
                for (int pc = 0; pc < NP; ++pc) acc[pc] = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* hrow = hs + (size_t)st * STAGE_FLOATS + (size_t)r * ROW_FLOATS + 4 * sub;
#pragma unroll
                for (int e = 0; e < kTmaSlots; ++e) {  // CSR order; padding slots carry weight 0
                    const float pe = __shfl_sync(gmask, pq, grp * 8 + e);
#pragma unroll
                    for (int pc = 0; pc < NP; ++pc) {
                        const float4 hv = *reinterpret_cast<const float4*>(hrow + e * C + 32 * pc);
                        acc[pc].x = fmaf(pe, hv.x, acc[pc].x);
                        acc[pc].y = fmaf(pe, hv.y, acc[pc].y);
                        acc[pc].z = fmaf(pe, hv.z, acc[pc].z);
                        acc[pc].w = fmaf(pe, hv.w, acc[pc].w);
                    }
                }
                if (row < N) {
                    float* orow = out + (size_t)row * C + 4 * sub;
#pragma unroll
                    for (int pc = 0; pc < NP; ++pc)
                        *reinterpret_cast<float4*>(orow + 32 * pc) =
                            make_float4(acc[pc].x + bv[pc].x, acc[pc].y + bv[pc].y, acc[pc].z + bv[pc].z, acc[pc].w + bv[pc].w);
                    if (sub == 0) {
                        m_out[row] = mx;
                        z_out[row] = zs;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
        }
    }
}

// ---- host side: tensor map over h[N, C] (row-major fp32), box = one row, for tile::gather4
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static int g_gat_tma = -1;  // BG_GAT_TMA: 1 = use the TMA kernel where eligible (default set in bg_gat.cu's dispatch), 0 = never
int gat_tma_enabled() {
    if (g_gat_tma < 0) g_gat_tma = getenv("BG_GAT_TMA") ? atoi(getenv("BG_GAT_TMA")) : 0;
    return g_gat_tma;
}

template <int C>
static int launch_tma(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias, float* out, float* m,
                      float* z, float slope, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled();
    BG_REQUIRE(enc != nullptr, BG_EUNSUPPORTED, "bg_gat_fwd_tma: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)g->N};
    const cuuint64_t gstride[1] = {(cuuint64_t)C * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)C, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(h), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BG_REQUIRE(rc == CUDA_SUCCESS, BG_ECUDA, "bg_gat_fwd_tma: cuTensorMapEncodeTiled failed (%d)", (int)rc);
    constexpr size_t smem = (size_t)kTmaStages * kTmaRows * kTmaSlots * C * sizeof(float) + (size_t)kTmaStages * kTmaRows * kTmaSlots * sizeof(float) +
                            2 * kTmaStages * sizeof(uint64_t) + 256;
    static bool once = (cudaFuncSetAttribute(gat_fwd_tma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    const int nblocks = (int)ceil_div(g->N, kTmaRows);
    const int cap = kSMs * (C <= 64 ? 2 : 1);
    const unsigned grid = (unsigned)(nblocks < cap ? nblocks : cap);
    launch_k(gat_fwd_tma_kernel<C>, grid, kTmaThreads, smem, st, map, g->rowptr, g->col, s, d, bias, out, m, z, (int)g->N, (int)g->E, slope, nblocks);
    return check_launch("bg_gat_fwd_tma");
}

// BG_OK when launched, 1 when not eligible (C not 64 / 128, a row with more than 8 in-edges, misaligned h), < 0 on error.
int gat_fwd_tma_try(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias, float* out, float* m,
                    float* z, int C, float slope, cudaStream_t st) {
    if ((C != 64 && C != 128) || g->max_deg > kTmaSlots || g->max_deg <= 0) return 1;
    if ((reinterpret_cast<uintptr_t>(h) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return 1;
    return C == 64 ? launch_tma<64>(g, h, s, d, bias, out, m, z, slope, st) : launch_tma<128>(g, h, s, d, bias, out, m, z, slope, st);
}

}  // namespace bg

// The TMA-gather variant of bg_gat_fwd, callable on its own (benchmarks / tests): same arguments and results; returns
// BG_EUNSUPPORTED when the shape is not eligible (C other than 64 / 128, a row with more than 8 in-edges incl. the self loop).
extern "C" int bg_gat_fwd_tma(const BgGraph* g, const float* h, const float* s, const float* d, const float* bias, float* out,
                              float* m, float* z, int32_t C, float slope, void* stream) {
    BG_REQUIRE(g && g->rowptr && g->col && h && s && d && out && m && z, BG_EINVAL, "bg_gat_fwd_tma: null pointer");
    BG_REQUIRE(g->N > 0 && g->N < ((int64_t)1 << 31) - 64, BG_EINVAL, "bg_gat_fwd_tma: N out of range");
    const int rc = bg::gat_fwd_tma_try(g, h, s, d, bias, out, m, z, C, slope, bg::as_stream(stream));
    BG_REQUIRE(rc != 1, BG_EUNSUPPORTED, "bg_gat_fwd_tma: needs C in {64, 128}, max in-degree <= 8 (got C=%d, max_deg=%d) and 16-byte aligned h / out",
               (int)C, (int)g->max_deg);
    return rc;
}
extern "C" int bg_set_gat_tma(int32_t on) {
    const int prev = bg::gat_tma_enabled();
    bg::g_gat_tma = on ? 1 : 0;
    return prev;
}
