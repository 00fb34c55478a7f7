// Dense layers on the 5th-generation tensor cores (tcgen05 + TMEM): fp32-accurate via the 3xTF32 split (default), or
// bf16 operands with fp32 accumulation (BG_DENSE_TC=bf16 / bg_set_dense_tc(2): the stated reduced-precision mode).
//
// y[128 rows, BN] = epilogue(X[128, K] W[BN, K]^T): the 128/64-wide Linear(+LayerNorm+LeakyReLU/ReLU)(+attention dots)
// layers of the generator / discriminator (reference models.py:49-66,92-113 and the conv `lin`s).
//
// TF32 keeps 10 mantissa bits, a single TF32 product is ~1e-3 accurate - not enough for the rel-1e-5 parity mode.  Every
// fp32 operand is therefore split into hi = a with its low 13 mantissa bits cleared (exactly a TF32 number) and
// lo = a - hi (exact in fp32, <= 13 significant bits), and the accumulator receives hi*hi + lo*hi + hi*lo (the dropped
// lo*lo term is ~2^-22 relative).  Operands are staged in shared memory K-major in the canonical SWIZZLE_128B layout
// (one 128-byte row = 32 fp32 = 4 UMMA k-steps), accumulators live in TMEM (128 lanes x BN columns fp32), the epilogue
// reads them back with tcgen05.ld, one thread per row (bias, LayerNorm, activation, attention dots, store).
//
// Accumulation accuracy (3xTF32 mode).  The tensor core adds every k-step into the fp32 TMEM accumulator with truncation
// (round toward zero): one chain over K = 128..524 is 16..66 one-sided roundings on the hi*hi products alone, and the bias
// does not average out (measured round 1: logits 5e-5 of max through 33 layers against 2.4e-5 for fp32 FFMA arithmetic).  The
// kernel therefore keeps FOUR accumulators in TMEM (4 x BN columns): the hi*hi products go round-robin by k-step into three
// of them (a third of the roundings each, on a third of the magnitude), all lo*hi / hi*lo correction products into the
// fourth (they are 2^-11 of the result: their roundings vanish), and the epilogue adds the four in fp32 round-to-nearest.
//
// bf16 mode: operands rounded to bf16 (RN) in the stage loader, ONE kind::f16 MMA per 16-wide k-step, same SWIZZLE_128B
// staging (a 128-byte row then holds 64 k), single accumulator.  Tolerance stated in tests/test_models_gpu.py.
//
// Pipeline per CTA: 2 shared-memory stages.  All 256 threads load / split / store the next stage while the tensor core
// consumes the previous one; one thread issues the 12 MMAs of a stage and a tcgen05.commit that frees it.
#include <stdlib.h>
#include <string.h>

#include "bg_common.cuh"

namespace bg {

constexpr int TC_M = 128;        // rows per CTA = UMMA M
constexpr int TC_STAGES = 2;

struct SegViewTc {
    int nseg;
    int off[BG_MAX_SEG + 1];
    BgSeg seg[BG_MAX_SEG];
};
struct SegColTc {
    const float* base;
    const int32_t* gather;
    int ld;
    bool ones;
};
__device__ __forceinline__ SegColTc seg_resolve_tc(const SegViewTc& sv, int k, int K) {
    SegColTc c{nullptr, nullptr, 0, false};
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
__device__ __forceinline__ float seg_load_tc(const SegColTc& c, int64_t row) {
    if (c.base == nullptr) return c.ones ? 1.f : 0.f;
    const int64_t r = c.gather ? (int64_t)__ldg(c.gather + row) : row;
    return __ldg(c.base + r * c.ld);
}

struct DenseTcParams {
    int64_t N;
    SegViewTc x;
    int K;
    const float* W;  // [BN, K] row-major
    const float *bias, *gamma, *beta, *att_src, *att_dst;
    int act;
    float* out;
    int64_t ld_out;
    float *xhat, *rstd, *s, *d;
};

// ---- raw PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for one swizzle atom along K)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
    return d;
}
// byte offset of element (row, k) inside a [rows][32 fp32] SWIZZLE_128B tile (Swizzle<3,4,3>: 16-byte chunk ^= row & 7)
__device__ __forceinline__ int swz(int row, int k) { return row * 128 + ((((k >> 2) ^ (row & 7)) << 4) | ((k & 3) << 2)); }

#define BG_TMEM_LD32(v, addr)                                                                                                      \
    asm volatile(                                                                                                                 \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22," \
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                                             \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),   \
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),     \
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),     \
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                                       \
        : "r"(addr))

// 32 accumulator columns [c0, c0+32) of this thread's row: the NACC partial accumulators (BN columns apart) summed in fp32
// round-to-nearest, in a fixed order
template <int NACC, int BN>
__device__ __forceinline__ void tmem_row32(uint32_t taddr, int c0, float (&out)[32]) {
    uint32_t v[32];
    BG_TMEM_LD32(v, taddr + c0);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = __uint_as_float(v[j]);
#pragma unroll
    for (int a = 1; a < NACC; ++a) {
        BG_TMEM_LD32(v, taddr + a * BN + c0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) out[j] += __uint_as_float(v[j]);
    }
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // d = {hi -> upper half, lo -> lower half}
    return r;
}

// MODE 0: 3xTF32 (fp32-accurate), MODE 1: bf16 operands / fp32 accumulate
template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1) dense_tc_kernel(const DenseTcParams p) {
    pdl_prologue();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int SK = MODE ? 64 : 32;          // k per shared-memory stage (one 128-byte swizzle row of the operand type)
    constexpr int NV = MODE ? 2 : 1;            // float4 loads per (thread, row): 16 bytes of the staged operand type
    constexpr int NACC = MODE ? 1 : 4;          // TMEM accumulators (see the header comment)
    constexpr int NSPLIT = MODE ? 1 : 2;        // hi / lo copies of every operand tile
    constexpr int A_BYTES = TC_M * 128, B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = NSPLIT * (A_BYTES + B_BYTES);
    constexpr int TMEM_COLS = NACC * BN;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // the epilogue re-uses the operand stages as its staging area (parameters, row partials, padded output tile): the mbarriers
    // sit behind whichever of the two is larger
    constexpr int EPI_BYTES = (5 * BN + 4 * TC_M + TC_M * (BN + 1)) * 4;
    constexpr int BAR_OFF = ((TC_STAGES * STAGE_BYTES > EPI_BYTES ? TC_STAGES * STAGE_BYTES : EPI_BYTES) + 15) & ~15;
    uint64_t* empty_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [TC_STAGES]
    uint64_t* accum_bar = empty_bar + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TC_M;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(empty_bar + s, 1);
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {  // TMEM: NACC x BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // instruction descriptor: D=F32, A=B=TF32 (kind::tf32) or BF16 (kind::f16), both K-major, N=BN, M=128
    constexpr uint32_t FMT = MODE ? 1u : 2u;
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    const int nkb = (p.K + SK - 1) / SK;
    // Loader mapping: a thread owns one 16-byte chunk of the staged 128-byte row (4 consecutive k in fp32 / tf32, 8 in bf16),
    // rows tid/8 + 32 i: 128-bit global loads (coalesced: 8 threads = one row), 128-bit swizzled shared stores (8 distinct
    // chunks per row: conflict-free).  Needs every segment width / row stride / pointer to be a multiple of 4 floats
    // (checked on the host), so a float4 never straddles two segments.
    const int chunk = tid & 7;
    const int rbase = tid >> 3;  // 0..31
    struct Regs {
        float4 a[TC_M / 32][NV], b[BN / 32][NV];
    };
    auto fetch = [&](int kb, Regs& R) {  // global -> registers for k-block kb, issued TWO iterations ahead of its use
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int kg = kb * SK + chunk * (4 * NV) + 4 * v;
            const SegColTc xc = seg_resolve_tc(p.x, kg, p.K);
#pragma unroll
            for (int i = 0; i < TC_M / 32; ++i) {
                const int64_t r = row0 + rbase + 32 * i;
                R.a[i][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < p.N && xc.base) {
                    const int64_t rr = xc.gather ? (int64_t)__ldg(xc.gather + r) : r;
                    R.a[i][v] = __ldg(reinterpret_cast<const float4*>(xc.base + rr * xc.ld));
                }
            }
#pragma unroll
            for (int i = 0; i < BN / 32; ++i)
                R.b[i][v] = kg < p.K ? __ldg(reinterpret_cast<const float4*>(p.W + (int64_t)(rbase + 32 * i) * p.K + kg))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto stage_store = [&](const float4 (&v)[NV], uint8_t* tile, int tile_bytes, int row) {
        const int o = row * 128 + ((chunk ^ (row & 7)) << 4);
        if constexpr (MODE == 0) {
            const float in[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
            float hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // round-to-nearest TF32 split: hi + lo == a up to 2^-22 |a|, both unbiased
                uint32_t hb, lb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[u]));
                hi[u] = __uint_as_float(hb);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(in[u] - hi[u]));
                lo[u] = __uint_as_float(lb);
            }
            *reinterpret_cast<float4*>(tile + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(tile + tile_bytes + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        } else {
            uint4 q;
            q.x = pack_bf16(v[0].x, v[0].y);
            q.y = pack_bf16(v[0].z, v[0].w);
            q.z = pack_bf16(v[NV - 1].x, v[NV - 1].y);
            q.w = pack_bf16(v[NV - 1].z, v[NV - 1].w);
            *reinterpret_cast<uint4*>(tile + o) = q;
        }
    };
    auto process = [&](int kb, Regs& R) {
        const int s = kb % TC_STAGES;
        uint8_t* st = smem + s * STAGE_BYTES;
        if (kb >= TC_STAGES) mbar_wait(empty_bar + s, ((kb / TC_STAGES) - 1) & 1);  // MMAs that read this stage retired
#pragma unroll
        for (int i = 0; i < TC_M / 32; ++i) stage_store(R.a[i], st, A_BYTES, rbase + 32 * i);
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) stage_store(R.b[i], st + NSPLIT * A_BYTES, B_BYTES, rbase + 32 * i);
        if (kb + 2 < nkb) fetch(kb + 2, R);  // refill this register set: in flight for two full iterations
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_hi = smem_u32(st), b_hi = a_hi + NSPLIT * A_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // one UMMA k-step = 32 bytes of the K-major row (8 tf32 / 16 bf16)
                const uint32_t adv = ks * 32;
                if constexpr (MODE == 0) {
                    const uint32_t a_lo = a_hi + A_BYTES, b_lo = b_hi + B_BYTES;
                    const int g = kb * 4 + ks;  // global k-step: hi*hi round-robin over accumulators 0..2, corrections into 3
                    tc_mma_tf32(tmem_base + (uint32_t)((g % 3) * BN), make_desc(a_hi + adv), make_desc(b_hi + adv), IDESC, g >= 3);
                    tc_mma_tf32(tmem_base + 3u * BN, make_desc(a_lo + adv), make_desc(b_hi + adv), IDESC, g != 0);
                    tc_mma_tf32(tmem_base + 3u * BN, make_desc(a_hi + adv), make_desc(b_lo + adv), IDESC, 1);
                } else {
                    tc_mma_bf16(tmem_base, make_desc(a_hi + adv), make_desc(b_hi + adv), IDESC, (kb | ks) != 0);
                }
            }
            tc_commit(empty_bar + s);               // frees the stage when these MMAs have read it
            if (kb == nkb - 1) tc_commit(accum_bar);  // accumulators complete
        }
    };
    Regs R0, R1;
    fetch(0, R0);
    if (nkb > 1) fetch(1, R1);
    for (int kb = 0; kb < nkb; kb += 2) {
        process(kb, R0);
        if (kb + 1 < nkb) process(kb + 1, R1);
    }

    // ---- epilogue: all 8 warps.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (the hardware's lane quarter of a warp) = rows, and the
    // column half w / 4: each thread pulls its BN / 2 accumulator columns out of TMEM ONCE (summing the NACC partial
    // accumulators), the LayerNorm row statistics and the attention dots are completed across the two halves through shared
    // memory, and results leave through a padded shared-memory tile so that global stores are coalesced (a warp writes 128
    // consecutive bytes of one row).  Round 1/2a: warps 0..3 only, three TMEM sweeps per row (mean, variance, output) and
    // 16-byte stores 512 bytes apart - the epilogue was ~20 of the kernel's ~35 us at N = 15 k (profiles/r02b_summary.md).
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    {
        constexpr int HB = BN / 2, TLD = BN + 1;
        float* sbias = reinterpret_cast<float*>(smem);          // operand stages are free: every MMA has completed
        float* sgam = sbias + BN;
        float* sbet = sgam + BN;
        float* sas = sbet + BN;
        float* sad = sas + BN;
        float* red = sad + BN;                                    // [2][2][TC_M]
        float* tile = red + 4 * TC_M;                             // [TC_M][TLD]
        for (int c = tid; c < BN; c += kThreads) {
            sbias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
            sgam[c] = p.gamma ? __ldg(p.gamma + c) : 1.f;
            sbet[c] = p.gamma ? __ldg(p.beta + c) : 0.f;
            sas[c] = p.att_src ? __ldg(p.att_src + c) : 0.f;
            sad[c] = p.att_src ? __ldg(p.att_dst + c) : 0.f;
        }
        const int q4 = warp & 3, half = warp >> 2;
        const int r_local = q4 * 32 + lane, c0 = half * HB;
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        float v[HB];
#pragma unroll
        for (int ch = 0; ch < HB / 32; ++ch) {
            float t32[32];
            tmem_row32<NACC, BN>(taddr, c0 + 32 * ch, t32);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[32 * ch + j] = t32[j];
        }
        __syncthreads();  // parameters staged
#pragma unroll
        for (int j = 0; j < HB; ++j) v[j] += sbias[c0 + j];
        // coalesced copy of the staged tile to a row-major global tensor
        auto flush_tile = [&](float* dst, int64_t ld) {
            __syncthreads();
            for (int idx = tid; idx < TC_M * BN; idx += kThreads) {
                const int r = idx / BN, c = idx % BN;
                if (row0 + r < p.N) dst[(row0 + r) * ld + c] = tile[r * TLD + c];
            }
            __syncthreads();
        };
        if (p.gamma) {
            float sm = 0.f;
#pragma unroll
            for (int j = 0; j < HB; ++j) sm += v[j];
            red[half * TC_M + r_local] = sm;
            __syncthreads();
            const float mean = (red[r_local] + red[TC_M + r_local]) / (float)BN;
            float vs = 0.f;
#pragma unroll
            for (int j = 0; j < HB; ++j) vs = fmaf(v[j] - mean, v[j] - mean, vs);
            red[(2 + half) * TC_M + r_local] = vs;
            __syncthreads();
            const float rs = 1.f / sqrtf((red[2 * TC_M + r_local] + red[3 * TC_M + r_local]) / (float)BN + 1e-5f);
            if (p.rstd && half == 0 && row0 + r_local < p.N) p.rstd[row0 + r_local] = rs;
#pragma unroll
            for (int j = 0; j < HB; ++j) v[j] = (v[j] - mean) * rs;
            if (p.xhat) {  // normalised value saved for the backward
#pragma unroll
                for (int j = 0; j < HB; ++j) tile[r_local * TLD + c0 + j] = v[j];
                flush_tile(p.xhat, BN);
            }
#pragma unroll
            for (int j = 0; j < HB; ++j) v[j] = fmaf(v[j], sgam[c0 + j], sbet[c0 + j]);
        }
        if (p.act == BG_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < HB; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
        } else if (p.act == BG_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < HB; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
        }
        if (p.att_src) {
            float ss = 0.f, dd = 0.f;
#pragma unroll
            for (int j = 0; j < HB; ++j) {
                ss = fmaf(v[j], sas[c0 + j], ss);
                dd = fmaf(v[j], sad[c0 + j], dd);
            }
            __syncthreads();  // red is reused
            red[half * TC_M + r_local] = ss;
            red[(2 + half) * TC_M + r_local] = dd;
            __syncthreads();
            if (half == 0 && row0 + r_local < p.N) {
                p.s[row0 + r_local] = red[r_local] + red[TC_M + r_local];
                p.d[row0 + r_local] = red[2 * TC_M + r_local] + red[3 * TC_M + r_local];
            }
        }
#pragma unroll
        for (int j = 0; j < HB; ++j) tile[r_local * TLD + c0 + j] = v[j];
        flush_tile(p.out, p.ld_out);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

// 0: off (FFMA kernels only), 1: 3xTF32 (default), 2: bf16 operands.  -1: read BG_DENSE_TC on first use.
static int g_dense_tc = -1;
static int dense_tc_mode() {
    if (g_dense_tc < 0) {
        const char* e = getenv("BG_DENSE_TC");
        g_dense_tc = !e ? 1 : (!strcmp(e, "bf16") || !strcmp(e, "2")) ? 2 : (atoi(e) ? 1 : 0);
    }
    return g_dense_tc;
}

template <int BN, int MODE>
static void launch_dense_tc(const DenseTcParams& p, unsigned grid, cudaStream_t st) {
    constexpr int stage_bytes = TC_STAGES * (MODE ? 1 : 2) * (TC_M * 128 + BN * 128);
    constexpr int epi_bytes = (5 * BN + 4 * TC_M + TC_M * (BN + 1)) * 4;  // epilogue: parameters, row partials, padded output tile
    constexpr int smem = (stage_bytes > epi_bytes ? stage_bytes : epi_bytes) + 16 + 1024 + 64;
    static bool once = (cudaFuncSetAttribute(dense_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), true);
    (void)once;
    launch_k(dense_tc_kernel<BN, MODE>, grid, kThreads, smem, st, p);
}

// Returns BG_OK if the layer was launched on the tensor-core path, 1 if the shape is not eligible (caller falls back to
// the FFMA kernel), negative on error.
int dense_tc_try(const BgDense* a, int K, cudaStream_t st) {
    const int mode = dense_tc_mode();
    if (!mode) return 1;
    if (a->w_sk != 1 || a->w_so != K) return 1;                 // plain [Cout, K] row-major weights only
    // measured crossover against the FFMA kernel: 128-wide layers from K=128, 64-wide layers from K=256
    if (!((a->Cout == 128 && K >= 128) || (a->Cout == 64 && K >= 256)) || a->N < 128) return 1;
    if (K % 4) return 1;
    for (int q = 0; q < a->nseg; ++q) {  // 128-bit loads: widths, row strides and pointers in units of 4 floats; no ones-segment
        const BgSeg& sg = a->seg[q];
        if (!sg.ptr || (sg.width & 3) || (sg.ld & 3) || (reinterpret_cast<uintptr_t>(sg.ptr) & 15)) return 1;
    }
    if (reinterpret_cast<uintptr_t>(a->W) & 15) return 1;
    if ((a->ld_out & 3) || (reinterpret_cast<uintptr_t>(a->out) & 15)) return 1;
    DenseTcParams p;
    p.N = a->N;
    p.x.nseg = a->nseg;
    p.x.off[0] = 0;
    for (int q = 0; q < BG_MAX_SEG; ++q) {
        p.x.seg[q] = q < a->nseg ? a->seg[q] : BgSeg{nullptr, nullptr, 0, 0};
        p.x.off[q + 1] = p.x.off[q] + (q < a->nseg ? a->seg[q].width : 0);
    }
    p.K = K;
    p.W = a->W; p.bias = a->bias; p.gamma = a->ln_gamma; p.beta = a->ln_beta;
    p.att_src = a->att_src; p.att_dst = a->att_dst; p.act = a->act;
    p.out = a->out; p.ld_out = a->ld_out; p.xhat = a->xhat; p.rstd = a->rstd; p.s = a->s; p.d = a->d;
    const unsigned grid = (unsigned)ceil_div(a->N, TC_M);
    if (a->Cout == 128) {
        if (mode == 2) launch_dense_tc<128, 1>(p, grid, st);
        else launch_dense_tc<128, 0>(p, grid, st);
    } else {
        if (mode == 2) launch_dense_tc<64, 1>(p, grid, st);
        else launch_dense_tc<64, 0>(p, grid, st);
    }
    return check_launch("bg_dense_fwd(tcgen05)");
}

}  // namespace bg

// 0 = FFMA kernels only, 1 = tcgen05 3xTF32 (fp32-accurate, default), 2 = tcgen05 bf16 operands (reduced precision).
// Returns the previous mode.
extern "C" int bg_set_dense_tc(int32_t mode) {
    const int prev = bg::dense_tc_mode();
    bg::g_dense_tc = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
    return prev;
}
