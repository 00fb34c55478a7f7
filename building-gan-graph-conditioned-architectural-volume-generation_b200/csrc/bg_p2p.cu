// Data-parallel gradient exchange fused with the optimiser step: ONE kernel per model update that (1) waits until every
// rank's flat gradient bucket is complete, (2) reads all W buckets straight out of the peers' memory over NVLink / NVSwitch
// (one-shot all-reduce: every rank sums the same W values in the same rank order, so all ranks compute bit-identical
// averages), (3) applies Adam to the local flat parameter / moment buffers, and (4) tells the peers it is done reading.
//
// Why not ncclAllReduce + a scale kernel + bg_adam_flat.  The payload is tiny (discriminator 63 KB, generator 1.1 MB;
// reference train.py:36-37 / trainer.py:481,495 - SURVEY section 8e) and the step issues 6 of them: round 1 measured
// ~200 us per collective on the step's critical path (inter-rank skew at the barrier + NCCL's launch / protocol latency +
// a separate scale launch), 10-20x what moving 63 KB over NVLink costs.  Here the exchange is loads from peer memory inside
// the kernel that needs the result; the only synchronisation is two flag rounds (ready / done) over the same links.
//
// Memory model.  Buckets and flags live in symmetric memory (host side: torch.distributed._symmetric_memory; this file
// only sees raw peer pointers).  flags of rank q: uint32 [2 W]: slot p = "rank p's bucket is complete, epoch e", slot
// W + p = "rank p has finished reading, epoch e".  Epochs only grow (one launch = +2), compared as signed differences, so
// the flags never need resetting; the epoch lives in device memory and is advanced by the kernel itself => the launch is
// CUDA-graph capturable (no host-side state).  Signals are st.release.sys / ld.acquire.sys; peer gradient loads are
// ld.relaxed.sys (never served from a stale local L1 line).
#include <algorithm>
#include <math.h>

#include "bg_common.cuh"

namespace bg {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// Bounded spin: a peer that never makes the matching call (a rank that died, a call sequence that differs between ranks)
// must end in a loud launch failure, not in a silent hang of every rank.  ~2e7 polls of a system-scope load: tens of seconds.
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t want) {
    for (unsigned int polls = 0; (int32_t)(ld_acquire_sys(flag) - want) < 0; ++polls)
        if (polls > 20000000u) {
            printf("bg_p2p_allreduce_adam: timed out waiting for a peer (flag %p = %u, want %u): ranks out of step\n", (const void*)flag,
                   ld_acquire_sys(flag), want);
            __trap();
        }
}

struct AdamConsts {
    double lr, b1, b2d;
    float b2, w1, w2, eps, wd, step_size, bc2_sqrt;
};

__global__ void __launch_bounds__(kThreads) p2p_allreduce_adam_kernel(const BgPeers P, uint32_t* __restrict__ epoch,
                                                                      unsigned int* __restrict__ ticket, float4* __restrict__ p,
                                                                      float4* __restrict__ m, float4* __restrict__ v,
                                                                      float4* __restrict__ gavg, int64_t n4, AdamConsts A,
                                                                      const int64_t* __restrict__ step_dev) {
    pdl_prologue();
    const int W = P.world, me = P.rank;
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch);
    uint32_t* my_flags = P.flags[me];
    // ---- ready round: my bucket was written by kernels that completed before this one started (stream order)
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(P.flags[threadIdx.x] + me, e + 1);
    }
    if (threadIdx.x < W) wait_flag(my_flags + threadIdx.x, e + 1);
    __syncthreads();
    if (step_dev) {
        const double t = (double)*step_dev;
        A.step_size = (float)(A.lr / (1.0 - pow(A.b1, t)));
        A.bc2_sqrt = (float)sqrt(1.0 - pow(A.b2d, t));
    }
    const float inv_w = 1.f / (float)W;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
        float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < W; ++r) {  // rank order: the same sum, bit for bit, on every rank
            const float4 t = ld_peer4(reinterpret_cast<const float4*>(P.grad[r]) + i);
            gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
        }
        gg.x *= inv_w; gg.y *= inv_w; gg.z *= inv_w; gg.w *= inv_w;
        if (gavg) gavg[i] = gg;
        if (p) {
            float4 pp = p[i], mm = m[i], vv = v[i];
            float* Pp = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // identical to adam_flat_kernel (bg_dense.cu)
                const float gr = A.wd != 0.f ? fmaf(A.wd, Pp[k], G[k]) : G[k];
                const float diff = gr - M[k];
                M[k] = A.w1 < 0.5f ? M[k] + A.w1 * diff : gr - diff * (1.f - A.w1);
                V[k] = V[k] * A.b2 + A.w2 * gr * gr;
                const float denom = sqrtf(V[k]) / A.bc2_sqrt + A.eps;
                Pp[k] = Pp[k] - A.step_size * (M[k] / denom);
            }
            p[i] = pp; m[i] = mm; v[i] = vv;
        }
    }
    // ---- done round: the last CTA of this rank tells every peer that this rank has finished reading, waits for the same from
    // every peer (nobody may overwrite a bucket another rank is still reading) and advances the epoch
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
        if (last) *ticket = 0u;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < W) {
        st_release_sys(P.flags[threadIdx.x] + W + me, e + 2);
        wait_flag(my_flags + W + threadIdx.x, e + 2);
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch = e + 2;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_p2p_allreduce_adam(const BgPeers* peers, uint32_t* epoch, uint32_t* ticket, float* p, float* m, float* v,
                                     float* gavg, int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay,
                                     int64_t step, const int64_t* step_dev, void* stream) {
    BG_REQUIRE(peers && epoch && ticket, BG_EINVAL, "bg_p2p_allreduce_adam: null pointer");
    BG_REQUIRE(peers->world >= 1 && peers->world <= BG_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world, BG_EINVAL,
               "bg_p2p_allreduce_adam: world %d / rank %d out of range (max %d ranks)", peers->world, peers->rank, BG_MAX_PEERS);
    for (int r = 0; r < peers->world; ++r)
        BG_REQUIRE(peers->grad[r] && peers->flags[r] && ((uintptr_t)peers->grad[r] & 15) == 0, BG_EINVAL,
                   "bg_p2p_allreduce_adam: peer %d has a null / misaligned bucket or no flags", r);
    BG_REQUIRE((p && m && v) || (!p && !m && !v && gavg), BG_EINVAL,
               "bg_p2p_allreduce_adam: pass p, m, v together (all-reduce + Adam) or none of them with gavg (all-reduce only)");
    BG_REQUIRE(n >= 0 && n % 4 == 0, BG_EINVAL, "bg_p2p_allreduce_adam: n=%lld must be a multiple of 4 (flat buckets are padded)", (long long)n);
    BG_REQUIRE((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)gavg) & 15) == 0, BG_EINVAL, "bg_p2p_allreduce_adam: buffers must be 16-byte aligned");
    BG_REQUIRE(!p || step_dev || step >= 1, BG_EINVAL, "bg_p2p_allreduce_adam: step=%lld (counts from 1)", (long long)step);
    AdamConsts A{lr, beta1, beta2, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay, 0.f, 1.f};
    if (p && !step_dev) {
        A.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
        A.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    }
    // few, fat CTAs: the payload is <= 1.1 MB and every CTA polls the ready flags
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(ceil_div(n / 4, (int64_t)kThreads * 4), kSMs));
    launch_k(p2p_allreduce_adam_kernel, (int)grid, kThreads, 0, as_stream(stream), *peers, epoch, reinterpret_cast<unsigned int*>(ticket),
             reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), reinterpret_cast<float4*>(gavg), n / 4, A,
             step_dev);
    return check_launch("bg_p2p_allreduce_adam");
}
