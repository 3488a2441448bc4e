// Weight gradients of the tcgen05 training path (autograd of nn.Linear inside train.py:115):
//   dWaug[m][n] = sum_r AT[m][r] * BT[n][r]      over all rows r of both branches, fp32-accurate 3xTF32,
// for up to three layers per launch.  AT = pre-activation gradients, BT = (layer input | 1), both tile-blocked
// FEATURE-MAJOR in HBM ([tile][feature][128 rows], one contiguous block per 128-row tile of the row-tile kernels
// that wrote it), which is exactly the K-major operand layout of this GEMM (its reduction runs over rows) and
// lets a CTA stream whole tiles sequentially.  n == Kin is the constant-1 row of BT, i.e. the bias gradient.
//
// Warp-specialised, no CTA-wide barrier in the steady state:
//   warps 0..14  stream 32-row slabs with 16-byte cp.async straight into the MMA core-matrix image
//                [chunk of 4 rows][feature][4] of a 3-stage ring, then split the chunks THEY copied into the
//                lo image (x_lo = x - trunc_tf32(x)), fence.proxy.async and arrive on the stage's `full` barrier;
//   warp 15      waits for `full`, one elected lane issues per 8-row k-step
//                    D[128 x 2Nb] (+)= A_hi x [B_hi ; B_lo]^T      (B_hi and B_lo stacked along N: one read of A_hi)
//                    D[128 x  Nb]  += A_lo x  B_hi^T
//                and commits to the stage's `empty` barrier (the producers' licence to refill it).
// The accumulator of a layer lives in TMEM for the whole layer (two buffers, so the next layer's MMAs start while
// this one is drained); the drain adds the two halves and writes this CTA's partial into the [grid][param_count]
// layout the FFMA kernels use, so pcvae_reduce_grads is unchanged and the result is deterministic.
#include <cuda_pipeline.h>

#include "pcvae_tc.cuh"
#include "pcvae_train.cuh"

namespace pcvae {
namespace tc {

constexpr int SLAB = 32, WG_STAGES = 3, WG_CH = SLAB / 4;
constexpr int WG_AROWS = 128, WG_BROWS = 224;                                 // image rows: MMA M, MMA N (hi + lo)
constexpr int WG_ACS = (WG_AROWS + 1) * 4, WG_BCS = (WG_BROWS + 1) * 4;      // chunk strides in floats (+1 row: conflict-free copies)
constexpr int WG_A_FLOATS = WG_CH * WG_ACS, WG_B_FLOATS = WG_CH * WG_BCS;
constexpr int WG_STAGE_FLOATS = 2 * WG_A_FLOATS + WG_B_FLOATS;               // A_hi | A_lo | B (hi rows [0,Nb), lo rows [Nb,2Nb))
constexpr int WG_PRODUCERS = NT - 32;
constexpr int WG_MMA_WARP = NWARP - 1;
constexpr int WG_DRAIN_WARPS = 12;                                            // 3 column parts x 4 TMEM lane quarters
constexpr int WG_MAXJOBS = 3;
constexpr int WG_IPT = 4;                                                     // 16-byte copies per producer thread and slab (at most)

constexpr int WG_TROWS = 128, WG_SPT = WG_TROWS / SLAB;                        // rows per scratch tile, slabs per tile

struct WgradArgs {
    WgradJob job[WG_MAXJOBS];
    int njobs;
    long nvt;                             // 128-row tiles in the scratch (rows past the batch carry zero gradients)
    float* gp; long P;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(NT, 1) k_wgrad_tc(const WgradArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full_bar[WG_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[WG_STAGES];
    __shared__ __align__(8) uint64_t done_bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* gp = a.gp + (long)blockIdx.x * a.P;
    const long mine_t = blockIdx.x < a.nvt ? (a.nvt - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;    // tiles of this CTA
    const long mine = mine_t * WG_SPT;                                                                // slabs per layer
    if (mine == 0) {                                    // no rows for this CTA: its partials are zero
        for (int j = 0; j < a.njobs; ++j) {
            const WgradJob& J = a.job[j];
            for (int i = tid; i < J.Ma * (J.Kin + 1); i += NT) {
                const int m = i / (J.Kin + 1), n = i - m * (J.Kin + 1);
                if (n < J.Kin) gp[J.W_off + m * J.Kin + n] = 0.f; else gp[J.b_off + m] = 0.f;
            }
        }
        return;
    }
    for (int i = tid; i < WG_STAGES * WG_STAGE_FLOATS; i += NT) smem[i] = 0.f;   // rows the copies never touch start as zero
    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(WG_PRODUCERS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar[s])), "r"(1));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int nslab = (int)mine;                        // slabs per layer for this CTA; its sequence runs layer after layer
    // All cursors below advance incrementally (no divisions in the steady state): slab g of the sequence uses stage
    // g % 3, and the mbarrier phase of a stage's k-th use has parity k & 1.

    if (warp != WG_MMA_WARP) {
        // ------------------------------ producers ------------------------------
        int is_j = 0, is_i = 0, is_s = 0;               // next slab to copy: layer, slab in the layer, stage
        auto issue_next = [&]() {                       // cp.async the next slab of the sequence into its stage
            if (is_j < a.njobs) {
                const WgradJob& J = a.job[is_j];
                const long vt = blockIdx.x + (long)(is_i / WG_SPT) * gridDim.x;      // scratch tile, 32-row slab inside it
                const int r0 = (is_i % WG_SPT) * SLAB;
                const float* At = J.AT + vt * (long)(J.Fa * WG_TROWS) + r0;
                const float* Bt = J.BT + vt * (long)(J.Fb * WG_TROWS) + r0 - (long)J.Ma * WG_TROWS;
                float* st = smem + is_s * WG_STAGE_FLOATS;
                float* stB = st + 2 * WG_A_FLOATS - J.Ma * 4;
                const int items = WG_CH * (J.Ma + J.Kin + 1);
#pragma unroll
                for (int k = 0; k < WG_IPT; ++k) {
                    const int idx = tid + k * WG_PRODUCERS;
                    const int f = idx / WG_CH, c = idx % WG_CH;               // feature row, 4-row chunk
                    if (idx < items) {
                        if (f < J.Ma) __pipeline_memcpy_async(st + c * WG_ACS + f * 4, At + f * WG_TROWS + 4 * c, 16);
                        else __pipeline_memcpy_async(stB + c * WG_BCS + f * 4, Bt + f * WG_TROWS + 4 * c, 16);
                    }
                }
                if (++is_i == nslab) { is_i = 0; ++is_j; }
                is_s = is_s == WG_STAGES - 1 ? 0 : is_s + 1;
            }
            __pipeline_commit();
        };
        int pf_j = 0, pf_t = 0;                         // next tile to pull towards L2 (both operands, contiguous)
        auto prefetch_next = [&]() {
            if (pf_j < a.njobs) {
                const WgradJob& J = a.job[pf_j];
                const long vt = blockIdx.x + (long)pf_t * gridDim.x;
                const char* pa = reinterpret_cast<const char*>(J.AT + vt * (long)(J.Fa * WG_TROWS));
                const char* pb = reinterpret_cast<const char*>(J.BT + vt * (long)(J.Fb * WG_TROWS));
                const int la = J.Ma * (WG_TROWS * 4 / 128), lb = (J.Kin + 1) * (WG_TROWS * 4 / 128);    // 128-byte lines
                for (int l = tid; l < la + lb; l += WG_PRODUCERS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(l < la ? pa + (long)l * 128 : pb + (long)(l - la) * 128));
                if (++pf_t == (int)mine_t) { pf_t = 0; ++pf_j; }
            }
        };
        prefetch_next();                                // tile 0 (its first slabs are requested right away anyway)
        prefetch_next();                                // tile 1
        issue_next();
        issue_next();
        int s = 0;                                      // stage of the current slab
        uint32_t ph = 0;                                // parity of the current pass over the ring
        bool first = true;
        for (int j = 0; j < a.njobs; ++j) {
            const WgradJob& J = a.job[j];
            const int items = WG_CH * (J.Ma + J.Kin + 1);
            for (int i = 0; i < nslab; ++i) {
                __pipeline_wait_prior(1);               // this thread's copies of the current slab have landed
                float* st = smem + s * WG_STAGE_FLOATS;
                float* stB = st + 2 * WG_A_FLOATS - J.Ma * 4;
                float4 v[WG_IPT];
                float* hp[WG_IPT];
#pragma unroll
                for (int k = 0; k < WG_IPT; ++k) {      // lo images of the chunks this thread copied
                    const int idx = tid + k * WG_PRODUCERS;
                    const int f = idx / WG_CH, c = idx % WG_CH;
                    hp[k] = f < J.Ma ? st + c * WG_ACS + f * 4 : stB + c * WG_BCS + f * 4;
                    if (idx < items) v[k] = *reinterpret_cast<const float4*>(hp[k]);
                }
#pragma unroll
                for (int k = 0; k < WG_IPT; ++k) {
                    const int idx = tid + k * WG_PRODUCERS;
                    if (idx < items) {
                        float* lo = hp[k] + ((idx / WG_CH) < J.Ma ? WG_A_FLOATS : J.Nb * 4);
                        *reinterpret_cast<float4*>(lo) = make_float4(tf32_lo(v[k].x), tf32_lo(v[k].y), tf32_lo(v[k].z), tf32_lo(v[k].w));
                    }
                }
                fence_async_smem();
                mbar_arrive(&full_bar[s]);
                // refill the stage the PREVIOUS slab used (slab g + 2 goes there) once its MMAs have read it
                if (!first) {
                    const int sp = s == 0 ? WG_STAGES - 1 : s - 1;
                    if (is_j < a.njobs) mbar_wait(&empty_bar[sp], s == 0 ? ph ^ 1u : ph);
                }
                first = false;
                if ((i % WG_SPT) == 0) prefetch_next();
                issue_next();
                if (i == nslab - 1 && warp < WG_DRAIN_WARPS) {
                    // the layer's last slab is on its way: drain its accumulator once the MMAs have finished.
                    // Row m of dWaug = TMEM lane; the layer's columns are split over three warps per lane quarter.
                    mbar_wait(&done_bar[j & 1], (uint32_t)((j >> 1) & 1));
                    tc_fence_after();
                    const int q = warp & 3, part = warp >> 2;
                    const int m = 32 * q + lane;
                    const uint32_t acc = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)((j & 1) * 256);
                    const int nb4 = J.Nb / 4, per = (nb4 + 2) / 3;
                    for (int u = part * per; u < min(nb4, (part + 1) * per); ++u) {
                        float x[4], w[4];
                        tmem_ld4(acc + 4 * u, x);
                        tmem_ld4(acc + J.Nb + 4 * u, w);
                        if (m < J.Ma) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int n = 4 * u + e;
                                if (n < J.Kin) gp[J.W_off + m * J.Kin + n] = x[e] + w[e];
                                else if (n == J.Kin) gp[J.b_off + m] = x[e] + w[e];
                            }
                        }
                    }
                    tc_fence_before();
                }
                if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
            }
        }
        __pipeline_wait_prior(0);
    } else {
        // ------------------------------ MMA issuer ------------------------------
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < a.njobs; ++j) {
            const int Nb = a.job[j].Nb;
            const uint32_t acc = tmem + (uint32_t)((j & 1) * 256);
            const uint32_t id2 = make_idesc(128, 2 * Nb), id1 = make_idesc(128, Nb);
            for (int i = 0; i < nslab; ++i) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    float* st = smem + s * WG_STAGE_FLOATS;
                    const uint64_t dAh = make_desc(smem_u32(st), WG_ACS * 4, 128), dAl = make_desc(smem_u32(st + WG_A_FLOATS), WG_ACS * 4, 128);
                    const uint64_t dB = make_desc(smem_u32(st + 2 * WG_A_FLOATS), WG_BCS * 4, 128);
                    constexpr uint64_t sa = (2 * WG_ACS * 4) >> 4, sb = (2 * WG_BCS * 4) >> 4;
#pragma unroll
                    for (int ks = 0; ks < SLAB / 8; ++ks) {
                        mma_tf32_ss(acc, dAh + ks * sa, dB + ks * sb, id2, (i > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32_ss(acc, dAl + ks * sa, dB + ks * sb, id1, 1);
                    }
                    mma_commit(&empty_bar[s]);
                    if (i == nslab - 1) mma_commit(&done_bar[j & 1]);
                }
                __syncwarp();
                if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

}  // namespace tc

int wgrad_tc_launch(const WgradJob* jobs, int njobs, long nvt, float* gp, long P, int grid, cudaStream_t st) {
    if (njobs < 1 || njobs > tc::WG_MAXJOBS) return fail(PCVAE_EINVAL, "wgrad_tc: %d layers in one launch (1..%d)", njobs, tc::WG_MAXJOBS);
    tc::WgradArgs a{};
    for (int j = 0; j < njobs; ++j) {
        a.job[j] = jobs[j];
        if (jobs[j].Ma > tc::WG_AROWS || 2 * jobs[j].Nb > tc::WG_BROWS || jobs[j].Nb % 16 || jobs[j].Kin + 1 > jobs[j].Nb ||
            jobs[j].Ma > jobs[j].Fa || jobs[j].Kin + 1 > jobs[j].Fb)
            return fail(PCVAE_EINVAL, "wgrad_tc: layer %d shape (%d x %d, N %d) not supported", j, jobs[j].Ma, jobs[j].Kin, jobs[j].Nb);
        if (tc::WG_CH * (jobs[j].Ma + jobs[j].Kin + 1) > tc::WG_IPT * tc::WG_PRODUCERS)
            return fail(PCVAE_EINVAL, "wgrad_tc: layer %d has too many feature rows (%d)", j, jobs[j].Ma + jobs[j].Kin + 1);
    }
    a.njobs = njobs; a.nvt = nvt; a.gp = gp; a.P = P;
    const size_t sm = (size_t)tc::WG_STAGES * tc::WG_STAGE_FLOATS * sizeof(float) + 128;
    cudaError_t e = cudaFuncSetAttribute(tc::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    tc::k_wgrad_tc<<<grid, NT, sm, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "wgrad_tc: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // namespace pcvae
