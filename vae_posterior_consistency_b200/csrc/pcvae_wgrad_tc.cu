// Weight gradients of the tcgen05 training path (autograd of nn.Linear inside train.py:115):
//   dWaug[m][n] = sum_r A[r][m] * B[r][n]      over all rows r of both branches, fp32-accurate 3xTF32,
// for up to three layers per launch.  A = pre-activation gradients, B = (layer input | 1); n == Kin is the constant-1
// column of B, i.e. the bias gradient.  The reduction runs over ROWS, so both operands are needed K-major over rows.
// The row-tile kernels write their scratch slab-blocked feature-major, [tile][row / 32][feature][row % 32] (a warp's
// 32 rows of one feature are one 128-byte store, and a 32-row slab of a buffer with F feature rows is ONE contiguous
// block of 32 F floats).
//
// Warp-specialised, mbarrier-pipelined, no CTA-wide barrier in the steady state:
//   warp 14      one lane streams slabs into a 4-stage ring of RAW blocks with two bulk async copies per slab
//                (cp.async.bulk, completion counted in bytes on the stage's `land` barrier);
//   warps 0..9   wait for `land`, read the raw block (conflict-free LDS.128: four rows of one feature), release the raw
//                stage, and write the MMA operand images of a 2-stage ring: the hi image in the K-major no-swizzle
//                core-matrix order [chunk of 4 rows][feature][4] and the lo image (x_lo = x - trunc_tf32(x)) beside it;
//                fence.proxy.async and arrive on the image stage's `full` barrier;
//   warp 15      waits for `full`, one elected lane issues per 8-row k-step
//                    D[128 x 2Nb] (+)= A_hi x [B_hi ; B_lo]^T      (B_hi and B_lo stacked along N: one read of A_hi)
//                    D[128 x  Nb]  += A_lo x  B_hi^T
//                and commits to the image stage's `empty` barrier;
//   warps 10..13 (one per TMEM lane quarter) drain a layer's accumulator when its last MMA has completed.
// The accumulator of a layer lives in TMEM for the whole layer (two buffers, so the next layer's MMAs run while this
// one is drained); the drain adds the two halves and writes this CTA's partial into the [grid][param_count] layout
// the FFMA kernels use, so pcvae_reduce_grads is unchanged and the result is deterministic.
#include "pcvae_tc.cuh"
#include "pcvae_train.cuh"

namespace pcvae {
namespace tc {

constexpr int SLAB = 32, WG_CH = SLAB / 4;                                    // 8 chunks of 4 rows per slab
constexpr int WG_RSTAGES = 4, WG_ISTAGES = 2;                                 // raw ring, image ring
constexpr int WG_FMAX = 104;                                                  // largest feature pitch of a scratch buffer
constexpr int WG_RAW_FLOATS = WG_FMAX * SLAB;                                 // one raw operand block [feature][32 rows]
constexpr int WG_RSTAGE_FLOATS = 2 * WG_RAW_FLOATS;                           // raw A | raw B
constexpr int WG_BROWS = 224;                                                 // B image rows per chunk: hi [0,Nb) + lo [Nb,2Nb)
constexpr int WG_ACS = (WG_FMAX + 1) * 4, WG_BCS = (WG_BROWS + 1) * 4;        // chunk strides in floats (+1 row: conflict-free writes)
constexpr int WG_A_FLOATS = WG_CH * WG_ACS, WG_B_FLOATS = WG_CH * WG_BCS;
constexpr int WG_ISTAGE_FLOATS = 2 * WG_A_FLOATS + WG_B_FLOATS;               // A_hi | A_lo | B
constexpr int WG_SMEM_FLOATS = WG_RSTAGES * WG_RSTAGE_FLOATS + WG_ISTAGES * WG_ISTAGE_FLOATS;
constexpr int WG_MMA_WARP = NWARP - 1, WG_COPY_WARP = NWARP - 2;
constexpr int WG_DRAIN_WARP0 = NWARP - 6;                                     // warps 10..13: one per TMEM lane quarter
constexpr int WG_CONVERTERS = WG_DRAIN_WARP0 * 32;                            // 320 threads
constexpr int WG_IPT = 3;                                                     // float4 items per converter thread, operand and slab
constexpr int WG_MAXJOBS = 3;
constexpr int WG_TROWS = 128, WG_SPT = WG_TROWS / SLAB;                        // rows per scratch tile, slabs per tile

struct WgradArgs {
    WgradJob job[WG_MAXJOBS];
    int njobs;
    long nvt;                             // 128-row tiles in the scratch (rows past the batch carry zero gradients)
    float* gp; long P;
    int* status;                          // tensor-core status word (tc_status_ptr)
};

__global__ void __launch_bounds__(NT, 1) k_wgrad_tc(const WgradArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t land_bar[WG_RSTAGES];
    __shared__ __align__(8) uint64_t rfree_bar[WG_RSTAGES];
    __shared__ __align__(8) uint64_t full_bar[WG_ISTAGES];
    __shared__ __align__(8) uint64_t empty_bar[WG_ISTAGES];
    __shared__ __align__(8) uint64_t done_bar[2];
    __shared__ __align__(8) uint64_t drained_bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* gp = a.gp + (long)blockIdx.x * a.P;
    const int mine_t = blockIdx.x < a.nvt ? (int)((a.nvt - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;    // tiles of this CTA
    const int nslab = mine_t * WG_SPT;                                                                       // slabs per layer
    if (nslab == 0) {                                   // no rows for this CTA: its partials are zero
        pdl_wait();
        for (int j = 0; j < a.njobs; ++j) {
            const WgradJob& J = a.job[j];
            for (int i = tid; i < J.Ma * (J.Kin + 1); i += NT) {
                const int m = i / (J.Kin + 1), n = i - m * (J.Kin + 1);
                if (n < J.Kin) gp[J.W_off + m * J.Kin + n] = 0.f; else gp[J.b_off + m] = 0.f;
            }
        }
        return;
    }
    for (int i = tid; i < WG_SMEM_FLOATS; i += NT) smem[i] = 0.f;               // image rows nobody writes start as zero
    if (tid == 0) {
        for (int s = 0; s < WG_RSTAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&land_bar[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&rfree_bar[s])), "r"(WG_CONVERTERS));
        }
        for (int s = 0; s < WG_ISTAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(WG_CONVERTERS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar[s])), "r"(1));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar[1])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&drained_bar[0])), "r"(128));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&drained_bar[1])), "r"(128));
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
    pdl_wait();                                           // both operands come from the kernels before
    const uint32_t tmem = tmem_slot;
    // Slab g of this CTA's sequence (layer after layer) uses raw stage g % 4 and image stage g % 2; the k-th use of a
    // stage completes the phase of parity k & 1 on each of its barriers.  All cursors advance incrementally.
    float* const raw0 = smem;
    float* const img0 = smem + WG_RSTAGES * WG_RSTAGE_FLOATS;

    if (warp == WG_COPY_WARP) {
        // ------------------------------ bulk-copy lane ------------------------------
        if (lane == 0) {
            int r = 0;
            uint32_t pr = 0;                             // parity of the current pass over the raw ring
            bool wrapped = false;
            // The gradient operand A was written (and B's producer ran) just before this launch, last tiles last: tiles
            // are taken in DESCENDING order so that the first reads find what is still in L2, and the activation
            // operand B (written a few kernels ago, read exactly once here) is marked evict-first so that it does not
            // push the not-yet-read gradient tiles out.
            const uint64_t pol_b = l2_policy_evict_first();
            for (int j = 0; j < a.njobs; ++j) {
                const WgradJob J = a.job[j];
                const uint32_t abytes = (uint32_t)(J.Fa * SLAB * 4), bbytes = (uint32_t)(min(J.Fb, J.Nb) * SLAB * 4);
                for (int t = mine_t - 1; t >= 0; --t) {
                    const long vt = blockIdx.x + (long)t * gridDim.x;
                    const float* At = J.AT + vt * (long)(J.Fa * WG_TROWS);
                    const float* Bt = J.BT + vt * (long)(J.Fb * WG_TROWS);
                    for (int sl = 0; sl < WG_SPT; ++sl) {
                        if (wrapped) mbar_wait(&rfree_bar[r], pr ^ 1u, a.status, 5);      // the converters have read the stage's previous slab
                        float* st = raw0 + r * WG_RSTAGE_FLOATS;
                        mbar_expect_tx(&land_bar[r], abytes + bbytes);
                        bulk_g2s(st, At + sl * (SLAB * J.Fa), abytes, &land_bar[r]);
                        bulk_g2s_hint(st + WG_RAW_FLOATS, Bt + sl * (SLAB * J.Fb), bbytes, &land_bar[r], pol_b);
                        if (++r == WG_RSTAGES) { r = 0; pr ^= 1u; wrapped = true; }
                    }
                }
            }
        }
    } else if (warp == WG_MMA_WARP) {
        // ------------------------------ MMA issuer ------------------------------
        int im = 0;
        uint32_t pi = 0;
        for (int j = 0; j < a.njobs; ++j) {
            const int Nb = a.job[j].Nb;
            const uint32_t acc = tmem + (uint32_t)((j & 1) * 256);
            // M = 64 for the layers with at most 64 outputs: the tensor core fetches shared-memory operands at ~64 B/cycle
            // and the A tile (M x 8 rows x 4 B per k-step, twice: hi and lo) is most of what a k-step reads for these
            // layers -- an M = 128 tile that is half padding costs the same fetch as a full one
            const int Mm = a.job[j].Ma <= 64 ? 64 : 128;
            const uint32_t id2 = make_idesc(Mm, 2 * Nb), id1 = make_idesc(Mm, Nb);
            constexpr uint64_t sa = (2 * WG_ACS * 4) >> 4, sb = (2 * WG_BCS * 4) >> 4;
            if (j >= 2) {                                // this accumulator buffer was layer j-2's: wait until it is drained
                mbar_wait(&drained_bar[j & 1], (uint32_t)(((j >> 1) - 1) & 1), a.status, 5);
                tc_fence_after();
            }
            for (int i = 0; i < nslab; ++i) {
                if (j == a.njobs - 1 && i == nslab - WG_SPT) pdl_trigger();   // last tile of the last layer: the next kernel may take the SM when this CTA exits
                mbar_wait(&full_bar[im], pi, a.status, 5);
                tc_fence_after();
                if (elect_one()) {
                    float* st = img0 + im * WG_ISTAGE_FLOATS;
                    const uint64_t dAh = make_desc(smem_u32(st), WG_ACS * 4, 128), dAl = make_desc(smem_u32(st + WG_A_FLOATS), WG_ACS * 4, 128);
                    const uint64_t dB = make_desc(smem_u32(st + 2 * WG_A_FLOATS), WG_BCS * 4, 128);
#pragma unroll
                    for (int ks = 0; ks < SLAB / 8; ++ks) {
                        mma_tf32_ss(acc, dAh + ks * sa, dB + ks * sb, id2, (i > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32_ss(acc, dAl + ks * sa, dB + ks * sb, id1, 1);
                    }
                    mma_commit(&empty_bar[im]);
                    if (i == nslab - 1) mma_commit(&done_bar[j & 1]);
                }
                __syncwarp();
                if (++im == WG_ISTAGES) { im = 0; pi ^= 1u; }
            }
        }
    } else if (warp >= WG_DRAIN_WARP0) {
        // ------------------------------ drain ------------------------------
        // Row m of dWaug = TMEM lane (this warp's quarter); columns [0,Nb) hold A_hi B_hi + A_lo B_hi, [Nb,2Nb) A_hi B_lo.
        for (int j = 0; j < a.njobs; ++j) {
            const WgradJob J = a.job[j];
            mbar_wait(&done_bar[j & 1], (uint32_t)((j >> 1) & 1), a.status, 5);
            tc_fence_after();
            const int q = warp & 3;
            // M = 128: accumulator row m is TMEM lane m; M = 64: row 16 i + r (r < 16) is lane 32 i + r (the first 16
            // lanes of every lane quarter)
            const bool m64 = J.Ma <= 64;
            const int m = m64 ? 16 * q + lane : 32 * q + lane;
            const bool mine = m64 ? (lane < 16 && m < J.Ma) : (m < J.Ma);
            const uint32_t acc = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)((j & 1) * 256);
            float* wrow = gp + J.W_off + m * J.Kin;
            for (int c = 0; c < J.Kin + 1; c += 4) {
                float x[4], w[4];
                tmem_ld4(acc + c, x);
                tmem_ld4(acc + J.Nb + c, w);
                if (mine) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int n = c + e;
                        if (n < J.Kin) wrow[n] = x[e] + w[e];
                        else if (n == J.Kin) gp[J.b_off + m] = x[e] + w[e];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&drained_bar[j & 1]);
        }
    } else {
        // ------------------------------ converters ------------------------------
        int r = 0, im = 0;
        uint32_t pr = 0, pi = 0;
        bool iwrapped = false;
        for (int j = 0; j < a.njobs; ++j) {
            const WgradJob J = a.job[j];
            const int Bn = min(J.Fb, J.Nb);              // B feature rows copied per slab
            const int loB = J.Nb * 4;
            // item idx = (feature f = idx / 8, chunk c = idx % 8): raw float offset 4 idx, image offset c * stride + 4 f
            int imgA[WG_IPT], imgB[WG_IPT];              // -1 = none
#pragma unroll
            for (int k = 0; k < WG_IPT; ++k) {
                const int idx = tid + k * WG_CONVERTERS, f = idx >> 3, c = idx & 7;
                imgA[k] = idx < WG_CH * J.Fa ? c * WG_ACS + f * 4 : -1;
                imgB[k] = idx < WG_CH * Bn ? 2 * WG_A_FLOATS + c * WG_BCS + f * 4 : -1;
            }
            for (int i = 0; i < nslab; ++i) {
                const float* raw = raw0 + r * WG_RSTAGE_FLOATS + tid * 4;
                float* st = img0 + im * WG_ISTAGE_FLOATS;
                mbar_wait(&land_bar[r], pr, a.status, 5);
                float4 va[WG_IPT], vb[WG_IPT];
#pragma unroll
                for (int k = 0; k < WG_IPT; ++k) {
                    if (imgA[k] >= 0) va[k] = *reinterpret_cast<const float4*>(raw + k * (WG_CONVERTERS * 4));
                    if (imgB[k] >= 0) vb[k] = *reinterpret_cast<const float4*>(raw + WG_RAW_FLOATS + k * (WG_CONVERTERS * 4));
                }
                if (iwrapped) mbar_wait(&empty_bar[im], pi ^ 1u, a.status, 5);            // the MMAs of the image stage's previous slab are done
#pragma unroll
                for (int k = 0; k < WG_IPT; ++k) {
                    if (imgA[k] >= 0) {
                        *reinterpret_cast<float4*>(st + imgA[k]) = va[k];
                        *reinterpret_cast<float4*>(st + imgA[k] + WG_A_FLOATS) =
                            make_float4(tf32_lo(va[k].x), tf32_lo(va[k].y), tf32_lo(va[k].z), tf32_lo(va[k].w));
                    }
                    if (imgB[k] >= 0) {
                        *reinterpret_cast<float4*>(st + imgB[k]) = vb[k];
                        *reinterpret_cast<float4*>(st + imgB[k] + loB) =
                            make_float4(tf32_lo(vb[k].x), tf32_lo(vb[k].y), tf32_lo(vb[k].z), tf32_lo(vb[k].w));
                    }
                }
                fence_async_smem();
                mbar_arrive(&full_bar[im]);
                mbar_arrive(&rfree_bar[r]);              // the raw block has been consumed: the copy lane may refill the stage
                if (++r == WG_RSTAGES) { r = 0; pr ^= 1u; }
                if (++im == WG_ISTAGES) { im = 0; pi ^= 1u; iwrapped = true; }
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
        const WgradJob& J = jobs[j];
        if (J.Fa > tc::WG_FMAX || J.Fb > tc::WG_FMAX ||  2 * J.Nb > tc::WG_BROWS || J.Nb % 16 ||
            J.Ma > J.Fa || J.Kin + 1 > J.Fb || J.Kin + 1 > J.Nb || tc::WG_CH * J.Fa > tc::WG_IPT * tc::WG_CONVERTERS)
            return fail(PCVAE_EINVAL, "wgrad_tc: layer %d shape (%d/%d x %d/%d, N %d) not supported", j, J.Ma, J.Fa, J.Kin, J.Fb, J.Nb);
    }
    a.njobs = njobs; a.nvt = nvt; a.gp = gp; a.P = P; a.status = tc_status_ptr();
    const size_t sm = (size_t)tc::WG_SMEM_FLOATS * sizeof(float) + 128;
    cudaError_t e = cudaFuncSetAttribute(tc::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    e = launch_tc(tc::k_wgrad_tc, grid, NT, sm, st, true, a);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "wgrad_tc: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // namespace pcvae
