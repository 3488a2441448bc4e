// tcgen05 (5th-gen tensor core) version of the reward main kernel for the MLP family.
//
// Same mathematics and the same tile / sample-loop structure as k_reward_main<MLP> in
// pcvae_reward.cu (64 (row,candidate) pairs x {without, with target} = 128 tail evaluations per
// tile, per-pair accumulator summed in the reference's order), but the two dense layers of the
// tail run on the tensor cores:
//
//   layer 2:  D2[128 x 64]  = A2[128 x 104] * B2[64 x 104]^T     (100 inputs + bias column, 50 -> 64 outputs)
//   layer 3:  D3[128 x 32]  = A3[128 x 56]  * B3[32 x 56]^T      ( 50 inputs + bias column, 20 -> 32 outputs)
//
// with `tcgen05.mma.cta_group::1.kind::tf32` issued by one thread.  The layer-2 operand A2 lives in
// TMEM, double-buffered (row = TMEM lane, feature = TMEM column; written with `tcgen05.st`, read by
// the MMA in its A-from-TMEM form), accumulators are in TMEM (`tcgen05.ld`), the layer-3 operand
// ReLU(D2) and the weights sit in shared memory in the canonical K-major no-swizzle core-matrix
// layout.  FP32 accuracy is kept by the 3xTF32 split
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   x_lo = x - trunc_tf32(x)
// (the tensor core ignores the 13 low mantissa bits of a tf32 operand, so the raw fp32 word serves
// as x_hi); the reward is a cancellation of two KLs and needs ~1e-6 relative activations.
// Biases ride along as an extra K column (A[:,100] = 1, B2[n][100] = b2[n]; output column 50 of
// layer 2 is forced to 1 to become the bias column of layer 3).
//
// TMEM columns (all 512): buffer b of A2 at 208*b: [0,104) hi, [104,208) lo;  D2 [416,480);  D3 [480,512).
//
// Software pipeline over the samples m (two mbarriers): after epilogue2(m) one thread issues MMA3(m) and
// MMA2(m+1) back to back, so the tensor pipe runs while the CUDA cores do the KL epilogue of sample m
// and construct the operand of sample m+2:
//   tensor pipe :  MMA2(m) | MMA3(m) MMA2(m+1) | MMA3(m+1) MMA2(m+2) | ...
//   CUDA cores  :  ... epilogue2(m) | KL(m) construct(m+2) | epilogue2(m+1) | KL(m+1) construct(m+3) | ...
#include <cuda_pipeline.h>

#include "pcvae_reward.cuh"
#include "pcvae_tc.cuh"

namespace pcvae {
namespace tc {

constexpr int ROWS = 128;                 // UMMA M
constexpr int K2 = 104;                   // layer-2 reduction (100 + bias + pad)
constexpr int C2 = K2 / 4;                // 16-byte K chunks of B2
constexpr int N2 = 64;                    // layer-2 outputs (50 + ones column + pad)
constexpr int K3 = 56, C3 = K3 / 4;       // layer-3 reduction (50 + bias + pad)
constexpr int N3 = 32;                    // layer-3 outputs (20 + pad)
constexpr int TMEM_COLS = 512;
constexpr int A2_COLS = 2 * K2;           // hi + lo of one A2 buffer
constexpr int COL_D2 = 2 * A2_COLS, COL_D3 = COL_D2 + N2;
constexpr int A_CHUNK = ROWS * 4;         // floats per 16-byte K chunk of the layer-3 operand in shared memory
constexpr int C3W = N2 / 4;               // chunks epilogue 2 writes (all 64 columns of D2)
constexpr int KG = K2 / 4;                // features per thread in the construct phase (26)
constexpr int NBUF = 5;                   // prefetch ring for v / t / baseT
constexpr int ISSUER_WARP = 4;             // MMA-issuing warp (one elected lane); warps 0-3 carry the KL epilogue

__global__ void __launch_bounds__(NT, 1) k_reward_main_tc(const RewardArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D;
    float* B2_hi = smem;                             // [C2][64][4]
    float* B2_lo = B2_hi + C2 * N2 * 4;
    float* B3_hi = B2_lo + C2 * N2 * 4;              // [C3][32][4]
    float* B3_lo = B3_hi + C3 * N3 * 4;
    float* A3_hi = B3_lo + C3 * N3 * 4;              // [C3W][128][4]
    float* A3_lo = A3_hi + C3W * A_CHUNK;
    float* wT_s = A3_lo + C3W * A_CHUNK;             // [K2]
    float* b0_s = wT_s + K2;                         // [40][64]
    float* bT_s = b0_s + BASEW * NPAIR;              // [NBUF][64][40]
    float* v_s = bT_s + NBUF * NPAIR * BASEW;        // [NBUF][64]
    float* t_s = v_s + NBUF * NPAIR;                 // [NBUF][64]
    float* kl_s = t_s + NBUF * NPAIR;                // [2][128]
    int* pn_s = reinterpret_cast<int*>(kl_s + 2 * ROWS); // [64]
    int* pu_s = pn_s + NPAIR;                        // [64]
    uint64_t* bar2 = reinterpret_cast<uint64_t*>(pu_s + NPAIR);
    uint64_t* bar3 = bar2 + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar3 + 1);

    // ---- one-time setup: weights (hi / lo, K-major core-matrix layout), barriers, TMEM ----
    const float* th = a.theta;
    for (int i = tid; i < C2 * N2 * 4; i += NT) {
        const int c = i / (N2 * 4), n = (i >> 2) % N2, k = 4 * c + (i & 3);
        float w = 0.f;
        if (n < H2 && k < H1) w = th[a.L.W2 + n * H1 + k];
        else if (n < H2 && k == H1) w = th[a.L.b2 + n];
        else if (n == H2 && k == H1) w = 1.0f;           // ones column -> bias column of layer 3
        B2_hi[i] = w;
        B2_lo[i] = tf32_lo(w);
    }
    for (int i = tid; i < C3 * N3 * 4; i += NT) {
        const int c = i / (N3 * 4), n = (i >> 2) % N3, k = 4 * c + (i & 3);
        float w = 0.f;
        if (n < LAT2 && k < H2) w = th[a.L.W3 + n * H2 + k];
        else if (n < LAT2 && k == H2) w = th[a.L.b3 + n];
        B3_hi[i] = w;
        B3_lo[i] = tf32_lo(w);
    }
    for (int k = tid; k < K2; k += NT) wT_s[k] = (k < H1) ? th[a.L.W1 + (long)k * D + (D - 1)] : 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar2)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar3)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t ph2 = 0, ph3 = 0;

    constexpr uint32_t IDESC2 = make_idesc(ROWS, N2), IDESC3 = make_idesc(ROWS, N3);
    constexpr uint32_t SBO = 128, B2_LBO = N2 * 16, B3_LBO = N3 * 16, A3_LBO = A_CHUNK * 4;
    // descriptors of k-step 0; a k-step (8 tf32 = two 16-byte chunks) advances the 16-byte-unit address field
    const uint64_t dB2h = make_desc(smem_u32(B2_hi), B2_LBO, SBO), dB2l = make_desc(smem_u32(B2_lo), B2_LBO, SBO);
    const uint64_t dB3h = make_desc(smem_u32(B3_hi), B3_LBO, SBO), dB3l = make_desc(smem_u32(B3_lo), B3_LBO, SBO);
    const uint64_t dA3h = make_desc(smem_u32(A3_hi), A3_LBO, SBO), dA3l = make_desc(smem_u32(A3_lo), A3_LBO, SBO);
    constexpr uint64_t B2_STEP = (2 * B2_LBO) >> 4, B3_STEP = (2 * B3_LBO) >> 4, A3_STEP = (2 * A3_LBO) >> 4;

    auto issue_mma2 = [&](int buf) {      // D2 = A2[buf] * B2^T, 3xTF32 (single thread)
        const uint32_t ah = tmem + A2_COLS * buf, al = ah + K2;
#pragma unroll
        for (int ks = 0; ks < K2 / 8; ++ks) {
            mma_tf32_ts(tmem + COL_D2, al + 8 * ks, dB2h + ks * B2_STEP, IDESC2, ks > 0);
            mma_tf32_ts(tmem + COL_D2, ah + 8 * ks, dB2l + ks * B2_STEP, IDESC2, 1);
            mma_tf32_ts(tmem + COL_D2, ah + 8 * ks, dB2h + ks * B2_STEP, IDESC2, 1);
        }
        mma_commit(bar2);
    };
    auto issue_mma3 = [&]() {             // D3 = A3 * B3^T, 3xTF32 (single thread)
#pragma unroll
        for (int ks = 0; ks < K3 / 8; ++ks) {
            mma_tf32_ss(tmem + COL_D3, dA3l + ks * A3_STEP, dB3h + ks * B3_STEP, IDESC3, ks > 0);
            mma_tf32_ss(tmem + COL_D3, dA3h + ks * A3_STEP, dB3l + ks * B3_STEP, IDESC3, 1);
            mma_tf32_ss(tmem + COL_D3, dA3h + ks * A3_STEP, dB3h + ks * B3_STEP, IDESC3, 1);
        }
        mma_commit(bar3);
    };

    // thread -> (TMEM lane quarter q, column group cg): row r of the tile, features [KG*cg, KG*cg + KG)
    const int q = warp & 3, cg = warp >> 2;
    const int row = 32 * q + lane, pi = row & (NPAIR - 1);
    const bool withT = row >= NPAIR;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    const int k0 = KG * cg;

    const int ptot = a.off[a.N];
    const int ntiles = (ptot + NPAIR - 1) / NPAIR;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * NPAIR;
        if (tid < NPAIR) {
            int n = 0, u = 0;
            if (p0 + tid < ptot) { const int pr = a.pairs[p0 + tid]; n = pr / CANDP; u = pr - n * CANDP; }
            pn_s[tid] = n;
            pu_s[tid] = u;
        }
        __syncthreads();
        auto prefetch = [&](int m) {
            if (m < a.M) {
                const int buf = m % NBUF;
                for (int c = tid; c < NPAIR * (BASEW / 4); c += NT) {
                    const int i = c / (BASEW / 4), qq = c - i * (BASEW / 4);
                    __pipeline_memcpy_async(bT_s + (buf * NPAIR + i) * BASEW + 4 * qq,
                                            a.baseT + ((long)pn_s[i] * a.M + m) * BASEW + 4 * qq, 16);
                }
                if (tid < NPAIR) {
                    const float* r = a.im + (long)m * a.im_ss + (long)pn_s[tid] * D;
                    __pipeline_memcpy_async(v_s + buf * NPAIR + tid, r + pu_s[tid], 4);
                    __pipeline_memcpy_async(t_s + buf * NPAIR + tid, r + (D - 1), 4);
                }
            }
            __pipeline_commit();       // one group per call (possibly empty) keeps wait_prior counts uniform
        };
        prefetch(0);
        prefetch(1);
        prefetch(2);
        // per-thread row state for the whole tile: h0[k] and W1[k][u] of this row's pair
        float H0r[KG], Ur[KG];
        {
            const int n = pn_s[pi], u = pu_s[pi];
            const float* h0 = a.base_in + (long)n * H1;
#pragma unroll
            for (int j = 0; j < KG; ++j) {
                const int k = k0 + j;
                float hv = 0.f, uv = 0.f;
                if (k < H1) { hv = h0[k]; uv = __ldg(th + a.L.W1 + (long)k * D + u); }
                else if (k == H1) hv = 1.0f;          // bias column
                H0r[j] = hv;
                Ur[j] = uv;
            }
        }
        for (int idx = tid; idx < BASEW * NPAIR; idx += NT) {
            const int f = idx / NPAIR, i = idx - f * NPAIR;
            b0_s[idx] = a.base0[(long)pn_s[i] * BASEW + f];
        }
        float acc = 0.f;

        // A2(m) = relu(h0 + v w_u [+ t w_T]) -> TMEM buffer m&1 as hi / lo tf32 operands
        auto construct = [&](int m) {
            const int buf = m % NBUF;
            const float v = v_s[buf * NPAIR + pi], t = withT ? t_s[buf * NPAIR + pi] : 0.f;
            const uint32_t ah0 = lane_addr + A2_COLS * (m & 1) + k0, al0 = ah0 + K2;
            float hi[16], lo[16];
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 2);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        const float h = fmaxf(fmaf(t, wT_s[k0 + j0 + j], fmaf(v, Ur[j0 + j], H0r[j0 + j])), 0.f);
                        hi[j] = h;
                        lo[j] = tf32_lo(h);
                    }
                if (part == 0) { tmem_st16(ah0 + j0, hi); tmem_st16(al0 + j0, lo); }
                else if (part == 1) { tmem_st8(ah0 + j0, hi); tmem_st8(al0 + j0, lo); }
                else { tmem_st2(ah0 + j0, hi); tmem_st2(al0 + j0, lo); }
            }
            tmem_st_wait();
        };
        // KL of sample m against the base posteriors (warps 0-3, one tile row per thread)
        auto kl_epilogue = [&](int m) {
            const int buf = m % NBUF;
            mbar_wait(bar3, ph3, a.status, 4);
            tc_fence_after();
            float o[20];
            tmem_ld16(lane_addr + COL_D3, o);
            tmem_ld4(lane_addr + COL_D3 + 16, o + 16);
            float s = 0.f;
            if (!withT) {
#pragma unroll
                for (int l = 0; l < LAT; ++l) {
                    const float dm = o[l] - b0_s[l * NPAIR + pi];
                    s += (((dm * dm) * b0_s[(2 * LAT + l) * NPAIR + pi] + expf(o[LAT + l]) * b0_s[(3 * LAT + l) * NPAIR + pi] - 1.0f) -
                          o[LAT + l]) + b0_s[(LAT + l) * NPAIR + pi];
                }
            } else {
                const float* bt = bT_s + (buf * NPAIR + pi) * BASEW;
#pragma unroll
                for (int l = 0; l < LAT; ++l) {
                    const float dm = o[l] - bt[l];
                    s += (((dm * dm) * bt[2 * LAT + l] + expf(o[LAT + l]) * bt[3 * LAT + l] - 1.0f) - o[LAT + l]) + bt[LAT + l];
                }
            }
            kl_s[(m & 1) * ROWS + row] = 0.5f * s;
            tc_fence_before();
        };

        // ---- prologue: A2(0), start MMA2(0), A2(1) ----
        __pipeline_wait_prior(1);          // samples 0 and 1 landed
        __syncthreads();
        construct(0);
        tc_fence_before();
        __syncthreads();
        if (warp == ISSUER_WARP) {
            tc_fence_after();
            if (elect_one()) issue_mma2(0);
            __syncwarp();
        }
        if (a.M > 1) construct(1);

        // One CTA barrier per sample.  In iteration m the tensor pipe runs MMA3(m) and MMA2(m+1) while the CUDA
        // cores build the operand of sample m+2; the KL epilogue of sample m-1 is hidden under the wait for MMA2(m).
        for (int m = 0; m < a.M; ++m) {
            prefetch(m + 3);
            __pipeline_wait_prior(1);      // this thread's copies of sample m+2 landed (visible to all after the barrier)
            if (m > 0) {
                if (cg == 0) kl_epilogue(m - 1);
                ph3 ^= 1;
            }
            mbar_wait(bar2, ph2, a.status, 4);          // MMA2(m) (and everything issued before it) complete
            ph2 ^= 1;
            tc_fence_after();
            // ---- epilogue 2: D2 -> ReLU -> hi / lo operand of layer 3 in shared memory ----
            {
                float d[16];
                tmem_ld16(lane_addr + COL_D2 + 16 * cg, d);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float4 hi, lo;
                    hi.x = fmaxf(d[4 * jj + 0], 0.f); hi.y = fmaxf(d[4 * jj + 1], 0.f);
                    hi.z = fmaxf(d[4 * jj + 2], 0.f); hi.w = fmaxf(d[4 * jj + 3], 0.f);
                    lo.x = tf32_lo(hi.x); lo.y = tf32_lo(hi.y); lo.z = tf32_lo(hi.z); lo.w = tf32_lo(hi.w);
                    *reinterpret_cast<float4*>(A3_hi + (4 * cg + jj) * A_CHUNK + row * 4) = hi;
                    *reinterpret_cast<float4*>(A3_lo + (4 * cg + jj) * A_CHUNK + row * 4) = lo;
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();               // A3(m), A2(m+1), kl(m-1) and the prefetched sample m+2 are complete and visible
            if (warp == ISSUER_WARP) {
                tc_fence_after();
                if (elect_one()) {
                    issue_mma3();
                    if (m + 1 < a.M) issue_mma2((m + 1) & 1);
                }
                __syncwarp();
            }
            if (m > 0 && tid < NPAIR) {    // approx_KL += KL_I; approx_KL -= KL_II   (sample m-1)
                const float* kb = kl_s + ((m - 1) & 1) * ROWS;
                acc += kb[tid];
                acc -= kb[NPAIR + tid];
            }
            if (m + 2 < a.M) construct(m + 2);
        }
        if (cg == 0) kl_epilogue(a.M - 1);
        ph3 ^= 1;
        __pipeline_wait_prior(0);
        __syncthreads();
        if (tid < NPAIR) {
            const float* kb = kl_s + ((a.M - 1) & 1) * ROWS;
            acc += kb[tid];
            acc -= kb[NPAIR + tid];
            if (p0 + tid < ptot) a.R[(long)pn_s[tid] * (D - 1) + pu_s[tid]] = acc / (float)a.M;   // evaluate.py:540
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

static size_t smem_bytes() {
    size_t f = 2 * (size_t)C2 * N2 * 4 + 2 * (size_t)C3 * N3 * 4 + 2 * (size_t)C3W * A_CHUNK + K2 + BASEW * NPAIR +
               NBUF * NPAIR * BASEW + 2 * NBUF * NPAIR + 2 * ROWS + 2 * NPAIR + 8;
    return f * sizeof(float) + 128;
}

}  // namespace tc

int reward_main_tc_launch(const RewardArgs& a, int grid, cudaStream_t st) {
    const size_t sm = tc::smem_bytes();
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "reward_main_tc: shared memory %zu B exceeds %d", sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(tc::k_reward_main_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reward_main_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    tc::k_reward_main_tc<<<grid, NT, sm, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reward_main_tc: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // namespace pcvae
