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
// with `tcgen05.mma.cta_group::1.kind::tf32` issued by one thread, operands in shared memory in
// the canonical K-major no-swizzle layout (8-row x 16-byte core matrices), accumulators in TMEM,
// read back with `tcgen05.ld.32x32b`.  FP32 accuracy is kept by the 3xTF32 split
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   x_lo = x - trunc_tf32(x)
// (the tensor core ignores the 13 low mantissa bits of a tf32 operand, so the raw fp32 word serves
// as x_hi); the reward is a cancellation of two KLs and needs ~1e-6 relative activations.
// Biases ride along as an extra K column (A[:,100] = 1, B2[n][100] = b2[n]; output column 50 of
// layer 2 is forced to 1 to become the bias column of layer 3).
//
// Per sample m:  construct A2 (registers hold h0 / W1[:,u] of the thread's row for the whole tile)
//   -> MMA2 (39 instr) -> epilogue 2 (TMEM -> ReLU -> hi/lo split -> A3) -> MMA3 (21 instr)
//   -> epilogue 3 (TMEM -> KL vs base posteriors) -> in-order accumulate.
#include <cuda_pipeline.h>

#include "pcvae_reward.cuh"

namespace pcvae {
namespace tc {

constexpr int ROWS = 128;                 // UMMA M
constexpr int K2 = 104, C2 = K2 / 4;      // layer-2 reduction (100 + bias + pad), 16-byte chunks
constexpr int N2 = 64;                    // layer-2 outputs (50 + ones column + pad)
constexpr int K3 = 56, C3 = K3 / 4;       // layer-3 reduction (50 + bias + pad)
constexpr int C3W = N2 / 4;               // chunks epilogue 2 writes (all 64 columns)
constexpr int N3 = 32;                    // layer-3 outputs (20 + pad)
constexpr int TMEM_COLS = 128;
constexpr int A_CHUNK = ROWS * 4;         // floats per K-chunk of an A operand
constexpr int MAXCH = (C2 + 3) / 4;       // K-chunks per thread in the construct phase (7)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_lo(float v) {
    return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start address,
// leading byte offset (between the two 16-byte K chunks of one MMA), stride byte offset (between 8-row groups)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc),
        "r"(accumulate));
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (int i = 0; i < (1 << 20); ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;   // bounded: never hang the GPU; the caller's results will be wrong and tests catch it
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(NT, 1) k_reward_main_tc(const RewardArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D;
    float* A_hi = smem;                              // [C2][128][4]  (layer 3 reuses the first C3W chunks)
    float* A_lo = A_hi + C2 * A_CHUNK;
    float* B2_hi = A_lo + C2 * A_CHUNK;              // [C2][64][4]
    float* B2_lo = B2_hi + C2 * N2 * 4;
    float* B3_hi = B2_lo + C2 * N2 * 4;              // [C3][32][4]
    float* B3_lo = B3_hi + C3 * N3 * 4;
    float* wT_s = B3_lo + C3 * N3 * 4;               // [K2]
    float* b0_s = wT_s + K2;                         // [40][64]
    float* bT_s = b0_s + BASEW * NPAIR;              // [2][64][40]
    float* v_s = bT_s + 2 * NPAIR * BASEW;           // [2][64]
    float* t_s = v_s + 2 * NPAIR;                    // [2][64]
    float* kl_s = t_s + 2 * NPAIR;                   // [128]
    int* pn_s = reinterpret_cast<int*>(kl_s + ROWS); // [64]
    int* pu_s = pn_s + NPAIR;                        // [64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(pu_s + NPAIR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    // ---- one-time setup: weights (hi / lo, K-major core-matrix layout), barrier, TMEM ----
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
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
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
    uint32_t phase = 0;

    constexpr uint32_t IDESC2 = make_idesc(ROWS, N2), IDESC3 = make_idesc(ROWS, N3);
    const uint32_t aHi = smem_u32(A_hi), aLo = smem_u32(A_lo);
    const uint32_t b2Hi = smem_u32(B2_hi), b2Lo = smem_u32(B2_lo), b3Hi = smem_u32(B3_hi), b3Lo = smem_u32(B3_lo);
    constexpr uint32_t A_LBO = A_CHUNK * 4, SBO = 128, B2_LBO = N2 * 16, B3_LBO = N3 * 16;

    const int row = tid & (ROWS - 1), kq = tid >> 7;     // construct: row of the tile, K-chunk phase
    const int pi = row & (NPAIR - 1);
    const bool withT = row >= NPAIR;

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
        auto prefetch = [&](int m, int buf) {
            for (int c = tid; c < NPAIR * (BASEW / 4); c += NT) {
                const int i = c / (BASEW / 4), q = c - i * (BASEW / 4);
                __pipeline_memcpy_async(bT_s + (buf * NPAIR + i) * BASEW + 4 * q,
                                        a.baseT + ((long)pn_s[i] * a.M + m) * BASEW + 4 * q, 16);
            }
            if (tid < NPAIR) {
                const float* r = a.im + (long)m * a.im_ss + (long)pn_s[tid] * D;
                __pipeline_memcpy_async(v_s + buf * NPAIR + tid, r + pu_s[tid], 4);
                __pipeline_memcpy_async(t_s + buf * NPAIR + tid, r + (D - 1), 4);
            }
            __pipeline_commit();
        };
        prefetch(0, 0);
        // per-thread row state for the whole tile: h0[k] and W1[k][u] of this row's pair, k in the thread's chunks
        float H0r[MAXCH][4], Ur[MAXCH][4];
        {
            const int n = pn_s[pi], u = pu_s[pi];
            const float* h0 = a.base_in + (long)n * H1;
#pragma unroll
            for (int j = 0; j < MAXCH; ++j) {
                const int c = kq + 4 * j;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 4 * c + e;
                    float hv = 0.f, uv = 0.f;
                    if (c < C2) {
                        if (k < H1) { hv = h0[k]; uv = __ldg(th + a.L.W1 + (long)k * D + u); }
                        else if (k == H1) hv = 1.0f;      // bias column
                    }
                    H0r[j][e] = hv;
                    Ur[j][e] = uv;
                }
            }
        }
        for (int idx = tid; idx < BASEW * NPAIR; idx += NT) {
            const int f = idx / NPAIR, i = idx - f * NPAIR;
            b0_s[idx] = a.base0[(long)pn_s[i] * BASEW + f];
        }
        float acc = 0.f;
        __pipeline_wait_prior(0);
        __syncthreads();

        for (int m = 0; m < a.M; ++m) {
            const int buf = m & 1;
            if (m + 1 < a.M) prefetch(m + 1, buf ^ 1);
            // ---- construct A2 = relu(h0 + v w_u [+ t w_T]) as hi / lo tf32 operands ----
            {
                const float v = v_s[buf * NPAIR + pi], t = withT ? t_s[buf * NPAIR + pi] : 0.f;
#pragma unroll
                for (int j = 0; j < MAXCH; ++j) {
                    const int c = kq + 4 * j;
                    if (c < C2) {
                        const float4 wt = *reinterpret_cast<const float4*>(wT_s + 4 * c);
                        float4 hi, lo;
                        hi.x = fmaxf(fmaf(t, wt.x, fmaf(v, Ur[j][0], H0r[j][0])), 0.f);
                        hi.y = fmaxf(fmaf(t, wt.y, fmaf(v, Ur[j][1], H0r[j][1])), 0.f);
                        hi.z = fmaxf(fmaf(t, wt.z, fmaf(v, Ur[j][2], H0r[j][2])), 0.f);
                        hi.w = fmaxf(fmaf(t, wt.w, fmaf(v, Ur[j][3], H0r[j][3])), 0.f);
                        lo.x = tf32_lo(hi.x); lo.y = tf32_lo(hi.y); lo.z = tf32_lo(hi.z); lo.w = tf32_lo(hi.w);
                        *reinterpret_cast<float4*>(A_hi + c * A_CHUNK + row * 4) = hi;
                        *reinterpret_cast<float4*>(A_lo + c * A_CHUNK + row * 4) = lo;
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            // ---- layer 2 on the tensor cores ----
            if (tid == 0) {
                tc_fence_after();
#pragma unroll 1
                for (int ks = 0; ks < K2 / 8; ++ks) {
                    const uint64_t dAh = make_desc(aHi + ks * 2 * A_LBO, A_LBO, SBO), dAl = make_desc(aLo + ks * 2 * A_LBO, A_LBO, SBO);
                    const uint64_t dBh = make_desc(b2Hi + ks * 2 * B2_LBO, B2_LBO, SBO), dBl = make_desc(b2Lo + ks * 2 * B2_LBO, B2_LBO, SBO);
                    mma_tf32(tmem, dAl, dBh, IDESC2, ks > 0);
                    mma_tf32(tmem, dAh, dBl, IDESC2, 1);
                    mma_tf32(tmem, dAh, dBh, IDESC2, 1);
                }
                mma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            // ---- epilogue 2: TMEM -> ReLU -> hi/lo -> A3 (rows = TMEM lanes; 16 columns per thread) ----
            {
                const int q = warp & 3, cg = warp >> 2, r = 32 * q + lane;
                float d[16];
                tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * cg), d);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = 4 * cg + jj;
                    float4 hi, lo;
                    hi.x = fmaxf(d[4 * jj + 0], 0.f); hi.y = fmaxf(d[4 * jj + 1], 0.f);
                    hi.z = fmaxf(d[4 * jj + 2], 0.f); hi.w = fmaxf(d[4 * jj + 3], 0.f);
                    lo.x = tf32_lo(hi.x); lo.y = tf32_lo(hi.y); lo.z = tf32_lo(hi.z); lo.w = tf32_lo(hi.w);
                    *reinterpret_cast<float4*>(A_hi + c * A_CHUNK + r * 4) = hi;
                    *reinterpret_cast<float4*>(A_lo + c * A_CHUNK + r * 4) = lo;
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            // ---- layer 3 on the tensor cores ----
            if (tid == 0) {
                tc_fence_after();
#pragma unroll 1
                for (int ks = 0; ks < K3 / 8; ++ks) {
                    const uint64_t dAh = make_desc(aHi + ks * 2 * A_LBO, A_LBO, SBO), dAl = make_desc(aLo + ks * 2 * A_LBO, A_LBO, SBO);
                    const uint64_t dBh = make_desc(b3Hi + ks * 2 * B3_LBO, B3_LBO, SBO), dBl = make_desc(b3Lo + ks * 2 * B3_LBO, B3_LBO, SBO);
                    mma_tf32(tmem + N2, dAl, dBh, IDESC3, ks > 0);
                    mma_tf32(tmem + N2, dAh, dBl, IDESC3, 1);
                    mma_tf32(tmem + N2, dAh, dBh, IDESC3, 1);
                }
                mma_commit(bar);
            }
            // ---- epilogue 3 (warps 0-3): TMEM -> KL against the base posterior of the row's variant ----
            if (warp < 4) {
                mbar_wait(bar, phase);
                tc_fence_after();
                const int r = 32 * warp + lane, i = r & (NPAIR - 1);
                float o[20];
                tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)N2, o);
                tmem_ld4(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(N2 + 16), o + 16);
                float s = 0.f;
                if (r < NPAIR) {
#pragma unroll
                    for (int l = 0; l < LAT; ++l) {
                        const float dm = o[l] - b0_s[l * NPAIR + i];
                        s += (((dm * dm) * b0_s[(2 * LAT + l) * NPAIR + i] + expf(o[LAT + l]) * b0_s[(3 * LAT + l) * NPAIR + i] - 1.0f) -
                              o[LAT + l]) + b0_s[(LAT + l) * NPAIR + i];
                    }
                } else {
                    const float* bt = bT_s + (buf * NPAIR + i) * BASEW;
#pragma unroll
                    for (int l = 0; l < LAT; ++l) {
                        const float dm = o[l] - bt[l];
                        s += (((dm * dm) * bt[2 * LAT + l] + expf(o[LAT + l]) * bt[3 * LAT + l] - 1.0f) - o[LAT + l]) + bt[LAT + l];
                    }
                }
                kl_s[r] = 0.5f * s;
                tc_fence_before();
            }
            phase ^= 1;
            __syncthreads();
            if (tid < NPAIR) { acc += kl_s[tid]; acc -= kl_s[NPAIR + tid]; }     // approx_KL += KL_I; approx_KL -= KL_II
            __pipeline_wait_prior(0);
            __syncthreads();
        }
        if (tid < NPAIR && p0 + tid < ptot) a.R[(long)pn_s[tid] * (D - 1) + pu_s[tid]] = acc / (float)a.M;   // evaluate.py:540
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
    size_t f = 2 * (size_t)C2 * A_CHUNK + 2 * (size_t)C2 * N2 * 4 + 2 * (size_t)C3 * N3 * 4 + K2 + BASEW * NPAIR +
               2 * NPAIR * BASEW + 4 * NPAIR + ROWS + 2 * NPAIR + 8;
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
