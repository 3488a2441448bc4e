// tcgen05 (5th-gen tensor core) version of the fused decoder + loss + decoder-backward step (k_dec in
// pcvae_train.cu, mode PCVAE_DEC_TRAIN): same mathematics (src/models/VAE.py:397-467 forward / loss and the
// autograd of train.py:115), 128-row tiles, every dense product on the tensor cores in fp32-accurate 3xTF32
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   x_lo = x - trunc_tf32(x).
//
// Three kernels, because the weight images of the forward and of the data-gradient products do not fit one
// SM's shared memory together (tf32 operands cannot be read transposed from a no-swizzle image, so the
// data-gradient products need their own transposed copy), and the weight-gradient products need BOTH operands
// in shared memory:
//
//   k_dec_fwd_tc   F4: [128 x 16]  z|1  x W4aug^T -> 64     F5: [128 x 56] h4|1 x W5aug^T -> 112
//                  F6: [128 x 104] h5|1 x W6aug^T -> round16(D);  sigmoid, masked NLL sums, dL/d(pre-sigmoid)
//   k_dec_bwd_tc   X6: [128 x 104] dpre6 x W6 -> 112   X5: [128 x 104] dpre5 x W5 -> 64   X4: [128 x 56] dpre4 x W4 -> 16
//                  KL sums, d_mean / d_logvar (reparameterisation folded in)
//   k_wgrad_tc     dWaug[m][n] = sum_rows dpre[row][m] * (act|1)[row][n]  for layers 6, 5, 4 (bias = the 1 column)
//
// Biases ride along as an extra K column (the augmented weights hold b at column `in`, and one extra output
// row generates the constant-1 column of the next layer).  In the first two kernels the activation operand
// lives in TENSOR MEMORY (written by the epilogue threads with tcgen05.st, consumed by the MMA in its
// A-from-TMEM form), accumulators are in TMEM, and the weights sit in shared memory once per CTA as hi / lo
// images in the canonical K-major no-swizzle core-matrix layout [k/4][n][4].  Between the kernels the
// activations travel through HBM feature-major ([feature][row]: a warp's 32 rows make one 128-byte store), which
// is exactly the K-major operand layout of the weight-gradient GEMM (its reduction runs over rows); that
// kernel streams 32-row slabs with cp.async through a 3-stage ring and keeps its accumulator in TMEM for the
// whole launch.  Its per-CTA partials land in the [grid][param_count] layout k_dec uses, so
// pcvae_reduce_grads is unchanged.
//
// TMEM columns (all 512), both row-tile kernels:  RA [0,224): 112 hi + 112 lo    RB [224,336): 56 hi + 56 lo
//                                                 ACC1 [336,448)                 ACC2 [448,512)
#include <cuda_pipeline.h>

#include "pcvae_tc.cuh"
#include "pcvae_train.cuh"

namespace pcvae {
namespace tc {

constexpr int ROWS = 128;
constexpr int RA_HI = 0, RA_LO = 112, RB_HI = 224, RB_LO = 280, ACC1 = 336, ACC2 = 448;
constexpr int DEC_ISSUER_WARP = 4;
// forward images  [K/4 chunks][N rows][4]
constexpr int F4_C = 4, F4_N = 64;        // K = 16 (z|1), 50 outputs + the constant-1 generator
constexpr int F5_C = 14, F5_N = 112;      // K = 56 (h4|1), 100 outputs + the constant-1 generator
constexpr int F6_C = 26;                  // K = 104 (h5|1), round16(D) outputs
// data-gradient images (transposed weights)
constexpr int X6_C = 26, X6_N = 112;      // K = 104 (d), 100 inputs k
constexpr int X5_C = 26, X5_N = 64;       // K = 104 (n), 50 inputs k
constexpr int X4_C = 14, X4_N = 16;       // K = 56 (n), 10 inputs k

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 3xTF32 product: activation operand in TMEM (hi at a_hi, lo at a_lo), weight image (K-major) in shared memory;
// a k-step (8 tf32) is two 16-byte chunks of the image
__device__ __forceinline__ void issue_3x(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint64_t b_step,
                                         int ksteps, uint32_t idesc) {
    for (int ks = 0; ks < ksteps; ++ks) {
        mma_tf32_ts(acc, a_lo + 8 * ks, b_hi + ks * b_step, idesc, ks > 0);
        mma_tf32_ts(acc, a_hi + 8 * ks, b_lo + ks * b_step, idesc, 1);
        mma_tf32_ts(acc, a_hi + 8 * ks, b_hi + ks * b_step, idesc, 1);
    }
}

__device__ __forceinline__ void load_mask4(const void* m, long gi, int kind, float* o) {
    if (kind == PCVAE_MASK_U8) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const unsigned char*>(m) + gi);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = ((w >> (8 * j)) & 0xFFu) ? 1.f : 0.f;
    } else {
        const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(m) + gi);
        o[0] = v.x != 0.f ? 1.f : 0.f; o[1] = v.y != 0.f ? 1.f : 0.f; o[2] = v.z != 0.f ? 1.f : 0.f; o[3] = v.w != 0.f ? 1.f : 0.f;
    }
}

// hi / lo image of an augmented weight matrix; w(n, k) supplies element (row n, column k)
template <typename F>
__device__ __forceinline__ void build_image(float* hi, float* lo, int chunks, int nrows, int tid, F w) {
    for (int i = tid; i < chunks * nrows * 4; i += NT) {
        const int c = i / (nrows * 4), n = (i >> 2) % nrows, k = 4 * c + (i & 3);
        const float v = w(n, k);
        hi[i] = v;
        lo[i] = tf32_lo(v);
    }
}

struct TileCtx {
    uint32_t tmem, lane_addr, ph;
    int q, cg, row, c28, c16;
};

// barrier + (one elected lane) MMA issue + commit; every thread then waits for the batch
template <typename Issue>
__device__ __forceinline__ void run_mma(TileCtx& cx, uint64_t* bar, int warp, Issue&& issue) {
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == DEC_ISSUER_WARP) {
        tc_fence_after();
        if (elect_one()) {
            issue();
            mma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, cx.ph);
    cx.ph ^= 1;
    tc_fence_after();
}

__device__ __forceinline__ void tc_setup(TileCtx& cx, uint64_t* bar, uint32_t* slot, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cx.tmem = *slot;
    cx.ph = 0;
    cx.q = warp & 3;
    cx.cg = warp >> 2;
    cx.row = 32 * cx.q + lane;
    cx.lane_addr = cx.tmem + ((uint32_t)(32 * cx.q) << 16);
    cx.c28 = 28 * cx.cg;
    cx.c16 = 16 * cx.cg;
}

__device__ __forceinline__ void tc_teardown(const TileCtx& cx, int tid) {
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "r"(512));
    }
}

// 28 accumulator columns of this thread, in parts of 16 / 8 / 4
__device__ __forceinline__ void ld_part(uint32_t addr, int part, float* v) {
    if (part == 0) tmem_ld16(addr, v);
    else if (part == 1) tmem_ld8(addr + 16, v);
    else tmem_ld4(addr + 24, v);
}
__device__ __forceinline__ void st_part(uint32_t addr, int part, const float* v) {
    if (part == 0) tmem_st16(addr, v);
    else if (part == 1) tmem_st8(addr + 16, v);
    else tmem_st4(addr + 24, v);
}

// ------------------------------------------------------------------------------------------------
// forward + loss
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) k_dec_fwd_tc(const DecArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ float red_s[NWARP][PCVAE_NSUMS];
    __shared__ __align__(8) uint64_t bar_s;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D, N6 = (D + 15) & ~15;
    float* W4h = smem;
    float* W4l = W4h + F4_C * F4_N * 4;
    float* W5h = W4l + F4_C * F4_N * 4;
    float* W5l = W5h + F5_C * F5_N * 4;
    float* W6h = W5l + F5_C * F5_N * 4;
    float* W6l = W6h + F6_C * N6 * 4;
    const float* th = a.theta;
    const Layout L = a.L;
    build_image(W4h, W4l, F4_C, F4_N, tid, [&](int n, int k) {
        if (n < G1 && k < LAT) return th[L.W4 + n * LAT + k];
        if (n < G1 && k == LAT) return th[L.b4 + n];
        return (n == G1 && k == LAT) ? 1.0f : 0.0f;            // constant-1 output -> bias column of layer 5
    });
    build_image(W5h, W5l, F5_C, F5_N, tid, [&](int n, int k) {
        if (n < G2 && k < G1) return th[L.W5 + n * G1 + k];
        if (n < G2 && k == G1) return th[L.b5 + n];
        return (n == G2 && k == G1) ? 1.0f : 0.0f;            // constant-1 output -> bias column of layer 6
    });
    build_image(W6h, W6l, F6_C, N6, tid, [&](int n, int k) {
        if (n < D && k < G2) return th[L.W6 + n * G2 + k];
        if (n < D && k == G2) return th[L.b6 + n];
        return 0.0f;
    });
    TileCtx cx;
    tc_setup(cx, &bar_s, &tmem_slot, tid);
    const uint32_t tmem = cx.tmem, lane_addr = cx.lane_addr;
    const int cg = cx.cg, row = cx.row, c28 = cx.c28, c16 = cx.c16;

    const uint32_t cs4 = F4_N * 16, cs5 = F5_N * 16, cs6 = N6 * 16;     // chunk strides (LBO); 8-row groups are 128 B apart (SBO)
    const uint64_t f4h = make_desc(smem_u32(W4h), cs4, 128), f4l = make_desc(smem_u32(W4l), cs4, 128);
    const uint64_t f5h = make_desc(smem_u32(W5h), cs5, 128), f5l = make_desc(smem_u32(W5l), cs5, 128);
    const uint64_t f6h = make_desc(smem_u32(W6h), cs6, 128), f6l = make_desc(smem_u32(W6l), cs6, 128);
    const uint64_t fs4 = (2 * cs4) >> 4, fs5 = (2 * cs5) >> 4, fs6 = (2 * cs6) >> 4;
    const uint32_t idF4 = make_idesc(ROWS, F4_N), idF5 = make_idesc(ROWS, F5_N), idF6 = make_idesc(ROWS, N6);

    // NLL constants exactly as k_dec / torch.distributions.Normal compute them
    const float scale = expf(a.x_logvar * 0.5f);
    const float var = scale * scale;
    const float inv2var = 1.0f / (2.0f * var);
    const float inv_var = 1.0f / var;
    const float log_scale = logf(scale);
    const float alpha = a.alpha, ls = a.loss_scale;
    float s_req = 0.f, s_rep = 0.f, s_red = 0.f, s_imp = 0.f, s_sse = 0.f;
    const long R2P = a.R2P;

    const int ntiles = (a.B + ROWS - 1) / ROWS;
    const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = t * ROWS;
        const int grow = row0 + row;
        const bool ok = grow < a.B;
        {   // pull the next tile of this CTA towards L2 while this one is processed
            const int tn = t + gridDim.x;
            if (tn < ntiles) {
                const long r0 = (long)tn * ROWS, nrows = min((long)ROWS, (long)a.B - r0);
                prefetch_l2(a.x + r0 * D, nrows * D * 4, tid);
                for (int b = 0; b < a.nbr; ++b) prefetch_l2((const char*)a.mask[b] + r0 * D * msz, nrows * D * msz, tid);
            }
        }
        for (int br = 0; br < a.nbr; ++br) {
            const long wrow = (long)br * a.B + grow;          // column in the feature-major scratch
            // ---- z | 1 -> RA, HBM ----
            if (cg == 0) {
                float v[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
                if (ok) {
                    const float2* zp = reinterpret_cast<const float2*>(a.z[br] + (long)grow * LAT);
#pragma unroll
                    for (int j = 0; j < LAT / 2; ++j) { const float2 p2 = zp[j]; v[2 * j] = p2.x; v[2 * j + 1] = p2.y; }
                }
                v[LAT] = 1.0f;
#pragma unroll
                for (int j = 0; j < 16; ++j) lo[j] = tf32_lo(v[j]);
                tmem_st16(lane_addr + RA_HI, v);
                tmem_st16(lane_addr + RA_HI + 16, lo);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < TCW_Z; ++j) a.ws_zT[j * R2P + wrow] = v[j];
                }
            }
            run_mma(cx, &bar_s, warp, [&] { issue_3x(tmem + ACC2, tmem + RA_HI, tmem + RA_HI + 16, f4h, f4l, fs4, F4_C / 2, idF4); });

            // ---- h4 = relu(acc4) | 1 -> RB, HBM ----
            uint32_t m4 = 0;                                  // relu mask of this thread's 16 h4 columns
            {
                float v[16], lo[16];
                tmem_ld16(lane_addr + ACC2 + c16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (v[j] > 0.f) m4 |= 1u << j; else v[j] = 0.f;
                    lo[j] = tf32_lo(v[j]);
                }
                if (cg < 3) { tmem_st16(lane_addr + RB_HI + c16, v); tmem_st16(lane_addr + RB_LO + c16, lo); }
                else { tmem_st8(lane_addr + RB_HI + c16, v); tmem_st8(lane_addr + RB_LO + c16, lo); }
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c16 + j < TCW_H4) a.ws_h4T[(c16 + j) * R2P + wrow] = v[j];
                }
            }
            run_mma(cx, &bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RB_HI, tmem + RB_LO, f5h, f5l, fs5, F5_C / 2, idF5); });

            // ---- h5 = relu(acc5) | 1 -> RA, HBM ----
            uint32_t m5 = 0;                                  // relu mask of this thread's 28 h5 columns
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                float v[16], lo[16];
                ld_part(lane_addr + ACC1 + c28, part, v);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        if (v[j] > 0.f) m5 |= 1u << (j0 + j); else v[j] = 0.f;
                        lo[j] = tf32_lo(v[j]);
                    }
                st_part(lane_addr + RA_HI + c28, part, v);
                st_part(lane_addr + RA_LO + c28, part, lo);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < cnt && c28 + j0 + j < TCW_H5) a.ws_h5T[(c28 + j0 + j) * R2P + wrow] = v[j];
                }
            }
            if (ok) {
                a.ws_relu[wrow * 8 + cg] = m5;
                a.ws_relu[wrow * 8 + 4 + cg] = m4;
            }
            run_mma(cx, &bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RA_HI, tmem + RA_LO, f6h, f6l, fs6, F6_C / 2, idF6); });

            // ---- x_hat = sigmoid(acc6): loss terms, dL/d(pre-sigmoid) -> HBM ----
            {
                float* xo = a.xhat[br];
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                    float v[16];
                    ld_part(lane_addr + ACC1 + c28, part, v);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (4 * g >= cnt) continue;
                        const int c = c28 + j0 + 4 * g;
                        float dpre[4] = {0.f, 0.f, 0.f, 0.f};
                        if (ok && c < D) {
                            const long gi = (long)grow * D + c;
                            const float4 xv4 = *reinterpret_cast<const float4*>(a.x + gi);
                            const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
                            float m[4], mp[4] = {0.f, 0.f, 0.f, 0.f}, xh[4];
                            load_mask4(a.mask[0], gi, a.mask_kind, m);
                            if (a.nbr > 1) load_mask4(a.mask[1], gi, a.mask_kind, mp);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                xh[j] = 1.0f / (1.0f + expf(-v[4 * g + j]));
                                const float diff = xv[j] - xh[j];
                                const float nll = fmaf(diff * diff, inv2var, log_scale);
                                float coef;
                                if (br == 0) {
                                    s_req += m[j] * nll;
                                    s_red += m[j] * (1.f - mp[j]) * nll;
                                    s_imp += (1.f - m[j]) * nll;
                                    s_sse += (1.f - m[j]) * diff * diff;
                                    coef = (1.f - alpha) * m[j] + alpha * m[j] * (1.f - mp[j]);
                                } else {
                                    s_rep += mp[j] * nll;
                                    coef = alpha * mp[j];
                                }
                                dpre[j] = coef * (xh[j] - xv[j]) * inv_var * ls * xh[j] * (1.f - xh[j]);
                            }
                            if (xo) *reinterpret_cast<float4*>(xo + gi) = make_float4(xh[0], xh[1], xh[2], xh[3]);
                        }
                        if (ok && c < TCW_H5) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) a.ws_dp6T[(c + j) * R2P + wrow] = dpre[j];
                        }
                    }
                }
            }
            tc_fence_before();      // the next branch overwrites RA / ACC2 only after its own barrier
        }
    }

    {
        float v[PCVAE_NSUMS] = {s_req, s_rep, 0.f, 0.f, 0.f, s_red, s_imp, s_sse};     // KL sums come from k_dec_bwd_tc
#pragma unroll
        for (int j = 0; j < PCVAE_NSUMS; ++j) {
            const float w = warp_sum(v[j]);
            if (lane == 0) red_s[warp][j] = w;
        }
        __syncthreads();
        if (tid < PCVAE_NSUMS && tid != PCVAE_S_KL_Q && tid != PCVAE_S_KL_P && tid != PCVAE_S_KL_REG) {
            float s = 0.f;
            for (int w = 0; w < NWARP; ++w) s += red_s[w][tid];
            a.sums_partials[blockIdx.x * PCVAE_NSUMS + tid] = s;
        }
    }
    tc_teardown(cx, tid);
}

// ------------------------------------------------------------------------------------------------
// data gradients + latent-space terms
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) k_dec_bwd_tc(const DecArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ float red_s[NWARP][3];
    __shared__ __align__(8) uint64_t bar_s;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D;
    float* T6h = smem;
    float* T6l = T6h + X6_C * X6_N * 4;
    float* T5h = T6l + X6_C * X6_N * 4;
    float* T5l = T5h + X5_C * X5_N * 4;
    float* T4h = T5l + X5_C * X5_N * 4;
    float* T4l = T4h + X4_C * X4_N * 4;
    const float* th = a.theta;
    const Layout L = a.L;
    // transposed weights: image row = layer INPUT index, image column (reduction) = layer OUTPUT index
    build_image(T6h, T6l, X6_C, X6_N, tid, [&](int k, int d) { return (k < G2 && d < D) ? th[L.W6 + d * G2 + k] : 0.0f; });
    build_image(T5h, T5l, X5_C, X5_N, tid, [&](int k, int n) { return (k < G1 && n < G2) ? th[L.W5 + n * G1 + k] : 0.0f; });
    build_image(T4h, T4l, X4_C, X4_N, tid, [&](int k, int n) { return (k < LAT && n < G1) ? th[L.W4 + n * LAT + k] : 0.0f; });
    TileCtx cx;
    tc_setup(cx, &bar_s, &tmem_slot, tid);
    const uint32_t tmem = cx.tmem, lane_addr = cx.lane_addr;
    const int cg = cx.cg, row = cx.row, c28 = cx.c28, c16 = cx.c16;

    const uint32_t cs6 = X6_N * 16, cs5 = X5_N * 16, cs4 = X4_N * 16;
    const uint64_t x6h = make_desc(smem_u32(T6h), cs6, 128), x6l = make_desc(smem_u32(T6l), cs6, 128);
    const uint64_t x5h = make_desc(smem_u32(T5h), cs5, 128), x5l = make_desc(smem_u32(T5l), cs5, 128);
    const uint64_t x4h = make_desc(smem_u32(T4h), cs4, 128), x4l = make_desc(smem_u32(T4l), cs4, 128);
    const uint64_t xs6 = (2 * cs6) >> 4, xs5 = (2 * cs5) >> 4, xs4 = (2 * cs4) >> 4;
    const uint32_t idX6 = make_idesc(ROWS, X6_N), idX5 = make_idesc(ROWS, X5_N), idX4 = make_idesc(ROWS, X4_N);

    const float alpha = a.alpha, ls = a.loss_scale;
    float s_klq = 0.f, s_klp = 0.f, s_klr = 0.f;
    const long R2P = a.R2P;

    const int ntiles = (a.B + ROWS - 1) / ROWS;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = t * ROWS;
        const int grow = row0 + row;
        const bool ok = grow < a.B;
        for (int br = 0; br < a.nbr; ++br) {
            const long wrow = (long)br * a.B + grow;
            uint32_t m5 = 0, m4 = 0;
            if (ok) { m5 = a.ws_relu[wrow * 8 + cg]; m4 = a.ws_relu[wrow * 8 + 4 + cg]; }
            // ---- dpre6 (HBM) -> RA ----
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                float v[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        const int c = c28 + j0 + j;
                        v[j] = (ok && c < TCW_H5) ? a.ws_dp6T[c * R2P + wrow] : 0.f;
                        lo[j] = tf32_lo(v[j]);
                    }
                st_part(lane_addr + RA_HI + c28, part, v);
                st_part(lane_addr + RA_LO + c28, part, lo);
            }
            run_mma(cx, &bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RA_HI, tmem + RA_LO, x6h, x6l, xs6, X6_C / 2, idX6); });

            // ---- dpre5 = dh5 * relu'(h5) -> RA, HBM ----
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                float v[16], lo[16];
                ld_part(lane_addr + ACC1 + c28, part, v);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        if (!((m5 >> (j0 + j)) & 1u) || c28 + j0 + j >= G2) v[j] = 0.f;     // column G2 is the bias column
                        lo[j] = tf32_lo(v[j]);
                    }
                st_part(lane_addr + RA_HI + c28, part, v);
                st_part(lane_addr + RA_LO + c28, part, lo);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < cnt && c28 + j0 + j < TCW_H5) a.ws_dp5T[(c28 + j0 + j) * R2P + wrow] = v[j];
                }
            }
            run_mma(cx, &bar_s, warp, [&] { issue_3x(tmem + ACC2, tmem + RA_HI, tmem + RA_LO, x5h, x5l, xs5, X5_C / 2, idX5); });

            // ---- dpre4 = dh4 * relu'(h4) -> RB, HBM ----
            {
                float v[16], lo[16];
                tmem_ld16(lane_addr + ACC2 + c16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (!((m4 >> j) & 1u) || c16 + j >= G1) v[j] = 0.f;
                    lo[j] = tf32_lo(v[j]);
                }
                if (cg < 3) { tmem_st16(lane_addr + RB_HI + c16, v); tmem_st16(lane_addr + RB_LO + c16, lo); }
                else { tmem_st8(lane_addr + RB_HI + c16, v); tmem_st8(lane_addr + RB_LO + c16, lo); }
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c16 + j < TCW_H4) a.ws_dp4T[(c16 + j) * R2P + wrow] = v[j];
                }
            }
            run_mma(cx, &bar_s, warp, [&] { issue_3x(tmem + ACC2, tmem + RB_HI, tmem + RB_LO, x4h, x4l, xs4, X4_C / 2, idX4); });

            // ---- latent-space terms: KL sums, d_mean / d_logvar (one row per thread of column group 0) ----
            if (cg == 0) {
                float dzv[16];
                tmem_ld16(lane_addr + ACC2, dzv);
                if (ok) {
#pragma unroll
                    for (int l = 0; l < LAT; ++l) {
                        const long gi = (long)grow * LAT + l;
                        const float mq = a.mean[0][gi], lq = a.logvar[0][gi];
                        const float eq = expf(lq);
                        float mp_ = 0.f, lp = 0.f, ep = 1.f;
                        if (a.nbr > 1) { mp_ = a.mean[1][gi]; lp = a.logvar[1][gi]; ep = expf(lp); }
                        const float dmu = mq - mp_;
                        if (br == 0) {
                            s_klq += 0.5f * (eq + mq * mq - 1.f - lq);
                            if (a.nbr > 1) {
                                s_klp += 0.5f * (ep + mp_ * mp_ - 1.f - lp);
                                s_klr += 0.5f * (expf(lq - lp) + dmu * dmu / ep - 1.f - (lq - lp));
                            }
                        }
                        const float dz = dzv[l];
                        float gm, gv;
                        if (br == 0) {
                            gm = (1.f - alpha) * a.beta_w * mq;
                            gv = (1.f - alpha) * a.beta_w * 0.5f * (eq - 1.f);
                            if (a.nbr > 1) {
                                gm += alpha * dmu / ep;
                                gv += alpha * 0.5f * (expf(lq - lp) - 1.f);
                            }
                            gm = fmaf(gm, ls, dz);
                            gv = fmaf(gv, ls, dz * 0.5f * expf(lq * 0.5f) * (a.eps[0] ? a.eps[0][gi] : 0.f));
                        } else {
                            gm = alpha * a.beta_w * mp_ - alpha * dmu / ep;
                            gv = alpha * a.beta_w * 0.5f * (ep - 1.f) + alpha * 0.5f * (1.f - (eq + dmu * dmu) / ep);
                            gm = fmaf(gm, ls, dz);
                            gv = fmaf(gv, ls, dz * 0.5f * expf(lp * 0.5f) * (a.eps[1] ? a.eps[1][gi] : 0.f));
                        }
                        a.d_mean[br][gi] = gm;
                        a.d_logvar[br][gi] = gv;
                    }
                }
            }
            tc_fence_before();
        }
    }

    {
        const float v[3] = {s_klq, s_klp, s_klr};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float w = warp_sum(v[j]);
            if (lane == 0) red_s[warp][j] = w;
        }
        __syncthreads();
        if (tid < 3) {
            float s = 0.f;
            for (int w = 0; w < NWARP; ++w) s += red_s[w][tid];
            const int slot = tid == 0 ? PCVAE_S_KL_Q : (tid == 1 ? PCVAE_S_KL_P : PCVAE_S_KL_REG);
            a.sums_partials[blockIdx.x * PCVAE_NSUMS + slot] = s;
        }
    }
    tc_teardown(cx, tid);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient  dWaug[m][n] = sum_r AT[m][r] * BT[n][r]  over all rows of both branches, 3xTF32, both operands
// K-major from shared memory (K = rows).  32-row slabs are streamed with cp.async through a 3-stage ring; the
// accumulator lives in TMEM for the whole launch and is written once, as this CTA's partial, into
// gp[cta][W_off + m*Kin + n] (n < Kin) and gp[cta][b_off + m] (n == Kin: the constant-1 row of BT).
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
    const float* AT; int Ma;              // [>= Ma][R2P]: pre-activation gradients, feature-major
    const float* BT; int Kin;             // [>= Kin + 1][R2P]: layer input | 1, feature-major
    int Nb;                               // MMA N: round16(Kin + 1)
    long R2P;                             // row pitch (multiple of 32; columns >= the real row count are zero)
    float* gp; long P; int W_off, b_off;
};

constexpr int SLAB = 32, WG_STAGES = 3, WG_CH = SLAB / 4;
constexpr int WG_ACS = (128 + 1) * 4, WG_BCS = (112 + 1) * 4;       // chunk strides in floats (+1 row: conflict-free cp.async writes)
constexpr int WG_A_FLOATS = WG_CH * WG_ACS, WG_B_FLOATS = WG_CH * WG_BCS;
constexpr int WG_STAGE_FLOATS = 2 * WG_A_FLOATS + 2 * WG_B_FLOATS;

__global__ void __launch_bounds__(NT, 1) k_wgrad_tc(const WgradArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t free_bar[WG_STAGES];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* gp = a.gp + (long)blockIdx.x * a.P;
    const long nslab = a.R2P / SLAB;
    const long mine = blockIdx.x < nslab ? (nslab - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (mine == 0) {                                    // no rows for this CTA: its partial is zero
        for (int i = tid; i < a.Ma * (a.Kin + 1); i += NT) {
            const int m = i / (a.Kin + 1), n = i - m * (a.Kin + 1);
            if (n < a.Kin) gp[a.W_off + m * a.Kin + n] = 0.f; else gp[a.b_off + m] = 0.f;
        }
        return;
    }
    for (int i = tid; i < WG_STAGES * WG_STAGE_FLOATS; i += NT) smem[i] = 0.f;   // rows the copies never touch stay zero
    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&free_bar[s])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc(128, a.Nb);
    const int nb1 = a.Kin + 1;                          // real rows of BT
    const int items = WG_CH * (a.Ma + nb1);             // 16-byte copies per slab

    auto stage_ptr = [&](int s) { return smem + s * WG_STAGE_FLOATS; };   // Ahi | Alo | Bhi | Blo
    auto load_slab = [&](long i) {                        // cp.async this CTA's i-th slab into stage i % 3
        if (i < mine) {
            float* st = stage_ptr((int)(i % WG_STAGES));
            const long r0 = (blockIdx.x + i * gridDim.x) * SLAB;
            for (int idx = tid; idx < items; idx += NT) {
                const int f = idx / WG_CH, c = idx - f * WG_CH;          // feature row, 4-row chunk
                if (f < a.Ma) __pipeline_memcpy_async(st + c * WG_ACS + f * 4, a.AT + (long)f * a.R2P + r0 + 4 * c, 16);
                else __pipeline_memcpy_async(st + 2 * WG_A_FLOATS + c * WG_BCS + (f - a.Ma) * 4, a.BT + (long)(f - a.Ma) * a.R2P + r0 + 4 * c, 16);
            }
        }
        __pipeline_commit();
    };
    uint32_t free_ph[WG_STAGES] = {0, 0, 0};
    load_slab(0);
    load_slab(1);
    for (long i = 0; i < mine; ++i) {
        const int s = (int)(i % WG_STAGES);
        if (i >= 1) {                                     // stage (i+2)%3 was read by the MMAs of slab i-1
            const int sp = (int)((i + 2) % WG_STAGES);
            mbar_wait(&free_bar[sp], free_ph[sp]);
            free_ph[sp] ^= 1;
        }
        load_slab(i + 2);
        __pipeline_wait_prior(2);                         // slab i landed (this thread's copies)
        __syncthreads();
        float* st = stage_ptr(s);
        for (int idx = tid; idx < items; idx += NT) {     // lo images
            const int c = idx / (a.Ma + nb1), f = idx - c * (a.Ma + nb1);
            float* hi = f < a.Ma ? st + c * WG_ACS + f * 4 : st + 2 * WG_A_FLOATS + c * WG_BCS + (f - a.Ma) * 4;
            float* lo = hi + (f < a.Ma ? WG_A_FLOATS : WG_B_FLOATS);
            const float4 v = *reinterpret_cast<const float4*>(hi);
            *reinterpret_cast<float4*>(lo) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == DEC_ISSUER_WARP) {
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dAh = make_desc(smem_u32(st), WG_ACS * 4, 128), dAl = make_desc(smem_u32(st + WG_A_FLOATS), WG_ACS * 4, 128);
                const uint64_t dBh = make_desc(smem_u32(st + 2 * WG_A_FLOATS), WG_BCS * 4, 128),
                               dBl = make_desc(smem_u32(st + 2 * WG_A_FLOATS + WG_B_FLOATS), WG_BCS * 4, 128);
                constexpr uint64_t sa = (2 * WG_ACS * 4) >> 4, sb = (2 * WG_BCS * 4) >> 4;
#pragma unroll
                for (int ks = 0; ks < SLAB / 8; ++ks) {
                    mma_tf32_ss(tmem, dAl + ks * sa, dBh + ks * sb, idesc, (i > 0 || ks > 0) ? 1u : 0u);
                    mma_tf32_ss(tmem, dAh + ks * sa, dBl + ks * sb, idesc, 1);
                    mma_tf32_ss(tmem, dAh + ks * sa, dBh + ks * sb, idesc, 1);
                }
                mma_commit(&free_bar[s]);
                if (i == mine - 1) mma_commit(&done_bar);
            }
            __syncwarp();
        }
    }
    __pipeline_wait_prior(0);
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    {   // accumulator row m = TMEM lane; this thread's quarter of the columns
        const int q = warp & 3, cgp = warp >> 2;
        const int m = 32 * q + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
        const int nbc = a.Nb / 4;
        const int per = (nbc + 3) / 4 * 4;                // columns per group, multiple of 4
        for (int c = cgp * per; c < min(a.Nb, (cgp + 1) * per); c += 4) {
            float v[4];
            tmem_ld4(lane_addr + c, v);
            if (m < a.Ma) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = c + j;
                    if (n < a.Kin) gp[a.W_off + m * a.Kin + n] = v[j];
                    else if (n == a.Kin) gp[a.b_off + m] = v[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
    }
}

static size_t dec_fwd_tc_smem(int D) {
    const int N6 = (D + 15) & ~15;
    return (size_t)2 * (F4_C * F4_N * 4 + F5_C * F5_N * 4 + F6_C * N6 * 4) * sizeof(float) + 128;
}
static size_t dec_bwd_tc_smem() {
    return (size_t)2 * (X6_C * X6_N * 4 + X5_C * X5_N * 4 + X4_C * X4_N * 4) * sizeof(float) + 128;
}

}  // namespace tc

bool dec_tc_supported(const Layout& L) { return L.D % 4 == 0 && L.D >= 4 && L.D <= 104; }

template <typename Kern, typename Args>
static int tc_launch(Kern kern, const Args& args, size_t sm, int grid, cudaStream_t st, const char* name) {
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "%s: shared memory %zu B exceeds %d", name, sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    kern<<<grid, NT, sm, st>>>(args);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: launch: %s", name, cudaGetErrorString(e));
    return PCVAE_OK;
}

int dec_tc_launch(const DecArgs& a, int grid, cudaStream_t st) {
    const long R2 = (long)a.nbr * a.B;
    if (a.R2P > R2) {   // zero the padding columns [R2, R2P) of every feature row (read by the weight-gradient slabs)
        cudaError_t e = cudaMemset2DAsync(a.ws_zT + R2, a.R2P * sizeof(float), 0, (a.R2P - R2) * sizeof(float), TCW_FEATS, st);
        if (e != cudaSuccess) return fail(PCVAE_ECUDA, "dec_tc: cudaMemset2DAsync: %s", cudaGetErrorString(e));
    }
    if (int rc = tc_launch(tc::k_dec_fwd_tc, a, tc::dec_fwd_tc_smem(a.L.D), grid, st, "dec_fwd_tc")) return rc;
    if (int rc = tc_launch(tc::k_dec_bwd_tc, a, tc::dec_bwd_tc_smem(), grid, st, "dec_bwd_tc")) return rc;
    const size_t wsm = (size_t)tc::WG_STAGES * tc::WG_STAGE_FLOATS * sizeof(float) + 128;
    tc::WgradArgs w6{a.ws_dp6T, a.L.D, a.ws_h5T, G2, 112, a.R2P, a.gp, a.L.total, a.L.W6, a.L.b6};
    tc::WgradArgs w5{a.ws_dp5T, G2, a.ws_h4T, G1, 64, a.R2P, a.gp, a.L.total, a.L.W5, a.L.b5};
    tc::WgradArgs w4{a.ws_dp4T, G1, a.ws_zT, LAT, 16, a.R2P, a.gp, a.L.total, a.L.W4, a.L.b4};
    if (int rc = tc_launch(tc::k_wgrad_tc, w6, wsm, grid, st, "wgrad6")) return rc;
    if (int rc = tc_launch(tc::k_wgrad_tc, w5, wsm, grid, st, "wgrad5")) return rc;
    return tc_launch(tc::k_wgrad_tc, w4, wsm, grid, st, "wgrad4");
}

}  // namespace pcvae
