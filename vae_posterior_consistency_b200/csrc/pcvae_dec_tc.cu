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
//   k_wgrad_tc     dWaug[m][n] = sum_rows dpre[row][m] * (act|1)[row][n]  for layers 6, 5, 4 (bias = the 1 column;
//                  pcvae_wgrad_tc.cu, one launch for the three layers)
//
// Biases ride along as an extra K column (the augmented weights hold b at column `in`, and one extra output
// row generates the constant-1 column of the next layer).  In the first two kernels the activation operand
// lives in TENSOR MEMORY (written by the epilogue threads with tcgen05.st, consumed by the MMA in its
// A-from-TMEM form), accumulators are in TMEM, and the weights sit in shared memory once per CTA as hi / lo
// images in the canonical K-major no-swizzle core-matrix layout [k/4][n][4].  Between the kernels the
// activations travel through HBM feature-major ([feature][row]: a warp's 32 rows make one 128-byte store), which
// is exactly the K-major operand layout of the weight-gradient GEMM (its reduction runs over rows); that
// kernel streams 32-row slabs with bulk async copies through a 4-stage ring (pcvae_wgrad_tc.cu) and keeps its
// accumulator in TMEM for a whole layer.  Its per-CTA partials land in the [grid][param_count] layout k_dec
// uses, so pcvae_reduce_grads is unchanged.
//
// TMEM columns (all 512), both row-tile kernels:  RA [0,224): 112 hi + 112 lo    RB [224,336): 56 hi + 56 lo
//                                                 ACC1 [336,448)                 ACC2 [448,512)
#include <cuda_pipeline.h>

#include "pcvae_tc_tile.cuh"
#include "pcvae_train.cuh"

namespace pcvae {
namespace tc {

// sigmoid with the two special-function instructions only: ex2.approx(-v log2 e) and rcp.approx(1 + t).  1 + t lies in
// [1, inf], so no denormal handling is needed; relative error ~2^-21 (the loss tolerance is 1e-4 relative)
__device__ __forceinline__ float sigmoid_fast(float v) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return r;
}

// weight images of k_dec_fwd_tc at `base`: W4 hi | lo, W5 hi | lo, W6 hi | lo
__device__ __forceinline__ void dec_fwd_images(float* base, const float* __restrict__ th, const Layout& L, int tid) {
    const int D = L.D, N6 = (D + 15) & ~15;
    float* W4h = base;
    float* W4l = W4h + F4_C * F4_N * 4;
    float* W5h = W4l + F4_C * F4_N * 4;
    float* W5l = W5h + F5_C * F5_N * 4;
    float* W6h = W5l + F5_C * F5_N * 4;
    float* W6l = W6h + F6_C * N6 * 4;
    zero_images(base, 2 * (F4_C * F4_N * 4 + F5_C * F5_N * 4 + F6_C * N6 * 4), tid);
    __syncthreads();
    image_linear(W4h, W4l, F4_N, th + L.W4, th + L.b4, G1, LAT, true, tid);      // constant-1 output -> bias column of layer 5
    image_linear(W5h, W5l, F5_N, th + L.W5, th + L.b5, G2, G1, true, tid);       // constant-1 output -> bias column of layer 6
    image_linear(W6h, W6l, N6, th + L.W6, th + L.b6, D, G2, false, tid);
}
// weight images of k_dec_bwd_tc at `base`: W6^T hi | lo, W5^T hi | lo, W4^T hi | lo
__device__ __forceinline__ void dec_bwd_images(float* base, const float* __restrict__ th, const Layout& L, int tid) {
    float* T6h = base;
    float* T6l = T6h + X6_C * X6_N * 4;
    float* T5h = T6l + X6_C * X6_N * 4;
    float* T5l = T5h + X5_C * X5_N * 4;
    float* T4h = T5l + X5_C * X5_N * 4;
    float* T4l = T4h + X4_C * X4_N * 4;
    zero_images(base, 2 * (X6_C * X6_N * 4 + X5_C * X5_N * 4 + X4_C * X4_N * 4), tid);
    __syncthreads();
    image_linear_T(T6h, T6l, X6_N, th + L.W6, L.D, G2, tid);
    image_linear_T(T5h, T5l, X5_N, th + L.W5, G2, G1, tid);
    image_linear_T(T4h, T4l, X4_N, th + L.W4, G1, LAT, tid);
}

// ------------------------------------------------------------------------------------------------
// forward + loss
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) k_dec_fwd_tc(const DecArgs a, const int stage_inputs) {
    __shared__ __align__(8) uint64_t img_bar;
    extern __shared__ __align__(128) float smem[];
    __shared__ float red_s[NWARP][PCVAE_NSUMS];
    __shared__ __align__(8) uint64_t bar_s;
    __shared__ __align__(8) uint64_t in_bar;
    __shared__ __align__(8) uint64_t desc_s[6];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D, N6 = (D + 15) & ~15;
    float* W4h = smem;
    float* W4l = W4h + F4_C * F4_N * 4;
    float* W5h = W4l + F4_C * F4_N * 4;
    float* W5l = W5h + F5_C * F5_N * 4;
    float* W6h = W5l + F5_C * F5_N * 4;
    float* W6l = W6h + F6_C * N6 * 4;
    // loss-input staging (uint8 masks): the item's 128 x D tile of x and of both masks, requested with bulk async copies
    // at the start of the item and read from shared memory by the loss epilogue
    float* xin = W6l + F6_C * N6 * 4;
    const uint8_t* m0s = reinterpret_cast<const uint8_t*>(xin + ROWS * D);
    const uint8_t* m1s = m0s + ROWS * D;
    const float* th = a.theta;
    const Layout L = a.L;
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&in_bar)), "r"(1));
    if (a.wimg_fwd) fetch_images(smem, a.wimg_fwd, (uint32_t)dec_fwd_image_floats(D) * 4, &img_bar, tid, a.status);
    else dec_fwd_images(smem, th, L, tid);
    TileCtx cx;
    tc_setup(cx, &bar_s, &tmem_slot, tid, a.status);
    const uint32_t tmem = cx.tmem, lane_addr = cx.lane_addr;
    const int cg = cx.cg, row = cx.row, c28 = cx.c28, c16 = cx.c16;

    const uint32_t cs4 = F4_N * 16, cs5 = F5_N * 16, cs6 = N6 * 16;     // chunk strides (LBO); 8-row groups are 128 B apart (SBO)
    const uint64_t f4h = make_desc(smem_u32(W4h), cs4, 128), f4l = make_desc(smem_u32(W4l), cs4, 128);
    const uint64_t f5h = make_desc(smem_u32(W5h), cs5, 128), f5l = make_desc(smem_u32(W5l), cs5, 128);
    const uint64_t f6h = make_desc(smem_u32(W6h), cs6, 128), f6l = make_desc(smem_u32(W6l), cs6, 128);
    if (tid == 0) { st_desc(&desc_s[0], f4h); st_desc(&desc_s[1], f4l); st_desc(&desc_s[2], f5h); st_desc(&desc_s[3], f5l); st_desc(&desc_s[4], f6h); st_desc(&desc_s[5], f6l); }
    __syncthreads();
    // opaque run-time copies (loaded back with volatile loads): nothing for ptxas to fold or hoist
    const uint64_t r_f4h = ld_desc(&desc_s[0]);
    const uint64_t r_f4l = ld_desc(&desc_s[1]);
    const uint64_t r_f5h = ld_desc(&desc_s[2]);
    const uint64_t r_f5l = ld_desc(&desc_s[3]);
    const uint64_t r_f6h = ld_desc(&desc_s[4]);
    const uint64_t r_f6l = ld_desc(&desc_s[5]);
    const uint64_t fs4 = (2 * cs4) >> 4, fs5 = (2 * cs5) >> 4, fs6 = (2 * cs6) >> 4;
    const uint32_t idF4 = make_idesc(ROWS, F4_N), idF5 = make_idesc(ROWS, F5_N), idF6 = make_idesc(ROWS, N6);

    // NLL constants exactly as k_dec / torch.distributions.Normal compute them
    const float scale = expf(a.x_logvar * 0.5f);
    const float var = scale * scale;
    const float inv2var = 1.0f / (2.0f * var);
    const float inv_var = 1.0f / var;
    const float log_scale = logf(scale);
    const float alpha = a.alpha, ls = a.loss_scale;
    // dL/dx_hat = coef * (x_hat - x) / var * loss_scale with coef = (1 - alpha) m + alpha m (1 - m_p) on the q branch and
    // alpha m_p on the p branch (VAE.py:432-445): three constants, selected by the mask bits
    const float coef_q_both = (1.f - alpha) * inv_var * ls, coef_q_only = ((1.f - alpha) + alpha) * inv_var * ls,
                coef_p = alpha * inv_var * ls;
    float s_req = 0.f, s_rep = 0.f, s_red = 0.f, s_imp = 0.f, s_sse = 0.f;

    const int ntiles = (a.B + ROWS - 1) / ROWS;
    const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
    // z of the item in hand (column group 0 only): loaded one item ahead, under the F6 MMAs of the previous item
    float2 zreg[LAT / 2];
    auto load_z = [&](int wi) {
        const int ti = a.nbr == 2 ? (wi >> 1) : wi, bi = a.nbr == 2 ? (wi & 1) : 0;
        const int gr = ti * ROWS + row;
#pragma unroll
        for (int j = 0; j < LAT / 2; ++j) zreg[j] = make_float2(0.f, 0.f);
        if (cg == 0 && wi < ntiles * a.nbr && gr < a.B) {
            const float2* zp = reinterpret_cast<const float2*>((bi ? a.z[1] : a.z[0]) + (long)gr * LAT);
#pragma unroll
            for (int j = 0; j < LAT / 2; ++j) zreg[j] = zp[j];
        }
    };
    pdl_wait();                                           // z, mean, logvar come from the encoder kernel before
    load_z(blockIdx.x);
    uint32_t in_ph = 0;
    auto issue_in = [&](int ti) -> bool {                // full tiles only (a ragged last tile takes the direct global loads)
        if (!stage_inputs || (ti + 1) * ROWS > a.B) return false;
        if (tid == 0) {
            const uint32_t xb = (uint32_t)(ROWS * D * 4), mb_ = (uint32_t)(ROWS * D);
            mbar_expect_tx(&in_bar, xb + a.nbr * mb_);
            bulk_g2s(xin, a.x + (long)ti * ROWS * D, xb, &in_bar);
            bulk_g2s(reinterpret_cast<float*>(const_cast<uint8_t*>(m0s)),
                     reinterpret_cast<const float*>(static_cast<const uint8_t*>(a.mask[0]) + (long)ti * ROWS * D), mb_, &in_bar);
            if (a.nbr > 1)
                bulk_g2s(reinterpret_cast<float*>(const_cast<uint8_t*>(m1s)),
                         reinterpret_cast<const float*>(static_cast<const uint8_t*>(a.mask[1]) + (long)ti * ROWS * D), mb_, &in_bar);
        }
        return true;
    };
    // work items = (tile, branch) pairs, tile-major, strided over the CTAs: the two branches of a tile run on
    // neighbouring CTAs at about the same time (x and the masks are shared through L2) and the load is balanced
    for (int w = blockIdx.x; w < ntiles * a.nbr; w += gridDim.x) {
        if (w + gridDim.x >= ntiles * a.nbr) pdl_trigger();   // this CTA's last item: the next kernel may take the SM when it exits
        const int t = a.nbr == 2 ? (w >> 1) : w, br = a.nbr == 2 ? (w & 1) : 0;
        const int row0 = t * ROWS;
        const int grow = row0 + row;
        const bool ok = grow < a.B;
        {   // pull the next item of this CTA towards L2 while this one is processed
            const int wn = w + gridDim.x;
            if (wn < ntiles * a.nbr) {
                const long r0 = (long)(a.nbr == 2 ? (wn >> 1) : wn) * ROWS, nrows = min((long)ROWS, (long)a.B - r0);
                prefetch_l2(a.x + r0 * D, nrows * D * 4, tid);
                for (int b = 0; b < a.nbr; ++b) prefetch_l2((const char*)a.mask[b] + r0 * D * msz, nrows * D * msz, tid);
            }
        }
        {
            const long vt = (long)br * ntiles + t;            // tile of the scratch: [vt][row / 32][feature][row % 32]
            float* zT = a.ws_zT + (long)vt * (TCW_Z * ROWS) + (row >> 5) * (32 * TCW_Z) + (row & 31);
            float* h4T = a.ws_h4T + (long)vt * (TCW_H4 * ROWS) + (row >> 5) * (32 * TCW_H4) + (row & 31);
            float* h5T = a.ws_h5T + (long)vt * (TCW_H5 * ROWS) + (row >> 5) * (32 * TCW_H5) + (row & 31);
            float* dp6T = a.ws_dp6T + (long)vt * (TCW_H5 * ROWS) + (row >> 5) * (32 * TCW_H5) + (row & 31);
            unsigned* reluT = a.ws_relu + (vt * ROWS + row) * 8;
            // ---- z | 1 -> RA, HBM ----
            // (the scratch copies of every stage go out under the MMA batch the stage starts: in front of the barrier the
            // LSU queue throttles the warps on their way to it)
            float zv[16];
            if (cg == 0) {
                float lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) zv[j] = 0.f;
#pragma unroll
                for (int j = 0; j < LAT / 2; ++j) { zv[2 * j] = zreg[j].x; zv[2 * j + 1] = zreg[j].y; }     // zero for rows past the batch
                zv[LAT] = 1.0f;
#pragma unroll
                for (int j = 0; j < 16; ++j) lo[j] = tf32_lo(zv[j]);
                tmem_st16(lane_addr + RA_HI, zv);
                tmem_st16(lane_addr + RA_HI + 16, lo);
            }
            mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC2, tmem + RA_HI, tmem + RA_HI + 16, r_f4h, r_f4l, fs4, F4_C / 2, idF4); });
            if (cg == 0) {
#pragma unroll
                for (int j = 0; j < TCW_Z; ++j) zT[j * 32] = zv[j];
            }
            // the barrier inside mma_kick ends the previous item's loss epilogue: the staging buffer is free
            const bool staged = issue_in(t);
            mma_wait(cx, &bar_s);

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
                mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RB_HI, tmem + RB_LO, r_f5h, r_f5l, fs5, F5_C / 2, idF5); });
                scratch_store(h4T, c16, TCW_H4, v, 16);
            }
            mma_wait(cx, &bar_s);

            // ---- h5 = relu(acc5) | 1 -> RA, HBM ----
            uint32_t m5 = 0;                                  // relu mask of this thread's 28 h5 columns
            float acc[28];
            tmem_ld28(lane_addr + ACC1 + c28, acc);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                float lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        if (acc[j0 + j] > 0.f) m5 |= 1u << (j0 + j); else acc[j0 + j] = 0.f;
                        lo[j] = tf32_lo(acc[j0 + j]);
                    }
                st_part(lane_addr + RA_HI + c28, part, acc + j0);
                st_part(lane_addr + RA_LO + c28, part, lo);
            }
            mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RA_HI, tmem + RA_LO, r_f6h, r_f6l, fs6, F6_C / 2, idF6); });
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                scratch_store(h5T, c28 + j0, TCW_H5, acc + j0, cnt);
            }
            reluT[cg] = m5;
            reluT[4 + cg] = m4;
            // while the tensor pipe runs F6: this thread's 28 entries of x and of the two masks (as bits)
            if (staged) { mbar_wait(&in_bar, in_ph, a.status, 3); in_ph ^= 1u; }
            float xr[28];
            uint32_t mb = 0, mpb = 0;
#pragma unroll
            for (int g = 0; g < 7; ++g) {
                const int c = c28 + 4 * g;
                float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (staged && c < D) {                         // full tile, inputs in shared memory
                    x4 = *reinterpret_cast<const float4*>(xin + row * D + c);
                    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(m0s + row * D + c);
                    const uint32_t w1 = a.nbr > 1 ? *reinterpret_cast<const uint32_t*>(m1s + row * D + c) : 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        mb |= (((w0 >> (8 * j)) & 0xFFu) ? 1u : 0u) << (4 * g + j);
                        mpb |= (((w1 >> (8 * j)) & 0xFFu) ? 1u : 0u) << (4 * g + j);
                    }
                } else if (ok && c < D) {
                    const long gi = (long)grow * D + c;
                    x4 = *reinterpret_cast<const float4*>(a.x + gi);
                    float m[4];
                    load_mask4(a.mask[0], gi, a.mask_kind, m);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mb |= (m[j] != 0.f ? 1u : 0u) << (4 * g + j);
                    if (a.nbr > 1) {
                        load_mask4(a.mask[1], gi, a.mask_kind, m);
#pragma unroll
                        for (int j = 0; j < 4; ++j) mpb |= (m[j] != 0.f ? 1u : 0u) << (4 * g + j);
                    }
                }
                xr[4 * g] = x4.x; xr[4 * g + 1] = x4.y; xr[4 * g + 2] = x4.z; xr[4 * g + 3] = x4.w;
            }
            load_z(w + gridDim.x);                            // the next item's z, also under the F6 MMAs
            mma_wait(cx, &bar_s);

            // ---- x_hat = sigmoid(acc6): loss terms, dL/d(pre-sigmoid) -> HBM ----
            // Branch-free per element: the item fixes a primary mask P (q branch: mask, p branch: mask_p), a secondary
            // mask S (q: mask_p, p: all ones) and two gradient weights; NLL sums are taken per item and folded into the
            // running sums of the item's branch afterwards.
            //   q: RE_q over P, RE_d over P & ~S, RE_q_imputed / SSE over ~P; dL/dx_hat weight (1 - alpha) on P & S, 1 on P & ~S
            //   p: RE_p over P;                                              dL/dx_hat weight alpha on P       (VAE.py:432-445)
            {
                float* xo = a.xhat[br];
                const uint32_t Pm = br == 0 ? mb : mpb, Sm = br == 0 ? mpb : 0xFFFFFFFFu;
                const float wA = br == 0 ? coef_q_both : coef_p, wB = br == 0 ? coef_q_only : coef_p;
                float n_p = 0.f, n_pns = 0.f, n_np = 0.f, q_np = 0.f;
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
                            float xh[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int e = j0 + 4 * g + j;
                                const bool pb = (Pm >> e) & 1u, sb = (Sm >> e) & 1u;
                                xh[j] = sigmoid_fast(v[4 * g + j]);
                                const float diff = xr[e] - xh[j];
                                const float d2 = diff * diff;
                                const float nll = fmaf(d2, inv2var, log_scale);
                                if (pb) n_p += nll;
                                if (pb && !sb) n_pns += nll;
                                if (!pb) { n_np += nll; q_np += d2; }
                                const float coef = pb ? (sb ? wA : wB) : 0.f;
                                dpre[j] = (coef * -diff) * (xh[j] * (1.f - xh[j]));
                            }
                            if (xo) *reinterpret_cast<float4*>(xo + (long)grow * D + c) = make_float4(xh[0], xh[1], xh[2], xh[3]);
                        }
                        if (c < TCW_H5) {
                            float* __restrict__ p6 = dp6T + (size_t)c * 32;
#pragma unroll
                            for (int j = 0; j < 4; ++j) p6[j * 32] = dpre[j];
                        }
                    }
                }
                if (br == 0) { s_req += n_p; s_red += n_pns; s_imp += n_np; s_sse += q_np; }
                else s_rep += n_p;
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
    __shared__ __align__(8) uint64_t img_bar;
    extern __shared__ __align__(128) float smem[];
    __shared__ float red_s[NWARP][3];
    __shared__ __align__(8) uint64_t bar_s;
    __shared__ __align__(8) uint64_t in_bar;
    __shared__ __align__(8) uint64_t desc_s[6];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D;
    float* T6h = smem;
    float* T6l = T6h + X6_C * X6_N * 4;
    float* T5h = T6l + X6_C * X6_N * 4;
    float* T5l = T5h + X5_C * X5_N * 4;
    float* T4h = T5l + X5_C * X5_N * 4;
    float* T4l = T4h + X4_C * X4_N * 4;
    // dpre6 staging: the item's [4 slabs][104 features][32 rows] block of the scratch (contiguous, written by
    // k_dec_fwd_tc) arrives by one bulk async copy requested an item ahead, so stage 0 reads shared memory instead of
    // waiting for 28 global loads per thread
    float* d6s = T4l + X4_C * X4_N * 4;
    const float* th = a.theta;
    const Layout L = a.L;
    const int ntiles = (a.B + ROWS - 1) / ROWS;
    uint32_t in_ph = 0;
    auto issue_in = [&](int wi) {
        if (wi < ntiles * a.nbr && tid == 0) {
            const int ti = a.nbr == 2 ? (wi >> 1) : wi, bi = a.nbr == 2 ? (wi & 1) : 0;
            const uint32_t bytes = (uint32_t)(TCW_H5 * ROWS * 4);
            mbar_expect_tx(&in_bar, bytes);
            bulk_g2s(d6s, a.ws_dp6T + ((long)bi * ntiles + ti) * (TCW_H5 * ROWS), bytes, &in_bar);
        }
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&in_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (a.wimg_bwd) fetch_images(smem, a.wimg_bwd, (uint32_t)dec_bwd_image_floats() * 4, &img_bar, tid, a.status);
    else dec_bwd_images(smem, th, L, tid);
    TileCtx cx;
    tc_setup(cx, &bar_s, &tmem_slot, tid, a.status);
    const uint32_t tmem = cx.tmem, lane_addr = cx.lane_addr;
    const int cg = cx.cg, row = cx.row, c28 = cx.c28, c16 = cx.c16;

    const uint32_t cs6 = X6_N * 16, cs5 = X5_N * 16, cs4 = X4_N * 16;
    const uint64_t x6h = make_desc(smem_u32(T6h), cs6, 128), x6l = make_desc(smem_u32(T6l), cs6, 128);
    const uint64_t x5h = make_desc(smem_u32(T5h), cs5, 128), x5l = make_desc(smem_u32(T5l), cs5, 128);
    const uint64_t x4h = make_desc(smem_u32(T4h), cs4, 128), x4l = make_desc(smem_u32(T4l), cs4, 128);
    if (tid == 0) { st_desc(&desc_s[0], x6h); st_desc(&desc_s[1], x6l); st_desc(&desc_s[2], x5h); st_desc(&desc_s[3], x5l); st_desc(&desc_s[4], x4h); st_desc(&desc_s[5], x4l); }
    __syncthreads();
    // opaque run-time copies (loaded back with volatile loads): nothing for ptxas to fold or hoist
    const uint64_t r_x6h = ld_desc(&desc_s[0]);
    const uint64_t r_x6l = ld_desc(&desc_s[1]);
    const uint64_t r_x5h = ld_desc(&desc_s[2]);
    const uint64_t r_x5l = ld_desc(&desc_s[3]);
    const uint64_t r_x4h = ld_desc(&desc_s[4]);
    const uint64_t r_x4l = ld_desc(&desc_s[5]);
    const uint64_t xs6 = (2 * cs6) >> 4, xs5 = (2 * cs5) >> 4, xs4 = (2 * cs4) >> 4;
    const uint32_t idX6 = make_idesc(ROWS, X6_N), idX5 = make_idesc(ROWS, X5_N), idX4 = make_idesc(ROWS, X4_N);

    const float alpha = a.alpha, ls = a.loss_scale;
    float s_klq = 0.f, s_klp = 0.f, s_klr = 0.f;

    pdl_wait();                                           // dpre6, the ReLU masks and the latent statistics come from the kernels before
    issue_in(blockIdx.x);
    for (int w = blockIdx.x; w < ntiles * a.nbr; w += gridDim.x) {      // (tile, branch) items as in k_dec_fwd_tc
        if (w + gridDim.x >= ntiles * a.nbr) pdl_trigger();   // this CTA's last item
        const int t = a.nbr == 2 ? (w >> 1) : w, br = a.nbr == 2 ? (w & 1) : 0;
        const int row0 = t * ROWS;
        const int grow = row0 + row;
        const bool ok = grow < a.B;
        {
            const long vt = (long)br * ntiles + t;
            float* dp5T = a.ws_dp5T + (long)vt * (TCW_H5 * ROWS) + (row >> 5) * (32 * TCW_H5) + (row & 31);
            float* dp4T = a.ws_dp4T + (long)vt * (TCW_H4 * ROWS) + (row >> 5) * (32 * TCW_H4) + (row & 31);
            const unsigned* reluT = a.ws_relu + (vt * ROWS + row) * 8;
            uint32_t m5 = 0, m4 = 0;
            if (ok) { m5 = reluT[cg]; m4 = reluT[4 + cg]; }
            // ---- dpre6 (staged in shared memory) -> RA ----
            mbar_wait(&in_bar, in_ph, a.status, 3);
            in_ph ^= 1u;
            {
                float d6[28];
                {
                    const float* __restrict__ p6 = d6s + (row >> 5) * (32 * TCW_H5) + (row & 31) + c28 * 32;
                    const int nv = ok ? min(28, TCW_H5 - c28) : 0;
#pragma unroll
                    for (int j = 0; j < 28; ++j) d6[j] = j < nv ? p6[j * 32] : 0.f;
                }
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                    float v[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < cnt) { v[j] = d6[j0 + j]; lo[j] = tf32_lo(v[j]); }
                    st_part(lane_addr + RA_HI + c28, part, v);
                    st_part(lane_addr + RA_LO + c28, part, lo);
                }
            }
            // latent-space inputs of this thread's (up to four) latents, see the last epilogue of the item: requested here
            // so that their latency hides under the three MMA stages instead of being exposed after the last one
            float lat_mq[4] = {0.f, 0.f, 0.f, 0.f}, lat_lq[4] = {0.f, 0.f, 0.f, 0.f}, lat_mp[4] = {0.f, 0.f, 0.f, 0.f},
                  lat_lp[4] = {0.f, 0.f, 0.f, 0.f}, lat_e[4] = {0.f, 0.f, 0.f, 0.f};
            if (cg < 3 && ok) {
                const long g0 = (long)grow * LAT + 4 * cg;
                auto ld4 = [&](const float* p, float* o) {
                    const float2 u = *reinterpret_cast<const float2*>(p + g0);
                    o[0] = u.x; o[1] = u.y;
                    if (cg < 2) { const float2 w2 = *reinterpret_cast<const float2*>(p + g0 + 2); o[2] = w2.x; o[3] = w2.y; }
                };
                ld4(a.mean[0], lat_mq);
                ld4(a.logvar[0], lat_lq);
                if (a.nbr > 1) { ld4(a.mean[1], lat_mp); ld4(a.logvar[1], lat_lp); }
                if (a.eps[br]) ld4(a.eps[br], lat_e);
            }
            mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RA_HI, tmem + RA_LO, r_x6h, r_x6l, xs6, X6_C / 2, idX6); });
            issue_in(w + gridDim.x);        // every thread has read the staging buffer (barrier inside mma_kick)
            mma_wait(cx, &bar_s);

            // ---- dpre5 = dh5 * relu'(h5) -> RA, HBM ----
            const uint32_t k5 = m5 & col_bits(c28, G2);                  // column G2 is the bias column
            float acc[28];
            tmem_ld28(lane_addr + ACC1 + c28, acc);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                float lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        if (!(k5 & (1u << (j0 + j)))) acc[j0 + j] = 0.f;
                        lo[j] = tf32_lo(acc[j0 + j]);
                    }
                st_part(lane_addr + RA_HI + c28, part, acc + j0);
                st_part(lane_addr + RA_LO + c28, part, lo);
            }
            mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC2, tmem + RA_HI, tmem + RA_LO, r_x5h, r_x5l, xs5, X5_C / 2, idX5); });
            // scratch copies under the MMA batch (see k_dec_fwd_tc)
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                scratch_store(dp5T, c28 + j0, TCW_H5, acc + j0, cnt);
            }
            mma_wait(cx, &bar_s);

            // ---- dpre4 = dh4 * relu'(h4) -> RB, HBM ----
            {
                float v[16], lo[16];
                tmem_ld16(lane_addr + ACC2 + c16, v);
                const uint32_t k4 = m4 & col_bits(c16, G1);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (!(k4 & (1u << j))) v[j] = 0.f;
                    lo[j] = tf32_lo(v[j]);
                }
                if (cg < 3) { tmem_st16(lane_addr + RB_HI + c16, v); tmem_st16(lane_addr + RB_LO + c16, lo); }
                else { tmem_st8(lane_addr + RB_HI + c16, v); tmem_st8(lane_addr + RB_LO + c16, lo); }
                mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC2, tmem + RB_HI, tmem + RB_LO, r_x4h, r_x4l, xs4, X4_C / 2, idX4); });
                scratch_store(dp4T, c16, TCW_H4, v, 16);
            }
            mma_wait(cx, &bar_s);

            // ---- latent-space terms: KL sums, d_mean / d_logvar; the row's 10 latents are split over its column
            //      groups (4 + 4 + 2 + 0) so that no warp waits for a single group doing all of them ----
            if (cg < 3) {
                float dzv[4];
                tmem_ld4(lane_addr + ACC2 + 4 * cg, dzv);
                if (ok) {
#pragma unroll
                    for (int li = 0; li < 4; ++li) {
                        const int l = 4 * cg + li;
                        if (l >= LAT) continue;
                        const long gi = (long)grow * LAT + l;
                        const float mq = lat_mq[li], lq = lat_lq[li];
                        const float eq = expf(lq);
                        float mp_ = 0.f, lp = 0.f, ep = 1.f;
                        if (a.nbr > 1) { mp_ = lat_mp[li]; lp = lat_lp[li]; ep = expf(lp); }
                        const float dmu = mq - mp_;
                        if (br == 0) {
                            s_klq += 0.5f * (eq + mq * mq - 1.f - lq);
                            if (a.nbr > 1) {
                                s_klp += 0.5f * (ep + mp_ * mp_ - 1.f - lp);
                                s_klr += 0.5f * (expf(lq - lp) + dmu * dmu / ep - 1.f - (lq - lp));
                            }
                        }
                        const float dz = dzv[li];
                        float gm, gv;
                        if (br == 0) {
                            gm = (1.f - alpha) * a.beta_w * mq;
                            gv = (1.f - alpha) * a.beta_w * 0.5f * (eq - 1.f);
                            if (a.nbr > 1) {
                                gm += alpha * dmu / ep;
                                gv += alpha * 0.5f * (expf(lq - lp) - 1.f);
                            }
                            gm = fmaf(gm, ls, dz);
                            gv = fmaf(gv, ls, dz * 0.5f * expf(lq * 0.5f) * lat_e[li]);
                        } else {
                            gm = alpha * a.beta_w * mp_ - alpha * dmu / ep;
                            gv = alpha * a.beta_w * 0.5f * (ep - 1.f) + alpha * 0.5f * (1.f - (eq + dmu * dmu) / ep);
                            gm = fmaf(gm, ls, dz);
                            gv = fmaf(gv, ls, dz * 0.5f * expf(lp * 0.5f) * lat_e[li]);
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

static size_t dec_fwd_tc_smem(int D, int nbr, bool stage_inputs) {
    const int N6 = (D + 15) & ~15;
    return (size_t)2 * (F4_C * F4_N * 4 + F5_C * F5_N * 4 + F6_C * N6 * 4) * sizeof(float) + 128 +
           (stage_inputs ? (size_t)ROWS * D * (4 + nbr) : 0);
}
static size_t dec_bwd_tc_smem() {
    return (size_t)2 * (X6_C * X6_N * 4 + X5_C * X5_N * 4 + X4_C * X4_N * 4) * sizeof(float) + 128 +
           (size_t)TCW_H5 * ROWS * sizeof(float);
}

}  // namespace tc

bool dec_tc_supported(const Layout& L) { return L.D % 4 == 0 && L.D >= 4 && L.D <= 104; }

template <typename Kern, typename Args>
static int tc_launch(Kern kern, const Args& args, size_t sm, int grid, cudaStream_t st, const char* name) {
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "%s: shared memory %zu B exceeds %d", name, sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    e = launch_tc(kern, grid, NT, sm, st, true, args);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: launch: %s", name, cudaGetErrorString(e));
    return PCVAE_OK;
}

int dec_tc_launch(const DecArgs& a_in, int grid, cudaStream_t st) {
    DecArgs a = a_in;
    a.status = tc_status_ptr();
    prof_mark(st);
    {   // loss inputs are staged through shared memory when the masks are bytes and the tile fits beside the weight images
        const bool stage_inputs = a.mask_kind == PCVAE_MASK_U8 && tc::dec_fwd_tc_smem(a.L.D, a.nbr, true) + 1024 <= (size_t)MAX_SMEM;
        const size_t sm = tc::dec_fwd_tc_smem(a.L.D, a.nbr, stage_inputs);
        if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "dec_fwd_tc: shared memory %zu B exceeds %d", sm, MAX_SMEM);
        cudaError_t e = cudaFuncSetAttribute(tc::k_dec_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return fail(PCVAE_ECUDA, "dec_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        e = launch_tc(tc::k_dec_fwd_tc, grid, NT, sm, st, true, a, stage_inputs ? 1 : 0);
        if (e != cudaSuccess) return fail(PCVAE_ECUDA, "dec_fwd_tc: launch: %s", cudaGetErrorString(e));
    }
    prof_mark(st);
    if (int rc = tc_launch(tc::k_dec_bwd_tc, a, tc::dec_bwd_tc_smem(), grid, st, "dec_bwd_tc")) return rc;
    prof_mark(st);
    const WgradJob jobs[3] = {{a.ws_dp6T, TCW_H5, a.L.D, a.ws_h5T, TCW_H5, G2, 112, a.L.W6, a.L.b6},
                              {a.ws_dp5T, TCW_H5, G2, a.ws_h4T, TCW_H4, G1, 64, a.L.W5, a.L.b5},
                              {a.ws_dp4T, TCW_H4, G1, a.ws_zT, TCW_Z, LAT, 16, a.L.W4, a.L.b4}};
    const int rc = wgrad_tc_launch(jobs, 3, a.nvt, a.gp, a.L.total, grid, st);
    prof_mark(st);
    return rc;
}

}  // namespace pcvae
