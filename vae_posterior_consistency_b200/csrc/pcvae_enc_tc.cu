// tcgen05 (5th-gen tensor core) version of the zero-impute MLP encoder and its backward (k_enc_fwd / k_enc_bwd in
// pcvae_train.cu, family MLP): same mathematics (Reg_VAE.encoder / vanilla_VAE.encoder, src/models/VAE.py:387-395,
// 1155-1163, and the autograd of train.py:115), 128-row tiles, every dense product on the tensor cores in
// fp32-accurate 3xTF32 (see pcvae_dec_tc.cu for the scheme; the scaffolding is shared through pcvae_tc_tile.cuh).
//
//   k_enc_fwd_tc   E1: [128 x round8(D+1)] x*mask|1 x W1aug^T -> 112     E2: [128 x 104] h1|1 x W2aug^T -> 64
//                  E3: [128 x 56] h2|1 x W3aug^T -> 32 (mean | logvar);   z = mean + eps * exp(logvar / 2)
//   k_enc_bwd_tc   X3: [128 x 24] (d_mean|d_logvar) x W3 -> 64    X2: [128 x 56] dpre2 x W2 -> 112
//                  (no data gradient for the first layer: x is an input)
//   k_wgrad_tc     dWaug for layers 1, 2, 3 from the feature-major scratch (pcvae_wgrad_tc.cu, one launch)
//
// The backward kernel needs 240 TMEM columns and 62 KB of shared memory, so two of its CTAs share an SM.
#include <cstdlib>

#include "pcvae_tc_tile.cuh"
#include "pcvae_train.cuh"

namespace pcvae {
namespace tc {

// TMEM map of the backward kernel (256 columns): d3 hi [0,24) lo [32,56); acc of X3 [64,128); dpre2 hi [128,184)
// lo [184,240); acc of X2 [0,112) -- it aliases d3 and the X3 accumulator, both dead by then
constexpr int B_D3H = 0, B_D3L = 32, B_AC2 = 64, B_RBH = 128, B_RBL = 184, B_AC1 = 0, B_COLS = 256;

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// weight images of k_enc_bwd_tc at `base`: W3^T hi | lo, W2^T hi | lo
__device__ __forceinline__ void enc_bwd_images(float* base, const float* __restrict__ th, const Layout& L, int tid) {
    float* T3h = base;
    float* T3l = T3h + Y3_C * Y3_N * 4;
    float* T2h = T3l + Y3_C * Y3_N * 4;
    float* T2l = T2h + Y2_C * Y2_N * 4;
    zero_images(base, 2 * (Y3_C * Y3_N * 4 + Y2_C * Y2_N * 4), tid);
    __syncthreads();
    image_linear_T(T3h, T3l, Y3_N, th + L.W3, LAT2, H2, tid);
    image_linear_T(T2h, T2l, Y2_N, th + L.W2, H2, H1, tid);
}

// weight images of k_enc_fwd_tc at `base` (shared memory): W1 hi | lo, W2 hi | lo, W3 hi | lo
__device__ __forceinline__ void enc_fwd_images(float* base, const float* __restrict__ th, const Layout& L, int tid) {
    const int D = L.D, K1 = (D + 8) & ~7, C1 = K1 / 4;
    float* W1h = base;
    float* W1l = W1h + C1 * E1_N * 4;
    float* W2h = W1l + C1 * E1_N * 4;
    float* W2l = W2h + E2_C * E2_N * 4;
    float* W3h = W2l + E2_C * E2_N * 4;
    float* W3l = W3h + E3_C * E3_N * 4;
    zero_images(base, 2 * (C1 * E1_N * 4 + E2_C * E2_N * 4 + E3_C * E3_N * 4), tid);
    __syncthreads();
    image_linear(W1h, W1l, E1_N, th + L.W1, th + L.b1, H1, D, true, tid);        // constant-1 output -> bias column of layer 2
    image_linear(W2h, W2l, E2_N, th + L.W2, th + L.b2, H2, H1, true, tid);       // constant-1 output -> bias column of layer 3
    // mean rows -> accumulator columns [0, 10), logvar rows -> [16, 26): both start on a 4-column boundary, so the latent
    // epilogue can be split over the column groups with aligned 4-column TMEM loads
    image_linear(W3h, W3l, E3_N, th + L.W3, th + L.b3, LAT, H2, false, tid);
    image_linear(W3h + 16 * 4, W3l + 16 * 4, E3_N, th + L.W3 + LAT * H2, th + L.b3 + LAT, LAT, H2, false, tid);
}
__global__ void __launch_bounds__(NT, 1) k_enc_fwd_tc(const EncFwdArgs a, const int stage_inputs_mode) {
    const int stage_inputs = stage_inputs_mode & 1;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar_s;
    __shared__ __align__(8) uint64_t in_bar;
    __shared__ __align__(8) uint64_t desc_s[6];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int D = a.L.D, K1 = (D + 8) & ~7, C1 = K1 / 4;
    float* W1h = smem;
    float* W1l = W1h + C1 * E1_N * 4;
    float* W2h = W1l + C1 * E1_N * 4;
    float* W2l = W2h + E2_C * E2_N * 4;
    float* W3h = W2l + E2_C * E2_N * 4;
    float* W3l = W3h + E3_C * E3_N * 4;
    // input staging (uint8 masks): the 128 x D tile of x and of the item's mask, fetched one item ahead with two bulk
    // async copies, so that stage 0 reads shared memory instead of issuing 14 uncoalesced global loads per thread
    float* xin = W3l + E3_C * E3_N * 4;
    const uint8_t* min_ = reinterpret_cast<const uint8_t*>(xin + ROWS * D);
    const float* th = a.theta;
    const Layout L = a.L;
    __shared__ __align__(8) uint64_t img_bar;
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&in_bar)), "r"(1));
    if (a.wimg) fetch_images(smem, a.wimg, (uint32_t)enc_fwd_image_floats(D) * 4, &img_bar, tid, a.tw.status);
    else enc_fwd_images(smem, th, L, tid);
    TileCtx cx;
    tc_setup(cx, &bar_s, &tmem_slot, tid, a.tw.status);
    const uint32_t tmem = cx.tmem, lane_addr = cx.lane_addr;
    const int cg = cx.cg, row = cx.row, c28 = cx.c28, c16 = cx.c16;

    const uint32_t cs1 = E1_N * 16, cs2 = E2_N * 16, cs3 = E3_N * 16;   // chunk strides (LBO); 8-row groups are 128 B apart (SBO)
    const uint64_t e1h = make_desc(smem_u32(W1h), cs1, 128), e1l = make_desc(smem_u32(W1l), cs1, 128);
    const uint64_t e2h = make_desc(smem_u32(W2h), cs2, 128), e2l = make_desc(smem_u32(W2l), cs2, 128);
    const uint64_t e3h = make_desc(smem_u32(W3h), cs3, 128), e3l = make_desc(smem_u32(W3l), cs3, 128);
    if (tid == 0) { st_desc(&desc_s[0], e1h); st_desc(&desc_s[1], e1l); st_desc(&desc_s[2], e2h); st_desc(&desc_s[3], e2l); st_desc(&desc_s[4], e3h); st_desc(&desc_s[5], e3l); }
    __syncthreads();
    // opaque run-time copies (loaded back with volatile loads): nothing for ptxas to fold or hoist
    const uint64_t r_e1h = ld_desc(&desc_s[0]);
    const uint64_t r_e1l = ld_desc(&desc_s[1]);
    const uint64_t r_e2h = ld_desc(&desc_s[2]);
    const uint64_t r_e2l = ld_desc(&desc_s[3]);
    const uint64_t r_e3h = ld_desc(&desc_s[4]);
    const uint64_t r_e3l = ld_desc(&desc_s[5]);
    const uint64_t es1 = (2 * cs1) >> 4, es2 = (2 * cs2) >> 4, es3 = (2 * cs3) >> 4;
    const uint32_t idE1 = make_idesc(ROWS, E1_N), idE2 = make_idesc(ROWS, E2_N), idE3 = make_idesc(ROWS, E3_N);
    const EncTcWs tw = a.tw;
    const bool save = tw.inT != nullptr;

    const int ntiles = (a.B + ROWS - 1) / ROWS;
    const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
    uint32_t in_ph = 0;
    // request the inputs of item (ti, bi); full tiles only (a ragged last tile takes the direct global loads)
    auto issue_in = [&](int ti, int bi) -> bool {
        if (!stage_inputs || ti >= ntiles || (ti + 1) * ROWS > a.B) return false;
        if (tid == 0) {
            const uint32_t xb = (uint32_t)(ROWS * D * 4), mb = (uint32_t)(ROWS * D);
            mbar_expect_tx(&in_bar, xb + mb);
            bulk_g2s(xin, a.x + bi * a.x_bs + (long)ti * ROWS * D, xb, &in_bar);
            bulk_g2s(reinterpret_cast<float*>(const_cast<uint8_t*>(min_)),
                     reinterpret_cast<const float*>(static_cast<const uint8_t*>(bi ? a.mask[1] : a.mask[0]) + (long)ti * ROWS * D), mb, &in_bar);
        }
        return true;
    };
    // The j-th work item of this CTA.  Flat order (default, bit 1 of `mode`): (tile, branch) pairs, tile-major, strided
    // over the CTAs as in k_dec_fwd_tc (7 instead of 8 items on the busiest CTA at 1024 items).  Nested order
    // (PCVAE_ENC_FLAT=0, kept as a cross-check): tiles strided over the CTAs, both branches of a tile back to back.
    const bool flat = (stage_inputs_mode & 2) != 0;
    const int nitems = ntiles * a.nbr;
    auto item_of = [&](int j, int& t_, int& br_) -> bool {
        if (flat) {
            const int w = blockIdx.x + j * gridDim.x;
            if (w >= nitems) return false;
            t_ = a.nbr == 2 ? (w >> 1) : w;
            br_ = a.nbr == 2 ? (w & 1) : 0;
            return true;
        }
        const int tj = j / a.nbr;
        br_ = j - tj * a.nbr;
        t_ = blockIdx.x + tj * gridDim.x;
        return t_ < ntiles;
    };
    bool staged = false;
    {
        int t0, b0;
        if (item_of(0, t0, b0)) staged = issue_in(t0, b0);
    }
    for (int j = 0;; ++j) {
        int t, br, tn = 0, bn = 0;
        if (!item_of(j, t, br)) break;
        const bool has_next = item_of(j + 1, tn, bn);
        if (!has_next) pdl_trigger();                     // this CTA's last item: the next kernel may take the SM when it exits
        const int grow = t * ROWS + row;
        const bool ok = grow < a.B;
        if (has_next && (flat || bn == 0)) {   // pull the next tile of this CTA towards L2 while this one is processed
            const long r0 = (long)tn * ROWS, nrows = min((long)ROWS, (long)a.B - r0);
            if (a.x_bs == 0) prefetch_l2(a.x + r0 * D, nrows * D * 4, tid);
            for (int b = 0; b < a.nbr; ++b)
                if (!flat || b == bn) {
                    if (a.x_bs != 0) prefetch_l2(a.x + b * a.x_bs + r0 * D, nrows * D * 4, tid);
                    prefetch_l2((const char*)(b ? a.mask[1] : a.mask[0]) + r0 * D * msz, nrows * D * msz, tid);
                }
        }
        {
            const long vt = (long)br * ntiles + t;            // tile of the scratch: [vt][row / 32][feature][row % 32]
            float* inT = tw.inT + (long)vt * (ETW_IN * ROWS) + (row >> 5) * (32 * ETW_IN) + (row & 31);
            float* h1T = tw.h1T + (long)vt * (ETW_H1 * ROWS) + (row >> 5) * (32 * ETW_H1) + (row & 31);
            float* h2T = tw.h2T + (long)vt * (ETW_H2 * ROWS) + (row >> 5) * (32 * ETW_H2) + (row & 31);
            unsigned* reluT = tw.relu + (vt * ROWS + row) * 8;
            // ---- x * mask | 1 -> RA, HBM: all loads of the thread in flight before the first TMEM store ----
            if (staged) { mbar_wait(&in_bar, in_ph, a.tw.status, 3); in_ph ^= 1u; }
            {
                float xm[28];
#pragma unroll
                for (int g = 0; g < 7; ++g) {
                    const int c = c28 + 4 * g;
                    float xv[4] = {0.f, 0.f, 0.f, 0.f};
                    if (c < D) {
                        if (staged) {                          // full tile, inputs in shared memory
                            const float4 x4 = *reinterpret_cast<const float4*>(xin + row * D + c);
                            const uint32_t mw = *reinterpret_cast<const uint32_t*>(min_ + row * D + c);
                            xv[0] = x4.x * ((mw & 0xFFu) ? 1.f : 0.f); xv[1] = x4.y * ((mw & 0xFF00u) ? 1.f : 0.f);
                            xv[2] = x4.z * ((mw & 0xFF0000u) ? 1.f : 0.f); xv[3] = x4.w * ((mw & 0xFF000000u) ? 1.f : 0.f);
                        } else if (ok) {
                            const long gi = (long)grow * D + c;
                            const float4 x4 = *reinterpret_cast<const float4*>(a.x + br * a.x_bs + gi);
                            float m[4];
                            load_mask4<true>(br ? a.mask[1] : a.mask[0], gi, a.mask_kind, m);
                            xv[0] = x4.x * m[0]; xv[1] = x4.y * m[1]; xv[2] = x4.z * m[2]; xv[3] = x4.w * m[3];
                        }
                    } else if (c == D) {
                        xv[0] = 1.0f;                          // bias column
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) xm[4 * g + j] = xv[j];
                }
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                    float v[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < cnt) { v[j] = xm[j0 + j]; lo[j] = tf32_lo(v[j]); }
                    st_part(lane_addr + RA_HI + c28, part, v);
                    st_part(lane_addr + RA_LO + c28, part, lo);
                }
                mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RA_HI, tmem + RA_LO, r_e1h, r_e1l, es1, K1 / 8, idE1); });
                // The scratch copies for the weight-gradient kernel go out UNDER the MMA batch they do not feed (the values are
                // still in registers): issued in front of the barrier they cost 13 of this kernel's 67 us -- the LSU queue
                // throttles the warps on their way to the barrier.
                if (save) {
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                        scratch_store(inT, c28 + j0, D + 1, xm + j0, cnt);
                    }
                }
            }
            // every thread has read its inputs (barrier inside mma_kick): the staging buffer is free for the next item
            staged = has_next ? issue_in(tn, bn) : false;
            mma_wait(cx, &bar_s);

            // ---- h1 = relu(acc1) | 1 -> RA, HBM (scratch stores under the E2 MMAs) ----
            uint32_t m1 = 0;                                  // relu mask of this thread's 28 h1 columns
            float acc[28];
            tmem_ld28(lane_addr + ACC1 + c28, acc);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                float lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        if (acc[j0 + j] > 0.f) m1 |= 1u << (j0 + j); else acc[j0 + j] = 0.f;
                        lo[j] = tf32_lo(acc[j0 + j]);
                    }
                st_part(lane_addr + RA_HI + c28, part, acc + j0);
                st_part(lane_addr + RA_LO + c28, part, lo);
            }
            mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC2, tmem + RA_HI, tmem + RA_LO, r_e2h, r_e2l, es2, E2_C / 2, idE2); });
            if (save) {
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
                    scratch_store(h1T, c28 + j0, ETW_H1, acc + j0, cnt);
                }
            }
            mma_wait(cx, &bar_s);

            // ---- h2 = relu(acc2) | 1 -> RB, HBM (scratch stores under the E3 MMAs) ----
            uint32_t m2 = 0;                                  // relu mask of this thread's 16 h2 columns
            float h2v[16];
            {
                float lo[16];
                tmem_ld16(lane_addr + ACC2 + c16, h2v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (h2v[j] > 0.f) m2 |= 1u << j; else h2v[j] = 0.f;
                    lo[j] = tf32_lo(h2v[j]);
                }
                if (cg < 3) { tmem_st16(lane_addr + RB_HI + c16, h2v); tmem_st16(lane_addr + RB_LO + c16, lo); }
                else { tmem_st8(lane_addr + RB_HI + c16, h2v); tmem_st8(lane_addr + RB_LO + c16, lo); }
            }
            // the row's 10 latents are split over its column groups (4 + 4 + 2 + 0) so that no warp waits for one group
            // doing all of them; the noise of the thread's latents is requested before the E3 MMAs are waited for
            const int l0 = 4 * cg;
            const long gl = (long)grow * LAT + l0;
            float2 e01 = make_float2(0.f, 0.f), e23 = make_float2(0.f, 0.f);
            if (cg < 3 && ok && a.z[br] && a.eps[br]) {
                e01 = *reinterpret_cast<const float2*>(a.eps[br] + gl);
                if (cg < 2) e23 = *reinterpret_cast<const float2*>(a.eps[br] + gl + 2);
            }
            mma_kick(&bar_s, warp, [&] { issue_3x(tmem + ACC1, tmem + RB_HI, tmem + RB_LO, r_e3h, r_e3l, es3, E3_C / 2, idE3); });
            if (save) {
                scratch_store(h2T, c16, ETW_H2, h2v, 16);
                reluT[cg] = m1;
                reluT[4 + cg] = m2;
            }
            mma_wait(cx, &bar_s);

            // ---- mean | logvar, reparameterisation ----
            if (cg < 3) {
                float mv[4], lv[4];
                tmem_ld4(lane_addr + ACC1 + l0, mv);
                tmem_ld4(lane_addr + ACC1 + 16 + l0, lv);
                if (ok) {
                    *reinterpret_cast<float2*>(a.mean[br] + gl) = make_float2(mv[0], mv[1]);
                    *reinterpret_cast<float2*>(a.logvar[br] + gl) = make_float2(lv[0], lv[1]);
                    if (cg < 2) {
                        *reinterpret_cast<float2*>(a.mean[br] + gl + 2) = make_float2(mv[2], mv[3]);
                        *reinterpret_cast<float2*>(a.logvar[br] + gl + 2) = make_float2(lv[2], lv[3]);
                    }
                    if (a.z[br]) {
                        float zz[4] = {mv[0], mv[1], mv[2], mv[3]};
                        if (a.eps[br]) {
                            zz[0] = fmaf(e01.x, expf(lv[0] * 0.5f), zz[0]);
                            zz[1] = fmaf(e01.y, expf(lv[1] * 0.5f), zz[1]);
                            if (cg < 2) {
                                zz[2] = fmaf(e23.x, expf(lv[2] * 0.5f), zz[2]);
                                zz[3] = fmaf(e23.y, expf(lv[3] * 0.5f), zz[3]);
                            }
                        }
                        *reinterpret_cast<float2*>(a.z[br] + gl) = make_float2(zz[0], zz[1]);
                        if (cg < 2) *reinterpret_cast<float2*>(a.z[br] + gl + 2) = make_float2(zz[2], zz[3]);
                    }
                }
            }
            tc_fence_before();      // the next branch overwrites RA / ACC2 only after its own barrier
        }
    }
    tc_teardown(cx, tid);
}

// ------------------------------------------------------------------------------------------------
// backward: data gradients down to the first layer's pre-activation
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2) k_enc_bwd_tc(const EncBwdArgs a) {
    __shared__ __align__(8) uint64_t img_bar;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar_s;
    __shared__ __align__(8) uint64_t desc_s[6];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* T3h = smem;
    float* T3l = T3h + Y3_C * Y3_N * 4;
    float* T2h = T3l + Y3_C * Y3_N * 4;
    float* T2l = T2h + Y2_C * Y2_N * 4;
    const float* th = a.theta;
    const Layout L = a.L;
    if (a.wimg) fetch_images(smem, a.wimg, (uint32_t)enc_bwd_image_floats() * 4, &img_bar, tid, a.tw.status);
    else enc_bwd_images(smem, th, L, tid);
    TileCtx cx;
    tc_setup(cx, &bar_s, &tmem_slot, tid, a.tw.status, B_COLS);
    const uint32_t tmem = cx.tmem, lane_addr = cx.lane_addr;
    const int cg = cx.cg, row = cx.row, c28 = cx.c28, c16 = cx.c16;

    const uint32_t cs3 = Y3_N * 16, cs2 = Y2_N * 16;
    const uint64_t y3h = make_desc(smem_u32(T3h), cs3, 128), y3l = make_desc(smem_u32(T3l), cs3, 128);
    const uint64_t y2h = make_desc(smem_u32(T2h), cs2, 128), y2l = make_desc(smem_u32(T2l), cs2, 128);
    if (tid == 0) { st_desc(&desc_s[0], y3h); st_desc(&desc_s[1], y3l); st_desc(&desc_s[2], y2h); st_desc(&desc_s[3], y2l); }
    __syncthreads();
    // opaque run-time copies (loaded back with volatile loads): nothing for ptxas to fold or hoist
    const uint64_t r_y3h = ld_desc(&desc_s[0]);
    const uint64_t r_y3l = ld_desc(&desc_s[1]);
    const uint64_t r_y2h = ld_desc(&desc_s[2]);
    const uint64_t r_y2l = ld_desc(&desc_s[3]);
    const uint64_t ys3 = (2 * cs3) >> 4, ys2 = (2 * cs2) >> 4;
    const uint32_t idY3 = make_idesc(ROWS, Y3_N), idY2 = make_idesc(ROWS, Y2_N);
    const EncTcWs tw = a.tw;

    const int nvt = ((a.B + ROWS - 1) / ROWS) * a.nbr;
    const int ntiles = (a.B + ROWS - 1) / ROWS;
    pdl_wait();                                           // d_mean / d_logvar and the ReLU masks come from the kernels before
    for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x) {
        if (vt + gridDim.x >= nvt) pdl_trigger();         // this CTA's last tile
        const int br = vt / ntiles, t = vt - br * ntiles;
        const int grow = t * ROWS + row;
        const bool ok = grow < a.B;
        float* dp1T = tw.dp1T + (long)vt * (ETW_H1 * ROWS) + (row >> 5) * (32 * ETW_H1) + (row & 31);      // scratch tile vt: [row / 32][feature][row % 32]
        float* dp2T = tw.dp2T + (long)vt * (ETW_H2 * ROWS) + (row >> 5) * (32 * ETW_H2) + (row & 31);
        float* dp3T = tw.dp3T + (long)vt * (ETW_DP3 * ROWS) + (row >> 5) * (32 * ETW_DP3) + (row & 31);
        const unsigned* reluT = tw.relu + ((long)vt * ROWS + row) * 8;
        uint32_t m1 = 0, m2 = 0;
        if (ok) { m1 = reluT[cg]; m2 = reluT[4 + cg]; }
        // ---- dpre3 = (d_mean | d_logvar) -> TMEM, HBM (the scratch stores of a stage go out under its MMA batch) ----
        float v3[24];
        if (cg == 0) {
            float lo[24];
#pragma unroll
            for (int j = 0; j < 24; ++j) v3[j] = 0.f;
            if (ok) {
                const long gi = (long)grow * LAT;
                const float2* dm = reinterpret_cast<const float2*>(a.d_mean[br] + gi);
                const float2* dv = reinterpret_cast<const float2*>(a.d_logvar[br] + gi);
#pragma unroll
                for (int l = 0; l < LAT / 2; ++l) {
                    const float2 p = dm[l], q2 = dv[l];
                    v3[2 * l] = p.x; v3[2 * l + 1] = p.y; v3[LAT + 2 * l] = q2.x; v3[LAT + 2 * l + 1] = q2.y;
                }
                if (a.d_z[br]) {          // reparameterisation backward folded in (z = mean + eps * exp(logvar / 2))
#pragma unroll
                    for (int l = 0; l < LAT; ++l) {
                        const float dz = a.d_z[br][gi + l];
                        v3[l] += dz;
                        if (a.eps[br]) v3[LAT + l] = fmaf(dz * 0.5f * expf(a.logvar[br][gi + l] * 0.5f), a.eps[br][gi + l], v3[LAT + l]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 24; ++j) lo[j] = tf32_lo(v3[j]);
            tmem_st16(lane_addr + B_D3H, v3);
            tmem_st8(lane_addr + B_D3H + 16, v3 + 16);
            tmem_st16(lane_addr + B_D3L, lo);
            tmem_st8(lane_addr + B_D3L + 16, lo + 16);
        }
        mma_kick(&bar_s, warp, [&] { issue_3x(tmem + B_AC2, tmem + B_D3H, tmem + B_D3L, r_y3h, r_y3l, ys3, Y3_C / 2, idY3); });
        if (cg == 0) {
#pragma unroll
            for (int j = 0; j < LAT2; ++j) dp3T[j * 32] = v3[j];
        }
        mma_wait(cx, &bar_s);

        // ---- dpre2 = dh2 * relu'(h2) -> TMEM, HBM ----
        {
            float v[16], lo[16];
            tmem_ld16(lane_addr + B_AC2 + c16, v);
            const uint32_t k2 = m2 & col_bits(c16, H2);                  // column H2 is the bias column
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (!(k2 & (1u << j))) v[j] = 0.f;
                lo[j] = tf32_lo(v[j]);
            }
            if (cg < 3) { tmem_st16(lane_addr + B_RBH + c16, v); tmem_st16(lane_addr + B_RBL + c16, lo); }
            else { tmem_st8(lane_addr + B_RBH + c16, v); tmem_st8(lane_addr + B_RBL + c16, lo); }
            mma_kick(&bar_s, warp, [&] { issue_3x(tmem + B_AC1, tmem + B_RBH, tmem + B_RBL, r_y2h, r_y2l, ys2, Y2_C / 2, idY2); });
            scratch_store(dp2T, c16, ETW_H2, v, 16);
        }
        mma_wait(cx, &bar_s);

        // ---- dpre1 = dh1 * relu'(h1) -> HBM ----
        const uint32_t k1 = m1 & col_bits(c28, H1);                      // column H1 is the bias column
        float acc[28];
        tmem_ld28(lane_addr + B_AC1 + c28, acc);
#pragma unroll
        for (int part = 0; part < 3; ++part) {
            const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < cnt) v[j] = acc[j0 + j];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < cnt && !(k1 & (1u << (j0 + j)))) v[j] = 0.f;
            scratch_store(dp1T, c28 + j0, ETW_H1, v, cnt);
        }
        tc_fence_before();
        __syncthreads();            // the X2 accumulator aliases the columns the next tile writes first
        tc_fence_after();
    }
    tc_teardown(cx, tid, B_COLS);
}

static size_t enc_fwd_tc_smem(int D, bool stage_inputs) {
    const int C1 = ((D + 8) & ~7) / 4;
    return (size_t)2 * (C1 * E1_N * 4 + E2_C * E2_N * 4 + E3_C * E3_N * 4) * sizeof(float) + 128 +
           (stage_inputs ? (size_t)ROWS * D * 5 : 0);
}
static size_t enc_bwd_tc_smem() { return (size_t)2 * (Y3_C * Y3_N * 4 + Y2_C * Y2_N * 4) * sizeof(float) + 128; }

}  // namespace tc

bool enc_tc_supported(const Layout& L) { return L.fam == PCVAE_FAMILY_MLP && !L.aug && L.D % 4 == 0 && L.D >= 4 && L.D <= 100; }

void enc_tc_carve(float* w, long rows, int nbr, EncTcWs* tw) {
    const long nvt = tc_nvt(rows, nbr), n = nvt * 128;
    tw->nvt = nvt;
    tw->status = tc_status_ptr();
    tw->inT = w;  w += n * ETW_IN;
    tw->h1T = w;  w += n * ETW_H1;
    tw->h2T = w;  w += n * ETW_H2;
    tw->dp1T = w; w += n * ETW_H1;
    tw->dp2T = w; w += n * ETW_H2;
    tw->dp3T = w; w += n * ETW_DP3;
    tw->relu = reinterpret_cast<unsigned*>(w);
}

template <typename Kern, typename Args>
static int enc_tc_go(Kern kern, const Args& args, size_t sm, int grid, cudaStream_t st, const char* name, bool dependent = false) {
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "%s: shared memory %zu B exceeds %d", name, sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    e = launch_tc(kern, grid, NT, sm, st, dependent, args);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: launch: %s", name, cudaGetErrorString(e));
    return PCVAE_OK;
}

int enc_fwd_tc_launch(const EncFwdArgs& a, int grid, cudaStream_t st) {
    // inputs are staged through shared memory when the masks are bytes and the tile fits beside the weight images
    const bool stage_inputs = a.mask_kind == PCVAE_MASK_U8 && tc::enc_fwd_tc_smem(a.L.D, true) <= (size_t)MAX_SMEM;
    const size_t sm = tc::enc_fwd_tc_smem(a.L.D, stage_inputs);
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "enc_fwd_tc: shared memory %zu B exceeds %d", sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(tc::k_enc_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "enc_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    prof_mark(st);
    static const int flat_items = [] { const char* e = getenv("PCVAE_ENC_FLAT"); return e && e[0] == '0' ? 0 : 2; }();
    tc::k_enc_fwd_tc<<<grid, NT, sm, st>>>(a, (stage_inputs ? 1 : 0) | flat_items);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "enc_fwd_tc: launch: %s", cudaGetErrorString(e));
    prof_mark(st);
    return PCVAE_OK;
}

int enc_bwd_tc_launch(const EncBwdArgs& a, int grid, cudaStream_t st) {
    prof_mark(st);
    if (a.B > 0)
        if (int rc = enc_tc_go(tc::k_enc_bwd_tc, a, tc::enc_bwd_tc_smem(), 2 * grid, st, "enc_bwd_tc", true)) return rc;
    prof_mark(st);
    const int D = a.L.D;
    const WgradJob jobs[3] = {{a.tw.dp1T, ETW_H1, H1, a.tw.inT, ETW_IN, D, (D + 16) & ~15, a.L.W1, a.L.b1},
                              {a.tw.dp2T, ETW_H2, H2, a.tw.h1T, ETW_H1, H1, 112, a.L.W2, a.L.b2},
                              {a.tw.dp3T, ETW_DP3, LAT2, a.tw.h2T, ETW_H2, H2, 64, a.L.W3, a.L.b3}};
    const int rc = wgrad_tc_launch(jobs, 3, a.tw.nvt, a.gp, a.L.total, grid, st);
    prof_mark(st);
    return rc;
}

}  // namespace pcvae
