// k_enc_fwd_tc2: the tcgen05 encoder forward (pcvae_enc_tc.cu: k_enc_fwd_tc, same mathematics, same scratch layout) with
// TWO work items in flight per SM.  Round 1's kernel kept one 128-row item per CTA in tensor memory (operand hi + lo
// images of a 104-wide layer and its accumulator take 336 of the 512 columns), so every
// tcgen05.ld -> ALU -> tcgen05.st -> fence -> barrier -> MMA -> commit -> wait chain of an item was exposed and the tensor
// pipe idled 75 % of the time.  Here the CTA is two independent HALVES of 8 warps (4 TMEM lane quarters x 2 column
// groups); each half owns 256 TMEM columns, its own mbarrier and its own named barrier, takes its own (tile, branch)
// items and runs the same layer pipeline -- while one half is in a register epilogue the other half's MMAs run.  The
// weight images in shared memory are shared by both halves.  An item fits 256 columns because its operand reaches
// tensor memory in K-CHUNKS of 56 columns (7 k-steps): chunk a -> MMAs (accumulate = false) -> chunk b over the same
// columns -> MMAs (accumulate = true); the accumulator of layer 2 aliases the part of layer 1's accumulator that has
// already been read.
//
//   per-half TMEM columns:  OP hi [0,56)  OP lo [56,112)  ACC2 [112,176)  ACC1 [120,232)  ACC3 [176,208)
//   (ACC1 column j sits at 120 + j: ACC2 overlaps ACC1 columns 0..55 only, which every thread has read before the
//    layer-2 MMAs are issued; ACC1 columns 56..111 stay intact until they are read under those MMAs)
#include <cstdlib>

#include "pcvae_tc_tile.cuh"
#include "pcvae_train.cuh"

namespace pcvae {
namespace tc {

constexpr int F1_N = 112;                  // E1: 100 outputs + the constant-1 generator; K = round8(D + 1)
constexpr int F2_C = 26, F2_N = 64;        // E2: K = 104 (h1|1), 50 outputs + the constant-1 generator
constexpr int F3_C = 14, F3_N = 32;        // E3: K = 56 (h2|1), mean at columns 0..9, logvar at 16..25
constexpr int H_OPH = 0, H_OPL = 56, H_ACC2 = 112, H_ACC1 = 120, H_ACC3 = 176, H_COLS = 256;
constexpr int KCH = 7;                     // k-steps (of 8 columns) per operand chunk

__device__ __forceinline__ void half_sync(int h) { asm volatile("bar.sync %0, 256;" ::"r"(h + 1) : "memory"); }

// MMAs of k-steps [ks0, ks0 + nks) of a layer: operand chunk at columns 0.. of the OP region, weight image k-step ks0 + i
__device__ __forceinline__ void issue_3x_chunk(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint64_t b_step,
                                               int ks0, int nks, uint32_t idesc, bool accumulate) {
    for (int i = 0; i < nks; ++i) {
        const uint64_t o = (uint64_t)(ks0 + i) * b_step;
        mma_tf32_ts(acc, a_lo + 8 * i, b_hi + o, idesc, (i > 0 || accumulate) ? 1u : 0u);
        mma_tf32_ts(acc, a_hi + 8 * i, b_lo + o, idesc, 1);
        mma_tf32_ts(acc, a_hi + 8 * i, b_hi + o, idesc, 1);
    }
}

template <typename Issue>
__device__ __forceinline__ void half_kick(uint64_t* bar, int h, int wh, Issue&& issue) {
    tmem_st_wait();
    tc_fence_before();
    half_sync(h);
    if (wh == 0) {
        tc_fence_after();
        if (elect_one()) {
            issue();
            mma_commit(bar);
        }
        __syncwarp();
    }
}

struct HalfCtx { uint32_t ph; int* status; };
__device__ __forceinline__ void half_wait(HalfCtx& hx, uint64_t* bar) {
    mbar_wait(bar, hx.ph, hx.status, 2);
    hx.ph ^= 1u;
    tc_fence_after();
}

// 28 operand columns of this thread (hi / lo split) into the OP region at local column c0, parts of 16 / 8 / 4; the
// values also go to the feature-major scratch (features first .. of the row's slot, features >= limit skipped)
__device__ __forceinline__ void put28(uint32_t lane_addr, int c0, const float* v28, float* __restrict__ sc_col0, int first, int limit,
                                      bool save) {
#pragma unroll
    for (int part = 0; part < 3; ++part) {
        const int j0 = part == 0 ? 0 : (part == 1 ? 16 : 24), cnt = part == 0 ? 16 : (part == 1 ? 8 : 4);
        float v[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < cnt) { v[j] = v28[j0 + j]; lo[j] = tf32_lo(v[j]); }
        st_part(lane_addr + H_OPH + c0, part, v);
        st_part(lane_addr + H_OPL + c0, part, lo);
        if (save) scratch_store(sc_col0, first + j0, limit, v, cnt);
    }
}

__global__ void __launch_bounds__(NT, 1) k_enc_fwd_tc2(const EncFwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar_s[2];
    __shared__ __align__(8) uint64_t desc_s[6];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.L.D, K1 = (D + 8) & ~7, C1 = K1 / 4;
    float* W1h = smem;
    float* W1l = W1h + C1 * F1_N * 4;
    float* W2h = W1l + C1 * F1_N * 4;
    float* W2l = W2h + F2_C * F2_N * 4;
    float* W3h = W2l + F2_C * F2_N * 4;
    float* W3l = W3h + F3_C * F3_N * 4;
    const float* th = a.theta;
    const Layout L = a.L;
    zero_images(smem, 2 * (C1 * F1_N * 4 + F2_C * F2_N * 4 + F3_C * F3_N * 4), tid);
    __syncthreads();
    image_linear(W1h, W1l, F1_N, th + L.W1, th + L.b1, H1, D, true, tid);
    image_linear(W2h, W2l, F2_N, th + L.W2, th + L.b2, H2, H1, true, tid);
    image_linear(W3h, W3l, F3_N, th + L.W3, th + L.b3, LAT, H2, false, tid);
    image_linear(W3h + 16 * 4, W3l + 16 * 4, F3_N, th + L.W3 + LAT * H2, th + L.b3 + LAT, LAT, H2, false, tid);
    // tensor memory (all 512 columns, 256 per half) and one mbarrier per half
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_s[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_s[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const uint32_t cs1 = F1_N * 16, cs2 = F2_N * 16, cs3 = F3_N * 16;   // chunk strides (LBO); 8-row groups are 128 B apart (SBO)
    if (tid == 0) {
        st_desc(&desc_s[0], make_desc(smem_u32(W1h), cs1, 128)); st_desc(&desc_s[1], make_desc(smem_u32(W1l), cs1, 128));
        st_desc(&desc_s[2], make_desc(smem_u32(W2h), cs2, 128)); st_desc(&desc_s[3], make_desc(smem_u32(W2l), cs2, 128));
        st_desc(&desc_s[4], make_desc(smem_u32(W3h), cs3, 128)); st_desc(&desc_s[5], make_desc(smem_u32(W3l), cs3, 128));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // opaque run-time copies of the descriptors (see pcvae_tc_tile.cuh: st_desc / ld_desc)
    const uint64_t r_e1h = ld_desc(&desc_s[0]), r_e1l = ld_desc(&desc_s[1]);
    const uint64_t r_e2h = ld_desc(&desc_s[2]), r_e2l = ld_desc(&desc_s[3]);
    const uint64_t r_e3h = ld_desc(&desc_s[4]), r_e3l = ld_desc(&desc_s[5]);
    const uint64_t es1 = (2 * cs1) >> 4, es2 = (2 * cs2) >> 4, es3 = (2 * cs3) >> 4;
    const uint32_t idE1 = make_idesc(ROWS, F1_N), idE2 = make_idesc(ROWS, F2_N), idE3 = make_idesc(ROWS, F3_N);
    const EncTcWs tw = a.tw;
    const bool save = tw.inT != nullptr;

    const int h = warp >> 3, wh = warp & 7, q = wh & 3, cg = wh >> 2;
    const int row = 32 * q + lane, c28 = 28 * cg;
    const uint32_t tbase = tmem + (uint32_t)(H_COLS * h);
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
    uint64_t* bar = &bar_s[h];
    HalfCtx hx{0u, tw.status};
    const int ks1 = K1 / 8, ks1a = ks1 < KCH ? ks1 : KCH, ks1b = ks1 - ks1a;     // k-steps of layer 1: chunk a, chunk b
    const int ntiles = (a.B + ROWS - 1) / ROWS, nitems = ntiles * a.nbr;
    const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
    const int htid = tid & 255;

    for (int w = 2 * blockIdx.x + h; w < nitems; w += 2 * gridDim.x) {
        const int t = a.nbr == 2 ? (w >> 1) : w, br = a.nbr == 2 ? (w & 1) : 0;
        const int grow = t * ROWS + row;
        const bool ok = grow < a.B;
        {   // pull this half's next tile towards L2
            const int wn = w + 2 * gridDim.x;
            if (wn < nitems) {
                const int tn = a.nbr == 2 ? (wn >> 1) : wn, bn = a.nbr == 2 ? (wn & 1) : 0;
                const long r0 = (long)tn * ROWS, nrows = min((long)ROWS, (long)a.B - r0);
                const char* px = reinterpret_cast<const char*>(a.x + r0 * D);
                for (long off = (long)htid * 128; off < nrows * D * 4; off += 256 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + off));
                const char* pm = reinterpret_cast<const char*>(bn ? a.mask[1] : a.mask[0]) + r0 * D * msz;
                for (long off = (long)htid * 128; off < nrows * D * msz; off += 256 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm + off));
            }
        }
        const long vt = (long)br * ntiles + t;            // tile of the scratch: [vt][row / 32][feature][row % 32]
        float* inT = tw.inT + (long)vt * (ETW_IN * ROWS) + (row >> 5) * (32 * ETW_IN) + (row & 31);
        float* h1T = tw.h1T + (long)vt * (ETW_H1 * ROWS) + (row >> 5) * (32 * ETW_H1) + (row & 31);
        float* h2T = tw.h2T + (long)vt * (ETW_H2 * ROWS) + (row >> 5) * (32 * ETW_H2) + (row & 31);
        unsigned* reluT = tw.relu + (vt * ROWS + row) * 8;

        // ---- x * mask | 1: this thread's 28 columns of chunk a (c28..) and of chunk b (56 + c28..), all loads in flight ----
        float xa[28], xb[28];
        const void* mk = br ? a.mask[1] : a.mask[0];
#pragma unroll
        for (int half_ = 0; half_ < 2; ++half_) {
            float* dst = half_ ? xb : xa;
#pragma unroll
            for (int g = 0; g < 7; ++g) {
                const int c = 56 * half_ + c28 + 4 * g;
                float xv[4] = {0.f, 0.f, 0.f, 0.f};
                if (c < D) {
                    if (ok) {
                        const long gi = (long)grow * D + c;
                        const float4 x4 = *reinterpret_cast<const float4*>(a.x + gi);
                        float m[4];
                        load_mask4<true>(mk, gi, a.mask_kind, m);
                        xv[0] = x4.x * m[0]; xv[1] = x4.y * m[1]; xv[2] = x4.z * m[2]; xv[3] = x4.w * m[3];
                    }
                } else if (c == D) {
                    xv[0] = 1.0f;                          // bias column
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[4 * g + j] = xv[j];
            }
        }
        put28(lane_addr, c28, xa, inT, c28, D + 1, save);
        half_kick(bar, h, wh, [&] { issue_3x_chunk(tbase + H_ACC1, tbase + H_OPH, tbase + H_OPL, r_e1h, r_e1l, es1, 0, ks1a, idE1, false); });
        half_wait(hx, bar);
        if (ks1b > 0) {
            put28(lane_addr, c28, xb, inT, 56 + c28, D + 1, save);
            half_kick(bar, h, wh, [&] { issue_3x_chunk(tbase + H_ACC1, tbase + H_OPH, tbase + H_OPL, r_e1h, r_e1l, es1, ks1a, ks1b, idE1, true); });
            half_wait(hx, bar);
        }

        // ---- h1 = relu(acc1) | 1: columns c28.. (chunk a of layer 2) and 56 + c28.. (chunk b) ----
        uint32_t m1a = 0, m1b = 0;
        float ha[28], hb[28];
        tmem_ld28(lane_addr + H_ACC1 + c28, ha);
        tmem_ld28(lane_addr + H_ACC1 + 56 + c28, hb);
#pragma unroll
        for (int j = 0; j < 28; ++j) {
            if (ha[j] > 0.f) m1a |= 1u << j; else ha[j] = 0.f;
            if (hb[j] > 0.f) m1b |= 1u << j; else hb[j] = 0.f;
        }
        put28(lane_addr, c28, ha, h1T, c28, ETW_H1, save);
        half_kick(bar, h, wh, [&] { issue_3x_chunk(tbase + H_ACC2, tbase + H_OPH, tbase + H_OPL, r_e2h, r_e2l, es2, 0, KCH, idE2, false); });
        half_wait(hx, bar);
        put28(lane_addr, c28, hb, h1T, 56 + c28, ETW_H1, save);
        half_kick(bar, h, wh, [&] { issue_3x_chunk(tbase + H_ACC2, tbase + H_OPH, tbase + H_OPL, r_e2h, r_e2l, es2, KCH, F2_C / 2 - KCH, idE2, true); });
        // the row's noise while the MMAs run: latents 0..3 + 8, 9 (column group 0) or 4..7 (column group 1)
        const int l0 = 4 * cg;
        const long gl = (long)grow * LAT;
        const bool want_z = ok && a.z[br] && a.eps[br];
        float e4[4] = {0.f, 0.f, 0.f, 0.f}, e2[2] = {0.f, 0.f};
        if (want_z) {
            const float2 p0 = *reinterpret_cast<const float2*>(a.eps[br] + gl + l0), p1 = *reinterpret_cast<const float2*>(a.eps[br] + gl + l0 + 2);
            e4[0] = p0.x; e4[1] = p0.y; e4[2] = p1.x; e4[3] = p1.y;
            if (cg == 0) { const float2 p2 = *reinterpret_cast<const float2*>(a.eps[br] + gl + 8); e2[0] = p2.x; e2[1] = p2.y; }
        }
        half_wait(hx, bar);

        // ---- h2 = relu(acc2) | 1: columns 32 cg .. 32 cg + 31 (K = 56: one chunk) ----
        uint32_t m2 = 0;
        {
            float v[32], lo[16];
            tmem_ld16(lane_addr + H_ACC2 + 32 * cg, v);
            tmem_ld16(lane_addr + H_ACC2 + 32 * cg + 16, v + 16);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (v[j] > 0.f) m2 |= 1u << j; else v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) lo[j] = tf32_lo(v[j]);
            tmem_st16(lane_addr + H_OPH + 32 * cg, v);
            tmem_st16(lane_addr + H_OPL + 32 * cg, lo);
#pragma unroll
            for (int j = 0; j < 16; ++j) lo[j] = tf32_lo(v[16 + j]);
            if (cg == 0) { tmem_st16(lane_addr + H_OPH + 16, v + 16); tmem_st16(lane_addr + H_OPL + 16, lo); }
            else { tmem_st8(lane_addr + H_OPH + 48, v + 16); tmem_st8(lane_addr + H_OPL + 48, lo); }
            if (save) {
                scratch_store(h2T, 32 * cg, ETW_H2, v, 16);
                scratch_store(h2T, 32 * cg + 16, ETW_H2, v + 16, 16);
                reluT[cg] = m1a;
                reluT[2 + cg] = m1b;
                reluT[4 + 2 * cg] = m2 & 0xFFFFu;
                reluT[5 + 2 * cg] = m2 >> 16;
            }
        }
        half_kick(bar, h, wh, [&] { issue_3x_chunk(tbase + H_ACC3, tbase + H_OPH, tbase + H_OPL, r_e3h, r_e3l, es3, 0, F3_C / 2, idE3, false); });
        half_wait(hx, bar);

        // ---- mean | logvar, reparameterisation ----
        {
            float mv[4], lv[4];
            tmem_ld4(lane_addr + H_ACC3 + l0, mv);
            tmem_ld4(lane_addr + H_ACC3 + 16 + l0, lv);
            if (ok) {
                *reinterpret_cast<float2*>(a.mean[br] + gl + l0) = make_float2(mv[0], mv[1]);
                *reinterpret_cast<float2*>(a.mean[br] + gl + l0 + 2) = make_float2(mv[2], mv[3]);
                *reinterpret_cast<float2*>(a.logvar[br] + gl + l0) = make_float2(lv[0], lv[1]);
                *reinterpret_cast<float2*>(a.logvar[br] + gl + l0 + 2) = make_float2(lv[2], lv[3]);
                if (a.z[br]) {
                    float zz[4] = {mv[0], mv[1], mv[2], mv[3]};
                    if (a.eps[br]) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) zz[j] = fmaf(e4[j], expf(lv[j] * 0.5f), zz[j]);
                    }
                    *reinterpret_cast<float2*>(a.z[br] + gl + l0) = make_float2(zz[0], zz[1]);
                    *reinterpret_cast<float2*>(a.z[br] + gl + l0 + 2) = make_float2(zz[2], zz[3]);
                }
            }
            if (cg == 0) {
                tmem_ld4(lane_addr + H_ACC3 + 8, mv);
                tmem_ld4(lane_addr + H_ACC3 + 24, lv);
                if (ok) {
                    *reinterpret_cast<float2*>(a.mean[br] + gl + 8) = make_float2(mv[0], mv[1]);
                    *reinterpret_cast<float2*>(a.logvar[br] + gl + 8) = make_float2(lv[0], lv[1]);
                    if (a.z[br]) {
                        float z0 = mv[0], z1 = mv[1];
                        if (a.eps[br]) { z0 = fmaf(e2[0], expf(lv[0] * 0.5f), z0); z1 = fmaf(e2[1], expf(lv[1] * 0.5f), z1); }
                        *reinterpret_cast<float2*>(a.z[br] + gl + 8) = make_float2(z0, z1);
                    }
                }
            }
        }
        tc_fence_before();      // the half's next item overwrites OP / ACC only after its own barrier
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

static size_t enc_fwd_tc2_smem(int D) {
    const int C1 = ((D + 8) & ~7) / 4;
    return (size_t)2 * (C1 * F1_N * 4 + F2_C * F2_N * 4 + F3_C * F3_N * 4) * sizeof(float) + 128;
}

}  // namespace tc

int enc_fwd_tc2_launch(const EncFwdArgs& a, int grid, cudaStream_t st) {
    const size_t sm = tc::enc_fwd_tc2_smem(a.L.D);
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "enc_fwd_tc2: shared memory %zu B exceeds %d", sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(tc::k_enc_fwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "enc_fwd_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    prof_mark(st);
    tc::k_enc_fwd_tc2<<<grid, NT, sm, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "enc_fwd_tc2: launch: %s", cudaGetErrorString(e));
    prof_mark(st);
    return PCVAE_OK;
}

}  // namespace pcvae
