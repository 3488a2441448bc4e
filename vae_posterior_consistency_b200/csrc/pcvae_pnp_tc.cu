// PNP (EDDI) set encoder with its MLP tail on the tensor cores (Reg_EDDI / vanilla_EDDI.encoder and its backward,
// src/models/VAE.py:719-741, 903-925; emb_dim a multiple of 4).
//
// The encoder is  agg = sum_d m_d * relu(x_d * A_d + C_d)  (collapsed per-feature embedding + masked sum-pool,
// element-wise work with a ReLU inside the sum, CUDA cores)  followed by the MLP K -> 100 -> 50 -> 2L
// (dense, tensor cores).  The FFMA kernels k_enc_fwd<PNP> / k_enc_bwd<PNP> do both on the CUDA cores in one launch each
// (220 + 369 us at batch 65 536 x 2 branches, K = 20).  Here the two halves are separate launches:
//
//   forward    k_pnp_embed_fwd : agg[branch][row][K] (row-major) + a mask of ones          (pcvae_tile.cuh: pnp_embed)
//              k_enc_fwd_tc    : the zero-impute MLP encoder kernel run with obs_dim = K, x = agg of the branch
//                                (EncFwdArgs::x_bs), mask = ones; its scratch keeps agg | 1, h1, h2 for the backward
//   backward   k_enc_bwd_tc + k_wgrad_tc : dpre3, dpre2, dpre1 and dW1..dW3, db1..db3 exactly as for the MLP family
//              k_pnp_embed_bwd : d_agg = dpre1 * W1 (K x 100, FFMA), dA / dC accumulated per CTA (pnp_embed_bwd), then the
//                                chain rule through the collapsed tables to type_pars1, type_bias1, pnp_encoder1
//
// Parameter offsets are those of the PNP layout, so every gradient lands where the FFMA path puts it.
#include <cooperative_groups.h>
#include <cuda_pipeline.h>

#include "pcvae_tc.cuh"
#include "pcvae_train.cuh"

namespace pcvae {

// The two element-wise kernels are latency-bound (a tile is load -> barrier -> compute -> barrier -> combine): 32-row
// tiles keep their shared memory under 100 KB so that two CTAs share an SM and one's barriers hide under the other's math.
constexpr int TMP = 32;           // rows per tile of the two element-wise kernels
constexpr int PP = TMP + 4;       // feature-major pitch in shared memory
constexpr int EMB_SEG = 12;       // feature segments pnp_embed may split D into (partial sums in shared memory)
constexpr int EMB_CTAS = 2;       // CTAs per SM

struct PnpEmbArgs {
    Layout L;
    int B, nbr, mask_kind;
    const float* theta;
    const float* x;
    const void* mask[2];
    const float* ac;          // collapsed tables A | C
    float* agg;               // [nbr][B][K]
    unsigned char* ones;      // [B][K]
    const float* dp1T;        // backward: [nvt][row / 32][ETW_H1][row % 32]
    float* gp;                // backward: [grid][P]
};

__global__ void __launch_bounds__(NT, EMB_CTAS) k_pnp_embed_fwd(const PnpEmbArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int D = a.L.D, K = a.L.K, K4 = round4(K);
    float* in_s = smem;                      // [D][PP]  x
    float* ms_s = in_s + D * PP;             // [D][PP]  mask
    float* A_s = ms_s + D * PP;              // [D][K4]
    float* C_s = A_s + D * K4;               // [D][K4]
    float* agg_s = C_s + D * K4;             // [K4][PP]
    float* part_s = agg_s + K4 * PP;         // [EMB_SEG][K4][PP]
    for (int i = tid; i < 2 * D * K4; i += NT) A_s[i] = a.ac[i];
    // the mask of ones the tensor-core kernel multiplies the pooled embedding with (it is the MLP family's kernel)
    for (long i = (long)blockIdx.x * NT + tid; i < ((long)a.B * K + 3) / 4; i += (long)gridDim.x * NT)
        reinterpret_cast<unsigned*>(a.ones)[i] = 0x01010101u;
    __syncthreads();
    const int ntiles = (a.B + TMP - 1) / TMP;
    for (int vt = blockIdx.x; vt < ntiles * a.nbr; vt += gridDim.x) {
        const int br = vt / ntiles, row0 = (vt - br * ntiles) * TMP;
        const float* __restrict__ x = a.x;
        const void* __restrict__ mk = a.mask[br];
        {   // L2 prefetch of this CTA's next tile while the current one is being processed
            const int nvt = vt + gridDim.x;
            if (nvt < ntiles * a.nbr) {
                const int nbr_ = nvt / ntiles, nrow0 = (nvt - nbr_ * ntiles) * TMP;
                const long nrows = min(TMP, a.B - nrow0);
                prefetch_l2(x + (long)nrow0 * D, nrows * D * 4, tid);
                const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
                prefetch_l2((const char*)a.mask[nbr_] + (long)nrow0 * D * msz, nrows * D * msz, tid);
            }
        }
        tile_elems<TMP, 8, XM>(D, row0, a.B, tid,
            [&](int d, int r, bool ok) {
                XM v{0.f, 0.f};
                if (ok) {
                    const long gi = (long)(row0 + r) * D + d;
                    v.x = x[gi];
                    v.m = load_mask(mk, gi, a.mask_kind);
                }
                return v;
            },
            [&](int d, int r, bool, XM v) { in_s[d * PP + r] = v.x; ms_s[d * PP + r] = v.m; });
        __syncthreads();
        pnp_embed<TMP>(in_s, ms_s, A_s, C_s, agg_s, part_s, EMB_SEG * K4 * PP, D, K4, tid);
        __syncthreads();
        float* out = a.agg + ((long)br * a.B + row0) * K;
        for (int i = tid; i < TMP * K; i += NT) {
            const int r = i / K, k = i - r * K;
            if (row0 + r < a.B) out[i] = agg_s[k * PP + r];
        }
        __syncthreads();
    }
}

// Forward, uint8 masks and obs_dim % 4 == 0: the tile of x and of the mask is staged ROW-MAJOR by two bulk async copies (a
// 32-row tile of either is one contiguous block of the batch; double-buffered, requested a tile ahead), and the thread
// mapping is chosen so that row-major is what it wants: lane = row, warp = (quad of embedding columns, segment of the
// features).  A thread reads 4 features of its row with one LDS.128 (rows are 25 16-byte units apart: conflict-free per
// quarter-warp) and their mask bytes with one LDS.32, the table entries as warp-wide broadcasts: 2 LDS.128 + 12
// arithmetic instructions per feature for 4 (feature, column) pairs, no transposition, no staging instructions.  The
// feature-major version above spends a quarter of its instructions staging the tile.  Segment partials are combined in
// segment order.
__global__ void __launch_bounds__(NT, EMB_CTAS) k_pnp_embed_fwd_rm(const PnpEmbArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t in_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D, K = a.L.K, K4 = round4(K);
    float* xt = smem;                                        // [2][TMP][D]
    unsigned char* mt = reinterpret_cast<unsigned char*>(xt + 2 * TMP * D);   // [2][TMP][D] bytes
    float* A_s = reinterpret_cast<float*>(mt + 2 * TMP * D);  // [D][K4]
    float* C_s = A_s + D * K4;                               // [D][K4]
    float* part_s = C_s + D * K4;                            // [nseg][K4][PP]
    for (int i = tid; i < 2 * D * K4; i += NT) A_s[i] = a.ac[i];
    // the mask of ones the tensor-core kernel multiplies the pooled embedding with (it is the MLP family's kernel)
    for (long i = (long)blockIdx.x * NT + tid; i < ((long)a.B * K + 3) / 4; i += (long)gridDim.x * NT)
        reinterpret_cast<unsigned*>(a.ones)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&in_bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&in_bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tc::fence_async_smem();
    }
    __syncthreads();
    const int ngs = K4 / 4, nseg = NWARP / ngs;               // column quads, feature segments (K <= 32: nseg >= 2)
    const int kq = warp % ngs, sg = warp / ngs;
    const int dq = D / 4, qper = (dq + nseg - 1) / nseg;      // feature quads per segment
    const int q0 = sg * qper, q1 = min(dq, q0 + qper);
    const int ntiles = (a.B + TMP - 1) / TMP, nvt = ntiles * a.nbr;
    // full tiles arrive by bulk copies; a ragged last tile (its mask block need not be a multiple of 16 bytes) is copied by
    // the threads when its turn comes
    auto full_tile = [&](int vt) { return vt < nvt && ((vt % ntiles) + 1) * TMP <= a.B; };
    auto request = [&](int vt, int buf) {
        if (full_tile(vt) && tid == 0) {
            const int br = vt / ntiles, row0 = (vt - br * ntiles) * TMP;
            tc::mbar_expect_tx(&in_bar[buf], (uint32_t)(TMP * D * 5));
            tc::bulk_g2s(xt + buf * TMP * D, a.x + (long)row0 * D, (uint32_t)(TMP * D * 4), &in_bar[buf]);
            tc::bulk_g2s(reinterpret_cast<float*>(mt + buf * TMP * D),
                         reinterpret_cast<const float*>(static_cast<const unsigned char*>(a.mask[br]) + (long)row0 * D),
                         (uint32_t)(TMP * D), &in_bar[buf]);
        }
    };
    request(blockIdx.x, 0);
    uint32_t ph[2] = {0, 0};
    int it = 0;
    for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x, ++it) {
        const int buf = it & 1;
        const int br = vt / ntiles, row0 = (vt - br * ntiles) * TMP;
        request(vt + gridDim.x, buf ^ 1);                    // the other buffer was released by the barrier that ended the last tile
        if (full_tile(vt)) {
            tc::mbar_wait(&in_bar[buf], ph[buf], nullptr, 0);
            ph[buf] ^= 1u;
        } else {
            const int nrows = a.B - row0;
            const float* xs = a.x + (long)row0 * D;
            const unsigned char* ms = static_cast<const unsigned char*>(a.mask[br]) + (long)row0 * D;
            for (int i = tid; i < nrows * D; i += NT) {
                xt[buf * TMP * D + i] = xs[i];
                mt[buf * TMP * D + i] = ms[i];
            }
            __syncthreads();
        }
        const bool ok = row0 + lane < a.B;
        if (sg < nseg) {
            const float* xr = xt + (buf * TMP + lane) * D;
            const unsigned char* mr = mt + (buf * TMP + lane) * D;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (ok) {
#pragma unroll 2
                for (int q = q0; q < q1; ++q) {
                    const float4 x4 = lds4(xr + 4 * q);
                    const unsigned mw = *reinterpret_cast<const unsigned*>(mr + 4 * q);
                    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float m = ((mw >> (8 * j)) & 0xFFu) ? 1.f : 0.f;
                        const int d = 4 * q + j;
                        const float4 av = lds4(A_s + d * K4 + 4 * kq), cv = lds4(C_s + d * K4 + 4 * kq);
                        acc[0] = fmaf(m, fmaxf(fmaf(xv[j], av.x, cv.x), 0.f), acc[0]);
                        acc[1] = fmaf(m, fmaxf(fmaf(xv[j], av.y, cv.y), 0.f), acc[1]);
                        acc[2] = fmaf(m, fmaxf(fmaf(xv[j], av.z, cv.z), 0.f), acc[2]);
                        acc[3] = fmaf(m, fmaxf(fmaf(xv[j], av.w, cv.w), 0.f), acc[3]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) part_s[(sg * K4 + 4 * kq + j) * PP + lane] = acc[j];
        }
        __syncthreads();
        // pooled embedding of the tile, row-major [row][K]: warp = row, lane = column
        float* out = a.agg + ((long)br * a.B + row0) * K;
        for (int r = warp; r < TMP; r += NWARP) {
            if (lane < K && row0 + r < a.B) {
                float sum = part_s[lane * PP + r];
                for (int g2 = 1; g2 < nseg; ++g2) sum += part_s[(g2 * K4 + lane) * PP + r];
                out[(long)r * K + lane] = sum;
            }
        }
        __syncthreads();
    }
}

// Launched as clusters of EMB_CTAS CTAs: the gradient partials are [grid][P] with grid = one row per SM-sized slot
// (pcvae_grid_ctas), so the CTAs of a cluster add their dA / dC tables into rank 0's shared memory (distributed shared
// memory, fixed order) and rank 0 writes the cluster's row.
// RM (uint8 masks, obs_dim % 4 == 0): x and the mask are staged ROW-MAJOR by bulk async copies as in k_pnp_embed_fwd_rm --
// the (feature, column-quad) owner threads read x[row][d] with lanes along d, which is conflict-free for any pitch -- instead
// of being transposed into feature-major tiles by the threads (16.6 % of this kernel's instructions).
template <bool RM>
__global__ void __cluster_dims__(EMB_CTAS, 1, 1) __launch_bounds__(NT, EMB_CTAS) k_pnp_embed_bwd(const PnpEmbArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t in_bar[2];
    constexpr int RB = 1;
    const int tid = threadIdx.x;
    const int D = a.L.D, K = a.L.K, K4 = round4(K);
    // RM: xt [2][TMP][D] floats | mt [2][TMP][D] bytes;  otherwise: in_s [D][PP] | ms_s [D][PP]
    float* in_s = smem;
    float* ms_s = in_s + D * PP;
    float* xt = smem;
    unsigned char* mt = reinterpret_cast<unsigned char*>(xt + 2 * TMP * D);
    float* dp1_s = RM ? reinterpret_cast<float*>(mt + 2 * TMP * D) : ms_s + D * PP;   // [100][PP]  dL/d(pre1)
    float* agg_s = dp1_s + H1 * PP;          // [K4][PP]   dL/d(agg)
    float* W1_s = agg_s + K4 * PP;           // [K][100]
    float* A_s = W1_s + K * H1;              // [D][K4] (A then C)
    float* C_s = A_s + D * K4;
    float* dA_s = C_s + D * K4;              // [D][K4] (dA then dC)
    float* dC_s = dA_s + D * K4;
    stage_linear(W1_s, nullptr, a.theta + a.L.W1, nullptr, K, H1, H1, tid);
    for (int i = tid; i < 2 * D * K4; i += NT) { A_s[i] = a.ac[i]; dA_s[i] = 0.f; }
    zero_floats(agg_s, K4 * PP, tid);
    if (RM && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&in_bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&in_bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tc::fence_async_smem();
    }
    __syncthreads();
    const int ngs = K4 / 4;
    const int ntiles = (a.B + TMP - 1) / TMP, nt128 = (a.B + 127) / 128, nvt = ntiles * a.nbr;
    auto full_tile = [&](int vt) { return vt < nvt && ((vt % ntiles) + 1) * TMP <= a.B; };
    auto request = [&](int vt, int buf) {
        if (RM && full_tile(vt) && tid == 0) {
            const int br = vt / ntiles, row0 = (vt - br * ntiles) * TMP;
            tc::mbar_expect_tx(&in_bar[buf], (uint32_t)(TMP * D * 5));
            tc::bulk_g2s(xt + buf * TMP * D, a.x + (long)row0 * D, (uint32_t)(TMP * D * 4), &in_bar[buf]);
            tc::bulk_g2s(reinterpret_cast<float*>(mt + buf * TMP * D),
                         reinterpret_cast<const float*>(static_cast<const unsigned char*>(a.mask[br]) + (long)row0 * D),
                         (uint32_t)(TMP * D), &in_bar[buf]);
        }
    };
    request(blockIdx.x, 0);
    uint32_t ph[2] = {0, 0};
    int it = 0;
    for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x, ++it) {
        const int buf = it & 1;
        const int br = vt / ntiles, t64 = vt - br * ntiles, row0 = t64 * TMP;
        const float* __restrict__ x = a.x;
        const void* __restrict__ mk = a.mask[br];
        request(vt + gridDim.x, buf ^ 1);                    // released by the barrier that ended the last tile
        {   // the dpre1 rows of this tile: one 32-row slab [feature][32] of a 128-row tile of the tensor-core scratch
            static_assert(TMP == 32, "one slab per tile");
            const float* src = a.dp1T + ((long)br * nt128 + (t64 >> 2)) * (ETW_H1 * 128) + (long)(t64 & 3) * (32 * ETW_H1);
            for (int i = tid; i < H1 * (TMP / 4); i += NT) {
                const int f = i / (TMP / 4), c = i - f * (TMP / 4);
                __pipeline_memcpy_async(dp1_s + f * PP + 4 * c, src + f * 32 + 4 * c, 16);
            }
            __pipeline_commit();
        }
        if (!RM) {
            tile_elems<TMP, 8, XM>(D, row0, a.B, tid,
                [&](int d, int r, bool ok) {
                    XM v{0.f, 0.f};
                    if (ok) {
                        const long gi = (long)(row0 + r) * D + d;
                        v.x = x[gi];
                        v.m = load_mask(mk, gi, a.mask_kind);
                    }
                    return v;
                },
                [&](int d, int r, bool, XM v) { in_s[d * PP + r] = v.x; ms_s[d * PP + r] = v.m; });
        } else if (full_tile(vt)) {
            tc::mbar_wait(&in_bar[buf], ph[buf], nullptr, 0);
            ph[buf] ^= 1u;
        } else {                                             // ragged last tile: copied by the threads, the rest masked off
            const int nrows = a.B - row0;
            const float* xs = x + (long)row0 * D;
            const unsigned char* ms = static_cast<const unsigned char*>(mk) + (long)row0 * D;
            for (int i = tid; i < TMP * D; i += NT) {
                xt[buf * TMP * D + i] = i < nrows * D ? xs[i] : 0.f;
                mt[buf * TMP * D + i] = i < nrows * D ? ms[i] : (unsigned char)0;
            }
        }
        __pipeline_wait_prior(0);
        __syncthreads();
        gemm_dx<TMP, RB, false, 2>(dp1_s, W1_s, agg_s, K, H1, tid);      // agg_s <- dL/d(agg)
        __syncthreads();
        if (!RM) {
            pnp_embed_bwd<TMP>(in_s, ms_s, agg_s, A_s, C_s, dA_s, dC_s, D, K, K4, tid);
        } else {
            // as pnp_embed_bwd (pcvae_tile.cuh), reading x and the mask bytes row-major: lanes run along the feature
            for (int item = tid; item < D * ngs; item += NT) {
                const int jg = item / D, d = item - jg * D, j0 = 4 * jg;
                const float4 a4 = lds4(A_s + d * K4 + j0), c4 = lds4(C_s + d * K4 + j0);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
                float accA[4] = {0.f, 0.f, 0.f, 0.f}, accC[4] = {0.f, 0.f, 0.f, 0.f};
                const float* xc = xt + buf * TMP * D + d;
                const unsigned char* mc = mt + buf * TMP * D + d;
#pragma unroll 2
                for (int r = 0; r < TMP; r += 4) {
                    float xv[4], mv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { xv[i] = xc[(r + i) * D]; mv[i] = mc[(r + i) * D] ? 1.f : 0.f; }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 g4 = lds4(agg_s + (j0 + jj) * PP + r);
                        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float tval = mv[i] * gv[i];
                            if (fmaf(xv[i], av[jj], cv[jj]) > 0.f) {
                                accC[jj] += tval;
                                accA[jj] = fmaf(tval, xv[i], accA[jj]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (j0 + jj < K) { dA_s[d * K4 + j0 + jj] += accA[jj]; dC_s[d * K4 + j0 + jj] += accC[jj]; }
            }
        }
        __syncthreads();
    }
    {
        namespace cg = cooperative_groups;
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();                                         // every CTA of the cluster has finished its tiles
        for (unsigned rk = 1; rk < cl.num_blocks(); ++rk) {
            if (cl.block_rank() == rk) {
                float* dst = cl.map_shared_rank(dA_s, 0);
                for (int i = tid; i < 2 * D * K4; i += NT) dst[i] += dA_s[i];
            }
            cl.sync();
        }
        if (cl.block_rank() != 0) return;
    }
    // chain rule through the collapsed tables back to type_pars1, type_bias1, pnp_encoder1 (as k_enc_bwd<PNP>)
    float* gp = a.gp + (long)(blockIdx.x / EMB_CTAS) * a.L.total;
    const float* th = a.theta;
    for (int i = tid; i < D * K; i += NT) {           // dE[d][q] = sum_j dA[d][j] We[j][1+q]
        const int d = i / K, q = i - d * K;
        float s = 0.f;
        for (int j = 0; j < K; ++j) s = fmaf(dA_s[d * K4 + j], th[a.L.We + j * (K + 2) + 1 + q], s);
        gp[a.L.E + i] = s;
    }
    for (int d = tid; d < D; d += NT) {               // dbE[d] = sum_j dC[d][j] We[j][K+1]
        float s = 0.f;
        for (int j = 0; j < K; ++j) s = fmaf(dC_s[d * K4 + j], th[a.L.We + j * (K + 2) + K + 1], s);
        gp[a.L.bE + d] = s;
    }
    for (int i = tid; i < K * (K + 2); i += NT) {     // dWe[j][c]
        const int j = i / (K + 2), c = i - j * (K + 2);
        float s = 0.f;
        if (c == 0) for (int d = 0; d < D; ++d) s += dA_s[d * K4 + j];
        else if (c <= K) for (int d = 0; d < D; ++d) s = fmaf(dA_s[d * K4 + j], th[a.L.E + d * K + (c - 1)], s);
        else for (int d = 0; d < D; ++d) s = fmaf(dC_s[d * K4 + j], th[a.L.bE + d], s);
        gp[a.L.We + i] = s;
    }
    for (int j = tid; j < K; j += NT) {               // dbe[j] = sum_d dC[d][j]
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += dC_s[d * K4 + j];
        gp[a.L.be + j] = s;
    }
}

static size_t emb_fwd_smem(const Layout& L) {
    const int K4 = round4(L.K);
    return ((size_t)2 * L.D * PP + 2 * L.D * K4 + (size_t)(1 + EMB_SEG) * K4 * PP) * sizeof(float);
}
static size_t emb_fwd_rm_smem(const Layout& L) {
    const int K4 = round4(L.K), nseg = NWARP / (K4 / 4);
    return (size_t)2 * TMP * L.D * 5 + ((size_t)2 * L.D * K4 + (size_t)nseg * K4 * PP) * sizeof(float) + 128;
}
static size_t emb_bwd_smem(const Layout& L, bool rm) {
    const int K4 = round4(L.K);
    const size_t stage = rm ? (size_t)2 * TMP * L.D * 5 : (size_t)2 * L.D * PP * sizeof(float);
    return stage + ((size_t)H1 * PP + K4 * PP + (size_t)L.K * H1 + 4 * L.D * K4) * sizeof(float) + 128;
}

// the MLP tail as the tensor-core kernels see it: obs_dim = emb_dim, weights where the PNP layout keeps them
static Layout tail_layout(const Layout& L) {
    Layout T = L;
    T.fam = PCVAE_FAMILY_MLP;
    T.D = L.K;
    T.K = 0;
    T.aug = 0;
    return T;
}

Layout pnp_tail_layout(const Layout& L) { return tail_layout(L); }

bool pnp_tc_supported(const Layout& L) {
    return L.fam == PCVAE_FAMILY_PNP && L.K % 4 == 0 && L.K >= 4 && L.K <= MAX_K && enc_tc_supported(tail_layout(L)) &&
           emb_fwd_smem(L) <= (size_t)MAX_SMEM && emb_bwd_smem(L, false) <= (size_t)MAX_SMEM;
}

// agg [nbr][rows][K] floats, then the ones mask [rows][K] bytes (both rounded up to 16 bytes)
long pnp_tc_extra_floats(const Layout& L, long rows, int nbr) {
    return ((long)nbr * rows * L.K + 3) / 4 * 4 + ((long)rows * L.K + 15) / 16 * 4;
}

template <typename Kern>
static int emb_go(Kern kern, const PnpEmbArgs& args, size_t sm, int grid, cudaStream_t st, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    kern<<<EMB_CTAS * grid, NT, sm, st>>>(args);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: launch: %s", name, cudaGetErrorString(e));
    return PCVAE_OK;
}

static void emb_args(PnpEmbArgs* e, const Layout& L, int B, int nbr, int mask_kind, const float* theta, const float* x,
                     const void* const* mask, const float* ac, float* extra) {
    e->L = L; e->B = B; e->nbr = nbr; e->mask_kind = mask_kind; e->theta = theta; e->x = x;
    e->mask[0] = mask[0]; e->mask[1] = mask[1]; e->ac = ac;
    e->agg = extra;
    e->ones = reinterpret_cast<unsigned char*>(extra + ((long)nbr * B * L.K + 3) / 4 * 4);
}

int pnp_enc_fwd_tc_launch(const EncFwdArgs& a, float* extra, int grid, cudaStream_t st) {
    PnpEmbArgs e{};
    emb_args(&e, a.L, a.B, a.nbr, a.mask_kind, a.theta, a.x, a.mask, a.ac, extra);
    prof_mark(st);
    // row-major staging by bulk copies needs byte masks and 16-byte rows
    const bool rm = a.mask_kind == PCVAE_MASK_U8 && a.L.D % 4 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.mask[0]) & 15) == 0 && (a.nbr < 2 || (reinterpret_cast<uintptr_t>(a.mask[1]) & 15) == 0);
    if (rm) { if (int rc = emb_go(k_pnp_embed_fwd_rm, e, emb_fwd_rm_smem(a.L), grid, st, "pnp_embed_fwd")) return rc; }
    else if (int rc = emb_go(k_pnp_embed_fwd, e, emb_fwd_smem(a.L), grid, st, "pnp_embed_fwd")) return rc;
    EncFwdArgs t = a;
    t.L = tail_layout(a.L);
    t.x = e.agg;
    t.x_bs = (long)a.B * a.L.K;
    t.mask[0] = t.mask[1] = e.ones;
    t.mask_kind = PCVAE_MASK_U8;
    t.act_ws = nullptr;
    return enc_fwd_tc_launch(t, grid, st);
}

int pnp_enc_bwd_tc_launch(const EncBwdArgs& a, float* extra, int grid, cudaStream_t st) {
    EncBwdArgs t = a;
    t.L = tail_layout(a.L);
    if (int rc = enc_bwd_tc_launch(t, grid, st)) return rc;
    PnpEmbArgs e{};
    emb_args(&e, a.L, a.B, a.nbr, a.mask_kind, a.theta, a.x, a.mask, a.ac, extra);
    e.dp1T = a.tw.dp1T;
    e.gp = a.gp;
    const bool rm = a.mask_kind == PCVAE_MASK_U8 && a.L.D % 4 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.mask[0]) & 15) == 0 && (a.nbr < 2 || (reinterpret_cast<uintptr_t>(a.mask[1]) & 15) == 0);
    const int rc = rm ? emb_go(k_pnp_embed_bwd<true>, e, emb_bwd_smem(a.L, true), grid, st, "pnp_embed_bwd")
                      : emb_go(k_pnp_embed_bwd<false>, e, emb_bwd_smem(a.L, false), grid, st, "pnp_embed_bwd");
    prof_mark(st);
    return rc;
}

}  // namespace pcvae
