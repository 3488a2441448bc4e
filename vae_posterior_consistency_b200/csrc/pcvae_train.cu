// Training-path kernels: encoder forward, decoder + loss (+ backward), encoder backward,
// stand-alone loss terms, partial reductions and Adam.  See include/pcvae_b200.h for the
// reference code each entry point replaces.
#include <cuda_pipeline.h>

#include "pcvae_internal.cuh"
#include "pcvae_train.cuh"

namespace pcvae {

// ====================================================================================
// PNP collapsed tables  A[d][j] = We[j][0] + sum_i E[d][i] We[j][1+i],
//                       C[d][j] = bE[d] We[j][K+1] + be[j]          
// stored as ac[0..D*K4) = A, ac[D*K4..2*D*K4) = C with K4 = round4(K), pads zero.
// ====================================================================================
__global__ void k_pnp_tables(Layout L, const float* __restrict__ theta, float* __restrict__ ac) {
    const int D = L.D, K = L.K, K4 = round4(K);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D * K4; i += gridDim.x * blockDim.x) {
        const int d = i / K4, j = i - d * K4;
        float a = 0.f, c = 0.f;
        if (j < K) {
            const float* We = theta + L.We + j * (K + 2);
            a = We[0];
            for (int q = 0; q < K; ++q) a = fmaf(theta[L.E + d * K + q], We[1 + q], a);
            c = fmaf(theta[L.bE + d], We[K + 1], theta[L.be + j]);
        }
        ac[i] = a;
        ac[D * K4 + i] = c;
    }
}

void pnp_tables_launch(const Layout& L, const float* theta, float* ac, cudaStream_t st) {
    k_pnp_tables<<<8, 256, 0, st>>>(L, theta, ac);
}

// ====================================================================================
// Encoder forward
// ====================================================================================
template <int FAM, int TM>
__global__ void __launch_bounds__(NT, 1) k_enc_fwd(const EncFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RB = 1;   // 4x4 register blocks (see pcvae_tile.cuh)
    constexpr int P = TM + 4;
    const int tid = threadIdx.x;
    const int D = a.L.D, K = a.L.K, K4 = round4(K);
    const int IN1 = (FAM == PCVAE_FAMILY_MLP) ? (a.L.aug ? 2 * D : D) : K;
    const int INR = (FAM == PCVAE_FAMILY_MLP) ? IN1 : D;          // feature rows of in_s
    float* W1_s = smem;
    float* b1_s = W1_s + IN1 * H1;
    float* W2_s = b1_s + H1;
    float* b2_s = W2_s + H1 * H2P;
    float* W3_s = b2_s + H2P;
    float* b3_s = W3_s + H2 * LAT2;
    float* in_s = b3_s + LAT2;            // [INR][P]  MLP: x*mask (, mask)   PNP: x
    float* h1_s = in_s + INR * P;         // [100][P]
    float* h2_s = h1_s + H1 * P;          // [52][P]
    float* o_s = h2_s + H2P * P;          // [20][P]
    float* ms_s = o_s + LAT2 * P;         // PNP [D][P]
    float* A_s = ms_s + D * P;            // PNP [D][K4]
    float* C_s = A_s + D * K4;            // PNP [D][K4]
    float* agg_s = C_s + D * K4;          // PNP [K4][P]

    stage_linear(W1_s, b1_s, a.theta + a.L.W1, a.theta + a.L.b1, IN1, H1, H1, tid);
    stage_linear(W2_s, b2_s, a.theta + a.L.W2, a.theta + a.L.b2, H1, H2, H2P, tid);
    stage_linear(W3_s, b3_s, a.theta + a.L.W3, a.theta + a.L.b3, H2, LAT2, LAT2, tid);
    if (FAM == PCVAE_FAMILY_PNP)
        for (int i = tid; i < 2 * D * K4; i += NT) A_s[i] = a.ac[i];
    __syncthreads();

    const int ntiles = (a.B + TM - 1) / TM;
    const int act_tile = enc_act_feats(FAM, K) * TM;
    for (int vt = blockIdx.x; vt < ntiles * a.nbr; vt += gridDim.x) {
        const int br = vt / ntiles, row0 = (vt - br * ntiles) * TM;
        const float* __restrict__ x = a.x;
        const void* __restrict__ mk = a.mask[br];
        {   // L2 prefetch of this CTA's next tile while the current one is being processed
            const int nvt = vt + gridDim.x;
            if (nvt < ntiles * a.nbr) {
                const int nbr_ = nvt / ntiles, nrow0 = (nvt - nbr_ * ntiles) * TM;
                const long nrows = min(TM, a.B - nrow0);
                prefetch_l2(x + (long)nrow0 * D, nrows * D * 4, tid);
                const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
                prefetch_l2((const char*)a.mask[nbr_] + (long)nrow0 * D * msz, nrows * D * msz, tid);
            }
        }
        tile_elems<TM, 16, XM>(D, row0, a.B, tid,
            [&](int d, int r, bool ok) {
                XM v{0.f, 0.f};
                if (ok) {
                    const long gi = (long)(row0 + r) * D + d;
                    v.x = x[gi];
                    v.m = load_mask(mk, gi, a.mask_kind);
                }
                return v;
            },
            [&](int d, int r, bool, XM v) {
                if (FAM == PCVAE_FAMILY_MLP) { in_s[d * P + r] = v.x * v.m; if (a.L.aug) in_s[(D + d) * P + r] = v.m; }
                else { in_s[d * P + r] = v.x; ms_s[d * P + r] = v.m; }
            });
        __syncthreads();
        if (FAM == PCVAE_FAMILY_PNP) {
            pnp_embed<TM>(in_s, ms_s, A_s, C_s, agg_s, h1_s, (H1 + H2P) * P, D, K4, tid);
            __syncthreads();
            gemm_fwd<TM, RB, ACT_RELU>(agg_s, W1_s, b1_s, h1_s, K, H1, tid);
        } else {
            gemm_fwd<TM, RB, ACT_RELU>(in_s, W1_s, b1_s, h1_s, IN1, H1, tid);
        }
        __syncthreads();
        gemm_fwd<TM, RB, ACT_RELU, 2>(h1_s, W2_s, b2_s, h2_s, H1, H2P, tid);
        __syncthreads();
        gemm_fwd<TM, RB, ACT_NONE, 2>(h2_s, W3_s, b3_s, o_s, H2, LAT2, tid);
        __syncthreads();
        // outputs (row-major [B][L]); reparameterisation z = mean + eps * exp(logvar/2)  (VAE.py:390-392)
        for (int i = tid; i < TM * LAT; i += NT) {
            const int r = i / LAT, l = i - r * LAT;
            if (row0 + r < a.B) {
                const long gi = (long)(row0 + r) * LAT + l;
                const float mu = o_s[l * P + r], lv = o_s[(LAT + l) * P + r];
                a.mean[br][gi] = mu;
                a.logvar[br][gi] = lv;
                if (a.z[br]) a.z[br][gi] = a.eps[br] ? fmaf(a.eps[br][gi], expf(lv * 0.5f), mu) : mu;
            }
        }
        if (a.act_ws) {
            // saved activations, tile-blocked feature-major [feat][TM]: (agg,) h1, h2
            float* ws = a.act_ws + (long)vt * act_tile;
            int f0 = 0;
            if (FAM == PCVAE_FAMILY_PNP) {
                for (int i = tid; i < K4 * (TM / 4); i += NT) {
                    const int f = i / (TM / 4), c = i - f * (TM / 4);
                    sts4(ws + f * TM + 4 * c, lds4(agg_s + f * P + 4 * c));
                }
                f0 = K4;
            }
            for (int i = tid; i < (H1 + H2P) * (TM / 4); i += NT) {
                const int f = i / (TM / 4), c = i - f * (TM / 4);
                const float* src = (f < H1) ? (h1_s + f * P) : (h2_s + (f - H1) * P);
                sts4(ws + (f0 + f) * TM + 4 * c, lds4(src + 4 * c));
            }
        }
        __syncthreads();
    }
}

// ====================================================================================
// Encoder backward
// ====================================================================================
template <int FAM, int TM>
__global__ void __launch_bounds__(NT, 1) k_enc_bwd(const EncBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RB = 1;   // 4x4 register blocks (see pcvae_tile.cuh)
    constexpr int P = TM + 4;
    const int tid = threadIdx.x;
    const int D = a.L.D, K = a.L.K, K4 = round4(K);
    const int IN1 = (FAM == PCVAE_FAMILY_MLP) ? (a.L.aug ? 2 * D : D) : K;
    const int INR = (FAM == PCVAE_FAMILY_MLP) ? IN1 : D;               // feature rows of in_s
    float* W2_s = smem;                        // [100][52]
    float* W3_s = W2_s + H1 * H2P;             // [50][20]
    float* dW1_s = W3_s + H2 * LAT2;           // [IN1][100]
    float* db1_s = dW1_s + IN1 * H1;
    float* dW2_s = db1_s + H1;                 // [100][52]
    float* db2_s = dW2_s + H1 * H2P;
    float* dW3_s = db2_s + H2P;                // [50][20]
    float* db3_s = dW3_s + H2 * LAT2;
    float* in_s = db3_s + LAT2;                // [INR][P]
    float* h1_s = in_s + INR * P;              // [100][P]
    float* h2_s = h1_s + H1 * P;               // [52][P]
    float* d3_s = h2_s + H2P * P;              // [20][P]
    float* ms_s = d3_s + LAT2 * P;             // PNP [D][P]
    float* agg_s = ms_s + D * P;               // PNP [K4][P]
    float* W1_s = agg_s + K4 * P;              // PNP [K][100]
    float* A_s = W1_s + K * H1;                // PNP [D][K4] (A then C)
    float* C_s = A_s + D * K4;
    float* dA_s = C_s + D * K4;                // PNP [D][K4] (dA then dC)
    float* dC_s = dA_s + D * K4;

    stage_linear(W2_s, nullptr, a.theta + a.L.W2, nullptr, H1, H2, H2P, tid);
    stage_linear(W3_s, nullptr, a.theta + a.L.W3, nullptr, H2, LAT2, LAT2, tid);
    zero_floats(dW1_s, (int)(in_s - dW1_s), tid);
    if (FAM == PCVAE_FAMILY_PNP) {
        stage_linear(W1_s, nullptr, a.theta + a.L.W1, nullptr, K, H1, H1, tid);
        for (int i = tid; i < 2 * D * K4; i += NT) { A_s[i] = a.ac[i]; dA_s[i] = 0.f; }
    }
    __syncthreads();

    const int ntiles = (a.B + TM - 1) / TM;
    const int act_tile = enc_act_feats(FAM, K) * TM;
    for (int vt = blockIdx.x; vt < ntiles * a.nbr; vt += gridDim.x) {
        const int br = vt / ntiles, row0 = (vt - br * ntiles) * TM;
        const float* __restrict__ x = a.x;
        const void* __restrict__ mk = a.mask[br];
        {   // L2 prefetch of this CTA's next tile while the current one is being processed
            const int nvt = vt + gridDim.x;
            if (nvt < ntiles * a.nbr) {
                const int nbr_ = nvt / ntiles, nrow0 = (nvt - nbr_ * ntiles) * TM;
                const long nrows = min(TM, a.B - nrow0);
                prefetch_l2(x + (long)nrow0 * D, nrows * D * 4, tid);
                const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
                prefetch_l2((const char*)a.mask[nbr_] + (long)nrow0 * D * msz, nrows * D * msz, tid);
            }
        }
        tile_elems<TM, 16, XM>(D, row0, a.B, tid,
            [&](int d, int r, bool ok) {
                XM v{0.f, 0.f};
                if (ok) {
                    const long gi = (long)(row0 + r) * D + d;
                    v.x = x[gi];
                    v.m = load_mask(mk, gi, a.mask_kind);
                }
                return v;
            },
            [&](int d, int r, bool, XM v) {
                if (FAM == PCVAE_FAMILY_MLP) { in_s[d * P + r] = v.x * v.m; if (a.L.aug) in_s[(D + d) * P + r] = v.m; }
                else { in_s[d * P + r] = v.x; ms_s[d * P + r] = v.m; }
            });
        {
            const float* ws = a.act_ws + (long)vt * act_tile;
            int f0 = 0;
            if (FAM == PCVAE_FAMILY_PNP) {
                for (int i = tid; i < K4 * (TM / 4); i += NT) {
                    const int f = i / (TM / 4), c = i - f * (TM / 4);
                    __pipeline_memcpy_async(agg_s + f * P + 4 * c, ws + f * TM + 4 * c, 16);
                }
                f0 = K4;
            }
            for (int i = tid; i < (H1 + H2P) * (TM / 4); i += NT) {
                const int f = i / (TM / 4), c = i - f * (TM / 4);
                float* dst = (f < H1) ? (h1_s + f * P) : (h2_s + (f - H1) * P);
                __pipeline_memcpy_async(dst + 4 * c, ws + (f0 + f) * TM + 4 * c, 16);
            }
            __pipeline_commit();
        }
#pragma unroll
        for (int i = tid; i < TM * LAT; i += NT) {
            const int r = i / LAT, l = i - r * LAT;
            float dm = 0.f, dv = 0.f;
            if (row0 + r < a.B) {
                const long gi = (long)(row0 + r) * LAT + l;
                dm = a.d_mean[br][gi];
                dv = a.d_logvar[br][gi];
                if (a.d_z[br]) {
                    const float dz = a.d_z[br][gi];
                    dm += dz;
                    if (a.eps[br]) dv = fmaf(dz * 0.5f * expf(a.logvar[br][gi] * 0.5f), a.eps[br][gi], dv);
                }
            }
            d3_s[l * P + r] = dm;
            d3_s[(LAT + l) * P + r] = dv;
        }
        __pipeline_wait_prior(0);
        __syncthreads();
        gemm_dw<TM, 2>(h2_s, d3_s, dW3_s, H2, LAT2, LAT2, tid);
        bias_dw<TM>(d3_s, db3_s, LAT2, tid);
        __syncthreads();
        gemm_dx<TM, RB, true, 2>(d3_s, W3_s, h2_s, H2, LAT2, tid);     // h2_s <- dL/d(pre2)
        __syncthreads();
        gemm_dw<TM, 2>(h1_s, h2_s, dW2_s, H1, H2, H2P, tid);
        bias_dw<TM>(h2_s, db2_s, H2, tid);
        __syncthreads();
        gemm_dx<TM, RB, true>(h2_s, W2_s, h1_s, H1, H2P, tid);      // h1_s <- dL/d(pre1)
        __syncthreads();
        if (FAM == PCVAE_FAMILY_MLP) {
            gemm_dw<TM>(in_s, h1_s, dW1_s, IN1, H1, H1, tid);
            bias_dw<TM>(h1_s, db1_s, H1, tid);
        } else {
            gemm_dw<TM, 2>(agg_s, h1_s, dW1_s, K, H1, H1, tid);
            bias_dw<TM>(h1_s, db1_s, H1, tid);
            __syncthreads();
            gemm_dx<TM, RB, false, 2>(h1_s, W1_s, agg_s, K, H1, tid);  // agg_s <- dL/d(agg)
            __syncthreads();
            pnp_embed_bwd<TM>(in_s, ms_s, agg_s, A_s, C_s, dA_s, dC_s, D, K, K4, tid);
        }
        __syncthreads();
    }

    // flush per-CTA partial gradients (encoder slice of the flat layout)
    float* gp = a.gp + (long)blockIdx.x * a.L.total;
    flush_linear_grad(dW1_s, db1_s, gp + a.L.W1, gp + a.L.b1, IN1, H1, H1, tid);
    flush_linear_grad(dW2_s, db2_s, gp + a.L.W2, gp + a.L.b2, H1, H2, H2P, tid);
    flush_linear_grad(dW3_s, db3_s, gp + a.L.W3, gp + a.L.b3, H2, LAT2, LAT2, tid);
    if (FAM == PCVAE_FAMILY_PNP) {
        // chain rule through the collapsed tables back to type_pars1, type_bias1, pnp_encoder1
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
}

// ====================================================================================
// Decoder (+ loss, + backward)
// ====================================================================================

template <int TM>
__global__ void __launch_bounds__(NT, 1) k_dec(const DecArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RB = 1;   // 4x4 register blocks (see pcvae_tile.cuh)
    __shared__ float red_s[NWARP][PCVAE_NSUMS];
    constexpr int P = TM + 4;
    const int tid = threadIdx.x;
    const int D = a.L.D, DP = round4(D);
    const bool bwd = (a.mode == PCVAE_DEC_TRAIN || a.mode == PCVAE_DEC_BWD);
    const bool lossy = (a.mode == PCVAE_DEC_TRAIN || a.mode == PCVAE_DEC_EVAL);
    float* W4_s = smem;                     // [10][52]
    float* b4_s = W4_s + LAT * G1P;
    float* W5_s = b4_s + G1P;               // [50][100]
    float* b5_s = W5_s + G1 * G2;
    float* W6_s = b5_s + G2;                // [100][DP]
    float* b6_s = W6_s + G2 * DP;
    float* z_s = b6_s + DP;                 // [12][P]
    float* g1_s = z_s + LATP * P;           // [52][P]
    float* g2_s = g1_s + G1P * P;           // [100][P]
    float* xh_s = g2_s + G2 * P;            // [DP][P]
    float* dW4_s = xh_s + DP * P;           // gradient accumulators (bwd modes only)
    float* db4_s = dW4_s + LAT * G1P;
    float* dW5_s = db4_s + G1P;
    float* db5_s = dW5_s + G1 * G2;
    float* dW6_s = db5_s + G2;
    float* db6_s = dW6_s + G2 * DP;
    float* dend_s = db6_s + DP;

    stage_linear(W4_s, b4_s, a.theta + a.L.W4, a.theta + a.L.b4, LAT, G1, G1P, tid);
    stage_linear(W5_s, b5_s, a.theta + a.L.W5, a.theta + a.L.b5, G1, G2, G2, tid);
    stage_linear(W6_s, b6_s, a.theta + a.L.W6, a.theta + a.L.b6, G2, D, DP, tid);
    if (bwd) zero_floats(dW4_s, (int)(dend_s - dW4_s), tid);
    zero_floats(z_s, LATP * P, tid);
    __syncthreads();

    // NLL constants exactly as torch.distributions.Normal computes them (scale = exp(lv/2),
    // var = scale^2, log_prob = -(d^2)/(2 var) - log(scale) - log(sqrt(2 pi)))
    const float scale = expf(a.x_logvar * 0.5f);
    const float var = scale * scale;
    const float inv2var = 1.0f / (2.0f * var);
    const float inv_var = 1.0f / var;
    const float log_scale = logf(scale);
    const float alpha = a.alpha, ls = a.loss_scale;

    float s_req = 0.f, s_rep = 0.f, s_klq = 0.f, s_klp = 0.f, s_klr = 0.f, s_red = 0.f, s_imp = 0.f, s_sse = 0.f;

    const int ntiles = (a.B + TM - 1) / TM;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = t * TM;
        if (lossy) {
            const long nrows = min(TM, a.B - row0);
            const int msz = a.mask_kind == PCVAE_MASK_U8 ? 1 : 4;
            prefetch_l2(a.x + (long)row0 * D, nrows * D * 4, tid);
            for (int b = 0; b < a.nbr; ++b) prefetch_l2((const char*)a.mask[b] + (long)row0 * D * msz, nrows * D * msz, tid);
        }
        for (int br = 0; br < a.nbr; ++br) {
#pragma unroll
            for (int i = tid; i < TM * LAT; i += NT) {
                const int r = i / LAT, l = i - r * LAT;
                z_s[l * P + r] = (row0 + r < a.B) ? a.z[br][(long)(row0 + r) * LAT + l] : 0.f;
            }
            __syncthreads();
            gemm_fwd<TM, RB, ACT_RELU, 2>(z_s, W4_s, b4_s, g1_s, LAT, G1P, tid);
            __syncthreads();
            gemm_fwd<TM, RB, ACT_RELU>(g1_s, W5_s, b5_s, g2_s, G1, G2, tid);
            __syncthreads();
            gemm_fwd<TM, RB, ACT_SIGMOID>(g2_s, W6_s, b6_s, xh_s, G2, DP, tid);
            __syncthreads();
            // ---- element-wise: outputs, loss terms, dL/d(pre-sigmoid) ----
            {
                float* __restrict__ xo = a.xhat[br];
                const float* __restrict__ x = a.x;
                const void* __restrict__ m0 = a.mask[0];
                const void* __restrict__ m1 = a.mask[1];
                const float* __restrict__ dxh = a.d_xhat[br];
                struct E { float xv, m, mp; };
                tile_elems<TM, 7, E>(DP, row0, a.B, tid,
                    [&](int d, int r, bool ok) {
                        E e{0.f, 0.f, 0.f};
                        if (ok && d < D) {
                            const long gi = (long)(row0 + r) * D + d;
                            if (lossy) {
                                e.xv = x[gi];
                                e.m = load_mask(m0, gi, a.mask_kind) != 0.f ? 1.f : 0.f;
                                if (a.nbr > 1) e.mp = load_mask(m1, gi, a.mask_kind) != 0.f ? 1.f : 0.f;
                            } else if (a.mode == PCVAE_DEC_BWD) {
                                e.xv = dxh[gi];
                            }
                        }
                        return e;
                    },
                    [&](int d, int r, bool ok, E e) {
                        const float xh = xh_s[d * P + r];
                        float dpre = 0.f;
                        if (ok && d < D) {
                            if (xo) xo[(long)(row0 + r) * D + d] = xh;
                            if (lossy) {
                                const float diff = e.xv - xh;
                                const float nll = fmaf(diff * diff, inv2var, log_scale);
                                float coef;
                                if (br == 0) {
                                    s_req += e.m * nll;
                                    s_red += e.m * (1.f - e.mp) * nll;
                                    s_imp += (1.f - e.m) * nll;
                                    s_sse += (1.f - e.m) * diff * diff;
                                    coef = (1.f - alpha) * e.m + alpha * e.m * (1.f - e.mp);
                                } else {
                                    s_rep += e.mp * nll;
                                    coef = alpha * e.mp;
                                }
                                dpre = coef * (xh - e.xv) * inv_var * ls * xh * (1.f - xh);
                            } else if (a.mode == PCVAE_DEC_BWD) {
                                dpre = e.xv * xh * (1.f - xh);
                            }
                        }
                        if (bwd) xh_s[d * P + r] = dpre;
                    });
            }
            if (bwd) {
                __syncthreads();
                gemm_dw<TM>(g2_s, xh_s, dW6_s, G2, D, DP, tid);
                bias_dw<TM>(xh_s, db6_s, D, tid);
                __syncthreads();
                gemm_dx<TM, RB, true>(xh_s, W6_s, g2_s, G2, DP, tid);
                __syncthreads();
                gemm_dw<TM, 2>(g1_s, g2_s, dW5_s, G1, G2, G2, tid);
                bias_dw<TM>(g2_s, db5_s, G2, tid);
                __syncthreads();
                gemm_dx<TM, RB, true, 2>(g2_s, W5_s, g1_s, G1, G2, tid);
                __syncthreads();
                gemm_dw<TM, 2>(z_s, g1_s, dW4_s, LAT, G1, G1P, tid);
                bias_dw<TM>(g1_s, db4_s, G1, tid);
                __syncthreads();
                gemm_dx<TM, RB, false, 2>(g1_s, W4_s, z_s, LAT, G1P, tid);   // z_s <- dL/dz
                __syncthreads();
            }
            // ---- latent-space terms: KL sums, d_mean / d_logvar, d_z ----
            if (a.mode == PCVAE_DEC_BWD) {
                for (int i = tid; i < TM * LAT; i += NT) {
                    const int r = i / LAT, l = i - r * LAT;
                    if (row0 + r < a.B) a.d_z[br][(long)(row0 + r) * LAT + l] = z_s[l * P + r];
                }
            } else if (lossy) {
                for (int i = tid; i < TM * LAT; i += NT) {
                    const int r = i / LAT, l = i - r * LAT;
                    if (row0 + r >= a.B) continue;
                    const long gi = (long)(row0 + r) * LAT + l;
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
                    if (a.mode == PCVAE_DEC_TRAIN) {
                        const float dz = z_s[l * P + r];
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
            __syncthreads();
        }
    }

    if (lossy) {
        float v[PCVAE_NSUMS] = {s_req, s_rep, s_klq, s_klp, s_klr, s_red, s_imp, s_sse};
#pragma unroll
        for (int j = 0; j < PCVAE_NSUMS; ++j) {
            const float w = warp_sum(v[j]);
            if ((tid & 31) == 0) red_s[tid >> 5][j] = w;
        }
        __syncthreads();
        if (tid < PCVAE_NSUMS) {
            float s = 0.f;
            for (int w = 0; w < NWARP; ++w) s += red_s[w][tid];
            a.sums_partials[blockIdx.x * PCVAE_NSUMS + tid] = s;
        }
    }
    if (bwd) {
        float* gp = a.gp + (long)blockIdx.x * a.L.total;
        flush_linear_grad(dW4_s, db4_s, gp + a.L.W4, gp + a.L.b4, LAT, G1, G1P, tid);
        flush_linear_grad(dW5_s, db5_s, gp + a.L.W5, gp + a.L.b5, G1, G2, G2, tid);
        flush_linear_grad(dW6_s, db6_s, gp + a.L.W6, gp + a.L.b6, G2, D, DP, tid);
    }
}

// ====================================================================================
// Stand-alone loss terms (module API)
// ====================================================================================
struct LossArgs {
    int B, D, Lat, nbr, mask_kind;
    const float* x;
    const void* mask[2];
    const float* xhat[2];
    const float* mean[2];
    const float* logvar[2];
    float alpha, beta_w, x_logvar, loss_scale;
    float* sums_partials;
    float* d_xhat[2];
    float* d_mean[2];
    float* d_logvar[2];
};

__global__ void __launch_bounds__(NT) k_loss_terms(const LossArgs a) {
    __shared__ float red_s[NWARP][PCVAE_NSUMS];
    const int tid = threadIdx.x;
    const float scale = expf(a.x_logvar * 0.5f);
    const float var = scale * scale;
    const float inv2var = 1.0f / (2.0f * var), inv_var = 1.0f / var, log_scale = logf(scale);
    const float alpha = a.alpha, ls = a.loss_scale;
    float v[PCVAE_NSUMS] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long nel = (long)a.B * a.D;
    for (long i = (long)blockIdx.x * NT + tid; i < nel; i += (long)gridDim.x * NT) {
        const float xv = a.x[i];
        const float m = load_mask(a.mask[0], i, a.mask_kind) != 0.f ? 1.f : 0.f;
        const float mp = (a.nbr > 1) ? (load_mask(a.mask[1], i, a.mask_kind) != 0.f ? 1.f : 0.f) : 0.f;
        {
            const float xh = a.xhat[0][i];
            const float diff = xv - xh;
            const float nll = fmaf(diff * diff, inv2var, log_scale);
            v[PCVAE_S_RE_Q] += m * nll;
            v[PCVAE_S_RE_D] += m * (1.f - mp) * nll;
            v[PCVAE_S_RE_IMP] += (1.f - m) * nll;
            v[PCVAE_S_SSE_UNOBS] += (1.f - m) * diff * diff;
            if (a.d_xhat[0]) a.d_xhat[0][i] = ((1.f - alpha) * m + alpha * m * (1.f - mp)) * (xh - xv) * inv_var * ls;
        }
        if (a.nbr > 1) {
            const float xh = a.xhat[1][i];
            const float diff = xv - xh;
            v[PCVAE_S_RE_P] += mp * fmaf(diff * diff, inv2var, log_scale);
            if (a.d_xhat[1]) a.d_xhat[1][i] = alpha * mp * (xh - xv) * inv_var * ls;
        }
    }
    const long nl = (long)a.B * a.Lat;
    for (long i = (long)blockIdx.x * NT + tid; i < nl; i += (long)gridDim.x * NT) {
        const float mq = a.mean[0][i], lq = a.logvar[0][i], eq = expf(lq);
        v[PCVAE_S_KL_Q] += 0.5f * (eq + mq * mq - 1.f - lq);
        float gmq = (1.f - alpha) * a.beta_w * mq, gvq = (1.f - alpha) * a.beta_w * 0.5f * (eq - 1.f);
        if (a.nbr > 1) {
            const float mp_ = a.mean[1][i], lp = a.logvar[1][i], ep = expf(lp);
            const float dmu = mq - mp_, er = expf(lq - lp);
            v[PCVAE_S_KL_P] += 0.5f * (ep + mp_ * mp_ - 1.f - lp);
            v[PCVAE_S_KL_REG] += 0.5f * (er + dmu * dmu / ep - 1.f - (lq - lp));
            gmq += alpha * dmu / ep;
            gvq += alpha * 0.5f * (er - 1.f);
            if (a.d_mean[1]) {
                a.d_mean[1][i] = (alpha * a.beta_w * mp_ - alpha * dmu / ep) * ls;
                a.d_logvar[1][i] = (alpha * a.beta_w * 0.5f * (ep - 1.f) + alpha * 0.5f * (1.f - (eq + dmu * dmu) / ep)) * ls;
            }
        }
        if (a.d_mean[0]) { a.d_mean[0][i] = gmq * ls; a.d_logvar[0][i] = gvq * ls; }
    }
#pragma unroll
    for (int j = 0; j < PCVAE_NSUMS; ++j) {
        const float w = warp_sum(v[j]);
        if ((tid & 31) == 0) red_s[tid >> 5][j] = w;
    }
    __syncthreads();
    if (tid < PCVAE_NSUMS) {
        float s = 0.f;
        for (int w = 0; w < NWARP; ++w) s += red_s[w][tid];
        a.sums_partials[blockIdx.x * PCVAE_NSUMS + tid] = s;
    }
}

// ====================================================================================
// Reductions and Adam
// ====================================================================================
__global__ void k_reduce_sums(const float* __restrict__ sp, int grid, double nll_const, double* __restrict__ sums) {
    const int j = threadIdx.x;
    if (j >= PCVAE_NSUMS) return;
    double s = 0.0;
    for (int c = 0; c < grid; ++c) s += (double)sp[c * PCVAE_NSUMS + j];
    if (j == PCVAE_S_RE_Q || j == PCVAE_S_RE_P || j == PCVAE_S_RE_D || j == PCVAE_S_RE_IMP) s += nll_const;
    sums[j] = s;
}

__global__ void k_reduce_grads(const float* __restrict__ gp, int grid, long P, long begin, long end,
                               float* __restrict__ grad, int accumulate) {
    for (long i = begin + (long)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (long)gridDim.x * blockDim.x) {
        float s = 0.f;
        int c = 0;
        for (; c + 8 <= grid; c += 8) {      // eight loads in flight, summed in CTA order
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = gp[(long)(c + j) * P + i];
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j];
        }
        for (; c < grid; ++c) s += gp[(long)c * P + i];
        grad[i] = accumulate ? grad[i] + s : s;
    }
}

__global__ void k_adam(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                       float* __restrict__ v, long n, float lr_bc1, float inv_sqrt_bc2, float b1, float b2, float eps) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float g = grad[i];
        const float mi = m[i] + (g - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1-beta1)
        const float vi = fmaf(b2, v[i], (1.f - b2) * g * g);      // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        theta[i] -= lr_bc1 * (mi / denom);
    }
}

// reduce + Adam (+ loss sums) in one launch: a block owns 32 consecutive parameters; its eight warps sum interleaved
// eighths of the per-CTA partials (128-byte coalesced loads, always in the same order), warp 0 combines the eight
// sums in a fixed order and applies torch.optim.Adam; block 0 also reduces the loss sums.
// DEV: the 1-based Adam step is *state + 1 (device counter of completed steps, so that a CUDA graph can replay the
// launch); lr_bc1 then carries the plain learning rate and the bias corrections are computed here.  The last block to
// finish (ticket state[1]) advances the counter: every block has read it by then.
template <bool DEV>
__global__ void __launch_bounds__(256) k_reduce_adam(const float* __restrict__ gp, int grid, long P, float* __restrict__ grad,
                                                     float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                                                     float lr_bc1, float inv_sqrt_bc2, float b1, float b2, float eps,
                                                     const float* __restrict__ sp, double nll_const, double* __restrict__ sums,
                                                     unsigned long long* __restrict__ state) {
    __shared__ float red[8][32];
    __shared__ float bc_s[2];
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    // launched as a programmatic dependent of the weight-gradient kernel (launch_tc): its CTAs may be resident before the
    // partials are complete; returns at once after an ordinary launch
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (DEV) {
        if (threadIdx.x == 0) {
            const double step = (double)(*reinterpret_cast<volatile unsigned long long*>(state) + 1ull);
            const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
            bc_s[0] = (float)((double)lr_bc1 / bc1);
            bc_s[1] = (float)(1.0 / sqrt(bc2));
        }
    }
    const long i = (long)blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < P) {
        int c = part;
        for (; c + 24 < grid; c += 32) {                // four loads in flight per thread, summed in CTA order
            const float v0 = gp[(long)c * P + i], v1 = gp[(long)(c + 8) * P + i], v2 = gp[(long)(c + 16) * P + i],
                        v3 = gp[(long)(c + 24) * P + i];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; c < grid; c += 8) s += gp[(long)c * P + i];
    }
    red[part][lane] = s;
    __syncthreads();
    if (DEV) { lr_bc1 = bc_s[0]; inv_sqrt_bc2 = bc_s[1]; }
    if (part == 0 && i < P) {
        float g = red[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) g += red[q][lane];
        grad[i] = g;
        const float mi = m[i] + (g - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1-beta1)
        const float vi = fmaf(b2, v[i], (1.f - b2) * g * g);      // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        m[i] = mi;
        v[i] = vi;
        theta[i] -= lr_bc1 * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    }
    if (DEV) {
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned* ticket = reinterpret_cast<unsigned*>(state + 1);
            if (atomicAdd(ticket, 1u) == gridDim.x - 1) {         // last block: every block has read the counter
                *ticket = 0u;
                __threadfence();
                *reinterpret_cast<volatile unsigned long long*>(state) = *reinterpret_cast<volatile unsigned long long*>(state) + 1ull;
            }
        }
    }
    if (sp && blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + PCVAE_NSUMS) {
        const int j = threadIdx.x - 32;
        double a = 0.0;
        for (int c = 0; c < grid; ++c) a += (double)sp[c * PCVAE_NSUMS + j];
        if (j == PCVAE_S_RE_Q || j == PCVAE_S_RE_P || j == PCVAE_S_RE_D || j == PCVAE_S_RE_IMP) a += nll_const;
        sums[j] = a;
        if (DEV) sums[PCVAE_NSUMS + j] += a;              // running totals since the caller last zeroed them
    }
}

__global__ void k_ffma_probe(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace pcvae

// ========================================================================================
// C ABI
// ========================================================================================
using namespace pcvae;

static int g_train_tc = 1;

// block `which` (0 enc fwd, 1 enc bwd, 2 dec fwd, 3 dec bwd) of a caller's weight-image buffer, or nullptr
static const float* img_block(const Layout& L, const float* images, int which) {
    if (!images) return nullptr;
    WeightImages w;
    Layout Lenc;
    bool enc, dec;
    if (!pcvae::weight_images_plan(L, &w, &Lenc, &enc, &dec)) return nullptr;
    if (which < 2 && !enc) return nullptr;
    if (which >= 2 && !dec) return nullptr;
    const long off[4] = {w.enc_fwd, w.enc_bwd, w.dec_fwd, w.dec_bwd};
    return images + off[which];
}

// which blocks exist for this model and where they start (every block is a multiple of 4 floats)
bool pcvae::weight_images_plan(const Layout& L, WeightImages* w, Layout* Lenc, bool* enc, bool* dec) {
    *enc = false;
    *Lenc = L;
    if (pnp_tc_supported(L)) { *Lenc = pnp_tail_layout(L); *enc = true; }
    else if (enc_tc_supported(L)) *enc = true;
    *dec = dec_tc_supported(L);
    long o = 0;
    w->enc_fwd = o; if (*enc) o += tc::enc_fwd_image_floats(Lenc->D);
    w->enc_bwd = o; if (*enc) o += tc::enc_bwd_image_floats();
    w->dec_fwd = o; if (*dec) o += tc::dec_fwd_image_floats(L.D);
    w->dec_bwd = o; if (*dec) o += tc::dec_bwd_image_floats();
    w->total = o;
    return *enc || *dec;
}

static size_t enc_fwd_smem(const Layout& L) {
    const int P = TM_TRAIN + 4, K4 = round4(L.K);
    const int in1 = L.fam == PCVAE_FAMILY_MLP ? (L.aug ? 2 * L.D : L.D) : L.K;
    const int inr = L.fam == PCVAE_FAMILY_MLP ? in1 : L.D;
    size_t f = (size_t)in1 * H1 + H1 + H1 * H2P + H2P + H2 * LAT2 + LAT2 + (size_t)(inr + H1 + H2P + LAT2) * P;
    if (L.fam == PCVAE_FAMILY_PNP) f += (size_t)L.D * P + 2 * L.D * K4 + K4 * P;
    return f * sizeof(float);
}

static size_t enc_bwd_smem(const Layout& L) {
    const int P = TM_TRAIN + 4, K4 = round4(L.K);
    const int in1 = L.fam == PCVAE_FAMILY_MLP ? (L.aug ? 2 * L.D : L.D) : L.K;
    const int inr = L.fam == PCVAE_FAMILY_MLP ? in1 : L.D;
    size_t f = (size_t)H1 * H2P + H2 * LAT2 + ((size_t)in1 * H1 + H1 + H1 * H2P + H2P + H2 * LAT2 + LAT2) +
               (size_t)(inr + H1 + H2P + LAT2) * P;
    if (L.fam == PCVAE_FAMILY_PNP) f += (size_t)L.D * P + K4 * P + L.K * H1 + 4 * L.D * K4;
    return f * sizeof(float);
}

static size_t dec_smem(const Layout& L, bool bwd) {
    const int P = TM_TRAIN + 4, DP = round4(L.D);
    size_t w = (size_t)LAT * G1P + G1P + G1 * G2 + G2 + G2 * DP + DP;
    size_t f = w + (size_t)(LATP + G1P + G2 + DP) * P + (bwd ? w : 0);
    return f * sizeof(float);
}

template <typename Kern, typename Args>
static int launch(Kern kern, size_t smem, int grid, cudaStream_t st, const char* name, const Args& args) {
    if (smem > MAX_SMEM) return fail(PCVAE_EINVAL, "%s: needs %zu B shared memory (> %d): obs_dim too large", name, smem, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    kern<<<grid, NT, smem, st>>>(args);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: launch: %s", name, cudaGetErrorString(e));
    return PCVAE_OK;
}

extern "C" {

int pcvae_abi_version(void) { return PCVAE_ABI_VERSION; }
const char* pcvae_last_error(void) { return err_buf(); }

long pcvae_param_count(const pcvae_model* m) {
    Layout L;
    return make_layout(m, &L) ? L.total : -1;
}

int pcvae_param_offsets(const pcvae_model* m, long* offsets) {
    Layout L;
    if (!make_layout(m, &L)) return -1;
    int n = 0;
    if (L.fam == PCVAE_FAMILY_PNP) { offsets[n++] = L.E; offsets[n++] = L.bE; offsets[n++] = L.We; offsets[n++] = L.be; }
    const int rest[12] = {L.W1, L.b1, L.W2, L.b2, L.W3, L.b3, L.W4, L.b4, L.W5, L.b5, L.W6, L.b6};
    for (int i = 0; i < 12; ++i) offsets[n++] = rest[i];
    offsets[n] = L.total;
    return n;
}

long pcvae_decoder_offset(const pcvae_model* m) {
    Layout L;
    return make_layout(m, &L) ? L.W4 : -1;
}

int pcvae_grid_ctas(void) {
    int g = 0;
    return device_ok(&g) == PCVAE_OK ? g : -1;
}

size_t pcvae_enc_act_ws_floats(const pcvae_model* m, int rows, int n_branch) {
    if (!m || rows < 0) return 0;
    const long ntiles = (rows + TM_TRAIN - 1) / TM_TRAIN;
    return (size_t)ntiles * n_branch * enc_act_feats(m->family, m->emb_dim) * TM_TRAIN;
}

int pcvae_enc_fwd(const pcvae_enc_fwd_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "enc_fwd: null params");
    Layout L;
    if (!make_layout(&p->model, &L)) return PCVAE_EINVAL;
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 0 || p->n_branch < 1 || p->n_branch > 2) return fail(PCVAE_EINVAL, "enc_fwd: bad rows/n_branch");
    if (p->rows == 0) return PCVAE_OK;
    if (!p->theta || !p->x) return fail(PCVAE_EINVAL, "enc_fwd: null theta/x");
    for (int b = 0; b < p->n_branch; ++b)
        if (!p->mask[b] || !p->mean[b] || !p->logvar[b]) return fail(PCVAE_EINVAL, "enc_fwd: null mask/mean/logvar for branch %d", b);
    cudaStream_t st = (cudaStream_t)stream;
    EncFwdArgs a{};
    a.L = L; a.B = p->rows; a.nbr = p->n_branch; a.mask_kind = p->mask_kind; a.theta = p->theta; a.x = p->x;
    for (int b = 0; b < 2; ++b) { a.mask[b] = p->mask[b]; a.eps[b] = p->eps[b]; a.mean[b] = p->mean[b]; a.logvar[b] = p->logvar[b]; a.z[b] = p->z[b]; }
    a.act_ws = p->act_ws; a.ac = p->pnp_ac;
    if (g_train_tc && p->tc_workspace && pnp_tc_supported(L)) {
        // pooled embedding on the CUDA cores, MLP tail on the tensor cores (pcvae_pnp_tc.cu)
        const long base = etw_floats(p->rows, p->n_branch), need = base + pnp_tc_extra_floats(L, p->rows, p->n_branch);
        if (p->tc_workspace_floats < need)
            return fail(PCVAE_EINVAL, "enc_fwd: tc_workspace has %ld floats, needs %ld", p->tc_workspace_floats, need);
        if (!p->pnp_ac) return fail(PCVAE_EINVAL, "enc_fwd: PNP family needs pnp_ac workspace");
        enc_tc_carve(p->tc_workspace, p->rows, p->n_branch, &a.tw);
        a.wimg = img_block(L, p->weight_images, 0);
        pnp_tables_launch(L, p->theta, p->pnp_ac, st);
        return pnp_enc_fwd_tc_launch(a, p->tc_workspace + base, grid, st);
    }
    if (g_train_tc && p->tc_workspace && enc_tc_supported(L)) {
        const long need = etw_floats(p->rows, p->n_branch);
        if (p->tc_workspace_floats < need)
            return fail(PCVAE_EINVAL, "enc_fwd: tc_workspace has %ld floats, needs %ld", p->tc_workspace_floats, need);
        enc_tc_carve(p->tc_workspace, p->rows, p->n_branch, &a.tw);
        a.wimg = img_block(L, p->weight_images, 0);
        return enc_fwd_tc_launch(a, grid, st);
    }
    if (L.fam == PCVAE_FAMILY_PNP) {
        if (!p->pnp_ac) return fail(PCVAE_EINVAL, "enc_fwd: PNP family needs pnp_ac workspace");
        pnp_tables_launch(L, p->theta, p->pnp_ac, st);
        return launch(k_enc_fwd<PCVAE_FAMILY_PNP, TM_TRAIN>, enc_fwd_smem(L), grid, st, "enc_fwd", a);
    }
    return launch(k_enc_fwd<PCVAE_FAMILY_MLP, TM_TRAIN>, enc_fwd_smem(L), grid, st, "enc_fwd", a);
}

int pcvae_enc_bwd(const pcvae_enc_bwd_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "enc_bwd: null params");
    Layout L;
    if (!make_layout(&p->model, &L)) return PCVAE_EINVAL;
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 0 || p->n_branch < 1 || p->n_branch > 2) return fail(PCVAE_EINVAL, "enc_bwd: bad rows/n_branch");
    if (!p->theta || !p->grad_partials) return fail(PCVAE_EINVAL, "enc_bwd: null theta/grad_partials");
    const bool use_pnp_tc = g_train_tc && p->tc_workspace && p->rows > 0 && pnp_tc_supported(L);
    const bool use_tc = use_pnp_tc || (g_train_tc && p->tc_workspace && p->rows > 0 && enc_tc_supported(L));
    if (p->rows > 0 && (!p->x || (!p->act_ws && !use_tc))) return fail(PCVAE_EINVAL, "enc_bwd: null x/act_ws");
    for (int b = 0; b < p->n_branch && p->rows > 0; ++b)
        if (!p->mask[b] || !p->d_mean[b] || !p->d_logvar[b]) return fail(PCVAE_EINVAL, "enc_bwd: null mask/d_mean/d_logvar for branch %d", b);
    cudaStream_t st = (cudaStream_t)stream;
    EncBwdArgs a{};
    a.L = L; a.B = p->rows; a.nbr = p->n_branch; a.mask_kind = p->mask_kind; a.theta = p->theta; a.x = p->x;
    for (int b = 0; b < 2; ++b) {
        a.mask[b] = p->mask[b]; a.d_mean[b] = p->d_mean[b]; a.d_logvar[b] = p->d_logvar[b];
        a.d_z[b] = p->d_z[b]; a.eps[b] = p->eps[b]; a.logvar[b] = p->logvar[b];
        if (p->d_z[b] && p->eps[b] && !p->logvar[b]) return fail(PCVAE_EINVAL, "enc_bwd: d_z and eps given without logvar for branch %d", b);
    }
    a.act_ws = p->act_ws; a.ac = p->pnp_ac; a.gp = p->grad_partials;
    if (use_pnp_tc) {
        const long base = etw_floats(p->rows, p->n_branch), need = base + pnp_tc_extra_floats(L, p->rows, p->n_branch);
        if (p->tc_workspace_floats < need)
            return fail(PCVAE_EINVAL, "enc_bwd: tc_workspace has %ld floats, needs %ld", p->tc_workspace_floats, need);
        if (!p->pnp_ac) return fail(PCVAE_EINVAL, "enc_bwd: PNP family needs pnp_ac tables");
        enc_tc_carve(p->tc_workspace, p->rows, p->n_branch, &a.tw);
        a.wimg = img_block(L, p->weight_images, 1);
        return pnp_enc_bwd_tc_launch(a, p->tc_workspace + base, grid, st);
    }
    if (use_tc) {
        const long need = etw_floats(p->rows, p->n_branch);
        if (p->tc_workspace_floats < need)
            return fail(PCVAE_EINVAL, "enc_bwd: tc_workspace has %ld floats, needs %ld", p->tc_workspace_floats, need);
        enc_tc_carve(p->tc_workspace, p->rows, p->n_branch, &a.tw);
        a.wimg = img_block(L, p->weight_images, 1);
        return enc_bwd_tc_launch(a, grid, st);
    }
    if (L.fam == PCVAE_FAMILY_PNP) {
        if (!p->pnp_ac) return fail(PCVAE_EINVAL, "enc_bwd: PNP family needs pnp_ac tables");
        return launch(k_enc_bwd<PCVAE_FAMILY_PNP, TM_TRAIN>, enc_bwd_smem(L), grid, st, "enc_bwd", a);
    }
    return launch(k_enc_bwd<PCVAE_FAMILY_MLP, TM_TRAIN>, enc_bwd_smem(L), grid, st, "enc_bwd", a);
}

int pcvae_dec(const pcvae_dec_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "dec: null params");
    Layout L;
    if (!make_layout(&p->model, &L)) return PCVAE_EINVAL;
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->mode < PCVAE_DEC_FWD || p->mode > PCVAE_DEC_EVAL) return fail(PCVAE_EINVAL, "dec: bad mode %d", p->mode);
    if (p->rows < 0 || p->n_branch < 1 || p->n_branch > 2) return fail(PCVAE_EINVAL, "dec: bad rows/n_branch");
    if (!p->theta) return fail(PCVAE_EINVAL, "dec: null theta");
    const bool bwd = p->mode == PCVAE_DEC_TRAIN || p->mode == PCVAE_DEC_BWD;
    const bool lossy = p->mode == PCVAE_DEC_TRAIN || p->mode == PCVAE_DEC_EVAL;
    if (bwd && !p->grad_partials) return fail(PCVAE_EINVAL, "dec: null grad_partials");
    if (lossy && !p->sums_partials) return fail(PCVAE_EINVAL, "dec: null sums_partials");
    if (p->rows > 0) {
        for (int b = 0; b < p->n_branch; ++b) {
            if (!p->z[b]) return fail(PCVAE_EINVAL, "dec: null z[%d]", b);
            if (p->mode == PCVAE_DEC_FWD && !p->xhat[b]) return fail(PCVAE_EINVAL, "dec: null xhat[%d]", b);
            if (p->mode == PCVAE_DEC_BWD && (!p->d_xhat[b] || !p->d_z[b])) return fail(PCVAE_EINVAL, "dec: null d_xhat/d_z[%d]", b);
            if (lossy && (!p->mask[b] || !p->mean[b] || !p->logvar[b])) return fail(PCVAE_EINVAL, "dec: null mask/mean/logvar[%d]", b);
            if (p->mode == PCVAE_DEC_TRAIN && (!p->d_mean[b] || !p->d_logvar[b])) return fail(PCVAE_EINVAL, "dec: null d_mean/d_logvar[%d]", b);
        }
        if (lossy && !p->x) return fail(PCVAE_EINVAL, "dec: null x");
    }
    DecArgs a{};
    a.L = L; a.mode = p->mode; a.B = p->rows; a.nbr = p->n_branch; a.mask_kind = p->mask_kind; a.theta = p->theta; a.x = p->x;
    for (int b = 0; b < 2; ++b) {
        a.z[b] = p->z[b]; a.xhat[b] = p->xhat[b]; a.mask[b] = p->mask[b]; a.mean[b] = p->mean[b]; a.logvar[b] = p->logvar[b];
        a.eps[b] = p->eps[b]; a.d_mean[b] = p->d_mean[b]; a.d_logvar[b] = p->d_logvar[b]; a.d_xhat[b] = p->d_xhat[b]; a.d_z[b] = p->d_z[b];
    }
    a.alpha = p->alpha; a.beta_w = p->beta_w; a.x_logvar = p->x_logvar; a.loss_scale = p->loss_scale;
    a.sums_partials = p->sums_partials; a.gp = p->grad_partials;
    if (p->mode == PCVAE_DEC_TRAIN && g_train_tc && dec_tc_supported(L) && p->tc_workspace && p->rows > 0) {
        const long need = tcw_floats(p->rows, p->n_branch), nvt = tc_nvt(p->rows, p->n_branch), R2P = nvt * 128;
        if (p->tc_workspace_floats < need)
            return fail(PCVAE_EINVAL, "dec: tc_workspace has %ld floats, needs %ld", p->tc_workspace_floats, need);
        float* w = p->tc_workspace;
        a.nvt = nvt;
        a.ws_zT = w;   w += R2P * TCW_Z;
        a.ws_h4T = w;  w += R2P * TCW_H4;
        a.ws_h5T = w;  w += R2P * TCW_H5;
        a.ws_dp6T = w; w += R2P * TCW_H5;
        a.ws_dp5T = w; w += R2P * TCW_H5;
        a.ws_dp4T = w; w += R2P * TCW_H4;
        a.ws_relu = reinterpret_cast<unsigned*>(w);
        a.wimg_fwd = img_block(L, p->weight_images, 2);
        a.wimg_bwd = img_block(L, p->weight_images, 3);
        return dec_tc_launch(a, grid, (cudaStream_t)stream);
    }
    return launch(k_dec<TM_TRAIN>, dec_smem(L, bwd), grid, (cudaStream_t)stream, "dec", a);
}

long pcvae_dec_tc_workspace_floats(const pcvae_model* m, int rows, int n_branch) {
    Layout L;
    if (!m || !make_layout(m, &L) || !dec_tc_supported(L) || rows < 0 || n_branch < 1) return 0;
    return tcw_floats(rows, n_branch);
}

long pcvae_enc_tc_workspace_floats(const pcvae_model* m, int rows, int n_branch) {
    Layout L;
    if (!g_train_tc || !m || !make_layout(m, &L) || rows < 1 || n_branch < 1) return 0;
    if (pnp_tc_supported(L)) return etw_floats(rows, n_branch) + pnp_tc_extra_floats(L, rows, n_branch);
    if (!enc_tc_supported(L)) return 0;
    return etw_floats(rows, n_branch);
}

long pcvae_weight_images_floats(const pcvae_model* m) {
    Layout L, Lenc;
    WeightImages w;
    bool enc, dec;
    if (!g_train_tc || !m || !make_layout(m, &L) || !weight_images_plan(L, &w, &Lenc, &enc, &dec)) return 0;
    return w.total;
}

int pcvae_build_weight_images(const pcvae_model* m, const float* theta, float* images, void* stream) {
    if (!m || !theta || !images) return fail(PCVAE_EINVAL, "build_weight_images: null pointer");
    Layout L, Lenc;
    if (!make_layout(m, &L)) return PCVAE_EINVAL;
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    WeightImages w;
    bool enc, dec;
    if (!weight_images_plan(L, &w, &Lenc, &enc, &dec)) return fail(PCVAE_EINVAL, "build_weight_images: this model has no tensor-core training kernels");
    return build_weight_images_launch(L, Lenc, enc, dec, w, theta, images, (cudaStream_t)stream);
}

int pcvae_set_train_tensor_cores(int enable) {
    const int prev = g_train_tc;
    g_train_tc = enable ? 1 : 0;
    return prev;
}

int pcvae_loss_terms(const pcvae_loss_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "loss_terms: null params");
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 0 || p->obs_dim < 1 || p->latent_dim < 1 || p->n_branch < 1 || p->n_branch > 2)
        return fail(PCVAE_EINVAL, "loss_terms: bad sizes");
    if (!p->sums_partials) return fail(PCVAE_EINVAL, "loss_terms: null sums_partials");
    if (p->rows > 0) {
        if (!p->x) return fail(PCVAE_EINVAL, "loss_terms: null x");
        for (int b = 0; b < p->n_branch; ++b)
            if (!p->mask[b] || !p->xhat[b] || !p->mean[b] || !p->logvar[b]) return fail(PCVAE_EINVAL, "loss_terms: null input for branch %d", b);
    }
    LossArgs a{};
    a.B = p->rows; a.D = p->obs_dim; a.Lat = p->latent_dim; a.nbr = p->n_branch; a.mask_kind = p->mask_kind; a.x = p->x;
    for (int b = 0; b < 2; ++b) {
        a.mask[b] = p->mask[b]; a.xhat[b] = p->xhat[b]; a.mean[b] = p->mean[b]; a.logvar[b] = p->logvar[b];
        a.d_xhat[b] = p->d_xhat[b]; a.d_mean[b] = p->d_mean[b]; a.d_logvar[b] = p->d_logvar[b];
    }
    a.alpha = p->alpha; a.beta_w = p->beta_w; a.x_logvar = p->x_logvar; a.loss_scale = p->loss_scale;
    a.sums_partials = p->sums_partials;
    k_loss_terms<<<grid, NT, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "loss_terms: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_reduce_sums(const float* sums_partials, int grid, int rows, int obs_dim, double* sums, void* stream) {
    if (!sums_partials || !sums || grid < 1) return fail(PCVAE_EINVAL, "reduce_sums: bad arguments");
    const double c = 0.5 * 1.8378770664093453 * (double)rows * (double)obs_dim;   // 0.5*log(2*pi) per entry
    k_reduce_sums<<<1, 32, 0, (cudaStream_t)stream>>>(sums_partials, grid, c, sums);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reduce_sums: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_reduce_grads(const float* grad_partials, int grid, long param_count, long begin, long end, float* grad,
                       int accumulate, void* stream) {
    if (!grad_partials || !grad || grid < 1 || begin < 0 || end > param_count || begin > end)
        return fail(PCVAE_EINVAL, "reduce_grads: bad arguments");
    if (begin == end) return PCVAE_OK;
    const int blocks = (int)((end - begin + 255) / 256);
    k_reduce_grads<<<blocks, 256, 0, (cudaStream_t)stream>>>(grad_partials, grid, param_count, begin, end, grad, accumulate);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reduce_grads: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int step, float lr,
                    float beta1, float beta2, float eps, void* stream) {
    if (!theta || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) return fail(PCVAE_EINVAL, "adam_step: bad arguments");
    if (n == 0) return PCVAE_OK;
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const int blocks = (int)((n + 255) / 256);
    k_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(theta, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1),
                                                      (float)(1.0 / sqrt(bc2)), beta1, beta2, eps);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "adam_step: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_reduce_adam(const float* grad_partials, int grid, long param_count, float* grad, float* theta, float* exp_avg,
                      float* exp_avg_sq, int step, float lr, float beta1, float beta2, float eps, const float* sums_partials,
                      int rows, int obs_dim, double* sums, void* stream) {
    if (!grad_partials || !grad || !theta || !exp_avg || !exp_avg_sq || grid < 1 || param_count < 1 || step < 1)
        return fail(PCVAE_EINVAL, "reduce_adam: bad arguments");
    if ((sums_partials == nullptr) != (sums == nullptr)) return fail(PCVAE_EINVAL, "reduce_adam: sums_partials and sums go together");
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const double c = 0.5 * 1.8378770664093453 * (double)rows * (double)obs_dim;   // 0.5*log(2*pi) per entry
    const int blocks = (int)((param_count + 31) / 32);
    cudaError_t e = launch_tc(k_reduce_adam<false>, blocks, 256, 0, (cudaStream_t)stream, true, grad_partials, grid, param_count, grad,
                              theta, exp_avg, exp_avg_sq, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps,
                              sums_partials, c, sums, (unsigned long long*)nullptr);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reduce_adam: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_reduce_adam_dev(const float* grad_partials, int grid, long param_count, float* grad, float* theta, float* exp_avg,
                          float* exp_avg_sq, unsigned long long* step_state, float lr, float beta1, float beta2, float eps,
                          const float* sums_partials, int rows, int obs_dim, double* sums, void* stream) {
    if (!grad_partials || !grad || !theta || !exp_avg || !exp_avg_sq || grid < 1 || param_count < 1 || !step_state)
        return fail(PCVAE_EINVAL, "reduce_adam_dev: bad arguments");
    if ((sums_partials == nullptr) != (sums == nullptr)) return fail(PCVAE_EINVAL, "reduce_adam_dev: sums_partials and sums go together");
    const double c = 0.5 * 1.8378770664093453 * (double)rows * (double)obs_dim;
    const int blocks = (int)((param_count + 31) / 32);
    cudaError_t e = launch_tc(k_reduce_adam<true>, blocks, 256, 0, (cudaStream_t)stream, true, grad_partials, grid, param_count, grad,
                              theta, exp_avg, exp_avg_sq, lr, 1.0f, beta1, beta2, eps, sums_partials, c, sums, step_state);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reduce_adam_dev: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_ffma_probe(float* scratch, int iters, double* flops, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (!scratch || iters < 1) return fail(PCVAE_EINVAL, "ffma_probe: bad arguments");
    const int blocks = grid * 4, threads = 512;
    k_ffma_probe<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "ffma_probe: launch: %s", cudaGetErrorString(e));
    if (flops) *flops = 2.0 * 8.0 * 16.0 * (double)iters * blocks * threads;
    return PCVAE_OK;
}

}  // extern "C"
