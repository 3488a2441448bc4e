// MIWAE / Reg_MIWAE (Student-t decoder + importance-weighted bound), reference src/models/VAE.py:3011-3134, 3137-3301
// (SURVEY.md section 8f item 4).  The 128-wide ReLU layers run on the generic dense kernels (pcvae_dense.cu); this file
// holds what is specific to the family:
//   * the output heads of encoder and decoder (VAE.py:3047-3049, 3061-3066) and their backward, from the RAW layer
//     output so that softplus' / sigmoid' are exact;
//   * z = mean + scale * eps over S samples per row (VAE.py:3054-3056) and its backward;
//   * the loss (VAE.py:3068-3110, 3197-3263): Student-t log-likelihood per (row, sample, feature), the reference's
//     un-transposed [B*S] -> [S, B] reshape of the per-(row, sample) likelihoods, log p(z) - log q(z|x) of a
//     loss-internal draw, logsumexp over samples, the KL / likelihood regularisers of Reg_MIWAE, the importance-weighted
//     imputation of llh_eval, and every gradient in closed form (oracle/pcvae_oracle.py: miwae_loss_closed_form_grads,
//     reg_miwae_loss_closed_form_grads are the specification, checked against autograd of the reference formulas).
// Element-wise / reduction kernels (HBM-bound); all reductions run in a fixed order -> deterministic.
#include "pcvae_internal.cuh"
#include "pcvae_special.cuh"

namespace pcvae {

constexpr float HALF_LOG_PI_F = 0.57236494292470008707f;

__device__ __forceinline__ float mw_softplus(float v) { return v > 20.f ? v : log1pf(expf(v)); }   // nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float mw_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }

// ---- heads -----------------------------------------------------------------------------------------
// ENC: raw [R][2W] -> out0 = raw[:, :W] (mean), out1 = softplus(raw[:, W:]) (scale)
// DEC: raw [R][3W] -> out0 = sigmoid, out1 = softplus + 0.001, out2 = softplus + 3
__global__ void k_miwae_heads(const float* __restrict__ raw, long R, int W, int mode, float* __restrict__ o0,
                              float* __restrict__ o1, float* __restrict__ o2) {
    const int C = mode == PCVAE_MIWAE_HEADS_ENC ? 2 : 3;
    const long n = R * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long r = i / ((long)W * C);
        const int c = (int)(i - r * W * C), chunk = c / W, w = c - chunk * W;
        const float v = raw[i];
        const long o = r * W + w;
        if (mode == PCVAE_MIWAE_HEADS_ENC) {
            if (chunk == 0) o0[o] = v; else o1[o] = mw_softplus(v);
        } else {
            if (chunk == 0) o0[o] = mw_sigmoid(v);
            else if (chunk == 1) o1[o] = mw_softplus(v) + 0.001f;
            else o2[o] = mw_softplus(v) + 3.0f;
        }
    }
}

// d_raw from the gradients of the head outputs (a null d* counts as zero); softplus'(v) = sigmoid(v), 1 above the threshold
__global__ void k_miwae_heads_bwd(const float* __restrict__ raw, long R, int W, int mode, const float* __restrict__ d0,
                                  const float* __restrict__ d1, const float* __restrict__ d2, float* __restrict__ d_raw) {
    const int C = mode == PCVAE_MIWAE_HEADS_ENC ? 2 : 3;
    const long n = R * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long r = i / ((long)W * C);
        const int c = (int)(i - r * W * C), chunk = c / W, w = c - chunk * W;
        const float v = raw[i];
        const long o = r * W + w;
        const float* d = chunk == 0 ? d0 : (chunk == 1 ? d1 : d2);
        const float g = d ? d[o] : 0.f;
        float f;
        if (chunk == 0) {
            if (mode == PCVAE_MIWAE_HEADS_ENC) f = 1.f;
            else { const float s = mw_sigmoid(v); f = s * (1.f - s); }
        } else {
            f = v > 20.f ? 1.f : mw_sigmoid(v);
        }
        d_raw[i] = g * f;
    }
}

// ---- latent sampling -------------------------------------------------------------------------------
__global__ void k_miwae_sample_z(const float* __restrict__ mean, const float* __restrict__ scale, const float* __restrict__ eps,
                                 float* __restrict__ z, long n, int S, int L) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long b = i / ((long)S * L);
        const int l = (int)(i % L);
        z[i] = eps ? fmaf(eps[i], scale[b * L + l], mean[b * L + l]) : mean[b * L + l];
    }
}

__global__ void k_miwae_sample_z_bwd(const float* __restrict__ dz, const float* __restrict__ eps, float* __restrict__ d_mean,
                                     float* __restrict__ d_scale, int B, int S, int L) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * L; i += gridDim.x * blockDim.x) {
        const int b = i / L, l = i - b * L;
        float sm = 0.f, ss = 0.f;
        for (int s = 0; s < S; ++s) {
            const long j = ((long)b * S + s) * L + l;
            const float g = dz[j];
            sm += g;
            if (eps) ss = fmaf(g, eps[j], ss);
        }
        d_mean[i] = sm;
        d_scale[i] = ss;
    }
}

// ---- loss ------------------------------------------------------------------------------------------
struct MiwaeArgs {
    int B, S, D, L, reg, mask_kind, rowwise;
    const float* x;
    const void* mask[2];             // [0] = mask (q branch), [1] = mask_p (p branch)
    const float* xm[2];
    const float* xs[2];
    const float* df[2];
    const float* mean[2];
    const float* scale[2];
    const float* eps2[2];
    float alpha;
    float* lpx;       // [2][3][B*S]: masked log-likelihood of the branch; q only: on mask & ~mask_p; q only: on ~mask
    float* lw;        // [2][B*S]   lw[j*S + i], then the softmax weights
    float* glpx;      // [2][B*S]   dloss / d(lpx of the branch), flat (b, s) order
    float* bstat;     // [2][B][2]  lse of column j; KL(q || p) of row j (slot [0][j][1])
    double* out;
    float* xm_imp;
    float* d_xm[2];
    float* d_xs[2];
    float* d_df[2];
    float* d_mean[2];
    float* d_scale[2];
};

__device__ __forceinline__ float mw_mask(const void* m, int kind, long i) {
    if (kind == PCVAE_MASK_U8) return static_cast<const unsigned char*>(m)[i] ? 1.f : 0.f;
    return static_cast<const float*>(m)[i] != 0.f ? 1.f : 0.f;
}

// StudentT(df, loc, scale).log_prob(x) as torch writes it (distributions/studentT.py)
__device__ __forceinline__ float student_t_logp(float x, float loc, float sc, float df, float* y_out, float* A_out) {
    const float y = (x - loc) / sc;
    const float q = y * y / df;
    const float Z = logf(sc) + 0.5f * logf(df) + HALF_LOG_PI_F + lgammaf(0.5f * df) - lgammaf(0.5f * (df + 1.f));
    *y_out = y;
    *A_out = 1.f + q;
    return -0.5f * (df + 1.f) * log1pf(q) - Z;
}

__device__ __forceinline__ float warp_sum_fixed(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per (branch, row, sample): the three masked sums of the Student-t log-likelihood over the features
__global__ void __launch_bounds__(256) k_miwae_rows(const MiwaeArgs a) {
    const int lane = threadIdx.x & 31;
    const long BS = (long)a.B * a.S, nrow = BS * (a.reg ? 2 : 1);
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long it = wid; it < nrow; it += nw) {
        const int br = (int)(it / BS);
        const long row = it - (long)br * BS;
        const int b = (int)(row / a.S);
        float s_main = 0.f, s_reg = 0.f, s_imp = 0.f;
        for (int d = lane; d < a.D; d += 32) {
            const long xi = (long)b * a.D + d, j = row * a.D + d;
            float y, A;
            const float lp = student_t_logp(a.x[xi], a.xm[br][j], a.xs[br][j], a.df[br][j], &y, &A);
            const float m0 = mw_mask(a.mask[0], a.mask_kind, xi);
            if (br == 0) {
                s_main = fmaf(lp, m0, s_main);
                s_imp = fmaf(lp, 1.f - m0, s_imp);
                if (a.reg) s_reg = fmaf(lp, m0 * (1.f - mw_mask(a.mask[1], a.mask_kind, xi)), s_reg);
            } else {
                s_main = fmaf(lp, mw_mask(a.mask[1], a.mask_kind, xi), s_main);
            }
        }
        s_main = warp_sum_fixed(s_main); s_reg = warp_sum_fixed(s_reg); s_imp = warp_sum_fixed(s_imp);
        if (lane == 0) {
            float* o = a.lpx + (long)br * 3 * BS;
            o[row] = s_main; o[BS + row] = s_reg; o[2 * BS + row] = s_imp;
        }
    }
}

constexpr int MW_NT = 128, MW_MAXL = 16;

// one block per (branch, column j): logsumexp over the S entries of column j of the "[samples, rows]" matrix
//   lw[i][j] = lpx_flat[rowwise ? j*S + i : i*B + j] + log p(z_ji) - log q(z_ji | x_j),   z_ji = mean_j + scale_j * eps2_ji
// (VAE.py:3078-3092: the likelihood matrix is the row-major [B*S] vector viewed as [S, B] WITHOUT a transpose, the
// prior / posterior terms are transposed properly), softmax weights, gradients of the bound with respect to lpx, mean and
// scale, the KL regulariser's gradient, and the importance-weighted imputation of row j.
__global__ void __launch_bounds__(MW_NT) k_miwae_cols(const MiwaeArgs a) {
    __shared__ float red[MW_NT / 32][2 * MW_MAXL + 2];
    __shared__ float mu_s[MW_MAXL], sc_s[MW_MAXL], lsc_s[MW_MAXL];
    __shared__ float bc[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbr = a.reg ? 2 : 1;
    const long BS = (long)a.B * a.S;
    for (int item = blockIdx.x; item < nbr * a.B; item += gridDim.x) {
        const int br = item / a.B, j = item - br * a.B;
        const int L = a.L, S = a.S;
        __syncthreads();
        if (tid < L) {
            mu_s[tid] = a.mean[br][(long)j * L + tid];
            sc_s[tid] = a.scale[br][(long)j * L + tid];
            lsc_s[tid] = logf(sc_s[tid]);
        }
        __syncthreads();
        const float* lpx = a.lpx + (long)br * 3 * BS;
        float* lw = a.lw + (long)br * BS + (long)j * S;
        const float* e2 = a.eps2[br] + (long)j * S * L;
        // pass 1: lw and its maximum
        float mx = -INFINITY;
        for (int i = tid; i < S; i += MW_NT) {
            float t = 0.f;
            for (int l = 0; l < L; ++l) {
                const float e = e2[(long)i * L + l];
                const float z = fmaf(sc_s[l], e, mu_s[l]);
                t += -0.5f * z * z + 0.5f * e * e + lsc_s[l];
            }
            const long k = a.rowwise ? (long)j * S + i : (long)i * a.B + j;
            const float v = lpx[k] + t;
            lw[i] = v;
            mx = fmaxf(mx, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp][0] = mx;
        __syncthreads();
        if (tid == 0) { float m = red[0][0]; for (int w = 1; w < MW_NT / 32; ++w) m = fmaxf(m, red[w][0]); bc[0] = m; }
        __syncthreads();
        mx = bc[0];
        // pass 2: sum of exponentials
        float se = 0.f;
        for (int i = tid; i < S; i += MW_NT) se += expf(lw[i] - mx);
        se = warp_sum_fixed(se);
        __syncthreads();
        if (lane == 0) red[warp][0] = se;
        __syncthreads();
        if (tid == 0) {
            float s = red[0][0];
            for (int w = 1; w < MW_NT / 32; ++w) s += red[w][0];
            bc[1] = mx + logf(s);
            a.bstat[((long)br * a.B + j) * 2] = bc[1];
        }
        __syncthreads();
        const float lse = bc[1];
        // pass 3: weights, gradient of the bound: dloss/dlw[i][j] = -coef_br * w / B
        const float coef = a.reg ? (br == 0 ? 1.f - a.alpha : a.alpha) : 1.f;
        const float gscale = -coef / (float)a.B;
        float gm[MW_MAXL], gs[MW_MAXL];
#pragma unroll
        for (int l = 0; l < MW_MAXL; ++l) { gm[l] = 0.f; gs[l] = 0.f; }
        const bool want = a.d_mean[br] != nullptr;
        for (int i = tid; i < S; i += MW_NT) {
            const float w = expf(lw[i] - lse);
            lw[i] = w;
            if (want) {
                const float g = gscale * w;
                const long k = a.rowwise ? (long)j * S + i : (long)i * a.B + j;
                a.glpx[(long)br * BS + k] = g;
#pragma unroll
                for (int l = 0; l < MW_MAXL; ++l)
                    if (l < L) {
                        const float e = e2[(long)i * L + l];
                        const float z = fmaf(sc_s[l], e, mu_s[l]);
                        gm[l] = fmaf(g, -z, gm[l]);
                        gs[l] = fmaf(g, -z * e + 1.f / sc_s[l], gs[l]);
                    }
            }
        }
        if (want) {
#pragma unroll
            for (int l = 0; l < MW_MAXL; ++l)
                if (l < L) {
                    const float m = warp_sum_fixed(gm[l]), s = warp_sum_fixed(gs[l]);
                    if (lane == 0) { red[warp][2 + l] = m; red[warp][2 + MW_MAXL + l] = s; }
                }
            __syncthreads();
            if (tid < L) {
                float m = 0.f, s = 0.f;
                for (int w = 0; w < MW_NT / 32; ++w) { m += red[w][2 + tid]; s += red[w][2 + MW_MAXL + tid]; }
                if (a.reg) {
                    // KL(N(mean_q, scale_q) || N(mean_p, scale_p)).mean() over [B, L] (VAE.py:3246, 3265-3270), weight alpha
                    const float mq = a.mean[0][(long)j * L + tid], sq = a.scale[0][(long)j * L + tid];
                    const float mp = a.mean[1][(long)j * L + tid], sp = a.scale[1][(long)j * L + tid];
                    const float dm = mq - mp, n = (float)a.B * (float)L;
                    if (br == 0) {
                        m += a.alpha * (dm / (sp * sp) / n);
                        s += a.alpha * ((-1.f / sq + sq / (sp * sp)) / n);
                    } else {
                        m += a.alpha * (-dm / (sp * sp) / n);
                        s += a.alpha * ((1.f / sp - (sq * sq + dm * dm) / (sp * sp * sp)) / n);
                    }
                }
                a.d_mean[br][(long)j * L + tid] = m;
                a.d_scale[br][(long)j * L + tid] = s;
            }
        }
        if (a.reg && br == 0 && tid == 0) {                 // KL of row j, torch's kl_divergence(Normal, Normal) form
            float kl = 0.f;
            for (int l = 0; l < L; ++l) {
                const float mq = a.mean[0][(long)j * L + l], sq = a.scale[0][(long)j * L + l];
                const float mp = a.mean[1][(long)j * L + l], sp = a.scale[1][(long)j * L + l];
                const float vr = (sq / sp) * (sq / sp), t1 = ((mq - mp) / sp) * ((mq - mp) / sp);
                kl += 0.5f * (vr + t1 - 1.f - logf(vr));
            }
            a.bstat[((long)0 * a.B + j) * 2 + 1] = kl;
        }
        // importance-weighted imputation of row j (q branch): xm_imp[j][d] = sum_i w[i] * xm[j][i][d]  (VAE.py:3097-3099)
        if (br == 0 && a.xm_imp) {
            __syncthreads();                                 // all weights of the column are in lw
            const float* xm = a.xm[0] + (long)j * S * a.D;
            for (int d = tid; d < a.D; d += MW_NT) {
                float acc = 0.f;
                for (int i = 0; i < S; ++i) acc = fmaf(lw[i], xm[(long)i * a.D + d], acc);
                a.xm_imp[(long)j * a.D + d] = acc;
            }
        }
    }
}

// scalars (one block): neg_bound of each branch, KL_reg, reg_like, the imputed-likelihood scalar of MIWAE's llh_eval, loss
__global__ void __launch_bounds__(256) k_miwae_finish(const MiwaeArgs a) {
    __shared__ double red[256];
    const int tid = threadIdx.x;
    const long BS = (long)a.B * a.S;
    auto block_sum = [&](double v) -> double {
        red[tid] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        const double r = red[0];
        __syncthreads();
        return r;
    };
    double v = 0.0;
    for (int j = tid; j < a.B; j += 256) v += (double)a.bstat[((long)0 * a.B + j) * 2];
    const double nb_q = -block_sum(v) / a.B;
    double nb_p = 0.0, kl = 0.0, reg_like = 0.0;
    if (a.reg) {
        v = 0.0;
        for (int j = tid; j < a.B; j += 256) v += (double)a.bstat[((long)1 * a.B + j) * 2];
        nb_p = -block_sum(v) / a.B;
        v = 0.0;
        for (int j = tid; j < a.B; j += 256) v += (double)a.bstat[((long)0 * a.B + j) * 2 + 1];
        kl = block_sum(v) / ((double)a.B * a.L);
        v = 0.0;
        for (long r = tid; r < BS; r += 256) v += (double)a.lpx[BS + r];
        reg_like = block_sum(v) / (double)BS;
    }
    v = 0.0;
    for (long r = tid; r < BS; r += 256) v += (double)a.lpx[2 * BS + r];
    const double imp = block_sum(v) / ((double)a.B * 5000.0);          // VAE.py:3100: "/ (x.shape[0] * 5000)"
    if (tid == 0) {
        const double al = a.alpha;
        a.out[0] = a.reg ? nb_q + al * (kl - nb_q + nb_p - reg_like) : nb_q;
        a.out[1] = nb_q; a.out[2] = nb_p; a.out[3] = kl; a.out[4] = reg_like; a.out[5] = imp;
    }
}

// gradients with respect to the decoder heads' outputs: one thread per (branch, row, sample, feature)
__global__ void __launch_bounds__(256) k_miwae_grads(const MiwaeArgs a) {
    const long BS = (long)a.B * a.S, per = BS * a.D, n = per * (a.reg ? 2 : 1);
    const float greg = a.reg ? -a.alpha / (float)BS : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int br = (int)(i / per);
        const long j = i - (long)br * per;
        const long row = j / a.D;
        const int d = (int)(j - row * a.D), b = (int)(row / a.S);
        const long xi = (long)b * a.D + d;
        const float m0 = mw_mask(a.mask[0], a.mask_kind, xi);
        float g;
        if (br == 0) {
            g = a.glpx[row] * m0;
            if (a.reg) g = fmaf(greg, m0 * (1.f - mw_mask(a.mask[1], a.mask_kind, xi)), g);
        } else {
            g = a.glpx[BS + row] * mw_mask(a.mask[1], a.mask_kind, xi);
        }
        float o_m = 0.f, o_s = 0.f, o_d = 0.f;
        if (g != 0.f) {
            const float xs = a.xs[br][j], df = a.df[br][j];
            const float y = (a.x[xi] - a.xm[br][j]) / xs;
            const float A = 1.f + y * y / df;
            const float c = (df + 1.f) * y / (xs * df * A);
            o_m = g * c;
            o_s = g * (c * y - 1.f / xs);
            o_d = g * (-0.5f * log1pf(y * y / df) + 0.5f * (df + 1.f) * y * y / (df * df * A)
                       - (0.5f / df - 0.5f * digamma_half_step(df)));
        }
        a.d_xm[br][j] = o_m;
        a.d_xs[br][j] = o_s;
        a.d_df[br][j] = o_d;
    }
}

static int ew_blocks(long n, int grid) {
    long b = (n + 255) / 256;
    const long cap = (long)grid * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pcvae

using namespace pcvae;

extern "C" {

int pcvae_miwae_heads(const float* raw, long rows, int width, int mode, float* out0, float* out1, float* out2, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (mode != PCVAE_MIWAE_HEADS_ENC && mode != PCVAE_MIWAE_HEADS_DEC) return fail(PCVAE_EINVAL, "miwae_heads: unknown mode %d", mode);
    if (rows < 0 || width < 1) return fail(PCVAE_EINVAL, "miwae_heads: bad sizes");
    if (rows == 0) return PCVAE_OK;
    if (!raw || !out0 || !out1 || (mode == PCVAE_MIWAE_HEADS_DEC && !out2)) return fail(PCVAE_EINVAL, "miwae_heads: null pointer");
    const long n = rows * width * (mode == PCVAE_MIWAE_HEADS_ENC ? 2 : 3);
    k_miwae_heads<<<ew_blocks(n, grid), 256, 0, (cudaStream_t)stream>>>(raw, rows, width, mode, out0, out1, out2);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "miwae_heads: launch: %s", cudaGetErrorString(e));
}

int pcvae_miwae_heads_bwd(const float* raw, long rows, int width, int mode, const float* d0, const float* d1, const float* d2,
                          float* d_raw, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (mode != PCVAE_MIWAE_HEADS_ENC && mode != PCVAE_MIWAE_HEADS_DEC) return fail(PCVAE_EINVAL, "miwae_heads_bwd: unknown mode %d", mode);
    if (rows < 0 || width < 1) return fail(PCVAE_EINVAL, "miwae_heads_bwd: bad sizes");
    if (rows == 0) return PCVAE_OK;
    if (!raw || !d_raw) return fail(PCVAE_EINVAL, "miwae_heads_bwd: null pointer");
    const long n = rows * width * (mode == PCVAE_MIWAE_HEADS_ENC ? 2 : 3);
    k_miwae_heads_bwd<<<ew_blocks(n, grid), 256, 0, (cudaStream_t)stream>>>(raw, rows, width, mode, d0, d1, d2, d_raw);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "miwae_heads_bwd: launch: %s", cudaGetErrorString(e));
}

int pcvae_miwae_sample_z(const float* mean, const float* scale, const float* eps, float* z, int rows, int samples, int latent_dim,
                         void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 0 || samples < 1 || latent_dim < 1) return fail(PCVAE_EINVAL, "miwae_sample_z: bad sizes");
    if (rows == 0) return PCVAE_OK;
    if (!mean || !scale || !z) return fail(PCVAE_EINVAL, "miwae_sample_z: null pointer");
    const long n = (long)rows * samples * latent_dim;
    k_miwae_sample_z<<<ew_blocks(n, grid), 256, 0, (cudaStream_t)stream>>>(mean, scale, eps, z, n, samples, latent_dim);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "miwae_sample_z: launch: %s", cudaGetErrorString(e));
}

int pcvae_miwae_sample_z_bwd(const float* d_z, const float* eps, float* d_mean, float* d_scale, int rows, int samples,
                             int latent_dim, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 0 || samples < 1 || latent_dim < 1) return fail(PCVAE_EINVAL, "miwae_sample_z_bwd: bad sizes");
    if (rows == 0) return PCVAE_OK;
    if (!d_z || !d_mean || !d_scale) return fail(PCVAE_EINVAL, "miwae_sample_z_bwd: null pointer");
    k_miwae_sample_z_bwd<<<ew_blocks((long)rows * latent_dim, grid), 256, 0, (cudaStream_t)stream>>>(d_z, eps, d_mean, d_scale, rows,
                                                                                                 samples, latent_dim);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "miwae_sample_z_bwd: launch: %s", cudaGetErrorString(e));
}

size_t pcvae_miwae_loss_workspace_bytes(int rows, int samples) {
    if (rows < 1 || samples < 1) return 0;
    const size_t BS = (size_t)rows * samples;
    return (10 * BS + 4 * (size_t)rows) * sizeof(float);
}

int pcvae_miwae_loss(const pcvae_miwae_loss_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "miwae_loss: null params");
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 1 || p->samples < 1 || p->obs_dim < 1 || p->latent_dim < 1 || p->latent_dim > MW_MAXL)
        return fail(PCVAE_EINVAL, "miwae_loss: sizes (rows >= 1, samples >= 1, latent_dim 1..%d)", MW_MAXL);
    if (p->mask_kind != PCVAE_MASK_U8 && p->mask_kind != PCVAE_MASK_F32) return fail(PCVAE_EINVAL, "miwae_loss: unknown mask kind");
    const int nbr = p->regularised ? 2 : 1;
    if (!p->x || !p->mask || (p->regularised && !p->mask_p) || !p->out || !p->workspace) return fail(PCVAE_EINVAL, "miwae_loss: null pointer");
    if (p->workspace_bytes < pcvae_miwae_loss_workspace_bytes(p->rows, p->samples)) return fail(PCVAE_EINVAL, "miwae_loss: workspace too small");
    const bool grads = p->d_xm[0] != nullptr;
    MiwaeArgs a{};
    a.B = p->rows; a.S = p->samples; a.D = p->obs_dim; a.L = p->latent_dim; a.reg = p->regularised ? 1 : 0;
    a.mask_kind = p->mask_kind; a.rowwise = p->rowwise ? 1 : 0;
    a.x = p->x; a.mask[0] = p->mask; a.mask[1] = p->mask_p; a.alpha = p->regularised ? p->alpha : 0.f;
    for (int br = 0; br < nbr; ++br) {
        if (!p->xm[br] || !p->xs[br] || !p->df[br] || !p->mean[br] || !p->scale[br] || !p->eps2[br])
            return fail(PCVAE_EINVAL, "miwae_loss: branch %d has a null input", br);
        a.xm[br] = p->xm[br]; a.xs[br] = p->xs[br]; a.df[br] = p->df[br]; a.mean[br] = p->mean[br]; a.scale[br] = p->scale[br];
        a.eps2[br] = p->eps2[br];
        if (grads && (!p->d_xm[br] || !p->d_xs[br] || !p->d_df[br] || !p->d_mean[br] || !p->d_scale[br]))
            return fail(PCVAE_EINVAL, "miwae_loss: give every gradient buffer of branch %d or none", br);
        if (grads) { a.d_xm[br] = p->d_xm[br]; a.d_xs[br] = p->d_xs[br]; a.d_df[br] = p->d_df[br]; a.d_mean[br] = p->d_mean[br]; a.d_scale[br] = p->d_scale[br]; }
    }
    const size_t BS = (size_t)a.B * a.S;
    float* w = static_cast<float*>(p->workspace);
    a.lpx = w; w += 6 * BS;
    a.lw = w; w += 2 * BS;
    a.glpx = w; w += 2 * BS;
    a.bstat = w;
    a.out = p->out; a.xm_imp = p->xm_imputed;
    cudaStream_t st = (cudaStream_t)stream;
    const long nrow = (long)BS * nbr;
    long rb = (nrow * 32 + 255) / 256;
    if (rb > (long)grid * 8) rb = (long)grid * 8;
    k_miwae_rows<<<(int)rb, 256, 0, st>>>(a);
    int cb = nbr * a.B;
    if (cb > grid * 8) cb = grid * 8;
    k_miwae_cols<<<cb, MW_NT, 0, st>>>(a);
    k_miwae_finish<<<1, 256, 0, st>>>(a);
    if (grads) k_miwae_grads<<<ew_blocks(nrow * a.D, grid), 256, 0, st>>>(a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "miwae_loss: launch: %s", cudaGetErrorString(e));
}

}  // extern "C"
