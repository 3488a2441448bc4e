// not-MIWAE MNAR importance-sampling pieces: latent sampling over S samples per row and the
// REG_notMIWAE_v2 / notMIWAE_myversion loss with its gradients (src/models/VAE.py:2377-2505,
// 2748-2847).  Element-wise / reduction kernels (HBM-bound); the 128-wide dense layers are in
// pcvae_dense.cu.  All reductions run in a fixed order -> deterministic.
#include "pcvae_internal.cuh"

namespace pcvae {

constexpr float HALF_LOG_2PI_F = 0.91893853320467274178f;

__global__ void k_mnar_sample_z(const float* __restrict__ mean, const float* __restrict__ logvar,
                                const float* __restrict__ eps, float* __restrict__ z, long n, int S, int L) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long b = i / ((long)S * L);
        const int l = (int)(i % L);
        const float mu = mean[b * L + l], lv = logvar[b * L + l];
        z[i] = eps ? fmaf(eps[i], expf(lv * 0.5f), mu) : mu;
    }
}

__global__ void k_mnar_sample_z_bwd(const float* __restrict__ dz, const float* __restrict__ logvar,
                                    const float* __restrict__ eps, float* __restrict__ d_mean,
                                    float* __restrict__ d_logvar, int B, int S, int L) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * L; i += gridDim.x * blockDim.x) {
        const int b = i / L, l = i - b * L;
        const float hs = 0.5f * expf(logvar[i] * 0.5f);
        float sm = 0.f, sv = 0.f;
        for (int s = 0; s < S; ++s) {
            const long j = ((long)b * S + s) * L + l;
            const float g = dz[j];
            sm += g;
            if (eps) sv = fmaf(g * hs, eps[j], sv);
        }
        d_mean[i] = sm;
        d_logvar[i] = sv;
    }
}

struct MnarArgs {
    int B, S, D, L, reg;
    const float* x;
    const float* mask;
    const float* mask_p;
    const float* xm[2];
    const float* xlv[2];
    const float* mean[2];
    const float* logvar[2];
    const float* eps_kl;
    const float* W;
    const float* b;
    float alpha;
    float* rowv;      // [5][B*S]: lw_q, lw_p, re_q, re_d, (scratch)
    float* bstat;     // [B][8]: lse_q, lse_p, lse_neg, sum_re_q, sum_re_d, kl_reg_row
    double* out;
    float* xm_imp;
    float* d_xm[2];
    float* d_xlv[2];
    float* d_mean[2];
    float* d_logvar[2];
    float* dWb_part;  // [grid][2][128]
    float* d_W;
    float* d_b;
};

__device__ __forceinline__ float softplusf(float v) { return v > 20.f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// one warp per (row, sample): masked Gaussian NLL sums, Bernoulli self-masking log-prob, KL -> l_w
__global__ void __launch_bounds__(256) k_mnar_rows(const MnarArgs a) {
    const int lane = threadIdx.x & 31;
    const long nrow = (long)a.B * a.S;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long row = wid; row < nrow; row += nw) {
        const int b = (int)(row / a.S);
        float re_q = 0.f, re_d = 0.f, re_p = 0.f, logp = 0.f;
        for (int d = lane; d < a.D; d += 32) {
            const float x = a.x[(long)b * a.D + d], m = a.mask[(long)b * a.D + d];
            const long j = row * a.D + d;
            {
                const float xm = a.xm[0][j], xlv = a.xlv[0][j];
                const float scale = expf(xlv * m * 0.5f);
                const float diff = x * m - xm * m;
                const float nll = diff * diff / (2.f * scale * scale) + logf(scale) + HALF_LOG_2PI_F;
                re_q += nll;
                const float mixed = xm * (1.f - m) + x * m;
                const float lg = -softplusf(a.W[d]) * (mixed - a.b[d]);
                logp -= fmaxf(lg, 0.f) - lg * m + log1pf(expf(-fabsf(lg)));     // -BCEWithLogits(lg, m)
                if (a.reg) {
                    const float md = m * (1.f - a.mask_p[(long)b * a.D + d]);
                    const float sc = expf(xlv * md * 0.5f);
                    const float df = x * md - xm * md;
                    re_d += df * df / (2.f * sc * sc) + logf(sc) + HALF_LOG_2PI_F;
                }
            }
            if (a.reg) {
                const float mp = a.mask_p[(long)b * a.D + d];
                const float xm = a.xm[1][j], xlv = a.xlv[1][j];
                const float scale = expf(xlv * mp * 0.5f);
                const float diff = x * mp - xm * mp;
                re_p += diff * diff / (2.f * scale * scale) + logf(scale) + HALF_LOG_2PI_F;
            }
        }
        // KL terms over the latent dimension
        float kl_q = 0.f, kl_p = 0.f;
        for (int l = lane; l < a.L; l += 32) {
            const float mu = a.mean[0][b * a.L + l], lv = a.logvar[0][b * a.L + l];
            if (a.reg) {
                kl_q += 0.5f * (expf(lv) + mu * mu - 1.f - lv);
                const float mp = a.mean[1][b * a.L + l], lp = a.logvar[1][b * a.L + l];
                kl_p += 0.5f * (expf(lp) + mp * mp - 1.f - lp);
            } else {
                // Monte-Carlo KL from the second draw z' (VAE.py:2791-2798): log q(z'|x) - log p(z')
                const float e = a.eps_kl[row * a.L + l];
                const float std = expf(lv * 0.5f);
                const float z2 = fmaf(e, std, mu);
                const float dq = z2 - mu;
                const float log_q = -(dq * dq) / (2.f * std * std) - logf(std) - HALF_LOG_2PI_F;
                const float log_p = -(z2 * z2) * 0.5f - HALF_LOG_2PI_F;
                kl_q += log_q - log_p;
            }
        }
        re_q = warp_sum(re_q); re_d = warp_sum(re_d); re_p = warp_sum(re_p); logp = warp_sum(logp);
        kl_q = warp_sum(kl_q); kl_p = warp_sum(kl_p);
        if (lane == 0) {
            a.rowv[row] = re_q + kl_q - logp;
            a.rowv[nrow + row] = re_p + kl_p;
            a.rowv[2 * nrow + row] = re_q;
            a.rowv[3 * nrow + row] = re_d;
        }
    }
}

// one CTA per row b: logsumexp over the S samples of l_w_q, l_w_p and -l_w_q; per-row sums
__global__ void __launch_bounds__(256) k_mnar_lse(const MnarArgs a) {
    __shared__ float red[8][6];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long nrow = (long)a.B * a.S;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const float* lq = a.rowv + (long)b * a.S;
        const float* lp = a.rowv + nrow + (long)b * a.S;
        const float* rq = a.rowv + 2 * nrow + (long)b * a.S;
        const float* rd = a.rowv + 3 * nrow + (long)b * a.S;
        float mq = -INFINITY, mp = -INFINITY, mn = -INFINITY;
        for (int s = tid; s < a.S; s += 256) { mq = fmaxf(mq, lq[s]); mp = fmaxf(mp, lp[s]); mn = fmaxf(mn, -lq[s]); }
        for (int o = 16; o > 0; o >>= 1) {
            mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, o));
            mp = fmaxf(mp, __shfl_xor_sync(0xffffffffu, mp, o));
            mn = fmaxf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) { red[w][0] = mq; red[w][1] = mp; red[w][2] = mn; }
        __syncthreads();
        mq = red[0][0]; mp = red[0][1]; mn = red[0][2];
        for (int i = 1; i < 8; ++i) { mq = fmaxf(mq, red[i][0]); mp = fmaxf(mp, red[i][1]); mn = fmaxf(mn, red[i][2]); }
        __syncthreads();
        float sq = 0.f, sp = 0.f, sn = 0.f, srq = 0.f, srd = 0.f;
        for (int s = tid; s < a.S; s += 256) {
            sq += expf(lq[s] - mq); sp += expf(lp[s] - mp); sn += expf(-lq[s] - mn);
            srq += rq[s]; srd += rd[s];
        }
        sq = warp_sum(sq); sp = warp_sum(sp); sn = warp_sum(sn); srq = warp_sum(srq); srd = warp_sum(srd);
        if (lane == 0) { red[w][0] = sq; red[w][1] = sp; red[w][2] = sn; red[w][3] = srq; red[w][4] = srd; }
        __syncthreads();
        if (tid == 0) {
            float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 5; ++j) t[j] += red[i][j];
            float* bs = a.bstat + (long)b * 8;
            bs[0] = mq + logf(t[0]);
            bs[1] = mp + logf(t[1]);
            bs[2] = mn + logf(t[2]);
            bs[3] = t[3];
            bs[4] = t[4];
            float klr = 0.f;
            if (a.reg)
                for (int l = 0; l < a.L; ++l) {
                    const float uq = a.mean[0][b * a.L + l], vq = a.logvar[0][b * a.L + l];
                    const float up = a.mean[1][b * a.L + l], vp = a.logvar[1][b * a.L + l];
                    const float dm = uq - up;
                    klr += 0.5f * (expf(vq - vp) + dm * dm / expf(vp) - 1.f - (vq - vp));
                }
            bs[5] = klr;
        }
        __syncthreads();
    }
}

// final scalars (single thread, fixed order, fp64)
__global__ void k_mnar_final(const MnarArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double lq = 0, lp = 0, rq = 0, rd = 0, klr = 0;
    const double logS = log((double)a.S);
    for (int b = 0; b < a.B; ++b) {
        const float* bs = a.bstat + (long)b * 8;
        lq += (double)bs[0] - logS; lp += (double)bs[1] - logS; rq += bs[3]; rd += bs[4]; klr += bs[5];
    }
    const double B = a.B, BS = (double)a.B * a.S;
    const double loss_q = lq / B, loss_p = lp / B;
    double loss = loss_q;
    if (a.reg) loss = loss_q + (double)a.alpha * (klr / (B * a.L) - loss_q + loss_p + rd / BS);
    a.out[0] = loss; a.out[1] = rq / BS; a.out[2] = loss_q; a.out[3] = loss_p;
}

// xm_imputed[b][d] = sum_s softmax_s(-l_w_q) xm_q[b][s][d]      (VAE.py:2458-2461, 2811-2812)
__global__ void k_mnar_impute(const MnarArgs a) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)a.B * a.D; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / a.D), d = (int)(i - (long)b * a.D);
        const float lse = a.bstat[(long)b * 8 + 2];
        float acc = 0.f;
        for (int s = 0; s < a.S; ++s) {
            const long row = (long)b * a.S + s;
            acc = fmaf(expf(-a.rowv[row] - lse), a.xm[0][row * a.D + d], acc);
        }
        a.xm_imp[i] = acc;
    }
}

// gradients w.r.t. the decoder heads and the self-masking parameters; one warp per (row, sample)
__global__ void __launch_bounds__(256) k_mnar_grads(const MnarArgs a) {
    __shared__ float part[8][2][128];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long nrow = (long)a.B * a.S;
    const long wid = ((long)blockIdx.x * blockDim.x + tid) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    const float invB = 1.f / (float)a.B, invBS = 1.f / ((float)a.B * (float)a.S);
    const float cq = a.reg ? (1.f - a.alpha) : 1.f;
    float gW[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
    for (long row = wid; row < nrow; row += nw) {
        const int b = (int)(row / a.S);
        const float wq = cq * expf(a.rowv[row] - a.bstat[(long)b * 8]) * invB;              // dloss/dl_w_q
        const float wp = a.reg ? a.alpha * expf(a.rowv[nrow + row] - a.bstat[(long)b * 8 + 1]) * invB : 0.f;
        const float wd = a.reg ? a.alpha * invBS : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int d = lane + 32 * c;
            if (d >= a.D) continue;
            const float x = a.x[(long)b * a.D + d], m = a.mask[(long)b * a.D + d];
            const long j = row * a.D + d;
            const float xm = a.xm[0][j], xlv = a.xlv[0][j];
            const float scale = expf(xlv * 0.5f), var = scale * scale;
            const float diff = xm - x;
            // NLL(m): d/dxm = m (xm - x)/var ; d/dxlv = m (0.5 - (x - xm)^2 / (2 var))
            float gm = wq * m * diff / var;
            float gv = wq * m * (0.5f - diff * diff / (2.f * var));
            if (a.reg) {
                const float md = m * (1.f - a.mask_p[(long)b * a.D + d]);
                gm += wd * md * diff / var;
                gv += wd * md * (0.5f - diff * diff / (2.f * var));
            }
            // -log p(s|x): d/dlogit = sigmoid(logit) - m ; logit = -softplus(W)(x~ - b), x~ = xm(1-m) + x m
            const float spw = softplusf(a.W[d]);
            const float mixed = xm * (1.f - m) + x * m;
            const float lg = -spw * (mixed - a.b[d]);
            const float dl = wq * (sigmoidf_(lg) - m);
            gm += dl * (-spw) * (1.f - m);
            gW[c] += dl * (-(mixed - a.b[d])) * sigmoidf_(a.W[d]);
            gb[c] += dl * spw;
            a.d_xm[0][j] = gm;
            a.d_xlv[0][j] = gv;
            if (a.reg) {
                const float mp = a.mask_p[(long)b * a.D + d];
                const float xp = a.xm[1][j], xv = a.xlv[1][j];
                const float sc = expf(xv * 0.5f), vr = sc * sc;
                const float df = xp - x;
                a.d_xm[1][j] = wp * mp * df / vr;
                a.d_xlv[1][j] = wp * mp * (0.5f - df * df / (2.f * vr));
            }
        }
    }
    for (int c = 0; c < 4; ++c) {
        const int d = lane + 32 * c;
        part[w][0][d] = gW[c];
        part[w][1][d] = gb[c];
    }
    __syncthreads();
    for (int i = tid; i < 2 * 128; i += 256) {
        const int k = i >> 7, d = i & 127;
        float s = 0.f;
        for (int q = 0; q < 8; ++q) s += part[q][k][d];
        a.dWb_part[((long)blockIdx.x * 2 + k) * 128 + d] = s;
    }
}

// d_mean / d_logvar of both branches [B][L] and the reduction of the dW / db partials
__global__ void k_mnar_latent_grads(const MnarArgs a, int grid_parts) {
    const float invB = 1.f / (float)a.B;
    const float cq = a.reg ? (1.f - a.alpha) : 1.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.B * a.L; i += gridDim.x * blockDim.x) {
        const int b = i / a.L, l = i - b * a.L;
        const float uq = a.mean[0][i], vq = a.logvar[0][i];
        float gm, gv;
        if (a.reg) {
            const float up = a.mean[1][i], vp = a.logvar[1][i];
            const float ep = expf(vp), eq = expf(vq), dm = uq - up;
            const float kr = a.alpha * invB / (float)a.L;                      // KL_reg is a mean over [B,S,L]
            gm = cq * invB * uq + kr * dm / ep;
            gv = cq * invB * 0.5f * (eq - 1.f) + kr * 0.5f * (expf(vq - vp) - 1.f);
            a.d_mean[1][i] = a.alpha * invB * up - kr * dm / ep;
            a.d_logvar[1][i] = a.alpha * invB * 0.5f * (ep - 1.f) + kr * 0.5f * (1.f - (eq + dm * dm) / ep);
        } else {
            // MC KL: KL_s = sum_l [-eps'^2/2 - lv/2 + z'^2/2] (+const); d/dmu = z', d/dlv = -1/2 + z' std eps'/2
            const float std = expf(vq * 0.5f), lse = a.bstat[(long)b * 8];
            gm = 0.f; gv = 0.f;
            for (int s = 0; s < a.S; ++s) {
                const long row = (long)b * a.S + s;
                const float wq = expf(a.rowv[row] - lse) * invB;
                const float e = a.eps_kl[row * a.L + l];
                const float z2 = fmaf(e, std, uq);
                gm = fmaf(wq, z2, gm);
                gv = fmaf(wq, -0.5f + 0.5f * z2 * std * e, gv);
            }
        }
        a.d_mean[0][i] = gm;
        a.d_logvar[0][i] = gv;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * a.D; i += gridDim.x * blockDim.x) {
        const int k = i / a.D, d = i - k * a.D;
        float s = 0.f;
        for (int c = 0; c < grid_parts; ++c) s += a.dWb_part[((long)c * 2 + k) * 128 + d];
        (k == 0 ? a.d_W : a.d_b)[d] = s;
    }
}

struct MnarWs { size_t rowv, bstat, part, total; };
static MnarWs mnar_ws(int B, int S, int grid) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    MnarWs w; size_t o = 0;
    w.rowv = o; o = al(o + (size_t)4 * B * S * 4);
    w.bstat = o; o = al(o + (size_t)B * 8 * 4);
    w.part = o; o = al(o + (size_t)grid * 2 * 128 * 4);
    w.total = o;
    return w;
}
constexpr int MNAR_GRID = 296;   // CTAs of the warp-per-row kernels (2 per SM on a 148-SM part)

}  // namespace pcvae

using namespace pcvae;

extern "C" {

int pcvae_mnar_sample_z(const float* mean, const float* logvar, const float* eps, float* z, int rows, int samples,
                        int latent_dim, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 0 || samples < 1 || latent_dim < 1) return fail(PCVAE_EINVAL, "mnar_sample_z: bad sizes");
    if (rows == 0) return PCVAE_OK;
    if (!mean || !logvar || !z) return fail(PCVAE_EINVAL, "mnar_sample_z: null pointer");
    const long n = (long)rows * samples * latent_dim;
    k_mnar_sample_z<<<grid * 4, 256, 0, (cudaStream_t)stream>>>(mean, logvar, eps, z, n, samples, latent_dim);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "mnar_sample_z: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_mnar_sample_z_bwd(const float* d_z, const float* logvar, const float* eps, float* d_mean, float* d_logvar,
                            int rows, int samples, int latent_dim, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 0 || samples < 1 || latent_dim < 1) return fail(PCVAE_EINVAL, "mnar_sample_z_bwd: bad sizes");
    if (rows == 0) return PCVAE_OK;
    if (!d_z || !logvar || !d_mean || !d_logvar) return fail(PCVAE_EINVAL, "mnar_sample_z_bwd: null pointer");
    k_mnar_sample_z_bwd<<<(rows * latent_dim + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_z, logvar, eps, d_mean,
                                                                                           d_logvar, rows, samples, latent_dim);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "mnar_sample_z_bwd: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

size_t pcvae_mnar_loss_workspace_bytes(int rows, int samples, int obs_dim) {
    (void)obs_dim;
    if (rows < 0 || samples < 1) return 0;
    return mnar_ws(rows, samples, MNAR_GRID).total;
}

int pcvae_mnar_loss(const pcvae_mnar_loss_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "mnar_loss: null params");
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 1 || p->samples < 1 || p->obs_dim < 1 || p->obs_dim > MAX_D || p->latent_dim < 1 || p->latent_dim > 32)
        return fail(PCVAE_EINVAL, "mnar_loss: bad sizes");
    const int nb = p->regularised ? 2 : 1;
    if (!p->x || !p->mask || !p->W || !p->b || !p->out || !p->workspace) return fail(PCVAE_EINVAL, "mnar_loss: null pointer");
    if (p->regularised && !p->mask_p) return fail(PCVAE_EINVAL, "mnar_loss: regularised loss needs mask_p");
    if (!p->regularised && !p->eps_kl) return fail(PCVAE_EINVAL, "mnar_loss: notMIWAE_myversion loss needs eps_kl");
    for (int i = 0; i < nb; ++i)
        if (!p->xm[i] || !p->xlv[i] || !p->mean[i] || !p->logvar[i]) return fail(PCVAE_EINVAL, "mnar_loss: null branch %d input", i);
    const MnarWs w = mnar_ws(p->rows, p->samples, MNAR_GRID);
    if (p->workspace_bytes < w.total) return fail(PCVAE_EWORKSPACE, "mnar_loss: workspace %zu < %zu bytes", p->workspace_bytes, w.total);
    const bool grads = p->d_xm[0] != nullptr;
    if (grads) {
        for (int i = 0; i < nb; ++i)
            if (!p->d_xm[i] || !p->d_xlv[i] || !p->d_mean[i] || !p->d_logvar[i]) return fail(PCVAE_EINVAL, "mnar_loss: null gradient output (branch %d)", i);
        if (!p->d_W || !p->d_b) return fail(PCVAE_EINVAL, "mnar_loss: null d_W/d_b");
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)p->workspace;
    MnarArgs a{};
    a.B = p->rows; a.S = p->samples; a.D = p->obs_dim; a.L = p->latent_dim; a.reg = p->regularised ? 1 : 0;
    a.x = p->x; a.mask = p->mask; a.mask_p = p->mask_p; a.eps_kl = p->eps_kl; a.W = p->W; a.b = p->b; a.alpha = p->alpha;
    for (int i = 0; i < 2; ++i) {
        a.xm[i] = p->xm[i]; a.xlv[i] = p->xlv[i]; a.mean[i] = p->mean[i]; a.logvar[i] = p->logvar[i];
        a.d_xm[i] = p->d_xm[i]; a.d_xlv[i] = p->d_xlv[i]; a.d_mean[i] = p->d_mean[i]; a.d_logvar[i] = p->d_logvar[i];
    }
    a.rowv = (float*)(ws + w.rowv); a.bstat = (float*)(ws + w.bstat); a.dWb_part = (float*)(ws + w.part);
    a.out = p->out; a.xm_imp = p->xm_imputed; a.d_W = p->d_W; a.d_b = p->d_b;
    k_mnar_rows<<<MNAR_GRID, 256, 0, st>>>(a);
    k_mnar_lse<<<min(p->rows, MNAR_GRID), 256, 0, st>>>(a);
    k_mnar_final<<<1, 32, 0, st>>>(a);
    if (p->xm_imputed) k_mnar_impute<<<(p->rows * p->obs_dim + 255) / 256, 256, 0, st>>>(a);
    if (grads) {
        k_mnar_grads<<<MNAR_GRID, 256, 0, st>>>(a);
        k_mnar_latent_grads<<<(p->rows * p->latent_dim + 255) / 256, 256, 0, st>>>(a, MNAR_GRID);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "mnar_loss: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // extern "C"
