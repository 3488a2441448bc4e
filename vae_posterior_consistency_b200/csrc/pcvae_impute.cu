// Importance-weighted MNAR imputation in ONE pass over the [rows, samples] grid: what
// eval_vae_mnar (reference src/experiment_main/evaluate.py:13-69) gets from its row-by-row calls of
// model.forward + model.loss(llh_eval=True) with S = valid_k = 10 000 samples (src/models/VAE.py:2377-2396, 2398-2461
// for REG_notMIWAE_v2; 2748-2770, 2772-2823 for notMIWAE_myversion):
//     z_s = mean + exp(logvar / 2) eps_s;  h = ELU(W2 ELU(W1 z_s + b1) + b2);  xm_s = sigmoid(Wm h + bm),
//     xlv_s = hardtanh_[-10, 0](Wv h + bv);  l_w,s = RE_s + KL_s - log p(mask | x~_s);   x_imputed = sum_s softmax(-l_w)_s xm_s
// The decoder outputs [S, D] of a row are never written to memory: a work item is (row, chunk of samples); the CTA
// keeps the decoder weights in shared memory, runs 64-sample tiles through the three dense layers (register-blocked
// FP32 FFMA, the tile machinery of pcvae_tile.cuh), evaluates the loss terms of the tile and folds it into a running
// (max, sum of exponentials, weighted sum of xm) -- an online softmax.  Chunks of a row are merged by a second,
// tiny kernel in chunk order (deterministic).  Noise: `eps` (and `eps_kl` for the Monte-Carlo KL of
// notMIWAE_myversion) either come from the caller ([rows][S][L], parity mode: drawn on the host in the reference's
// order) or are generated here with Philox (throughput mode: no per-row host draws, no [rows][S][L] buffer).
#include "pcvae_internal.cuh"
#include "pcvae_philox.cuh"

namespace pcvae {

constexpr int IM_TM = 64, IM_P = IM_TM + 4, IM_H = 128, IM_MAXL = 16, IM_MAXD = 64;
constexpr float IM_HALF_LOG_2PI = 0.91893853320467274178f;

struct ImputeArgs {
    int N, S, D, L, reg, chunk, nsplit;
    const float* W1; const float* b1;      // seq_decoder.0  [128][L]
    const float* W2; const float* b2;      // seq_decoder.2  [128][128]
    const float* Wm; const float* bm;      // x_mean.0       [D][128]
    const float* Wv; const float* bv;      // x_logvar.0     [D][128]
    const float* sW; const float* sb;      // self-masking slope (before softplus) and offset [D]
    const float* x; const float* mask;     // [N][D]
    const float* mean; const float* logvar;   // [N][L]
    const float* eps; const float* eps_kl;    // [N][S][L] or null (Philox)
    unsigned long long seed, offset;
    float* part;                           // [N][nsplit][2 + D]
    float* xm_imp;                         // [N][D]
};

__device__ __forceinline__ float im_softplus(float v) { return v > 20.f ? v : log1pf(expf(v)); }

__global__ void __launch_bounds__(NT, 1) k_mnar_impute(const ImputeArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int D = a.D, L = a.L, NP2 = round4(2 * D);
    float* W1_s = smem;                      // [L][128]
    float* b1_s = W1_s + L * IM_H;           // [128]
    float* W2_s = b1_s + IM_H;               // [128][128]
    float* b2_s = W2_s + IM_H * IM_H;        // [128]
    float* Wo_s = b2_s + IM_H;               // [128][NP2]   x_mean rows | x_logvar rows
    float* bo_s = Wo_s + IM_H * NP2;         // [NP2]
    float* z_s = bo_s + NP2;                 // [IM_MAXL][P]  latent tile, feature-major
    float* h1_s = z_s + IM_MAXL * IM_P;      // [128][P]      layer 1, later the head outputs [NP2][P]
    float* h2_s = h1_s + IM_H * IM_P;        // [128][P]
    float* row_s = h2_s + IM_H * IM_P;       // x[D], mask[D], softplus(sW)[D], sb[D]
    float* lat_s = row_s + 4 * IM_MAXD;      // mean[L], std[L], logstd[L]
    float* nl_s = lat_s + 3 * IM_MAXL;       // [TM]  -l_w of the tile, then exp(-l_w - max)
    float* acc_s = nl_s + IM_TM;             // [D]   running weighted sum of xm
    float* st_s = acc_s + IM_MAXD;           // running max, running sum, tile scale
    stage_linear(W1_s, b1_s, a.W1, a.b1, L, IM_H, IM_H, tid);
    stage_linear(W2_s, b2_s, a.W2, a.b2, IM_H, IM_H, IM_H, tid);
    // heads side by side: outputs [0, D) = x_mean rows, [D, 2D) = x_logvar rows
    for (int i = tid; i < IM_H * NP2; i += NT) {
        const int k = i / NP2, n = i - k * NP2;
        Wo_s[i] = n < D ? __ldg(a.Wm + (long)n * IM_H + k) : (n < 2 * D ? __ldg(a.Wv + (long)(n - D) * IM_H + k) : 0.f);
    }
    for (int n = tid; n < NP2; n += NT) bo_s[n] = n < D ? __ldg(a.bm + n) : (n < 2 * D ? __ldg(a.bv + n - D) : 0.f);
    __syncthreads();
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const int nitems = a.N * a.nsplit;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int n = item / a.nsplit, c = item - n * a.nsplit;
        const int s_begin = c * a.chunk, s_end = min(a.S, s_begin + a.chunk);
        __syncthreads();
        if (tid < D) {
            row_s[tid] = a.x[(long)n * D + tid];
            row_s[IM_MAXD + tid] = a.mask[(long)n * D + tid];
            row_s[2 * IM_MAXD + tid] = im_softplus(a.sW[tid]);
            row_s[3 * IM_MAXD + tid] = a.sb[tid];
            acc_s[tid] = 0.f;
        }
        if (tid < L) {
            const float lv = a.logvar[(long)n * L + tid];
            lat_s[tid] = a.mean[(long)n * L + tid];
            lat_s[IM_MAXL + tid] = expf(lv * 0.5f);
            lat_s[2 * IM_MAXL + tid] = lv;
        }
        if (tid == 0) { st_s[0] = -INFINITY; st_s[1] = 0.f; }
        __syncthreads();
        for (int s0 = s_begin; s0 < s_end; s0 += IM_TM) {
            const int valid = min(IM_TM, s_end - s0);
            // ---- latent tile z[l][r] and the KL term of every sample ----
            // regularised: analytic KL(q || N(0, I)), the same for every sample (VAE.py:2431-2432); otherwise the Monte-Carlo
            // KL log q(z'|x) - log p(z') of a SECOND draw z' (VAE.py:2791-2798), kept per sample in nl_s
            if (tid < IM_TM) {
                const int r = tid;
                float kl = 0.f;
                const long so = ((long)n * a.S + s0 + r) * L;
                for (int l = 0; l < L; ++l) {
                    const float mu = lat_s[l], sd = lat_s[IM_MAXL + l], lv = lat_s[2 * IM_MAXL + l];
                    float e = 0.f, e2 = 0.f;
                    if (r < valid) {
                        if (a.eps) {
                            e = a.eps[so + l];
                            if (!a.reg) e2 = a.eps_kl[so + l];
                        } else {
                            const unsigned long long ctr = a.offset + (unsigned long long)(s0 + r);
                            const uint4 rr = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)l, (uint32_t)ctr, (uint32_t)(ctr >> 32)), key);
                            const float2 g = box_muller(rr.x, rr.y);
                            e = g.x; e2 = g.y;
                        }
                    }
                    z_s[l * IM_P + r] = fmaf(e, sd, mu);
                    if (a.reg) {
                        kl += 0.5f * (expf(lv) + mu * mu - 1.f - lv);
                    } else {
                        const float z2 = fmaf(e2, sd, mu), dq = z2 - mu;
                        kl += (-(dq * dq) / (2.f * sd * sd) - logf(sd) - IM_HALF_LOG_2PI) - (-(z2 * z2) * 0.5f - IM_HALF_LOG_2PI);
                    }
                }
                nl_s[r] = kl;
            }
            __syncthreads();
            gemm_fwd<IM_TM, 1, ACT_ELU>(z_s, W1_s, b1_s, h1_s, L, IM_H, tid);
            __syncthreads();
            gemm_fwd<IM_TM, 1, ACT_ELU>(h1_s, W2_s, b2_s, h2_s, IM_H, IM_H, tid);
            __syncthreads();
            gemm_fwd<IM_TM, 1, ACT_NONE>(h2_s, Wo_s, bo_s, h1_s, IM_H, NP2, tid);     // raw heads over h1 (dead)
            __syncthreads();
            // ---- loss terms: 8 threads per sample, features strided by 8; xm written back in place ----
            {
                const int r = tid >> 3, sub = tid & 7;
                float re = 0.f, logp = 0.f;
                for (int d = sub; d < D; d += 8) {
                    const float xm = 1.0f / (1.0f + expf(-h1_s[d * IM_P + r]));
                    const float xlv = fminf(fmaxf(h1_s[(D + d) * IM_P + r], -10.0f), 0.0f);
                    h1_s[d * IM_P + r] = xm;
                    const float x = row_s[d], m = row_s[IM_MAXD + d];
                    const float scale = expf(xlv * m * 0.5f);
                    const float diff = x * m - xm * m;
                    re += diff * diff / (2.f * scale * scale) + logf(scale) + IM_HALF_LOG_2PI;
                    const float mixed = xm * (1.f - m) + x * m;
                    const float lg = -row_s[2 * IM_MAXD + d] * (mixed - row_s[3 * IM_MAXD + d]);
                    logp -= fmaxf(lg, 0.f) - lg * m + log1pf(expf(-fabsf(lg)));
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    re += __shfl_xor_sync(0xffffffffu, re, o);
                    logp += __shfl_xor_sync(0xffffffffu, logp, o);
                }
                if (sub == 0) nl_s[r] = r < valid ? -(re + nl_s[r] - logp) : -INFINITY;
            }
            __syncthreads();
            // ---- online softmax: new maximum, rescale, add the tile ----
            if (tid < 32) {
                float mx = fmaxf(nl_s[lane], nl_s[lane + 32]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float m_old = st_s[0], m_new = fmaxf(m_old, mx);
                const float w0 = expf(nl_s[lane] - m_new), w1 = expf(nl_s[lane + 32] - m_new);
                float sum = w0 + w1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                __syncwarp();
                nl_s[lane] = w0; nl_s[lane + 32] = w1;
                if (lane == 0) {
                    const float sc = m_old == -INFINITY ? 0.f : expf(m_old - m_new);
                    st_s[0] = m_new; st_s[1] = st_s[1] * sc + sum; st_s[2] = sc;
                }
            }
            __syncthreads();
            {
                const int d = tid >> 2, q = tid & 3;             // four threads per feature, 16 samples each
                float acc = 0.f;
                if (d < D) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc = fmaf(nl_s[16 * q + i], h1_s[d * IM_P + 16 * q + i], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);     // every lane takes part (D need not be a multiple of 8)
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if (d < D && q == 0) acc_s[d] = fmaf(acc_s[d], st_s[2], acc);
            }
            __syncthreads();
        }
        float* out = a.part + ((long)n * a.nsplit + c) * (2 + D);
        if (tid == 0) { out[0] = st_s[0]; out[1] = st_s[1]; }
        if (tid < D) out[2 + tid] = acc_s[tid];
    }
}

// chunks of a row merged in chunk order: x_imputed = sum_c acc_c e^{m_c - M} / sum_c l_c e^{m_c - M}
__global__ void k_mnar_impute_merge(const ImputeArgs a) {
    const long total = (long)a.N * a.D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int n = (int)(i / a.D), d = (int)(i - (long)n * a.D);
        const float* p = a.part + (long)n * a.nsplit * (2 + a.D);
        float M = -INFINITY;
        for (int c = 0; c < a.nsplit; ++c) M = fmaxf(M, p[c * (2 + a.D)]);
        float den = 0.f, num = 0.f;
        for (int c = 0; c < a.nsplit; ++c) {
            const float* pc = p + c * (2 + a.D);
            const float w = pc[0] == -INFINITY ? 0.f : expf(pc[0] - M);
            den = fmaf(pc[1], w, den);
            num = fmaf(pc[2 + d], w, num);
        }
        a.xm_imp[i] = num / den;
    }
}

static size_t impute_smem(int D, int L) {
    const int NP2 = round4(2 * D);
    const size_t f = (size_t)L * IM_H + IM_H + (size_t)IM_H * IM_H + IM_H + (size_t)IM_H * NP2 + NP2 + (size_t)IM_MAXL * IM_P +
                     2 * (size_t)IM_H * IM_P + 4 * IM_MAXD + 3 * IM_MAXL + IM_TM + IM_MAXD + 4;
    return f * sizeof(float);
}

// samples per work item: enough items to fill the grid twice, chunks a multiple of the tile
static void impute_split(int N, int S, int grid, int* chunk, int* nsplit) {
    int want = (2 * grid + N - 1) / N;
    if (want < 1) want = 1;
    int ch = (S + want - 1) / want;
    ch = (ch + IM_TM - 1) / IM_TM * IM_TM;
    if (ch < 4 * IM_TM) ch = 4 * IM_TM;
    *chunk = ch;
    *nsplit = (S + ch - 1) / ch;
}

}  // namespace pcvae

using namespace pcvae;

extern "C" {

size_t pcvae_mnar_impute_workspace_bytes(int rows, int samples, int obs_dim) {
    int grid;
    if (device_ok(&grid) != PCVAE_OK || rows < 1 || samples < 1 || obs_dim < 1) return 0;
    int chunk, nsplit;
    impute_split(rows, samples, grid, &chunk, &nsplit);
    return (size_t)rows * nsplit * (2 + obs_dim) * sizeof(float);
}

int pcvae_mnar_impute(const pcvae_mnar_impute_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "mnar_impute: null params");
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 1 || p->samples < 1 || p->obs_dim < 1 || p->obs_dim > IM_MAXD || p->latent_dim < 1 || p->latent_dim > IM_MAXL)
        return fail(PCVAE_EINVAL, "mnar_impute: sizes (rows >= 1, samples >= 1, obs_dim 1..%d, latent_dim 1..%d)", IM_MAXD, IM_MAXL);
    if (!p->dec0_W || !p->dec0_b || !p->dec2_W || !p->dec2_b || !p->xmean_W || !p->xmean_b || !p->xlogvar_W || !p->xlogvar_b ||
        !p->W || !p->b || !p->x || !p->mask || !p->mean || !p->logvar || !p->xm_imputed || !p->workspace)
        return fail(PCVAE_EINVAL, "mnar_impute: null pointer");
    if (p->eps && !p->regularised && !p->eps_kl) return fail(PCVAE_EINVAL, "mnar_impute: eps_kl is needed beside eps when regularised = 0");
    ImputeArgs a{};
    a.N = p->rows; a.S = p->samples; a.D = p->obs_dim; a.L = p->latent_dim; a.reg = p->regularised ? 1 : 0;
    impute_split(a.N, a.S, grid, &a.chunk, &a.nsplit);
    if (p->workspace_bytes < (size_t)a.N * a.nsplit * (2 + a.D) * sizeof(float)) return fail(PCVAE_EINVAL, "mnar_impute: workspace too small");
    a.W1 = p->dec0_W; a.b1 = p->dec0_b; a.W2 = p->dec2_W; a.b2 = p->dec2_b; a.Wm = p->xmean_W; a.bm = p->xmean_b;
    a.Wv = p->xlogvar_W; a.bv = p->xlogvar_b; a.sW = p->W; a.sb = p->b; a.x = p->x; a.mask = p->mask; a.mean = p->mean;
    a.logvar = p->logvar; a.eps = p->eps; a.eps_kl = p->eps_kl; a.seed = p->seed; a.offset = p->offset;
    a.part = static_cast<float*>(p->workspace); a.xm_imp = p->xm_imputed;
    const size_t sm = impute_smem(a.D, a.L);
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "mnar_impute: shared memory %zu B exceeds %d", sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(k_mnar_impute, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "mnar_impute: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int items = a.N * a.nsplit;
    k_mnar_impute<<<items < grid ? items : grid, NT, sm, (cudaStream_t)stream>>>(a);
    const long total = (long)a.N * a.D;
    k_mnar_impute_merge<<<(int)((total + 255) / 256 < 4 * grid ? (total + 255) / 256 : 4 * grid), 256, 0, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "mnar_impute: launch: %s", cudaGetErrorString(e));
}

}  // extern "C"
