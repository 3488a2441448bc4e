// Active-selection information reward (src/experiment_main/evaluate.py:416-425, 514-634)
// for one acquisition step: every (test row, candidate feature, MC sample) triple in one
// pass, using the incremental-encoder identity of SURVEY.md A.5:
//
//   R[n,u] = 1/M sum_m [ kl(tail(h0 + v w_u) || tail(h0)) - kl(tail(h0 + v w_u + t w_T) || tail(h0 + t w_T)) ]
//
// with h0 = W1 (x*mask) + b1 per row, v = im[m,n,u], t = im[m,n,D-1] (PNP: h0 -> masked
// sum-pool agg0, v w_u -> relu(v A_u + C_u)).  Three kernels:
//   k_reward_prep : per row tile: h0 (or agg0), tail(h0), tail(h0 + t_m w_T) for every m,
//                   R := -1e4, candidate lists
//   k_scan/k_pairs: compact (row, candidate) pair list
//   k_reward_main : 64 pairs x {with/without target} = 128 tail evaluations per tile, the
//                   sample loop inside with the per-pair accumulator in a register, summed
//                   in the reference's order (acc += KL_I; acc -= KL_II, evaluate.py:537-538)
//                   so results do not depend on tiling or on how rows are sharded over GPUs.
#include <cuda_pipeline.h>

#include "pcvae_internal.cuh"
#include "pcvae_reward.cuh"

namespace pcvae {

__device__ __forceinline__ void write_base(const float* o_s, int P, float* __restrict__ dst, long stride,
                                           int row0, int N, int TM, int tid) {
    for (int i = tid; i < TM * LAT; i += NT) {
        const int r = i / LAT, l = i - r * LAT;
        if (row0 + r < N) {
            float* d = dst + (long)(row0 + r) * stride;
            const float mu = o_s[l * P + r], lv = o_s[(LAT + l) * P + r];
            d[l] = mu;
            d[LAT + l] = lv;
            d[2 * LAT + l] = 1.0f / expf(lv * 0.5f);
            d[3 * LAT + l] = 1.0f / expf(lv);
        }
    }
}

template <int FAM>
__global__ void __launch_bounds__(NT, 1) k_reward_prep(const RewardArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RB = 1;
    constexpr int TM = TM_TRAIN, P = TM + 4;
    const int tid = threadIdx.x;
    const int D = a.L.D, K = a.L.K, K4 = round4(K);
    const int IN1 = (FAM == PCVAE_FAMILY_MLP) ? D : K;
    const int INW = (FAM == PCVAE_FAMILY_MLP) ? H1 : K4;
    float* W1_s = smem;
    float* b1_s = W1_s + IN1 * H1;
    float* W2_s = b1_s + H1;
    float* b2_s = W2_s + H1 * H2P;
    float* W3_s = b2_s + H2P;
    float* b3_s = W3_s + H2 * LAT2;
    float* in_s = b3_s + LAT2;            // [D][P]
    float* h1_s = in_s + D * P;           // [100][P]
    float* h2_s = h1_s + H1 * P;          // [52][P]
    float* o_s = h2_s + H2P * P;          // [20][P]
    float* t_s = o_s + LAT2 * P;          // [TM]
    float* h0_s = t_s + TM;               // MLP [100][P]
    float* ms_s = h0_s;                   // PNP [D][P]  (aliases h0_s: families are exclusive)
    float* A_s = ms_s + D * P;            // PNP [D][K4] then C
    float* C_s = A_s + D * K4;
    float* agg_s = C_s + D * K4;          // PNP [K4][P]
    float* aggT_s = agg_s + K4 * P;       // PNP [K4][P]

    stage_linear(W1_s, b1_s, a.theta + a.L.W1, a.theta + a.L.b1, IN1, H1, H1, tid);
    stage_linear(W2_s, b2_s, a.theta + a.L.W2, a.theta + a.L.b2, H1, H2, H2P, tid);
    stage_linear(W3_s, b3_s, a.theta + a.L.W3, a.theta + a.L.b3, H2, LAT2, LAT2, tid);
    if (FAM == PCVAE_FAMILY_PNP)
        for (int i = tid; i < 2 * D * K4; i += NT) A_s[i] = a.ac[i];
    __syncthreads();

    auto tail = [&](float* dst, long stride, int row0) {
        gemm_fwd<TM, RB, ACT_RELU, 2>(h1_s, W2_s, b2_s, h2_s, H1, H2P, tid);
        __syncthreads();
        gemm_fwd<TM, RB, ACT_NONE, 2>(h2_s, W3_s, b3_s, o_s, H2, LAT2, tid);
        __syncthreads();
        write_base(o_s, P, dst, stride, row0, a.N, TM, tid);
    };

    const int ntiles = (a.N + TM - 1) / TM;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = t * TM;
        tile_elems<TM, 16, XM>(D, row0, a.N, tid,
            [&](int d, int r, bool ok) {
                XM v{0.f, 0.f};
                if (ok) {
                    const long gi = (long)(row0 + r) * D + d;
                    v.x = a.x[gi];
                    v.m = load_mask(a.mask, gi, a.mask_kind);
                }
                return v;
            },
            [&](int d, int r, bool ok, XM v) {
                if (ok && d < D - 1) a.R[(long)(row0 + r) * (D - 1) + d] = -1e4f;     // evaluate.py:391
                if (FAM == PCVAE_FAMILY_MLP) in_s[d * P + r] = v.x * v.m;
                else { in_s[d * P + r] = v.x; ms_s[d * P + r] = v.m; }
            });
        // candidate list of each row: features u < D-1 with mask == 0   (evaluate.py:418)
        if (tid < TM && row0 + tid < a.N) {
            const long n = row0 + tid;
            int c = 0;
            for (int u = 0; u < D - 1; ++u)
                if (load_mask(a.mask, n * D + u, a.mask_kind) == 0.f) a.cand[n * CANDP + c++] = (uint8_t)u;
            a.cnt[n] = c;
        }
        __syncthreads();
        if (FAM == PCVAE_FAMILY_MLP) {
            gemm_fwd<TM, RB, ACT_NONE>(in_s, W1_s, b1_s, h0_s, D, H1, tid);        // h0 = W1 (x*mask) + b1
            __syncthreads();
            for (int i = tid; i < H1 * TM; i += NT) {
                const int k = i / TM, r = i - k * TM;
                const float v = h0_s[k * P + r];
                h1_s[k * P + r] = fmaxf(v, 0.f);
            }
            tile_elems<TM, 1, int>(H1, row0, a.N, tid, [](int, int, bool) { return 0; },
                [&](int k, int r, bool ok, int) {
                    if (ok) a.base_in[(long)(row0 + r) * INW + k] = h0_s[k * P + r];
                });
        } else {
            pnp_embed<TM>(in_s, ms_s, A_s, C_s, agg_s, h1_s, (H1 + H2P) * P, D, K4, tid);      // agg0
            __syncthreads();
            tile_elems<TM, 1, int>(K4, row0, a.N, tid, [](int, int, bool) { return 0; },
                [&](int j, int r, bool ok, int) {
                    if (ok) a.base_in[(long)(row0 + r) * INW + j] = agg_s[j * P + r];
                });
            gemm_fwd<TM, RB, ACT_RELU>(agg_s, W1_s, b1_s, h1_s, K, H1, tid);
        }
        __syncthreads();
        tail(a.base0, BASEW, row0);
        for (int m = 0; m < a.M; ++m) {
            if (tid < TM) t_s[tid] = (row0 + tid < a.N) ? a.im[(long)m * a.im_ss + (long)(row0 + tid) * D + (D - 1)] : 0.f;
            __syncthreads();
            if (FAM == PCVAE_FAMILY_MLP) {
                const float* wT = a.theta + a.L.W1 + (D - 1);           // W1[k][D-1], stride D
                for (int i = tid; i < H1 * TM; i += NT) {
                    const int k = i / TM, r = i - k * TM;
                    h1_s[k * P + r] = fmaxf(fmaf(t_s[r], __ldg(wT + (long)k * D), h0_s[k * P + r]), 0.f);
                }
            } else {
                for (int i = tid; i < K4 * TM; i += NT) {
                    const int j = i / TM, r = i - j * TM;
                    const float e = fmaxf(fmaf(t_s[r], A_s[(D - 1) * K4 + j], C_s[(D - 1) * K4 + j]), 0.f);
                    aggT_s[j * P + r] = agg_s[j * P + r] + e;
                }
                __syncthreads();
                gemm_fwd<TM, RB, ACT_RELU>(aggT_s, W1_s, b1_s, h1_s, K, H1, tid);
            }
            __syncthreads();
            tail(a.baseT + (long)m * BASEW, (long)a.M * BASEW, row0);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_scan(const int* __restrict__ cnt, int* __restrict__ off, int N) {
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int chunk = (N + 1023) / 1024;
    const int b = min(tid * chunk, N), e = min(b + chunk, N);
    int s = 0;
    for (int j = b; j < e; ++j) s += cnt[j];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (tid >= o) ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - s;
    for (int j = b; j < e; ++j) { off[j] = run; run += cnt[j]; }
    if (tid == 1023) off[N] = part[1023];
}

__global__ void k_pairs(const int* __restrict__ cnt, const int* __restrict__ off, const uint8_t* __restrict__ cand,
                        int* __restrict__ pairs, int N) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const int c = cnt[n], o = off[n];
        for (int j = 0; j < c; ++j) pairs[o + j] = n * CANDP + cand[(long)n * CANDP + j];
    }
}

template <int FAM>
__global__ void __launch_bounds__(NT, 1) k_reward_main(const RewardArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RB = 1;
    constexpr int TM = TM_REWARD, P = TM + 4, NP_ = NPAIR;
    const int tid = threadIdx.x;
    const int D = a.L.D, K = a.L.K, K4 = round4(K);
    const int INW = (FAM == PCVAE_FAMILY_MLP) ? H1 : K4;
    float* W2_s = smem;
    float* b2_s = W2_s + H1 * H2P;
    float* W3_s = b2_s + H2P;
    float* b3_s = W3_s + H2 * LAT2;
    float* colT_s = b3_s + LAT2;                 // [INW] target column: W1[:,D-1] (MLP) / A[D-1] (PNP)
    float* colTC_s = colT_s + INW;               // [INW] PNP: C[D-1]
    float* H0_s = colTC_s + INW;                 // [INW][64]
    float* U_s = H0_s + INW * NP_;               // [INW][64]  W1[:,u] (MLP) / A[u] (PNP)
    float* h_s = U_s + INW * NP_;                // [100][P]
    float* h2_s = h_s + H1 * P;                  // [52][P]
    float* o_s = h2_s + H2P * P;                 // [20][P]
    float* b0_s = o_s + LAT2 * P;                // [40][64]
    float* bT_s = b0_s + BASEW * NP_;            // [2][64][40]
    float* v_s = bT_s + 2 * NP_ * BASEW;         // [2][64]
    float* t_s = v_s + 2 * NP_;                  // [2][64]
    float* term_s = t_s + 2 * NP_;               // [20][64]
    int* pn_s = reinterpret_cast<int*>(term_s + 2 * LAT * NP_);   // [64]
    int* pu_s = pn_s + NP_;                      // [64]
    float* hin_s = reinterpret_cast<float*>(pu_s + NP_);          // PNP [K4][P]
    float* W1_s = hin_s + K4 * P;                // PNP [K][100]
    float* b1_s = W1_s + K * H1;                 // PNP [100]
    float* UC_s = b1_s + H1;                     // PNP [K4][64]  C[u]

    stage_linear(W2_s, b2_s, a.theta + a.L.W2, a.theta + a.L.b2, H1, H2, H2P, tid);
    stage_linear(W3_s, b3_s, a.theta + a.L.W3, a.theta + a.L.b3, H2, LAT2, LAT2, tid);
    if (FAM == PCVAE_FAMILY_MLP) {
        for (int k = tid; k < H1; k += NT) colT_s[k] = a.theta[a.L.W1 + (long)k * D + (D - 1)];
    } else {
        stage_linear(W1_s, b1_s, a.theta + a.L.W1, a.theta + a.L.b1, K, H1, H1, tid);
        for (int j = tid; j < K4; j += NT) {
            colT_s[j] = a.ac[(D - 1) * K4 + j];
            colTC_s[j] = a.ac[D * K4 + (D - 1) * K4 + j];
        }
    }
    __syncthreads();

    const int ptot = a.off[a.N];
    const int ntiles = (ptot + NP_ - 1) / NP_;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * NP_;
        if (tid < NP_) {
            int n = 0, u = 0;
            if (p0 + tid < ptot) { const int pr = a.pairs[p0 + tid]; n = pr / CANDP; u = pr - n * CANDP; }
            pn_s[tid] = n;
            pu_s[tid] = u;
        }
        __syncthreads();
        auto prefetch = [&](int m, int buf) {
            for (int c = tid; c < NP_ * (BASEW / 4); c += NT) {
                const int i = c / (BASEW / 4), q = c - i * (BASEW / 4);
                __pipeline_memcpy_async(bT_s + (buf * NP_ + i) * BASEW + 4 * q,
                                        a.baseT + ((long)pn_s[i] * a.M + m) * BASEW + 4 * q, 16);
            }
            if (tid < NP_) {
                const float* row = a.im + (long)m * a.im_ss + (long)pn_s[tid] * D;
                __pipeline_memcpy_async(v_s + buf * NP_ + tid, row + pu_s[tid], 4);
                __pipeline_memcpy_async(t_s + buf * NP_ + tid, row + (D - 1), 4);
            }
            __pipeline_commit();
        };
        prefetch(0, 0);
        for (int idx = tid; idx < INW * NP_; idx += NT) {
            const int k = idx / NP_, i = idx - k * NP_;
            H0_s[idx] = a.base_in[(long)pn_s[i] * INW + k];
            if (FAM == PCVAE_FAMILY_MLP) {
                U_s[idx] = __ldg(a.theta + a.L.W1 + (long)k * D + pu_s[i]);
            } else {
                U_s[idx] = a.ac[pu_s[i] * K4 + k];
                UC_s[idx] = a.ac[D * K4 + pu_s[i] * K4 + k];
            }
        }
        for (int idx = tid; idx < BASEW * NP_; idx += NT) {
            const int f = idx / NP_, i = idx - f * NP_;
            b0_s[idx] = a.base0[(long)pn_s[i] * BASEW + f];
        }
        float acc = 0.f;
        __pipeline_wait_prior(0);
        __syncthreads();

        for (int m = 0; m < a.M; ++m) {
            const int buf = m & 1;
            if (m + 1 < a.M) prefetch(m + 1, buf ^ 1);
            const float* vb = v_s + buf * NP_;
            const float* tb = t_s + buf * NP_;
            if (FAM == PCVAE_FAMILY_MLP) {
                for (int idx = tid; idx < H1 * NP_; idx += NT) {
                    const int k = idx / NP_, i = idx - k * NP_;
                    const float hA = fmaf(vb[i], U_s[idx], H0_s[idx]);
                    const float hB = fmaf(tb[i], colT_s[k], hA);
                    h_s[k * P + i] = fmaxf(hA, 0.f);
                    h_s[k * P + NP_ + i] = fmaxf(hB, 0.f);
                }
            } else {
                for (int idx = tid; idx < K4 * NP_; idx += NT) {
                    const int j = idx / NP_, i = idx - j * NP_;
                    const float eu = fmaxf(fmaf(vb[i], U_s[idx], UC_s[idx]), 0.f);
                    const float eT = fmaxf(fmaf(tb[i], colT_s[j], colTC_s[j]), 0.f);
                    hin_s[j * P + i] = H0_s[idx] + eu;
                    hin_s[j * P + NP_ + i] = (H0_s[idx] + eT) + eu;
                }
                __syncthreads();
                gemm_fwd<TM, RB, ACT_RELU>(hin_s, W1_s, b1_s, h_s, K, H1, tid);
            }
            __syncthreads();
            gemm_fwd<TM, RB, ACT_RELU>(h_s, W2_s, b2_s, h2_s, H1, H2P, tid);
            __syncthreads();
            gemm_fwd<TM, RB, ACT_NONE, 2>(h2_s, W3_s, b3_s, o_s, H2, LAT2, tid);
            __syncthreads();
            // KL terms, evaluate.py:582-583 / 631-632 (divide by std, not variance)
            for (int idx = tid; idx < 2 * LAT * NP_; idx += NT) {
                const int i = idx & (NP_ - 1), rest = idx / NP_;
                const int w = rest / LAT, l = rest - w * LAT;
                const float mu_i = o_s[l * P + w * NP_ + i], lv_i = o_s[(LAT + l) * P + w * NP_ + i];
                float mu, lv, rstd, rvar;
                if (w == 0) {
                    mu = b0_s[l * NP_ + i]; lv = b0_s[(LAT + l) * NP_ + i];
                    rstd = b0_s[(2 * LAT + l) * NP_ + i]; rvar = b0_s[(3 * LAT + l) * NP_ + i];
                } else {
                    const float* bt = bT_s + (buf * NP_ + i) * BASEW;
                    mu = bt[l]; lv = bt[LAT + l]; rstd = bt[2 * LAT + l]; rvar = bt[3 * LAT + l];
                }
                const float d = mu_i - mu;
                term_s[rest * NP_ + i] = (((d * d) * rstd + expf(lv_i) * rvar - 1.0f) - lv_i) + lv;
            }
            __syncthreads();
            if (tid < 2 * NP_) {
                const int i = tid >> 1, w = tid & 1;
                float s = 0.f;
#pragma unroll
                for (int l = 0; l < LAT; ++l) s += term_s[(w * LAT + l) * NP_ + i];
                const float kl = 0.5f * s;
                const float other = __shfl_xor_sync(0xffffffffu, kl, 1);
                if (w == 0) { acc += kl; acc -= other; }      // approx_KL += KL_I; approx_KL -= KL_II
            }
            __pipeline_wait_prior(0);
            __syncthreads();
        }
        if (tid < 2 * NP_ && (tid & 1) == 0) {
            const int i = tid >> 1;
            if (p0 + i < ptot) a.R[(long)pn_s[i] * (D - 1) + pu_s[i]] = acc / (float)a.M;   // evaluate.py:540
        }
        __syncthreads();
    }
}

static size_t prep_smem(const Layout& L) {
    const int P = TM_TRAIN + 4, K4 = round4(L.K);
    const int in1 = L.fam == PCVAE_FAMILY_MLP ? L.D : L.K;
    size_t f = (size_t)in1 * H1 + H1 + H1 * H2P + H2P + H2 * LAT2 + LAT2 + (size_t)(L.D + H1 + H2P + LAT2) * P + TM_TRAIN;
    if (L.fam == PCVAE_FAMILY_MLP) f += (size_t)H1 * P;
    else f += (size_t)L.D * P + 2 * L.D * K4 + 2 * K4 * P;
    return f * sizeof(float);
}

static size_t main_smem(const Layout& L) {
    const int P = TM_REWARD + 4, K4 = round4(L.K);
    const int inw = L.fam == PCVAE_FAMILY_MLP ? H1 : K4;
    size_t f = (size_t)H1 * H2P + H2P + H2 * LAT2 + LAT2 + 2 * inw + 2 * (size_t)inw * NPAIR + (size_t)(H1 + H2P + LAT2) * P +
               BASEW * NPAIR + 2 * NPAIR * BASEW + 4 * NPAIR + 2 * LAT * NPAIR + 2 * NPAIR;
    if (L.fam == PCVAE_FAMILY_PNP) f += (size_t)K4 * P + L.K * H1 + H1 + (size_t)K4 * NPAIR;
    return f * sizeof(float);
}

struct WsPlan { size_t base_in, base0, baseT, cnt, off, cand, pairs, total; };

static WsPlan plan_ws(const Layout& L, int N, int M) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const int inw = L.fam == PCVAE_FAMILY_MLP ? H1 : round4(L.K);
    WsPlan w;
    size_t o = 0;
    w.base_in = o; o = al(o + (size_t)N * inw * 4);
    w.base0 = o; o = al(o + (size_t)N * BASEW * 4);
    w.baseT = o; o = al(o + (size_t)N * M * BASEW * 4);
    w.cnt = o; o = al(o + (size_t)N * 4);
    w.off = o; o = al(o + (size_t)(N + 1) * 4);
    w.cand = o; o = al(o + (size_t)N * CANDP);
    w.pairs = o; o = al(o + (size_t)N * (L.D > 1 ? L.D - 1 : 1) * 4);
    w.total = o;
    return w;
}

}  // namespace pcvae

using namespace pcvae;

static int g_reward_tc = 1;   // MLP family: 1 = warp-specialised tcgen05 kernel (default), 2 = lock-step tcgen05 kernel, 0 = FFMA

extern "C" {

int pcvae_set_reward_tensor_cores(int enable) {
    const int prev = g_reward_tc;
    g_reward_tc = enable == 2 ? 2 : (enable ? 1 : 0);
    return prev;
}

size_t pcvae_reward_workspace_bytes(const pcvae_model* m, int rows, int samples) {
    Layout L;
    if (!make_layout(m, &L) || rows < 0 || samples < 1) return 0;
    return plan_ws(L, rows, samples).total;
}

int pcvae_reward_chain(const pcvae_reward_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "reward_chain: null params");
    Layout L;
    if (!make_layout(&p->model, &L)) return PCVAE_EINVAL;
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (L.D < 2) return fail(PCVAE_EINVAL, "reward_chain: obs_dim must be >= 2 (last column is the target)");
    if (p->rows < 0 || p->samples < 1) return fail(PCVAE_EINVAL, "reward_chain: bad rows/samples");
    if (p->rows == 0) return PCVAE_OK;
    if ((long)p->rows * CANDP > 2147483647L) return fail(PCVAE_EINVAL, "reward_chain: too many rows for one call (shard by rows)");
    if (!p->theta || !p->x || !p->mask || !p->im || !p->R || !p->workspace) return fail(PCVAE_EINVAL, "reward_chain: null pointer");
    const WsPlan w = plan_ws(L, p->rows, p->samples);
    if (p->workspace_bytes < w.total) return fail(PCVAE_EWORKSPACE, "reward_chain: workspace %zu < %zu bytes", p->workspace_bytes, w.total);
    if (L.aug) return fail(PCVAE_EINVAL, "reward_chain: the mask-augmented family has no reward kernel (active_learning.py never selects it)");
    if (L.fam == PCVAE_FAMILY_PNP && !p->pnp_ac) return fail(PCVAE_EINVAL, "reward_chain: PNP family needs pnp_ac workspace");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)p->workspace;
    RewardArgs a{};
    a.L = L; a.N = p->rows; a.M = p->samples; a.mask_kind = p->mask_kind;
    a.theta = p->theta; a.x = p->x; a.mask = p->mask; a.im = p->im; a.im_ss = p->im_sample_stride; a.R = p->R;
    a.base_in = (float*)(ws + w.base_in); a.base0 = (float*)(ws + w.base0); a.baseT = (float*)(ws + w.baseT);
    a.cnt = (int*)(ws + w.cnt); a.off = (int*)(ws + w.off); a.cand = (uint8_t*)(ws + w.cand); a.pairs = (int*)(ws + w.pairs);
    a.ac = p->pnp_ac;
    a.status = tc_status_ptr();
    cudaError_t e;
    const size_t s1 = prep_smem(L), s2 = main_smem(L);
    if (s1 > MAX_SMEM || s2 > MAX_SMEM) return fail(PCVAE_EINVAL, "reward_chain: shared memory %zu/%zu B exceeds %d", s1, s2, MAX_SMEM);
    if (L.fam == PCVAE_FAMILY_PNP) {
        pnp_tables_launch(L, p->theta, p->pnp_ac, st);
        cudaFuncSetAttribute(k_reward_prep<PCVAE_FAMILY_PNP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
        cudaFuncSetAttribute(k_reward_main<PCVAE_FAMILY_PNP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2);
        k_reward_prep<PCVAE_FAMILY_PNP><<<grid, NT, s1, st>>>(a);
    } else {
        cudaFuncSetAttribute(k_reward_prep<PCVAE_FAMILY_MLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
        cudaFuncSetAttribute(k_reward_main<PCVAE_FAMILY_MLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2);
        k_reward_prep<PCVAE_FAMILY_MLP><<<grid, NT, s1, st>>>(a);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(PCVAE_ECUDA, "reward_chain: prep launch: %s", cudaGetErrorString(e));
    k_scan<<<1, 1024, 0, st>>>(a.cnt, a.off, a.N);
    k_pairs<<<(a.N + 255) / 256, 256, 0, st>>>(a.cnt, a.off, a.cand, a.pairs, a.N);
    if (L.fam == PCVAE_FAMILY_PNP) k_reward_main<PCVAE_FAMILY_PNP><<<grid, NT, s2, st>>>(a);
    else if (g_reward_tc == 1) { if (int rc = reward_main_ws_launch(a, grid, st)) return rc; }
    else if (g_reward_tc == 2) { if (int rc = reward_main_tc_launch(a, grid, st)) return rc; }
    else k_reward_main<PCVAE_FAMILY_MLP><<<grid, NT, s2, st>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(PCVAE_ECUDA, "reward_chain: main launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // extern "C"
