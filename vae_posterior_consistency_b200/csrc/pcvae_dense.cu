// Generic dense layer on row tiles (weights resident in shared memory) -- forward and backward.
// Building block of the 128-wide not-MIWAE MNAR networks (src/models/VAE.py:2342-2363, 2706-2730).
#include "pcvae_internal.cuh"

namespace pcvae {

// Padded output count = pitch of the weight image W_s[K][NP].  A pitch that is a multiple of 32 floats puts W_s[k][n] for
// consecutive k into one bank: gemm_dx / gemm_dw then serialise 32-way (128-wide layers: 40 us for ONE 64-row tile).
// With pitch % 32 == 4 the eight 16-byte reads of a quarter-warp hit eight different bank groups (as 100 and 52 do).
__host__ __device__ inline int dense_pitch(int N) {
    const int np = round4(N);
    return (np % 32 == 0) ? np + 4 : np;
}

struct DenseArgs {
    int R, K, N, act;
    const float* x;
    const float* mask;
    const float* W;
    const float* b;
    float* y;
    const float* yin;
    const float* dy;
    float* dx;
    float* dWp;
    float* dbp;
};

// TM rows per tile: 64, or 32 when the batch has so few rows that 64-row tiles would leave most SMs idle (a CTA's time is
// rows x K x N on one SM: the 256-row encoder layers of the MNAR networks take half as long on twice as many SMs)
template <int ACT, int TM>
__global__ void __launch_bounds__(NT, 1) k_dense_fwd(const DenseArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int P = TM + 4, RB = 1;
    const int tid = threadIdx.x, K = a.K, N = a.N, NP = dense_pitch(N);
    float* W_s = smem;                 // [K][NP]
    float* b_s = W_s + K * NP;         // [NP]
    float* in_s = b_s + NP;            // [K][P]
    float* out_s = in_s + K * P;       // [NP][P]
    stage_linear(W_s, b_s, a.W, a.b, K, N, NP, tid);
    __syncthreads();
    const int ntiles = (a.R + TM - 1) / TM;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = t * TM;
        tile_elems<TM, 16, XM>(K, row0, a.R, tid,
            [&](int k, int r, bool ok) {
                XM v{0.f, 1.f};
                if (ok) {
                    const long gi = (long)(row0 + r) * K + k;
                    v.x = a.x[gi];
                    if (a.mask) v.m = a.mask[gi];
                }
                return v;
            },
            [&](int k, int r, bool, XM v) { in_s[k * P + r] = v.x * v.m; });
        __syncthreads();
        gemm_fwd<TM, RB, ACT>(in_s, W_s, b_s, out_s, K, NP, tid);
        __syncthreads();
        tile_elems<TM, 1, int>(N, row0, a.R, tid, [](int, int, bool) { return 0; },
            [&](int n, int r, bool ok, int) { if (ok) a.y[(long)(row0 + r) * N + n] = out_s[n * P + r]; });
        __syncthreads();
    }
}

__device__ __forceinline__ float act_grad_from_output(float y, int act) {
    if (act == ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == ACT_SIGMOID) return y * (1.f - y);
    if (act == ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
    if (act == ACT_HARDTANH) return (y > -10.f && y < 0.f) ? 1.f : 0.f;
    return 1.f;
}

template <int TM>
__global__ void __launch_bounds__(NT, 1) k_dense_bwd(const DenseArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int P = TM + 4, RB = 1;
    const int tid = threadIdx.x, K = a.K, N = a.N, NP = dense_pitch(N);
    float* W_s = smem;                 // [K][NP]
    float* dW_s = W_s + K * NP;        // [K][NP]
    float* db_s = dW_s + K * NP;       // [NP]
    float* in_s = db_s + NP;           // [K][P]   x*mask, then dx
    float* dy_s = in_s + K * P;        // [NP][P]  dL/d(pre-activation)
    stage_linear(W_s, nullptr, a.W, nullptr, K, N, NP, tid);
    zero_floats(dW_s, K * NP + NP, tid);
    zero_floats(dy_s, NP * P, tid);
    __syncthreads();
    const int ntiles = (a.R + TM - 1) / TM;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = t * TM;
        tile_elems<TM, 16, XM>(K, row0, a.R, tid,
            [&](int k, int r, bool ok) {
                XM v{0.f, 1.f};
                if (ok) {
                    const long gi = (long)(row0 + r) * K + k;
                    v.x = a.x[gi];
                    if (a.mask) v.m = a.mask[gi];
                }
                return v;
            },
            [&](int k, int r, bool, XM v) { in_s[k * P + r] = v.x * v.m; });
        tile_elems<TM, 16, XM>(N, row0, a.R, tid,
            [&](int n, int r, bool ok) {
                XM v{0.f, 0.f};
                if (ok) {
                    const long gi = (long)(row0 + r) * N + n;
                    v.x = a.dy[gi];
                    v.m = a.yin[gi];
                }
                return v;
            },
            [&](int n, int r, bool ok, XM v) { dy_s[n * P + r] = ok ? v.x * act_grad_from_output(v.m, a.act) : 0.f; });
        __syncthreads();
        gemm_dw<TM>(in_s, dy_s, dW_s, K, N, NP, tid);
        bias_dw<TM>(dy_s, db_s, N, tid);
        if (a.dx) {
            __syncthreads();
            gemm_dx<TM, RB, false>(dy_s, W_s, in_s, K, NP, tid);
            __syncthreads();
            tile_elems<TM, 1, int>(K, row0, a.R, tid, [](int, int, bool) { return 0; },
                [&](int k, int r, bool ok, int) {
                    if (ok) {
                        const long gi = (long)(row0 + r) * K + k;
                        a.dx[gi] = in_s[k * P + r] * (a.mask ? a.mask[gi] : 1.f);
                    }
                });
        }
        __syncthreads();
    }
    flush_linear_grad(dW_s, db_s, a.dWp + (long)blockIdx.x * N * K, a.dbp + (long)blockIdx.x * N, K, N, NP, tid);
}

// dW[i] = sum_c dWp[c][i], db[j] = sum_c dbp[c][j] over the CTAs that ran, in CTA order (deterministic)
__global__ void k_dense_reduce(const float* __restrict__ dWp, const float* __restrict__ dbp, int live, int nW, int nb,
                               float* __restrict__ dW, float* __restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nW + nb) return;
    const float* src = i < nW ? dWp + i : dbp + (i - nW);
    const long stride = i < nW ? nW : nb;
    float s = 0.f;
    int c = 0;
    for (; c + 8 <= live; c += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = src[(long)(c + j) * stride];
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
    }
    for (; c < live; ++c) s += src[(long)c * stride];
    if (i < nW) dW[i] = s; else db[i - nW] = s;
}

}  // namespace pcvae

using namespace pcvae;

static size_t dense_fwd_smem(int K, int N, int TM) {
    const int NP = dense_pitch(N), P = TM + 4;
    return ((size_t)K * NP + NP + (size_t)K * P + (size_t)NP * P) * sizeof(float);
}
static size_t dense_bwd_smem(int K, int N, int TM) {
    const int NP = dense_pitch(N), P = TM + 4;
    return (2 * (size_t)K * NP + NP + (size_t)K * P + (size_t)NP * P) * sizeof(float);
}

// rows per tile: 32 when 64-row tiles would occupy at most half of the SMs
static int tile_rows(int grid, int rows) { return 2 * ((rows + TM_TRAIN - 1) / TM_TRAIN) <= grid ? 32 : TM_TRAIN; }
// CTAs that own at least one tile (one CTA still runs for rows == 0 so that the partials are zeroed)
static int live_ctas(int grid, int rows) {
    const int tm = tile_rows(grid, rows), ntiles = (rows + tm - 1) / tm;
    return ntiles < 1 ? 1 : (ntiles < grid ? ntiles : grid);
}

template <typename Kern>
static int launch_dense(Kern kern, size_t smem, int grid, cudaStream_t st, const char* name, const DenseArgs& a) {
    if (smem > MAX_SMEM) return fail(PCVAE_EINVAL, "%s: needs %zu B shared memory (> %d)", name, smem, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    kern<<<grid, NT, smem, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "%s: launch: %s", name, cudaGetErrorString(e));
    return PCVAE_OK;
}

extern "C" {

int pcvae_dense_fwd(const pcvae_dense_fwd_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "dense_fwd: null params");
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 0 || p->in_dim < 1 || p->in_dim > MAX_D || p->out_dim < 1 || p->out_dim > MAX_D)
        return fail(PCVAE_EINVAL, "dense_fwd: sizes outside [1,%d]", MAX_D);
    if (p->rows == 0) return PCVAE_OK;
    if (!p->x || !p->W || !p->b || !p->y) return fail(PCVAE_EINVAL, "dense_fwd: null pointer");
    DenseArgs a{};
    a.R = p->rows; a.K = p->in_dim; a.N = p->out_dim; a.act = p->act; a.x = p->x; a.mask = p->mask; a.W = p->W; a.b = p->b; a.y = p->y;
    cudaStream_t st = (cudaStream_t)stream;
    const int tm = tile_rows(grid, a.R), live = live_ctas(grid, a.R);
    const size_t sm = dense_fwd_smem(a.K, a.N, tm);
#define PCVAE_DENSE_FWD(ACT) (tm == 32 ? launch_dense(k_dense_fwd<ACT, 32>, sm, live, st, "dense_fwd", a) \
                                       : launch_dense(k_dense_fwd<ACT, TM_TRAIN>, sm, live, st, "dense_fwd", a))
    switch (p->act) {
        case PCVAE_ACT_NONE: return PCVAE_DENSE_FWD(ACT_NONE);
        case PCVAE_ACT_RELU: return PCVAE_DENSE_FWD(ACT_RELU);
        case PCVAE_ACT_SIGMOID: return PCVAE_DENSE_FWD(ACT_SIGMOID);
        case PCVAE_ACT_ELU: return PCVAE_DENSE_FWD(ACT_ELU);
        case PCVAE_ACT_HARDTANH_M10_0: return PCVAE_DENSE_FWD(ACT_HARDTANH);
    }
#undef PCVAE_DENSE_FWD
    return fail(PCVAE_EINVAL, "dense_fwd: unknown activation %d", p->act);
}

int pcvae_dense_bwd(const pcvae_dense_bwd_params* p, void* stream) {
    if (!p) return fail(PCVAE_EINVAL, "dense_bwd: null params");
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (p->rows < 0 || p->in_dim < 1 || p->in_dim > MAX_D || p->out_dim < 1 || p->out_dim > MAX_D)
        return fail(PCVAE_EINVAL, "dense_bwd: sizes outside [1,%d]", MAX_D);
    if (p->act < PCVAE_ACT_NONE || p->act > PCVAE_ACT_HARDTANH_M10_0) return fail(PCVAE_EINVAL, "dense_bwd: unknown activation %d", p->act);
    if (!p->W || !p->dW_partials || !p->db_partials) return fail(PCVAE_EINVAL, "dense_bwd: null pointer");
    if (p->rows > 0 && (!p->x || !p->y || !p->dy)) return fail(PCVAE_EINVAL, "dense_bwd: null x/y/dy");
    DenseArgs a{};
    a.R = p->rows; a.K = p->in_dim; a.N = p->out_dim; a.act = p->act; a.x = p->x; a.mask = p->mask; a.W = p->W;
    a.yin = p->y; a.dy = p->dy; a.dx = p->dx; a.dWp = p->dW_partials; a.dbp = p->db_partials;
    if ((p->dW == nullptr) != (p->db == nullptr)) return fail(PCVAE_EINVAL, "dense_bwd: give both dW and db or neither");
    const int tm = tile_rows(grid, a.R), live = live_ctas(grid, a.R);
    if (int rc = tm == 32 ? launch_dense(k_dense_bwd<32>, dense_bwd_smem(a.K, a.N, tm), live, (cudaStream_t)stream, "dense_bwd", a)
                          : launch_dense(k_dense_bwd<TM_TRAIN>, dense_bwd_smem(a.K, a.N, tm), live, (cudaStream_t)stream, "dense_bwd", a))
        return rc;
    if (p->dW) {
        const int nW = a.N * a.K, n = nW + a.N;
        k_dense_reduce<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a.dWp, a.dbp, live, nW, a.N, p->dW, p->db);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(PCVAE_ECUDA, "dense_bwd: reduce launch: %s", cudaGetErrorString(e));
    }
    return PCVAE_OK;
}

}  // extern "C"
