// Row-tile building blocks shared by every kernel of the hot path (sm_100a).
//
// A CTA (NT threads, persistent, one per SM) owns a tile of TM table rows.  All
// activations of the tile live in shared memory FEATURE-MAJOR: buf[feature][row] with a
// row pitch of TM+4 floats (16-byte aligned rows; +4 skews consecutive features by four
// banks so that eight lanes reading eight consecutive features are conflict-free).
// Weights live in shared memory as W_s[in][NP] (NP = outputs padded to a multiple of 4,
// pad columns zero) for the whole kernel, so one layout serves the forward product
// (reduce over `in`, vector loads along `out`), the data-gradient product (reduce over
// `out` in chunks of 4) and the weight-gradient accumulation.
//
// The dense layers here are 10..128 wide in true FP32 (north_star: masks/indices
// bit-exact, fp32 losses to 1e-4), so they are register-blocked FFMA mini-GEMMs; each
// thread keeps an 8x4 (rows x outputs) block and per reduction step issues 3 LDS.128
// for 32 FFMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcvae {

constexpr int NT = 256;          // threads per CTA
constexpr int NWARP = NT / 32;

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

__host__ __device__ constexpr int round4(int v) { return (v + 3) & ~3; }

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    if (act == ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
    return v;
}

// ---------------------------------------------------------------------------------
// Stage an nn.Linear weight [N][K] (row-major, global) into W_s[K][NP] (+ bias[NP]).
// Global reads are coalesced; pad columns are zeroed.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void stage_linear(float* W_s, float* b_s, const float* __restrict__ W,
                                             const float* __restrict__ b, int K, int N, int NP, int tid) {
    if (NP != N)
        for (int i = tid; i < K * (NP - N); i += NT) {
            int k = i / (NP - N), n = N + i % (NP - N);
            W_s[k * NP + n] = 0.0f;
        }
    for (int i = tid; i < N * K; i += NT) {
        int n = i / K, k = i - n * K;
        W_s[k * NP + n] = __ldg(W + i);
    }
    if (b_s)
        for (int n = tid; n < NP; n += NT) b_s[n] = (n < N) ? __ldg(b + n) : 0.0f;
}

__device__ __forceinline__ void zero_floats(float* p, int n, int tid) {
    for (int i = tid; i < n; i += NT) p[i] = 0.0f;
}

// ---------------------------------------------------------------------------------
// C[n][r] = act(b[n] + sum_k A[k][r] * W[k][n]),  n < NP (pad outputs get act(0)).
// lane -> (row group rg, output group); a thread owns rows {4rg..4rg+3} and
// {TM/2+4rg..+3} and 4 consecutive outputs.
// ---------------------------------------------------------------------------------
template <int TM, int ACT>
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ A_s, const float* __restrict__ W_s,
                                         const float* __restrict__ b_s, float* __restrict__ C_s,
                                         int K, int NP, int tid) {
    constexpr int P = TM + 4;
    constexpr int RG = TM / 8;       // row groups: 8 (TM=64) / 16 (TM=128)
    constexpr int NGW = 32 / RG;     // output groups per warp
    const int lane = tid & 31, warp = tid >> 5;
    const int rg = lane % RG, ngl = lane / RG;
    const int r0 = 4 * rg, r1 = TM / 2 + 4 * rg;
    for (int ng = warp * NGW + ngl; ng < NP / 4; ng += NWARP * NGW) {
        const int n0 = 4 * ng;
        float acc[8][4];
        {
            float4 bv = lds4(b_s + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i][0] = bv.x; acc[i][1] = bv.y; acc[i][2] = bv.z; acc[i][3] = bv.w; }
        }
        const float* ap = A_s + r0;
        const float* wp = W_s + n0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 a0 = lds4(ap + k * P);
            const float4 a1 = lds4(ap + k * P + (r1 - r0));
            const float4 w = lds4(wp + k * NP);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 o0, o1;
            o0.x = act_apply(acc[0][j], ACT); o0.y = act_apply(acc[1][j], ACT);
            o0.z = act_apply(acc[2][j], ACT); o0.w = act_apply(acc[3][j], ACT);
            o1.x = act_apply(acc[4][j], ACT); o1.y = act_apply(acc[5][j], ACT);
            o1.z = act_apply(acc[6][j], ACT); o1.w = act_apply(acc[7][j], ACT);
            sts4(C_s + (n0 + j) * P + r0, o0);
            sts4(C_s + (n0 + j) * P + r1, o1);
        }
    }
}

// ---------------------------------------------------------------------------------
// Data gradient: out[k][r] = (sum_n dY[n][r] * W[k][n]) * (RELU_MASK ? out[k][r] > 0 : 1)
// for k < Kout.  `out` holds the forward activation of the layer input on entry (for
// the ReLU mask) and is overwritten in place.  n runs over NP (pad rows of dY must be
// finite; pad weights are zero).
// ---------------------------------------------------------------------------------
template <int TM, bool RELU_MASK>
__device__ __forceinline__ void gemm_dx(const float* __restrict__ dY_s, const float* __restrict__ W_s,
                                        float* __restrict__ out_s, int Kout, int NP, int tid) {
    constexpr int P = TM + 4;
    constexpr int RG = TM / 8;
    constexpr int NGW = 32 / RG;
    constexpr int KB = 4 * NGW;      // outputs (k) covered by one warp per iteration
    const int lane = tid & 31, warp = tid >> 5;
    const int rg = lane % RG, kgl = lane / RG;
    const int r0 = 4 * rg, r1 = TM / 2 + 4 * rg;
    for (int kb = warp * KB; kb < Kout; kb += NWARP * KB) {
        int kk[4];
        const float* wrow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            kk[j] = kb + kgl + NGW * j;
            wrow[j] = W_s + min(kk[j], Kout - 1) * NP;
        }
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 1
        for (int n = 0; n < NP; n += 4) {
            float wv[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = lds4(wrow[j] + n);
                wv[j][0] = w.x; wv[j][1] = w.y; wv[j][2] = w.z; wv[j][3] = w.w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 a0 = lds4(dY_s + (n + q) * P + r0);
                const float4 a1 = lds4(dY_s + (n + q) * P + r1);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j][q], acc[i][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (kk[j] < Kout) {
                float* o = out_s + kk[j] * P;
                float4 v0 = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                float4 v1 = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
                if (RELU_MASK) {
                    const float4 h0 = lds4(o + r0), h1 = lds4(o + r1);
                    v0.x = h0.x > 0.f ? v0.x : 0.f; v0.y = h0.y > 0.f ? v0.y : 0.f;
                    v0.z = h0.z > 0.f ? v0.z : 0.f; v0.w = h0.w > 0.f ? v0.w : 0.f;
                    v1.x = h1.x > 0.f ? v1.x : 0.f; v1.y = h1.y > 0.f ? v1.y : 0.f;
                    v1.z = h1.z > 0.f ? v1.z : 0.f; v1.w = h1.w > 0.f ? v1.w : 0.f;
                }
                sts4(o + r0, v0);
                sts4(o + r1, v1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// Weight gradient: dW_s[k][n] += sum_r X[k][r] * dY[n][r]  (k < K, n < N), reduction over
// the TM rows of the tile in chunks of 4.  A half-warp owns a 20x20 block (5x5 per
// thread, features interleaved by 4 so that loads are conflict-free / broadcast); each
// (k,n) is owned by exactly one thread, so the shared-memory accumulate is race-free.
// Rows past the end of the table must have dY == 0.
// ---------------------------------------------------------------------------------
template <int TM>
__device__ __forceinline__ void gemm_dw(const float* __restrict__ X_s, const float* __restrict__ dY_s,
                                        float* __restrict__ dW_s, int K, int N, int NP, int tid) {
    constexpr int P = TM + 4;
    const int hw = tid >> 4, hl = tid & 15;
    const int kl = hl & 3, nl = hl >> 2;
    const int nbk = (K + 19) / 20, nbn = (N + 19) / 20;
    for (int blk = hw; blk < nbk * nbn; blk += NT / 16) {
        const int kb = (blk % nbk) * 20, nb = (blk / nbk) * 20;
        const float* xr[5];
        const float* yr[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            xr[i] = X_s + min(kb + kl + 4 * i, K - 1) * P;
            yr[i] = dY_s + min(nb + nl + 4 * i, N - 1) * P;
        }
        float acc[5][5];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
        for (int r = 0; r < TM; r += 4) {
            float4 xv[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) xv[i] = lds4(xr[i] + r);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float4 yv = lds4(yr[j] + r);
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    acc[i][j] = fmaf(xv[i].x, yv.x, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].y, yv.y, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].z, yv.z, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].w, yv.w, acc[i][j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int k = kb + kl + 4 * i;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int n = nb + nl + 4 * j;
                if (k < K && n < N) dW_s[k * NP + n] += acc[i][j];
            }
        }
    }
}

// db_s[n] += sum_r dY[n][r]
template <int TM>
__device__ __forceinline__ void bias_dw(const float* __restrict__ dY_s, float* __restrict__ db_s, int N, int tid) {
    constexpr int P = TM + 4;
    for (int n = tid; n < N; n += NT) {
        float s = 0.0f;
#pragma unroll 4
        for (int r = 0; r < TM; r += 4) {
            const float4 v = lds4(dY_s + n * P + r);
            s += (v.x + v.y) + (v.z + v.w);
        }
        db_s[n] += s;
    }
}

// Write a shared-memory gradient accumulator dW_s[K][NP] (+ bias) to global memory in the
// nn.Linear [N][K] layout (coalesced stores).
__device__ __forceinline__ void flush_linear_grad(const float* dW_s, const float* db_s, float* __restrict__ gW,
                                                  float* __restrict__ gb, int K, int N, int NP, int tid) {
    for (int i = tid; i < N * K; i += NT) {
        int n = i / K, k = i - n * K;
        gW[i] = dW_s[k * NP + n];
    }
    if (gb)
        for (int n = tid; n < N; n += NT) gb[n] = db_s[n];
}

// ---------------------------------------------------------------------------------
// Tile element visitor: global row-major [rows][D] <-> shared feature-major [D][P].  A warp
// step covers 4 rows x 8 features: each row segment is one 32-byte sector of global memory
// and the 32 shared-memory accesses hit 32 distinct banks.  Loads are issued in batches of
// up to BATCH independent requests before the first use, so a tile costs a couple of memory
// round trips instead of one per element (the first profile showed `long_scoreboard` as the
// top stall with one-at-a-time loads).
//   load(d, r, ok) -> T      (ok: row0 + r < B)          use(d, r, ok, T)
// ---------------------------------------------------------------------------------
template <int TM, int BATCH, typename T, typename LoadF, typename UseF>
__device__ __forceinline__ void tile_elems(int D, int row0, int B, int tid, LoadF load, UseF use) {
    const int lane = tid & 31, warp = tid >> 5;
    const int dl = lane & 7, rl = lane >> 3;
    const int ndb = (D + 7) >> 3;
    for (int rb = warp; rb < TM / 4; rb += NWARP) {
        const int r = rb * 4 + rl;
        const bool ok = (row0 + r) < B;
        for (int db0 = 0; db0 < ndb; db0 += BATCH) {
            T v[BATCH];
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                const int d = (db0 + j) * 8 + dl;
                if (d < D) v[j] = load(d, r, ok);
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                const int d = (db0 + j) * 8 + dl;
                if (d < D) use(d, r, ok, v[j]);
            }
        }
    }
}

struct XM { float x, m; };

// prefetch [base, base+bytes) into L2, one request per 128-byte line, spread over the CTA
__device__ __forceinline__ void prefetch_l2(const void* base, long bytes, int tid) {
    const char* p = reinterpret_cast<const char*>(base);
    for (long off = (long)tid * 128; off < bytes; off += (long)NT * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

// agg[j][r] = sum_d m[d][r] * relu(x[d][r] * A[d][j] + C[d][j])     (VAE.py:726-733)
template <int TM>
__device__ __forceinline__ void pnp_embed(const float* __restrict__ xs, const float* __restrict__ ms,
                                          const float* __restrict__ A_s, const float* __restrict__ C_s,
                                          float* __restrict__ agg_s, int D, int K4, int tid) {
    constexpr int P = TM + 4;
    constexpr int RG = TM / 8;
    constexpr int NGW = 32 / RG;
    const int lane = tid & 31, warp = tid >> 5;
    const int rg = lane % RG, ngl = lane / RG;
    const int r0 = 4 * rg, r1 = TM / 2 + 4 * rg;
    for (int ng = warp * NGW + ngl; ng < K4 / 4; ng += NWARP * NGW) {
        const int n0 = 4 * ng;
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int d = 0; d < D; ++d) {
            const float4 x0 = lds4(xs + d * P + r0), x1 = lds4(xs + d * P + r1);
            const float4 m0 = lds4(ms + d * P + r0), m1 = lds4(ms + d * P + r1);
            const float4 a = lds4(A_s + d * K4 + n0), c = lds4(C_s + d * K4 + n0);
            const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] = fmaf(mv[i], fmaxf(fmaf(xv[i], av[j], cv[j]), 0.f), acc[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sts4(agg_s + (n0 + j) * P + r0, make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]));
            sts4(agg_s + (n0 + j) * P + r1, make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]));
        }
    }
}


__device__ __forceinline__ float load_mask(const void* __restrict__ m, long idx, int kind) {
    if (kind == 0) return reinterpret_cast<const uint8_t*>(m)[idx] ? 1.0f : 0.0f;
    return reinterpret_cast<const float*>(m)[idx];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pcvae
