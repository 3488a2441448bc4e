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
// for 32 FFMA (RB=2) or a 4x4 block with 2 LDS.128 per 16 FFMA (RB=1, twice the warps).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcvae {

constexpr int NT = 512;          // threads per CTA (16 warps, <= 128 registers each)
constexpr int NWARP = NT / 32;

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_ELU = 3, ACT_HARDTANH = 4 };

__host__ __device__ constexpr int round4(int v) { return (v + 3) & ~3; }

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    if (act == ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
    if (act == ACT_ELU) return v > 0.0f ? v : expm1f(v);
    if (act == ACT_HARDTANH) return fminf(fmaxf(v, -10.0f), 0.0f);
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
    // Lanes run along n: the shared-memory stores of a warp are 32 consecutive floats (a scatter with lanes along k puts
    // the whole warp into one bank when NP % 32 == 0 -- 8 us per launch for the 128 x 128 layers of the MNAR networks).
    // Each thread reads four consecutive k of its row (one 16-byte load when aligned); the loads of four items are in
    // flight before the first store.
    const int lane = tid & 31, warp = tid >> 5;
    const int kq = (K + 3) >> 2, ng = (N + 31) >> 5, items = kq * ng;
    const bool vec = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    for (int base = warp; base < items; base += 4 * NWARP) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = base + u * NWARP;
            const int q = it / ng, n = (it - q * ng) * 32 + lane, k0 = 4 * q;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[u][j] = 0.f;
            if (it < items && n < N) {
                const float* w = W + (long)n * K + k0;
                if (vec) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(w));
                    v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (k0 + j < K) v[u][j] = __ldg(w + j);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = base + u * NWARP;
            const int q = it / ng, n = (it - q * ng) * 32 + lane, k0 = 4 * q;
            if (it < items && n < N) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k0 + j < K) W_s[(k0 + j) * NP + n] = v[u][j];
            }
        }
    }
    if (b_s)
        for (int n = tid; n < NP; n += NT) b_s[n] = (n < N) ? __ldg(b + n) : 0.0f;
}

__device__ __forceinline__ void zero_floats(float* p, int n, int tid) {
    for (int i = tid; i < n; i += NT) p[i] = 0.0f;
}

// ---------------------------------------------------------------------------------
// C[n][r] = act(b[n] + sum_k A[k][r] * W[k][n]),  n < NP (pad outputs get act(0)).
// lane -> (row group rg, output group); a thread owns RB chunks of 4 rows
// ({c*TM/RB + 4rg .. +3}, c < RB) and 4 consecutive outputs: per reduction step
// RB+1 LDS.128 feed 16*RB FFMA.
// ---------------------------------------------------------------------------------
template <int TM, int RB, int ACT, int RN = 4>
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ A_s, const float* __restrict__ W_s,
                                         const float* __restrict__ b_s, float* __restrict__ C_s,
                                         int K, int NP, int tid) {
    constexpr int P = TM + 4;
    constexpr int RG = TM / (4 * RB);   // row groups
    constexpr int NGW = 32 / RG;        // output groups per warp
    constexpr int CS = TM / RB;         // row offset between a thread's chunks
    static_assert(RG <= 32 && 32 % RG == 0, "row groups must tile a warp");
    static_assert(RN == 4 || RN == 2, "RN: outputs per thread");
    const int lane = tid & 31, warp = tid >> 5;
    const int rg = lane % RG, ngl = lane / RG;
    const int r0 = 4 * rg;
    for (int ng = warp * NGW + ngl; ng < NP / RN; ng += NWARP * NGW) {
        const int n0 = RN * ng;
        float acc[4 * RB][RN];
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const float bv = b_s[n0 + j];
#pragma unroll
            for (int i = 0; i < 4 * RB; ++i) acc[i][j] = bv;
        }
        const float* ap = A_s + r0;
        const float* wp = W_s + n0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float wv[RN];
            if (RN == 4) { const float4 w = lds4(wp + k * NP); wv[0] = w.x; wv[1] = w.y; wv[RN - 2] = w.z; wv[RN - 1] = w.w; }
            else { const float2 w = *reinterpret_cast<const float2*>(wp + k * NP); wv[0] = w.x; wv[1] = w.y; }
#pragma unroll
            for (int c = 0; c < RB; ++c) {
                const float4 a = lds4(ap + k * P + c * CS);
                const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < RN; ++j) acc[4 * c + i][j] = fmaf(av[i], wv[j], acc[4 * c + i][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < RN; ++j)
#pragma unroll
            for (int c = 0; c < RB; ++c) {
                float4 o;
                o.x = act_apply(acc[4 * c + 0][j], ACT); o.y = act_apply(acc[4 * c + 1][j], ACT);
                o.z = act_apply(acc[4 * c + 2][j], ACT); o.w = act_apply(acc[4 * c + 3][j], ACT);
                sts4(C_s + (n0 + j) * P + r0 + c * CS, o);
            }
    }
}

// ---------------------------------------------------------------------------------
// Data gradient: out[k][r] = (sum_n dY[n][r] * W[k][n]) * (RELU_MASK ? out[k][r] > 0 : 1)
// for k < Kout.  `out` holds the forward activation of the layer input on entry (for
// the ReLU mask) and is overwritten in place.  n runs over NP (pad rows of dY must be
// finite; pad weights are zero).
// ---------------------------------------------------------------------------------
template <int TM, int RB, bool RELU_MASK, int KPT = 4>
__device__ __forceinline__ void gemm_dx(const float* __restrict__ dY_s, const float* __restrict__ W_s,
                                        float* __restrict__ out_s, int Kout, int NP, int tid) {
    constexpr int P = TM + 4;
    constexpr int RG = TM / (4 * RB);
    constexpr int NGW = 32 / RG;
    constexpr int CS = TM / RB;
    constexpr int KB = KPT * NGW;    // outputs (k) covered by one warp per iteration
    const int lane = tid & 31, warp = tid >> 5;
    const int rg = lane % RG, kgl = lane / RG;
    const int r0 = 4 * rg;
    for (int kb = warp * KB; kb < Kout; kb += NWARP * KB) {
        int kk[KPT];
        const float* wrow[KPT];
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            kk[j] = kb + kgl + NGW * j;
            wrow[j] = W_s + min(kk[j], Kout - 1) * NP;
        }
        float acc[4 * RB][KPT];
#pragma unroll
        for (int i = 0; i < 4 * RB; ++i)
#pragma unroll
            for (int j = 0; j < KPT; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
        for (int n = 0; n < NP; n += 4) {
            float wv[KPT][4];
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                const float4 w = lds4(wrow[j] + n);
                wv[j][0] = w.x; wv[j][1] = w.y; wv[j][2] = w.z; wv[j][3] = w.w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int c = 0; c < RB; ++c) {
                    const float4 a = lds4(dY_s + (n + q) * P + r0 + c * CS);
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < KPT; ++j) acc[4 * c + i][j] = fmaf(av[i], wv[j][q], acc[4 * c + i][j]);
                }
        }
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            if (kk[j] < Kout) {
                float* o = out_s + kk[j] * P + r0;
#pragma unroll
                for (int c = 0; c < RB; ++c) {
                    float4 v = make_float4(acc[4 * c][j], acc[4 * c + 1][j], acc[4 * c + 2][j], acc[4 * c + 3][j]);
                    if (RELU_MASK) {
                        const float4 h = lds4(o + c * CS);
                        v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f;
                        v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
                    }
                    sts4(o + c * CS, v);
                }
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
template <int TM, int RS = 1>
__device__ __forceinline__ void gemm_dw(const float* __restrict__ X_s, const float* __restrict__ dY_s,
                                        float* __restrict__ dW_s, int K, int N, int NP, int tid) {
    // RS == 2: the tile's rows are split in two halves handled by different half-warps (for layers
    // with too few 20x20 blocks to occupy the CTA); the halves add into dW_s one after the other,
    // separated by a __syncthreads(), so the result stays deterministic.  Call from all threads.
    constexpr int P = TM + 4;
    constexpr int RROWS = TM / RS;
    const int hw = tid >> 4, hl = tid & 15;
    const int kl = hl & 3, nl = hl >> 2;
    const int nbk = (K + 19) / 20, nbn = (N + 19) / 20;
    const int nblk = nbk * nbn;
    static_assert(RS == 1 || RS == 2, "row split");
    for (int it0 = 0; it0 < nblk * RS; it0 += NT / 16) {
        const int item = it0 + hw;
        const bool active = item < nblk * RS;
        const int blk = active ? item % nblk : 0, half = active ? item / nblk : 0;
        const int kb = (blk % nbk) * 20, nb = (blk / nbk) * 20;
        float acc[5][5];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) acc[i][j] = 0.0f;
        if (active) {
            const float* xr[5];
            const float* yr[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                xr[i] = X_s + min(kb + kl + 4 * i, K - 1) * P + half * RROWS;
                yr[i] = dY_s + min(nb + nl + 4 * i, N - 1) * P + half * RROWS;
            }
#pragma unroll 2
            for (int r = 0; r < RROWS; r += 4) {
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
        }
#pragma unroll
        for (int h = 0; h < RS; ++h) {
            if (active && half == h) {
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
            if (RS > 1 && h + 1 < RS) __syncthreads();
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
    // lanes along n (conflict-free shared-memory reads, see stage_linear), four consecutive k per thread and store
    const int lane = tid & 31, warp = tid >> 5;
    const int kq = (K + 3) >> 2, ng = (N + 31) >> 5, items = kq * ng;
    const bool vec = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(gW) & 15) == 0);
    for (int it = warp; it < items; it += NWARP) {
        const int q = it / ng, n = (it - q * ng) * 32 + lane, k0 = 4 * q;
        if (n < N) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (k0 + j < K) ? dW_s[(k0 + j) * NP + n] : 0.f;
            float* g = gW + (long)n * K + k0;
            if (vec) {
                *reinterpret_cast<float4*>(g) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k0 + j < K) g[j] = v[j];
            }
        }
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
// The feature loop is split into SEG segments so that all warps have work; segment partials
// go to `scratch` ([SEG][K4][P] floats) and are summed in a fixed order (deterministic).
// Contains __syncthreads(): call from all threads.
template <int TM>
__device__ __forceinline__ void pnp_embed(const float* __restrict__ xs, const float* __restrict__ ms,
                                          const float* __restrict__ A_s, const float* __restrict__ C_s,
                                          float* __restrict__ agg_s, float* __restrict__ scratch, int scratch_floats,
                                          int D, int K4, int tid) {
    constexpr int P = TM + 4;
    constexpr int RG = TM / 4;
    constexpr int NGW = 32 / RG;
    const int lane = tid & 31, warp = tid >> 5;
    const int rg = lane % RG, ngl = lane / RG;
    const int r0 = 4 * rg;
    const int ngs = K4 / 4;                               // output groups
    int seg = (NWARP * NGW) / ngs;                        // feature segments that fit the CTA
    seg = max(1, min(seg, min(D, scratch_floats / (K4 * P))));
    const int dper = (D + seg - 1) / seg;
    for (int item = warp * NGW + ngl; item < ngs * seg; item += NWARP * NGW) {
        const int ng = item % ngs, sg = item / ngs;
        const int n0 = 4 * ng, d0 = sg * dper, d1 = min(D, d0 + dper);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int d = d0; d < d1; ++d) {
            const float4 x0 = lds4(xs + d * P + r0), m0 = lds4(ms + d * P + r0);
            const float4 a = lds4(A_s + d * K4 + n0), c = lds4(C_s + d * K4 + n0);
            const float xv[4] = {x0.x, x0.y, x0.z, x0.w}, mv[4] = {m0.x, m0.y, m0.z, m0.w};
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] = fmaf(mv[i], fmaxf(fmaf(xv[i], av[j], cv[j]), 0.f), acc[i][j]);
        }
        float* dst = scratch + sg * K4 * P;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            sts4(dst + (n0 + j) * P + r0, make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]));
    }
    __syncthreads();
    for (int i = tid; i < K4 * (TM / 4); i += NT) {
        const int j = i / (TM / 4), c4 = 4 * (i - j * (TM / 4));
        float4 s = lds4(scratch + j * P + c4);
        for (int sg = 1; sg < seg; ++sg) {
            const float4 v = lds4(scratch + sg * K4 * P + j * P + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        sts4(agg_s + j * P + c4, s);
    }
}

// dC[d][j] += sum_r t,  dA[d][j] += sum_r t * x   with  t = m * [x A + C > 0] * dagg[j][r].
// One thread per (feature d, group of 4 embedding columns): consecutive lanes take consecutive
// features (conflict-free with the +4 pitch), the dagg loads are warp broadcasts; each (d,j) has
// exactly one owner, so the shared-memory accumulate is race-free and deterministic.
template <int TM>
__device__ __forceinline__ void pnp_embed_bwd(const float* __restrict__ xs, const float* __restrict__ ms,
                                              const float* __restrict__ dagg_s, const float* __restrict__ A_s,
                                              const float* __restrict__ C_s, float* __restrict__ dA_s,
                                              float* __restrict__ dC_s, int D, int K, int K4, int tid) {
    constexpr int P = TM + 4;
    const int ngs = K4 / 4;
    for (int item = tid; item < D * ngs; item += NT) {
        const int jg = item / D, d = item - jg * D, j0 = 4 * jg;
        const float4 a4 = lds4(A_s + d * K4 + j0), c4 = lds4(C_s + d * K4 + j0);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
        float accA[4] = {0.f, 0.f, 0.f, 0.f}, accC[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int r = 0; r < TM; r += 4) {
            const float4 x4 = lds4(xs + d * P + r), m4 = lds4(ms + d * P + r);
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 g4 = lds4(dagg_s + (j0 + jj) * P + r);
                const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float tval = mv[i] * gv[i];
                    if (fmaf(xv[i], av[jj], cv[jj]) > 0.f) {       // predicated adds: one select fewer per (row, feature, column)
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
