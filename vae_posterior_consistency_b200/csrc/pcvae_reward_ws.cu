// Warp-specialised tcgen05 reward main kernel for the MLP family (the default; pcvae_reward_tc.cu keeps the
// lock-step version as the in-process cross-check, pcvae_reward.cu the FP32 FFMA one).
//
// Mathematics, tile shape (64 (row,candidate) pairs x {without, with target} = 128 tail evaluations), the 3xTF32
// operand split and the order of the scalar arithmetic are those of k_reward_main_tc; layer 3 sums its three partial
// products in two accumulators instead of one, so R agrees with that kernel to the last few ulps of the layer-3
// outputs, not bit for bit (tests/test_gpu_parity.py).  What changes is who does what and when, and where the layer-3
// operand lives.  The lock-step kernel runs construct -> MMA2 -> epilogue 2 -> MMA3 -> KL with all 16 warps doing
// each phase together between CTA barriers (tensor pipe 43 % active).  Here each warpgroup owns one phase for the
// whole launch and the phases of consecutive samples overlap through mbarriers; no CTA-wide barrier after the prologue:
//
//   warpgroup 0     constructor    features [0, 56)   of A2(g) = relu(h0 + v w_u [+ t w_T]) | 1  -> TMEM (K-steps 0-6)
//   warpgroup 1     constructor    features [56, 104) of A2(g)                                   -> TMEM (K-steps 7-12)
//                                  (the second half has the shorter deadline -- MMA3 + the first half of the next MMA2 --
//                                  and gets the smaller share: 44 real features + the bias column)
//   warpgroup 2     epilogue 2     D2(g) -> ReLU -> hi written over D2 in place, lo next to it (TMEM), then
//                   + KL           D3(g-1) -> KL against the base posteriors -> per-pair accumulator -> R
//   warpgroup 3     issuer         one elected lane of warp 12 issues every tcgen05.mma and tcgen05.commit and does
//                                  nothing else (warps 13-15 only give their registers away): a tcgen05.mma does not
//                                  retire from the issuing warp until the tensor pipe accepts it (measured: issuing
//                                  MMA2 holds the warp for about as long as MMA2 runs), so an issuer that shares its
//                                  warp with other work serialises that work with the MMAs
//
//   tensor pipe:  MMA2(g+1) K-steps 0-6 | K-steps 7-12 | MMA3(g) | MMA2(g+2) ...
//
// Both layers take their A operand from tensor memory.  (A first version kept the layer-3 operand in shared memory as
// the lock-step kernel does: MMA3 then ran at ~70 cycles per instruction instead of 16 -- the tensor core fetches
// shared-memory operands at ~64 B/cycle and an M = 128 A tile is 4 KB per K-step -- and took as long as MMA2.)
// TMEM (512 columns): A2 hi [0,104) lo [104,208), ONE buffer: the constructors of the first K-half refill it for sample
// g+1 as soon as the first seven K-steps of MMA2(g) have completed (tcgen05.commit between the halves), while the other
// half is still being read; X[b] = D2 accumulator / layer-3 operand hi [.., +64) and lo [+64, +120), two buffers at 208
// and 328; D3 at 448 (64 columns: layer 3 issues a_hi * [b_hi; b_lo] as ONE N = 64 instruction and a_lo * b_hi as an
// N = 32 one -- 14 instead of 21 instructions per sample, an M = 128 tf32 MMA costs ~40 cycles for any N <= 64 -- and the
// KL adds the two column halves).  X alternates by sample and the in-order tensor pipe orders MMA3(g) before
// MMA2(g+2), so epilogue 2 needs no "empty" handshake.
//
// g counts the samples of all tiles of this CTA (tile-major), so the pipeline does not drain at tile ends.  The
// constructors keep h0 and w_u of their row in registers (112 / 88 per thread): setmaxnreg moves registers from
// warpgroup 3 to warpgroups 0,1.  Each role prefetches its own per-sample inputs (v, t; base posteriors of the
// with-target rows) with cp.async into a private ring: no cross-thread hand-over, no barrier.
//
// Handshakes (producer -> consumer, arrivals per phase):
//   full_H[h]   constructors of half h (128)    -> issuer          empty_H[h]   tcgen05.commit after the half's K-steps -> constructors
//   full_D2[b]  tcgen05.commit after MMA2       -> epilogue 2      full_A3[b]   epilogue 2 (128)          -> issuer
//   full_D3     tcgen05.commit after MMA3       -> KL              (no empty_D3: the KL of sample g-1 precedes the full_A3(g)
//                                                                   arrival in the same threads, so full_A3(g) implies D3 was read)
// Every wait is bounded in time (status word + a CTA-wide abort flag: after one time-out all waits fall through).
#include <cuda_pipeline.h>

#include <type_traits>

#include "pcvae_reward.cuh"
#include "pcvae_tc.cuh"

namespace pcvae {
namespace rws {

using namespace tc;

constexpr int ROWS = 128;                 // UMMA M
constexpr int K2 = 104, C2 = K2 / 4;      // layer-2 reduction (100 + bias + pad), 16-byte K chunks of B2
constexpr int N2 = 64;                    // layer-2 outputs (50 + ones column + pad)
constexpr int K3 = 56, C3 = K3 / 4;       // layer-3 reduction (50 + bias + pad)
constexpr int N3 = 32;                    // layer-3 outputs (20 + pad)
constexpr int TMEM_COLS = 512;
constexpr int COL_X = 2 * K2, X_COLS = N2 + K3;   // X[b] at COL_X + b * X_COLS: accumulator / hi [0,64), lo [64,120)
constexpr int COL_D3 = COL_X + 2 * X_COLS;        // D3: [0,32) = hi*hi + lo*hi, [32,64) = hi*lo (one buffer)
static_assert(COL_D3 + 2 * N3 == TMEM_COLS, "TMEM map");
constexpr int KH0 = 56, KH1 = K2 - KH0;   // features of the two constructor warpgroups (K-steps 0-6 and 7-12)
constexpr int VT_NB = 4, VT_AHEAD = 3;    // ring of (v, t) per constructor thread
constexpr int BT_NB = 3, BT_AHEAD = 2;    // ring of base posteriors per with-target row
constexpr int BPITCH = BASEW + 4;         // 44 floats: 16-byte reads of consecutive rows hit distinct bank groups
constexpr int NCON = 256;                 // constructor threads
// setmaxnreg: the kernel launches with 128 registers per thread (512 threads; the launcher checks it); warpgroup 3
// keeps 56 (it frees 128 * 72 = 9 216), warpgroups 0 and 1 grow to 160 (they take 256 * 32 = 8 192)
constexpr int REG_LAUNCH = 128, REG_CON = 160, REG_ISSUE = 56;

enum { FULL_H = 0, EMPTY_H = 2, FULL_D2 = 4, FULL_A3 = 6, FULL_D3 = 8, NBAR = 9 };

struct Ctl {
    int* status;
    volatile int* abort_s;
};

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded in TIME (2 s), not in polls: try_wait is given a suspend-time hint, so how long one poll takes is up to the
// hardware; a warp parked in try_wait leaves the issue slots of its scheduler to the warps that have work
__device__ __forceinline__ void wait_on(uint64_t* bar, uint32_t parity, const Ctl& c) {
    const uint32_t addr = smem_u32(bar);
    uint64_t t0 = 0;
#pragma unroll 1
    for (int i = 0;; ++i) {            // not unrolled: the hot loops of all roles have to stay in the instruction cache
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(100000u) : "memory");
        if (ok) return;
        if ((i & 15) == 15) {
            if (*c.abort_s) return;
            const uint64_t t = globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) break;
        }
    }
    *c.abort_s = 1;
    if (c.status) { *reinterpret_cast<volatile int*>(c.status) = 5; __threadfence_system(); }
}

__device__ __forceinline__ void st4(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
}
// 16 / 8 accumulator columns without the wait (the caller waits once for all its loads)
__device__ __forceinline__ void ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void ld4_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void ld2_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int R>
__device__ __forceinline__ void regs_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void regs_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

__global__ void __launch_bounds__(NT, 1) k_reward_main_ws(const RewardArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.L.D;
    float* B2_hi = smem;                             // [C2][64][4]
    float* B2_lo = B2_hi + C2 * N2 * 4;
    float* B3_s = B2_lo + C2 * N2 * 4;               // [C3][64][4]
    float* wT_s = B3_s + 2 * C3 * N3 * 4;            // [K2]
    float* b0_s = wT_s + K2;                         // [64][44]   base posterior of the pair's row (without target)
    float* bT_s = b0_s + NPAIR * BPITCH;             // [BT_NB][64][44]   base posteriors with the sampled target
    float* vt_s = bT_s + BT_NB * NPAIR * BPITCH;     // [VT_NB][2][256]
    float* kl_s = vt_s + VT_NB * 2 * NCON;           // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(kl_s + 2 * ROWS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    int* abort_s = reinterpret_cast<int*>(tmem_slot + 1);

    // ---- one-time setup: weights (hi / lo, K-major core-matrix layout), barriers, TMEM ----
    const float* th = a.theta;
    for (int i = tid; i < C2 * N2 * 4; i += NT) {
        const int c = i / (N2 * 4), n = (i >> 2) % N2, k = 4 * c + (i & 3);
        float w = 0.f;
        if (n < H2 && k < H1) w = th[a.L.W2 + n * H1 + k];
        else if (n < H2 && k == H1) w = th[a.L.b2 + n];
        else if (n == H2 && k == H1) w = 1.0f;           // ones column -> bias column of layer 3
        B2_hi[i] = w;
        B2_lo[i] = tf32_lo(w);
    }
    // layer 3: hi and lo images stacked along N ([C3][64][4]: rows 0-31 hi, rows 32-63 lo), so that a_hi * [b_hi; b_lo]
    // is ONE N = 64 instruction (an M = 128 tf32 MMA costs ~40 cycles for any N <= 64: the A read from tensor memory)
    for (int i = tid; i < C3 * N3 * 4; i += NT) {
        const int c = i / (N3 * 4), n = (i >> 2) % N3, k = 4 * c + (i & 3);
        float w = 0.f;
        if (n < LAT2 && k < H2) w = th[a.L.W3 + n * H2 + k];
        else if (n < LAT2 && k == H2) w = th[a.L.b3 + n];
        B3_s[(c * 2 * N3 + n) * 4 + (i & 3)] = w;
        B3_s[(c * 2 * N3 + N3 + n) * 4 + (i & 3)] = tf32_lo(w);
    }
    for (int k = tid; k < K2; k += NT) wT_s[k] = (k < H1) ? th[a.L.W1 + (long)k * D + (D - 1)] : 0.f;
    if (tid == 0) {
        auto init = [&](int b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + b)), "r"(count)); };
        for (int b = 0; b < 2; ++b) {
            init(FULL_H + b, 128); init(EMPTY_H + b, 1); init(FULL_D2 + b, 1); init(FULL_A3 + b, 128);
        }
        init(FULL_D3, 1);
        *abort_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const Ctl ctl{a.status, abort_s};

    const int wg = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane, pi = row & (NPAIR - 1);
    const bool withT = row >= NPAIR;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);

    const int ptot = a.off[a.N];
    const int ntiles = (ptot + NPAIR - 1) / NPAIR;
    const int nt_cta = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int M = a.M;
    const int G = nt_cta * M;                         // samples this CTA walks through, tile-major
    // (row n, candidate u) of this thread's pair in the it-th tile of this CTA; pairs past the end evaluate pair (0, 0)
    auto pair_of = [&](int it, int& n, int& u) -> bool {
        const int p = (blockIdx.x + it * gridDim.x) * NPAIR + pi;
        n = 0; u = 0;
        if (p >= ptot) return false;
        const int pr = a.pairs[p];
        n = pr / CANDP;
        u = pr - n * CANDP;
        return true;
    };

    if (wg < 2) {
        // ------------------------------------------------------------------------------------------
        // constructors: warpgroup 0 features [0, 56), warpgroup 1 features [56, 104) of this thread's row
        // ------------------------------------------------------------------------------------------
        regs_inc<REG_CON>();
        auto role = [&](auto k0c, auto knc) {
            constexpr int K0 = decltype(k0c)::value, KN = decltype(knc)::value;
            const int h = K0 == 0 ? 0 : 1;
            float* vt_mine = vt_s + tid;                 // slot s: v at [s][0][tid], t at [s][1][tid]
            int pf_it = 0, pf_m = 0, pf_n = 0, pf_u = 0, pf_slot = 0;
            if (nt_cta > 0) pair_of(0, pf_n, pf_u);
            auto prefetch = [&]() {
                if (pf_it < nt_cta) {
                    const float* r = a.im + (long)pf_m * a.im_ss + (long)pf_n * D;
                    __pipeline_memcpy_async(vt_mine + (pf_slot * 2) * NCON, r + pf_u, 4);
                    if (withT) __pipeline_memcpy_async(vt_mine + (pf_slot * 2 + 1) * NCON, r + (D - 1), 4);
                    pf_slot = pf_slot + 1 == VT_NB ? 0 : pf_slot + 1;
                    if (++pf_m == M) {
                        pf_m = 0;
                        if (++pf_it < nt_cta) pair_of(pf_it, pf_n, pf_u);
                    }
                }
                __pipeline_commit();       // one group per call (possibly empty) keeps wait_prior counts uniform
            };
#pragma unroll
            for (int i = 0; i < VT_AHEAD; ++i) prefetch();
            uint32_t g = 0;
            int slot = 0;
            const uint32_t ah0 = lane_addr + K0, al0 = ah0 + K2;
            for (int it = 0; it < nt_cta; ++it) {
                int n, u;
                pair_of(it, n, u);
                float H0r[KN], Ur[KN];
                {
                    const float* h0 = a.base_in + (long)n * H1;
#pragma unroll
                    for (int j = 0; j < KN; ++j) {
                        const int k = K0 + j;            // compile-time: the bias and pad columns cost no registers
                        float hv = 0.f, uv = 0.f;
                        if (k < H1) { hv = h0[k]; uv = __ldg(th + a.L.W1 + (long)k * D + u); }
                        else if (k == H1) hv = 1.0f;          // bias column
                        H0r[j] = hv;
                        Ur[j] = uv;
                    }
                }
                for (int m = 0; m < M; ++m, ++g) {
                    prefetch();
                    __pipeline_wait_prior(VT_AHEAD);          // the copies of sample g landed
                    const float v = vt_mine[(slot * 2) * NCON];
                    const float t = withT ? vt_mine[(slot * 2 + 1) * NCON] : 0.f;
                    slot = slot + 1 == VT_NB ? 0 : slot + 1;
                    if (g >= 1) {                             // the K-steps of MMA2(g - 1) that read this half have completed
                        wait_on(bars + EMPTY_H + h, (g - 1) & 1u, ctl);
                        tc_fence_after();
                    }
#pragma unroll
                    for (int j0 = 0; j0 < KN; j0 += 16) {
                        const int cnt = KN - j0 >= 16 ? 16 : 8;
                        float hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < cnt) hi[j] = fmaf(v, Ur[j0 + j], H0r[j0 + j]);
                        // rows without the target (TMEM lane quarters 0, 1) have t = 0: fmaf(0, w, pre) == pre, and the
                        // warp-uniform branch skips the products and the loads of w_T
                        if (withT) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                if (j < cnt) {
                                    const float4 w4 = *reinterpret_cast<const float4*>(wT_s + K0 + j0 + j);
                                    hi[j] = fmaf(t, w4.x, hi[j]); hi[j + 1] = fmaf(t, w4.y, hi[j + 1]);
                                    hi[j + 2] = fmaf(t, w4.z, hi[j + 2]); hi[j + 3] = fmaf(t, w4.w, hi[j + 3]);
                                }
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < cnt) {
                                hi[j] = fmaxf(hi[j], 0.f);
                                lo[j] = tf32_lo(hi[j]);
                            }
                        if (cnt == 16) { tmem_st16(ah0 + j0, hi); tmem_st16(al0 + j0, lo); }
                        else { tmem_st8(ah0 + j0, hi); tmem_st8(al0 + j0, lo); }
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(bars + FULL_H + h);
                }
            }
            __pipeline_wait_prior(0);
        };
        using IC0 = std::integral_constant<int, 0>;
        using ICA = std::integral_constant<int, KH0>;
        using ICB = std::integral_constant<int, KH1>;
        if (wg == 0) role(IC0{}, ICA{});
        else role(ICA{}, ICB{});
    } else if (wg == 2) {
        // ------------------------------------------------------------------------------------------
        // epilogue 2 of sample g, then the KL of sample g - 1 (off the tensor pipe's critical path: MMA3(g) needs only
        // the operand written by epilogue 2).  Keeps the launch allocation of 128 registers.
        // ------------------------------------------------------------------------------------------
        float* b0_mine = b0_s + pi * BPITCH;
        float* bT_mine = bT_s + pi * BPITCH;            // slot s at + s * NPAIR * BPITCH
        int pf_it = 0, pf_m = 0, pf_n = 0, pf_u = 0, pf_slot = 0;
        if (nt_cta > 0) pair_of(0, pf_n, pf_u);
        auto prefetch = [&]() {                          // base posteriors of a with-target row, BT_AHEAD samples ahead
            if (withT && pf_it < nt_cta) {
                const float* src = a.baseT + ((long)pf_n * M + pf_m) * BASEW;
                float* dst = bT_mine + pf_slot * NPAIR * BPITCH;
#pragma unroll
                for (int c = 0; c < BASEW / 4; ++c) __pipeline_memcpy_async(dst + 4 * c, src + 4 * c, 16);
                pf_slot = pf_slot + 1 == BT_NB ? 0 : pf_slot + 1;
                if (++pf_m == M) {
                    pf_m = 0;
                    if (++pf_it < nt_cta) pair_of(pf_it, pf_n, pf_u);
                }
            }
            __pipeline_commit();
        };
#pragma unroll
        for (int i = 0; i < BT_AHEAD; ++i) prefetch();
        // KL state: walks the same (tile, sample) sequence one sample behind epilogue 2
        uint32_t kg = 0;
        int k_it = 0, k_m = 0, k_n = 0, k_u = 0, slot = 0;
        bool k_valid = false;
        float acc = 0.f;
        auto kl_step = [&]() {
            if (k_m == 0) {
                k_valid = pair_of(k_it, k_n, k_u);
                if (!withT) {                              // this thread is the only reader of its b0 row: no barrier
                    const float4* src = reinterpret_cast<const float4*>(a.base0 + (long)k_n * BASEW);
#pragma unroll
                    for (int c = 0; c < BASEW / 4; ++c) *reinterpret_cast<float4*>(b0_mine + 4 * c) = src[c];
                }
                acc = 0.f;
            }
            prefetch();
            __pipeline_wait_prior(BT_AHEAD);
            const float* bt = withT ? bT_mine + slot * NPAIR * BPITCH : b0_mine;
            slot = slot + 1 == BT_NB ? 0 : slot + 1;
            float bv[BASEW];
#pragma unroll
            for (int c = 0; c < BASEW / 4; ++c) {
                const float4 x4 = *reinterpret_cast<const float4*>(bt + 4 * c);
                bv[4 * c] = x4.x; bv[4 * c + 1] = x4.y; bv[4 * c + 2] = x4.z; bv[4 * c + 3] = x4.w;
            }
            wait_on(bars + FULL_D3, kg & 1u, ctl);
            tc_fence_after();
            uint32_t o[LAT2], o2[LAT2];
            ld16_nowait(lane_addr + COL_D3, o);
            ld4_nowait(lane_addr + COL_D3 + 16, o + 16);
            ld16_nowait(lane_addr + COL_D3 + N3, o2);
            ld4_nowait(lane_addr + COL_D3 + N3 + 16, o2 + 16);
            ld_wait();
            tc_fence_before();
#pragma unroll
            for (int l = 0; l < LAT2; ++l) o[l] = __float_as_uint(__uint_as_float(o[l]) + __uint_as_float(o2[l]));
            float sum = 0.f;
#pragma unroll
            for (int l = 0; l < LAT; ++l) {
                const float mu = __uint_as_float(o[l]), lv = __uint_as_float(o[LAT + l]);
                const float dm = mu - bv[l];
                sum += (((dm * dm) * bv[2 * LAT + l] + expf(lv) * bv[3 * LAT + l] - 1.0f) - lv) + bv[LAT + l];
            }
            float* kb = kl_s + (kg & 1u) * ROWS;
            kb[row] = 0.5f * sum;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (!withT) {              // approx_KL += KL_I; approx_KL -= KL_II   (evaluate.py:537-538)
                acc += kb[pi];
                acc -= kb[NPAIR + pi];
            }
            ++kg;
            if (++k_m == M) {
                if (!withT && k_valid) a.R[(long)k_n * (D - 1) + k_u] = acc / (float)M;   // evaluate.py:540
                k_m = 0;
                ++k_it;
            }
        };
        {       // lo columns [48, 56) of both layer-3 operand buffers: 48, 49 are rewritten per sample, 50-55 stay zero
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            tmem_st8(lane_addr + COL_X + N2 + 48, z);
            tmem_st8(lane_addr + COL_X + X_COLS + N2 + 48, z);
            tmem_st_wait();
        }
        for (uint32_t g = 0; g < (uint32_t)G; ++g) {
            const uint32_t b = g & 1u;
            const uint32_t xa = lane_addr + COL_X + X_COLS * b;
            wait_on(bars + FULL_D2 + b, (g >> 1) & 1u, ctl);
            tc_fence_after();
            // columns [0, 50) only: column 50 is the constant 1 the MMA itself produced (its hi is already in place, its
            // lo is 0) and columns 51-55 are exact zeros (zero weight rows); their lo columns were zeroed once above
            uint32_t d[H2 + 2];
            ld16_nowait(xa, d);
            ld16_nowait(xa + 16, d + 16);
            ld16_nowait(xa + 32, d + 32);
            ld2_nowait(xa + 48, d + 48);
            ld_wait();
            // ReLU; hi goes back over the accumulator columns it came from (this thread's lane, its own columns), lo next to it
#pragma unroll
            for (int j0 = 0; j0 < H2; j0 += 16) {
                const int cnt = H2 - j0 >= 16 ? 16 : 2;
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < cnt) {
                        hi[j] = fmaxf(__uint_as_float(d[j0 + j]), 0.f);
                        lo[j] = tf32_lo(hi[j]);
                    }
                if (cnt == 16) { tmem_st16(xa + j0, hi); tmem_st16(xa + N2 + j0, lo); }
                else { tmem_st2(xa + j0, hi); tmem_st2(xa + N2 + j0, lo); }
            }
            tmem_st_wait();
            // KL of sample g - 1 BEFORE the arrival: MMA3(g) overwrites the D3 this reads, and MMA3(g) waits for full_A3(g).
            // The issuer is busy with MMA2(g + 1) for longer than both take.
            if (g >= 1) kl_step();
            tc_fence_before();
            mbar_arrive(bars + FULL_A3 + b);
        }
        if (G > 0) kl_step();
        __pipeline_wait_prior(0);
    } else {
        // ------------------------------------------------------------------------------------------
        // warpgroup 3: its registers go to the constructors; one elected lane of warp 12 is the issuer
        // ------------------------------------------------------------------------------------------
        regs_dec<REG_ISSUE>();
        if (warp == 12) {
            constexpr uint32_t IDESC2 = make_idesc(ROWS, N2), IDESC3 = make_idesc(ROWS, N3), IDESC3S = make_idesc(ROWS, 2 * N3);
            constexpr uint32_t SBO = 128, B2_LBO = N2 * 16, B3_LBO = 2 * N3 * 16;
            // descriptors of k-step 0; a k-step (8 tf32 = two 16-byte chunks) advances the 16-byte-unit address field
            const uint64_t dB2h = make_desc(smem_u32(B2_hi), B2_LBO, SBO), dB2l = make_desc(smem_u32(B2_lo), B2_LBO, SBO);
            const uint64_t dB3 = make_desc(smem_u32(B3_s), B3_LBO, SBO);
            constexpr uint64_t B2_STEP = (2 * B2_LBO) >> 4, B3_STEP = (2 * B3_LBO) >> 4;
            const uint32_t ah = tmem, al = tmem + K2;
            // D2 = A2 * B2^T into X[b], 3xTF32, K-steps [KS0, KS1) (single thread)
            auto issue_mma2 = [&](uint32_t b, auto ks0c, auto ks1c) {
                constexpr int KS0 = decltype(ks0c)::value, KS1 = decltype(ks1c)::value;
                const uint32_t dcol = tmem + COL_X + X_COLS * b;
#pragma unroll
                for (int ks = KS0; ks < KS1; ++ks) {
                    mma_tf32_ts(dcol, al + 8 * ks, dB2h + ks * B2_STEP, IDESC2, ks > 0);
                    mma_tf32_ts(dcol, ah + 8 * ks, dB2l + ks * B2_STEP, IDESC2, 1);
                    mma_tf32_ts(dcol, ah + 8 * ks, dB2h + ks * B2_STEP, IDESC2, 1);
                }
            };
            using IC0 = std::integral_constant<int, 0>;
            using ICH = std::integral_constant<int, KH0 / 8>;
            using ICE = std::integral_constant<int, K2 / 8>;
            // sample g: first half as soon as its features are written, then the second; the commit between them
            // hands the first half of the A2 buffer back to its constructors while the second is still being read
            auto mma2_of = [&](uint32_t g) {
                const uint32_t b = g & 1u;
                wait_on(bars + FULL_H, g & 1u, ctl);
                tc_fence_after();
                issue_mma2(b, IC0{}, ICH{});
                mma_commit(bars + EMPTY_H);
                wait_on(bars + FULL_H + 1, g & 1u, ctl);
                tc_fence_after();
                issue_mma2(b, ICH{}, ICE{});
                mma_commit(bars + EMPTY_H + 1);
                mma_commit(bars + FULL_D2 + b);
            };
            // D3 = relu(D2)[b] * B3^T, 3xTF32 as a_hi * [b_hi; b_lo] (N = 64) + a_lo * b_hi (N = 32), A from tensor memory
            auto mma3_of = [&](uint32_t g) {
                const uint32_t b = g & 1u;
                const uint32_t xh = tmem + COL_X + X_COLS * b, xl = xh + N2, dcol = tmem + COL_D3;
                wait_on(bars + FULL_A3 + b, (g >> 1) & 1u, ctl);      // epilogue 2 wrote the operand (and the KL read D3(g - 1))
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < K3 / 8; ++ks) {
                    mma_tf32_ts(dcol, xh + 8 * ks, dB3 + ks * B3_STEP, IDESC3S, ks > 0);
                    mma_tf32_ts(dcol, xl + 8 * ks, dB3 + ks * B3_STEP, IDESC3, 1);
                }
                mma_commit(bars + FULL_D3);
            };
            if (elect_one()) {
                // MMA2(0), then per sample MMA2(g + 1) | MMA3(g); one loop so that the issue code exists once
                for (uint32_t g = 0; g <= (uint32_t)G; ++g) {
                    if (g < (uint32_t)G) mma2_of(g);
                    if (g >= 1) mma3_of(g - 1);
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

static size_t smem_bytes() {
    size_t f = 2 * (size_t)C2 * N2 * 4 + 2 * (size_t)C3 * N3 * 4 + K2 + (size_t)NPAIR * BPITCH +
               (size_t)BT_NB * NPAIR * BPITCH + (size_t)VT_NB * 2 * NCON + 2 * ROWS;
    return f * sizeof(float) + NBAR * sizeof(uint64_t) + 16 + 128;
}

}  // namespace rws

int reward_main_ws_launch(const RewardArgs& a, int grid, cudaStream_t st) {
    const size_t sm = rws::smem_bytes();
    if (sm > MAX_SMEM) return fail(PCVAE_EINVAL, "reward_main_ws: shared memory %zu B exceeds %d", sm, MAX_SMEM);
    cudaError_t e = cudaFuncSetAttribute(rws::k_reward_main_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reward_main_ws: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    // setmaxnreg.inc waits for registers that only exist if the launch allocation is what the budget assumes
    static int regs = -1;
    if (regs < 0) {
        cudaFuncAttributes fa;
        if ((e = cudaFuncGetAttributes(&fa, rws::k_reward_main_ws)) != cudaSuccess) return fail(PCVAE_ECUDA, "reward_main_ws: cudaFuncGetAttributes: %s", cudaGetErrorString(e));
        regs = fa.numRegs;
    }
    if (regs != rws::REG_LAUNCH) return fail(PCVAE_EINVAL, "reward_main_ws: built with %d registers per thread, the register budget needs %d", regs, rws::REG_LAUNCH);
    rws::k_reward_main_ws<<<grid, NT, sm, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "reward_main_ws: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // namespace pcvae
