// Row-tile scaffolding shared by the tcgen05 training kernels (decoder: pcvae_dec_tc.cu, encoder: pcvae_enc_tc.cu):
// 128-row tiles, 512 threads = 4 TMEM lane quarters x 4 column groups, activation operands in TENSOR MEMORY as
// hi / lo tf32 images, accumulators in TMEM, weights in shared memory as K-major no-swizzle core-matrix images.
//
// TMEM columns (all 512):  RA [0,224): 112 hi + 112 lo    RB [224,336): 56 hi + 56 lo
//                          ACC1 [336,448)                 ACC2 [448,512)
#pragma once
#include "pcvae_tc.cuh"

namespace pcvae {
namespace tc {

constexpr int ROWS = 128;
constexpr int RA_HI = 0, RA_LO = 112, RB_HI = 224, RB_LO = 280, ACC1 = 336, ACC2 = 448;
constexpr int DEC_ISSUER_WARP = 4;

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 3xTF32 product: activation operand in TMEM (hi at a_hi, lo at a_lo), weight image (K-major) in shared memory;
// a k-step (8 tf32) is two 16-byte chunks of the image
__device__ __forceinline__ void issue_3x(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint64_t b_step,
                                         int ksteps, uint32_t idesc) {
    for (int ks = 0; ks < ksteps; ++ks) {
        mma_tf32_ts(acc, a_lo + 8 * ks, b_hi + ks * b_step, idesc, ks > 0);
        mma_tf32_ts(acc, a_hi + 8 * ks, b_lo + ks * b_step, idesc, 1);
        mma_tf32_ts(acc, a_hi + 8 * ks, b_hi + ks * b_step, idesc, 1);
    }
}

// four mask entries: uint8 -> 0/1; float32 -> 0/1 (RAW = false) or the stored values (RAW = true: the encoders multiply
// x by the mask as it is, VAE.py:388)
template <bool RAW = false>
__device__ __forceinline__ void load_mask4(const void* m, long gi, int kind, float* o) {
    if (kind == PCVAE_MASK_U8) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const unsigned char*>(m) + gi);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = ((w >> (8 * j)) & 0xFFu) ? 1.f : 0.f;
    } else {
        const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(m) + gi);
        if (RAW) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
        else { o[0] = v.x != 0.f ? 1.f : 0.f; o[1] = v.y != 0.f ? 1.f : 0.f; o[2] = v.z != 0.f ? 1.f : 0.f; o[3] = v.w != 0.f ? 1.f : 0.f; }
    }
}

// hi / lo image of a weight matrix in the K-major no-swizzle core-matrix layout [k/4][row][4].  The caller zeroes the
// images first (zero_images), then scatters the real entries: global reads run along the nn.Linear rows (coalesced),
// the shared-memory writes take whatever banks they hit (a few dozen elements per thread, once per launch).
__device__ __forceinline__ void zero_images(float* base, int floats, int tid) {
    float4* p = reinterpret_cast<float4*>(base);
    for (int i = tid; i < floats / 4; i += NT) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void image_put(float* hi, float* lo, int nrows, int row, int col, float v) {
    const int dst = ((col >> 2) * nrows + row) * 4 + (col & 3);
    hi[dst] = v;
    lo[dst] = tf32_lo(v);
}
// One nn.Linear weight matrix W [N][K] (row-major) into an image.  A warp takes one 16-byte chunk column of the image
// at a time and its lanes run along the image rows, so every shared-memory store is a conflict-free STS.128 (a
// scalar scatter along K puts the eight chunks of a warp into the same banks: image rows are 16 B apart, chunks
// nrows * 16 B); all loads of a thread are requested before its first store.
//   forward   : image row n, chunk c = W[n][4c .. 4c+3]   (one 16-byte load when W + n*K is 16-byte aligned)
//   transposed: image row k, chunk c = W[4c .. 4c+3][k]   (four loads that are coalesced along k)
template <bool TRANSPOSED>
__device__ __forceinline__ void image_scatter(float* hi, float* lo, int nrows, const float* __restrict__ W, int N, int K, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int R = TRANSPOSED ? K : N;                     // image rows
    const int Cn = TRANSPOSED ? N : K;                    // image columns (reduction index)
    const int nch = (Cn + 3) >> 2, rg = (R + 31) >> 5;    // chunks, row groups of 32
    const bool vec = !TRANSPOSED && ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    for (int base = warp; base < nch * rg; base += 4 * NWARP) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = base + u * NWARP;
            const int c = it / rg, r = (it - c * rg) * 32 + lane;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it < nch * rg && r < R) {
                if (vec) {
                    v[u] = __ldg(reinterpret_cast<const float4*>(W + (long)r * K + 4 * c));
                } else if (!TRANSPOSED) {
                    const float* w = W + (long)r * K + 4 * c;
                    v[u].x = __ldg(w);
                    if (4 * c + 1 < K) v[u].y = __ldg(w + 1);
                    if (4 * c + 2 < K) v[u].z = __ldg(w + 2);
                    if (4 * c + 3 < K) v[u].w = __ldg(w + 3);
                } else {
                    const float* w = W + (long)(4 * c) * K + r;
                    v[u].x = __ldg(w);
                    if (4 * c + 1 < N) v[u].y = __ldg(w + K);
                    if (4 * c + 2 < N) v[u].z = __ldg(w + 2 * K);
                    if (4 * c + 3 < N) v[u].w = __ldg(w + 3 * K);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = base + u * NWARP;
            const int c = it / rg, r = (it - c * rg) * 32 + lane;
            if (it < nch * rg && r < R) {
                const int dst = (c * nrows + r) * 4;
                *reinterpret_cast<float4*>(hi + dst) = v[u];
                *reinterpret_cast<float4*>(lo + dst) = make_float4(tf32_lo(v[u].x), tf32_lo(v[u].y), tf32_lo(v[u].z), tf32_lo(v[u].w));
            }
        }
    }
}
// forward image of y = W x + b with W [N][K] row-major: image row n, column k; the bias sits at column K and, when
// `one` is set, image row N holds a 1 at column K (it regenerates the constant-1 column for the next layer)
__device__ __forceinline__ void image_linear(float* hi, float* lo, int nrows, const float* __restrict__ W,
                                             const float* __restrict__ b, int N, int K, bool one, int tid) {
    image_scatter<false>(hi, lo, nrows, W, N, K, tid);
    __syncthreads();            // when K % 4 != 0 the bias column shares its 16-byte chunk with the last weight columns
    for (int n = tid; n < N; n += NT) image_put(hi, lo, nrows, n, K, __ldg(b + n));
    if (one && tid == 0) image_put(hi, lo, nrows, N, K, 1.0f);
}
// data-gradient image (transposed weights): image row = layer INPUT index k, image column (reduction) = OUTPUT index n
__device__ __forceinline__ void image_linear_T(float* hi, float* lo, int nrows, const float* __restrict__ W, int N, int K, int tid) {
    image_scatter<true>(hi, lo, nrows, W, N, K, tid);
}

// The shared-memory matrix descriptors of a kernel live in SHARED MEMORY and the issuing lane reads the two it needs
// with volatile loads right before each MMA batch.  Descriptors built from immediates in registers are miscompiled
// by ptxas 12.9.86 in some code shapes: the constant high word (stride-byte-offset and version bits) of descriptor
// pairs of different layers is merged into one uniform register whose UMOV is placed behind a later use inside the
// work-item loop, so the FIRST item of a CTA issues its first MMA batch with SBO = 0 (image rows >= 8 alias rows
// 0..7; found with profiles/r01_ncu_summary.md "descriptor high word", guarded by tests/test_abi.py's SASS check).
__device__ __forceinline__ void st_desc(uint64_t* slot, uint64_t d) { *reinterpret_cast<volatile uint64_t*>(slot) = d; }
__device__ __forceinline__ uint64_t ld_desc(const uint64_t* slot) {
    uint64_t v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(slot)) : "memory");
    return v;
}

// Prebuilt weight images (pcvae_build_weight_images): the whole block of a kernel arrives by bulk async copies -- one L2
// round trip instead of the dependent loads, splits and scattered stores of the build (~5 us of a 60 us kernel).
// Initialises `bar` (one use), contains __syncthreads().
__device__ __forceinline__ void fetch_images(float* dst, const float* __restrict__ src, uint32_t bytes, uint64_t* bar, int tid, int* status) {
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        mbar_expect_tx(bar, bytes);
        for (uint32_t off = 0; off < bytes; off += 32768) {
            const uint32_t nb = bytes - off < 32768 ? bytes - off : 32768;
            bulk_g2s(dst + off / 4, src + off / 4, nb, bar);
        }
    }
    __syncthreads();
    mbar_wait(bar, 0, status, 3);
}

struct TileCtx {
    uint32_t tmem, lane_addr, ph;
    int q, cg, row, c28, c16;
    int* status;                      // tensor-core status word of the device (bounded waits report here)
};

// barrier + (one elected lane) MMA issue + commit; every thread then waits for the batch.  mma_kick / mma_wait are the
// two halves: independent work (global loads for the next epilogue) placed between them overlaps the tensor pipe.
template <typename Issue>
__device__ __forceinline__ void mma_kick(uint64_t* bar, int warp, Issue&& issue) {
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == DEC_ISSUER_WARP) {
        tc_fence_after();
        if (elect_one()) {
            issue();
            mma_commit(bar);
        }
        __syncwarp();
    }
}
__device__ __forceinline__ void mma_wait(TileCtx& cx, uint64_t* bar) {
    mbar_wait(bar, cx.ph, cx.status, 2);
    cx.ph ^= 1;
    tc_fence_after();
}
template <typename Issue>
__device__ __forceinline__ void run_mma(TileCtx& cx, uint64_t* bar, int warp, Issue&& issue) {
    mma_kick(bar, warp, issue);
    mma_wait(cx, bar);
}

__device__ __forceinline__ void tc_setup(TileCtx& cx, uint64_t* bar, uint32_t* slot, int tid, int* status, uint32_t ncols = 512) {
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cx.tmem = *slot;
    cx.ph = 0;
    cx.status = status;
    cx.q = warp & 3;
    cx.cg = warp >> 2;
    cx.row = 32 * cx.q + lane;
    cx.lane_addr = cx.tmem + ((uint32_t)(32 * cx.q) << 16);
    cx.c28 = 28 * cx.cg;
    cx.c16 = 16 * cx.cg;
}

__device__ __forceinline__ void tc_teardown(const TileCtx& cx, int tid, uint32_t ncols = 512) {
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "r"(ncols));
    }
}

// all 28 accumulator columns of this thread with the three loads in flight together (one TMEM round trip instead of three)
__device__ __forceinline__ void tmem_ld28(uint32_t taddr, float* v) {
    uint32_t r[28];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
                 : "r"(taddr + 16));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]) : "r"(taddr + 24));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 28; ++i) v[i] = __uint_as_float(r[i]);
}

// 28 accumulator columns of this thread, in parts of 16 / 8 / 4
__device__ __forceinline__ void ld_part(uint32_t addr, int part, float* v) {
    if (part == 0) tmem_ld16(addr, v);
    else if (part == 1) tmem_ld8(addr + 16, v);
    else tmem_ld4(addr + 24, v);
}
__device__ __forceinline__ void st_part(uint32_t addr, int part, const float* v) {
    if (part == 0) tmem_st16(addr, v);
    else if (part == 1) tmem_st8(addr + 16, v);
    else tmem_st4(addr + 24, v);
}

// bits j < 32 with first + j < limit (the columns of a thread's run that are real features)
__device__ __forceinline__ uint32_t col_bits(int first, int limit) {
    const int n = limit - first;
    return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// `cnt` consecutive features (first, first + 1, ...) of this thread's row into the feature-major scratch; `col0` points
// at the row's slot of feature 0, features >= limit are skipped.  One 64-bit base and immediate offsets (feature f + 1
// lives 32 floats after feature f); `first` is warp-uniform, so the all-valid test is a uniform branch and the common
// case carries no per-element predicate or address arithmetic.
__device__ __forceinline__ void scratch_store(float* __restrict__ col0, int first, int limit, const float* v, int cnt) {
    float* __restrict__ p = col0 + (size_t)first * 32;
    if (first + cnt <= limit) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < cnt) p[j * 32] = v[j];
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < cnt && j < limit - first) p[j * 32] = v[j];
    }
}

}  // namespace tc
}  // namespace pcvae
