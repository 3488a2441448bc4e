// Device-side batch preparation for throughput mode: gather a batch
// of table rows by a permutation index (what DataLoader(shuffle=True) + default collate do on
// the host, src/utils/loaders.py:342-352,389-397), draw the per-batch sub-mask
// mask_p = mask & (u < 1 - p/100) (src/utils/utils.py:36-39, src/experiment_main/train.py:53-55)
// and the N(0,1) noise of the two rsample() calls (src/models/VAE.py:390-392) with Philox.
// Parity mode feeds host-generated mask_p / eps instead (the CPU generators cannot be
// reproduced by Philox), see DESIGN.md.
#include <curand_kernel.h>

#include "pcvae_internal.cuh"
#include "pcvae_philox.cuh"

namespace pcvae {

__global__ void k_gather_rows(const float* __restrict__ table, const void* __restrict__ mtable,
                              const long* __restrict__ idx, float* __restrict__ x, void* __restrict__ mask,
                              int B, int D, int mask_kind) {
    const long total = (long)B * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / D), d = (int)(i - (long)b * D);
        const long src = idx[b] * D + d;
        x[i] = table[src];
        if (mask_kind == PCVAE_MASK_U8) reinterpret_cast<uint8_t*>(mask)[i] = reinterpret_cast<const uint8_t*>(mtable)[src];
        else reinterpret_cast<float*>(mask)[i] = reinterpret_cast<const float*>(mtable)[src];
    }
}

__global__ void k_draw_submask(const uint8_t* __restrict__ mask, uint8_t* __restrict__ mask_p, long n, float keep,
                               unsigned long long seed, unsigned long long offset) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nthreads = (long)gridDim.x * blockDim.x;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, t, offset, &st);
    for (long i = 4 * t; i < n; i += 4 * nthreads) {
        const float4 u = curand_uniform4(&st);          // (0,1]
        const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i + j < n) mask_p[i + j] = (mask[i + j] && (1.0f - uv[j]) < keep) ? 1 : 0;   // rand() in [0,1) < keep
    }
}

__global__ void k_draw_normal(float* __restrict__ out, long n, unsigned long long seed, unsigned long long offset) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nthreads = (long)gridDim.x * blockDim.x;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, t, offset, &st);
    for (long i = 4 * t; i < n; i += 4 * nthreads) {
        const float4 g = curand_normal4(&st);
        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i + j < n) out[i + j] = gv[j];
    }
}

// Fused batch preparation (one launch instead of three): a warp per table row copies the row of x and of the
// mask with 16-byte / 4-byte vector accesses and draws the sub-mask for its entries; a second, flat loop draws the
// 2 x 10 standard-normal values of every row (Box-Muller), one thread per (row, group of four values).  Philox counter =
// (row, lane | 64 + group, offset), key = seed: what a row gets does not depend on how rows are dealt to warps.
// A warp has PREP_R rows in flight (their indices first, then all their loads: two dependent memory round trips per
// PREP_R rows), and the grid is exactly what is resident at once (prep_grid) -- no second, partial wave.
// Requires D % 4 == 0 and D <= 128, uint8 masks.
// (build-time knobs for tools/build_variant.sh; measured: 4 rows, 4 CTAs per SM = 26.9 us, profiles/r02_ncu_summary.md)
#ifndef PCVAE_PREP_R
#define PCVAE_PREP_R 4
#endif
#ifndef PCVAE_PREP_MINB
#define PCVAE_PREP_MINB 4
#endif
constexpr int PREP_R = PCVAE_PREP_R;
__global__ void __launch_bounds__(256, PCVAE_PREP_MINB) k_prep_batch(const float* __restrict__ table, const uint8_t* __restrict__ mtable,
                                                    const long* __restrict__ idx, float* __restrict__ x,
                                                    uint8_t* __restrict__ mask, uint8_t* __restrict__ mask_p,
                                                    float* __restrict__ eps, int B, int D, int n_eps, float keep,
                                                    unsigned long long seed, unsigned long long offset,
                                                    const unsigned long long* __restrict__ step_state, long n_batches) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int D4 = D >> 2;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    if (step_state) {        // replayable launch: batch number and Philox offset come from the device step counter
        const unsigned long long st = *step_state;
        idx += (long)(st % (unsigned long long)n_batches) * B;
        offset += st * 8ull;
    }
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int base = w0; base < B; base += PREP_R * warps) {
        long s[PREP_R];
#pragma unroll
        for (int k = 0; k < PREP_R; ++k) {
            const int b = base + k * warps;
            s[k] = b < B ? idx[b] : -1;
        }
        if (lane < D4) {
            float4 v[PREP_R];
            uint32_t m[PREP_R];
#pragma unroll
            for (int k = 0; k < PREP_R; ++k)
                if (s[k] >= 0) {
                    v[k] = reinterpret_cast<const float4*>(table + s[k] * D)[lane];
                    m[k] = reinterpret_cast<const uint32_t*>(mtable + s[k] * D)[lane];
                }
#pragma unroll
            for (int k = 0; k < PREP_R; ++k) {
                if (s[k] < 0) continue;
                const int b = base + k * warps;
                reinterpret_cast<float4*>(x + (long)b * D)[lane] = v[k];
                reinterpret_cast<uint32_t*>(mask + (long)b * D)[lane] = m[k];
                const uint4 r = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)lane, (uint32_t)offset, (uint32_t)(offset >> 32)), key);
                const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
                uint32_t mp = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (((m[k] >> (8 * j)) & 0xFFu) && u01(rv[j]) < keep) mp |= 1u << (8 * j);    // rand() in [0,1) < keep
                reinterpret_cast<uint32_t*>(mask_p + (long)b * D)[lane] = mp;
            }
        }
    }
    // n_eps * 10 normals per row, four per thread: item = (row, group q), the counter a lane q of the row's warp used to take
    const int Q = (10 * n_eps + 3) / 4;
    const long items = (long)B * Q, nthreads = (long)gridDim.x * blockDim.x;
    for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += nthreads) {
        const int b = (int)(it / Q), q = (int)(it - (long)b * Q);
        const uint4 r = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(64 + q), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
        float gv[4];
        const float r0 = sqrtf(-2.0f * logf(u01_open(r.x))), r1 = sqrtf(-2.0f * logf(u01_open(r.z)));
        float s0, c0, s1, c1;
        sincospif(2.0f * u01(r.y), &s0, &c0);
        sincospif(2.0f * u01(r.w), &s1, &c1);
        gv[0] = r0 * c0; gv[1] = r0 * s0; gv[2] = r1 * c1; gv[3] = r1 * s1;
        const int e0 = 4 * q;                            // entry in the row's [n_eps][10] block
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = e0 + j, br = e / 10, l = e - br * 10;
            if (br < n_eps) eps[((long)br * B + b) * 10 + l] = gv[j];
        }
    }
}

// CTAs of k_prep_batch that are resident at once on this device
static int prep_grid(int sms) {
    static int per_sm = 0;
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_prep_batch, 256, 0) != cudaSuccess || n < 1) n = 4;
        per_sm = n;
    }
    return sms * per_sm;
}

// Host-streamed batches: the mask arrives bit-packed (W = ceil(D / 32) words per row, bit j of word w is
// mask[row][32 w + j]) and, optionally, x as the stream of its OBSERVED entries only (row-major, row r starting at
// vals[row_off[r]]; entries under a zero mask bit never reach the training loss, VAE.py:388, 411-445).  A warp per
// row expands both into the dense layouts the row-tile kernels read, and draws the sub-mask and the noise exactly
// as k_prep_batch does (same Philox counters), so a step fed this way equals a step fed by k_prep_batch on the
// same rows.  Requires D % 4 == 0 and D <= 128.
__global__ void __launch_bounds__(256) k_prep_packed(const uint32_t* __restrict__ bits, const float* __restrict__ vals,
                                                     const uint32_t* __restrict__ row_off, float* __restrict__ x,
                                                     uint8_t* __restrict__ mask, uint8_t* __restrict__ mask_p,
                                                     float* __restrict__ eps, int B, int D, int n_eps, float keep,
                                                     unsigned long long seed, unsigned long long offset) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int D4 = D >> 2, W = (D + 31) >> 5;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        if (lane < D4) {
            const uint32_t* rb = bits + (long)b * W;
            const int wi = lane >> 3, sh = (lane & 7) * 4;
            const uint32_t word = rb[wi];
            const uint32_t nib = (word >> sh) & 0xFu;
            const uint32_t m = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);   // four 0/1 bytes
            reinterpret_cast<uint32_t*>(mask + (long)b * D)[lane] = m;
            if (vals) {
                int before = __popc(word & ((1u << sh) - 1u));
                for (int w = 0; w < wi; ++w) before += __popc(rb[w]);
                const float* v = vals + row_off[b] + before;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                int k = 0;
                if (nib & 1u) o.x = v[k++];
                if (nib & 2u) o.y = v[k++];
                if (nib & 4u) o.z = v[k++];
                if (nib & 8u) o.w = v[k++];
                reinterpret_cast<float4*>(x + (long)b * D)[lane] = o;
            }
            if (mask_p) {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)lane, (uint32_t)offset, (uint32_t)(offset >> 32)), key);
                const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
                uint32_t mp = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (((nib >> j) & 1u) && u01(rv[j]) < keep) mp |= 1u << (8 * j);
                reinterpret_cast<uint32_t*>(mask_p + (long)b * D)[lane] = mp;
            }
        }
        if (lane < (10 * n_eps + 3) / 4) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(64 + lane), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
            float gv[4];
            const float r0 = sqrtf(-2.0f * logf(u01_open(r.x))), r1 = sqrtf(-2.0f * logf(u01_open(r.z)));
            float s0, c0, s1, c1;
            sincospif(2.0f * u01(r.y), &s0, &c0);
            sincospif(2.0f * u01(r.w), &s1, &c1);
            gv[0] = r0 * c0; gv[1] = r0 * s0; gv[2] = r1 * c1; gv[3] = r1 * s1;
            const int e0 = 4 * lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = e0 + j, br = e / 10, l = e - br * 10;
                if (br < n_eps) eps[((long)br * B + b) * 10 + l] = gv[j];
            }
        }
    }
}

}  // namespace pcvae

using namespace pcvae;

extern "C" {

int pcvae_gather_rows(const float* table, const void* mask_table, const long* idx, float* x, void* mask, int rows,
                      int obs_dim, int mask_kind, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 0 || obs_dim < 1) return fail(PCVAE_EINVAL, "gather_rows: bad sizes");
    if (rows == 0) return PCVAE_OK;
    if (!table || !mask_table || !idx || !x || !mask) return fail(PCVAE_EINVAL, "gather_rows: null pointer");
    k_gather_rows<<<grid * 8, 256, 0, (cudaStream_t)stream>>>(table, mask_table, idx, x, mask, rows, obs_dim, mask_kind);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "gather_rows: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_prep_batch(const float* table, const uint8_t* mask_table, const long* idx, float* x, uint8_t* mask,
                     uint8_t* mask_p, float* eps, int rows, int obs_dim, int n_eps, float keep_prob,
                     unsigned long long seed, unsigned long long offset, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 0 || obs_dim < 4 || obs_dim > 128 || obs_dim % 4 || n_eps < 0 || n_eps > 2 || !(keep_prob >= 0.f && keep_prob <= 1.f))
        return fail(PCVAE_EINVAL, "prep_batch: bad arguments (obs_dim must be a multiple of 4, <= 128; n_eps 0..2)");
    if (rows == 0) return PCVAE_OK;
    if (!table || !mask_table || !idx || !x || !mask || !mask_p || (n_eps > 0 && !eps)) return fail(PCVAE_EINVAL, "prep_batch: null pointer");
    k_prep_batch<<<prep_grid(grid), 256, 0, (cudaStream_t)stream>>>(table, mask_table, idx, x, mask, mask_p, eps, rows, obs_dim, n_eps,
                                                              keep_prob, seed, offset, nullptr, 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "prep_batch: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_prep_batch_dev(const float* table, const uint8_t* mask_table, const long* idx_batches, long n_batches, float* x,
                         uint8_t* mask, uint8_t* mask_p, float* eps, int rows, int obs_dim, int n_eps, float keep_prob,
                         unsigned long long seed, unsigned long long offset0, const unsigned long long* step_state, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 1 || obs_dim < 4 || obs_dim > 128 || obs_dim % 4 || n_eps < 0 || n_eps > 2 || !(keep_prob >= 0.f && keep_prob <= 1.f) ||
        n_batches < 1)
        return fail(PCVAE_EINVAL, "prep_batch_dev: bad arguments (obs_dim must be a multiple of 4, <= 128; n_eps 0..2; n_batches >= 1)");
    if (!table || !mask_table || !idx_batches || !x || !mask || !mask_p || (n_eps > 0 && !eps) || !step_state)
        return fail(PCVAE_EINVAL, "prep_batch_dev: null pointer");
    k_prep_batch<<<prep_grid(grid), 256, 0, (cudaStream_t)stream>>>(table, mask_table, idx_batches, x, mask, mask_p, eps, rows, obs_dim, n_eps,
                                                              keep_prob, seed, offset0, step_state, n_batches);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "prep_batch_dev: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_prep_packed(const uint32_t* mask_bits, const float* vals, const uint32_t* row_off, float* x, uint8_t* mask,
                      uint8_t* mask_p, float* eps, int rows, int obs_dim, int n_eps, float keep_prob,
                      unsigned long long seed, unsigned long long offset, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (rows < 0 || obs_dim < 4 || obs_dim > 128 || obs_dim % 4 || n_eps < 0 || n_eps > 2 || !(keep_prob >= 0.f && keep_prob <= 1.f))
        return fail(PCVAE_EINVAL, "prep_packed: bad arguments (obs_dim must be a multiple of 4, <= 128; n_eps 0..2)");
    if (rows == 0) return PCVAE_OK;
    if (!mask_bits || !mask || (n_eps > 0 && !eps)) return fail(PCVAE_EINVAL, "prep_packed: null pointer");
    if ((vals != nullptr) != (row_off != nullptr) || (vals && !x))
        return fail(PCVAE_EINVAL, "prep_packed: vals, row_off and x go together (all three or none)");
    k_prep_packed<<<grid * 8, 256, 0, (cudaStream_t)stream>>>(mask_bits, vals, row_off, x, mask, mask_p, eps, rows, obs_dim, n_eps,
                                                               keep_prob, seed, offset);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "prep_packed: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_draw_submask(const uint8_t* mask, uint8_t* mask_p, long n, float keep_prob, unsigned long long seed,
                       unsigned long long offset, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (n < 0 || !(keep_prob >= 0.f && keep_prob <= 1.f)) return fail(PCVAE_EINVAL, "draw_submask: bad arguments");
    if (n == 0) return PCVAE_OK;
    if (!mask || !mask_p) return fail(PCVAE_EINVAL, "draw_submask: null pointer");
    k_draw_submask<<<grid * 4, 256, 0, (cudaStream_t)stream>>>(mask, mask_p, n, keep_prob, seed, offset);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "draw_submask: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_draw_normal(float* out, long n, unsigned long long seed, unsigned long long offset, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (n < 0) return fail(PCVAE_EINVAL, "draw_normal: bad arguments");
    if (n == 0) return PCVAE_OK;
    if (!out) return fail(PCVAE_EINVAL, "draw_normal: null pointer");
    k_draw_normal<<<grid * 4, 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "draw_normal: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

}  // extern "C"
