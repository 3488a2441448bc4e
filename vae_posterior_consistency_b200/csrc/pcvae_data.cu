// Device-side batch preparation for throughput mode: gather a batch
// of table rows by a permutation index (what DataLoader(shuffle=True) + default collate do on
// the host, src/utils/loaders.py:342-352,389-397), draw the per-batch sub-mask
// mask_p = mask & (u < 1 - p/100) (src/utils/utils.py:36-39, src/experiment_main/train.py:53-55)
// and the N(0,1) noise of the two rsample() calls (src/models/VAE.py:390-392) with Philox.
// Parity mode feeds host-generated mask_p / eps instead (the CPU generators cannot be
// reproduced by Philox), see DESIGN.md.
#include <curand_kernel.h>

#include "pcvae_internal.cuh"

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
