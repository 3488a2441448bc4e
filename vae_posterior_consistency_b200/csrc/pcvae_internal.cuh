// Internal declarations shared by the translation units of libpcvae_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include "pcvae_b200.h"
#include "pcvae_tile.cuh"

namespace pcvae {

// fixed layer widths of the in-scope families (src/models/VAE.py:366-376, 692-701)
constexpr int H1 = 100, H2 = 50, H2P = 52;       // encoder hidden widths (H2 padded to 4)
constexpr int LAT = 10, LAT2 = 20, LATP = 12;     // latent, 2*latent, latent padded
constexpr int G1 = 50, G1P = 52, G2 = 100;        // decoder hidden widths
constexpr int TM_TRAIN = 64;                       // rows per tile in the training kernels
constexpr int TM_REWARD = 128;                     // tail evaluations per tile in the reward kernel
constexpr int MAX_SMEM = 232448;                   // 227 KB dynamic shared memory per CTA (sm_100)
constexpr int MAX_D = 128, MAX_K = 32;

// offsets (floats) of every parameter tensor in the flat vector, see pcvae_b200.h
struct Layout {
    int fam, D, K;                     // fam: PCVAE_FAMILY_MLP or _PNP (MLP_MASK is MLP with aug = 1)
    int aug;                           // 1: first encoder layer reads [x*mask, mask] (2D inputs), VAE.py:547
    int E, bE, We, be;                 // PNP only
    int W1, b1, W2, b2, W3, b3;        // encoder MLP (first layer input = D or K)
    int W4, b4, W5, b5, W6, b6;        // decoder
    int total;
};

__host__ __device__ inline int enc_act_feats(int fam, int K) {
    return (fam == PCVAE_FAMILY_PNP ? round4(K) : 0) + H1 + H2P;
}

char* err_buf();
int fail(int code, const char* fmt, ...);
bool make_layout(const pcvae_model* m, Layout* L);
// PCVAE_OK and the SM count when the current device is a compute-capability-10.x part
int device_ok(int* n_sm);
// device pointer of the current device's tensor-core status word (pinned mapped host memory; nullptr if it could not be
// set up): kernels store a non-zero code there when a bounded mbarrier wait runs out, device_ok() reports it
int* tc_status_ptr();
// records the next armed profiling event (pcvae_profile_events) on `st`, if any
void prof_mark(cudaStream_t st);
// PCVAE_PDL=0 turns programmatic dependent launch off (every kernel then waits for its predecessor's completion)
bool pdl_enabled();
// kernel<<<grid, block, smem, st>>>(args...); with `dependent` the launch carries the programmatic-stream-serialization
// attribute: the kernel MUST call tc::pdl_wait() before it touches anything an earlier kernel of the stream wrote
template <typename... KArgs, typename... Args>
inline cudaError_t launch_tc(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool dependent, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (dependent && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
// fills the PNP collapsed tables A,C (2*D*round4(K) floats) from theta
void pnp_tables_launch(const Layout& L, const float* theta, float* ac, cudaStream_t st);

}  // namespace pcvae
