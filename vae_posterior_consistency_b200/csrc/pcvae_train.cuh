// Argument blocks shared by the FFMA training kernels (pcvae_train.cu) and their tcgen05 versions.
#pragma once
#include "pcvae_internal.cuh"

namespace pcvae {

struct DecArgs {
    Layout L;
    int mode, B, nbr, mask_kind;
    const float* theta;
    const float* z[2];
    float* xhat[2];
    const float* x;
    const void* mask[2];
    const float* mean[2];
    const float* logvar[2];
    const float* eps[2];
    float alpha, beta_w, x_logvar, loss_scale;
    float* sums_partials;
    float* d_mean[2];
    float* d_logvar[2];
    const float* d_xhat[2];
    float* d_z[2];
    float* gp;
    // tensor-core path only: activation / pre-activation-gradient buffers for the weight-gradient GEMMs, tile-blocked
    // feature-major [vt][feature][128 rows] (vt = branch * ntiles + tile; one 128-row tile of a kernel is one contiguous
    // block, rows past the batch hold zero gradients), and the ReLU masks
    float* ws_zT;     // [nvt][16][128]   (z, 1, 0...)
    float* ws_h4T;    // [nvt][56][128]   (relu h4, 1, 0...)
    float* ws_h5T;    // [nvt][104][128]  (relu h5, 1, 0...)
    float* ws_dp6T;   // [nvt][104][128]  dL/d(pre-sigmoid)
    float* ws_dp5T;   // [nvt][104][128]
    float* ws_dp4T;   // [nvt][56][128]
    unsigned* ws_relu; // [nvt*128][8]: bit masks of h5 > 0 (4 column groups x 28) and h4 > 0 (4 x 16)
    long nvt;
    int* status;      // tensor-core status word (tc_status_ptr), set by the launchers
    const float* wimg_fwd;   // tensor-core path, optional: prebuilt weight images of k_dec_fwd_tc / k_dec_bwd_tc
    const float* wimg_bwd;   // (pcvae_build_weight_images); nullptr: every CTA builds them from theta
};

// tensor-core encoder scratch (pcvae_enc_tc.cu): tile-blocked feature-major [vt][feature][128 rows] like the decoder's
struct EncTcWs {
    float* inT;       // [nvt][104][128]  (x*mask, 1 at feature D, 0...)
    float* h1T;       // [nvt][104][128]  (relu h1, 1, 0...)
    float* h2T;       // [nvt][56][128]   (relu h2, 1, 0...)
    float* dp1T;      // [nvt][104][128]  dL/d(pre1)
    float* dp2T;      // [nvt][56][128]
    float* dp3T;      // [nvt][24][128]   (d_mean | d_logvar)
    unsigned* relu;   // [nvt*128][8]: bit masks of h1 > 0 (4 column groups x 28) and h2 > 0 (4 x 16)
    long nvt;
    int* status;      // tensor-core status word (tc_status_ptr), set by the launchers
};

struct EncFwdArgs {
    Layout L;
    int B, nbr, mask_kind;
    const float* theta;
    const float* x;
    const void* mask[2];
    const float* eps[2];
    float* mean[2];
    float* logvar[2];
    float* z[2];
    float* act_ws;
    const float* ac;
    EncTcWs tw;       // tensor-core path only
    long x_bs;        // tensor-core path only: floats between the x of branch 0 and of branch 1 (0: both branches read
                      // the same x; PNP family: the pooled embeddings differ per branch, pcvae_pnp_tc.cu)
    const float* wimg;   // tensor-core path, optional: prebuilt weight images of k_enc_fwd_tc
};

struct EncBwdArgs {
    Layout L;
    int B, nbr, mask_kind;
    const float* theta;
    const float* x;
    const void* mask[2];
    const float* act_ws;
    const float* d_mean[2];
    const float* d_logvar[2];
    const float* ac;
    float* gp;   // [grid][P]
    const float* d_z[2];
    const float* eps[2];
    const float* logvar[2];
    EncTcWs tw;       // tensor-core path only
    const float* wimg;   // tensor-core path, optional: prebuilt weight images of k_enc_bwd_tc
};

// Shapes of the weight images of the row-tile kernels ([K / 4 chunks][N rows][4], K-major no-swizzle core-matrix order)
namespace tc {
// pcvae_enc_tc.cu, forward images
constexpr int E1_N = 112;                  // 100 outputs + the constant-1 generator; K = round8(D + 1)
constexpr int E2_C = 26, E2_N = 64;        // K = 104 (h1|1), 50 outputs + the constant-1 generator
constexpr int E3_C = 14, E3_N = 32;        // K = 56 (h2|1), mean | logvar
// pcvae_enc_tc.cu, data-gradient images (transposed weights)
constexpr int Y3_C = 6, Y3_N = 64;         // K = 24 (n over 2L), 50 inputs k
constexpr int Y2_C = 14, Y2_N = 112;       // K = 56 (n over 50), 100 inputs k
// pcvae_dec_tc.cu, forward images
constexpr int F4_C = 4, F4_N = 64;        // K = 16 (z|1), 50 outputs + the constant-1 generator
constexpr int F5_C = 14, F5_N = 112;      // K = 56 (h4|1), 100 outputs + the constant-1 generator
constexpr int F6_C = 26;                  // K = 104 (h5|1), round16(D) outputs
// pcvae_dec_tc.cu, data-gradient images (transposed weights)
constexpr int X6_C = 26, X6_N = 112;      // K = 104 (d), 100 inputs k
constexpr int X5_C = 26, X5_N = 64;       // K = 104 (n), 50 inputs k
constexpr int X4_C = 14, X4_N = 16;       // K = 56 (n), 10 inputs k
__host__ __device__ inline int enc_fwd_image_floats(int D) { return 2 * ((((D + 8) & ~7) / 4) * E1_N * 4 + E2_C * E2_N * 4 + E3_C * E3_N * 4); }
__host__ __device__ inline int enc_bwd_image_floats() { return 2 * (Y3_C * Y3_N * 4 + Y2_C * Y2_N * 4); }
__host__ __device__ inline int dec_fwd_image_floats(int D) { return 2 * (F4_C * F4_N * 4 + F5_C * F5_N * 4 + F6_C * ((D + 15) & ~15) * 4); }
__host__ __device__ inline int dec_bwd_image_floats() { return 2 * (X6_C * X6_N * 4 + X5_C * X5_N * 4 + X4_C * X4_N * 4); }
}  // namespace tc

// Prebuilt weight images (pcvae_weight_images_floats / pcvae_build_weight_images): one buffer, four blocks in this order,
// each exactly the shared-memory weight region of its kernel.  `Lenc` is the layout the encoder kernels run with (the
// MLP tail of the PNP family: pcvae_pnp_tc.cu).
struct WeightImages {
    long enc_fwd, enc_bwd, dec_fwd, dec_bwd, total;      // float offsets
};
bool weight_images_plan(const Layout& L, WeightImages* w, Layout* Lenc, bool* enc, bool* dec);
// pcvae_tc_images.cu: all blocks in ONE launch, a thread per 16-byte chunk of an image
int build_weight_images_launch(const Layout& L, const Layout& Lenc, bool enc, bool dec, const WeightImages& w, const float* theta,
                               float* images, cudaStream_t st);

// pcvae_dec_tc.cu
constexpr int TCW_Z = 16, TCW_H4 = 56, TCW_H5 = 104;            // pitches of the buffers above
constexpr int TCW_FEATS = TCW_Z + 2 * TCW_H4 + 3 * TCW_H5;      // feature rows of the scratch
inline long tc_nvt(long rows, int nbr) { return nbr * ((rows + 127) / 128); }      // 128-row tiles over all branches
inline long tcw_floats(long rows, int nbr) { return (TCW_FEATS + 8) * 128 * tc_nvt(rows, nbr); }
bool dec_tc_supported(const Layout& L);
int dec_tc_launch(const DecArgs& a, int grid, cudaStream_t st);

// pcvae_enc_tc.cu
constexpr int ETW_IN = 104, ETW_H1 = 104, ETW_H2 = 56, ETW_DP3 = 24;          // pitches (feature rows) of EncTcWs
constexpr int ETW_FEATS = ETW_IN + 2 * ETW_H1 + 2 * ETW_H2 + ETW_DP3;
inline long etw_floats(long rows, int nbr) { return (ETW_FEATS + 8) * 128 * tc_nvt(rows, nbr); }
bool enc_tc_supported(const Layout& L);
void enc_tc_carve(float* w, long rows, int nbr, EncTcWs* tw);
int enc_fwd_tc_launch(const EncFwdArgs& a, int grid, cudaStream_t st);
int enc_bwd_tc_launch(const EncBwdArgs& a, int grid, cudaStream_t st);

// pcvae_pnp_tc.cu: PNP (EDDI) set encoder = masked pooled embedding on the CUDA cores + the MLP tail on the tensor-core
// encoder kernels (the tail sees obs_dim = emb_dim, x = pooled embedding per branch, mask = ones)
bool pnp_tc_supported(const Layout& L);
Layout pnp_tail_layout(const Layout& L);                                // the MLP tail as the tensor-core encoder kernels see it
long pnp_tc_extra_floats(const Layout& L, long rows, int nbr);       // workspace floats beyond etw_floats()
int pnp_enc_fwd_tc_launch(const EncFwdArgs& a, float* extra, int grid, cudaStream_t st);
int pnp_enc_bwd_tc_launch(const EncBwdArgs& a, float* extra, int grid, cudaStream_t st);

// pcvae_wgrad_tc.cu: dWaug[m][n] = sum_rows AT[m][row] * BT[n][row] for up to three layers in one launch
struct WgradJob {
    const float* AT; int Fa, Ma;          // [nvt][Fa][128]: pre-activation gradients (Ma <= Fa real features)
    const float* BT; int Fb, Kin;         // [nvt][Fb][128]: layer input | 1 (Kin + 1 <= Fb real features)
    int Nb;                               // MMA N: round16(Kin + 1)
    int W_off, b_off;                     // where dW [Ma][Kin] and db [Ma] go in the flat layout
};
int wgrad_tc_launch(const WgradJob* jobs, int njobs, long nvt, float* gp, long P, int grid, cudaStream_t st);

}  // namespace pcvae
