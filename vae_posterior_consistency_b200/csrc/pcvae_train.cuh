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
    // tensor-core path only: feature-major ("transposed") activation / pre-activation-gradient buffers [feature][R2P]
    // (R2P = nbr*B rounded up to 32, padding columns zero) for the weight-gradient GEMMs, and the ReLU masks
    float* ws_zT;     // [16][R2P]   (z, 1, 0...)
    float* ws_h4T;    // [56][R2P]   (relu h4, 1, 0...)
    float* ws_h5T;    // [104][R2P]  (relu h5, 1, 0...)
    float* ws_dp6T;   // [104][R2P]  dL/d(pre-sigmoid)
    float* ws_dp5T;   // [104][R2P]
    float* ws_dp4T;   // [56][R2P]
    unsigned* ws_relu; // [nbr*B][8]: bit masks of h5 > 0 (4 column groups x 28) and h4 > 0 (4 x 16)
    long R2P;
};

// tensor-core encoder scratch (pcvae_enc_tc.cu): feature-major [feature][R2P] like the decoder's
struct EncTcWs {
    float* inT;       // [104][R2P]  (x*mask, 1 at feature D, 0...)
    float* h1T;       // [104][R2P]  (relu h1, 1, 0...)
    float* h2T;       // [56][R2P]   (relu h2, 1, 0...)
    float* dp1T;      // [104][R2P]  dL/d(pre1)
    float* dp2T;      // [56][R2P]
    float* dp3T;      // [24][R2P]   (d_mean | d_logvar)
    unsigned* relu;   // [nbr*B][8]: bit masks of h1 > 0 (4 column groups x 28) and h2 > 0 (4 x 16)
    long R2P;
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
};

// pcvae_dec_tc.cu
constexpr int TCW_Z = 16, TCW_H4 = 56, TCW_H5 = 104;            // pitches of the buffers above
constexpr int TCW_FEATS = TCW_Z + 2 * TCW_H4 + 3 * TCW_H5;      // feature rows of the scratch
inline long tcw_r2p(long rows, int nbr) { return (rows * nbr + 31) / 32 * 32; }
inline long tcw_floats(long rows, int nbr) { return TCW_FEATS * tcw_r2p(rows, nbr) + 8 * rows * nbr; }
bool dec_tc_supported(const Layout& L);
int dec_tc_launch(const DecArgs& a, int grid, cudaStream_t st);

// pcvae_enc_tc.cu
constexpr int ETW_IN = 104, ETW_H1 = 104, ETW_H2 = 56, ETW_DP3 = 24;          // pitches (feature rows) of EncTcWs
constexpr int ETW_FEATS = ETW_IN + 2 * ETW_H1 + 2 * ETW_H2 + ETW_DP3;
inline long etw_floats(long rows, int nbr) { return ETW_FEATS * tcw_r2p(rows, nbr) + 8 * rows * nbr; }
bool enc_tc_supported(const Layout& L);
void enc_tc_carve(float* w, long rows, int nbr, EncTcWs* tw);
int enc_fwd_tc_launch(const EncFwdArgs& a, int grid, cudaStream_t st);
int enc_bwd_tc_launch(const EncBwdArgs& a, int grid, cudaStream_t st);

// pcvae_wgrad_tc.cu: dWaug[m][n] = sum_rows AT[m][row] * BT[n][row] for up to three layers in one launch
struct WgradJob {
    const float* AT; int Ma;              // [>= Ma][R2P]: pre-activation gradients, feature-major
    const float* BT; int Kin;             // [>= Kin + 1][R2P]: layer input | 1, feature-major
    int Nb;                               // MMA N: round16(Kin + 1)
    int W_off, b_off;                     // where dW [Ma][Kin] and db [Ma] go in the flat layout
};
int wgrad_tc_launch(const WgradJob* jobs, int njobs, long R2P, float* gp, long P, int grid, cudaStream_t st);

}  // namespace pcvae
