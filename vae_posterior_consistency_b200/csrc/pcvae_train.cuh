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

// pcvae_dec_tc.cu
constexpr int TCW_Z = 16, TCW_H4 = 56, TCW_H5 = 104;            // pitches of the buffers above
constexpr int TCW_FEATS = TCW_Z + 2 * TCW_H4 + 3 * TCW_H5;      // feature rows of the scratch
inline long tcw_r2p(long rows, int nbr) { return (rows * nbr + 31) / 32 * 32; }
inline long tcw_floats(long rows, int nbr) { return TCW_FEATS * tcw_r2p(rows, nbr) + 8 * rows * nbr; }
bool dec_tc_supported(const Layout& L);
int dec_tc_launch(const DecArgs& a, int grid, cudaStream_t st);

}  // namespace pcvae
