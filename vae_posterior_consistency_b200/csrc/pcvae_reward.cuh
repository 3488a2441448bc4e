// Shared declarations of the reward kernels (pcvae_reward.cu: FP32 FFMA path, all families;
// pcvae_reward_tc.cu: tcgen05 3xTF32 tensor-core path of the main kernel, MLP family).
#pragma once
#include "pcvae_internal.cuh"

namespace pcvae {

constexpr int BASEW = 40;      // per base: mean[10], logvar[10], 1/std[10], 1/var[10]
constexpr int CANDP = 128;     // candidate list pitch (bytes)
constexpr int NPAIR = TM_REWARD / 2;

struct RewardArgs {
    Layout L;
    int N, M, mask_kind;
    const float* theta;
    const float* x;
    const void* mask;
    const float* im;
    long im_ss;
    float* R;
    float* base_in;   // [N][INW]   h0 pre-activation (MLP, INW=100) or agg0 (PNP, INW=K4)
    float* base0;     // [N][40]
    float* baseT;     // [N][M][40]
    int* cnt;         // [N]
    int* off;         // [N+1]
    uint8_t* cand;    // [N][CANDP]
    int* pairs;       // [N*(D-1)]  n*128+u
    const float* ac;  // PNP tables
    int* status;      // tensor-core status word (tc_status_ptr), set by the launcher
};


// tcgen05 version of k_reward_main for the MLP family; returns a PCVAE_* code.
int reward_main_tc_launch(const RewardArgs& a, int grid, cudaStream_t st);
// warp-specialised tcgen05 version (pcvae_reward_ws.cu): same arithmetic in the same order, phases of consecutive
// samples overlapped through mbarriers; bit-identical R
int reward_main_ws_launch(const RewardArgs& a, int grid, cudaStream_t st);

}  // namespace pcvae
