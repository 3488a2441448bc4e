/*
 * pcvae_b200.h -- C ABI of the B200-native partial-VAE posterior-consistency hot path.
 *
 * One shared library (libpcvae_b200.so, built by nvcc for sm_100a only).  The
 * reference (stschia/VAE-posterior-consistency) has no FFI of its own: its seam is
 * the Python API of src/models/VAE.py plus the functions of
 * src/experiment_main/{train,evaluate}.py that call it.  Every entry
 * point below names the reference code it replaces; INTEGRATION.md shows the
 * ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C: device pointers, sizes, scalars; no torch types.
 *  - the caller owns every buffer (inputs, outputs, workspaces); the library keeps
 *    no device memory.
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *    internal synchronisation, no default-stream use; reentrant.
 *  - return 0 on success, a PCVAE_E* code otherwise; the message of the last
 *    failure on the calling thread is pcvae_last_error().  Never throws/exits.
 *  - there is NO CPU fallback and no multi-backend dispatch: a device that is not
 *    compute capability 10.x yields PCVAE_EDEVICE.
 *  - fp32 storage, fp32 accumulation and fp32-ACCURATE products everywhere.  The tensor-core kernels (default for the
 *    zero-impute MLP family, the MLP tail of the PNP family and the reward of the MLP family) issue every dense
 *    product as three tcgen05.mma.kind::tf32 instructions on hi / lo operand splits (a*b ~= a_hi*b_hi + a_hi*b_lo +
 *    a_lo*b_hi, ~2^-21 relative per product); the FFMA kernels (other shapes and families, and the cross-check of the
 *    tensor-core ones: pcvae_set_train_tensor_cores / pcvae_set_reward_tensor_cores) use plain fp32 FMAs.  No result
 *    is ever computed at TF32 / BF16 precision.  Masks are 0/1 valued, either uint8 (torch.bool) or float32.
 *
 * Flat parameter vector `theta` (and the gradient / Adam vectors of the same
 * layout): the reference's trainable parameters in state_dict order
 * (SURVEY.md A.1), each in its nn.Linear [out][in] row-major layout:
 *   family MLP (Reg_VAE, vanilla_VAE; src/models/VAE.py:366-376):
 *     seq_encoder.0.{weight[100,D],bias[100]} .2.{[50,100],[50]} .4.{[2L,50],[2L]}
 *     seq_decoder.0.{[50,L],[50]} .2.{[100,50],[100]} .4.{[D,100],[D]}
 *   family MLP_MASK (Reg_VAE_mask, vanilla_VAE_mask; src/models/VAE.py:526-537, 1011-1022): as MLP, but the first
 *     encoder layer reads [x*mask, mask] (VAE.py:547): seq_encoder.0.weight[100,2D]
 *   family PNP (Reg_EDDI, vanilla_EDDI; src/models/VAE.py:687-709):
 *     type_pars1[D,K], type_bias1[D,1], pnp_encoder1.0.{[K,K+2],[K]},
 *     pnp_encoder2.0.{[100,K],[100]} .2.{[50,100],[50]} .4.{[2L,50],[2L]}, seq_decoder.* as above
 * L (latent_dim) must be 10 (the reference hard-codes 10 at VAE.py:724).
 */
#ifndef PCVAE_B200_H
#define PCVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define PCVAE_ABI_VERSION 1

enum { PCVAE_OK = 0, PCVAE_EINVAL = 1, PCVAE_EDEVICE = 2, PCVAE_ECUDA = 3, PCVAE_EWORKSPACE = 4 };
enum { PCVAE_FAMILY_MLP = 0, PCVAE_FAMILY_PNP = 1, PCVAE_FAMILY_MLP_MASK = 2 };
enum { PCVAE_MASK_U8 = 0, PCVAE_MASK_F32 = 1 };

/* number of loss partial sums produced by pcvae_dec_loss / pcvae_loss_terms */
#define PCVAE_NSUMS 8
/* indices into the sums vector (double[PCVAE_NSUMS]) */
enum {
    PCVAE_S_RE_Q = 0,   /* NLL(x, xhat_q; mask)            VAE.py:422-423 */
    PCVAE_S_RE_P = 1,   /* NLL(x, xhat_p; mask_p)          VAE.py:424-426 */
    PCVAE_S_KL_Q = 2,   /* KL(q || N(0,1))                 VAE.py:427,476-478 */
    PCVAE_S_KL_P = 3,   /* KL(p || N(0,1))                 VAE.py:428 */
    PCVAE_S_KL_REG = 4, /* KL(q || p)                      VAE.py:442,469-474 */
    PCVAE_S_RE_D = 5,   /* NLL(x, xhat_q; mask & ~mask_p)  VAE.py:444-446 */
    PCVAE_S_RE_IMP = 6, /* NLL(x, xhat_q; ~mask)           VAE.py:413-414 */
    PCVAE_S_SSE_UNOBS = 7 /* sum ((xhat_q - x) * ~mask)^2  evaluate.py:232-234 */
};

typedef struct {
    int family;      /* PCVAE_FAMILY_* */
    int obs_dim;     /* D  (1..128) */
    int emb_dim;     /* K  (PNP only, 1..32) */
    int latent_dim;  /* L  (must be 10) */
} pcvae_model;

int pcvae_abi_version(void);
const char* pcvae_last_error(void);

/* number of floats in theta for a model; -1 on invalid model */
long pcvae_param_count(const pcvae_model* m);
/* offsets (in floats) of each parameter tensor in theta, state_dict order; returns the
 * number of tensors written (12 MLP, 16 PNP) or -1. `offsets` needs room for 17 longs
 * (the last entry is the total). */
int pcvae_param_offsets(const pcvae_model* m, long* offsets);
/* offset of the first decoder parameter (seq_decoder.0.weight) in theta */
long pcvae_decoder_offset(const pcvae_model* m);

/* ------------------------------------------------------------------------
 * Encoder forward.  Replaces Reg_VAE.encoder / vanilla_VAE.encoder
 * (src/models/VAE.py:387-395, 1155-1163) and Reg_EDDI.encoder / vanilla_EDDI.encoder
 * (VAE.py:719-741, 903-925; PNP embedding in the collapsed form of SURVEY.md A.3).
 * Up to two "branches" (the q pass on `mask[0]`, the p pass on `mask[1]`,
 * VAE.py:503-506) share one launch and one read of x.
 * --------------------------------------------------------------------- */
typedef struct {
    pcvae_model model;
    int rows;               /* B */
    int n_branch;           /* 1 or 2 */
    int mask_kind;          /* PCVAE_MASK_* */
    const float* theta;
    const float* x;         /* [B][D] */
    const void* mask[2];    /* [B][D] per branch */
    const float* eps[2];    /* [B][L] standard-normal draws, or NULL: z = mean (sample=False) */
    float* mean[2];         /* out [B][L] */
    float* logvar[2];       /* out [B][L] */
    float* z[2];            /* out [B][L], may be NULL */
    float* act_ws;          /* optional: saved hidden activations for pcvae_enc_bwd,
                               pcvae_enc_act_ws_floats() floats; NULL = do not save */
    float* pnp_ac;          /* PNP only: workspace 2*D*K floats for the collapsed A,C tables */
    /* optional: scratch of pcvae_enc_tc_workspace_floats() floats.  When given (and tensor cores are enabled and
     * the shape is supported) the dense products run on tcgen05 in 3xTF32 -- same results to fp32 rounding --
     * and everything pcvae_enc_bwd needs is saved there instead of act_ws (which may then be NULL). */
    float* tc_workspace;
    long tc_workspace_floats;
    /* optional, tensor-core path: the buffer pcvae_build_weight_images() filled from THIS theta (see below); NULL = every
     * CTA builds the operand images of the weights from theta at launch */
    const float* weight_images;
} pcvae_enc_fwd_params;

size_t pcvae_enc_act_ws_floats(const pcvae_model* m, int rows, int n_branch);
/* floats of the encoder tc_workspace for `rows` x `n_branch`; 0 when this (family, obs_dim) has no tensor-core
 * encoder (PNP family, obs_dim not a multiple of 4 or > 100, or pcvae_set_train_tensor_cores(0)) */
long pcvae_enc_tc_workspace_floats(const pcvae_model* m, int rows, int n_branch);
int pcvae_enc_fwd(const pcvae_enc_fwd_params* p, void* stream);

/* ------------------------------------------------------------------------
 * Encoder backward (autograd of the encoders above, train.py:115): given
 * dL/dmean, dL/dlogvar per branch and the activations saved by pcvae_enc_fwd,
 * accumulates per-CTA partial parameter gradients into `grad_partials`
 * ([pcvae_grid_ctas()][param_count] floats, encoder slice only) -- reduce with
 * pcvae_reduce_grads.
 * --------------------------------------------------------------------- */
typedef struct {
    pcvae_model model;
    int rows, n_branch, mask_kind;
    const float* theta;
    const float* x;
    const void* mask[2];
    const float* act_ws;        /* from pcvae_enc_fwd */
    const float* d_mean[2];     /* [B][L] */
    const float* d_logvar[2];   /* [B][L] */
    const float* pnp_ac;        /* PNP: A,C tables from pcvae_enc_fwd */
    float* grad_partials;       /* [grid][P] */
    /* optional reparameterisation backward (z = mean + eps*exp(logvar/2), VAE.py:390-392): when
       d_z[b] is non-NULL, d_mean += d_z and d_logvar += d_z * 0.5 * exp(logvar/2) * eps are
       folded in by the kernel (eps[b] NULL means sample=False, z = mean). */
    const float* d_z[2];        /* [B][L] or NULL */
    const float* eps[2];        /* [B][L] or NULL */
    const float* logvar[2];     /* [B][L], required when d_z[b] and eps[b] are given */
    /* the tc_workspace pcvae_enc_fwd filled for these rows (then act_ws is not read), or NULL */
    float* tc_workspace;
    long tc_workspace_floats;
    const float* weight_images; /* optional, as in pcvae_enc_fwd_params */
} pcvae_enc_bwd_params;
int pcvae_enc_bwd(const pcvae_enc_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------
 * Decoder (+ loss, + backward).  One kernel, four modes:
 *  PCVAE_DEC_FWD    z -> xhat                      (decoder(), VAE.py:397-401)
 *  PCVAE_DEC_BWD    z, d_xhat -> d_z, dW partials  (its autograd)
 *  PCVAE_DEC_TRAIN  reparameterise, decode both branches, masked Gaussian NLL + KL +
 *                   posterior-consistency loss, and the backward through the decoder
 *                   down to d_mean/d_logvar (forward()+loss()+backward() of
 *                   Reg_VAE/Reg_EDDI/vanilla_*, VAE.py:403-467, 496-507; train.py:87-116)
 *  PCVAE_DEC_EVAL   loss sums only (stage='evaluate', VAE.py:410-420) and RMSE sums
 * Loss weights: L = (1-alpha)(RE_q + beta_w KL_q) + alpha (KL_reg + RE_p + beta_w KL_p + RE_d),
 * gradients are of L * loss_scale (loss_scale = 1/B, VAE.py:452).  vanilla_* = one branch,
 * alpha = 0.
 * --------------------------------------------------------------------- */
enum { PCVAE_DEC_FWD = 0, PCVAE_DEC_BWD = 1, PCVAE_DEC_TRAIN = 2, PCVAE_DEC_EVAL = 3 };
typedef struct {
    pcvae_model model;
    int mode;
    int rows, n_branch, mask_kind;
    const float* theta;
    const float* z[2];          /* [B][L] latent samples per branch (q, p) */
    float* xhat[2];             /* FWD/EVAL/TRAIN: optional out [B][D] */
    /* TRAIN / EVAL */
    const float* x;             /* [B][D] */
    const void* mask[2];        /* mask, mask_p */
    const float* mean[2];       /* [B][L] */
    const float* logvar[2];
    const float* eps[2];        /* TRAIN: the draws used for z (needed for d_logvar) */
    float alpha, beta_w, x_logvar, loss_scale;
    float* sums_partials;       /* [grid][PCVAE_NSUMS] floats */
    float* d_mean[2];           /* TRAIN out [B][L] */
    float* d_logvar[2];
    /* BWD */
    const float* d_xhat[2];     /* [B][D] */
    float* d_z[2];              /* BWD out [B][L] */
    float* grad_partials;       /* TRAIN/BWD: [grid][P], decoder slice only */
    /* TRAIN, optional: scratch of pcvae_dec_tc_workspace_floats() floats.  When given (and tensor cores are
     * enabled and the shape is supported) the dense products run on tcgen05 in 3xTF32 -- same results to fp32
     * rounding -- and the pre-activation gradients / activations for the weight-gradient GEMMs go through it. */
    float* tc_workspace;
    long tc_workspace_floats;
    const float* weight_images; /* TRAIN, optional, as in pcvae_enc_fwd_params */
} pcvae_dec_params;
int pcvae_dec(const pcvae_dec_params* p, void* stream);
/* floats of tc_workspace for `rows` x `n_branch`; 0 when this (family, obs_dim) has no tensor-core decoder */
long pcvae_dec_tc_workspace_floats(const pcvae_model* m, int rows, int n_branch);
/* process-wide switch for the tcgen05 training kernels (default 1); returns the previous value */
int pcvae_set_train_tensor_cores(int enable);
/* Prebuilt weight images for the tensor-core training kernels.  Each of the four row-tile kernels of a step keeps the hi / lo
 * tf32 operand images of three weight matrices in shared memory; built from theta by every CTA at launch that is ~5 us
 * of a 60 us kernel (dependent L2 round trips, splits, scattered stores).  pcvae_build_weight_images() builds all of them
 * ONCE per theta into `images` (pcvae_weight_images_floats() floats; two launches of two CTAs -- run it on a side stream
 * while the batch is prepared); the kernels then fetch their block with bulk async copies.  Pass the buffer in the
 * weight_images field of the encoder / decoder calls that run with the same theta.  Results are bit-identical with and
 * without.  pcvae_weight_images_floats() is 0 when the model has no tensor-core training kernels. */
long pcvae_weight_images_floats(const pcvae_model* m);
int pcvae_build_weight_images(const pcvae_model* m, const float* theta, float* images, void* stream);
/* Per-kernel timing hooks for benchmarks: arm `n` caller-created CUDA events (cudaEvent_t[n], timing enabled) on the
 * calling thread.  Each tensor-core launcher (pcvae_enc_fwd, pcvae_dec TRAIN, pcvae_enc_bwd) then records the next
 * event on its stream before its first kernel and after every kernel, until the events run out.  Pass NULL to
 * disarm.  Returns how many events of the previous arming were recorded.  No effect on results. */
int pcvae_profile_events(void** events, int n);

/* ------------------------------------------------------------------------
 * Stand-alone loss terms for the module API (`model.loss(...)` called on tensors the
 * caller already holds; VAE.py:403-467, 749-817, 933-964, 1171-1208).  Computes the
 * PCVAE_NSUMS sums and, when the d_* pointers are non-NULL, the gradients of
 * L*loss_scale w.r.t. xhat_q, xhat_p, mean/logvar of q and p.
 * --------------------------------------------------------------------- */
typedef struct {
    int rows, obs_dim, latent_dim, n_branch, mask_kind;
    const float* x;
    const void* mask[2];
    const float* xhat[2];
    const float* mean[2];
    const float* logvar[2];
    float alpha, beta_w, x_logvar, loss_scale;
    float* sums_partials;       /* [grid][PCVAE_NSUMS] */
    float* d_xhat[2];
    float* d_mean[2];
    float* d_logvar[2];
} pcvae_loss_params;
int pcvae_loss_terms(const pcvae_loss_params* p, void* stream);

/* grid size (number of CTAs = number of partial rows) every kernel above uses on the
 * current device: one persistent CTA per SM. */
int pcvae_grid_ctas(void);

/* sums[PCVAE_NSUMS] (double, device) = sum over CTAs of sums_partials, plus the
 * B*D*0.5*log(2*pi) constant every NLL carries. */
int pcvae_reduce_sums(const float* sums_partials, int grid, int rows, int obs_dim,
                      double* sums, void* stream);
/* grad[i] (+)= sum_c grad_partials[c][i] for i in [begin, end) ; deterministic order */
int pcvae_reduce_grads(const float* grad_partials, int grid, long param_count, long begin,
                       long end, float* grad, int accumulate, void* stream);

/* torch.optim.Adam (train.py:21; lr 1e-3, betas .9/.999, eps 1e-8, no weight decay, no
 * amsgrad) on the flat vectors; `step` is the 1-based step count. */
int pcvae_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq,
                    long n, int step, float lr, float beta1, float beta2, float eps,
                    void* stream);
/* Single-GPU training: pcvae_reduce_grads over all parameters, pcvae_adam_step and (when sums_partials / sums are
 * non-NULL) pcvae_reduce_sums in ONE launch.  `grad` receives the reduced gradient.  The partials of a parameter
 * are summed in a fixed order (eight interleaved chains, then combined), so results are deterministic. */
int pcvae_reduce_adam(const float* grad_partials, int grid, long param_count, float* grad, float* theta,
                      float* exp_avg, float* exp_avg_sq, int step, float lr, float beta1, float beta2, float eps,
                      const float* sums_partials, int rows, int obs_dim, double* sums, void* stream);

/* ------------------------------------------------------------------------
 * Data-parallel training (optional): the tail of a step in ONE launch on every rank --
 * deterministic reduce of the rank's partials, exchange of the reduced gradient with the other GPUs of the node over
 * NVLink peer memory, sum over the ranks in rank order (bit-identical on all ranks) and Adam.  It stands where
 * `train_loss.backward(); optimizer.step()` stand in train.py:114-116 once the batch is split over GPUs, and replaces
 * pcvae_reduce_grads + an NCCL all-reduce + pcvae_adam_step.  Each rank (one process per GPU) allocates an exchange
 * buffer, publishes its 64-byte CUDA IPC handle (any host channel, e.g. torch.distributed.all_gather_object), opens the
 * handles of the others, and passes all `world` base pointers (its own at index `rank`) with every call.  `seq` is the
 * 1-based number of the call on this buffer and must advance by one per call on every rank in lockstep.  A rank that
 * does not arrive within ~4 s makes the waiting ranks set *status = 1 and SKIP their Adam update (no hang, no update from
 * stale slots); check the status word on the host (dist.PeerExchange.check, called by train() at every epoch end).
 * --------------------------------------------------------------------- */
#define PCVAE_DP_MAX_WORLD 16
size_t pcvae_dp_exchange_bytes(long param_count, int world);
int pcvae_dp_exchange_alloc(long param_count, int world, void** buffer, unsigned char* ipc_handle_64);
int pcvae_dp_exchange_open(const unsigned char* ipc_handle_64, void** peer_buffer);
int pcvae_dp_exchange_close(void* peer_buffer);
int pcvae_dp_exchange_free(void* buffer);
typedef struct {
    const float* grad_partials; int grid; long param_count;     /* as pcvae_reduce_adam */
    float* grad; float* theta; float* exp_avg; float* exp_avg_sq;
    int step; float lr, beta1, beta2, eps;
    const float* sums_partials; int rows, obs_dim; double* sums; /* this rank's loss sums (optional pair) */
    int world, rank; unsigned seq;
    void* peer_buffers[PCVAE_DP_MAX_WORLD];
    int* status;                                                  /* device int, zero-initialised by the caller */
    unsigned long long* step_state;                               /* optional (CUDA-graph replay): seq and the Adam step are
                                                                     step_state[0] + 1 and the launch advances it, as in
                                                                     pcvae_reduce_adam_dev; `step` and `seq` are then ignored */
} pcvae_dp_params;
int pcvae_dp_reduce_adam(const pcvae_dp_params* p, void* stream);
/* The same kernel body with `world` ranks EMULATED on one GPU (tests on a single-GPU box): ranks[r] describes rank r
 * (its own partials, weights, moments; peer_buffers = `world` plain cudaMalloc'ed exchange buffers on this GPU, the same
 * array in every entry).  One cooperative launch plays all ranks -- separate launches that wait for one another on one
 * GPU are not guaranteed to overlap.  Results are bit-identical to `world` real ranks (the per-chunk arithmetic does
 * not depend on the grid). */
#define PCVAE_DP_EMU_MAX_WORLD 4
int pcvae_dp_reduce_adam_emulated(const pcvae_dp_params* const* ranks, int world, void* stream);

/* ------------------------------------------------------------------------
 * Active-selection information reward for ONE acquisition step, all (row, candidate,
 * sample) triples in one launch.  Replaces the candidate loop around R_lindley_chain and
 * chaini_I / chaini_II (src/experiment_main/evaluate.py:416-425, 514-634) using the
 * incremental-encoder identity of SURVEY.md A.5.  Candidates with mask[n][u] != 0 keep
 * R = -1e4 (evaluate.py:391).  Precondition (guaranteed by active_learning_func,
 * evaluate.py:360-361): the target column D-1 is unobserved, mask[:, D-1] == 0.
 * Rows are independent: shard by rows across GPUs with no collective.
 * --------------------------------------------------------------------- */
typedef struct {
    pcvae_model model;
    int rows;               /* N */
    int samples;            /* M */
    int mask_kind;
    const float* theta;
    const float* x;         /* [N][D] */
    const void* mask;       /* [N][D] */
    const float* im;        /* [M][N_total][D] imputations (decoder means); row stride below */
    long im_sample_stride;  /* floats between consecutive samples (N_total*D) */
    float* R;               /* out [N][D-1] */
    void* workspace;        /* pcvae_reward_workspace_bytes() */
    size_t workspace_bytes;
    float* pnp_ac;          /* PNP: 2*D*K floats */
} pcvae_reward_params;
size_t pcvae_reward_workspace_bytes(const pcvae_model* m, int rows, int samples);
int pcvae_reward_chain(const pcvae_reward_params* p, void* stream);
/* Select the main reward kernel of the MLP family: 1 (default) = tcgen05 tensor cores with the fp32-accurate
 * 3xTF32 operand split, warp-specialised (csrc/pcvae_reward_ws.cu: constructor / epilogue / KL warpgroups
 * pipelined through mbarriers); 2 = the same arithmetic in lock-step phases (csrc/pcvae_reward_tc.cu, kept as
 * the cross-check: R is bit-identical to 1); 0 = FP32 FFMA (csrc/pcvae_reward.cu).  All compute the same
 * function to fp32 accuracy; the PNP family always uses the FFMA kernel.  Returns the previous value. */
int pcvae_set_reward_tensor_cores(int enable);

/* ------------------------------------------------------------------------
 * Throughput-mode batch preparation on the device.
 *  pcvae_gather_rows  : x[b] = table[idx[b]], mask[b] = mask_table[idx[b]]  -- the batch a
 *                       DataLoader(shuffle=True) yields (src/utils/loaders.py:342-352,389-397)
 *                       for the permutation `idx` (int64, from torch's RandomSampler).
 *  pcvae_draw_submask : mask_p = mask & (u < keep_prob), u ~ U[0,1) Philox4x32-10
 *                       (create_missing_uci, src/utils/utils.py:36-39; train.py:53-55).
 *  pcvae_draw_normal  : standard-normal draws for rsample() (VAE.py:390-392).
 * Philox streams are statistically, not bitwise, equivalent to the reference's NumPy /
 * torch CPU generators; parity mode passes host-drawn mask_p / eps instead.
 * --------------------------------------------------------------------- */
int pcvae_gather_rows(const float* table, const void* mask_table, const long* idx, float* x, void* mask,
                      int rows, int obs_dim, int mask_kind, void* stream);
int pcvae_draw_submask(const uint8_t* mask, uint8_t* mask_p, long n, float keep_prob,
                       unsigned long long seed, unsigned long long offset, void* stream);
int pcvae_draw_normal(float* out, long n, unsigned long long seed, unsigned long long offset, void* stream);
/* The three steps above in ONE launch (uint8 masks, obs_dim % 4 == 0, obs_dim <= 128): x, mask gathered by idx,
 * mask_p drawn, and eps[n_eps][rows][10] standard normals (n_eps = 2 for the regularised families, 1 for vanilla). */
int pcvae_prep_batch(const float* table, const uint8_t* mask_table, const long* idx, float* x, uint8_t* mask,
                     uint8_t* mask_p, float* eps, int rows, int obs_dim, int n_eps, float keep_prob,
                     unsigned long long seed, unsigned long long offset, void* stream);

/* Replayable variants for a CUDA graph of the whole training step (launch-bound small batches, train.py's default batch
 * of 64; also removes the launch gaps at large batches): the per-step scalars come from a device counter
 * step_state[0] = number of completed steps (step_state[1] is a ticket word; both zero-initialised by the caller).
 *   pcvae_prep_batch_dev : rows of batch (step % n_batches) of idx_batches[n_batches][rows], Philox offset
 *                          offset0 + 8 * step                          (pcvae_prep_batch with idx + ..., offset0 + 8 step)
 *   pcvae_reduce_adam_dev: Adam step number step + 1 (bias corrections computed on the device); the last block of the
 *                          launch then advances step_state[0]          (pcvae_reduce_adam with step + 1).  `sums` holds
 *                          2 * PCVAE_NSUMS doubles here: the step's sums, then running totals (+= every launch; the
 *                          caller zeroes them), so that an epoch's loss needs no per-step host work.  The same holds
 *                          for pcvae_dp_reduce_adam when its step_state is set. */
int pcvae_prep_batch_dev(const float* table, const uint8_t* mask_table, const long* idx_batches, long n_batches, float* x,
                         uint8_t* mask, uint8_t* mask_p, float* eps, int rows, int obs_dim, int n_eps, float keep_prob,
                         unsigned long long seed, unsigned long long offset0, const unsigned long long* step_state, void* stream);
int pcvae_reduce_adam_dev(const float* grad_partials, int grid, long param_count, float* grad, float* theta, float* exp_avg,
                          float* exp_avg_sq, unsigned long long* step_state, float lr, float beta1, float beta2, float eps,
                          const float* sums_partials, int rows, int obs_dim, double* sums, void* stream);

/* Host-streamed batches (a table kept in host memory; the e2e path of bench.py).  What DataLoader + collate hand to
 * train.py:36-58 as (data_sample, mask) arrives here in a compact form built once by the loader:
 *   mask_bits[rows][ceil(obs_dim / 32)]  bit j of word w = mask[row][32 w + j]           (always)
 *   vals[nnz], row_off[rows]             x's OBSERVED entries only, row-major; row r starts at vals[row_off[r]]
 *                                        (optional pair: entries under a zero mask bit never reach the training
 *                                        loss, VAE.py:388, 411-445; pass NULL, NULL when x was copied dense)
 * One launch writes the dense uint8 mask, x (zeros at unobserved entries; only with vals), mask_p (may be NULL) and
 * eps[n_eps][rows][10], drawing exactly what pcvae_prep_batch draws for the same (row, seed, offset).
 * obs_dim % 4 == 0, obs_dim <= 128. */
int pcvae_prep_packed(const uint32_t* mask_bits, const float* vals, const uint32_t* row_off, float* x, uint8_t* mask,
                      uint8_t* mask_p, float* eps, int rows, int obs_dim, int n_eps, float keep_prob,
                      unsigned long long seed, unsigned long long offset, void* stream);

/* ------------------------------------------------------------------------
 * Generic dense layer y = act(x W^T + b) on row tiles (weights resident in shared memory), forward
 * and backward.  Building block of the 128-wide not-MIWAE MNAR networks (REG_notMIWAE_v2 /
 * notMIWAE_myversion, src/models/VAE.py:2342-2363, 2706-2730); in_dim, out_dim <= 128.
 * `mask` (fp32, optional) fuses the zero-imputation x*mask of the first encoder layer (VAE.py:2378, 2749).
 * --------------------------------------------------------------------- */
enum { PCVAE_ACT_NONE = 0, PCVAE_ACT_RELU = 1, PCVAE_ACT_SIGMOID = 2, PCVAE_ACT_ELU = 3,
       PCVAE_ACT_HARDTANH_M10_0 = 4 /* nn.Hardtanh(-10, 0), VAE.py:2363 */ };
typedef struct {
    int rows, in_dim, out_dim, act;
    const float* x;        /* [rows][in_dim] */
    const float* mask;     /* optional [rows][in_dim] */
    const float* W;        /* [out_dim][in_dim] */
    const float* b;        /* [out_dim] */
    float* y;              /* out [rows][out_dim] */
} pcvae_dense_fwd_params;
int pcvae_dense_fwd(const pcvae_dense_fwd_params* p, void* stream);
typedef struct {
    int rows, in_dim, out_dim, act;
    const float* x;        /* layer input (before the optional mask) */
    const float* mask;
    const float* y;        /* layer output (activation derivative is taken from it) */
    const float* dy;       /* [rows][out_dim] */
    const float* W;
    float* dx;             /* optional out [rows][in_dim] */
    float* dW_partials;    /* workspace [pcvae_grid_ctas()][out_dim*in_dim]; only the first
                              min(pcvae_grid_ctas(), ceil(rows/64)) rows are written */
    float* db_partials;    /* workspace [pcvae_grid_ctas()][out_dim], same rule */
    float* dW;             /* optional out [out_dim][in_dim]: when dW and db are both given the call also sums the */
    float* db;             /* optional out [out_dim]          written partial rows into them, in CTA order         */
} pcvae_dense_bwd_params;
int pcvae_dense_bwd(const pcvae_dense_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------
 * not-MIWAE MNAR importance-sampling pieces (S samples per row).
 *  pcvae_mnar_sample_z      z[b][s][l] = mean[b][l] + exp(logvar[b][l]/2) eps[b][s][l]   (VAE.py:2382-2387)
 *  pcvae_mnar_sample_z_bwd  d_mean = sum_s d_z, d_logvar = sum_s d_z * 0.5 exp(logvar/2) eps
 *  pcvae_mnar_loss          REG_notMIWAE_v2.loss (regularised=1, VAE.py:2398-2471) or notMIWAE_myversion.loss
 *                           (regularised=0, VAE.py:2772-2823): per-(row,sample) masked Gaussian NLL with
 *                           per-entry variance, KL, Bernoulli self-masking term -softplus(W)(x~-b),
 *                           logsumexp over samples; llh_eval imputation sum_s softmax(-l_w) xm (VAE.py:2458-2461);
 *                           and the gradients of the loss w.r.t. every input.
 * out[0]=loss, out[1]=RE_q.mean(), out[2]=loss_q, out[3]=loss_p (doubles, device).
 * --------------------------------------------------------------------- */
int pcvae_mnar_sample_z(const float* mean, const float* logvar, const float* eps, float* z, int rows,
                        int samples, int latent_dim, void* stream);
int pcvae_mnar_sample_z_bwd(const float* d_z, const float* logvar, const float* eps, float* d_mean,
                            float* d_logvar, int rows, int samples, int latent_dim, void* stream);
typedef struct {
    int rows, samples, obs_dim, latent_dim, regularised;
    const float* x;            /* [B][D] */
    const float* mask;         /* [B][D] fp32 */
    const float* mask_p;       /* [B][D] fp32 (regularised only) */
    const float* xm[2];        /* [B][S][D] x_mean head, q and p branch */
    const float* xlv[2];       /* [B][S][D] x_logvar head */
    const float* mean[2];      /* [B][L] */
    const float* logvar[2];    /* [B][L] */
    const float* eps_kl;       /* regularised=0: the second N(0,1) draw [B][S][L] of the MC KL (VAE.py:2791-2798) */
    const float* W;            /* [D] self-masking slope (before softplus) */
    const float* b;            /* [D] self-masking offset */
    float alpha;
    void* workspace;           /* pcvae_mnar_loss_workspace_bytes() */
    size_t workspace_bytes;
    double* out;               /* [4] */
    float* xm_imputed;         /* optional [B][D] */
    float* d_xm[2];            /* optional gradients, NULL = forward only */
    float* d_xlv[2];
    float* d_mean[2];
    float* d_logvar[2];
    float* d_W;                /* [D] */
    float* d_b;                /* [D] */
} pcvae_mnar_loss_params;
size_t pcvae_mnar_loss_workspace_bytes(int rows, int samples, int obs_dim);
int pcvae_mnar_loss(const pcvae_mnar_loss_params* p, void* stream);

/* ------------------------------------------------------------------------
 * Importance-weighted MNAR imputation in one pass over the [rows, samples] grid: replaces
 * eval_vae_mnar's row-by-row model.forward + model.loss(llh_eval=True) calls (src/experiment_main/evaluate.py:27-49) for
 * REG_notMIWAE_v2 (regularised = 1, src/models/VAE.py:2377-2396, 2398-2461) and notMIWAE_myversion (regularised = 0,
 * VAE.py:2748-2770, 2772-2823):  x_imputed[n] = sum_s softmax_s(-l_w[n][s]) x_mean[n][s][:],  l_w = RE + KL - log p(mask | x~).
 * The [samples, obs_dim] decoder outputs of a row never reach memory (online softmax per chunk of samples, chunks merged
 * in order).  mean / logvar [rows][L] are the encoder statistics (pcvae_dense_fwd x 4).  Noise: eps (and, for
 * regularised = 0, eps_kl: the second draw of the Monte-Carlo KL, VAE.py:2791-2798) as [rows][samples][L] arrays drawn by
 * the caller, or both NULL: drawn in the kernel with Philox from (seed, offset).  obs_dim <= 64, latent_dim <= 16.
 * --------------------------------------------------------------------- */
typedef struct {
    int rows, samples, obs_dim, latent_dim, regularised;
    const float* dec0_W; const float* dec0_b;         /* seq_decoder.0  [128][L], [128] */
    const float* dec2_W; const float* dec2_b;         /* seq_decoder.2  [128][128], [128] */
    const float* xmean_W; const float* xmean_b;       /* x_mean.0       [D][128], [D] */
    const float* xlogvar_W; const float* xlogvar_b;   /* x_logvar.0     [D][128], [D] */
    const float* W; const float* b;                   /* self-masking slope (before softplus) and offset, [D] */
    const float* x; const float* mask;                /* [rows][D] fp32 */
    const float* mean; const float* logvar;           /* [rows][L] */
    const float* eps; const float* eps_kl;            /* [rows][samples][L] or NULL */
    unsigned long long seed, offset;                  /* Philox key / counter base when eps is NULL */
    void* workspace; size_t workspace_bytes;          /* pcvae_mnar_impute_workspace_bytes() */
    float* xm_imputed;                                /* out [rows][D] */
} pcvae_mnar_impute_params;
size_t pcvae_mnar_impute_workspace_bytes(int rows, int samples, int obs_dim);
int pcvae_mnar_impute(const pcvae_mnar_impute_params* p, void* stream);

/* ------------------------------------------------------------------------
 * MIWAE / Reg_MIWAE (Student-t decoder + importance-weighted bound), reference src/models/VAE.py:3011-3134, 3137-3301;
 * dispatched by src/utils/loaders.py:135-147, 234-245, trained at src/experiment_main/train.py:102-113, evaluated by
 * eval_miwae, src/experiment_main/evaluate.py:72-133.  The 128-wide ReLU layers are
 * pcvae_dense_fwd / pcvae_dense_bwd calls (PCVAE_ACT_RELU / PCVAE_ACT_NONE); these entries are the rest:
 *  pcvae_miwae_heads        encoder (VAE.py:3047-3049): raw [rows][2W] -> out0 = raw[:, :W], out1 = softplus(raw[:, W:]);
 *                           decoder (VAE.py:3061-3066): raw [rows][3W] -> out0 = sigmoid, out1 = softplus + 0.001,
 *                           out2 = softplus + 3   (nn.Softplus: beta 1, threshold 20)
 *  pcvae_miwae_heads_bwd    d_raw from the gradients of the outputs (NULL = zero), derivatives taken from raw
 *  pcvae_miwae_sample_z     z[b][s][l] = mean[b][l] + scale[b][l] eps[b][s][l] (VAE.py:3054-3056; eps NULL: z = mean)
 *  pcvae_miwae_sample_z_bwd d_mean = sum_s d_z, d_scale = sum_s d_z eps
 *  pcvae_miwae_loss         MIWAE.loss (regularised = 0, VAE.py:3068-3110) / Reg_MIWAE.loss (regularised = 1,
 *                           VAE.py:3197-3263): StudentT(df, loc, scale).log_prob(x) per (row, sample, feature), masked sums,
 *                           the reference's [B*S] -> [S, B] reshape WITHOUT a transpose of the per-(row, sample)
 *                           likelihoods (VAE.py:3078-3081; rowwise = 1 instead treats every row as its own batch, which is
 *                           what eval_miwae's per-row calls amount to), log p(z) - log q(z|x) of the loss-internal draw
 *                           z = mean + scale eps2, -mean(logsumexp over samples); Reg: loss = nb_q + alpha (KL_reg - nb_q +
 *                           nb_p - reg_like); llh_eval imputation sum_k softmax_k x_mean[., k, :]; every gradient.
 * out[0] = loss, [1] = neg_bound_q, [2] = neg_bound_p, [3] = KL_reg, [4] = reg_like, [5] = sum of the log-likelihood on the
 * unobserved entries / (rows * 5000) (third return value of MIWAE.loss with llh_eval, VAE.py:3100); doubles, device.
 * --------------------------------------------------------------------- */
enum { PCVAE_MIWAE_HEADS_ENC = 0, PCVAE_MIWAE_HEADS_DEC = 1 };
int pcvae_miwae_heads(const float* raw, long rows, int width, int mode, float* out0, float* out1, float* out2, void* stream);
int pcvae_miwae_heads_bwd(const float* raw, long rows, int width, int mode, const float* d0, const float* d1, const float* d2,
                          float* d_raw, void* stream);
int pcvae_miwae_sample_z(const float* mean, const float* scale, const float* eps, float* z, int rows, int samples,
                         int latent_dim, void* stream);
int pcvae_miwae_sample_z_bwd(const float* d_z, const float* eps, float* d_mean, float* d_scale, int rows, int samples,
                             int latent_dim, void* stream);
typedef struct {
    int rows, samples, obs_dim, latent_dim, regularised;
    int mask_kind;             /* PCVAE_MASK_U8 (torch.bool bytes; the reference needs bool for `~new_mask`) or _F32 */
    int rowwise;               /* 0: the reference's reshape (training); 1: every row its own batch (eval_miwae) */
    const float* x;            /* [B][D] */
    const void* mask;          /* [B][D] */
    const void* mask_p;        /* [B][D] (regularised only) */
    const float* xm[2];        /* [B][S][D] Student-t loc, q and p branch */
    const float* xs[2];        /* [B][S][D] scale */
    const float* df[2];        /* [B][S][D] degrees of freedom */
    const float* mean[2];      /* [B][L] */
    const float* scale[2];     /* [B][L] */
    const float* eps2[2];      /* [B][S][L] N(0,1) draws of the loss-internal z (VAE.py:3086-3088, 3218-3220, 3239-3241) */
    float alpha;
    void* workspace;           /* pcvae_miwae_loss_workspace_bytes() */
    size_t workspace_bytes;
    double* out;               /* [6] */
    float* xm_imputed;         /* optional [B][D] */
    float* d_xm[2];            /* optional gradients, NULL = forward only */
    float* d_xs[2];
    float* d_df[2];
    float* d_mean[2];          /* direct terms only (through log p(z) - log q(z|x) and KL_reg) */
    float* d_scale[2];
} pcvae_miwae_loss_params;
size_t pcvae_miwae_loss_workspace_bytes(int rows, int samples);
int pcvae_miwae_loss(const pcvae_miwae_loss_params* p, void* stream);

/* FP32 FFMA peak probe used by bench.py for the roofline denominator: runs `iters`
 * dependent-chain-free FMA rounds on every SM; returns 0 and the FLOP count in *flops. */
int pcvae_ffma_probe(float* scratch, int iters, double* flops, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* PCVAE_B200_H */
