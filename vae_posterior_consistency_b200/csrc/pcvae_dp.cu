// Data-parallel training step tail in ONE launch: deterministic reduce of this GPU's per-CTA gradient partials, exchange
// of the reduced gradient with the other GPUs of the node over NVLink peer memory, fixed-order sum over the ranks and
// torch.optim.Adam -- what `loss.backward(); all_reduce(grad); optimizer.step()` would be for the reference's train.py:
// 114-116 under data parallelism.  Replaces pcvae_reduce_grads + NCCL all-reduce + pcvae_adam_step.
//
// Every rank owns an EXCHANGE BUFFER (cudaMalloc'ed by pcvae_dp_exchange_alloc, opened on the peers through CUDA IPC):
//     float    data [2][world][P32]     slot [seq & 1][r] holds rank r's reduced gradient of call number seq
//     unsigned flag [world][nblocks]    flag [r][b] = seq once block b of rank r has delivered its 32 parameters
// Block b owns the chunks of 32 consecutive parameters b, b + gridDim, ... on every rank (same grid on all ranks).  After
// its local reduce of a chunk it PUSHES the 32 values into slot [rank] of every rank's buffer (peer stores over NVLink);
// after its last chunk it raises flag [rank][b] on every rank with release semantics; it then polls only its OWN copy
// of the flags (local memory, acquire) until all ranks have delivered block b, sums the slots in rank order -- the same
// order on every rank, so all ranks compute bit-identical gradients and weights -- and applies Adam.  The grid is sized
// so that all blocks are resident at once (no second wave queues behind blocks that wait for other ranks) and no block
// waits for another block of its own grid, so the launch cannot deadlock on residency; a rank that never arrives is
// detected by a bounded wait (status word, no hang).  The two data slots alternate by call
// number: a rank can overwrite slot s only in call seq + 2, which it reaches only after every rank has delivered call
// seq + 1, i.e. after every rank has finished reading call seq.
#include <cstring>

#include "pcvae_internal.cuh"

namespace pcvae {

constexpr int DP_MAX_WORLD = PCVAE_DP_MAX_WORLD;

struct DpArgs {
    const float* gp; int grid; long P;
    float* grad; float* theta; float* m; float* v;
    float lr_bc1, inv_sqrt_bc2, b1, b2, eps;
    const float* sp; double nll_const; double* sums;
    int world, rank; unsigned seq;
    long P32; int nblocks;
    char* peer[DP_MAX_WORLD];
    int* status;
    unsigned long long* state;      // optional device step counter (CUDA-graph replay): seq = Adam step = state[0] + 1
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One block of one rank.  `bid` / `nbid` are the block's index and the block count WITHIN its rank: blockIdx.x / gridDim.x
// in a real multi-GPU launch (k_dp_reduce_adam), a slice of the grid when several ranks are emulated in one cooperative
// launch on one GPU (k_dp_reduce_adam_emulated).
__device__ __forceinline__ void dp_block(DpArgs a, const int bid, const int nbid) {
    __shared__ float red[8][32];
    __shared__ float bc_s[2];
    __shared__ int late_s;
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    if (threadIdx.x == 0) late_s = 0;
    // launched as a programmatic dependent of the weight-gradient kernel (launch_tc): blocks may be resident before the
    // partials are complete; returns at once after an ordinary or cooperative launch
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.state) {                                        // replayable launch: call number and Adam step from the device
        const unsigned long long st = *reinterpret_cast<volatile unsigned long long*>(a.state) + 1ull;
        a.seq = (unsigned)st;
        if (threadIdx.x == 0) {
            const double bc1 = 1.0 - pow((double)a.b1, (double)st), bc2 = 1.0 - pow((double)a.b2, (double)st);
            bc_s[0] = (float)((double)a.lr_bc1 / bc1);
            bc_s[1] = (float)(1.0 / sqrt(bc2));
        }
    }
    const long P = a.P;
    const long slot_floats = a.P32;
    const long data_floats = 2 * (long)a.world * slot_floats;
    const long slot0 = (long)(a.seq & 1u) * a.world * slot_floats;
    // ---- phase 1: this block's chunks of 32 parameters (chunk c = blockIdx.x, + gridDim.x, ...): reduce the rank's
    //      partials in the order of k_reduce_adam and push the 32 values into slot [rank] of every rank ----
    for (int c = bid; c < a.nblocks; c += nbid) {
        const long i = (long)c * 32 + lane;
        float s = 0.f;
        if (i < P) {
            // warp `part` sums partials part, part + 8, ... in that order (the order of k_reduce_adam); all loads of up to
            // 160 partials are in flight before the first add: a block owns only 2 - 4 chunks and they are taken one after
            // the other, so with four loads in flight this phase was ~17 us of dependent round trips
            for (int q0 = part; q0 < a.grid; q0 += 160) {
                float v[20];
#pragma unroll
                for (int k = 0; k < 20; ++k) {
                    const int q = q0 + 8 * k;
                    v[k] = q < a.grid ? a.gp[(long)q * P + i] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 20; ++k)
                    if (q0 + 8 * k < a.grid) s += v[k];
            }
        }
        red[part][lane] = s;
        __syncthreads();
        if (part == 0) {
            float g = red[0][lane];
#pragma unroll
            for (int q = 1; q < 8; ++q) g += red[q][lane];
            const long my = slot0 + (long)a.rank * slot_floats + (long)c * 32 + lane;
            for (int r = 0; r < a.world; ++r) reinterpret_cast<float*>(a.peer[r])[my] = g;      // one 128-byte store per rank
        }
        __syncthreads();
    }
    // ---- phase 2: ONE flag per (rank, block): lane r of warp 0 raises it on rank r with release semantics (the warp's
    //      pushes of all chunks come first), then waits (acquire) for rank r's flag in the local buffer.  All blocks of
    //      the grid are resident at once (the launcher sizes the grid for that), so no block waits for a second wave ----
    if (part == 0) {
        __syncwarp();
        if (lane < a.world) {
            unsigned* f = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(a.peer[lane]) + data_floats) +
                          (long)a.rank * a.nblocks + bid;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(a.seq) : "memory");
            const unsigned* w = reinterpret_cast<const unsigned*>(reinterpret_cast<const float*>(a.peer[a.rank]) + data_floats) +
                                (long)lane * a.nblocks + bid;
            const unsigned long long t0 = globaltimer_ns();
            unsigned spins = 0, seen;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(w) : "memory");
                if ((int)(seen - a.seq) >= 0) break;     // flags only grow (wrap-safe comparison)
                __nanosleep(100);
                if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > 4000000000ull) { atomicExch(a.status, 1); late_s = 1; break; }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // a rank that did not deliver within the bound: the slots hold stale data, so this block applies NO update (the status
    // word is set; the host raises on it, dist.PeerExchange.check); weights stay what they were
    const bool late = late_s != 0;
    if (a.state) { a.lr_bc1 = bc_s[0]; a.inv_sqrt_bc2 = bc_s[1]; }
    // ---- phase 3: sum the slots in rank order (identical on every rank) and apply Adam; the block's chunks are dealt
    //      over its eight warps ----
    {
        int j = 0;
        for (int c = bid; c < a.nblocks && !late; c += nbid, ++j) {
            if ((j & 7) != part) continue;
            const long i = (long)c * 32 + lane;
            const volatile float* mine = reinterpret_cast<const volatile float*>(a.peer[a.rank]) + slot0 + (long)c * 32 + lane;
            float tot = 0.f;
            for (int r = 0; r < a.world; ++r) tot += mine[(long)r * slot_floats];
            if (i < P) {
                a.grad[i] = tot;
                const float mi = a.m[i] + (tot - a.m[i]) * (1.f - a.b1);
                const float vi = fmaf(a.b2, a.v[i], (1.f - a.b2) * tot * tot);
                a.m[i] = mi;
                a.v[i] = vi;
                a.theta[i] -= a.lr_bc1 * (mi / (sqrtf(vi) * a.inv_sqrt_bc2 + a.eps));
            }
        }
    }
    if (a.state) {                                        // the last block advances the step counter (every block has read it)
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned* ticket = reinterpret_cast<unsigned*>(a.state + 1);
            if (atomicAdd(ticket, 1u) == (unsigned)nbid - 1) {
                *ticket = 0u;
                __threadfence();
                *reinterpret_cast<volatile unsigned long long*>(a.state) = *reinterpret_cast<volatile unsigned long long*>(a.state) + 1ull;
            }
        }
    }
    if (a.sp && bid == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + PCVAE_NSUMS) {   // this rank's loss sums
        const int j = threadIdx.x - 32;
        double acc = 0.0;
        for (int c = 0; c < a.grid; ++c) acc += (double)a.sp[c * PCVAE_NSUMS + j];
        if (j == PCVAE_S_RE_Q || j == PCVAE_S_RE_P || j == PCVAE_S_RE_D || j == PCVAE_S_RE_IMP) acc += a.nll_const;
        a.sums[j] = acc;
        if (a.state) a.sums[PCVAE_NSUMS + j] += acc;      // running totals since the caller last zeroed them
    }
}

__global__ void __launch_bounds__(256) k_dp_reduce_adam(DpArgs a) { dp_block(a, blockIdx.x, gridDim.x); }

// Test vehicle (pcvae_dp_reduce_adam_emulated): `world` ranks whose buffers all live on ONE GPU, run as one cooperative
// launch -- blocks [r * nbid, (r + 1) * nbid) play rank r.  Separate launches that wait for one another on one GPU are
// not guaranteed to run concurrently (B200_PROFILING.md); a cooperative launch is.
struct DpEmuArgs { DpArgs r[PCVAE_DP_EMU_MAX_WORLD]; int nbid; };
__global__ void __launch_bounds__(256) k_dp_reduce_adam_emulated(DpEmuArgs e) {
    const int r = blockIdx.x / e.nbid;
    dp_block(e.r[r], blockIdx.x - r * e.nbid, e.nbid);
}

static long dp_p32(long P) { return (P + 31) / 32 * 32; }

}  // namespace pcvae

using namespace pcvae;

extern "C" {

size_t pcvae_dp_exchange_bytes(long param_count, int world) {
    if (param_count < 1 || world < 1 || world > DP_MAX_WORLD) return 0;
    const long P32 = dp_p32(param_count);
    return (size_t)2 * world * P32 * sizeof(float) + (size_t)world * (P32 / 32) * sizeof(unsigned);
}

int pcvae_dp_exchange_alloc(long param_count, int world, void** buffer, unsigned char* ipc_handle_64) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    const size_t bytes = pcvae_dp_exchange_bytes(param_count, world);
    if (!bytes || !buffer || !ipc_handle_64) return fail(PCVAE_EINVAL, "dp_exchange_alloc: bad arguments (world 1..%d)", DP_MAX_WORLD);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "dp_exchange_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(PCVAE_ECUDA, "dp_exchange_alloc: %s", cudaGetErrorString(e)); }
    memcpy(ipc_handle_64, &h, 64);
    *buffer = p;
    return PCVAE_OK;
}

int pcvae_dp_exchange_open(const unsigned char* ipc_handle_64, void** peer_buffer) {
    if (!ipc_handle_64 || !peer_buffer) return fail(PCVAE_EINVAL, "dp_exchange_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "dp_exchange_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    *peer_buffer = p;
    return PCVAE_OK;
}

int pcvae_dp_exchange_close(void* peer_buffer) {
    if (!peer_buffer) return PCVAE_OK;
    const cudaError_t e = cudaIpcCloseMemHandle(peer_buffer);
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "dp_exchange_close: %s", cudaGetErrorString(e));
}

int pcvae_dp_exchange_free(void* buffer) {
    if (!buffer) return PCVAE_OK;
    const cudaError_t e = cudaFree(buffer);
    return e == cudaSuccess ? PCVAE_OK : fail(PCVAE_ECUDA, "dp_exchange_free: %s", cudaGetErrorString(e));
}

static int dp_convert(const pcvae_dp_params* p, DpArgs& a) {
    if (!p || !p->grad_partials || !p->grad || !p->theta || !p->exp_avg || !p->exp_avg_sq || p->grid < 1 || p->param_count < 1 ||
        (p->step < 1 && !p->step_state) || !p->status)
        return fail(PCVAE_EINVAL, "dp_reduce_adam: bad arguments");
    if (p->world < 1 || p->world > DP_MAX_WORLD || p->rank < 0 || p->rank >= p->world || (p->seq == 0 && !p->step_state))
        return fail(PCVAE_EINVAL, "dp_reduce_adam: world %d (1..%d), rank %d, seq %u (>= 1)", p->world, DP_MAX_WORLD, p->rank, p->seq);
    if ((p->sums_partials == nullptr) != (p->sums == nullptr)) return fail(PCVAE_EINVAL, "dp_reduce_adam: sums_partials and sums go together");
    a = DpArgs{};
    for (int r = 0; r < p->world; ++r) {
        if (!p->peer_buffers[r]) return fail(PCVAE_EINVAL, "dp_reduce_adam: peer buffer %d is null", r);
        a.peer[r] = static_cast<char*>(p->peer_buffers[r]);
    }
    const int host_step = p->step_state ? 1 : p->step;
    const double bc1 = p->step_state ? 1.0 : 1.0 - pow((double)p->beta1, host_step);       // device: corrections in the kernel
    const double bc2 = p->step_state ? 1.0 : 1.0 - pow((double)p->beta2, host_step);
    a.state = p->step_state;
    a.gp = p->grad_partials; a.grid = p->grid; a.P = p->param_count;
    a.grad = p->grad; a.theta = p->theta; a.m = p->exp_avg; a.v = p->exp_avg_sq;
    a.lr_bc1 = (float)(p->lr / bc1); a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2)); a.b1 = p->beta1; a.b2 = p->beta2; a.eps = p->eps;
    a.sp = p->sums_partials; a.sums = p->sums;
    a.nll_const = 0.5 * 1.8378770664093453 * (double)p->rows * (double)p->obs_dim;
    a.world = p->world; a.rank = p->rank; a.seq = p->seq;
    a.P32 = dp_p32(p->param_count); a.nblocks = (int)(a.P32 / 32);
    a.status = p->status;
    return PCVAE_OK;
}

int pcvae_dp_reduce_adam(const pcvae_dp_params* p, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    DpArgs a;
    if (int rc = dp_convert(p, a)) return rc;
    // every block must be resident at once (a block waits for the other ranks before it ends): four per SM (1 024 threads,
    // under 64 registers each).  All ranks must launch the same grid, i.e. be the same GPU model.
    const int blocks = a.nblocks < 4 * grid ? a.nblocks : 4 * grid;
    const cudaError_t e = launch_tc(k_dp_reduce_adam, blocks, 256, 0, (cudaStream_t)stream, true, a);
    if (e != cudaSuccess) return fail(PCVAE_ECUDA, "dp_reduce_adam: launch: %s", cudaGetErrorString(e));
    return PCVAE_OK;
}

int pcvae_dp_reduce_adam_emulated(const pcvae_dp_params* const* ranks, int world, void* stream) {
    int grid;
    if (int rc = device_ok(&grid)) return rc;
    if (!ranks || world < 1 || world > PCVAE_DP_EMU_MAX_WORLD)
        return fail(PCVAE_EINVAL, "dp_reduce_adam_emulated: world %d (1..%d)", world, PCVAE_DP_EMU_MAX_WORLD);
    DpEmuArgs e{};
    for (int r = 0; r < world; ++r) {
        if (!ranks[r] || ranks[r]->world != world || ranks[r]->rank != r)
            return fail(PCVAE_EINVAL, "dp_reduce_adam_emulated: entry %d must describe rank %d of %d", r, r, world);
        if (int rc = dp_convert(ranks[r], e.r[r])) return rc;
        if (e.r[r].nblocks != e.r[0].nblocks) return fail(PCVAE_EINVAL, "dp_reduce_adam_emulated: ranks differ in param_count");
    }
    // all world * nbid blocks must be co-resident: ask the occupancy calculator, then launch cooperatively (the runtime
    // refuses a cooperative grid that does not fit instead of letting it deadlock)
    int per_sm = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_reduce_adam_emulated, 256, 0);
    if (err != cudaSuccess || per_sm < 1) return fail(PCVAE_ECUDA, "dp_reduce_adam_emulated: occupancy query: %s", cudaGetErrorString(err));
    int nbid = (per_sm * grid) / world;
    if (nbid > 2 * grid) nbid = 2 * grid;
    if (nbid > e.r[0].nblocks) nbid = e.r[0].nblocks;
    if (nbid < 1) return fail(PCVAE_EINVAL, "dp_reduce_adam_emulated: %d ranks do not fit on this GPU", world);
    e.nbid = nbid;
    void* kargs[] = {&e};
    err = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(k_dp_reduce_adam_emulated), dim3(world * nbid), dim3(256), kargs, 0,
                                      (cudaStream_t)stream);
    if (err != cudaSuccess) return fail(PCVAE_ECUDA, "dp_reduce_adam_emulated: launch: %s", cudaGetErrorString(err));
    return PCVAE_OK;
}

}  // extern "C"
