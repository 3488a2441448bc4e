"""Data-parallel training tail on two GPUs of one node (skipped on a single-GPU box): see tests/dp_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_peer_exchange_adam_matches_nccl_and_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(here, "dp_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "dp_worker ok" in out.stdout


def _dp_params(L, KR, eng, gp, grad, theta, m, v, step, sp, rows, sums, world, rank, seq, bufs, status, state=None):
    import ctypes as C
    p = L.DpParams()
    p.grad_partials, p.grid, p.param_count = gp.data_ptr(), eng.grid, eng.P
    p.grad, p.theta, p.exp_avg, p.exp_avg_sq = grad.data_ptr(), theta.data_ptr(), m.data_ptr(), v.data_ptr()
    p.step, p.lr, p.beta1, p.beta2, p.eps = step, 1e-3, 0.9, 0.999, 1e-8
    p.sums_partials, p.rows, p.obs_dim, p.sums = sp.data_ptr(), rows, eng.D, sums.data_ptr()
    p.world, p.rank, p.seq = world, rank, seq
    for r in range(world):
        p.peer_buffers[r] = bufs[r].data_ptr()
    p.status = status.data_ptr()
    p.step_state = None if state is None else state.data_ptr()
    return p


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("device_counted", [False, True])
def test_dp_exchange_kernel_with_emulated_ranks_is_bit_exact(world, device_counted):
    """pcvae_dp_reduce_adam's kernel body with `world` ranks emulated on ONE GPU (one cooperative launch; separate
    launches that wait for one another are not safe on one GPU): against pcvae_reduce_adam per rank (same reduction
    order) + rank-ordered sum + pcvae_adam_step, bit for bit, over several calls on the same exchange buffers."""
    import ctypes as C
    from vae_posterior_consistency_b200 import kernels as KR, lib as L
    lib = L.load()
    dev = torch.device("cuda", 0)
    eng = KR.Engine(L.FAMILY_MLP, 100, 0, dev)
    P, grid = eng.P, eng.grid
    g = torch.Generator(device=dev).manual_seed(17 + world)
    nbytes = lib.pcvae_dp_exchange_bytes(P, world)
    assert nbytes > 0
    bufs = [torch.zeros((nbytes + 3) // 4, dtype=torch.float32, device=dev) for _ in range(world)]
    theta0 = torch.randn(P, device=dev, generator=g)
    th = [theta0.clone() for _ in range(world)]
    ms = [torch.zeros(P, device=dev) for _ in range(world)]
    vs = [torch.zeros(P, device=dev) for _ in range(world)]
    grads = [torch.empty(P, device=dev) for _ in range(world)]
    sums = [torch.zeros(2 * L.NSUMS, dtype=torch.float64, device=dev) for _ in range(world)]
    status = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
    state = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)] if device_counted else [None] * world
    th_ref, m_ref, v_ref = theta0.clone(), torch.zeros(P, device=dev), torch.zeros(P, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rows = 4096
    for call in range(1, 5):
        gps = [torch.randn(grid, P, device=dev, generator=g) * (0.01 * (r + 1)) for r in range(world)]
        sps = [torch.rand(grid, L.NSUMS, device=dev, generator=g) for _ in range(world)]
        params = [_dp_params(L, KR, eng, gps[r], grads[r], th[r], ms[r], vs[r], call, sps[r], rows, sums[r], world, r, call,
                             bufs, status[r], state[r]) for r in range(world)]
        arr = (C.POINTER(L.DpParams) * world)(*[C.pointer(p) for p in params])
        L.check(lib.pcvae_dp_reduce_adam_emulated(arr, world, st), "pcvae_dp_reduce_adam_emulated")
        # reference: each rank's reduced gradient in the order of pcvae_reduce_adam, summed in rank order, one Adam step
        tot = None
        for r in range(world):
            g_r, scratch = torch.empty(P, device=dev), [torch.zeros(P, device=dev) for _ in range(3)]
            s_r = torch.empty(L.NSUMS, dtype=torch.float64, device=dev)
            L.check(lib.pcvae_reduce_adam(gps[r].data_ptr(), grid, P, g_r.data_ptr(), scratch[0].data_ptr(), scratch[1].data_ptr(),
                                          scratch[2].data_ptr(), 1, 1e-3, 0.9, 0.999, 1e-8, sps[r].data_ptr(), rows, eng.D,
                                          s_r.data_ptr(), st), "pcvae_reduce_adam")
            tot = (torch.zeros(P, device=dev) + g_r) if tot is None else tot + g_r
            assert torch.equal(sums[r][:L.NSUMS], s_r)
        L.check(lib.pcvae_adam_step(th_ref.data_ptr(), tot.data_ptr(), m_ref.data_ptr(), v_ref.data_ptr(), P, call, 1e-3, 0.9,
                                    0.999, 1e-8, st), "pcvae_adam_step")
        torch.cuda.synchronize()
        for r in range(world):
            assert int(status[r].item()) == 0
            assert torch.equal(grads[r], tot), f"call {call}: rank {r} gradient differs from the rank-ordered sum"
            assert torch.equal(ms[r], m_ref) and torch.equal(vs[r], v_ref)
            assert torch.equal(th[r], th[0]), f"call {call}: rank {r} weights differ from rank 0"
            if device_counted:
                # the bias corrections are formed on the device from the float learning rate (pcvae_reduce_adam_dev does
                # the same): equal to the host-side double computation to the last bit or so of lr / (1 - beta1^t)
                assert int(state[r][0].item()) == call
                torch.testing.assert_close(th[r], th_ref, rtol=1e-6, atol=1e-9)
            else:
                assert torch.equal(th[r], th_ref), f"call {call}: rank {r} weights differ from reduce + sum + Adam"
    assert not torch.equal(th_ref, theta0)
