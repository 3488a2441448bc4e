"""Worker of tests/test_gpu_dp.py (one process per GPU, launched by torch.distributed.run): the fused
reduce + NVLink peer exchange + Adam kernel (pcvae_dp_reduce_adam) against the reduce -> NCCL all-reduce -> Adam
sequence and against a single-GPU step on the whole batch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pcvae_oracle as O                                        # noqa: E402
from vae_posterior_consistency_b200 import kernels as KR, lib as L          # noqa: E402
from vae_posterior_consistency_b200.dist import row_block                   # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", local)
    B, D, steps = 4096 + 37, 100, 4
    p = O.init_params("mlp", D, seed=1)
    g = torch.Generator().manual_seed(5)
    xs = [torch.rand(B, D, generator=g) for _ in range(steps)]
    ms = [torch.rand(B, D, generator=g) < 0.7 for _ in range(steps)]
    mps = [m & (torch.rand(B, D, generator=g) < 0.7) for m in ms]
    eqs = [torch.randn(B, 10, generator=g) for _ in range(steps)]
    eps = [torch.randn(B, 10, generator=g) for _ in range(steps)]
    lo, hi = row_block(B, world, rank)

    def run(mode):
        theta = KR.flatten_params(p, L.FAMILY_MLP, dev)
        if mode == "single":
            tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta, regularised=True)
            sl = slice(0, B)
        else:
            os.environ["PCVAE_DP"] = mode
            tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta, regularised=True, dist_group=dist.group.WORLD, world_size=world)
            assert (tr.xch is not None) == (mode == "peer")
            sl = slice(lo, hi)
        losses = []
        for s in range(steps):
            loss = tr.step(xs[s][sl].to(dev), ms[s][sl].to(dev), mps[s][sl].to(dev), eqs[s][sl].to(dev), eps[s][sl].to(dev),
                           global_rows=B)
            if mode != "single":
                dist.all_reduce(loss)                      # local sums are per rank: the global loss is their sum
            losses.append(float(loss))
        torch.cuda.synchronize()
        if tr.xch is not None:
            tr.xch.check()
            tr.xch.close()
        return tr.theta.clone(), tr.grad.clone(), losses

    th_peer, g_peer, l_peer = run("peer")
    th_nccl, g_nccl, l_nccl = run("nccl")
    th_one, g_one, l_one = run("single")
    # every rank holds bit-identical weights after the fused path
    gathered = [torch.empty_like(th_peer) for _ in range(world)]
    dist.all_gather(gathered, th_peer)
    for r in range(world):
        assert torch.equal(gathered[r], gathered[0]), f"rank {r} diverged from rank 0"
    scale = float(g_one.abs().max())
    torch.testing.assert_close(g_peer, g_nccl, rtol=1e-4, atol=1e-6 * scale)
    torch.testing.assert_close(g_peer, g_one, rtol=2e-3, atol=2e-5 * scale)
    torch.testing.assert_close(th_peer, th_nccl, rtol=0, atol=2e-5)
    torch.testing.assert_close(th_peer, th_one, rtol=0, atol=2e-4)       # Adam normalises: small gradient differences move a weight by < lr
    for a, b, c in zip(l_peer, l_nccl, l_one):
        assert abs(a - b) <= 1e-5 * abs(b) and abs(a - c) <= 1e-4 * abs(c), (a, b, c)
    # a second trainer on fresh buffers starts its own sequence (no stale flags)
    th2, _, _ = run("peer")
    assert torch.equal(th2, th_peer)
    # the same tail inside a CUDA graph of the whole step (device-counted seq / Adam step): replay == eager launches
    os.environ["PCVAE_DP"] = "peer"
    T, Bl, nb = 5000, 512, 6
    gt = torch.Generator(device=dev).manual_seed(11)
    table = torch.rand(T, D, device=dev, generator=gt)
    mtable = torch.rand(T, D, device=dev, generator=gt) < 0.7
    gi = torch.Generator(device=dev).manual_seed(100 + rank)                # every rank gathers its own rows
    idx = torch.stack([torch.randperm(T, device=dev, generator=gi)[:Bl] for _ in range(nb)])
    res = []
    for graph in (True, False):
        tr = KR.GraphedFusedTrainer(L.FAMILY_MLP, D, 0, KR.flatten_params(p, L.FAMILY_MLP, dev), table, mtable, Bl, nb,
                                    keep=0.7, seed=5, dist_group=dist.group.WORLD, world_size=world, global_rows=Bl * world)
        tr.set_batches(idx)
        if graph:
            tr.capture(warmup=2)
            for _ in range(7):
                tr.step_graph()
        else:
            for _ in range(9):
                tr.step_eager_dev()
        torch.cuda.synchronize()
        tr.xch.check()
        res.append((tr.theta.clone(), float(tr.total), int(tr.state[0])))
        tr.xch.close()
    assert res[0][2] == res[1][2] == 9
    assert torch.equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
    gathered = [torch.empty_like(res[0][0]) for _ in range(world)]
    dist.all_gather(gathered, res[0][0])
    for r in range(world):
        assert torch.equal(gathered[r], gathered[0]), f"graph replay: rank {r} diverged from rank 0"
    assert not torch.equal(res[0][0], KR.flatten_params(p, L.FAMILY_MLP, dev))          # it did train
    dist.barrier()
    if rank == 0:
        print(f"dp_worker ok: world {world}, losses {l_peer}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
