"""Host-side multi-rank logic on CPU with the gloo backend, world_size 2: row-block partitioning, the
disjoint-block merge used by the acquisition loop, and the data-parallel gradient rule (slice of the same
global batch, loss scaled by 1/global_rows, all-reduce sum) -- the oracle stands in for the kernels."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pcvae_oracle as O
    from vae_posterior_consistency_b200 import dist as D
    torch.set_num_threads(1)
    ws, rk, group = D.world()
    assert (ws, rk) == (world, rank)
    # --- row blocks + merge (acquisition loop histories) ---
    n = 11
    lo, hi = D.row_block(n, ws, rk)
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    mine = torch.zeros(n, 3)
    mine[lo:hi] = full[lo:hi]
    D.merge_row_blocks(mine, group)
    ok_merge = torch.equal(mine, full)
    # --- data-parallel gradient rule ---
    B, Dm = 23, 7
    p = O.init_params("mlp", Dm, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, Dm, generator=g)
    mask = torch.rand(B, Dm, generator=g) < 0.7
    mask_p = mask & (torch.rand(B, Dm, generator=g) < 0.7)
    eq, ep = torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g)
    lo, hi = D.row_block(B, ws, rk)
    loss_l, grads_l, _ = O.train_step(p, x[lo:hi], mask[lo:hi], mask_p[lo:hi], eq[lo:hi], ep[lo:hi])
    scale = (hi - lo) / B                                  # oracle divides by local rows; rule is 1/global_rows
    flat = torch.cat([grads_l[k].reshape(-1) for k in O.trainable_names(p)]) * scale
    D.allreduce_grads(flat, group)
    loss = loss_l * scale
    dist.all_reduce(loss)
    ref_loss, ref_grads, _ = O.train_step(p, x, mask, mask_p, eq, ep)
    ref_flat = torch.cat([ref_grads[k].reshape(-1) for k in O.trainable_names(p)])
    ok_dp = torch.allclose(flat, ref_flat, rtol=1e-4, atol=1e-6) and torch.allclose(loss, ref_loss, rtol=1e-5)
    q.put((rank, ok_merge, ok_dp))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_row_blocks_tile_exactly():
    sys.path.insert(0, ROOT)
    from vae_posterior_consistency_b200.dist import row_block
    for n in (0, 1, 7, 8, 100000, 2000):
        for w in (1, 2, 4, 8):
            blocks = [row_block(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_graph_epoch_host_logic_without_a_gpu():
    """train._graph_epoch (throughput mode, CUDA-graph replay): host side only -- the sampler's full batches go to the
    trainer as one index table, the first epoch's warm-up steps count as steps, later epochs replay every batch, and the
    host RNG is consumed as iter(DataLoader) would (base seed, then the sampler's permutation)."""
    import torch
    from torch.utils.data import DataLoader, TensorDataset
    from vae_posterior_consistency_b200 import train as T

    class FakeTrainer:
        def __init__(self, B):
            self.B, self.graph, self.step_count, self.calls = B, None, 0, []
            self.table = self.mtable = None
            self._total = torch.zeros((), dtype=torch.float64)

        @property
        def total(self):
            return self._total

        def reset_total(self):
            self._total = torch.zeros((), dtype=torch.float64)

        def set_batches(self, idx):
            self.calls.append(("set", idx.clone(), self.step_count))

        def capture(self, warmup=3):
            self.graph = object()
            self.step_count += warmup
            self._total += warmup
            self.calls.append(("capture", warmup))

        def step_graph(self):
            self.step_count += 1
            self._total += 1
            self.calls.append(("replay",))

    N, B = 64, 8
    loader = DataLoader(TensorDataset(torch.arange(N)), batch_size=B, shuffle=True)
    tr = FakeTrainer(B)
    torch.manual_seed(5)
    tot1 = T._graph_epoch(tr, loader, torch.device("cpu"), True, 10)
    tot2 = T._graph_epoch(tr, loader, torch.device("cpu"), True, 10)
    state_after = torch.get_rng_state()
    assert float(tot1) == N // B and float(tot2) == N // B and tr.step_count == 2 * (N // B)
    kinds = [c[0] for c in tr.calls]
    assert kinds == ["set", "capture"] + ["replay"] * (N // B - 3) + ["set"] + ["replay"] * (N // B)
    assert tr.calls[1] == ("capture", 3)
    # the index tables are the epochs' permutations, batch by batch, and the RNG stream is the DataLoader's
    torch.manual_seed(5)
    ref = []
    for _ in range(2):
        ref.append(torch.stack([b[0] for b in loader]))
    assert torch.equal(torch.get_rng_state(), state_after)
    sets = [c for c in tr.calls if c[0] == "set"]
    assert torch.equal(sets[0][1], ref[0]) and sets[0][2] == 0
    assert torch.equal(sets[1][1], ref[1]) and sets[1][2] == N // B
