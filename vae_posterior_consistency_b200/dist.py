"""Multi-GPU plumbing (one process per GPU, torch.distributed).  The hot path shards by rows:
 * active-selection reward / acquisition loop: contiguous row blocks, no data-path collective
   (SURVEY.md section 8e); histories are merged at the end (disjoint blocks: sum == gather);
 * data-parallel training: every rank takes a contiguous slice of the SAME global batch, scales
   its loss by 1/global_rows, and the flat gradient vector is all-reduced (sum) before Adam.
"""
import torch


def world():
    """(world_size, rank, group) of the default process group, (1, 0, None) when not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank(), dist.group.WORLD
    return 1, 0, None


def row_block(n_rows: int, world_size: int, rank: int):
    """[lo, hi) of the contiguous row block owned by `rank`; blocks tile [0, n_rows) exactly."""
    return (rank * n_rows) // world_size, ((rank + 1) * n_rows) // world_size


def merge_row_blocks(t: torch.Tensor, group, device=None):
    """Every rank holds `t` with only its own row block filled (zeros elsewhere): all-reduce(sum) is a gather."""
    import torch.distributed as dist
    buf = t.to(device) if device is not None else t
    dist.all_reduce(buf, group=group)
    if buf is not t:
        t.copy_(buf.cpu())
    return t


def allreduce_grads(flat_grad: torch.Tensor, group):
    import torch.distributed as dist
    dist.all_reduce(flat_grad, group=group)
    return flat_grad
