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


def all_reduce_sum(t: torch.Tensor, group):
    """In-place sum over the ranks.  NCCL reduces device tensors directly; a gloo group (CPU tests, or several ranks
    sharing one GPU) takes the tensor through host memory."""
    import torch.distributed as dist
    if t.is_cuda and dist.get_backend(group) == "gloo":
        h = t.cpu()
        dist.all_reduce(h, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, group=group)
    return t


def merge_row_blocks(t: torch.Tensor, group, device=None):
    """Every rank holds `t` with only its own row block filled (zeros elsewhere): all-reduce(sum) is a gather."""
    import torch.distributed as dist
    if dist.get_backend(group) == "gloo" or device is None:
        dist.all_reduce(t, group=group)
        return t
    buf = t.to(device)
    dist.all_reduce(buf, group=group)
    t.copy_(buf.cpu())
    return t


def allreduce_grads(flat_grad: torch.Tensor, group):
    import torch.distributed as dist
    dist.all_reduce(flat_grad, group=group)
    return flat_grad


class PeerExchange:
    """Exchange buffers of pcvae_dp_reduce_adam (include/pcvae_b200.h): one cudaMalloc'ed buffer per rank, opened on
    every other rank of the node through CUDA IPC; the 64-byte handles travel over torch.distributed.  Collective:
    every rank of `group` must construct it (and later call `FusedTrainer.step`) in lockstep."""

    def __init__(self, param_count: int, group, device, _barrier=True):
        import ctypes as C
        import torch.distributed as dist
        from . import lib as L
        self.lib = L.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > L.DP_MAX_WORLD:
            raise L.PcvaeError(f"PeerExchange: world size {self.world} > {L.DP_MAX_WORLD}")
        self.device = torch.device(device)
        self.P = param_count
        self.seq = 0
        self.own = C.c_void_p()
        self.opened = []
        handle = C.create_string_buffer(64)
        mine = None
        try:
            with torch.cuda.device(self.device):
                L.check(self.lib.pcvae_dp_exchange_alloc(param_count, self.world, C.byref(self.own), handle),
                        "pcvae_dp_exchange_alloc")
            mine = bytes(handle.raw)
        except L.PcvaeError:               # still take part in the collective below: every rank must see the failure
            pass
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        if any(h is None for h in handles):
            self.close()
            raise L.PcvaeError("PeerExchange: exchange buffer could not be allocated on rank(s) "
                               f"{[r for r, h in enumerate(handles) if h is None]}")
        self.ptrs = [None] * self.world
        with torch.cuda.device(self.device):
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.ptrs[r] = self.own.value
                else:
                    q = C.c_void_p()
                    L.check(self.lib.pcvae_dp_exchange_open(h, C.byref(q)), "pcvae_dp_exchange_open")
                    self.ptrs[r] = q.value
                    self.opened.append(q)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        if _barrier:                                    # every buffer is zeroed and mapped before the first push
            dist.barrier(group=group)                   # (create_or_none: its all-reduce is that barrier)

    @classmethod
    def create_or_none(cls, param_count: int, group, device):
        """Collective constructor: the exchange is used only if EVERY rank could set it up (peer access / CUDA IPC can be
        unavailable, e.g. GPUs of different nodes); otherwise all ranks agree on None and the caller keeps the NCCL
        all-reduce.  Both are GPU paths; this is a choice of collective, not a fallback of the arithmetic."""
        import sys
        import torch.distributed as dist
        xch, err = None, None
        try:
            xch = cls(param_count, group, device, _barrier=False)
        except Exception as e:                      # noqa: BLE001 -- any set-up failure must reach the vote below
            err = e
        ok = torch.tensor([1 if xch is not None else 0], device=device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            return xch
        if xch is not None:
            xch.close()
        if dist.get_rank(group) == 0:
            print(f"pcvae: NVLink peer exchange unavailable ({err}); using the NCCL all-reduce", file=sys.stderr)
        return None

    def next_seq(self) -> int:
        self.seq += 1
        return self.seq

    def check(self):
        """Host-side check of the bounded wait (synchronises)."""
        from . import lib as L
        if int(self.status.item()) != 0:
            raise L.PcvaeError("pcvae_dp_reduce_adam: a rank did not deliver its gradient within the wait bound")

    def close(self):
        for q in self.opened:
            self.lib.pcvae_dp_exchange_close(q)
        self.opened = []
        if self.own is not None and self.own.value:
            self.lib.pcvae_dp_exchange_free(self.own)
        self.own = None
