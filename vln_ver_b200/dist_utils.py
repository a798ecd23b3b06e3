"""Host-side helpers of the data-parallel path: panoramas shard
across ranks, weights are replicated, the forward needs no collective and the only exchange is
the gradient all-reduce (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """contiguous, balanced [lo, hi) slice of `n_items` panoramas owned by `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device='cpu'):
    """device-timed milliseconds -> max over ranks (the bench's timing rule)."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def allreduce_mean_(tensors):
    """in-place mean over ranks of a list of gradient tensors (one flat bucket)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat)
    flat /= dist.get_world_size()
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return tensors


class FlatGradients:
    """All gradients of a parameter list as views into ONE flat fp32 buffer, reduced with one collective.

    The graphed training step (vln_ver_b200/graph.py) cannot carry torch DDP's reducer through a CUDA-graph capture
    (its bucket hooks invalidate the capture); what the data-parallel path needs is only SURVEY 8(e)'s single
    exchange -- the mean of the gradients over ranks -- so the step zeroes this buffer, lets autograd accumulate
    into the views, and calls `allreduce_mean_()` (NCCL AVG on the GPUs: capturable; sum / world on gloo)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            assert p.dtype == torch.float32, 'master parameters are fp32'
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def allreduce_mean_(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.flat
        if self.flat.is_cuda:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat)
            self.flat /= dist.get_world_size()
        return self.flat
