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
    """All gradients of a parameter list as views into ONE flat fp32 buffer, reduced bucket by bucket.

    The graphed training step (vln_ver_b200/graph.py) cannot carry torch DDP's reducer through a CUDA-graph capture
    (its bucket hooks invalidate the capture); what the data-parallel path needs is only SURVEY 8(e)'s single
    exchange -- the mean of the gradients over ranks -- so the step zeroes this buffer, lets autograd accumulate
    into the views, and calls `allreduce_mean_()` (NCCL AVG on the GPUs: capturable; sum / world on gloo).

    `bucket_bytes`: split the buffer into contiguous buckets of about that size (parameter order = flat order).
    With `overlap=True` every bucket is reduced on a side stream as soon as autograd has accumulated the last of its
    gradients (post-accumulate hooks; the event record / stream wait are captured into the CUDA graph as edges), so
    only the bucket that becomes ready last -- the query embedding, whose gradient is layer 0's d(query) -- is exposed;
    `allreduce_mean_()` then reduces what has not been launched and joins the side stream."""

    def __init__(self, params, bucket_bytes=None, overlap=False):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        spans, off = [], 0
        for p in self.params:
            assert p.dtype == torch.float32, 'master parameters are fp32'
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            spans.append((off, off + p.numel()))
            off += p.numel()
        # contiguous buckets [lo, hi) of the flat buffer and the parameters each one waits for
        self.buckets, self.bucket_of = [], {}
        limit = (bucket_bytes // 4) if bucket_bytes else total
        lo = cnt = 0
        for i, (a, b) in enumerate(spans):
            cnt += 1
            self.bucket_of[i] = len(self.buckets)
            if b - lo >= limit or i == len(spans) - 1:
                self.buckets.append([lo, b, cnt])
                lo, cnt = b, 0
        self.overlap = bool(overlap) and len(self.buckets) > 1 and self._distributed()
        self.pending = [n for _, _, n in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.comm = torch.cuda.Stream(device=dev) if (self.overlap and self.flat.is_cuda) else None
        self.hooks = []
        if self.overlap:
            for i, p in enumerate(self.params):
                self.hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(self.bucket_of[i])))

    @staticmethod
    def _distributed():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _make_hook(self, k):
        def hook(_param):
            self.pending[k] -= 1
            if self.pending[k] == 0:
                self._reduce_bucket(k, side=True)
        return hook

    def _reduce_bucket(self, k, side):
        lo, hi, _ = self.buckets[k]
        view = self.flat[lo:hi]
        self.launched[k] = True
        if self.comm is not None and side:
            cur = torch.cuda.current_stream(self.flat.device)
            self.comm.wait_stream(cur)                  # the gradients of this bucket are complete on `cur`
            with torch.cuda.stream(self.comm):
                dist.all_reduce(view, op=dist.ReduceOp.AVG)
        elif self.flat.is_cuda:
            dist.all_reduce(view, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(view)
            view /= dist.get_world_size()

    def close(self):
        """detach from the parameters (their gradients stay views of the flat buffer until reassigned)."""
        for h in self.hooks:
            h.remove()
        self.hooks = []

    def zero_(self):
        self.flat.zero_()

    def allreduce_mean_(self):
        if not self._distributed():
            return self.flat
        for k in range(len(self.buckets)):              # whatever the hooks did not launch (all of it without overlap)
            if not self.launched[k]:
                self._reduce_bucket(k, side=False)
        if self.comm is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm)
        self.pending = [n for _, _, n in self.buckets]
        self.launched = [False] * len(self.buckets)
        return self.flat
