"""Host -> device ingest of view-feature batches (the step in front of the path).

The reference reads each panorama's ViT tokens from HDF5 inside `forward` and copies them synchronously
(projects/mmdet3d_plugin/bevformer/detectors/voxelformer.py:317-325).  Here a batch is a dict of PINNED
host tensors (`feats (Ncam, B, 196, C)`, `l2i (B, Ncam, 4, 4)`, `sh (B, 3)`, labels ...) and
`DevicePrefetcher` keeps one batch in flight on a copy stream: the H2D copy of batch i+1 is enqueued
before batch i's kernels are launched, so the copy engine runs under the compute of the previous step
and the step only ever waits on an event.
"""
import torch


def pin(batch):
    """dict of CPU tensors -> the same dict in page-locked memory (async copies need it)."""
    return {k: (v if v.is_pinned() else v.pin_memory()) for k, v in batch.items()}


class DevicePrefetcher:
    """Iterate host batches as device batches, copying one batch ahead on a side stream.

        for batch in DevicePrefetcher(host_batches, device):
            outs = head(batch['feats'], None, lidar2img=batch['l2i'], originshift=batch['sh'])

    Every batch is copied exactly once; nothing is copied beyond the last batch of `host_batches`."""

    def __init__(self, host_batches, device):
        self.host_batches = host_batches
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise ValueError('DevicePrefetcher copies to a CUDA device')
        self.stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0

    def _start(self, hb):
        with torch.cuda.stream(self.stream):
            db = {k: v.to(self.device, non_blocking=True) for k, v in hb.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.h2d_bytes += sum(v.numel() * v.element_size() for v in hb.values())
        return db, ev

    def __iter__(self):
        it = iter(self.host_batches)
        nxt = next(it, None)
        pending = self._start(nxt) if nxt is not None else None
        while pending is not None:
            db, ev = pending
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in db.values():
                t.record_stream(cur)              # allocated on the copy stream, consumed on the compute stream
            nxt = next(it, None)
            pending = self._start(nxt) if nxt is not None else None
            yield db
