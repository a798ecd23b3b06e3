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

    Every batch is copied exactly once; nothing is copied beyond the last batch of `host_batches`.  Batches of the same
    shapes land in TWO device buffer sets used alternately (no allocation inside the loop: an allocation on the copy
    stream can end in cudaMalloc / cudaFree, which synchronise the device); a yielded batch is valid until the
    next-but-one batch is requested.  `reuse_buffers=False` gives every batch fresh tensors instead."""

    def __init__(self, host_batches, device, reuse_buffers=True):
        self.host_batches = host_batches
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise ValueError('DevicePrefetcher copies to a CUDA device')
        self.stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.reuse = reuse_buffers
        self._sets = [None, None]            # device buffer sets
        self._done = [None, None]            # event on the compute stream: the consumer of the set's last batch ran
        self._n = 0

    def _buffers(self, k, hb):
        cur = self._sets[k]
        if cur is not None and cur.keys() == hb.keys() and all(
                cur[n].shape == v.shape and cur[n].dtype == v.dtype for n, v in hb.items()):
            return cur
        self._sets[k] = {n: torch.empty(v.shape, dtype=v.dtype, device=self.device) for n, v in hb.items()}
        self._done[k] = None
        return self._sets[k]

    def _start(self, hb):
        k = self._n & 1
        self._n += 1
        if self.reuse:
            db = self._buffers(k, hb)                       # (allocated on the current stream, before the copy)
            if self._done[k] is not None:
                self.stream.wait_event(self._done[k])       # the previous contents have been consumed
            else:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for n, v in hb.items():
                    db[n].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
        else:
            with torch.cuda.stream(self.stream):
                db = {n: v.to(self.device, non_blocking=True) for n, v in hb.items()}
                ev = torch.cuda.Event()
                ev.record(self.stream)
        self.h2d_bytes += sum(v.numel() * v.element_size() for v in hb.values())
        return db, ev, k

    def __iter__(self):
        it = iter(self.host_batches)
        nxt = next(it, None)
        pending = self._start(nxt) if nxt is not None else None
        while pending is not None:
            db, ev, k = pending
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            if not self.reuse:
                for t in db.values():
                    t.record_stream(cur)          # allocated on the copy stream, consumed on the compute stream
            nxt = next(it, None)
            pending = self._start(nxt) if nxt is not None else None
            yield db
            if self.reuse:                        # everything the consumer enqueued on its stream precedes this event
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.device))
                self._done[k] = done


# --------------------------------------------------------------------------- file formats either side of the path
def view_keys(sample_idx, num_cams=6):
    """HDF5 dataset names of one panorama's view features, `<scan>_<viewpoint>_i<elevation>_<heading>`
    (voxelformer.py:317-318).  The shipped code reads elevation 1 x 6 headings (:287-288); its commented-out
    loop (:284-286) is the 18-view order: elevation-major (0, 1, 2), then heading 0..5 -- the camera order of
    `VoxelFormerEncoder._camera_tensors`."""
    scan, vp = sample_idx.split('_')
    if num_cams == 6:
        elevations = (1,)
    elif num_cams == 18:
        elevations = (0, 1, 2)
    else:
        raise ValueError(f'{num_cams} views: the feature files hold 6 headings x 3 elevations')
    return ['%s_%s_i%s_%s' % (scan, vp, e, deg) for e in elevations for deg in range(6)]


def _open_h5(path, mode='r'):
    """h5py.File when h5py is installed, otherwise vln_ver_b200.h5min.File (the subset of the container format
    these two files use: root-group datasets, contiguous or gzip-chunked, fp16 / fp32 / fp64)."""
    try:
        import h5py
        return h5py.File(path, mode)
    except ImportError:
        from . import h5min
        return h5min.File(path, mode)


class ViewFeatureStore:
    """`VoxelFormer.get_image_feature` (voxelformer.py:317-325) for whole batches: datasets are (1, 197, C) ViT
    tokens in fp16 or fp32; the CLS token is dropped and the rest cast to fp32 (`f[key][:, 1:, :].astype(np.float32)`);
    every dataset is read once and kept (the reference's `_feature_store`).  `batch()` assembles the
    (Ncam, B, 196, C) layout the head takes, in pinned memory, ready for `DevicePrefetcher`.
    `opener(path)` must return a mapping `key -> array-like` usable as a context manager (default: h5py.File)."""

    def __init__(self, opener=None, pinned=True):
        self.opener = opener or _open_h5
        self.pinned = pinned
        self._feature_store = {}

    def get(self, img_ft_file, key):
        import numpy as np
        ft = self._feature_store.get((img_ft_file, key))
        if ft is None:
            with self.opener(img_ft_file) as f:
                ft = np.asarray(f[key])[:, 1:, :].astype(np.float32)          # (1, 196, C)
            self._feature_store[(img_ft_file, key)] = ft
        return ft

    def panorama(self, img_ft_file, sample_idx, num_cams=6):
        import numpy as np
        return np.array([self.get(img_ft_file, k) for k in view_keys(sample_idx, num_cams)])    # (Ncam, 1, 196, C)

    def batch(self, img_metas, num_cams=6):
        """img_metas: dicts with `file_name` and `sample_idx` (voxelformer.py:282-288) -> (Ncam, B, 196, C) fp32."""
        import numpy as np
        feats = np.concatenate([self.panorama(m['file_name'], m['sample_idx'], num_cams) for m in img_metas], axis=1)
        t = torch.from_numpy(np.ascontiguousarray(feats))
        return t.pin_memory() if self.pinned and torch.cuda.is_available() else t


def export_bev_embed(path, img_metas, bev_embed, embed_dims, bev_z, bev_h, bev_w, opener=None):
    """The `getbev` export of VoxelFormerOccupancyHead.forward (HEAD:627-638): one gzip-compressed float64 dataset
    of shape (C, Z, H, W) per panorama, keyed by `sample_idx`, appended to `path`.  `bev_embed` is the
    (Nq, bs, C) tensor of the default branch; as in the reference the (Nq, C) block of a panorama is
    REINTERPRETED as (C, Z, H, W) (`.view`, not a permute -- SURVEY.md A4.3).  The reference handles bs = 1 and
    keys by img_metas[0]; a batch writes one dataset per panorama."""
    import os
    nq, bs, c = bev_embed.shape
    assert c == embed_dims and nq == bev_z * bev_h * bev_w and len(img_metas) == bs
    per_sample = bev_embed.detach().permute(1, 0, 2).contiguous().view(bs, embed_dims, bev_z, bev_h, bev_w)
    data = per_sample.double().cpu().numpy()
    opener = opener or _open_h5
    outf = opener(path, 'a' if os.path.exists(path) else 'w')
    try:
        for b, meta in enumerate(img_metas):
            key = meta['sample_idx']
            outf.create_dataset(key, data[b].shape, dtype='float', compression='gzip')
            outf[key][...] = data[b]
    finally:
        outf.close()
