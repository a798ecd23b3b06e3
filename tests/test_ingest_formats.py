"""CPU tier for SURVEY.md 8(f) row N4: the file formats either side of the path -- view-feature ingest
(voxelformer.py:317-325) and the `getbev` export (HEAD:627-638) -- through an in-memory stand-in for the HDF5
container (h5py is not installed in this image; the adapter never re-implements the container format)."""
import numpy as np
import pytest
import torch

from vln_ver_b200 import ingest


class FakeH5(dict):
    """dict-backed stand-in with the slice of h5py.File's interface the reference uses."""
    files = {}

    def __init__(self, path, mode='r'):
        super().__init__()
        self.path, self.opens = path, 0
        if mode in ('r', 'a'):
            self.update(FakeH5.files.get(path, {}))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def create_dataset(self, key, shape, dtype=None, compression=None):
        assert compression == 'gzip' and dtype == 'float'
        self[key] = np.zeros(shape, dtype=np.float64)
        return self[key]

    def close(self):
        FakeH5.files[self.path] = dict(self)


def test_view_keys_follow_the_reference_naming():
    assert ingest.view_keys('scanA_vp7', 6) == ['scanA_vp7_i1_%d' % d for d in range(6)]
    k18 = ingest.view_keys('scanA_vp7', 18)
    assert len(k18) == 18 and k18[0] == 'scanA_vp7_i0_0' and k18[6] == 'scanA_vp7_i1_0' and k18[17] == 'scanA_vp7_i2_5'
    with pytest.raises(ValueError):
        ingest.view_keys('scanA_vp7', 12)


def test_feature_store_drops_cls_casts_and_caches():
    rng = np.random.default_rng(0)
    data = {k: rng.standard_normal((1, 197, 32)).astype(np.float16)
            for s in ('scanA_vp0', 'scanB_vp3') for k in ingest.view_keys(s, 18)}
    opened = []

    def opener(path, mode='r'):
        opened.append(path)
        f = FakeH5(path)
        f.update(data)
        return f
    store = ingest.ViewFeatureStore(opener=opener, pinned=False)
    metas = [dict(file_name='feats.h5', sample_idx='scanA_vp0'), dict(file_name='feats.h5', sample_idx='scanB_vp3')]
    batch = store.batch(metas, num_cams=18)
    assert batch.shape == (18, 2, 196, 32) and batch.dtype == torch.float32
    want = data['scanB_vp3_i2_4'][0, 1:].astype(np.float32)                  # CLS token dropped, fp32
    assert torch.equal(batch[16, 1], torch.from_numpy(want))
    n = len(opened)
    assert n == 36
    store.batch(metas, num_cams=18)                                           # second pass: served from the cache
    assert len(opened) == n
    six = store.batch(metas[:1], num_cams=6)
    assert six.shape == (6, 1, 196, 32) and torch.equal(six[:, 0], batch[6:12, 0])


def test_getbev_export_matches_the_reference_reinterpretation(tmp_path):
    C, Z, H, W, bs = 8, 2, 3, 3, 2
    bev = torch.randn(Z * H * W, bs, C)
    path = str(tmp_path / 'bev.h5')
    FakeH5.files.clear()
    metas = [dict(sample_idx='scanA_vp0'), dict(sample_idx='scanA_vp1')]
    ingest.export_bev_embed(path, metas, bev, C, Z, H, W, opener=FakeH5)
    out = FakeH5.files[path]
    assert sorted(out) == ['scanA_vp0', 'scanA_vp1']
    for b, m in enumerate(metas):
        # HEAD:633-634 for one panorama: the (Nq, 1, C) tensor viewed as (1, C, Z, H, W), squeezed, float64
        ref = bev[:, b:b + 1].contiguous().view(1, C, Z, H, W).squeeze().double().numpy()
        assert out[m['sample_idx']].dtype == np.float64 and out[m['sample_idx']].shape == (C, Z, H, W)
        assert np.array_equal(out[m['sample_idx']], ref)


def test_feature_file_and_getbev_export_through_real_hdf5_files(tmp_path):
    """end to end through the default opener (h5py if installed, else vln_ver_b200.h5min): a feature file with the
    reference's layout is written, read back by ViewFeatureStore, and a getbev export is appended twice."""
    from vln_ver_b200.ingest import _open_h5
    rng = np.random.default_rng(1)
    feats = {k: rng.standard_normal((1, 197, 16)).astype(np.float16) for k in ingest.view_keys('scanA_vp0', 18)}
    fpath = str(tmp_path / 'feats.hdf5')
    with _open_h5(fpath, 'w') as f:
        for k, v in feats.items():
            f.create_dataset(k, data=v)
    store = ingest.ViewFeatureStore(pinned=False)
    batch = store.batch([dict(file_name=fpath, sample_idx='scanA_vp0')], num_cams=18)
    assert batch.shape == (18, 1, 196, 16)
    assert torch.equal(batch[7, 0], torch.from_numpy(feats['scanA_vp0_i1_1'][0, 1:].astype(np.float32)))
    C, Z, H, W = 8, 2, 3, 3
    bpath = str(tmp_path / 'bev.hdf5')
    bev0, bev1 = torch.randn(Z * H * W, 1, C), torch.randn(Z * H * W, 1, C)
    ingest.export_bev_embed(bpath, [dict(sample_idx='scanA_vp0')], bev0, C, Z, H, W)
    ingest.export_bev_embed(bpath, [dict(sample_idx='scanA_vp1')], bev1, C, Z, H, W)      # HEAD:628-629: append
    with _open_h5(bpath, 'r') as f:
        assert sorted(f.keys()) == ['scanA_vp0', 'scanA_vp1']
        for key, bev in (('scanA_vp0', bev0), ('scanA_vp1', bev1)):
            got = np.asarray(f[key])
            assert got.dtype == np.float64 and got.shape == (C, Z, H, W)
            assert np.array_equal(got, bev.contiguous().view(1, C, Z, H, W).squeeze().double().numpy())
