"""CPU tier: vln_ver_b200/h5min.py, the minimal HDF5 container used for the feature / getbev files when h5py is
absent.  The READER is checked against a file written by a real HDF5 library (the MATLAB 7.3
sample scipy ships: superblock v0 behind a 512-byte user block, symbol-table group, v1 object header, v2 layout);
the WRITER by round trips through that reader and by structural checks of what it lays out."""
import glob
import os
import struct

import numpy as np
import pytest

from vln_ver_b200 import h5min


def _matlab_sample():
    import scipy.io
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), 'matlab', 'tests', 'data', 'testhdf5_7.4_GLNX86.mat'))
    return hits[0] if hits else None


@pytest.mark.skipif(_matlab_sample() is None, reason='scipy test data not installed')
def test_reader_on_a_file_written_by_libhdf5():
    with h5min.File(_matlab_sample()) as f:
        assert f.keys() == ['testdouble']
        a = f['testdouble']
    assert a.dtype == np.float64 and a.shape == (9, 1)
    assert np.allclose(a[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_round_trip_contiguous_and_gzip(tmp_path):
    rng = np.random.default_rng(0)
    data = {'f16': rng.standard_normal((1, 197, 24)).astype(np.float16),
            'f32': rng.standard_normal((3, 5)).astype(np.float32),
            'i64': rng.integers(-9, 9, (7, 2)),
            'u8': rng.integers(0, 255, (4,), dtype=np.uint8),
            'empty': np.zeros((0, 3), np.float32)}
    big = rng.standard_normal((8, 2, 15, 15))
    path = str(tmp_path / 'a.h5')
    with h5min.File(path, 'w') as f:
        for k, v in data.items():
            f.create_dataset(k, data=v)
        ds = f.create_dataset('bev', big.shape, dtype='float', compression='gzip')     # HEAD:635-636
        ds[...] = big
        assert 'bev' in f and np.array_equal(f['bev'][1:3], big[1:3])
        with pytest.raises(ValueError):
            f.create_dataset('bev', (1,), dtype='float')
    with h5min.File(path) as f:
        assert f.keys() == sorted(list(data) + ['bev'])
        for k, v in data.items():
            got = f[k]
            assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
        assert f['bev'].dtype == np.float64 and np.array_equal(f['bev'], big)
        with pytest.raises(KeyError):
            f['nope']
        with pytest.raises(h5min.H5Error):
            f.create_dataset('x', (1,))
    # the gzip dataset really is compressed on disk
    const = str(tmp_path / 'c.h5')
    with h5min.File(const, 'w') as f:
        f.create_dataset('z', (64, 4, 15, 15), dtype='float', compression='gzip')
    assert os.path.getsize(const) < 64 * 4 * 15 * 15 * 8 // 20


def test_append_and_many_datasets_across_symbol_nodes(tmp_path):
    path = str(tmp_path / 'm.h5')
    n = 5000                                               # > 2 * MAX_LEAF_K links: several SNODs under the root B-tree
    names = ['scan%04d_vp_i%d_%d' % (i // 18, (i // 6) % 3, i % 6) for i in range(n)]
    with h5min.File(path, 'w') as f:
        for i, k in enumerate(names):
            f.create_dataset(k, data=np.full((2,), i, np.float32))
    with h5min.File(path, 'a') as f:
        f.create_dataset('zz_last', data=np.arange(3, dtype=np.float32))
        f.create_dataset('00_first', data=np.arange(4, dtype=np.float32))
    with h5min.File(path) as f:
        assert len(f) == n + 2 and f.keys()[0] == '00_first' and f.keys()[-1] == 'zz_last'
        for i in (0, 1, 2047, 2048, 2049, 4095, 4096, n - 1):
            assert f[names[i]][0] == i
        assert np.array_equal(f['00_first'], np.arange(4, dtype=np.float32))


def test_written_structures_follow_the_format_specification(tmp_path):
    path = str(tmp_path / 's.h5')
    with h5min.File(path, 'w') as f:
        f.create_dataset('b', data=np.arange(6, dtype=np.float32).reshape(2, 3))
        f.create_dataset('a', (4, 2), dtype='float', compression='gzip')
    d = open(path, 'rb').read()
    assert d[:8] == b'\x89HDF\r\n\x1a\n' and d[8] == 0 and d[13] == 8 and d[14] == 8          # superblock v0, 8-byte fields
    leaf_k, internal_k = struct.unpack_from('<HH', d, 16)
    base, _, eof, _ = struct.unpack_from('<QQQQ', d, 24)
    assert base == 0 and eof == len(d)
    root, cache = struct.unpack_from('<QI', d, 64)
    tree, heap = struct.unpack_from('<QQ', d, 80)
    assert cache == 1 and d[tree:tree + 4] == b'TREE' and d[heap:heap + 4] == b'HEAP'
    assert d[root] == 1 and root % 8 == 0                                                   # v1 object header, aligned
    used, = struct.unpack_from('<H', d, tree + 6)
    snod, = struct.unpack_from('<Q', d, tree + 32)
    assert used == 1 and d[snod:snod + 4] == b'SNOD'
    nsym, = struct.unpack_from('<H', d, snod + 6)
    seg, = struct.unpack_from('<Q', d, heap + 24)
    offs = [struct.unpack_from('<Q', d, snod + 8 + 40 * i)[0] for i in range(nsym)]
    names = [d[seg + o:d.index(b'\0', seg + o)].decode() for o in offs]
    assert names == ['a', 'b']                                                              # links sorted by name
    assert snod + 8 + 2 * leaf_k * 40 <= len(d) and tree + 24 + 8 + 2 * internal_k * 16 <= len(d)   # full-size nodes
    r = h5min._Reader(d)
    types = [m[0] for m in r.messages(r.links(r.root_header)['a'])]
    assert types == [0x0001, 0x0003, 0x0005, 0x000B, 0x0008]                                # space, type, fill, filters, layout
    for m in r.messages(r.links(r.root_header)['a']):
        assert m[1] % 8 == 0 and m[2] % 8 == 0                                              # 8-byte aligned messages


def test_unsupported_files_fail_loudly(tmp_path):
    p = str(tmp_path / 'x.h5')
    open(p, 'wb').write(b'not hdf5' * 100)
    with pytest.raises(h5min.H5Error):
        h5min.File(p)
    v2 = bytearray(b'\x89HDF\r\n\x1a\n' + bytes([2]) + b'\0' * 87)
    open(p, 'wb').write(bytes(v2))
    with pytest.raises(h5min.H5Error, match='superblock version 2'):
        h5min.File(p)
