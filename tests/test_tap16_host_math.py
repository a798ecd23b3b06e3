"""CPU tier: the tap arithmetic of the forward sampler generations on 16-cell image rows
(vln_ver_b200/csrc/tap16.cuh, the header sca_tc6.cu / sca_tc7.cu include) compiled with g++ into a test-only harness
and checked against the bilinear weights of F.grid_sample(bilinear, zeros, align_corners=False) -- the sampling rule
of MSDeformableAttention3D (M/multi_scale_deformable_attn_function.py:29-53).  NOT a product path: libver_b200.so has no
host implementation; the kernels themselves are checked in tests/test_gpu_parity.py on the B200."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_harness', 'tap16_host.cpp')
LIB = os.path.join(HERE, 'host_harness', 'tap16_host.so')


@pytest.fixture(scope='module')
def harness():
    hdr = os.path.join(os.path.dirname(HERE), 'vln_ver_b200', 'csrc', 'tap16.cuh')
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++', SRC, '-o', LIB], check=True)
    lib = ctypes.CDLL(LIB)
    lib.tap16_row.restype = ctypes.c_uint
    lib.tap16_row.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    return lib


def _row(lib, px, py, aw, Sh, Sw):
    """dense (Sh, Sw) interpolation row of the points (pixel coordinates px, py; weights aw) from the harness."""
    x1 = np.ascontiguousarray(px + 1.0, dtype=np.float32)
    y1 = np.ascontiguousarray(py + 1.0, dtype=np.float32)
    a = np.ascontiguousarray(aw, dtype=np.float32)
    row = np.zeros(Sh * 16, dtype=np.float32)
    km = lib.tap16_row(x1.ctypes.data, y1.ctypes.data, a.ctypes.data, len(a), Sh, Sw, row.ctypes.data)
    assert not km & 0x80000000, 'weight reached the sink word'
    row = row.reshape(Sh, 16)
    assert np.all(row[:, 0] >= 0) and np.all(row[:, Sw + 1:] >= 0)
    return row[:, 1:Sw + 1], row, km


def _grid_sample_row(px, py, aw, Sh, Sw):
    """the same row from grid_sample itself: sample the one-hot images."""
    eye = torch.eye(Sh * Sw, dtype=torch.float64).view(1, Sh * Sw, Sh, Sw)
    gx = (torch.from_numpy(px).double() + 0.5) / Sw * 2 - 1          # pixel = loc * size - 0.5, grid = 2 loc - 1
    gy = (torch.from_numpy(py).double() + 0.5) / Sh * 2 - 1
    grid = torch.stack([gx, gy], -1).view(1, 1, -1, 2)
    s = torch.nn.functional.grid_sample(eye, grid, mode='bilinear', padding_mode='zeros', align_corners=False)
    return (s[0, :, 0, :] * torch.from_numpy(aw).double()[None]).sum(1).view(Sh, Sw).numpy()


@pytest.mark.parametrize('Sh,Sw', [(14, 14), (7, 10), (9, 5), (2, 2), (13, 14)])
def test_tap16_rows_are_the_bilinear_weights(harness, Sh, Sw):
    rng = np.random.default_rng(Sh * 100 + Sw)
    for trial in range(200):
        n = 8
        px = rng.uniform(-3.0, Sw + 2.0, n).astype(np.float32)
        py = rng.uniform(-3.0, Sh + 2.0, n).astype(np.float32)
        if trial % 4 == 0:        # integers, borders, far outside, coinciding points
            px[:4] = [0.0, Sw - 1.0, -1.0, float(Sw)]
            py[:4] = [0.0, Sh - 1.0, 0.5, -1.0]
            px[4], py[4] = px[5], py[5]
            px[6], py[6] = 1e6, -1e6
        aw = rng.uniform(0.05, 1.0, n).astype(np.float32)
        got, full, km = _row(harness, px, py, aw, Sh, Sw)
        want = _grid_sample_row(px, py, aw, Sh, Sw)
        assert np.abs(got - want).max() < 2e-6 * max(1.0, np.abs(want).max())
        # weight of corners outside the map lands on padding cells only; every touched image row is in the chunk mask
        touched = np.nonzero(np.abs(full).sum(1))[0]
        assert all((km >> int(y)) & 1 for y in touched)


def test_tap16_nan_and_inf_contribute_nothing(harness):
    px = np.array([np.nan, np.inf, 3.25, -np.inf], dtype=np.float32)
    py = np.array([2.0, 2.0, np.nan, np.inf], dtype=np.float32)
    aw = np.ones(4, dtype=np.float32)
    got, full, _ = _row(harness, px, py, aw, 14, 14)
    # NaN / inf y: zero weight.  NaN x with a finite y is clamped to the left padding cell; +inf x to the right one.
    assert np.abs(got).max() == 0.0 and np.isfinite(full).all()
