"""CPU tier: the tap arithmetic of the 3-D sampler kernels (vln_ver_b200/csrc/trilinear.cuh, the
header msda3d.cu includes) compiled with g++ into a test-only harness and checked against the
oracle (the reference-owned voxel_multi_scale_deformable_attn_pytorch restated, pinned in
test_oracle_pins_reference.py).  This is NOT a product path: libver_b200.so has no host
implementation; the kernels themselves are checked in tests/test_gpu_parity.py on the B200."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import ver_ref
from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_harness', 'msda3d_host.cpp')
LIB = os.path.join(HERE, 'host_harness', 'msda3d_host.so')


@pytest.fixture(scope='module')
def harness():
    hdr = os.path.join(os.path.dirname(HERE), 'vln_ver_b200', 'csrc', 'trilinear.cuh')
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++', SRC, '-o', LIB],
                       check=True)
    return ctypes.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_case(seed, Bv, shapes, NH, Dh, Nq, NP):
    g = torch.Generator().manual_seed(seed)
    NL = len(shapes)
    S = sum(d * h * w for d, h, w in shapes)
    value = torch.randn(Bv, S, NH, Dh, generator=g)
    loc = torch.rand(Bv, Nq, NH, NL, NP, 3, generator=g) * 1.5 - 0.25        # some points outside
    d, h, w = shapes[0]
    loc[0, 0, 0, 0, 0] = torch.tensor([0.0, 0.0, 0.0])                       # volume corner
    loc[0, 0, 0, 0, 1] = torch.tensor([1.0, 1.0, 1.0])
    loc[0, 0, 0, 0, 2] = torch.tensor([-3.0, 0.5, 0.5])                      # far outside
    loc[0, 0, 0, 0, 3 % NP] = torch.tensor([0.5, 0.5, float('nan')])         # NaN -> contributes nothing
    wts = torch.rand(Bv, Nq, NH, NL, NP, generator=g)
    wts = wts / wts.sum((-1, -2), keepdim=True)
    gout = torch.randn(Bv, Nq, NH * Dh, generator=g)
    return value, loc, wts, gout, S


@pytest.mark.parametrize('shapes,NH,Dh,Nq,NP', [
    ([(3, 5, 7)], 4, 8, 13, 4),
    ([(4, 6, 6)], 8, 96, 9, 4),
    ([(3, 4, 5), (2, 2, 3)], 2, 40, 7, 3),
])
def test_tap_arithmetic_matches_oracle(harness, shapes, NH, Dh, Nq, NP):
    Bv = 2
    value, loc, wts, gout, S = make_case(7, Bv, shapes, NH, Dh, Nq, NP)
    NL = len(shapes)
    sh = np.asarray(shapes, dtype=np.int32)
    v, l, w, go = (x.numpy().copy() for x in (value, loc, wts, gout))
    out = np.empty((Bv, Nq, NH * Dh), np.float32)
    harness.msda3d_host_forward(_p(v), _p(sh), NL, _p(l), _p(w), _p(out), Bv, S, NH, Dh, Nq, NP)
    gv, gl, gw = np.empty_like(v), np.empty_like(l), np.empty_like(w)
    harness.msda3d_host_backward(_p(v), _p(sh), NL, _p(l), _p(w), _p(go), _p(gv), _p(gl), _p(gw),
                                 Bv, S, NH, Dh, Nq, NP)
    # oracle in fp64; the NaN location is replaced by a far-outside one (grid_sample would
    # propagate the NaN; the kernels treat it as outside -- documented in include/ver_b200.h)
    loc64 = loc.double().clone()
    loc64[torch.isnan(loc64)] = -5.0
    v64 = value.double().requires_grad_(True)
    l64 = loc64.requires_grad_(True)
    w64 = wts.double().requires_grad_(True)
    ref = ver_ref.voxel_multi_scale_deformable_attn_pytorch(v64, torch.tensor(shapes), l64, w64)
    gv_r, gl_r, gw_r = torch.autograd.grad(ref, (v64, l64, w64), gout.double())
    assert rel_err(torch.from_numpy(out), ref) < 1e-5
    assert rel_err(torch.from_numpy(gv), gv_r) < 1e-5
    assert rel_err(torch.from_numpy(gw), gw_r) < 1e-5
    # d/d loc is one-sided where a coordinate sits exactly on a voxel centre / the padding border
    # (the planted corner points): exclude those two, as test_msda_golden does for the 2-D op
    gl_t, gl_ref = torch.from_numpy(gl).double(), gl_r.clone()
    gl_t[0, 0, 0, 0, :4] = 0
    gl_ref[0, 0, 0, 0, :4] = 0
    assert rel_err(gl_t, gl_ref) < 1e-5
