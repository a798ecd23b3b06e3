"""CPU tier for SURVEY.md 8(f) row N1: the lattice form of the head's `up_sample` stack
(vln_ver_b200/upsample.py) equals the three ConvTranspose3d as the reference writes them (HEAD:254-258),
values and every gradient, in fp64; and the structural fact it rests on (odd rows / columns of every
layer's output are the bare bias)."""
import pytest
import torch
import torch.nn as nn

from vln_ver_b200.upsample import lattice_supported, up_sample_lattice


def stack(C, n=3, seed=0, bias=True):
    torch.manual_seed(seed)
    return nn.Sequential(*[nn.ConvTranspose3d(C, C, (3, 5, 5), stride=(1, 2, 2), padding=(2, 4, 4), dilation=(2, 2, 2),
                                              output_padding=(0, 1, 1), bias=bias) for _ in range(n)]).double()


@pytest.mark.parametrize('shape', [(2, 6, 4, 5, 6), (1, 6, 1, 1, 1), (1, 6, 5, 2, 7), (1, 6, 4, 15, 15)])
def test_lattice_equals_dense_fp64(shape):
    convs = stack(shape[1])
    x = torch.randn(*shape, dtype=torch.float64, requires_grad=True)
    ref = convs(x)
    y = up_sample_lattice(x, convs)
    assert y.shape == ref.shape == (shape[0], shape[1], shape[2], 8 * shape[3], 8 * shape[4])
    assert (ref - y).abs().max().item() < 1e-13
    g = torch.randn_like(ref)
    params = [x] + list(convs.parameters())
    for a, b in zip(torch.autograd.grad(y, params, g), torch.autograd.grad(ref, params, g)):
        assert (a - b).abs().max().item() < 1e-12 * max(1.0, b.abs().max().item())


def test_odd_rows_and_columns_are_the_bare_bias():
    convs = stack(4, n=1)
    y = convs(torch.randn(1, 4, 3, 5, 5, dtype=torch.float64))
    b = convs[0].bias.view(1, 4, 1, 1, 1)
    assert torch.equal(y[:, :, :, 1::2, :], b.expand_as(y[:, :, :, 1::2, :]))
    assert torch.equal(y[:, :, :, :, 1::2], b.expand_as(y[:, :, :, :, 1::2]))
    assert not torch.equal(y[:, :, :, 0::2, 0::2], b.expand_as(y[:, :, :, 0::2, 0::2]))


def test_no_bias_and_unsupported_hyper_parameters():
    convs = stack(4, n=2, bias=False)
    x = torch.randn(1, 4, 2, 3, 3, dtype=torch.float64)
    assert (convs(x) - up_sample_lattice(x, convs)).abs().max().item() < 1e-13
    other = nn.Sequential(nn.ConvTranspose3d(4, 4, 3, stride=2))
    assert lattice_supported(convs) and not lattice_supported(other)
    with pytest.raises(ValueError):
        up_sample_lattice(x.float(), other)


# ------------------------------------------------------------------ GEMM + col2im execution
import ctypes      # noqa: E402
import os          # noqa: E402
import subprocess  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def col_harness():
    """csrc/convt_index.cuh (the header col2im.cu includes) compiled with g++ into a test-only harness."""
    src = os.path.join(HERE, 'host_harness', 'col2im_host.cpp')
    lib = os.path.join(HERE, 'host_harness', 'col2im_host.so')
    hdr = os.path.join(os.path.dirname(HERE), 'vln_ver_b200', 'csrc', 'convt_index.cuh')
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-x', 'c++', src, '-o', lib], check=True)
    return ctypes.CDLL(lib)


class _HostCol2Im(torch.autograd.Function):
    """TEST-ONLY stand-in for ops.convt_col2im_fn running the kernels' index header on the host (fp64)."""
    harness = None

    @staticmethod
    def forward(ctx, cols, Z, Hi, Wi, s):
        B, n_in, taps, C = cols.shape
        ctx.dims = (B, Z, Hi, Wi, s, C)
        cols = cols.contiguous()
        out = torch.empty(B, Z * s * Hi * s * Wi, C, dtype=torch.float64)
        _HostCol2Im.harness.col2im_host(ctypes.c_void_p(cols.data_ptr()), ctypes.c_void_p(out.data_ptr()), B, Z, Hi, Wi, s, C)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Z, Hi, Wi, s, C = ctx.dims
        g = g.contiguous()
        gc = torch.empty(B, Z * Hi * Wi, 75, C, dtype=torch.float64)
        _HostCol2Im.harness.im2col_host(ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(gc.data_ptr()), B, Z, Hi, Wi, s, C)
        return gc, None, None, None, None


@pytest.mark.parametrize('shape', [(2, 8, 4, 5, 6), (1, 8, 1, 1, 1), (1, 8, 5, 2, 7), (1, 8, 4, 15, 15)])
def test_gemm_col2im_form_equals_dense_fp64(col_harness, shape):
    """the GEMM + col2im execution (weight matrix layout, tap order, the kernels' index relation and its adjoint)
    against the three ConvTranspose3d as written: values and every gradient."""
    from vln_ver_b200.upsample import up_sample_gemm
    _HostCol2Im.harness = col_harness
    convs = stack(shape[1], seed=3)
    x = torch.randn(*shape, dtype=torch.float64, requires_grad=True)
    ref = convs(x)
    y = up_sample_gemm(x, convs, col2im=_HostCol2Im.apply)
    assert y.shape == ref.shape
    assert (ref - y).abs().max().item() < 1e-13
    g = torch.randn_like(ref)
    params = [x] + list(convs.parameters())
    for a, b in zip(torch.autograd.grad(y, params, g), torch.autograd.grad(ref, params, g)):
        assert (a - b).abs().max().item() < 1e-12 * max(1.0, b.abs().max().item())


def test_col2im_library_entry_points_validate_arguments():
    from vln_ver_b200 import _lib
    rc = _lib.lib.ver_convt_col2im(1, None, None, 1, 1, 1, 1, 2, 768, None)
    assert rc == -1 and b'null' in _lib.lib.ver_last_error()
    buf = ctypes.c_void_p(4096)
    assert _lib.lib.ver_convt_col2im(1, buf, buf, 1, 1, 1, 1, 3, 768, None) == -1        # stride
    assert _lib.lib.ver_convt_im2col(1, buf, buf, 1, 1, 1, 1, 2, 12, None) == -1         # channels % 8


def test_dense_execution_and_the_choice_between_them():
    import vln_ver_b200 as V
    from vln_ver_b200.upsample import pick_execution, up_sample, up_sample_dense
    convs = stack(6, seed=5)
    x = torch.randn(1, 6, 2, 3, 3, dtype=torch.float64)
    assert torch.equal(up_sample_dense(x, convs), convs(x))
    with pytest.raises(V.VerError):                        # the product dispatcher has no CPU path
        up_sample(x, convs)
    assert pick_execution(torch.float32, 900) == 'gemm'
    assert pick_execution(torch.float16, 900) == 'dense'                   # one shipped-size panorama
    assert pick_execution(torch.float16, 8 * 900) == 'gemm'
