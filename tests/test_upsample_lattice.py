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


# ------------------------------------------------------------------ occ_proj on the reinterpreted volume
@pytest.mark.parametrize('C,Z,H,W,out_f', [(12, 4, 3, 3, 10), (24, 2, 3, 6, 7)])
def test_occ_proj_from_lattice_equals_the_reinterpreted_dense_linear(C, Z, H, W, out_f):
    """HEAD:564-572 on the up-sampled volume (raw .view to (Z, X, Y, C), permute, flatten, Linear) against the
    lattice evaluation that never builds the volume: values and every gradient in fp64."""
    import torch.nn.functional as F
    from vln_ver_b200.upsample import occ_proj_from_lattice
    convs = stack(C, seed=9)
    torch.manual_seed(1)
    lin = nn.Linear(Z * C, out_f).double()
    x = torch.randn(2, C, Z, H, W, dtype=torch.float64, requires_grad=True)
    Xo, Yo = 8 * H, 8 * W
    ref = F.linear(convs(x).contiguous().view(2, Z, Xo, Yo, C).permute(0, 2, 3, 1, 4).flatten(3), lin.weight, lin.bias)
    e, b = up_sample_lattice(x, convs, assemble=False)
    y = occ_proj_from_lattice(e, b, lin.weight, lin.bias, Xo, Yo)
    assert y.shape == ref.shape and (ref - y).abs().max().item() < 1e-13
    g = torch.randn_like(ref)
    params = [x, lin.weight, lin.bias] + list(convs.parameters())
    for a, c in zip(torch.autograd.grad(y, params, g), torch.autograd.grad(ref, params, g)):
        assert (a - c).abs().max().item() < 1e-12 * max(1.0, c.abs().max().item())


def test_occ_proj_plan_at_the_shipped_shape():
    """vocc.py: 768 channels, 4 x 120 x 120 after up-sampling, rows of 4 x 768 inputs: five data patterns (y mod 5)
    with 720 / 768 / 816 data positions -- a quarter of the dense GEMM."""
    from vln_ver_b200.upsample import _occ_proj_plan, occ_proj_plan_supported
    assert occ_proj_plan_supported(768, 4, 120, 120, 120, 120)
    assert not occ_proj_plan_supported(768, 2, 24, 24, 24, 24)          # runs would straddle channels
    groups, chan = _occ_proj_plan(768, 4, 120, 120, 120, 120, 'cpu')
    assert sorted(len(c) for _, c, _ in groups) == [720, 720, 768, 816, 816]
    assert all(len(r) == 2880 for r, _, _ in groups) and chan.shape == (14400, 4)
    for r, _, _ in groups:
        assert len(set((r % 120 % 5).tolist())) == 1                     # a pattern is a residue of y mod 5


def test_head_tail_modes_agree_with_the_oracle_on_cpu(monkeypatch):
    """The shipped head tail (HEAD:551-580) through every CPU-runnable combination of up_sample_mode /
    occ_proj_mode against the oracle.  The head's LayerNorm goes through the CUDA library on the product path;
    here -- test only -- it is replaced by torch's so the host logic can run on CPU tensors."""
    import torch.nn.functional as F
    import vln_ver_b200 as V
    from oracle import ver_ref
    from vln_ver_b200.modules import voxelformer_occupancy_head as H
    monkeypatch.setattr(H, 'apply_layernorm', lambda ln, y: F.layer_norm(y, ln.normalized_shape, ln.weight, ln.bias, ln.eps))
    grid = (4, 3, 3)
    cfg = V.vocc_head_cfg(*grid, num_cams=6, embed_dims=768, only_occ=False, refine_occ=True,
                          occupancy_size=[0.5, 0.5, 0.5], occ_dims=8, num_layers=1, with_decoder=False)
    torch.manual_seed(3)
    head = V.build_head(cfg).eval()
    with torch.no_grad():
        for p in head.up_sample.parameters():
            p.mul_(0.5)
    assert (head.occ_xdim, head.occ_ydim, head.occ_zdim) == (24, 24, 7)
    bev = torch.randn(2, 36, 768)
    sd = {k: v.detach() for k, v in head.state_dict().items()}
    with torch.no_grad():
        ref = ver_ref.occ_head(sd, '', bev, *grid, 24, 24, 7, occ_dims=8, refine_occ=True, only_occ=False)
        outs = {}
        for up, occ in (('dense', 'dense'), ('lattice', 'dense'), ('lattice', 'lattice')):
            head.up_sample_mode, head.occ_proj_mode = up, occ
            outs[up, occ] = head._occupancy_tail(bev, 2)
    for k, y in outs.items():
        assert y.shape == ref.shape == (2, 7 * 24 * 24, 16)
        assert ((y - ref).abs().max() / ref.abs().max()).item() < 1e-4, k


def test_head_tail_without_refine_matches_the_golden_fixture_on_cpu(monkeypatch):
    """the only_occ head tail (the bench / smoke path: per-voxel occ_proj and the column occ_proj branch) against the
    fixture from the unmodified reference head, run on CPU with the LayerNorm swapped for torch's (test only)."""
    import torch.nn.functional as F
    import vln_ver_b200 as V
    from vln_ver_b200.config import per_voxel_occupancy_size
    from vln_ver_b200.modules import voxelformer_occupancy_head as H
    from conftest import load_golden, rel_err, sub
    monkeypatch.setattr(H, 'apply_layernorm', lambda ln, y: F.layer_norm(y, ln.normalized_shape, ln.weight, ln.bias, ln.eps))
    g = load_golden('head.npz')
    for tag in ('pervoxel', 'column'):
        c, sd = sub(g, tag), sub(g, tag + '.sd')
        grid = tuple(c['grid'].tolist())
        osz = per_voxel_occupancy_size(*grid) if tag == 'pervoxel' else [2.0, 2.0, 0.5]
        cfg = V.vocc_head_cfg(*grid, num_cams=6, embed_dims=32, only_occ=True, refine_occ=False,
                              occupancy_size=osz, occ_dims=16, num_layers=1)
        head = V.build_head(cfg).eval()
        head.load_state_dict(sd, strict=False)
        with torch.no_grad():
            y = head._occupancy_tail(c['bev_embed'], 1)
        assert rel_err(y, c['occupancy_preds']) < 1e-5
