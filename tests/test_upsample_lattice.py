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
