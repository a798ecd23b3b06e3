"""CPU tier: the C-ABI library loads and exports every symbol declared in include/*.h,
argument validation works without touching a GPU, ops refuse CPU tensors, and the
registry boundary (B1) builds the vocc.py head tree with the reference's state_dict keys."""
import ctypes
import glob
import os
import re

import pytest
import torch

import vln_ver_b200 as V
from vln_ver_b200 import _lib, ops, registry, synth
from vln_ver_b200.config import per_voxel_occupancy_size

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        text = re.sub(r'/\*.*?\*/', '', open(h).read(), flags=re.S)
        names += re.findall(r'\b(ver_[a-z0-9_]+)\s*\(', text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 13
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(dll, s), f'{s} declared in include/ but not exported'
    assert set(_lib.EXPORTED) == set(syms)
    assert dll.ver_abi_version() == _lib.ABI_VERSION


def test_argument_validation_without_gpu():
    rc = _lib.lib.ver_msda_forward(0, None, None, 1, None, None, None, 1, 1, 1, 1, 1, 1, None)
    assert rc == -1 and b'null' in _lib.lib.ver_last_error()
    rc = _lib.lib.ver_sca_forward(7, None, 0, None, 0, None, None, None, *([1] * 10), None)
    assert rc == -1
    with pytest.raises(_lib.VerError):
        _lib.check(rc)


def test_ops_refuse_cpu_tensors():
    with pytest.raises(V.VerError):
        ops.point_sampling(torch.zeros(1, 6, 4, 4), torch.zeros(1, 3), synth.PC_RANGE, 2, 2, 2)
    with pytest.raises(V.VerError):
        ops.ms_deform_attn_forward(torch.zeros(1, 4, 1, 8), [[2, 2]], None, torch.zeros(1, 1, 1, 1, 1, 2),
                                   torch.zeros(1, 1, 1, 1, 1))


def test_vocc_head_tree_builds_with_reference_state_dict_keys():
    head = V.build_head(V.vocc_head_cfg())          # shipped vocc.py values
    head.init_weights()
    keys = set(head.state_dict().keys())
    expect = ['voxel_embedding.weight', 'positional_encoding.row_embed.weight',
              'positional_encoding.col_embed.weight', 'positional_encoding.z_embed.weight',
              'transformer.level_embeds', 'transformer.cams_embeds', 'occ_proj.weight', 'occ_proj.bias',
              'up_sample.0.weight', 'up_sample.2.bias']
    for i in range(3):
        p = f'transformer.encoder.layers.{i}.'
        expect += [p + 'attentions.0.output_proj.weight',
                   p + 'attentions.0.deformable_attention.sampling_offsets.bias',
                   p + 'attentions.0.deformable_attention.attention_weights.weight',
                   p + 'attentions.0.deformable_attention.value_proj.weight',
                   p + 'ffns.0.layers.0.0.weight', p + 'ffns.0.layers.1.bias',
                   p + 'norms.0.weight', p + 'norms.1.bias']
    expect += [f'occ_branches.{i}.weight' for i in (0, 1, 3, 4, 6)]
    missing = [k for k in expect if k not in keys]
    assert not missing, missing
    assert head.state_dict()['voxel_embedding.weight'].shape == (900, 768)
    assert head.state_dict()['transformer.level_embeds'].shape == (4, 768)
    assert head.state_dict()['occ_proj.weight'].shape == (128 * 35, 4 * 768)
    assert (head.occ_xdim, head.occ_ydim, head.occ_zdim, head.voxel_num) == (120, 120, 35, 504000)
    enc_params = sum(p.numel() for p in head.transformer.encoder.layers[0].parameters())
    assert enc_params == 3693504
    # offset bias init: head h points along (cos, sin)(2 pi h / 8) / max|.|, scaled by p + 1
    b = head.transformer.encoder.layers[0].attentions[0].deformable_attention.sampling_offsets.bias
    b = b.view(8, 1, 8, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.arange(1., 9.)) and torch.allclose(b[2, 0, 3], torch.tensor([0., 4.]), atol=1e-6)


def test_registry_semantics_and_errors():
    with pytest.raises(KeyError):
        registry.build_from_cfg(dict(type='Nope'), registry.ATTENTION)
    with pytest.raises(ValueError):
        registry.build_attention(dict(type='MSDeformableAttention3D', embed_dims=100, num_heads=8))
    sca = registry.build_attention(dict(type='SpatialCrossAttention', embed_dims=256, num_cams=18,
                                        deformable_attention=dict(type='MSDeformableAttention3D',
                                                                  embed_dims=256, num_levels=1)))
    assert sca.fp16_enabled is False and sca.deformable_attention.output_proj is None
    with pytest.raises(NameError):
        sca(torch.zeros(1, 4, 256), torch.zeros(18, 196, 1, 256), None, residual=torch.zeros(1, 4, 256),
            spatial_shapes=[[14, 14]], reference_points_cam=torch.zeros(18, 1, 4, 1, 2),
            bev_mask=torch.zeros(18, 1, 4, 1, dtype=torch.bool))


def test_per_voxel_head_dims_and_synth_rig():
    for grid in [(8, 20, 20), (16, 40, 40), (20, 20, 20), (40, 40, 40), (16, 80, 80), (4, 15, 15)]:
        s = per_voxel_occupancy_size(*grid)
        pc = synth.PC_RANGE
        assert int((pc[3] - pc[0]) / s[0]) == grid[2] and int((pc[5] - pc[2]) / s[2]) == grid[0]
    l2i, sh = synth.make_rig(2, 18, (8, 20, 20))
    assert l2i.shape == (2, 18, 4, 4) and sh.shape == (2, 3)
    l2i2, _ = synth.make_rig(2, 18, (8, 20, 20))
    assert (l2i == l2i2).all()                         # seeded


def test_reference_point_grid_matches_oracle_bit_for_bit():
    from oracle import ver_ref
    a = V.VoxelFormerEncoder.get_reference_points(4, 15, 15, dim='3d', bs=2, device='cpu')
    assert torch.equal(a, ver_ref.get_reference_points_3d(4, 15, 15, bs=2))


def test_half_cache_follows_parameter_versions():
    """fused_layer.half_of: one fp16 copy per optimizer step -- refreshed when the parameter is updated in place
    (optimizer.step bumps _version) or replaced, reused otherwise; concatenations follow all their parts."""
    from vln_ver_b200 import fused_layer as F
    a = torch.nn.Parameter(torch.randn(4, 8))
    b = torch.nn.Parameter(torch.randn(2, 8))
    h1 = F.half_of(a)
    assert h1.dtype == torch.float16 and F.half_of(a) is h1
    cat1 = F.half_of(a, b)
    assert cat1.shape == (6, 8) and F.half_of(a, b) is cat1
    with torch.no_grad():
        a.add_(1.0)
    h2 = F.half_of(a)
    assert h2 is not h1 and torch.equal(h2, a.detach().half())
    cat2 = F.half_of(a, b)
    assert cat2 is not cat1 and torch.equal(cat2[:4], a.detach().half())
    f = F.f32_cat(a, b)
    assert f.dtype == torch.float32 and F.f32_cat(a, b) is f
    # single-parameter copies are refreshed TOGETHER: the first stale lookup casts every registered copy that is out of date
    hb = F.half_of(b)
    with torch.no_grad():
        a.mul_(0.5)
        b.add_(2.0)
    assert torch.equal(F.half_of(a), a.detach().half())
    assert F._SINGLES[id(b)][2] == (F._GENERATION[0], b._version, b.data_ptr())         # b already refreshed, by a's lookup
    assert F.half_of(b) is not hb and torch.equal(F.half_of(b), b.detach().half())


def test_weight_cache_is_invalidated_explicitly_and_by_hooks():
    """the fp16 weight copies are keyed on the autograd version counter, which `.data` writes do
    not bump.  Contract now: such writes must be followed by invalidate_weight_cache(); optimizer steps and
    load_state_dict are hooked; entries die with their parameters (weak references)."""
    import gc

    import torch
    from vln_ver_b200 import fused_layer as FL
    FL.invalidate_weight_cache()
    lin = torch.nn.Linear(8, 8)
    a = FL.half_of(lin.weight)
    assert FL.half_of(lin.weight) is a                              # cached
    with torch.no_grad():
        lin.weight.add_(1.0)                                        # bumps the version counter
    b = FL.half_of(lin.weight)
    assert b is not a and torch.equal(b, lin.weight.detach().half())
    lin.weight.data.mul_(2.0)                                       # bypasses the version counter ...
    assert FL.half_of(lin.weight) is b                              # ... so the copy is stale by construction
    FL.invalidate_weight_cache()                                    # the documented remedy
    c = FL.half_of(lin.weight)
    assert c is not b and torch.equal(c, lin.weight.detach().half())
    # hooks: optimizer.step() and load_state_dict() invalidate
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    FL.install_cache_hooks(lin, opt)
    d = FL.half_of(lin.weight)
    lin.weight.grad = torch.ones_like(lin.weight)
    lin.bias.grad = torch.ones_like(lin.bias)
    opt.step()
    assert FL.half_of(lin.weight) is not d
    e = FL.half_of(lin.weight)
    lin.load_state_dict({k: v.clone() * 0 for k, v in lin.state_dict().items()})
    f = FL.half_of(lin.weight)
    assert f is not e and float(f.abs().max()) == 0.0
    # no leak: entries go away with the module
    n_before = len(FL._HALF_CACHE) + len(FL._SINGLES)
    del lin, opt, a, b, c, d, e, f
    gc.collect()
    assert len(FL._HALF_CACHE) + len(FL._SINGLES) < max(n_before, 1)


def test_spatial_shapes_are_read_once():
    import torch
    from vln_ver_b200 import ops
    assert ops.shapes_to_host([[14, 14]]) == [[14, 14]]
    assert ops.shapes_to_host(torch.tensor([[14, 14], [7, 7]])) == [[14, 14], [7, 7]]
    assert ops._shapes_arg(torch.tensor([[14, 14]]))[1] == 1
