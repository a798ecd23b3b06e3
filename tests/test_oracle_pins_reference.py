"""CPU tier, build container only: re-runs the UNMODIFIED reference (imported from
/root/reference through oracle/mmcv_shim.py) and checks that (a) the committed golden
fixtures are what the reference produces and (b) the oracle restatement agrees with it."""

import pytest
import torch

from oracle import mmcv_shim, ver_ref
from vln_ver_b200 import synth
from conftest import load_golden, rel_err, sub

pytestmark = pytest.mark.skipif(not mmcv_shim.reference_available(),
                                reason='/root/reference is only present in the build container')
PC = synth.PC_RANGE


def test_unmodified_point_sampling_reproduces_fixture():
    from oracle import gen_golden
    enc_mod = mmcv_shim.import_reference('bevformer.modules.voxel_encoder')
    g = load_golden('point_sampling_6cam.npz')
    c = sub(g, 'g4x15x15')
    ref_3d, rpc, mask = gen_golden.run_unmodified_point_sampling(
        enc_mod, 4, 15, 15, c['lidar2img'][0].numpy(), c['originshift'][0].numpy())
    assert torch.equal(rpc, c['rpc']) and torch.equal(mask, c['mask'])


def test_unmodified_sca_18_views_matches_oracle_on_fresh_inputs():
    sca_mod = mmcv_shim.import_reference('bevformer.modules.spatial_cross_attention')
    torch.manual_seed(7)
    C, ncam, grid = 64, 18, (2, 6, 6)
    m = sca_mod.SpatialCrossAttention(
        embed_dims=C, num_cams=ncam, pc_range=PC, dropout=0.1, batch_first=True,
        deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C, num_points=8,
                                  num_levels=1)).eval()
    l2i, sh = synth.make_rig(1, ncam, grid, seed=77)
    rpc, mask = ver_ref.point_sampling_batched(*grid, PC, torch.from_numpy(l2i), torch.from_numpy(sh))
    q = torch.randn(1, 72, C)
    v = torch.randn(ncam, 196, 1, C)
    with torch.no_grad():
        y = m(q, v, v, reference_points_cam=rpc, bev_mask=mask, spatial_shapes=torch.tensor([[14, 14]]),
              level_start_index=torch.tensor([0]))
        y2 = ver_ref.sca_forward(dict(m.state_dict()), '', q, v, rpc, mask, torch.tensor([[14, 14]]))
    assert rel_err(y2, y) < 1e-6


def test_reference_3d_sampler_at_depth_one_equals_2d_restatement():
    vtsa = mmcv_shim.import_reference('bevformer.modules.voxel_temporal_self_attention')
    g = torch.Generator().manual_seed(3)
    v = torch.randn(2, 35, 4, 8, generator=g, dtype=torch.float64)
    loc = torch.rand(2, 11, 4, 1, 6, 2, generator=g, dtype=torch.float64) * 1.4 - 0.2
    w = torch.rand(2, 11, 4, 1, 6, generator=g, dtype=torch.float64)
    loc3 = torch.cat([loc, torch.full_like(loc[..., :1], 0.5)], -1)
    a = vtsa.voxel_multi_scale_deformable_attn_pytorch(v, [(1, 5, 7)], loc3, w)
    b = ver_ref.multi_scale_deformable_attn_pytorch(v, torch.tensor([[5, 7]]), loc, w)
    assert (a - b).abs().max().item() < 1e-13


def test_reference_3d_sampler_equals_restatement_on_fresh_inputs():
    """N2/N3: the reference-owned 3-D sampler vs oracle/ver_ref.py, values and gradients, two levels."""
    vtsa = mmcv_shim.import_reference('bevformer.modules.voxel_temporal_self_attention')
    g = torch.Generator().manual_seed(5)
    shapes = torch.tensor([[3, 5, 7], [2, 3, 4]])
    S = 3 * 5 * 7 + 2 * 3 * 4
    v = torch.randn(2, S, 4, 8, generator=g, dtype=torch.float64, requires_grad=True)
    loc = (torch.rand(2, 11, 4, 2, 3, 3, generator=g, dtype=torch.float64) * 1.6 - 0.3).requires_grad_(True)
    w = torch.rand(2, 11, 4, 2, 3, generator=g, dtype=torch.float64, requires_grad=True)
    a = vtsa.voxel_multi_scale_deformable_attn_pytorch(v, shapes, loc, w)
    b = ver_ref.voxel_multi_scale_deformable_attn_pytorch(v, shapes, loc, w)
    assert torch.equal(a, b)
    go = torch.randn(a.shape, generator=g, dtype=torch.float64)
    for x, y in zip(torch.autograd.grad(a, (v, loc, w), go), torch.autograd.grad(b, (v, loc, w), go)):
        assert torch.equal(x, y)


def test_unmodified_decoder_reproduces_fixture():
    """the committed decoder fixture is what the unmodified VoxelDetectionTransformerDecoder /
    VoxelCustomMSDeformableAttention / VoxelTemporalSelfAttention produce from the stored weights."""
    from oracle import gen_golden
    vd = mmcv_shim.import_reference('bevformer.modules.voxel_decoder')
    vtsa = mmcv_shim.import_reference('bevformer.modules.voxel_temporal_self_attention')
    mmcv_shim.register_detr_decoder_layer()
    g = load_golden('decoder_c64.npz')
    grid = [int(x) for x in g['grid']]
    ss = torch.tensor([grid])
    a, d, t = sub(g, 'attn'), sub(g, 'dec'), sub(g, 'tsa')
    attn = vd.VoxelCustomMSDeformableAttention(embed_dims=64, num_levels=1, batch_first=False).eval()
    attn.load_state_dict(sub(g, 'attn.sd'))
    dec = mmcv_shim.build_transformer_layer_sequence(gen_golden.decoder_cfg(64, 2)).eval()
    dec.load_state_dict(sub(g, 'dec.sd'))
    regs = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 10))
                                for _ in range(2)]).eval()
    regs.load_state_dict(sub(g, 'reg.sd'))
    tsa = vtsa.VoxelTemporalSelfAttention(embed_dims=64, num_levels=1, batch_first=True).eval()
    tsa.sampling_offsets.bias.data = torch.zeros(tsa.sampling_offsets.out_features)
    tsa.load_state_dict(sub(g, 'tsa.sd'))
    with torch.no_grad():
        y = attn(a['query'], key=None, value=a['value'], query_pos=a['query_pos'], reference_points=a['ref'],
                 spatial_shapes=ss, level_start_index=torch.tensor([0]))
        hs, refs = dec(query=a['query'], key=None, value=a['value'], query_pos=a['query_pos'],
                       reference_points=d['ref'], reg_branches=regs, cls_branches=None, spatial_shapes=ss,
                       level_start_index=torch.tensor([0]))
        yt = tsa(t['query'], query_pos=t['query_pos'], reference_points=t['ref'], spatial_shapes=ss,
                 level_start_index=torch.tensor([0]))
    assert torch.equal(y, a['out']) and torch.equal(hs, d['hs']) and torch.equal(refs, d['refs'])
    assert torch.equal(yt, t['out'])


def test_shipped_temporal_self_attention_cannot_run_as_initialised():
    """the shipped init assigns a 2-component bias to a 3-component Linear."""
    vtsa = mmcv_shim.import_reference('bevformer.modules.voxel_temporal_self_attention')
    m = vtsa.VoxelTemporalSelfAttention(embed_dims=64, num_levels=1).eval()
    assert m.sampling_offsets.bias.numel() != m.sampling_offsets.out_features
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 8, 64), reference_points=torch.zeros(2, 8, 1, 3), spatial_shapes=torch.tensor([[2, 2, 2]]))
