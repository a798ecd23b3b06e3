"""CPU tier for SURVEY.md 8(f) rows N2 / N3 (detection decoder, temporal self-attention):
(a) the oracle restatements against the golden vectors gen_golden.py produced from the UNMODIFIED
    reference classes;
(b) the host logic of the product modules (registry build from the vocc.py-shaped cfg, state_dict keys,
    permutes, reference-point refinement) against the same vectors.  The product sampler is CUDA-only
    (no CPU path exists in the product), so for (b) -- and only inside these tests -- the op is
    monkeypatched with the oracle's sampler; the kernels themselves are checked on the B200 in
    tests/test_gpu_parity.py and their tap arithmetic in tests/test_msda3d_host_math.py."""
import pytest
import torch

import vln_ver_b200 as V
from oracle import ver_ref
from vln_ver_b200 import _lib, ops, registry
from conftest import load_golden, rel_err, sub


@pytest.fixture(scope='module')
def gold():
    g = load_golden('decoder_c64.npz')
    return g, [int(x) for x in g['grid']]


def test_msda3d_restatement_matches_reference():
    g = load_golden('msda3d_cases.npz')
    for name in ('small', 'dh96', 'two_level'):
        c = sub(g, name)
        v = c['value'].double().requires_grad_(True)
        l = c['loc'].double().requires_grad_(True)
        w = c['w'].double().requires_grad_(True)
        out = ver_ref.voxel_multi_scale_deformable_attn_pytorch(v, c['shapes'], l, w)
        assert rel_err(out, c['out']) < 1e-12
        gv, gl, gw = torch.autograd.grad(out, (v, l, w), c['gout'].double())
        assert rel_err(gv, c['gvalue']) < 1e-12
        assert rel_err(gl, c['gloc']) < 1e-12
        assert rel_err(gw, c['gw']) < 1e-12


def test_decoder_restatements_match_unmodified_modules(gold):
    g, grid = gold
    ss = torch.tensor([grid])
    a = sub(g, 'attn')
    y = ver_ref.voxel_custom_msda_forward(sub(g, 'attn.sd'), '', a['query'], a['value'], a['ref'], ss,
                                          query_pos=a['query_pos'])
    assert rel_err(y, a['out']) < 1e-6
    d = sub(g, 'dec')
    regs = _reg_branches(g)
    hs, refs = ver_ref.decoder_forward(sub(g, 'dec.sd'), '', a['query'], a['value'], a['query_pos'], d['ref'], ss,
                                       num_layers=2, reg_branches=regs)
    assert rel_err(hs, d['hs']) < 2e-6 and rel_err(refs, d['refs']) < 1e-6
    t = sub(g, 'tsa')
    yt = ver_ref.temporal_self_attention_forward(sub(g, 'tsa.sd'), '', t['query'], t['ref'], ss,
                                                 query_pos=t['query_pos'])
    assert rel_err(yt, t['out']) < 1e-6


def _reg_branches(g, C=64, L=2):
    regs = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(C, C), torch.nn.ReLU(), torch.nn.Linear(C, 10))
                                for _ in range(L)]).eval()
    regs.load_state_dict(sub(g, 'reg.sd'))
    return regs


def decoder_cfg(C=64, L=2):
    """vocc.py:137-158 (decoder=dict(...)) at the fixture's width."""
    return dict(
        type='VoxelDetectionTransformerDecoder', num_layers=L, return_intermediate=True,
        transformerlayers=dict(
            type='DetrTransformerDecoderLayer',
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='VoxelCustomMSDeformableAttention', embed_dims=C, num_levels=1)],
            ffn_cfgs=dict(type='FFN', embed_dims=C, feedforward_channels=1024, num_fcs=2, ffn_drop=0.,
                          act_cfg=dict(type='ReLU', inplace=True)),
            feedforward_channels=2 * C, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))


@pytest.fixture
def oracle_sampler(monkeypatch):
    """TEST-ONLY stand-in for the CUDA sampler so that module plumbing can run on CPU tensors."""
    monkeypatch.setattr(ops, 'voxel_multi_scale_deformable_attn',
                        lambda v, shapes, loc, w: ver_ref.voxel_multi_scale_deformable_attn_pytorch(
                            v, torch.as_tensor(shapes), loc, w))


def test_product_sampler_has_no_cpu_path():
    with pytest.raises(V.VerError):
        ops.voxel_multi_scale_deformable_attn(torch.zeros(1, 8, 1, 8), [[2, 2, 2]], torch.zeros(1, 1, 1, 1, 1, 3),
                                              torch.zeros(1, 1, 1, 1, 1))
    rc = _lib.lib.ver_msda3d_forward(0, None, None, 1, None, None, None, 1, 1, 1, 1, 1, 1, None)
    assert rc == -1 and b'null' in _lib.lib.ver_last_error()


def test_attention_module_host_logic(gold, oracle_sampler):
    g, grid = gold
    a = sub(g, 'attn')
    m = registry.build_attention(dict(type='VoxelCustomMSDeformableAttention', embed_dims=64, num_levels=1,
                                      batch_first=False)).eval()
    assert m.sampling_offsets.out_features == 8 * 1 * 4 * 3 and m.fp16_enabled is False
    b = m.sampling_offsets.bias.view(8, 1, 4, 3)        # (cos, sin, cos+sin)/max|.| scaled by p+1
    assert torch.allclose(b[0, 0, :, 0], torch.arange(1., 5.)) and torch.allclose(b[0, 0, :, 2], torch.arange(1., 5.))
    m.load_state_dict(sub(g, 'attn.sd'))                # identical keys
    with torch.no_grad():
        y = m(a['query'], key=None, value=a['value'], query_pos=a['query_pos'], reference_points=a['ref'],
              spatial_shapes=torch.tensor([grid]), level_start_index=torch.tensor([0]))
    assert rel_err(y, a['out']) < 1e-6
    with pytest.raises(ValueError):
        m(a['query'], value=a['value'], reference_points=a['ref'][..., :2], spatial_shapes=[grid])
    with pytest.raises(ValueError):
        registry.build_attention(dict(type='VoxelCustomMSDeformableAttention', embed_dims=100, num_heads=8))


def test_decoder_host_logic(gold, oracle_sampler):
    g, grid = gold
    a, d = sub(g, 'attn'), sub(g, 'dec')
    dec = registry.build_transformer_layer_sequence(decoder_cfg()).eval()
    assert dec.layers[0].ffns[0].layers[0][0].out_features == 128       # deprecated kwarg wins (vocc.py:156)
    assert dec.layers[0].attentions[0].batch_first is False
    dec.load_state_dict(sub(g, 'dec.sd'))               # identical keys (mmcv MultiheadAttention: attn.in_proj_*)
    with torch.no_grad():
        hs, refs = dec(query=a['query'], key=None, value=a['value'], query_pos=a['query_pos'],
                       reference_points=d['ref'], reg_branches=_reg_branches(g), cls_branches=None,
                       spatial_shapes=[grid], level_start_index=[0])
    assert hs.shape == (2, 10, 2, 64) and refs.shape == (2, 2, 10, 3)
    assert rel_err(hs, d['hs']) < 2e-6 and rel_err(refs, d['refs']) < 1e-6
    # without box refinement the reference points pass through unchanged
    with torch.no_grad():
        _, refs0 = dec(query=a['query'], key=None, value=a['value'], query_pos=a['query_pos'],
                       reference_points=d['ref'], reg_branches=None, spatial_shapes=[grid], level_start_index=[0])
    assert torch.equal(refs0[1], d['ref'])


def test_temporal_self_attention_host_logic(gold, oracle_sampler):
    g, grid = gold
    t = sub(g, 'tsa')
    m = registry.build_attention(dict(type='VoxelTemporalSelfAttention', embed_dims=64, num_levels=1)).eval()
    assert m.sampling_offsets.bias.numel() == m.sampling_offsets.out_features == 2 * 8 * 1 * 4 * 3
    m.load_state_dict(sub(g, 'tsa.sd'))
    with torch.no_grad():
        y = m(t['query'], query_pos=t['query_pos'], reference_points=t['ref'], spatial_shapes=torch.tensor([grid]),
              level_start_index=torch.tensor([0]))
    assert rel_err(y, t['out']) < 1e-6


def test_transformer_forward_decodes(oracle_sampler, monkeypatch):
    """VoxelPerceptionTransformer.forward = get_voxel_features + the decoder half
    (M/voxel_transformer.py:246-301) against ver_ref.transformer_decode; the encoder (CUDA-only) is
    replaced by a fixed volume here."""
    C, grid, bs = 64, (4, 6, 6), 2
    Nv = grid[0] * grid[1] * grid[2]
    torch.manual_seed(3)
    tr = registry.build_transformer(dict(
        type='VoxelPerceptionTransformer', num_cams=6, embed_dims=C, decoder_on_bev=False,
        encoder=dict(type='VoxelFormerEncoder', num_layers=1, pc_range=[-6, -6, -1.5, 6, 6, 2],
                     transformerlayers=dict(
                         type='VoxelFormerLayer',
                         attn_cfgs=[dict(type='SpatialCrossAttention', pc_range=[-6, -6, -1.5, 6, 6, 2], embed_dims=C,
                                         deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                                   num_points=8, num_levels=1))],
                         ffn_cfgs=dict(type='FFN', embed_dims=C, feedforward_channels=1024, num_fcs=2,
                                       ffn_drop=0., act_cfg=dict(type='ReLU', inplace=True)),
                         feedforward_channels=2 * C, ffn_dropout=0.1,
                         operation_order=('cross_attn', 'norm', 'ffn', 'norm'))),
        decoder=decoder_cfg(C, 2))).eval()
    tr.init_weights()
    assert tr.reference_points.out_features == 3
    volume = torch.randn(bs, Nv, C)
    monkeypatch.setattr(tr, 'get_voxel_features', lambda *a, **k: volume)
    oqe = torch.randn(10, 2 * C)
    with torch.no_grad():
        vox, hs, init_ref, refs = tr(torch.zeros(6, bs, 196, C), torch.zeros(Nv, C), oqe, *grid)
        sd = dict(tr.state_dict())
        vox2, hs2, init2, refs2 = ver_ref.transformer_decode(sd, '', volume, oqe, *grid, num_layers=2)
    assert vox.shape == (Nv, bs, C) and torch.equal(vox, vox2)
    assert rel_err(init_ref, init2) < 1e-6 and rel_err(hs, hs2) < 2e-6 and rel_err(refs, refs2) < 1e-6


def test_head_with_decoder_state_dict_keys_and_detection_tail():
    """vocc.py default head (only_occ=False): the built module tree has exactly the reference's state_dict
    keys (recorded from the unmodified head by gen_golden.gen_head_detection), and the detection tail
    (cls / reg branches, box placement into pc_range, HEAD:583-611) reproduces the unmodified head's outputs."""
    g = load_golden('head_detection_c32.npz')
    grid = [int(x) for x in g['grid']]
    cfg = V.vocc_head_cfg(*grid, num_cams=6, embed_dims=32, only_occ=False, refine_occ=False,
                          occupancy_size=[2.0, 2.0, 0.5], occ_dims=16, num_layers=1, num_decoder_layers=2,
                          num_query=10)
    head = V.build_head(cfg).eval()
    head.init_weights()
    assert sorted(head.state_dict().keys()) == [str(k) for k in g['state_dict_keys']]
    assert head.cls_out_channels == 17 and head.code_size == 10
    missing, unexpected = head.load_state_dict(sub(g, 'sd'), strict=False)
    assert not unexpected
    c = {k: torch.from_numpy(v) for k, v in g.items() if not k.startswith(('sd.', 'state_dict'))}
    with torch.no_grad():
        cls, boxes, layouts = head._detection_tail(c['hs'], c['init_ref'], c['inter_refs'])
        cls2, boxes2 = ver_ref.detection_tail(sub(g, 'sd'), '', c['hs'], c['init_ref'], c['inter_refs'],
                                              head.pc_range)
    assert layouts is None
    assert rel_err(cls, c['all_cls_scores']) < 1e-6 and rel_err(boxes, c['all_bbox_preds']) < 1e-6
    assert rel_err(cls2, c['all_cls_scores']) < 1e-6 and rel_err(boxes2, c['all_bbox_preds']) < 1e-6
    # the shipped vocc.py tree: 6 decoder layers, 100 queries, 768 channels
    full = V.build_head(V.vocc_head_cfg())
    keys = set(full.state_dict().keys())
    for k in ('transformer.decoder.layers.5.attentions.0.attn.in_proj_weight',
              'transformer.decoder.layers.0.attentions.1.sampling_offsets.weight',
              'transformer.decoder.layers.3.ffns.0.layers.0.0.weight', 'transformer.reference_points.weight',
              'query_embedding.weight', 'cls_branches.5.6.bias', 'reg_branches.0.4.weight', 'layout_branches.2.0.weight'):
        assert k in keys, k
    sd = full.state_dict()
    assert sd['transformer.decoder.layers.0.attentions.1.sampling_offsets.weight'].shape == (8 * 1 * 4 * 3, 768)
    assert sd['transformer.decoder.layers.0.ffns.0.layers.0.0.weight'].shape == (1536, 768)
    assert sd['query_embedding.weight'].shape == (100, 1536) and sd['reg_branches.0.4.weight'].shape == (10, 768)
