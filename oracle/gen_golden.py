"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the
UNMODIFIED reference classes from /root/reference (through oracle/mmcv_shim.py).

Run in the build container (the reference does not exist on the GPU box):
    python -m oracle.gen_golden
The fixtures are committed; tests compare (a) the oracle restatement and (b) the
CUDA path against them.
"""
import json
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mmcv_shim, ver_ref          # noqa: E402
from vln_ver_b200 import synth                 # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
PC = synth.PC_RANGE


def sd_np(module):
    return {k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def perturb(module, seed):
    """zero-initialised offset/weight projections -> query dependent."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        if name.endswith('sampling_offsets.weight') or name.endswith('attention_weights.weight'):
            with torch.no_grad():
                p.add_(torch.randn(p.shape, generator=g) * 0.02)


# --------------------------------------------------------------------------- A5
def gen_msda():
    vtsa = mmcv_shim.import_reference('bevformer.modules.voxel_temporal_self_attention')
    ref3d = vtsa.voxel_multi_scale_deformable_attn_pytorch
    g = torch.Generator().manual_seed(11)
    cases = {}
    for name, (Bv, H, W, NH, Dh, Nq, NP) in {
            'small': (3, 14, 14, 8, 12, 37, 8),
            'dh96': (2, 14, 14, 8, 96, 19, 8),
            'rect': (2, 5, 9, 4, 8, 23, 4)}.items():
        value = torch.randn(Bv, H * W, NH, Dh, generator=g, dtype=torch.float64)
        loc = torch.rand(Bv, Nq, NH, 1, NP, 2, generator=g, dtype=torch.float64) * 1.6 - 0.3
        # exact-border and far-outside locations (zero padding, align_corners=False)
        loc[0, 0, 0, 0, 0] = torch.tensor([0.0, 0.0])
        loc[0, 0, 0, 0, 1] = torch.tensor([1.0, 1.0])
        loc[0, 0, 0, 0, 2] = torch.tensor([0.5 / W, 0.5 / H])          # pixel centre (0,0)
        loc[0, 0, 0, 0, 3] = torch.tensor([-3.0, 7.0])
        loc[0, 1, 0, 0, 0] = torch.tensor([(W - 0.5) / W, (H - 0.5) / H])  # last pixel centre
        w = torch.rand(Bv, Nq, NH, 1, NP, generator=g, dtype=torch.float64)
        w = w / w.sum(-1, keepdim=True)
        gout = torch.randn(Bv, Nq, NH * Dh, generator=g, dtype=torch.float64)
        value.requires_grad_(True); loc.requires_grad_(True); w.requires_grad_(True)
        # reference-owned 3-D sampler at depth 1 (z = 0.5 -> grid z = 0)
        loc3 = torch.cat([loc, torch.full_like(loc[..., :1], 0.5)], -1)
        out3 = ref3d(value, [(1, H, W)], loc3, w)
        out2 = ver_ref.multi_scale_deformable_attn_pytorch(
            value, torch.tensor([[H, W]]), loc, w)
        assert (out3 - out2).abs().max().item() < 1e-12, (out3 - out2).abs().max()
        gv, gl, gw = torch.autograd.grad(out3, (value, loc, w), gout)
        cases.update({
            f'{name}.value': value.detach().numpy().astype(np.float32),
            f'{name}.shape': np.array([H, W], np.int64),
            f'{name}.loc': loc.detach().numpy().astype(np.float32),
            f'{name}.w': w.detach().numpy().astype(np.float32),
            f'{name}.gout': gout.numpy().astype(np.float32),
        })
        # golden outputs recomputed in fp64 from the fp32-rounded inputs so the
        # fixture is self-consistent
        v32 = torch.from_numpy(cases[f'{name}.value']).double().requires_grad_(True)
        l32 = torch.from_numpy(cases[f'{name}.loc']).double().requires_grad_(True)
        w32 = torch.from_numpy(cases[f'{name}.w']).double().requires_grad_(True)
        g32 = torch.from_numpy(cases[f'{name}.gout']).double()
        l3 = torch.cat([l32, torch.full_like(l32[..., :1], 0.5)], -1)
        o = ref3d(v32, [(1, H, W)], l3, w32)
        gv, gl, gw = torch.autograd.grad(o, (v32, l32, w32), g32)
        cases.update({f'{name}.out': o.detach().numpy(), f'{name}.gvalue': gv.numpy(),
                      f'{name}.gloc': gl.numpy(), f'{name}.gw': gw.numpy()})
    np.savez_compressed(os.path.join(OUT, 'msda_cases.npz'), **cases)
    print('msda_cases ok')


# --------------------------------------------------------------------------- A1 / A2
def run_unmodified_point_sampling(enc_mod, bev_z, bev_h, bev_w, l2i6, shift, num_points_in_voxel=4):
    """Calls the UNMODIFIED VoxelFormerEncoder.point_sampling (reads the literal
    'path to/...' files, M/voxel_encoder.py:122,133) from a temp cwd."""
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, 'path to', 'camera_parameters', 'world2pixel'))
        scan, vp = 'scanA', 'vp0'
        data = {f'{vp}_i1_{d}': l2i6[d].astype(np.float64).tolist() for d in range(6)}
        with open(os.path.join(td, 'path to', 'camera_parameters', 'world2pixel', scan + '.json'), 'w') as f:
            json.dump(data, f)
        with open(os.path.join(td, 'path to', 'scanvp2cord.pkl'), 'wb') as f:
            pickle.dump({scan + '_' + vp: [float(x) for x in shift]}, f)
        os.chdir(td)
        try:
            Enc = enc_mod.VoxelFormerEncoder
            ref_3d = Enc.get_reference_points(bev_z, bev_h, bev_w, num_points_in_voxel, dim='3d',
                                              bs=1, device='cpu', dtype=torch.float32)
            dummy = type('E', (), {})()
            rpc, mask = Enc.point_sampling(dummy, ref_3d, PC, [dict(sample_idx=f'{scan}_{vp}')])
        finally:
            os.chdir(cwd)
    return ref_3d, rpc, mask


def gen_point_sampling(enc_mod):
    out = {}
    for tag, grid, seed in [('g4x15x15', (4, 15, 15), 21), ('g8x20x20', (8, 20, 20), 22)]:
        l2i, sh = synth.make_rig(1, 6, grid, seed=seed)
        ref_3d, rpc, mask = run_unmodified_point_sampling(enc_mod, *grid, l2i[0], sh[0])
        # restatement must agree bit for bit
        r2 = ver_ref.get_reference_points_3d(*grid)
        assert torch.equal(r2, ref_3d)
        rpc2, mask2 = ver_ref.point_sampling(r2, PC, torch.from_numpy(l2i[0]), torch.from_numpy(sh[0]))
        assert torch.equal(rpc, rpc2) and torch.equal(mask, mask2)
        idx = ver_ref.visible_indexes(mask)
        out.update({f'{tag}.lidar2img': l2i, f'{tag}.originshift': sh,
                    f'{tag}.ref_3d': ref_3d.numpy(), f'{tag}.rpc': rpc.numpy(),
                    f'{tag}.mask': mask.numpy(),
                    f'{tag}.index_len': np.array([len(i) for i in idx], np.int64),
                    f'{tag}.index_cat': torch.cat(idx).numpy().astype(np.int64)})
    np.savez_compressed(os.path.join(OUT, 'point_sampling_6cam.npz'), **out)
    print('point_sampling ok')


# --------------------------------------------------------------------------- A3 / A4
def gen_sca(sca_mod):
    out = {}
    for tag, ncam, grid, C, seed in [('c6', 6, (4, 15, 15), 256, 31), ('c18', 18, (4, 10, 10), 256, 32)]:
        torch.manual_seed(seed)
        m = sca_mod.SpatialCrossAttention(
            embed_dims=C, num_cams=ncam, pc_range=PC, dropout=0.1, batch_first=True,
            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                      num_points=8, num_levels=1)).eval()
        perturb(m, seed + 100)
        l2i, sh = synth.make_rig(1, ncam, grid, seed=seed)
        rpc, mask = ver_ref.point_sampling_batched(*grid, PC, torch.from_numpy(l2i), torch.from_numpy(sh))
        Nq = grid[0] * grid[1] * grid[2]
        query = torch.randn(1, Nq, C)
        value = torch.randn(ncam, 196, 1, C) * 0.5
        ss = torch.tensor([[14, 14]])
        with torch.no_grad():
            y = m(query, value, value, reference_points_cam=rpc, bev_mask=mask,
                  spatial_shapes=ss, level_start_index=torch.tensor([0]))
            sd = {k: v for k, v in m.state_dict().items()}
            y2 = ver_ref.sca_forward(sd, '', query, value, rpc, mask, ss)
        assert torch.allclose(y, y2, atol=1e-6), (y - y2).abs().max()
        out.update({f'{tag}.query': query.numpy(), f'{tag}.value': value.numpy(),
                    f'{tag}.lidar2img': l2i, f'{tag}.originshift': sh,
                    f'{tag}.grid': np.array(grid), f'{tag}.out': y.numpy()})
        out.update({f'{tag}.sd.{k}': v for k, v in sd_np(m).items()})
    np.savez_compressed(os.path.join(OUT, 'sca.npz'), **out)
    print('sca ok')


# --------------------------------------------------------------------------- A6 / A7
def encoder_cfg(C, ffn, num_layers=3):
    return dict(
        type='VoxelFormerEncoder', num_layers=num_layers, pc_range=PC, num_points_in_voxel=4,
        return_intermediate=False,
        transformerlayers=dict(
            type='VoxelFormerLayer',
            attn_cfgs=[dict(type='SpatialCrossAttention', pc_range=PC,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_points=8, num_levels=1),
                            embed_dims=C)],
            ffn_cfgs=dict(type='FFN', embed_dims=C, feedforward_channels=1024, num_fcs=2,
                          ffn_drop=0., act_cfg=dict(type='ReLU', inplace=True)),
            feedforward_channels=ffn, ffn_dropout=0.1,
            operation_order=('cross_attn', 'norm', 'ffn', 'norm')))


def run_unmodified_encoder(enc, bev_query, value, grid, l2i6, shift, bev_pos):
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, 'path to', 'camera_parameters', 'world2pixel'))
        data = {f'vp0_i1_{d}': l2i6[d].astype(np.float64).tolist() for d in range(6)}
        with open(os.path.join(td, 'path to', 'camera_parameters', 'world2pixel', 'scanA.json'), 'w') as f:
            json.dump(data, f)
        with open(os.path.join(td, 'path to', 'scanvp2cord.pkl'), 'wb') as f:
            pickle.dump({'scanA_vp0': [float(x) for x in shift]}, f)
        os.chdir(td)
        try:
            with torch.no_grad():
                y = enc(bev_query, value, value, bev_z=grid[0], bev_h=grid[1], bev_w=grid[2],
                        bev_pos=bev_pos, spatial_shapes=torch.tensor([[14, 14]]),
                        level_start_index=torch.tensor([0]), prev_bev=None,
                        shift=bev_query.new_tensor([[0., 0., 0.]]),
                        img_metas=[dict(sample_idx='scanA_vp0')])
        finally:
            os.chdir(cwd)
    return y


def gen_encoder():
    C, grid, seed = 256, (4, 15, 15), 41
    torch.manual_seed(seed)
    enc = mmcv_shim.build_transformer_layer_sequence(encoder_cfg(C, C, num_layers=2)).eval()
    for p in enc.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
    for m in enc.modules():
        if type(m).__name__ == 'MSDeformableAttention3D':
            m.init_weights()
    perturb(enc, seed + 100)
    Nq = grid[0] * grid[1] * grid[2]
    l2i, sh = synth.make_rig(1, 6, grid, seed=seed)
    bev_query = torch.randn(Nq, 1, C)
    value = torch.randn(6, 196, 1, C) * 0.5
    bev_pos = torch.zeros(Nq, 1, C)
    y = run_unmodified_encoder(enc, bev_query, value, grid, l2i[0], sh[0], bev_pos)
    sd = dict(enc.state_dict())
    y2 = ver_ref.encoder_forward(sd, '', bev_query, value, *grid, PC, torch.from_numpy(l2i),
                                 torch.from_numpy(sh), torch.tensor([[14, 14]]), num_layers=2)
    assert torch.allclose(y, y2, atol=2e-6), (y - y2).abs().max()
    out = {'bev_query': bev_query.numpy(), 'value': value.numpy(), 'lidar2img': l2i,
           'originshift': sh, 'grid': np.array(grid), 'out': y.numpy()}
    out.update({f'sd.{k}': v for k, v in sd_np(enc).items()})
    np.savez_compressed(os.path.join(OUT, 'encoder_6cam_c256.npz'), **out)
    print('encoder ok')


# --------------------------------------------------------------------------- A8
def gen_transformer():
    """Unmodified VoxelPerceptionTransformer.get_voxel_features at its literal
    (6, 1, 14, 14, 768) shape.  Weights are NOT stored (44 MB): they are regenerated
    from `manual_seed(seed)` + the reference's own init_weights + perturb(); the
    fixture holds inputs' seeds and a row-subsampled output."""
    vt = mmcv_shim.import_reference('bevformer.modules.voxel_transformer')
    C, grid, seed = 768, (4, 15, 15), 51
    torch.manual_seed(seed)
    tr = vt.VoxelPerceptionTransformer(
        num_cams=6, embed_dims=C, rotate_prev_bev=True, use_shift=True, use_can_bus=True,
        decoder_on_bev=False, encoder=encoder_cfg(C, 2 * C), decoder=None).eval()
    tr.init_weights()
    perturb(tr, seed + 100)
    Nq = grid[0] * grid[1] * grid[2]
    l2i, sh = synth.make_rig(1, 6, grid, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    feats = torch.randn(6, 1, 196, C, generator=g) * 0.5
    bev_queries = torch.randn(Nq, C, generator=g)
    bev_pos = torch.zeros(1, C, *grid)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, 'path to', 'camera_parameters', 'world2pixel'))
        data = {f'vp0_i1_{d}': l2i[0, d].astype(np.float64).tolist() for d in range(6)}
        with open(os.path.join(td, 'path to', 'camera_parameters', 'world2pixel', 'scanA.json'), 'w') as f:
            json.dump(data, f)
        with open(os.path.join(td, 'path to', 'scanvp2cord.pkl'), 'wb') as f:
            pickle.dump({'scanA_vp0': [float(x) for x in sh[0]]}, f)
        os.chdir(td)
        try:
            with torch.no_grad():
                y = tr.get_voxel_features(feats, bev_queries, *grid, bev_pos=bev_pos,
                                          img_metas=[dict(sample_idx='scanA_vp0')])
        finally:
            os.chdir(cwd)
    sd = dict(tr.state_dict())
    with torch.no_grad():
        y2 = ver_ref.get_voxel_features(sd, '', feats, bev_queries, *grid, PC,
                                        torch.from_numpy(l2i), torch.from_numpy(sh))
    assert torch.allclose(y, y2, atol=5e-6), (y - y2).abs().max()
    rows = np.arange(0, Nq, 7)
    np.savez_compressed(
        os.path.join(OUT, 'transformer_6cam_c768.npz'), seed=np.array(seed), grid=np.array(grid),
        lidar2img=l2i, originshift=sh, rows=rows, out_rows=y[0, rows].numpy(),
        out_abs_sum=np.array(y.double().abs().sum().item()))
    print('transformer ok')
    return tr, sd


# --------------------------------------------------------------------------- A9 / A10 / A12
def gen_head():
    pe = mmcv_shim.import_reference('bevformer.modules.voxel_positional_embedding')
    mmcv_shim.import_reference('bevformer.modules.voxel_transformer')
    hd = mmcv_shim.import_reference('bevformer.dense_heads.voxelformer_occupancy_head')
    C, grid, seed = 32, (4, 6, 6), 61
    occ_size = [12.0 / grid[2], 12.0 / grid[1], 3.5 / grid[0]]      # bev_z == occ_zdim branch
    out = {}
    for tag, osz in [('pervoxel', occ_size), ('column', [2.0, 2.0, 0.5])]:
        torch.manual_seed(seed)
        head = hd.VoxelFormerOccupancyHead(
            bev_h=grid[1], bev_w=grid[2], bev_z=grid[0], num_query=10, num_classes=17, in_channels=C,
            sync_cls_avg_factor=True, with_box_refine=True, as_two_stage=False,
            point_cloud_range=PC, occupancy_size=osz, occ_dims=16, occupancy_classes=16,
            only_occ=True, only_det=False, refine_occ=False,
            transformer=mmcv_shim.ConfigDict(
                type='VoxelPerceptionTransformer', num_cams=6, embed_dims=C,
                encoder=encoder_cfg(C, 2 * C, num_layers=1), decoder=None),
            bbox_coder=dict(type='NMSFreeCoder', pc_range=PC, max_num=50, num_classes=17),
            positional_encoding=dict(type='VoxelLearnedPositionalEncoding', num_feats=C // 2,
                                     row_num_embed=grid[1], col_num_embed=grid[2], z_num_embed=grid[0]),
            loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
            loss_bbox=dict(type='L1Loss', loss_weight=0.25),
            loss_iou=dict(type='GIoULoss', loss_weight=0.0),
            loss_occupancy=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25,
                                loss_weight=1.0)).eval()
        Nq = grid[0] * grid[1] * grid[2]
        bev_embed = torch.randn(1, Nq, C)

        class _Stub(torch.nn.Module):          # the head only consumes get_voxel_features()
            decoder = None

            def get_voxel_features(self, *a, **k):
                return bev_embed
        head.transformer = _Stub()
        with torch.no_grad():
            outs = head(torch.zeros(6, 1, 196, C), [dict(sample_idx='scanA_vp0')])
            pos = head.positional_encoding(torch.zeros(1, *grid))
            logits = outs['occupancy_preds']
            # make some voxels occupied for the decode test
            logits = logits + torch.randn_like(logits) * 3.0
            dec = head.get_occupancy_prediction(dict(occupancy_preds=logits.clone(), flow_preds=None))
        sd = {k: v for k, v in head.state_dict().items()}
        y2 = ver_ref.occ_head(sd, '', bev_embed, *grid, head.occ_xdim, head.occ_ydim, head.occ_zdim,
                              occ_dims=16, refine_occ=False, only_occ=True)
        assert torch.allclose(outs['occupancy_preds'], y2, atol=1e-6)
        pos2 = ver_ref.positional_encoding(sd, 'positional_encoding.', 1, *grid)
        assert torch.equal(pos, pos2)
        dec2 = ver_ref.get_occupancy_prediction(logits)
        assert torch.equal(dec['occupancy_preds'], dec2)
        keep = {k: v.numpy() for k, v in sd.items()
                if k.startswith(('occ_', 'positional_encoding', 'voxel_embedding'))}
        out.update({f'{tag}.bev_embed': bev_embed.numpy(), f'{tag}.grid': np.array(grid),
                    f'{tag}.occ_dims3': np.array([head.occ_xdim, head.occ_ydim, head.occ_zdim]),
                    f'{tag}.occupancy_preds': outs['occupancy_preds'].numpy(),
                    f'{tag}.pos': pos.numpy(), f'{tag}.decode_logits': logits.numpy(),
                    f'{tag}.decode': dec['occupancy_preds'].numpy()})
        out.update({f'{tag}.sd.{k}': v for k, v in keep.items()})
    np.savez_compressed(os.path.join(OUT, 'head.npz'), **out)
    print('head ok')


# --------------------------------------------------------------------------- N2 / N3
DEC_C, DEC_GRID, DEC_NQ, DEC_BS, DEC_LAYERS = 64, (4, 6, 6), 10, 2, 2


def decoder_cfg(C, num_layers):
    """vocc.py:137-158 at a small width."""
    return dict(
        type='VoxelDetectionTransformerDecoder', num_layers=num_layers, return_intermediate=True,
        transformerlayers=dict(
            type='DetrTransformerDecoderLayer',
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='VoxelCustomMSDeformableAttention', embed_dims=C, num_levels=1)],
            ffn_cfgs=dict(type='FFN', embed_dims=C, feedforward_channels=1024, num_fcs=2, ffn_drop=0.,
                          act_cfg=dict(type='ReLU', inplace=True)),
            feedforward_channels=2 * C, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))


def randomize(module, seed):
    """every weight matrix xavier-uniform, every bias / 1-D parameter small random: nothing on the
    decoder path may stay at a zero init, or the fixture would not exercise it."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() > 1:
                bound = (6.0 / (p.shape[0] + p.shape[1])) ** 0.5
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            elif name.endswith('sampling_offsets.bias') or 'norms' in name and name.endswith('weight'):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)


def gen_msda3d():
    vtsa = mmcv_shim.import_reference('bevformer.modules.voxel_temporal_self_attention')
    ref3d = vtsa.voxel_multi_scale_deformable_attn_pytorch
    g = torch.Generator().manual_seed(71)
    out = {}
    for name, (Bv, shapes, NH, Dh, Nq, NP) in {
            'small': (3, [(3, 5, 7)], 4, 8, 13, 4),
            'dh96': (2, [(4, 6, 6)], 8, 96, 9, 4),
            'two_level': (2, [(3, 4, 5), (2, 2, 3)], 2, 40, 7, 3)}.items():
        NL = len(shapes)
        S = sum(d * h * w for d, h, w in shapes)
        D, H, W = shapes[0]
        value = torch.randn(Bv, S, NH, Dh, generator=g)
        loc = torch.rand(Bv, Nq, NH, NL, NP, 3, generator=g) * 1.6 - 0.3
        loc[0, 0, 0, 0, 0] = torch.tensor([0.0, 0.0, 0.0])                 # corner of the padded volume
        loc[0, 0, 0, 0, 1] = torch.tensor([1.0, 1.0, 1.0])
        loc[0, 0, 0, 0, 2] = torch.tensor([0.5 / W, 0.5 / H, 0.5 / D])     # centre of voxel (0,0,0)
        loc[0, 1, 0, 0, 0] = torch.tensor([-3.0, 7.0, 0.5])                # far outside
        w = torch.rand(Bv, Nq, NH, NL, NP, generator=g)
        w = w / w.sum((-1, -2), keepdim=True)
        gout = torch.randn(Bv, Nq, NH * Dh, generator=g)
        v64 = value.double().requires_grad_(True)
        l64 = loc.double().requires_grad_(True)
        w64 = w.double().requires_grad_(True)
        ss = torch.tensor(shapes)
        y = ref3d(v64, ss, l64, w64)
        gv, gl, gw = torch.autograd.grad(y, (v64, l64, w64), gout.double())
        v2 = value.double().requires_grad_(True)
        y2 = ver_ref.voxel_multi_scale_deformable_attn_pytorch(v2, ss, l64, w64)
        assert torch.equal(y, y2)
        out.update({f'{name}.value': value.numpy(), f'{name}.loc': loc.numpy(), f'{name}.w': w.numpy(),
                    f'{name}.gout': gout.numpy(), f'{name}.shapes': np.array(shapes),
                    f'{name}.out': y.detach().numpy(), f'{name}.gvalue': gv.numpy(),
                    f'{name}.gloc': gl.numpy(), f'{name}.gw': gw.numpy()})
    np.savez_compressed(os.path.join(OUT, 'msda3d_cases.npz'), **out)
    print('msda3d ok')


def gen_decoder():
    """Unmodified VoxelCustomMSDeformableAttention, VoxelDetectionTransformerDecoder (on the shim's
    MultiheadAttention + a DetrTransformerDecoderLayer that subclasses the reference's own copy of
    BaseTransformerLayer) and VoxelTemporalSelfAttention (with a sampling_offsets bias of the Linear's
    own size, see ver_ref.temporal_self_attention_forward)."""
    vd = mmcv_shim.import_reference('bevformer.modules.voxel_decoder')
    vtsa = mmcv_shim.import_reference('bevformer.modules.voxel_temporal_self_attention')
    mmcv_shim.register_detr_decoder_layer()
    C, grid, nq, bs, L = DEC_C, DEC_GRID, DEC_NQ, DEC_BS, DEC_LAYERS
    Nv = grid[0] * grid[1] * grid[2]
    ss = torch.tensor([list(grid)])
    g = torch.Generator().manual_seed(81)
    out = {}

    # -- the attention module alone
    attn = vd.VoxelCustomMSDeformableAttention(embed_dims=C, num_levels=1, batch_first=False).eval()
    randomize(attn, 82)
    query = torch.randn(nq, bs, C, generator=g)
    query_pos = torch.randn(nq, bs, C, generator=g) * 0.5
    value = torch.randn(Nv, bs, C, generator=g)
    ref = torch.rand(bs, nq, 1, 3, generator=g)
    with torch.no_grad():
        y = attn(query, key=None, value=value, query_pos=query_pos, reference_points=ref, spatial_shapes=ss,
                 level_start_index=torch.tensor([0]))
        y2 = ver_ref.voxel_custom_msda_forward(dict(attn.state_dict()), '', query, value, ref, ss,
                                               query_pos=query_pos)
    assert torch.allclose(y, y2, atol=1e-6), (y - y2).abs().max()
    out.update({'attn.query': query.numpy(), 'attn.query_pos': query_pos.numpy(), 'attn.value': value.numpy(),
                'attn.ref': ref.numpy(), 'attn.out': y.numpy(), 'grid': np.array(grid)})
    out.update({f'attn.sd.{k}': v for k, v in sd_np(attn).items()})

    # -- the decoder with box refinement
    dec = mmcv_shim.build_transformer_layer_sequence(decoder_cfg(C, L)).eval()
    randomize(dec, 83)
    regs = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(C, C), torch.nn.ReLU(),
                                                    torch.nn.Linear(C, 10)) for _ in range(L)]).eval()
    randomize(regs, 84)
    ref3 = torch.rand(bs, nq, 3, generator=g)
    with torch.no_grad():
        hs, refs = dec(query=query, key=None, value=value, query_pos=query_pos, reference_points=ref3,
                       reg_branches=regs, cls_branches=None, spatial_shapes=ss,
                       level_start_index=torch.tensor([0]))
        hs2, refs2 = ver_ref.decoder_forward(dict(dec.state_dict()), '', query, value, query_pos, ref3, ss,
                                             num_layers=L, reg_branches=regs)
    assert torch.allclose(hs, hs2, atol=2e-6), (hs - hs2).abs().max()
    assert torch.allclose(refs, refs2, atol=1e-6)
    out.update({'dec.ref': ref3.numpy(), 'dec.hs': hs.numpy(), 'dec.refs': refs.numpy()})
    out.update({f'dec.sd.{k}': v for k, v in sd_np(dec).items()})
    out.update({f'reg.sd.{k}': v for k, v in sd_np(regs).items()})

    # -- temporal self-attention (N3)
    tsa = vtsa.VoxelTemporalSelfAttention(embed_dims=C, num_levels=1, batch_first=True).eval()
    tsa.sampling_offsets.bias.data = torch.zeros(tsa.sampling_offsets.out_features)
    randomize(tsa, 85)
    vq = torch.randn(bs, Nv, C, generator=g)
    vpos = torch.randn(bs, Nv, C, generator=g) * 0.5
    zs, ys, xs = torch.meshgrid(*[(torch.arange(n) + 0.5) / n for n in grid], indexing='ij')
    ref_vox = torch.stack((xs, ys, zs), -1).view(1, Nv, 1, 3).repeat(bs * 2, 1, 1, 1)
    with torch.no_grad():
        yt = tsa(vq, query_pos=vpos, reference_points=ref_vox, spatial_shapes=ss,
                 level_start_index=torch.tensor([0]))
        yt2 = ver_ref.temporal_self_attention_forward(dict(tsa.state_dict()), '', vq, ref_vox, ss, query_pos=vpos)
    assert torch.allclose(yt, yt2, atol=1e-6), (yt - yt2).abs().max()
    out.update({'tsa.query': vq.numpy(), 'tsa.query_pos': vpos.numpy(), 'tsa.ref': ref_vox.numpy(),
                'tsa.out': yt.numpy()})
    out.update({f'tsa.sd.{k}': v for k, v in sd_np(tsa).items()})
    np.savez_compressed(os.path.join(OUT, 'decoder_c64.npz'), **out)
    print('decoder ok')


def gen_head_detection():
    """Unmodified VoxelFormerOccupancyHead.forward, default branch (only_occ=False, HEAD:537-625), with the
    transformer replaced by fixed outputs: pins the detection tail (cls / reg branches + box placement) and
    the occupancy tail fed from the (Nq, bs, C) layout."""
    for mod in ('voxel_positional_embedding', 'spatial_cross_attention', 'voxel_encoder', 'voxel_decoder',
                'voxel_transformer'):
        mmcv_shim.import_reference('bevformer.modules.' + mod)
    hd = mmcv_shim.import_reference('bevformer.dense_heads.voxelformer_occupancy_head')
    mmcv_shim.register_detr_decoder_layer()
    C, grid, seed, nq, L = 32, (4, 6, 6), 91, 10, 2
    torch.manual_seed(seed)
    head = hd.VoxelFormerOccupancyHead(
        bev_h=grid[1], bev_w=grid[2], bev_z=grid[0], num_query=nq, num_classes=17, in_channels=C,
        sync_cls_avg_factor=True, with_box_refine=True, as_two_stage=False, point_cloud_range=PC,
        occupancy_size=[2.0, 2.0, 0.5], occ_dims=16, occupancy_classes=16, only_occ=False, only_det=False,
        refine_occ=False,
        transformer=mmcv_shim.ConfigDict(
            type='VoxelPerceptionTransformer', num_cams=6, embed_dims=C, decoder_on_bev=False,
            encoder=encoder_cfg(C, 2 * C, num_layers=1), decoder=decoder_cfg(C, L)),
        bbox_coder=dict(type='NMSFreeCoder', pc_range=PC, max_num=50, num_classes=17),
        positional_encoding=dict(type='VoxelLearnedPositionalEncoding', num_feats=C // 2,
                                 row_num_embed=grid[1], col_num_embed=grid[2], z_num_embed=grid[0]),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
        loss_bbox=dict(type='L1Loss', loss_weight=0.25), loss_iou=dict(type='GIoULoss', loss_weight=0.0),
        loss_occupancy=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)).eval()
    randomize(head, seed + 1)
    Nq = grid[0] * grid[1] * grid[2]
    g = torch.Generator().manual_seed(seed + 2)
    bev_embed = torch.randn(Nq, 1, C, generator=g)
    hs = torch.randn(L, nq, 1, C, generator=g)
    init_ref = torch.rand(1, nq, 3, generator=g)
    inter_refs = torch.rand(L, 1, nq, 3, generator=g)
    inter_refs[0, 0, 0] = torch.tensor([0.0, 1.0, 0.5])       # clamped by inverse_sigmoid's eps

    class _Stub(torch.nn.Module):
        def __init__(self, decoder):
            super().__init__()
            self.decoder = decoder

        def forward(self, *a, **k):
            return bev_embed, hs, init_ref, inter_refs
    all_keys = sorted(head.state_dict().keys())               # with the real transformer tree
    head.transformer = _Stub(head.transformer.decoder)
    with torch.no_grad():
        outs = head(torch.zeros(6, 1, 196, C), [dict(sample_idx='scanA_vp0')])
    sd = dict(head.state_dict())
    cls2, box2 = ver_ref.detection_tail(sd, '', hs, init_ref, inter_refs, PC)
    assert torch.allclose(outs['all_cls_scores'], cls2, atol=1e-6)
    assert torch.allclose(outs['all_bbox_preds'], box2, atol=1e-6)
    occ2 = ver_ref.occ_head(sd, '', bev_embed.permute(1, 0, 2), *grid, head.occ_xdim, head.occ_ydim, head.occ_zdim,
                            occ_dims=16, refine_occ=False, only_occ=False)
    assert torch.allclose(outs['occupancy_preds'], occ2, atol=1e-6)
    keep = {k: v.numpy() for k, v in sd.items()
            if k.startswith(('cls_branches', 'reg_branches', 'occ_', 'query_embedding'))}
    out = {'bev_embed': bev_embed.numpy(), 'hs': hs.numpy(), 'init_ref': init_ref.numpy(),
           'inter_refs': inter_refs.numpy(), 'grid': np.array(grid), 'all_cls_scores': outs['all_cls_scores'].numpy(),
           'all_bbox_preds': outs['all_bbox_preds'].numpy(), 'occupancy_preds': outs['occupancy_preds'].numpy(),
           'state_dict_keys': np.array(all_keys)}
    out.update({f'sd.{k}': v for k, v in keep.items()})
    np.savez_compressed(os.path.join(OUT, 'head_detection_c32.npz'), **out)
    print('head detection ok')


def refine_upsample_weights(i):
    """Weights of up_sample layer i of the refine_occ fixture (133 M parameters: regenerated from a seed by the
    generator AND by the tests instead of being stored)."""
    g = torch.Generator().manual_seed(7000 + i)
    w = torch.randn(768, 768, 3, 5, 5, generator=g) * 0.01
    b = torch.randn(768, generator=g) * 0.1
    return w, b


def gen_head_refine():
    """Unmodified VoxelFormerOccupancyHead.forward, default branch with refine_occ=True (HEAD:551-580): the raw
    `.view` reinterpretations (:558, :564), the three ConvTranspose3d(768, 768) of `up_sample` (:254-258, the
    channel count is hard-coded, so C = 768) and the column-wise occ_proj (bev_z != occ_zdim, :570-576) on a tiny
    2 x 3 x 3 grid -> 7 x 24 x 24 occupancy cells.  Pins the oracle's refine_occ tail (it was only
    compared with the restatement)."""
    for mod in ('voxel_positional_embedding', 'spatial_cross_attention', 'voxel_encoder', 'voxel_decoder',
                'voxel_transformer'):
        mmcv_shim.import_reference('bevformer.modules.' + mod)
    hd = mmcv_shim.import_reference('bevformer.dense_heads.voxelformer_occupancy_head')
    mmcv_shim.register_detr_decoder_layer()
    C, grid, seed, nq, L = 768, (2, 3, 3), 131, 4, 1
    torch.manual_seed(seed)
    head = hd.VoxelFormerOccupancyHead(
        bev_h=grid[1], bev_w=grid[2], bev_z=grid[0], num_query=nq, num_classes=17, in_channels=C,
        sync_cls_avg_factor=True, with_box_refine=True, as_two_stage=False, point_cloud_range=PC,
        occupancy_size=[0.5, 0.5, 0.5], occ_dims=16, occupancy_classes=16, only_occ=False, only_det=False,
        refine_occ=True,
        transformer=mmcv_shim.ConfigDict(
            type='VoxelPerceptionTransformer', num_cams=6, embed_dims=C, decoder_on_bev=False,
            encoder=encoder_cfg(C, 2 * C, num_layers=1), decoder=decoder_cfg(C, L)),
        bbox_coder=dict(type='NMSFreeCoder', pc_range=PC, max_num=50, num_classes=17),
        positional_encoding=dict(type='VoxelLearnedPositionalEncoding', num_feats=C // 2,
                                 row_num_embed=grid[1], col_num_embed=grid[2], z_num_embed=grid[0]),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
        loss_bbox=dict(type='L1Loss', loss_weight=0.25), loss_iou=dict(type='GIoULoss', loss_weight=0.0),
        loss_occupancy=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)).eval()
    assert (head.occ_xdim, head.occ_ydim, head.occ_zdim) == (24, 24, 7)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for i, conv in enumerate(head.up_sample):
            w, b = refine_upsample_weights(i)
            conv.weight.copy_(w)
            conv.bias.copy_(b)
        for n, p in head.named_parameters():
            if n.startswith(('occ_proj', 'occ_branches')):
                p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.2))
    Nq = grid[0] * grid[1] * grid[2]
    bev_embed = torch.randn(Nq, 1, C, generator=g)
    hs = torch.randn(L, nq, 1, C, generator=g)
    init_ref = torch.rand(1, nq, 3, generator=g)
    inter_refs = torch.rand(L, 1, nq, 3, generator=g)

    class _Stub(torch.nn.Module):
        def __init__(self, decoder):
            super().__init__()
            self.decoder = decoder

        def forward(self, *a, **k):
            return bev_embed, hs, init_ref, inter_refs
    head.transformer = _Stub(head.transformer.decoder)
    with torch.no_grad():
        outs = head(torch.zeros(6, 1, 196, C), [dict(sample_idx='scanA_vp0')])
    sd = dict(head.state_dict())
    occ2 = ver_ref.occ_head(sd, '', bev_embed.permute(1, 0, 2), *grid, head.occ_xdim, head.occ_ydim, head.occ_zdim,
                            occ_dims=16, refine_occ=True, only_occ=False)
    err = (outs['occupancy_preds'] - occ2).abs().max().item() / outs['occupancy_preds'].abs().max().item()
    assert err < 1e-6, err
    out = {'bev_embed': bev_embed.numpy(), 'grid': np.array(grid), 'occ_dims3': np.array([24, 24, 7]),
           'occupancy_preds': outs['occupancy_preds'].numpy()}
    out.update({f'sd.{k}': v.numpy() for k, v in sd.items() if k.startswith(('occ_proj', 'occ_branches'))})
    np.savez_compressed(os.path.join(OUT, 'head_refine_c768.npz'), **out)
    print('head refine ok, restatement vs unmodified head', err)


def main():
    assert mmcv_shim.reference_available(), 'needs /root/reference'
    os.makedirs(OUT, exist_ok=True)
    mmcv_shim.install()
    sca_mod = mmcv_shim.import_reference('bevformer.modules.spatial_cross_attention')
    enc_mod = mmcv_shim.import_reference('bevformer.modules.voxel_encoder')
    gen_msda()
    gen_point_sampling(enc_mod)
    gen_sca(sca_mod)
    gen_encoder()
    gen_transformer()
    gen_head()
    gen_msda3d()
    gen_decoder()
    gen_head_detection()
    gen_head_refine()


if __name__ == '__main__':
    main()
