"""GPU tier (B200) for SURVEY.md 8(f) rows N2 / N3: the 3-D deformable sampler kernels
(ver_msda3d_forward/backward through the C ABI), the detection decoder and the temporal self-attention
module, against the golden vectors from the unmodified reference and the CPU oracle on seeded inputs.
Tolerances: max-norm relative error <= 1e-5 fp32, <= 1e-3 fp16 storage (BASELINE.json north_star)."""
import pytest
import torch

import vln_ver_b200 as V
from oracle import ver_ref
from vln_ver_b200 import ops, registry
from conftest import load_golden, rel_err, sub
from test_decoder_host import _reg_branches, decoder_cfg

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = {torch.float32: 1e-5, torch.float16: 1e-3}


def cuda(t):
    return t.to(DEV)


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


# ------------------------------------------------------------------ the sampler (operator level)
@pytest.mark.parametrize('name', ['small', 'dh96', 'two_level'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_msda3d_golden(name, dtype):
    c = sub(load_golden('msda3d_cases.npz'), name)
    shapes = c['shapes'].tolist()
    v = cuda(c['value']).to(dtype).requires_grad_(True)
    l = cuda(c['loc']).requires_grad_(True)
    w = cuda(c['w']).requires_grad_(True)
    n0 = V.launch_count()
    out = ops.voxel_multi_scale_deformable_attn(v, shapes, l, w)
    tol = TOL[dtype]
    if dtype == torch.float16:      # reference recomputed from the fp16-rounded operands: kernel error only
        v64 = v.detach().double().cpu().requires_grad_(True)
        l64, w64 = c['loc'].double().requires_grad_(True), c['w'].double().requires_grad_(True)
        ref = ver_ref.voxel_multi_scale_deformable_attn_pytorch(v64, c['shapes'], l64, w64)
        go = c['gout'].to(dtype).double()
        gv_r, gl_r, gw_r = torch.autograd.grad(ref, (v64, l64, w64), go)
    else:
        ref, gv_r, gl_r, gw_r, go = c['out'], c['gvalue'], c['gloc'], c['gw'], c['gout']
    assert rel_err(out, ref) < tol
    out.backward(cuda(go).to(dtype))
    assert V.launch_count() - n0 == 3            # forward, memset + backward
    assert rel_err(v.grad, gv_r) < 2 * tol
    assert rel_err(w.grad, gw_r) < 2 * tol
    # d/d loc is one-sided exactly on a voxel centre / the padding border (planted at [0, 0, 0, 0, 0:3])
    lg, lr = l.grad.clone().cpu().double(), gl_r.clone().double()
    lg[0, 0, 0, 0, :3] = 0
    lr[0, 0, 0, 0, :3] = 0
    assert rel_err(lg, lr) < 2 * tol


@pytest.mark.parametrize('Dh', [16, 32, 64, 96, 128, 200])
def test_msda3d_random_vs_oracle(Dh):
    """every channels-per-lane specialisation (1, 2, 3, 4, 8), a ragged Dh, multiple levels."""
    g = torch.Generator().manual_seed(Dh)
    shapes = [(4, 9, 11), (2, 5, 6)]
    S = sum(d * h * w for d, h, w in shapes)
    Bv, Nq, NH, NP = 3, 157, 4, 4
    v = torch.randn(Bv, S, NH, Dh, generator=g) * 0.5
    loc = torch.rand(Bv, Nq, NH, 2, NP, 3, generator=g) * 1.4 - 0.2
    w = torch.rand(Bv, Nq, NH, 2 * NP, generator=g).softmax(-1).view(Bv, Nq, NH, 2, NP)
    go = torch.randn(Bv, Nq, NH * Dh, generator=g)
    v64, l64, w64 = (x.double().requires_grad_(True) for x in (v, loc, w))
    ref = ver_ref.voxel_multi_scale_deformable_attn_pytorch(v64, torch.tensor(shapes), l64, w64)
    gv_r, gl_r, gw_r = torch.autograd.grad(ref, (v64, l64, w64), go.double())
    out = ops.voxel_ms_deform_attn_forward(cuda(v), shapes, cuda(loc), cuda(w))
    gv, gl, gw = ops.voxel_ms_deform_attn_backward(cuda(v), shapes, cuda(loc), cuda(w), cuda(go))
    assert rel_err(out, ref) < 1e-5
    assert rel_err(gv, gv_r) < 2e-5 and rel_err(gl, gl_r) < 2e-5 and rel_err(gw, gw_r) < 2e-5


def test_msda3d_edge_cases():
    shapes = [(2, 3, 4)]
    v = torch.randn(1, 24, 2, 8, device=DEV)
    # every location outside the padded volume, NaN included: zero output, zero gradients
    loc = torch.full((1, 5, 2, 1, 4, 3), 3.0, device=DEV)
    loc[0, 0, 0, 0, 0, 0] = float('nan')
    w = torch.full((1, 5, 2, 1, 4), 0.25, device=DEV)
    out = ops.voxel_ms_deform_attn_forward(v, shapes, loc, w)
    gv, gl, gw = ops.voxel_ms_deform_attn_backward(v, shapes, loc, w, torch.ones_like(out))
    assert (out == 0).all() and (gv == 0).all() and (gl == 0).all() and (gw == 0).all()
    # a location on a voxel centre returns that voxel exactly; weights are linear
    loc = torch.tensor([(1 + 0.5) / 4, (2 + 0.5) / 3, (1 + 0.5) / 2], device=DEV).expand(1, 1, 2, 1, 1, 3).contiguous()
    w = torch.full((1, 1, 2, 1, 1), 2.0, device=DEV)
    out = ops.voxel_ms_deform_attn_forward(v, shapes, loc, w)
    assert torch.equal(out.view(2, 8), 2.0 * v[0, (1 * 3 + 2) * 4 + 1])
    # bad arguments surface as VerError with the reference's check
    with pytest.raises(V.VerError, match='num_value'):
        ops.voxel_ms_deform_attn_forward(v, [(2, 3, 5)], loc, w)
    with pytest.raises(V.VerError):
        ops.voxel_ms_deform_attn_forward(v, shapes, loc[..., :2].contiguous(), w)


def test_msda3d_full_size_properties():
    """config-2 volume (16x40x40, 768 channels), every voxel a query (the temporal self-attention shape):
    too large for the CPU oracle in seconds, so size-independent properties -- linearity in the weights and in the
    volume, and a partition-of-unity check (a constant volume sampled strictly inside returns the constant)."""
    D, H, W, NH, Dh, NP = 16, 40, 40, 8, 96, 4
    Nv = D * H * W
    g = torch.Generator(device=DEV).manual_seed(9)
    v = torch.randn(2, Nv, NH, Dh, device=DEV, generator=g)
    loc = torch.rand(2, Nv, NH, 1, NP, 3, device=DEV, generator=g)
    lo = torch.tensor([0.5 / W, 0.5 / H, 0.5 / D], device=DEV)
    loc_in = lo + loc * (1 - 2 * lo)                                     # strictly inside the voxel centres
    w = torch.rand(2, Nv, NH, NP, device=DEV, generator=g).softmax(-1).view(2, Nv, NH, 1, NP)
    shapes = [(D, H, W)]
    a = ops.voxel_ms_deform_attn_forward(v, shapes, loc_in, w)
    ones = ops.voxel_ms_deform_attn_forward(torch.ones_like(v), shapes, loc_in, w)
    assert (ones - 1).abs().max().item() < 1e-5
    b = ops.voxel_ms_deform_attn_forward(3.0 * v, shapes, loc_in, 0.5 * w)
    assert rel_err(b, 1.5 * a) < 1e-6
    # backward/forward duality: <out, go> differentiated wrt the volume equals the scatter of go
    go = torch.randn_like(a)
    gv, _, gw = ops.voxel_ms_deform_attn_backward(v, shapes, loc_in, w, go)
    assert abs((gv * v).sum().item() - (a * go).sum().item()) < 1e-3 * (a * go).abs().sum().item()
    assert abs((gw * w).sum().item() - (a * go).sum().item()) < 1e-3 * (a * go).abs().sum().item()


# ------------------------------------------------------------------ modules
def test_attention_module_golden():
    g = load_golden('decoder_c64.npz')
    grid = [int(x) for x in g['grid']]
    a = sub(g, 'attn')
    m = registry.build_attention(dict(type='VoxelCustomMSDeformableAttention', embed_dims=64, num_levels=1,
                                      batch_first=False)).to(DEV).eval()
    m.load_state_dict(sub(g, 'attn.sd'))
    with torch.no_grad():
        y = m(cuda(a['query']), key=None, value=cuda(a['value']), query_pos=cuda(a['query_pos']),
              reference_points=cuda(a['ref']), spatial_shapes=torch.tensor([grid]), level_start_index=torch.tensor([0]))
    assert rel_err(y, a['out']) < 1e-5


def test_decoder_golden_and_gradients():
    g = load_golden('decoder_c64.npz')
    grid = [int(x) for x in g['grid']]
    a, d = sub(g, 'attn'), sub(g, 'dec')
    dec = registry.build_transformer_layer_sequence(decoder_cfg()).to(DEV).eval()
    dec.load_state_dict(sub(g, 'dec.sd'))
    regs = _reg_branches(g).to(DEV)
    value = cuda(a['value']).requires_grad_(True)
    n0 = V.launch_count()
    hs, refs = dec(query=cuda(a['query']), key=None, value=value, query_pos=cuda(a['query_pos']),
                   reference_points=cuda(d['ref']), reg_branches=regs, cls_branches=None,
                   spatial_shapes=[grid], level_start_index=[0])
    assert V.launch_count() - n0 == 2, 'one sampler launch per decoder layer'
    assert rel_err(hs, d['hs']) < 1e-5 and rel_err(refs, d['refs']) < 1e-5
    # gradient wrt the voxel volume against the oracle's autograd
    go = torch.randn(hs.shape, generator=torch.Generator().manual_seed(1))
    hs.backward(cuda(go))
    v_ref = a['value'].clone().requires_grad_(True)
    sd = {k: v for k, v in sub(g, 'dec.sd').items()}
    hs_r, _ = ver_ref.decoder_forward(sd, '', a['query'], v_ref, a['query_pos'], d['ref'], torch.tensor([grid]),
                                      num_layers=2, reg_branches=_reg_branches(g))
    hs_r.backward(go)
    assert rel_err(value.grad, v_ref.grad) < 1e-4


def test_temporal_self_attention_golden():
    g = load_golden('decoder_c64.npz')
    grid = [int(x) for x in g['grid']]
    t = sub(g, 'tsa')
    m = registry.build_attention(dict(type='VoxelTemporalSelfAttention', embed_dims=64, num_levels=1)).to(DEV).eval()
    m.load_state_dict(sub(g, 'tsa.sd'))
    with torch.no_grad():
        y = m(cuda(t['query']), query_pos=cuda(t['query_pos']), reference_points=cuda(t['ref']),
              spatial_shapes=torch.tensor([grid]), level_start_index=torch.tensor([0]))
    assert rel_err(y, t['out']) < 1e-5


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_transformer_forward_lift_encode_decode(dtype):
    """VoxelPerceptionTransformer.forward end to end on the GPU -- 18 views lifted and encoded, then the
    box queries decoded out of the voxel volume -- against the oracle (get_voxel_features + transformer_decode)."""
    from vln_ver_b200 import synth
    C, grid, ncam, B = 256, (4, 8, 8), 18, 2       # Dh = 32: the fused SCA kernels cover Dh in {32, 64, 96, 128}
    cfg = V.vocc_head_cfg(*grid, num_cams=ncam, embed_dims=C, num_layers=1, ffn_dims=2 * C)['transformer']
    cfg['decoder'] = decoder_cfg(C, 2)
    torch.manual_seed(5)
    tr = registry.build_transformer(cfg)
    tr.init_weights()
    sd = {k: v.detach().clone() for k, v in tr.state_dict().items()}
    l2i, sh = synth.make_rig(B, ncam, grid, seed=4)
    feats = torch.from_numpy(synth.make_features(B, ncam, dim=C, seed=6))
    l2i, sh = torch.from_numpy(l2i), torch.from_numpy(sh)
    Nq = grid[0] * grid[1] * grid[2]
    queries = torch.randn(Nq, C)
    oqe = torch.randn(10, 2 * C)
    with torch.no_grad():
        vox_r = ver_ref.get_voxel_features(sd, '', feats, queries, *grid, synth.PC_RANGE, l2i, sh, num_layers=1)
        _, hs_r, init_r, refs_r = ver_ref.transformer_decode(sd, '', vox_r, oqe, *grid, num_layers=2)
    tr = tr.to(DEV).eval()
    V.set_compute_dtype(tr, dtype)
    with torch.no_grad():
        vox, hs, init_ref, refs = tr(cuda(feats), cuda(queries), cuda(oqe), *grid, lidar2img=cuda(l2i),
                                     originshift=cuda(sh))
    tol = 1e-5 if dtype == torch.float32 else 5e-3      # fp16: storage rounding through 3 stacked layers
    assert rel_err(vox.permute(1, 0, 2), vox_r) < tol
    assert rel_err(hs, hs_r) < tol and rel_err(init_ref, init_r) < 1e-5 and rel_err(refs, refs_r) < 1e-5


def test_head_default_forward_with_decoder_vs_oracle():
    """vocc.py default head (only_occ=False, with_box_refine=True) on the GPU: lift + encode + decode + the
    detection and occupancy tails, against the oracle composed from its pinned pieces."""
    import torch.nn.functional as F
    from vln_ver_b200 import synth
    C, grid, ncam, B, L, nq = 256, (4, 8, 8), 18, 2, 2, 10
    cfg = V.vocc_head_cfg(*grid, num_cams=ncam, embed_dims=C, only_occ=False, refine_occ=False,
                          occupancy_size=[1.5, 1.5, 0.5], occ_dims=16, num_layers=1, num_decoder_layers=L,
                          num_query=nq)
    torch.manual_seed(11)
    head = V.build_head(cfg)
    head.init_weights()
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():       # reg branches start near zero after xavier; make the refinement visible
        for p in head.reg_branches.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    sd = {k: v.detach().clone() for k, v in head.state_dict().items()}
    l2i, sh = synth.make_rig(B, ncam, grid, seed=14)
    feats = torch.from_numpy(synth.make_features(B, ncam, dim=C, seed=15))
    l2i, sh = torch.from_numpy(l2i), torch.from_numpy(sh)

    def reg(l):
        def f(x):
            for i in (0, 2):
                x = F.relu(F.linear(x, sd[f'reg_branches.{l}.{i}.weight'], sd[f'reg_branches.{l}.{i}.bias']))
            return F.linear(x, sd[f'reg_branches.{l}.4.weight'], sd[f'reg_branches.{l}.4.bias'])
        return f
    with torch.no_grad():
        vox_r = ver_ref.get_voxel_features(sd, 'transformer.', feats, sd['voxel_embedding.weight'], *grid,
                                           synth.PC_RANGE, l2i, sh, num_layers=1)
        bev_r, hs_r, init_r, refs_r = ver_ref.transformer_decode(
            sd, 'transformer.', vox_r, sd['query_embedding.weight'], *grid, num_layers=L,
            reg_branches=[reg(l) for l in range(L)])
        cls_r, box_r = ver_ref.detection_tail(sd, '', hs_r, init_r, refs_r, head.pc_range)
        occ_r = ver_ref.occ_head(sd, '', vox_r, *grid, head.occ_xdim, head.occ_ydim, head.occ_zdim, occ_dims=16,
                                 refine_occ=False, only_occ=True)
    head = head.to(DEV).eval()
    n0 = V.launch_count()
    with torch.no_grad():
        outs = head(cuda(feats), None, lidar2img=cuda(l2i), originshift=cuda(sh))
    assert V.launch_count() - n0 >= L + 3
    assert outs['all_cls_scores'].shape == (L, B, nq, 17) and outs['all_bbox_preds'].shape == (L, B, nq, 10)
    assert rel_err(outs['bev_embed'], bev_r) < 1e-5
    assert rel_err(outs['all_cls_scores'], cls_r) < 1e-5
    assert rel_err(outs['all_bbox_preds'], box_r) < 1e-5
    assert rel_err(outs['occupancy_preds'], occ_r) < 1e-5
    assert outs['all_layout_preds'] is None
    # default-branch loss(): the occupancy term of the last decoder layer (HEAD:1324-1332, :977-986)
    gts = [torch.from_numpy(x) for x in synth.make_occ_gt(B, head.voxel_num, frac=0.2, seed=77)]
    loss_ref = ver_ref.occupancy_loss(occ_r, gts)
    losses = head.loss(None, None, None, [cuda(t) for t in gts], None, outs)
    assert set(losses) == {'loss_occupancy', 'loss_flow'} and float(losses['loss_flow']) == 0.0
    assert abs(losses['loss_occupancy'].item() - loss_ref.item()) < 1e-5 * abs(loss_ref.item())
    with pytest.raises(NotImplementedError):
        head.loss(None, None, None, None, None, dict(outs, occupancy_preds=None))
