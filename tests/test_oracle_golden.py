"""CPU tier: the oracle restatement (oracle/ver_ref.py) against the committed golden
vectors that oracle/gen_golden.py produced from the UNMODIFIED reference."""
import torch

from oracle import ver_ref
from vln_ver_b200 import synth
from conftest import load_golden, rel_err, sub

PC = synth.PC_RANGE


def test_msda_restatement_matches_reference_3d_sampler():
    g = load_golden('msda_cases.npz')
    for name in ('small', 'dh96', 'rect'):
        c = sub(g, name)
        v = c['value'].double().requires_grad_(True)
        l = c['loc'].double().requires_grad_(True)
        w = c['w'].double().requires_grad_(True)
        out = ver_ref.multi_scale_deformable_attn_pytorch(v, c['shape'][None], l, w)
        assert rel_err(out, c['out']) < 1e-12
        gv, gl, gw = torch.autograd.grad(out, (v, l, w), c['gout'].double())
        assert rel_err(gv, c['gvalue']) < 1e-12
        assert rel_err(gl, c['gloc']) < 1e-12
        assert rel_err(gw, c['gw']) < 1e-12


def test_point_sampling_bit_exact():
    g = load_golden('point_sampling_6cam.npz')
    for tag, grid in (('g4x15x15', (4, 15, 15)), ('g8x20x20', (8, 20, 20))):
        c = sub(g, tag)
        ref = ver_ref.get_reference_points_3d(*grid)
        assert torch.equal(ref, c['ref_3d'])
        rpc, mask = ver_ref.point_sampling(ref, PC, c['lidar2img'][0], c['originshift'][0])
        assert torch.equal(rpc, c['rpc'])
        assert torch.equal(mask, c['mask'])
        idx = ver_ref.visible_indexes(mask)
        assert [len(i) for i in idx] == c['index_len'].tolist()
        assert torch.equal(torch.cat(idx), c['index_cat'])


def test_sca_matches_unmodified_module():
    g = load_golden('sca.npz')
    for tag in ('c6', 'c18'):
        c = sub(g, tag)
        sd = sub(g, tag + '.sd')
        grid = c['grid'].tolist()
        rpc, mask = ver_ref.point_sampling_batched(*grid, PC, c['lidar2img'], c['originshift'])
        y = ver_ref.sca_forward(sd, '', c['query'], c['value'], rpc, mask, torch.tensor([[14, 14]]))
        assert rel_err(y, c['out']) < 1e-6


def test_encoder_matches_unmodified_module():
    g = load_golden('encoder_6cam_c256.npz')
    c = {k: torch.from_numpy(v) for k, v in g.items() if not k.startswith('sd.')}
    sd = sub(g, 'sd')
    grid = c['grid'].tolist()
    y = ver_ref.encoder_forward(sd, '', c['bev_query'], c['value'], *grid, PC, c['lidar2img'],
                                c['originshift'], torch.tensor([[14, 14]]), num_layers=2)
    assert rel_err(y, c['out']) < 2e-6


def test_head_positional_decode():
    g = load_golden('head.npz')
    for tag in ('pervoxel', 'column'):
        c = sub(g, tag)
        sd = sub(g, tag + '.sd')
        grid = c['grid'].tolist()
        ox, oy, oz = c['occ_dims3'].tolist()
        y = ver_ref.occ_head(sd, '', c['bev_embed'], *grid, ox, oy, oz, occ_dims=16, refine_occ=False,
                             only_occ=True)
        assert rel_err(y, c['occupancy_preds']) < 1e-6
        assert torch.equal(ver_ref.positional_encoding(sd, 'positional_encoding.', 1, *grid), c['pos'])
        assert torch.equal(ver_ref.get_occupancy_prediction(c['decode_logits']), c['decode'])


def test_focal_loss_restatement_gradient_is_finite_and_matches_numeric():
    torch.manual_seed(0)
    x = torch.randn(50, 16, dtype=torch.float64, requires_grad=True)
    gt = torch.from_numpy(synth.make_occ_gt(1, 50, frac=0.3)[0])
    loss = ver_ref.occupancy_loss(x[None], [gt])
    loss.backward()
    eps = 1e-6
    xp = x.detach().clone(); xp[3, 5] += eps
    xm = x.detach().clone(); xm[3, 5] -= eps
    num = (ver_ref.occupancy_loss(xp[None], [gt]) - ver_ref.occupancy_loss(xm[None], [gt])) / (2 * eps)
    assert abs(num.item() - x.grad[3, 5].item()) < 1e-7


def test_refine_occ_tail_shapes():
    """default-branch head tail (raw .view reinterpretations, 3x ConvTranspose3d) -- restated only."""
    torch.manual_seed(0)
    C, grid = 768, (2, 3, 3)
    sd = {}
    for i in range(3):
        sd[f'up_sample.{i}.weight'] = torch.randn(768, 768, 3, 5, 5) * 0.01
        sd[f'up_sample.{i}.bias'] = torch.zeros(768)
    sd['occ_proj.weight'] = torch.randn(16 * 7, 2 * C) * 0.01
    sd['occ_proj.bias'] = torch.zeros(16 * 7)
    for i in (0, 3):
        sd[f'occ_branches.{i}.weight'] = torch.randn(16, 16) * 0.1
        sd[f'occ_branches.{i}.bias'] = torch.zeros(16)
        sd[f'occ_branches.{i + 1}.weight'] = torch.ones(16)
        sd[f'occ_branches.{i + 1}.bias'] = torch.zeros(16)
    sd['occ_branches.6.weight'] = torch.randn(16, 16) * 0.1
    sd['occ_branches.6.bias'] = torch.zeros(16)
    bev = torch.randn(1, 18, C)
    y = ver_ref.occ_head(sd, '', bev, *grid, 24, 24, 7, occ_dims=16, refine_occ=True, only_occ=False)
    assert y.shape == (1, 7 * 24 * 24, 16)


def _refine_fixture_state_dict(g):
    """state_dict of the refine_occ fixture: the small tensors are stored, the 133 M up_sample weights are
    regenerated from the generator's seeds (oracle/gen_golden.py::refine_upsample_weights)."""
    from oracle.gen_golden import refine_upsample_weights
    sd = sub(g, 'sd')
    for i in range(3):
        sd[f'up_sample.{i}.weight'], sd[f'up_sample.{i}.bias'] = refine_upsample_weights(i)
    return sd


def test_refine_occ_tail_pinned_to_the_unmodified_head():
    """HEAD:551-580 with refine_occ=True (raw .view at :558 / :564, three ConvTranspose3d(768, 768), column
    occ_proj): the restatement reproduces the output of the UNMODIFIED VoxelFormerOccupancyHead.forward recorded by
    oracle/gen_golden.py::gen_head_refine (this tail was compared with the restatement only)."""
    g = load_golden('head_refine_c768.npz')
    sd = _refine_fixture_state_dict(g)
    grid = g['grid'].tolist()
    ox, oy, oz = g['occ_dims3'].tolist()
    bev = torch.from_numpy(g['bev_embed'])                       # (Nq, 1, C) as the transformer returns it
    with torch.no_grad():
        y = ver_ref.occ_head(sd, '', bev.permute(1, 0, 2), *grid, ox, oy, oz, occ_dims=16, refine_occ=True,
                             only_occ=False)
    assert rel_err(y, torch.from_numpy(g['occupancy_preds'])) < 1e-5
