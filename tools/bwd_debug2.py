"""Repeat new-vs-old backward comparisons (debug): CPU-generated test-like data, several seeds."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import _lib, ops, synth
B, ncam, grid, NH, Dh = 2, 18, (8, 20, 20), 8, 96
quick = '--quick' in sys.argv
if quick:
    B, grid = 1, (4, 8, 8)
Nq = grid[0] * grid[1] * grid[2]
_lib.lib.ver_debug_bwd_variant.argtypes = [ctypes.c_int]
for seed in ([5] if quick else [5, 6, 7]):
    l2i, sh = synth.make_rig(B, ncam, grid, seed=11)
    rpc, mask, bits, count = ops.point_sampling(torch.from_numpy(l2i).cuda(), torch.from_numpy(sh).cuda(), synth.PC_RANGE, *grid)
    vis = ops.Visibility(rpc, mask, bits, count, grid)
    g = torch.Generator().manual_seed(seed)
    value = (torch.randn(B * ncam, 196, NH * Dh, generator=g) * 0.5).half().cuda()
    logits = torch.randn(B * Nq, 192, generator=g)
    logits[:, :128] *= 2.0
    logits = logits.cuda()
    gout = torch.randn(B, Nq, NH * Dh, generator=g).half().cuda()
    res = []
    for variant in (0, 1, 0):
        _lib.lib.ver_debug_bwd_variant(variant)
        vc = value.clone().requires_grad_(True)
        lc = logits.clone().requires_grad_(True)
        out = ops.sca_sample_tc(vc, lc, vis, 14, 14, NH, 8)
        out.backward(gout)
        torch.cuda.synchronize()
        res.append((vc.grad.float().clone(), lc.grad.clone()))
    _lib.lib.ver_debug_bwd_variant(0)
    sc = res[1][1].abs().max().item()
    print(f'seed {seed}: max|glogits| {sc:.3f}; new-old {(res[0][1]-res[1][1]).abs().max().item()/sc:.2e}; '
          f'new-new {(res[0][1]-res[2][1]).abs().max().item()/sc:.2e}; gvalue new-old '
          f'{(res[0][0]-res[1][0]).abs().max().item()/res[1][0].abs().max().item():.2e}')
    d = (res[0][1] - res[1][1]).abs()
    w = d.amax(1).argmax().item()
    print('   worst row', w, 'b', w // Nq, 'count', int(count.view(-1)[w]), 'col', d[w].argmax().item(),
          'new', res[0][1][w, d[w].argmax()].item(), 'old', res[1][1][w, d[w].argmax()].item())
    cols = d.amax(0)
    print('   max diff: offsets', cols[:128].max().item() / sc, 'attention', cols[128:].max().item() / sc)
