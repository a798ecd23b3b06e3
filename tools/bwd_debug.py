"""Localise differences between the two tensor-core backward kernels (debug)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import _lib, ops, synth

B, ncam, grid, NH, Dh = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 18, (8, 20, 20), 8, 96
Nq = grid[0] * grid[1] * grid[2]
l2i, sh = synth.make_rig(B, ncam, grid, seed=11)
rpc, mask, bits, count = ops.point_sampling(torch.from_numpy(l2i).cuda(), torch.from_numpy(sh).cuda(), synth.PC_RANGE, *grid)
vis = ops.Visibility(rpc, mask, bits, count, grid)
g = torch.Generator(device='cuda').manual_seed(0)
value = (torch.randn(B * ncam, 196, NH * Dh, device='cuda', generator=g) * 0.5).half()
logits = torch.randn(B * Nq, 192, device='cuda', generator=g)
logits[:, :128] *= 2
vimg = ops.value_image(value, NH)
counts, index = vis.index
gs = torch.randn(B, Nq, NH * Dh, device='cuda', generator=g).half()
_lib.lib.ver_debug_bwd_variant.argtypes = [ctypes.c_int]
outs = []
for variant in (0, 1):
    _lib.lib.ver_debug_bwd_variant(variant)
    gvalue = torch.empty(B * ncam, 196, NH * Dh, device='cuda')
    glogits = torch.empty_like(logits)
    _lib.check(_lib.lib.ver_sca_backward(_lib.VER_F16, vimg.data_ptr(), ops.VER_LAYOUT_TC_IMAGE, logits.data_ptr(), 192,
                                         rpc.data_ptr(), bits.data_ptr(), counts.data_ptr(), index.data_ptr(),
                                         gs.data_ptr(), gvalue.data_ptr(), glogits.data_ptr(), B, ncam, *grid, 14, 14,
                                         NH, Dh, 8, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    outs.append((gvalue, glogits))
_lib.lib.ver_debug_bwd_variant(0)
new, old = outs[0][1], outs[1][1]
d = (new - old).abs()
scale = old.abs().max().item()
print('max |old|', scale, 'max diff', d.max().item())
off = d[:, :128].view(Nq, NH, 8, 2)
att = d[:, 128:].view(Nq, NH, 8)
print('offset-x max diff per point', off[..., 0].amax((0, 1)).tolist())
print('offset-y max diff per point', off[..., 1].amax((0, 1)).tolist())
print('attention max diff per point', att.amax((0, 1)).tolist())
print('per head max diff', d[:, :128].view(Nq, NH, 16).amax((0, 2)).tolist())
dvv = d.amax(1).view(B, Nq)
for bb in range(B):
    print('panorama', bb, 'max diff', dvv[bb].max().item(), 'mean', dvv[bb].mean().item())
    for c in range(ncam):
        k = int(counts[bb, c])
        ids = index[bb, c, :k].long()
        if k:
            print(f'   cam {c:2d}: {k:4d} hits, max diff over its voxels {dvv[bb][ids].max().item():.3e}')
worst = d.amax(1).argmax().item()
print('worst row', worst, 'b', worst // Nq, 'count', int(count.view(-1)[worst]), 'bits', hex(int(bits.view(-1)[worst]) & 0xffffffff))
print('new', new[worst, :16].tolist())
print('old', old[worst, :16].tolist())
gv = (outs[0][0] - outs[1][0]).abs().view(B * ncam, -1).amax(1)
print('grad_value max diff per view', gv.tolist())
