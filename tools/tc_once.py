#!/usr/bin/env python
"""Launch the sorted-row forward sampler a few times at the benchmark shape (for ncu captures):
    ncu --set full --import-source on -k regex:sca_fwd_tc -s 1 -c 1 -o gpurun_out/x python tools/tc_once.py [variant]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import _lib, ops, synth  # noqa: E402


def main():
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    B, ncam, grid, NH, Dh = 8, 18, (16, 40, 40), 8, 96
    Nq = grid[0] * grid[1] * grid[2]
    l2i, sh = synth.make_rig(B, ncam, grid, seed=1235)
    rpc, mask, bits, count = ops.point_sampling(torch.from_numpy(l2i).cuda(), torch.from_numpy(sh).cuda(),
                                                synth.PC_RANGE, *grid)
    vis = ops.Visibility(rpc, mask, bits, count, grid)
    g = torch.Generator(device='cuda').manual_seed(0)
    value = (torch.randn(B * ncam, 196, NH * Dh, device='cuda', generator=g) * 0.5).half()
    logits = torch.randn(B * Nq, 192, device='cuda', generator=g)
    logits[:, :128] *= 2
    order, smask, tu = vis.order
    vimg = ops.value_image(value, NH)
    slots = torch.empty((B, Nq, NH * Dh), dtype=torch.float16, device='cuda')
    vimg16 = ops.value_image16(value, NH, 14, 14) if variant in (6, 7) else None
    for _ in range(3):
        if variant in (6, 7):          # generations on 16-cell image rows
            _lib.check(_lib.lib.ver_sca_forward_sorted16(vimg16.data_ptr(), logits.data_ptr(), 192, rpc.data_ptr(),
                                                         order.data_ptr(), smask.data_ptr(), tu.data_ptr(),
                                                         slots.data_ptr(), B, ncam, Nq, 14, 14, NH, Dh, 8, variant,
                                                         torch.cuda.current_stream().cuda_stream))
            continue
        _lib.check(_lib.lib.ver_sca_forward_sorted(vimg.data_ptr(), logits.data_ptr(), 192, rpc.data_ptr(),
                                                   order.data_ptr(), smask.data_ptr(), tu.data_ptr(),
                                                   slots.data_ptr(), B, ncam, Nq, 14, 14, NH, Dh, 8, variant,
                                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print('ok', float(slots.float().abs().mean()))


if __name__ == '__main__':
    main()
