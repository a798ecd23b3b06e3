#!/usr/bin/env python
"""Micro-benchmark of the 3-D deformable sampler kernels (csrc/msda3d.cu) at the two shapes the path uses:
 - decoder:   100 box queries per panorama reading the 16x40x40x768 volume (VoxelCustomMSDeformableAttention)
 - self-attn: every voxel a query, (previous | current) volumes (VoxelTemporalSelfAttention)
CUDA events, L2 flushed between iterations.  Prints one JSON line per shape."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import ops  # noqa: E402


def bench(name, Bv, grid, Nq, NH=8, Dh=96, NP=4, dtype=torch.float16, iters=10):
    D, H, W = grid
    S = D * H * W
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    v = torch.randn(Bv, S, NH, Dh, device=dev, generator=g).to(dtype)
    # reference points: the query's own voxel centre when every voxel is a query (self-attention, the
    # reference's ref_2d grid, M/voxel_encoder.py:47-76), anywhere in the volume for the box queries
    if Nq == S:
        zs, ys, xs = torch.meshgrid(*[(torch.arange(n, device=dev) + 0.5) / n for n in grid], indexing='ij')
        base = torch.stack((xs, ys, zs), -1).view(1, S, 1, 1, 1, 3).expand(Bv, -1, -1, -1, -1, -1)
    else:
        base = torch.rand(Bv, Nq, 1, 1, 1, 3, device=dev, generator=g)
    loc = (base + (torch.rand(Bv, Nq, NH, 1, NP, 3, device=dev, generator=g) - 0.5) * 0.2).contiguous()
    w = torch.rand(Bv, Nq, NH, 1, NP, device=dev, generator=g).softmax(-1)
    go = torch.randn(Bv, Nq, NH * Dh, device=dev, generator=g).to(dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [list(grid)]
    tf, tb = [], []
    for i in range(iters + 3):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = ops.voxel_ms_deform_attn_forward(v, shapes, loc, w)
        e[1].record()
        ops.voxel_ms_deform_attn_backward(v, shapes, loc, w, go)
        e[2].record()
        torch.cuda.synchronize()
        if i >= 3:
            tf.append(e[0].elapsed_time(e[1]))
            tb.append(e[1].elapsed_time(e[2]))
    b = v.element_size()
    rows = Bv * Nq
    fwd_bytes = Bv * S * NH * Dh * b + rows * NH * NP * 16 + rows * NH * Dh * b      # volume once + loc/w + out
    gather_bytes = rows * NH * NP * 8 * Dh * b                                      # L2-side gather traffic
    f, bw = sum(tf) / len(tf), sum(tb) / len(tb)
    print(json.dumps({'shape': name, 'Bv': Bv, 'grid': grid, 'Nq': Nq, 'dtype': str(dtype).split('.')[-1],
                      'fwd_us': round(f * 1e3, 1), 'bwd_us': round(bw * 1e3, 1),
                      'fwd_hbm_GBps_algorithmic': round(fwd_bytes / f / 1e6, 1),
                      'fwd_gather_GBps_L2': round(gather_bytes / f / 1e6, 1)}), flush=True)


if __name__ == '__main__':
    assert torch.cuda.is_available()
    for dt in (torch.float32, torch.float16):
        bench('decoder 100 queries x 8 panoramas', 8, (16, 40, 40), 100, dtype=dt)
        bench('self-attn all voxels, 2 volumes', 2, (16, 40, 40), 25600, dtype=dt)
