#!/usr/bin/env python
"""up_sample stack of the shipped vocc.py head (3 x ConvTranspose3d 768->768, 4x15x15 -> 4x120x120):
the stack as written (cuDNN, 1.67 TFLOP / panorama) against the lattice form (vln_ver_b200/upsample.py,
0.478 TFLOP).  CUDA events; prints one JSON line per (dtype, batch, direction)."""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200.upsample import up_sample_gemm, up_sample_lattice  # noqa: E402


def timeit(f, iters=5, warm=2):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    convs = nn.Sequential(*[nn.ConvTranspose3d(768, 768, (3, 5, 5), stride=(1, 2, 2), padding=(2, 4, 4),
                                               dilation=(2, 2, 2), output_padding=(0, 1, 1)) for _ in range(3)]).cuda()
    with torch.no_grad():
        for p in convs.parameters():
            p.mul_(0.5)
    for dtype in (torch.float16, torch.float32):
        for bs in (1, 8):
            x = torch.randn(bs, 768, 4, 15, 15, device='cuda').to(dtype)
            wl = [(c.weight.detach().to(dtype), c.bias.detach().to(dtype)) for c in convs]

            def dense(inp=x):
                y = inp
                for (w, b), c in zip(wl, convs):
                    y = nn.functional.conv_transpose3d(y, w, b, stride=c.stride, padding=c.padding,
                                                       output_padding=c.output_padding, dilation=c.dilation)
                return y

            def lattice(inp=x):
                return up_sample_lattice(inp, convs, dtype=dtype)

            def gemm(inp=x):
                return up_sample_gemm(inp, convs, dtype=dtype)
            with torch.no_grad():
                ref, y, yg = dense(), lattice(), gemm()
                err = ((ref.float() - y.float()).abs().max() / ref.float().abs().max()).item()
                errg = ((ref.float() - yg.float()).abs().max() / ref.float().abs().max()).item()
                t_d, t_l, t_g = timeit(dense), timeit(lattice), timeit(gemm)
            flop_d, flop_l = 1.672e12 * bs, 0.478e12 * bs
            print(json.dumps({'dtype': str(dtype).split('.')[-1], 'batch': bs, 'pass': 'forward',
                              'dense_ms': round(t_d, 3), 'lattice_cudnn_ms': round(t_l, 3), 'lattice_gemm_ms': round(t_g, 3),
                              'speedup_gemm_vs_dense': round(t_d / t_g, 2),
                              'dense_TFLOPs': round(flop_d / t_d / 1e9, 1), 'lattice_cudnn_TFLOPs': round(flop_l / t_l / 1e9, 1),
                              'lattice_gemm_TFLOPs': round(flop_l / t_g / 1e9, 1),
                              'max_rel_diff_cudnn': err, 'max_rel_diff_gemm': errg}), flush=True)
            if bs == 1:
                xg = x.clone().requires_grad_(True)

                def fb(fn):
                    def run():
                        for p in convs.parameters():
                            p.grad = None
                        xg.grad = None
                        y = up_sample_lattice(xg, convs, dtype=dtype) if fn == 'lattice' else \
                            up_sample_gemm(xg, convs, dtype=dtype) if fn == 'gemm' else \
                            nn.Sequential(*convs)(xg) if dtype == torch.float32 else None
                        if y is None:       # fp16 dense with autograd through the casts
                            y = xg
                            for c in convs:
                                y = nn.functional.conv_transpose3d(y, c.weight.to(dtype), c.bias.to(dtype), stride=c.stride,
                                                                   padding=c.padding, output_padding=c.output_padding,
                                                                   dilation=c.dilation)
                        y.float().square().mean().backward()
                    return run
                t_d, t_l = timeit(fb('dense'), iters=3, warm=1), timeit(fb('lattice'), iters=3, warm=1)
                t_g = timeit(fb('gemm'), iters=3, warm=1)
                print(json.dumps({'dtype': str(dtype).split('.')[-1], 'batch': bs, 'pass': 'forward+backward',
                                  'dense_ms': round(t_d, 3), 'lattice_cudnn_ms': round(t_l, 3),
                                  'lattice_gemm_ms': round(t_g, 3), 'speedup_gemm_vs_dense': round(t_d / t_g, 2)}),
                      flush=True)


if __name__ == '__main__':
    main()
