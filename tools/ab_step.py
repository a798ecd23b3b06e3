#!/usr/bin/env python
"""A/B of the training step (headline configuration, CUDA-graph replay) under different routings of the projections
(fused_layer.TC_GEMM) inside ONE process on ONE box -- box-to-box spread of the step time is larger than the effects.
    python tools/ab_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vln_ver_b200 import fused_layer as FL  # noqa: E402

CASES = [('library GEMMs only', dict(logits=False, ffn1=False, value_proj=False, output_proj=False, ffn2=False, ffn2_bwd=False)),
         ('+ logits', dict(logits=True)),
         ('+ ffn1 (bias+ReLU+dropout epilogue)', dict(ffn1=True)),
         ('+ ffn2 backward (mask + colsum epilogue)', dict(ffn2_bwd=True)),
         ('+ output_proj, ffn2, value_proj forward', dict(output_proj=True, ffn2=True, value_proj=True)),
         ('default', None)]


def main():
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    default = dict(FL.TC_GEMM)
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    results = {}
    for rnd in range(rounds):
        state = dict(default)
        for name, delta in CASES:
            if delta is None:
                state = dict(default)
            else:
                state.update(delta)
            FL.TC_GEMM.clear()
            FL.TC_GEMM.update(state)
            c = bench.Config('train', bench.PER_GPU_BATCH, bench.GRID, dev, 0, 1, 0, n_pool=3)
            c.capture(3)
            ms, _ = c.timed_device(10, 3)
            results.setdefault(name, []).append(ms / 10)
            print(f'round {rnd}: {name:45s} {ms / 10:7.3f} ms/step  ({c.graph_note})', flush=True)
            c.release()
    print('--- best of rounds')
    for name, v in results.items():
        print(f'{name:45s} {min(v):7.3f} ms/step')


if __name__ == '__main__':
    main()
