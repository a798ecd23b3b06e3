#!/usr/bin/env python
"""Inference forward of the SHIPPED vocc.py head (SURVEY.md 8(d) "separately report the shipped vocc.py head"):
6 views, 15x15x4 voxels, 3 encoder layers, 6 decoder layers x 100 box queries (3-D deformable sampler),
refine_occ=True (up_sample 15x15 -> 120x120, occ_proj Linear(3072, 4480)), 504 000 occupancy logits x 16.
Times the whole head forward with the up_sample stack as written (cuDNN) and in lattice form (GEMM + col2im
kernel), fp16 storage and fp32.  CUDA events; one JSON line per configuration."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vln_ver_b200 as V            # noqa: E402
from vln_ver_b200 import synth      # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    grid, ncam = (4, 15, 15), 6
    torch.manual_seed(0)
    head = V.build_head(V.vocc_head_cfg())          # the shipped values
    head.init_weights()
    head = head.cuda().eval()
    for dtype in (torch.float16, torch.float32):
        V.set_compute_dtype(head, dtype)
        for bs in (1, 8):
            l2i, sh = synth.make_rig(bs, ncam, grid, seed=3)
            feats = torch.from_numpy(synth.make_features(bs, ncam, dim=768, seed=4)).cuda()
            l2i, sh = torch.from_numpy(l2i).cuda(), torch.from_numpy(sh).cuda()
            res = {}
            outs = {}
            for mode in ('dense', 'auto', 'gemm+occ_lattice'):
                head.up_sample_mode = 'gemm' if '+' in mode else mode
                head.occ_proj_mode = 'lattice' if '+' in mode else 'dense'

                def run():
                    with torch.no_grad():
                        return head(feats, None, lidar2img=l2i, originshift=sh)
                for _ in range(2):
                    o = run()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0 = V.launch_count()
                a.record()
                for _ in range(5):
                    o = run()
                b.record()
                torch.cuda.synchronize()
                res[mode] = a.elapsed_time(b) / 5
                outs[mode] = o
                launches = (V.launch_count() - n0) // 5
            d = (outs['auto']['occupancy_preds'] - outs['dense']['occupancy_preds']).abs().max().item()
            ref = outs['dense']['occupancy_preds'].abs().max().item()
            print(json.dumps({'workload': 'shipped vocc.py head, inference forward', 'dtype': str(dtype).split('.')[-1],
                              'batch': bs, 'ms_up_sample_as_written': round(res['dense'], 3),
                              'ms_lattice_form': round(res['auto'], 3),
                              'ms_gemm_up_sample_and_lattice_occ_proj': round(res['gemm+occ_lattice'], 3),
                              'occ_lattice_max_rel_diff': (outs['gemm+occ_lattice']['occupancy_preds']
                                                           - outs['dense']['occupancy_preds']).abs().max().item() / ref,
                              'panoramas_per_s_as_written': round(bs / res['dense'] * 1e3, 1),
                              'panoramas_per_s_lattice': round(bs / res['auto'] * 1e3, 1),
                              'libver_launches_per_forward': launches,
                              'occupancy_max_rel_diff': d / ref,
                              'shapes': {k: list(v.shape) for k, v in outs['auto'].items() if v is not None}}), flush=True)


if __name__ == '__main__':
    main()
