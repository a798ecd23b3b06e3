"""Worker of tests/test_gpu_ddp.py (launched by torch.distributed.run, one rank per GPU, NCCL):
DDP over the fused fp16 encoder layer (vln_ver_b200.fused_layer.VoxelLayerFunction) must reproduce the gradient of
the combined batch on one GPU.  Rank r owns panorama r; DDP averages the per-rank gradients, the single-GPU loss is
the mean over the panoramas (HEAD.loss_only_occupancy), so both are (g_0 + ... + g_{W-1}) / W."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vln_ver_b200 as V  # noqa: E402
from vln_ver_b200 import synth  # noqa: E402
from vln_ver_b200.config import per_voxel_occupancy_size  # noqa: E402

LS = 4096.0


def build(grid, ncam, C):
    torch.manual_seed(0)
    cfg = V.vocc_head_cfg(*grid, num_cams=ncam, embed_dims=C, only_occ=True, refine_occ=False,
                          occupancy_size=per_voxel_occupancy_size(*grid), num_layers=2, occ_dims=32)
    head = V.build_head(cfg)
    head.init_weights()
    g = torch.Generator().manual_seed(100)
    for n, p in head.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            with torch.no_grad():
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
        if n.startswith('positional_encoding'):
            p.requires_grad_(False)
    for m in head.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return head


def grads_of(head, model, feats, l2i, sh, gts, dev):
    head.zero_grad(set_to_none=True)
    outs = model(feats.to(dev), None, lidar2img=l2i.to(dev), originshift=sh.to(dev))
    loss = head.loss_only_occupancy(None, None, None, [t.to(dev) for t in gts], None, outs)['loss_occupancy']
    (loss * LS).backward()
    return {n: p.grad.float() / LS for n, p in head.named_parameters() if p.grad is not None}, loss.item()


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    grid, ncam, C = (4, 8, 8), 18, 256
    l2i, sh = synth.make_rig(world, ncam, grid, seed=19)
    feats = torch.from_numpy(synth.make_features(world, ncam, dim=C, seed=20))
    l2i, sh = torch.from_numpy(l2i), torch.from_numpy(sh)
    head0 = build(grid, ncam, C)
    gts = [torch.from_numpy(x) for x in synth.make_occ_gt(world, head0.voxel_num, frac=0.2)]
    # ---- DDP: one panorama per rank
    head = build(grid, ncam, C).to(dev).train()
    V.set_compute_dtype(head, torch.float16)
    model = torch.nn.parallel.DistributedDataParallel(head, device_ids=[local], gradient_as_bucket_view=True,
                                                      static_graph=True)
    n0 = V.launch_count()
    g_ddp, _ = grads_of(head, model, feats[:, rank:rank + 1], l2i[rank:rank + 1], sh[rank:rank + 1],
                        gts[rank:rank + 1], dev)
    assert V.launch_count() > n0
    # ---- the combined batch on this GPU, no DDP
    single = build(grid, ncam, C).to(dev).train()
    V.set_compute_dtype(single, torch.float16)
    g_one, _ = grads_of(single, single, feats, l2i, sh, gts, dev)
    worst = 0.0
    for n, g in g_one.items():
        ref = g.abs().max().item()
        if ref < 1e-12:
            continue
        err = (g_ddp[n] - g).abs().max().item() / ref
        worst = max(worst, err)
        assert err < 5e-3, (n, err)
    # every rank holds the same averaged gradient
    for n, g in sorted(g_ddp.items())[:8]:
        t = g.clone()
        dist.broadcast(t, 0)
        assert torch.equal(t, g), n
    dist.barrier()
    if rank == 0:
        print(f'DDP_PARITY_OK world={world} params={len(g_one)} max_rel_err={worst:.2e}', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
