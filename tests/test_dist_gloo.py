"""CPU tier, world_size 2 over gloo: the data-parallel contract of SURVEY.md 8(e) -- panoramas
shard across ranks, gradients are averaged, and the result equals the single-process gradient of
the batch-mean loss.  The model on each rank is the CPU oracle restatement (the sm_100a modules
have no CPU path by design)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vln_ver_b200 import dist_utils, synth
from vln_ver_b200.config import per_voxel_occupancy_size

GRID, NCAM, C, B = (2, 4, 4), 6, 64, 4


def _setup():
    import vln_ver_b200 as V
    torch.manual_seed(0)
    cfg = V.vocc_head_cfg(*GRID, num_cams=NCAM, embed_dims=C, only_occ=True, refine_occ=False,
                          occupancy_size=per_voxel_occupancy_size(*GRID), num_layers=1, occ_dims=16)
    head = V.build_head(cfg)
    head.init_weights()
    l2i, sh = synth.make_rig(B, NCAM, GRID, seed=4)
    feats = torch.from_numpy(synth.make_features(B, NCAM, dim=C, seed=5))
    gts = [torch.from_numpy(x) for x in synth.make_occ_gt(B, head.voxel_num, frac=0.3)]
    return head, torch.from_numpy(l2i), torch.from_numpy(sh), feats, gts


def _loss(sd, head, feats, l2i, sh, gts):
    from oracle import ver_ref
    bev = ver_ref.get_voxel_features(sd, 'transformer.', feats, sd['voxel_embedding.weight'], *GRID,
                                     synth.PC_RANGE, l2i, sh, num_layers=1)
    occ = ver_ref.occ_head(sd, '', bev, *GRID, head.occ_xdim, head.occ_ydim, head.occ_zdim, occ_dims=16,
                           refine_occ=False, only_occ=True)
    return ver_ref.occupancy_loss(occ, gts)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    head, l2i, sh, feats, gts = _setup()
    lo, hi = dist_utils.shard_range(B, rank, world)
    sd = {k: v.detach().clone().double().requires_grad_(True) for k, v in head.state_dict().items()}
    loss = _loss(sd, head, feats[:, lo:hi].double(), l2i[lo:hi], sh[lo:hi], gts[lo:hi])
    loss.backward()
    names = [k for k, v in sd.items() if v.grad is not None]
    local = {k: sd[k].grad.detach().float().clone() for k in names}      # this rank's gradient, before any exchange
    # the flat bucket of the graphed step (fp32 master gradients as views into one buffer, one collective)
    p32 = [torch.nn.Parameter(sd[k].detach().float()) for k in names]
    bucket = dist_utils.FlatGradients(p32)
    bucket.zero_()
    for p, k in zip(p32, names):
        p.grad += sd[k].grad.float()              # accumulate into the views, as autograd does
    bucket.allreduce_mean_()
    flat_ok = all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in p32)
    dist_utils.allreduce_mean_([sd[k].grad for k in names])
    for p, k in zip(p32, names):                  # same mean through both paths (fp32 vs fp64)
        assert flat_ok and torch.allclose(p.grad.double(), sd[k].grad, rtol=1e-5, atol=1e-7), k
    # the bucketed, overlapped form of the same exchange: every bucket is reduced from autograd's post-accumulate
    # hooks as soon as its last gradient has arrived (bench.py's multi-GPU step); same mean
    q32 = [torch.nn.Parameter(sd[k].detach().float()) for k in names]
    over = dist_utils.FlatGradients(q32, bucket_bytes=4096, overlap=True)
    assert over.overlap and len(over.buckets) > 2 and over.buckets[0][0] == 0 and over.buckets[-1][1] == over.flat.numel()
    for _ in range(2):                            # twice: the per-step bookkeeping resets
        over.zero_()
        sum((q * local[k]).sum() for q, k in zip(q32, names)).backward()     # d/dq = this rank's gradient
        assert any(over.launched)                 # hooks fired during backward
        over.allreduce_mean_()
        assert not any(over.launched) and over.pending == [n for _, _, n in over.buckets]
        for q, p in zip(q32, p32):
            assert torch.allclose(q.grad, p.grad, rtol=1e-6, atol=1e-8)
    ms = dist_utils.max_over_ranks(10.0 + rank)
    if rank == 0:
        torch.save({'grads': {k: sd[k].grad for k in names}, 'ms': ms}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [dist_utils.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gradient_mean_equals_full_batch(tmp_path):
    out = str(tmp_path / 'r0.pt')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got['ms'] == 11.0                       # max over ranks
    head, l2i, sh, feats, gts = _setup()
    sd = {k: v.detach().clone().double().requires_grad_(True) for k, v in head.state_dict().items()}
    _loss(sd, head, feats.double(), l2i, sh, gts).backward()
    for k, g in got['grads'].items():
        ref = sd[k].grad
        assert torch.allclose(g, ref, rtol=1e-9, atol=1e-12), k


def test_flat_gradient_buckets_partition_the_buffer():
    """FlatGradients without a process group: gradients are views of one flat buffer, the buckets are contiguous,
    cover it exactly and count every parameter once; without a group allreduce_mean_() is the identity."""
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(n)) for n in (1000, 10, 300, 5000, 7, 64)]
    fg = dist_utils.FlatGradients(params, bucket_bytes=4000, overlap=True)
    assert not fg.overlap and fg.hooks == []                       # no process group: nothing to overlap
    assert fg.buckets[0][0] == 0 and fg.buckets[-1][1] == fg.flat.numel() == sum(p.numel() for p in params)
    assert all(a[1] == b[0] for a, b in zip(fg.buckets, fg.buckets[1:]))
    assert sum(n for _, _, n in fg.buckets) == len(params)
    assert all(fg.buckets[fg.bucket_of[i]][0] <= off < fg.buckets[fg.bucket_of[i]][1]
               for i, off in enumerate([0, 1000, 1010, 1310, 6310, 6317]))
    fg.zero_()
    sum((p * (i + 1.0)).sum() for i, p in enumerate(params)).backward()
    for i, p in enumerate(params):
        assert p.grad.data_ptr() >= fg.flat.data_ptr() and torch.all(p.grad == i + 1.0)
    before = fg.flat.clone()
    assert fg.allreduce_mean_() is fg.flat and torch.equal(fg.flat, before)
    one = dist_utils.FlatGradients([torch.nn.Parameter(torch.randn(5))])
    assert len(one.buckets) == 1
