#!/usr/bin/env python
"""Benchmark of the VER 2D->3D lift+encode hot path (BASELINE.json metric:
panoramas/sec, 18-view -> voxel lift+encode; HBM GB/s of the fused sampler).

    python bench.py --gpus N --steps K --warmup W            # sm_100a arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json configs[1]): vocc.py-shaped forward + backward + AdamW step, batch 8
panoramas per GPU, 18 views x 196 tokens x 768, 16x40x40 voxels, fp16 storage / fp32 accumulate,
synthetic features + random occupancy labels, per-voxel occupancy head (SURVEY.md 8(d) "head
consistency").  N > 1: one rank per GPU, panoramas sharded, DDP gradient all-reduce (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = (16, 40, 40)
NCAM = 18
PER_GPU_BATCH = 8
EMBED = 768
LOSS_SCALE = 1024.0
METRIC = 'panoramas/sec (18-view->voxel lift+encode, fwd+bwd+optimizer step)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=PER_GPU_BATCH, help='panoramas per GPU')
    ap.add_argument('--grid', type=int, nargs=3, default=list(GRID), metavar=('Z', 'H', 'W'))
    ap.add_argument('--mode', default='train', choices=['train', 'infer'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def workload_name(args):
    z, h, w = args.grid
    what = 'fwd+bwd+AdamW' if args.mode == 'train' else 'inference fwd'
    return (f'vocc.py {what}, batch={args.batch}/GPU, {NCAM} views x196x{EMBED}, {h}x{w}x{z} voxels, '
            f'fp16 storage/fp32 accumulate, per-voxel occ head, synthetic')


# ----------------------------------------------------------------------------- model
def perturb(head, seed=101):
    g = torch.Generator().manual_seed(seed)
    for n, p in head.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            with torch.no_grad():
                p.add_(torch.randn(p.shape, generator=g) * 0.02)


def build_model(grid):
    import vln_ver_b200 as V
    from vln_ver_b200.config import per_voxel_occupancy_size
    torch.manual_seed(0)
    cfg = V.vocc_head_cfg(*grid, num_cams=NCAM, embed_dims=EMBED, only_occ=True, refine_occ=False,
                          occupancy_size=per_voxel_occupancy_size(*grid))
    head = V.build_head(cfg)
    head.init_weights()
    perturb(head)
    # parameters that never receive a gradient on this path are excluded statically
    # (the reference needs find_unused_parameters=True for them, mmdet_train.py:76-80)
    for n, p in head.named_parameters():
        if n.startswith('positional_encoding'):
            p.requires_grad_(False)
    return head


def make_batches(n_batches, batch, grid, rank, voxel_num):
    from vln_ver_b200 import synth
    out = []
    for i in range(n_batches):
        seed = 1000 * rank + 10 * i
        l2i, sh = synth.make_rig(batch, NCAM, grid, seed=1235 + seed)
        feats = synth.make_features(batch, NCAM, dim=EMBED, seed=1234 + seed)
        gts = synth.make_occ_gt(batch, voxel_num, seed=1236 + seed)
        out.append(dict(feats=torch.from_numpy(feats), l2i=torch.from_numpy(l2i), sh=torch.from_numpy(sh),
                        gts=torch.from_numpy(np.stack(gts))))
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                    '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(',')]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == 'Active' for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                'reasons': reasons, 'samples': len(self.samples)}


# ----------------------------------------------------------------------------- sm_100a arm
def run_b200(args):
    import torch.distributed as dist
    import vln_ver_b200 as V
    from vln_ver_b200 import ops

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, (world, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    ddp = world > 1
    if ddp:
        dist.init_process_group('nccl', device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    grid = tuple(args.grid)
    Nq = grid[0] * grid[1] * grid[2]
    train = args.mode == 'train'

    head = build_model(grid).to(dev)
    V.set_compute_dtype(head, torch.float16)
    model = head
    if train:
        head.train()
        if ddp:
            model = torch.nn.parallel.DistributedDataParallel(
                head, device_ids=[local], gradient_as_bucket_view=True, static_graph=True)
        params = [p for p in head.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True)   # vocc.py:261-268
    else:
        head.eval()

    n_pool = 3
    pool_host = make_batches(n_pool, args.batch, grid, rank, head.voxel_num)
    from vln_ver_b200.ingest import pin
    pool_host = [pin(b) for b in pool_host]
    pool_dev = [{k: v.to(dev) for k, v in b.items()} for b in pool_host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in pool_host[0].values())

    sampler_events = []

    def step(batch, timed_sampler=False):
        """one pass of the hot path over one batch; returns the loss tensor (device)."""
        ops.PROFILE_EVENTS = sampler_events if timed_sampler else None
        if not train:
            with torch.no_grad():
                outs = model(batch['feats'], None, lidar2img=batch['l2i'], originshift=batch['sh'])
            return outs['occupancy_preds'].float().mean()
        outs = model(batch['feats'], None, lidar2img=batch['l2i'], originshift=batch['sh'])
        loss = head.loss_only_occupancy(None, None, None, list(batch['gts']), None, outs)['loss_occupancy']
        opt.zero_grad(set_to_none=True)
        (loss * LOSS_SCALE).backward()
        torch._foreach_mul_([p.grad for p in params if p.grad is not None], 1.0 / LOSS_SCALE)
        torch.nn.utils.clip_grad_norm_(params, 300.0)                                # vocc.py:270
        opt.step()
        return loss.detach()

    def barrier():
        if ddp:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_one, steps, warmup):
        for i in range(warmup):
            run_one(i)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = V.launch_count()
        t0.record()
        for i in range(steps):
            run_one(warmup + i)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if ddp:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), V.launch_count() - n0

    # ---- (1) device-resident inputs: `value`
    with ClockSampler(local) as clocks:
        ms_dev, launches = timed(lambda i: step(pool_dev[i % n_pool], timed_sampler=True),
                                 args.steps, args.warmup)
        clk = clocks.summary()
    ops.PROFILE_EVENTS = None
    # sampler forward duration from the events recorded inside the timed steps
    torch.cuda.synchronize()
    ev = sampler_events[-3 * args.steps:] if len(sampler_events) >= 3 * args.steps else sampler_events
    sampler_ms = float(np.mean([a.elapsed_time(b) for a, b in ev])) if ev else None

    # ---- (2) end to end through the public API with HOST buffers: `e2e`
    # every step's inputs start in pinned host memory and are copied inside the timed region (exactly K
    # copies for K steps); vln_ver_b200.ingest.DevicePrefetcher keeps the copy of step i+1 on a side stream
    # under the compute of step i.  Every step's loss is read device -> host inside the region.
    from vln_ver_b200.ingest import DevicePrefetcher

    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_pass(first, n):
        """n steps from host batches; every step's loss is read back on the host through a pinned buffer
        + event, one step behind the launch front (the host enqueues step i+1 while step i runs)."""
        pf = DevicePrefetcher((pool_host[(first + j) % n_pool] for j in range(n)), dev)
        pending, last = None, None
        for i, batch in enumerate(pf):
            loss = step(batch)
            buf = loss_host[i % 2]
            buf.copy_(loss.detach().float().reshape(1), non_blocking=True)      # device -> host, 4 bytes
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[0].synchronize()
                last = pending[1].item()
            pending = (ev, buf)
        if pending is not None:
            pending[0].synchronize()
            last = pending[1].item()
        return pf.h2d_bytes

    e2e_pass(0, args.warmup)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    copied = e2e_pass(args.warmup, args.steps)
    t1.record()
    barrier()
    assert copied == h2d_bytes * args.steps
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if ddp:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_e2e = ms.item()

    pano = args.batch * world * args.steps
    value = pano / (ms_dev / 1e3)
    e2e_value = pano / (ms_e2e / 1e3)

    result = None
    if rank == 0:
        # ---- roofline of the fused forward sampler (dominant HBM-side kernel of the lift)
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        counts = [int(ops.point_sampling(b['l2i'], b['sh'], head.transformer.encoder.pc_range, *grid)[3]
                      .sum().item()) for b in pool_dev]
        P = float(np.mean(counts))                       # visible (b, cam, voxel) pairs per launch
        B = args.batch
        bv = 2
        bytes_min = (B * NCAM * 208 * EMBED * bv        # value operand images (196 px padded to 208), read once
                     + B * Nq * 192 * 4                 # offset/weight logits, once per VOXEL
                     + P * 8 + B * Nq * 4               # reference points per hit, visibility bits
                     + B * Nq * EMBED * bv)             # slots written once
        bytes_survey = (B * NCAM * 196 * EMBED * bv + P * (512 + 256) + P * 4
                        + B * Nq * EMBED * bv + B * Nq * 4)      # SURVEY.md 8(d) fused formula
        roof = None
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'sampler_traffic.json')
        if os.path.exists(tpath) and args.batch == PER_GPU_BATCH and tuple(args.grid) == GRID:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj['dram_bytes_per_launch'], tj['source']
        if sampler_ms:
            ach = bytes_min / (sampler_ms * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'kernel': 'sca_fwd_tc4_kernel<96, 8> (tcgen05 fused SCA sampler forward, visibility-sorted rows, A operand in TMEM)',
                    'achieved': round(ach, 1),
                    'peak': peak, 'unit': 'GB/s', 'frac': round(ach / peak, 4), 'traffic': traffic, 'traffic_source': traffic_src,
                    'peak_source': peak_src, 'launch_ms': round(sampler_ms, 4),
                    'algorithmic_bytes': int(bytes_min),
                    'achieved_survey_formula': round(bytes_survey / (sampler_ms * 1e-3) / 1e9, 1),
                    'hits_per_launch': P,
                    'note': 'algorithmic bytes = compulsory traffic of the per-voxel-logit formulation '
                            '(smaller than SURVEY 8(d) fused formula, also given); timed with CUDA events '
                            'around the launch inside the timed steps'}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference(args, steps=1, warmup=0)
        result = {
            'metric': METRIC, 'value': round(value, 2), 'unit': 'panoramas/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms_dev / args.steps, 3),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16',
            'data': 'synthetic', 'impl': 'b200',
            'config': {'workload': workload_name(args), 'global_batch': args.batch * world,
                       'parallelism': f'dp{world}', 'l2': 'inputs+activations per step (>1 GB) exceed the 126 MB L2; '
                                                         f'{n_pool} distinct batches cycled'},
            'e2e': {'value': round(e2e_value, 2), 'unit': 'panoramas/s', 'h2d_bytes_per_step': int(h2d_bytes),
                    'd2h_bytes_per_step': 4, 'ms_per_step': round(ms_e2e / args.steps, 3)},
            'gpu_launches': int(launches), 'clocks': clk, 'roofline': roof, 'cpu_baseline': cpu,
        }
        print(json.dumps(result), flush=True)
    if ddp:
        dist.barrier()
        dist.destroy_process_group()
    return result


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference(args, steps, warmup, budget_s=240.0):
    """The reference's CPU path for the same workload, as the oracle PORT (oracle/ver_ref.py: the
    reference's PyTorch-CPU ops restated; /root/reference itself cannot travel to the GPU box).
    One step = fwd+bwd of ONE panorama (bounded sample).  Threads: min(cores, 16) -- measured: the
    op mix (grid_sample on padded rebatches, index_put loops) gets SLOWER beyond that (127 s/step with
    128 threads vs 33 s with 8).  If `steps` would exceed `budget_s` the number of timed steps is cut
    and reported."""
    from oracle import ver_ref
    from vln_ver_b200 import synth
    cores = os.cpu_count() or 1
    threads = min(cores, 16)
    torch.set_num_threads(threads)
    grid = tuple(args.grid)
    head = build_model(grid)
    train = args.mode == 'train'
    sd = {k: v.detach().clone().requires_grad_(train and v.is_floating_point())
          for k, v in head.state_dict().items()}
    batch = make_batches(1, 1, grid, 0, head.voxel_num)[0]

    def one():
        with torch.set_grad_enabled(train):
            bev = ver_ref.get_voxel_features(sd, 'transformer.', batch['feats'], sd['voxel_embedding.weight'],
                                             *grid, synth.PC_RANGE, batch['l2i'], batch['sh'])
            occ = ver_ref.occ_head(sd, '', bev, *grid, head.occ_xdim, head.occ_ydim, head.occ_zdim,
                                   occ_dims=head.occ_dims, refine_occ=False, only_occ=True)
            if train:
                loss = ver_ref.occupancy_loss(occ, [batch['gts'][0]])
                loss.backward()
                for v in sd.values():
                    v.grad = None
    tw = time.perf_counter()
    for _ in range(warmup):
        one()
        if time.perf_counter() - tw > budget_s / 4:       # warm-up is bounded too
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one()
        done += 1
        dt = time.perf_counter() - t0
        if dt / done * (done + 1) > budget_s:
            break
    dt = time.perf_counter() - t0
    return {'value': round(done / dt, 4), 'unit': 'panoramas/s', 'cores': threads, 'host_cores': cores,
            'kind': 'port', 'steps_timed': done,
            'sample': f'{done} step(s) x 1 panorama {"fwd+bwd" if train else "fwd"} of the same model/grid '
                      f'(oracle port of the reference PyTorch-CPU path, fp32, {threads} threads), {dt:.1f} s',
            'ms_per_step': round(dt / done * 1e3, 1)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cpu = cpu_reference(args, steps=args.steps, warmup=args.warmup)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    print(json.dumps({
        'metric': METRIC, 'value': cpu['value'], 'unit': 'panoramas/s', 'n_gpus': world, 'steps': cpu['steps_timed'],
        'warmup': args.warmup, 'ms_per_step': cpu['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': workload_name(args), 'global_batch': 1, 'parallelism': 'cpu'},
        'cpu_baseline': cpu, 'gpu_launches': 0,
        'e2e': {'value': cpu['value'], 'unit': 'panoramas/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}}), flush=True)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
