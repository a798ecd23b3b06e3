#!/usr/bin/env python
"""Bring-up check of sca_fwd_tc6_kernel on the GPU box: tc6 against tc4 on small and full-size problems, with the
kernel's wait watchdog in no-trap mode so that a protocol error is reported (ver_debug_tc6) instead of hanging or
killing the context; phase timers of the full-size launch.   timeout -s KILL 180 python tools/tc6_check.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import _lib, ops, synth  # noqa: E402

L = _lib.lib
for _n in ('ver_debug_tc6', 'ver_debug_tc7'):
    getattr(L, _n).restype = ctypes.c_int
    getattr(L, _n).argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_uint)]
CODES = {1: 'control: built', 2: 'control: V landed', 3: 'control: accumulator drained', 4: 'control: V buffer free',
         5: 'control: drain', 6: 'epilogue: accumulator full', 7: 'builder: batch retired'}
PHASES = {0: 'B: setup', 1: 'B: item top', 2: 'B: wait MMA retire', 3: 'B: un-tap (tc6) / copy + un-tap (tc7)', 4: 'B: taps', 6: 'B: fences + arrive',
          7: 'B: item end', 9: 'C: wait built', 10: 'C: wait V', 11: 'C: MMA issue', 12: 'C: V buffer wait + TMA',
          13: 'C: wait drained accumulator', 16: 'E: wait full accumulator', 17: 'E: TMEM->slots'}


def debug(flags, tag=None, which=('ver_debug_tc6', 'ver_debug_tc7')):
    """set the flags of both kernels; returns (aborted, timers of the LAST kernel in `which`) of what ran since the last call"""
    torch.cuda.synchronize()
    aborted, timers = 0, None
    for name in which:
        t = (ctypes.c_ulonglong * 32)()
        d = (ctypes.c_uint * 8)()
        a = getattr(L, name)(flags, t, d)
        if a and tag:
            print(f'{tag} [{name}]: WAIT TIMED OUT: {CODES.get(d[0], d[0])}, block {d[1]}, thread {d[2]}, words {d[3]} {d[4]}',
                  flush=True)
        aborted |= a
        timers = list(t)
    return aborted, timers


def case(B, grid, Dh, time_it, sh=14, sw=14, seed=1235):
    ncam, NH = 18, 8
    Nq = grid[0] * grid[1] * grid[2]
    S = sh * sw
    l2i, shf = synth.make_rig(B, ncam, grid, seed=seed)
    rpc, mask, bits, count = ops.point_sampling(torch.from_numpy(l2i).cuda(), torch.from_numpy(shf).cuda(),
                                                synth.PC_RANGE, *grid)
    vis = ops.Visibility(rpc, mask, bits, count, grid)
    g = torch.Generator(device='cuda').manual_seed(seed)
    value = (torch.randn(B * ncam, S, NH * Dh, device='cuda', generator=g) * 0.5).half()
    logits = torch.randn(B * Nq, 192, device='cuda', generator=g)
    logits[:, :128] *= 2
    order, smask, tu = vis.order
    vimg = ops.value_image(value, NH)
    vimg16 = ops.value_image16(value, NH, sh, sw)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream

    def run4(slots):
        _lib.check(L.ver_sca_forward_sorted(vimg.data_ptr(), logits.data_ptr(), 192, rpc.data_ptr(), order.data_ptr(),
                                            smask.data_ptr(), tu.data_ptr(), slots.data_ptr(), B, ncam, Nq, sh, sw, NH,
                                            Dh, 8, 4, st))

    def run16(variant):
        def run(slots):
            _lib.check(L.ver_sca_forward_sorted16(vimg16.data_ptr(), logits.data_ptr(), 192, rpc.data_ptr(),
                                                  order.data_ptr(), smask.data_ptr(), tu.data_ptr(), slots.data_ptr(), B,
                                                  ncam, Nq, sh, sw, NH, Dh, 8, variant, st))
        return run
    run6, run7 = run16(6), run16(7)
    outs = {}
    for name, run in (('tc4', run4), ('tc7', run7), ('tc6', run6)):
        slots = torch.full((B, Nq, NH * Dh), float('nan'), dtype=torch.float16, device='cuda')
        run(slots)
        if debug(2, f'B={B} grid={grid} Dh={Dh} {name}')[0]:
            return False
        outs[name] = slots.clone()
        if time_it:
            ts = []
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for _ in range(8):
                flush.zero_()
                ev[0].record()
                run(slots)
                ev[1].record()
                torch.cuda.synchronize()
                ts.append(ev[0].elapsed_time(ev[1]))
            ts.sort()
            print(f'  {name}: median {ts[4] * 1e3:.1f} us, min {ts[0] * 1e3:.1f} us', flush=True)
            if debug(2, 'timing loop')[0]:
                return False
    a = outs['tc4'].float()
    ok = True
    for name in ('tc7', 'tc6'):
        b = outs[name].float()
        nan = int(torch.isnan(b).sum().item())
        d = ((a - b).abs().max() / a.abs().max()).item()
        print(f'B={B} grid={grid} Dh={Dh} map {sh}x{sw}: max |{name} - tc4| / max |tc4| = {d:.2e}, NaNs: {nan}', flush=True)
        ok &= nan == 0 and d < 2e-3
    print(f'  tc7 == tc6 bit for bit: {torch.equal(outs["tc7"], outs["tc6"])}')
    if time_it:
        for name, run, hook in (('tc7', run7, 'ver_debug_tc7'), ('tc6', run6, 'ver_debug_tc6')):
            debug(3)
            slots = torch.empty((B, Nq, NH * Dh), dtype=torch.float16, device='cuda')
            flush.zero_()
            run(slots)
            _, t = debug(2, which=(hook,))
            print(f'  {name} phase timers (SM cycles summed over 148 CTAs / 148):')
            for k, pname in PHASES.items():
                print(f'    [{k:2d}] {pname:40s} {t[k] / 148:10.0f}')
    return ok


def main():
    assert L.ver_debug_tc6(2, None, None) == 0 and L.ver_debug_tc7(2, None, None) == 0      # watchdog: report, do not trap
    ok = True
    ok &= case(1, (3, 5, 7), 32, False)
    ok &= case(2, (8, 20, 20), 96, False)
    ok &= case(1, (3, 11, 13), 64, False, seed=7)
    ok &= case(2, (4, 9, 9), 32, False, sh=7, sw=10, seed=11)
    ok &= case(1, (4, 9, 9), 64, False, sh=9, sw=5, seed=13)
    if ok:
        ok &= case(8, (16, 40, 40), 96, True)
    print('tc6_check:', 'OK' if ok else 'FAILED', flush=True)
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
