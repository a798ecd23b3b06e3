#!/usr/bin/env python
"""Bring-up check of sca_fwd_tc5_kernel on the GPU box: tc5 against tc4 on a small and the full-size problem,
with the kernel's wait watchdog in no-trap mode so that a protocol error is reported (ver_debug_tc5_diag)
instead of hanging or killing the context.   timeout -s KILL 120 python tools/tc5_check.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import _lib, ops, synth  # noqa: E402

L = _lib.lib
L.ver_debug_tc5_timing.restype = ctypes.c_int
L.ver_debug_tc5_timing.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
L.ver_debug_tc5_diag.restype = ctypes.c_int
L.ver_debug_tc5_diag.argtypes = [ctypes.POINTER(ctypes.c_uint)]
CODES = {1: 'control: built', 2: 'control: V landed', 3: 'control: accumulator drained', 4: 'control: V buffer free',
         5: 'control: drain', 6: 'epilogue: accumulator full', 7: 'worker: operand retired'}


def diag(tag):
    torch.cuda.synchronize()
    out = (ctypes.c_uint * 8)()
    aborted = L.ver_debug_tc5_diag(out)
    if aborted:
        print(f'{tag}: WAIT TIMED OUT: {CODES.get(out[0], out[0])}, block {out[1]}, thread {out[2]}, '
              f'words {out[3]} {out[4]}', flush=True)
    return aborted


def case(B, grid, Dh, time_it):
    ncam, NH = 18, 8
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
    outs = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for variant in (4, 5):
        slots = torch.full((B, Nq, NH * Dh), float('nan'), dtype=torch.float16, device='cuda')

        def launch():
            _lib.check(L.ver_sca_forward_sorted(vimg.data_ptr(), logits.data_ptr(), 192, rpc.data_ptr(),
                                                order.data_ptr(), smask.data_ptr(), tu.data_ptr(), slots.data_ptr(),
                                                B, ncam, Nq, 14, 14, NH, Dh, 8, variant,
                                                torch.cuda.current_stream().cuda_stream))
        launch()
        if diag(f'B={B} grid={grid} Dh={Dh} variant {variant}'):
            return False
        outs[variant] = slots.clone()
        if time_it:
            ts = []
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for _ in range(8):
                flush.zero_()
                ev[0].record()
                launch()
                ev[1].record()
                torch.cuda.synchronize()
                ts.append(ev[0].elapsed_time(ev[1]))
            ts.sort()
            print(f'  variant {variant}: median {ts[4] * 1e3:.1f} us, min {ts[0] * 1e3:.1f} us', flush=True)
            if diag('timing loop'):
                return False
    a, b = outs[4].float(), outs[5].float()
    nan = int(torch.isnan(b).sum().item())
    d = ((a - b).abs().max() / a.abs().max()).item()
    print(f'B={B} grid={grid} Dh={Dh}: max |tc5 - tc4| / max |tc4| = {d:.2e}, NaNs in tc5 output: {nan}', flush=True)
    return nan == 0 and d < 1e-3


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'knobs':
        # bottleneck experiments (results invalid): 2 = no trap, +4 = no MMAs issued, +8 = no tcgen05.st
        for flags in (2, 6, 10, 14):
            assert L.ver_debug_tc5_timing(flags, None) == 0
            print('debug flags', flags, flush=True)
            case(8, (16, 40, 40), 96, True)
        return
    assert L.ver_debug_tc5_timing(2, None) == 0          # watchdog: report, do not trap
    ok = True
    for B, grid, Dh, t in ((1, (3, 5, 7), 32, False), (2, (8, 20, 20), 96, False), (3, (3, 11, 13), 64, False),
                           (8, (16, 40, 40), 96, True)):
        ok = case(B, grid, Dh, t) and ok
        if not ok:
            break
    print('tc5_check:', 'OK' if ok else 'FAILED', flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
