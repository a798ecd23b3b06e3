#!/usr/bin/env python
"""Per-phase SM-clock breakdown of sca_fwd_tc_kernel (debug timers compiled into the kernel, enabled
through the undocumented ver_debug_tc_timing hook).  Run on the GPU box:  python tools/tc_timing.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import _lib, ops, synth  # noqa: E402

NAMES = {0: 'setup (barriers, tmem alloc, ids, lists)', 1: 'logits -> softmax/offsets', 2: 'B: refs + wait mma retire',
         3: 'B: zero + bar.sync', 4: 'B: taps (atomics)', 5: 'B: fences + arrive', 6: 'B: drain', 7: 'epilogue',
         8: 'M: prologue idle', 9: 'M: wait built', 10: 'M: wait V', 11: 'M: issue', 12: 'epilogue: wait last MMAs', 13: 'B: tap arithmetic',
         16: 'bwd: zero A + ids', 17: 'bwd: G gather', 18: 'bwd: A rows', 19: 'bwd: fence+MMA', 20: 'bwd: dots dump+taps',
         21: 'bwd: softmax bwd + atomics', 22: 'bwd: dV write'}


def main():
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
    # ---- sorted-row forward: sca_fwd_tc4_kernel (A operand in TMEM) and sca_fwd_tc3_kernel (A image in smem)
    def hook(name):
        f = getattr(_lib.lib, name)
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
        return f
    names3 = {0: 'W: setup', 1: 'W: item top (logits, softmax)', 4: 'W: tap arithmetic', 2: 'W: wait MMA retire',
              3: 'W: un-tap', 5: 'W: taps (smem RMW)', 6: 'W: fences + arrive', 7: 'W: epilogue wait MMA',
              8: 'W: epilogue TMEM->slots', 9: 'C: wait built', 10: 'C: wait V', 11: 'C: MMA issue',
              12: 'C: V buffer wait + TMA'}
    names4 = {0: 'W: setup', 1: 'W: item top (slot, prefetch, softmax)', 4: 'W: taps (arithmetic + RMW)',
              2: 'W: wait MMA retire', 5: 'W: copy scratch->TMEM + zero', 6: 'W: fences + arrive', 7: 'W: item end',
              9: 'C: wait built', 10: 'C: wait V', 13: 'C: wait drained accumulator', 11: 'C: MMA issue',
              12: 'C: V buffer wait + TMA', 16: 'E: wait full accumulator', 17: 'E: TMEM->slots'}
    order, smask, tu = vis.order
    ncams = sum(bin(int(x) & 0xffffffff).count('1') for x in tu.flatten().tolist())
    hits = int(count.sum().item())
    print(f'sorted tiles: {tu.numel()} tiles, {ncams} (tile, camera) products = {ncams * 128 / hits:.3f} x hits '
          f'({hits} hits, {hits / (B * Nq):.3f} per voxel)')
    vimg = ops.value_image(value, NH)
    slots = torch.empty((B, Nq, NH * Dh), dtype=torch.float16, device='cuda')

    cur = [0]

    def launch3():
        _lib.check(_lib.lib.ver_sca_forward_sorted(vimg.data_ptr(), logits.data_ptr(), 192, rpc.data_ptr(),
                                                   order.data_ptr(), smask.data_ptr(), tu.data_ptr(),
                                                   slots.data_ptr(), B, ncam, Nq, 14, 14, NH, Dh, 8, cur[0],
                                                   torch.cuda.current_stream().cuda_stream))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ref_out = None
    names5 = {0: 'W: setup', 1: 'W: item top (slot, prefetch, softmax)', 4: 'W: taps (arithmetic + RMW)',
              2: 'W: wait operand free (MMA J-3)', 5: 'W: copy issue scratch->TMEM + un-tap', 6: 'W: st drain + arrive',
              7: 'W: tail', 9: 'C: wait built', 10: 'C: wait V', 13: 'C: wait drained accumulator', 11: 'C: MMA issue',
              12: 'C: V buffer wait + TMA', 16: 'E: wait full accumulator', 17: 'E: TMEM->slots'}
    variants = [(5, 'sca_fwd_tc5_kernel', 'ver_debug_tc5_timing', names5),
                (4, 'sca_fwd_tc4_kernel', 'ver_debug_tc4_timing', names4)]
    if '--tc3' in sys.argv:
        variants.append((3, 'sca_fwd_tc3_kernel', 'ver_debug_tc3_timing', names3))
    if '--bwd-only' in sys.argv:
        variants = []
    for variant, kname, tname, names in variants:
        cur[0] = variant
        f3 = hook(tname)
        for _ in range(3):
            launch3()
        torch.cuda.synchronize()
        if ref_out is None:
            ref_out = slots.clone()
        else:
            d = (slots.float() - ref_out.float()).abs().max().item() / ref_out.float().abs().max().item()
            print(f'  max |this - tc5| / max |tc5| = {d:.2e}')
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ts = []
        for i in range(10):
            flush.zero_()
            ev[0].record()
            launch3()
            ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        ts.sort()
        print(f'{kname} alone (L2 flushed): median {ts[5] * 1e3:.1f} us, min {ts[0] * 1e3:.1f} us')
        assert f3(1, None) == 0
        launch3()
        torch.cuda.synchronize()
        out3 = (ctypes.c_ulonglong * 32)()
        assert f3(0, out3) == 0
        nct = min(148, B * NH * ((Nq + 255) // 256))
        print(f'{nct} persistent CTAs, cycles per CTA (thread 0 of group 0 / control thread / epilogue warp 8):')
        for i in sorted(names):
            print(f'  [{i:2d}] {names[i]:40s} {out3[i] / nct:10.0f}')
        print(f'  worker total {sum(out3[i] for i in range(9)) / nct:.0f}, control total '
              f'{sum(out3[i] for i in range(9, 14)) / nct:.0f}')
    if '--fwd-only' in sys.argv:
        return
    # ---- backward: sca_bwd_tc2_kernel (two threads per hit, lane-interleaved Dots staging) vs sca_bwd_tc_kernel
    names_b2 = {16: 'un-tap + hit ids', 17: 'G image stores, Dots MMA issue', 18: "A' rows",
                19: 'fences, dV^T MMA issue, wait Dots MMAs', 20: 'Dots dump', 21: 'tap read-back, softmax bwd, atomics',
                23: 'wait dV^T MMAs', 22: 'dV^T -> grad_value'}
    _lib.lib.ver_debug_bwd_variant.restype = ctypes.c_int
    _lib.lib.ver_debug_bwd_variant.argtypes = [ctypes.c_int]
    counts, index = vis.index
    gs = torch.randn(B, Nq, NH * Dh, device='cuda', generator=g).half()
    gvalue = torch.empty(B * ncam, 196, NH * Dh, device='cuda')
    glogits = torch.empty_like(logits)

    def launch_bwd():
        _lib.check(_lib.lib.ver_sca_backward(_lib.VER_F16, vimg.data_ptr(), ops.VER_LAYOUT_TC_IMAGE, logits.data_ptr(), 192,
                                             rpc.data_ptr(), bits.data_ptr(), counts.data_ptr(), index.data_ptr(),
                                             gs.data_ptr(), gvalue.data_ptr(), glogits.data_ptr(), B, ncam, *grid, 14, 14,
                                             NH, Dh, 8, torch.cuda.current_stream().cuda_stream))
    ref_b = None
    for variant, kname in ((0, 'sca_bwd_tc2_kernel'), (1, 'sca_bwd_tc_kernel')):
        _lib.lib.ver_debug_bwd_variant(variant)
        for _ in range(2):
            launch_bwd()
        torch.cuda.synchronize()
        if ref_b is None:
            ref_b = (gvalue.clone(), glogits.clone())
        else:
            dv = (gvalue - ref_b[0]).abs().max().item() / ref_b[0].abs().max().item()
            dl = (glogits - ref_b[1]).abs().max().item() / ref_b[1].abs().max().item()
            print(f'  max-norm difference old vs new backward: grad_value {dv:.2e}, grad_logits {dl:.2e}')
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ts = []
        for i in range(8):
            flush.zero_()
            ev[0].record()
            launch_bwd()
            ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        ts.sort()
        print(f'{kname} alone (L2 flushed, includes the glogits memset): median {ts[4] * 1e3:.1f} us, min {ts[0] * 1e3:.1f} us')
        if variant == 0:
            fb = hook('ver_debug_bwd2_timing')
            assert fb(1, None) == 0
            launch_bwd()
            torch.cuda.synchronize()
            outb = (ctypes.c_ulonglong * 32)()
            assert fb(0, outb) == 0
            bc = B * ncam * NH
            print(f'  {bc} CTAs, cycles per CTA (thread 0):')
            for i in (16, 18, 17, 19, 20, 21, 23, 22):
                print(f'  [{i:2d}] {names_b2[i]:42s} {outb[i] / bc:10.0f}')
    _lib.lib.ver_debug_bwd_variant(0)
    if '--no-legacy' in sys.argv or '--bwd-only' in sys.argv:
        return
    ops.TC_FORWARD = 'block'
    fn = _lib.lib.ver_debug_tc_timing
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
    for _ in range(2):
        ops.sca_sample_tc(value, logits, vis, 14, 14, NH, 8)
    torch.cuda.synchronize()
    assert fn(1, None) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.sca_sample_tc(value, logits, vis, 14, 14, NH, 8)
    e1.record()
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 32)()
    assert fn(0, out) == 0
    ctas = (Nq // 256) * NH * B
    print(f'launch {e0.elapsed_time(e1):.3f} ms (includes value_image), {ctas} CTAs, cycles per CTA:')
    tot_b = sum(out[i] for i in range(8))
    for i in range(14):
        print(f'  [{i:2d}] {NAMES[i]:42s} {out[i] / ctas:10.0f}')
    print(f'  builder thread total {tot_b / ctas:.0f} cycles/CTA')
    # backward
    v = value.clone().requires_grad_(True)
    lg = logits.clone().requires_grad_(True)
    o = ops.sca_sample_tc(v, lg, vis, 14, 14, NH, 8)
    go = torch.randn_like(o)
    o.backward(go, retain_graph=True)
    torch.cuda.synchronize()
    assert fn(1, None) == 0
    e0.record()
    o.backward(go)
    e1.record()
    torch.cuda.synchronize()
    assert fn(0, out) == 0
    bc = B * ncam * NH
    print(f'backward {e0.elapsed_time(e1):.3f} ms (includes casts), {bc} CTAs, cycles per CTA:')
    for i in range(16, 23):
        print(f'  [{i:2d}] {NAMES[i]:42s} {out[i] / bc:10.0f}')


if __name__ == '__main__':
    main()
