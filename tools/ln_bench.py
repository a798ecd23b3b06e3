"""Micro-benchmark of the row kernels of the training path (CUDA events, inputs larger than L2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import fused_layer as F, ops
from vln_ver_b200._lib import VER_F16, check, lib

import ctypes
lib.ver_debug_ln_bwd_tma.restype = ctypes.c_int
lib.ver_debug_ln_bwd_tma.argtypes = [ctypes.c_int]
R, C, FF = 204800, 768, 1536
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(R, C, device=dev, generator=g).half()
res = torch.randn(R, C, device=dev, generator=g).half()
dy = torch.randn(R, C, device=dev, generator=g).half()
gam = torch.ones(C, device=dev)
bet = torch.zeros(C, device=dev)
h = torch.randn(R, FF, device=dev, generator=g).half()
dh = torch.randn(R, FF, device=dev, generator=g).half()


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for p in (0.0, 0.1):
    y, z, st, bits = F._ln_fwd(x, res, gam, bet, p, 1e-5, 1234, True)
    t = timeit(lambda: F._ln_fwd(x, res, gam, bet, p, 1e-5, 1234, True))
    print(f'dropout_add_ln_fwd p={p}: {t:.1f} us  ({4 * R * C * 2 / t / 1e3:.0f} GB/s incl. allocation of outputs)')
    outs = {}
    for tma in (1, 0):
        lib.ver_debug_ln_bwd_tma(tma)
        outs[tma] = F._ln_bwd(dy, z, st, gam, p, 1234, bits)
        t = timeit(lambda: F._ln_bwd(dy, z, st, gam, p, 1234, bits))
        if bits is not None:
            t2 = timeit(lambda: F._ln_bwd(dy, z, st, gam, p, 1234, None))
            print(f'  (mask regenerated instead of read: {t2:.1f} us)')
        print(f'dropout_add_ln_bwd p={p} {"TMA ring" if tma else "register loads"}: {t:.1f} us  '
              f'({4 * R * C * 2 / t / 1e3:.0f} GB/s incl. partial sums + fold)')
    lib.ver_debug_ln_bwd_tma(1)
    print('  TMA-fed == register-fed:', all(torch.equal(a, b) for a, b in zip(outs[1], outs[0])))
    hh = h.clone()
    t = timeit(lambda: check(lib.ver_relu_dropout_fwd(VER_F16, hh.data_ptr(), hh.data_ptr(), hh.numel(), p, 99, None,
                                                      torch.cuda.current_stream().cuda_stream)))
    print(f'relu_dropout_fwd p={p}: {t:.1f} us  ({2 * R * FF * 2 / t / 1e3:.0f} GB/s)')
    d2 = dh.clone()
    t = timeit(lambda: F._relu_dropout_bwd_(d2, h, p))
    print(f'relu_dropout_bwd (+colsum) p={p}: {t:.1f} us  ({3 * R * FF * 2 / t / 1e3:.0f} GB/s)')
a = torch.empty_like(x)
t = timeit(lambda: a.copy_(x))
print(f'torch copy fp16 R x C: {t:.1f} us ({2 * R * C * 2 / t / 1e3:.0f} GB/s)')
