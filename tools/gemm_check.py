#!/usr/bin/env python
"""Bring-up + timing of the tcgen05 Linear (csrc/gemm_tc.cu) against torch (cuBLAS) on the layer's shapes.
    timeout -s KILL 120 python tools/gemm_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vln_ver_b200 import _lib, ops  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def check(M, N, K, epi, p=0.0):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    x = (torch.randn(M, K, device='cuda', generator=g) * 0.5).half()
    w = (torch.randn(N, K, device='cuda', generator=g) * 0.05).half()
    b = torch.randn(N, device='cuda', generator=g)
    out = ops.linear_tc(x, w, b, epilogue=epi, p=p, seed=1234)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    if epi == ops.LINEAR_BIAS_RELU_DROPOUT_F16:
        ref = torch.relu(ref)
        if p > 0:
            # the standalone kernel on the bias-only output draws the same mask (same seed, same element index)
            h = ops.linear_tc(x, w, b, epilogue=ops.LINEAR_BIAS_F16)
            _lib.check(_lib.lib.ver_relu_dropout_fwd(_lib.VER_F16, h.data_ptr(), h.data_ptr(), h.numel(), p, 1234,
                                                     ops._seed_epoch(h.device).data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream))
            keep = h != 0                                          # kept AND positive
            pos = ref.half() > 0
            frac = 1.0 - keep[pos].float().mean().item()          # dropped fraction among the positive activations
            assert abs(frac - p) < 0.02, frac
            ref = torch.where(keep, ref / (1 - p), torch.zeros_like(ref))
    e = rel(out, ref)
    print(f'M={M:6d} N={N:4d} K={K:4d} epilogue {epi} p={p}: rel err {e:.2e}', flush=True)
    return e < 2e-3


def bench(M, N, K, epi, name):
    g = torch.Generator(device='cuda').manual_seed(1)
    x = (torch.randn(M, K, device='cuda', generator=g) * 0.5).half()
    w = (torch.randn(N, K, device='cuda', generator=g) * 0.05).half()
    b = torch.randn(N, device='cuda', generator=g)
    b16 = b.half()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def t(fn):
        for _ in range(3):
            fn()
        ts = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    def lib_fn():
        if epi == ops.LINEAR_BIAS_F32:
            return torch.addmm(b, x, w.t(), out_dtype=torch.float32)
        h = torch.addmm(b16, x, w.t())
        if epi == ops.LINEAR_BIAS_RELU_DROPOUT_F16:
            _lib.check(_lib.lib.ver_relu_dropout_fwd(_lib.VER_F16, h.data_ptr(), h.data_ptr(), h.numel(), 0.1, 7,
                                                     None, torch.cuda.current_stream().cuda_stream))
        return h
    t_own = t(lambda: ops.linear_tc(x, w, b, epilogue=epi, p=0.1 if epi == 2 else 0.0, seed=7))
    t_lib = t(lib_fn)
    fl = 2.0 * M * N * K
    print(f'{name:28s} M={M:6d} N={N:4d} K={K:4d}: tcgen05 {t_own * 1e3:7.1f} us ({fl / t_own / 1e9:6.0f} TFLOP/s)   '
          f'library{" + relu_dropout pass" if epi == 2 else ""} {t_lib * 1e3:7.1f} us ({fl / t_lib / 1e9:6.0f} TFLOP/s)',
          flush=True)


def bench_bwd(M, N, K, p=0.1):
    """dX of FFN2 + ReLU / dropout backward + column sums: fused two-CTA GEMM against library GEMM + ver_relu_dropout_bwd"""
    from vln_ver_b200 import fused_layer as FL
    g = torch.Generator(device='cuda').manual_seed(2)
    dy = (torch.randn(M, K, device='cuda', generator=g) * 0.1).half()
    w2 = (torch.randn(K, N, device='cuda', generator=g) * 0.05).half()
    h = (torch.relu(torch.randn(M, N, device='cuda', generator=g)) * (torch.rand(M, N, device='cuda', generator=g) >= p) / (1 - p)).half()
    w2t = w2.t().contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def t(fn):
        for _ in range(3):
            fn()
        ts = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    def own():
        da, part = ops.linear_relu_dropout_bwd(dy, w2t, h, p)
        return FL._fold_rows(part)

    def lib_fn():
        dh = torch.mm(dy, w2)
        return FL._relu_dropout_bwd_(dh, h, p)
    t_own, t_lib = t(own), t(lib_fn)
    fl = 2.0 * M * N * K
    print(f'FFN2 dX + ReLU/dropout bwd + colsum M={M:6d} N={N:4d} K={K:4d}: tcgen05 fused {t_own * 1e3:7.1f} us '
          f'({fl / t_own / 1e9:6.0f} TFLOP/s)   library GEMM + relu_dropout_bwd pass {t_lib * 1e3:7.1f} us', flush=True)


def main():
    _lib.lib.ver_debug_gemm_variant.argtypes = [__import__('ctypes').c_int]
    ok = True
    # two-CTA kernel (N % 256 == 0, fp16 out) first: ragged M incl. a fully out-of-range second half
    _lib.lib.ver_debug_gemm_variant(2)
    for M, N, K, epi, p in ((256, 256, 64, 0, 0.0), (300, 256, 128, 0, 0.0), (5000, 768, 768, 0, 0.0),
                            (3000, 1536, 768, 2, 0.1), (2500, 768, 1536, 0, 0.0), (28224, 768, 768, 0, 0.0)):
        ok = check(M, N, K, epi, p) and ok
        if not ok:
            print('two-CTA kernel FAILED', flush=True)
            sys.exit(1)
    print('two-CTA kernel OK', flush=True)
    _lib.lib.ver_debug_gemm_variant(1)
    for M, N, K, epi, p in ((256, 256, 64, 0, 0.0), (1000, 256, 128, 0, 0.0), (300, 128, 192, 0, 0.0),
                            (4096, 192, 768, 1, 0.0), (5000, 768, 768, 0, 0.0), (3000, 1536, 768, 2, 0.0),
                            (3000, 1536, 768, 2, 0.1), (2500, 768, 1536, 0, 0.0)):
        ok = check(M, N, K, epi, p) and ok
        if not ok:
            break
    print('gemm_check:', 'OK' if ok else 'FAILED', flush=True)
    if not ok:
        sys.exit(1)
    rows = 8 * 25600
    for variant, tag in ((2, 'two-CTA'), (1, 'one-CTA')):
        _lib.lib.ver_debug_gemm_variant(variant)
        print(f'--- tcgen05 column = {tag} kernel where it applies', flush=True)
        bench(8 * 18 * 196, 768, 768, 0, 'value_proj')
        bench(rows, 768, 768, 0, 'output_proj')
        bench(rows, 1536, 768, 2, 'FFN1 + ReLU + dropout')
        bench(rows, 768, 1536, 0, 'FFN2')
    _lib.lib.ver_debug_gemm_variant(0)
    bench(rows, 192, 768, 1, 'offsets+weights logits')
    bench_bwd(rows, 1536, 768)
    return
    bench(8 * 18 * 196, 768, 768, 0, 'value_proj')
    bench(rows, 192, 768, 1, 'offsets+weights logits')
    bench(rows, 768, 768, 0, 'output_proj')
    bench(rows, 1536, 768, 2, 'FFN1 + ReLU + dropout')
    bench(rows, 768, 1536, 0, 'FFN2')


if __name__ == '__main__':
    main()
