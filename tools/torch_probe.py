"""GPU-side probe of torch features the host code wants to rely on (run on the box, prints findings)."""
import torch
dev = 'cuda'
a = torch.randn(4096, 768, device=dev, dtype=torch.float16)
w = torch.randn(192, 768, device=dev, dtype=torch.float16)
b32 = torch.randn(192, device=dev)
try:
    y = torch.mm(a, w.t(), out_dtype=torch.float32)
    print('mm out_dtype fp32 OK', y.dtype, (y - a.float() @ w.float().t()).abs().max().item())
except Exception as e:  # noqa: BLE001
    print('mm out_dtype FAILED', repr(e)[:200])
try:
    y = torch.addmm(b32, a, w.t(), out_dtype=torch.float32)
    print('addmm out_dtype fp32 OK', y.dtype, (y - (a.float() @ w.float().t() + b32)).abs().max().item())
except Exception as e:  # noqa: BLE001
    print('addmm out_dtype FAILED', repr(e)[:300])
try:
    g = torch.randn(4096, 192, device=dev, dtype=torch.float16)
    dw = torch.mm(g.t(), a, out_dtype=torch.float32)
    print('mm(g.t(), a) out_dtype OK', dw.dtype, dw.shape)
except Exception as e:  # noqa: BLE001
    print('mm transposed out_dtype FAILED', repr(e)[:300])


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


R = 204800
x = torch.randn(R, 768, device=dev, dtype=torch.float16)
h = torch.randn(R, 1536, device=dev, dtype=torch.float16)
w1 = torch.randn(1536, 768, device=dev, dtype=torch.float16) * 0.02
w2 = torch.randn(768, 1536, device=dev, dtype=torch.float16) * 0.02
wl = torch.randn(192, 768, device=dev, dtype=torch.float16) * 0.02
b1 = torch.zeros(1536, device=dev, dtype=torch.float16)
print('linear 768->1536 fwd us', timeit(lambda: torch.nn.functional.linear(x, w1, b1)))
print('linear 1536->768 fwd us', timeit(lambda: torch.nn.functional.linear(h, w2)))
print('mm dx 1536->768 (dy@W) us', timeit(lambda: torch.mm(h, w1)))
print('mm dW (dy^T@x) fp16 us', timeit(lambda: torch.mm(h.t(), x)))
try:
    print('mm dW (dy^T@x) fp32 out us', timeit(lambda: torch.mm(h.t(), x, out_dtype=torch.float32)))
    print('logits mm fp32 out us', timeit(lambda: torch.mm(x, wl.t(), out_dtype=torch.float32)))
except Exception as e:  # noqa: BLE001
    print('timing out_dtype FAILED', repr(e)[:200])
print('logits mm fp16 out us', timeit(lambda: torch.mm(x, wl.t())))
print('colsum fp16 768 us', timeit(lambda: x.sum(0)))
print('colsum fp16 1536 us', timeit(lambda: h.sum(0)))
ones = torch.ones(1, R, device=dev, dtype=torch.float16)
print('colsum via mm ones fp16 768 us', timeit(lambda: torch.mm(ones, x)))
print('copy fp16 768 us', timeit(lambda: x.clone()))
