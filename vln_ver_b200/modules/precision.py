"""Storage/GEMM precision policy of the lift path.

`compute_dtype = None` (default): everything follows the input dtype -- fp32, the
reference's own precision (vocc.py has no fp16 key).
`compute_dtype = torch.float16`: fp32 master parameters, fp16 activations / value maps /
GEMM operands with fp32 accumulation; sampling locations, attention weights, LayerNorm
statistics, the loss and every reduction stay fp32 (BASELINE.json configs 2-5).
"""
import torch
import torch.nn.functional as F


class PrecisionMixin:
    compute_dtype = None

    @staticmethod
    def _linear(x, layer, cd):
        w, b = layer.weight, layer.bias
        if (cd == torch.float16 and x.is_cuda and torch.is_grad_enabled()
                and (w.requires_grad or x.requires_grad) and w.dtype == torch.float32):
            # fp16-storage training: cached fp16 weights + hand-written backward (bias gradient by column-sum kernel)
            from ..fused_layer import linear_f16
            return linear_f16(x, layer)
        if x.dtype != cd:
            x = x.to(cd)
        if w.dtype != cd:
            w = w.to(cd)
            b = b.to(cd) if b is not None else None
        return F.linear(x, w, b)


def set_compute_dtype(module, dtype):
    """Set the GEMM/storage dtype on every vln_ver_b200 module below `module`."""
    assert dtype in (None, torch.float32, torch.float16)
    if dtype == torch.float32:
        dtype = None
    for m in module.modules():
        if isinstance(m, PrecisionMixin):
            m.compute_dtype = dtype
    return module
