"""SURVEY.md 8(f) row N1 -- the `up_sample` stack of the shipped vocc.py head (HEAD:254-258, applied at
HEAD:557-560): three ConvTranspose3d(768, 768, kernel (3,5,5), stride (1,2,2), padding (2,4,4),
dilation (2,2,2), output_padding (0,1,1)), 15x15 -> 120x120 laterally, 1.67 TFLOP per panorama as written.

Structure the reference's hyper-parameters imply (and cuDNN cannot see): along H and W a transposed
convolution writes output index o = 2 i - 4 + 2 k -- ALWAYS EVEN.  Every odd row / column of a layer's output
is the bare bias, and the data lives on the even-even lattice, which has the size of the layer's INPUT grid.
So with X_l = b_l + S(e_l)  (S = zero-stuffing onto the even-even lattice, e_0 = the voxel volume):

    e_1 = convT(e_0, W_1; stride 1, padding 2, dilation (2,1,1))                      (lattice H x W)
    e_l = convT(e_{l-1}, W_l; stride (1,2,2), padding 2, dilation (2,1,1), output_padding (0,1,1))
          + convT(1, U_l; stride 1, padding 2, dilation (2,1,1)),   U_l[k] = W_l[:, :, k]^T b_{l-1}      (l >= 2)

where the second term is the response to the constant bias field of the previous layer (a one-input-channel
convolution of a ones field: boundary-exact, negligible cost).  The data-carrying work is
(H W + H W + 4 H W) Z x 75 taps x 768^2 MACs = 0.478 TFLOP at 15x15x4 instead of 1.672: 3.5x fewer FLOPs, the
same values up to fp32 summation order (tests: fp64 equality on CPU, fp32 on the GPU against the oracle).
The dense (bs, C, Z, 8H, 8W) tensor HEAD:564 reinterprets is materialised once at the end.

Two executions of the lattice form:
  * `up_sample_lattice`  -- the small transposed convolutions as library calls; runs anywhere torch does: it is what
    the CPU tests use to pin the algebra in fp64 (and an A/B arm of tools/upsample_bench.py), not a product path;
  * `up_sample_gemm`     -- channels-last, each layer = one library GEMM  cols = e @ W  (M = lattice positions,
    K = 768, N = 75 * 768) + the hand-written gather `ver_convt_col2im` (csrc/col2im.cu); its backward is the
    adjoint kernel `ver_convt_im2col` + the GEMM's own autograd.  Measured on B200 (profiles/r01w, r01x): cuDNN
    runs the stack AS WRITTEN at 0.95-1.1 PFLOP/s in fp16, but the small lattice convolutions only at 0.11-0.41;
    the GEMM form keeps the 3.5x FLOP saving at GEMM speed.
`up_sample` picks per dtype / batch from those measurements (see its docstring).
"""
import torch
import torch.nn.functional as F


def _check(conv):
    ok = (tuple(conv.kernel_size) == (3, 5, 5) and tuple(conv.stride) == (1, 2, 2) and tuple(conv.padding) == (2, 4, 4)
          and tuple(conv.dilation) == (2, 2, 2) and tuple(conv.output_padding) == (0, 1, 1) and conv.groups == 1)
    if not ok:
        raise ValueError('lattice up-sampling is derived for the vocc.py ConvTranspose3d hyper-parameters '
                         '(HEAD:254-258); got ' + repr(conv))


def lattice_supported(convs):
    try:
        for c in convs:
            _check(c)
        return True
    except (ValueError, AttributeError):
        return False


def up_sample_lattice(x, convs, dtype=None, assemble=True):
    """x (bs, C, Z, H, W) -> (bs, C, Z, 2^L H, 2^L W), equal to nn.Sequential(*convs)(x).
    assemble=False: return (e, bias) instead -- the last layer's lattice data (bs, C, Z, 2^(L-1) H, 2^(L-1) W)
    and its bias, i.e. the dense result is bias everywhere plus e on the even-even lattice."""
    dtype = dtype or x.dtype
    acc = torch.float64 if dtype == torch.float64 else torch.float32      # precision of the bias-field constants
    e = x.to(dtype)
    prev_bias = None
    for layer, conv in enumerate(convs):
        _check(conv)
        w, b = conv.weight, conv.bias
        if b is None:
            b = w.new_zeros(w.shape[1])
        wd = w.to(dtype)
        if layer == 0:
            e = F.conv_transpose3d(e, wd, None, stride=1, padding=(2, 2, 2), dilation=(2, 1, 1))
        else:
            e = F.conv_transpose3d(e, wd, None, stride=(1, 2, 2), padding=(2, 2, 2), output_padding=(0, 1, 1),
                                   dilation=(2, 1, 1))
            # response to the previous layer's constant bias field (master-precision weights: a C-term dot per tap)
            u = torch.einsum('iokhw,i->okhw', w.to(acc), prev_bias.to(acc)).unsqueeze(0)          # (1, Cout, 3, 5, 5)
            ones = torch.ones((1, 1) + tuple(e.shape[2:]), dtype=acc, device=e.device)
            c = F.conv_transpose3d(ones, u, None, stride=1, padding=(2, 2, 2), dilation=(2, 1, 1))
            e = e + c.to(dtype)
        prev_bias = b
    if not assemble:
        return e, prev_bias
    bs, C, Z, H, W = e.shape
    out = prev_bias.to(dtype).view(1, C, 1, 1, 1).expand(bs, C, Z, 2 * H, 2 * W).contiguous()
    out[:, :, :, 0::2, 0::2] += e
    return out


def _bias_field_response(conv, prev_bias, Z, H, W, acc, device):
    """convT(constant field prev_bias) sampled on the output lattice -> (Z*H*W, Cout), boundary exact."""
    u = torch.einsum('iokhw,i->okhw', conv.weight.to(acc), prev_bias.to(acc)).unsqueeze(0)
    ones = torch.ones((1, 1, Z, H, W), dtype=acc, device=device)
    c = F.conv_transpose3d(ones, u, None, stride=1, padding=(2, 2, 2), dilation=(2, 1, 1))      # (1, Cout, Z, H, W)
    return c.flatten(2).transpose(1, 2)[0]


_WMAT_CACHE = {}


def _weight_matrix(weight, dtype):
    """(Cin, Cout, 3, 5, 5) -> (Cin, 75 * Cout), column = tap * Cout + channel, in `dtype`.  Under no_grad
    (inference) the 88 MB fp16 matrix is cached per parameter version, like fused_layer.half_of."""
    cin, cout = weight.shape[:2]
    if torch.is_grad_enabled() and weight.requires_grad:
        return weight.to(dtype).permute(0, 2, 3, 4, 1).reshape(cin, 75 * cout)
    import weakref
    from .fused_layer import _GENERATION            # invalidate_weight_cache() drops these copies too
    key = (id(weight), dtype)
    hit = _WMAT_CACHE.get(key)
    tag = (_GENERATION[0], weight._version, weight.data_ptr(), weight.device)
    if hit is None or hit[0] != tag or hit[2]() is not weight:
        ref = weakref.ref(weight, lambda _r, k=key: _WMAT_CACHE.pop(k, None))
        hit = (tag, weight.detach().to(dtype).permute(0, 2, 3, 4, 1).reshape(cin, 75 * cout).contiguous(), ref)
        _WMAT_CACHE[key] = hit
    return hit[1]


def up_sample_gemm(x, convs, dtype=None, col2im=None, assemble=True):
    """Same contract as up_sample_lattice (including assemble=False); GEMM + col2im execution (CUDA).  `col2im(cols, Z, Hi, Wi, s)` defaults to
    the libver_b200 kernel; tests inject a torch restatement to check the algebra on CPU."""
    if col2im is None:
        from . import ops
        col2im = ops.convt_col2im_fn
    dtype = dtype or x.dtype
    acc = torch.float64 if dtype == torch.float64 else torch.float32
    bs, C, Z, H, W = x.shape
    e = x.to(dtype).flatten(2).transpose(1, 2).contiguous()                  # (bs, Z*H*W, C) channels-last
    prev_bias = None
    for layer, conv in enumerate(convs):
        _check(conv)
        cin, cout = conv.weight.shape[:2]
        wmat = _weight_matrix(conv.weight, dtype)
        cols = (e.reshape(-1, cin) @ wmat).view(bs, Z * H * W, 75, cout)
        s = 1 if layer == 0 else 2
        e = col2im(cols, Z, H, W, s)                                          # (bs, Z*(sH)*(sW), Cout)
        H, W = s * H, s * W
        if layer > 0:
            e = e + _bias_field_response(conv, prev_bias, Z, H, W, acc, e.device).to(dtype)
        prev_bias = conv.bias if conv.bias is not None else conv.weight.new_zeros(cout)
    C = e.shape[-1]
    if not assemble:
        return e.view(bs, Z, H, W, C).permute(0, 4, 1, 2, 3), prev_bias
    out = prev_bias.to(dtype).view(1, C, 1, 1, 1).expand(bs, C, Z, 2 * H, 2 * W).contiguous()
    out[:, :, :, 0::2, 0::2] += e.view(bs, Z, H, W, C).permute(0, 4, 1, 2, 3)
    return out


def up_sample_dense(x, convs, dtype=None):
    """The stack exactly as the reference writes it (three library ConvTranspose3d calls) in `dtype`."""
    dtype = dtype or x.dtype
    y = x.to(dtype)
    for conv in convs:
        y = F.conv_transpose3d(y, conv.weight.to(dtype), None if conv.bias is None else conv.bias.to(dtype),
                               stride=conv.stride, padding=conv.padding, output_padding=conv.output_padding,
                               dilation=conv.dilation)
    return y


def pick_execution(dtype, positions):
    """Which execution `up_sample` uses on the GPU, from the measurements in profiles/r01x_upsample_bench.txt:
    fp32 -> GEMM + col2im lattice form (4-6x faster forward, 18x faster forward+backward than the stack as
    written); fp16 -> cuDNN runs the stack as written at ~1 PFLOP/s, which the lattice form only beats from
    ~8 panoramas of 4x15x15 on (9.2 vs 12.2 ms at 8), so smaller inputs stay dense.
    `positions` = batch * Z * H * W of the input volume."""
    if dtype == torch.float16 and positions < 8 * 900:
        return 'dense'
    return 'gemm'


def up_sample(x, convs, dtype=None):
    """HEAD:557-560 `self.up_sample(bev_for_occ)` on the product path: CUDA tensors only, like every other op of the
    package (`up_sample_lattice` is what the CPU tests call to pin the algebra)."""
    if not x.is_cuda:
        from ._lib import VerError
        raise VerError('vln_ver_b200.upsample.up_sample needs a CUDA tensor (no CPU fallback on the product path)')
    dtype = dtype or x.dtype
    how = pick_execution(dtype, x.shape[0] * x.shape[2] * x.shape[3] * x.shape[4])
    return {'dense': up_sample_dense, 'gemm': up_sample_gemm}[how](x, convs, dtype)


# --------------------------------------------------------------------------- occ_proj on the reinterpreted volume
_OCC_PLAN_CACHE = {}


def _occ_proj_plan(C, Z, Hf, Wf, Xo, Yo, device):
    """Index plan of HEAD:564-572 for an up-sampled volume (C, Z, Hf, Wf) that is bias everywhere except on the
    even-even lattice.  HEAD:564 REINTERPRETS that memory as (Z, Xo, Yo, C); row (x, y) of the permuted /
    flattened tensor (:570) is the concatenation over z' of the C-element run starting at flat offset
    ((z' Xo + x) Yo + y) C.  Returns, per distinct data pattern: the rows that have it, the positions inside the
    Z*C-wide row that carry data, and for each (row, position) the flat index into the lattice tensor
    (C, Z, Hf/2, Wf/2); plus, for every (row, z'), the channel whose bias fills the run."""
    key = (C, Z, Hf, Wf, Xo, Yo, str(device))
    if key in _OCC_PLAN_CACHE:
        return _OCC_PLAN_CACHE[key]
    V = Z * Hf * Wf
    if not occ_proj_plan_supported(C, Z, Hf, Wf, Xo, Yo):
        raise ValueError('occ_proj lattice plan: a C-element run must stay inside one channel of the volume')
    rows = torch.arange(Xo * Yo)
    rho = torch.arange(Z)[None, :] * (Xo * Yo) + rows[:, None]                    # (rows, Z) run index
    chan = rho // (V // C)                                                        # bias channel of each run
    start = (rho % (V // C)) * C                                                  # offset of the run inside its channel
    j = torch.arange(C)
    off = start[:, :, None] + j[None, None, :]                                    # (rows, Z, C) offset inside the channel
    z, rem = off // (Hf * Wf), off % (Hf * Wf)
    h, w = rem // Wf, rem % Wf
    data = (h % 2 == 0) & (w % 2 == 0)                                            # (rows, Z, C)
    src = ((chan[:, :, None] * Z + z) * (Hf // 2) + h // 2) * (Wf // 2) + w // 2  # index into (C, Z, Hf/2, Wf/2)
    data2 = data.reshape(len(rows), Z * C)
    src2 = src.reshape(len(rows), Z * C)
    # group rows by their data pattern (at the vocc.py shape: y mod 5, five patterns of 180-204 positions per run)
    patterns, inverse = torch.unique(data2, dim=0, return_inverse=True)
    groups = []
    for p in range(patterns.shape[0]):
        r = torch.nonzero(inverse == p)[:, 0]
        cols = torch.nonzero(patterns[p])[:, 0]
        groups.append((r.to(device), cols.to(device), src2[r][:, cols].to(device)))
    plan = (groups, chan.to(device))
    _OCC_PLAN_CACHE[key] = plan
    return plan


def occ_proj_plan_supported(C, Z, Hf, Wf, Xo, Yo):
    """a C-element run of the reinterpreted memory must stay inside one channel of the (C, Z, Hf, Wf) volume"""
    return (Z * Hf * Wf) % C == 0 and Xo * Yo == Hf * Wf and Hf % 2 == 0 and Wf % 2 == 0


def occ_proj_from_lattice(e, last_bias, weight, bias, Xo, Yo, dtype=None):
    """`occ_proj` of HEAD:564-572 applied to the up-sampled volume WITHOUT materialising it:
    equals  F.linear(X.view(bs, Z, Xo, Yo, C).permute(0, 2, 3, 1, 4).flatten(3), weight, bias)  for
    X = last_bias + (e on the even-even lattice), e (bs, C, Z, Hf/2, Wf/2) from up_sample_*(assemble=False).
    Three quarters of every input row are bias constants: their contribution is a per-run scalar times the row sums
    of a weight block, and the GEMM runs on the data positions only (K = 720-816 instead of 3072 at the vocc.py
    shape).  Returns (bs, Xo, Yo, out_features)."""
    dtype = dtype or e.dtype
    acc = torch.float64 if dtype == torch.float64 else torch.float32
    bs, C, Z, H2, W2 = e.shape
    groups, chan = _occ_proj_plan(C, Z, 2 * H2, 2 * W2, Xo, Yo, e.device)
    out_f = weight.shape[0]
    # constant part: sum over z' of last_bias[channel of the run] * (sum of the weight block's columns)
    wsum = weight.to(acc).view(out_f, Z, C).sum(-1)                               # (out, Z)
    const = last_bias.to(acc)[chan] @ wsum.t()                                    # (rows, out)
    if bias is not None:
        const = const + bias.to(acc)
    out = const.to(dtype).unsqueeze(0).repeat(bs, 1, 1)                           # (bs, rows, out)
    e_flat = e.reshape(bs, -1).to(dtype)
    wt = weight.to(dtype).t()                                                     # (Z*C, out)
    for r, cols, src in groups:
        vals = e_flat[:, src.reshape(-1)].view(bs, src.shape[0], src.shape[1])    # (bs, rows_p, K_p)
        out[:, r] += vals @ wt[cols]
    return out.view(bs, Xo, Yo, out_f)
