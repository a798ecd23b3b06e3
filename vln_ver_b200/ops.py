"""torch-facing wrappers of the C ABI (include/ver_b200.h).

torch is plumbing here: device memory, the current stream and autograd graph
bookkeeping.  Every function requires CUDA tensors and raises otherwise -- there is
no CPU or eager fallback (the CPU restatement lives in oracle/ and is test-only).
"""
import ctypes
import os
from ctypes import c_double, c_int32, c_void_p

import torch
from torch.autograd.function import Function, once_differentiable

from ._lib import VER_F16, VER_F32, VerError, check, lib

IMG_W, IMG_H = 1280.0, 1024.0     # hard-coded in the reference, M/voxel_encoder.py:179-180
PROFILE_EVENTS = None             # set to a list by bench.py to collect (start, end) CUDA events of the forward sampler
PROFILE_EVENTS_BWD = None         # same for the backward sampler (fused layer / SCASampleTCFunction)


class nvtx_range:
    """NVTX range around a phase of the path (SURVEY section 5 tracing row) -- visible in nsys / ncu timelines.
    Off unless VER_NVTX=1: the markers are host calls on the launch path."""
    ON = os.environ.get('VER_NVTX', '0') == '1'

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if nvtx_range.ON:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if nvtx_range.ON:
            torch.cuda.nvtx.range_pop()


def nvtx_push(name):
    if nvtx_range.ON:
        torch.cuda.nvtx.range_push(name)


def nvtx_pop():
    if nvtx_range.ON:
        torch.cuda.nvtx.range_pop()


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VerError('vln_ver_b200 ops need CUDA tensors (no CPU fallback); got a '
                           f'{t.device} tensor')


def _code(dtype):
    if dtype == torch.float32:
        return VER_F32
    if dtype == torch.float16:
        return VER_F16
    raise VerError(f'unsupported storage dtype {dtype} (float32 / float16 only)')


def _c(t, dtype=None):
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------ A1 + A2, K8
def point_sampling(lidar2img, originshift, pc_range, bev_z, bev_h, bev_w, img_w=IMG_W, img_h=IMG_H):
    """Batched VoxelFormerEncoder.get_reference_points('3d') + point_sampling math
    (M/voxel_encoder.py:53-83, :136-195).  lidar2img (B, Ncam, 4, 4), originshift (B, 3).
    Returns reference_points_cam (Ncam, B, Nq, 1, 2) fp32, bev_mask (Ncam, B, Nq, 1) bool,
    vis_bits (B, Nq) int32 (bit c = camera c sees the voxel; None if Ncam > 32), count (B, Nq) int32."""
    _need_cuda(lidar2img, originshift)
    lidar2img = _c(lidar2img, torch.float32)
    originshift = _c(originshift, torch.float32)
    B, Ncam = lidar2img.shape[:2]
    assert lidar2img.shape[2:] == (4, 4) and originshift.shape == (B, 3)
    Nq = bev_z * bev_h * bev_w
    dev = lidar2img.device
    rpc = torch.empty((Ncam, B, Nq, 1, 2), dtype=torch.float32, device=dev)
    mask = torch.empty((Ncam, B, Nq, 1), dtype=torch.uint8, device=dev)
    bits = torch.empty((B, Nq), dtype=torch.int32, device=dev) if Ncam <= 32 else None
    count = torch.empty((B, Nq), dtype=torch.int32, device=dev)
    pc = (c_double * 6)(*[float(x) for x in pc_range])
    check(lib.ver_point_sampling_f32(_ptr(lidar2img), _ptr(originshift), pc, B, Ncam, bev_z, bev_h,
                                     bev_w, float(img_w), float(img_h), _ptr(rpc), _ptr(mask),
                                     _ptr(bits), _ptr(count), _stream()))
    return rpc, mask.view(torch.bool), bits, count


def visible_index(bev_mask):
    """Device-side replacement of the per-camera nonzero() of
    M/spatial_cross_attention.py:138-142.  bev_mask (Ncam, B, Nq[, 1]) bool ->
    counts (B, Ncam) int32, index (B, Ncam, Nq) int32 ascending, -1 padded."""
    _need_cuda(bev_mask)
    if bev_mask.dim() == 4:
        assert bev_mask.shape[-1] == 1, 'D (points per voxel) is 1 on this path (SURVEY A4.1)'
        bev_mask = bev_mask[..., 0]
    m = _c(bev_mask.view(torch.uint8) if bev_mask.dtype == torch.bool else bev_mask.to(torch.uint8))
    Ncam, B, Nq = m.shape
    counts = torch.empty((B, Ncam), dtype=torch.int32, device=m.device)
    index = torch.empty((B, Ncam, Nq), dtype=torch.int32, device=m.device)
    check(lib.ver_visible_index(_ptr(m), B, Ncam, Nq, _ptr(counts), _ptr(index), _stream()))
    return counts, index


# ------------------------------------------------------------------ A5 (operator boundary)
_SHAPES_HOST = {}


def shapes_to_host(spatial_shapes):
    """`spatial_shapes` as a Python list of lists.  The reference hands a (num_levels, 2|3) int64 TENSOR to every
    layer (M/voxel_encoder.py:257-284); reading it is a device -> host synchronisation, so the value is read ONCE
    per tensor (keyed on storage, version and shape) instead of once per layer per step."""
    if not isinstance(spatial_shapes, torch.Tensor):
        return [list(int(v) for v in s) for s in spatial_shapes]
    if not spatial_shapes.is_cuda:
        return spatial_shapes.tolist()
    key = (spatial_shapes.data_ptr(), spatial_shapes._version, tuple(spatial_shapes.shape), spatial_shapes.device.index)
    hit = _SHAPES_HOST.get(key)
    if hit is None:
        if len(_SHAPES_HOST) > 64:
            _SHAPES_HOST.clear()
        hit = _SHAPES_HOST[key] = spatial_shapes.tolist()          # the one D2H copy
    return hit


def _shapes_arg(spatial_shapes):
    spatial_shapes = shapes_to_host(spatial_shapes)
    flat = [int(v) for hw in spatial_shapes for v in hw]
    return (c_int32 * len(flat))(*flat), len(flat) // 2


def ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                           sampling_locations, attention_weights, im2col_step=64):
    """Drop-in for mmcv._ext.ms_deform_attn_forward
    (M/multi_scale_deformable_attn_function.py:118-124).  value (Bv, S, NH, Dh) fp32|fp16;
    sampling_locations (Bv, Nq, NH, NL, NP, 2); attention_weights (Bv, Nq, NH, NL, NP).
    `value_level_start_index` and `im2col_step` are accepted for signature parity
    (level starts are recomputed from the shapes; there is no im2col batching)."""
    _need_cuda(value, sampling_locations, attention_weights)
    value = _c(value)
    loc = _c(sampling_locations, torch.float32)
    w = _c(attention_weights, torch.float32)
    Bv, S, NH, Dh = value.shape
    _, Nq, _, NL, NP, _ = loc.shape
    shapes, nl = _shapes_arg(value_spatial_shapes)
    assert nl == NL
    out = torch.empty((Bv, Nq, NH * Dh), dtype=value.dtype, device=value.device)
    check(lib.ver_msda_forward(_code(value.dtype), _ptr(value), shapes, NL, _ptr(loc), _ptr(w),
                               _ptr(out), Bv, S, NH, Dh, Nq, NP, _stream()))
    return out


def ms_deform_attn_backward(value, value_spatial_shapes, value_level_start_index,
                            sampling_locations, attention_weights, grad_output, grad_value,
                            grad_sampling_loc, grad_attn_weight, im2col_step=64):
    """Drop-in for mmcv._ext.ms_deform_attn_backward (:150-160): fills the three
    caller-provided buffers (fp32)."""
    _need_cuda(value, sampling_locations, attention_weights, grad_output, grad_value)
    value = _c(value)
    loc = _c(sampling_locations, torch.float32)
    w = _c(attention_weights, torch.float32)
    go = _c(grad_output, value.dtype)
    Bv, S, NH, Dh = value.shape
    _, Nq, _, NL, NP, _ = loc.shape
    shapes, _ = _shapes_arg(value_spatial_shapes)
    for g in (grad_value, grad_sampling_loc, grad_attn_weight):
        assert g.is_contiguous() and g.dtype == torch.float32
    check(lib.ver_msda_backward(_code(value.dtype), _ptr(value), shapes, NL, _ptr(loc), _ptr(w),
                                _ptr(go), _ptr(grad_value), _ptr(grad_sampling_loc),
                                _ptr(grad_attn_weight), Bv, S, NH, Dh, Nq, NP, _stream()))


class MultiScaleDeformableAttnFunction(Function):
    """Mirror of MultiScaleDeformableAttnFunction_fp32
    (M/multi_scale_deformable_attn_function.py:90-163): same positional signature,
    same saved tensors, same gradient tuple."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index,
                sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.shapes = shapes_to_host(value_spatial_shapes)
        out = ms_deform_attn_forward(value, ctx.shapes, value_level_start_index,
                                     sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, sampling_locations, attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value = torch.empty(value.shape, dtype=torch.float32, device=value.device)
        grad_sampling_loc = torch.empty(sampling_locations.shape, dtype=torch.float32,
                                        device=value.device)
        grad_attn_weight = torch.empty(attention_weights.shape, dtype=torch.float32,
                                       device=value.device)
        ms_deform_attn_backward(value, ctx.shapes, None, sampling_locations, attention_weights,
                                grad_output.contiguous(), grad_value, grad_sampling_loc,
                                grad_attn_weight, ctx.im2col_step)
        return (grad_value.to(value.dtype), None, None,
                grad_sampling_loc.to(sampling_locations.dtype),
                grad_attn_weight.to(attention_weights.dtype), None)


# aliases under the reference's names (spatial_cross_attention.py:24-25)
MultiScaleDeformableAttnFunction_fp32 = MultiScaleDeformableAttnFunction
MultiScaleDeformableAttnFunction_fp16 = MultiScaleDeformableAttnFunction


# ------------------------------------------------------------------ N2 / N3: 3-D (voxel volume) sampler
def _shapes3_arg(spatial_shapes):
    spatial_shapes = shapes_to_host(spatial_shapes)
    flat = [int(v) for dhw in spatial_shapes for v in dhw]
    if len(flat) % 3:
        raise VerError('3-D spatial_shapes must be (num_levels, 3) = (d, h, w)')
    return (c_int32 * len(flat))(*flat), len(flat) // 3


def voxel_ms_deform_attn_forward(value, value_spatial_shapes, sampling_locations, attention_weights):
    """sm_100a replacement of voxel_multi_scale_deformable_attn_pytorch
    (M/voxel_temporal_self_attention.py:275-335; same argument order and meaning).
    value (Bv, S, NH, Dh) fp32|fp16, S = sum d*h*w; sampling_locations (Bv, Nq, NH, NL, NP, 3) = (x, y, z);
    attention_weights (Bv, Nq, NH, NL, NP) -> (Bv, Nq, NH*Dh)."""
    _need_cuda(value, sampling_locations, attention_weights)
    value = _c(value)
    loc = _c(sampling_locations, torch.float32)
    w = _c(attention_weights, torch.float32)
    Bv, S, NH, Dh = value.shape
    _, Nq, _, NL, NP, last = loc.shape
    if last != 3:
        raise VerError(f'sampling_locations last dim must be 3 (x, y, z), got {last}')
    assert loc.shape[0] == Bv and loc.shape[2] == NH and w.shape == loc.shape[:-1]
    shapes, nl = _shapes3_arg(value_spatial_shapes)
    assert nl == NL, (nl, NL)
    out = torch.empty((Bv, Nq, NH * Dh), dtype=value.dtype, device=value.device)
    check(lib.ver_msda3d_forward(_code(value.dtype), _ptr(value), shapes, NL, _ptr(loc), _ptr(w), _ptr(out),
                                 Bv, S, NH, Dh, Nq, NP, _stream()))
    return out


def voxel_ms_deform_attn_backward(value, value_spatial_shapes, sampling_locations, attention_weights,
                                  grad_output):
    """Gradients of voxel_ms_deform_attn_forward: (grad_value, grad_loc, grad_w), all fp32."""
    _need_cuda(value, sampling_locations, attention_weights, grad_output)
    value = _c(value)
    loc = _c(sampling_locations, torch.float32)
    w = _c(attention_weights, torch.float32)
    go = _c(grad_output, value.dtype)
    Bv, S, NH, Dh = value.shape
    _, Nq, _, NL, NP, _ = loc.shape
    shapes, _ = _shapes3_arg(value_spatial_shapes)
    gv = torch.empty(value.shape, dtype=torch.float32, device=value.device)
    gl = torch.empty(loc.shape, dtype=torch.float32, device=value.device)
    gw = torch.empty(w.shape, dtype=torch.float32, device=value.device)
    check(lib.ver_msda3d_backward(_code(value.dtype), _ptr(value), shapes, NL, _ptr(loc), _ptr(w), _ptr(go),
                                  _ptr(gv), _ptr(gl), _ptr(gw), Bv, S, NH, Dh, Nq, NP, _stream()))
    return gv, gl, gw


class VoxelMultiScaleDeformableAttnFunction(Function):
    """Autograd node of the 3-D sampler (the reference differentiates through F.grid_sample)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, sampling_locations, attention_weights):
        ctx.shapes = shapes_to_host(value_spatial_shapes)
        out = voxel_ms_deform_attn_forward(value, ctx.shapes, sampling_locations, attention_weights)
        ctx.save_for_backward(value, sampling_locations, attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, loc, w = ctx.saved_tensors
        gv, gl, gw = voxel_ms_deform_attn_backward(value, ctx.shapes, loc, w, grad_output.contiguous())
        return gv.to(value.dtype), None, gl.to(loc.dtype), gw.to(w.dtype)


def voxel_multi_scale_deformable_attn(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Differentiable call with the reference function's signature."""
    return VoxelMultiScaleDeformableAttnFunction.apply(value, value_spatial_shapes, sampling_locations,
                                                       attention_weights)


# ------------------------------------------------------------------ N1: col2im of the lattice-form up_sample
def convt_col2im(cols, Z, Hi, Wi, s):
    """cols (B, Z*Hi*Wi, 75, C) -> (B, Z*(s Hi)*(s Wi), C): ver_convt_col2im (HEAD:254-258 in lattice form)."""
    _need_cuda(cols)
    cols = _c(cols)
    B, n_in, taps, C = cols.shape
    assert taps == 75 and n_in == Z * Hi * Wi, (cols.shape, Z, Hi, Wi)
    out = torch.empty((B, Z * s * Hi * s * Wi, C), dtype=cols.dtype, device=cols.device)
    check(lib.ver_convt_col2im(_code(cols.dtype), _ptr(cols), _ptr(out), B, Z, Hi, Wi, s, C, _stream()))
    return out


def convt_im2col(grad_out, Z, Hi, Wi, s):
    """adjoint of convt_col2im: grad_out (B, Z*(s Hi)*(s Wi), C) -> (B, Z*Hi*Wi, 75, C)."""
    _need_cuda(grad_out)
    grad_out = _c(grad_out)
    B, n_out, C = grad_out.shape
    assert n_out == Z * s * Hi * s * Wi
    gcols = torch.empty((B, Z * Hi * Wi, 75, C), dtype=grad_out.dtype, device=grad_out.device)
    check(lib.ver_convt_im2col(_code(grad_out.dtype), _ptr(grad_out), _ptr(gcols), B, Z, Hi, Wi, s, C, _stream()))
    return gcols


class ConvTCol2ImFunction(Function):
    @staticmethod
    def forward(ctx, cols, Z, Hi, Wi, s):
        ctx.dims = (Z, Hi, Wi, s)
        return convt_col2im(cols, Z, Hi, Wi, s)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        return convt_im2col(grad_out.contiguous(), *ctx.dims), None, None, None, None


def convt_col2im_fn(cols, Z, Hi, Wi, s):
    return ConvTCol2ImFunction.apply(cols, Z, Hi, Wi, s)


# ------------------------------------------------------------------ fused SCA sampler
class Visibility:
    """Per-forward camera geometry products shared by the three encoder layers."""

    def __init__(self, rpc, mask, bits, count, grid):
        self.rpc, self.mask, self.bits, self.count, self.grid = rpc, mask, bits, count, grid
        self._index = None
        self._order = None

    @property
    def index(self):
        if self._index is None:
            self._index = visible_index(self.mask)
        return self._index

    @property
    def order(self):
        """(order, smask, tile_union): voxels stably sorted by camera bit set, see visibility_order()."""
        if self._order is None:
            self._order = visibility_order(self.bits)
        return self._order


def visibility_order(bits):
    """Device-side counterpart of the per-camera rebatch (M/spatial_cross_attention.py:138-154) for the
    tensor-core sampler: bits (B, Nq) int32 -> order (B, Nq) int32, smask (B, Nq) int32,
    tile_union (B, ceil(Nq/128)) int32."""
    _need_cuda(bits)
    bits = _c(bits, torch.int32)
    B, Nq = bits.shape
    need = ctypes.c_size_t(0)
    check(lib.ver_visibility_order_workspace(B, Nq, ctypes.byref(need)))
    ws = torch.empty((need.value,), dtype=torch.uint8, device=bits.device)
    order = torch.empty((B, Nq), dtype=torch.int32, device=bits.device)
    smask = torch.empty((B, Nq), dtype=torch.int32, device=bits.device)
    tile_union = torch.empty((B, (Nq + 127) // 128), dtype=torch.int32, device=bits.device)
    check(lib.ver_visibility_order(_ptr(bits), B, Nq, _ptr(order), _ptr(smask), _ptr(tile_union), _ptr(ws),
                                   need.value, _stream()))
    return order, smask, tile_union


class SCASampleFunction(Function):
    """slots = fused sampler(value, logits) -- see ver_sca_forward in include/ver_b200.h."""

    @staticmethod
    def forward(ctx, value, logits, vis, Sh, Sw, NH, NP, head_major=False):
        """value: (Bv, S, NH, Dh) / (Bv, S, C) [mmcv layout] or (Bv, NH, S, Dh) if head_major."""
        _need_cuda(value, logits)
        value = _c(value)
        logits = _c(logits, torch.float32)
        Bv = value.shape[0]
        S = value.shape[2] if head_major else value.shape[1]
        C = value[0].numel() // S
        Z, H, W = vis.grid
        Ncam, B = vis.rpc.shape[:2]
        assert Bv == B * Ncam and S == Sh * Sw
        Dh = C // NH
        layout = 1 if head_major else 0
        Nq = Z * H * W
        assert logits.shape[0] == B * Nq
        if vis.bits is None:
            raise VerError('fused SCA needs Ncam <= 32')
        slots = torch.empty((B, Nq, C), dtype=value.dtype, device=value.device)
        prof = PROFILE_EVENTS
        if prof is not None:          # bench.py: CUDA events on the launching stream around the kernel
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(lib.ver_sca_forward(_code(value.dtype), _ptr(value), layout, _ptr(logits), logits.shape[1],
                                  _ptr(vis.rpc), _ptr(vis.bits), _ptr(slots), B, Ncam, Z, H, W, Sh, Sw,
                                  NH, Dh, NP, _stream()))
        if prof is not None:
            e1.record()
            prof.append((e0, e1))
        ctx.save_for_backward(value, logits)
        ctx.vis, ctx.dims = vis, (B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, layout)
        return slots

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_slots):
        value, logits = ctx.saved_tensors
        vis = ctx.vis
        B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, layout = ctx.dims
        counts, index = vis.index
        gs = _c(grad_slots, value.dtype)
        gvalue = torch.empty(value.shape, dtype=torch.float32, device=value.device)
        glogits = torch.empty(logits.shape, dtype=torch.float32, device=value.device)
        if logits.shape[1] > NH * NP * 3:
            glogits[:, NH * NP * 3:].zero_()
        check(lib.ver_sca_backward(_code(value.dtype), _ptr(value), layout, _ptr(logits), logits.shape[1],
                                   _ptr(vis.rpc), _ptr(vis.bits), _ptr(counts), _ptr(index), _ptr(gs),
                                   _ptr(gvalue), _ptr(glogits), B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP,
                                   _stream()))
        return gvalue.to(value.dtype), glogits, None, None, None, None, None, None


def sca_sample(value, logits, vis, Sh, Sw, NH, NP, head_major=False):
    return SCASampleFunction.apply(value, logits, vis, Sh, Sw, NH, NP, head_major)


def value_image(value, NH):
    """fp16 value maps (Bv, S, C) -> tcgen05 operand images (Bv, NH, Dh/8, SP/8, 8, 8)."""
    _need_cuda(value)
    assert value.dtype == torch.float16
    value = _c(value)
    Bv, S = value.shape[:2]
    Dh = value[0].numel() // S // NH
    SP = (S + 15) // 16 * 16
    vimg = torch.empty((Bv, NH, Dh // 8, SP // 8, 8, 8), dtype=torch.float16, device=value.device)
    check(lib.ver_value_image_f16(_ptr(value), _ptr(vimg), Bv, S, NH, Dh, _stream()))
    return vimg


def value_image16(value, NH, Sh, Sw):
    """fp16 value maps (Bv, Sh*Sw, C) -> operand images with 16-cell image rows (Bv, NH, Dh/8, 2*Sh, 8, 8):
    cell y*16 + x + 1 holds pixel (y, x), the rest of an image row is zero (sca_fwd_tc6_kernel's layout)."""
    _need_cuda(value)
    assert value.dtype == torch.float16
    value = _c(value)
    Bv, S = value.shape[:2]
    assert S == Sh * Sw
    Dh = value[0].numel() // S // NH
    vimg = torch.empty((Bv, NH, Dh // 8, 2 * Sh, 8, 8), dtype=torch.float16, device=value.device)
    check(lib.ver_value_image16_f16(_ptr(value), _ptr(vimg), Bv, Sh, Sw, NH, Dh, _stream()))
    return vimg


def sca_forward_sorted16(vimg16, logits, vis, Sh, Sw, NH, NP, variant=0):
    """Forward sampler on 16-cell image rows: slots (B, Nq, C) fp16.  variant 0 / 7 = sca_fwd_tc7_kernel (A operand in
    TMEM), 6 = sca_fwd_tc6_kernel (A in the shared-memory operand)."""
    _need_cuda(vimg16, logits)
    Ncam, B = vis.rpc.shape[:2]
    Z, H, W = vis.grid
    Nq = Z * H * W
    Dh = vimg16.shape[2] * 8
    if not lib.ver_tc6_supported(Ncam, Sh, Sw, Dh, NP):
        raise VerError('sca_forward_sorted16: shape not covered (Ncam <= 32, Sh, Sw <= 14, Dh in 32/64/96, NP in 4/8)')
    order, smask, tile_union = vis.order
    slots = torch.empty((B, Nq, NH * Dh), dtype=torch.float16, device=logits.device)
    check(lib.ver_sca_forward_sorted16(_ptr(vimg16), _ptr(logits), logits.shape[1], _ptr(vis.rpc), _ptr(order),
                                       _ptr(smask), _ptr(tile_union), _ptr(slots), B, Ncam, Nq, Sh, Sw, NH, Dh, NP,
                                       int(variant), _stream()))
    return slots


VER_LAYOUT_TC_IMAGE = 2
# 'sorted': visibility-sorted rows, the fastest measured kernel generation that covers the shape
# (sca_fwd_tc4_kernel); 'sorted3' / 'sorted4' / 'sorted5': force a generation (tests and tools/ A-B timing only;
# sorted5 = sca_fwd_tc5_kernel: three A operands in TMEM, two issuing threads, bounded waits);
# 'sorted16' / 'sorted16_6': the generations on 16-cell image rows with two builder threads per row -- sca_fwd_tc7_kernel (A
# operand in TMEM: 389 us against 408 us for sorted4, but it needs its own value image, ver_value_image16_f16, which the backward
# cannot share) / sca_fwd_tc6_kernel (A in the shared-memory operand, 564 us); 'block': sca_fwd_tc_kernel (4x8x8 voxel blocks)
TC_FORWARD = os.environ.get('VER_TC_FORWARD', 'sorted')      # (the environment override is for A/B runs of bench.py)
_SORTED_VARIANT = {'sorted': 0, 'sorted3': 3, 'sorted4': 4, 'sorted5': 5, 'sorted16': 0, 'sorted16_6': 0}
_SORTED16_VARIANT = {'sorted16': 7, 'sorted16_6': 6}      # sca_fwd_tc7_kernel / sca_fwd_tc6_kernel


class SCASampleTCFunction(Function):
    """Tensor-core (tcgen05) fused sampler, fp16 maps.  `value` is (Bv, S, C) in the mmcv layout (what
    value_proj produces); the operand image is built here and kept for backward."""

    @staticmethod
    def forward(ctx, value, logits, vis, Sh, Sw, NH, NP):
        _need_cuda(value, logits)
        logits = _c(logits, torch.float32)
        Bv, S = value.shape[:2]
        C = value[0].numel() // S
        Z, H, W = vis.grid
        Ncam, B = vis.rpc.shape[:2]
        assert Bv == B * Ncam and S == Sh * Sw
        Dh, Nq = C // NH, Z * H * W
        assert logits.shape[0] == B * Nq
        if vis.bits is None:
            raise VerError('fused SCA needs Ncam <= 32')
        vimg = value_image(value, NH)
        slots = torch.empty((B, Nq, C), dtype=torch.float16, device=value.device)
        prof = PROFILE_EVENTS
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if (TC_FORWARD in _SORTED16_VARIANT and logits.shape[1] % 4 == 0 and lib.ver_tc6_supported(Ncam, Sh, Sw, Dh, NP)):
            slots = sca_forward_sorted16(value_image16(value, NH, Sh, Sw), logits, vis, Sh, Sw, NH, NP,
                                         _SORTED16_VARIANT[TC_FORWARD])
        elif TC_FORWARD in _SORTED_VARIANT and NP % 4 == 0 and logits.shape[1] % 4 == 0:
            order, smask, tile_union = vis.order
            check(lib.ver_sca_forward_sorted(_ptr(vimg), _ptr(logits), logits.shape[1], _ptr(vis.rpc), _ptr(order),
                                             _ptr(smask), _ptr(tile_union), _ptr(slots), B, Ncam, Nq, Sh, Sw,
                                             NH, Dh, NP, _SORTED_VARIANT[TC_FORWARD], _stream()))
        else:
            check(lib.ver_sca_forward(VER_F16, _ptr(vimg), VER_LAYOUT_TC_IMAGE, _ptr(logits), logits.shape[1],
                                      _ptr(vis.rpc), _ptr(vis.bits), _ptr(slots), B, Ncam, Z, H, W, Sh, Sw,
                                      NH, Dh, NP, _stream()))
        if prof is not None:
            e1.record()
            prof.append((e0, e1))
        ctx.save_for_backward(vimg, logits)
        ctx.vis, ctx.dims, ctx.vshape = vis, (B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP), value.shape
        return slots

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_slots):
        vimg, logits = ctx.saved_tensors
        vis = ctx.vis
        B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP = ctx.dims
        counts, index = vis.index
        gs = _c(grad_slots, torch.float16)
        gvalue = torch.empty(ctx.vshape, dtype=torch.float32, device=vimg.device)    # mmcv layout
        glogits = torch.empty(logits.shape, dtype=torch.float32, device=vimg.device)
        if logits.shape[1] > NH * NP * 3:
            glogits[:, NH * NP * 3:].zero_()
        check(lib.ver_sca_backward(VER_F16, _ptr(vimg), VER_LAYOUT_TC_IMAGE, _ptr(logits), logits.shape[1],
                                   _ptr(vis.rpc), _ptr(vis.bits), _ptr(counts), _ptr(index), _ptr(gs),
                                   _ptr(gvalue), _ptr(glogits), B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP,
                                   _stream()))
        return gvalue.to(torch.float16), glogits, None, None, None, None, None


def sca_sample_tc(value, logits, vis, Sh, Sw, NH, NP):
    return SCASampleTCFunction.apply(value, logits, vis, Sh, Sw, NH, NP)


def tc_supported(dtype, Ncam, S, Dh, NP):
    """shapes the tcgen05 sampler covers (the rest takes the gather kernels); the shared-memory
    budget of the forward kernel (2 value images + 2 interpolation-matrix halves) caps Dh at 96 for
    14x14 maps."""
    if not (dtype == torch.float16 and Ncam <= 32 and 1 <= NP <= 8 and S <= 256 and Dh in (32, 64, 96, 128)):
        return False
    SP = (S + 15) // 16 * 16
    return 2 * Dh * SP * 2 + 2 * 128 * SP * 2 + 36 * 1024 <= 227 * 1024


# ------------------------------------------------------------------ A8 prologue, A6 epilogue
def feat_embed(feats, cams_embeds, level_embed, dtype=torch.float32):
    """(Ncam, B, S, C) fp32 + cams_embeds[cam] + level_embeds[0] -> (B*Ncam, S, C) `dtype`
    (M/voxel_transformer.py:146-168 + M/spatial_cross_attention.py:158-161).  No autograd:
    the embeddings' gradients are taken by the autograd wrapper in modules/."""
    _need_cuda(feats, level_embed)
    feats = _c(feats, torch.float32)
    Ncam, B, S, C = feats.shape
    out = torch.empty((B * Ncam, S, C), dtype=dtype, device=feats.device)
    ce = _c(cams_embeds.detach(), torch.float32) if cams_embeds is not None else None
    le = _c(level_embed.detach(), torch.float32)
    check(lib.ver_feat_embed(_code(dtype), _ptr(feats), _ptr(ce), _ptr(le), _ptr(out), Ncam, B, S, C,
                             _stream()))
    return out


def add_layernorm(x, residual, gamma, beta, eps=1e-5):
    """y = LayerNorm(x + residual) (inference; no autograd)."""
    _need_cuda(x, gamma, beta)
    x = _c(x)
    r = _c(residual, x.dtype) if residual is not None else None
    C = x.shape[-1]
    y = torch.empty_like(x)
    check(lib.ver_add_layernorm(_code(x.dtype), _ptr(x), _ptr(r), _ptr(_c(gamma.detach(), torch.float32)),
                                _ptr(_c(beta.detach(), torch.float32)), _ptr(y), x.numel() // C, C,
                                float(eps), _stream()))
    return y


# ---- dropout keys.  key = f(torch seed, host counter) + device epoch word.  The host counter gives every dropout
# site of every eager step its own key; the epoch word lives in device memory and is ADDED by the kernels when they
# run, so a step captured in a CUDA graph (whose host-side keys are frozen into the graph) still draws fresh masks
# on every replay -- the captured step calls advance_dropout_epoch() once.  Both parts are in dropout_rng_state()
# so that a checkpoint can restore the mask sequence.
_SEED_COUNTER = [0]
_SEED_EPOCH = {}          # device index -> int64 tensor (1,)


def _next_seed():
    """A fresh 64-bit Philox key per dropout site and step, derived from torch's global seed."""
    _SEED_COUNTER[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _SEED_COUNTER[0] * 0xD1B54A32D192ED03) & (2 ** 64 - 1)


def _seed_epoch(device):
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    t = _SEED_EPOCH.get(idx)
    if t is None:
        t = _SEED_EPOCH[idx] = torch.zeros(1, dtype=torch.int64, device=f'cuda:{idx}')
    return t


def advance_dropout_epoch(device=None):
    """Bump the device-side epoch word (a stream-ordered in-place add: capturable).  Call once per training step
    from inside a captured step; harmless (and unnecessary) in eager mode."""
    _seed_epoch(device if device is not None else torch.cuda.current_device()).add_(0x632BE59BD9B4E019)


def dropout_rng_state():
    return {'counter': _SEED_COUNTER[0], 'epoch': {i: int(t.item()) for i, t in _SEED_EPOCH.items()}}


def set_dropout_rng_state(state):
    _SEED_COUNTER[0] = int(state['counter'])
    for i, v in state.get('epoch', {}).items():
        _seed_epoch(f'cuda:{int(i)}').fill_(int(v))


class DropoutAddLayerNormFunction(Function):
    """y = LayerNorm(residual + dropout(x)): one pass forward, one pass backward
    (the 'norm' steps of VoxelFormerLayer fused with the residual adds / dropouts in front of them)."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, p, eps, training):
        _need_cuda(x, weight, bias)
        x = _c(x)
        r = _c(residual, x.dtype) if residual is not None else None
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty_like(x)
        need_bwd = any(ctx.needs_input_grad[:4])
        z = torch.empty_like(x) if need_bwd else None
        stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device) if need_bwd else None
        p_eff = float(p) if training else 0.0
        seed = _next_seed()
        w32, b32 = _c(weight.detach(), torch.float32), _c(bias.detach(), torch.float32)
        check(lib.ver_dropout_add_layernorm_fwd(_code(x.dtype), _ptr(x), _ptr(r), _ptr(w32), _ptr(b32), _ptr(y),
                                                _ptr(z), _ptr(stats), rows, C, float(eps), p_eff, seed,
                                                _ptr(_seed_epoch(x.device)), _stream()))
        if need_bwd:
            ctx.save_for_backward(z, stats, w32)
            ctx.p, ctx.seed, ctx.has_res, ctx.wdtype = p_eff, seed, residual is not None, weight.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        z, stats, w32 = ctx.saved_tensors
        C = z.shape[-1]
        rows = z.numel() // C
        dy = _c(dy, z.dtype)
        dx = torch.empty_like(z)
        dres = torch.empty_like(z) if ctx.has_res else None
        nb = lib.ver_dropout_add_layernorm_bwd_blocks(rows)
        dgp = torch.empty((nb, C), dtype=torch.float32, device=z.device)
        dbp = torch.empty((nb, C), dtype=torch.float32, device=z.device)
        check(lib.ver_dropout_add_layernorm_bwd(_code(z.dtype), _ptr(dy), _ptr(z), _ptr(stats), _ptr(w32),
                                                _ptr(dx), _ptr(dres), _ptr(dgp), _ptr(dbp), None, rows, C, ctx.p,
                                                ctx.seed, _ptr(_seed_epoch(z.device)), _stream()))
        return dx, dres, dgp.sum(0).to(ctx.wdtype), dbp.sum(0).to(ctx.wdtype), None, None, None


def dropout_add_layernorm(x, residual, weight, bias, p=0.0, eps=1e-5, training=False):
    return DropoutAddLayerNormFunction.apply(x, residual, weight, bias, p, eps, training)


class ReluDropoutFunction(Function):
    """h = dropout(relu(a)) in place (mmcv FFN's Linear -> ReLU(inplace) -> Dropout)."""

    @staticmethod
    def forward(ctx, a, p, training):
        _need_cuda(a)
        assert a.is_contiguous() and a.numel() % 8 == 0
        p_eff = float(p) if training else 0.0
        check(lib.ver_relu_dropout_fwd(_code(a.dtype), _ptr(a), _ptr(a), a.numel(), p_eff, _next_seed(),
                                       _ptr(_seed_epoch(a.device)), _stream()))
        ctx.mark_dirty(a)
        ctx.save_for_backward(a)
        ctx.p = p_eff
        return a

    @staticmethod
    @once_differentiable
    def backward(ctx, dh):
        h, = ctx.saved_tensors
        dh = _c(dh, h.dtype)
        da = torch.empty_like(h)
        check(lib.ver_relu_dropout_bwd(_code(h.dtype), _ptr(dh), _ptr(h), _ptr(da), h.numel(), ctx.p, 0, None,
                                       _stream()))
        return da, None, None


def relu_dropout_(a, p=0.0, training=False):
    return ReluDropoutFunction.apply(a, p, training)


# ------------------------------------------------------------------ K5: tcgen05 Linear with fused epilogue
LINEAR_BIAS_F16, LINEAR_BIAS_F32, LINEAR_BIAS_RELU_DROPOUT_F16 = 0, 1, 2


def linear_tc_supported(x, weight):
    return (x.is_cuda and x.dtype == torch.float16 and weight.dtype == torch.float16 and x.dim() == 2
            and bool(lib.ver_linear_supported(x.shape[0], weight.shape[0], x.shape[1])))


def linear_tc(x, weight, bias=None, epilogue=LINEAR_BIAS_F16, p=0.0, seed=0):
    """out = epilogue(x @ weight^T + bias) on the hand-written tcgen05 GEMM (csrc/gemm_tc.cu; no autograd).
    x (M, K) fp16, weight (N, K) fp16 (nn.Linear layout), bias (N,) fp32 or None."""
    _need_cuda(x, weight)
    assert x.dtype == torch.float16 and weight.dtype == torch.float16 and x.dim() == 2 and weight.dim() == 2
    x, weight = _c(x), _c(weight)
    M, K = x.shape
    N = weight.shape[0]
    assert weight.shape[1] == K
    b32 = _c(bias.detach(), torch.float32) if bias is not None else None
    out = torch.empty((M, N), dtype=torch.float32 if epilogue == LINEAR_BIAS_F32 else torch.float16, device=x.device)
    check(lib.ver_linear_f16(int(epilogue), _ptr(x), K, _ptr(weight), K, _ptr(b32), _ptr(out), N, M, N, K, float(p),
                             int(seed), _ptr(_seed_epoch(x.device)) if p > 0 else None, _stream()))
    return out


def linear_relu_dropout_bwd(dy, w_t, h, p):
    """da = (dy @ w_t^T) * [h > 0] / (1 - p) and its column sums as partial rows (fold -> bias gradient).
    dy (M, K) fp16; w_t (N, K) fp16 = W2^T; h (M, N) fp16 -> da (M, N) fp16, part (row blocks, N) fp32."""
    _need_cuda(dy, w_t, h)
    dy, w_t, h = _c(dy), _c(w_t), _c(h)
    M, K = dy.shape
    N = w_t.shape[0]
    assert w_t.shape[1] == K and h.shape == (M, N) and dy.dtype == w_t.dtype == h.dtype == torch.float16
    da = torch.empty((M, N), dtype=torch.float16, device=dy.device)
    part = torch.empty((lib.ver_linear_bwd_colsum_rows(M), N), dtype=torch.float32, device=dy.device)
    check(lib.ver_linear_relu_dropout_bwd_f16(_ptr(dy), K, _ptr(w_t), K, _ptr(h), _ptr(da), N, _ptr(part), M, N, K,
                                              float(p), _stream()))
    return da, part


# ------------------------------------------------------------------ A11, A12
class _FocalFunction(Function):
    @staticmethod
    def forward(ctx, logits, occ_gt, gamma, alpha):
        _need_cuda(logits, occ_gt)
        x = _c(logits, torch.float32)
        N, Ccls = x.shape
        if occ_gt.dim() == 1:            # dense class targets (N,), values in [0, Ccls]
            assert occ_gt.shape[0] == N
            gt, n_gt = None, -1
            dense = _c(occ_gt, torch.int32)
        else:                            # sparse (n, 2) (flat index, class)
            gt = _c(occ_gt, torch.int64)
            n_gt = gt.shape[0]
            dense = torch.empty((N,), dtype=torch.int32, device=x.device)
        loss = torch.empty((1,), dtype=torch.float32, device=x.device)
        npos = torch.empty((1,), dtype=torch.int32, device=x.device)
        grad = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        check(lib.ver_focal_loss(_ptr(x), _ptr(gt), n_gt, _ptr(dense), _ptr(loss), _ptr(npos),
                                 _ptr(grad), N, Ccls, float(gamma), float(alpha), _stream()))
        ctx.grad = grad
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(npos)
        return loss, npos

    @staticmethod
    @once_differentiable
    def backward(ctx, gloss, _gnpos):
        return (ctx.grad * gloss).to(ctx.in_dtype), None, None, None


def occupancy_focal_loss(logits, occ_gt, gamma=2.0, alpha=0.25, loss_weight=1.0):
    """HEAD:1386-1444 for one panorama: dense target from sparse (index, class) GT, sigmoid
    focal loss summed and divided by avg_factor = #(gt < classes), nan_to_num'ed."""
    loss_sum, npos = _FocalFunction.apply(logits, occ_gt, gamma, alpha)
    loss = loss_weight * loss_sum[0] / npos[0].to(torch.float32)
    return torch.nan_to_num(loss)


def occupancy_decode(logits, threshold=0.25):
    """get_occupancy_prediction (HEAD:1505-1524) -> (n_occ, 2) int64 (flat index, class).
    One host sync at the end to size the result (the reference's torch.where syncs too)."""
    _need_cuda(logits)
    x = _c(logits.reshape(-1, logits.shape[-1]), torch.float32)
    N, Ccls = x.shape
    pairs = torch.empty((N, 2), dtype=torch.int64, device=x.device)
    cnt = torch.empty((1,), dtype=torch.int32, device=x.device)
    scratch = torch.empty(((N + 1023) // 1024 + 1,), dtype=torch.int32, device=x.device)
    check(lib.ver_occupancy_decode(_ptr(x), N, Ccls, float(threshold), _ptr(pairs), _ptr(cnt),
                                   _ptr(scratch), _stream()))
    return pairs[:int(cnt.item())]
