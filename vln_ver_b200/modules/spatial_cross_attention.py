"""SpatialCrossAttention / MSDeformableAttention3D -- registry-compatible mirrors of
projects/mmdet3d_plugin/bevformer/modules/spatial_cross_attention.py (:31-176, :179-402)
running on libver_b200's sm_100a kernels.

Same class names, constructor kwargs, forward signatures, state_dict keys and error
behaviour; different execution: SCA never builds the padded per-camera rebatch, it
computes the offset / attention-weight projections once per voxel (they are the same
row for every camera that sees the voxel) and hands them to the fused sampler.
"""
import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..registry import ATTENTION, BaseModule, build_attention, constant_init, xavier_init
from .precision import PrecisionMixin


@ATTENTION.register_module()
class SpatialCrossAttention(PrecisionMixin, BaseModule):
    """Cross attention of voxel queries to the camera views that see them.

    Args mirror the reference (spatial_cross_attention.py:45-59): embed_dims, num_cams,
    pc_range, dropout, init_cfg, batch_first, deformable_attention (cfg dict)."""

    def __init__(self, embed_dims=256, num_cams=6, pc_range=None, dropout=0.1, init_cfg=None,
                 batch_first=False,
                 deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=256,
                                           num_levels=4),
                 **kwargs):
        super().__init__(init_cfg)
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        self.fp16_enabled = False
        self.deformable_attention = build_attention(deformable_attention)
        self.embed_dims = embed_dims
        self.num_cams = num_cams
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.batch_first = batch_first
        self.sampler = 'auto'          # 'auto': tcgen05 sampler for fp16 maps, gather otherwise; 'gather'
        self.init_weight()

    def init_weight(self):
        xavier_init(self.output_proj, distribution='uniform', bias=0.)

    # ------------------------------------------------------------------
    def _visibility(self, reference_points_cam, bev_mask, visibility, grid):
        if visibility is not None:
            return visibility
        # caller came through the plain reference signature: derive the per-voxel camera
        # bit set from bev_mask (tiny integer reduction; plumbing, not the hot path)
        Ncam, bs, Nq, D = bev_mask.shape
        if D != 1:
            raise NotImplementedError('num_Z_anchors (D) must be 1 on the VER path '
                                      '(get_reference_points ignores num_points_in_voxel)')
        if Ncam > 32:
            raise ops.VerError('fused SCA needs num_cams <= 32')
        m = bev_mask[..., 0].to(torch.int64)
        shifts = torch.arange(Ncam, device=m.device, dtype=torch.int64).view(Ncam, 1, 1)
        bits64 = (m << shifts).sum(0)
        bits = torch.where(bits64 >= 2 ** 31, bits64 - 2 ** 32, bits64).to(torch.int32)
        count = m.sum(0).to(torch.int32)
        if grid is None:
            grid = (1, 1, Nq)
        return ops.Visibility(reference_points_cam.to(torch.float32).contiguous(),
                              bev_mask, bits.contiguous(), count, grid)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, reference_points_cam=None,
                bev_mask=None, level_start_index=None, flag='encoder', **kwargs):
        """query (bs, Nq, C); key/value (Ncam, S, bs, C); reference_points_cam
        (Ncam, bs, Nq, 1, 2); bev_mask (Ncam, bs, Nq, 1) -> (bs, Nq, C)
        (reference forward: spatial_cross_attention.py:76-176)."""
        if key is None:
            key = query
        if value is None:
            value = key
        if residual is None:
            inp_residual = query
        if query_pos is not None:
            query = query + query_pos
        if residual is not None:
            # the reference leaves `inp_residual` / `slots` undefined in this case (:129-131,:176)
            raise NameError("name 'inp_residual' is not defined")

        bs, num_query, _ = query.size()
        num_cams, l, vb, embed_dims = key.shape
        da = self.deformable_attention
        hw = kwargs.get('spatial_hw')
        if hw is None:
            hw = [int(v) for v in ops.shapes_to_host(spatial_shapes)[0]]
        if da.num_levels != 1:
            raise NotImplementedError('fused SCA supports num_levels == 1 (vocc.py:58)')
        Sh, Sw = hw
        assert Sh * Sw == l, (Sh, Sw, l)       # spatial_cross_attention.py:334
        vis = self._visibility(reference_points_cam, bev_mask, kwargs.get('visibility'),
                               kwargs.get('voxel_grid'))

        cd = self.compute_dtype or query.dtype
        # value_proj on (bs*Ncam, S, C): view b*Ncam + cam (:158-161, :336)
        v = value.permute(2, 0, 1, 3).reshape(bs * num_cams, l, embed_dims)
        v = self._linear(v, da.value_proj, cd)
        use_tc = self.sampler != 'gather' and ops.tc_supported(
            v.dtype, num_cams, l, embed_dims // da.num_heads, da.num_points)
        if not use_tc:
            # head-major maps [Bv][NH][S][Dh]: each (view, head) map is one contiguous block, i.e. a
            # single bulk (TMA) copy into shared memory instead of 196 row copies
            v = v.view(bs * num_cams, l, da.num_heads, -1).permute(0, 2, 1, 3).contiguous()
        # one GEMM for sampling_offsets (+) attention_weights, once per VOXEL (:340-343)
        w_cat = torch.cat([da.sampling_offsets.weight, da.attention_weights.weight], 0)
        b_cat = torch.cat([da.sampling_offsets.bias, da.attention_weights.bias], 0)
        q2 = query.reshape(bs * num_query, embed_dims)
        if cd == torch.float32:
            logits = F.linear(q2.float(), w_cat, b_cat)
        else:
            # low-precision product only for the data-dependent part; bias joins in fp32
            logits = F.linear(q2.to(cd), w_cat.to(cd)).float() + b_cat
        if use_tc:      # fp16 maps: interpolation-matrix x value on the tcgen05 tensor cores
            slots = ops.sca_sample_tc(v, logits, vis, Sh, Sw, da.num_heads, da.num_points)
        else:           # fp32 (parity mode) / unsupported shapes: shared-memory gather kernels
            slots = ops.sca_sample(v, logits, vis, Sh, Sw, da.num_heads, da.num_points, head_major=True)
        slots = self._linear(slots, self.output_proj, cd)
        if kwargs.get('return_unfused', False):
            # VoxelFormerLayer fuses dropout + residual + the following LayerNorm into one kernel
            return slots, inp_residual
        return self.dropout(slots) + inp_residual.to(slots.dtype)


@ATTENTION.register_module()
class MSDeformableAttention3D(PrecisionMixin, BaseModule):
    """Deformable attention with per-anchor reference points
    (spatial_cross_attention.py:179-402); sampling runs on ops.MultiScaleDeformableAttnFunction."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=8, im2col_step=64,
                 dropout=0.1, batch_first=True, norm_cfg=None, init_cfg=None):
        super().__init__(init_cfg)
        if embed_dims % num_heads != 0:
            raise ValueError(f'embed_dims must be divisible by num_heads, '
                             f'but got {embed_dims} and {num_heads}')
        dim_per_head = embed_dims // num_heads
        self.norm_cfg = norm_cfg
        self.batch_first = batch_first
        self.output_proj = None
        self.fp16_enabled = False
        if dim_per_head % 32 != 0:
            warnings.warn('MSDeformableAttention3D: dims per head that are multiples of 32 '
                          '(32/64/96/128) take the shared-memory staged sm_100a kernels; '
                          f'{dim_per_head} falls back to the generic kernel.')
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        """offset bias = head-wise compass directions scaled by (point index + 1); zero
        offset/attention weights; xavier value_proj (reference :254-273)."""
        constant_init(self.sampling_offsets, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(
            self.num_heads, 1, 1, 2).repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid_init.view(-1)
        constant_init(self.attention_weights, val=0., bias=0.)
        xavier_init(self.value_proj, distribution='uniform', bias=0.)
        xavier_init(self.output_proj, distribution='uniform', bias=0.)
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        bs, num_value, _ = value.shape
        shapes = ops.shapes_to_host(spatial_shapes)
        assert sum(int(h) * int(w) for h, w in shapes) == num_value

        cd = self.compute_dtype or query.dtype
        value = self._linear(value, self.value_proj, cd)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, self.num_heads, -1)
        sampling_offsets = F.linear(query.float(), self.sampling_offsets.weight,
                                    self.sampling_offsets.bias).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
        attention_weights = F.linear(query.float(), self.attention_weights.weight,
                                     self.attention_weights.bias).view(
            bs, num_query, self.num_heads, self.num_levels * self.num_points).softmax(-1)
        attention_weights = attention_weights.view(bs, num_query, self.num_heads, self.num_levels,
                                                   self.num_points)
        if reference_points.shape[-1] == 2:
            normalizer = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32,
                                      device=query.device)
            bs, num_query, num_Z_anchors, xy = reference_points.shape
            reference_points = reference_points[:, :, None, None, None, :, :]
            sampling_offsets = sampling_offsets / normalizer[None, None, None, :, None, :]
            num_all_points = sampling_offsets.shape[4]
            assert num_all_points % num_Z_anchors == 0
            sampling_offsets = sampling_offsets.view(
                bs, num_query, self.num_heads, self.num_levels, num_all_points // num_Z_anchors,
                num_Z_anchors, xy)
            sampling_locations = (reference_points + sampling_offsets).view(
                bs, num_query, self.num_heads, self.num_levels, num_all_points, xy)
        elif reference_points.shape[-1] == 4:
            assert False
        else:
            raise ValueError(f'Last dim of reference_points must be'
                             f' 2 or 4, but get {reference_points.shape[-1]} instead.')
        output = ops.MultiScaleDeformableAttnFunction.apply(
            value, shapes, level_start_index, sampling_locations, attention_weights, self.im2col_step)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return output
