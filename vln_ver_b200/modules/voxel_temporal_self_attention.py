"""SURVEY.md 8(f) row N3 -- VoxelTemporalSelfAttention, mirror of
projects/mmdet3d_plugin/bevformer/modules/voxel_temporal_self_attention.py (:26-273): every voxel
attends into the (previous | current) voxel volume with 3-D deformable sampling.

Difference from the shipped module, on purpose: the reference's init_weights
builds a 2-component offset bias (:113-124) for a Linear with 3 components per point (:99-100) and
assigns it over `.bias.data`, so its first forward fails inside F.linear.  Here the bias is built with
the 3-component directions of VoxelCustomMSDeformableAttention.init_weights
(M/voxel_decoder.py:212-229), which is what the 3-D sampler needs; everything else -- parameters,
state_dict keys, forward arithmetic -- is the reference's.  The sampler runs in libver_b200.so
(ver_msda3d_forward/backward).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..registry import ATTENTION, BaseModule, constant_init, xavier_init
from .precision import PrecisionMixin
from .voxel_decoder import _shape_list


@ATTENTION.register_module()
class VoxelTemporalSelfAttention(PrecisionMixin, BaseModule):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, num_bev_queue=2,
                 im2col_step=64, dropout=0.1, batch_first=True, norm_cfg=None, init_cfg=None):
        super().__init__(init_cfg)
        if embed_dims % num_heads != 0:
            raise ValueError(f'embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}')
        self.norm_cfg = norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.fp16_enabled = False
        self.im2col_step = im2col_step
        self.embed_dims, self.num_levels, self.num_heads = embed_dims, num_levels, num_heads
        self.num_points, self.num_bev_queue = num_points, num_bev_queue
        self.sampling_offsets = nn.Linear(embed_dims * num_bev_queue,
                                          num_bev_queue * num_heads * num_levels * num_points * 3)
        self.attention_weights = nn.Linear(embed_dims * num_bev_queue,
                                           num_bev_queue * num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        constant_init(self.sampling_offsets, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin(), thetas.cos() + thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(
            self.num_heads, 1, 1, 3).repeat(1, self.num_levels * self.num_bev_queue, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid_init.view(-1)
        constant_init(self.attention_weights, val=0., bias=0.)
        xavier_init(self.value_proj, distribution='uniform', bias=0.)
        xavier_init(self.output_proj, distribution='uniform', bias=0.)
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, flag='decoder', **kwargs):
        """query (bs, Nq, C) (batch_first); value None -> the current volume stacked twice (:180-182) or
        (bs*2, Nq, C) = [previous, current]; reference_points (bs*2, Nq, num_levels, 3);
        spatial_shapes (num_levels, 3) = (d, h, w) -> (bs, Nq, C)."""
        if value is None:
            assert self.batch_first
            bs, len_bev, c = query.shape
            value = torch.stack([query, query], 1).reshape(bs * 2, len_bev, c)
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, embed_dims = query.shape
        _, num_value, _ = value.shape
        shapes = _shape_list(spatial_shapes)
        assert sum(int(d) * int(h) * int(w) for d, h, w in shapes) == num_value
        assert self.num_bev_queue == 2
        nq2 = self.num_bev_queue

        cd = self.compute_dtype or query.dtype
        query = torch.cat([value[:bs], query], -1)
        value = self._linear(value, self.value_proj, cd)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.reshape(bs * nq2, num_value, self.num_heads, -1)
        q32 = query.float()
        sampling_offsets = F.linear(q32, self.sampling_offsets.weight, self.sampling_offsets.bias).view(
            bs, num_query, self.num_heads, nq2, self.num_levels, self.num_points, 3)
        attention_weights = F.linear(q32, self.attention_weights.weight, self.attention_weights.bias).view(
            bs, num_query, self.num_heads, nq2, self.num_levels * self.num_points).softmax(-1).view(
            bs, num_query, self.num_heads, nq2, self.num_levels, self.num_points)
        attention_weights = attention_weights.permute(0, 3, 1, 2, 4, 5).reshape(
            bs * nq2, num_query, self.num_heads, self.num_levels, self.num_points).contiguous()
        sampling_offsets = sampling_offsets.permute(0, 3, 1, 2, 4, 5, 6).reshape(
            bs * nq2, num_query, self.num_heads, self.num_levels, self.num_points, 3)
        if reference_points.shape[-1] == 3:
            normalizer = torch.tensor([[w, h, d] for d, h, w in shapes], dtype=torch.float32, device=query.device)
            sampling_locations = reference_points[:, :, None, :, None, :].float() \
                + sampling_offsets / normalizer[None, None, None, :, None, :]
        else:
            raise ValueError(f'Last dim of reference_points must be'
                             f' 2 or 4, but get {reference_points.shape[-1]} instead.')
        output = ops.voxel_multi_scale_deformable_attn(value, shapes, sampling_locations, attention_weights)
        # mean over the (previous, current) pair (:262-266), written without the reference's permute round trip
        output = output.view(bs, nq2, num_query, embed_dims).mean(1)
        output = self._linear(output, self.output_proj, cd)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return self.dropout(output).to(identity.dtype) + identity
