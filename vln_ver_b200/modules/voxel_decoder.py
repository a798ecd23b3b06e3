"""SURVEY.md 8(f) row N2 -- the detection decoder that follows the encoder in
VoxelPerceptionTransformer.forward: mirror of
projects/mmdet3d_plugin/bevformer/modules/voxel_decoder.py (:53-337) plus the two mmcv classes
vocc.py:139-145 names by string (`DetrTransformerDecoderLayer`, `MultiheadAttention`; mmcv-full
1.4.0 is not vendored by the reference, so they are provided here with mmcv's semantics and are
registered only when the real mmcv is absent).

The sampling itself -- 100 box queries x 8 heads x 4 points trilinearly reading the encoded voxel
volume -- runs in libver_b200.so (`ver_msda3d_forward/backward`, csrc/msda3d.cu); the reference
runs it as F.grid_sample on a 5-D tensor (M/voxel_temporal_self_attention.py:275-335).
"""
import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..registry import (ATTENTION, HAVE_MMCV, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, BaseModule,
                        ModuleList, build_transformer_layer, constant_init, xavier_init)
from .custom_base_transformer_layer import MyCustomBaseTransformerLayer
from .precision import PrecisionMixin


def inverse_sigmoid(x, eps=1e-5):
    """log(x / (1 - x)) with both terms clamped to eps (reference :35-50)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _shape_list(spatial_shapes):
    return ops.shapes_to_host(spatial_shapes)


# --------------------------------------------------------------------------- mmcv pieces named by vocc.py
class MultiheadAttention(BaseModule):
    """mmcv 1.4.0 `MultiheadAttention`: nn.MultiheadAttention with positional encodings added to
    query / key, proj_drop + dropout_layer on the output and the identity added.  The deprecated
    `dropout=` kwarg of vocc.py:145 sets both attn_drop and dropout_layer.drop_prob.
    state_dict keys: attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.{weight,bias}."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None, batch_first=False, **kwargs):
        super().__init__(init_cfg)
        dropout_layer = dict(dropout_layer) if dropout_layer else None
        if 'dropout' in kwargs:
            warnings.warn('The arguments `dropout` in MultiheadAttention has been deprecated, now you can '
                          'separately set `attn_drop`(float), proj_drop(float), and `dropout_layer`(dict) ')
            attn_drop = kwargs['dropout']
            dropout_layer = dropout_layer or dict(type='Dropout')
            dropout_layer['drop_prob'] = kwargs.pop('dropout')
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Dropout(dropout_layer.get('drop_prob', 0.5)) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None:
            if query_pos.shape == key.shape:
                key_pos = query_pos
            else:
                warnings.warn(f'position encoding of key is missing in {self.__class__.__name__}.')
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class BaseTransformerLayer(MyCustomBaseTransformerLayer):
    """mmcv `BaseTransformerLayer`: the constructor is the one the reference copied into
    MyCustomBaseTransformerLayer (M/custom_base_transformer_layer.py:72-163) with mmcv's
    batch_first=False default; forward is the plain operation-order walk (:165-260)."""

    def __init__(self, *args, batch_first=False, **kwargs):
        super().__init__(*args, batch_first=batch_first, **kwargs)

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [attn_masks.clone() for _ in range(self.num_attn)]
            warnings.warn(f'Use same attn_mask in all attentions in {self.__class__.__name__} ')
        else:
            assert len(attn_masks) == self.num_attn, \
                f'The length of attn_masks {len(attn_masks)} must be equal to the number of attention in ' \
                f'operation_order {self.num_attn}'
        for layer in self.operation_order:
            if layer == 'self_attn':
                query = self.attentions[attn_index](
                    query, query, query, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=query_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=query_key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'norm':
                norm = self.norms[norm_index]
                query = norm(query.to(norm.weight.dtype))      # LayerNorm in the parameters' precision
                norm_index += 1
            elif layer == 'cross_attn':
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=key_pos, attn_mask=attn_masks[attn_index], key_padding_mask=key_padding_mask,
                    **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'ffn':
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query


class DetrTransformerDecoderLayer(BaseTransformerLayer):
    """mmdet 2.14 `DetrTransformerDecoderLayer` (vocc.py:139): BaseTransformerLayer + the 6-op check."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), ffn_num_fcs=2, **kwargs):
        super().__init__(attn_cfgs=attn_cfgs, feedforward_channels=feedforward_channels,
                         ffn_dropout=ffn_dropout, operation_order=operation_order, act_cfg=act_cfg,
                         norm_cfg=norm_cfg, ffn_num_fcs=ffn_num_fcs, **kwargs)
        assert len(operation_order) == 6
        assert set(operation_order) == set(['self_attn', 'norm', 'cross_attn', 'ffn'])


if not HAVE_MMCV:
    ATTENTION.register_module()(MultiheadAttention)
    TRANSFORMER_LAYER.register_module()(BaseTransformerLayer)
    TRANSFORMER_LAYER.register_module()(DetrTransformerDecoderLayer)


# --------------------------------------------------------------------------- N2 attention
@ATTENTION.register_module()
class VoxelCustomMSDeformableAttention(PrecisionMixin, BaseModule):
    """Deformable cross-attention of the box queries into the voxel volume (reference :135-337).
    Same constructor, parameters (sampling_offsets 3 components per point, attention_weights,
    value_proj, output_proj), init and forward signature; `voxel_multi_scale_deformable_attn_pytorch`
    (:315-316) is replaced by the sm_100a kernel."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64, dropout=0.1,
                 batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__(init_cfg)
        if embed_dims % num_heads != 0:
            raise ValueError(f'embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}')
        self.norm_cfg = norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.fp16_enabled = False
        self.im2col_step = im2col_step
        self.embed_dims, self.num_levels, self.num_heads, self.num_points = embed_dims, num_levels, num_heads, num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 3)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        """offset bias: per head (cos t, sin t, cos t + sin t) normalised by its max-abs component and
        scaled by (point index + 1) (reference :212-229)."""
        constant_init(self.sampling_offsets, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin(), thetas.cos() + thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(
            self.num_heads, 1, 1, 3).repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid_init.view(-1)
        constant_init(self.attention_weights, val=0., bias=0.)
        xavier_init(self.value_proj, distribution='uniform', bias=0.)
        xavier_init(self.output_proj, distribution='uniform', bias=0.)
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, flag='decoder', **kwargs):
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
        shapes = _shape_list(spatial_shapes)
        assert sum(int(d) * int(h) * int(w) for d, h, w in shapes) == num_value

        cd = self.compute_dtype or query.dtype
        value = self._linear(value, self.value_proj, cd)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, self.num_heads, -1)
        q32 = query.float()
        sampling_offsets = F.linear(q32, self.sampling_offsets.weight, self.sampling_offsets.bias).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points, 3)
        attention_weights = F.linear(q32, self.attention_weights.weight, self.attention_weights.bias).view(
            bs, num_query, self.num_heads, self.num_levels * self.num_points).softmax(-1).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points)
        if reference_points.shape[-1] == 3:
            normalizer = torch.tensor([[w, h, d] for d, h, w in shapes], dtype=torch.float32, device=query.device)
            sampling_locations = reference_points[:, :, None, :, None, :].float() \
                + sampling_offsets / normalizer[None, None, None, :, None, :]
        else:
            raise ValueError(f'Last dim of reference_points must be'
                             f' 2 or 4, but get {reference_points.shape[-1]} instead.')
        output = ops.voxel_multi_scale_deformable_attn(value, shapes, sampling_locations, attention_weights)
        output = self._linear(output, self.output_proj, cd)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        # the (small) query stream keeps the identity's dtype; only the volume GEMM + sampling follow cd
        return self.dropout(output).to(identity.dtype) + identity


# --------------------------------------------------------------------------- N2 decoder
@TRANSFORMER_LAYER_SEQUENCE.register_module()
class VoxelDetectionTransformerDecoder(BaseModule):
    """`num_layers` decoder layers with iterative reference-point refinement (reference :53-132).
    Built like mmcv's TransformerLayerSequence: `transformerlayers` is one layer cfg (deep-copied
    num_layers times) or a list of them."""

    def __init__(self, *args, transformerlayers=None, num_layers=None, return_intermediate=False,
                 init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            import copy
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(build_transformer_layer(transformerlayers[i]))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm
        self.return_intermediate = return_intermediate
        self.fp16_enabled = False

    def forward(self, query, *args, reference_points=None, reg_branches=None, key_padding_mask=None, **kwargs):
        """query (num_query, bs, C); reference_points (bs, num_query, 3) in (0, 1); kwargs carry
        key=None, value (num_value, bs, C), query_pos, spatial_shapes [[Z, H, W]], level_start_index.
        Returns (stack of per-layer outputs, stack of per-layer reference points) when
        return_intermediate, else (output, reference_points)."""
        output = query
        intermediate, intermediate_reference_points = [], []
        for lid, layer in enumerate(self.layers):
            reference_points_input = reference_points[..., :3].unsqueeze(2)     # (bs, nq, num_levels=1, 3)
            output = layer(output, *args, reference_points=reference_points_input,
                           key_padding_mask=key_padding_mask, **kwargs)
            output = output.permute(1, 0, 2)
            if reg_branches is not None:
                tmp = reg_branches[lid](output)
                assert reference_points.shape[-1] == 3
                new_reference_points = torch.zeros_like(reference_points)
                new_reference_points[..., :2] = tmp[..., :2] + inverse_sigmoid(reference_points[..., :2])
                new_reference_points[..., 2:3] = tmp[..., 4:5] + inverse_sigmoid(reference_points[..., 2:3])
                reference_points = new_reference_points.sigmoid().detach()
            output = output.permute(1, 0, 2)
            if self.return_intermediate:
                intermediate.append(output)
                intermediate_reference_points.append(reference_points)
        if self.return_intermediate:
            return torch.stack(intermediate), torch.stack(intermediate_reference_points)
        return output, reference_points
