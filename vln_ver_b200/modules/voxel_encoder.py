"""VoxelFormerEncoder / VoxelFormerLayer -- mirrors of
projects/mmdet3d_plugin/bevformer/modules/voxel_encoder.py (:31-296, :300-464).

Differences in execution only: camera geometry (A1+A2) is one sm_100a kernel per
forward for the whole batch (no per-forward JSON/pickle parse on the hot path when the
matrices are supplied as tensors), its products are shared by the three layers, and the
layers call the fused SCA sampler.
"""
import copy
import json
import os
import pickle
import warnings

import torch

from .. import ops
from ..registry import (TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, BaseModule, ModuleList,
                        build_transformer_layer)
from .custom_base_transformer_layer import MyCustomBaseTransformerLayer
from .precision import PrecisionMixin
from .spatial_cross_attention import SpatialCrossAttention


def apply_layernorm(norm, x):
    """LayerNorm step of the layer.  Inference: the fused sm_100a row kernel (fp32 statistics
    whatever the storage type).  Training: torch's differentiable LayerNorm, evaluated in fp32
    when the activations are stored in fp16."""
    if x.is_cuda and not (torch.is_grad_enabled() and (x.requires_grad or norm.weight.requires_grad)):
        return ops.add_layernorm(x, None, norm.weight, norm.bias, norm.eps)
    if x.is_cuda and x.dtype == torch.float16 and x.shape[-1] % 8 == 0 and x.shape[-1] <= 1024:
        # fp16 storage, training: the fused row kernel (fp32 statistics, one pass forward, one pass backward)
        return ops.dropout_add_layernorm(x, None, norm.weight, norm.bias, 0.0, norm.eps, False)
    if x.dtype != norm.weight.dtype:
        return norm(x.float()).to(x.dtype)
    return norm(x)


class TransformerLayerSequence(BaseModule):
    """mmcv.cnn.bricks.transformer.TransformerLayerSequence (1.4.0) semantics."""

    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(build_transformer_layer(transformerlayers[i]))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class VoxelFormerEncoder(PrecisionMixin, TransformerLayerSequence):
    """Stack of VoxelFormerLayers over (bs, Nq, C) voxel queries.

    Camera metadata: each `img_metas[b]` may carry `lidar2img` (Ncam x 4 x 4) and
    `originshift` (3,) directly; otherwise `sample_idx = "<scan>_<vp>"` is resolved through
    `world2pixel_dir/<scan>.json` and `scanvp2cord_path` exactly like the reference
    (voxel_encoder.py:121-135; the defaults are the reference's literal relative paths).
    Device-resident batches can skip img_metas with `lidar2img=(B,Ncam,4,4)`,
    `originshift=(B,3)` keyword tensors.
    """
    world2pixel_dir = 'path to/camera_parameters/world2pixel/'
    scanvp2cord_path = 'path to/scanvp2cord.pkl'
    image_size = (1280.0, 1024.0)        # voxel_encoder.py:179-180

    def __init__(self, *args, pc_range=None, num_points_in_pillar=None, num_points_in_voxel=1,
                 return_intermediate=False, dataset_type='nuscenes', **kwargs):
        super().__init__(*args, **kwargs)
        self.return_intermediate = return_intermediate
        self.num_points_in_voxel = num_points_in_voxel
        self.pc_range = pc_range
        self.fp16_enabled = False
        self._file_cache = {}

    # ------------------------------------------------------------------ A1
    @staticmethod
    def get_reference_points(bev_z, bev_h, bev_w, num_points_in_voxel=1, dim='3d', bs=1,
                             device='cuda', dtype=torch.float):
        """Normalised voxel centres (reference :53-115).  '3d' -> (bs, 1, Nq, 3); '2d' ->
        (bs, Nq, 1, 3).  Tiny index arithmetic kept in torch for API parity; the hot path
        recomputes the same values inside ver_point_sampling_f32."""
        z = (torch.arange(bev_z, device=device, dtype=dtype) + 0.5) / bev_z
        y = (torch.arange(bev_h, device=device, dtype=dtype) + 0.5) / bev_h
        x = (torch.arange(bev_w, device=device, dtype=dtype) + 0.5) / bev_w
        grid = torch.stack([x.view(1, 1, bev_w).expand(bev_z, bev_h, bev_w),
                            y.view(1, bev_h, 1).expand(bev_z, bev_h, bev_w),
                            z.view(bev_z, 1, 1).expand(bev_z, bev_h, bev_w)], -1).reshape(-1, 3)
        if dim == '3d':
            return grid[None, None].repeat(bs, 1, 1, 1)
        if dim == '2d':
            return grid[None, :, None].repeat(bs, 1, 1, 1)
        raise ValueError(dim)

    # ------------------------------------------------------------------ A2
    def _camera_tensors(self, img_metas, num_cams, device, lidar2img=None, originshift=None):
        if lidar2img is not None:
            return (torch.as_tensor(lidar2img, dtype=torch.float32, device=device),
                    torch.as_tensor(originshift, dtype=torch.float32, device=device))
        mats, shifts = [], []
        for meta in img_metas:
            if 'lidar2img' in meta:
                mats.append(torch.as_tensor(meta['lidar2img'], dtype=torch.float64))
                shifts.append(torch.as_tensor(meta['originshift'], dtype=torch.float64))
                continue
            scan, vp = meta['sample_idx'].split('_')
            key = os.path.join(self.world2pixel_dir, scan + '.json')
            if key not in self._file_cache:
                with open(key, 'r') as f:
                    self._file_cache[key] = json.load(f)
            if self.scanvp2cord_path not in self._file_cache:
                with open(self.scanvp2cord_path, 'rb') as f:
                    self._file_cache[self.scanvp2cord_path] = pickle.load(f)
            world2pixel = self._file_cache[key]
            elevations = ['i1'] if num_cams == 6 else ['i0', 'i1', 'i2']
            names = [f'{vp}_{e}_{deg}' for e in elevations for deg in range(6)]
            mats.append(torch.tensor([world2pixel[n] for n in names], dtype=torch.float64))
            shifts.append(torch.tensor(self._file_cache[self.scanvp2cord_path][scan + '_' + vp],
                                       dtype=torch.float64))
        return (torch.stack(mats).to(torch.float32).to(device),
                torch.stack(shifts).to(torch.float32).to(device))

    def point_sampling(self, reference_points, pc_range, img_metas, num_cams=None, **cam):
        """Same contract as the reference (:118-195): returns reference_points_cam
        (Ncam, bs, Nq, D, 2) fp32 and bev_mask (Ncam, bs, Nq, D) bool.  `reference_points`
        (bs, 1, Nq, 3) only fixes bs / device; the kernel regenerates the identical grid from
        (bev_z, bev_h, bev_w) carried in `cam['voxel_grid']`."""
        vis = self.visibility(cam['voxel_grid'], reference_points.shape[0], img_metas,
                              reference_points.device, num_cams=num_cams,
                              lidar2img=cam.get('lidar2img'), originshift=cam.get('originshift'))
        return vis.rpc, vis.mask

    def visibility(self, grid, bs, img_metas, device, num_cams=None, lidar2img=None, originshift=None):
        if num_cams is None:
            num_cams = self.layers[0].attentions[0].num_cams if hasattr(
                self.layers[0].attentions[0], 'num_cams') else 6
        l2i, shift = self._camera_tensors(img_metas, num_cams, device, lidar2img, originshift)
        if l2i.shape[0] != bs:
            raise ValueError(f'{l2i.shape[0]} camera rigs for a batch of {bs} panoramas')
        rpc, mask, bits, count = ops.point_sampling(l2i, shift, self.pc_range, *grid,
                                                    img_w=self.image_size[0], img_h=self.image_size[1])
        return ops.Visibility(rpc, mask, bits, count, tuple(grid))

    # ------------------------------------------------------------------ A7
    def forward(self, bev_query, key, value, *args, bev_z=None, bev_h=None, bev_w=None, bev_pos=None,
                spatial_shapes=None, level_start_index=None, valid_ratios=None, prev_bev=None,
                shift=0., **kwargs):
        """bev_query (Nq, bs, C); key/value (Ncam, S, bs, C) -> (bs, Nq, C)
        (or (num_layers, bs, Nq, C) with return_intermediate), reference :197-296."""
        if prev_bev is not None:
            raise NotImplementedError('temporal prev_bev is not on the vocc.py path (prev_bev=None)')
        output = bev_query
        intermediate = []
        bs = bev_query.size(1)
        num_cams = key.shape[0]
        vis = kwargs.pop('visibility', None)
        if vis is None:
            vis = self.visibility((bev_z, bev_h, bev_w), bs, kwargs.get('img_metas'), bev_query.device,
                                  num_cams=num_cams, lidar2img=kwargs.pop('lidar2img', None),
                                  originshift=kwargs.pop('originshift', None))
        hw = ops.shapes_to_host(spatial_shapes)[0]
        bev_query = bev_query.permute(1, 0, 2)
        if bev_pos is not None:
            bev_pos = bev_pos.permute(1, 0, 2)
        for lid, layer in enumerate(self.layers):
            output = layer(bev_query, key, value, *args, bev_pos=bev_pos, ref_2d=None, ref_3d=None,
                           bev_z=bev_z, bev_h=bev_h, bev_w=bev_w, spatial_shapes=spatial_shapes,
                           level_start_index=level_start_index, reference_points_cam=vis.rpc,
                           bev_mask=vis.mask, prev_bev=prev_bev, visibility=vis,
                           spatial_hw=[int(hw[0]), int(hw[1])], voxel_grid=(bev_z, bev_h, bev_w),
                           **kwargs)
            bev_query = output
            if self.return_intermediate:
                intermediate.append(output)
        if self.return_intermediate:
            return torch.stack(intermediate)
        return output


@TRANSFORMER_LAYER.register_module()
class VoxelFormerLayer(MyCustomBaseTransformerLayer):
    """One encoder layer; vocc.py uses operation_order ('cross_attn','norm','ffn','norm')
    (reference :300-464)."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), ffn_num_fcs=2,
                 **kwargs):
        super().__init__(attn_cfgs=attn_cfgs, feedforward_channels=feedforward_channels,
                         ffn_dropout=ffn_dropout, operation_order=operation_order, act_cfg=act_cfg,
                         norm_cfg=norm_cfg, ffn_num_fcs=ffn_num_fcs, **kwargs)
        self.fp16_enabled = False
        self.fuse_epilogues = True
        self.fuse_layer = True          # fp16 storage on CUDA: the whole layer as one autograd node (fused_layer.py)

    def _fused_layer_forward(self, query, value, query_pos, kwargs):
        """The vocc.py layer ('cross_attn', 'norm', 'ffn', 'norm') with fp16 storage as ONE autograd node with a
        hand-written backward (vln_ver_b200/fused_layer.py); returns None when the configuration is not covered."""
        from .. import fused_layer
        if not (self.fuse_layer and not self.pre_norm and query.is_cuda and query_pos is None
                and tuple(self.operation_order) == ('cross_attn', 'norm', 'ffn', 'norm')):
            return None
        attn, ffn = self.attentions[0], self.ffns[0]
        if not (isinstance(attn, SpatialCrossAttention) and attn.sampler != 'gather'
                and attn.compute_dtype == torch.float16 and getattr(ffn, 'compute_dtype', None) == torch.float16
                and getattr(ffn, 'fusable', lambda: False)()):
            return None
        vis = kwargs.get('visibility')
        hw = kwargs.get('spatial_hw')
        da = attn.deformable_attention
        if vis is None or hw is None or da.num_levels != 1 or value is None or value.dim() != 4:
            return None
        num_cams, l, bs, C = value.shape
        if query.shape[0] != bs or hw[0] * hw[1] != l:
            return None
        q2 = query.to(torch.float16).reshape(bs * query.shape[1], C)
        feat = value.to(torch.float16).permute(2, 0, 1, 3).reshape(bs * num_cams * l, C)
        first, last, last_drop = ffn.layers[0], ffn.layers[1], ffn.layers[2]
        if not fused_layer.supported(q2, feat, vis, da.num_heads, da.num_points, l, first[0].out_features):
            return None
        n1, n2 = self.norms[0], self.norms[1]
        cfg = dict(NH=da.num_heads, NP=da.num_points, Sh=int(hw[0]), Sw=int(hw[1]), training=self.training,
                   p_attn=attn.dropout.p, p_ffn=first[2].p, p_out=last_drop.p, eps1=n1.eps, eps2=n2.eps)
        params = (da.value_proj.weight, da.value_proj.bias, da.sampling_offsets.weight, da.sampling_offsets.bias,
                  da.attention_weights.weight, da.attention_weights.bias, attn.output_proj.weight,
                  attn.output_proj.bias, n1.weight, n1.bias, first[0].weight, first[0].bias, last.weight, last.bias,
                  n2.weight, n2.bias)
        out = fused_layer.voxel_layer(q2, feat, vis, cfg, params)
        return out.view(bs, query.shape[1], C)

    def forward(self, query, key=None, value=None, bev_pos=None, query_pos=None, key_pos=None,
                attn_masks=None, query_key_padding_mask=None, key_padding_mask=None, ref_2d=None,
                ref_3d=None, bev_z=None, bev_h=None, bev_w=None, reference_points_cam=None, mask=None,
                spatial_shapes=None, level_start_index=None, prev_bev=None, **kwargs):
        fused = self._fused_layer_forward(query, value, query_pos, kwargs)
        if fused is not None:
            return fused
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
            warnings.warn(f'Use same attn_mask in all attentions in {self.__class__.__name__} ')
        else:
            assert len(attn_masks) == self.num_attn
        order = self.operation_order
        skip_norm = False
        for op_i, layer in enumerate(order):
            # fused epilogue: [cross_attn|ffn] followed by 'norm' -> dropout + residual + LayerNorm in
            # one kernel (ver_dropout_add_layernorm_*); same arithmetic as the separate steps below
            fuse = (self.fuse_epilogues and not self.pre_norm and query.is_cuda
                    and op_i + 1 < len(order) and order[op_i + 1] == 'norm'
                    and query.shape[-1] % 8 == 0 and query.shape[-1] <= 1024)
            if layer == 'norm' and skip_norm:
                skip_norm = False
                norm_index += 1
                continue
            if fuse and layer == 'cross_attn' and isinstance(self.attentions[attn_index], SpatialCrossAttention):
                attn, norm = self.attentions[attn_index], self.norms[norm_index]
                proj, resid = attn(
                    query, key, value, None, query_pos=query_pos, key_pos=key_pos, reference_points=ref_3d,
                    reference_points_cam=reference_points_cam, mask=mask, attn_mask=attn_masks[attn_index],
                    key_padding_mask=key_padding_mask, spatial_shapes=spatial_shapes,
                    level_start_index=level_start_index, return_unfused=True, **kwargs)
                query = ops.dropout_add_layernorm(proj, resid.to(proj.dtype), norm.weight, norm.bias,
                                                  attn.dropout.p, norm.eps, attn.training)
                attn_index += 1
                identity = query
                skip_norm = True
                continue
            if fuse and layer == 'ffn' and getattr(self.ffns[ffn_index], 'fusable', lambda: False)():
                ffn, norm = self.ffns[ffn_index], self.norms[norm_index]
                out, p_last = ffn.forward_unfused_tail(query)
                query = ops.dropout_add_layernorm(out, query.to(out.dtype), norm.weight, norm.bias, p_last,
                                                  norm.eps, ffn.training)
                ffn_index += 1
                skip_norm = True
                continue
            if layer == 'self_attn':
                query = self.attentions[attn_index](
                    query, prev_bev, prev_bev, identity if self.pre_norm else None, query_pos=bev_pos,
                    key_pos=bev_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=query_key_padding_mask, reference_points=ref_2d,
                    spatial_shapes=[[bev_h, bev_w]], level_start_index=[0], **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'norm':
                query = apply_layernorm(self.norms[norm_index], query)
                norm_index += 1
            elif layer == 'cross_attn':
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=key_pos, reference_points=ref_3d, reference_points_cam=reference_points_cam,
                    mask=mask, attn_mask=attn_masks[attn_index], key_padding_mask=key_padding_mask,
                    spatial_shapes=spatial_shapes, level_start_index=level_start_index, **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'ffn':
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query
