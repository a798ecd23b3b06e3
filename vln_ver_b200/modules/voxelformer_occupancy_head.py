"""VoxelFormerOccupancyHead -- occupancy part of
projects/mmdet3d_plugin/bevformer/dense_heads/voxelformer_occupancy_head.py (HEAD):
query embedding -> VoxelPerceptionTransformer -> [up_sample] -> occ_proj -> occ_branches
(HEAD:300-308, :323-352, :551-580), sigmoid focal occupancy loss (HEAD:1386-1444) and the
sparse decode (HEAD:1505-1524).  With a decoder in the transformer cfg (vocc.py default,
only_occ=False) the DETR-style forward tail is built too: query embeddings,
cls / reg / layout branches and the box decoding of HEAD:368-625.  The detection LOSSES
(Hungarian assignment, L1 / GIoU) are a different task and are not built.
"""
import copy

import torch
import torch.nn as nn

from .. import ops
from ..registry import (HAVE_MMCV, HEADS, LOSSES, BaseModule, bias_init_with_prob,
                        build_loss, build_positional_encoding, build_transformer)
from ..upsample import (lattice_supported, occ_proj_from_lattice, occ_proj_plan_supported, up_sample, up_sample_gemm,
                        up_sample_lattice)
from .precision import PrecisionMixin
from .voxel_decoder import inverse_sigmoid
from .voxel_encoder import apply_layernorm


class FocalLoss(nn.Module):
    """mmdet FocalLoss(use_sigmoid=True) surface (vocc.py:190-195) on ver_focal_loss."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, 'Only sigmoid focal loss supported now.'
        self.use_sigmoid, self.gamma, self.alpha = use_sigmoid, gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        """pred (N, C) logits, target (N,) class ids in [0, C] (C = background)."""
        assert weight is None, 'per-sample weights are not used on the occupancy path'
        reduction = reduction_override if reduction_override else self.reduction
        assert reduction == 'mean'
        loss_sum, _ = ops._FocalFunction.apply(pred, target, self.gamma, self.alpha)
        denom = avg_factor if avg_factor is not None else pred.numel()
        return self.loss_weight * loss_sum[0] / denom


class _CfgHolder(nn.Module):
    """Stands in for detection-only losses so vocc.py's loss_cls / loss_bbox / loss_iou dicts build."""

    def __init__(self, **cfg):
        super().__init__()
        self.cfg = cfg
        self.use_sigmoid = cfg.get('use_sigmoid', False)
        self.loss_weight = cfg.get('loss_weight', 1.0)

    def forward(self, *a, **k):
        raise NotImplementedError('detection losses are outside the lift+encode hot path')


if not HAVE_MMCV or LOSSES.get('FocalLoss') is None:
    LOSSES.register_module(name='FocalLoss', module=FocalLoss)
    for _n in ('L1Loss', 'GIoULoss'):
        LOSSES.register_module(name=_n, module=type(_n, (_CfgHolder,), {}))


@HEADS.register_module()
class VoxelFormerOccupancyHead(PrecisionMixin, BaseModule):
    def __init__(self, *args, with_box_refine=True, as_two_stage=False, transformer=None,
                 bbox_coder=None, num_cls_fcs=2, code_weights=None, bev_h=120, bev_w=120, bev_z=4,
                 num_layout_query=10, getbev=None, occupancy_size=[0.1, 0.1, 0.1],
                 point_cloud_range=[-6.0, -6.0, -1.5, 6.0, 6.0, 2.0], loss_layout=None,
                 loss_occupancy=None, loss_flow=None, flow_gt_dimension=2, occ_dims=16, det_dims=None,
                 num_occ_fcs=2, occupancy_classes=1, only_occ=False, only_det=False, add_layout=False,
                 with_occupancy_flow=False, with_color_render=False, occ_weights=None,
                 flow_weights=None, occ_loss_type='focal_loss', occ_head_type='mlp',
                 occ_head_network=None, refine_occ=False,
                 # DETRHead kwargs that vocc.py passes (HEAD:87-195)
                 num_classes=17, in_channels=768, num_query=100, num_reg_fcs=2,
                 sync_cls_avg_factor=False, positional_encoding=None, loss_cls=None, loss_bbox=None,
                 loss_iou=None, train_cfg=None, test_cfg=None, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        self.bev_h, self.bev_w, self.bev_z = bev_h, bev_w, bev_z
        self.fp16_enabled = False
        self.only_occ, self.only_det, self.add_layout = only_occ, only_det, add_layout
        self.occ_loss_type, self.occ_head_type = occ_loss_type, occ_head_type
        self.refine_occ = refine_occ
        # 'auto': lattice form (GEMM + col2im kernel on CUDA); 'gemm' / 'lattice': force one execution of it;
        # 'dense': the three ConvTranspose3d as written (A/B switch for tests and tools/upsample_bench.py)
        self.up_sample_mode = 'auto'
        # 'lattice': occ_proj straight from the lattice data, skipping the bias constants (upsample.occ_proj_from_lattice;
        # pinned in fp64 on CPU, not yet measured on the GPU -> off by default)
        self.occ_proj_mode = 'dense'
        self.getbev = getbev
        self.with_box_refine, self.as_two_stage = with_box_refine, as_two_stage
        self.num_classes, self.in_channels, self.num_query = num_classes, in_channels, num_query
        self.num_reg_fcs, self.num_layout_query = num_reg_fcs, num_layout_query
        self.code_size = kwargs.get('code_size', 10)                          # HEAD:111-114
        self.layout_range = [-50.0, -50.0, -5.0, 50.0, 50.0, 5.0]             # HEAD:91
        # mmdet DETRHead: sigmoid classification has no background column
        self.cls_out_channels = num_classes if (loss_cls or {}).get('use_sigmoid', False) else num_classes + 1
        self.occ_weights = occ_weights
        self.pc_range = (bbox_coder or {}).get('pc_range', point_cloud_range)
        self.real_w = self.pc_range[3] - self.pc_range[0]
        self.real_h = self.pc_range[4] - self.pc_range[1]
        self.real_z = self.pc_range[5] - self.pc_range[2]
        self.occupancy_size = occupancy_size
        self.point_cloud_range = point_cloud_range
        # HEAD:144-146 -- Python float division then int() truncation
        self.occ_xdim = int((point_cloud_range[3] - point_cloud_range[0]) / occupancy_size[0])
        self.occ_ydim = int((point_cloud_range[4] - point_cloud_range[1]) / occupancy_size[1])
        self.occ_zdim = int((point_cloud_range[5] - point_cloud_range[2]) / occupancy_size[2])
        self.occ_dims, self.num_occ_fcs = occ_dims, num_occ_fcs
        self.occupancy_classes = occupancy_classes
        self.voxel_num = self.occ_xdim * self.occ_ydim * self.occ_zdim
        self.bev_num = self.bev_h * self.bev_w * self.bev_z
        transformer = copy.deepcopy(dict(transformer))
        if self.only_occ:
            transformer['decoder'] = None                                     # HEAD:160-161
        self.positional_encoding = build_positional_encoding(positional_encoding)
        self.transformer = build_transformer(transformer)
        self.embed_dims = self.transformer.embed_dims
        assert positional_encoding['num_feats'] * 2 == self.embed_dims
        self.loss_occupancy = build_loss(loss_occupancy)
        self.loss_cls = build_loss(loss_cls) if (loss_cls and not only_occ) else None
        # frozen nn.Parameter in the reference too (HEAD:115-123, :161-162): part of its state_dict
        self.code_weights = nn.Parameter(torch.tensor(
            list(code_weights) if code_weights is not None else [1.0] * 8 + [0.0, 0.0]), requires_grad=False)
        self._init_layers()

    def _init_layers(self):
        """detection branches when a decoder exists (HEAD:179-231), then voxel_embedding, occ_proj,
        occ_branches, up_sample (HEAD:226-258)."""
        if self.transformer.decoder is not None:
            C = self.embed_dims

            def mlp(out, norm):
                layers = []
                for _ in range(self.num_reg_fcs):
                    layers += [nn.Linear(C, C)] + ([nn.LayerNorm(C), nn.ReLU(inplace=True)] if norm else [nn.ReLU()])
                return nn.Sequential(*layers, nn.Linear(C, out))
            num_pred = self.transformer.decoder.num_layers + (1 if self.as_two_stage else 0)

            def clones(m):
                if self.with_box_refine:
                    return nn.ModuleList([copy.deepcopy(m) for _ in range(num_pred)])
                return nn.ModuleList([m for _ in range(num_pred)])
            self.cls_branches = clones(mlp(self.cls_out_channels, True))
            self.reg_branches = clones(mlp(self.code_size, False))
            self.layout_branches = clones(mlp(self.code_size, False))
            if not self.as_two_stage:
                self.query_embedding = nn.Embedding(self.num_query, C * 2)
                self.query_layout_embedding = nn.Embedding(self.num_layout_query, C * 2)
        self.voxel_embedding = nn.Embedding(self.bev_num, self.embed_dims)
        if self.bev_z == self.occ_zdim:
            self.occ_proj = nn.Linear(self.embed_dims, self.occ_dims)
        else:
            self.occ_proj = nn.Linear(self.bev_z * self.embed_dims, self.occ_dims * self.occ_zdim)
        occ_branch = []
        for _ in range(self.num_occ_fcs):
            occ_branch += [nn.Linear(self.occ_dims, self.occ_dims), nn.LayerNorm(self.occ_dims),
                           nn.ReLU(inplace=True)]
        occ_branch.append(nn.Linear(self.occ_dims, self.occupancy_classes))
        self.occ_branches = nn.Sequential(*occ_branch)
        if self.refine_occ:
            # 8x lateral upsampling of the voxel volume, library conv
            self.up_sample = nn.Sequential(*[
                nn.ConvTranspose3d(768, 768, (3, 5, 5), stride=(1, 2, 2), padding=(2, 4, 4),
                                   dilation=(2, 2, 2), output_padding=(0, 1, 1)) for _ in range(3)])

    def init_weights(self):
        self.transformer.init_weights()
        self.positional_encoding.init_weights()
        if self.loss_cls is not None and self.loss_cls.use_sigmoid and hasattr(self, 'cls_branches'):
            for m in self.cls_branches:                                        # HEAD:272-276
                nn.init.constant_(m[-1].bias, bias_init_with_prob(0.01))
        if self.loss_occupancy.use_sigmoid:
            nn.init.constant_(self.occ_branches[-1].bias, bias_init_with_prob(0.01))

    # ------------------------------------------------------------------ A10
    def _occupancy_tail(self, bev_embed, bs):
        """bev_embed (bs, Nq, C) -> occupancy logits (bs, occ_z*occ_y*occ_x, classes).
        Per-sample semantics of the raw `.view`s at HEAD:334 / :558 / :564."""
        cd = self.compute_dtype or bev_embed.dtype
        C = self.embed_dims
        x = bev_embed.contiguous()
        # the add_layout branch of the reference never up-samples (HEAD:459-470), the only_occ branch neither (:334)
        if self.refine_occ and not self.only_occ and not self.add_layout:
            x = x.view(bs, C, self.bev_z, self.bev_h, self.bev_w)
            w_dtype = cd
            if (self.occ_proj_mode == 'lattice' and self.bev_z != self.occ_zdim and lattice_supported(self.up_sample)
                    and self.up_sample_mode != 'dense'
                    and occ_proj_plan_supported(C, self.bev_z, 8 * self.bev_h, 8 * self.bev_w, self.occ_xdim,
                                                self.occ_ydim)):
                fn = up_sample_lattice if self.up_sample_mode == 'lattice' else up_sample_gemm
                e, last_bias = fn(x, self.up_sample, dtype=w_dtype, assemble=False)
                occ = occ_proj_from_lattice(e, last_bias, self.occ_proj.weight, self.occ_proj.bias, self.occ_xdim,
                                            self.occ_ydim, dtype=w_dtype)
                occ = occ.view(bs, self.occ_xdim, self.occ_ydim, self.occ_zdim, self.occ_dims).permute(0, 3, 1, 2, 4)
                return self._occ_branches_tail(occ.reshape(bs, -1, self.occ_dims), cd)
            mode = self.up_sample_mode if lattice_supported(self.up_sample) else 'dense'
            if mode != 'dense':
                # same values with 3.5x fewer FLOPs: the data of every layer lives on the even-even
                # lattice of its output, the rest is the bare bias (vln_ver_b200/upsample.py)
                fn = {'auto': up_sample, 'gemm': up_sample_gemm, 'lattice': up_sample_lattice}[mode]
                x = fn(x, self.up_sample, dtype=w_dtype)
            else:
                for conv in self.up_sample:
                    x = nn.functional.conv_transpose3d(
                        x.to(w_dtype), conv.weight.to(w_dtype), conv.bias.to(w_dtype), stride=conv.stride,
                        padding=conv.padding, output_padding=conv.output_padding, dilation=conv.dilation)
            x = x.contiguous().view(bs, self.bev_z, self.occ_xdim, self.occ_ydim, C)
            lat = (self.occ_xdim, self.occ_ydim)
        else:
            x = x.view(bs, self.bev_z, self.bev_h, self.bev_w, C)
            lat = (self.bev_h, self.bev_w)
        if self.bev_z == self.occ_zdim:
            occ = self._linear(x, self.occ_proj, cd)
        else:
            x = x.permute(0, 2, 3, 1, 4).flatten(3)
            occ = self._linear(x, self.occ_proj, cd)
            occ = occ.view(bs, lat[0], lat[1], self.occ_zdim, self.occ_dims).permute(0, 3, 1, 2, 4)
        return self._occ_branches_tail(occ.reshape(bs, -1, self.occ_dims), cd)

    def _occ_branches_tail(self, y, cd):
        """occ_branches (HEAD:242-248, applied at :580) on (bs, voxels, occ_dims)."""
        for layer in self.occ_branches:
            if isinstance(layer, nn.Linear):
                y = self._linear(y, layer, cd)
            elif isinstance(layer, nn.LayerNorm):
                y = apply_layernorm(layer, y)
            else:
                y = layer(y)
        return y.float()

    def forward(self, mlvl_feats, img_metas, prev_bev=None, only_bev=False, **cam):
        """mlvl_feats (Ncam, bs, S, C).  `cam` may carry device tensors lidar2img (bs, Ncam, 4, 4)
        and originshift (bs, 3) instead of img_metas lookups.  Returns the reference's dict
        (HEAD:357-367 / :615-625); detection entries are None."""
        num_cam, bs, _, _ = mlvl_feats.shape
        voxel_queries = self.voxel_embedding.weight
        needs_pos = any('self_attn' in l.operation_order for l in self.transformer.encoder.layers)
        voxel_pos = None
        if needs_pos:       # A9: only a self-attention layer ever reads it
            voxel_pos = self.positional_encoding(
                torch.zeros((bs, self.bev_z, self.bev_h, self.bev_w), device=voxel_queries.device))
        feat_args = dict(grid_length=(self.real_h / self.bev_h, self.real_w / self.bev_w), bev_pos=voxel_pos,
                         img_metas=img_metas, prev_bev=prev_bev, **cam)
        if only_bev or self.only_occ or self.transformer.decoder is None:
            bev_embed = self.transformer.get_voxel_features(
                mlvl_feats, voxel_queries, self.bev_z, self.bev_h, self.bev_w, **feat_args)
            if only_bev:
                return bev_embed
            return {'bev_embed': bev_embed if self.only_occ else bev_embed.permute(1, 0, 2),
                    'all_cls_scores': None, 'all_bbox_preds': None, 'all_layout_preds': None,
                    'occupancy_preds': self._occupancy_tail(bev_embed, bs), 'flow_preds': None,
                    'enc_cls_scores': None, 'enc_bbox_preds': None, 'enc_occupancy_preds': None}
        # lift + encode + decode (HEAD:368-625; only_det :370-430, add_layout :431-535, default :537-625)
        bev_embed, hs, init_reference, inter_references = self.transformer(
            mlvl_feats, voxel_queries, self.query_embedding.weight, self.bev_z, self.bev_h, self.bev_w,
            reg_branches=self.reg_branches if self.with_box_refine else None,
            cls_branches=self.cls_branches if self.as_two_stage else None, **feat_args)
        cls, boxes, layouts = self._detection_tail(hs, init_reference, inter_references)
        if self.getbev is not None:                  # HEAD:627-638: append the voxel features to an HDF5 file
            from ..ingest import export_bev_embed
            export_bev_embed(self.getbev, img_metas, bev_embed, self.embed_dims, self.bev_z, self.bev_h, self.bev_w)
        return {'bev_embed': bev_embed, 'all_cls_scores': cls, 'all_bbox_preds': boxes,
                'all_layout_preds': layouts if self.add_layout else None,
                'occupancy_preds': None if self.only_det else self._occupancy_tail(bev_embed.permute(1, 0, 2), bs),
                'flow_preds': None, 'enc_cls_scores': None, 'enc_bbox_preds': None, 'enc_occupancy_preds': None}

    def _detection_tail(self, hs, init_reference, inter_references):
        """hs (num_dec, num_query, bs, C) -> class scores (num_dec, bs, num_query, cls_out) and boxes
        (num_dec, bs, num_query, code_size) with centre x, y (slots 0, 1) and z (slot 4) offset from the
        layer's input reference point, squashed and mapped into pc_range (HEAD:590-611); layout boxes
        likewise into layout_range (HEAD:497-511)."""
        hs = hs.permute(0, 2, 1, 3).float()
        classes, coords, layouts = [], [], []

        def place(t, reference, rng):
            xy = (t[..., 0:2] + reference[..., 0:2]).sigmoid()
            z = (t[..., 4:5] + reference[..., 2:3]).sigmoid()
            return torch.cat([xy[..., 0:1] * (rng[3] - rng[0]) + rng[0], xy[..., 1:2] * (rng[4] - rng[1]) + rng[1],
                              t[..., 2:4], z * (rng[5] - rng[2]) + rng[2], t[..., 5:]], -1)
        for lvl in range(hs.shape[0]):
            reference = inverse_sigmoid(init_reference if lvl == 0 else inter_references[lvl - 1])
            assert reference.shape[-1] == 3
            classes.append(self.cls_branches[lvl](hs[lvl]))
            coords.append(place(self.reg_branches[lvl](hs[lvl]), reference, self.pc_range))
            if self.add_layout:
                layouts.append(place(self.layout_branches[lvl](hs[lvl]), reference, self.layout_range))
        return torch.stack(classes), torch.stack(coords), (torch.stack(layouts) if layouts else None)

    # ------------------------------------------------------------------ A11
    def loss_only_occupancy(self, gt_bboxes_list, gt_labels_list, point_coords, occ_gts, flow_gts,
                            preds_dicts, gt_bboxes_ignore=None, img_metas=None):
        """occ_gts: per panorama an (n, 2) int64 tensor (flat index, class) -- the reference's
        `occ_gts[0][0]` (HEAD:1408) generalised to a batch; the batch loss is the mean of the
        per-panorama losses."""
        preds = preds_dicts['occupancy_preds']
        losses = []
        for b in range(preds.shape[0]):
            gt = occ_gts[b]
            gt = gt[0] if isinstance(gt, (list, tuple)) else gt
            losses.append(ops.occupancy_focal_loss(
                preds[b].reshape(-1, self.occupancy_classes), gt.to(preds.device),
                gamma=self.loss_occupancy.gamma, alpha=self.loss_occupancy.alpha,
                loss_weight=self.loss_occupancy.loss_weight))
        loss = torch.stack(losses).mean()
        return {'loss_occupancy': loss, 'loss_flow': torch.zeros_like(loss)}

    def loss(self, gt_bboxes_list, gt_labels_list, point_coords, occ_gts, flow_gts, preds_dicts,
             gt_bboxes_ignore=None, img_metas=None):
        """Default-branch loss (only_occ=False, HEAD:1284-1385) restricted to the hot path's share of it: the
        occupancy term the reference computes in the LAST decoder layer's loss_single (dense target filled with
        class `occupancy_classes`, sparse GT scattered in, avg_factor = number of occupied voxels, sigmoid focal loss,
        nan_to_num -- HEAD:1324-1332 and :977-986; the same arithmetic as loss_only_occupancy, :1386-1444) and the
        zero flow term (`loss_flow = zeros_like`, :982).  Keys follow the reference's loss dict (:1355-1357: the last
        decoder layer's entries are un-prefixed).  The detection terms (loss_cls / loss_bbox per decoder layer:
        Hungarian assignment, L1 / GIoU) are outside the lift+encode path and are not computed:
        asking for them raises."""
        if preds_dicts.get('occupancy_preds') is None:
            raise NotImplementedError('only the occupancy / flow terms of the default-branch loss are on the hot path; '
                                      'the detection losses (HEAD:1336-1353) are out of scope')
        return self.loss_only_occupancy(gt_bboxes_list, gt_labels_list, point_coords, occ_gts, flow_gts, preds_dicts,
                                        gt_bboxes_ignore=gt_bboxes_ignore, img_metas=img_metas)

    # ------------------------------------------------------------------ A12
    def get_occupancy_prediction(self, occ_results, occ_threshold=0.25):
        if self.occ_loss_type != 'focal_loss':
            raise NotImplementedError(self.occ_loss_type)
        occ_results['occupancy_preds'] = ops.occupancy_decode(
            occ_results['occupancy_preds'].reshape(-1, self.occupancy_classes), occ_threshold)
        occ_results['flow_preds'] = None
        return occ_results
