"""VoxelPerceptionTransformer -- mirror of
projects/mmdet3d_plugin/bevformer/modules/voxel_transformer.py (:24-301): get_voxel_features
(encoder side, the hot path) and forward (encoder + detection decoder).

get_voxel_features (A8): the per-view tokens get cams_embeds + level_embeds added and are
re-laid out (Ncam, B, S, C) -> (B*Ncam, S, C) by ONE kernel (ver_feat_embed) instead of the
reference's reshape/permute/add/permute chain; the result is handed to the encoder as a
(Ncam, S, B, C) *view* so the public signature of the encoder / SCA is unchanged.
"""
import torch
import torch.nn as nn
from torch.autograd.function import Function, once_differentiable
from torch.nn.init import normal_

from .. import ops
from ..registry import TRANSFORMER, BaseModule, build_transformer_layer_sequence, xavier_init
from .precision import PrecisionMixin
from .spatial_cross_attention import MSDeformableAttention3D
from .voxel_decoder import VoxelDetectionTransformerDecoder  # noqa: F401  (registered there; re-exported)


class _FeatEmbed(Function):
    @staticmethod
    def forward(ctx, feats, cams_embeds, level_embed, dtype):
        ctx.shape = feats.shape
        ctx.has_cams = cams_embeds is not None
        return ops.feat_embed(feats, cams_embeds, level_embed, dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        Ncam, B, S, C = ctx.shape
        g = g.view(B, Ncam, S, C).float()
        g_feats = g.permute(1, 0, 2, 3) if ctx.needs_input_grad[0] else None
        g_cams = g.sum((0, 2)) if ctx.has_cams and ctx.needs_input_grad[1] else None
        g_level = g.sum((0, 1, 2)) if ctx.needs_input_grad[2] else None
        return g_feats, g_cams, g_level, None


@TRANSFORMER.register_module()
class VoxelPerceptionTransformer(PrecisionMixin, BaseModule):
    def __init__(self, num_feature_levels=4, num_cams=6, two_stage_num_proposals=300, encoder=None,
                 decoder=None, embed_dims=256, rotate_prev_bev=True, use_shift=True, use_can_bus=True,
                 can_bus_norm=True, use_cams_embeds=True, rotate_center=[100, 100],
                 decoder_on_bev=False, voxel_2_bev_type='mlp', bev_z=1, **kwargs):
        super().__init__(**kwargs)
        self.encoder = build_transformer_layer_sequence(encoder)
        self.decoder = build_transformer_layer_sequence(decoder) if decoder is not None else None
        self.embed_dims = embed_dims
        self.num_feature_levels = num_feature_levels
        self.num_cams = num_cams
        self.fp16_enabled = False
        self.rotate_prev_bev, self.use_shift, self.use_can_bus = rotate_prev_bev, use_shift, use_can_bus
        self.can_bus_norm, self.use_cams_embeds = can_bus_norm, use_cams_embeds
        self.decoder_on_bev, self.voxel_2_bev_type, self.bev_z = decoder_on_bev, voxel_2_bev_type, bev_z
        self.two_stage_num_proposals = two_stage_num_proposals
        self.init_layers()
        self.rotate_center = rotate_center

    def init_layers(self):
        self.level_embeds = nn.Parameter(torch.Tensor(self.num_feature_levels, self.embed_dims))
        self.cams_embeds = nn.Parameter(torch.Tensor(self.num_cams, self.embed_dims))
        if self.decoder is not None:
            self.reference_points = nn.Linear(self.embed_dims, 3)

    def init_weights(self):
        """xavier on every >1-D parameter, then the attention modules' own init, then N(0,1)
        embeddings (reference :99-116)."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformableAttention3D):
                m.init_weights()
        normal_(self.level_embeds)
        normal_(self.cams_embeds)
        if self.decoder is not None:
            xavier_init(self.reference_points, distribution='uniform', bias=0.)

    def get_voxel_features(self, mlvl_feats, bev_queries, bev_z, bev_h, bev_w, grid_length=[0.512, 0.512],
                           bev_pos=None, prev_bev=None, **kwargs):
        """mlvl_feats (Ncam, bs, S, C) ViT tokens (CLS dropped); bev_queries (Nq, C) ->
        voxel features (bs, Nq, C)   (reference :119-185)."""
        num_cam, bs, S, C = mlvl_feats.shape
        if self.use_cams_embeds and num_cam != self.num_cams:
            raise ValueError(f'{num_cam} views but cams_embeds has {self.num_cams} rows')
        h = w = int(round(S ** 0.5))
        if h * w != S:
            raise ValueError(f'{S} tokens per view is not a square map')
        cd = self.compute_dtype or torch.float32
        # cast the (Nq, C) table once, THEN broadcast over the batch (the reference repeats, :137)
        bev_queries = bev_queries.to(cd).unsqueeze(1).expand(-1, bs, -1)
        if bev_pos is not None:
            bev_pos = bev_pos.flatten(2).permute(2, 0, 1)
        feat = _FeatEmbed.apply(mlvl_feats, self.cams_embeds if self.use_cams_embeds else None,
                                self.level_embeds[0], cd)                     # (bs*Ncam, S, C)
        feat_flatten = feat.view(bs, num_cam, S, C).permute(1, 2, 0, 3)      # (Ncam, S, bs, C) view
        spatial_shapes = [[h, w]]
        return self.encoder(bev_queries.to(cd), feat_flatten, feat_flatten, bev_z=bev_z, bev_h=bev_h,
                            bev_w=bev_w, bev_pos=bev_pos, spatial_shapes=spatial_shapes,
                            level_start_index=[0], prev_bev=prev_bev, shift=None, **kwargs)

    def forward(self, mlvl_feats, bev_queries, object_query_embed, bev_z, bev_h, bev_w,
                grid_length=[0.512, 0.512], bev_pos=None, reg_branches=None, cls_branches=None,
                prev_bev=None, **kwargs):
        voxel_embed = self.get_voxel_features(mlvl_feats, bev_queries, bev_z, bev_h, bev_w,
                                              grid_length=grid_length, bev_pos=bev_pos,
                                              prev_bev=prev_bev, **kwargs)
        if self.decoder is None:
            return voxel_embed.permute(1, 0, 2), None, None, None
        if self.decoder_on_bev:
            raise NotImplementedError('decoder_on_bev=True (voxel2bev MLP / pool, reference :262-285) is not on '
                                      'the vocc.py path (vocc.py:113 sets decoder_on_bev=False)')
        # decoder half (reference :246-301): object queries -> reference points -> 6 decoder layers
        # reading the encoded voxel volume through the 3-D deformable sampler
        bs = mlvl_feats.shape[1]
        query_pos, query = torch.split(object_query_embed, self.embed_dims, dim=1)
        query_pos = query_pos.unsqueeze(0).expand(bs, -1, -1)
        query = query.unsqueeze(0).expand(bs, -1, -1)
        reference_points = self.reference_points(query_pos).sigmoid()
        init_reference_out = reference_points
        query = query.permute(1, 0, 2)
        query_pos = query_pos.permute(1, 0, 2)
        voxel_embed = voxel_embed.permute(1, 0, 2)
        inter_states, inter_references = self.decoder(
            query=query, key=None, value=voxel_embed, query_pos=query_pos, reference_points=reference_points,
            reg_branches=reg_branches, cls_branches=cls_branches, spatial_shapes=[[bev_z, bev_h, bev_w]],
            level_start_index=[0], **kwargs)
        return voxel_embed, inter_states, init_reference_out, inter_references
