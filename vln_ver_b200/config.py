"""The `pts_bbox_head` sub-tree of projects/configs/verformer/vocc.py (:87-195) as a
function of the quantities the benchmark configs vary (grid, views, head variant).
Everything else is the shipped value."""
import copy

POINT_CLOUD_RANGE = [-6.0, -6.0, -1.5, 6.0, 6.0, 2.0]      # vocc.py:9


def vocc_head_cfg(bev_z=4, bev_h=15, bev_w=15, num_cams=6, embed_dims=768, only_occ=False,
                  refine_occ=True, occupancy_size=(0.1, 0.1, 0.1), occ_dims=128, num_layers=3,
                  pc_range=POINT_CLOUD_RANGE, with_decoder=None, ffn_dims=None, num_decoder_layers=6,
                  num_query=100):
    """Defaults reproduce vocc.py (15x15x4, refine_occ=True, full decoder tree).  For the
    per-voxel head of the grid sweeps use `occupancy_size=per_voxel_occupancy_size(...)`,
    `refine_occ=False`, `only_occ=True`."""
    _dim_ = embed_dims
    with_decoder = (not only_occ) if with_decoder is None else with_decoder
    decoder = dict(
        type='VoxelDetectionTransformerDecoder', num_layers=num_decoder_layers, return_intermediate=True,
        transformerlayers=dict(
            type='DetrTransformerDecoderLayer',
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=_dim_, num_heads=8, dropout=0.1),
                       dict(type='VoxelCustomMSDeformableAttention', embed_dims=_dim_, num_levels=1)],
            ffn_cfgs=dict(type='FFN', embed_dims=768, feedforward_channels=1024, num_fcs=2, ffn_drop=0.,
                          act_cfg=dict(type='ReLU', inplace=True)),
            feedforward_channels=_dim_ * 2, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))
    cfg = dict(
        type='VoxelFormerOccupancyHead', bev_h=bev_h, bev_w=bev_w, bev_z=bev_z, getbev=None,
        num_query=num_query, num_classes=17, in_channels=_dim_, sync_cls_avg_factor=True,
        with_box_refine=True, as_two_stage=False, point_cloud_range=list(pc_range),
        occupancy_size=list(occupancy_size), occ_dims=occ_dims, occupancy_classes=16,
        only_occ=only_occ, only_det=False, refine_occ=refine_occ,
        transformer=dict(
            type='VoxelPerceptionTransformer', rotate_prev_bev=True, use_shift=True, use_can_bus=True,
            embed_dims=_dim_, decoder_on_bev=False, num_cams=num_cams,
            encoder=dict(
                type='VoxelFormerEncoder', num_layers=num_layers, pc_range=list(pc_range),
                num_points_in_voxel=4, return_intermediate=False,
                transformerlayers=dict(
                    type='VoxelFormerLayer',
                    attn_cfgs=[dict(type='SpatialCrossAttention', pc_range=list(pc_range),
                                    num_cams=num_cams,
                                    deformable_attention=dict(type='MSDeformableAttention3D',
                                                              embed_dims=_dim_, num_points=8,
                                                              num_levels=1),
                                    embed_dims=_dim_)],
                    feedforward_channels=ffn_dims or _dim_ * 2, ffn_dropout=0.1,
                    operation_order=('cross_attn', 'norm', 'ffn', 'norm'))),
            decoder=decoder if with_decoder else None),
        bbox_coder=dict(type='NMSFreeCoder', post_center_range=[-10, -10, -5.0, 10, 10, 5.0],
                        pc_range=list(pc_range), max_num=50, voxel_size=[0.2, 0.2, 8], num_classes=17),
        positional_encoding=dict(type='VoxelLearnedPositionalEncoding', num_feats=_dim_ // 2,
                                 row_num_embed=bev_h, col_num_embed=bev_w, z_num_embed=bev_z),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
        loss_bbox=dict(type='L1Loss', loss_weight=0.25),
        loss_iou=dict(type='GIoULoss', loss_weight=0.0),
        loss_occupancy=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0))
    if embed_dims != 768:
        # the layer's default ffn_cfgs hard-codes embed_dims=768 (custom_base_transformer_layer.py:74-81)
        cfg['transformer']['encoder']['transformerlayers']['ffn_cfgs'] = dict(
            type='FFN', embed_dims=_dim_, feedforward_channels=1024, num_fcs=2, ffn_drop=0.,
            act_cfg=dict(type='ReLU', inplace=True))
        if with_decoder:
            decoder['transformerlayers']['ffn_cfgs']['embed_dims'] = _dim_      # vocc.py:148 hard-codes 768 too
    return copy.deepcopy(cfg)


def per_voxel_occupancy_size(bev_z, bev_h, bev_w, pc_range=POINT_CLOUD_RANGE):
    """occupancy_size that makes occ grid == voxel grid (bev_z == occ_zdim branch, HEAD:236-237)."""
    import math
    size = []
    for ext, n in ((pc_range[3] - pc_range[0], bev_w), (pc_range[4] - pc_range[1], bev_h),
                   (pc_range[5] - pc_range[2], bev_z)):
        s = ext / n
        while int(ext / s) != n:          # guard against int() truncation of ext / s
            s = math.nextafter(s, 0.0)
        size.append(s)
    return size
