"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of VER's 2D->3D
volumetric-lifting hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product package
never does.

Every function cites the reference lines it follows.  Shorthands:
  M/   = projects/mmdet3d_plugin/bevformer/modules/
  HEAD = projects/mmdet3d_plugin/bevformer/dense_heads/voxelformer_occupancy_head.py

Pinning status (see tests/test_oracle_pins_reference.py, oracle/gen_golden.py):
  * multi_scale_deformable_attn_pytorch : third-party (mmcv-full==1.4.0, not vendored).
    Pinned to the reference-owned 3-D function M/voxel_temporal_self_attention.py:275-335
    at depth 1 and to the golden vectors generated from it.
  * point_sampling / SCA / MSDA3D / encoder / positional encoding / only_occ head:
    pinned against the UNMODIFIED reference classes imported from /root/reference
    through oracle/mmcv_shim.py (golden fixtures under tests/golden/).
  * refine_occ (default-branch) head tail, focal loss (mmdet 2.14 FocalLoss CPU path),
    occupancy decode: restated; the reference ships no test for them ("parity
    unpinned" by the reference itself).
  * SURVEY 8(f) N2/N3 (bottom of this file): the 3-D sampler is the reference's own function
    (bit-equal, values and gradients); VoxelCustomMSDeformableAttention, VoxelDetectionTransformerDecoder,
    VoxelTemporalSelfAttention and the head's detection tail are pinned against the unmodified classes
    (tests/golden/decoder_c64.npz, head_detection_c32.npz); mmcv's MultiheadAttention wrapper is
    third-party (mmcv-full==1.4.0) and restated.

Batched semantics: the reference only runs bs=1.  For B>1 the oracle
is "run the bs=1 reference on each panorama with its own camera matrices and
concatenate".
"""
import torch
import torch.nn.functional as F

IMG_W, IMG_H = 1280, 1024          # hard-coded at M/voxel_encoder.py:179-180


# --------------------------------------------------------------------------- A5 / K3
# Storage-rounding emulation (tests only).  The product's fp16-storage mode rounds activations to fp16 at fixed
# points (value_proj output, sampler output, every GEMM output, every LayerNorm output, the view tokens and the
# query table); the reference has no such mode.  With STORAGE_ROUND = lambda t: t.half().to(t.dtype) the functions
# below round at the same points, which separates "storage rounding" from "arithmetic differences" when the fp16
# product is compared with the fp64 oracle (tests/test_gpu_parity.py::test_fp16_residual_is_storage_rounding).
STORAGE_ROUND = None


def _r(t):
    return t if STORAGE_ROUND is None else STORAGE_ROUND(t)


def multi_scale_deformable_attn_pytorch(value, value_spatial_shapes,
                                        sampling_locations, attention_weights):
    """mmcv 1.4.0 `multi_scale_deformable_attn_pytorch` (2-D), restated line by
    line from the reference-owned 3-D version
    M/voxel_temporal_self_attention.py:275-335 with (H, W) in place of (D, H, W).
    Call site in the reference: M/spatial_cross_attention.py:396-398.

    value (bs, num_keys, num_heads, Dh); value_spatial_shapes (num_levels, 2) = (h, w);
    sampling_locations (bs, nq, num_heads, num_levels, num_points, 2) = (x, y) in [0,1];
    attention_weights (bs, nq, num_heads, num_levels, num_points) -> (bs, nq, num_heads*Dh)
    """
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, num_heads, num_levels, num_points, _ = sampling_locations.shape
    value_list = value.split([int(H_) * int(W_) for H_, W_ in value_spatial_shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for level, (H_, W_) in enumerate(value_spatial_shapes):
        H_, W_ = int(H_), int(W_)
        # bs, H_*W_, num_heads, embed_dims -> bs*num_heads, embed_dims, H_, W_
        value_l_ = value_list[level].flatten(2).transpose(1, 2).reshape(
            bs * num_heads, embed_dims, H_, W_)
        # bs, nq, num_heads, num_points, 2 -> bs*num_heads, nq, num_points, 2
        sampling_grid_l_ = sampling_grids[:, :, :, level].transpose(1, 2).flatten(0, 1)
        sampling_value_l_ = F.grid_sample(
            value_l_, sampling_grid_l_, mode='bilinear', padding_mode='zeros',
            align_corners=False)
        sampling_value_list.append(sampling_value_l_)
    attention_weights = attention_weights.transpose(1, 2).reshape(
        bs * num_heads, 1, num_queries, num_levels * num_points)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) *
              attention_weights).sum(-1).view(bs, num_heads * embed_dims, num_queries)
    return output.transpose(1, 2).contiguous()


# --------------------------------------------------------------------------- A1
def get_reference_points_3d(bev_z, bev_h, bev_w, bs=1, dtype=torch.float32):
    """M/voxel_encoder.py:53-83 (dim='3d').  Returns (bs, 1, Nq, 3), last dim (x, y, z),
    flat order z-major then h then w; `num_points_in_voxel` is ignored (D == 1)."""
    zs = torch.linspace(0.5, bev_z - 0.5, bev_z, dtype=dtype).view(1, bev_z, 1, 1).expand(
        1, bev_z, bev_h, bev_w) / bev_z
    ys = torch.linspace(0.5, bev_h - 0.5, bev_h, dtype=dtype).view(1, bev_h, 1).expand(
        1, bev_z, bev_h, bev_w) / bev_h
    xs = torch.linspace(0.5, bev_w - 0.5, bev_w, dtype=dtype).view(1, 1, bev_w).expand(
        1, bev_z, bev_h, bev_w) / bev_w
    ref_3d = torch.stack((xs, ys, zs), -1)
    ref_3d = ref_3d.permute(0, 4, 1, 2, 3).flatten(2).permute(0, 2, 1)
    return ref_3d[None].repeat(bs, 1, 1, 1)


# --------------------------------------------------------------------------- A2
def point_sampling(reference_points, pc_range, lidar2img, originshift):
    """M/voxel_encoder.py:136-195 (the math after the JSON/pickle reads), for ONE
    panorama.  reference_points (1, D, Nq, 3) fp32; lidar2img (Ncam, 4, 4);
    originshift (3,).  Returns reference_points_cam (Ncam, 1, Nq, D, 2) fp32 and
    bev_mask (Ncam, 1, Nq, D) bool."""
    originshifts = reference_points.new_tensor(originshift)
    lidar2img = reference_points.new_tensor(lidar2img)[None]          # (1, Ncam, 4, 4)
    reference_points = reference_points.clone()
    reference_points[..., 0:1] = reference_points[..., 0:1] * \
        (pc_range[3] - pc_range[0]) + pc_range[0] + originshifts[0]
    reference_points[..., 1:2] = reference_points[..., 1:2] * \
        (pc_range[4] - pc_range[1]) + pc_range[1] + originshifts[1]
    reference_points[..., 2:3] = reference_points[..., 2:3] * \
        (pc_range[5] - pc_range[2]) + pc_range[2] + originshifts[2]
    reference_points = torch.cat(
        (reference_points, torch.ones_like(reference_points[..., :1])), -1)
    reference_points = reference_points.permute(1, 0, 2, 3)
    D, B, num_query = reference_points.size()[:3]
    num_cam = lidar2img.size(1)
    reference_points = reference_points.view(
        D, B, 1, num_query, 4).repeat(1, 1, num_cam, 1, 1).unsqueeze(-1)
    lidar2img = lidar2img.view(1, B, num_cam, 1, 4, 4).repeat(D, 1, 1, num_query, 1, 1)
    reference_points_cam = torch.matmul(lidar2img.to(torch.float32),
                                        reference_points.to(torch.float32)).squeeze(-1)
    eps = 1e-5
    bev_mask = (reference_points_cam[..., 2:3] > eps)
    reference_points_cam = reference_points_cam[..., 0:2] / torch.maximum(
        reference_points_cam[..., 2:3], torch.ones_like(reference_points_cam[..., 2:3]) * eps)
    reference_points_cam[..., 0] /= IMG_W
    reference_points_cam[..., 1] /= IMG_H
    bev_mask = (bev_mask & (reference_points_cam[..., 1:2] > 0.0)
                & (reference_points_cam[..., 1:2] < 1.0)
                & (reference_points_cam[..., 0:1] < 1.0)
                & (reference_points_cam[..., 0:1] > 0.0))
    bev_mask = torch.nan_to_num(bev_mask)
    reference_points_cam = reference_points_cam.permute(2, 1, 3, 0, 4)
    bev_mask = bev_mask.permute(2, 1, 3, 0, 4).squeeze(-1)
    return reference_points_cam, bev_mask


def point_sampling_batched(bev_z, bev_h, bev_w, pc_range, lidar2img, originshift):
    """Batched A1+A2: lidar2img (B, Ncam, 4, 4), originshift (B, 3) ->
    reference_points_cam (Ncam, B, Nq, 1, 2), bev_mask (Ncam, B, Nq, 1)."""
    B = lidar2img.shape[0]
    ref_3d = get_reference_points_3d(bev_z, bev_h, bev_w, bs=1)
    rpcs, masks = [], []
    for b in range(B):
        r, m = point_sampling(ref_3d, pc_range, lidar2img[b], originshift[b])
        rpcs.append(r)
        masks.append(m)
    return torch.cat(rpcs, 1), torch.cat(masks, 1)


def visible_indexes(bev_mask):
    """Per-camera visible-voxel index tensors of ONE panorama,
    M/spatial_cross_attention.py:138-142.  bev_mask (Ncam, 1, Nq, D)."""
    return [m[0].sum(-1).nonzero().squeeze(-1) for m in bev_mask]


# --------------------------------------------------------------------------- A4
def msda3d_forward(sd, pre, query, value, reference_points, spatial_shapes,
                   num_heads=8, num_levels=1, num_points=8):
    """MSDeformableAttention3D.forward, M/spatial_cross_attention.py:275-402
    (batch_first=True, CPU dispatch :396-398).  `sd[pre + 'value_proj.weight']` ..."""
    bs, num_query, _ = query.shape
    bs, num_value, _ = value.shape
    assert int((spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum()) == num_value
    value = _r(F.linear(value, sd[pre + 'value_proj.weight'], sd[pre + 'value_proj.bias']))
    value = value.view(bs, num_value, num_heads, -1)
    sampling_offsets = F.linear(query, sd[pre + 'sampling_offsets.weight'],
                                sd[pre + 'sampling_offsets.bias']).view(
        bs, num_query, num_heads, num_levels, num_points, 2)
    attention_weights = F.linear(query, sd[pre + 'attention_weights.weight'],
                                 sd[pre + 'attention_weights.bias']).view(
        bs, num_query, num_heads, num_levels * num_points)
    attention_weights = attention_weights.softmax(-1)
    attention_weights = attention_weights.view(bs, num_query, num_heads, num_levels, num_points)
    assert reference_points.shape[-1] == 2
    offset_normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
    bs, num_query, num_Z_anchors, xy = reference_points.shape
    reference_points = reference_points[:, :, None, None, None, :, :]
    sampling_offsets = sampling_offsets / offset_normalizer[None, None, None, :, None, :]
    bs, num_query, num_heads, num_levels, num_all_points, xy = sampling_offsets.shape
    sampling_offsets = sampling_offsets.view(
        bs, num_query, num_heads, num_levels, num_all_points // num_Z_anchors, num_Z_anchors, xy)
    sampling_locations = reference_points + sampling_offsets
    bs, num_query, num_heads, num_levels, num_points, num_Z_anchors, xy = sampling_locations.shape
    assert num_all_points == num_points * num_Z_anchors
    sampling_locations = sampling_locations.view(
        bs, num_query, num_heads, num_levels, num_all_points, xy)
    return multi_scale_deformable_attn_pytorch(
        value, spatial_shapes, sampling_locations, attention_weights)


# --------------------------------------------------------------------------- A3
def sca_forward_single(sd, pre, query, value, reference_points_cam, bev_mask, spatial_shapes,
                       **kw):
    """SpatialCrossAttention.forward for ONE panorama (bs=1), eval mode (dropout off),
    M/spatial_cross_attention.py:76-176.  query (1, Nq, C); value (Ncam, S, 1, C);
    reference_points_cam (Ncam, 1, Nq, D, 2); bev_mask (Ncam, 1, Nq, D)."""
    inp_residual = query
    slots = torch.zeros_like(query)
    bs, num_query, C = query.size()
    assert bs == 1
    num_cams = value.shape[0]
    D = reference_points_cam.size(3)
    indexes = visible_indexes(bev_mask)
    max_len = max([len(each) for each in indexes])
    queries_rebatch = query.new_zeros([bs, num_cams, max_len, C])
    reference_points_rebatch = reference_points_cam.new_zeros([bs, num_cams, max_len, D, 2])
    for j in range(bs):
        for i, reference_points_per_img in enumerate(reference_points_cam):
            idx = indexes[i]
            queries_rebatch[j, i, :len(idx)] = query[j, idx]
            reference_points_rebatch[j, i, :len(idx)] = reference_points_per_img[j, idx]
    l = value.shape[1]
    value = value.permute(2, 0, 1, 3).reshape(bs * num_cams, l, C)
    queries = msda3d_forward(
        sd, pre + 'deformable_attention.',
        queries_rebatch.view(bs * num_cams, max_len, C), value,
        reference_points_rebatch.view(bs * num_cams, max_len, D, 2), spatial_shapes,
        **kw).view(bs, num_cams, max_len, C)
    for j in range(bs):
        for i, idx in enumerate(indexes):
            slots[j, idx] += queries[j, i, :len(idx)]
    count = bev_mask.sum(-1) > 0
    count = count.permute(1, 2, 0).sum(-1)
    count = torch.clamp(count, min=1.0)
    slots = _r(slots / count[..., None])
    slots = _r(F.linear(slots, sd[pre + 'output_proj.weight'], sd[pre + 'output_proj.bias']))
    return slots + inp_residual


def sca_forward(sd, pre, query, value, reference_points_cam, bev_mask, spatial_shapes, **kw):
    """Batched SCA = per-panorama bs=1 reference, concatenated (R3).
    query (B, Nq, C); value (Ncam, S, B, C); rpc (Ncam, B, Nq, D, 2); mask (Ncam, B, Nq, D)."""
    outs = []
    for b in range(query.shape[0]):
        outs.append(sca_forward_single(
            sd, pre, query[b:b + 1], value[:, :, b:b + 1], reference_points_cam[:, b:b + 1],
            bev_mask[:, b:b + 1], spatial_shapes, **kw))
    return torch.cat(outs, 0)


# --------------------------------------------------------------------------- A6 / A7
def ffn_forward(sd, pre, x):
    """mmcv 1.4.0 FFN (Linear-ReLU-[Dropout]-Linear-[Dropout] + identity), eval mode;
    built at M/custom_base_transformer_layer.py:157-158 with vocc.py:134-135."""
    h = _r(F.relu(F.linear(x, sd[pre + 'layers.0.0.weight'], sd[pre + 'layers.0.0.bias'])))
    return x + _r(F.linear(h, sd[pre + 'layers.1.weight'], sd[pre + 'layers.1.bias']))


def layer_forward(sd, pre, query, value, reference_points_cam, bev_mask, spatial_shapes,
                  operation_order=('cross_attn', 'norm', 'ffn', 'norm'), **kw):
    """VoxelFormerLayer.forward, M/voxel_encoder.py:344-464, pre_norm=False."""
    norm_index = attn_index = ffn_index = 0
    C = query.shape[-1]
    for op in operation_order:
        if op == 'cross_attn':
            query = sca_forward(sd, f'{pre}attentions.{attn_index}.', query, value,
                                reference_points_cam, bev_mask, spatial_shapes, **kw)
            attn_index += 1
        elif op == 'norm':
            query = _r(F.layer_norm(query, (C,), sd[f'{pre}norms.{norm_index}.weight'],
                                    sd[f'{pre}norms.{norm_index}.bias'], 1e-5))
            norm_index += 1
        elif op == 'ffn':
            query = ffn_forward(sd, f'{pre}ffns.{ffn_index}.', query)
            ffn_index += 1
        else:
            raise NotImplementedError(op)
    return query


def encoder_forward(sd, pre, bev_query, value, bev_z, bev_h, bev_w, pc_range, lidar2img,
                    originshift, spatial_shapes, num_layers=3, **kw):
    """VoxelFormerEncoder.forward, M/voxel_encoder.py:197-296 (prev_bev=None).
    bev_query (Nq, B, C); value (Ncam, S, B, C) -> (B, Nq, C)."""
    rpc, mask = point_sampling_batched(bev_z, bev_h, bev_w, pc_range, lidar2img, originshift)
    q = bev_query.permute(1, 0, 2)
    for lid in range(num_layers):
        q = layer_forward(sd, f'{pre}layers.{lid}.', q, value, rpc, mask, spatial_shapes, **kw)
    return q


# --------------------------------------------------------------------------- A8
def get_voxel_features(sd, pre, mlvl_feats, bev_queries, bev_z, bev_h, bev_w, pc_range,
                       lidar2img, originshift, num_layers=3, use_cams_embeds=True, **kw):
    """VoxelPerceptionTransformer.get_voxel_features, M/voxel_transformer.py:119-185,
    generalised from the literal `reshape(6, 1, ...)` (:146) to (Ncam, B).
    mlvl_feats (Ncam, B, 196, C); bev_queries (Nq, C) -> (B, Nq, C)."""
    num_cam, bs, S, C = mlvl_feats.shape
    h = w = int(round(S ** 0.5))
    bev_queries = _r(bev_queries).unsqueeze(1).repeat(1, bs, 1)
    feat = mlvl_feats.reshape(num_cam, bs, h, w, C).permute(1, 0, 4, 2, 3)
    feat = feat.flatten(3).permute(1, 0, 3, 2)
    if use_cams_embeds:
        feat = feat + sd[pre + 'cams_embeds'][:, None, None, :].to(feat.dtype)
    feat = _r(feat + sd[pre + 'level_embeds'][None, None, 0:1, :].to(feat.dtype))
    spatial_shapes = torch.as_tensor([(h, w)], dtype=torch.long)
    feat_flatten = feat.permute(0, 2, 1, 3)                       # (Ncam, S, B, C)
    return encoder_forward(sd, pre + 'encoder.', bev_queries, feat_flatten, bev_z, bev_h, bev_w,
                           pc_range, lidar2img, originshift, spatial_shapes,
                           num_layers=num_layers, **kw)


# --------------------------------------------------------------------------- A9
def positional_encoding(sd, pre, bs, d, h, w):
    """VoxelLearnedPositionalEncoding.forward, M/voxel_positional_embedding.py:43-71."""
    x_embed = sd[pre + 'col_embed.weight'][:w]
    y_embed = sd[pre + 'row_embed.weight'][:h]
    z_embed = sd[pre + 'z_embed.weight'][:d]
    xyz = (x_embed[None, None].repeat(d, h, 1, 1) + y_embed[None, :, None, :].repeat(d, 1, w, 1)
           + z_embed[:, None, None, :].repeat(1, h, w, 1))
    return xyz.permute(3, 0, 1, 2).unsqueeze(0).repeat(bs, 1, 1, 1, 1)


# --------------------------------------------------------------------------- A10
def _conv_transpose_stack(sd, pre, x):
    for i in range(3):
        x = F.conv_transpose3d(x, sd[f'{pre}up_sample.{i}.weight'], sd[f'{pre}up_sample.{i}.bias'],
                               stride=(1, 2, 2), padding=(2, 4, 4), output_padding=(0, 1, 1),
                               dilation=(2, 2, 2))
    return x


def occ_head_single(sd, pre, bev_embed, bev_z, bev_h, bev_w, occ_xdim, occ_ydim, occ_zdim,
                    occ_dims=128, refine_occ=False, only_occ=False, num_occ_fcs=2):
    """Occupancy part of VoxelFormerOccupancyHead.forward for ONE panorama.
    only_occ branch HEAD:323-352 takes bev_embed (1, Nq, C); default branch HEAD:551-580
    takes the (Nq, 1, C) tensor and applies the raw `.view` reinterpretations (A4.3)."""
    C = bev_embed.shape[-1]
    bs = 1
    if refine_occ and not only_occ:
        x = bev_embed.contiguous().view(bs, C, bev_z, bev_h, bev_w)          # HEAD:558
        x = _conv_transpose_stack(sd, pre, x)                                # HEAD:560
        x = x.contiguous().view(bs, bev_z, occ_xdim, occ_ydim, C)            # HEAD:564
    else:
        x = bev_embed.contiguous().view(bs, bev_z, bev_h, bev_w, C)          # HEAD:334 / :566
    if bev_z == occ_zdim:
        occ_pred = _r(F.linear(x, sd[pre + 'occ_proj.weight'], sd[pre + 'occ_proj.bias']))
    else:
        x = x.permute(0, 2, 3, 1, 4).flatten(3)
        occ_pred = _r(F.linear(x, sd[pre + 'occ_proj.weight'], sd[pre + 'occ_proj.bias']))
        if refine_occ and not only_occ:
            occ_pred = occ_pred.view(bs, occ_xdim, occ_ydim, occ_zdim, occ_dims)
        else:
            occ_pred = occ_pred.view(bs, bev_h, bev_w, occ_zdim, occ_dims)
        occ_pred = occ_pred.permute(0, 3, 1, 2, 4)
    occ_pred = occ_pred.reshape(bs, occ_zdim, -1, occ_dims)
    occ_pred = occ_pred.reshape(bs, -1, occ_dims)
    y = occ_pred
    for i in range(num_occ_fcs):                                             # HEAD:242-248
        y = _r(F.linear(y, sd[f'{pre}occ_branches.{3 * i}.weight'], sd[f'{pre}occ_branches.{3 * i}.bias']))
        y = _r(F.layer_norm(y, (occ_dims,), sd[f'{pre}occ_branches.{3 * i + 1}.weight'],
                            sd[f'{pre}occ_branches.{3 * i + 1}.bias'], 1e-5))
        y = F.relu(y)
    k = 3 * num_occ_fcs
    return _r(F.linear(y, sd[f'{pre}occ_branches.{k}.weight'], sd[f'{pre}occ_branches.{k}.bias']))


def occ_head(sd, pre, bev_embed_bnc, *args, only_occ=False, **kw):
    """Batched head: bev_embed (B, Nq, C) from the encoder; per-sample bs=1 reference.
    Default branch receives voxel_embed.permute(1,0,2) = (Nq, 1, C) (M/voxel_transformer.py:262)."""
    outs = []
    for b in range(bev_embed_bnc.shape[0]):
        e = bev_embed_bnc[b:b + 1]
        if not only_occ:
            e = e.permute(1, 0, 2)
        outs.append(occ_head_single(sd, pre, e, *args, only_occ=only_occ, **kw))
    return torch.cat(outs, 0)


# --------------------------------------------------------------------------- A11
def dense_occupancy_target(occ_gt, voxel_num, occupancy_classes=16):
    """HEAD:1326-1330 / :1405-1409: fill with class `occupancy_classes` ("empty"),
    scatter the sparse (index, class) ground truth."""
    gt = torch.full((voxel_num,), occupancy_classes, dtype=torch.long)
    gt[occ_gt[:, 0].long()] = occ_gt[:, 1].long()
    return gt


def sigmoid_focal_loss(pred, target, gamma=2.0, alpha=0.25, avg_factor=None, loss_weight=1.0):
    """mmdet 2.14 FocalLoss(use_sigmoid=True) CPU path (py_sigmoid_focal_loss +
    weight_reduce_loss 'mean' with avg_factor); cfg vocc.py:190-195, call HEAD:981 / :1425.
    pred (N, Ccls) logits; target (N,) int64 in [0, Ccls], Ccls = background."""
    num_classes = pred.size(1)
    t = F.one_hot(target, num_classes=num_classes + 1)[:, :num_classes].type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * t + p * (1 - t)
    focal_weight = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(pred, t, reduction='none') * focal_weight
    loss = loss.sum() / avg_factor if avg_factor is not None else loss.mean()
    return loss_weight * loss


def occupancy_loss(occupancy_preds, occ_gt_list, occupancy_classes=16, **kw):
    """Per-panorama HEAD:1386-1444 (loss_only_occupancy), mean over the batch is NOT taken
    by the reference (bs=1); batched oracle = mean of per-panorama losses."""
    losses = []
    for b, occ_gt in enumerate(occ_gt_list):
        preds = occupancy_preds[b].reshape(-1, occupancy_classes)
        gt = dense_occupancy_target(occ_gt, preds.shape[0], occupancy_classes)
        avg = (gt < occupancy_classes).sum() * 1.0
        losses.append(torch.nan_to_num(sigmoid_focal_loss(preds, gt, avg_factor=avg, **kw)))
    return torch.stack(losses).mean()


# --------------------------------------------------------------------------- A12
def get_occupancy_prediction(occupancy_preds, occupancy_classes=16, occ_threshold=0.25):
    """HEAD:1505-1524 (focal_loss branch): sigmoid, append threshold column, argmax,
    keep rows with argmax < classes -> (n_occ, 2) int64 (flat index, class)."""
    p = occupancy_preds.reshape(-1, occupancy_classes).sigmoid()
    p = torch.cat((p, torch.ones_like(p)[:, :1] * occ_threshold), dim=-1)
    occ_class = p.argmax(dim=-1)
    occ_index, = torch.where(occ_class < occupancy_classes)
    return torch.stack([occ_index, occ_class[occ_index]], dim=-1)


# =========================================================================== N2 / N3
def voxel_multi_scale_deformable_attn_pytorch(value, value_spatial_shapes, sampling_locations,
                                              attention_weights):
    """The reference-owned 3-D sampler, M/voxel_temporal_self_attention.py:275-335, restated
    (pinned to the unmodified function in tests/test_oracle_pins_reference.py).

    value (bs, num_keys, num_heads, Dh); value_spatial_shapes (num_levels, 3) = (d, h, w);
    sampling_locations (bs, nq, num_heads, num_levels, num_points, 3) = (x, y, z) in [0, 1];
    attention_weights (bs, nq, num_heads, num_levels, num_points) -> (bs, nq, num_heads*Dh).
    The reference's bare `.squeeze()` (:319) is `.squeeze(2)` here: identical unless some other
    dimension has size 1, where the reference mis-shapes."""
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, _, num_levels, num_points, _ = sampling_locations.shape
    shapes = [(int(d), int(h), int(w)) for d, h, w in value_spatial_shapes]
    value_list = value.split([d * h * w for d, h, w in shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampled = []
    for level, (d, h, w) in enumerate(shapes):
        value_l = value_list[level].flatten(2).transpose(1, 2).reshape(bs * num_heads, embed_dims, d, h, w)
        grid_l = sampling_grids[:, :, :, level].transpose(1, 2).flatten(0, 1).unsqueeze(1)
        sampled.append(F.grid_sample(value_l, grid_l, mode='bilinear', padding_mode='zeros',
                                     align_corners=False).squeeze(2))      # (bs*nh, Dh, nq, np)
    attention_weights = attention_weights.transpose(1, 2).reshape(
        bs * num_heads, 1, num_queries, num_levels * num_points)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * attention_weights).sum(-1)
    return out.view(bs, num_heads * embed_dims, num_queries).transpose(1, 2).contiguous()


def inverse_sigmoid(x, eps=1e-5):
    """M/voxel_decoder.py:35-50."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def voxel_custom_msda_forward(sd, pre, query, value, reference_points, spatial_shapes, identity=None,
                              query_pos=None, num_heads=8, num_levels=1, num_points=4):
    """VoxelCustomMSDeformableAttention.forward, M/voxel_decoder.py:236-322, batch_first=False, eval.
    query (nq, bs, C); value (num_value, bs, C); reference_points (bs, nq, num_levels, 3);
    spatial_shapes (num_levels, 3) = (d, h, w) -> (nq, bs, C)."""
    if identity is None:
        identity = query
    if query_pos is not None:
        query = query + query_pos
    query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
    bs, num_query, _ = query.shape
    _, num_value, _ = value.shape
    assert int((spatial_shapes[:, 0] * spatial_shapes[:, 1] * spatial_shapes[:, 2]).sum()) == num_value
    value = _r(F.linear(value, sd[pre + 'value_proj.weight'], sd[pre + 'value_proj.bias']))
    value = value.view(bs, num_value, num_heads, -1)
    offsets = F.linear(query, sd[pre + 'sampling_offsets.weight'], sd[pre + 'sampling_offsets.bias']).view(
        bs, num_query, num_heads, num_levels, num_points, 3)
    weights = F.linear(query, sd[pre + 'attention_weights.weight'], sd[pre + 'attention_weights.bias']).view(
        bs, num_query, num_heads, num_levels * num_points).softmax(-1).view(
        bs, num_query, num_heads, num_levels, num_points)
    if reference_points.shape[-1] != 3:
        raise ValueError(f'Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.')
    normalizer = torch.stack([spatial_shapes[..., 2], spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
    locations = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
    out = voxel_multi_scale_deformable_attn_pytorch(value, spatial_shapes, locations, weights)
    out = F.linear(out, sd[pre + 'output_proj.weight'], sd[pre + 'output_proj.bias'])
    return out.permute(1, 0, 2) + identity


def multihead_self_attention_forward(sd, pre, query, query_pos, num_heads=8):
    """mmcv 1.4.0 MultiheadAttention as the 'self_attn' op of DetrTransformerDecoderLayer
    (vocc.py:141-145; called with key = value = query, key_pos = query_pos at
    M/custom_base_transformer_layer.py:223-234), eval mode.  query (nq, bs, C)."""
    nq, bs, C = query.shape
    qk = query if query_pos is None else query + query_pos
    Wq, Wk, Wv = sd[pre + 'attn.in_proj_weight'].chunk(3, 0)
    bq, bk, bv = sd[pre + 'attn.in_proj_bias'].chunk(3, 0)
    dh = C // num_heads

    def heads(x, W, b):                       # (nq, bs, C) -> (bs, nh, nq, dh)
        return F.linear(x, W, b).view(nq, bs, num_heads, dh).permute(1, 2, 0, 3)
    q, k, v = heads(qk, Wq, bq), heads(qk, Wk, bk), heads(query, Wv, bv)
    att = (q @ k.transpose(-1, -2) / dh ** 0.5).softmax(-1)
    out = (att @ v).permute(2, 0, 1, 3).reshape(nq, bs, C)
    out = F.linear(out, sd[pre + 'attn.out_proj.weight'], sd[pre + 'attn.out_proj.bias'])
    return query + out


def decoder_layer_forward(sd, pre, query, value, query_pos, reference_points, spatial_shapes,
                          operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'), **kw):
    """DetrTransformerDecoderLayer.forward = mmcv BaseTransformerLayer.forward (the reference's
    copy: M/custom_base_transformer_layer.py:165-260), pre_norm=False, batch_first=False."""
    C = query.shape[-1]
    norm_index = attn_index = ffn_index = 0
    for op in operation_order:
        if op == 'self_attn':
            query = multihead_self_attention_forward(sd, f'{pre}attentions.{attn_index}.', query, query_pos,
                                                     num_heads=kw.get('num_heads', 8))
            attn_index += 1
        elif op == 'cross_attn':
            query = voxel_custom_msda_forward(sd, f'{pre}attentions.{attn_index}.', query, value,
                                              reference_points, spatial_shapes, query_pos=query_pos, **kw)
            attn_index += 1
        elif op == 'norm':
            query = _r(F.layer_norm(query, (C,), sd[f'{pre}norms.{norm_index}.weight'],
                                    sd[f'{pre}norms.{norm_index}.bias'], 1e-5))
            norm_index += 1
        elif op == 'ffn':
            query = ffn_forward(sd, f'{pre}ffns.{ffn_index}.', query)
            ffn_index += 1
    return query


def decoder_forward(sd, pre, query, value, query_pos, reference_points, spatial_shapes, num_layers=6,
                    reg_branches=None, return_intermediate=True, **kw):
    """VoxelDetectionTransformerDecoder.forward, M/voxel_decoder.py:68-132.
    query / query_pos (nq, bs, C); value (num_value, bs, C); reference_points (bs, nq, 3) in (0,1);
    reg_branches: None or a list of callables (bs, nq, C) -> (bs, nq, >=5)."""
    output = query
    inter, inter_ref = [], []
    for lid in range(num_layers):
        output = decoder_layer_forward(sd, f'{pre}layers.{lid}.', output, value, query_pos,
                                       reference_points[..., :3].unsqueeze(2), spatial_shapes, **kw)
        output = output.permute(1, 0, 2)
        if reg_branches is not None:
            tmp = reg_branches[lid](output)
            assert reference_points.shape[-1] == 3
            new_ref = torch.zeros_like(reference_points)
            new_ref[..., :2] = tmp[..., :2] + inverse_sigmoid(reference_points[..., :2])
            new_ref[..., 2:3] = tmp[..., 4:5] + inverse_sigmoid(reference_points[..., 2:3])
            reference_points = new_ref.sigmoid().detach()
        output = output.permute(1, 0, 2)
        if return_intermediate:
            inter.append(output)
            inter_ref.append(reference_points)
    if return_intermediate:
        return torch.stack(inter), torch.stack(inter_ref)
    return output, reference_points


def transformer_decode(sd, pre, voxel_embed, object_query_embed, bev_z, bev_h, bev_w, num_layers=6,
                       reg_branches=None, **kw):
    """The decoder half of VoxelPerceptionTransformer.forward, M/voxel_transformer.py:246-301
    (decoder_on_bev=False).  voxel_embed (bs, Nq, C) from get_voxel_features;
    object_query_embed (num_query, 2C) -> (voxel_embed (Nq, bs, C), inter_states, init_reference,
    inter_references)."""
    bs, _, C = voxel_embed.shape
    query_pos, query = torch.split(object_query_embed, C, dim=1)
    query_pos = query_pos.unsqueeze(0).expand(bs, -1, -1)
    query = query.unsqueeze(0).expand(bs, -1, -1)
    reference_points = F.linear(query_pos, sd[pre + 'reference_points.weight'],
                                sd[pre + 'reference_points.bias']).sigmoid()
    init_reference_out = reference_points
    query, query_pos = query.permute(1, 0, 2), query_pos.permute(1, 0, 2)
    voxel_embed = voxel_embed.permute(1, 0, 2)
    inter_states, inter_references = decoder_forward(
        sd, pre + 'decoder.', query, voxel_embed, query_pos, reference_points,
        torch.tensor([[bev_z, bev_h, bev_w]]), num_layers=num_layers, reg_branches=reg_branches, **kw)
    return voxel_embed, inter_states, init_reference_out, inter_references


def temporal_self_attention_forward(sd, pre, query, reference_points, spatial_shapes, query_pos=None,
                                    num_heads=8, num_levels=1, num_points=4, num_bev_queue=2):
    """VoxelTemporalSelfAttention.forward, M/voxel_temporal_self_attention.py:129-273, with
    value=None (no history: the current volume is stacked twice, :180-182), batch_first=True, eval.
    query (bs, Nq, C); reference_points (bs*2, Nq, num_levels, 3) -> (bs, Nq, C).
    NOTE the shipped module cannot run as initialised: init_weights assigns a 2-component bias of
    num_heads*num_levels*num_bev_queue*num_points*2 elements (:113-124) to a Linear with
    ...*3 outputs (:99-100); the arithmetic below is what forward computes once
    sampling_offsets.bias has the Linear's own size."""
    bs, len_bev, C = query.shape
    value = torch.stack([query, query], 1).reshape(bs * 2, len_bev, C)
    identity = query
    if query_pos is not None:
        query = query + query_pos
    bs, num_query, embed_dims = query.shape
    _, num_value, _ = value.shape
    assert int((spatial_shapes[:, 0] * spatial_shapes[:, 1] * spatial_shapes[:, 2]).sum()) == num_value
    assert num_bev_queue == 2
    query = torch.cat([value[:bs], query], -1)
    value = F.linear(value, sd[pre + 'value_proj.weight'], sd[pre + 'value_proj.bias'])
    value = value.reshape(bs * num_bev_queue, num_value, num_heads, -1)
    offsets = F.linear(query, sd[pre + 'sampling_offsets.weight'], sd[pre + 'sampling_offsets.bias']).view(
        bs, num_query, num_heads, num_bev_queue, num_levels, num_points, 3)
    weights = F.linear(query, sd[pre + 'attention_weights.weight'], sd[pre + 'attention_weights.bias']).view(
        bs, num_query, num_heads, num_bev_queue, num_levels * num_points).softmax(-1).view(
        bs, num_query, num_heads, num_bev_queue, num_levels, num_points)
    weights = weights.permute(0, 3, 1, 2, 4, 5).reshape(
        bs * num_bev_queue, num_query, num_heads, num_levels, num_points).contiguous()
    offsets = offsets.permute(0, 3, 1, 2, 4, 5, 6).reshape(
        bs * num_bev_queue, num_query, num_heads, num_levels, num_points, 3)
    if reference_points.shape[-1] != 3:
        raise ValueError(f'Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.')
    normalizer = torch.stack([spatial_shapes[..., 2], spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
    locations = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
    out = voxel_multi_scale_deformable_attn_pytorch(value, spatial_shapes, locations, weights)
    out = out.permute(1, 2, 0).view(num_query, embed_dims, bs, num_bev_queue).mean(-1).permute(2, 0, 1)
    out = F.linear(out, sd[pre + 'output_proj.weight'], sd[pre + 'output_proj.bias'])
    return out + identity


def detection_tail(sd, pre, hs, init_reference, inter_references, pc_range, num_reg_fcs=2):
    """The per-decoder-layer classification / box outputs of VoxelFormerOccupancyHead.forward,
    HEAD:583-611 (default branch; identical code at :385-412 for only_det).  hs (num_dec, nq, bs, C);
    init_reference (bs, nq, 3); inter_references (num_dec, bs, nq, 3).
    cls_branches[l] = [Linear, LayerNorm, ReLU] x num_reg_fcs + Linear; reg_branches[l] = [Linear, ReLU] x
    num_reg_fcs + Linear (HEAD:181-197)."""
    hs = hs.permute(0, 2, 1, 3)
    C = hs.shape[-1]
    outputs_classes, outputs_coords = [], []
    for lvl in range(hs.shape[0]):
        reference = inverse_sigmoid(init_reference if lvl == 0 else inter_references[lvl - 1])
        x = hs[lvl]
        for i in range(num_reg_fcs):
            x = F.linear(x, sd[f'{pre}cls_branches.{lvl}.{3 * i}.weight'], sd[f'{pre}cls_branches.{lvl}.{3 * i}.bias'])
            x = F.relu(F.layer_norm(x, (C,), sd[f'{pre}cls_branches.{lvl}.{3 * i + 1}.weight'],
                                    sd[f'{pre}cls_branches.{lvl}.{3 * i + 1}.bias'], 1e-5))
        k = 3 * num_reg_fcs
        outputs_classes.append(F.linear(x, sd[f'{pre}cls_branches.{lvl}.{k}.weight'], sd[f'{pre}cls_branches.{lvl}.{k}.bias']))
        tmp = hs[lvl]
        for i in range(num_reg_fcs):
            tmp = F.relu(F.linear(tmp, sd[f'{pre}reg_branches.{lvl}.{2 * i}.weight'], sd[f'{pre}reg_branches.{lvl}.{2 * i}.bias']))
        k = 2 * num_reg_fcs
        tmp = F.linear(tmp, sd[f'{pre}reg_branches.{lvl}.{k}.weight'], sd[f'{pre}reg_branches.{lvl}.{k}.bias']).clone()
        assert reference.shape[-1] == 3
        tmp[..., 0:2] += reference[..., 0:2]
        tmp[..., 0:2] = tmp[..., 0:2].sigmoid()
        tmp[..., 4:5] += reference[..., 2:3]
        tmp[..., 4:5] = tmp[..., 4:5].sigmoid()
        tmp[..., 0:1] = tmp[..., 0:1] * (pc_range[3] - pc_range[0]) + pc_range[0]
        tmp[..., 1:2] = tmp[..., 1:2] * (pc_range[4] - pc_range[1]) + pc_range[1]
        tmp[..., 4:5] = tmp[..., 4:5] * (pc_range[5] - pc_range[2]) + pc_range[2]
        outputs_coords.append(tmp)
    return torch.stack(outputs_classes), torch.stack(outputs_coords)
