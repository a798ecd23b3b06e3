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

Batched semantics: the reference only runs bs=1.  For B>1 the oracle
is "run the bs=1 reference on each panorama with its own camera matrices and
concatenate".
"""
import torch
import torch.nn.functional as F

IMG_W, IMG_H = 1280, 1024          # hard-coded at M/voxel_encoder.py:179-180


# --------------------------------------------------------------------------- A5 / K3
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
    value = F.linear(value, sd[pre + 'value_proj.weight'], sd[pre + 'value_proj.bias'])
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
    slots = slots / count[..., None]
    slots = F.linear(slots, sd[pre + 'output_proj.weight'], sd[pre + 'output_proj.bias'])
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
    h = F.relu(F.linear(x, sd[pre + 'layers.0.0.weight'], sd[pre + 'layers.0.0.bias']))
    return x + F.linear(h, sd[pre + 'layers.1.weight'], sd[pre + 'layers.1.bias'])


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
            query = F.layer_norm(query, (C,), sd[f'{pre}norms.{norm_index}.weight'],
                                 sd[f'{pre}norms.{norm_index}.bias'], 1e-5)
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
    bev_queries = bev_queries.unsqueeze(1).repeat(1, bs, 1)
    feat = mlvl_feats.reshape(num_cam, bs, h, w, C).permute(1, 0, 4, 2, 3)
    feat = feat.flatten(3).permute(1, 0, 3, 2)
    if use_cams_embeds:
        feat = feat + sd[pre + 'cams_embeds'][:, None, None, :].to(feat.dtype)
    feat = feat + sd[pre + 'level_embeds'][None, None, 0:1, :].to(feat.dtype)
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
        occ_pred = F.linear(x, sd[pre + 'occ_proj.weight'], sd[pre + 'occ_proj.bias'])
    else:
        x = x.permute(0, 2, 3, 1, 4).flatten(3)
        occ_pred = F.linear(x, sd[pre + 'occ_proj.weight'], sd[pre + 'occ_proj.bias'])
        if refine_occ and not only_occ:
            occ_pred = occ_pred.view(bs, occ_xdim, occ_ydim, occ_zdim, occ_dims)
        else:
            occ_pred = occ_pred.view(bs, bev_h, bev_w, occ_zdim, occ_dims)
        occ_pred = occ_pred.permute(0, 3, 1, 2, 4)
    occ_pred = occ_pred.reshape(bs, occ_zdim, -1, occ_dims)
    occ_pred = occ_pred.reshape(bs, -1, occ_dims)
    y = occ_pred
    for i in range(num_occ_fcs):                                             # HEAD:242-248
        y = F.linear(y, sd[f'{pre}occ_branches.{3 * i}.weight'], sd[f'{pre}occ_branches.{3 * i}.bias'])
        y = F.layer_norm(y, (occ_dims,), sd[f'{pre}occ_branches.{3 * i + 1}.weight'],
                         sd[f'{pre}occ_branches.{3 * i + 1}.bias'], 1e-5)
        y = F.relu(y)
    k = 3 * num_occ_fcs
    return F.linear(y, sd[f'{pre}occ_branches.{k}.weight'], sd[f'{pre}occ_branches.{k}.bias'])


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
