"""TEST / BASELINE INFRASTRUCTURE (not product code): a state_dict of the lift+encode path initialised by the
reference's own rules, built with plain torch so that bench.py's CPU reference arm never imports the product.

Keys = SURVEY.md A3 (prefix `pts_bbox_head.` stripped).  Initialisation rules restated from the reference:
  * VoxelPerceptionTransformer.init_weights (M/voxel_transformer.py:99-116): xavier_uniform_ on every >1-D
    parameter of the transformer, then MSDeformableAttention3D.init_weights, then N(0, 1) level / camera embeddings;
  * MSDeformableAttention3D.init_weights (M/spatial_cross_attention.py:255-273): sampling_offsets weight 0, bias =
    head direction (cos, sin)(2 pi h / NH) / max|.| scaled by (p + 1); attention_weights 0; value_proj xavier, bias 0;
  * SpatialCrossAttention.init_weight (:71-73): output_proj xavier, bias 0;
  * nn.Embedding default N(0, 1) for voxel_embedding (HEAD:226-227); nn.Linear / nn.LayerNorm defaults for the head's
    occ_proj / occ_branches (HEAD:236-248; the head's init_weights does not touch them);
  * + the bench perturbation of SURVEY 8(d): N(0, 0.02) on sampling_offsets.weight / attention_weights.weight so that
    offsets and weights depend on the query.
Numerically this is the same DISTRIBUTION as the product's `build_head(...).init_weights()`, not the same draws:
the CPU arm is a timing baseline (its run time does not depend on the weight values)."""
import math

import torch
from torch import nn


def offsets_bias(num_heads=8, num_levels=1, num_points=8):
    """M/spatial_cross_attention.py:258-269."""
    thetas = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2).repeat(1, num_levels, num_points, 1)
    for i in range(num_points):
        grid[:, :, i, :] *= i + 1
    return grid.view(-1)


def lift_encode_state_dict(grid, num_cams=18, embed=768, ffn=1536, num_layers=3, num_heads=8, num_points=8,
                           occ_dims=128, classes=16, seed=0, perturb=0.02):
    g = torch.Generator().manual_seed(seed)
    z, h, w = grid

    def xavier(*shape):
        t = torch.empty(*shape)
        fan_out, fan_in = shape[0], shape[1]
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return t.uniform_(-a, a, generator=g)

    def linear_default(out_f, in_f):          # nn.Linear.reset_parameters: kaiming_uniform(a = sqrt 5), bias U(+-1/sqrt fan_in)
        bound = 1.0 / math.sqrt(in_f)
        return (torch.empty(out_f, in_f).uniform_(-bound, bound, generator=g),
                torch.empty(out_f).uniform_(-bound, bound, generator=g))

    sd = {'voxel_embedding.weight': torch.randn(z * h * w, embed, generator=g),
          'transformer.level_embeds': torch.randn(4, embed, generator=g),
          'transformer.cams_embeds': torch.randn(num_cams, embed, generator=g)}
    for l in range(num_layers):
        p = f'transformer.encoder.layers.{l}.'
        a = p + 'attentions.0.'
        d = a + 'deformable_attention.'
        sd[d + 'sampling_offsets.weight'] = torch.randn(num_heads * num_points * 2, embed, generator=g) * perturb
        sd[d + 'sampling_offsets.bias'] = offsets_bias(num_heads, 1, num_points)
        sd[d + 'attention_weights.weight'] = torch.randn(num_heads * num_points, embed, generator=g) * perturb
        sd[d + 'attention_weights.bias'] = torch.zeros(num_heads * num_points)
        sd[d + 'value_proj.weight'] = xavier(embed, embed)
        sd[d + 'value_proj.bias'] = torch.zeros(embed)
        sd[a + 'output_proj.weight'] = xavier(embed, embed)
        sd[a + 'output_proj.bias'] = torch.zeros(embed)
        sd[p + 'ffns.0.layers.0.0.weight'] = xavier(ffn, embed)
        sd[p + 'ffns.0.layers.0.0.bias'] = linear_default(ffn, embed)[1]
        sd[p + 'ffns.0.layers.1.weight'] = xavier(embed, ffn)
        sd[p + 'ffns.0.layers.1.bias'] = linear_default(embed, ffn)[1]
        for n in (0, 1):
            sd[p + f'norms.{n}.weight'] = torch.ones(embed)
            sd[p + f'norms.{n}.bias'] = torch.zeros(embed)
    sd['occ_proj.weight'], sd['occ_proj.bias'] = linear_default(occ_dims, embed)
    for i, (o, n_in) in ((0, (occ_dims, occ_dims)), (3, (occ_dims, occ_dims)), (6, (classes, occ_dims))):
        sd[f'occ_branches.{i}.weight'], sd[f'occ_branches.{i}.bias'] = linear_default(o, n_in)
    for i in (1, 4):
        sd[f'occ_branches.{i}.weight'] = torch.ones(occ_dims)
        sd[f'occ_branches.{i}.bias'] = torch.zeros(occ_dims)
    return sd


def check_against(module_state_dict, sd):
    """Same key set (on the path) and shapes as a module's state_dict -- used by tests/."""
    for k, v in sd.items():
        assert k in module_state_dict, k
        assert tuple(module_state_dict[k].shape) == tuple(v.shape), (k, module_state_dict[k].shape, v.shape)
