"""Importing this package registers the lift-path modules (the reference does the same as a
side effect of importing its plugin package, tools/train.py:114-126)."""
from .custom_base_transformer_layer import FFN, MyCustomBaseTransformerLayer          # noqa: F401
from .multi_scale_deformable_attn_function import (MultiScaleDeformableAttnFunction_fp16,  # noqa: F401
                                                   MultiScaleDeformableAttnFunction_fp32)
from .precision import set_compute_dtype                                              # noqa: F401
from .spatial_cross_attention import MSDeformableAttention3D, SpatialCrossAttention   # noqa: F401
from .voxel_encoder import VoxelFormerEncoder, VoxelFormerLayer                       # noqa: F401
from .voxel_positional_embedding import VoxelLearnedPositionalEncoding                # noqa: F401
from .voxel_decoder import (BaseTransformerLayer, DetrTransformerDecoderLayer, MultiheadAttention,  # noqa: F401
                            VoxelCustomMSDeformableAttention, VoxelDetectionTransformerDecoder)
from .voxel_temporal_self_attention import VoxelTemporalSelfAttention                  # noqa: F401
from .voxel_transformer import VoxelPerceptionTransformer                             # noqa: F401
from .voxelformer_occupancy_head import FocalLoss, VoxelFormerOccupancyHead           # noqa: F401
