"""Name-compatible home of the autograd wrappers of
projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py
(the `_fp16` class :15-87 is never selected by the reference, spatial_cross_attention.py:388-391;
both names resolve to the same sm_100a-backed Function here)."""
from ..ops import (MultiScaleDeformableAttnFunction, MultiScaleDeformableAttnFunction_fp16,  # noqa: F401
                   MultiScaleDeformableAttnFunction_fp32, ms_deform_attn_backward,
                   ms_deform_attn_forward)
