"""VoxelLearnedPositionalEncoding -- mirror of
projects/mmdet3d_plugin/bevformer/modules/voxel_positional_embedding.py:10-79.
Computed for API parity; the cross-attention-only encoder of vocc.py never reads it
(SURVEY.md A9), so the head skips it on the fast path."""
import torch
import torch.nn as nn

from ..registry import POSITIONAL_ENCODING, BaseModule


@POSITIONAL_ENCODING.register_module()
class VoxelLearnedPositionalEncoding(BaseModule):
    def __init__(self, num_feats, row_num_embed=50, col_num_embed=50, z_num_embed=16,
                 init_cfg=dict(type='Uniform', layer='Embedding')):
        super().__init__(init_cfg)
        self.num_feats = num_feats
        self.row_embed = nn.Embedding(row_num_embed, num_feats * 2)
        self.col_embed = nn.Embedding(col_num_embed, num_feats * 2)
        self.z_embed = nn.Embedding(z_num_embed, num_feats * 2)
        self.row_num_embed, self.col_num_embed, self.z_num_embed = row_num_embed, col_num_embed, z_num_embed

    def init_weights(self):
        # mmcv init_cfg dict(type='Uniform', layer='Embedding'): U(0, 1) on every Embedding
        for m in (self.row_embed, self.col_embed, self.z_embed):
            nn.init.uniform_(m.weight, 0, 1)
        self._is_init = True

    def forward(self, mask):
        """mask (bs, d, h, w) -> (bs, 2*num_feats, d, h, w) = col[x] + row[y] + z[z]."""
        d, h, w = mask.shape[-3:]
        dev = mask.device
        x = self.col_embed(torch.arange(w, device=dev))
        y = self.row_embed(torch.arange(h, device=dev))
        z = self.z_embed(torch.arange(d, device=dev))
        pos = (x[None, None, :, :] + y[None, :, None, :]) + z[:, None, None, :]
        return pos.permute(3, 0, 1, 2).unsqueeze(0).repeat(mask.shape[0], 1, 1, 1, 1)

    def __repr__(self):
        return (f'{self.__class__.__name__}(num_feats={self.num_feats}, '
                f'row_num_embed={self.row_num_embed}, col_num_embed={self.col_num_embed})')
