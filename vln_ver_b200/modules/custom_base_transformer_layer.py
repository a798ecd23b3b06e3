"""MyCustomBaseTransformerLayer + mmcv-1.4.0-compatible FFN
(projects/mmdet3d_plugin/bevformer/modules/custom_base_transformer_layer.py:72-163)."""
import copy
import warnings

import torch.nn as nn

from ..registry import (FEEDFORWARD_NETWORK, HAVE_MMCV, BaseModule, ConfigDict, ModuleList, Sequential,
                        build_attention, build_feedforward_network, build_norm_layer)
from .precision import PrecisionMixin

_ACTIVATIONS = {'ReLU': nn.ReLU, 'GELU': nn.GELU}


class FFN(PrecisionMixin, BaseModule):
    """[Linear-act-Dropout] x (num_fcs-1), Linear, Dropout, (+ identity).  Parameter names
    `layers.0.0.*`, `layers.1.*` as in mmcv."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2, f'num_fcs should be no less than 2. got {num_fcs}.'
        self.embed_dims, self.feedforward_channels, self.num_fcs = embed_dims, feedforward_channels, num_fcs
        self.act_cfg = act_cfg
        act = dict(act_cfg)
        self.activate = _ACTIVATIONS[act.pop('type')](**act)
        layers, in_channels = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(in_channels, feedforward_channels), self.activate,
                                     nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = nn.Dropout(dropout_layer.get('drop_prob', 0.5)) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        cd = self.compute_dtype or x.dtype
        out = x
        for layer in self.layers:
            if isinstance(layer, nn.Linear):
                out = self._linear(out, layer, cd)
            elif isinstance(layer, nn.Sequential):
                out = layer[2](layer[1](self._linear(out, layer[0], cd)))
            else:
                out = layer(out)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity.to(out.dtype) + self.dropout_layer(out)

    def forward_unfused_tail(self, x):
        """Fast path used by VoxelFormerLayer when the FFN is followed by 'norm' (num_fcs == 2, ReLU):
        returns (pre-dropout output of the last Linear, final dropout p) so that the caller can fuse
        dropout + identity add + LayerNorm into one kernel; the inner ReLU + Dropout is one kernel too."""
        from .. import ops
        cd = self.compute_dtype or x.dtype
        first, last, last_drop = self.layers[0], self.layers[1], self.layers[2]
        a = self._linear(x, first[0], cd)
        h = ops.relu_dropout_(a, first[2].p, self.training)
        return self._linear(h, last, cd), last_drop.p

    def fusable(self):
        return (self.num_fcs == 2 and self.add_identity and isinstance(self.activate, nn.ReLU)
                and isinstance(self.dropout_layer, nn.Identity))


if not HAVE_MMCV:
    FEEDFORWARD_NETWORK.register_module()(FFN)


class MyCustomBaseTransformerLayer(BaseModule):
    """Builds attentions / FFNs / norms from cfg; maps the deprecated
    feedforward_channels / ffn_dropout / ffn_num_fcs kwargs into ffn_cfgs
    (reference :72-163).  Unlike the reference the default ffn_cfgs dict is copied, not
    mutated in place: same built modules, no cross-instance leakage."""

    def __init__(self, attn_cfgs=None,
                 ffn_cfgs=dict(type='FFN', embed_dims=768, feedforward_channels=1024, num_fcs=2,
                               ffn_drop=0., act_cfg=dict(type='ReLU', inplace=True)),
                 operation_order=None, norm_cfg=dict(type='LN'), init_cfg=None, batch_first=True,
                 **kwargs):
        ffn_cfgs = copy.deepcopy(ffn_cfgs)
        deprecated_args = dict(feedforward_channels='feedforward_channels', ffn_dropout='ffn_drop',
                               ffn_num_fcs='num_fcs')
        for ori_name, new_name in deprecated_args.items():
            if ori_name in kwargs:
                warnings.warn(f'The arguments `{ori_name}` in BaseTransformerLayer has been deprecated, '
                              f'now you should set `{new_name}` and other FFN related arguments to a '
                              f'dict named `ffn_cfgs`. ')
                ffn_cfgs[new_name] = kwargs[ori_name]
        super().__init__(init_cfg)
        self.batch_first = batch_first
        allowed = {'self_attn', 'norm', 'ffn', 'cross_attn'}
        assert set(operation_order) & allowed == set(operation_order), \
            f'The operation_order of {self.__class__.__name__} should contains all four operation ' \
            f"type {['self_attn', 'norm', 'ffn', 'cross_attn']}"
        num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        else:
            assert num_attn == len(attn_cfgs), \
                f'The length of attn_cfg {num_attn} is not consistent with the number of attention' \
                f'{len(attn_cfgs)} in operation_order {operation_order}.'
            attn_cfgs = [copy.deepcopy(c) for c in attn_cfgs]
        self.num_attn = num_attn
        self.operation_order = operation_order
        self.norm_cfg = norm_cfg
        self.pre_norm = operation_order[0] == 'norm'
        self.attentions = ModuleList()
        index = 0
        for operation_name in operation_order:
            if operation_name in ['self_attn', 'cross_attn']:
                if 'batch_first' in attn_cfgs[index]:
                    assert self.batch_first == attn_cfgs[index]['batch_first']
                else:
                    attn_cfgs[index]['batch_first'] = self.batch_first
                attention = build_attention(attn_cfgs[index])
                attention.operation_name = operation_name
                self.attentions.append(attention)
                index += 1
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = ModuleList()
        num_ffns = operation_order.count('ffn')
        if isinstance(ffn_cfgs, dict):
            ffn_cfgs = [copy.deepcopy(ConfigDict(ffn_cfgs)) for _ in range(num_ffns)]
        assert len(ffn_cfgs) == num_ffns
        for ffn_index in range(num_ffns):
            if 'embed_dims' not in ffn_cfgs[ffn_index]:
                ffn_cfgs[ffn_index]['embed_dims'] = self.embed_dims
            else:
                assert ffn_cfgs[ffn_index]['embed_dims'] == self.embed_dims
            self.ffns.append(build_feedforward_network(ffn_cfgs[ffn_index]))
        self.norms = ModuleList()
        for _ in range(operation_order.count('norm')):
            self.norms.append(build_norm_layer(norm_cfg, self.embed_dims)[1])
