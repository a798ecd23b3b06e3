"""Registry plumbing for the drop-in boundary B1.

If the real mmcv / mmdet are importable the modules register into THEIR registries
(`ATTENTION.register_module()` etc., exactly like
projects/mmdet3d_plugin/bevformer/modules/spatial_cross_attention.py:31), so
`build_from_cfg` on the `model=dict(...)` tree of vocc.py finds them.  Otherwise a
local registry with the same `register_module()` / `build()` semantics is used, and
the small mmcv pieces the reference relies on (BaseModule, FFN, LayerNorm builder,
xavier/constant init, ConfigDict) are provided here with mmcv 1.4.0 behaviour.
"""
import copy
import math

import torch.nn as nn


def _real(modname):
    """import `modname` unless it is absent or the oracle's test shim."""
    try:
        mod = __import__(modname, fromlist=['_'])
    except Exception:  # noqa: BLE001
        return None
    return None if getattr(mod, '__ver_b200_shim__', False) else mod


_mmcv_registry = _real('mmcv.cnn.bricks.registry') if _real('mmcv') else None
HAVE_MMCV = _mmcv_registry is not None


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if not force and key in self._module_dict:
                raise KeyError(f'{key} is already registered in {self._name}')
            self._module_dict[key] = cls
            return cls
        return _register(module) if module is not None else _register

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
    args = copy.copy(dict(cfg))
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    else:
        obj_cls = obj_type
    return obj_cls(**args)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = v


if HAVE_MMCV:                                         # pragma: no cover (mmcv absent in this image)
    ATTENTION = _mmcv_registry.ATTENTION
    FEEDFORWARD_NETWORK = _mmcv_registry.FEEDFORWARD_NETWORK
    POSITIONAL_ENCODING = _mmcv_registry.POSITIONAL_ENCODING
    TRANSFORMER_LAYER = _mmcv_registry.TRANSFORMER_LAYER
    TRANSFORMER_LAYER_SEQUENCE = _mmcv_registry.TRANSFORMER_LAYER_SEQUENCE
    _b = _real('mmdet.models.utils.builder')
    TRANSFORMER = _b.TRANSFORMER if _b else Registry('Transformer')
    _m = _real('mmdet.models')
    HEADS = _m.HEADS if _m else Registry('head')
    LOSSES = _m.LOSSES if _m else Registry('loss')
    from mmcv.runner.base_module import BaseModule, ModuleList, Sequential    # noqa: F401
else:
    ATTENTION = Registry('attention')
    FEEDFORWARD_NETWORK = Registry('feed-forward Network')
    POSITIONAL_ENCODING = Registry('position encoding')
    TRANSFORMER_LAYER = Registry('transformerLayer')
    TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
    TRANSFORMER = Registry('Transformer')
    HEADS = Registry('head')
    LOSSES = Registry('loss')

    class BaseModule(nn.Module):
        """mmcv.runner.BaseModule contract: `init_cfg`, `init_weights()`, `_is_init`."""

        def __init__(self, init_cfg=None):
            super().__init__()
            self._is_init = False
            self.init_cfg = copy.deepcopy(init_cfg)

        @property
        def is_init(self):
            return self._is_init

        def init_weights(self):
            for m in self.children():
                if hasattr(m, 'init_weights'):
                    m.init_weights()
            self._is_init = True

    class ModuleList(BaseModule, nn.ModuleList):
        def __init__(self, modules=None, init_cfg=None):
            BaseModule.__init__(self, init_cfg)
            nn.ModuleList.__init__(self, modules)

    class Sequential(BaseModule, nn.Sequential):
        def __init__(self, *args, init_cfg=None):
            BaseModule.__init__(self, init_cfg)
            nn.Sequential.__init__(self, *args)


def build_attention(cfg, default_args=None):
    return build_from_cfg(cfg, ATTENTION, default_args)


def build_feedforward_network(cfg, default_args=None):
    return build_from_cfg(cfg, FEEDFORWARD_NETWORK, default_args)


def build_positional_encoding(cfg, default_args=None):
    return build_from_cfg(cfg, POSITIONAL_ENCODING, default_args)


def build_transformer_layer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER, default_args)


def build_transformer_layer_sequence(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER_SEQUENCE, default_args)


def build_transformer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER, default_args)


def build_loss(cfg):
    return build_from_cfg(cfg, LOSSES)


def build_norm_layer(cfg, num_features, postfix=''):
    """mmcv build_norm_layer for dict(type='LN'): ('ln', nn.LayerNorm(C, eps=1e-5))."""
    cfg = dict(cfg)
    layer_type = cfg.pop('type')
    if layer_type != 'LN':
        raise KeyError(f'Unrecognized norm type {layer_type} (only LN is on the VER lift path)')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    layer = nn.LayerNorm(num_features, **cfg)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return 'ln' + str(postfix), layer


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    """mmcv.cnn.xavier_init -- a no-op on None, which the reference relies on
    (spatial_cross_attention.py:272 passes output_proj=None)."""
    assert distribution in ['uniform', 'normal']
    if hasattr(module, 'weight') and module.weight is not None:
        (nn.init.xavier_uniform_ if distribution == 'uniform' else nn.init.xavier_normal_)(
            module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def bias_init_with_prob(prior_prob):
    return float(-math.log((1 - prior_prob) / prior_prob))
