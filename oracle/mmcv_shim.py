"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the mmcv / mmdet / mmdet3d
symbols that the reference's hot-path files import, so those files can be
imported UNMODIFIED from /root/reference inside the build container.

Nothing in the product package (vln_ver_b200/) may import this file; it is used
by oracle/gen_golden.py (fixture generation) and by tests that pin the oracle
restatement (oracle/ver_ref.py) against the reference's own Python.

Semantics restated from mmcv-full==1.4.0 / mmdet==2.14.0 (pinned by the
reference at docs/install.md:27-33; source NOT under /root/reference):
  * Registry.register_module()/build_from_cfg      (mmcv/utils/registry.py)
  * BaseModule / ModuleList / Sequential           (mmcv/runner/base_module.py)
  * xavier_init / constant_init (no-op on None)    (mmcv/cnn/utils/weight_init.py)
  * FFN, TransformerLayerSequence, build_* helpers (mmcv/cnn/bricks/transformer.py)
  * build_norm_layer(dict(type='LN'), C) = nn.LayerNorm(C, eps=1e-5)
  * mmdet FocalLoss(use_sigmoid=True) CPU path (py_sigmoid_focal_loss)
  * mmdet DETRHead.__init__ (only what VoxelFormerOccupancyHead relies on)
  * MultiheadAttention wrapper                      (mmcv/cnn/bricks/transformer.py); DetrTransformerDecoderLayer
    is built on the REFERENCE's own copy of BaseTransformerLayer (register_detr_decoder_layer)
"""
import copy
import importlib.util
import math
import os
import sys
import types
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get('VER_REFERENCE_ROOT', '/root/reference')


# --------------------------------------------------------------------------- utils
class ConfigDict(dict):
    """addict-like dict with attribute access (mmcv.utils.ConfigDict)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return ConfigDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(ConfigDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, ConfigDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            self._module_dict[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry._name} registry')
    else:
        obj_cls = obj_type
    return obj_cls(**args)


def deprecated_api_warning(name_dict, cls_name=None):
    def wrapper(fn):
        return fn
    return wrapper


def digit_version(v):
    out = []
    for x in v.split('+')[0].split('.'):
        num = ''.join(ch for ch in x if ch.isdigit())
        out.append(int(num) if num else 0)
    return tuple(out)


TORCH_VERSION = torch.__version__


class _ExtLoader:
    @staticmethod
    def load_ext(name, funcs):
        class _Missing:
            def __getattr__(self, item):
                raise RuntimeError(f'mmcv._ext.{item} is not available (CPU oracle shim)')
        return _Missing()


# --------------------------------------------------------------------------- runner
class BaseModule(nn.Module):
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


def _identity_decorator(*dargs, **dkwargs):
    def deco(fn):
        return fn
    return deco


force_fp32 = _identity_decorator
auto_fp16 = _identity_decorator


# --------------------------------------------------------------------------- cnn
def xavier_init(module, gain=1, bias=0, distribution='normal'):
    assert distribution in ['uniform', 'normal']
    if hasattr(module, 'weight') and module.weight is not None:
        if distribution == 'uniform':
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def bias_init_with_prob(prior_prob):
    return float(-math.log((1 - prior_prob) / prior_prob))


Linear = nn.Linear

ATTENTION = Registry('attention')
FEEDFORWARD_NETWORK = Registry('feed-forward Network')
POSITIONAL_ENCODING = Registry('position encoding')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
TRANSFORMER = Registry('Transformer')       # mmdet.models.utils.builder.TRANSFORMER
HEADS = Registry('head')
LOSSES = Registry('loss')
BBOX_CODERS = Registry('bbox_coder')

_ACT = {'ReLU': nn.ReLU, 'GELU': nn.GELU}


def build_activation_layer(cfg):
    cfg = dict(cfg)
    return _ACT[cfg.pop('type')](**cfg)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    t = cfg.pop('type')
    assert t == 'LN', t
    cfg.setdefault('eps', 1e-5)
    cfg.pop('requires_grad', None)
    return 'ln' + str(postfix), nn.LayerNorm(num_features, **cfg)


def build_dropout(cfg):
    cfg = dict(cfg)
    t = cfg.pop('type')
    assert t == 'Dropout'
    return nn.Dropout(cfg.get('drop_prob', 0.5))


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


@FEEDFORWARD_NETWORK.register_module()
class FFN(BaseModule):
    """mmcv 1.4.0 FFN: [Linear-act-Dropout]*(num_fcs-1), Linear, Dropout, + identity."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0.,
                 dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2
        self.embed_dims = embed_dims
        self.feedforward_channels = feedforward_channels
        self.num_fcs = num_fcs
        self.act_cfg = act_cfg
        self.activate = build_activation_layer(act_cfg)
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(Linear(in_channels, feedforward_channels),
                                     self.activate, nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


@ATTENTION.register_module()
class MultiheadAttention(BaseModule):
    """mmcv 1.4.0 `mmcv.cnn.bricks.transformer.MultiheadAttention` restated (mmcv is not
    vendored in the reference): a wrapper of nn.MultiheadAttention that adds the positional
    encodings to query / key, applies proj_drop + dropout_layer and adds the identity.
    vocc.py:141-145 builds it with embed_dims, num_heads=8, dropout=0.1 (deprecated kwarg:
    becomes attn_drop and dropout_layer.drop_prob)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        dropout_layer = dict(dropout_layer) if dropout_layer else dropout_layer
        if 'dropout' in kwargs:
            attn_drop = kwargs['dropout']
            dropout_layer['drop_prob'] = kwargs.pop('dropout')
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


def register_detr_decoder_layer():
    """mmcv's `DetrTransformerDecoderLayer` (vocc.py:139) = `BaseTransformerLayer` with the
    6-op order check.  The reference owns a verbatim copy of BaseTransformerLayer
    (M/custom_base_transformer_layer.py:37-260, `MyCustomBaseTransformerLayer`), so the shim class
    subclasses THAT (reference-owned forward) with mmcv's batch_first=False default."""
    if TRANSFORMER_LAYER.get('DetrTransformerDecoderLayer') is not None:
        return TRANSFORMER_LAYER.get('DetrTransformerDecoderLayer')
    base = import_reference('bevformer.modules.custom_base_transformer_layer').MyCustomBaseTransformerLayer

    class DetrTransformerDecoderLayer(base):
        def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                     act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), ffn_num_fcs=2,
                     **kwargs):
            kwargs.setdefault('batch_first', False)
            super().__init__(attn_cfgs=attn_cfgs, feedforward_channels=feedforward_channels,
                             ffn_dropout=ffn_dropout, operation_order=operation_order, act_cfg=act_cfg,
                             norm_cfg=norm_cfg, ffn_num_fcs=ffn_num_fcs, **kwargs)
            assert len(operation_order) == 6
            assert set(operation_order) == set(['self_attn', 'norm', 'cross_attn', 'ffn'])

    TRANSFORMER_LAYER.register_module()(DetrTransformerDecoderLayer)
    return DetrTransformerDecoderLayer


class TransformerLayerSequence(BaseModule):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(build_transformer_layer(transformerlayers[i]))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm


# --------------------------------------------------------------------------- mmdet bits
def py_sigmoid_focal_loss(pred, target, weight=None, gamma=2.0, alpha=0.25,
                          reduction='mean', avg_factor=None):
    """mmdet 2.14 py_sigmoid_focal_loss + weight_reduce_loss (target already one-hot)."""
    pred_sigmoid = pred.sigmoid()
    target = target.type_as(pred)
    pt = (1 - pred_sigmoid) * target + pred_sigmoid * (1 - target)
    focal_weight = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(pred, target, reduction='none') * focal_weight
    if weight is not None:
        if weight.shape != loss.shape:
            weight = weight.view(-1, 1)
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        return loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction == 'none':
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


@LOSSES.register_module()
class FocalLoss(nn.Module):
    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean',
                 loss_weight=1.0):
        super().__init__()
        assert use_sigmoid
        self.use_sigmoid = use_sigmoid
        self.gamma, self.alpha = gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        reduction = reduction_override if reduction_override else self.reduction
        num_classes = pred.size(1)
        target = F.one_hot(target, num_classes=num_classes + 1)[:, :num_classes]
        return self.loss_weight * py_sigmoid_focal_loss(
            pred, target, weight, gamma=self.gamma, alpha=self.alpha,
            reduction=reduction, avg_factor=avg_factor)


@LOSSES.register_module()
class L1Loss(nn.Module):
    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight


@LOSSES.register_module()
class GIoULoss(nn.Module):
    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight


def build_loss(cfg):
    return build_from_cfg(cfg, LOSSES)


def build_head(cfg):
    return build_from_cfg(cfg, HEADS)


class _Coder:
    def __init__(self, pc_range=None, **kwargs):
        self.pc_range = pc_range
        self.__dict__.update(kwargs)


def build_bbox_coder(cfg, **default_args):
    cfg = dict(cfg)
    cfg.pop('type')
    return _Coder(**cfg)


class DETRHead(BaseModule):
    """Only the constructor plumbing VoxelFormerOccupancyHead needs (mmdet 2.14
    DETRHead.__init__): loss_cls / positional_encoding / transformer are built
    from cfg, embed_dims is taken from the transformer, then _init_layers()."""
    _version = 2

    def __init__(self, num_classes, in_channels, num_query=100, num_reg_fcs=2,
                 transformer=None, sync_cls_avg_factor=False,
                 positional_encoding=None, loss_cls=None, loss_bbox=None, loss_iou=None,
                 train_cfg=None, test_cfg=None, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        self.bg_cls_weight = 0
        self.sync_cls_avg_factor = sync_cls_avg_factor
        self.num_query = num_query
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.num_reg_fcs = num_reg_fcs
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.fp16_enabled = False
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)
        self.loss_iou = build_loss(loss_iou)
        self.cls_out_channels = num_classes if self.loss_cls.use_sigmoid else num_classes + 1
        self.act_cfg = transformer.get('act_cfg', dict(type='ReLU', inplace=True))
        self.activate = build_activation_layer(self.act_cfg)
        self.positional_encoding = build_positional_encoding(positional_encoding)
        self.transformer = build_transformer(transformer)
        self.embed_dims = self.transformer.embed_dims
        num_feats = positional_encoding['num_feats']
        assert num_feats * 2 == self.embed_dims
        self._init_layers()


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def multi_apply(func, *args, **kwargs):
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def reduce_mean(t):
    return t


# --------------------------------------------------------------------------- install
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__ver_b200_shim__ = True
    sys.modules[name] = m
    return m


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    m.__ver_b200_shim__ = True
    sys.modules[name] = m
    return m


_INSTALLED = False


def install():
    """Put the fake mmcv/mmdet/mmdet3d modules and the synthetic parent packages
    of the reference's plugin into sys.modules (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    from oracle import ver_ref   # the restated 2-D sampler (pinned in tests)

    _pkg('mmcv').__dict__.update(ConfigDict=ConfigDict, deprecated_api_warning=deprecated_api_warning)
    _mod('mmcv.utils', ConfigDict=ConfigDict, build_from_cfg=build_from_cfg,
         deprecated_api_warning=deprecated_api_warning, ext_loader=_ExtLoader,
         TORCH_VERSION=TORCH_VERSION, digit_version=digit_version, Registry=Registry,
         to_2tuple=lambda x: x if isinstance(x, (tuple, list)) else (x, x))
    _mod('mmcv.utils.ext_loader', load_ext=_ExtLoader.load_ext)
    _mod('mmcv.runner', force_fp32=force_fp32, auto_fp16=auto_fp16, BaseModule=BaseModule,
         ModuleList=ModuleList, Sequential=Sequential)
    _mod('mmcv.runner.base_module', BaseModule=BaseModule, ModuleList=ModuleList, Sequential=Sequential)
    _pkg('mmcv.cnn').__dict__.update(
        xavier_init=xavier_init, constant_init=constant_init, Linear=Linear,
        bias_init_with_prob=bias_init_with_prob, build_activation_layer=build_activation_layer,
        build_norm_layer=build_norm_layer)
    _pkg('mmcv.cnn.bricks')
    _mod('mmcv.cnn.bricks.registry', ATTENTION=ATTENTION, FEEDFORWARD_NETWORK=FEEDFORWARD_NETWORK,
         POSITIONAL_ENCODING=POSITIONAL_ENCODING, TRANSFORMER_LAYER=TRANSFORMER_LAYER,
         TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE)
    _mod('mmcv.cnn.bricks.transformer', build_attention=build_attention,
         build_feedforward_network=build_feedforward_network,
         build_positional_encoding=build_positional_encoding,
         build_transformer_layer=build_transformer_layer,
         build_transformer_layer_sequence=build_transformer_layer_sequence,
         TransformerLayerSequence=TransformerLayerSequence, FFN=FFN,
         POSITIONAL_ENCODING=POSITIONAL_ENCODING, ATTENTION=ATTENTION)
    _pkg('mmcv.ops')
    _mod('mmcv.ops.multi_scale_deform_attn',
         multi_scale_deformable_attn_pytorch=ver_ref.multi_scale_deformable_attn_pytorch)

    _pkg('mmdet')
    _mod('mmdet.core', multi_apply=multi_apply, reduce_mean=reduce_mean)
    _pkg('mmdet.models').__dict__.update(HEADS=HEADS)
    _pkg('mmdet.models.utils')
    _mod('mmdet.models.utils.builder', TRANSFORMER=TRANSFORMER)
    _mod('mmdet.models.utils.transformer', inverse_sigmoid=inverse_sigmoid)
    _mod('mmdet.models.dense_heads', DETRHead=DETRHead)
    _pkg('mmdet3d')
    _pkg('mmdet3d.core')
    _pkg('mmdet3d.core.bbox')
    _mod('mmdet3d.core.bbox.coders', build_bbox_coder=build_bbox_coder)
    _pkg('mmdet3d.models')
    _mod('mmdet3d.models.builder', build_loss=build_loss, build_head=build_head)
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            _mod('h5py')
    # voxel_decoder.py:9-12 imports cv2 / matplotlib at module level and never uses them
    for name in ('cv2', 'matplotlib'):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                _pkg(name)
                if name == 'matplotlib':
                    sys.modules['matplotlib'].pyplot = _mod('matplotlib.pyplot')

    # synthetic parents: __path__ points into the reference so its (broken)
    # package __init__ files are never executed
    plug = os.path.join(REFERENCE_ROOT, 'projects', 'mmdet3d_plugin')
    _pkg('projects', os.path.join(REFERENCE_ROOT, 'projects'))
    _pkg('projects.mmdet3d_plugin', plug)
    _pkg('projects.mmdet3d_plugin.bevformer', os.path.join(plug, 'bevformer'))
    _pkg('projects.mmdet3d_plugin.bevformer.modules', os.path.join(plug, 'bevformer', 'modules'))
    _pkg('projects.mmdet3d_plugin.bevformer.dense_heads', os.path.join(plug, 'bevformer', 'dense_heads'))
    _pkg('projects.mmdet3d_plugin.models')
    _pkg('projects.mmdet3d_plugin.models.utils')
    _mod('projects.mmdet3d_plugin.models.utils.bricks', run_time=lambda name: (lambda fn: fn))
    _mod('projects.mmdet3d_plugin.models.utils.visual', save_tensor=lambda *a, **k: None)
    _pkg('projects.mmdet3d_plugin.core')
    _pkg('projects.mmdet3d_plugin.core.bbox')
    _mod('projects.mmdet3d_plugin.core.bbox.util', normalize_bbox=lambda *a, **k: None)
    # `voxel_transformer.py:19` imports the absent `decoder.py` (R8)
    _mod('projects.mmdet3d_plugin.bevformer.modules.decoder',
         CustomMSDeformableAttention=type('CustomMSDeformableAttention', (nn.Module,), {}))
    _INSTALLED = True


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'projects', 'mmdet3d_plugin'))


def import_reference(modname):
    """import e.g. 'bevformer.modules.spatial_cross_attention' unmodified."""
    install()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return importlib.import_module('projects.mmdet3d_plugin.' + modname)
