"""Stand-ins for the third-party packages the reference imports at module load (mmcv / mmdet / mmdet3d / mmcls / cv2 /
matplotlib), so that the reference's OWN Python files under /root/reference can be imported and executed in this image —
used only by make_reference_golden.py (fixture generation; never imported by the product, the tests or the bench).

Only import-time plumbing is stubbed (registries, BaseModule, fp16 decorators, init helpers).  Third-party ARITHMETIC the
reference calls (mmcv's pure-torch multi_scale_deformable_attn_pytorch) is supplied from its published semantics by the
caller and named in the fixture metadata.
"""
import sys
import types

import torch.nn as nn


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


class Registry:
    def __init__(self, name):
        self.name, self.classes = name, {}

    def register_module(self, *a, **k):
        def deco(cls):
            self.classes[cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg, **kw):
        cfg = dict(cfg)
        return self.classes[cfg.pop('type')](**cfg, **kw)


def _identity_decorator(*a, **k):
    def deco(f):
        return f
    return deco


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None, *a, **k):
        super().__init__()


def _xavier_init(m, gain=1, bias=0, distribution='normal'):          # mmcv guards on hasattr(module, 'weight') the same way
    if getattr(m, 'weight', None) is not None:
        (nn.init.xavier_uniform_ if distribution == 'uniform' else nn.init.xavier_normal_)(m.weight, gain=gain)
    if getattr(m, 'bias', None) is not None:
        nn.init.constant_(m.bias, bias)


def _constant_init(m, val, bias=0):
    if getattr(m, 'weight', None) is not None:
        nn.init.constant_(m.weight, val)
    if getattr(m, 'bias', None) is not None:
        nn.init.constant_(m.bias, bias)


def install():
    for n in ['mmcv', 'mmcv.cnn', 'mmcv.cnn.bricks', 'mmcv.cnn.bricks.registry', 'mmcv.cnn.bricks.transformer', 'mmcv.runner',
              'mmcv.runner.base_module', 'mmcv.utils', 'mmcv.ops', 'mmcv.ops.multi_scale_deform_attn', 'mmdet', 'mmdet.core',
              'mmdet.models', 'mmdet3d', 'mmdet3d.models', 'mmdet3d.models.builder', 'mmcls', 'mmcls.models', 'cv2', 'matplotlib',
              'matplotlib.pyplot']:
        _mod(n)
    S = sys.modules
    S['mmcv.runner'].BaseModule = BaseModule
    S['mmcv.runner'].force_fp32 = _identity_decorator
    S['mmcv.runner'].auto_fp16 = _identity_decorator
    S['mmcv.runner.base_module'].BaseModule = BaseModule
    S['mmcv.runner.base_module'].ModuleList = nn.ModuleList
    S['mmcv.runner.base_module'].Sequential = nn.Sequential
    S['mmcv.cnn'].build_conv_layer = None
    S['mmcv.cnn'].xavier_init = _xavier_init
    S['mmcv.cnn'].constant_init = _constant_init
    for r in ('ATTENTION', 'TRANSFORMER_LAYER', 'TRANSFORMER_LAYER_SEQUENCE'):
        setattr(S['mmcv.cnn.bricks.registry'], r, Registry(r))
    S['mmcv.cnn.bricks.transformer'].TransformerLayerSequence = BaseModule
    S['mmcv.utils'].ext_loader = types.SimpleNamespace(load_ext=lambda *a, **k: None)
    S['mmcv.utils'].ConfigDict = dict
    S['mmcv.utils'].build_from_cfg = lambda cfg, reg, default_args=None: reg.build(cfg)
    S['mmcv.utils'].deprecated_api_warning = _identity_decorator
    S['mmcv.utils'].to_2tuple = lambda x: (x, x)
    S['mmdet.core'].multi_apply = lambda f, *a, **k: tuple(map(list, zip(*map(f, *a))))
    S['mmdet.core'].reduce_mean = lambda t: t
    regs = {r: Registry(r) for r in ('HEADS', 'BACKBONES', 'NECKS', 'DETECTORS')}
    for r, v in regs.items():
        setattr(S['mmdet.models'], r, v)
    b = S['mmdet3d.models.builder']
    S['mmdet3d.models'].builder = b
    b.MIDDLE_ENCODERS = Registry('MIDDLE_ENCODERS')
    b.build_backbone = lambda cfg: None if cfg is None else regs['BACKBONES'].build(cfg)
    b.build_head = lambda cfg: regs['HEADS'].build(cfg)
    S['matplotlib'].use = lambda *a, **k: None
    return regs
