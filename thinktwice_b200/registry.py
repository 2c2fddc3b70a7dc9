"""Registry shim with mmcv's `Registry.register_module()` / `build(cfg)` behaviour.

The reference registers its modules into mmdet / mmdet3d registries (encoder_decoder_framework.py:23,
lss.py:286,351, lidarnet.py:24,61, thinktwice_decoder.py:262) and builds them from config dicts with a
`type` key (`mmdet3d.models.builder.build_backbone/build_head/build_model`).  mmcv is not installed in
this image, so the same contract is provided here; when mmdet IS importable the classes are also
registered there, which makes `build_model(cfg.model)` in the reference's agent
and trainer pick up these implementations unchanged.
"""


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise TypeError(f'cfg must be a dict with a "type" key, got {cfg!r}')
        args = dict(cfg)
        t = args.pop('type')
        cls = self.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f'{t} is not in the {self.name} registry')
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        return cls(**args)


DETECTORS = Registry('detector')
BACKBONES = Registry('backbone')
NECKS = Registry('neck')
HEADS = Registry('head')
MIDDLE_ENCODERS = Registry('middle_encoder')


def build_backbone(cfg, **kw):
    return BACKBONES.build(cfg, kw or None)


def build_head(cfg, **kw):
    return HEADS.build(cfg, kw or None)


def build_model(cfg, train_cfg=None, test_cfg=None):
    """mmdet3d.models.build_model(cfg.model, train_cfg=..., test_cfg=...) (thinktwice_agent.py:168)."""
    extra = {}
    if train_cfg is not None:
        extra['train_cfg'] = train_cfg
    if test_cfg is not None:
        extra['test_cfg'] = test_cfg
    return DETECTORS.build(cfg, extra or None)


def register_into_mmdet():
    """If the OpenMMLab stack is importable, publish the classes into its registries (force=True)."""
    try:
        from mmdet.models import BACKBONES as B, DETECTORS as D, HEADS as H   # noqa
    except Exception:
        return False
    for reg, theirs in ((BACKBONES, B), (DETECTORS, D), (HEADS, H)):
        for k, v in reg.module_dict.items():
            theirs.register_module(name=k, force=True, module=v)
    return True
