"""mmcv-free loader for mmcv-style python configs (SURVEY.md §5 "Config / flags").

The reference reads `open_loop_training/configs/thinktwice.py` through `mmcv.Config.fromfile`
(train.py:116, thinktwice_agent.py:154): a python file exec'd into a dict, `_base_` files merged
underneath, and every nested dict given attribute access.  mmcv is absent here, so this restates
that contract: `Config.fromfile(path)` -> object with item *and* attribute access.
"""
import os


class ConfigDict(dict):
    """dict with attribute access (mmcv.ConfigDict behaviour used by the path: `config.pred_len`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


def _merge(base, over):
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get('_delete_', False):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load(path):
    path = os.path.abspath(path)
    scope = {'__file__': path}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), scope)
    cfg = {k: v for k, v in scope.items()
           if not k.startswith('__') and not callable(v) and not isinstance(v, type(os))}
    bases = cfg.pop('_base_', [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


class Config(ConfigDict):
    @staticmethod
    def fromfile(path):
        return Config(_wrap(_load(path)))

    def merge_from_dict(self, options):
        """`--cfg-options a.b=c` style overrides (train.py:79-88)."""
        for key, val in options.items():
            node = self
            parts = key.split('.')
            for p in parts[:-1]:
                node = node.setdefault(p, ConfigDict())
            node[parts[-1]] = _wrap(val)


DEFAULT_CONFIG = os.path.join(os.path.dirname(__file__), 'configs', 'thinktwice.py')
PLUMBING_CONFIG = os.path.join(os.path.dirname(__file__), 'configs', 'plumbing.py')
