"""Restatement of mmcv 1.x's checkpoint loader (mmcv/runner/checkpoint.py: `load_state_dict`, `load_checkpoint`) — the
call the reference agent makes at leaderboard/team_code/thinktwice_agent.py:170.  TEST INFRASTRUCTURE ONLY (mmcv is not
installed in this image); restated from its published source:

  * `load_checkpoint(model, filename, map_location, strict=False)` reads the file with torch.load, takes
    checkpoint['state_dict'] when present, strips a leading 'module.' from every key and calls `load_state_dict`;
  * `load_state_dict(module, state_dict, strict=False)` does NOT call `module.load_state_dict`: it walks `module._modules`
    recursively and calls `_load_from_state_dict(state_dict, prefix, local_metadata, True, missing, unexpected, err)` on
    every sub-module, then reports `unexpected` / `missing` (minus 'num_batches_tracked') keys — as a warning only when
    strict is False, which is the agent's case.  A model whose parameters are not reachable through `_modules` under the
    checkpoint's names therefore loads NOTHING and only warns.
"""
import torch


def load_state_dict(module, state_dict, strict=False):
    unexpected_keys, all_missing_keys, err_msg = [], [], []
    metadata = getattr(state_dict, '_metadata', None)
    state_dict = state_dict.copy()
    if metadata is not None:
        state_dict._metadata = metadata

    def load(mod, prefix=''):
        local_metadata = {} if metadata is None else metadata.get(prefix[:-1], {})
        mod._load_from_state_dict(state_dict, prefix, local_metadata, True, all_missing_keys, unexpected_keys, err_msg)
        for name, child in mod._modules.items():
            if child is not None:
                load(child, prefix + name + '.')

    load(module)
    missing_keys = [k for k in all_missing_keys if 'num_batches_tracked' not in k]
    # torch (and mmcv) compute "unexpected" per module only for keys under that module's own prefix; a key nobody
    # consumed at all is found by the set difference below (mmcv relies on the root call for this)
    consumed = set(n for n, _ in module.named_parameters()) | set(n for n, _ in module.named_buffers())
    unexpected_keys = sorted(set(unexpected_keys) | (set(state_dict.keys()) - consumed))
    if strict and (missing_keys or unexpected_keys or err_msg):
        raise RuntimeError(f'missing {missing_keys[:5]} unexpected {unexpected_keys[:5]} {err_msg[:2]}')
    return missing_keys, unexpected_keys, err_msg


def load_checkpoint(model, filename, map_location=None, strict=False):
    checkpoint = torch.load(filename, map_location=map_location, weights_only=False)
    if not isinstance(checkpoint, dict):
        raise RuntimeError(f'No state_dict found in checkpoint file {filename}')
    state_dict = checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint
    if list(state_dict.keys())[0].startswith('module.'):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    missing, unexpected, err = load_state_dict(model, state_dict, strict)
    checkpoint['_load_report'] = dict(missing=missing, unexpected=unexpected, err=err)
    return checkpoint
