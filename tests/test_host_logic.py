"""CPU tests of the host-side orchestration: the whole forward is walked with the C-ABI calls stubbed out, so
wrong names / shapes / weight keys in the Python layer surface without a GPU.  (Numerics are the GPU tests' job.)"""
import ctypes as C

import pytest
import torch


class _FakeLib:
    def __getattr__(self, name):
        def f(*a):
            return 4096 if name.endswith('_bytes') else 0
        return f


@pytest.fixture
def stubbed(monkeypatch):
    from thinktwice_b200 import lib
    calls = []
    monkeypatch.setattr(lib, 'load', lambda: _FakeLib())
    monkeypatch.setattr(lib, 'require_cuda', lambda dev: None)
    monkeypatch.setattr(lib, '_p', lambda t, off=0: C.c_void_p(0))
    monkeypatch.setattr(lib, '_stream', lambda: C.c_void_p(0))
    monkeypatch.setattr(lib, 'call', lambda name, *a: calls.append(name))
    import thinktwice_b200.engine as engine
    import thinktwice_b200.lss as lss
    import thinktwice_b200.lidarnet as lidarnet
    import thinktwice_b200.thinktwice_decoder as dec
    for mod in (engine, lss, lidarnet, dec):
        monkeypatch.setattr(mod, '_p', lib._p, raising=False)
        monkeypatch.setattr(mod, '_stream', lib._stream, raising=False)
    return calls


@pytest.mark.parametrize('which,B', [('plumbing', 1), ('plumbing', 2)])
def test_forward_walk_with_stubbed_library(stubbed, which, B):
    from thinktwice_b200.config import Config, PLUMBING_CONFIG
    from thinktwice_b200.registry import build_model
    from thinktwice_b200.synthetic import make_batch
    cfg = Config.fromfile(PLUMBING_CONFIG)
    model = build_model(cfg.model)
    model.prepare('cpu')
    batch = make_batch(cfg, B, seed=0, num_points=300)
    pred = model.forward_inference(batch)
    K, T = cfg.model.train_cfg['refine_num'], 4
    assert pred['pred_wp'].shape == (B, K + 1, T, 2)
    assert pred['mu_branches'].shape == (B, K + 1, 2) and pred['future_sigma'].shape == (B, K + 1, 3, 2)
    assert pred['refine_BEV_feature'].shape == (B, K, 32, 21, 21)
    assert pred['refine_future_BEV_feature'].shape == (B, K, T, 32, 21, 21)
    assert pred['refine_flattned_BEV_feature'].shape == (B, K, 256)
    assert {'tt_lift_splat', 'tt_voxelize_mean', 'tt_sparse_rulebook', 'tt_msda_forward', 'tt_look_project',
            'tt_look_rebatch', 'tt_look_reduce', 'tt_dcn_im2col', 'tt_gru_input'} <= set(stubbed)
    seg = model.last_cam_feat['seg']
    assert (seg.N, seg.H, seg.W, seg.C) == (4 * B, 128, 128, 12)


def test_state_dict_roundtrip_with_oracle_names():
    from oracle.model import EncoderDecoder as Oracle
    from thinktwice_b200.config import Config, PLUMBING_CONFIG
    from thinktwice_b200.registry import build_model
    cfg = Config.fromfile(PLUMBING_CONFIG)
    o = Oracle(**{k: v for k, v in cfg.model.items() if k != 'type'})
    m = build_model(cfg.model)
    sd_o, sd_m = o.state_dict(), m.state_dict()
    assert set(sd_o) == set(sd_m)
    assert all(sd_o[k].shape == sd_m[k].shape for k in sd_o)
    assert torch.equal(sd_o['img_encoder.frustum'], sd_m['img_encoder.frustum'])      # same fp32 frustum buffer
    assert torch.equal(sd_o['img_encoder.voxel_coord'], sd_m['img_encoder.voxel_coord'])
    m.load_state_dict(sd_o)
