"""CPU: the oracle reproduces the committed golden vectors (tests/golden, made by make_golden.py)."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_oracle_ops_match_golden_vectors():
    from oracle import decoder, lidar, voxel_pool
    o = np.load(os.path.join(G, 'ops.npz'))
    out = voxel_pool.voxel_pooling_ref(torch.from_numpy(o['vp_geom']), torch.from_numpy(o['vp_feats']), torch.tensor([6, 5, 1]))
    assert np.allclose(out.numpy(), o['vp_out'], atol=1e-5)
    m = decoder.msda_pytorch(torch.from_numpy(o['msda_value']), torch.tensor([(6, 8), (3, 4), (2, 2), (1, 2)]),
                             torch.from_numpy(o['msda_loc']), torch.from_numpy(o['msda_aw']))
    assert np.allclose(m.numpy(), o['msda_out'], atol=1e-5)
    v, c, n = lidar.hard_voxelize(torch.from_numpy(o['vox_pts']), [0.2, 0.2, 0.4], [0.0, 0.0, 0.0, 2.0, 2.0, 0.8], 3, 1000)
    assert np.array_equal(c.numpy(), o['vox_coors']) and np.array_equal(n.numpy(), o['vox_num'])
    assert np.allclose((v.sum(1) / n.float().view(-1, 1)).numpy(), o['vox_mean'], atol=1e-6)
    conv = lidar.SparseConvBase(4, 8, 3, stride=2, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(o['sp_weight']))
        y = conv(lidar.SparseTensor(torch.from_numpy(o['sp_feats']), torch.from_numpy(o['sp_coords']), (7, 10, 9), 1))
    assert np.allclose(y.dense().numpy(), o['sp_dense'], atol=1e-5)


@pytest.mark.parametrize('seed', [0])
def test_oracle_plumbing_forward_matches_golden(seed):
    import sys
    sys.path.insert(0, G)
    from make_golden import plumbing
    got, ref = plumbing(seed), np.load(os.path.join(G, f'plumbing_seed{seed}.npz'))
    for k in ref.files:
        scale = np.abs(ref[k]).max() + 1e-12
        assert np.abs(got[k] - ref[k]).max() / scale < 1e-4, k
