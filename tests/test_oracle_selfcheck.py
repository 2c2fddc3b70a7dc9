"""Oracle self-checks that stand in for the golden vectors the reference does not ship."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import camera, decoder, lidar, model, voxel_pool

torch.manual_seed(0)


def test_voxel_pool_index_add_equals_kernel_loop():
    g = torch.Generator().manual_seed(1)
    geom = torch.randint(-2, 8, (2, 3, 5, 4, 3), generator=g, dtype=torch.int32)
    geom[..., 2] = torch.randint(-1, 2, geom[..., 2].shape, generator=g, dtype=torch.int32)
    feats = torch.randn(2, 3, 5, 4, 7, generator=g)
    vn = torch.tensor([6, 5, 1])
    a = voxel_pool.voxel_pooling_ref(geom, feats, vn)
    b, memo = voxel_pool.voxel_pooling_loops(geom, feats, vn)
    assert a.shape == (2, 7, 5, 6)
    assert torch.allclose(a, b, atol=1e-5)
    inside = (memo[..., 0] >= 0).sum()
    assert 0 < inside < geom[..., 0].numel()


def test_voxel_pool_empty_and_all_outside():
    vn = torch.tensor([4, 4, 1])
    out = voxel_pool.voxel_pooling_ref(torch.zeros(1, 0, 3, dtype=torch.int32), torch.zeros(1, 0, 5), vn)
    assert out.shape == (1, 5, 4, 4) and out.abs().sum() == 0
    geom = torch.full((1, 10, 3), 9, dtype=torch.int32)
    out = voxel_pool.voxel_pooling_ref(geom, torch.ones(1, 10, 5), vn)
    assert out.abs().sum() == 0


def test_rot_flip_is_anti_transpose():
    x = torch.randn(2, 3, 5, 5)
    y = model.anti_transpose(x)
    H = W = 5
    for i in range(H):
        for j in range(W):
            assert torch.equal(y[:, :, i, j], x[:, :, H - 1 - j, W - 1 - i])


def test_resnet50_matches_torchvision_trunk():
    import torchvision
    tv = torchvision.models.resnet50(weights=None).eval()
    mine = camera.ResNet50().eval()
    missing = mine.load_state_dict({k: v for k, v in tv.state_dict().items() if not k.startswith('fc.')}, strict=True)
    x = torch.randn(1, 3, 64, 96)
    with torch.no_grad():
        outs = mine(x)
        y = tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))
        ref = []
        for l in (tv.layer1, tv.layer2, tv.layer3, tv.layer4):
            y = l(y)
            ref.append(y)
    for a, b in zip(outs, ref):
        assert torch.allclose(a, b, atol=1e-5)


def test_dcn_zero_offset_is_grouped_conv():
    m = camera.DeformConv2dPack(16, 16, groups=4).eval()
    x = torch.randn(2, 16, 9, 11)
    with torch.no_grad():
        y = m(x)                                   # conv_offset is zero-init -> zero offsets
        ref = F.conv2d(x, m.weight, padding=1, groups=4)
    assert torch.allclose(y, ref, atol=1e-5)


def test_msda_matches_bruteforce_bilinear():
    g = torch.Generator().manual_seed(3)
    shapes = [(6, 8), (3, 4)]
    bs, nh, dh, nq, P = 2, 2, 4, 5, 3
    nk = sum(h * w for h, w in shapes)
    value = torch.randn(bs, nk, nh, dh, generator=g)
    loc = torch.rand(bs, nq, nh, len(shapes), P, 2, generator=g) * 1.4 - 0.2     # some outside [0,1]
    aw = torch.rand(bs, nq, nh, len(shapes), P, generator=g)
    out = decoder.msda_pytorch(value, torch.tensor(shapes), loc, aw)
    ref = torch.zeros(bs, nq, nh, dh)
    starts = np.cumsum([0] + [h * w for h, w in shapes])
    for b in range(bs):
        for q in range(nq):
            for h in range(nh):
                for l, (H, W) in enumerate(shapes):
                    for p in range(P):
                        x = float(loc[b, q, h, l, p, 0]) * W - 0.5
                        y = float(loc[b, q, h, l, p, 1]) * H - 0.5
                        x0, y0 = int(np.floor(x)), int(np.floor(y))
                        for dy in (0, 1):
                            for dx in (0, 1):
                                xi, yi = x0 + dx, y0 + dy
                                if 0 <= xi < W and 0 <= yi < H:
                                    wgt = (1 - abs(x - xi)) * (1 - abs(y - yi))
                                    ref[b, q, h] += aw[b, q, h, l, p] * wgt * value[b, starts[l] + yi * W + xi, h]
    assert torch.allclose(out.view(bs, nq, nh, dh), ref, atol=1e-5)


def _dense_from(st):
    return st.dense()


def test_sparse_conv_equals_masked_dense_conv3d():
    g = torch.Generator().manual_seed(5)
    shape = (7, 10, 9)
    B, Cin, Cout = 2, 3, 4
    mask = torch.rand(B, *shape, generator=g) < 0.15
    coords = mask.nonzero()
    feats = torch.randn(coords.shape[0], Cin, generator=g)
    x = lidar.SparseTensor(feats, coords, shape, B)
    dense_in = x.dense()
    # strided SparseConv3d
    for k, s, p in [(3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)]:
        conv = lidar.SparseConvBase(Cin, Cout, k, stride=s, padding=p)
        with torch.no_grad():
            y = conv(x)
            w = conv.weight.permute(0, 4, 1, 2, 3)                    # (Cout, Cin, kd, kh, kw)
            ref = F.conv3d(dense_in, w, stride=s, padding=p)
            active = F.conv3d(mask[:, None].float(), torch.ones(1, 1, *conv.k), stride=s, padding=p) > 0
        assert torch.allclose(y.dense(), ref, atol=1e-5)
        assert y.coords.shape[0] == int(active.sum())                   # output set = sites reached by an input
    # submanifold
    conv = lidar.SparseConvBase(Cin, Cout, 3, padding=1, subm=True)
    with torch.no_grad():
        y = conv(x)
        ref = F.conv3d(dense_in, conv.weight.permute(0, 4, 1, 2, 3), padding=1) * mask[:, None]
    assert torch.allclose(y.dense(), ref, atol=1e-5)
    assert torch.equal(y.coords, x.coords)


def test_hard_voxelize_matches_literal_loop():
    g = torch.Generator().manual_seed(7)
    pts = torch.rand(400, 5, generator=g)
    pts[:, :3] = pts[:, :3] * torch.tensor([2.4, 2.4, 1.2]) - torch.tensor([0.2, 0.2, 0.1])   # some outside
    pts[:100, :3] = pts[100:200, :3] * 0.999 + 0.0001                                          # crowd some voxels
    args = ([0.2, 0.2, 0.4], [0.0, 0.0, 0.0, 2.0, 2.0, 0.8], 3, 60)
    v, c, n = lidar.hard_voxelize(pts, *args)
    v2, c2, n2 = lidar.hard_voxelize_loop(pts.numpy(), *args)
    assert v.shape[0] == 60                                                                    # max_voxels cap hit
    assert torch.equal(c, c2) and torch.equal(n, n2) and torch.equal(v, v2)
    assert int(n.max()) == 3


def test_frame_coupling_of_look_module_is_reproduced():
    """frame j's output depends on the batch it is in (msda:338-342)."""
    from thinktwice_b200.config import Config, PLUMBING_CONFIG
    from thinktwice_b200.synthetic import make_batch
    cfg = Config.fromfile(PLUMBING_CONFIG)
    m = model.EncoderDecoder(**{k: v for k, v in cfg.model.items() if k != 'type'})
    model.init_oracle_weights(m, 0)
    b2 = make_batch(cfg, 2, seed=0, num_points=500)
    model.calibrate_bn(m, b2)
    b1 = make_batch(cfg, 1, seed=0, num_points=500)
    with torch.no_grad():
        p2 = m.forward_inference(b2)
        p1 = m.forward_inference(b1)
    # coarse prediction (no Look module) is per-frame ...
    assert torch.allclose(p2['pred_wp'][0, 0], p1['pred_wp'][0, 0], atol=1e-4)
    # ... the refined one is not
    assert not torch.allclose(p2['pred_wp'][0, 1], p1['pred_wp'][0, 1], atol=1e-4)
