"""SURVEY §8f row f4 — the training-side native ops through the C ABI against torch autograd of the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize('channels_last_grad,Cc', [(False, 80), (True, 80), (False, 6), (True, 6)])
def test_voxel_pooling_backward_kernel_matches_the_reference_backward(channels_last_grad, Cc):
    """ops/voxel_pooling/voxel_pooling.py:57-69: the gradient of a kept point is the output gradient at its recorded (b, y, x); dropped
    points get zero.  Checked against autograd through the oracle's index_add restatement AND the literal indexing of the reference."""
    from oracle.voxel_pool import voxel_pooling_ref
    from thinktwice_b200.ops.voxel_pooling import voxel_pooling
    from thinktwice_b200.ops.voxel_pooling.voxel_pooling import last_pos_memo
    g = torch.Generator().manual_seed(5)
    B, P, X, Y, Z = 2, 5000, 21, 21, 1                                    # C = 80: four channels per thread; C = 6: the scalar path
    geom = torch.stack([torch.randint(-3, X + 3, (B, P), generator=g), torch.randint(-3, Y + 3, (B, P), generator=g),
                        torch.randint(-1, Z + 1, (B, P), generator=g)], -1).int()
    feats = torch.randn(B, P, Cc, generator=g)
    gout = torch.randn(B, Cc, Y, X, generator=g)
    f_ref = feats.clone().requires_grad_(True)
    voxel_pooling_ref(geom, f_ref, (X, Y, Z)).backward(gout)
    f = feats.cuda().requires_grad_(True)
    out = voxel_pooling(geom.cuda().contiguous(), f, torch.tensor([X, Y, Z]))
    go = gout.cuda()
    if channels_last_grad:                                              # the (B, C, Y, X) view of an NHWC buffer: what permute() upstream produces
        go = go.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    out.backward(go)
    assert torch.equal(f.grad.cpu(), f_ref.grad)                        # a pure gather: bit-exact
    memo = last_pos_memo().cpu()
    kept = (memo != -1)[..., 0]
    lit = torch.zeros(B, P, Cc)
    lit[kept] = gout[memo[kept][..., 0].long(), :, memo[kept][..., 1].long(), memo[kept][..., 2].long()]
    assert torch.equal(f.grad.cpu(), lit) and 0 < int(kept.sum()) < B * P


@pytest.mark.parametrize('heads,dh,levels', [(8, 32, 4), (4, 16, 2), (2, 48, 3)])
def test_ms_deform_attn_forward_and_backward_match_autograd_of_the_oracle(heads, dh, levels):
    """mmcv's ms_deform_attn_forward / backward contract (msda:141-147, 172-183).  Sampling locations reach outside [0, 1] (zero
    padding, partially visible bilinear footprints); gradients of value (atomics), locations and weights vs float64 autograd."""
    from oracle.decoder import msda_pytorch
    from thinktwice_b200.ops.ms_deform_attn import MultiScaleDeformableAttnFunction_fp32
    g = torch.Generator().manual_seed(11 + heads)
    shapes = [(14, 28), (7, 14), (4, 7), (2, 4)][:levels]
    starts, k = [], 0
    for h, w in shapes:
        starts.append(k)
        k += h * w
    bs, nq, P = 2, 37, 8
    value = torch.randn(bs, k, heads, dh, generator=g)
    loc = torch.rand(bs, nq, heads, levels, P, 2, generator=g) * 1.3 - 0.15
    aw = torch.rand(bs, nq, heads, levels, P, generator=g).flatten(-2).softmax(-1).view(bs, nq, heads, levels, P)
    gout = torch.randn(bs, nq, heads * dh, generator=g)
    v64, l64, a64 = (t.double().requires_grad_(True) for t in (value, loc, aw))
    ref = msda_pytorch(v64, shapes, l64, a64)
    ref.backward(gout.double())
    vc, lc, ac = (t.cuda().requires_grad_(True) for t in (value, loc, aw))
    ss = torch.tensor(shapes, dtype=torch.long, device='cuda')
    out = MultiScaleDeformableAttnFunction_fp32.apply(vc, ss, torch.tensor(starts, dtype=torch.long, device='cuda'), lc, ac, 64)
    assert rel(out, ref) < 1e-5
    out.backward(gout.cuda())
    assert rel(vc.grad, v64.grad) < 1e-5
    assert rel(ac.grad, a64.grad) < 1e-5
    # the sampled value is piecewise bilinear in the location: its derivative jumps where a pixel coordinate crosses an integer, and a
    # float32 coordinate within rounding of an integer falls on the other side than the float64 reference (seen: 18.999999 vs 19.0).
    # Those samples are compared through float32 autograd of the oracle (same side of the jump), all the others against float64.
    near = torch.zeros(loc.shape[:-1], dtype=torch.bool)
    for l, (h, w) in enumerate(shapes):
        for axis, size in ((0, w), (1, h)):
            pix = loc[:, :, :, l, :, axis].double() * size - 0.5
            near[:, :, :, l] |= (pix - pix.round()).abs() < 1e-4
    v32, l32, a32 = (t.clone().requires_grad_(True) for t in (value, loc, aw))
    msda_pytorch(v32, shapes, l32, a32).backward(gout)
    far = (~near)[..., None].expand_as(loc)
    scale = float(l64.grad.abs().max())
    assert float((lc.grad.cpu().double() - l64.grad)[far].abs().max()) < 1e-4 * scale     # differences of neighbouring values scaled by the map size
    assert float((lc.grad.cpu() - l32.grad)[~far].abs().max() if (~far).any() else 0.0) < 1e-3 * scale
    assert float(lc.grad.abs().max()) > 0 and int(near.sum()) < near.numel() // 100
