"""CPU checks of the weight repacking (thinktwice_b200/weights.py): each packed layout is replayed with plain torch
GEMMs exactly the way the kernels index it and compared with the torch convolution it stands for."""
import torch
import torch.nn.functional as F

from thinktwice_b200.weights import Packer, tf32_split

CPU = torch.device('cpu')


def test_tf32_split_is_exact_to_21_bits_and_tf32_representable():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(4096, generator=g) * torch.logspace(-6, 6, 4096)
    hi, lo = tf32_split(w)
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0            # 13 low mantissa bits clear
    assert int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float(((hi.double() + lo.double() - w.double()).abs() / w.double().abs()).max()) < 2.0 ** -21
    assert float(((hi.double() - w.double()).abs() / w.double().abs()).max()) <= 2.0 ** -11


def test_rowpacked_stem_layout_reproduces_conv7x7_s2():
    """the kernels read, for kernel row kh, 32 contiguous floats = 8 pixels x 4 floats of the zero-bordered image."""
    g = torch.Generator().manual_seed(1)
    N, H, W, Cout = 2, 12, 20, 8
    x = torch.randn(N, 3, H, W, generator=g, dtype=torch.float64)
    sd = {'c.weight': torch.randn(Cout, 3, 7, 7, generator=g), 'b.weight': torch.rand(Cout, generator=g) + 0.5,
          'b.bias': torch.randn(Cout, generator=g), 'b.running_mean': torch.randn(Cout, generator=g),
          'b.running_var': torch.rand(Cout, generator=g) + 0.5}
    pw = Packer(sd, CPU, tc_mode=0).conv_rowpacked('c', bn='b')
    assert (pw.Cin, pw.KH, pw.KW, pw.alg_k) == (32, 7, 1, 147)
    pad = torch.zeros(N, H + 6, W + 8, 4, dtype=torch.float64)
    pad[:, 3:3 + H, 3:3 + W, :3] = x.permute(0, 2, 3, 1)
    flat = pad.reshape(N, H + 6, (W + 8) * 4)
    OH, OW = H // 2, W // 2
    out = torch.zeros(N, OH, OW, Cout, dtype=torch.float64)
    wk = pw.w.double().view(7, 32, Cout)
    for oh in range(OH):
        for ow in range(OW):
            for kh in range(7):
                slab = flat[:, 2 * oh + kh, 2 * ow * 4: 2 * ow * 4 + 32]       # x_ld = 4 floats per pixel, Cin = 32
                out[:, oh, ow] += slab @ wk[kh]
    out += pw.bias.double()
    s = sd['b.weight'].double() / (sd['b.running_var'].double() + 1e-5).sqrt()
    ref = F.conv2d(x, sd['c.weight'].double(), stride=2, padding=3) * s.view(1, -1, 1, 1) + \
        (sd['b.bias'].double() - sd['b.running_mean'].double() * s).view(1, -1, 1, 1)
    assert float((out.permute(0, 3, 1, 2) - ref).abs().max()) < 1e-5


def test_group_gemms_reproduce_grouped_conv_over_group_major_columns():
    """tt_dcn_im2col writes columns [group][tap][Cin_g]; with zero offsets they are a plain unfold."""
    g = torch.Generator().manual_seed(2)
    N, C, H, W, G = 1, 16, 6, 7, 4
    x = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    sd = {'d.weight': torch.randn(C, C // G, 3, 3, generator=g)}
    packs = Packer(sd, CPU, tc_mode=0).conv_group_gemms('d', groups=G)
    cols = F.unfold(x, 3, padding=1).view(N, G, C // G, 9, H * W).permute(0, 4, 1, 3, 2)   # [n][pix][g][tap][c]
    outs = [cols[:, :, i].reshape(N, H * W, -1) @ p.w.double() for i, p in enumerate(packs)]
    got = torch.cat(outs, -1).permute(0, 2, 1).reshape(N, C, H, W)
    ref = F.conv2d(x, sd['d.weight'].double(), padding=1, groups=G)
    assert all(p.Cin == 9 * C // G and p.Cout == C // G for p in packs)
    assert float((got - ref).abs().max()) < 1e-5


def test_cout_pad_appends_zero_channels_and_tc_planes_follow_the_kernel_layout():
    g = torch.Generator().manual_seed(3)
    sd = {'c.weight': torch.randn(18, 8, 3, 3, generator=g), 'c.bias': torch.randn(18, generator=g)}
    pw = Packer(sd, CPU, tc_mode=3).conv('c', cout_pad=32)
    assert pw.Cout == 32 and pw.w.shape == (72, 32) and pw.bias.shape == (32,)
    assert float(pw.w[:, 18:].abs().max()) == 0 and float(pw.bias[18:].abs().max()) == 0
    hi, lo = pw.w_tc[0], pw.w_tc[1]                                        # [Cout][taps][Cin]
    assert hi.shape == (32, 9, 8)
    assert torch.allclose((hi + lo)[:18], sd['c.weight'].permute(0, 2, 3, 1).reshape(18, 9, 8), rtol=0, atol=1e-6)
    # SIMT layout [tap * Cin + c][co] agrees with the tensor-core planes
    assert torch.allclose(pw.w.view(9, 8, 32).permute(2, 0, 1), hi + lo, rtol=0, atol=1e-6)


def test_spconv_pack_has_tap_major_k_and_tc_planes():
    g = torch.Generator().manual_seed(4)
    Cout, Cin = 32, 32
    sd = {'s.weight': torch.randn(Cout, 3, 3, 3, Cin, generator=g), 'n.weight': torch.ones(Cout), 'n.bias': torch.zeros(Cout),
          'n.running_mean': torch.zeros(Cout), 'n.running_var': torch.ones(Cout)}
    pw, k = Packer(sd, CPU, tc_mode=3).spconv('s', 'n', eps=0.0)
    assert k == (3, 3, 3) and pw.w.shape == (27 * Cin, Cout) and pw.w_tc.shape == (2, Cout, 27, Cin)
    w = sd['s.weight'].reshape(Cout, 27, Cin)
    assert torch.allclose(pw.w.view(27, Cin, Cout).permute(2, 0, 1), w, atol=1e-6)
    assert torch.allclose(pw.w_tc[0] + pw.w_tc[1], w, atol=1e-6)
