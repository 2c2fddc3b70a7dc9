"""Oracle: camera -> BEV encoder (LSS).  TEST INFRASTRUCTURE ONLY.

Restates open_loop_training/code/model_code/backbones/lss.py (whole file) plus
the mmdet / mmcv pieces it builds from config (ResNet-50, PAFPN, BasicBlock,
DeformConv2dPack).  state_dict keys equal the reference's.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.ops import deform_conv2d

from .voxel_pool import voxel_pooling_ref


# ----------------------------------------------------------------------------
# mmdet ResNet-50 (style='pytorch', out_indices 0-3) — cfg thinktwice.py:141-148
# ----------------------------------------------------------------------------
class Bottleneck(nn.Module):
    def __init__(self, cin, planes, stride, down):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if down:
            self.downsample = nn.Sequential(
                nn.Conv2d(cin, planes * 4, 1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * 4))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + idt)


class ResNet50(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, n, stride) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]):
            blocks = []
            for b in range(n):
                blocks.append(Bottleneck(cin, planes, stride if b == 0 else 1, b == 0))
                cin = planes * 4
            setattr(self, f'layer{li + 1}', nn.Sequential(*blocks))

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.max_pool2d(x, 3, stride=2, padding=1)
        outs = []
        for li in range(4):
            x = getattr(self, f'layer{li + 1}')(x)
            outs.append(x)
        return outs


class ConvModule(nn.Module):
    """mmcv ConvModule with norm_cfg=None, act_cfg=None: just `.conv`."""
    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding)

    def forward(self, x):
        return self.conv(x)


class PAFPN(nn.Module):
    """mmdet PAFPN as re-stated in-tree at lss.py:286-348 (4 in, 4 out, no extra levels)."""
    def __init__(self, in_channels=(256, 512, 1024, 2048), out_channels=256):
        super().__init__()
        n = len(in_channels)
        self.lateral_convs = nn.ModuleList([ConvModule(c, out_channels, 1) for c in in_channels])
        self.fpn_convs = nn.ModuleList([ConvModule(out_channels, out_channels, 3, padding=1) for _ in range(n)])
        self.downsample_convs = nn.ModuleList(
            [ConvModule(out_channels, out_channels, 3, stride=2, padding=1) for _ in range(n - 1)])
        self.pafpn_convs = nn.ModuleList(
            [ConvModule(out_channels, out_channels, 3, padding=1) for _ in range(n - 1)])

    def forward(self, inputs):
        lat = [l(x) for l, x in zip(self.lateral_convs, inputs)]
        for i in range(len(lat) - 1, 0, -1):                       # lss.py:301-305 top-down, nearest
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
        inter = [c(x) for c, x in zip(self.fpn_convs, lat)]        # lss.py:309-311
        for i in range(len(inter) - 1):                            # lss.py:314-315 bottom-up
            inter[i + 1] = inter[i + 1] + self.downsample_convs[i](inter[i])
        outs = [inter[0]] + [self.pafpn_convs[i - 1](inter[i]) for i in range(1, len(inter))]
        return tuple(outs)


class BasicBlock(nn.Module):
    """mmdet.models.backbones.resnet.BasicBlock(c, c) — lss.py:185-187."""
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + x)


class DeformConv2dPack(nn.Module):
    """mmcv DCNv1 pack (lss.py:189-197): offsets from a zero-init 3x3 conv, grouped weight, no bias."""
    def __init__(self, cin, cout, groups=4):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, 3, 3))
        self.conv_offset = nn.Conv2d(cin, 18, 3, padding=1, bias=True)
        nn.init.kaiming_uniform_(self.weight, nonlinearity='relu')
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        return deform_conv2d(x, self.conv_offset(x), self.weight, padding=1)


class ASPPModule(nn.Module):                                       # lss.py:20-46
    def __init__(self, cin, planes, k, padding, dilation):
        super().__init__()
        self.atrous_conv = nn.Conv2d(cin, planes, k, padding=padding, dilation=dilation, bias=False)
        self.bn = nn.BatchNorm2d(planes)

    def forward(self, x):
        return F.relu(self.bn(self.atrous_conv(x)))


class ASPP(nn.Module):                                             # lss.py:49-118
    def __init__(self, cin, mid):
        super().__init__()
        self.aspp1 = ASPPModule(cin, mid, 1, 0, 1)
        self.aspp2 = ASPPModule(cin, mid, 3, 6, 6)
        self.aspp3 = ASPPModule(cin, mid, 3, 12, 12)
        self.aspp4 = ASPPModule(cin, mid, 3, 18, 18)
        self.global_avg_pool = nn.Sequential(
            nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(cin, mid, 1, bias=False), nn.BatchNorm2d(mid), nn.ReLU())
        self.conv1 = nn.Conv2d(mid * 5, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)

    def forward(self, x):
        x5 = self.global_avg_pool(x)
        x5 = F.interpolate(x5, size=x.shape[2:], mode='bilinear', align_corners=True)
        y = torch.cat((self.aspp1(x), self.aspp2(x), self.aspp3(x), self.aspp4(x), x5), 1)
        return F.relu(self.bn1(self.conv1(y)))                     # dropout(0.5) is identity in eval


class Mlp(nn.Module):                                              # lss.py:121-143
    def __init__(self, cin, hid, cout):
        super().__init__()
        self.fc1 = nn.Linear(cin, hid)
        self.fc2 = nn.Linear(hid, cout)

    def forward(self, x):
        return self.fc2(F.relu(self.fc1(x)))


class SELayer(nn.Module):                                          # lss.py:146-158
    def __init__(self, c):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, c, 1)
        self.conv_expand = nn.Conv2d(c, c, 1)

    def forward(self, x, x_se):
        return x * torch.sigmoid(self.conv_expand(F.relu(self.conv_reduce(x_se))))


class DepthNet(nn.Module):                                         # lss.py:161-240
    def __init__(self, cin, mid, context_channels, depth_channels):
        super().__init__()
        self.reduce_conv = nn.Sequential(nn.Conv2d(cin, mid, 3, padding=1), nn.BatchNorm2d(mid), nn.ReLU())
        self.context_conv = nn.Conv2d(mid, context_channels, 1)
        self.bn = nn.BatchNorm1d(22)
        self.depth_mlp = Mlp(22, mid, mid)
        self.depth_se = SELayer(mid)
        self.context_mlp = Mlp(22, mid, mid)
        self.context_se = SELayer(mid)
        self.depth_conv = nn.Sequential(
            BasicBlock(mid), BasicBlock(mid), BasicBlock(mid), ASPP(mid, mid),
            DeformConv2dPack(mid, mid, groups=4), nn.Conv2d(mid, depth_channels, 1))

    @staticmethod
    def mlp_input(mats):                                           # lss.py:206-231
        intr = mats['intrin_mats'][:, -1:, ..., :3, :3]
        B, N = intr.shape[0], intr.shape[2]
        ida = mats['ida_mats'][:, -1:]
        s2e = mats['sensor2ego_mats'][:, -1:, ..., :3, :]
        v = torch.stack([intr[:, 0:1, ..., 0, 0], intr[:, 0:1, ..., 1, 1], intr[:, 0:1, ..., 0, 2],
                         intr[:, 0:1, ..., 1, 2], ida[:, 0:1, ..., 0, 0], ida[:, 0:1, ..., 0, 1],
                         ida[:, 0:1, ..., 0, 3], ida[:, 0:1, ..., 1, 0], ida[:, 0:1, ..., 1, 1],
                         ida[:, 0:1, ..., 1, 3]], -1)
        v = torch.cat([v, s2e.reshape(B, 1, N, -1)], -1)
        return v.reshape(-1, v.shape[-1])                          # (B*N, 22)

    def forward(self, x, mats):
        m = self.bn(self.mlp_input(mats))
        x = self.reduce_conv(x)
        context = self.context_conv(self.context_se(x, self.context_mlp(m)[..., None, None]))
        depth = self.depth_conv(self.depth_se(x, self.depth_mlp(m)[..., None, None]))
        return torch.cat([depth, context], 1)


class UnetLayer(nn.Module):                                        # lss.py:243-258
    def __init__(self, cin, cmid, cout):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cout, 2, stride=2)
        self.conv_relu = nn.Sequential(nn.Conv2d(cmid, cout, 3, padding=1), nn.ReLU())

    def forward(self, x1, x2):
        return self.conv_relu(torch.cat((self.up(x1), x2), 1))


class UNet(nn.Module):                                             # lss.py:260-282
    def __init__(self, n_class, fpn_ch):
        super().__init__()
        self.unet_layer4 = UnetLayer(fpn_ch[3], 256 + fpn_ch[2], 256)
        self.unet_layer3 = UnetLayer(256, 256 + fpn_ch[1], 256)
        self.unet_layer2 = UnetLayer(256, 128 + fpn_ch[0], 128)
        self.unet_layer0 = nn.Sequential(
            nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
            nn.Conv2d(128, 64, 3, padding=1, bias=False), nn.ReLU(),
            nn.Conv2d(64, 64, 3, padding=1, bias=False))
        self.conv_last = nn.Conv2d(64, n_class, 1)

    def forward(self, feats):
        e1, e2, e3, e4 = feats
        d = self.unet_layer4(e4, e3)
        d = self.unet_layer3(d, e2)
        d = self.unet_layer2(d, e1)
        return self.conv_last(self.unet_layer0(d))


def seg_to_feat_stack(cin):                                        # lss.py:409-438
    spec = [(cin, 64, 1, 1), (64, 16, 1, 1), (16, 32, 3, 2), (32, 32, 1, 1),
            (32, 64, 3, 2), (64, 64, 1, 1), (64, 128, 3, 2)]
    layers = []
    for ci, co, k, s in spec:
        layers += [nn.Conv2d(ci, co, k, padding=k // 2, stride=s), nn.BatchNorm2d(co), nn.ReLU()]
    return nn.Sequential(*layers)


class LSS(nn.Module):
    """lss.py:351-724."""
    def __init__(self, x_bound, y_bound, z_bound, d_bound, final_dim, downsample_factor, output_channels,
                 depth_net_conf, seg_net_conf, queue_len=1, fpn_in_channels=(256, 256, 256, 256), **_unused):
        super().__init__()
        self.downsample_factor = downsample_factor
        self.d_bound = d_bound
        self.final_dim = final_dim
        self.output_channels = output_channels
        self.queue_len = queue_len
        if queue_len != 1:
            self.bev_multiframe_merge = nn.Conv2d(256 * queue_len, 256, 3, padding=1, bias=False)
        rows = [x_bound, y_bound, z_bound]
        self.register_buffer('voxel_size', torch.Tensor([r[2] for r in rows]))
        self.register_buffer('voxel_coord', torch.Tensor([r[0] + r[2] / 2.0 for r in rows]))
        self.register_buffer('voxel_num', torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows]))
        self.register_buffer('frustum', self.create_frustum())
        self.depth_channels = self.frustum.shape[0]
        self.img_backbone = ResNet50()
        self.img_neck = PAFPN()
        self.neck_conv = nn.Conv2d(256, depth_net_conf['in_channels'], 1)
        self.depth_net = DepthNet(depth_net_conf['in_channels'], depth_net_conf['mid_channels'],
                                  output_channels, self.depth_channels)
        self.seg_net = UNet(seg_net_conf['out_channels'], list(fpn_in_channels))
        self.seg_res_to_image_feature = seg_to_feat_stack(seg_net_conf['out_channels'])
        self.merge_seg_and_image = nn.Conv2d(256 + 128, 256, 3, padding=1)

    def create_frustum(self):                                      # lss.py:454-471
        H, W = self.final_dim
        fH, fW = H // self.downsample_factor, W // self.downsample_factor
        d = torch.arange(*self.d_bound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
        D = d.shape[0]
        x = torch.linspace(0, W - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
        y = torch.linspace(0, H - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
        return torch.stack((x, y, d, torch.ones_like(d)), -1)

    def get_geometry(self, sensor2ego, intrin, ida):               # lss.py:474-512 (bda_mat is None)
        B, N = sensor2ego.shape[:2]
        p = self.frustum
        p = ida.view(B, N, 1, 1, 1, 4, 4).inverse().matmul(p.unsqueeze(-1))
        p = torch.cat((p[..., :2, :] * p[..., 2:3, :], p[..., 2:, :]), 5)
        comb = sensor2ego.matmul(torch.inverse(intrin))
        p = comb.view(B, N, 1, 1, 1, 4, 4).matmul(p).squeeze(-1)
        return p[..., :3]

    def geom_index(self, geom):                                    # lss.py:630-631 (trunc toward zero)
        return ((geom - (self.voxel_coord - self.voxel_size / 2.0)) / self.voxel_size).int()

    def single_sweep(self, sweep_index, imgs, mats, keep=None):    # lss.py:542-621
        B, _, N = imgs.shape[:3]
        x = imgs.reshape(B * N, *imgs.shape[3:])
        fpn = self.img_neck(self.img_backbone(x))
        src = self.neck_conv(fpn[2])
        df = self.depth_net(src, mats)
        depth = df[:, :self.depth_channels]
        prob = depth.softmax(1)
        feat = df[:, self.depth_channels:self.depth_channels + self.output_channels]
        seg = self.seg_net(fpn)
        feat = self.merge_seg_and_image(torch.cat((feat, self.seg_res_to_image_feature(seg.detach())), 1))
        lifted = prob.unsqueeze(1) * feat.unsqueeze(2)             # (BN, C, D, H, W) — the 514 MB tensor
        lifted = lifted.reshape(B, N, *lifted.shape[1:]).permute(0, 1, 3, 4, 5, 2).contiguous()
        geom = self.get_geometry(mats['sensor2ego_mats'][:, sweep_index], mats['intrin_mats'][:, sweep_index],
                                 mats['ida_mats'][:, sweep_index])
        bev = voxel_pooling_ref(self.geom_index(geom).contiguous(), lifted, self.voxel_num)
        if keep is not None:
            keep.update(depth_prob=prob, img_feature=feat, geom=geom, img_feats=src, depth_feature=df)
        return dict(fpn_feats=fpn, bev=bev.contiguous(), depth=depth, seg=seg)

    @staticmethod
    def build_mats(img_metas, num_cams):                           # lss.py:667-687
        intr, ida, s2e = [], [], []
        for b in range(len(img_metas)):
            _i, _a, _s = [], [], []
            for t in range(len(img_metas[0])):
                m = img_metas[b][t]
                k = torch.zeros((num_cams, 4, 4))
                k[:, :3, :3] = m['cam_intrinsic']
                k[:, 3, 3] = 1
                _i.append(k)
                _a.append(m['ida_mats'])
                _s.append(m['currlidar2keycam'].permute(0, 2, 1))  # quirk: transpose, not inverse (lss.py:677)
            intr.append(torch.stack(_i)); ida.append(torch.stack(_a)); s2e.append(torch.stack(_s))
        return dict(intrin_mats=torch.stack(intr), ida_mats=torch.stack(ida), sensor2ego_mats=torch.stack(s2e))

    def forward(self, img, img_metas, keep=None):                  # lss.py:635-724
        if img.dim() == 5:
            img = img.unsqueeze(1)
        T, N = img.shape[1], img.shape[2]
        mats = {k: v.to(img.device) for k, v in self.build_mats(img_metas, N).items()}
        key = self.single_sweep(-1, img[:, -1:], mats, keep)
        outs = dict(seg=key['seg'], depth=key['depth'], fpn_feats=key['fpn_feats'])
        outs['lidar2img'] = torch.stack([m[-1]['lidar2img'] for m in img_metas], 0)
        outs['ida_mat'] = mats['ida_mats'][:, -1].clone()
        bevs = [key['bev']]
        for s in range(1, T):                                      # quirk: mats index -s (lss.py:712-716)
            bevs.append(self.single_sweep(-s, img[:, -(s + 1):img.shape[1] - s], mats)['bev'])
        assert len(bevs) == self.queue_len
        bev = torch.cat(bevs, 1)
        if self.queue_len > 1:
            bev = self.bev_multiframe_merge(bev)
        outs['bev'] = bev
        return outs
