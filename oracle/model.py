"""Oracle: EncoderDecoder framework (fusion, pyramid, orchestration).  TEST INFRASTRUCTURE ONLY.

Restates open_loop_training/code/encoder_decoder_framework.py:25-138,194-250 and
code/utils.py:84-121 (SEModule / SEBasicBlock).  `forward_inference(batch)` has the
reference signature and returns the reference's pred dict.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .camera import LSS
from .decoder import ThinkTwiceDecoder
from .lidar import LidarNet


def anti_transpose(x):
    """torch.rot90(torch.flip(x, [2]), 1, [2, 3]) — framework:241,246."""
    return torch.rot90(torch.flip(x, dims=[2]), 1, dims=[2, 3])


class SEModule(nn.Module):                                         # code/utils.py:84-96
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Conv2d(c, c, 1)
        self.fc2 = nn.Conv2d(c, c, 1)

    def forward(self, x):
        s = 0.5 * x.mean((2, 3), keepdim=True) + 0.5 * x.amax((2, 3), keepdim=True)
        return x * self.fc2(F.relu(self.fc1(s))).sigmoid()


class SEBasicBlock(nn.Module):                                     # code/utils.py:99-121
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c * 2, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(c * 2)
        self.conv2 = nn.Conv2d(c * 2, c, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(c)
        self.se = SEModule(c)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        return F.relu(self.se(y) + x)


def _strip(cfg):
    return {k: v for k, v in cfg.items() if k != 'type'}


class EncoderDecoder(nn.Module):
    def __init__(self, img_encoder, decoder, lidar_encoder, num_cams=4, train_cfg=None, **_unused):
        super().__init__()
        self.config = train_cfg
        self.num_cams = num_cams
        self.img_encoder = LSS(**_strip(img_encoder))
        self.lidar_encoder = LidarNet(**_strip(lidar_encoder))
        def cbr(cin, cout, stride, last_act):
            l = [nn.Conv2d(cin, 256, 3, padding=1, bias=False, stride=stride), nn.BatchNorm2d(256), nn.ReLU(),
                 nn.Conv2d(256, cout, 3, padding=1, bias=False, stride=stride), nn.BatchNorm2d(cout)]
            return nn.Sequential(*(l + ([nn.ReLU()] if last_act else [])))
        self.conv_cam = cbr(256, 256, 1, False)                    # framework:85-91
        self.conv_lidar = cbr(512, 256, 2, True)                   # :94-101
        self.conv_fusion = cbr(512, 256, 1, False)                 # :103-109
        self._256_to_32 = nn.Conv2d(256, 32, 3, padding=1)
        self.MLP21, self.MLP10, self.MLP4, self.MLP2 = SEBasicBlock(32), SEBasicBlock(64), SEBasicBlock(128), SEBasicBlock(256)
        self.conv21_10 = nn.Conv2d(32, 64, 3, stride=2)
        self.conv10_4 = nn.Conv2d(64, 128, 3, stride=2)
        self.conv4_2 = nn.Conv2d(128, 256, 3, stride=1)
        self.output_fc = nn.Sequential(nn.Linear(1024, 512), nn.ReLU(), nn.BatchNorm1d(512), nn.Linear(512, 256), nn.ReLU())
        self.measurements_encoder = nn.Sequential(nn.Linear(9, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU())
        self.decoder = ThinkTwiceDecoder(**_strip(decoder))

    def get_fusion_feat(self, cam_bev, lidar_feat):                # framework:213-235
        lidar_hi = lidar_feat.clone()
        cam = F.relu(self.conv_cam(cam_bev) + cam_bev)
        pts = self.conv_lidar(lidar_feat)
        bev = F.relu(self.conv_fusion(torch.cat([cam, pts], 1)) + cam + pts)
        f21 = self.MLP21(F.relu(self._256_to_32(bev)))
        f10 = self.MLP10(F.relu(self.conv21_10(f21)))
        f4 = self.MLP4(F.relu(self.conv10_4(f10)))
        f2 = self.MLP2(F.relu(self.conv4_2(f4)))
        return self.output_fc(f2.flatten(1)), f21, [None, None, f21, f10, f4, f2], lidar_hi

    def forward_inference(self, batch, dead_work=False, keep=None):  # framework:194-210, 238-250
        target_point = batch['target_point'].to(torch.float32)
        speed = batch['speed'].to(torch.float32).view(-1, 1) / 12.
        state = torch.cat([speed, target_point, batch['target_command']], -1)
        kc = {} if keep is not None else None
        cam = self.img_encoder(batch['img'], batch['img_metas'], kc)
        cam['bev'] = anti_transpose(cam['bev'])
        meas = self.measurements_encoder(state)
        kl = {} if keep is not None else None
        lidar = [anti_transpose(x) for x in self.lidar_encoder(batch['points'][:, -1], kl)]
        flat, bev32, mid, lidar_hi = self.get_fusion_feat(cam['bev'], lidar[0])
        kd = {} if keep is not None else None
        pred = self.decoder(flat, bev32, meas, self, [cam['lidar2img'], cam['ida_mat'], cam['fpn_feats'], lidar_hi],
                            dead_work, kd)
        if keep is not None:
            keep.update(cam=cam, cam_keep=kc, lidar=lidar, lidar_keep=kl, flat=flat, bev32=bev32, mid=mid,
                        meas=meas, dec_keep=kd)
        return pred


# ----------------------------------------------------------------------------
# Weight initialisation + BN calibration for well-conditioned synthetic runs
# (SURVEY.md §8d "Weights").  Not reference behaviour: test fixture.
# ----------------------------------------------------------------------------
def init_oracle_weights(model, seed=0):
    """He-style random init of every parameter from one seeded CPU generator (deterministic order)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda shape, std: torch.randn(tuple(shape), generator=g) * std
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, nn.ConvTranspose2d):
                m.weight.copy_(rn(m.weight.shape, (2.0 / m.weight.shape[0]) ** 0.5))   # k2s2: fan_in = Cin
                if m.bias is not None:
                    m.bias.copy_(rn(m.bias.shape, 0.02))
            elif isinstance(m, (nn.Conv2d, nn.Linear)):
                m.weight.copy_(rn(m.weight.shape, (2.0 / m.weight[0].numel()) ** 0.5))
                if m.bias is not None:
                    m.bias.copy_(rn(m.bias.shape, 0.02))
            elif isinstance(m, nn.LayerNorm):
                m.weight.copy_(1 + rn(m.weight.shape, 0.02))
                m.bias.copy_(rn(m.bias.shape, 0.02))
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.copy_(1 + rn(m.weight.shape, 0.05))
                m.bias.copy_(rn(m.bias.shape, 0.05))
            elif isinstance(getattr(m, 'weight', None), nn.Parameter) and m.weight.dim() == 5:
                w = m.weight                                       # sparse conv (Cout, kd, kh, kw, Cin)
                w.copy_(rn(w.shape, (2.0 / (w[0].numel() / 3.0)) ** 0.5))   # ~1/3 of the taps are occupied
            elif hasattr(m, 'conv_offset') and isinstance(getattr(m, 'weight', None), nn.Parameter):
                m.weight.copy_(rn(m.weight.shape, (2.0 / m.weight[0].numel()) ** 0.5))  # DCN main weight
        for name, p in model.named_parameters():
            if name.endswith('conv_offset.weight'):
                p.copy_(rn(p.shape, 0.01))                         # reference zero-inits; non-zero exercises sampling
            elif name.endswith(('temporal_embedding', 'cams_embeds', 'static_embedding', 'level_embeds')):
                p.copy_(rn(p.shape, 0.02))
        for m in model.modules():
            if hasattr(m, 'sampling_offsets') and hasattr(m, 'init_weights'):
                m.init_weights()                                   # reference grid bias (msda:403-421)
                m.sampling_offsets.weight.copy_(rn(m.sampling_offsets.weight.shape, 0.02))
                m.attention_weights.weight.copy_(rn(m.attention_weights.weight.shape, 0.05))
                m.value_proj.weight.copy_(rn(m.value_proj.weight.shape, (1.0 / 256) ** 0.5))
    return model


@torch.no_grad()
def calibrate_bn(model, batch):
    """One eval-mode pass in which every BatchNorm first sets running stats from its own input."""
    handles = []

    def pre(m, inp):
        x = inp[0]
        dims = [0] + list(range(2, x.dim()))
        n = x.numel() // x.shape[1]
        if n >= 16:
            mean = x.mean(dims)
            var = x.var(dims, unbiased=False)
        else:                                                      # tiny batches: scalar stats keep scale sane
            mean = torch.full((x.shape[1],), float(x.mean()))
            var = torch.full((x.shape[1],), float(x.var(unbiased=False)))
        var = torch.where(var < 1e-6, torch.ones_like(var), var)
        m.running_mean.copy_(mean.to(m.running_mean.device))
        m.running_var.copy_(var.to(m.running_var.device))

    for m in model.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            handles.append(m.register_forward_pre_hook(pre))
    model.eval()
    model.forward_inference(batch)
    for h in handles:
        h.remove()
    return model
