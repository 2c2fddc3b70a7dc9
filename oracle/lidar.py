"""Oracle: LiDAR -> BEV encoder.  TEST INFRASTRUCTURE ONLY.

Restates open_loop_training/code/model_code/backbones/lidarnet.py:24-96 and the
third-party modules its config (configs/thinktwice.py:159-193) instantiates:
mmcv hard Voxelization, mmdet3d HardSimpleVFE, SparseEncoder (spconv SubMConv3d /
SparseConv3d, restated as explicit rulebook gather-matmul-scatter), SECOND and
SECONDFPN.  Sparse conv weights use the spconv-2.x layout (Cout, kd, kh, kw, Cin).
"""
import itertools

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# mmcv.ops.Voxelization (hard, deterministic CPU semantics) — called via
# MVXTwoStageDetector.voxelize at lidarnet.py:88
# ----------------------------------------------------------------------------
def grid_size_of(voxel_size, pc_range):
    vs = np.asarray(voxel_size, np.float32)
    r = np.asarray(pc_range, np.float32)
    return np.round((r[3:] - r[:3]) / vs).astype(np.int64)          # (x, y, z) counts


def hard_voxelize(points, voxel_size, pc_range, max_points, max_voxels):
    """points (N, F) f32 -> voxels (M, max_points, F), coors (M, 3) [z, y, x] int32, num_points (M,).

    First-seen voxel order, first `max_points` points per voxel by input order, voxels past
    `max_voxels` dropped.  Vectorised restatement of mmcv's hard_voxelize CPU loop.
    """
    pts = points.detach().cpu()
    N, Fd = pts.shape
    vs = torch.tensor(voxel_size, dtype=torch.float32)
    lo = torch.tensor(pc_range[:3], dtype=torch.float32)
    gs = torch.from_numpy(grid_size_of(voxel_size, pc_range))
    c = torch.floor((pts[:, :3] - lo) / vs).long()                  # (x, y, z)
    ok = ((c >= 0) & (c < gs)).all(1)
    idx = torch.nonzero(ok).squeeze(1)
    c = c[idx]
    key = (c[:, 2] * gs[1] + c[:, 1]) * gs[0] + c[:, 0]
    uniq, inv = torch.unique(key, return_inverse=True)
    # first-seen order of voxels
    first = torch.full((uniq.numel(),), N, dtype=torch.long).scatter_reduce_(0, inv, idx, 'amin')
    order = torch.argsort(first)
    rank_of = torch.empty_like(order)
    rank_of[order] = torch.arange(order.numel())
    vid = rank_of[inv]                                              # voxel id per kept point
    M = min(int(order.numel()), max_voxels)
    keep = vid < M
    idx, vid, c = idx[keep], vid[keep], c[keep]
    # slot of each point inside its voxel = number of earlier points of that voxel (stable sort by voxel)
    o = torch.argsort(vid, stable=True)
    sv = vid[o]
    start = torch.zeros(M + 1, dtype=torch.long)
    start[1:] = torch.cumsum(torch.bincount(sv, minlength=M), 0)
    slot = torch.arange(sv.numel()) - start[sv]
    sel = slot < max_points
    voxels = pts.new_zeros(M, max_points, Fd)
    voxels[sv[sel], slot[sel]] = pts[idx[o][sel]]
    num = torch.bincount(sv, minlength=M).clamp(max=max_points).int()
    coors = torch.zeros(M, 3, dtype=torch.int32)
    coors[vid] = torch.stack([c[:, 2], c[:, 1], c[:, 0]], 1).int()
    return voxels, coors, num


def hard_voxelize_loop(points, voxel_size, pc_range, max_points, max_voxels):
    """Literal per-point loop (mmcv hard_voxelize_kernel); small cases only — pins hard_voxelize()."""
    pts = np.asarray(points, np.float32)
    gs = grid_size_of(voxel_size, pc_range)
    vs = np.asarray(voxel_size, np.float32)
    lo = np.asarray(pc_range[:3], np.float32)
    table, voxels, coors, num = {}, [], [], []
    for p in pts:
        c = np.floor((p[:3] - lo) / vs).astype(np.int64)
        if (c < 0).any() or (c >= gs).any():
            continue
        k = (int(c[2]), int(c[1]), int(c[0]))
        v = table.get(k, -1)
        if v == -1:
            if len(voxels) >= max_voxels:
                continue
            v = len(voxels)
            table[k] = v
            voxels.append(np.zeros((max_points, pts.shape[1]), np.float32))
            coors.append(k)
            num.append(0)
        if num[v] < max_points:
            voxels[v][num[v]] = p
            num[v] += 1
    return (torch.from_numpy(np.stack(voxels)), torch.tensor(coors, dtype=torch.int32),
            torch.tensor(num, dtype=torch.int32))


# ----------------------------------------------------------------------------
# spconv restated: explicit rulebooks
# ----------------------------------------------------------------------------
class SparseTensor:
    def __init__(self, feats, coords, shape, batch_size):
        self.feats, self.coords, self.shape, self.batch_size = feats, coords.long(), tuple(shape), batch_size

    def keys(self, coords=None, shape=None):
        c = self.coords if coords is None else coords
        D, H, W = self.shape if shape is None else shape
        return ((c[:, 0] * D + c[:, 1]) * H + c[:, 2]) * W + c[:, 3]

    def dense(self):
        D, H, W = self.shape
        out = self.feats.new_zeros(self.batch_size, D, H, W, self.feats.shape[1])
        c = self.coords
        out[c[:, 0], c[:, 1], c[:, 2], c[:, 3]] = self.feats
        return out.permute(0, 4, 1, 2, 3).contiguous()


def _triple(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


class SparseConvBase(nn.Module):
    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, subm=False):
        super().__init__()
        self.k, self.s, self.p, self.subm = _triple(kernel_size), _triple(stride), _triple(padding), subm
        self.weight = nn.Parameter(torch.empty(cout, *self.k, cin))
        nn.init.kaiming_uniform_(self.weight.view(cout, -1), a=5 ** 0.5)

    def forward(self, x: SparseTensor):
        dev = x.coords.device
        k, s, p = (torch.tensor(v, device=dev) for v in (self.k, self.s, self.p))
        in_shape = torch.tensor(x.shape, device=dev)
        if self.subm:
            out_shape, out_coords = x.shape, x.coords
        else:
            out_shape = tuple((x.shape[i] + 2 * self.p[i] - self.k[i]) // self.s[i] + 1 for i in range(3))
            oshape = torch.tensor(out_shape, device=dev)
            cand = []
            for tap in itertools.product(*(range(v) for v in self.k)):
                num = x.coords[:, 1:] + p - torch.tensor(tap, device=dev)
                o = torch.div(num, s, rounding_mode='floor')
                ok = ((num % s) == 0).all(1) & (o >= 0).all(1) & (o < oshape).all(1)
                cand.append(torch.cat([x.coords[ok, :1], o[ok]], 1))
            cand = torch.cat(cand, 0)
            okeys = torch.unique(x.keys(cand, out_shape))
            D, H, W = out_shape
            out_coords = torch.stack([okeys // (D * H * W), (okeys // (H * W)) % D, (okeys // W) % H, okeys % W], 1)
        in_keys = x.keys()
        sk, so = torch.sort(in_keys)
        out = x.feats.new_zeros(out_coords.shape[0], self.weight.shape[0])
        for tap in itertools.product(*(range(v) for v in self.k)):
            src = out_coords[:, 1:] * s - p + torch.tensor(tap, device=dev)
            ok = (src >= 0).all(1) & (src < in_shape).all(1)
            qk = x.keys(torch.cat([out_coords[:, :1], src], 1))
            pos = torch.searchsorted(sk, qk).clamp(max=max(sk.numel() - 1, 0))
            hit = ok & (sk[pos] == qk) if sk.numel() else torch.zeros_like(ok)
            oi = torch.nonzero(hit).squeeze(1)
            if oi.numel():
                w = self.weight[:, tap[0], tap[1], tap[2], :]          # (Cout, Cin)
                out.index_add_(0, oi, x.feats[so[pos[oi]]] @ w.t())
        return SparseTensor(out, out_coords, out_shape, x.batch_size)


class SparseSeq(nn.Module):
    """make_sparse_convmodule(order=(conv, norm, act)): {0: conv, 1: BN1d(eps 1e-3)} + ReLU."""
    def __init__(self, conv, cout):
        super().__init__()
        self.add_module('0', conv)
        self.add_module('1', nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01))

    def forward(self, x):
        y = getattr(self, '0')(x)
        y.feats = F.relu(getattr(self, '1')(y.feats))
        return y


class SparseBasicBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = SparseConvBase(c, c, 3, padding=1, subm=True)
        self.bn1 = nn.BatchNorm1d(c, eps=1e-3, momentum=0.01)
        self.conv2 = SparseConvBase(c, c, 3, padding=1, subm=True)
        self.bn2 = nn.BatchNorm1d(c, eps=1e-3, momentum=0.01)

    def forward(self, x):
        y = self.conv1(x)
        y.feats = F.relu(self.bn1(y.feats))
        y = self.conv2(y)
        y.feats = F.relu(self.bn2(y.feats) + x.feats)
        return y


class SparseEncoder(nn.Module):
    """mmdet3d SparseEncoder(block_type='basicblock') as configured at thinktwice.py:167-176;
    forward = lidarnet.py:28-58."""
    def __init__(self, in_channels, sparse_shape, output_channels, encoder_channels, encoder_paddings,
                 base_channels=16, **_unused):
        super().__init__()
        self.sparse_shape = tuple(sparse_shape)
        self.conv_input = SparseSeq(SparseConvBase(in_channels, base_channels, 3, padding=1, subm=True), base_channels)
        self.encoder_layers = nn.Module()
        cin = base_channels
        n_stage = len(encoder_channels)
        for i, blocks in enumerate(encoder_channels):
            stage = []
            for j, cout in enumerate(blocks):
                pad = encoder_paddings[i][j]
                if j == len(blocks) - 1 and i != n_stage - 1:
                    stage.append(SparseSeq(SparseConvBase(cin, cout, 3, stride=2, padding=pad), cout))
                else:
                    stage.append(SparseBasicBlock(cout))
                cin = cout
            self.encoder_layers.add_module(f'encoder_layer{i + 1}', nn.Sequential(*stage))
        self.conv_out = SparseSeq(SparseConvBase(cin, output_channels, (3, 1, 1), stride=(2, 1, 1), padding=0),
                                  output_channels)

    def forward(self, voxel_features, coors, batch_size):
        x = SparseTensor(voxel_features, coors, self.sparse_shape, batch_size)
        x = self.conv_input(x)
        for stage in self.encoder_layers.children():
            x = stage(x)
        d = self.conv_out(x).dense()
        N, C, D, H, W = d.shape
        return d.view(N, C * D, H, W)


class SECOND(nn.Module):                                           # cfg thinktwice.py:177-184
    def __init__(self, in_channels=256, out_channels=(128, 256), layer_nums=(5, 5), layer_strides=(1, 2), **_u):
        super().__init__()
        blocks, cin = [], in_channels
        for cout, n, s in zip(out_channels, layer_nums, layer_strides):
            seq = [nn.Conv2d(cin, cout, 3, stride=s, padding=1, bias=False),
                   nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01), nn.ReLU()]
            for _ in range(n):
                seq += [nn.Conv2d(cout, cout, 3, padding=1, bias=False),
                        nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01), nn.ReLU()]
            blocks.append(nn.Sequential(*seq))
            cin = cout
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        outs = []
        for b in self.blocks:
            x = b(x)
            outs.append(x)
        return outs


class SECONDFPN(nn.Module):                                        # cfg thinktwice.py:185-192
    def __init__(self, in_channels=(128, 256), out_channels=(256, 256), upsample_strides=(1, 2), **_u):
        super().__init__()
        de = []
        for cin, cout, s in zip(in_channels, out_channels, upsample_strides):
            if s > 1:
                up = nn.ConvTranspose2d(cin, cout, s, stride=s, bias=False)
            else:
                up = nn.Conv2d(cin, cout, 1, stride=1, bias=False)
            de.append(nn.Sequential(up, nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01), nn.ReLU()))
        self.deblocks = nn.ModuleList(de)

    def forward(self, xs):
        return [torch.cat([d(x) for d, x in zip(self.deblocks, xs)], 1)]


class LidarNet(nn.Module):
    """lidarnet.py:61-96 (MVXTwoStageDetector reduced to the members the forward touches)."""
    def __init__(self, pts_voxel_layer, pts_voxel_encoder, pts_middle_encoder, pts_backbone, pts_neck, **_unused):
        super().__init__()
        self.vcfg = dict(pts_voxel_layer)
        self.num_features = pts_voxel_encoder['num_features']
        self.pts_middle_encoder = SparseEncoder(**{k: v for k, v in pts_middle_encoder.items() if k != 'type'})
        self.pts_backbone = SECOND(**{k: v for k, v in pts_backbone.items() if k != 'type'})
        self.pts_neck = SECONDFPN(**{k: v for k, v in pts_neck.items() if k != 'type'})

    def voxelize(self, pts):                                       # MVXTwoStageDetector.voxelize
        mv = self.vcfg['max_voxels']
        mv = mv[0] if self.training else mv[1]
        vs, cs, ns = [], [], []
        for i, res in enumerate(pts):
            v, c, n = hard_voxelize(res, self.vcfg['voxel_size'], self.vcfg['point_cloud_range'],
                                    self.vcfg['max_num_points'], mv)
            vs.append(v); ns.append(n); cs.append(F.pad(c, (1, 0), value=i))
        return torch.cat(vs), torch.cat(ns), torch.cat(cs)

    def forward(self, pts, keep=None):
        dev = pts.device
        voxels, num_points, coors = self.voxelize(pts)
        voxels, num_points, coors = voxels.to(dev), num_points.to(dev), coors.to(dev)
        feats = voxels[:, :, :self.num_features].sum(1) / num_points.type_as(voxels).view(-1, 1)   # HardSimpleVFE
        batch_size = int(coors[-1, 0]) + 1
        # spconv's behaviour for z >= sparse_shape[0] is undefined; the oracle drops such voxels
        ok = coors[:, 1] < self.pts_middle_encoder.sparse_shape[0]
        x = self.pts_middle_encoder(feats[ok], coors[ok], batch_size)
        if keep is not None:
            keep.update(voxel_feats=feats, coors=coors, sparse_out=x)
        return self.pts_neck(self.pts_backbone(x))
