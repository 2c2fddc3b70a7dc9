"""Camera -> BEV encoder on the B200 op library.

Mirror of the reference `LSS` backbone (open_loop_training/code/model_code/backbones/lss.py:351-724):
same constructor arguments, same forward(img, img_metas) contract, same output dict keys.  All tensor
math runs in libtt_b200 (channels-last, BN folded, concat by channel offset, fused lift-splat).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import lib
from .engine import FMap
from .lib import ACT_RELU, LiftSplatDesc, _p
from .registry import BACKBONES
from .weights import bn_affine


@BACKBONES.register_module()
class LSS(nn.Module):
    """nn.Module so that the reference-named parameters ('img_encoder.img_backbone.conv1.weight', ...) live in THIS
    module's tree: mmcv's recursive checkpoint loader (thinktwice_agent.py:170) finds them by walking `_modules`."""

    def __init__(self, x_bound, y_bound, z_bound, d_bound, final_dim, downsample_factor, output_channels,
                 img_backbone_conf=None, img_neck_conf=None, depth_net_conf=None, seg_net_conf=None, queue_len=1,
                 fpn_in_channels=(256, 256, 256, 256), prefix='img_encoder.'):
        super().__init__()
        self.prefix = prefix
        self.d_bound, self.final_dim = list(d_bound), tuple(final_dim)
        self.downsample_factor, self.output_channels, self.queue_len = downsample_factor, output_channels, queue_len
        rows = [x_bound, y_bound, z_bound]
        # buffers exactly as lss.py:386-398 (fp32 arithmetic of the reference)
        self.register_buffer('voxel_size', torch.Tensor([r[2] for r in rows]))
        self.register_buffer('voxel_coord', torch.Tensor([r[0] + r[2] / 2.0 for r in rows]))
        self.register_buffer('voxel_num', torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows]))
        self.fH, self.fW = final_dim[0] // downsample_factor, final_dim[1] // downsample_factor
        self.frustum_d = torch.arange(*d_bound, dtype=torch.float)
        self.frustum_u = torch.linspace(0, final_dim[1] - 1, self.fW, dtype=torch.float)
        self.frustum_v = torch.linspace(0, final_dim[0] - 1, self.fH, dtype=torch.float)
        self.depth_channels = self.frustum_d.numel()
        self.mid = depth_net_conf['mid_channels']
        self.n_seg = seg_net_conf['out_channels']
        self.register_buffer('frustum', self.make_frustum())

    def make_frustum(self):
        D, fH, fW = self.depth_channels, self.fH, self.fW
        return torch.stack((self.frustum_u.view(1, 1, fW).expand(D, fH, fW), self.frustum_v.view(1, fH, 1).expand(D, fH, fW),
                            self.frustum_d.view(D, 1, 1).expand(D, fH, fW), torch.ones(D, fH, fW)), -1)

    # ------------------------------------------------------------------ weight preparation
    def prepare(self, pk, eng):
        self.eng = eng
        p = self.prefix
        w = self.w = {}
        b = p + 'img_backbone.'
        if pk.tc_mode == 4:                                          # f16s engine: 8 halves per pixel, 64-half K slabs
            w['stem'] = pk.conv_rowpacked(b + 'conv1', bn=b + 'bn1', cpad=8, kslab=64)
        else:
            w['stem'] = pk.conv_rowpacked(b + 'conv1', bn=b + 'bn1', cpad=4, kslab=32)   # 7x7x3 s2: 7 row-packed K slabs
        self.blocks = []
        for li, n in enumerate([3, 4, 6, 3]):
            for i in range(n):
                q = f'{b}layer{li + 1}.{i}.'
                blk = dict(c1=pk.conv(q + 'conv1', bn=q + 'bn1'), c2=pk.conv(q + 'conv2', bn=q + 'bn2'),
                           c3=pk.conv(q + 'conv3', bn=q + 'bn3'), stride=2 if (i == 0 and li > 0) else 1,
                           down=pk.conv(q + 'downsample.0', bn=q + 'downsample.1') if i == 0 else None,
                           stage=li, last=(i == n - 1))
                self.blocks.append(blk)
        nk = p + 'img_neck.'
        for grp, cnt in (('lateral_convs', 4), ('fpn_convs', 4), ('downsample_convs', 3), ('pafpn_convs', 3)):
            w[grp] = [pk.conv(f'{nk}{grp}.{i}.conv') for i in range(cnt)]
        w['neck_conv'] = pk.conv(p + 'neck_conv')
        d = p + 'depth_net.'
        w['reduce'] = pk.conv(d + 'reduce_conv.0', bn=d + 'reduce_conv.1')
        w['context_conv'] = pk.conv(d + 'context_conv')
        s22 = bn_affine(pk.sd, d + 'bn', 1e-5)
        for m in ('depth', 'context'):
            w[m + '_fc1'] = pk.linear(f'{d}{m}_mlp.fc1', in_affine=s22, cin_pad=24)
            w[m + '_fc2'] = pk.linear(f'{d}{m}_mlp.fc2')
            w[m + '_red'] = pk.conv1x1_as_linear(f'{d}{m}_se.conv_reduce')
            w[m + '_exp'] = pk.conv1x1_as_linear(f'{d}{m}_se.conv_expand')
        w['bb'] = [(pk.conv(f'{d}depth_conv.{i}.conv1', bn=f'{d}depth_conv.{i}.bn1'),
                    pk.conv(f'{d}depth_conv.{i}.conv2', bn=f'{d}depth_conv.{i}.bn2')) for i in range(3)]
        a = d + 'depth_conv.3.'
        w['aspp'] = [pk.conv(f'{a}aspp{i + 1}.atrous_conv', bn=f'{a}aspp{i + 1}.bn') for i in range(4)]
        w['aspp_gap'] = pk.conv(a + 'global_avg_pool.1', bn=a + 'global_avg_pool.2')
        w['aspp_out'] = pk.conv(a + 'conv1', bn=a + 'bn1')
        w['dcn_off'] = pk.conv(d + 'depth_conv.4.conv_offset', cout_pad=32)   # 18 -> 32 channels: tcgen05 instead of SIMT
        w['dcn'] = pk.conv_group_gemms(d + 'depth_conv.4', groups=4)   # one GEMM per group over the sampled columns
        w['depth_out'] = pk.conv(d + 'depth_conv.5')
        u = p + 'seg_net.'
        for l in (4, 3, 2):
            w[f'u{l}_up'] = pk.convT(f'{u}unet_layer{l}.up')
            w[f'u{l}_conv'] = pk.conv(f'{u}unet_layer{l}.conv_relu.0')
        w['u0_a'] = pk.conv(u + 'unet_layer0.1'); w['u0_b'] = pk.conv(u + 'unet_layer0.3')
        w['seg_last'] = pk.conv(u + 'conv_last', cout_pad=32)          # 12 -> 32 zero-padded channels: tcgen05 instead of SIMT
        self.seg_classes = int(pk.sd[u + 'conv_last.weight'].shape[0])
        r = p + 'seg_res_to_image_feature.'
        w['s2f'] = [pk.conv(f'{r}{3 * i}', bn=f'{r}{3 * i + 1}', cout_pad=32) for i in range(7)]   # s2f.1 is 64 -> 16
        self.s2f_channels = [int(pk.sd[f'{r}{3 * i}.weight'].shape[0]) for i in range(7)]
        w['merge'] = pk.conv(p + 'merge_seg_and_image')
        if self.queue_len != 1:
            w['sweep_merge'] = pk.conv(p + 'bev_multiframe_merge')
        dev = eng.device
        self.fu, self.fv, self.fd = self.frustum_u.to(dev), self.frustum_v.to(dev), self.frustum_d.to(dev)
        # lower bound of the voxel grid exactly as lss.py:630 evaluates it in fp32
        # host copies of the (checkpoint-loadable) geometry buffers: the forward never reads a device scalar back
        self.h_vsize, self.h_vnum = self.voxel_size.float().cpu(), [int(v) for v in self.voxel_num.cpu()]
        self.lower = (self.voxel_coord.float().cpu() - self.h_vsize / 2.0)

    # ------------------------------------------------------------------ host-side matrices (lss.py:667-687, 496-502)
    @staticmethod
    def build_mats(img_metas, num_cams):
        intr, ida, s2e = [], [], []
        for b in range(len(img_metas)):
            _i, _a, _s = [], [], []
            for t in range(len(img_metas[0])):
                m = img_metas[b][t]
                k = torch.zeros((num_cams, 4, 4))
                k[:, :3, :3] = m['cam_intrinsic']
                k[:, 3, 3] = 1
                _i.append(k); _a.append(m['ida_mats'])
                _s.append(m['currlidar2keycam'].permute(0, 2, 1))      # the reference transposes (lss.py:677)
            intr.append(torch.stack(_i)); ida.append(torch.stack(_a)); s2e.append(torch.stack(_s))
        return dict(intrin_mats=torch.stack(intr).float(), ida_mats=torch.stack(ida).float(),
                    sensor2ego_mats=torch.stack(s2e).float())

    @staticmethod
    def depthnet_mlp_input(mats):                                      # lss.py:206-231
        intr = mats['intrin_mats'][:, -1:, ..., :3, :3]
        B, N = intr.shape[0], intr.shape[2]
        ida = mats['ida_mats'][:, -1:]
        s2e = mats['sensor2ego_mats'][:, -1:, ..., :3, :]
        v = torch.stack([intr[:, 0:1, ..., 0, 0], intr[:, 0:1, ..., 1, 1], intr[:, 0:1, ..., 0, 2], intr[:, 0:1, ..., 1, 2],
                         ida[:, 0:1, ..., 0, 0], ida[:, 0:1, ..., 0, 1], ida[:, 0:1, ..., 0, 3], ida[:, 0:1, ..., 1, 0],
                         ida[:, 0:1, ..., 1, 1], ida[:, 0:1, ..., 1, 3]], -1)
        v = torch.cat([v, s2e.reshape(B, 1, N, -1)], -1).reshape(B * N, 22)
        return torch.cat([v, v.new_zeros(B * N, 2)], 1)                # padded to 24 columns (vector loads)

    # ------------------------------------------------------------------ sub-graphs
    def _backbone(self, x):
        e, w = self.eng, self.w
        x = e.maxpool3x3s2(x, 'rs.pool', fmt='s')
        outs = []
        for bi, blk in enumerate(self.blocks):
            y = e.conv(x, blk['c1'], name='rs.y1', act=ACT_RELU, fmt='s')
            y = e.conv(y, blk['c2'], name='rs.y2', stride=blk['stride'], pad=1, act=ACT_RELU, fmt='s')
            idt = x if blk['down'] is None else e.conv(x, blk['down'], name='rs.idt', stride=blk['stride'], fmt='s')
            name = f"rs.c{blk['stage'] + 2}" if blk['last'] else f'rs.o{bi % 2}'
            x = e.conv(y, blk['c3'], name=name, act=ACT_RELU, res=idt, fmt='s')
            if blk['last']:
                outs.append(x)
        return outs

    def _pafpn(self, c, dst=None):
        """dst: optional list of 4 FMaps (or None entries) the final outputs are written into — the skip slices of the UNet's concat
        buffers, so that the skip connection costs no copy (lss.py:254-256 concatenates them)."""
        e, w = self.eng, self.w
        dst = dst or [None] * 4
        lat = [None] * 4
        lat[3] = e.conv(c[3], w['lateral_convs'][3], name='fpn.lat3', fmt='s')
        for i in (2, 1, 0):                                            # top-down, nearest x2 fused as residual
            lat[i] = e.conv(c[i], w['lateral_convs'][i], name=f'fpn.lat{i}', res=lat[i + 1], res_mode=lib.RES_UP2, fmt='s')
        inter = [e.conv(lat[i], w['fpn_convs'][i], out=dst[0] if i == 0 else None, name=f'fpn.int{i}', pad=1, fmt='s') for i in range(4)]
        for i in range(3):                                             # bottom-up: inter[i+1] += down(inter[i])
            inter[i + 1] = e.conv(inter[i], w['downsample_convs'][i], out=inter[i + 1], name=f'fpn.down{i}', stride=2,
                                  pad=1, res=inter[i + 1])
        outs = [inter[0]] + [e.conv(inter[i], w['pafpn_convs'][i - 1], out=dst[i], name=f'fpn.out{i}', pad=1, fmt='s') for i in (1, 2, 3)]
        return outs

    def _se_vec(self, m, which):
        e, w = self.eng, self.w
        h = e.linear(m, w[which + '_fc1'], name=f'dn.{which}.h1', act=ACT_RELU)
        h = e.linear(h, w[which + '_fc2'], name=f'dn.{which}.h2')
        h = e.linear(h, w[which + '_red'], name=f'dn.{which}.h3', act=ACT_RELU)
        return e.linear(h, w[which + '_exp'], name=f'dn.{which}.g')    # pre-sigmoid gate (BN, mid)

    def _depthnet(self, src, mlp_in, merge_in):
        e, w = self.eng, self.w
        BN, H, W = src.N, src.H, src.W
        x = e.conv(src, w['reduce'], name='dn.x', pad=1, act=ACT_RELU)
        gc, gd = self._se_vec(mlp_in, 'context'), self._se_vec(mlp_in, 'depth')
        cx = e.se_gate(x, gc, 'dn.cx')
        e.conv(cx, w['context_conv'], out=merge_in.slice(0, self.output_channels), name='dn.context')
        y = e.se_gate(x, gd, 'dn.dx')
        for i, (c1, c2) in enumerate(w['bb']):                          # 3 x BasicBlock
            t = e.conv(y, c1, name='dn.bb.t', pad=1, act=ACT_RELU, fmt='s')
            y = e.conv(t, c2, name=f'dn.bb.o{i % 2}', pad=1, act=ACT_RELU, res=y)
        cat = e.fmap('dn.aspp.cat', BN, H, W, 5 * self.mid)
        for i, dil in enumerate([1, 6, 12, 18]):
            e.conv(y, w['aspp'][i], out=cat.slice(i * self.mid, self.mid), name=f'dn.aspp{i}', pad=0 if i == 0 else dil,
                   dil=dil, act=ACT_RELU)
        g = e.global_avgpool(y, 'dn.aspp.gap')
        g = e.linear(g, w['aspp_gap'], name='dn.aspp.gapc', act=ACT_RELU)
        e.broadcast_rows(g, cat.slice(4 * self.mid, self.mid))
        y = e.conv(cat, w['aspp_out'], name='dn.aspp.out', act=ACT_RELU)
        off = e.conv(y, w['dcn_off'], name='dn.dcn.off', pad=1)
        col = e.out_map('dn.dcn.col', BN, H, W, 9 * self.mid, fmt='s')     # only the grouped GEMMs read the sampled columns
        lib.call('tt_dcn_im2col', _p(y.t), _p(off.t), off.ld, _p(col.t), BN, H, W, self.mid, 4)
        e.finish_out(col)
        yd = e.fmap('dn.dcn.out', BN, H, W, self.mid)
        cg = 9 * self.mid // len(w['dcn'])
        for g, wg in enumerate(w['dcn']):                               # grouped conv = one dense GEMM per group slice
            e.conv(col.slice(g * cg, cg), wg, out=yd.slice(g * wg.Cout, wg.Cout), name='dn.dcn.out')
        y = yd
        return e.conv(y, w['depth_out'], name='dn.depth', fmt='f')

    def _upcat(self, x, skip, ups, name):
        """cat[ConvTranspose2d_k2s2(x), skip] into one buffer (lss.py:254-256)."""
        e = self.eng
        cu = ups[0][0].Cout
        cat = e.fmap(name, x.N, 2 * x.H, 2 * x.W, cu + skip.C, fmt='s')
        up = cat.slice(0, cu)
        for i in range(2):
            for j in range(2):
                e.conv(x, ups[i][j], out=up, name=f'{name}.up{i}{j}', scatter=(2, i, 2, j))
        if not (skip.t is cat.t and skip.s is cat.s and skip.coff == cu):   # else the PAFPN wrote the skip straight into the slice
            e.copy_cols(skip, cat.slice(cu, skip.C))
        return cat

    def _unet(self, f):
        e, w = self.eng, self.w
        d = e.conv(self._upcat(f[3], f[2], w['u4_up'], 'un.cat4'), w['u4_conv'], name='un.d4', pad=1, act=ACT_RELU, fmt='s')
        d = e.conv(self._upcat(d, f[1], w['u3_up'], 'un.cat3'), w['u3_conv'], name='un.d3', pad=1, act=ACT_RELU, fmt='s')
        d = e.conv(self._upcat(d, f[0], w['u2_up'], 'un.cat2'), w['u2_conv'], name='un.d2', pad=1, act=ACT_RELU)
        d = e.upsample2x(d, 'un.up0', fmt='s')
        d = e.conv(d, w['u0_a'], name='un.d0a', pad=1, act=ACT_RELU, fmt='s')
        d = e.conv(d, w['u0_b'], name='un.d0b', pad=1, fmt='s')
        return e.conv(d, w['seg_last'], name='seg').slice(0, self.seg_classes)

    def _seg_to_feat(self, seg, out):
        e, w = self.eng, self.w
        x = seg
        spec = [(1, 1), (1, 1), (3, 2), (1, 1), (3, 2), (1, 1), (3, 2)]
        for i, (k, s) in enumerate(spec):
            last = i == len(spec) - 1
            x = e.conv(x, w['s2f'][i], out=out if last else None, name=f's2f.{i}', stride=s, pad=k // 2, act=ACT_RELU, fmt='s')
            x = x.slice(0, self.s2f_channels[i]) if not last else x
        return x

    def _single_sweep(self, imgs, s, bev_out):
        """lss.py:542-621 for one sweep; imgs (B, N, 3, H, W) NCHW on the device; s = sweep distance from the key frame."""
        e, w = self.eng, self.w
        B, N = imgs.shape[:2]
        raw = imgs.dtype == torch.uint8                              # (B, N, h, w, 3) camera frames: pre-processing fused into the staging
        if raw and self.pre is None:
            raise lib.TTError('raw uint8 frames need an AgentPreprocessor (EncoderDecoder.attach_preprocessor)')
        # stem (ResNet conv1 7x7 s2 p3 on 3 channels): the image goes channels-last (4 floats / pixel) into a buffer with
        # a physical zero border (3 px top / bottom / left, 5 right), so a tap ROW of 7 px x 4 floats is 28 contiguous
        # floats = one 32-float K slab of a 7x1 conv over 32 "channels" (x_ld = 4 < Cin = 32, row pitch x_hstride)
        im = imgs.reshape(B * N, *imgs.shape[2:])                      # (a view: every sweep is its own contiguous buffer)
        H0, W0 = self.final_dim if raw else im.shape[2:]
        if raw and not e.split:                                      # fp32 engines: the pre-processor writes the NCHW tensor they stage themselves
            im = self.pre.images(im, out=e.buf('img.f32', (B * N, 3, H0, W0)))
        Wp = W0 + 8
        if e.split:
            # scaled-split engine: 8 halves per pixel (3 channels + 5 zeros; TMA strides are multiples of 16 bytes), so a tap row
            # of 7 px is 56 halves inside ONE 64-half K slab; the planes are written straight from the NCHW image, the zero border
            # (zeroed once at allocation) is never touched
            rows = B * N * (H0 + 6) * Wp
            pbs = e.buf('img.nhwc#s8', (2, rows * 8), torch.float16, zero=True)
            if raw:                                                  # uint8 frame -> undistort / resize / crop / normalise -> planes, one kernel
                self.pre.images_to_stem(im, pbs, rows * 8, (H0 + 6, Wp, 3, 3))
            else:
                lib.call('tt_image_to_split8', _p(im.contiguous()), _p(pbs), C.c_longlong(rows * 8), B * N, im.shape[1], H0, W0, H0 + 6, Wp, 3, 3)
            x = e.conv(FMap(None, B * N, H0 + 6, W0, 64, ld=8, s=pbs), w['stem'], name='rs.stem', stride=2, act=ACT_RELU,
                       x_hstride=Wp * 8, x_nstride=(H0 + 6) * Wp * 8, fmt='f')    # only the max-pool (an fp32 kernel) reads it
        else:
            pb = e.nchw_to_nhwc_padded(im, 'img.nhwc', 4, 3, 3, 3, 5)
            x = e.conv(FMap(pb, B * N, H0 + 6, W0, 32, ld=4), w['stem'], name='rs.stem', stride=2, act=ACT_RELU,
                       x_hstride=Wp * 4, x_nstride=(H0 + 6) * Wp * 4)
        c = self._backbone(x)
        # the UNet's concat buffers exist before the PAFPN runs: its outputs 0..2 ARE the skip halves of cat2 / cat3 / cat4
        skips = []
        for lvl, (nm, cu) in enumerate((('un.cat2', w['u2_up'][0][0].Cout), ('un.cat3', w['u3_up'][0][0].Cout), ('un.cat4', w['u4_up'][0][0].Cout))):
            ci = c[lvl]
            skips.append(e.fmap(nm, ci.N, ci.H, ci.W, cu + 256, fmt='s').slice(cu, 256))
        fpn = self._pafpn(c, skips + [None])
        src = e.conv(fpn[2], w['neck_conv'], name='img_feats', fmt='s')
        mlp_in = e.wrap(e.static('in.mlp_in').view(B * N, 1, 1, 24))
        merge_in = e.fmap('merge_in', B * N, src.H, src.W, self.output_channels + 128, fmt='s')
        depth = self._depthnet(src, mlp_in, merge_in)
        seg = self._unet(fpn)
        self._seg_to_feat(seg, merge_in.slice(self.output_channels, 128))
        feat = e.conv(merge_in, w['merge'], name='img_feature', pad=1, fmt='f')
        m = e.static(f'in.lift_mats.{s}')                           # staged by stage(): ida^-1 and sensor2ego @ intrin^-1
        d = LiftSplatDesc()
        d.B, d.N, d.D, d.fH, d.fW, d.C = B, N, self.depth_channels, src.H, src.W, self.output_channels
        d.ld_d, d.d_coff, d.ld_c, d.c_coff = depth.ld, 0, feat.ld, 0
        d.lower, d.size = lib.f3(self.lower), lib.f3(self.h_vsize)
        d.X, d.Y, d.Z = self.h_vnum
        d.bev_ld, d.bev_coff, d.anti_transpose = bev_out.ld, 0, 0
        ws = e.buf('lift.ws', (lib.load().tt_lift_splat_workspace_bytes(C.byref(d)),), dtype=torch.uint8)
        lib.call('tt_lift_splat', C.byref(d), _p(depth.t), _p(feat.t), _p(m), _p(self.fu), _p(self.fv), _p(self.fd),
                 _p(bev_out.t, bev_out.coff), _p(ws))
        e.sync_split(bev_out)
        return dict(fpn_feats=fpn, depth=depth, seg=seg, img_feature=feat)

    # ------------------------------------------------------------------ forward (lss.py:635-724)
    def stage(self, img_metas, B, T, N):
        """Host half of LSS.forward: the python loops over img_metas (lss.py:667-687), the 4x4 inverses of
        get_geometry (lss.py:496,502) and DepthNet's 22 camera scalars (lss.py:206-231), uploaded into static
        device buffers.  Tiny, input-derived, and kept off the device path so the forward is graph-capturable."""
        e = self.eng
        mats = self.build_mats(img_metas, N)
        e.upload('in.mlp_in', self.depthnet_mlp_input(mats).contiguous())
        for s in range(T):
            idx = -1 if s == 0 else -s                              # key frame -> -1; history sweep s -> mats[-s] (lss.py:712-716)
            ida_inv = torch.inverse(mats['ida_mats'][:, idx])
            comb = mats['sensor2ego_mats'][:, idx].matmul(torch.inverse(mats['intrin_mats'][:, idx]))
            e.upload(f'in.lift_mats.{s}', torch.stack([ida_inv, comb], 2).reshape(B * N, 32).contiguous())
        return (torch.stack([m[-1]['lidar2img'] for m in img_metas], 0).float(), mats['ida_mats'][:, -1].clone())

    # ------------------------------------------------------------------ streaming BEV cache
    # lss.py:712-716 builds the history sweep's BEV from the PREVIOUS tick's images with the CURRENT key-frame matrices.  In closed loop
    # (thinktwice_agent.py: one forward per tick, queue of 2 frames) those images are the previous tick's key frame, and with a static rig
    # the matrices do not change from tick to tick — so that BEV is exactly the key-frame BEV the previous forward already computed.
    # With the cache on, a forward whose predecessor left a BEV behind copies it instead of re-running the camera encoder on the history
    # sweep (half of the camera branch).  The caller vouches for the stream being consecutive (reset_stream() at an episode start).
    stream_cache = False
    _cache_B = None
    pre = None                     # thinktwice_b200.preprocess.AgentPreprocessor for uint8 `img_raw` batches

    def cache_ready(self, B):
        return bool(self.stream_cache and self.queue_len == 2 and self._cache_B == B)

    def reset_stream(self):
        self._cache_B = None

    def _bev_bufs(self, B, T):
        e, Cb = self.eng, self.output_channels
        X, Y = self.h_vnum[0], self.h_vnum[1]
        bev_cat = e.fmap('bev_cat', B, Y, X, Cb * T)
        cache = e.fmap('bev.keycache', B, Y, X, Cb) if (self.stream_cache and T == 2) else None
        return bev_cat, cache

    def history_device(self, imgs, warm=False):
        """Device half, part 1: the history sweeps (lss.py:710-717) — needs every image but the key frame's.  imgs: one
        (B, N, 3, H, W) fp32 device tensor per sweep, oldest first, key frame last (entries may be None when `warm`).
        warm: take the history sweep's BEV from the streaming cache."""
        e, T, Cb = self.eng, len(imgs), self.output_channels
        assert T == self.queue_len, 'LSS.queue_len must be set correctly in config!'
        bev_cat, cache = self._bev_bufs(imgs[-1].shape[0], T)
        # history sweeps first (their buffers are recycled), key frame last so its FPN maps stay live.
        # bev_feature_list = [key, sweep 1, ...] (lss.py:697,717)
        if warm:
            e.copy_cols(cache, bev_cat.slice(Cb, Cb))
        else:
            for s in range(T - 1, 0, -1):
                self._single_sweep(imgs[T - 1 - s], s, bev_cat.slice(Cb * s, Cb))

    def key_device(self, imgs):
        """Device half, part 2: the key-frame sweep, then the sweep merge (lss.py:697-724)."""
        e, T, Cb = self.eng, len(imgs), self.output_channels
        bev_cat, cache = self._bev_bufs(imgs[-1].shape[0], T)
        key = self._single_sweep(imgs[-1], 0, bev_cat.slice(0, Cb))
        if cache is not None:
            e.copy_cols(bev_cat.slice(0, Cb), cache)                 # this tick's key-frame BEV is the next tick's history BEV
        bev = e.conv(bev_cat, self.w['sweep_merge'], name='bev', pad=1) if T > 1 else bev_cat
        return dict(bev=bev, seg=key['seg'], depth=key['depth'], fpn_feats=key['fpn_feats'], img_feature=key['img_feature'])

    def forward_device(self, img, warm=False):
        """Device half: img (B, T, N, 3, H, W) fp32 resident on the GPU, or the list of its T sweeps (B, N, 3, H, W)."""
        imgs = list(img) if isinstance(img, (list, tuple)) else [img[:, t].contiguous() for t in range(img.shape[1])]
        self.history_device(imgs, warm)
        return self.key_device(imgs)

    def forward(self, img, img_metas, timestamps=None, is_return_depth=False):
        if img.dim() == 5:
            img = img.unsqueeze(1)
        B, T, N = img.shape[:3]
        lidar2img, ida = self.stage(img_metas, B, T, N)
        outs = self.forward_device(img.to(self.eng.device))
        outs['lidar2img'], outs['ida_mat'] = lidar2img, ida
        return outs
