"""Parameter inventory of the model, in the reference's state_dict naming.

`param_spec(model_cfg)` lists every parameter / buffer (name, shape, kind) the forward path owns, so the
boundary classes can be built from a config alone and load an mmcv-format checkpoint
(`thinktwice_agent.py:170`) by name.  `ParamTree` materialises the list as a nested nn.Module whose
state_dict keys are exactly those names.
"""
import math

import torch
import torch.nn as nn


class Spec(list):
    def add(self, name, shape, kind):
        self.append((name, tuple(int(s) for s in shape), kind))

    def conv(self, n, cout, cin, k, bias=True, kw=None):
        self.add(n + '.weight', (cout, cin, k, kw if kw is not None else k), 'conv')
        if bias:
            self.add(n + '.bias', (cout,), 'bias')

    def convT(self, n, cin, cout, bias=True):
        self.add(n + '.weight', (cin, cout, 2, 2), 'convT')
        if bias:
            self.add(n + '.bias', (cout,), 'bias')

    def lin(self, n, cout, cin):
        self.add(n + '.weight', (cout, cin), 'linear')
        self.add(n + '.bias', (cout,), 'bias')

    def bn(self, n, c):
        self.add(n + '.weight', (c,), 'one')
        self.add(n + '.bias', (c,), 'zero')
        self.add(n + '.running_mean', (c,), 'buf_zero')
        self.add(n + '.running_var', (c,), 'buf_one')
        self.add(n + '.num_batches_tracked', (), 'buf_long')

    def ln(self, n, c):
        self.add(n + '.weight', (c,), 'one')
        self.add(n + '.bias', (c,), 'zero')

    def mlp(self, n, dims, idx_step=2):
        for i in range(len(dims) - 1):
            self.lin(f'{n}.{i * idx_step}', dims[i + 1], dims[i])


def _lss(s, p, cfg):
    D = int(round((cfg['d_bound'][1] - cfg['d_bound'][0]) / cfg['d_bound'][2]))
    fH, fW = cfg['final_dim'][0] // cfg['downsample_factor'], cfg['final_dim'][1] // cfg['downsample_factor']
    if cfg['queue_len'] != 1:
        s.conv(p + 'bev_multiframe_merge', 256, 256 * cfg['queue_len'], 3, bias=False)
    s.add(p + 'voxel_size', (3,), 'buf'); s.add(p + 'voxel_coord', (3,), 'buf')
    s.add(p + 'voxel_num', (3,), 'buf_long'); s.add(p + 'frustum', (D, fH, fW, 4), 'buf')
    b = p + 'img_backbone.'
    s.conv(b + 'conv1', 64, 3, 7, bias=False); s.bn(b + 'bn1', 64)
    cin = 64
    for li, (planes, n) in enumerate([(64, 3), (128, 4), (256, 6), (512, 3)]):
        for i in range(n):
            q = f'{b}layer{li + 1}.{i}.'
            s.conv(q + 'conv1', planes, cin, 1, bias=False); s.bn(q + 'bn1', planes)
            s.conv(q + 'conv2', planes, planes, 3, bias=False); s.bn(q + 'bn2', planes)
            s.conv(q + 'conv3', planes * 4, planes, 1, bias=False); s.bn(q + 'bn3', planes * 4)
            if i == 0:
                s.conv(q + 'downsample.0', planes * 4, cin, 1, bias=False); s.bn(q + 'downsample.1', planes * 4)
            cin = planes * 4
    nk = p + 'img_neck.'
    for i, c in enumerate(cfg['img_neck_conf']['in_channels']):
        s.conv(f'{nk}lateral_convs.{i}.conv', 256, c, 1)
    for i in range(4):
        s.conv(f'{nk}fpn_convs.{i}.conv', 256, 256, 3)
    for i in range(3):
        s.conv(f'{nk}downsample_convs.{i}.conv', 256, 256, 3)
    for i in range(3):
        s.conv(f'{nk}pafpn_convs.{i}.conv', 256, 256, 3)
    cin_d, mid = cfg['depth_net_conf']['in_channels'], cfg['depth_net_conf']['mid_channels']
    s.conv(p + 'neck_conv', cin_d, 256, 1)
    d = p + 'depth_net.'
    s.conv(d + 'reduce_conv.0', mid, cin_d, 3); s.bn(d + 'reduce_conv.1', mid)
    s.conv(d + 'context_conv', cfg['output_channels'], mid, 1)
    s.bn(d + 'bn', 22)
    for m in ('depth', 'context'):
        s.lin(f'{d}{m}_mlp.fc1', mid, 22); s.lin(f'{d}{m}_mlp.fc2', mid, mid)
        s.conv(f'{d}{m}_se.conv_reduce', mid, mid, 1); s.conv(f'{d}{m}_se.conv_expand', mid, mid, 1)
    for i in range(3):
        q = f'{d}depth_conv.{i}.'
        s.conv(q + 'conv1', mid, mid, 3, bias=False); s.bn(q + 'bn1', mid)
        s.conv(q + 'conv2', mid, mid, 3, bias=False); s.bn(q + 'bn2', mid)
    a = d + 'depth_conv.3.'
    for i, k in enumerate([1, 3, 3, 3]):
        s.conv(f'{a}aspp{i + 1}.atrous_conv', mid, mid, k, bias=False); s.bn(f'{a}aspp{i + 1}.bn', mid)
    s.conv(a + 'global_avg_pool.1', mid, mid, 1, bias=False); s.bn(a + 'global_avg_pool.2', mid)
    s.conv(a + 'conv1', mid, mid * 5, 1, bias=False); s.bn(a + 'bn1', mid)
    s.add(d + 'depth_conv.4.weight', (mid, mid // 4, 3, 3), 'conv')
    s.add(d + 'depth_conv.4.conv_offset.weight', (18, mid, 3, 3), 'zero'); s.add(d + 'depth_conv.4.conv_offset.bias', (18,), 'zero')
    s.conv(d + 'depth_conv.5', D, mid, 1)
    u = p + 'seg_net.'
    f = cfg['fpn_in_channels']
    ncls = cfg['seg_net_conf']['out_channels']
    s.convT(u + 'unet_layer4.up', f[3], 256); s.conv(u + 'unet_layer4.conv_relu.0', 256, 256 + f[2], 3)
    s.convT(u + 'unet_layer3.up', 256, 256); s.conv(u + 'unet_layer3.conv_relu.0', 256, 256 + f[1], 3)
    s.convT(u + 'unet_layer2.up', 256, 128); s.conv(u + 'unet_layer2.conv_relu.0', 128, 128 + f[0], 3)
    s.conv(u + 'unet_layer0.1', 64, 128, 3, bias=False); s.conv(u + 'unet_layer0.3', 64, 64, 3, bias=False)
    s.conv(u + 'conv_last', ncls, 64, 1)
    r = p + 'seg_res_to_image_feature.'
    for i, (ci, co, k) in enumerate([(ncls, 64, 1), (64, 16, 1), (16, 32, 3), (32, 32, 1), (32, 64, 3), (64, 64, 1), (64, 128, 3)]):
        s.conv(f'{r}{3 * i}', co, ci, k); s.bn(f'{r}{3 * i + 1}', co)
    s.conv(p + 'merge_seg_and_image', 256, 256 + 128, 3)


def _lidar(s, p, cfg):
    me = cfg['pts_middle_encoder']
    m = p + 'pts_middle_encoder.'

    def sp(n, cout, cin, k):
        k = k if isinstance(k, (list, tuple)) else (k, k, k)
        s.add(n + '.weight', (cout, k[0], k[1], k[2], cin), 'spconv')

    def bn1(n, c):
        s.bn(n, c)
    base = 16
    sp(m + 'conv_input.0', base, me['in_channels'], 3); bn1(m + 'conv_input.1', base)
    cin = base
    nst = len(me['encoder_channels'])
    for i, blocks in enumerate(me['encoder_channels']):
        for j, cout in enumerate(blocks):
            q = f'{m}encoder_layers.encoder_layer{i + 1}.{j}.'
            if j == len(blocks) - 1 and i != nst - 1:
                sp(q + '0', cout, cin, 3); bn1(q + '1', cout)
            else:
                sp(q + 'conv1', cout, cout, 3); bn1(q + 'bn1', cout); sp(q + 'conv2', cout, cout, 3); bn1(q + 'bn2', cout)
            cin = cout
    sp(m + 'conv_out.0', me['output_channels'], cin, (3, 1, 1)); bn1(m + 'conv_out.1', me['output_channels'])
    bb = cfg['pts_backbone']
    cin = bb['in_channels']
    for bi, (cout, n) in enumerate(zip(bb['out_channels'], bb['layer_nums'])):
        for k in range(n + 1):
            s.conv(f'{p}pts_backbone.blocks.{bi}.{3 * k}', cout, cin if k == 0 else cout, 3, bias=False)
            s.bn(f'{p}pts_backbone.blocks.{bi}.{3 * k + 1}', cout)
        cin = cout
    nk = cfg['pts_neck']
    for i, (ci, co, st) in enumerate(zip(nk['in_channels'], nk['out_channels'], nk['upsample_strides'])):
        if st > 1:
            s.convT(f'{p}pts_neck.deblocks.{i}.0', ci, co, bias=False)
        else:
            s.conv(f'{p}pts_neck.deblocks.{i}.0', co, ci, 1, bias=False)
        s.bn(f'{p}pts_neck.deblocks.{i}.1', co)


def _se_block(s, n, c):
    s.conv(n + '.conv1', 2 * c, c, 3); s.bn(n + '.bn1', 2 * c)
    s.conv(n + '.conv2', c, 2 * c, 3); s.bn(n + '.bn2', c)
    s.conv(n + '.se.fc1', c, c, 1); s.conv(n + '.se.fc2', c, c, 1)


def _framework(s):
    for n, cin in (('conv_cam', 256), ('conv_lidar', 512), ('conv_fusion', 512)):
        s.conv(n + '.0', 256, cin, 3, bias=False); s.bn(n + '.1', 256)
        s.conv(n + '.3', 256, 256, 3, bias=False); s.bn(n + '.4', 256)
    s.conv('_256_to_32', 32, 256, 3)
    for n, c in (('MLP21', 32), ('MLP10', 64), ('MLP4', 128), ('MLP2', 256)):
        _se_block(s, n, c)
    s.conv('conv21_10', 64, 32, 3); s.conv('conv10_4', 128, 64, 3); s.conv('conv4_2', 256, 128, 3)
    s.lin('output_fc.0', 512, 1024); s.bn('output_fc.2', 512); s.lin('output_fc.3', 256, 512)
    s.lin('measurements_encoder.0', 128, 9); s.lin('measurements_encoder.2', 128, 128)


def _decoder(s, p, cfg):
    T = cfg['pred_len']
    s.mlp(p + 'join_traj', [384, 512, 512, 256]); s.mlp(p + 'output_traj', [256, 512, 2 * T])
    s.mlp(p + 'join_ctrl', [384, 512, 512, 256]); s.mlp(p + 'speed_branch', [256, 256, 256, 1])
    s.mlp(p + 'value_branch_traj', [256, 256, 256, 1]); s.mlp(p + 'value_branch_ctrl', [256, 256, 256, 1])
    s.mlp(p + 'policy_head', [256, 512, 512]); s.mlp(p + 'dist_mu', [512, 512, 2 * T]); s.mlp(p + 'dist_sigma', [512, 512, 2 * T])
    for i in range(4):
        s.conv(f'{p}fpn_linear{i}', 256, cfg['FPN_out_channels'][i], 1)
    s.add(p + 'temporal_embedding', (T, 128), 'emb'); s.add(p + 'cams_embeds', (4, 256), 'emb')
    s.add(p + 'static_embedding', (4, 128), 'emb'); s.add(p + 'level_embeds', (4, 256), 'emb_zero')
    for k in range(cfg['refine_num']):
        q = f'{p}decoder_layers.{k}.'
        g = q + 'prediction_module.spatial_gru.'
        for n, cin in (('conv_update', 38), ('conv_reset', 38), ('conv_state_tilde', 38), ('conv_decoder', 32)):
            s.conv(f'{g}{n}.0', 32, cin, 3); s.conv(f'{g}{n}.2', 32, 32, 3)
        f = q + 'prediction_module.ffn.'                       # dead at inference (thinktwice_decoder.py:44-46)
        s.conv(f + '0', 64, 32, 1); s.conv(f + '2', 32, 64, 3); s.conv(f + '4', 32, 32, 1)
        c = q + 'look_module.cam_look_module.'
        s.lin(c + 'deformable_attention.sampling_offsets', 512, 256)
        s.lin(c + 'deformable_attention.attention_weights', 256, 256)
        s.lin(c + 'deformable_attention.value_proj', 256, 256)
        s.ln(c + 'query_linear.0', 1543); s.lin(c + 'query_linear.1', 512, 1543); s.lin(c + 'query_linear.3', 256, 512)
        s.ln(c + 'ffn.norm', 256); s.lin(c + 'ffn.w_1', 1024, 256); s.lin(c + 'ffn.w_2', 256, 1024)
        s.ln(c + 'output_proj.0', 1024); s.lin(c + 'output_proj.1', 512, 1024); s.lin(c + 'output_proj.3', 256, 512)
        l = q + 'look_module.'                                 # dead / unused branches, kept for checkpoint parity
        s.lin(l + 'lidar_look_module_atten.0', 256, 134); s.lin(l + 'lidar_look_module_atten.2', 512, 256)
        s.lin(l + 'lidar_look_module_MLP.0', 128, 512); s.lin(l + 'lidar_look_module_MLP.3', 256, 9 * 128)
        s.lin(l + 'look_feature_MLP.0', 512, 2048); s.lin(l + 'look_feature_MLP.2', 128, 512)
        s.ln(q + 'mlp.0', 1024); s.lin(q + 'mlp.1', 512, 1024); s.lin(q + 'mlp.4', 512, 512)
        s.mlp(q + 'traj_offset_module', [514, 256, 64, 2]); s.mlp(q + 'ctrl_offset_module', [516, 256, 64, 4])
        s.conv(q + 'BEV_feat_update_module.0', 128, 2080, 3); s.conv(q + 'BEV_feat_update_module.2', 32, 128, 3)
        s.lin(q + 'flattened_BEV_feat_update_module.0', 512, 2304); s.lin(q + 'flattened_BEV_feat_update_module.2', 256, 512)


def param_spec(model_cfg):
    s = Spec()
    _lss(s, 'img_encoder.', model_cfg['img_encoder'])
    _lidar(s, 'lidar_encoder.', model_cfg['lidar_encoder'])
    _framework(s)
    _decoder(s, 'decoder.', model_cfg['decoder']['config'])
    return s


def _init(t, kind, gen):
    with torch.no_grad():
        if kind in ('conv', 'linear'):                         # He-normal: keeps activations O(1) through ReLU stacks
            t.copy_(torch.randn(t.shape, generator=gen) * math.sqrt(2.0 / t[0].numel()))
        elif kind == 'convT':
            t.copy_(torch.randn(t.shape, generator=gen) * math.sqrt(1.0 / t.shape[0]))
        elif kind == 'spconv':
            t.copy_(torch.randn(t.shape, generator=gen) * math.sqrt(2.0 / (t[0].numel())))
        elif kind == 'emb':                                    # trunc_normal_(std=0.02) (thinktwice_decoder.py:373-376)
            t.copy_((torch.randn(t.shape, generator=gen) * 0.02).clamp_(-2, 2))
        elif kind in ('one', 'buf_one'):
            t.fill_(1)
        else:
            t.zero_()


class ParamTree(nn.Module):
    """nested container whose state_dict() keys are the spec names."""

    def __init__(self, spec, seed=0):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        for name, shape, kind in spec:
            parts = name.split('.')
            node = self
            for part in parts[:-1]:
                if not hasattr(node, part):
                    node.add_module(part, nn.Module())
                node = getattr(node, part)
            if kind.startswith('buf'):
                t = torch.zeros(shape, dtype=torch.long if kind == 'buf_long' else torch.float32)
                _init(t, kind, gen)
                node.register_buffer(parts[-1], t)
            else:
                t = torch.zeros(shape)
                _init(t, kind, gen)
                node.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
