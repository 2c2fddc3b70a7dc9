"""Oracle: stacked Look-Prediction-Refine decoder.  TEST INFRASTRUCTURE ONLY.

Restates open_loop_training/code/model_code/dense_heads/thinktwice_decoder.py:26-489,
multi_scale_deformable_attn_function.py:197-526 (with mmcv's pure-torch
multi_scale_deformable_attn_pytorch) and dense_heads/utils.py:53-106.  All reference
quirks are kept: batch-coupled max_len, "zero the first B rows, divide by B" (msda:338-341),
dead PredictionModule.ffn and dead LiDAR look branch (executed only when dead_work=True, as
the reference does, so the CPU baseline pays what the reference pays).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

ReLU = nn.ReLU


def msda_pytorch(value, spatial_shapes, sampling_locations, attention_weights):
    """mmcv.ops.multi_scale_deform_attn.multi_scale_deformable_attn_pytorch.
    value (bs, keys, heads, dh); spatial_shapes [(H, W)]; sampling_locations (bs, q, heads, L, P, 2) in [0,1]
    (x, y); attention_weights (bs, q, heads, L, P) -> (bs, q, heads*dh)."""
    bs, _, nh, dh = value.shape
    nq, L, P = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    vals = value.split([int(h) * int(w) for h, w in spatial_shapes], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for l, (h, w) in enumerate(spatial_shapes):
        v = vals[l].flatten(2).transpose(1, 2).reshape(bs * nh, dh, int(h), int(w))
        g = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros', align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(bs * nh, 1, nq, L * P)
    out = (torch.stack(sampled, -2).flatten(-2) * aw).sum(-1).view(bs, nh * dh, nq)
    return out.transpose(1, 2).contiguous()


class MSDeformableAttention3D(nn.Module):                          # msda:346-526
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=8):
        super().__init__()
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):                                        # msda:403-421
        nn.init.zeros_(self.sampling_offsets.weight)
        th = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        g = torch.stack([th.cos(), th.sin()], -1)
        g = (g / g.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            g[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = g.view(-1)
        nn.init.zeros_(self.attention_weights.weight); nn.init.zeros_(self.attention_weights.bias)
        nn.init.xavier_uniform_(self.value_proj.weight); nn.init.zeros_(self.value_proj.bias)

    def forward(self, query, value, reference_points, spatial_shapes):
        bs, nq, _ = query.shape
        nv = value.shape[1]
        value = self.value_proj(value).view(bs, nv, self.num_heads, -1)
        off = self.sampling_offsets(query).view(bs, nq, self.num_heads, self.num_levels, self.num_points, 2)
        aw = self.attention_weights(query).view(bs, nq, self.num_heads, self.num_levels * self.num_points)
        aw = aw.softmax(-1).view(bs, nq, self.num_heads, self.num_levels, self.num_points)
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)                 # (W, H)
        loc = reference_points[:, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
        return msda_pytorch(value, spatial_shapes, loc, aw)


class PositionwiseFeedForward(nn.Module):                          # msda:197-214
    def __init__(self, d_in, d_hid):
        super().__init__()
        self.norm = nn.LayerNorm(d_in)
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)

    def forward(self, x):
        return self.w_2(F.gelu(self.w_1(self.norm(x)))) + x


class SpatialCrossAttention(nn.Module):                            # msda:216-344
    def __init__(self, embed_dims=256, num_cams=4, query_dims=1543):
        super().__init__()
        self.embed_dims, self.num_cams = embed_dims, num_cams
        self.deformable_attention = MSDeformableAttention3D()
        self.query_linear = nn.Sequential(nn.LayerNorm(query_dims), nn.Linear(query_dims, 512), nn.GELU(),
                                          nn.Linear(512, embed_dims), nn.GELU())
        self.ffn = PositionwiseFeedForward(256, 1024)
        self.output_proj = nn.Sequential(nn.LayerNorm(num_cams * 256), nn.Linear(num_cams * 256, 512), nn.GELU(),
                                         nn.Linear(512, embed_dims))

    def forward(self, query, value, reference_points, spatial_shapes, indexes, keep=None):
        query = self.query_linear(query)
        bs, _, max_len, _ = query.shape
        l = value.shape[1]
        value = value.permute(2, 0, 1, 3).reshape(bs * self.num_cams, l, self.embed_dims)
        q = self.deformable_attention(query.view(bs * self.num_cams, max_len, self.embed_dims), value,
                                      reference_points.view(bs * self.num_cams, max_len, 2), spatial_shapes)
        q = self.ffn(q.view(bs, self.num_cams, max_len, self.embed_dims))
        if keep is not None:
            keep['sca_rows'] = q.clone()
        for j in range(bs):
            for i, per_cam in enumerate(indexes):                  # len(per_cam) == bs (a list of bs tensors): msda:338-341
                q[j, i, :len(per_cam)] = 0
                q[j, i] /= max(len(per_cam), 1.0)
        return self.output_proj(q.sum(-2).view(bs, -1))


class SpatialGRU(nn.Module):                                       # dense_heads/utils.py:53-106
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        def net(cin):
            return nn.Sequential(nn.Conv2d(cin, hidden_size, 3, padding=1), ReLU(),
                                 nn.Conv2d(hidden_size, hidden_size, 3, padding=1))
        self.conv_update = net(input_size + hidden_size)
        self.conv_reset = net(input_size + hidden_size)
        self.conv_state_tilde = net(input_size + hidden_size)
        self.conv_decoder = net(hidden_size)

    def forward(self, x, state):
        outs = []
        for t in range(x.shape[1]):
            xs = torch.cat([x[:, t], state], 1)
            u = torch.sigmoid(self.conv_update(xs))
            r = torch.sigmoid(self.conv_reset(xs))
            cand = self.conv_state_tilde(torch.cat([x[:, t], (1.0 - r) * state], 1))
            state = (1.0 - u) * state + u * cand
            outs.append(self.conv_decoder(state))
        return torch.stack(outs, 1)


class PredictionModule(nn.Module):                                 # thinktwice_decoder.py:26-46
    def __init__(self):
        super().__init__()
        self.spatial_gru = SpatialGRU(6, 32)
        self.ffn = nn.Sequential(nn.Conv2d(32, 64, 1), ReLU(), nn.Conv2d(64, 32, 3, padding=1), ReLU(), nn.Conv2d(32, 32, 1))

    def forward(self, bev, wp, ctrl, last_future, dead_work):
        x = torch.cat([wp, ctrl], 2)[..., None, None].repeat(1, 1, 1, bev.shape[-2], bev.shape[-1])
        fut = self.spatial_gru(x, bev).view(-1, *bev.shape[1:])
        if dead_work and last_future is not None:                  # result unused in the reference (:44-46)
            _ = self.ffn(fut) + last_future.view(-1, *bev.shape[1:])
        return fut


class LookModule(nn.Module):                                       # thinktwice_decoder.py:51-187
    def __init__(self):
        super().__init__()
        self.cam_look_module = SpatialCrossAttention()
        self.lidar_look_module_atten = nn.Sequential(nn.Linear(6 + 128, 256), ReLU(), nn.Linear(256, 512), nn.Sigmoid())
        self.lidar_look_module_MLP = nn.Sequential(nn.Linear(512, 128), ReLU(), nn.Flatten(start_dim=2),
                                                   nn.Linear(9 * 128, 256), ReLU())
        self.look_feature_MLP = nn.Sequential(nn.Linear(512 * 4, 512), ReLU(), nn.Linear(512, 128))
        self.point_cloud_range = [-8.0, -19.2, -4.0, 30.4, 19.2, 4.0]

    def lidar_look(self, wp, grid):                                # :79-85
        bs, T, _ = wp.shape
        r = self.point_cloud_range
        d = torch.tensor([0.0, -0.1, 0.1], device=wp.device)[None, None]
        rx = 1.0 - torch.clamp(((wp[..., 0] - r[0]) / (r[3] - r[0])).unsqueeze(-1) + d, min=0.0, max=1.0)
        ry = torch.clamp(((wp[..., 1] - r[1]) / (r[4] - r[1])).unsqueeze(-1) + d, min=0.0, max=1.0)
        rel = torch.stack([rx.unsqueeze(-1).repeat(1, 1, 1, 3), ry.unsqueeze(-2).repeat(1, 1, 3, 1)], -1)
        rel = rel.view(bs * T, -1, 1, 2) * 2 - 1
        s = F.grid_sample(grid.view(bs * T, *grid.shape[2:]), rel, align_corners=False)
        return s.view(bs, T, -1, 9).transpose(2, 3)

    def cam_ref_points_query(self, ref3d, lidar2img, ida, img_shape, query, mlvl_feats):      # :88-150
        ref = torch.cat((ref3d, torch.ones_like(ref3d[..., :1])), -1)
        B, nq = ref.shape[:2]
        N = lidar2img.size(1)
        ref = ref.view(B, 1, nq, 4).repeat(1, N, 1, 1).unsqueeze(-1)
        l2i = lidar2img.view(B, N, 1, 4, 4).repeat(1, 1, nq, 1, 1)
        idam = ida.view(B, N, 1, 4, 4).repeat(1, 1, nq, 1, 1)
        cam = torch.matmul(l2i, ref).squeeze(-1)
        eps = 1e-5
        cam2 = cam.clone()
        cam2[..., 0:2] = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.ones_like(cam[..., 2:3]) * eps)
        cam = torch.matmul(idam, cam2.unsqueeze(-1)).squeeze(-1)
        mask = cam[..., 2:3] > eps
        cam = cam[..., :2]
        cam[..., 0] /= img_shape[1]
        cam[..., 1] /= img_shape[0]
        mask = mask & (cam[..., 1:2] > 0.0) & (cam[..., 1:2] < 1.0) & (cam[..., 0:1] < 1.0) & (cam[..., 0:1] > 0.0)
        mask = mask.view(B, N, nq).permute(1, 0, 2)
        grid = cam.view(B * N, nq, 1, 2) * 2 - 1.0
        sampled = []
        for feat in mlvl_feats:
            C = feat.shape[1]
            s = F.grid_sample(feat, grid, align_corners=False)
            sampled.append(s.view(B, N, C, nq, 1).permute(0, 2, 3, 1, 4))
        sampled = torch.stack(sampled, -1).view(B, C, nq, N, len(mlvl_feats)).permute(0, 2, 3, 1, 4).reshape(B, nq, N, -1)
        indexes = [[mask[i, j].nonzero().squeeze(-1) for j in range(B)] for i in range(N)]
        max_len = max(len(s) for c in indexes for s in c)
        qd = query.shape[-1]
        q_re = query.new_zeros(B * N, max_len, qd + sampled.shape[-1])
        r_re = cam.new_zeros(B * N, max_len, 2)
        cam_p = cam.permute(1, 0, 2, 3)
        for i in range(N):
            for j in range(B):
                ix = indexes[i][j]
                if len(ix):
                    q_re[j * N + i, :len(ix)] = torch.cat([query[j, ix], sampled[j, ix, i]], -1)
                    r_re[j * N + i, :len(ix)] = cam_p[i, j, ix]
        return r_re.view(B, N, max_len, 2), q_re.view(B, N, max_len, -1), indexes

    def forward(self, wp, ctrl_sp, meas, flat, lidar2img, ida, img_size, mlvl_feats, fpn_flat, spatial_shapes,
                lidar_hi, temporal_embedding, static_embedding, dead_work, keep=None):         # :154-187
        B = wp.shape[0]
        static = torch.tensor([[5.0, 0.0], [0.0, -5.0], [0.0, 5.0], [-5.0, 0.0]], device=wp.device)[None].repeat(B, 1, 1)
        look = torch.cat([wp, static], 1)
        z = torch.linspace(-4, 10, 15, dtype=torch.float64, device=wp.device)[None, None, :, None].repeat(B, look.shape[1], 1, 1)
        look3d = torch.cat([look.unsqueeze(2).repeat(1, 1, 15, 1), z], -1).view(B, -1, 3).to(wp.dtype)
        in_ctrl = torch.cat([ctrl_sp.unsqueeze(2).repeat(1, 1, 15, 1).view(B, -1, 4),
                             torch.zeros(B, 4 * 15, 4, device=wp.device)], 1)
        emb = torch.cat([temporal_embedding[None, :, None].repeat(B, 1, 15, 1).view(B, -1, temporal_embedding.shape[-1]),
                         static_embedding[None, :, None].repeat(B, 1, 15, 1).view(B, -1, temporal_embedding.shape[-1])], 1)
        nq = look3d.shape[1]
        q = torch.cat([in_ctrl, look3d, emb, meas.unsqueeze(1).repeat(1, nq, 1), flat.unsqueeze(1).repeat(1, nq, 1)], -1)
        ref, q_re, indexes = self.cam_ref_points_query(look3d, lidar2img, ida, img_size, q, mlvl_feats)
        if keep is not None:
            keep.update(ref_rebatch=ref, query_rebatch=q_re, indexes=indexes)
        img_look = self.cam_look_module(q_re, fpn_flat, ref, spatial_shapes, indexes, keep)
        img_look = img_look.unsqueeze(1).repeat(1, temporal_embedding.shape[0], 1)
        if dead_work:                                              # computed then zeroed in the reference (:179-186)
            w = self.lidar_look_module_atten(torch.cat([wp, ctrl_sp, temporal_embedding[None].repeat(B, 1, 1)], -1))
            ll = self.lidar_look_module_MLP(self.lidar_look(wp, w[..., None, None] * lidar_hi.unsqueeze(1).float()))
            if keep is not None:
                keep['lidar_look'] = ll
        return torch.cat([img_look, img_look.new_zeros(B, temporal_embedding.shape[0], 256)], -1)


class ThinkTwiceDecoderLayer(nn.Module):                           # thinktwice_decoder.py:189-260
    def __init__(self):
        super().__init__()
        self.prediction_module = PredictionModule()
        self.look_module = LookModule()
        self.mlp = nn.Sequential(nn.LayerNorm(1024), nn.Linear(1024, 512), ReLU(), nn.Dropout(0.0), nn.Linear(512, 512), ReLU())
        self.traj_offset_module = nn.Sequential(nn.Linear(514, 256), ReLU(), nn.Linear(256, 64), ReLU(), nn.Linear(64, 2))
        self.ctrl_offset_module = nn.Sequential(nn.Linear(516, 256), ReLU(), nn.Linear(256, 64), ReLU(), nn.Linear(64, 4))
        self.BEV_feat_update_module = nn.Sequential(nn.Conv2d(512 * 4 + 32, 128, 3, padding=1), ReLU(), nn.Conv2d(128, 32, 3, padding=1))
        self.flattened_BEV_feat_update_module = nn.Sequential(nn.Linear(256 + 512 * 4, 512), ReLU(), nn.Linear(512, 256))

    def forward(self, bev, wp, ctrl, last_future, parent, grid2feat, meas, flat, lidar2img, ida, img_size, mlvl_feats,
                fpn_flat, spatial_shapes, lidar_hi, temb, semb, dead_work, keep=None):
        B = bev.shape[0]
        ctrl_sp = F.softplus(ctrl)
        fut = self.prediction_module(bev, wp, ctrl_sp, last_future, dead_work)
        fflat = grid2feat(fut, parent).view(B, 4, 256)
        fut = fut.view(B, -1, *bev.shape[1:])
        look = self.look_module(wp, ctrl_sp, meas, flat, lidar2img, ida, img_size, mlvl_feats, fpn_flat, spatial_shapes,
                                lidar_hi, temb, semb, dead_work, keep)
        a = self.mlp(torch.cat([fflat, look, temb[None].repeat(B, 1, 1), meas.unsqueeze(1).repeat(1, 4, 1)], -1))
        traj_off = self.traj_offset_module(torch.cat([wp, a], -1))
        ctrl_off = self.ctrl_offset_module(torch.cat([ctrl, a], -1))
        tiled = a.view(B, -1)[..., None, None].repeat(1, 1, bev.shape[2], bev.shape[3])
        new_bev = self.BEV_feat_update_module(torch.cat([bev, tiled], 1)) + bev
        new_flat = self.flattened_BEV_feat_update_module(torch.cat([flat, a.view(B, -1)], -1)) + flat
        if keep is not None:
            keep.update(all_future_feat=a, look_features=look, future_flat=fflat)
        return traj_off, ctrl_off, fut, new_bev, new_flat


def mlp(*dims, last_act=False):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2 or last_act:
            layers.append(ReLU())
    return nn.Sequential(*layers)


class ThinkTwiceDecoder(nn.Module):                                # thinktwice_decoder.py:262-489
    def __init__(self, config, bev_h=21, bev_w=21, **_unused):
        super().__init__()
        self.config = config
        self.pred_len = config['pred_len']
        self.join_traj = mlp(384, 512, 512, 256, last_act=True)
        self.output_traj = mlp(256, 512, 2 * self.pred_len)
        self.join_ctrl = mlp(384, 512, 512, 256, last_act=True)
        self.speed_branch = mlp(256, 256, 256, 1)
        self.value_branch_traj = mlp(256, 256, 256, 1)
        self.value_branch_ctrl = mlp(256, 256, 256, 1)
        self.policy_head = mlp(256, 512, 512, last_act=True)
        self.dist_mu = mlp(512, 512, 2 * self.pred_len)
        self.dist_sigma = mlp(512, 512, 2 * self.pred_len)
        for i in range(4):
            setattr(self, f'fpn_linear{i}', nn.Conv2d(config['FPN_out_channels'][i], 256, 1))
        self.temporal_embedding = nn.Parameter(torch.zeros(self.pred_len, 128))
        self.cams_embeds = nn.Parameter(torch.zeros(4, 256))
        self.static_embedding = nn.Parameter(torch.zeros(4, 128))
        self.level_embeds = nn.Parameter(torch.zeros(4, 256))
        self.decoder_layers = nn.ModuleList([ThinkTwiceDecoderLayer() for _ in range(config['refine_num'])])

    def transform_fpn_feats(self, mlvl):                           # :381-401
        shapes, flat = [], []
        for lvl, f in enumerate(mlvl):
            bn, c, h, w = f.shape
            f = f.view(bn // 4, 4, c, h, w).flatten(3).permute(1, 0, 3, 2)
            f = f + self.cams_embeds[:, None, None, :] + self.level_embeds[None, None, lvl:lvl + 1, :]
            shapes.append((h, w)); flat.append(f)
        flat = torch.cat(flat, 2)
        shapes = torch.as_tensor(shapes, dtype=torch.long, device=mlvl[0].device)
        return shapes, flat.permute(0, 2, 1, 3)                    # (cam, sum hw, bs, 256)

    @staticmethod
    def grid2feat(grid, parent):                                   # :405-415
        x = parent.MLP10(F.relu(parent.conv21_10(grid)))
        x = parent.MLP4(F.relu(parent.conv10_4(x)))
        x = parent.MLP2(F.relu(parent.conv4_2(x)))
        return parent.output_fc(x.flatten(1))

    def forward(self, flat, bev, meas, parent, look_meta, dead_work=False, keep=None):         # :419-489
        o = {'bev_feature': bev}
        o['pred_speed'] = self.speed_branch(flat)
        jt = self.join_traj(torch.cat([flat, meas], 1))
        o['pred_value_traj'] = self.value_branch_traj(jt); o['pred_features_traj'] = jt
        wps = [self.output_traj(jt).view(-1, self.pred_len, 2)]
        jc = self.join_ctrl(torch.cat([flat, meas], -1))
        o['pred_value_ctrl'] = self.value_branch_ctrl(jc); o['pred_features_ctrl'] = jc
        pol = self.policy_head(jc)
        ctrls = [torch.cat([self.dist_mu(pol).view(-1, self.pred_len, 2), self.dist_sigma(pol).view(-1, self.pred_len, 2)], -1)]
        lidar2img, ida, fpn, lidar_hi = look_meta
        lidar2img, ida = lidar2img.to(flat.device), ida.to(flat.device)
        mlvl = [getattr(self, f'fpn_linear{i}')(fpn[i]) for i in range(4)]
        shapes, fpn_flat = self.transform_fpn_feats(mlvl)
        fut, cur_bev, cur_flat = None, bev.clone(), flat.clone()
        s_bev, s_flat, s_fut = [], [], []
        keeps = [] if keep is not None else None
        for k, layer in enumerate(self.decoder_layers):
            wp, ctrl = wps[-1].detach(), ctrls[-1].detach()
            kk = {} if keep is not None else None
            t_off, c_off, fut, cur_bev, cur_flat = layer(
                cur_bev, wp, ctrl, fut, parent, self.grid2feat, meas, cur_flat, lidar2img, ida, self.config['img_size'],
                mlvl, fpn_flat, shapes, lidar_hi, self.temporal_embedding, self.static_embedding, dead_work, kk)
            wps.append(t_off.float() + wp.float()); ctrls.append(c_off.float() + ctrl.float())
            s_bev.append(cur_bev); s_flat.append(cur_flat); s_fut.append(fut)
            if keep is not None:
                keeps.append(kk)
        if keep is not None:
            keep.update(layers=keeps, mlvl_feats=mlvl, fpn_flat=fpn_flat)
        K = len(self.decoder_layers)
        B = flat.shape[0]
        o['refine_flattned_BEV_feature'] = torch.stack(s_flat, 1)
        o['refine_BEV_feature'] = torch.stack(s_bev, 1)
        o['refine_future_BEV_feature'] = torch.stack(s_fut, 1).view(B, self.pred_len, K, *bev.shape[1:]).transpose(1, 2)
        o['pred_wp'] = torch.stack(wps, 1)
        c = torch.clamp(F.softplus(torch.stack(ctrls, 1).float()), min=1e-3)
        o['mu_branches'] = c[:, :, 0, :2]; o['sigma_branches'] = c[:, :, 0, 2:]
        o['future_mu'] = c[:, :, 1:, :2]; o['future_sigma'] = c[:, :, 1:, 2:]
        return o
