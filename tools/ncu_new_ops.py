#!/usr/bin/env python
"""`ncu --set full` rows for the kernels of the "next" rows, at the sizes one closed-loop tick / one decoder layer
of the thinktwice.py config gives them:

  ncu --set full --clock-control none -o /tmp/r2_new_ops python tools/ncu_new_ops.py
  python tools/ncu_summary.py /tmp/r2_new_ops.ncu-rep gpurun_out/r2_kernels_new

  preprocess_u8_kernel : 8 frames (2 ticks x 4 cameras) 900 x 1600 uint8 -> fp32 NCHW, and -> the stem's operand planes
  lidar_stitch_kernel  : 2 x 20 000 points (and points_union_kernel: the same two clouds as a two-frame queue)
  voxel_pool_bwd_kernel: the lift-splat frustum of one sweep (4 x 80 x 28 x 56 points, 80 channels)
  msda_fwd / msda_bwd  : the Look module's attention of one decoder layer (4 cameras, 53 queries, 8 heads x 32, 4 FPN levels)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from thinktwice_b200.ops.ms_deform_attn import MultiScaleDeformableAttnFunction_fp32
    from thinktwice_b200.ops.voxel_pooling import voxel_pooling
    from thinktwice_b200.preprocess import AgentPreprocessor, union_metas
    conf = {'final_dim': (448, 896), 'H': 900, 'W': 1600, 'bot_pct_lim': (0.0, 0.0)}
    pre = AgentPreprocessor(dict(undistort=True, num_cams=4), conf, 'cuda:0')
    raw = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(2, 4, 900, 1600, 3), dtype=np.uint8)).cuda()
    g = torch.Generator().manual_seed(0)
    planes = torch.zeros(2, 8 * 454 * 904 * 8, dtype=torch.float16, device='cuda')
    prev, now = torch.randn(20000, 4, generator=g).cuda(), torch.randn(20000, 4, generator=g).cuda()
    rel = AgentPreprocessor.relative_matrix((1.0, 2.0, 0.3), (1.5, 2.2, 0.31))
    can = [np.zeros(18), np.zeros(18)]
    can[1][:2], can[1][-2] = (1.5, 0.4), 0.05
    metas = union_metas(can, torch.eye(4).expand(4, 4, 4).contiguous())
    B, P, Cc, X, Y = 1, 4 * 80 * 28 * 56, 80, 21, 21
    geom = torch.stack([torch.randint(-5, X + 5, (B, P), generator=g), torch.randint(-5, Y + 5, (B, P), generator=g),
                        torch.zeros(B, P, dtype=torch.long)], -1).int().cuda()
    feats = torch.randn(B, P, Cc, generator=g).cuda().requires_grad_(True)
    shapes = [(112, 224), (56, 112), (28, 56), (14, 28)]
    starts, k = [], 0
    for h, w in shapes:
        starts.append(k)
        k += h * w
    value = torch.randn(4, k, 8, 32, generator=g).cuda().requires_grad_(True)
    loc = torch.rand(4, 53, 8, 4, 8, 2, generator=g).cuda().requires_grad_(True)
    aw = torch.rand(4, 53, 8, 4, 8, generator=g).cuda().requires_grad_(True)
    ss, st = torch.tensor(shapes, device='cuda'), torch.tensor(starts, device='cuda')

    def step():
        pre.images(raw)
        pre.images_to_stem(raw.view(8, 900, 1600, 3), planes, planes.numel() // 2, (454, 904, 3, 3))
        pre.stitch_lidar(prev, now, rel)
        pre.union_points([prev, now], metas)
        out = voxel_pooling(geom, feats, torch.tensor([X, Y, 1]))
        out.backward(torch.ones_like(out))
        o = MultiScaleDeformableAttnFunction_fp32.apply(value, ss, st, loc, aw, 64)
        o.backward(torch.ones_like(o))
    step()                                                         # warm-up (allocations, module load) outside the profiled range
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print('done')


if __name__ == '__main__':
    main()
