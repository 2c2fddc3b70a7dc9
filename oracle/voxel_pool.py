"""Oracle: voxel pooling (the reference's only in-tree native op).  TEST INFRASTRUCTURE ONLY.

Restates open_loop_training/ops/voxel_pooling/voxel_pooling.py:10-55 and the
kernel src/voxel_pooling_forward_cuda.cu:9-36: every frustum point whose int
coords fall inside [0,X)x[0,Y)x[0,Z) adds its C-vector to out[b, y, x, :]; the
result is returned as (B, C, Y, X).  The CUDA reference sums with fp32 atomics
in nondeterministic order; here the order is the point order (index_add_).
"""
import numpy as np
import torch


def voxel_pooling_ref(geom_xyz, feats, voxel_num):
    """geom_xyz int32 (B, ..., 3); feats f32 (B, ..., C); voxel_num (3,) [X, Y, Z] -> (B, C, Y, X)."""
    B = feats.shape[0]
    C = feats.shape[-1]
    X, Y, Z = (int(v) for v in voxel_num)
    g = geom_xyz.reshape(B, -1, 3).long()
    f = feats.reshape(B, -1, C)
    out = f.new_zeros(B, Y * X, C)
    for b in range(B):
        x, y, z = g[b, :, 0], g[b, :, 1], g[b, :, 2]
        ok = (x >= 0) & (x < X) & (y >= 0) & (y < Y) & (z >= 0) & (z < Z)
        out[b].index_add_(0, (y * X + x)[ok], f[b][ok])
    return out.view(B, Y, X, C).permute(0, 3, 1, 2)


def voxel_pooling_loops(geom_xyz, feats, voxel_num):
    """Literal per-point loop of the kernel (voxel_pooling_forward_cuda.cu:18-35); small cases only.

    Also returns pos_memo (B, P, 3) = (batch, y, x) or -1, as the kernel records it for backward.
    """
    geom = np.asarray(geom_xyz).reshape(geom_xyz.shape[0], -1, 3)
    f = np.asarray(feats, dtype=np.float32).reshape(geom.shape[0], -1, feats.shape[-1])
    B, P, C = f.shape
    X, Y, Z = (int(v) for v in voxel_num)
    out = np.zeros((B, Y, X, C), np.float32)
    memo = -np.ones((B, P, 3), np.int32)
    for b in range(B):
        for p in range(P):
            x, y, z = (int(v) for v in geom[b, p])
            if x < 0 or x >= X or y < 0 or y >= Y or z < 0 or z >= Z:
                continue
            memo[b, p] = (b, y, x)
            out[b, y, x] += f[b, p]
    return torch.from_numpy(out).permute(0, 3, 1, 2), torch.from_numpy(memo)
