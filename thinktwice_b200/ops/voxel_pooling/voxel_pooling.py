"""voxel_pooling(geom_xyz, input_features, voxel_num) -> (B, C, Y, X).

Same signature, ownership and autograd contract as the reference's
open_loop_training/ops/voxel_pooling/voxel_pooling.py:10-72: int32 contiguous geometry, fp32 contiguous
features, gradient only w.r.t. the features through the recorded `pos_memo`.  The forward goes through the C ABI
(`tt_voxel_pooling_forward`, include/tt_b200.h) instead of the pybind module `voxel_pooling_ext`.
"""
import ctypes as C

import torch
from torch.autograd import Function

from ... import lib
from ...lib import _p

_last_memo = None


def last_pos_memo():
    return _last_memo


class VoxelPooling(Function):
    @staticmethod
    def forward(ctx, geom_xyz, input_features, voxel_num):
        global _last_memo
        assert geom_xyz.is_contiguous() and input_features.is_contiguous()
        if not (geom_xyz.is_cuda and input_features.is_cuda):
            raise lib.TTError('voxel_pooling needs CUDA tensors (no CPU fallback)')
        ctx.mark_non_differentiable(geom_xyz)
        vn = [int(v) for v in voxel_num]                       # host values (the reference syncs here too, :37-38)
        shape = input_features.shape
        geom = geom_xyz.reshape(geom_xyz.shape[0], -1, geom_xyz.shape[-1])
        feats = input_features.reshape(geom.shape[0], -1, input_features.shape[-1])
        assert geom.shape[1] == feats.shape[1] and geom.dtype == torch.int32 and feats.dtype == torch.float32
        B, P, Cc = feats.shape
        out = feats.new_zeros(B, vn[1], vn[0], Cc)
        memo = geom.new_ones(B, P, 3) * -1
        lib.call('tt_voxel_pooling_forward', B, P, Cc, vn[0], vn[1], vn[2], _p(geom), _p(feats), _p(out), _p(memo), None)
        ctx.save_for_backward(memo)
        ctx.in_shape = shape
        _last_memo = memo
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad_out):
        # voxel_pooling.py:57-69 (boolean-mask indexing on the host side of torch) as one gather kernel over the recorded positions
        (memo,) = ctx.saved_tensors
        if not grad_out.is_cuda or grad_out.dtype != torch.float32:
            raise lib.TTError('voxel_pooling backward needs an fp32 CUDA gradient (no CPU fallback)')
        B, P = memo.shape[:2]
        Cc = grad_out.shape[1]
        g = grad_out.new_empty(B, P, Cc)
        sb, sc, sy, sx = grad_out.stride()
        lib.call('tt_voxel_pooling_backward', B, P, Cc, _p(grad_out), C.c_longlong(sb), C.c_longlong(sc), C.c_longlong(sy), C.c_longlong(sx),
                 _p(memo), _p(g))
        return None, g.reshape(ctx.in_shape), None


voxel_pooling = VoxelPooling.apply
