"""voxel_pooling(geom_xyz, input_features, voxel_num) -> (B, C, Y, X).

Same signature, ownership and autograd contract as the reference's
open_loop_training/ops/voxel_pooling/voxel_pooling.py:10-72: int32 contiguous geometry, fp32 contiguous
features, gradient only w.r.t. the features through the recorded `pos_memo`.  The forward goes through the C ABI
(`tt_voxel_pooling_forward`, include/tt_b200.h) instead of the pybind module `voxel_pooling_ext`.
"""
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
        (memo,) = ctx.saved_tensors
        kept = (memo != -1)[..., 0]
        g = grad_out.new_zeros(memo.shape[0], memo.shape[1], grad_out.shape[1])
        mk = memo[kept]
        g[kept] = grad_out[mk[..., 0].long(), :, mk[..., 1].long(), mk[..., 2].long()]
        return None, g.reshape(ctx.in_shape), None


voxel_pooling = VoxelPooling.apply
