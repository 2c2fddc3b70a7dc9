"""Multi-scale deformable attention as an autograd Function on libtt_b200.

Mirror of the reference's `MultiScaleDeformableAttnFunction_fp32`
(open_loop_training/code/model_code/dense_heads/multi_scale_deformable_attn_function.py:120-195): same forward arguments, same saved
tensors, same gradient tuple — with `ext_module.ms_deform_attn_forward / _backward` (mmcv._ext) replaced by the C ABI entries
`tt_ms_deform_attn_forward / _backward` (include/tt_b200.h §9).  SURVEY.md §8f row f4.
"""
import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import lib
from ..lib import _p


def _desc(value, spatial_shapes, level_start_index, sampling_locations):
    bs, num_keys, heads, dh = value.shape
    _, nq, _, levels, points, _ = sampling_locations.shape
    d = lib.MsdaDesc()
    d.BN, d.rows_cap, d.heads, d.levels, d.points, d.dh, d.num_keys = bs, nq, heads, levels, points, dh, num_keys
    ss = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]        # host values (mmcv's op reads them on the host as well)
    st = [int(v) for v in level_start_index.tolist()]
    if levels > 4 or len(ss) != levels or len(st) != levels:
        raise lib.TTError('ms_deform_attn: 1..4 levels, one (h, w) and one start index per level')
    d.lvl_h, d.lvl_w, d.lvl_start = lib.i4([h for h, _ in ss]), lib.i4([w for _, w in ss]), lib.i4(st)
    return d


def ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step=None):
    for t in (value, sampling_locations, attention_weights):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise lib.TTError('ms_deform_attn needs fp32 CUDA tensors (no CPU fallback)')
    value, loc, aw = value.contiguous(), sampling_locations.contiguous(), attention_weights.contiguous()
    d = _desc(value, value_spatial_shapes, value_level_start_index, loc)
    out = value.new_empty(d.BN, d.rows_cap, d.heads * d.dh)
    lib.call('tt_ms_deform_attn_forward', C.byref(d), _p(value), _p(loc), _p(aw), _p(out))
    return out


def ms_deform_attn_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, grad_output,
                            grad_value, grad_sampling_loc, grad_attn_weight, im2col_step=None):
    """in-place like mmcv's op: grad_value must arrive zero-filled; the three gradients are written into the given tensors."""
    value, loc, aw = value.contiguous(), sampling_locations.contiguous(), attention_weights.contiguous()
    d = _desc(value, value_spatial_shapes, value_level_start_index, loc)
    for g in (grad_value, grad_sampling_loc, grad_attn_weight):
        assert g.is_contiguous() and g.is_cuda and g.dtype == torch.float32
    lib.call('tt_ms_deform_attn_backward', C.byref(d), _p(value), _p(loc), _p(aw), _p(grad_output.contiguous()), _p(grad_value),
             _p(grad_sampling_loc), _p(grad_attn_weight))


class MultiScaleDeformableAttnFunction_fp32(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        value, sampling_locations, attention_weights = value.float(), sampling_locations.float(), attention_weights.float()   # custom_fwd(cast_inputs=float32)
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value = torch.zeros_like(value)
        grad_sampling_loc = torch.zeros_like(sampling_locations)
        grad_attn_weight = torch.zeros_like(attention_weights)
        ms_deform_attn_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                                grad_output.contiguous(), grad_value, grad_sampling_loc, grad_attn_weight, im2col_step=ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None
