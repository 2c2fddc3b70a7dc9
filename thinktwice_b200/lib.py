"""ctypes binding of libtt_b200.so (include/tt_b200.h).

PyTorch is used only as the owner of device memory and streams: every wrapper takes tensors, passes
`data_ptr()` / the current CUDA stream to the C ABI and raises on a non-zero status.  There is no CPU
or eager fallback: if the library is missing or a tensor is not on a CUDA device this fails loudly.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtt_b200.so')

ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID, ACT_SOFTPLUS, ACT_SOFTPLUS_CLAMP = 0, 1, 2, 3, 4, 5
RES_NONE, RES_SAME, RES_UP2 = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TF32, IMPL_3XTF32, IMPL_F16S = 0, 1, 2, 3, 4


class TTError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('N', 'H', 'W', 'Cin', 'x_ld', 'x_coff')] + \
               [('x_nstride', C.c_longlong), ('y_nstride', C.c_longlong), ('x_hstride', C.c_longlong)] + \
               [(n, C.c_int) for n in ('Cout', 'KH', 'KW', 'stride', 'pad', 'dil', 'groups', 'OH', 'OW',
                                        'y_ld', 'y_coff', 'yH', 'yW', 'oy_mul', 'oy_add', 'ox_mul', 'ox_add',
                                        'act', 'bias_n_mod', 'res_mode', 'res_ld', 'res_coff', 'res_H', 'res_W',
                                        'res2_ld', 'res2_coff', 'taps', 'M', 'impl')]


class F16sIO(C.Structure):
    """tt_f16s_io (include/tt_b200.h): the pointers of one tt_conv2d_f16s call."""
    _fields_ = [('x_split', C.c_void_p), ('x_plane', C.c_longlong), ('w_split', C.c_void_p), ('bias', C.c_void_p),
                ('res', C.c_void_p), ('res_split', C.c_void_p), ('res_plane', C.c_longlong),
                ('res2', C.c_void_p), ('res2_split', C.c_void_p), ('res2_plane', C.c_longlong),
                ('y', C.c_void_p), ('y_split', C.c_void_p), ('y_plane', C.c_longlong)]


class LiftSplatDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('B', 'N', 'D', 'fH', 'fW', 'C', 'ld_d', 'd_coff', 'ld_c', 'c_coff')] + \
               [('lower', C.c_float * 3), ('size', C.c_float * 3)] + \
               [(n, C.c_int) for n in ('X', 'Y', 'Z', 'bev_ld', 'bev_coff', 'anti_transpose')]


class VoxelizeDesc(C.Structure):
    _fields_ = [('B', C.c_int), ('P', C.c_int), ('F', C.c_int), ('lower', C.c_float * 3), ('vsize', C.c_float * 3),
                ('grid', C.c_int * 3), ('zmax', C.c_int), ('max_points', C.c_int), ('max_voxels', C.c_int),
                ('cap', C.c_int)]


class RulebookDesc(C.Structure):
    _fields_ = [('B', C.c_int), ('in_shape', C.c_int * 3), ('out_shape', C.c_int * 3), ('k', C.c_int * 3),
                ('s', C.c_int * 3), ('p', C.c_int * 3), ('subm', C.c_int), ('cap_in', C.c_int), ('cap_out', C.c_int),
                ('table_size', C.c_int)]


class SparseConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('Cin', 'Cout', 'kvol', 'in_ld', 'out_ld', 'res_ld', 'cap_out', 'pair_cap', 'act', 'impl')]


class LookDesc(C.Structure):
    _fields_ = [('B', C.c_int), ('num_cams', C.c_int), ('num_query', C.c_int), ('T', C.c_int),
                ('img_w', C.c_float), ('img_h', C.c_float), ('levels', C.c_int),
                ('lvl_h', C.c_int * 4), ('lvl_w', C.c_int * 4), ('C', C.c_int), ('q_dim', C.c_int),
                ('emb_dim', C.c_int), ('meas_dim', C.c_int), ('flat_dim', C.c_int), ('max_len_cap', C.c_int)]


class MsdaDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('BN', 'rows_cap', 'heads', 'levels', 'points', 'dh')] + \
               [('lvl_h', C.c_int * 4), ('lvl_w', C.c_int * 4), ('lvl_start', C.c_int * 4), ('num_keys', C.c_int), ('value_ld', C.c_int),
                ('value_coff', C.c_int)]


class PreprocDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('n_img', 'H', 'W', 'newH', 'newW', 'outH', 'outW', 'crop_x', 'crop_y', 'undistort')] + \
               [('div', C.c_float), ('mean', C.c_float * 3), ('std', C.c_float * 3)] + \
               [(n, C.c_int) for n in ('pad_H', 'pad_W', 'pad_top', 'pad_left')]


_lib = None


def load():
    """dlopen the in-tree library; raises TTError when it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TTError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                      '(nvcc, sm_100a).  thinktwice_b200 has no CPU / eager fallback.')
    lib = C.CDLL(LIB_PATH)
    lib.tt_last_error.restype = C.c_char_p
    lib.tt_launch_count.restype = C.c_longlong
    for n in ('tt_voxel_pooling_workspace_bytes', 'tt_lift_splat_workspace_bytes', 'tt_voxelize_workspace_bytes',
              'tt_rulebook_workspace_bytes', 'tt_conv2d_workspace_bytes'):
        getattr(lib, n).restype = C.c_size_t
    if os.environ.get('TT_DEBUG'):                               # kernel-experiment knobs (tt_debug_set) for whole test / bench runs
        lib.tt_debug_set(int(os.environ['TT_DEBUG'], 0))
    _lib = lib
    return lib


EXPORTS = [
    'tt_version', 'tt_last_error', 'tt_launch_count', 'tt_voxel_pooling_workspace_bytes', 'tt_voxel_pooling_forward',
    'tt_lift_splat_workspace_bytes', 'tt_lift_splat', 'tt_conv2d', 'tt_conv2d_workspace_bytes', 'tt_nchw_to_nhwc', 'tt_nchw_to_nhwc_padded', 'tt_nhwc_to_nchw',
    'tt_maxpool3x3s2', 'tt_upsample2x_bilinear_ac', 'tt_global_avgpool', 'tt_broadcast_rows', 'tt_se_gate', 'tt_se_pool',
    'tt_se_apply', 'tt_anti_transpose', 'tt_copy2d', 'tt_layernorm', 'tt_eltwise', 'tt_fill', 'tt_dcn_im2col',
    'tt_voxelize_workspace_bytes', 'tt_voxelize_mean', 'tt_rulebook_workspace_bytes', 'tt_sparse_rulebook',
    'tt_sparse_conv', 'tt_sparse_to_bev', 'tt_conv2d_f16s', 'tt_split_f16', 'tt_merge_f16', 'tt_image_to_split8', 'tt_pointwise_f16s', 'tt_upsample2x_bilinear_ac_split', 'tt_f16s_saturation_count', 'tt_sparse_conv_f16s', 'tt_sparse_conv_os_f16s', 'tt_look_project', 'tt_look_rebatch', 'tt_msda_forward', 'tt_look_reduce', 'tt_gru_input',
    'tt_preprocess_u8', 'tt_lidar_stitch', 'tt_points_union', 'tt_voxel_pooling_backward', 'tt_ms_deform_attn_forward', 'tt_ms_deform_attn_backward',
]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, off=0):
    """device pointer of tensor `t` advanced by `off` elements (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise TTError('tt_b200 ops need CUDA tensors (there is no CPU fallback)')
    return C.c_void_p(t.data_ptr() + off * t.element_size())


def ref(struct):
    """pointer to a ctypes struct for a C-ABI call."""
    return C.byref(struct)


def require_cuda(dev):
    if torch.device(dev).type != 'cuda':
        raise TTError('thinktwice_b200 runs on a CUDA device only (no CPU fallback)')


def check(rc, name):
    if rc != 0:
        raise TTError(f'{name} failed ({rc}): {load().tt_last_error().decode()}')


def launch_count():
    return int(load().tt_launch_count())


def f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def i4(v):
    v = list(v) + [0] * (4 - len(v))
    return (C.c_int * 4)(*[int(x) for x in v])


# ---------------------------------------------------------------------------------------------
# thin typed wrappers (pointer arithmetic for channel offsets is done by the callers through `_p`)
# ---------------------------------------------------------------------------------------------
def call(name, *args):
    check(getattr(load(), name)(*args, _stream()), name)
