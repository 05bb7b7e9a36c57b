"""ctypes binding of libiiseg.so (the C ABI declared in include/iiseg.h).

There is no CPU fallback: if the library cannot be loaded, or a call fails, this
module raises.  The library is built in-tree by csrc/build.py (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libiiseg.so')


class IisegError(RuntimeError):
    pass


ABI_VERSION = 5      # IISEG_ABI_VERSION
MAX_SRC = 6          # IISEG_MAX_SRC
MAX_WGROUPS = 9      # IISEG_MAX_WGROUPS


class ConvDesc(C.Structure):
    """struct iiseg_conv_desc (include/iiseg.h)."""
    _fields_ = [
        ('src', C.c_void_p * MAX_SRC), ('C', C.c_int * MAX_SRC), ('Cs', C.c_int * MAX_SRC),
        ('N', C.c_int), ('H', C.c_int), ('W', C.c_int),
        ('src_image_stride', C.c_longlong), ('weight_ld', C.c_longlong), ('w_koff', C.c_int), ('w_groups', C.c_int), ('w_rows_total', C.c_int), ('w_group_koff', C.c_int * MAX_WGROUPS),
        ('w_group_row', C.c_int * MAX_WGROUPS),
        ('depool_mask', C.c_void_p), ('depool_UH', C.c_int), ('depool_UW', C.c_int), ('depool_h0', C.c_int), ('depool_w0', C.c_int),
        ('depool_out', C.c_void_p), ('depool_out_mask', C.c_void_p), ('depool_out_VH', C.c_int), ('depool_out_VW', C.c_int),
        ('depool_out_h0', C.c_int), ('depool_out_w0', C.c_int),
        ('depool_out_H2', C.c_int), ('depool_out_W2', C.c_int), ('depool_out_ph0', C.c_int), ('depool_out_pw0', C.c_int),
        ('weight', C.c_void_p), ('weight_npack', C.c_void_p), ('bias', C.c_void_p), ('post_scale', C.c_void_p), ('post_shift', C.c_void_p), ('post_mean', C.c_void_p),
        ('Cout', C.c_int), ('R', C.c_int), ('S', C.c_int), ('pad', C.c_int),
        ('oh0', C.c_int), ('ow0', C.c_int), ('OH', C.c_int), ('OW', C.c_int),
        ('out', C.c_void_p), ('out_stride', C.c_int), ('out_H', C.c_int), ('out_W', C.c_int), ('out_h0', C.c_int), ('out_w0', C.c_int),
        ('addend', C.c_void_p),
        ('AH', C.c_int), ('AW', C.c_int), ('ah0', C.c_int), ('aw0', C.c_int), ('addend_f32', C.c_int), ('addend_cs', C.c_int),
        ('pooled', C.c_void_p), ('pool_mask', C.c_void_p), ('pool_zmask', C.c_void_p), ('pool_H', C.c_int), ('pool_W', C.c_int),
        ('relu', C.c_int), ('split', C.c_int), ('out_f32', C.c_int), ('out_cs', C.c_int),
        ('upd_y', C.c_void_p), ('upd_y_bf16', C.c_void_p), ('upd_active', C.c_void_p), ('upd_norm_acc', C.c_void_p),
        ('upd_step', C.c_float), ('upd_step_dev', C.c_void_p), ('upd_C', C.c_int), ('upd_split', C.c_int), ('upd_cpad', C.c_int),
    ]


class DeconvDesc(C.Structure):
    """struct iiseg_deconv_desc (include/iiseg.h)."""
    _fields_ = [
        ('x', C.c_void_p), ('N', C.c_int), ('H', C.c_int), ('W', C.c_int),
        ('weight', C.c_void_p), ('bias', C.c_void_p), ('k', C.c_int), ('stride', C.c_int),
        ('oh0', C.c_int), ('ow0', C.c_int), ('OH', C.c_int), ('OW', C.c_int),
        ('addend', C.c_void_p), ('AH', C.c_int), ('AW', C.c_int), ('ah0', C.c_int), ('aw0', C.c_int),
        ('out', C.c_void_p),
    ]


class CtxConvDesc(C.Structure):
    """struct iiseg_ctx_conv_desc (include/iiseg.h)."""
    _fields_ = [
        ('in_', C.c_void_p), ('N', C.c_int), ('Cin', C.c_int), ('Hin', C.c_int), ('Win', C.c_int),
        ('in_h0', C.c_int), ('in_w0', C.c_int), ('check', C.c_int), ('dil', C.c_int),
        ('out', C.c_void_p), ('Cout', C.c_int), ('Hout', C.c_int), ('Wout', C.c_int), ('out_h0', C.c_int), ('out_w0', C.c_int),
        ('OH', C.c_int), ('OW', C.c_int),
        ('weight', C.c_void_p), ('bias', C.c_void_p), ('addend', C.c_void_p), ('active', C.c_void_p), ('relu', C.c_int),
        ('weight2', C.c_void_p), ('bias2', C.c_void_p), ('C2', C.c_int), ('in_nhwc', C.c_int), ('out_nhwc', C.c_int), ('stream', C.c_void_p),
    ]


_vp, _i, _f = C.c_void_p, C.c_int, C.c_float

# name -> (restype, argtypes); must list every symbol include/iiseg.h declares
SIGNATURES = {
    'iiseg_abi_version': (_i, []),
    'iiseg_conv_desc_size': (_i, []),
    'iiseg_conv_desc_last_offset': (_i, []),
    'iiseg_deconv_desc_size': (_i, []),
    'iiseg_last_error': (C.c_char_p, []),
    'iiseg_device_check': (_i, [_i]),
    'iiseg_read_diag': (_i, [_vp, _i]),
    'iiseg_launch_count': (C.c_int64, []),
    'iiseg_reserve_sms': (_i, [_i]),
    'iiseg_debug_read_timeline': (_i, [_vp, _i]),
    'iiseg_pack_nchw_f32_to_nhwc_bf16': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'iiseg_unpack_nhwc_bf16_to_nchw_f32': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'iiseg_unpack_nhwc_f32_to_nchw_f32': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'iiseg_widen_nhwc_bf16_to_f32': (_i, [_vp, _vp, C.c_longlong, _i, _i, _vp]),
    'iiseg_conv2d_fwd': (_i, [C.POINTER(ConvDesc), _vp]),
    'iiseg_maxpool2_mask_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'iiseg_unpool2_mask_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'iiseg_unpool2_mask_window_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'iiseg_deconv2d_fwd': (_i, [C.POINTER(DeconvDesc), _vp]),
    'iiseg_ctx_conv_desc_size': (_i, []),
    'iiseg_ctx_conv': (_i, [C.POINTER(CtxConvDesc)]),
    'iiseg_update_blocks': (_i, [_i, _i]),
    'iiseg_softmax_nchw': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'iiseg_softmax_update': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _i, _vp]),
    'iiseg_softmax_grad': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'iiseg_norm_finalize': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    'iiseg_norm_finalize_fixed': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    'iiseg_onehot_to_labels': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'iiseg_bn_relu_pack': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    'iiseg_channel_stats_chunks': (_i, [_i, _i, _i]),
    'iiseg_channel_stats': (_i, [_vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    'iiseg_maxpool2_f32': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    'iiseg_deconv_interleave': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    'iiseg_noise_pack': (_i, [_vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'iiseg_loss_grad': (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp]),
    'iiseg_sq_sum': (_i, [_vp, C.c_longlong, _vp, _vp]),
    'iiseg_broadcast_image': (_i, [_vp, C.c_longlong, _i, _vp]),
    'iiseg_add_bf16': (_i, [_vp, _vp, _vp, C.c_longlong, _vp]),
    'iiseg_ae_grad_add': (_i, [_vp, _vp, C.c_longlong, _vp, _vp]),
    'iiseg_loss_grad_terms': (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _i, _vp]),
    'iiseg_depool2_bwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'iiseg_pool2_relu_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'iiseg_transpose_shift': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, C.c_longlong, C.c_longlong, _i, C.c_longlong, _vp]),
    'iiseg_last_conv_plan': (_i, [_vp, _vp, _vp]),
    'iiseg_bias_grad': (_i, [_vp, C.c_longlong, _i, _vp, _i, _vp, _i, _vp]),
    'iiseg_sum_slabs': (_i, [_vp, _vp, _i, C.c_longlong, _vp]),
    'iiseg_rmsprop_pack': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _f, _vp]),
    'iiseg_adam_pack': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _f, _f, _f, _vp]),
    'iiseg_adam_advance': (_i, [_vp, _f, _f, _f, _vp]),
    'iiseg_metrics_accumulate': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
}

_lib = None


def load():
    """Load libiiseg.so and bind every declared symbol.  Raises if missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IisegError('%s not found: build it with `python -m iterative_inference_segm_b200.csrc.build` '
                             '(there is no CPU fallback)' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.iiseg_abi_version() != ABI_VERSION:
            raise IisegError('libiiseg ABI version mismatch (library %d, binding %d): rebuild with '
                             '`python -m iterative_inference_segm_b200.csrc.build --force`' % (lib.iiseg_abi_version(), ABI_VERSION))
        # the hand-mirrored descriptor structs must have the C layout: a stale library or an edited struct would otherwise
        # hand misaligned descriptors to the kernels
        if (C.sizeof(ConvDesc) != lib.iiseg_conv_desc_size() or ConvDesc.upd_cpad.offset != lib.iiseg_conv_desc_last_offset()
                or C.sizeof(DeconvDesc) != lib.iiseg_deconv_desc_size()
                or C.sizeof(CtxConvDesc) != lib.iiseg_ctx_conv_desc_size()):
            raise IisegError('ctypes mirror of iiseg_conv_desc / iiseg_deconv_desc does not match libiiseg.so '
                             '(conv %d vs %d bytes, last field at %d vs %d, deconv %d vs %d bytes): rebuild the library'
                             % (C.sizeof(ConvDesc), lib.iiseg_conv_desc_size(), ConvDesc.upd_cpad.offset,
                                lib.iiseg_conv_desc_last_offset(), C.sizeof(DeconvDesc), lib.iiseg_deconv_desc_size()))
        _lib = lib
    return _lib


def call(name, *args):
    """Call an int-status entry point; raise IisegError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise IisegError('%s failed (%d): %s' % (name, rc, lib.iiseg_last_error().decode()))


def read_diag():
    lib = load()
    buf = (C.c_int32 * 8)()
    n = lib.iiseg_read_diag(buf, 8)
    return list(buf[:n])


def launch_count():
    return int(load().iiseg_launch_count())
