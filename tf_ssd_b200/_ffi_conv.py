"""ctypes declarations of the network-forward entry points of ``include/ssd_b200.h``
(``ssd_conv2d`` and the HBM-bound layer kernels)."""

from __future__ import annotations

import ctypes as C

i, f, vp, i64 = C.c_int, C.c_float, C.c_void_p, C.c_int64

ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2


class ConvDesc(C.Structure):
    """``struct ssd_conv_desc`` (include/ssd_b200.h)."""
    _fields_ = [
        ("inp", vp), ("weight", vp), ("bias", vp), ("residual", vp), ("out0", vp), ("out1", vp),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
        ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("dilation", C.c_int32),
        ("pad_top", C.c_int32), ("pad_left", C.c_int32),
        ("act", C.c_int32), ("out_f32", C.c_int32), ("split", C.c_int32), ("reserved", C.c_int32),
        ("img_stride0", i64), ("pix_stride0", i64), ("img_stride1", i64), ("pix_stride1", i64),
    ]


class DwProjDesc(C.Structure):
    """``struct ssd_dwproj_desc`` (include/ssd_b200.h)."""
    _fields_ = [
        ("inp", vp), ("dw_weight", vp), ("dw_bias", vp), ("proj_weight", vp), ("proj_bias", vp), ("residual", vp), ("out", vp),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("Cout", C.c_int32), ("stride", C.c_int32), ("pad_top", C.c_int32), ("pad_left", C.c_int32),
        ("dw_act", C.c_int32), ("act", C.c_int32), ("reserved", C.c_int32),
    ]


class IrBlockDesc(C.Structure):
    """``struct ssd_irblock_desc`` (include/ssd_b200.h)."""
    _fields_ = [
        ("inp", vp), ("exp_weight", vp), ("exp_bias", vp), ("dw_weight", vp), ("dw_bias", vp), ("proj_weight", vp),
        ("proj_bias", vp), ("residual", vp), ("out", vp),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cexp", C.c_int32), ("Ho", C.c_int32),
        ("Wo", C.c_int32), ("Cout", C.c_int32), ("stride", C.c_int32), ("pad_top", C.c_int32), ("pad_left", C.c_int32),
        ("exp_act", C.c_int32), ("dw_act", C.c_int32), ("act", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
    ]


class StemDwProjDesc(C.Structure):
    """``struct ssd_stem_dwproj_desc`` (include/ssd_b200.h)."""
    _fields_ = [
        ("image", vp), ("stem_weight", vp), ("stem_bias", vp), ("dw_weight", vp), ("dw_bias", vp), ("proj_weight", vp),
        ("proj_bias", vp), ("out", vp),
        ("image_u8", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32),
        ("Cmid", C.c_int32), ("Cout", C.c_int32), ("pad_top", C.c_int32), ("pad_left", C.c_int32),
        ("stem_act", C.c_int32), ("dw_act", C.c_int32), ("act", C.c_int32), ("reserved", C.c_int32),
    ]


class AdamVar(C.Structure):
    """``struct ssd_adam_var`` (include/ssd_b200.h)."""
    _fields_ = [("w", vp), ("m", vp), ("v", vp), ("grad", vp), ("w16", vp), ("n", i64), ("l2", f), ("reserved", f)]


SIGNATURES = {
    "ssd_conv2d": (i, [C.POINTER(ConvDesc), vp]),
    "ssd_conv_chain": (i, [C.POINTER(ConvDesc), C.POINTER(C.c_int32), i, vp]),
    "ssd_conv_chain_supported": (i, [C.POINTER(ConvDesc), C.POINTER(C.c_int32), i]),
    "ssd_conv_chain_waves": (i, [C.POINTER(ConvDesc), C.POINTER(C.c_int32), i]),
    "ssd_depthwise3x3": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_dwproj": (i, [C.POINTER(DwProjDesc), vp]),
    "ssd_dwproj_supported": (i, [C.POINTER(DwProjDesc)]),
    "ssd_stem_dwproj": (i, [C.POINTER(StemDwProjDesc), vp]),
    "ssd_stem_dwproj_supported": (i, [C.POINTER(StemDwProjDesc)]),
    "ssd_irblock": (i, [C.POINTER(IrBlockDesc), vp]),
    "ssd_irblock_supported": (i, [C.POINTER(IrBlockDesc)]),
    "ssd_irblock_trace": (i, [vp]),
    "ssd_debug_irblock_mode": (i, [i]),
    "ssd_set_pdl": (i, [i]),
    "ssd_irblock_plan": (i, [C.POINTER(IrBlockDesc), C.POINTER(C.c_int32)]),
    "ssd_debug_trace": (i, [vp]),
    "ssd_debug_pair_mode": (i, [i]),
    "ssd_stem_conv3x3s2": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_stem_conv3x3s2_u8": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_stem_conv3x3": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_stem_conv3x3_u8": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_stem_conv3x3_f16c8": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_image_to_f16c8": (i, [vp, vp, i64, vp]),
    "ssd_image_u8_to_f16c8": (i, [vp, vp, i64, vp]),
    "ssd_preprocess_image": (i, [vp, i, i, vp, i, i, i, vp]),
    "ssd_flip_boxes": (i, [vp, i, vp]),
    "ssd_augment_workspace_bytes": (C.c_size_t, [i, i, i, i, i]),
    "ssd_augment_batch": (i, [vp, vp, vp, i, i, i, i, i, i, vp, vp, C.c_size_t, vp]),
    "ssd_maxpool": (i, [vp, vp, i, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_l2norm": (i, [vp, vp, vp, i64, i, vp]),
    # training
    "ssd_conv2d_wgrad": (i, [C.POINTER(ConvDesc), vp, i, vp, vp]),
    "ssd_relu_bwd": (i, [vp, vp, i64, vp]),
    "ssd_bias_grad": (i, [vp, vp, i64, i, i, vp]),
    "ssd_filter_flip_transpose": (i, [vp, vp, i, i, i, i, i, vp]),
    "ssd_upsample_zero": (i, [vp, vp, i, i, i, i, i, i, i, vp]),
    "ssd_maxpool_bwd": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_l2norm_bwd": (i, [vp, vp, vp, vp, vp, i64, i, i, vp]),
    "ssd_head_grad_gather": (i, [vp, vp, vp, i, i, i, i, i, i, i, vp]),
    "ssd_adam_step": (i, [vp, vp, vp, vp, vp, i64, f, f, f, f, f, f, vp, vp]),
    "ssd_adam_step_multi": (i, [vp, i, i64, f, f, f, f, f, vp, vp]),
    "ssd_grad_nonfinite_multi": (i, [vp, i, i64, vp, vp]),
    "ssd_adam_step_multi_guarded": (i, [vp, i, i64, f, f, f, f, f, vp, vp, vp]),
    "ssd_bn_workspace_bytes": (C.c_size_t, [i]),
    "ssd_bn_train_fwd": (i, [vp, vp, vp, vp, vp, i64, i, f, f, i, vp, vp, vp, vp, C.c_size_t, vp]),
    "ssd_bn_train_bwd": (i, [vp, vp, vp, vp, vp, i64, i, i, vp, vp, i, vp, vp, vp, C.c_size_t, vp]),
    "ssd_depthwise3x3_dgrad": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp]),
    "ssd_depthwise3x3_wgrad": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]),
}
