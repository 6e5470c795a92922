"""ctypes binding of libetb200.so (the C ABI declared in include/etb200.h).

There is no CPU fallback: importing works anywhere (so the host logic can be unit-tested), but every
compute call raises if the shared library or a CUDA device is missing.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libetb200.so")

ETB_MAX_LEVELS = 3
ETB_NA = 3
ETB_EMA_CHUNK = 4096

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
vp = C.c_void_p


class EtbEmaChunk(C.Structure):
    _fields_ = [("v", vp), ("m", vp), ("s", vp), ("n", C.c_int32), ("pad_", C.c_int32)]


class EtbSgdChunk(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("buf", vp), ("n", C.c_int32), ("group", C.c_int32)]


class EtbNmsParams(C.Structure):
    _fields_ = [("B", C.c_int32), ("P", C.c_int32), ("no", C.c_int32), ("conf_thres", C.c_float),
                ("iou_thres", C.c_float), ("max_nms", C.c_int32), ("max_det", C.c_int32), ("max_wh", C.c_float),
                ("need_cls_conf", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32)]


class EtbAssignLevels(C.Structure):
    _fields_ = [("nl", C.c_int32), ("nx", C.c_int32 * ETB_MAX_LEVELS), ("ny", C.c_int32 * ETB_MAX_LEVELS),
                ("anchors", (C.c_float * (ETB_NA * 2)) * ETB_MAX_LEVELS), ("anchor_t", C.c_float)]


class EtbAssignOut(C.Structure):
    _fields_ = [("idx", vp * ETB_MAX_LEVELS), ("tbox", vp * ETB_MAX_LEVELS), ("anch", vp * ETB_MAX_LEVELS),
                ("tcls", vp * ETB_MAX_LEVELS), ("tscore", vp * ETB_MAX_LEVELS), ("cnt", vp), ("cap", C.c_int32)]


class EtbLossParams(C.Structure):
    _fields_ = [("nl", C.c_int32), ("B", C.c_int32), ("na", C.c_int32), ("no", C.c_int32),
                ("nx", C.c_int32 * ETB_MAX_LEVELS), ("ny", C.c_int32 * ETB_MAX_LEVELS),
                ("balance", C.c_float * ETB_MAX_LEVELS), ("box_w", C.c_float), ("obj_w", C.c_float),
                ("cls_w", C.c_float), ("cp", C.c_float), ("cn", C.c_float), ("nsets", C.c_int32),
                ("ignore_obj", C.c_int32), ("with_bbox", C.c_int32), ("with_cls", C.c_int32)]


class EtbPackDesc(C.Structure):
    _fields_ = [("w", vp), ("out", vp), ("elems", C.c_int64), ("Cout", C.c_int32), ("Cin", C.c_int32), ("k", C.c_int32),
                ("mode", C.c_int32), ("ntaps", C.c_int32), ("out_ld", C.c_int32), ("kh", C.c_int8 * 12), ("kw", C.c_int8 * 12)]


class EtbFoldDesc(C.Structure):
    _fields_ = [("gamma", vp), ("beta", vp), ("mean", vp), ("var", vp), ("scale", vp), ("bias", vp), ("C", C.c_int32),
                ("eps", C.c_float)]


ETB_PACK_CHUNK = 4096


class EtbFocalParams(C.Structure):
    _fields_ = [("x", vp * ETB_MAX_LEVELS), ("dx", vp * ETB_MAX_LEVELS), ("M", C.c_int64 * ETB_MAX_LEVELS), ("nl", C.c_int32),
                ("label", C.c_int32)]


class EtbV8Levels(C.Structure):
    _fields_ = [("nl", C.c_int32), ("h", C.c_int32 * ETB_MAX_LEVELS), ("w", C.c_int32 * ETB_MAX_LEVELS),
                ("stride", C.c_float * ETB_MAX_LEVELS)]


class EtbConvParams(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
                ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("x_cstride", C.c_int32), ("y_cstride", C.c_int32), ("y_coffset", C.c_int32),
                ("res_cstride", C.c_int32), ("res_coffset", C.c_int32), ("act", C.c_int32), ("det_no", C.c_int32)]


_SIGS = {
    "etb_version": (C.c_int, []),
    "etb_last_error": (C.c_char_p, []),
    "etb_launch_count": (C.c_longlong, []),
    "etb_ema_table_count": (C.c_int64, [C.POINTER(C.c_int64), C.c_int32]),
    "etb_ema_table_fill": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64), C.c_int32,
                                     C.POINTER(EtbEmaChunk), C.c_int64]),
    "etb_ema_update": (C.c_int, [vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, vp]),
    "etb_ema_update_dev": (C.c_int, [vp, C.c_int64, vp, vp]),
    "etb_sgd_step": (C.c_int, [vp, C.c_int64, vp, C.c_int32, vp]),
    "etb_detect_decode": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, c_f32p, C.c_float, vp]),
    "etb_nms_workspace_bytes": (C.c_size_t, [C.POINTER(EtbNmsParams)]),
    "etb_nms_ssod": (C.c_int, [vp, C.POINTER(EtbNmsParams), vp, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    "etb_nms_val_workspace_bytes": (C.c_size_t, [C.POINTER(EtbNmsParams)]),
    "etb_nms_val": (C.c_int, [vp, C.POINTER(EtbNmsParams), vp, vp, vp, C.c_size_t, vp]),
    "etb_select_targets": (C.c_int, [vp, vp, C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32, vp, vp, vp]),
    "etb_build_targets": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.POINTER(EtbAssignLevels),
                                    C.POINTER(EtbAssignOut), vp]),
    "etb_bbox_ciou": (C.c_int, [vp, vp, C.c_int32, vp, vp]),
    "etb_loss_workspace_bytes": (C.c_size_t, [C.POINTER(EtbLossParams), C.c_int32]),
    "etb_loss_forward": (C.c_int, [C.POINTER(vp), C.POINTER(EtbLossParams), C.POINTER(EtbAssignOut), vp, vp,
                                   C.c_size_t, vp]),
    "etb_loss_backward": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(EtbLossParams),
                                    C.POINTER(EtbAssignOut), vp, vp, C.c_size_t, vp]),
    "etb_conv_workspace_bytes": (C.c_size_t, [C.POINTER(EtbConvParams)]),
    "etb_conv_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.POINTER(EtbConvParams), vp, C.c_size_t, vp]),
    "etb_dgrad_weight_elems": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "etb_pack_weight_dgrad": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_conv_dgrad": (C.c_int, [vp, vp, vp, C.POINTER(EtbConvParams), C.c_int32, vp]),
    "etb_conv_wgrad_workspace_bytes": (C.c_size_t, [C.POINTER(EtbConvParams)]),
    "etb_conv_wgrad": (C.c_int, [vp, vp, vp, C.POINTER(EtbConvParams), C.c_int32, vp, C.c_size_t, vp]),
    "etb_bn_partial_rows": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32]),
    "etb_bn_stats": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_int32, vp, C.c_int32, vp]),
    "etb_bn_finalize": (C.c_int, [vp, C.c_int32, C.c_int64, C.c_int32, vp, vp, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp]),
    "etb_bn_act_apply": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_bn_act_apply_res": (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_maxpool5_fwd": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_maxpool5_bwd": (C.c_int, [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_upsample2x_bwd": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_copy_slice_nhwc": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_bn_act_bwd_reduce": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, vp]),
    "etb_bn_act_bwd_finalize": (C.c_int, [vp, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp]),
    "etb_bn_act_bwd_apply": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       vp, vp]),
    "etb_stem_im2col": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp]),
    "etb_nchw_f32_to_nhwc_bf16": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_float, vp]),
    "etb_nhwc_bf16_to_nchw_f32": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_sppf_pool": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_upsample2x_nhwc": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, vp]),
    "etb_fold_bn": (C.c_int, [vp, vp, vp, vp, C.c_float, vp, vp, C.c_int32, vp]),
    "etb_pack_weight": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_pack_multi": (C.c_int, [vp, vp, C.c_int32, vp]),
    "etb_fold_bn_multi": (C.c_int, [vp, C.c_int32, vp]),
    "etb_pack_stem_weight": (C.c_int, [vp, vp, C.c_int32, vp]),
    "etb_detect_dy_rows": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "etb_detect_dy_pack": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "etb_column_sum": (C.c_int, [vp, C.c_int64, C.c_int32, vp, C.c_int32, vp]),
    "etb_netd_tail_rows": (C.c_int32, [C.c_int64]),
    "etb_netd_tail_fwd": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp]),
    "etb_netd_tail_bwd": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp]),
    "etb_domain_focal_workspace_bytes": (C.c_int64, []),
    "etb_domain_focal_fwd": (C.c_int, [C.POINTER(EtbFocalParams), vp, vp, C.c_int64, vp]),
    "etb_domain_focal_bwd": (C.c_int, [C.POINTER(EtbFocalParams), vp, vp]),
    "etb_val_process_batch": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, vp, C.c_int32, vp, vp, vp]),
    "etb_nms_boxes": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp, vp, vp]),
    "etb_bn_fused_rows": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32]),
    "etb_bn_fwd_fused": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, C.c_float, C.c_float, vp, vp, vp, vp, C.c_int32, vp, C.c_int32,
                                  C.c_int32, vp, C.c_int32, vp, vp]),
    "etb_bn_bwd_fused": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_int32, vp,
                                  C.c_int32, vp, vp]),
    "etb_stem_im2col_into": (C.c_int, [vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp]),
    "etb_tal_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "etb_tal_assign": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                C.c_float, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    "etb_v8_decode": (C.c_int, [vp, vp, C.POINTER(EtbV8Levels), C.c_int32, C.c_int32, C.c_int32, C.c_float, vp, vp, vp, vp, vp]),
}

_lib = None


def exported_symbols():
    """Names the header declares (used by the CPU test that checks the .so exports all of them)."""
    return sorted(_SIGS)


def register(name, restype, argtypes):
    _SIGS[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib():
    """Load libetb200.so (no compute is run).  Raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libetb200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU/PyTorch fallback for the B200 kernels)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().etb_last_error()
        raise RuntimeError("libetb200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))


def require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise RuntimeError("efficientteacher_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("expected a CUDA tensor, got %s" % (t.device,))


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
