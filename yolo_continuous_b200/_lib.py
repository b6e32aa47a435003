"""ctypes binding of include/yc_b200.h.  PyTorch is used for device memory and streams only."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YC_LIB_PATH") or os.path.join(HERE, "csrc", "libyc_b200.so")  # env override: kernel experiments

YC_MAX_LEVELS = 4
YC_MAX_ANCHORS = 4
YC_F32, YC_BF16 = 0, 1
YC_HEAD_IDETECT, YC_HEAD_IBIN, YC_HEAD_RAW = 0, 1, 2
YC_PATH_AUTO, YC_PATH_TCGEN05, YC_PATH_GENERIC = 0, 1, 2
YC_ERR_UNSUPPORTED = -2


class HeadLevel(C.Structure):
    _fields_ = [("x", C.c_void_p), ("blob", C.c_void_p), ("raw", C.c_void_p),
                ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("stride", C.c_float), ("anchor_wh", C.c_float * (YC_MAX_ANCHORS * 2)),
                ("stride_y", C.c_float)]


class HeadDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("path", C.c_int32), ("x_dtype", C.c_int32),
                ("nl", C.c_int32), ("na", C.c_int32), ("no", C.c_int32),
                ("bin_count", C.c_int32), ("bs", C.c_int32),
                ("level", HeadLevel * YC_MAX_LEVELS),
                ("z", C.c_void_p), ("bins", C.c_void_p), ("x_channels_last", C.c_int32)]


class NmsParams(C.Structure):
    _fields_ = [("bs", C.c_int32), ("rows", C.c_int32), ("row_stride", C.c_int32), ("nc", C.c_int32),
                ("conf_thres", C.c_float), ("nms_thres", C.c_double),
                ("write_corners", C.c_int32), ("correct_boxes", C.c_int32), ("letterbox", C.c_int32),
                ("input_h", C.c_int32), ("input_w", C.c_int32),
                ("image_hw", C.c_void_p), ("image_hw_stride", C.c_int32),
                ("box_div_w", C.c_float), ("box_div_h", C.c_float)]


EXPORTS = ["yc_last_error", "yc_version", "yc_device_check", "yc_head_pack_bytes", "yc_head_pack",
           "yc_head_forward", "yc_decode_box", "yc_nms_workspace_bytes", "yc_nms_batched",
           "yc_nms_single", "yc_box_iou", "yc_cvt_bbox", "yc_detect_fused",
           "yc_detect_fused_head", "yc_nms_from_candidates", "yc_nms_workspace_reset", "yc_detect_fused_head_noreset",
           "yc_letterbox_batch", "yc_format_detections", "yc_reserve_sms", "yc_copy_async",
           "yc_xchg_bytes", "yc_xchg_alloc", "yc_xchg_open", "yc_xchg_close", "yc_xchg_free", "yc_xchg_push", "yc_xchg_wait",
           "yc_xchg_state", "yc_match_detections"]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m yolo_continuous_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / eager fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.yc_last_error.restype = C.c_char_p
    lib.yc_version.restype = C.c_int
    lib.yc_device_check.argtypes = [C.c_int]
    lib.yc_head_pack_bytes.restype = C.c_size_t
    lib.yc_head_pack_bytes.argtypes = [C.c_int, C.c_int]
    lib.yc_head_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]
    lib.yc_head_forward.argtypes = [C.POINTER(HeadDesc), C.c_void_p]
    lib.yc_decode_box.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    lib.yc_nms_workspace_bytes.restype = C.c_size_t
    lib.yc_nms_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.yc_nms_batched.argtypes = [C.c_void_p, C.POINTER(NmsParams), C.c_void_p, C.c_size_t, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.yc_detect_fused.argtypes = [C.POINTER(HeadDesc), C.POINTER(NmsParams), C.c_void_p, C.c_size_t, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.yc_detect_fused_head.argtypes = [C.POINTER(HeadDesc), C.POINTER(NmsParams), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.yc_detect_fused_head_noreset.argtypes = lib.yc_detect_fused_head.argtypes
    lib.yc_nms_workspace_reset.argtypes = [C.POINTER(NmsParams), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.yc_nms_from_candidates.argtypes = [C.POINTER(NmsParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    lib.yc_reserve_sms.argtypes = [C.c_int]
    lib.yc_copy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.yc_xchg_bytes.restype = C.c_size_t
    lib.yc_xchg_bytes.argtypes = [C.c_int, C.c_int, C.c_size_t]
    lib.yc_xchg_alloc.argtypes = [C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    lib.yc_xchg_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.yc_xchg_close.argtypes = [C.c_void_p]
    lib.yc_xchg_free.argtypes = [C.c_void_p]
    lib.yc_xchg_push.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                 C.c_void_p]
    lib.yc_xchg_wait.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p]
    lib.yc_xchg_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.yc_letterbox_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.yc_format_detections.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
    lib.yc_nms_single.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.yc_box_iou.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.yc_match_detections.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.yc_cvt_bbox.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


lib = _load()


class YcError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib.yc_last_error().decode("utf-8", "replace")
        raise YcError(f"{what} failed with status {rc}: {msg}")


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise YcError(f"{name} must be a CUDA tensor: this path has no CPU fallback "
                      f"(got device {t.device})")
