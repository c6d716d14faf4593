"""ctypes binding of libb200det.so (the C ABI declared in include/b200det.h).

There is no CPU implementation and no fallback: if the shared library is
missing or a tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200det.so")

B200_LAYOUT_NCHW = 0
B200_LAYOUT_NHWC = 1
B200_MAX_LEVELS = 8
B200_NMS_MAX_SEG = 16384
B200_ERR_UNSUPPORTED = -4
B200_MATCH_SOFTMAX = 0
B200_MATCH_COLMAX = 1


class b200_level(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("height", ctypes.c_int32),
                ("width", ctypes.c_int32), ("spatial_scale", ctypes.c_float)]


class b200_rpn_level(ctypes.Structure):
    _fields_ = [("objectness", ctypes.c_void_p), ("box_regression", ctypes.c_void_p), ("anchors", ctypes.c_void_p),
                ("num_anchors", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("anchors_per_image", ctypes.c_int32)]


_lib = None

_vp, _i, _i64, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/b200det.h declares
SIGNATURES = {
    "b200_version": (_i, []),
    "b200_last_error_string": (ctypes.c_char_p, []),
    "b200_roi_align_forward": (_i, [ctypes.POINTER(b200_level), _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "b200_roi_align_forward_fast": (_i, [ctypes.POINTER(b200_level), _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "b200_roi_align_forward_ex": (_i, [ctypes.POINTER(b200_level), _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b200_roi_align_workspace_bytes": (ctypes.c_size_t, [_i64]),
    "b200_roi_align_forward_ws": (_i, [ctypes.POINTER(b200_level), _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "b200_roi_align_forward_bf16": (_i, [ctypes.POINTER(b200_level), _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "b200_roi_align_backward": (_i, [ctypes.POINTER(b200_level), _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "b200_roi_align_backward_ws": (_i, [ctypes.POINTER(b200_level), _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp, _vp, ctypes.c_size_t, _vp]),
    "b200_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "b200_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "b200_nms_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "b200_nms_batched": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _f, _i64, _vp, _vp, _vp, _sz, _vp]),
    "b200_rpn_candidates": (_i, [ctypes.POINTER(b200_rpn_level), _i, _i, _vp, _i, _f, _f, _f, _f, _vp, _vp, _vp]),
    "b200_box_candidates": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _f, _f, _f, _f, _f, _i64,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_select_detections": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "b200_select_topk": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _i, _vp, _vp, _vp, _vp]),
    "b200_embed_match": (_i, [_vp, _vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_linear_bf16": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "b200_embed_match_wide": (_i, [_vp, _vp, _i64, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "b200_mask_targets": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "b200_paste_masks": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _f, _vp, _vp]),
    "b200_colmax_decode": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "b200_roi_pool_forward": (_i, [_vp, _i, _i, _i, _i, _vp, _i64, _f, _i, _i, _vp, _vp, _vp]),
    "b200_roi_pool_backward": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}


def lib():
    """Load (once) and return the shared library; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m cvpr22_cross_modal_pseudo_labeling_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if hasattr(l, "b200_debug_set"):
            l.b200_debug_set.restype = None
            l.b200_debug_set.argtypes = [_i, _i, _i]
        if hasattr(l, "b200_debug_bwd"):
            l.b200_debug_bwd.restype = None
            l.b200_debug_bwd.argtypes = [_i]
        if hasattr(l, "b200_debug_rpn"):
            l.b200_debug_rpn.restype = None
            l.b200_debug_rpn.argtypes = [_i]
        if hasattr(l, "b200_debug_match"):
            l.b200_debug_match.restype = None
            l.b200_debug_match.argtypes = [_i]
        if hasattr(l, "b200_debug_nms"):
            l.b200_debug_nms.restype = None
            l.b200_debug_nms.argtypes = [_i]
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().b200_last_error_string().decode(errors="replace")
        if rc == -1:
            raise ValueError("%s: %s" % (what, msg))
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg))


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (the B200 path has no CPU implementation)" % name)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def debug_rpn(force_exact=False):
    """Test hook: make b200_rpn_candidates skip the sampled lower bound and run the exact bisection."""
    lib().b200_debug_rpn(int(force_exact))


def debug_set(force_generic=False, exact=True, variant=0):
    """Tuning / test hook (not part of the reference-facing ABI): select the generic RoIAlign kernel,
    the arithmetic of b200_roi_align_forward (exact / FMA), and a kernel variant:
      bits 0-3  occupancy variants of the marching kernels (1, 2; 8 = one channel chunk per CTA)
      bit 4     (16) fast math at the FPN box pooler shape: the separable marching kernel instead of
                the row-streaming kernel
      bit 5     (32) row-streaming kernel visits the RoIs in the order given (no ordering pass)
      bit 6     (64) row-streaming kernel without arithmetic (copy-pipeline probe; output is zeros)
      bit 7     (128) two copy warps instead of four; bit 8 (256) output tile not stored
      bit 9     (512) cycle counters of who waits for whom (read back with b200_debug_rows_stats)
      bits 10-12 (1024 * k) ring entries per wait group (default: 1 for fp32 maps, 4 for bf16 maps)
      bit 13    (8192) tiled gather (NHWC, sampling ratio != 2): 64 channels per tile instead of 32
      bit 14    (16384) tiled gather off: the plain per-(quad, bin) gather
    Process-wide; tests and probes reset it to (False, True, 0)."""
    lib().b200_debug_set(int(force_generic), int(exact), int(variant))


def debug_match(legacy=False):
    """Test hook: SOFTMAX scoring through the one-CTA-per-tile kernel (embed_match.cu) instead of the
    persistent kernel with overlapped epilogue (tc_gemm.cu)."""
    lib().b200_debug_match(int(legacy))


def debug_nms(mode=0):
    """Test hook: 0 (or False) = the library's choice (fused one-CTA-per-segment kernel when the longest segment fits
    in shared memory, except keep-all calls over long segments, which take the bitmask path), 1 (or True) = force the
    three-kernel bitmask path, 2 = force the fused kernel wherever it fits."""
    lib().b200_debug_nms(int(mode))


def debug_bwd(mode=0):
    """Test hook: the RoIAlign backward kernel for NHWC / sampling ratio 2 -- 0 (or False) = the marching kernel by
    tap row (default), 1 (or True) = the per-tap kernel, 2 = the marching kernel by output row (round 1)."""
    lib().b200_debug_bwd(int(mode))
