"""Drop-in for maskrcnn_benchmark.layers (reference layers/__init__.py:10-14), RoI hot path only."""
from .embed_match import caption_align, embed_logits, embed_match_softmax
from .linear import TensorCoreLinear, linear_bf16
from .nms import nms, nms_batched, select_topk
from .roi_align import ROIAlign, get_roi_align_math, roi_align, roi_align_multilevel, roi_align_with_mean, set_roi_align_math
from .roi_pool import ROIPool, roi_pool

__all__ = ["linear_bf16", "TensorCoreLinear", "embed_match_softmax", "embed_logits", "caption_align", "nms", "nms_batched", "select_topk", "roi_align", "roi_align_multilevel", "roi_align_with_mean", "ROIAlign", "set_roi_align_math", "get_roi_align_math", "roi_pool", "ROIPool"]
