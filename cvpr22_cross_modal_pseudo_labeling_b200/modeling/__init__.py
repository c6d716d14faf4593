"""Host-side mirrors of the reference's modeling classes that sit on the RoI hot path."""
from .box_coder import BoxCoder
from .poolers import LevelMapper, Pooler, make_pooler
from .pseudo_label import generate_pseudo_labels, pack_records
from .roi_heads.box_head.inference import PostProcessor, make_roi_box_post_processor
from .roi_heads.box_head.roi_box_predictors import FastRCNNPredictor, make_roi_box_predictor
from .roi_heads.mask_head.inference import Masker, MaskPostProcessor, make_roi_mask_post_processor
from .rpn.inference import RPNPostProcessor, make_rpn_postprocessor

__all__ = ["BoxCoder", "LevelMapper", "Pooler", "make_pooler", "RPNPostProcessor", "make_rpn_postprocessor",
           "PostProcessor", "make_roi_box_post_processor", "FastRCNNPredictor", "make_roi_box_predictor",
           "Masker", "MaskPostProcessor", "make_roi_mask_post_processor", "generate_pseudo_labels", "pack_records"]
