"""install(): swap the B200 hot path into an imported `maskrcnn_benchmark` (the reference).

The reference binds its native code in exactly four places on this path; each is replaced
by the object of the same name from this package (same call signature):

  maskrcnn_benchmark/layers/nms.py:8                    nms          -> layers.nms
  maskrcnn_benchmark/layers/roi_align.py:48,50          roi_align, ROIAlign
  maskrcnn_benchmark/layers/roi_pool.py:46,49           roi_pool, ROIPool
  maskrcnn_benchmark/structures/boxlist_ops.py:9        boxlist_nms
  maskrcnn_benchmark/modeling/poolers.py:45,124         Pooler, make_pooler     (fused multi-level kernel)
  maskrcnn_benchmark/modeling/rpn/inference.py:13       RPNPostProcessor        (one batched NMS per forward)
  maskrcnn_benchmark/modeling/roi_heads/box_head/inference.py:12       PostProcessor
  maskrcnn_benchmark/modeling/roi_heads/box_head/roi_box_predictors.py:8   FastRCNNPredictor
  maskrcnn_benchmark/modeling/roi_heads/mask_head/inference.py:11,168     MaskPostProcessor, Masker (one paste launch per image)

Call it after `import maskrcnn_benchmark` and before `build_detection_model(cfg)`.
"""
import importlib
import sys


def install(verbose=False):
    from . import layers, modeling, structures
    from .modeling.roi_heads.box_head import inference as box_inference
    from .modeling.roi_heads.box_head import roi_box_predictors
    from .modeling.roi_heads.mask_head import inference as mask_inference
    from .modeling.rpn import inference as rpn_inference

    patches = {
        "maskrcnn_benchmark.layers": dict(nms=layers.nms, roi_align=layers.roi_align, ROIAlign=layers.ROIAlign,
                                          roi_pool=layers.roi_pool, ROIPool=layers.ROIPool),
        "maskrcnn_benchmark.layers.nms": dict(nms=layers.nms),
        "maskrcnn_benchmark.layers.roi_align": dict(roi_align=layers.roi_align, ROIAlign=layers.ROIAlign),
        "maskrcnn_benchmark.layers.roi_pool": dict(roi_pool=layers.roi_pool, ROIPool=layers.ROIPool),
        "maskrcnn_benchmark.structures.boxlist_ops": dict(boxlist_nms=structures.boxlist_nms, _box_nms=layers.nms),
        "maskrcnn_benchmark.modeling.poolers": dict(Pooler=modeling.Pooler, make_pooler=modeling.make_pooler,
                                                    LevelMapper=modeling.LevelMapper, ROIAlign=layers.ROIAlign),
        "maskrcnn_benchmark.modeling.rpn.inference": dict(RPNPostProcessor=rpn_inference.RPNPostProcessor,
                                                          make_rpn_postprocessor=rpn_inference.make_rpn_postprocessor,
                                                          boxlist_nms=structures.boxlist_nms),
        "maskrcnn_benchmark.modeling.roi_heads.box_head.inference": dict(
            PostProcessor=box_inference.PostProcessor,
            make_roi_box_post_processor=box_inference.make_roi_box_post_processor,
            boxlist_nms=structures.boxlist_nms),
        "maskrcnn_benchmark.modeling.roi_heads.box_head.roi_box_predictors": dict(
            FastRCNNPredictor=roi_box_predictors.FastRCNNPredictor,
            make_roi_box_predictor=roi_box_predictors.make_roi_box_predictor),
        "maskrcnn_benchmark.modeling.roi_heads.mask_head.inference": dict(
            MaskPostProcessor=mask_inference.MaskPostProcessor, Masker=mask_inference.Masker,
            make_roi_mask_post_processor=mask_inference.make_roi_mask_post_processor),
    }
    done = []
    for mod_name, attrs in patches.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                continue  # that part of the reference is not importable in this environment
        for k, v in attrs.items():
            setattr(mod, k, v)
            done.append("%s.%s" % (mod_name, k))
    if verbose:
        print("\n".join(done))
    return done
