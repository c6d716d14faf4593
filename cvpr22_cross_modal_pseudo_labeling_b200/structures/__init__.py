from .bounding_box import BoxList
from .boxlist_ops import boxlist_iou, boxlist_nms, cat_boxlist, remove_small_boxes

__all__ = ["BoxList", "boxlist_nms", "boxlist_iou", "cat_boxlist", "remove_small_boxes"]
