"""B200-native RoI hot path for open-vocabulary pseudo-labeling.

Drop-in replacements for the RoI operators of hbdat/cvpr22_cross_modal_pseudo_labeling
(a maskrcnn-benchmark fork): `layers.nms / ROIAlign / roi_align / ROIPool`,
`structures.boxlist_nms`, `modeling.Pooler`, the RPN and box post-processors and the
embedding box predictor -- all backed by hand-written sm_100a CUDA kernels behind a
C ABI (include/b200det.h, libb200det.so).  No CPU implementation, no fallback.
"""
from . import _ext  # noqa: F401

__version__ = "0.1.0"


def build_kernels(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libb200det.so (in-tree)."""
    import importlib
    return importlib.import_module(__name__ + ".build").build(force=force, verbose=verbose)
