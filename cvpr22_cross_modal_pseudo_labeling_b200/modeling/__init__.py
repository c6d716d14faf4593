"""Host-side mirrors of the reference's modeling classes that sit on the RoI hot path."""
from .box_coder import BoxCoder
from .poolers import LevelMapper, Pooler, make_pooler

__all__ = ["BoxCoder", "LevelMapper", "Pooler", "make_pooler"]
