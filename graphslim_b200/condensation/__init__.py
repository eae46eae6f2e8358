from .gcond import GCond
from .gcondx import GCondX

__all__ = ["GCond", "GCondX"]
