"""GCondX on B200: drop-in for graphslim.condensation.gcondx.GCondX (structure-free GCond)."""
from .gcond import GCond


class GCondX(GCond):
    """graphslim/condensation/gcondx.py:17-79: identity synthetic adjacency, no PGE forward; the PGE optimiser
    is "stepped" (a no-op, PGE has no gradient) on every fifth outer step, features otherwise (:58-61)."""

    x_variant = True

    def _pge_turn(self, it, ol):
        return ol % 5 < 1
