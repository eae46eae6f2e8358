"""DosCondX on B200: drop-in for graphslim.condensation.doscondx.DosCondX (structure-free DosCond)."""
from .gcondx import GCondX


class DosCondX(GCondX):
    """graphslim/condensation/doscondx.py:19-63: identity synthetic adjacency, one matching step per outer step, the
    feature optimiser steps every time (:50-52); no PGE, no inner loop."""

    one_step = True
