"""DosCond on B200: drop-in for graphslim.condensation.doscond.DosCond ("Condensing Graphs via One-Step Gradient
Matching"), SURVEY.md section 8f-2 -- a host-loop variant over the GCond kernels."""
from .gcond import GCond


class DosCond(GCond):
    """graphslim/condensation/doscond.py:17-65: per outer step one PGE forward, one matching step against the freshly
    initialised condense model, then the PGE optimiser AND the feature optimiser both step (:55-56); the condense model
    is re-initialised every epoch and never trained (no inner loop, no pge.inference between steps).  The published
    adjacency is pge.inference(feat_syn) (:61)."""

    one_step = True
