"""Method registry for the reducers of this path (GCond / GCondX and their one-step siblings DosCond / DosCondX) (graphslim/reduction/registry.py:136-142).

``create_reducer('gcond'|'gcondx', setting, data, args)`` returns the B200 implementation; every other name is
forwarded to the reference registry when the reference package is importable, so this module can stand in for
``graphslim.reduction.create_reducer`` inside train_all.py.
"""
from importlib import import_module

_LOCAL = {
    "gcond": ("graphslim_b200.condensation.gcond", "GCond"),
    "gcondx": ("graphslim_b200.condensation.gcondx", "GCondX"),
    "doscond": ("graphslim_b200.condensation.doscond", "DosCond"),
    "doscondx": ("graphslim_b200.condensation.doscondx", "DosCondX"),
}


def normalize_method_name(method):
    return method.strip().replace("-", "_").lower()


def create_reducer(method, setting, data, args, **kwargs):
    key = normalize_method_name(method)
    if key in _LOCAL:
        mod, cls = _LOCAL[key]
        return getattr(import_module(mod), cls)(setting=setting, data=data, args=args, **kwargs)
    try:
        ref = import_module("graphslim.reduction")
    except ImportError as exc:
        raise ValueError(f"Unknown graph reduction method '{method}'. graphslim_b200 provides: "
                         f"{', '.join(sorted(_LOCAL))}") from exc
    return ref.create_reducer(method, setting, data, args, **kwargs)
