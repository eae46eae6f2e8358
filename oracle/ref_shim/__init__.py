"""ORACLE (test infrastructure, not product code) -- golden-vector generation only.

Import shim that lets the UNMODIFIED reference package under /root/reference
run in this container, where its third-party native dependencies
(torch_sparse, torch_geometric, torch_scatter, pygsp, networkit, dgl, ogb,
deeprobust, ...) are not installed.

* a meta-path finder fabricates every missing top-level package as an
  attribute-on-demand module;
* the handful of symbols the GCond path really executes get working CPU
  stand-ins: ``torch_sparse.SparseTensor`` / ``matmul`` (CSR SpMM with
  autograd), ``SparseTensor.sample_adj`` (C++ restatement in
  oracle/csrc/oracle_host.cpp driven by torch's own CPU generator),
  ``torch_geometric.loader.NeighborSampler`` and
  ``torch_geometric.utils.to_undirected``.

This module needs /root/reference and therefore cannot run on the GPU box; it
is used by ``oracle/make_goldens.py`` to produce the fixtures in tests/golden/.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GRAPHSLIM_REFERENCE_ROOT", "/root/reference")

_FAKE_TOP = (
    "torch_sparse", "torch_geometric", "torch_scatter", "torch_cluster", "pygsp", "networkit", "dgl", "ogb",
    "deeprobust", "gdown", "matplotlib", "wandb", "seaborn", "sortedcontainers_stub", "numba", "tensorboardX",
    "streamlit", "plotly", "pyvis", "community", "prettytable", "swanlab", "pyfpgrowth", "pyemd", "ot", "torch_spline_conv",
    "scikit_network", "sknetwork", "gntk", "tqdm_stub",
)


class _Anything:
    """Inert placeholder usable as base class, decorator, callable or namespace."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _FakeModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = type(name, (object,), {"__init__": lambda self, *a, **k: None,
                                     "__module__": self.__name__}) if name[:1].isupper() else _Anything()
        setattr(self, name, val)
        return val


def _requested_by_reference():
    """True when the import statement being resolved sits in a file of the reference tree."""
    f = sys._getframe(2)
    while f is not None:
        fn = f.f_code.co_filename
        if not fn.startswith("<frozen importlib") and "importlib" not in fn and fn != __file__:
            return fn.startswith(REFERENCE_ROOT)
        f = f.f_back
    return False


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Sits LAST on sys.meta_path, so it only sees imports nothing else could satisfy.

    While ``catch_all`` is on (during the import of the reference package) every such top-level
    name is fabricated and remembered; afterwards only the remembered names are served.
    """

    def __init__(self, tops):
        self.tops = set(tops)
        self.catch_all = False

    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in self.tops or (self.catch_all and top != "graphslim" and _requested_by_reference()):
            self.tops.add(top)
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _FakeModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        from . import stand_ins
        stand_ins.populate(module)


_installed = False


def install():
    """Make ``import graphslim`` resolve to the unmodified reference package."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "graphslim")):
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}; the shim only runs in the build "
                           "container (golden vectors are committed under tests/golden/)")
    missing = []
    for top in _FAKE_TOP:
        try:
            if importlib.util.find_spec(top) is None:
                missing.append(top)
        except (ImportError, ValueError):
            missing.append(top)
    from . import stand_ins  # noqa: F401  (import before the catch-all is armed)
    finder = _Finder(missing)
    sys.meta_path.append(finder)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    finder.catch_all = True
    try:
        importlib.import_module("graphslim.condensation.gcond")
        importlib.import_module("graphslim.condensation.gcondx")
        importlib.import_module("graphslim.dataset.loader")
        importlib.import_module("graphslim.reduction")
    finally:
        finder.catch_all = False
    _installed = True
