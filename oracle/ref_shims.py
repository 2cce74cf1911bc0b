"""TEST INFRASTRUCTURE ONLY — import shims for executing the *unmodified* pyliger reference.

This module is used only by ``tests/golden/make_golden.py`` (run in the development container,
where ``/root/reference`` is mounted) to generate the committed golden vectors.  It never copies
reference source; it only arranges ``sys.modules`` so that the three hot-path modules

    pyliger.factorization._iNMF_ANLS     (optimize_ALS)
    pyliger.factorization._online_iNMF   (online_iNMF, _update_A_B, _generate_idx, ...)
    pyliger.factorization._utilities     (nnlsm_blockpivot, _update_W_HALS, ...)

can be imported although ``anndata``, ``lazy_loader`` and ``h5sparse`` are not installed here and
``pyliger/__init__.py`` pulls in plotting / clustering dependencies that are absent.

Nothing under ``pyliger_b200/`` imports this file.
"""
import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get("PYLIGER_REFERENCE_SRC", "/root/reference/src")


class _MissingModule(types.ModuleType):
    def __getattr__(self, name):  # pragma: no cover - only hit if a stub is really used
        raise ImportError("module %s is not installed (stub)" % self.__name__)


def _lazy_load(name, error_on_import=False):
    try:
        return importlib.import_module(name)
    except ImportError:
        return _MissingModule(name)


class AnnData(object):
    """Duck-typed stand-in with the attributes the factorisation path touches."""

    def __init__(self, X=None, obs=None, var=None, uns=None, layers=None):
        self.X = X
        self.obs = obs
        self.var = var
        self.uns = {} if uns is None else uns
        self.layers = {} if layers is None else layers
        self.obsm = {}
        self.varm = {}
        self.isbacked = False

    @property
    def shape(self):
        return self.X.shape


def install():
    """Register the stub modules and the bypass packages. Idempotent."""
    if "pyliger" in sys.modules and getattr(sys.modules["pyliger"], "_b200_shim", False):
        return
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "pyliger")):
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_SRC)

    lazy = types.ModuleType("lazy_loader")
    lazy.load = _lazy_load
    sys.modules["lazy_loader"] = lazy

    ad = types.ModuleType("anndata")
    ad.AnnData = AnnData
    sys.modules["anndata"] = ad

    h5 = types.ModuleType("h5sparse")
    h5sub = types.ModuleType("h5sparse.h5sparse")

    class File(object):  # only used in isinstance checks on the in-memory path
        pass

    h5.File = File
    h5sub.File = File
    h5.h5sparse = h5sub
    sys.modules["h5sparse"] = h5
    sys.modules["h5sparse.h5sparse"] = h5sub

    for mod in ("numexpr", "h5py"):
        if mod not in sys.modules:
            try:
                importlib.import_module(mod)
            except ImportError:
                sys.modules[mod] = _MissingModule(mod)

    # third-party modules that pyliger/clustering/_utilities.py imports at module level but quantile_norm's exact-kNN
    # path never calls (annoy / pynndescent), plus numba, whose @njit on cluster_vote is an optimisation only
    if "numba" not in sys.modules:
        try:
            importlib.import_module("numba")
        except ImportError:
            nb = types.ModuleType("numba")
            nb.njit = lambda f=None, *a, **k: f if callable(f) else (lambda g: g)
            sys.modules["numba"] = nb
    for mod, attr in (("annoy", "AnnoyIndex"), ("pynndescent", "NNDescent")):
        if mod not in sys.modules:
            try:
                importlib.import_module(mod)
            except ImportError:
                m = types.ModuleType(mod)
                setattr(m, attr, type(attr, (), {}))
                sys.modules[mod] = m

    root = os.path.join(REFERENCE_SRC, "pyliger")
    pkg = types.ModuleType("pyliger")
    pkg.__path__ = [root]
    pkg._b200_shim = True
    sys.modules["pyliger"] = pkg
    for sub in ("factorization", "preprocessing", "tools", "clustering"):
        m = types.ModuleType("pyliger." + sub)
        m.__path__ = [os.path.join(root, sub)]
        sys.modules["pyliger." + sub] = m
        setattr(pkg, sub, m)


def load():
    """Return (optimize_ALS, online_iNMF, utilities module, online module, Liger class)."""
    install()
    from pyliger.factorization import _utilities as ref_util
    from pyliger.factorization import _online_iNMF as ref_online
    from pyliger.factorization._iNMF_ANLS import optimize_ALS
    from pyliger.pyliger import Liger

    return optimize_ALS, ref_online.online_iNMF, ref_util, ref_online, Liger
