"""Import and run the UNMODIFIED reference package (Teichlab/Genes2Genes v0.2.0) behind stubs for the
plotting / enrichment packages this image lacks (recipe: SURVEY.md section 8(c)).

Where the package comes from: ``/root/reference`` in the build container, else ``baseline/_ref`` (the
byte-for-byte copy ``baseline/install_reference.sh`` installs; it travels to the GPU box).  Used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by ``bench.py`` to time the
reference's own CPU path (``cpu_baseline`` / ``--impl reference``).  The product never imports this module.
"""
import importlib
import os
import sys
import types

import numpy as np
import pandas as pd

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    for root in (os.environ.get("G2G_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if root and os.path.isdir(os.path.join(root, "genes2genes")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "genes2genes"))


class _Dummy(types.ModuleType):
    """Module stub: any public attribute is a callable that returns another dummy."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)

        def _call(*a, **k):
            return None

        return _call


def _install_stubs():
    import scipy.sparse  # noqa: F401  (pre-import the real heavy packages)
    import sklearn  # noqa: F401
    import torch  # noqa: F401
    import tqdm

    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib.colors", "gseapy", "blitzgsea", "Levenshtein"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Dummy(name)
    if "tqdm.notebook" not in sys.modules or not hasattr(sys.modules["tqdm.notebook"], "tqdm_notebook"):
        m = types.ModuleType("tqdm.notebook")
        m.tqdm_notebook = tqdm.tqdm
        m.tqdm = tqdm.tqdm
        sys.modules["tqdm.notebook"] = m
    if "anndata" not in sys.modules:
        try:
            importlib.import_module("anndata")
        except Exception:
            ad = types.ModuleType("anndata")
            core = types.ModuleType("anndata._core")
            views = types.ModuleType("anndata._core.views")

            class SparseCSCView:  # only used in an isinstance() test, Main.py:230
                pass

            views.SparseCSCView = SparseCSCView
            core.views = views
            ad._core = core
            ad.AnnData = AnnDataStandIn
            sys.modules["anndata"] = ad
            sys.modules["anndata._core"] = core
            sys.modules["anndata._core.views"] = views
    # scipy >= 1.8 keeps scipy.sparse.csr / .csc as deprecated shims; make sure they resolve
    import scipy.sparse as sp
    for sub, cls in (("csr", "csr_matrix"), ("csc", "csc_matrix")):
        if not hasattr(sp, sub):
            m = types.ModuleType("scipy.sparse." + sub)
            setattr(m, cls, getattr(sp, cls))
            setattr(sp, sub, m)


class AnnDataStandIn:
    """The ~30 lines of AnnData surface the reference touches."""

    def __init__(self, X, obs, var_names):
        import scipy.sparse as sp
        self.X = X if sp.issparse(X) else sp.csr_matrix(np.asarray(X))
        self.obs = obs
        self.var_names = pd.Index(var_names)

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def shape(self):
        return self.X.shape

    def __getitem__(self, key):
        if isinstance(key, tuple):
            rows, cols = key
            assert isinstance(rows, slice) and rows == slice(None)
            idx = self.var_names.get_indexer(list(cols))
            return AnnDataStandIn(self.X[:, idx], self.obs, self.var_names[idx])
        rows = np.asarray(key)
        return AnnDataStandIn(self.X[rows], self.obs.iloc[rows], self.var_names)


def make_adata(X, time, obs_names=None, var_names=None):
    n, g = X.shape
    obs_names = obs_names if obs_names is not None else ["c%d" % i for i in range(n)]
    var_names = var_names if var_names is not None else ["g%d" % i for i in range(g)]
    obs = pd.DataFrame({"time": np.asarray(time, dtype=np.float64)}, index=pd.Index(obs_names))
    return AnnDataStandIn(X, obs, var_names)


def import_reference():
    """Return the reference's ``genes2genes`` package (imported unmodified)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("genes2genes")


def time_reference(X_ref, time_ref, X_query, time_query, n_bins, concurrent=False, n_processes=None):
    """Wall time of the reference's own ``RefQueryAligner(...)`` + ``align_all_pairs()`` on the given dense
    arrays (cells x genes): (seconds, alignment strings).  ``concurrent=True`` uses the reference's
    multiprocessing path (Main.py:438-448)."""
    import contextlib
    import io
    import time
    g2g = import_reference()
    genes = ["g%d" % i for i in range(X_ref.shape[1])]
    ar = make_adata(X_ref, time_ref, ["r%d" % i for i in range(X_ref.shape[0])], genes)
    aq = make_adata(X_query, time_query, ["q%d" % i for i in range(X_query.shape[0])], genes)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        t0 = time.perf_counter()
        aligner = g2g.Main.RefQueryAligner(ar, aq, genes, int(n_bins))
        if concurrent:
            aligner.align_all_pairs(concurrent=True, n_processes=n_processes)
        else:
            aligner.align_all_pairs()
        dt = time.perf_counter() - t0
    return dt, [a.alignment_str for a in aligner.results]
