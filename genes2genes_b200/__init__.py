"""genes2genes_b200 -- B200-native implementation of the Genes2Genes per-gene alignment hot path.

Drop-in for ``from genes2genes import Main``: ``Main.RefQueryAligner(adata_ref, adata_query,
gene_list, n_interpolation_points)``, ``align_all_pairs()``, ``results`` / ``results_map``.
Host code is Python/PyTorch over ``libg2g.so`` (hand-written sm_100a CUDA, C ABI in
``include/g2g.h``).  There is no CPU fallback: without the built library or a CUDA device the
aligner raises.
"""
from . import ClusterUtils  # noqa: F401
from . import Main  # noqa: F401
from . import OrgAlign  # noqa: F401
from . import TimeSeriesPreprocessor  # noqa: F401
from . import synthetic  # noqa: F401

__version__ = "0.1.0"
