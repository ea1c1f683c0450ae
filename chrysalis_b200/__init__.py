"""chrysalis_b200 - the detect_svgs -> pca -> aa hot path of rockdeme/chrysalis on B200 (sm_100a).

Re-exports mirror ``/root/reference/chrysalis/__init__.py:10-12`` for the path this package
replaces; everything else of the reference (plots, utils) is out of scope.
"""
from .core import aa, detect_svgs, local_morans_i, morans_i_permutation, pca
from .anndata_lite import AnnDataLite
from .utils import estimate_compartments, integrate_adatas

__all__ = ["detect_svgs", "pca", "aa", "morans_i_permutation", "local_morans_i", "integrate_adatas", "estimate_compartments", "AnnDataLite"]
__version__ = "0.1.0"
