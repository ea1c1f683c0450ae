"""Minimal AnnData-compatible container.

``anndata`` is not installed in the build image, and ``detect_svgs`` / ``pca`` / ``aa`` only
duck-type the object they are given (the attribute set the reference touches is listed in
SURVEY.md section 8b: ``X``, ``len()``, ``uns_keys()``, ``obsm``, ``var``, ``var_names``,
``uns``).  A real ``anndata.AnnData`` works unchanged; this class exists so tests, the
bench and ``smoke()`` can run without it.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


class AnnDataLite:
    def __init__(self, X, obsm=None, var_names=None, obs_names=None, uns=None):
        self.X = X
        n, g = X.shape
        var_names = [f"gene{j}" for j in range(g)] if var_names is None else list(var_names)
        obs_names = [f"spot{i}" for i in range(n)] if obs_names is None else list(obs_names)
        self.var = pd.DataFrame(index=pd.Index(var_names))
        self.obs = pd.DataFrame(index=pd.Index(obs_names))
        self.obsm = dict(obsm or {})
        self.varm = {}
        self.uns = dict(uns or {})

    @property
    def var_names(self):
        return self.var.index

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    @property
    def shape(self):
        return self.X.shape

    def __len__(self):
        return self.X.shape[0]

    def uns_keys(self):
        return list(self.uns.keys())

    def obsm_keys(self):
        return list(self.obsm.keys())

    def copy(self):
        out = AnnDataLite(self.X.copy(), {k: np.array(v, copy=True) for k, v in self.obsm.items()},
                          self.var_names, self.obs_names, dict(self.uns))
        out.var = self.var.copy()
        out.obs = self.obs.copy()
        return out
