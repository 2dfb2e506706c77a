"""Minimal AnnData stand-in (``X``, ``obs['time']``, ``var_names``, ``obs_names``, slicing).

``anndata`` is not installed in the build / GPU images; ``RefQueryAligner`` only touches this
surface, so real ``anndata.AnnData`` objects work the same way.
"""
import numpy as np
import pandas as pd


class AnnDataLite:
    def __init__(self, X, time, var_names=None, obs_names=None):
        self.X = X
        n, g = X.shape
        obs_names = ["cell%d" % i for i in range(n)] if obs_names is None else list(obs_names)
        var_names = ["g%d" % i for i in range(g)] if var_names is None else list(var_names)
        self.obs = pd.DataFrame({"time": np.asarray(time, dtype=np.float64)}, index=pd.Index(obs_names))
        self.var_names = pd.Index(var_names)

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def shape(self):
        return tuple(self.X.shape)

    def __getitem__(self, key):
        if isinstance(key, tuple):
            rows, cols = key
            idx = self.var_names.get_indexer(list(cols))
            X = self.X[:, idx] if rows == slice(None) else self.X[rows][:, idx]
            obs = self.obs if rows == slice(None) else self.obs.iloc[rows]
            return AnnDataLite(X, obs["time"].values, self.var_names[idx], obs.index)
        rows = np.asarray(key)
        return AnnDataLite(self.X[rows], self.obs["time"].values[rows], self.var_names, self.obs.index[rows])
