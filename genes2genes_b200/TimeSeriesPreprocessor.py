"""Host-side mirror of the reference's ``TimeSeriesPreprocessor`` interface.

The numeric work of the reference module (kernel weights, weighted moments, sampling) runs in
``libg2g.so`` for all genes at once; the classes here only carry the small per-dataset state
(sorted pseudotimes, interpolation points, adaptive kernel widths) and expose per-gene results
with the reference's attribute names.
"""
import numpy as np

N_SAMPLES = 50


class TrajectoryInterpolator:
    """Per-dataset interpolation state (reference: TimeSeriesPreprocessor.py:12-133).

    Only pseudotime is needed on the host: ``cell_pseudotimes`` (sorted), ``interpolation_points``
    and, in adaptive mode, the per-bin Gaussian-kernel denominators.  ``order`` is the argsort
    used to time-sort the cells (TimeSeriesPreprocessor.py:20).
    """

    def __init__(self, adata, n_bins, adaptive_kernel=True, kernel_WINDOW_SIZE=0.1, raising_degree=1,
                 gene_list=None, obs_names=None):
        """``adata``: an AnnData-like object, as in the reference (``TrajectoryInterpolator(adata, n_bins,
        ...)``, TimeSeriesPreprocessor.py:18), or just the pseudotime array of the cells (what the device path
        needs; ``gene_list`` / ``obs_names`` then come from the keyword arguments)."""
        self._adata = None
        if hasattr(adata, "obs"):
            self._adata = adata
            time = np.asarray(adata.obs["time"], dtype=np.float64)
            if obs_names is None:
                obs_names = getattr(adata, "obs_names", None)
            if gene_list is None:
                gene_list = getattr(adata, "var_names", None)
        else:
            time = np.asarray(adata, dtype=np.float64)
        self.n_bins = int(n_bins)
        self.order = np.argsort(time)
        self.cell_pseudotimes = time[self.order]
        self.interpolation_points = np.linspace(0, 1, self.n_bins)
        self.kernel_WINDOW_SIZE = kernel_WINDOW_SIZE
        self.adaptive_kernel = adaptive_kernel
        self.k = raising_degree
        self.gene_list = gene_list
        self._obs_names_in = obs_names
        self._obs_names = None
        self._weights = None
        self._mat = None

    @property
    def obs_names(self):
        """Cell names in pseudotime order, like the cells themselves (TimeSeriesPreprocessor.py:20)."""
        if self._obs_names is None and self._obs_names_in is not None:
            self._obs_names = np.asarray(self._obs_names_in)[self.order]
        return self._obs_names

    @property
    def adata(self):
        """The time-sorted AnnData (TimeSeriesPreprocessor.py:20); only when built from one."""
        if self._adata is None:
            raise AttributeError("this TrajectoryInterpolator was built from a pseudotime array, not an AnnData")
        return self._adata[self.order]

    @property
    def mat(self):
        """genes x cells CSR matrix of the time-sorted cells (TimeSeriesPreprocessor.py:28), built on first
        access; the device path streams the packed matrix instead."""
        if self._mat is None:
            from scipy.sparse import csr_matrix
            X = self.adata.X
            X = X.todense() if hasattr(X, "todense") else X
            if hasattr(X, "cpu"):
                X = X.cpu().numpy()
            self._mat = csr_matrix(np.asarray(X).transpose())
        return self._mat

    def run(self):
        if self.adaptive_kernel:
            self.reciprocal_cell_density_estimates = self.compute_cell_densities()
            self.adaptive_win_denoms = self.compute_adaptive_window_denominator()
        return self

    # -- adaptive kernel widths (TimeSeriesPreprocessor.py:53-119) --------------------------------
    def compute_cell_densities(self):
        """Reciprocal cell density around every interpolation point: cells inside
        [p[i-1], p[i+1]] (one-sided at the ends) per unit pseudotime, inverted; empty windows
        take the largest finite value."""
        p, tau = self.interpolation_points, self.cell_pseudotimes
        last = self.n_bins - 1
        lo = np.where(np.arange(self.n_bins) == 0, -np.inf, np.roll(p, 1))
        hi = np.where(np.arange(self.n_bins) == last, np.inf, np.roll(p, -1))
        width = np.full(self.n_bins, p[2] - p[0])
        width[0] = width[last] = p[1] - p[0]
        counts = np.array([np.count_nonzero((tau >= lo[i]) & (tau <= hi[i])) for i in range(self.n_bins)])
        self.cell_density_estimates = list(counts / width)
        recip = np.array([1.0 / d if d != 0 else np.inf for d in self.cell_density_estimates])
        if np.any(np.isinf(recip)):
            recip = np.where(np.isinf(recip), np.max(recip[np.isfinite(recip)]), recip)
        return recip

    def compute_adaptive_window_denominator(self):
        """Min-max scale the reciprocal densities to [0, 1], stretch by ``k``, turn them into window
        sizes relative to ``kernel_WINDOW_SIZE`` and shift so the least affected point keeps the
        base window; the squares are the Gaussian denominators."""
        r = np.asarray(self.reciprocal_cell_density_estimates, dtype=np.float64)
        span = r.max() - r.min()
        scale = 1.0 / span if span != 0 else 1.0
        scaled = (r * scale + (0.0 - r.min() * scale)) * self.k
        sizes = scaled * self.kernel_WINDOW_SIZE
        gap = np.abs(sizes - self.kernel_WINDOW_SIZE)
        residue = np.abs(self.kernel_WINDOW_SIZE - sizes[int(np.argmax(gap))])
        sizes = sizes + (residue / (self.k - 1) if self.k > 1 else residue)
        self.adaptive_window_sizes = sizes
        return [s ** 2 for s in sizes]

    def kernel_denominators(self):
        """Per-bin denominators handed to ``g2g_weights``."""
        if self.adaptive_kernel:
            return np.asarray(self.adaptive_win_denoms, dtype=np.float64)
        return np.full(self.n_bins, self.kernel_WINDOW_SIZE ** 2, dtype=np.float64)

    @property
    def cell_weight_mat(self):
        """[n_bins x cells] Gaussian weights as a DataFrame (TimeSeriesPreprocessor.py:122-133).
        Not used by the device path (the kernel regenerates weights from pseudotime); built on
        first access for downstream code that inspects it."""
        if self._weights is None:
            import pandas as pd
            d = np.abs(self.cell_pseudotimes[None, :] - self.interpolation_points[:, None])
            w = np.exp(-np.divide(d ** 2, self.kernel_denominators()[:, None]))
            self._weights = pd.DataFrame(w, index=self.interpolation_points, columns=self.obs_names)
        return self._weights


class SummaryTimeSeries:
    """Interpolated series of one gene (reference: TimeSeriesPreprocessor.py:198-213), a view over
    the batched device results.  ``data_bins`` -- the 50 samples per bin the reference draws -- are
    rebuilt on first access as ``mean + eps * std`` from the fixed epsilon table, bit-identical to
    ``Normal(mean, std).rsample()`` under the reference's seeding."""

    def __init__(self, gene_name, intpl_means, intpl_stds, mean_trend, std_trend, time_points, eps):
        self.gene_name = gene_name
        self.intpl_means = list(intpl_means)
        self.intpl_stds = list(intpl_stds)
        self.mean_trend = np.asarray(mean_trend)
        self.std_trend = np.asarray(std_trend)
        self.time_points = np.asarray(time_points)
        self._eps = eps
        self._bins = None

    @property
    def data_bins(self):
        if self._bins is None:
            mu = np.asarray(self.intpl_means)[:, None]
            sd = np.asarray(self.intpl_stds)[:, None]
            self._bins = list(mu + self._eps * sd)
        return self._bins

    @property
    def Y(self):
        return np.asarray(self.data_bins).flatten()

    @property
    def X(self):
        return np.repeat(self.time_points, self._eps.shape[1])

    def plot_mean_trend(self, color="midnightblue"):
        import seaborn as sb
        sb.lineplot(x=self.time_points, y=self.mean_trend, color=color, linewidth=4)

    def plot_std_trend(self, color="midnightblue"):
        import seaborn as sb
        sb.lineplot(x=self.time_points, y=self.std_trend, color=color, linewidth=4)
