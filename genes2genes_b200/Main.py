"""Drop-in for the per-gene alignment path of ``genes2genes.Main`` (reference v0.2.0):

    aligner = Main.RefQueryAligner(adata_ref, adata_query, gene_list, n_interpolation_points)
    aligner.align_all_pairs()
    aligner.results            # list of AligmentObj, gene_list order
    aligner.results_map[gene]  # dict gene -> AligmentObj

All genes are interpolated and aligned in one batch by ``libg2g.so`` on a B200; the objects here
are thin views over the batched result arrays, with the reference's attribute names
(SURVEY.md section 8(b)).  Heavy per-gene attributes (dense DP matrices, DE statistics) are
produced on first access.
"""
import re

import numpy as np

from . import OrgAlign as orgalign
from . import TimeSeriesPreprocessor
from . import _lib
from . import engine as _engine

__version__ = "v0.2.0-b200"

_FLAG_BAD_SIGMA = 0x10


class hcolors:
    MATCH = "\033[92m"
    INSERT = "\033[91m"
    DELETE = "\033[91m"
    STOP = "\033[0m"


# ---------------------------------------------------------------------------------------------------
# alignment-string post-processing (reference: DEAnalyser, Main.py:568-835)
# ---------------------------------------------------------------------------------------------------
def _index_digits(marks):
    """Index line under an alignment row: '*' consumes the next index (digits cycle 0..9),
    '^' repeats the previous one, '-' leaves a blank; every other character (the ANSI colour
    codes embedded in the row) is skipped (Main.py:579-595)."""
    out = []
    i = 0
    for ch in marks:
        if ch == "-":
            out.append(" ")
        elif ch == "^":
            out.append(str(i - 1))
        elif ch == "*":
            if i >= 10:
                i = 0
            out.append(str(i))
            i += 1
    return "".join(out)


def alignment_visual(s):
    """(al_visual, colored_alignment_str, regions) of a five-state string (Main.py:602-712).

    ``regions`` = (S_match, T_match, S_non_match, T_non_match, abstract_regions) from the
    five-state walk over maximal runs."""
    ref_row, qry_row, colored = [], [], []
    S_match, T_match, S_non, T_non, abstract = [], [], [], [], []
    i = j = 0
    for m in re.finditer(r"M+|W+|V+|D+|I+", s):
        ch, n = m.group(0)[0], len(m.group(0))
        if ch == "M":
            S_match.append([j, j + n - 1])
            T_match.append([i, i + n - 1])
            i += n
            j += n
            ref_row.append(hcolors.MATCH + "*" * n + hcolors.STOP)
            qry_row.append(hcolors.MATCH + "*" * n + hcolors.STOP)
            abstract.append("M")
        elif ch == "V":
            T_match.append([i, i + n - 1])
            i += n
            ref_row.append("^" * n)
            qry_row.append(hcolors.MATCH + "*" * n + hcolors.STOP)
            abstract.append("V")
        elif ch == "W":
            S_match.append([j, j + n - 1])
            j += n
            ref_row.append(hcolors.MATCH + "*" * n + hcolors.STOP)
            qry_row.append("^" * n)
            abstract.append("W")
        elif ch == "I":
            T_non.append([i, i + n - 1])
            i += n
            ref_row.append("-" * n)
            qry_row.append(hcolors.INSERT + "*" * n + hcolors.STOP)
        else:  # D
            S_non.append([j, j + n - 1])
            j += n
            ref_row.append(hcolors.DELETE + "*" * n + hcolors.STOP)
            qry_row.append("-" * n)
        colour = hcolors.MATCH if ch in "MWV" else hcolors.INSERT
        colored.append(colour + ch * n + hcolors.STOP)
    a1, a2 = "".join(ref_row), "".join(qry_row)
    index_line = "".join(str(k % 10) for k in range(len(s)))
    visual = (index_line + " Alignment index \n" + _index_digits(a1) + " Reference index" + "\n" + a1 + "\n" + a2
              + "\n" + _index_digits(a2) + " Query index" + "\n" + s + " 5-state string ")
    return visual, "".join(colored), (S_match, T_match, S_non, T_non, abstract)


def matched_time_points(s):
    """Pairs of (reference bin, query bin) matched through M, W or V (Main.py:716-773).

    The reference keeps two cursors that lag behind by one on the side a warp repeats; the
    bookkeeping below restates its transitions: ``i`` / ``j`` are the next unconsumed query /
    reference bins, and entering or leaving a warp run adjusts the repeated side."""
    i = j = 0
    S_pts, T_pts = [], []
    prev = ""
    for ch in s:
        if ch in "MID":
            # leaving a warp run: the side the warp was repeating has to move on
            if prev == "W":
                i += 1
            elif prev == "V":
                j += 1
        if ch == "M":
            T_pts.append(i)
            S_pts.append(j)
            i += 1
            j += 1
        elif ch == "W":  # reference advances, query bin repeats
            if prev == "V":
                i -= 1
                j += 1
            elif prev != "W":
                i -= 1
            T_pts.append(i)
            S_pts.append(j)
            j += 1
        elif ch == "V":  # query advances, reference bin repeats
            if prev == "W":
                j -= 1
                i += 1
            elif prev != "V":
                j -= 1
            T_pts.append(i)
            S_pts.append(j)
            i += 1
        elif ch == "I":
            i += 1
        elif ch == "D":
            j += 1
        prev = ch
    return np.array(S_pts), np.array(T_pts)


def _string_table(store):
    """All 5-state strings of a batch decoded at once (cached on the result store): (text, lengths,
    stride).  Per-gene access is then a plain slice."""
    tab = store.strings
    if tab is None:
        arr = store.host["align_str"]
        # bytes past a string's length are zero; latin-1 maps every byte to one character
        tab = (arr.tobytes().decode("latin-1"), store.host["align_len"].tolist(), int(arr.shape[1]))
        store.strings = tab
    return tab


class AligmentObj:
    """Alignment result of one gene (reference: Main.py:33-177; the class name keeps the
    reference's spelling).  A view over the aligner's batched arrays: constructing it costs O(1);
    ``alignment_str``, ``S``, ``T``, ``fwd_DP``, ``landscape_obj``, the region lists and the match
    percentages are materialised on first access and then cached as plain attributes."""

    _LAZY = {
        "alignment_str": "_load_str", "fwd_DP": "_load_core", "landscape_obj": "_load_core",
        "S": "_load_series", "T": "_load_series",
        "match_regions_S": "_load_regions", "match_regions_T": "_load_regions",
        "non_match_regions_S": "_load_regions", "non_match_regions_T": "_load_regions",
        "match_percentage": "compute_series_match_percentage",
        "match_percentage_S": "compute_series_match_percentage",
        "match_percentage_T": "compute_series_match_percentage",
    }

    # class-level defaults keep per-object construction to four attribute writes
    bwd_DP = None
    _host = None
    _visual = None
    _points = None
    _de = None

    def __init__(self, store, gene_idx, gene, host=None):
        self._store = store
        self._gene_idx = gene_idx
        self.gene = gene
        if host is not None:
            self._host = host

    def __getattr__(self, name):
        loader = AligmentObj._LAZY.get(name)
        if loader is None:
            raise AttributeError(name)
        getattr(self, loader)()
        return self.__dict__[name]

    def _results(self):
        return self._store.host if self._host is None else self._host

    def _load_str(self):
        idx = self._gene_idx
        if self._host is not None:   # single-gene re-alignment (align_single_pair)
            h = self._host
            self.alignment_str = h["align_str"][idx, :int(h["align_len"][idx])].tobytes().decode("ascii")
            return
        text, lens, stride = _string_table(self._store)
        self.alignment_str = text[idx * stride:idx * stride + lens[idx]]

    def _load_series(self):
        self.S, self.T = self._store.pairs[self.gene]

    def _load_core(self):
        st, idx, h = self._store, self._gene_idx, self._results()
        nT = st.n_bins
        dp = orgalign.DP5(st, idx, self.S, self.T, self.alignment_str, np.float64(h["opt_cost"][idx]),
                          st.state_params if self._host is None else self._host["state_params"])
        self.fwd_DP = dp
        self.landscape_obj = orgalign.AlignmentLandscape(dp, h["l_states"][idx].reshape(nT + 1, nT + 1), nT, nT)

    def _load_regions(self):
        regions = orgalign.legacy_matched_regions(self.alignment_str)
        self.match_regions_S, self.match_regions_T, self.non_match_regions_S, self.non_match_regions_T = regions

    # -- percentages (Main.py:103-127) ------------------------------------------------------------
    def compute_series_match_percentage(self):
        s = self.alignment_str
        n_m, n_v, n_w = s.count("M"), s.count("V"), s.count("W")
        self.match_percentage = round(((n_m + n_v + n_w) / len(s)) * 100, 2)
        n_bins = self._store.n_bins  # len(T.data_bins) == len(S.data_bins)
        self.match_percentage_T = round(((n_m + n_v) / n_bins) * 100, 2)
        self.match_percentage_S = round(((n_m + n_w) / n_bins) * 100, 2)

    def get_series_match_percentage(self):
        return self.match_percentage, self.match_percentage_S, self.match_percentage_T

    def get_opt_alignment_cost(self):
        return np.float64(self._results()["opt_cost"][self._gene_idx])

    # -- lazily derived fields (reference computes them eagerly in run_DEAnalyser, Main.py:129-138) --
    def _ensure_visual(self):
        if self._visual is None:
            self._visual = alignment_visual(self.alignment_str)
        return self._visual

    @property
    def al_visual(self):
        return self._ensure_visual()[0]

    @property
    def colored_alignment_str(self):
        return self._ensure_visual()[1]

    def _ensure_points(self):
        if self._points is None:
            self._points = matched_time_points(self.alignment_str)
        return self._points

    @property
    def match_points_S(self):
        return self._ensure_points()[0]

    @property
    def match_points_T(self):
        return self._ensure_points()[1]

    @property
    def l2fold_changes(self):
        """[log2(mean(S bin)/mean(T bin)), S bin, T bin] per matched pair (Main.py:777-783)."""
        s, t = self._ensure_points()
        with np.errstate(all="ignore"):
            return [[np.log2(np.mean(self.S.data_bins[a]) / np.mean(self.T.data_bins[b])), a, b]
                    for a, b in zip(s, t)]

    @property
    def matched_region_DE_info(self):
        """Per matched bin pair: Wilcoxon / KS / t-test p-values and the compression statistic
        (Main.py:791-835).  Computed with scipy on first access (the reference pays this for every
        gene inside ``align_all_pairs``)."""
        if self._de is None:
            store = self._store
            if getattr(store, "de", None) is not None and self._host is None:
                from . import DEStats
                self._de = DEStats.table_for_gene(store, self._gene_idx)   # batched (compute_DE_info)
            else:
                self._de = _de_info(self)
        return self._de

    # -- printing -----------------------------------------------------------------------------------
    def print(self):
        print("Fwd opt cost", self.fwd_DP.opt_cost)
        print("5-state alignment string: ", self.fwd_DP.alignment_str)
        print("Ref matched time points ranges: ", self.match_regions_S)
        print("Query matched time points ranges: ", self.match_regions_T)
        print("Ref mismatched time points ranges: ", self.non_match_regions_S)
        print("Query mismatched time points ranges: ", self.non_match_regions_T)

    def print_alignment(self):
        print(self.al_visual)
        print("Matched percentages: ")
        p1, p2, p3 = self.get_series_match_percentage()
        print("w.r.t alignment: ", p1, "%")
        print("w.r.t ref: ", p2, "%")
        print("w.r.t query: ", p3, "%")

    def plot_mean_trends(self):
        self.S.plot_mean_trend(color="forestgreen")
        self.T.plot_mean_trend(color="midnightblue")


def _de_info(a):
    import pandas as pd
    import scipy.stats as stats

    s, t = a.match_points_S, a.match_points_T
    S_bins, T_bins = a.S.data_bins, a.T.data_bins
    cost = _engine_cost_matrix(a)
    rows = []
    with np.errstate(all="ignore"):
        for k in range(len(s)):
            l2fc = np.log2(np.mean(S_bins[s[k]]) / np.mean(T_bins[t[k]]))
            # the reference tests the *reference-side* bins s[k] and t[k] for equality (Main.py:811-814)
            if not np.any(np.asarray(S_bins[s[k]]) - np.asarray(S_bins[t[k]])):
                pw = pk = pt = 0.0
            else:
                pw = np.round(stats.wilcoxon(S_bins[s[k]], T_bins[t[k]])[1], 6)
                pk = np.round(stats.ks_2samp(S_bins[s[k]], T_bins[t[k]])[1], 6)
                pt = np.round(stats.ttest_ind(S_bins[s[k]], T_bins[t[k]])[1], 6)
            rows.append([s[k], t[k], a.S.time_points[s[k]], a.T.time_points[t[k]],
                         np.round(cost[t[k], s[k]], 7), l2fc, pw, pk, pt])
    df = pd.DataFrame(rows, columns=["ref_bin", "query_bin", "ref_pseudotime", "query_pseudotime", "Delta",
                                     "l2fc", "wilcox", "ks2", "ttest"])
    return df


def _engine_cost_matrix(a):
    """Closed form of the MML match cost (device formula, csrc/g2g_dp.cu) for one gene on the host."""
    n = _engine.N_SAMPLES
    mS, sS = a.S.mean_trend[None, :], a.S.std_trend[None, :]
    mT, sT = a.T.mean_trend[:, None], a.T.std_trend[:, None]
    d2 = (mT - mS) ** 2
    term = (sT ** 2 + d2) / sS ** 2 + (sS ** 2 + d2) / sT ** 2 - 2.0
    return (n / 2.0 * term - 2.0 * _engine.model_constant(n) + np.log(sS) + np.log(sT)) / (4.0 * n)


# ---------------------------------------------------------------------------------------------------
class _LazyPairs(dict):
    """``aligner.pairs``: gene -> [S, T] interpolated series, filled on demand from the batch
    (the reference memoises ``run_interpolation`` per gene here, Main.py:417-418)."""

    def __init__(self, store):
        super().__init__()
        import weakref
        self._store = weakref.proxy(store)   # no reference cycle: a dropped aligner frees its staging buffer at once

    def __missing__(self, gene):
        st = self._store.series_pair(self._store.gene_index[gene])
        self[gene] = st
        return st


class _ResultStore:
    """Host copy of one batch's compact results plus the little state needed to serve every
    ``AligmentObj`` attribute.  It owns no batch device buffers and does not reference the aligner, so
    dropping the aligner frees its GPU memory at once while result objects stay usable."""

    def __init__(self, host, gene_list, n_bins, eps, points_ref, points_query, state_params, device):
        self.host = host
        self.gene_list = gene_list
        self.gene_index = {g: k for k, g in enumerate(gene_list)}
        self.n_bins = int(n_bins)
        self.eps = eps
        self.points_ref, self.points_query = points_ref, points_query
        self.state_params = tuple(state_params)
        self.device = device
        self.de = None
        self.strings = None
        self.pairs = _LazyPairs(self)

    def series_pair(self, idx):
        h = self.host
        ms, tr = h["musigma"][idx], h["trends"][idx]
        gene = self.gene_list[idx]
        S = TimeSeriesPreprocessor.SummaryTimeSeries(gene, ms[0], ms[1], tr[0], tr[1], self.points_ref, self.eps)
        T = TimeSeriesPreprocessor.SummaryTimeSeries(gene, ms[2], ms[3], tr[2], tr[3], self.points_query, self.eps)
        return [S, T]

    def device_trends(self, idx):
        import torch
        return torch.as_tensor(np.ascontiguousarray(self.host["trends"][idx:idx + 1])).to(self.device)

    def run_dp(self, idx, state_params, dump):
        import torch
        with torch.cuda.device(self.device):
            out = _engine.run_cost_dp(self.device_trends(idx), self.n_bins, _engine.transition_costs(state_params),
                                      _engine.start_costs(), _engine.model_constant(), dump=dump)
            torch.cuda.current_stream().synchronize()
        return out

    def dense_dump(self, idx, state_params):
        """Dense DP matrices / back-pointers / cost terms of one gene (K4 in dump mode)."""
        out = self.run_dp(idx, state_params, dump=True)
        d = {k: out[k][0].cpu().numpy() for k in ("cost", "mats", "bp", "L")}
        # null / match_len terms of OrgAlign.py:486-494 in closed form
        tr = self.host["trends"][idx]
        n = _engine.N_SAMPLES
        K = n / 2.0 * np.log(2 * np.pi) + 1.0 - n * np.log(0.001)
        C = _engine.model_constant(n)
        mS, sS, mT, sT = tr[0][None, :], tr[1][None, :], tr[2][:, None], tr[3][:, None]
        d2 = (mT - mS) ** 2
        I_RS = K + n * np.log(sS) + n / 2.0
        I_QT = K + n * np.log(sT) + n / 2.0
        I_QS = K + n * np.log(sS) + n * (sT ** 2 + d2) / (2 * sS ** 2)
        I_RT = K + n * np.log(sT) + n * (sS ** 2 + d2) / (2 * sT ** 2)
        mod_S, mod_T = C - np.log(sS), C - np.log(sT)
        null = (mod_S + I_RS + mod_T + I_QT) / (2 * n) + 0 * d2
        match = 0.5 * ((mod_S + I_QS + I_RS) + (mod_T + I_RT + I_QT)) / (2 * n)
        d["util"] = np.stack([null, match, d["cost"]])
        return d


class RefQueryAligner:
    """Entry point of the alignment (reference: Main.py:180-564).

    ``RefQueryAligner(adata_ref, adata_query, gene_list, n_interpolation_points[, adaptive_kernel])``.
    ``adata_*`` need ``.X`` (dense array, torch tensor or scipy sparse; cells x genes),
    ``.obs['time']`` in [0, 1], ``.var_names`` and ``.obs_names``.
    """

    #: scipy-sparse inputs at most this dense stay sparse on the device (K1s: work and memory ~ stored
    #: entries; measured break-even with the dense kernel near 9 % on B200); denser
    #: ones are densified on the device and streamed by K1 / K1m.  Precision: K1s sums ``w x^2`` un-shifted
    #: in f32, so the relative error of a bin's variance is about 1e-7 * (1 + mean^2 / variance).  At or below
    #: the default threshold mean^2 / variance is bounded by the density itself (a mostly-zero column), so the
    #: sparse path is as accurate as the dense one; raising the threshold (1.0 forces the sparse path, e.g.
    #: when the dense matrix would not fit in HBM) is only safe for data whose columns are not "high mean,
    #: low variance" -- dense, nearly constant genes lose variance digits to cancellation there.
    sparse_density_threshold = 0.08

    def __init__(self, *args, device=None, verbose=True, distributed=False):
        """``distributed=True`` (inside an initialised ``torch.distributed`` job, one process per GPU):
        this rank aligns its contiguous block of ``gene_list`` and ``align_all_pairs`` gathers every
        rank's result records to rank 0, whose ``results`` then cover the whole gene list; the other
        ranks keep the results of their own block."""
        self._verbose = verbose
        self._dist = None
        if distributed:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self._dist = (dist.get_rank(), dist.get_world_size())
                self._verbose = verbose and self._dist[0] == 0
        self._say("===============================================================================================================")
        self._say("Genes2Genes (" + __version__ + ")")
        self._say("Dynamic programming alignment of gene pseudotime trajectories using a bayesian information-theoretic framework")
        self._say("===============================================================================================================")
        if len(args) == 4:
            adaptive_kernel = False
        elif len(args) == 5:
            self._say("Running in adaptive interpolation mode")
            adaptive_kernel = args[4]
        else:
            raise TypeError("RefQueryAligner expects (adata_ref, adata_query, gene_list, n_interpolation_points"
                            "[, adaptive_kernel])")
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("genes2genes_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _lib.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        adata_ref, adata_query, gene_list, n_bins = args[0], args[1], args[2], args[3]
        self.adata_ref_full, self.adata_query_full = adata_ref, adata_query
        self._all_genes = list(gene_list)
        self._shard = (0, len(gene_list))
        self._capacity = None
        if self._dist is not None:
            from . import distributed as _distributed
            self._shard = _distributed.shard_range(len(gene_list), *self._dist)
            self._capacity = _distributed.max_shard(len(gene_list), self._dist[1])
            gene_list = self._all_genes[self._shard[0]:self._shard[1]]
        self._local_genes = list(gene_list)
        self.gene_list = gene_list
        self.n_artificial_time_points = int(n_bins)
        self.ref_time = np.asarray(adata_ref.obs["time"], dtype=np.float64)
        self.query_time = np.asarray(adata_query.obs["time"], dtype=np.float64)
        self._store = None
        self.state_params = [0.99, 0.1, 0.7]  # same defaults as the reference (Main.py:218)
        self.no_extreme_cases = False

        k = 1
        try:
            import scipy.sparse as _sp
            self._both_sparse = _sp.issparse(adata_ref.X) and _sp.issparse(adata_query.X)
        except Exception:
            self._both_sparse = False
        self._engine = None
        has_genes = len(self._local_genes) > 0   # a rank of a multi-GPU job may own no gene (fewer genes than ranks)
        # Order matters for the end-to-end time: the pseudotime sort of a dataset is all its upload needs, so each
        # host->device stream is started first and the kernel-width computation, the second dataset and the
        # engine's buffers are set up while the DMA engine works.
        with torch.cuda.device(self.device):
            up_r = self._begin_pack(adata_ref) if has_genes else None      # PCIe starts before the time sort
            self.TrajInt_R = TimeSeriesPreprocessor.TrajectoryInterpolator(
                self.ref_time, n_bins, adaptive_kernel=adaptive_kernel, raising_degree=k, gene_list=gene_list,
                obs_names=getattr(adata_ref, "obs_names", None))
            Xr = self._pack(adata_ref, self.TrajInt_R.order, up_r) if has_genes else None
            up_q = self._begin_pack(adata_query) if has_genes else None
            self.TrajInt_Q = TimeSeriesPreprocessor.TrajectoryInterpolator(
                self.query_time, n_bins, adaptive_kernel=adaptive_kernel, raising_degree=k, gene_list=gene_list,
                obs_names=getattr(adata_query, "obs_names", None))
            Xq = self._pack(adata_query, self.TrajInt_Q.order, up_q) if has_genes else None
            self.TrajInt_R.run()
            self.TrajInt_Q.run()
            if has_genes:
                if isinstance(Xr, _engine.CscMatrix) != isinstance(Xq, _engine.CscMatrix):
                    # one side fell over the density threshold: densify the sparse one too
                    self._both_sparse = False
                    if isinstance(Xr, _engine.CscMatrix):
                        Xr = self._pack(adata_ref, self.TrajInt_R.order)
                    else:
                        Xq = self._pack(adata_query, self.TrajInt_Q.order)
                self._engine = _engine.AlignmentEngine(
                    Xr, self.TrajInt_R.cell_pseudotimes, Xq, self.TrajInt_Q.cell_pseudotimes, len(self._local_genes),
                    self.n_artificial_time_points, self.TrajInt_R.kernel_denominators(),
                    self.TrajInt_Q.kernel_denominators(), self.state_params, self.device, capacity=self._capacity)
        self._say("Interpolator initialization completed")
        self._interp = None
        self._say("Aligner initialised to align trajectories of", adata_ref.shape[0], "reference cells &",
                  adata_query.shape[0], "query cells in terms of", len(self.gene_list), "genes")

    def _say(self, *a):
        if self._verbose:
            print(*a)

    # -- input packing ------------------------------------------------------------------------------
    def _local_columns(self, adata):
        """Where this aligner's (this rank's) genes sit in ``adata``: (col_lo, col_hi, idx) with ``idx`` the
        positions relative to ``col_lo`` (None when the genes are exactly the contiguous block).  Only the
        local genes need to be present in ``adata`` -- a rank of a multi-GPU job may load just its block."""
        var_names = adata.var_names
        local = self._local_genes
        if hasattr(var_names, "get_indexer"):          # pandas Index: one vectorised hash lookup
            cols = np.asarray(var_names.get_indexer(local), dtype=np.int64)
        else:
            pos = {str(g): k for k, g in enumerate(var_names)}
            cols = np.asarray([pos.get(g, -1) for g in local], dtype=np.int64)
        if len(cols) and cols.min() < 0:
            raise KeyError("gene %r of gene_list is not in var_names" % local[int(np.argmin(cols))])
        lo, hi = int(cols.min()), int(cols.max()) + 1
        if hi - lo == len(cols) and np.array_equal(cols, np.arange(lo, hi)):
            return lo, hi, None
        return lo, hi, (cols - lo).astype(np.int32)

    def _begin_pack(self, adata):
        """Start the host->device stream of a dense host ``adata.X`` (None for sparse / device inputs, which
        ``_pack`` handles in one go): the first chunks are in flight while the caller sorts the pseudotimes."""
        import torch
        X = adata.X
        try:
            import scipy.sparse as sp
            if sp.issparse(X):
                return None
        except Exception:
            pass
        if hasattr(X, "todense") and not hasattr(X, "__array_interface__") and not hasattr(X, "data_ptr"):
            return None
        if isinstance(X, np.matrix) or (torch.is_tensor(X) and X.device.type == "cuda"):
            return None
        col_lo, col_hi, idx = self._local_columns(adata)
        return _engine.HostBlockUpload(X, col_lo, col_hi, idx, self.device).start()

    def _pack(self, adata, order, started=None):
        """``adata[order][:, gene_list].X`` as a packed f32 device matrix (or a ``CscMatrix``)."""
        import torch
        identity = bool(np.all(order[1:] > order[:-1])) if len(order) > 1 else True
        if started is not None:
            return started.finish(None if identity else order)
        col_lo, col_hi, idx = self._local_columns(adata)
        X = adata.X
        row_order = None if identity else order
        try:
            import scipy.sparse as sp
            is_sparse = sp.issparse(X)
        except Exception:  # scipy always ships with the reference's dependencies; be permissive
            is_sparse = False
        if is_sparse:
            gene_idx = None
            if not (col_lo == 0 and col_hi == X.shape[1] and idx is None):
                gene_idx = (np.arange(col_lo, col_hi) if idx is None else col_lo + idx).astype(np.int32)
            X = X.tocsr()
            if not X.has_canonical_format:   # duplicate (row, col) entries add up, like .todense()
                X = X.copy()
                X.sum_duplicates()
            density = X.nnz / max(1, X.shape[0] * X.shape[1])
            if self._both_sparse and density <= self.sparse_density_threshold:
                return _engine.pack_csc(X, row_order, gene_idx, X.shape[1], self.device)
            return _engine.pack_csr(X, row_order, gene_idx, X.shape[1], self.device)
        if hasattr(X, "todense") and not hasattr(X, "__array_interface__") and not hasattr(X, "data_ptr"):
            X = X.todense()
        if isinstance(X, np.matrix):
            X = np.asarray(X)
        if torch.is_tensor(X) and X.device.type == "cuda":   # device-resident input: gather kernel
            blk = X[:, col_lo:col_hi]
            return _engine.pack_dense(blk, row_order, idx, self.device)
        return _engine.upload_host_block(X, row_order, col_lo, col_hi, idx, self.device)

    @property
    def ref_mat(self):
        return self._dense_frame(self.adata_ref_full)

    @property
    def query_mat(self):
        return self._dense_frame(self.adata_query_full)

    @staticmethod
    def _dense_frame(adata):
        """cells x genes DataFrame of the *un-subset* matrix, as the reference keeps (Main.py:229-247)."""
        import pandas as pd
        X = adata.X
        X = X.todense() if hasattr(X, "todense") else X
        if hasattr(X, "cpu"):
            X = X.cpu().numpy()
        return pd.DataFrame(np.asarray(X), columns=adata.var_names, index=adata.obs_names)

    # -- stages ---------------------------------------------------------------------------------------
    @property
    def _host(self):
        return None if self._store is None else self._store.host

    @property
    def pairs(self):
        """gene -> [S, T] interpolated series (reference: ``self.pairs``, Main.py:257, 417-418)."""
        if self._store is None:
            self._run_batch()
        return self._store.pairs

    def _run_batch(self, on_enqueued=None, store=None):
        """Enqueue the batch, copy the compact results back and build (or fill ``store``, a result store
        made ahead of time) the result store.  ``on_enqueued`` (single-process runs) is called after
        everything is queued and before the host waits for the device: host-side work done there
        overlaps the GPU."""
        import torch
        eng = self._engine
        with torch.cuda.device(self.device):
            if eng is not None:
                eng.set_state_params(self.state_params)
                if self._interp is None:
                    eng.run_interpolation(with_dp=True)   # interpolation + fix-up + DP, fused launch
                    self._interp = True
                else:
                    eng.run_dp()                          # new state_params: only the DP is repeated
            if self._dist is None:
                eng.begin_results_to_host()
                if on_enqueued is not None:
                    on_enqueued()
                host = eng.finish_results_to_host()
            else:
                if on_enqueued is not None:
                    on_enqueued()
                if eng is not None:
                    eng.redo_if_out_of_range()
                host = self._gather_results()
        if store is None:
            store = self._new_store(host)
        else:
            store.host = host
        self._store = store
        bad = np.nonzero(host["flags"] & _FLAG_BAD_SIGMA)[0]
        if len(bad):
            g = self.gene_list[int(bad[0])]
            raise ValueError("Expected parameter scale of the interpolated distribution to be positive "
                             "(GreaterThan(lower_bound=0.0)); gene %r has a zero or invalid interpolated "
                             "standard deviation (%d gene(s) affected)" % (g, len(bad)))
        return host

    def _new_store(self, host):
        return _ResultStore(host, self.gene_list, self.n_artificial_time_points,
                            _engine.epsilon_table(self.n_artificial_time_points),
                            self.TrajInt_R.interpolation_points, self.TrajInt_Q.interpolation_points,
                            self.state_params, self.device)

    peer_gather = True   # use rank 0's peer-mapped window when the box allows it (else NCCL / gloo gather)

    def _gather_results(self):
        """The one collective of the path: every rank's packed per-gene records go to rank 0 (equally sized
        byte buffers laid out for the largest shard): peer copies into rank 0's window
        (``distributed.PeerWindow``) on one box, an NCCL gather otherwise.  Returns rank 0's host arrays for ALL
        genes, the own block's on the other ranks."""
        import torch
        import torch.distributed as dist
        from . import distributed as D
        rank, world = self._dist
        eng, T = self._engine, self.n_artificial_time_points
        n_all = len(self._all_genes)
        cap = self._capacity
        n_bytes = D.record_bytes(cap, T)
        if eng is not None:
            send = eng.packed
        else:   # empty shard: join the collective with an all-zero buffer
            send = torch.zeros(max(n_bytes, 8), dtype=torch.uint8, device=self.device)
        window = D.peer_window(max(n_bytes, 8), self.device) if self.peer_gather else None
        on_host = window is None and dist.get_backend() != "nccl"   # gloo without peer mapping: host buffers
        recv = None
        if window is not None:
            # every rank copies its records into rank 0's window (one device-to-device copy over NVLink, no kernel);
            # first barrier: rank 0 has read the previous batch out of the window, second: every slot has landed
            dist.barrier()
            window.push(send)
            torch.cuda.current_stream().synchronize()
            dist.barrier()
            if rank == 0:
                recv = window.rows()
        else:
            if on_host:
                torch.cuda.current_stream().synchronize()
                send = send.cpu()
            recv = torch.empty((world, send.numel()), dtype=torch.uint8, device=send.device) if rank == 0 else None
            dist.gather(send, list(recv.unbind(0)) if rank == 0 else None, dst=0)
        if rank != 0:
            if eng is None:
                empty = np.zeros(max(n_bytes, 8), dtype=np.uint8)
                return _engine.HostResults([(empty, cap, 0, 0)], 0, T)
            return eng.results_to_host()
        segments = []
        for r in range(world):
            lo, hi = D.shard_range(n_all, r, world)
            segments.append((None, cap, hi - lo, lo))
        host_buf = None
        if on_host:
            raw = recv.numpy()
        else:
            host_buf = _engine._take_pinned(recv.numel())
            host_buf.copy_(recv.reshape(-1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            raw = host_buf.numpy().reshape(world, -1)
        segments = [(raw[r], c, held, lo) for r, (_, c, held, lo) in enumerate(segments)]
        self.gene_list = self._all_genes                    # rank 0 now holds every gene's result
        self._shard = (0, n_all)
        # per-key arrays covering every gene, filled straight from the (pinned) receive buffer: the small ones now,
        # the large ones on first access
        return _engine.HostResults(segments, n_all, T, pinned=host_buf)

    def _make_result(self, idx, host=None):
        return AligmentObj(self._store, idx, self.gene_list[idx], host)

    def run_interpolation(self, gene_idx):
        """[S, T] interpolated series of one gene (reference: Main.py:307-410)."""
        return self.pairs[self.gene_list[gene_idx]]

    def align_single_pair(self, gene_idx):
        """Align one gene with the current ``state_params`` (reference: Main.py:413-432); the
        interpolation of the batch is reused (the reference memoises it in ``pairs``)."""
        if self._store is None:
            self._run_batch()
        params = tuple(self.state_params)
        out = self._store.run_dp(gene_idx, params, dump=False)
        view = {k: _Shift(out[k].cpu().numpy(), gene_idx) for k in ("align_str", "align_len", "opt_cost", "l_states")}
        view["state_params"] = params
        return self._make_result(gene_idx, view)

    def align_all_pairs(self, concurrent=False, n_processes=None):
        """Align every gene of ``gene_list`` (reference: Main.py:434-457).  ``concurrent`` /
        ``n_processes`` are accepted for signature compatibility; the batch already runs all genes
        concurrently on the GPU."""
        self._say("Running gene-level alignment: \U0001F9EC")
        by_gene = None
        if self._dist is not None:
            # rank 0 serves every gene after the gather: its result views are built while the GPUs work
            genes = self._all_genes if self._dist[0] == 0 else self.gene_list
            store = _ResultStore(None, genes, self.n_artificial_time_points,
                                 _engine.epsilon_table(self.n_artificial_time_points),
                                 self.TrajInt_R.interpolation_points, self.TrajInt_Q.interpolation_points,
                                 self.state_params, self.device)
            built, by_gene = [], {}

            def build_views():
                built.extend(AligmentObj(store, k, g) for k, g in enumerate(genes))
                by_gene.update((a.gene, a) for a in built)
            self._run_batch(on_enqueued=build_views, store=store)   # rank 0's gene_list grows to all genes here
            self.results = built
        else:
            # the result objects are O(1) views: build them while the GPU is still working
            store = self._new_store(None)
            built, by_gene = [], {}

            def build_views():
                built.extend(AligmentObj(store, k, g) for k, g in enumerate(self.gene_list))
                by_gene.update((a.gene, a) for a in built)
            self._run_batch(on_enqueued=build_views, store=store)
            self.results = built
        text, lens, stride = _string_table(store)
        for k, a in enumerate(self.results):                    # strings are cheap to cut here, dear to cut lazily
            a.alignment_str = text[k * stride:k * stride + lens[k]]
        self.results_map = by_gene if by_gene is not None else {a.gene: a for a in self.results}
        self._say("Alignment completed! ✅")

    def compute_DE_info(self):
        """Compute ``matched_region_DE_info`` of every gene in one batch (rank statistics on the GPU,
        p-values from scipy's null distributions; reference: Main.py:791-835 per gene).  Without this
        call each result builds its own table with scipy on first access."""
        from . import DEStats
        if self._store is None:
            self._run_batch()
        DEStats.compute_all(self._store)
        return self._store.de

    # -- aggregate alignment (reference: Main.py:498-510) ---------------------------------------------
    def get_aggregate_alignment(self):
        """Cell-level average alignment over all genes; sets ``average_alignment`` and returns
        (string, path, match-count DataFrame).  The reference also draws the path on a heat map."""
        return self._aggregate(self.gene_list, True)

    def get_aggregate_alignment_for_subset(self, gene_subset):
        return self._aggregate(gene_subset, False)

    def _aggregate(self, genes, remember):
        from . import ClusterUtils
        average_alignment, alignment_path = ClusterUtils.get_cluster_average_alignments(self, genes)
        mat = ClusterUtils.get_pairwise_match_count_mat(self, genes)
        self._say("Average Alignment: ", alignment_visual(average_alignment)[1], "(cell-level)")
        if remember:
            pct = np.round(len(re.findall("M|W|V", average_alignment)) * 100 / len(average_alignment), 2)
            self._say("% similarity:", pct)
            self.average_alignment = average_alignment
        return average_alignment, alignment_path, mat

    # -- summaries ------------------------------------------------------------------------------------
    def get_match_stat_for_all_genes(self):
        import pandas as pd
        rows = [a.get_series_match_percentage() for a in self.results]
        df = pd.DataFrame(rows, columns=["match %", "match % S", "match % T"])
        df["cluster_id"] = getattr(self, "cluster_ids", [None] * len(rows))
        return df

    def get_stat_df(self, plot=True):
        """Per-gene summary table of the reference (``get_stat_df``, Main.py:459-491): gene, alignment
        similarity, optimal cost, log2 fold change of the mean expression, colour flag, sorted the same
        way; prints the mean similarity.  The two plots are drawn when matplotlib / seaborn are installed
        (``plot=False`` skips them)."""
        import pandas as pd
        ref_mat, query_mat = self.ref_mat, self.query_mat
        rows = []
        with np.errstate(all="ignore"):
            for g in self.gene_list:
                a = self.results_map[g]
                rgex, qgex = np.asarray(ref_mat.loc[:, g]), np.asarray(query_mat.loc[:, g])
                rows.append([g, a.match_percentage / 100, a.fwd_DP.opt_cost, np.log2(np.mean(rgex) / np.mean(qgex))])
        df = pd.DataFrame(rows, columns=["Gene", "alignment_similarity_percentage", "opt_alignment_cost", "l2fc"])
        df["color"] = np.repeat("green", df.shape[0])
        df.loc[df["alignment_similarity_percentage"] <= 0.5, "color"] = "red"
        df["abs_l2fc"] = np.abs(df["l2fc"])
        df = df.sort_values(["alignment_similarity_percentage", "abs_l2fc"], ascending=[True, False])
        print("Mean alignment similarity percentage (matched %): ")
        print(round(np.mean(df["alignment_similarity_percentage"]), 4) * 100, "%")
        if plot:
            try:
                import matplotlib.pyplot as plt
                import seaborn as sb
            except Exception:
                return df
            plt.subplots(1, 2, figsize=(10, 4))
            plt.subplot(1, 2, 1)
            sb.kdeplot(np.asarray([x * 100 for x in df["alignment_similarity_percentage"]]), fill=True,
                       label="Alignment Similarity %")
            plt.xlim([0, 100])
            plt.xlabel("Alignment similarity percentage", fontsize="12")
            plt.ylabel("Density", fontsize="12")
            plt.subplot(1, 2, 2)
            sb.scatterplot(x=df["alignment_similarity_percentage"] * 100, y=df["opt_alignment_cost"],
                           hue=df["color"], palette={"green": "green", "red": "red"}, legend=False)
            plt.xlabel("Alignment similarity percentage", fontsize="12")
            plt.ylabel("Optimal alignment cost (nits)", fontsize="12")
        return df

    def get_stat_table(self):
        """Gene / alignment similarity / optimal cost table (the numeric part of the reference's
        ``get_stat_df``, Main.py:459-477, without the plots)."""
        import pandas as pd
        return pd.DataFrame({
            "Gene": list(self.gene_list),
            "alignment_similarity_percentage": [a.match_percentage / 100 for a in self.results],
            "opt_alignment_cost": [a.fwd_DP.opt_cost for a in self.results],
        })


class _Shift:
    """Index adaptor so a one-gene result array can be addressed with the gene's global index."""

    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self._arr[(key[0] - self._idx,) + key[1:]]
        return self._arr[key - self._idx]
