"""Batched ``matched_region_DE_info`` for all genes at once.

The reference builds this table inside every ``AligmentObj.__init__`` with three scipy calls per matched
bin pair (Main.py:791-835) -- 10-16 % of its wall time.  Here the rank statistics of all pairs of all
genes come from one kernel (``g2g_de_statistics``), and the p-values are looked up with scipy's own null
distributions (exact signed-rank distribution, exact two-sample KS probabilities, Student t), vectorised
over pairs.  Values equal the reference's per-pair scipy results to rounding (they are rounded to six
decimals in the table, Main.py:823-825).
"""
import ctypes

import numpy as np

from . import _lib
from . import engine as _engine

N = _engine.N_SAMPLES
_KS_TABLE = {}


def _ks_exact_table(n=N):
    """p-value of the exact two-sided two-sample KS test for D = h / n, h = 0..n (n = m), from scipy
    itself on samples constructed to have that statistic."""
    if n not in _KS_TABLE:
        import scipy.stats as stats
        x = np.arange(n, dtype=np.float64)
        tab = np.ones(n + 1)
        for h in range(1, n + 1):
            tab[h] = stats.ks_2samp(x, x + (h - 0.5))[1]
        _KS_TABLE[n] = tab
    return _KS_TABLE[n]


def _wilcoxon_pvalues(r_plus, count, tie_term, n=N):
    """Two-sided p-values following scipy.stats.wilcoxon(method='auto', correction=False)."""
    from scipy import special
    r_plus = np.asarray(r_plus, dtype=np.float64)
    count = np.asarray(count, dtype=np.float64)
    p = np.empty_like(r_plus)
    exact = (tie_term == 0) & (count == n)
    if exact.any():
        try:
            from scipy.stats._wilcoxon import WilcoxonDistribution
            dist = WilcoxonDistribution(np.full(int(exact.sum()), n))
            rp = r_plus[exact]
            pe = 2 * np.minimum(dist.sf(np.floor(rp)), dist.cdf(np.ceil(rp)))
            p[exact] = np.clip(pe, 0, 1)
        except Exception:  # scipy without the vectorised distribution: exact null by dynamic programme
            counts = np.zeros(n * (n + 1) // 2 + 1)
            counts[0] = 1.0
            for k in range(1, n + 1):
                counts[k:] = counts[k:] + counts[:-k].copy()
            pmf = counts / counts.sum()
            cdf = np.cumsum(pmf)
            sf = np.cumsum(pmf[::-1])[::-1]
            rp = r_plus[exact]
            p[exact] = np.clip(2 * np.minimum(sf[np.floor(rp).astype(int)], cdf[np.ceil(rp).astype(int)]), 0, 1)
    approx = ~exact
    if approx.any():
        c = count[approx]
        mn = c * (c + 1.0) * 0.25
        with np.errstate(all="ignore"):
            se = np.sqrt((c * (c + 1.0) * (2.0 * c + 1.0) - tie_term[approx] / 2) / 24)
            z = (r_plus[approx] - mn) / se
            p[approx] = 2 * np.minimum(special.ndtr(z), special.ndtr(-z))
    return p


def compute_all(store):
    """Fill ``store.de`` with the DE table columns of every gene (arrays over all matched pairs plus
    per-gene offsets)."""
    import torch
    from scipy import special
    from .Main import matched_time_points

    host, T = store.host, store.n_bins
    G = len(store.gene_list)
    S_pts, T_pts, offs = [], [], [0]
    for g in range(G):
        n = int(host["align_len"][g])
        s = host["align_str"][g, :n].tobytes().decode("ascii")
        ps, pt = matched_time_points(s)
        S_pts.append(ps)
        T_pts.append(pt)
        offs.append(offs[-1] + len(ps))
    s_idx = np.concatenate(S_pts).astype(np.int32) if offs[-1] else np.zeros(0, np.int32)
    t_idx = np.concatenate(T_pts).astype(np.int32) if offs[-1] else np.zeros(0, np.int32)
    gene = np.repeat(np.arange(G, dtype=np.int32), np.diff(offs))
    P = len(gene)
    lib = _lib.load()
    dev = store.device
    with torch.cuda.device(dev):
        ms_d = torch.from_numpy(np.ascontiguousarray(host["musigma"])).to(dev)
        eps_d = torch.from_numpy(np.array(store.eps, dtype=np.float64)).to(dev)
        g_d, s_d, t_d = (torch.from_numpy(a).to(dev) for a in (gene, s_idx, t_idx))
        rp_d = torch.empty(P, dtype=torch.float64, device=dev)
        tie_d = torch.empty(P, dtype=torch.float64, device=dev)
        nz_d = torch.empty(P, dtype=torch.int32, device=dev)
        h_d = torch.empty(P, dtype=torch.int32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.g2g_de_statistics(ms_d.data_ptr(), eps_d.data_ptr(), T, g_d.data_ptr(), s_d.data_ptr(),
                                         t_d.data_ptr(), P, rp_d.data_ptr(), tie_d.data_ptr(), nz_d.data_ptr(),
                                         h_d.data_ptr(), st), "g2g_de_statistics")
        r_plus, tie, nz, h = (a.cpu().numpy() for a in (rp_d, tie_d, nz_d, h_d))
    ms, tr = host["musigma"], host["trends"]
    muS, sdS = ms[gene, 0, s_idx], ms[gene, 1, s_idx]
    muT, sdT = ms[gene, 2, t_idx], ms[gene, 3, t_idx]
    eps = store.eps
    e_mean, e_var = eps.mean(axis=1), eps.var(axis=1, ddof=1)
    with np.errstate(all="ignore"):
        # Student t (scipy.stats.ttest_ind, equal variances): bins are mu + sd * eps_row
        m1, m2 = muS + sdS * e_mean[s_idx], muT + sdT * e_mean[t_idx]
        v1, v2 = sdS ** 2 * e_var[s_idx], sdT ** 2 * e_var[t_idx]
        df = 2 * N - 2
        svar = ((N - 1) * v1 + (N - 1) * v2) / df
        tstat = (m1 - m2) / np.sqrt(svar * (2.0 / N))
        p_t = 2 * special.stdtr(df, -np.abs(tstat))
        p_w = _wilcoxon_pvalues(r_plus, nz, tie)
        p_k = _ks_exact_table()[h]
        l2fc = np.log2(tr[gene, 0, s_idx] / tr[gene, 2, t_idx])
        # Delta: match cost of the pair, closed form of the device kernel (csrc/g2g_dp.cu)
        mS, sS, mT, sT = tr[gene, 0, s_idx], tr[gene, 1, s_idx], tr[gene, 2, t_idx], tr[gene, 3, t_idx]
        d2 = (mT - mS) ** 2
        term = (sT ** 2 + d2) / sS ** 2 + (sS ** 2 + d2) / sT ** 2 - 2.0
        delta = (N / 2.0 * term - 2.0 * _engine.model_constant(N) + np.log(sS) + np.log(sT)) / (4.0 * N)
    # the reference tests the *reference-side* bins s and t for equality (Main.py:811-814): equal only
    # for s == t, where all three p-values are reported as 0
    same = s_idx == t_idx
    p_w, p_k, p_t = (np.where(same, 0.0, np.round(v, 6)) for v in (p_w, p_k, p_t))
    store.de = {"offsets": np.asarray(offs), "ref_bin": s_idx, "query_bin": t_idx,
                "ref_pseudotime": store.points_ref[s_idx], "query_pseudotime": store.points_query[t_idx],
                "Delta": np.round(delta, 7), "l2fc": l2fc, "wilcox": p_w, "ks2": p_k, "ttest": p_t}
    return store.de


COLUMNS = ["ref_bin", "query_bin", "ref_pseudotime", "query_pseudotime", "Delta", "l2fc", "wilcox", "ks2", "ttest"]


def table_for_gene(store, idx):
    import pandas as pd
    de = store.de
    lo, hi = de["offsets"][idx], de["offsets"][idx + 1]
    return pd.DataFrame({c: np.asarray(de[c][lo:hi], dtype=np.float64) for c in COLUMNS}, columns=COLUMNS)
