"""CPU oracle for the Genes2Genes per-gene alignment hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module; the
product package ``genes2genes_b200`` never does and fails loudly without its CUDA
library.

What it is: a plain numpy / pure-Python float64 restatement of the reference's
algorithm (``/root/reference/genes2genes`` v0.2.0), one function per reference
function, each citing the file:line it follows.  ``TSP.py`` abbreviates
``TimeSeriesPreprocessor.py``.

Parity pin: the reference has no tests of its own.  This oracle is pinned against
(i) outputs of the UNMODIFIED reference run in the build container on the bundled
tutorial data and on synthetic sparse data (``tests/golden/*.npz`` made by
``tests/golden/make_golden.py``) and (ii) the outputs the authors stored in
``notebooks/Tutorial.ipynb`` (``tests/golden/tutorial_published.json``); see
``tests/test_oracle_golden.py``.

Two levels are provided:
  * "faithful" per-gene functions that mirror the reference's arithmetic step by step
    (sums over the 50 samples of a bin, scalar DP loops);
  * "batched" numpy functions (closed-form cost, DP vectorised over genes) used to check
    the CUDA path at sizes where per-gene Python loops would take minutes.
"""
import math

import numpy as np

N_SAMPLES = 50  # TSP.py:179  generate_random_dataset(50, ...)
STATE_LETTERS = "MWVDI"  # candidate order everywhere in OrgAlign.py:352-389, 546-551, 743-744


# --------------------------------------------------------------------------------------
# Stage 0: fixed epsilon table  (MyFunctions.py:54-69 called from TSP.py:179 after TSP.py:184)
# --------------------------------------------------------------------------------------
def epsilon_table(n_bins, n_samples=N_SAMPLES, seed=1):
    """``eps[t, n]``: the standard-normal draws behind every ``data_bins[t][n]``.

    ``interpolate_gene_v2`` re-seeds torch with 1 (TSP.py:184) and then, bin by bin,
    ``generate_random_dataset`` makes ``n_samples`` *scalar* ``Normal(mu, sigma).rsample()``
    calls (MyFunctions.py:64-67); ``rsample`` is ``loc + eps * scale`` with ``eps`` one 0-d
    ``torch.normal`` draw.  So ``data_bins[t][n] == mu_t + eps[t, n] * sigma_t`` for every
    gene.  A private generator seeded identically yields the same mt19937 stream as the
    global one without touching global RNG state.
    """
    import torch

    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    zero = torch.zeros((), dtype=torch.float64)
    one = torch.ones((), dtype=torch.float64)
    out = np.empty((n_bins, n_samples), dtype=np.float64)
    for t in range(n_bins):
        for n in range(n_samples):
            out[t, n] = float(torch.normal(zero, one, generator=gen))
    return out


# --------------------------------------------------------------------------------------
# Stage I: interpolation  (TSP.py)
# --------------------------------------------------------------------------------------
def interpolation_points(n_bins):
    """TSP.py:23."""
    return np.linspace(0, 1, n_bins)


def sort_cells(time):
    """TSP.py:20 -- ``adata[np.argsort(adata.obs['time'])]``."""
    return np.argsort(np.asarray(time))


def adaptive_window_denominators(cell_pseudotimes, n_bins, kernel_window=0.1, k=1):
    """TSP.py:53-119 (compute_cell_densities + compute_adaptive_window_denominator).

    Returns the per-bin squared window sizes used as Gaussian-kernel denominators.
    The min-max scaling (sklearn ``MinMaxScaler`` at TSP.py:94-95) is restated directly.
    """
    p = interpolation_points(n_bins)
    tau = np.asarray(cell_pseudotimes, dtype=np.float64)
    range_mid = p[2] - p[0]
    range_corner = p[1] - p[0]
    dens = []
    for i in range(n_bins):
        if i == 0:
            logic = tau <= p[i + 1]
            rl = range_corner
        elif i == n_bins - 1:
            logic = tau >= p[i - 1]
            rl = range_corner
        else:
            logic = np.logical_and(tau <= p[i + 1], tau >= p[i - 1])
            rl = range_mid
        dens.append(np.count_nonzero(logic) / rl)
    with np.errstate(divide="ignore"):
        recip = np.asarray([1.0 / x if x != 0 else np.inf for x in dens])
    if np.any(np.isinf(recip)):
        recip = np.where(np.isinf(recip), np.max(recip[np.isfinite(recip)]), recip)
    lo, hi = recip.min(), recip.max()
    rng = hi - lo
    scale = 1.0 / rng if rng != 0 else 1.0  # sklearn's _handle_zeros_in_scale
    scaled = (recip * scale + (0.0 - lo * scale)) * k  # MinMaxScaler: X*scale_ + min_
    sizes = scaled * kernel_window
    temp = list(np.abs(sizes - np.repeat(kernel_window, n_bins)))
    least = temp.index(max(temp))
    residue = np.abs(kernel_window - sizes[least])
    if k > 1:
        sizes = sizes + (residue / (k - 1))
    else:
        sizes = sizes + residue
    return np.asarray([s ** 2 for s in sizes])


def weight_matrix(cell_pseudotimes, n_bins, denoms=None, kernel_window=0.1):
    """TSP.py:44-51 + 122-133: ``W[t, c] = exp(-(|tau_c - p_t|)**2 / denom_t)``.

    ``denoms is None`` is the non-adaptive kernel: one scalar denominator
    ``kernel_window**2`` (== 0.010000000000000002 for 0.1).
    """
    p = interpolation_points(n_bins)
    tau = np.asarray(cell_pseudotimes, dtype=np.float64)
    absdiff = np.abs(tau[None, :] - p[:, None])
    if denoms is None:
        return np.exp(-(absdiff ** 2) / kernel_window ** 2)
    return np.exp(-np.divide(absdiff ** 2, np.asarray(denoms)[:, None]))


def interpolate_gene(x, W, user_given_std, eps):
    """TSP.py:161-196 (compute_stat / interpolate_gene_v2) + 198-213 (SummaryTimeSeries).

    ``x``: dense float64 expression of one gene over the (time-sorted) cells;
    ``W``: [n_bins, n_cells]; returns a dict with the SummaryTimeSeries fields.
    """
    x = np.asarray(x, dtype=np.float64)
    n_bins, n = W.shape
    means = np.empty(n_bins)
    stds = np.empty(n_bins)
    bins = []
    for t in range(n_bins):
        row = W[t]
        if user_given_std[t] < 0:
            sw = np.sum(row)
            mean = np.dot(row, x) / sw
            real_mean = np.mean(x)
            wss = np.dot(row, (x - real_mean) ** 2)
            std = np.sqrt(wss / (sw * (n - 1) / n))
            std = std * (np.sum(row) / n)  # cell_densities[idx], TSP.py:191,174
        else:
            mean = 0.0
            std = user_given_std[t]
        means[t] = mean
        stds[t] = std
        bins.append(mean + eps[t] * std)  # Normal(mu, sigma).rsample() == loc + eps*scale
    return {
        "intpl_means": means,
        "intpl_stds": stds,
        "data_bins": bins,
        "mean_trend": np.asarray([np.mean(b) for b in bins]),  # TSP.py:205
        "std_trend": np.asarray([np.std(b) for b in bins]),  # TSP.py:206
    }


def extract_significant_regions_only(regions):
    """Main.py:261-291, restated with plain lists (same control flow and quirks)."""
    if len(regions) == 0:
        return []
    start = regions[0][0]
    adjacent = []
    filtered = []
    for k in range(len(regions)):
        if k != len(regions) - 1:
            if regions[k][1] != regions[k + 1][0]:
                if regions[k][1] - start > 0.2:
                    adjacent = adjacent + [regions[k][0], regions[k][1]]
                    filtered = filtered + adjacent
                start = regions[k + 1][0]
                adjacent = []
            else:
                adjacent = adjacent + [regions[k][0]]
                continue
        else:
            if len(adjacent) > 0:
                if regions[k][1] - start > 0.2:
                    adjacent = adjacent + [regions[k][0], regions[k][1]]
                    filtered = filtered + adjacent
    return filtered


def check_inconsistent_zero_region(x, tau, p):
    """Main.py:293-304: half-open windows [p[i-1], p[i]), flagged when nnz <= 3."""
    regions = []
    for i in range(1, len(p)):
        sel = np.logical_and(tau >= p[i - 1], tau < p[i])
        if np.count_nonzero(x[sel]) <= 3:
            regions.append([p[i - 1], p[i]])
    return extract_significant_regions_only(regions)


def run_interpolation(x_ref, x_query, tau_ref, tau_query, W_ref, W_query, eps):
    """Main.py:307-410 (the live branch, ``no_extreme_cases`` False).

    All arrays are over time-sorted cells.  Returns (S, T, branch) where branch is
    0 both ~zero, 1 ref ~zero, 2 query ~zero, 3 both expressed.
    """
    n_bins = W_ref.shape[0]
    p = interpolation_points(n_bins)
    ugs = np.repeat(-1.0, n_bins)
    nr = np.count_nonzero(x_ref)
    nq = np.count_nonzero(x_query)
    if nr <= 3 and nq <= 3:
        ugs = np.repeat(0.01, n_bins)
        S = interpolate_gene(x_ref, W_ref, ugs, eps)
        T = interpolate_gene(x_query, W_query, ugs, eps)
        branch = 0
    elif nr <= 3:
        T_temp = interpolate_gene(x_query, W_query, ugs, eps)
        regions = check_inconsistent_zero_region(x_query, tau_query, p)
        common = min(T_temp["intpl_stds"]) / 10
        if len(regions) != 0:
            ugs[np.isin(p, regions)] = common
            T = interpolate_gene(x_query, W_query, ugs, eps)
        else:
            T = T_temp
        S = interpolate_gene(x_ref, W_ref, np.repeat(common, n_bins), eps)
        branch = 1
    elif nq <= 3:
        S_temp = interpolate_gene(x_ref, W_ref, ugs, eps)
        regions = check_inconsistent_zero_region(x_ref, tau_ref, p)
        common = min(S_temp["intpl_stds"]) / 10
        if len(regions) != 0:
            ugs[np.isin(p, regions)] = common
            S = interpolate_gene(x_ref, W_ref, ugs, eps)
        else:
            S = S_temp
        T = interpolate_gene(x_query, W_query, np.repeat(common, n_bins), eps)
        branch = 2
    else:
        S_temp = interpolate_gene(x_ref, W_ref, ugs, eps)
        T_temp = interpolate_gene(x_query, W_query, ugs, eps)
        reg_S = check_inconsistent_zero_region(x_ref, tau_ref, p)
        reg_T = check_inconsistent_zero_region(x_query, tau_query, p)
        S, T = S_temp, T_temp
        if len(reg_S) != 0 or len(reg_T) != 0:
            common = min(min(S_temp["intpl_stds"]), min(T_temp["intpl_stds"])) / 10
            if len(reg_S) != 0:
                u = np.repeat(-1.0, n_bins)
                u[np.isin(p, reg_S)] = common
                S = interpolate_gene(x_ref, W_ref, u, eps)
            if len(reg_T) != 0:
                u = np.repeat(-1.0, n_bins)
                u[np.isin(p, reg_T)] = common
                T = interpolate_gene(x_query, W_query, u, eps)
        branch = 3
    return S, T, branch


# --------------------------------------------------------------------------------------
# Stage II: MML match cost  (MyFunctions.py:10-50, OrgAlign.py:471-503)
# --------------------------------------------------------------------------------------
def _run_dist_compute_v3(data, mu, sigma):
    """MyFunctions.py:31-50 (with :10-15, :21-24, :26-29 inlined), float64."""
    data = np.asarray(data, dtype=np.float64)
    N = len(data)
    det_fisher = (2 * (N ** 2)) / (sigma ** 4)
    sum_term = np.sum(((data - mu) / sigma) ** 2.0) / 2.0
    nll = ((N / 2.0) * np.log(2 * np.pi)) + (N * np.log(sigma)) + sum_term
    L = nll - (N * np.log(0.001))
    i_model = (np.log(5 / (36 * np.sqrt(3))) + (np.log(sigma) + np.log(15.0 * 3.0))
               + (0.5 * np.log(det_fisher)))
    return i_model, L + 1.0


def compute_cell_faithful(ref_bin, query_bin, mu_S, sd_S, mu_T, sd_T):
    """OrgAlign.py:476-501: returns (null, match_len, match_compression)."""
    I_ref_model, I_refdata_g_ref = _run_dist_compute_v3(ref_bin, mu_S, sd_S)
    I_query_model, I_querydata_g_query = _run_dist_compute_v3(query_bin, mu_T, sd_T)
    I_ref_model, I_querydata_g_ref = _run_dist_compute_v3(query_bin, mu_S, sd_S)
    I_query_model, I_refdata_g_query = _run_dist_compute_v3(ref_bin, mu_T, sd_T)
    n2 = len(query_bin) + len(ref_bin)
    len1 = (I_ref_model + I_querydata_g_ref + I_refdata_g_ref) / n2
    len2 = (I_query_model + I_refdata_g_query + I_querydata_g_query) / n2
    match_len = (len1 + len2) / 2.0
    null = (I_ref_model + I_refdata_g_ref + I_query_model + I_querydata_g_query) / n2
    return null, match_len, match_len - null


def cost_matrix_faithful(S, T):
    """Cost for every (query bin i, ref bin j): ``out[i, j]`` (OrgAlign.py:340, 471-503)."""
    nT, nS = len(T["data_bins"]), len(S["data_bins"])
    out = np.empty((3, nT, nS))
    for i in range(nT):
        for j in range(nS):
            out[:, i, j] = compute_cell_faithful(
                S["data_bins"][j], T["data_bins"][i],
                S["mean_trend"][j], S["std_trend"][j], T["mean_trend"][i], T["std_trend"][i])
    return out  # [null, match_len, cost]


C_MODEL = math.log(5 / (36 * math.sqrt(3))) + math.log(45.0) + 0.5 * math.log(2.0 * N_SAMPLES ** 2)
K_DATA = N_SAMPLES / 2.0 * math.log(2 * math.pi) + 1.0 - N_SAMPLES * math.log(0.001)


def cost_matrix_closed_form(mu_S, sd_S, mu_T, sd_T, n=N_SAMPLES):
    """Closed form of OrgAlign.py:471-503 in the four trend vectors.

    Uses ``sum_n ((x_n - mu)/sigma)**2 = n * (s_x**2 + (m_x - mu)**2) / sigma**2`` where
    ``(m_x, s_x)`` -- the bin's own sample mean / population std -- are exactly
    ``mean_trend`` / ``std_trend``.  Works on [..., nT] x [..., nS] batches; returns
    ``cost[..., i, j]``.
    """
    mu_S = np.asarray(mu_S)[..., None, :]
    sd_S = np.asarray(sd_S)[..., None, :]
    mu_T = np.asarray(mu_T)[..., :, None]
    sd_T = np.asarray(sd_T)[..., :, None]
    d2 = (mu_T - mu_S) ** 2
    c_model = math.log(5 / (36 * math.sqrt(3))) + math.log(45.0) + 0.5 * math.log(2.0 * n ** 2)
    a = (sd_T ** 2 + d2) / sd_S ** 2 + (sd_S ** 2 + d2) / sd_T ** 2 - 2.0
    return (n / 2.0 * a - 2.0 * c_model + np.log(sd_S) + np.log(sd_T)) / (4.0 * n)


# --------------------------------------------------------------------------------------
# Stage III: five-state DP  (OrgAlign.py)
# --------------------------------------------------------------------------------------
def transition_table(free_params=(0.99, 0.1, 0.7), prohibit_case=False):
    """OrgAlign.py:20-97.  ``I[to, from]`` with states ordered M,W,V,D,I."""
    P_mm, P_ii, P_mi = free_params
    k = (1.0 - P_mm) / 4.0
    P_wm = P_vm = P_im = P_dm = k
    if prohibit_case:
        P_wi, P_vi = 0.0, 0.0
    else:
        P_wi, P_vi = 0.0, P_ii
    P_di = 1.0 - P_ii - P_mi - P_wi - P_vi
    P_md, P_dd, P_id, P_wd, P_vd = P_mi, P_ii, P_di, P_vi, 0.0
    with np.errstate(divide="ignore"):
        nl = lambda v: -np.log(v)
        I_mm, I_wm, I_vm, I_im, I_dm = nl(P_mm), nl(P_wm), nl(P_vm), nl(P_im), nl(P_dm)
        I_ii, I_mi, I_wi, I_vi, I_di = nl(P_ii), nl(P_mi), nl(P_wi), nl(P_vi), nl(P_di)
        I_dd, I_md, I_wd, I_vd, I_id = nl(P_dd), nl(P_md), nl(P_wd), nl(P_vd), nl(P_id)
    I_ww, I_mw, I_vw, I_dw, I_iw = I_mm, I_wm, I_vm, I_dm, I_im
    I_vv, I_mv, I_wv, I_dv, I_iv = I_mm, I_vm, I_wm, I_dm, I_im
    # rows = destination state, columns = source state, both in M,W,V,D,I order
    return np.array([
        [I_mm, I_mw, I_mv, I_md, I_mi],  # OrgAlign.py:352-356
        [I_wm, I_ww, I_wv, I_wd, I_wi],  # :362-366
        [I_vm, I_vw, I_vv, I_vd, I_vi],  # :369-373
        [I_dm, I_dw, I_dv, I_dd, I_di],  # :377-381
        [I_im, I_iw, I_iv, I_id, I_ii],  # :384-388
    ], dtype=np.float64)


def start_costs():
    """OrgAlign.py:274-276, 304, 324, 336, 345: (-ln ProbI_init, -ln ProbD_init, -ln ProbM_first)."""
    ProbM = 0.9999
    ProbI = (1.0 - ProbM) / 2.0
    return np.array([-np.log(ProbI), -np.log(ProbI), -np.log(0.99)], dtype=np.float64)


def _first_min(vals):
    m = min(vals)
    return m, vals.index(m)


def dp5_fill(cost, trans, start):
    """OrgAlign.py:229-331 (init) + 333-460 (fill), forward run only.

    ``cost[i, j]`` is the match cost of query bin i / ref bin j.  Returns
    ``mats`` [5, nT+1, nS+1] (M,W,V,D,I), ``bp`` [5, nT+1, nS+1] int8 (argmin index of
    the predecessor state; -1 where the reference stores inf triples or leaves [0,0,0]
    untouched at (0,0)) and ``opt_cost``.
    """
    nT, nS = cost.shape
    inf = np.inf
    mats = np.zeros((5, nT + 1, nS + 1))
    bp = np.full((5, nT + 1, nS + 1), -1, dtype=np.int8)
    M, W, V, D, I = 0, 1, 2, 3, 4
    for j in range(1, nS + 1):
        mats[M, 0, j] = mats[W, 0, j] = mats[V, 0, j] = inf
        mats[I, 0, j] = inf
    for i in range(1, nT + 1):
        mats[M, i, 0] = mats[W, i, 0] = mats[V, i, 0] = inf
        mats[D, i, 0] = inf
    for i in range(1, nT + 1):
        mats[I, i, 0] = mats[I, i - 1, 0] + 0.0 + (start[0] if i == 1 else trans[I, I])
        bp[I, i, 0] = 4
    for j in range(1, nS + 1):
        mats[D, 0, j] = mats[D, 0, j - 1] + 0.0 + (start[1] if j == 1 else trans[D, D])
        bp[D, 0, j] = 3
    for i in range(1, nT + 1):
        for j in range(1, nS + 1):
            c = cost[i - 1, j - 1]
            if i == 1 and j == 1:
                tm = [mats[M, 0, 0] + c + start[2], inf, inf, inf, inf]
            else:
                tm = [mats[k, i - 1, j - 1] + c + trans[M, k] for k in range(5)]
            tw = [mats[k, i, j - 1] + c + trans[W, k] for k in range(5)]
            tv = [mats[k, i - 1, j] + c + trans[V, k] for k in range(5)]
            td = [mats[k, i, j - 1] + 0.0 + trans[D, k] for k in range(5)]
            ti = [mats[k, i - 1, j] + 0.0 + trans[I, k] for k in range(5)]
            for s, cand in ((M, tm), (W, tw), (V, tv), (D, td), (I, ti)):
                v, a = _first_min(cand)
                mats[s, i, j] = v
                bp[s, i, j] = a
    last = [mats[k, nT, nS] for k in range(5)]
    return mats, bp, min(last)


_MOVES = {0: (-1, -1), 1: (0, -1), 2: (-1, 0), 3: (0, -1), 4: (-1, 0)}  # OrgAlign.py:446-450


def backtrack(mats, bp):
    """OrgAlign.py:539-607: returns (alignment_str, tracked_path list of [i, j])."""
    nT, nS = mats.shape[1] - 1, mats.shape[2] - 1
    i, j = nT, nS
    last = [mats[k, i, j] for k in range(5)]
    state = last.index(min(last))
    s = ""
    path = []
    while not (i == 0 and j == 0):
        prev_state = int(bp[state, i, j])
        di, dj = _MOVES[state]
        s = STATE_LETTERS[state] + s
        path.append([i, j])
        i, j = i + di, j + dj
        state = prev_state
    return s, path


def landscape(mats):
    """OrgAlign.py:732-747: (L_matrix, L_matrix_states as int8)."""
    return mats.min(axis=0), np.argmin(mats, axis=0).astype(np.int8)


# --------------------------------------------------------------------------------------
# Batched (vectorised over genes) restatements, for larger parity shapes
# --------------------------------------------------------------------------------------
def moments_batch(X, W):
    """Batched TSP.py:161-174.  ``X`` [n_cells, G] (sorted cells), ``W`` [T, n_cells].

    Returns (means [G, T], stds [G, T]) for ``user_given_std < 0`` everywhere.
    """
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    sw = W.sum(axis=1)  # [T]
    means = (W @ X).T / sw[None, :]
    xbar = X.mean(axis=0)
    wss = (W @ ((X - xbar[None, :]) ** 2)).T
    stds = np.sqrt(wss / (sw * (n - 1) / n)[None, :]) * (sw / n)[None, :]
    return means, stds


def zero_region_mask_batch(X, tau, p):
    """Batched Main.py:293-304 + 261-291 + ``np.in1d(p, regions)``: bool [G, T]."""
    X = np.asarray(X)
    G = X.shape[1]
    T = len(p)
    cnt = np.zeros((G, T - 1), dtype=np.int64)
    for i in range(1, T):
        sel = np.logical_and(tau >= p[i - 1], tau < p[i])
        cnt[:, i - 1] = np.count_nonzero(X[sel], axis=0)
    flagged = cnt <= 3
    mask = np.zeros((G, T), dtype=bool)
    for g in range(G):
        regions = [[p[i - 1], p[i]] for i in range(1, T) if flagged[g, i - 1]]
        sig = extract_significant_regions_only(regions)
        if len(sig):
            mask[g] = np.isin(p, sig)
    return mask, cnt


def policy_batch(Xr, Xq, tau_r, tau_q, Wr, Wq, eps):
    """Batched Main.py:307-410 + TSP.py:203-206 trends.

    Returns dict of [G, T] arrays: mu_S, sd_S, mu_T, sd_T (final intpl values),
    mean/std trends for both sides, and ``branch`` [G].
    """
    T = Wr.shape[0]
    p = interpolation_points(T)
    mr, sr = moments_batch(Xr, Wr)
    mq, sq = moments_batch(Xq, Wq)
    nr = np.count_nonzero(np.asarray(Xr), axis=0)
    nq = np.count_nonzero(np.asarray(Xq), axis=0)
    zr, _ = zero_region_mask_batch(Xr, tau_r, p)
    zq, _ = zero_region_mask_batch(Xq, tau_q, p)
    G = mr.shape[0]
    mu_S, sd_S, mu_T, sd_T = mr.copy(), sr.copy(), mq.copy(), sq.copy()
    branch = np.full(G, 3, dtype=np.int8)
    for g in range(G):
        if nr[g] <= 3 and nq[g] <= 3:
            mu_S[g] = 0.0; sd_S[g] = 0.01; mu_T[g] = 0.0; sd_T[g] = 0.01
            branch[g] = 0
        elif nr[g] <= 3:
            common = sq[g].min() / 10
            mu_T[g, zq[g]] = 0.0; sd_T[g, zq[g]] = common
            mu_S[g] = 0.0; sd_S[g] = common
            branch[g] = 1
        elif nq[g] <= 3:
            common = sr[g].min() / 10
            mu_S[g, zr[g]] = 0.0; sd_S[g, zr[g]] = common
            mu_T[g] = 0.0; sd_T[g] = common
            branch[g] = 2
        else:
            if zr[g].any() or zq[g].any():
                common = min(sr[g].min(), sq[g].min()) / 10
                mu_S[g, zr[g]] = 0.0; sd_S[g, zr[g]] = common
                mu_T[g, zq[g]] = 0.0; sd_T[g, zq[g]] = common
    em = eps.mean(axis=1)[None, :]
    es = eps.std(axis=1)[None, :]
    return {
        "mu_S": mu_S, "sd_S": sd_S, "mu_T": mu_T, "sd_T": sd_T,
        "mean_trend_S": mu_S + sd_S * em, "std_trend_S": sd_S * es,
        "mean_trend_T": mu_T + sd_T * em, "std_trend_T": sd_T * es,
        "branch": branch, "nnz_ref": nr, "nnz_query": nq,
    }


def dp5_fill_batch(cost, trans, start):
    """``dp5_fill`` vectorised over a leading gene axis: ``cost`` [G, nT, nS].

    Bit-identical to the scalar version: the same float64 adds in the same order,
    first-minimum argmin (``np.argmin`` returns the first occurrence).
    """
    G, nT, nS = cost.shape
    inf = np.inf
    mats = np.zeros((5, G, nT + 1, nS + 1))
    bp = np.full((5, G, nT + 1, nS + 1), -1, dtype=np.int8)
    M, W, V, D, I = 0, 1, 2, 3, 4
    mats[[M, W, V, I], :, 0, 1:] = inf
    mats[[M, W, V, D], :, 1:, 0] = inf
    for i in range(1, nT + 1):
        mats[I, :, i, 0] = mats[I, :, i - 1, 0] + 0.0 + (start[0] if i == 1 else trans[I, I])
        bp[I, :, i, 0] = 4
    for j in range(1, nS + 1):
        mats[D, :, 0, j] = mats[D, :, 0, j - 1] + 0.0 + (start[1] if j == 1 else trans[D, D])
        bp[D, :, 0, j] = 3
    tr = trans
    for i in range(1, nT + 1):
        for j in range(1, nS + 1):
            c = cost[:, i - 1, j - 1]
            diag = mats[:, :, i - 1, j - 1]
            left = mats[:, :, i, j - 1]
            up = mats[:, :, i - 1, j]
            if i == 1 and j == 1:
                tm = np.full((5, G), inf)
                tm[0] = diag[0] + c + start[2]
            else:
                tm = (diag + c[None, :]) + tr[M][:, None]
            tw = (left + c[None, :]) + tr[W][:, None]
            tv = (up + c[None, :]) + tr[V][:, None]
            td = (left + 0.0) + tr[D][:, None]
            ti = (up + 0.0) + tr[I][:, None]
            for s, cand in ((M, tm), (W, tw), (V, tv), (D, td), (I, ti)):
                a = np.argmin(cand, axis=0)
                mats[s, :, i, j] = np.take_along_axis(cand, a[None, :], axis=0)[0]
                bp[s, :, i, j] = a
    opt = mats[:, :, nT, nS].min(axis=0)
    return mats.transpose(1, 0, 2, 3), bp.transpose(1, 0, 2, 3), opt


def align_from_trends(mean_S, std_S, mean_T, std_T, trans=None, start=None, cost=None):
    """Stages II+III for one gene from its trend vectors (closed-form cost unless given)."""
    trans = transition_table() if trans is None else trans
    start = start_costs() if start is None else start
    if cost is None:
        cost = cost_matrix_closed_form(mean_S, std_S, mean_T, std_T)
    mats, bp, opt = dp5_fill(cost, trans, start)
    s, path = backtrack(mats, bp)
    L, Ls = landscape(mats)
    return {"alignment_str": s, "path": path, "opt_cost": opt, "mats": mats, "bp": bp,
            "L": L, "L_states": Ls, "cost": cost}


def align_batch(Xr, Xq, time_r, time_q, n_bins, adaptive=False, free_params=(0.99, 0.1, 0.7)):
    """Whole path on [cells, genes] matrices; the batched oracle used by parity tests.

    Returns dict with per-gene strings, opt costs, L_states, trends and cost matrices.
    """
    order_r, order_q = sort_cells(time_r), sort_cells(time_q)
    tau_r, tau_q = np.asarray(time_r, dtype=np.float64)[order_r], np.asarray(time_q, dtype=np.float64)[order_q]
    Xr = np.asarray(Xr)[order_r]
    Xq = np.asarray(Xq)[order_q]
    den_r = adaptive_window_denominators(tau_r, n_bins) if adaptive else None
    den_q = adaptive_window_denominators(tau_q, n_bins) if adaptive else None
    Wr = weight_matrix(tau_r, n_bins, den_r)
    Wq = weight_matrix(tau_q, n_bins, den_q)
    eps = epsilon_table(n_bins)
    pol = policy_batch(Xr, Xq, tau_r, tau_q, Wr, Wq, eps)
    cost = cost_matrix_closed_form(pol["mean_trend_S"], pol["std_trend_S"],
                                   pol["mean_trend_T"], pol["std_trend_T"])
    trans, start = transition_table(free_params), start_costs()
    mats, bp, opt = dp5_fill_batch(cost, trans, start)
    strings, paths = [], []
    for g in range(cost.shape[0]):
        s, path = backtrack(mats[g], bp[g])
        strings.append(s)
        paths.append(path)
    out = dict(pol)
    out.update({"cost": cost, "mats": mats, "bp": bp, "opt_cost": opt, "alignment_str": strings,
                "paths": paths, "L_states": np.argmin(mats, axis=1).astype(np.int8),
                "L": mats.min(axis=1), "eps": eps})
    return out


def canonical_gap_blocks(s):
    """Sort every maximal run of I/D (parity tier P3(i))."""
    out, run = [], []
    for ch in s:
        if ch in "ID":
            run.append(ch)
        else:
            out.extend(sorted(run))
            run = []
            out.append(ch)
    out.extend(sorted(run))
    return "".join(out)


def align_gene_faithful(x_ref, x_query, tau_ref, tau_query, W_ref, W_query, eps,
                        free_params=(0.99, 0.1, 0.7)):
    """One gene through the faithful path (what ``align_single_pair`` does, Main.py:413-432)."""
    S, T, branch = run_interpolation(x_ref, x_query, tau_ref, tau_query, W_ref, W_query, eps)
    util = cost_matrix_faithful(S, T)
    res = align_from_trends(None, None, None, None, transition_table(free_params), start_costs(),
                            cost=util[2])
    res.update({"S": S, "T": T, "branch": branch, "util": util})
    return res


# --------------------------------------------------------------------------------------
# f4: alignment-string distances (ClusterUtils.py:21-38, 361-377)
# --------------------------------------------------------------------------------------
def levenshtein(a, b):
    """Plain edit distance (what ``Levenshtein.distance`` of the python-Levenshtein package returns;
    the package is a third-party dependency of ClusterUtils.py:8,27, absent from the reference tree)."""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def levenshtein_matrix(strings):
    """ClusterUtils.py:21-29."""
    return np.array([[levenshtein(a, b) / max(len(a), len(b)) for b in strings] for a in strings])


def binary_encode_alignment_path(al_str):
    """ClusterUtils.py:361-377."""
    S = ""
    T = ""
    for ch in al_str:
        if ch == "M":
            S += "1"; T += "1"
        elif ch == "W":
            S += "1"
        elif ch == "V":
            T += "1"
        elif ch == "D":
            S += "0"
        elif ch == "I":
            T += "0"
    return np.asarray([int(c) for c in S + T])


def hamming_matrix(strings):
    """ClusterUtils.py:31-38 (scipy cdist 'hamming' = mean of mismatching positions)."""
    enc = np.array([binary_encode_alignment_path(s) for s in strings])
    return (enc[:, None, :] != enc[None, :, :]).mean(axis=2)
