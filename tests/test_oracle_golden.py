"""Pin the CPU oracle (oracle/g2g_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, made by tests/golden/make_golden.py) and against the outputs the
authors stored in notebooks/Tutorial.ipynb (tests/golden/tutorial_published.json).

Parity tiers follow SURVEY.md section 8: P1 bitwise DP given the same cost matrix,
P2 numeric closeness of moments / trends / costs, P3 strings equal after canonicalising
[ID]+ gap blocks.
"""
import json
import os

import numpy as np
import pytest

from oracle import g2g_oracle as o
from genes2genes_b200 import synthetic

FIXTURES = ["tutorial_ref_T14.npz", "synthetic_sparse_ref.npz", "synthetic_T13_ref.npz"]


def _inputs(golden_dir, name):
    ref = np.load(os.path.join(golden_dir, name))
    if name.startswith("tutorial"):
        inp = np.load(os.path.join(golden_dir, "tutorial_inputs.npz"))
        Xr, tr, Xq, tq = inp["X_ref"], inp["time_ref"], inp["X_query"], inp["time_query"]
        genes = [str(g) for g in inp["genes"]]
        sel = [genes.index(str(g)) for g in ref["genes"]]
        return ref, Xr[:, sel], tr, Xq[:, sel], tq
    G, nr, nq, drop, seed = [int(v) for v in ref["gen"]]
    Xr, tr, Xq, tq, _ = synthetic.make_pair(G, nr, nq, dropout=drop / 100.0, seed=seed)
    return ref, Xr, tr, Xq, tq


def test_epsilon_table_reconstructs_reference_samples_bit_exactly(golden_dir):
    ref = np.load(os.path.join(golden_dir, "tutorial_ref_T14.npz"))
    eps = o.epsilon_table(14)
    for side in "ST":
        mu, sd, bins = ref["intpl_means_" + side], ref["intpl_stds_" + side], ref["data_bins_" + side]
        for g in range(bins.shape[0]):
            assert np.array_equal(mu[g][:, None] + eps * sd[g][:, None], bins[g])


@pytest.mark.parametrize("name", FIXTURES + ["tutorial_ref_adaptive.npz"])
def test_p1_dp_bitwise_given_reference_cost(golden_dir, name):
    ref = np.load(os.path.join(golden_dir, name))
    trans, start = o.transition_table(), o.start_costs()
    cost = ref["util"][:, 2]
    mats, bp, opt = o.dp5_fill_batch(cost, trans, start)
    assert np.array_equal(mats, ref["mats"])
    assert np.array_equal(bp, ref["bp"])
    assert np.array_equal(opt, ref["opt_cost"])
    assert np.array_equal(mats.min(axis=1), ref["L"])
    assert np.array_equal(np.argmin(mats, axis=1), ref["L_states"])
    for g in range(cost.shape[0]):
        s, path = o.backtrack(mats[g], bp[g])
        assert s == str(ref["alignment_str"][g])
        assert path == json.loads(str(ref["paths"][g]))
    # scalar restatement agrees with the batched one (first 5 genes)
    for g in range(min(5, cost.shape[0])):
        m1, b1, o1 = o.dp5_fill(cost[g], trans, start)
        assert np.array_equal(m1, mats[g]) and np.array_equal(b1, bp[g]) and o1 == opt[g]


@pytest.mark.parametrize("name", FIXTURES)
def test_p2_p3_batched_oracle_vs_reference(golden_dir, name):
    ref, Xr, tr, Xq, tq = _inputs(golden_dir, name)
    T = int(ref["n_bins"])
    b = o.align_batch(Xr, Xq, tr, tq, T)
    for k_or, k_ref in (("mu_S", "intpl_means_S"), ("sd_S", "intpl_stds_S"), ("mu_T", "intpl_means_T"),
                        ("sd_T", "intpl_stds_T"), ("mean_trend_S", "mean_trend_S"), ("std_trend_S", "std_trend_S"),
                        ("mean_trend_T", "mean_trend_T"), ("std_trend_T", "std_trend_T")):
        np.testing.assert_allclose(b[k_or], ref[k_ref], rtol=1e-10, atol=1e-13, err_msg=k_or)
    np.testing.assert_allclose(b["cost"], ref["util"][:, 2], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(b["opt_cost"], ref["opt_cost"], rtol=1e-10)
    G = len(ref["alignment_str"])
    canon = sum(o.canonical_gap_blocks(b["alignment_str"][g]) == o.canonical_gap_blocks(str(ref["alignment_str"][g]))
                for g in range(G))
    assert canon == G
    if name == "synthetic_sparse_ref.npz":  # every zero-expression branch of Main.py:344-406 is exercised
        assert set(np.unique(b["branch"])) == {0, 1, 2, 3}


def test_faithful_oracle_vs_reference(golden_dir):
    ref, Xr, tr, Xq, tq = _inputs(golden_dir, "synthetic_sparse_ref.npz")
    T = int(ref["n_bins"])
    eps = o.epsilon_table(T)
    orr, oq = o.sort_cells(tr), o.sort_cells(tq)
    taur, tauq = tr[orr], tq[oq]
    Wr, Wq = o.weight_matrix(taur, T), o.weight_matrix(tauq, T)
    for g in range(0, Xr.shape[1], 6):
        r = o.align_gene_faithful(Xr[orr, g].astype(np.float64), Xq[oq, g].astype(np.float64),
                                  taur, tauq, Wr, Wq, eps)
        np.testing.assert_allclose(r["S"]["intpl_means"], ref["intpl_means_S"][g], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(r["T"]["intpl_stds"], ref["intpl_stds_T"][g], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(r["util"], ref["util"][g], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r["opt_cost"], ref["opt_cost"][g], rtol=1e-11)
        assert o.canonical_gap_blocks(r["alignment_str"]) == o.canonical_gap_blocks(str(ref["alignment_str"][g]))


def test_adaptive_window_denominators(golden_dir):
    ref = np.load(os.path.join(golden_dir, "tutorial_ref_adaptive.npz"))
    inp = np.load(os.path.join(golden_dir, "tutorial_inputs.npz"))
    for side, key in (("ref", "R"), ("query", "Q")):
        tau = np.sort(inp["time_" + side])
        den = o.adaptive_window_denominators(tau, 14)
        np.testing.assert_allclose(den, ref["adaptive_win_denoms_" + key], rtol=1e-12)
        W = o.weight_matrix(tau, 14, den)
        np.testing.assert_allclose(W.sum(axis=1), ref["cell_weight_sum_" + key], rtol=1e-12)
    sel = [list(map(str, inp["genes"])).index(str(g)) for g in ref["genes"]]
    b = o.align_batch(inp["X_ref"][:, sel], inp["X_query"][:, sel], inp["time_ref"], inp["time_query"], 14,
                      adaptive=True)
    np.testing.assert_allclose(b["opt_cost"], ref["opt_cost"], rtol=1e-10)


def test_published_tutorial_outputs(golden_dir):
    """The authors' stored notebook outputs (Tutorial.ipynb:403-411, 485-486, 583, 722-725, 813-904)."""
    pub = json.load(open(os.path.join(golden_dir, "tutorial_published.json")))
    inp = np.load(os.path.join(golden_dir, "tutorial_inputs.npz"))
    genes = [str(g) for g in inp["genes"]]
    b = o.align_batch(inp["X_ref"], inp["X_query"], inp["time_ref"], inp["time_query"], 14)
    assert len(pub["strings"]) == 89
    for g, name in enumerate(genes):
        assert o.canonical_gap_blocks(pub["strings"][name]) == o.canonical_gap_blocks(b["alignment_str"][g])
    tnf = genes.index("TNF")
    assert round(float(b["opt_cost"][tnf]), 2) == pub["TNF_cost_nits"]
    s = b["alignment_str"][tnf]
    assert round(100.0 * sum(c in "MWV" for c in s) / len(s), 2) == pub["TNF_similarity_pct"]
    sims = [sum(c in "MWV" for c in s) / len(s) for s in b["alignment_str"]]
    assert abs(round(float(np.mean(sims)), 4) * 100 - pub["mean_similarity_pct"]) < 0.011
    for name, (sim, cost, _l2fc) in pub["stat_df_rows"].items():
        assert abs(float(b["opt_cost"][genes.index(name)]) - cost) < 5e-6


def test_significant_region_quirks():
    p = o.interpolation_points(14)
    w = lambda i: [p[i], p[i + 1]]
    # trailing singleton run is never kept, however long (Main.py:283)
    assert o.extract_significant_regions_only([w(0)]) == []
    # a run of 3 adjacent windows (length 3/13 > 0.2) ending at the last flagged window is kept
    assert o.extract_significant_regions_only([w(2), w(3), w(4)]) == [p[2], p[3], p[4], p[5]]
    # two windows (2/13 = 0.154 < 0.2) are not
    assert o.extract_significant_regions_only([w(2), w(3)]) == []
    # non-trailing run followed by a separate trailing singleton
    assert o.extract_significant_regions_only([w(0), w(1), w(2), w(8)]) == [p[0], p[1], p[2], p[3]]


def test_string_distances_against_scipy_and_known_values(golden_dir):
    """f4 oracle: Hamming equals scipy's cdist on the reference's binary encoding; Levenshtein known values."""
    from scipy.spatial import distance
    ref = np.load(os.path.join(golden_dir, "tutorial_ref_T14.npz"))
    strings = [str(s) for s in ref["alignment_str"]][:30]
    enc = np.array([o.binary_encode_alignment_path(s) for s in strings])
    assert enc.shape[1] == 28
    assert np.allclose(o.hamming_matrix(strings), distance.cdist(enc, enc, "hamming"))
    assert o.levenshtein("kitten", "sitting") == 3 and o.levenshtein("", "abc") == 3
    assert o.levenshtein("MMMDDI", "MMMIDD") == 2
