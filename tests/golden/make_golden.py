"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Writes, next to this script:
  tutorial_inputs.npz        the two bundled tutorial datasets (89 genes, 179/290 cells)
  tutorial_ref_T14.npz       reference outputs on them, n_bins=14 (all 89 genes)
  tutorial_ref_adaptive.npz  same, 5-argument adaptive-kernel mode (first 12 genes)
  synthetic_sparse_ref.npz   reference outputs on a 90%-dropout synthetic set that hits every
                             zero-expression branch of Main.py:344-406 (inputs are regenerated
                             from seeds by ``genes2genes_b200.synthetic``)
  synthetic_T13_ref.npz      a second n_bins (13) on a smaller synthetic set
  tutorial_published.json    outputs the AUTHORS stored in notebooks/Tutorial.ipynb
"""
import io
import json
import os
import re
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from baseline import ref_runner as rh  # noqa: E402
from _h5lite import read_h5ad_minimal  # noqa: E402
from genes2genes_b200 import synthetic  # noqa: E402


def run_reference(Xr, tr, Xq, tq, genes, n_bins, adaptive=None, n_data_bins=4, all_genes=None):
    g2g = rh.import_reference()
    all_genes = genes if all_genes is None else all_genes
    ar = rh.make_adata(Xr, tr, ["r%d" % i for i in range(len(tr))], all_genes)
    aq = rh.make_adata(Xq, tq, ["q%d" % i for i in range(len(tq))], all_genes)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all="ignore"):
        if adaptive is None:
            al = g2g.Main.RefQueryAligner(ar, aq, genes, n_bins)
        else:
            al = g2g.Main.RefQueryAligner(ar, aq, genes, n_bins, adaptive)
        al.align_all_pairs()
    G, T = len(genes), n_bins
    out = {
        "genes": np.asarray(genes),
        "n_bins": np.int64(n_bins),
        "alignment_str": np.asarray([a.alignment_str for a in al.results]),
        "opt_cost": np.asarray([float(a.fwd_DP.opt_cost) for a in al.results]),
        "match_percentage": np.asarray([a.get_series_match_percentage() for a in al.results]),
        "intpl_means_S": np.asarray([a.S.intpl_means for a in al.results], dtype=np.float64),
        "intpl_stds_S": np.asarray([a.S.intpl_stds for a in al.results], dtype=np.float64),
        "intpl_means_T": np.asarray([a.T.intpl_means for a in al.results], dtype=np.float64),
        "intpl_stds_T": np.asarray([a.T.intpl_stds for a in al.results], dtype=np.float64),
        "mean_trend_S": np.asarray([a.S.mean_trend for a in al.results]),
        "std_trend_S": np.asarray([a.S.std_trend for a in al.results]),
        "mean_trend_T": np.asarray([a.T.mean_trend for a in al.results]),
        "std_trend_T": np.asarray([a.T.std_trend for a in al.results]),
        "cell_weight_sum_R": np.asarray(al.TrajInt_R.cell_weight_mat.sum(axis=1)),
        "cell_weight_sum_Q": np.asarray(al.TrajInt_Q.cell_weight_mat.sum(axis=1)),
    }
    if adaptive:
        out["adaptive_win_denoms_R"] = np.asarray(al.TrajInt_R.adaptive_win_denoms)
        out["adaptive_win_denoms_Q"] = np.asarray(al.TrajInt_Q.adaptive_win_denoms)
    util = np.zeros((G, 3, T, T))
    mats = np.zeros((G, 5, T + 1, T + 1))
    bp = np.full((G, 5, T + 1, T + 1), -1, dtype=np.int8)
    Ls = np.zeros((G, T + 1, T + 1), dtype=np.int8)
    L = np.zeros((G, T + 1, T + 1))
    paths, mpS, mpT, al_visual, colored, regions, de_rows = [], [], [], [], [], [], []
    for g, a in enumerate(al.results):
        dp = a.fwd_DP
        for i in range(1, T + 1):
            for j in range(1, T + 1):
                util[g, :, i - 1, j - 1] = [float(v) for v in dp.DP_util_matrix[i, j]]
        for k, (m, b) in enumerate(((dp.DP_M_matrix, dp.backtrackers_M), (dp.DP_W_matrix, dp.backtrackers_W),
                                    (dp.DP_V_matrix, dp.backtrackers_V), (dp.DP_D_matrix, dp.backtrackers_D),
                                    (dp.DP_I_matrix, dp.backtrackers_I))):
            mats[g, k] = np.asarray(m)
            for i in range(T + 1):
                for j in range(T + 1):
                    v = b[i][j][2]
                    if np.isfinite(v) and not (i == 0 and j == 0):
                        # untouched [0,0,0] entries exist only at (0,0) and on borders handled below
                        bp[g, k, i, j] = int(v)
        # borders: the reference leaves [0,0,0] in M/W/V/I/D tables where it never writes; normalise
        # to -1 everywhere the DP does not define a predecessor (row 0 / col 0 except I col 0, D row 0)
        for k in range(5):
            bp[g, k, 0, :] = -1
            bp[g, k, :, 0] = -1
        bp[g, 4, 1:, 0] = 4
        bp[g, 3, 0, 1:] = 3
        Ls[g] = np.asarray(a.landscape_obj.L_matrix_states).astype(np.int8)
        L[g] = np.asarray(a.landscape_obj.L_matrix)
        paths.append(json.dumps([[int(x) for x in pr] for pr in dp.alignment_path]))
        mpS.append(json.dumps([int(x) for x in a.match_points_S]))
        mpT.append(json.dumps([int(x) for x in a.match_points_T]))
        al_visual.append(a.al_visual)
        colored.append(a.colored_alignment_str)
        regions.append(json.dumps([a.match_regions_S, a.match_regions_T, a.non_match_regions_S,
                                   a.non_match_regions_T]))
        de_rows.append(json.dumps(np.asarray(a.matched_region_DE_info, dtype=np.float64).tolist()))
    out.update({"util": util, "mats": mats, "bp": bp, "L_states": Ls, "L": L,
                "paths": np.asarray(paths), "match_points_S": np.asarray(mpS),
                "match_points_T": np.asarray(mpT), "al_visual": np.asarray(al_visual),
                "colored": np.asarray(colored), "regions": np.asarray(regions),
                "de_info": np.asarray(de_rows)})
    # aggregate (cell-level) alignment and pairwise match counts over all genes
    # (ClusterUtils.get_cluster_average_alignments / get_pairwise_match_count_mat, :435-518)
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        avg, avg_path = g2g.ClusterUtils.get_cluster_average_alignments(al, list(genes))
        mat = g2g.ClusterUtils.get_pairwise_match_count_mat(al, list(genes))
    out["aggregate_alignment"] = np.asarray(avg)
    out["aggregate_path"] = np.asarray(avg_path, dtype=np.int64)
    out["match_count_mat"] = np.asarray(mat, dtype=np.float64)
    # data_bins (bit-exactness of the epsilon table) for the first few genes, both sides
    nb = min(n_data_bins, G)
    out["data_bins_S"] = np.asarray([[np.asarray(b, dtype=np.float64) for b in al.results[g].S.data_bins]
                                     for g in range(nb)])
    out["data_bins_T"] = np.asarray([[np.asarray(b, dtype=np.float64) for b in al.results[g].T.data_bins]
                                     for g in range(nb)])
    return out


def published_from_notebook():
    nb = json.load(open(os.path.join(rh.REFERENCE_ROOT, "notebooks", "Tutorial.ipynb")))
    ansi = re.compile(r"\x1b\[[0-9;]*m")
    pub = {"strings": {}, "source": "notebooks/Tutorial.ipynb stored outputs"}
    for c in nb["cells"]:
        if c["cell_type"] != "code":
            continue
        src = "".join(c["source"])
        for o in c.get("outputs", []):
            txt = "".join(o.get("text", [])) if "text" in o else "".join(o.get("data", {}).get("text/plain", []))
            if "show_ordered_alignments" in src and "Alignment" in txt:
                for line in txt.splitlines():
                    parts = ansi.sub("", line).split()
                    if len(parts) == 2 and re.fullmatch(r"[MWVID]+", parts[1]):
                        pub["strings"][parts[0]] = parts[1]
            if "results_map['TNF']" in src and "al_visual" in src and "5-state string" in txt:
                pub["TNF_print"] = txt
            if "Optimal alignment cost" in txt:
                m = re.search(r"Optimal alignment cost: ([0-9.]+) nits", txt)
                s = re.search(r"Alignment similarity percentage: ([0-9.]+) %", txt)
                pub["TNF_cost_nits"] = float(m.group(1))
                pub["TNF_similarity_pct"] = float(s.group(1))
            if "Mean alignment similarity percentage" in txt:
                m = re.search(r"\n([0-9.]+) %", txt)
                pub["mean_similarity_pct"] = float(m.group(1))
            if "get_stat_df" in src and "opt_alignment_cost" in txt:
                rows = {}
                for line in txt.splitlines():
                    m = re.match(r"\s*\d+\s+(\S+)\s+([0-9.]+)\s+([0-9.]+)\s+(-?[0-9.]+)", line)
                    if m:
                        rows[m.group(1)] = [float(m.group(2)), float(m.group(3)), float(m.group(4))]
                pub["stat_df_rows"] = rows
    return pub


def main():
    d = os.path.join(rh.REFERENCE_ROOT, "notebooks", "data")
    Xr, tr, _, genes = read_h5ad_minimal(os.path.join(d, "adata_pam_local.h5ad"))
    Xq, tq, _, genes_q = read_h5ad_minimal(os.path.join(d, "adata_lps_local.h5ad"))
    assert genes == genes_q
    np.savez_compressed(os.path.join(HERE, "tutorial_inputs.npz"), X_ref=Xr, time_ref=tr, X_query=Xq,
                        time_query=tq, genes=np.asarray(genes))
    print("tutorial T=14 ...", flush=True)
    np.savez_compressed(os.path.join(HERE, "tutorial_ref_T14.npz"), **run_reference(Xr, tr, Xq, tq, genes, 14))
    print("tutorial adaptive ...", flush=True)
    np.savez_compressed(os.path.join(HERE, "tutorial_ref_adaptive.npz"),
                        **run_reference(Xr, tr, Xq, tq, genes[:12], 14, adaptive=True, all_genes=genes))
    print("synthetic sparse ...", flush=True)
    Xr, tr, Xq, tq, genes = synthetic.make_pair(96, 400, 330, dropout=0.92, seed=7)
    o = run_reference(Xr, tr, Xq, tq, genes, 14)
    o["gen"] = np.asarray([96, 400, 330, 92, 7])
    np.savez_compressed(os.path.join(HERE, "synthetic_sparse_ref.npz"), **o)
    print("synthetic T=13 ...", flush=True)
    Xr, tr, Xq, tq, genes = synthetic.make_pair(40, 230, 190, dropout=0.6, seed=3)
    o = run_reference(Xr, tr, Xq, tq, genes, 13)
    o["gen"] = np.asarray([40, 230, 190, 60, 3])
    np.savez_compressed(os.path.join(HERE, "synthetic_T13_ref.npz"), **o)
    json.dump(published_from_notebook(), open(os.path.join(HERE, "tutorial_published.json"), "w"), indent=1)
    print("done")


if __name__ == "__main__":
    main()
