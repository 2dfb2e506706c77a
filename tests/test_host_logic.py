"""CPU tests of the host side: alignment-string post-processing against the reference's golden
outputs, host tables against the oracle, the C-ABI surface, layout rules and the gloo gather."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from genes2genes_b200 import Main, OrgAlign, engine, _lib
from genes2genes_b200.TimeSeriesPreprocessor import SummaryTimeSeries, TrajectoryInterpolator
from oracle import g2g_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ["tutorial_ref_T14.npz", "synthetic_sparse_ref.npz", "synthetic_T13_ref.npz", "tutorial_ref_adaptive.npz"]


@pytest.mark.parametrize("name", FIXTURES)
def test_string_postprocessing_matches_reference(golden_dir, name):
    ref = np.load(os.path.join(golden_dir, name))
    for g, s in enumerate(str(x) for x in ref["alignment_str"]):
        vis, col, _ = Main.alignment_visual(s)
        assert vis == str(ref["al_visual"][g])
        assert col == str(ref["colored"][g])
        S, T = Main.matched_time_points(s)
        assert list(S) == json.loads(str(ref["match_points_S"][g]))
        assert list(T) == json.loads(str(ref["match_points_T"][g]))
        assert OrgAlign.legacy_matched_regions(s) == json.loads(str(ref["regions"][g]))
        assert OrgAlign.path_from_string(s) == json.loads(str(ref["paths"][g]))


def test_matched_points_equal_dp_cells_on_random_legal_strings():
    """For FSA-legal strings the matched bin pairs are the DP cells reached by M/W/V steps
    (SURVEY.md 8(a) row 14)."""
    rng = np.random.default_rng(0)
    for _ in range(3000):
        n = rng.integers(1, 30)
        s, prev = [], ""
        for k in range(n):
            opts = "MWVDI"
            if k == 0:
                opts = "MDI"
            elif prev == "I":
                opts = "MVDI"      # I -> W prohibited (P_wi = 0)
            elif prev == "D":
                opts = "MWDI"      # D -> V prohibited (P_vd = 0)
            prev = opts[rng.integers(len(opts))]
            s.append(prev)
        s = "".join(s)
        if s[0] in "WV":
            continue
        S, T = Main.matched_time_points(s)
        cells = OrgAlign.path_from_string(s)[::-1]
        want = [(c[1] - 1, c[0] - 1) for ch, c in zip(s, cells) if ch in "MWV"]
        assert list(zip(S.tolist(), T.tolist())) == want, s


def _store_from_reference(ref, T):
    """A result store filled with the reference's own outputs: exercises every AligmentObj view
    without a GPU."""
    G = len(ref["alignment_str"])
    strs = [str(s).encode() for s in ref["alignment_str"]]
    buf = np.zeros((G, 2 * T), dtype=np.uint8)
    for g, b in enumerate(strs):
        buf[g, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    host = {
        "align_str": buf, "align_len": np.array([len(b) for b in strs], dtype=np.int32),
        "opt_cost": ref["opt_cost"], "l_states": ref["L_states"].reshape(G, -1).astype(np.uint8),
        "musigma": np.stack([ref["intpl_means_S"], ref["intpl_stds_S"], ref["intpl_means_T"],
                             ref["intpl_stds_T"]], axis=1),
        "trends": np.stack([ref["mean_trend_S"], ref["std_trend_S"], ref["mean_trend_T"], ref["std_trend_T"]],
                           axis=1)}
    pts = np.linspace(0, 1, T)
    return Main._ResultStore(host, ["g%d" % k for k in range(G)], T, engine.epsilon_table(T), pts, pts,
                             (0.99, 0.1, 0.7), None)


def test_aligment_obj_fields_from_reference_strings(golden_dir):
    ref = np.load(os.path.join(golden_dir, "tutorial_ref_T14.npz"))
    store = _store_from_reference(ref, 14)
    results = [Main.AligmentObj(store, g, store.gene_list[g]) for g in range(89)]
    for g in range(0, 89, 7):
        a = results[g]
        assert a.alignment_str == str(ref["alignment_str"][g])
        assert tuple(a.get_series_match_percentage()) == tuple(ref["match_percentage"][g])
        assert [a.match_regions_S, a.match_regions_T, a.non_match_regions_S, a.non_match_regions_T] == \
            json.loads(str(ref["regions"][g]))
        assert (np.asarray(a.landscape_obj.L_matrix_states) == ref["L_states"][g].astype("<U1")).all()
        assert a.fwd_DP.alignment_path == json.loads(str(ref["paths"][g]))
        assert a.fwd_DP.opt_cost == ref["opt_cost"][g] == a.get_opt_alignment_cost()
        assert a.bwd_DP is None and a.fwd_DP.S_len == 14 and a.fwd_DP.T_len == 14
        a.cluster_id = 3                      # downstream code sets attributes on result objects
        assert a.cluster_id == 3
        with pytest.raises(AttributeError):
            a.no_such_attribute
        if g < ref["data_bins_S"].shape[0]:
            assert np.array_equal(np.asarray(a.S.data_bins), ref["data_bins_S"][g])
            assert np.array_equal(a.S.Y, ref["data_bins_S"][g].flatten())
        # DE statistics table (scipy on the host) vs the reference's
        want = np.asarray(json.loads(str(ref["de_info"][g])))
        got = a.matched_region_DE_info.values.astype(float)
        np.testing.assert_allclose(got[:, :4], want[:, :4], rtol=1e-12)
        np.testing.assert_allclose(got[:, 4], want[:, 4], atol=2e-7)     # Delta = round(cost, 7)
        np.testing.assert_allclose(got[:, 5:], want[:, 5:], rtol=1e-6, atol=1e-6)


def test_host_tables_equal_oracle():
    assert np.array_equal(engine.transition_costs(), o.transition_table())
    assert np.array_equal(engine.transition_costs((0.9, 0.2, 0.5), True), o.transition_table((0.9, 0.2, 0.5), True))
    assert np.array_equal(engine.start_costs(), o.start_costs())
    assert engine.model_constant() == o.C_MODEL
    assert np.array_equal(engine.epsilon_table(7), o.epsilon_table(7))
    fsa = OrgAlign.FiveStateMachine(0.99, 0.1, 0.7)
    assert fsa.I_di == 2.302585092994045 and fsa.I_ii == 2.3025850929940455 and np.isinf(fsa.I_wi)


def test_adaptive_denominators_equal_reference(golden_dir):
    ref = np.load(os.path.join(golden_dir, "tutorial_ref_adaptive.npz"))
    inp = np.load(os.path.join(golden_dir, "tutorial_inputs.npz"))
    for side, key in (("ref", "R"), ("query", "Q")):
        ti = TrajectoryInterpolator(inp["time_" + side], 14, adaptive_kernel=True).run()
        assert np.array_equal(np.asarray(ti.adaptive_win_denoms), ref["adaptive_win_denoms_" + key])
        np.testing.assert_allclose(ti.cell_weight_mat.sum(axis=1).values, ref["cell_weight_sum_" + key], rtol=1e-13)


def test_c_abi_exports_every_declared_symbol():
    """libg2g.so loads and exports exactly the entry points include/g2g.h declares."""
    header = open(os.path.join(ROOT, "include", "g2g.h")).read()
    declared = set(re.findall(r"\b(g2g_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.g2g_version() == 100
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (g2g_[a-z0-9_]+)", out))
    assert exported == declared


def test_layout_depends_on_cells_and_bins_only():
    lay = _lib.layout(20327, 14)
    assert lay.n_groups == 1 and lay.bins_per_group == 14 and lay.row_floats == 16 and lay.genes_per_tile == 64
    assert lay.n_pad == 20352 and lay.cells_per_item % 64 == 0
    assert lay.n_split == -(-lay.n_pad // lay.cells_per_item)
    lay = _lib.layout(100000, 50)
    assert lay.bins_per_group == 50 and lay.genes_per_tile == 32 and lay.row_floats == 52
    lay = _lib.layout(1000, 100)
    assert lay.n_groups == 2 and lay.bins_per_group == 50
    lay = _lib.layout(3, 2)
    assert lay.n_pad == 64 and lay.n_split == 1
    bad = _lib.Layout()
    assert _lib.load().g2g_layout(10, 1, ctypes.byref(bad)) < 0
    assert b"n_bins" in _lib.load().g2g_last_error()


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    assert lib.g2g_moments(None, 1, 10, 10, 14, None) < 0
    assert lib.g2g_cost_dp(None, 1, 300, None, None, 0.0, 50, None, None, None, None, None, None, None, None, None,
                           None) < 0


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from genes2genes_b200 import distributed as d
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
G, T = 10, 6
lo, hi = d.shard_range(G, rank, world)
per = -(-G // world)
class Src: pass
src = Src()
tmpl = d.record_templates(per, T)
rng = np.random.default_rng(rank)
for k, (shape, dtype) in tmpl.items():
    setattr(src, k, torch.from_numpy(rng.integers(0, 100, shape)).to(dtype))
g = d.ResultGather(src, dst=0)
g.run()
if rank == 0:
    res = g.host_results(per, T)
    assert len(res) == world
    for r in range(world):
        rr = np.random.default_rng(r)
        for k, (shape, dtype) in tmpl.items():
            want = torch.from_numpy(rr.integers(0, 100, shape)).to(dtype).numpy()
            assert np.array_equal(res[r][k], want), (r, k)
    print("GLOO_OK", d.record_bytes(per, T))
dist.destroy_process_group()
"""


def test_result_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert "GLOO_OK" in out.stdout, out.stdout + out.stderr


def test_shard_range_covers_all_genes():
    from genes2genes_b200 import distributed as d
    for G in (0, 1, 3, 7, 8, 9, 49, 1371, 20000):
        for world in (1, 2, 4, 8):
            spans = [d.shard_range(G, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            # balanced: no rank is empty while another holds two genes (G = 49 or 7 on 8 ranks, 9 on 4)
            assert max(sizes) - min(sizes) <= 1
            assert max(sizes) == d.max_shard(G, world)
            if G >= world:
                assert min(sizes) >= 1


def test_trajectory_interpolator_reference_signature_and_sorted_names():
    """``TrajectoryInterpolator(adata, n_bins, ...)`` like the reference (TSP.py:18); cells, their names
    and the weight-matrix columns are in pseudotime order (TSP.py:20, 51, 129)."""
    from genes2genes_b200.anndata_lite import AnnDataLite
    rng = np.random.default_rng(0)
    X = rng.random((6, 3)).astype(np.float32)
    t = np.array([0.9, 0.1, 0.5, 0.0, 1.0, 0.3])
    names = ["c%d" % i for i in range(6)]
    ad = AnnDataLite(X, t, ["a", "b", "c"], names)
    ti = TrajectoryInterpolator(ad, 4, adaptive_kernel=False).run()
    assert list(ti.obs_names) == ["c3", "c1", "c5", "c2", "c0", "c4"]
    assert list(ti.gene_list) == ["a", "b", "c"]
    W = ti.cell_weight_mat
    assert list(W.columns) == list(ti.obs_names)
    # the column of cell c0 (time 0.9) holds exp(-(0.9 - p)^2 / 0.1^2)
    np.testing.assert_allclose(W["c0"].values, np.exp(-(0.9 - np.linspace(0, 1, 4)) ** 2 / 0.1 ** 2), rtol=1e-12)
    assert ti.mat.shape == (3, 6)
    np.testing.assert_array_equal(np.asarray(ti.mat.todense()), X[np.argsort(t)].T)
    # from a bare pseudotime array (what the device path uses) there is no .mat
    with pytest.raises(AttributeError):
        TrajectoryInterpolator(t, 4, adaptive_kernel=False).mat


def test_mma_layout_and_table_size():
    lay = _lib.mma_layout(100000, 50)
    assert (lay.n_groups, lay.bins_per_group, lay.n_cols) == (1, 50, 64)
    assert lay.n_pad == 100032 and lay.cells_per_item == 4096 and lay.n_split == 25
    assert lay.table_bytes == (100032 // 64) * (64 * 256 + 256)
    lay = _lib.mma_layout(1500, 100)
    assert (lay.n_groups, lay.bins_per_group, lay.n_cols) == (2, 50, 64)
    lay = _lib.mma_layout(700, 30)
    assert lay.n_cols == 32 and lay.cells_per_item == 704 and lay.n_split == 1


def test_reference_runner_when_the_reference_is_present():
    """baseline/ref_runner.py runs the unmodified package (build container: /root/reference; GPU box:
    baseline/_ref) and agrees with the oracle's canonical strings."""
    from baseline import ref_runner
    if not ref_runner.reference_available():
        pytest.skip("reference package not present")
    from genes2genes_b200 import synthetic
    Xr, tr, Xq, tq, _ = synthetic.make_pair(3, 300, 250, dropout=0.8, seed=11)
    dt, strs = ref_runner.time_reference(Xr, tr, Xq, tq, 8)
    want = o.align_batch(Xr, Xq, tr, tq, 8)["alignment_str"]
    assert dt > 0 and len(strs) == 3
    assert [o.canonical_gap_blocks(a) for a in strs] == [o.canonical_gap_blocks(b) for b in want]


def test_average_alignment_walk_matches_reference(golden_dir):
    """Aggregate alignment walk fed the reference's own landscape states reproduces the reference's
    average alignment and path (ClusterUtils.py:455-518; Tutorial.ipynb prints 47.37 %)."""
    from genes2genes_b200 import ClusterUtils
    ref = np.load(os.path.join(golden_dir, "tutorial_ref_T14.npz"))
    L = ref["L_states"]
    counts = np.stack([(L == k).sum(axis=0) for k in range(5)])
    s, path = ClusterUtils.walk_average_alignment(counts)
    assert s == str(ref["aggregate_alignment"])
    assert path == ref["aggregate_path"].tolist()
    assert round(100 * sum(c in "MWV" for c in s) / len(s), 2) == 47.37
    # match-count matrix from matched points == the reference's
    mat = np.zeros((15, 15))
    for g in range(len(L)):
        S, T = Main.matched_time_points(str(ref["alignment_str"][g]))
        np.add.at(mat, (T + 1, S + 1), 1)
    assert np.array_equal(mat, ref["match_count_mat"])


def _region_mask_bit_scan(flagged, T, pts):
    """Python restatement of the kernel's formulation (csrc/g2g_fixup_body.cuh region_mask): every
    flagged window finds the bounds [a, b] of its run with bit scans, decides `keep`, and point t is
    listed iff window t or window t-1 lies in a kept run."""
    if flagged == 0:
        return 0
    last = flagged.bit_length() - 1
    full = (1 << 128) - 1
    kept = 0
    for t in range(T - 1):
        if (flagged >> t) & 1:
            zb = ~flagged & ((1 << t) - 1)
            a = zb.bit_length() if zb else 0
            za = ~flagged & full & ~((1 << (t + 1)) - 1)
            b = (za & -za).bit_length() - 2
            if (pts[b + 1] - pts[a]) > 0.2 and (b != last or b > a):
                kept |= 1 << t
    return (kept | (kept << 1)) & full


def test_region_mask_formulation_equals_the_reference_run_walk():
    """The lane-parallel region mask of the fix-up kernel against the oracle's restatement of
    extract_significant_regions_only (Main.py:261-291) + np.in1d(points, regions): exhaustive over all
    flag patterns for short series, random patterns up to 128 points."""
    import random
    rng = random.Random(7)
    for T in (2, 3, 5, 7, 13, 14, 29, 50, 64, 65, 100, 128):
        pts = np.linspace(0, 1, T)
        if T <= 14:
            patterns = range(1 << (T - 1))
        else:
            patterns = []
            for _ in range(600):
                dens = rng.choice([0.1, 0.5, 0.9, 0.98])
                patterns.append(sum(1 << k for k in range(T - 1) if rng.random() < dens))
        for f in patterns:
            regions = [[pts[k], pts[k + 1]] for k in range(T - 1) if (f >> k) & 1]
            listed = o.extract_significant_regions_only(regions)
            want = sum(1 << int(t) for t in np.nonzero(np.isin(pts, listed))[0])
            assert _region_mask_bit_scan(f, T, pts) == want, (T, bin(f))


def test_bench_reference_arm_prints_one_contract_line():
    """``bench.py --impl reference`` (the CPU arm, no GPU needed): exactly one line on stdout, JSON, with
    the keys of the measurement contract; library chatter must not reach stdout."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-genes-per-core", "1", "--workload", "C1"],
                         capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "gene_alignments_per_sec" and d["unit"] == "genes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and "workload" in d["config"]
    # the unmodified package when it is present (build container: /root/reference, GPU box: baseline/_ref)
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["sequential_value"] > 0 and d["scaling"] == "strong"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_host_results_assemble_rank_segments_lazily():
    """`engine.HostResults`: per-rank packed records laid out for the largest shard (capacity) become arrays
    covering every gene; the small arrays are copied at once, the large ones on first access, and the staging
    views are dropped when the last one has moved (the gather path of `RefQueryAligner(distributed=True)`)."""
    import torch
    from genes2genes_b200 import distributed as D
    n_all, world, T = 11, 4, 6
    cap = D.max_shard(n_all, world)
    rng = np.random.default_rng(0)
    want = {k: None for k in engine.AlignmentEngine.RESULT_KEYS}
    shapes = engine.result_shapes(n_all, T)
    np_dt = {torch.float64: np.float64, torch.int32: np.int32, torch.uint8: np.uint8}
    for k, (shape, dt) in shapes.items():
        want[k] = (rng.uniform(0, 200, size=shape)).astype(np_dt[dt])
    segments = []
    for r in range(world):
        lo, hi = D.shard_range(n_all, r, world)
        raw = np.zeros(D.record_bytes(cap, T), dtype=np.uint8)
        views = engine.unpack_results(raw, cap, T)
        for k in views:
            views[k][: hi - lo] = want[k][lo:hi]
        segments.append((raw, cap, hi - lo, lo))
    host = engine.HostResults(segments, n_all, T)
    assert set(host._arrays) == set(engine.HostResults.EAGER) and host._views is not None
    assert "trends" in host and "nope" not in host
    with pytest.raises(KeyError):
        host["nope"]
    for k in engine.AlignmentEngine.RESULT_KEYS:
        assert np.array_equal(host[k], want[k]), k
    assert host._views is None and host._segments is None      # staging released after the last array moved
    for raw, *_ in segments:                                    # the arrays own their memory
        raw[:] = 0
    assert np.array_equal(host["musigma"], want["musigma"])
