"""Tensor-core interpolation kernel (K1m, g2g_moments_mma) against the FFMA2 kernel (K1) and the f64 oracle:
errors of the interpolated means / stds and kernel times.  Run on the GPU box:

    python profiles/tools/k1m_check.py [--big]

Prints one JSON line per case (kept under profiles/ as the measured basis of the tensor-core decision)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from genes2genes_b200 import engine, synthetic  # noqa: E402
from genes2genes_b200.TimeSeriesPreprocessor import TrajectoryInterpolator  # noqa: E402
from oracle import g2g_oracle as o  # noqa: E402


def make_engine(Xr, tr, Xq, tq, T, moments, dev):
    tir = TrajectoryInterpolator(tr, T, adaptive_kernel=False).run()
    tiq = TrajectoryInterpolator(tq, T, adaptive_kernel=False).run()
    Xrd = engine.pack_dense(Xr, tir.order, None, dev)
    Xqd = engine.pack_dense(Xq, tiq.order, None, dev)
    return engine.AlignmentEngine(Xrd, tir.cell_pseudotimes, Xqd, tiq.cell_pseudotimes, Xr.shape[1], T,
                                  tir.kernel_denominators(), tiq.kernel_denominators(), device=dev, moments=moments)


def rel_err(got, want):
    scale = np.maximum(np.abs(want), np.abs(want).max(axis=1, keepdims=True) * 1e-2) + 1e-12
    return float((np.abs(got - want) / scale).max())


def time_k1(eng, flush, reps=5):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.time_moments(a, b)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def case(G, nr, nq, T, drop, seed, dev, flush, n_oracle=48):
    Xr, tr, Xq, tq, _ = synthetic.make_pair(G, nr, nq, dropout=drop, seed=seed)
    sel = np.unique(np.linspace(0, G - 1, min(G, n_oracle)).astype(int))
    ref = o.align_batch(np.ascontiguousarray(Xr[:, sel]), np.ascontiguousarray(Xq[:, sel]), tr, tq, T)
    out = {"G": G, "n_ref": nr, "n_query": nq, "T": T, "dropout": drop, "oracle_genes": int(len(sel))}
    hosts = {}
    for mode in ("ffma", "mma"):
        eng = make_engine(Xr, tr, Xq, tq, T, mode, dev)
        eng.run()
        h = eng.results_to_host()
        hosts[mode] = h
        out[mode] = {"variant_after_run": eng.variant, "k1_ms": time_k1(eng, flush)}
        for k, key in enumerate(("mu_S", "sd_S", "mu_T", "sd_T")):
            out[mode]["err_" + key] = rel_err(h["musigma"][sel, k], ref[key])
        strs = [h["align_str"][g, :h["align_len"][g]].tobytes().decode() for g in sel]
        out[mode]["canonical_equal"] = int(sum(o.canonical_gap_blocks(a) == o.canonical_gap_blocks(b)
                                               for a, b in zip(strs, ref["alignment_str"])))
        out[mode]["raw_equal"] = int(sum(a == b for a, b in zip(strs, ref["alignment_str"])))
        out[mode]["nnz_equal"] = bool(np.array_equal(h["nnz"][sel, 0], ref["nnz_ref"]) and
                                      np.array_equal(h["nnz"][sel, 1], ref["nnz_query"]))
        out[mode]["opt_cost_relerr"] = float(np.max(np.abs(h["opt_cost"][sel] - ref["opt_cost"]) /
                                                    np.abs(ref["opt_cost"])))
        del eng
    a, b = hosts["ffma"], hosts["mma"]
    out["mma_vs_ffma"] = {"nnz_equal": bool(np.array_equal(a["nnz"], b["nnz"])),
                          "flags_equal": bool(np.array_equal(a["flags"], b["flags"])),
                          "strings_equal": int(np.sum((a["align_len"] == b["align_len"]) &
                                                      np.all(a["align_str"] == b["align_str"], axis=1))),
                          "musigma_max_rel": float(np.max(np.abs(a["musigma"] - b["musigma"]) /
                                                          (np.abs(a["musigma"]) + 1e-3)))}
    print(json.dumps(out), flush=True)


def big_case(G, nr, nq, T, drop, dev, flush, n_oracle=16):
    """One rank's C4 block, generated on the device; a few genes are checked against the oracle."""
    Xr, tr, Xq, tq, _ = synthetic.make_pair_device(G, nr, nq, dev, dropout=drop, seed=100)
    sel = np.unique(np.linspace(0, G - 1, n_oracle).astype(int))
    Xr_h = Xr[:, torch.as_tensor(sel, device=dev)].cpu().numpy()
    Xq_h = Xq[:, torch.as_tensor(sel, device=dev)].cpu().numpy()
    ref = o.align_batch(Xr_h, Xq_h, tr, tq, T)
    out = {"G": G, "n_ref": nr, "n_query": nq, "T": T, "dropout": drop, "oracle_genes": int(len(sel)),
           "data": "device generator"}
    tir = TrajectoryInterpolator(tr, T, adaptive_kernel=False).run()
    tiq = TrajectoryInterpolator(tq, T, adaptive_kernel=False).run()
    Xrd = engine.pack_dense(Xr, tir.order, None, dev)
    Xqd = engine.pack_dense(Xq, tiq.order, None, dev)
    del Xr, Xq
    hosts = {}
    for mode in ("ffma", "mma"):
        eng = engine.AlignmentEngine(Xrd, tir.cell_pseudotimes, Xqd, tiq.cell_pseudotimes, G, T,
                                     tir.kernel_denominators(), tiq.kernel_denominators(), device=dev, moments=mode)
        eng.run()
        h = eng.results_to_host()
        hosts[mode] = h
        steps = []
        for _ in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.run(); b.record()
            torch.cuda.synchronize()
            steps.append(a.elapsed_time(b))
        out[mode] = {"variant_after_run": eng.variant, "k1_ms": time_k1(eng, flush), "step_ms": float(np.median(steps))}
        for k, key in enumerate(("mu_S", "sd_S", "mu_T", "sd_T")):
            out[mode]["err_" + key] = rel_err(h["musigma"][sel, k], ref[key])
        strs = [h["align_str"][g, :h["align_len"][g]].tobytes().decode() for g in sel]
        out[mode]["canonical_equal"] = int(sum(o.canonical_gap_blocks(a) == o.canonical_gap_blocks(b)
                                               for a, b in zip(strs, ref["alignment_str"])))
        out[mode]["nnz_equal"] = bool(np.array_equal(h["nnz"][sel, 0], ref["nnz_ref"]) and
                                      np.array_equal(h["nnz"][sel, 1], ref["nnz_query"]))
        del eng
    a, b = hosts["ffma"], hosts["mma"]
    ca = [o.canonical_gap_blocks(a["align_str"][g, :a["align_len"][g]].tobytes().decode()) for g in range(G)]
    cb = [o.canonical_gap_blocks(b["align_str"][g, :b["align_len"][g]].tobytes().decode()) for g in range(G)]
    out["mma_vs_ffma"] = {"nnz_equal": bool(np.array_equal(a["nnz"], b["nnz"])),
                          "flags_equal": bool(np.array_equal(a["flags"], b["flags"])),
                          "strings_equal": int(np.sum((a["align_len"] == b["align_len"]) &
                                                      np.all(a["align_str"] == b["align_str"], axis=1))),
                          "canonical_equal": int(sum(x == y for x, y in zip(ca, cb)))}
    alg = 4.0 * (nr + nq) * G
    for mode in ("ffma", "mma"):
        out[mode]["k1_GBps"] = alg / (out[mode]["k1_ms"] * 1e-3) / 1e9
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t0 = time.time()
    case(130, 700, 600, 30, 0.6, 1, dev, flush)
    case(300, 5000, 4000, 50, 0.9, 2, dev, flush)
    case(200, 9000, 7000, 56, 0.8, 3, dev, flush)
    case(70, 1500, 1100, 100, 0.6, 4, dev, flush)
    case(48, 900, 700, 128, 0.7, 5, dev, flush)
    if args.big:
        big_case(2500, 100000, 100000, 50, 0.9, dev, flush)
    print("done in %.1f s" % (time.time() - t0), file=sys.stderr)


if __name__ == "__main__":
    main()
