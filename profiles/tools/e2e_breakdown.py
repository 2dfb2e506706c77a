"""Where the end-to-end time of the public API goes (host phases, with CUDA syncs between them)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from genes2genes_b200 import Main, engine
from genes2genes_b200.anndata_lite import AnnDataLite
from genes2genes_b200.TimeSeriesPreprocessor import TrajectoryInterpolator

Xr, tr, Xq, tq, genes, T = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "C2", 0.9)
Xr_pin, Xq_pin = torch.from_numpy(Xr).pin_memory(), torch.from_numpy(Xq).pin_memory()
dev = torch.device("cuda", 0)

def sync():
    torch.cuda.synchronize()
    return time.perf_counter()

for rep in range(3):
    t = [sync()]
    tir = TrajectoryInterpolator(tr, T, adaptive_kernel=False).run(); tiq = TrajectoryInterpolator(tq, T, adaptive_kernel=False).run(); t.append(sync())
    Xrd = Xr_pin.to(dev, non_blocking=True); Xqd = Xq_pin.to(dev, non_blocking=True); t.append(sync())
    Xrp = engine.pack_dense(Xrd, tir.order, None, dev); Xqp = engine.pack_dense(Xqd, tiq.order, None, dev); t.append(sync())
    eng = engine.AlignmentEngine(Xrp, tir.cell_pseudotimes, Xqp, tiq.cell_pseudotimes, len(genes), T, tir.kernel_denominators(), tiq.kernel_denominators(), device=dev); t.append(sync())
    eng.run(); t.append(sync())
    host = eng.results_to_host(); t.append(sync())
    names = ["interpolators(argsort)", "H2D", "pack(sort rows)", "engine ctor(allocs)", "kernels", "D2H"]
    print("rep", rep, " ".join("%s=%.2fms" % (n, (b - a) * 1e3) for n, a, b in zip(names, t, t[1:])), "total=%.2fms" % ((t[-1] - t[0]) * 1e3))
for rep in range(3):
    t0 = sync()
    al = Main.RefQueryAligner(AnnDataLite(Xr_pin, tr, genes), AnnDataLite(Xq_pin, tq, genes), genes, T, verbose=False); t1 = sync()
    al.align_all_pairs(); t2 = sync()
    n = sum(a.alignment_str.count("M") for a in al.results); t3 = sync()
    print("api rep", rep, "ctor=%.2fms align_all_pairs=%.2fms touch=%.2fms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    t0 = sync(); ad = AnnDataLite(Xr_pin, tr, genes); t1 = sync(); print("  AnnDataLite ctor %.2fms" % ((t1 - t0) * 1e3))
