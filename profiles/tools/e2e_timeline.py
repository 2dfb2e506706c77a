"""Wall-clock timeline of one steady-state end-to-end call of the public API (C2 workload): where the
time between 'host arrays in pinned memory' and 'every result string touched' goes.  Run twice: as
the API runs (asynchronous), and with a device synchronisation after the constructor to expose the
copy + packing time separately."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from genes2genes_b200 import Main
from genes2genes_b200.anndata_lite import AnnDataLite

from genes2genes_b200 import synthetic
_G, _nr, _nq, T = synthetic.CONFIGS["C2"]
Xr, tr, Xq, tq, genes = synthetic.make_pair(_G, _nr, _nq, dropout=0.9, seed=100)
ar = AnnDataLite(torch.from_numpy(Xr).pin_memory(), tr, genes)
aq = AnnDataLite(torch.from_numpy(Xq).pin_memory(), tq, genes)


def call(sync_after_ctor):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    al = Main.RefQueryAligner(ar, aq, genes, T, verbose=False)
    t1 = time.perf_counter()
    if sync_after_ctor:
        torch.cuda.synchronize()
    t2 = time.perf_counter()
    al.align_all_pairs()
    t3 = time.perf_counter()
    n = sum(a.alignment_str.count("M") for a in al.results)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    return [round(1e3 * v, 3) for v in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)]


for _ in range(3):
    call(False)
for mode in (False, True, False, True):
    print("sync_after_ctor=%s  ctor %.3f  wait %.3f  align_all_pairs %.3f  touch %.3f  total %.3f ms" % ((mode,) + tuple(call(mode))))
# pure copy time of the same bytes
torch.cuda.synchronize()
t0 = time.perf_counter()
a = ar.X.to("cuda", non_blocking=True); b = aq.X.to("cuda", non_blocking=True)
torch.cuda.synchronize()
print("H2D alone %.3f ms for %.1f MB" % (1e3 * (time.perf_counter() - t0), (ar.X.numel() + aq.X.numel()) * 4 / 1e6))
