"""cProfile of one steady-state end-to-end call of the public API (C2 workload)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from genes2genes_b200 import Main
from genes2genes_b200.anndata_lite import AnnDataLite

from genes2genes_b200 import synthetic
_G, _nr, _nq, T = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
Xr, tr, Xq, tq, genes = synthetic.make_pair(_G, _nr, _nq, dropout=0.9, seed=100)
ar = AnnDataLite(torch.from_numpy(Xr).pin_memory(), tr, genes)
aq = AnnDataLite(torch.from_numpy(Xq).pin_memory(), tq, genes)

def call():
    al = Main.RefQueryAligner(ar, aq, genes, T, verbose=False)
    al.align_all_pairs()
    n = sum(a.alignment_str.count("M") for a in al.results)
    torch.cuda.synchronize()
    return n

for _ in range(3):
    call()
pr = cProfile.Profile()
pr.enable()
call()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
