"""Run the hot path and the CSR->CSC kernel with the debug build (make -C genes2genes_b200/csrc debug:
in-kernel bounds asserts, G2G_DEVICE_ASSERT).  A failed assert aborts the process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp, torch
from genes2genes_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "genes2genes_b200", "libg2g_debug.so")
from genes2genes_b200 import engine, synthetic
from genes2genes_b200.TimeSeriesPreprocessor import TrajectoryInterpolator

dev = torch.device("cuda", 0)
for G, n, T in ((2500, 20000, 50), (300, 5000, 14), (100, 6000, 100)):
    Xr, tr, Xq, tq, _ = synthetic.make_pair_device(G, n, n, dev, dropout=0.9, seed=100)
    tir, tiq = TrajectoryInterpolator(tr, T, adaptive_kernel=False).run(), TrajectoryInterpolator(tq, T, adaptive_kernel=False).run()
    eng = engine.AlignmentEngine(engine.pack_dense(Xr, tir.order, None, dev), tir.cell_pseudotimes,
                                 engine.pack_dense(Xq, tiq.order, None, dev), tiq.cell_pseudotimes, G, T,
                                 tir.kernel_denominators(), tiq.kernel_denominators(), device=dev)
    h = eng.run().results_to_host()
    print("debug build ok:", G, n, T, eng.variant, int(h["align_len"].sum()))
rng = np.random.default_rng(0)
dense = ((rng.uniform(size=(3000, 400)) < 0.07) * rng.uniform(0.5, 9, size=(3000, 400))).astype(np.float32)
c = engine.pack_csc(sp.csr_matrix(dense), rng.permutation(3000), rng.permutation(400)[:150], 400, dev)
torch.cuda.synchronize()
print("debug build ok: csr_to_csc", c.nnz)
