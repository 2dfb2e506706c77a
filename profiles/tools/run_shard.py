"""Run the hot path on one synthetic gene shard generated on the device (profiling driver).

    python profiles/tools/run_shard.py [--genes 2500] [--cells 100000] [--bins 50] [--moments auto] [--reps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from genes2genes_b200 import engine, synthetic  # noqa: E402
from genes2genes_b200.TimeSeriesPreprocessor import TrajectoryInterpolator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=2500)
ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--bins", type=int, default=50)
ap.add_argument("--moments", default="auto")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--lib", default=None, help="another build of libg2g.so (profiling variants)")
args = ap.parse_args()
if args.lib:
    from genes2genes_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(args.lib)
dev = torch.device("cuda", 0)
G, n, T = args.genes, args.cells, args.bins
Xr, tr, Xq, tq, _ = synthetic.make_pair_device(G, n, n, dev, dropout=0.9, seed=100)
tir = TrajectoryInterpolator(tr, T, adaptive_kernel=False).run()
tiq = TrajectoryInterpolator(tq, T, adaptive_kernel=False).run()
Xrd = engine.pack_dense(Xr, tir.order, None, dev)
Xqd = engine.pack_dense(Xq, tiq.order, None, dev)
del Xr, Xq
eng = engine.AlignmentEngine(Xrd, tir.cell_pseudotimes, Xqd, tiq.cell_pseudotimes, G, T,
                             tir.kernel_denominators(), tiq.kernel_denominators(), device=dev, moments=args.moments)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
for _ in range(args.reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.run(); b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
k1 = []
for _ in range(args.reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.time_moments(a, b)
    torch.cuda.synchronize()
    k1.append(a.elapsed_time(b))
fu = []
for _ in range(args.reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.time_fused(a, b)
    torch.cuda.synchronize()
    fu.append(a.elapsed_time(b))
import hashlib  # noqa: E402
eng.run()
torch.cuda.synchronize()
print("records sha", hashlib.sha256(eng.packed.cpu().numpy().tobytes()).hexdigest()[:16])
print("variant %s step ms %s k1 ms %s -> %.0f GB/s; fused ms %s" % (eng.variant, np.round(ms, 3), np.round(k1, 3),
                                                                    4.0 * 2 * n * G / (min(k1) * 1e-3) / 1e9, np.round(fu, 3)))
