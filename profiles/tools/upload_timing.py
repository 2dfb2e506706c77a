"""Wall time of the host->device + time-sort pipeline (engine.upload_host_block) alone, C2 matrices."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from genes2genes_b200 import engine, synthetic
G, nr, nq, T = synthetic.CONFIGS["C2"]
Xr, tr, Xq, tq, genes = synthetic.make_pair(G, nr, nq, dropout=0.9, seed=100)
Xr, Xq = torch.from_numpy(Xr).pin_memory(), torch.from_numpy(Xq).pin_memory()
dev = torch.device("cuda", 0)
orr, oq = np.argsort(tr), np.argsort(tq)
for chunk in (96 << 20, 32 << 20, 16 << 20, 8 << 20):
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a = engine.upload_host_block(Xr, orr, 0, G, None, dev, chunk_bytes=chunk)
        t1 = time.perf_counter()
        b = engine.upload_host_block(Xq, oq, 0, G, None, dev, chunk_bytes=chunk)
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
    print("chunk %3d MB: enqueue ref %.2f ms, query %.2f ms, drain %.2f ms, total %.2f ms" %
          (chunk >> 20, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0)))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = engine.upload_host_block(Xr, None, 0, G, None, dev); b = engine.upload_host_block(Xq, None, 0, G, None, dev)
    torch.cuda.synchronize(); t3 = time.perf_counter()
print("identity order (direct copy): %.2f ms" % (1e3 * (t3 - t0)))
