"""K1 (g2g_moments) alone at growing cell counts, T = 14, device-generated data: separates the fixed
cost of a launch (ramp, item quantisation) from the steady-state cost per cell, and compares an L2
flush that leaves dirty lines (memset) with one that leaves clean lines (memset, then a read pass)."""
import sys
import numpy as np
import torch
from genes2genes_b200 import engine

dev = torch.device("cuda", 0)
G, T = 1372, 14
flush_w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_r = torch.ones(64 << 20, dtype=torch.float32, device=dev)
den = np.full(T, 0.1 ** 2)
for scale in (1, 2, 4, 8):
    nr, nq = 20327 * scale, 17176 * scale
    gen = torch.Generator(device=dev).manual_seed(scale)
    sides = []
    for n in (nr, nq):
        X = torch.rand((n, G), device=dev, generator=gen)
        X = X * (torch.rand((n, G), device=dev, generator=gen) > 0.55)
        tau = np.sort(np.random.default_rng(n).uniform(0, 1, n)); tau[0], tau[-1] = 0.0, 1.0
        sides.append((X, tau))
    eng = engine.AlignmentEngine(sides[0][0], sides[0][1], sides[1][0], sides[1][1], G, T, den, den, device=dev)
    eng.run_interpolation()
    torch.cuda.synchronize()
    nbytes = (nr + nq) * G * 4
    for mode in ("dirty", "clean"):
        ts = []
        for rep in range(12):
            flush_w.zero_()
            if mode == "clean":
                flush_r.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.time_moments(a, b)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.mean(ts[2:]))
        print("scale %d  cells %7d  X %6.0f MB  flush=%s  K1 %.4f ms  %.0f GB/s  (%.3f us per 1000 cells)" %
              (scale, nr + nq, nbytes / 1e6, mode, t, nbytes / t / 1e6, t * 1e6 / (nr + nq)), flush=True)
    del eng, sides
    torch.cuda.empty_cache()
