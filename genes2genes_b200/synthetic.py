"""Deterministic synthetic single-cell trajectories.

Used by ``bench.py``, the tests and the golden-fixture script: there is no network for
real datasets, so every measured shape is generated from seeds.  Values are
``log1p``-normalised float32 like the reference's bundled ``.h5ad`` inputs.
"""
import numpy as np

__all__ = ["CONFIGS", "gene_params", "make_side", "make_pair", "make_side_device", "make_pair_device"]

# name -> (n_genes, n_ref_cells, n_query_cells, n_interpolation_points)   BASELINE.json "configs"
CONFIGS = {
    "C1": (89, 179, 290, 14),
    "C2": (1371, 20327, 17176, 14),
    "C3": (994, 3157, 890, 13),
    "C4": (20000, 100000, 100000, 50),
    "C5": (30000, 1000000, 1000000, 100),
    # one rank's gene block of C4 / C5 when sharded over 8 GPUs (C5 with the cell count cut to fit a bench host)
    "C4_8": (2500, 100000, 100000, 50),
    "C5_8s": (3750, 100000, 100000, 100),
}


def gene_params(n_genes, seed=0):
    """Per-gene bump parameters shared by the reference and the query side."""
    rng = np.random.default_rng(seed)
    return {
        "amp": rng.uniform(0.5, 3.0, n_genes),
        "ph": rng.uniform(0.0, 1.0, n_genes),
        "w": rng.uniform(0.1, 0.5, n_genes),
        # ~1% silent on ref only, ~1% on query only, ~1% on both (hits Main.py:344-374)
        "silent": rng.choice(4, size=n_genes, p=[0.97, 0.01, 0.01, 0.01]),
    }


def make_side(n_cells, params, seed, side, dropout=0.6, shift=0.0, row_chunk=8192, out=None):
    """One dataset: returns (X float32 [n_cells, n_genes], time float64 [n_cells]).

    ``side`` is 0 for the reference, 1 for the query (selects which genes are silent);
    ``shift`` moves every bump along pseudotime so ref and query differ.
    """
    rng = np.random.default_rng(seed)
    n_genes = len(params["amp"])
    tau = rng.uniform(0.0, 1.0, n_cells)
    tau = (tau - tau.min()) / (tau.max() - tau.min())  # exactly [0, 1]
    X = out if out is not None else np.empty((n_cells, n_genes), dtype=np.float32)
    silent = params["silent"]
    dead = (silent == 3) | (silent == (1 if side == 0 else 2))
    amp = params["amp"][None, :]
    ph = params["ph"][None, :] + shift
    w = params["w"][None, :]
    for r0 in range(0, n_cells, row_chunk):
        r1 = min(n_cells, r0 + row_chunk)
        t = tau[r0:r1, None]
        signal = amp * np.exp(-((t - ph) / w) ** 2)
        val = signal + rng.normal(0.0, 0.3, signal.shape)
        np.maximum(val, 0.0, out=val)
        drop = rng.uniform(0.0, 1.0, signal.shape) < dropout * np.exp(-signal)
        val[drop] = 0.0
        val[:, dead] = 0.0
        X[r0:r1] = np.log1p(val).astype(np.float32)
    return X, tau


def make_pair(n_genes, n_ref, n_query, dropout=0.6, seed=0):
    """(X_ref, time_ref, X_query, time_query, gene_names) with seeds 1 (ref) / 2 (query)."""
    params = gene_params(n_genes, seed)
    Xr, tr = make_side(n_ref, params, 1, 0, dropout=dropout)
    Xq, tq = make_side(n_query, params, 2, 1, dropout=dropout, shift=0.08)
    return Xr, tr, Xq, tq, ["g%d" % i for i in range(n_genes)]


_DEVICE_BLOCK = 256   # genes per random-number block of the device generator


def make_side_device(n_cells, params, seed, side, device, gene_lo=0, gene_hi=None, dropout=0.6, shift=0.0,
                     row_chunk=32768):
    """The same trajectory model as ``make_side`` generated on the GPU with torch (the whole-transcriptome
    shapes take minutes and tens of GB with numpy): returns (X float32 CUDA tensor [n_cells, ld] holding
    genes ``gene_lo:gene_hi`` of the dataset in the ORIGINAL cell order, time float64 numpy [n_cells]).

    Pseudotime comes from the same numpy stream as ``make_side``.  Noise and dropout of a gene are drawn
    per block of 256 absolute gene indices from a generator seeded by (seed, block, row chunk), so a gene's
    values do not depend on which gene range is generated: a rank's shard equals the matching columns
    of the single-GPU matrix bit for bit.  Values are NOT those of ``make_side`` (different generators)."""
    import torch
    n_genes = len(params["amp"])
    gene_hi = n_genes if gene_hi is None else gene_hi
    g_local = gene_hi - gene_lo
    ld = (g_local + 3) // 4 * 4
    rng = np.random.default_rng(seed)
    tau = rng.uniform(0.0, 1.0, n_cells)
    tau = (tau - tau.min()) / (tau.max() - tau.min())
    X = torch.zeros((n_cells, ld), dtype=torch.float32, device=device)
    silent = params["silent"]
    dead = (silent == 3) | (silent == (1 if side == 0 else 2))
    f32 = dict(dtype=torch.float32, device=device)
    tau_d = torch.as_tensor(tau, **f32)
    gen = torch.Generator(device=device)
    for b in range(gene_lo // _DEVICE_BLOCK, (gene_hi + _DEVICE_BLOCK - 1) // _DEVICE_BLOCK):
        b_lo, b_hi = b * _DEVICE_BLOCK, min(n_genes, (b + 1) * _DEVICE_BLOCK)
        lo, hi = max(b_lo, gene_lo), min(b_hi, gene_hi)
        amp = torch.as_tensor(params["amp"][b_lo:b_hi], **f32)[None, :]
        ph = torch.as_tensor(params["ph"][b_lo:b_hi] + shift, **f32)[None, :]
        w = torch.as_tensor(params["w"][b_lo:b_hi], **f32)[None, :]
        alive = torch.as_tensor(~dead[b_lo:b_hi], device=device)[None, :]
        for c, r0 in enumerate(range(0, n_cells, row_chunk)):
            r1 = min(n_cells, r0 + row_chunk)
            gen.manual_seed((int(seed) * 1000003 + b) * 4099 + c)
            t = tau_d[r0:r1, None]
            signal = amp * torch.exp(-((t - ph) / w) ** 2)
            noise = torch.randn((r1 - r0, b_hi - b_lo), generator=gen, **f32)
            u = torch.rand((r1 - r0, b_hi - b_lo), generator=gen, **f32)
            val = torch.clamp_min(signal + 0.3 * noise, 0.0)
            val = torch.where((u < dropout * torch.exp(-signal)) | ~alive, torch.zeros((), **f32), val)
            X[r0:r1, lo - gene_lo:hi - gene_lo] = torch.log1p(val)[:, lo - b_lo:hi - b_lo]
    return X, tau


def make_pair_device(n_genes, n_ref, n_query, device, gene_lo=0, gene_hi=None, dropout=0.6, seed=0):
    """Device twin of ``make_pair`` for the gene range [gene_lo, gene_hi):
    (X_ref, time_ref, X_query, time_query, names of the generated genes)."""
    params = gene_params(n_genes, seed)
    gene_hi = n_genes if gene_hi is None else gene_hi
    Xr, tr = make_side_device(n_ref, params, 1, 0, device, gene_lo, gene_hi, dropout=dropout)
    Xq, tq = make_side_device(n_query, params, 2, 1, device, gene_lo, gene_hi, dropout=dropout, shift=0.08)
    return Xr, tr, Xq, tq, ["g%d" % i for i in range(gene_lo, gene_hi)]
