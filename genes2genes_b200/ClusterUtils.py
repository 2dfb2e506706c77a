"""The two ``ClusterUtils`` functions every notebook calls right after alignment (SURVEY.md 8(f) row
f2): the aggregate (cell-level) alignment over a gene set and the pairwise match-count matrix.  The
per-gene counting loops of the reference (ClusterUtils.py:435-450, 471-490) run as two small CUDA
kernels (``csrc/g2g_aggregate.cu``) over the batch's result arrays; the walk over the (T+1) x (S+1)
grid stays on the host.  The rest of the reference's ``ClusterUtils`` (clustering, Levenshtein
distances) is out of scope.
"""
import ctypes

import numpy as np

from . import _lib

_STATE = "MWVDI"


def _gene_mask(aligner, gene_set):
    idx = aligner._store.gene_index
    mask = np.zeros(len(aligner.gene_list), dtype=np.uint8)
    for g in gene_set:
        mask[idx[g]] = 1
    return mask


def _device_counts(aligner, gene_set, which):
    """Run one of the two counting kernels on the aligner's device (result arrays are uploaded from
    the host copy: 1 byte per DP cell and gene)."""
    import torch
    store = aligner._store
    T = store.n_bins
    lib = _lib.load()
    host = store.host
    dev = store.device
    mask = _gene_mask(aligner, gene_set)
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        mask_d = torch.from_numpy(mask).to(dev)
        G = len(mask)
        if which == "states":
            src = torch.from_numpy(np.ascontiguousarray(host["l_states"])).to(dev)
            out = torch.empty((5, T + 1, T + 1), dtype=torch.int32, device=dev)
            _lib.check(lib.g2g_state_histogram(src.data_ptr(), G, T, mask_d.data_ptr(), out.data_ptr(), st),
                       "g2g_state_histogram")
        else:
            s = torch.from_numpy(np.ascontiguousarray(host["align_str"])).to(dev)
            n = torch.from_numpy(np.ascontiguousarray(host["align_len"])).to(dev)
            out = torch.empty((T + 1, T + 1), dtype=torch.int32, device=dev)
            _lib.check(lib.g2g_match_counts(s.data_ptr(), n.data_ptr(), G, T, mask_d.data_ptr(), out.data_ptr(), st),
                       "g2g_match_counts")
        return out.cpu().numpy()


def walk_average_alignment(counts, deterministic=True, rng=None):
    """Walk from (T_len, S_len) to (0, 0): at each cell take the most frequent landscape state over
    the gene set (ties -> lowest state index, like ``np.argmax``) and step M (-1,-1), W/D (0,-1),
    V/I (-1,0) (reference: ClusterUtils.py:463-518).  ``counts`` is int [5, T+1, S+1]."""
    i, j = counts.shape[1] - 1, counts.shape[2] - 1
    n = counts[:, 0, 0].sum()
    s, path = "", [[i, j]]
    guard = 4 * (i + j) + 8
    while not (i == 0 and j == 0) and guard > 0:
        guard -= 1
        freq = counts[:, i, j] / max(n, 1)
        if deterministic:
            cs = int(np.argmax(freq))
        else:
            rng = np.random.default_rng() if rng is None else rng
            cs = int(np.searchsorted(np.cumsum(freq), rng.random(), side="left"))
            cs = min(cs, 4)
        if cs == 0:
            i, j = i - 1, j - 1
        elif cs in (1, 3):
            j -= 1
        else:
            i -= 1
        s = _STATE[cs] + s
        path.append([i, j])
    return s, path


def get_cluster_average_alignments(aligner, gene_set, deterministic=True):
    """(average alignment string, tracked path) over ``gene_set`` (reference: ClusterUtils.py:455-518)."""
    counts = _device_counts(aligner, gene_set, "states")
    return walk_average_alignment(counts, deterministic)


def get_pairwise_match_count_mat(aligner, gene_set):
    """DataFrame [(T+1) x (S+1)]: how many genes of ``gene_set`` match query bin i-1 with reference bin
    j-1 (reference: ClusterUtils.py:435-450)."""
    import pandas as pd
    return pd.DataFrame(_device_counts(aligner, gene_set, "matches").astype(np.float64))
