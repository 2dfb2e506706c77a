"""The ``ClusterUtils`` functions on or next to the alignment hot path:
the aggregate (cell-level) alignment over a gene set and the pairwise match-count matrix, whose per-gene
counting loops (ClusterUtils.py:435-450, 471-490) run as two small CUDA kernels
(``csrc/g2g_aggregate.cu``) over the batch's result arrays while the walk over the (T+1) x (S+1) grid
stays on the host; and the Levenshtein / Hamming distance matrices between alignment strings
(ClusterUtils.py:21-38, 361-377; ``csrc/g2g_strdist.cu``).  The clustering itself (agglomerative
clustering, plots) is out of scope.
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
        cache = store.__dict__.setdefault("_device_arrays", {})   # device copies of the result arrays, made once

        def on_device(key):
            if key not in cache:
                cache[key] = torch.from_numpy(np.ascontiguousarray(host[key])).to(dev)
            return cache[key]
        if which == "states":
            src = on_device("l_states")
            out = torch.empty((5, T + 1, T + 1), dtype=torch.int32, device=dev)
            _lib.check(lib.g2g_state_histogram(src.data_ptr(), G, T, mask_d.data_ptr(), out.data_ptr(), st),
                       "g2g_state_histogram")
        else:
            s = on_device("align_str")
            n = on_device("align_len")
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


# ---------------------------------------------------------------------------------------------------
# alignment-string distance matrices (reference: ClusterUtils.py:21-38, 361-377)
# ---------------------------------------------------------------------------------------------------
def _string_batch(set_of_strings, device):
    import torch
    strings = [s.encode("ascii") for s in set_of_strings]
    lens = np.array([len(b) for b in strings], dtype=np.int32)
    stride = max(1, int(lens.max()))
    if stride > 256:
        raise ValueError("alignment strings longer than 256 characters are not supported")
    buf = np.zeros((len(strings), stride), dtype=np.uint8)
    for k, b in enumerate(strings):
        buf[k, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return torch.from_numpy(buf).to(dev), torch.from_numpy(lens).to(dev), lens, stride, dev


def _integer_distances(set_of_strings, metric, device=None):
    """u16 [n, n] integer distances from the device kernels (``csrc/g2g_strdist.cu``)."""
    import torch
    lib = _lib.load()
    s_d, l_d, lens, stride, dev = _string_batch(set_of_strings, device)
    n = len(lens)
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        ws_bytes = lib.g2g_strdist_workspace_bytes(n)
        ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.int64, device=dev)
        out = torch.empty((n, n), dtype=torch.int16, device=dev)
        if metric == "levenshtein":
            _lib.check(lib.g2g_levenshtein_matrix(s_d.data_ptr(), l_d.data_ptr(), n, stride, out.data_ptr(),
                                                  ws.data_ptr(), ws_bytes, st), "g2g_levenshtein_matrix")
            norm = None
        else:
            first = set_of_strings[0]
            n_bins = sum(first.count(c) for c in "MWD")
            if any(sum(s.count(c) for c in "MWD") != n_bins or sum(s.count(c) for c in "MVI") != n_bins
                   for s in set_of_strings):
                raise ValueError("hamming distance needs alignments of series with equal numbers of time points")
            _lib.check(lib.g2g_hamming_matrix(s_d.data_ptr(), l_d.data_ptr(), n, stride, n_bins, out.data_ptr(),
                                              ws.data_ptr(), ws_bytes, st), "g2g_hamming_matrix")
            norm = 2 * n_bins
        return out.cpu().numpy().view(np.uint16), lens, norm


def compute_levenshtein_dist_matrix(set_of_strings, device=None):
    """E[i, j] = Levenshtein(s_i, s_j) / max(len(s_i), len(s_j)) (reference: ClusterUtils.py:21-29)."""
    d, lens, _ = _integer_distances(list(set_of_strings), "levenshtein", device)
    return d / np.maximum(lens[:, None], lens[None, :]).astype(np.float64)


def compute_hamming_dist_matrix(set_of_strings, device=None):
    """Hamming distance (fraction of differing bits) between the binary path encodings
    (reference: ClusterUtils.py:31-38, 361-377)."""
    d, _, norm = _integer_distances(list(set_of_strings), "hamming", device)
    return d / float(norm)


def binary_encode_alignment_path(al_str):
    """Reference-side bits (M, W -> 1; D -> 0) then query-side bits (M, V -> 1; I -> 0)
    (reference: ClusterUtils.py:361-377)."""
    S = [1 if c in "MW" else 0 for c in al_str if c in "MWD"]
    T = [1 if c in "MV" else 0 for c in al_str if c in "MVI"]
    return np.asarray(S + T)
