"""Batched device pipeline behind ``Main.RefQueryAligner``: all genes of one (reference, query)
pair go through five kernel launches of ``libg2g.so``.

PyTorch is used for device memory, streams and (multi-GPU) ``torch.distributed`` only; every
kernel is hand-written CUDA behind the C ABI in ``include/g2g.h``.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import Side, SparseSide, c_f64

N_SAMPLES = 50  # reference: generate_random_dataset(50, ...)  TimeSeriesPreprocessor.py:179


_EPS_CACHE = {}


def epsilon_table(n_bins, n_samples=N_SAMPLES, seed=1):
    """The fixed standard-normal table behind every interpolated bin of the reference.

    The reference re-seeds torch with 1 at the top of each ``interpolate_gene_v2`` call
    (TimeSeriesPreprocessor.py:184) and then draws ``n_samples`` scalar normals per bin
    (MyFunctions.py:64-67), so ``data_bins[t][n] = mean_t + eps[t, n] * std_t`` with one table
    shared by all genes.  A private generator seeded the same way gives the same mt19937 stream
    without touching the global RNG state.
    """
    key = (int(n_bins), int(n_samples), int(seed))
    if key not in _EPS_CACHE:
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        zero = torch.zeros((), dtype=torch.float64)
        one = torch.ones((), dtype=torch.float64)
        out = np.empty((n_bins, n_samples), dtype=np.float64)
        for t in range(n_bins):
            for n in range(n_samples):
                out[t, n] = float(torch.normal(zero, one, generator=gen))
        out.setflags(write=False)
        _EPS_CACHE[key] = out
    return _EPS_CACHE[key]


def transition_costs(free_params=(0.99, 0.1, 0.7), prohibit_case=False):
    """-ln of the five-state machine's transition probabilities, ``I[to, from]`` with states in
    M, W, V, D, I order (reference: FiveStateMachine.__init__, OrgAlign.py:20-97).  Computed on
    the host with numpy so the table is bit-identical to the reference's."""
    p_mm, p_ii, p_mi = (float(v) for v in free_params)
    k = (1.0 - p_mm) / 4.0
    p_vi = 0.0 if prohibit_case else p_ii
    p_wi = 0.0
    p_di = 1.0 - p_ii - p_mi - p_wi - p_vi
    prob = np.array([
        # from:  M     W     V     D      I
        [p_mm, k, k, p_mi, p_mi],      # to M   (I_mm, I_mw, I_mv, I_md, I_mi)
        [k, p_mm, k, p_vi, p_wi],      # to W   (.., I_wd = -ln P_vi, I_wi = -ln 0)
        [k, k, p_mm, 0.0, p_vi],       # to V   (.., I_vd = -ln 0,   I_vi)
        [k, k, k, p_ii, p_di],         # to D   (.., I_dd, I_di)
        [k, k, k, p_di, p_ii],         # to I   (.., I_id, I_ii)
    ], dtype=np.float64)
    with np.errstate(divide="ignore"):
        return -np.log(prob)


def start_costs():
    """(-ln ProbI, -ln ProbD) of the DP borders (OrgAlign.py:274-276, 304, 324) and -ln 0.99 of
    the first match cell (OrgAlign.py:336, 345)."""
    prob_m = 0.9999
    prob_i = (1.0 - prob_m) / 2.0
    return np.array([-np.log(prob_i), -np.log(prob_i), -np.log(0.99)], dtype=np.float64)


def model_constant(n_samples=N_SAMPLES):
    """ln(5/(36 sqrt 3)) + ln(15*3) + 0.5 ln(2 N^2): the sigma-independent part of the MML model
    cost (MyFunctions.py:21-29, 41, 46)."""
    return math.log(5 / (36 * math.sqrt(3))) + math.log(45.0) + 0.5 * math.log(2.0 * n_samples ** 2)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def upload_small(arrays, device):
    """Several small host arrays -> device tensors with ONE asynchronous copy (a pageable source would
    make every ``.to(device)`` a blocking call queued behind the big matrix upload).  ``arrays`` maps
    name -> numpy array; 8-byte aligned slots in one pinned staging buffer."""
    specs, total = [], 0
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        specs.append((name, a, total))
        total += (a.nbytes + 7) // 8 * 8
    stage = _pinned_buffer(max(total, 8))   # waits (normally not at all) for the buffer's previous copy
    host = stage.numpy()
    for name, a, off in specs:
        host[off:off + a.nbytes] = a.view(np.uint8).reshape(-1)
    dev_buf = torch.empty(max(total, 8), dtype=torch.uint8, device=device)
    dev_buf.copy_(stage[:max(total, 8)], non_blocking=True)
    _mark_in_flight(stage)   # the next user of this staging buffer waits for the copy to leave the host
    out = {}
    for name, a, off in specs:
        out[name] = dev_buf[off:off + a.nbytes].view(_TORCH_DTYPES[a.dtype.type]).reshape(a.shape)
    return out


class DeviceSide:
    """Device-resident state of one dataset: packed expression matrix, weight table, workspaces.

    ``variant`` selects the interpolation kernel the buffers are laid out for: ``"ffma"`` (K1
    ``g2g_moments``, f32 weight table), ``"mma"`` (K1m ``g2g_moments_mma``, f16 hi/lo B-operand table) or ``"csc"`` (K1s, sparse input)."""

    def __init__(self, X, tau_dev, n_genes, n_bins, denom_dev, device, variant="ffma"):
        # X: dense f32 [n_cells, ld] (time-sorted rows, ld % 4 == 0) or a CscMatrix (gene-major sparse)
        self.csc = X if isinstance(X, CscMatrix) else None
        self.X = None if self.csc is not None else X
        self.n_cells = int(X.shape[0])
        self.ld = 0 if self.csc is not None else int(X.shape[1])
        self.n_genes, self.n_bins, self.device = n_genes, n_bins, device
        self.tau = tau_dev
        self.denom = denom_dev
        self.sum_w = torch.empty(n_bins, dtype=torch.float64, device=device)
        ws = _lib.load().g2g_weights_workspace_bytes(self.n_cells, n_bins)
        self.ws = torch.empty((ws + 7) // 8, dtype=torch.float64, device=device)
        self.ws_bytes = ws
        self.pilot = torch.empty(n_genes, dtype=torch.float32, device=device)
        self.g_stride = n_genes
        self.variant = None
        self.configure(variant)

    def configure(self, variant):
        """(Re)allocate the weight table and the partial-sum planes for ``variant``."""
        if variant == self.variant:
            return
        n_bins, n_genes, device = self.n_bins, self.n_genes, self.device
        self.wtab = self.part_f64 = self.part_i32 = None   # release before allocating the new set
        if variant == "mma":
            lay = _lib.mma_layout(self.n_cells, n_bins)
            self.wtab = torch.empty(int(lay.table_bytes), dtype=torch.uint8, device=device)
        else:
            lay = _lib.layout(self.n_cells, n_bins)
            self.wtab = torch.empty(lay.n_groups * lay.n_pad * lay.row_floats, dtype=torch.float32, device=device)
        self.layout = lay
        self.n_split = int(lay.n_split)
        self.part_f64 = torch.empty(self.n_split * (2 * n_bins + 1) * n_genes, dtype=torch.float64, device=device)
        self.part_i32 = torch.empty(self.n_split * n_bins * n_genes, dtype=torch.int32, device=device)
        self.variant = variant

    def c_side(self):
        mma = self.variant == "mma"
        return Side(self.X.data_ptr() if self.X is not None else None, self.n_cells, self.ld, self.wtab.data_ptr(),
                    self.pilot.data_ptr(), self.part_f64.data_ptr(), self.part_i32.data_ptr(),
                    self.n_split if mma else 0)

    def c_sparse_side(self):
        c = self.csc
        return SparseSide(c.col_ptr.data_ptr(), c.row_idx.data_ptr(), c.values.data_ptr(), self.n_cells,
                          self.wtab.data_ptr(), self.pilot.data_ptr(), self.part_f64.data_ptr(),
                          self.part_i32.data_ptr())


class CscMatrix:
    """Gene-major sparse expression matrix on the device: ``col_ptr`` i64 [G+1], ``row_idx`` i32 [nnz]
    (cell index in time-sorted order, ascending inside a column), ``values`` f32 [nnz]."""

    def __init__(self, col_ptr, row_idx, values, n_cells, n_genes):
        self.col_ptr, self.row_idx, self.values = col_ptr, row_idx, values
        self.shape = (int(n_cells), int(n_genes))

    @property
    def nnz(self):
        return int(self.values.numel())

    @property
    def device(self):
        return self.values.device


def pack_dense(X, row_order, gene_idx, device):
    """Device copy of ``X[row_order][:, gene_idx]`` as f32 with 16-byte aligned rows
    (``g2g_pack_rows``).  ``X`` may be a host array / tensor (copied first) or a CUDA tensor."""
    lib = _lib.load()
    if not torch.is_tensor(X):
        X = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
    if X.dtype != torch.float32:
        X = X.to(torch.float32)
    Xd = X.to(device, non_blocking=True).contiguous()
    n_cells, ld_src = int(Xd.shape[0]), int(Xd.shape[1])
    n_genes = ld_src if gene_idx is None else len(gene_idx)
    ld = (n_genes + 3) // 4 * 4
    identity_rows = row_order is None
    if identity_rows and gene_idx is None and ld == ld_src:
        return Xd
    out = torch.empty((n_cells, ld), dtype=torch.float32, device=device)
    small = {}
    if not identity_rows:
        small["ro"] = np.asarray(row_order, dtype=np.int64)
    if gene_idx is not None:
        small["gi"] = np.asarray(gene_idx, dtype=np.int32)
    small = upload_small(small, device) if small else {}
    ro, gi = small.get("ro"), small.get("gi")
    _lib.check(lib.g2g_pack_rows(_ptr(Xd), n_cells, n_genes, ld_src, _ptr(ro), _ptr(gi), _ptr(out), ld, _stream()),
               "g2g_pack_rows")
    return out


def _host_matrix_pointer(X):
    """(address, leading dimension in floats, keep-alive object) of a row-major f32 host matrix, or None
    when ``X`` is not one (other dtype / layout / device)."""
    if torch.is_tensor(X):
        if X.device.type == "cpu" and X.dtype == torch.float32 and X.dim() == 2 and (X.shape[1] == 1 or X.stride(1) == 1) \
                and X.stride(0) >= X.shape[1]:
            return X.data_ptr(), int(X.stride(0)), X
        return None
    if isinstance(X, np.ndarray) and not isinstance(X, np.matrix) and X.dtype == np.float32 and X.ndim == 2 \
            and (X.shape[1] == 1 or X.strides[1] == 4) and X.strides[0] % 4 == 0 and X.strides[0] >= 4 * X.shape[1]:
        return X.ctypes.data, X.strides[0] // 4, X
    return None


_COPY_STREAMS = {}


def _copy_stream(device):
    key = torch.device(device).index
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


class HostBlockUpload:
    """Packed device copy of ``X[row_order][:, col_lo:col_hi][:, gene_idx_rel]`` for a HOST matrix ``X``
    (row-major f32, cells x all genes; ideally pinned), in two steps so that the PCIe stream can start before
    the pseudotime order of the cells is known:

    ``start()`` issues the first chunk copies on a copy stream -- only the column block crosses PCIe, as
    strided DMA copies of ``chunk_bytes`` (``g2g_upload_block``; whole rows go as 1-D copies) into two staging
    chunks; ``finish(row_order)`` scatters every chunk to its time-sorted rows (``g2g_scatter_rows``) while
    the next one is in flight and returns the [cells x ld] matrix.  No whole-matrix second copy exists at any
    time.  ``row_order`` / ``gene_idx_rel`` may be None (identity)."""

    def __init__(self, X, col_lo, col_hi, gene_idx_rel, device, chunk_bytes=96 << 20):
        self.lib = _lib.load()
        host = _host_matrix_pointer(X)
        if host is None:   # other dtypes / layouts: make a f32 host copy of the column block only
            blk = X[:, col_lo:col_hi]
            blk = blk.numpy() if torch.is_tensor(blk) else np.asarray(blk)
            host = _host_matrix_pointer(np.ascontiguousarray(blk, dtype=np.float32))
            col_hi, col_lo = col_hi - col_lo, 0
        self.addr, self.ld_src, self.keep = host
        self.device = device
        self.n_cells = int(X.shape[0])
        self.col_lo, self.width = int(col_lo), int(col_hi - col_lo)
        self.gene_idx_rel = gene_idx_rel
        self.n_sel = self.width if gene_idx_rel is None else len(gene_idx_rel)
        self.ld = (self.n_sel + 3) // 4 * 4
        # whole rows: a chunk is one contiguous range of bytes (a strided copy of rows that are not 16-byte
        # multiples runs at 2/3 of the PCIe rate: profiles/tools/upload_timing.py)
        self.full_rows = self.col_lo == 0 and self.width == self.ld_src
        self.rows_per = min(max(256, int(chunk_bytes // max(4 * self.width, 1))), self.n_cells)
        self.cs = _copy_stream(device)
        self.stage = None
        self.landed = []          # events of the chunk copies issued so far
        self.next_chunk = 0
        self.n_chunks = -(-self.n_cells // self.rows_per)

    def _copy_chunk(self, k):
        r0 = k * self.rows_per
        rows = min(self.rows_per, self.n_cells - r0)
        buf = self.stage[k & 1]
        cs = ctypes.c_void_p(self.cs.cuda_stream)
        if self.full_rows:
            _lib.check(self.lib.g2g_upload_block(ctypes.c_void_p(self.addr + 4 * r0 * self.ld_src), rows * self.ld_src, 0, 1,
                                                 0, rows * self.ld_src, _ptr(buf), rows * self.ld_src, cs), "g2g_upload_block")
        else:
            _lib.check(self.lib.g2g_upload_block(ctypes.c_void_p(self.addr), self.ld_src, r0, rows, self.col_lo, self.width,
                                                 _ptr(buf), self.width, cs), "g2g_upload_block")
        ev = torch.cuda.Event()
        ev.record(self.cs)
        self.landed.append(ev)

    def start(self):
        if self.stage is None:
            self.stage = [torch.empty((self.rows_per, self.width), dtype=torch.float32, device=self.device)
                          for _ in range(min(2, self.n_chunks))]
            self.cs.wait_stream(torch.cuda.current_stream())
            while self.next_chunk < min(2, self.n_chunks):
                self._copy_chunk(self.next_chunk)
                self.next_chunk += 1
        return self

    def finish(self, row_order):
        self.start()
        cur = torch.cuda.current_stream()
        n_cells, ld = self.n_cells, self.ld
        out = torch.empty((n_cells, ld), dtype=torch.float32, device=self.device)
        if row_order is not None:
            rank = np.empty(n_cells, dtype=np.int64)
            rank[np.asarray(row_order)] = np.arange(n_cells, dtype=np.int64)
        else:
            rank = np.arange(n_cells, dtype=np.int64)
        small = {"rank": rank}
        if self.gene_idx_rel is not None:
            small["gi"] = np.asarray(self.gene_idx_rel, dtype=np.int32)
        small = upload_small(small, self.device)
        rank_dev, gi = small["rank"], small.get("gi")
        for k in range(self.n_chunks):
            r0 = k * self.rows_per
            rows = min(self.rows_per, n_cells - r0)
            cur.wait_event(self.landed[k])
            _lib.check(self.lib.g2g_scatter_rows(_ptr(self.stage[k & 1]), rows, self.n_sel, self.width,
                                                 ctypes.c_void_p(rank_dev[r0:].data_ptr()), _ptr(gi), _ptr(out), ld, _stream()),
                       "g2g_scatter_rows")
            if self.next_chunk < self.n_chunks:      # the chunk two ahead reuses this staging buffer
                done = torch.cuda.Event()
                done.record(cur)
                self.cs.wait_event(done)
                self._copy_chunk(self.next_chunk)
                self.next_chunk += 1
        for b in self.stage:   # the staging chunks are reused by the allocator only after the scatter kernels read them
            b.record_stream(cur)
        out._g2g_keep = self.keep   # the host block must outlive the asynchronous copies
        return out


def upload_host_block(X, row_order, col_lo, col_hi, gene_idx_rel, device, chunk_bytes=96 << 20):
    """``HostBlockUpload(...).start().finish(row_order)`` in one call."""
    return HostBlockUpload(X, col_lo, col_hi, gene_idx_rel, device, chunk_bytes).finish(row_order)


def pack_csr(csr, row_order, gene_idx, n_all_genes, device):
    """Densify a scipy CSR matrix (cells x all genes) on the device (``g2g_pack_csr``): only the
    non-zeros cross PCIe."""
    lib = _lib.load()
    n_cells = csr.shape[0]
    n_genes = n_all_genes if gene_idx is None else len(gene_idx)
    ld = (n_genes + 3) // 4 * 4
    indptr = torch.from_numpy(np.ascontiguousarray(csr.indptr, dtype=np.int64)).to(device, non_blocking=True)
    indices = torch.from_numpy(np.ascontiguousarray(csr.indices, dtype=np.int32)).to(device, non_blocking=True)
    data = torch.from_numpy(np.ascontiguousarray(csr.data, dtype=np.float32)).to(device, non_blocking=True)
    col_map = None
    if gene_idx is not None:
        cm = np.full(n_all_genes, -1, dtype=np.int32)
        cm[np.asarray(gene_idx)] = np.arange(n_genes, dtype=np.int32)
        col_map = torch.from_numpy(cm).to(device)
    ro = None if row_order is None else torch.as_tensor(np.ascontiguousarray(row_order, dtype=np.int64)).to(device)
    out = torch.empty((n_cells, ld), dtype=torch.float32, device=device)
    _lib.check(lib.g2g_pack_csr(_ptr(indptr), _ptr(indices), _ptr(data), n_cells, _ptr(ro), _ptr(col_map),
                                _ptr(out), ld, _stream()), "g2g_pack_csr")
    return out


def run_cost_dp(trends, n_bins, trans, start, c_model, cost_in=None, dump=False, out=None):
    """Enqueue K4 (``g2g_cost_dp``) for the genes of ``trends`` ([G, 4, T] f64 on the device) on the
    current stream.  Needs no other device state, so per-gene dense dumps can be produced long after
    the batch buffers are gone."""
    lib = _lib.load()
    T = int(n_bins)
    G = int(trends.shape[0])
    dev = trends.device
    if out is None:
        out = dict(align_str=torch.empty((G, 2 * T), dtype=torch.uint8, device=dev),
                   align_len=torch.empty(G, dtype=torch.int32, device=dev),
                   opt_cost=torch.empty(G, dtype=torch.float64, device=dev),
                   l_states=torch.empty((G, (T + 1) * (T + 1)), dtype=torch.uint8, device=dev))
    cost_out = mats = bp = l_out = None
    if dump:
        cost_out = torch.empty((G, T, T), dtype=torch.float64, device=dev)
        mats = torch.empty((G, 5, T + 1, T + 1), dtype=torch.float64, device=dev)
        bp = torch.empty((G, 5, T + 1, T + 1), dtype=torch.int8, device=dev)
        l_out = torch.empty((G, T + 1, T + 1), dtype=torch.float64, device=dev)
        out.update(cost=cost_out, mats=mats, bp=bp, L=l_out)
    trans_c = (c_f64 * 25)(*np.asarray(trans, dtype=np.float64).reshape(-1).tolist())
    start_c = (c_f64 * 3)(*np.asarray(start, dtype=np.float64).tolist())
    _lib.check(lib.g2g_cost_dp(_ptr(trends), G, T, trans_c, start_c, float(c_model), N_SAMPLES, _ptr(cost_in),
                               _ptr(out["align_str"]), _ptr(out["align_len"]), _ptr(out["opt_cost"]),
                               _ptr(out["l_states"]), _ptr(cost_out), _ptr(mats), _ptr(bp), _ptr(l_out),
                               _stream()), "g2g_cost_dp")
    return out


def pack_csc(csr, row_order, gene_idx, n_all_genes, device):
    """scipy CSR (cells x all genes, canonical: no duplicate entries) -> time-sorted, gene-selected
    ``CscMatrix`` on the device.  Only the stored entries cross PCIe; the transposition is
    ``g2g_csr_to_csc`` (a stable counting sort by column: rows stay in time order inside every column)."""
    lib = _lib.load()
    n_cells = int(csr.shape[0])
    n_genes = int(n_all_genes if gene_idx is None else len(gene_idx))
    indptr = torch.from_numpy(np.ascontiguousarray(csr.indptr, dtype=np.int64)).to(device, non_blocking=True)
    cols = torch.from_numpy(np.ascontiguousarray(csr.indices, dtype=np.int32)).to(device, non_blocking=True)
    vals = torch.from_numpy(np.ascontiguousarray(csr.data, dtype=np.float32)).to(device, non_blocking=True)
    ro = cm = None
    if row_order is not None:
        ro = torch.from_numpy(np.ascontiguousarray(row_order, dtype=np.int64)).to(device, non_blocking=True)
    if gene_idx is not None:
        cm_h = np.full(n_all_genes, -1, dtype=np.int32)
        cm_h[np.asarray(gene_idx)] = np.arange(n_genes, dtype=np.int32)
        cm = torch.from_numpy(cm_h).to(device, non_blocking=True)
    nnz = int(vals.numel())
    col_ptr = torch.empty(n_genes + 1, dtype=torch.int64, device=device)
    row_idx = torch.empty(max(nnz, 1), dtype=torch.int32, device=device)
    values = torch.empty(max(nnz, 1), dtype=torch.float32, device=device)
    ws_bytes = lib.g2g_csr_to_csc_workspace_bytes(n_cells, n_genes)
    ws = torch.empty(max((ws_bytes + 3) // 4, 1), dtype=torch.int32, device=device)
    _lib.check(lib.g2g_csr_to_csc(_ptr(indptr), _ptr(cols), _ptr(vals), n_cells, _ptr(ro), _ptr(cm), n_genes,
                                  _ptr(col_ptr), _ptr(row_idx), _ptr(values), _ptr(ws), ws_bytes, _stream()),
               "g2g_csr_to_csc")
    kept = nnz if gene_idx is None else int(col_ptr[-1].item())    # entries of the selected genes
    return CscMatrix(col_ptr, row_idx[:kept], values[:kept], n_cells, n_genes)


def result_shapes(n_rows, n_bins):
    """name -> (shape, dtype) of the per-gene result arrays for ``n_rows`` genes."""
    C, T = int(n_rows), int(n_bins)
    return {"opt_cost": ((C,), torch.float64), "trends": ((C, 4, T), torch.float64),
            "musigma": ((C, 4, T), torch.float64), "align_len": ((C,), torch.int32),
            "flags": ((C,), torch.int32), "nnz": ((C, 2), torch.int32),
            "align_str": ((C, 2 * T), torch.uint8), "l_states": ((C, (T + 1) * (T + 1)), torch.uint8)}


def unpack_results(raw, n_rows, n_bins, n_valid=None):
    """Views of a packed result buffer (numpy u8, laid out for ``n_rows`` genes): name -> array of the
    first ``n_valid`` genes."""
    shapes = result_shapes(n_rows, n_bins)
    host, off = {}, 0
    for k in AlignmentEngine.RESULT_KEYS:
        shape, dt = shapes[k]
        n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        arr = raw[off:off + n].view(_NP_DTYPES[dt]).reshape(shape)
        host[k] = arr if n_valid is None else arr[:n_valid]
        off += n
    return host


class HostResults:
    """name -> numpy array: the per-gene result arrays of a batch on the host (dict-like, read-only).

    The device->host copy lands in a pinned staging buffer.  The small arrays every result object needs
    (optimal costs, lengths, flags, counts, alignment strings) are copied out of it at once; the large ones
    (trends, interpolated means / stds, landscape states: 97 % of the bytes) move to their own arrays on first
    access, and the staging buffer goes back to the pool when the last of them has moved or this object dies.
    ``segments`` = [(raw u8 view laid out for ``capacity`` genes, capacity, genes held, first gene index)]:
    one for a single GPU, one per rank after a gather."""

    EAGER = ("opt_cost", "align_len", "flags", "nnz", "align_str")

    def __init__(self, segments, n_genes, n_bins, pinned=None):
        self._segments, self._n, self._T, self._pinned = segments, int(n_genes), int(n_bins), pinned
        self._arrays = {}
        self._shapes = result_shapes(self._n, self._T)
        self._views = [(unpack_results(raw, cap, self._T, held), lo, held) for raw, cap, held, lo in segments]
        for k in self.EAGER:
            self._materialize(k)

    def _materialize(self, key):
        shape, dt = self._shapes[key]
        out = np.empty(shape, dtype=_NP_DTYPES[dt])
        for views, lo, held in self._views:
            if held:
                out[lo:lo + held] = views[key]
        self._arrays[key] = out
        if len(self._arrays) == len(AlignmentEngine.RESULT_KEYS):
            self._release()
        return out

    def _release(self):
        self._views = self._segments = None
        if self._pinned is not None:
            _give_pinned(self._pinned)
            self._pinned = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def __getitem__(self, key):
        arr = self._arrays.get(key)
        if arr is None:
            if key not in self._shapes:
                raise KeyError(key)
            arr = self._materialize(key)
        return arr

    def __contains__(self, key):
        return key in self._shapes

    def __iter__(self):
        return iter(AlignmentEngine.RESULT_KEYS)

    def __len__(self):
        return len(AlignmentEngine.RESULT_KEYS)

    def keys(self):
        return list(AlignmentEngine.RESULT_KEYS)

    def values(self):
        return [self[k] for k in AlignmentEngine.RESULT_KEYS]

    def items(self):
        return [(k, self[k]) for k in AlignmentEngine.RESULT_KEYS]

    def get(self, key, default=None):
        return self[key] if key in self._shapes else default


class AlignmentEngine:
    """Runs the whole hot path for one (reference, query) pair of packed device matrices.

    ``run()`` enqueues K0 (weights) + K1a (pilot) + K1 (moments) and the fused K3 + K4 launch (fix-up /
    trends, cost + DP + traceback) and returns itself; the results are views of ``self.packed``.
    """

    #: smallest number of interpolation points the tensor-core interpolation kernel is used for: below it
    #: the FFMA2 register tile of g2g_moments is at the HBM / FP32 ridge already
    MMA_MIN_BINS = 30

    def __init__(self, X_ref, tau_ref, X_query, tau_query, n_genes, n_bins, denom_ref, denom_query,
                 free_params=(0.99, 0.1, 0.7), device=None, moments="auto", capacity=None):
        """``moments``: ``"auto"`` (tensor cores from ``MMA_MIN_BINS`` interpolation points, else the FFMA2
        kernel), ``"mma"`` or ``"ffma"``.  ``capacity`` (>= n_genes): rows the packed result arrays are
        laid out for, so that ranks with different gene counts share one record layout (multi-GPU gather)."""
        self.lib = _lib.load()
        _lib.check(self.lib.g2g_check_device(), "g2g_check_device")
        self.device = device if device is not None else X_ref.device
        self.n_genes, self.n_bins = int(n_genes), int(n_bins)
        self.sparse = isinstance(X_ref, CscMatrix)
        if self.sparse != isinstance(X_query, CscMatrix):
            raise ValueError("reference and query must both be dense or both be CscMatrix")
        T, G, dev = self.n_bins, self.n_genes, self.device
        self.points_host = np.linspace(0, 1, T)  # TimeSeriesPreprocessor.py:23
        self.eps = epsilon_table(T)
        f64 = dict(dtype=torch.float64, device=dev)
        small = upload_small({
            "tau_ref": np.asarray(tau_ref, dtype=np.float64), "tau_query": np.asarray(tau_query, dtype=np.float64),
            "denom_ref": np.asarray(denom_ref, dtype=np.float64),
            "denom_query": np.asarray(denom_query, dtype=np.float64), "points": self.points_host,
            "eps_mean": self.eps.mean(axis=1), "eps_std": self.eps.std(axis=1)}, dev)
        if moments not in ("auto", "mma", "ffma"):
            raise ValueError("moments must be 'auto', 'mma' or 'ffma'")
        if self.sparse:
            self.variant = "csc"
        elif moments == "mma" or (moments == "auto" and T >= self.MMA_MIN_BINS):
            self.variant = "mma"
        else:
            self.variant = "ffma"
        self.ref = DeviceSide(X_ref, small["tau_ref"], G, T, small["denom_ref"], dev, self.variant)
        self.query = DeviceSide(X_query, small["tau_query"], G, T, small["denom_query"], dev, self.variant)
        self.k1_status = torch.zeros(2, dtype=torch.int32, device=dev)
        #: SMs the persistent interpolation kernel leaves free (for a collective of the previous batch that runs
        #: on another stream while this batch computes; 0 = use every SM)
        self.sm_reserve = 0
        self.points, self.eps_mean, self.eps_std = small["points"], small["eps_mean"], small["eps_std"]
        self.set_state_params(free_params)
        self.start = start_costs()
        self.c_model = model_constant()
        # every per-gene result lives in ONE packed byte buffer (8-, 4-, then 1-byte element types, so
        # each view is aligned): the device->host copy and the multi-GPU gather move it as it is
        self.capacity = C = int(capacity) if capacity is not None else G
        if C < G:
            raise ValueError("capacity must be >= n_genes")
        shapes = result_shapes(C, T)
        sizes = [int(np.prod(shapes[k][0])) * torch.empty((), dtype=shapes[k][1]).element_size()
                 for k in self.RESULT_KEYS]
        self.packed = torch.zeros(max(sum(sizes), 8), dtype=torch.uint8, device=dev)
        off = 0
        for k, n in zip(self.RESULT_KEYS, sizes):
            full = self.packed[off:off + n].view(shapes[k][1]).reshape(shapes[k][0])
            setattr(self, k, full[:G])     # the kernels write the first G rows
            off += n
        self._bind_sides()
        self.fix_ws_bytes = self.lib.g2g_fixup_workspace_bytes(G, T)
        self.fix_ws = torch.empty((self.fix_ws_bytes + 7) // 8, **f64)
        self.launches_per_run = 0
        self._host_buf = None
        self._side_stream = torch.cuda.Stream(device=dev)
        self._side_stream2 = torch.cuda.Stream(device=dev)
        self._ev_fork = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()
        self._ev_join2 = torch.cuda.Event()

    def _bind_sides(self):
        self._sides = (Side * 2)(self.ref.c_side(), self.query.c_side())
        self._sparse_sides = (SparseSide * 2)(self.ref.c_sparse_side(), self.query.c_sparse_side()) \
            if self.sparse else None

    def use_variant(self, variant):
        """Switch the dense interpolation kernel (``"mma"`` / ``"ffma"``); buffers are re-laid out."""
        if self.sparse or variant == self.variant:
            return
        self.variant = variant
        self.ref.configure(variant)
        self.query.configure(variant)
        self._bind_sides()

    def set_state_params(self, free_params):
        self.free_params = tuple(float(v) for v in free_params)
        self.trans = transition_costs(self.free_params)

    # -- stages -----------------------------------------------------------------------------
    def run_moments(self):
        """First half of the hot path: K0 weight tables, K1a pilot means, K1m / K1 weighted moments."""
        return self.run_interpolation(with_dp=None)

    def run_fused(self):
        """Second half: the fused fix-up + cost + DP launch on the partial sums of ``run_moments``."""
        self._launch_fused()

    def run_interpolation(self, with_dp=False):
        """K0 / pilot / K1 and the per-gene fix-up.  ``with_dp=True`` finishes with the fused fix-up +
        cost + DP launch (``g2g_fixup_cost_dp``) instead of ``g2g_fixup_trends``: the whole hot path
        (``with_dp=None``: stop after the moments)."""
        lib, T, G = self.lib, self.n_bins, self.n_genes
        cur = torch.cuda.current_stream()
        # four independent small kernels: the weight tables of the two datasets go to two forked streams,
        # the pilot means stay on the current one
        self._ev_fork.record(cur)
        for s, stream in ((self.query, self._side_stream), (self.ref, self._side_stream2)):
            stream.wait_event(self._ev_fork)
            st = ctypes.c_void_p(stream.cuda_stream)
            weights = lib.g2g_weights_mma if self.variant == "mma" else lib.g2g_weights
            _lib.check(weights(_ptr(s.tau), s.n_cells, _ptr(self.points), _ptr(s.denom), T, _ptr(s.wtab),
                               _ptr(s.sum_w), _ptr(s.ws), s.ws_bytes, st), "g2g_weights")
        if not self.sparse:
            st = _stream()
            for s in (self.ref, self.query):
                _lib.check(lib.g2g_pilot_mean(_ptr(s.X), s.n_cells, G, s.ld, _ptr(s.pilot), st), "g2g_pilot_mean")
        self._ev_join.record(self._side_stream)
        self._ev_join2.record(self._side_stream2)
        cur.wait_event(self._ev_join)
        cur.wait_event(self._ev_join2)
        st = _stream()
        probe = self._probe
        if probe is not None:
            probe[0].record()
        self._launch_moments(st)
        if probe is not None:
            probe[1].record()
        if with_dp is None:
            return 5 if self.sparse else 7
        if with_dp:
            trans_c = (c_f64 * 25)(*np.asarray(self.trans, dtype=np.float64).reshape(-1).tolist())
            start_c = (c_f64 * 3)(*np.asarray(self.start, dtype=np.float64).tolist())
            _lib.check(lib.g2g_fixup_cost_dp(self._sides, _ptr(self.ref.sum_w), _ptr(self.query.sum_w),
                                             _ptr(self.points), _ptr(self.eps_mean), _ptr(self.eps_std), G, G, T,
                                             trans_c, start_c, float(self.c_model), N_SAMPLES, _ptr(self.musigma),
                                             _ptr(self.trends), _ptr(self.nnz), _ptr(self.flags),
                                             _ptr(self.align_str), _ptr(self.align_len), _ptr(self.opt_cost),
                                             _ptr(self.l_states), st), "g2g_fixup_cost_dp")
            if probe is not None:
                probe[2].record()
            return 6 if self.sparse else 8  # 2 x (weights + its per-bin sums[, pilot]) + moments + fused fix-up / DP
        _lib.check(lib.g2g_fixup_trends(self._sides, _ptr(self.ref.sum_w), _ptr(self.query.sum_w), _ptr(self.points),
                                        _ptr(self.eps_mean), _ptr(self.eps_std), G, G, T, _ptr(self.musigma),
                                        _ptr(self.trends), _ptr(self.nnz), _ptr(self.flags), _ptr(self.fix_ws),
                                        self.fix_ws_bytes, st), "g2g_fixup_trends")
        return 7 if self.sparse else 9  # 2 x (weights + sums[, pilot]) + moments + partial reduce + fix-up

    def run_dp(self, trends=None, cost_in=None, dump=False, gene_slice=None, force_new=False):
        """K4 on all genes (or ``gene_slice`` / the given ``trends``).  ``dump=True`` also returns the
        dense matrices; results go to fresh tensors unless this is the plain whole-batch run."""
        trends = self.trends if trends is None else trends
        if gene_slice is not None:
            trends = trends[gene_slice].contiguous()
        out = None
        if gene_slice is None and not dump and cost_in is None and not force_new and trends is self.trends:
            out = dict(align_str=self.align_str, align_len=self.align_len, opt_cost=self.opt_cost,
                       l_states=self.l_states)
        return run_cost_dp(trends, self.n_bins, self.trans, self.start, self.c_model, cost_in=cost_in, dump=dump,
                           out=out)

    def _launch_moments(self, st):
        lib, T, G = self.lib, self.n_bins, self.n_genes
        if self.variant == "csc":
            _lib.check(lib.g2g_moments_csc(self._sparse_sides, 2, G, G, T, st), "g2g_moments_csc")
        elif self.variant == "mma":
            _lib.check(lib.g2g_moments_mma_ex(self._sides, 2, G, G, T, _ptr(self.k1_status), int(self.sm_reserve), st),
                       "g2g_moments_mma")
        else:
            _lib.check(lib.g2g_moments(self._sides, 2, G, G, T, st), "g2g_moments")

    _probe = None

    def time_step(self, before_moments, after_moments, after_fused):
        """One whole step (``run()``) with three CUDA events recorded on the current stream around the moments
        launch and after the fused launch: the kernels' durations inside a step, not alone."""
        self._probe = (before_moments, after_moments, after_fused)
        try:
            self.run_interpolation(with_dp=True)
        finally:
            self._probe = None

    def time_moments(self, start_event, end_event):
        """Launch K1 alone between two CUDA events on the current stream (roofline measurement;
        the weight tables and pilot values of the last run are reused)."""
        start_event.record()
        self._launch_moments(_stream())
        end_event.record()

    def time_fused(self, start_event, end_event):
        """Launch the fused fix-up + cost + DP kernel alone between two CUDA events (its inputs are the partial
        sums of the last run)."""
        start_event.record()
        self._launch_fused()
        end_event.record()

    def _launch_fused(self):
        T, G = self.n_bins, self.n_genes
        trans_c = (c_f64 * 25)(*np.asarray(self.trans, dtype=np.float64).reshape(-1).tolist())
        start_c = (c_f64 * 3)(*np.asarray(self.start, dtype=np.float64).tolist())
        _lib.check(self.lib.g2g_fixup_cost_dp(self._sides, _ptr(self.ref.sum_w), _ptr(self.query.sum_w),
                                              _ptr(self.points), _ptr(self.eps_mean), _ptr(self.eps_std), G, G, T,
                                              trans_c, start_c, float(self.c_model), N_SAMPLES, _ptr(self.musigma),
                                              _ptr(self.trends), _ptr(self.nnz), _ptr(self.flags),
                                              _ptr(self.align_str), _ptr(self.align_len), _ptr(self.opt_cost),
                                              _ptr(self.l_states), _stream()), "g2g_fixup_cost_dp")

    def run(self, fused=True):
        """The whole hot path on the current stream.  ``fused=False`` runs the fix-up and the DP as the
        two separate entry points (same results bit for bit; one more launch and a workspace pass)."""
        if fused:
            self.launches_per_run = self.run_interpolation(with_dp=True)
        else:
            n = self.run_interpolation()
            self.run_dp()
            self.launches_per_run = n + 1
        return self

    RESULT_KEYS = ("opt_cost", "trends", "musigma", "align_len", "flags", "nnz", "align_str", "l_states")

    def results_to_host(self):
        """All result arrays in ONE device->host copy: they are packed into a byte buffer on the
        device (8-, 4-, then 1-byte element types so every view stays aligned) and land in a
        pinned host buffer."""
        self.begin_results_to_host()
        return self.finish_results_to_host()

    def begin_results_to_host(self):
        """Enqueue the packed device->host copy on the current stream and return at once."""
        self._host_buf = _take_pinned(self.packed.numel())   # private until finish_results_to_host returns it
        self._host_buf.copy_(self.packed, non_blocking=True)
        if self.variant == "mma":
            self._status_host = _take_pinned(8)
            self._status_host.copy_(self.k1_status.view(torch.uint8), non_blocking=True)
        self._pending_parts = [getattr(self, k) for k in self.RESULT_KEYS]

    def redo_if_out_of_range(self):
        """Tensor-core kernel only: read its status word (one small synchronous copy) and, if a value did not
        fit the f16 split (|x| >= 65504 or |x - pilot| >= 255) or the input is not finite, redo the batch with
        the f32 kernel, which has no range limit.  Returns True when it did.  (The single-GPU result copy
        checks the status word as part of ``finish_results_to_host``; the multi-GPU path calls this before
        the gather.)"""
        if self.variant != "mma":
            return False
        if not int(self.k1_status[0].item()) & 1:
            return False
        self.k1_status.zero_()
        self.use_variant("ffma")
        self.run_interpolation(with_dp=True)
        return True

    def finish_results_to_host(self):
        """Wait for the copy started by ``begin_results_to_host``; returns a ``HostResults`` (name -> numpy
        array) over the private pinned staging buffer."""
        torch.cuda.current_stream().synchronize()
        if self.variant == "mma":
            status = int(self._status_host.numpy().view(np.int32)[0])
            _give_pinned(self._status_host)
            self._status_host = None
            if status & 1:
                # a value did not fit the f16 split of the tensor-core kernel (|x - pilot| >= 255) or the input
                # is not finite: redo the batch with the f32 kernel, which has no range limit
                _give_pinned(self._host_buf)
                self.k1_status.zero_()
                self.use_variant("ffma")
                self.run_interpolation(with_dp=True)
                self.begin_results_to_host()
                return self.finish_results_to_host()
        self._pending_parts = None
        buf, self._host_buf = self._host_buf, None
        return HostResults([(buf.numpy(), self.capacity, self.n_genes, 0)], self.n_genes, self.n_bins, pinned=buf)


_PINNED = {}      # n_bytes -> [pinned buffer, event of the last copy that read or wrote it]  (upload staging)
_PINNED_FREE = {}  # n_bytes -> list of idle pinned result buffers


def _take_pinned(n_bytes):
    """A pinned host buffer nobody else holds (result copies): taken from / returned to a free list, so
    two live engines or interleaved begin / finish calls never share one."""
    free = _PINNED_FREE.setdefault(int(n_bytes), [])
    return free.pop() if free else torch.empty(int(n_bytes), dtype=torch.uint8, pin_memory=True)


def _give_pinned(buf):
    free = _PINNED_FREE.setdefault(int(buf.numel()), [])
    if len(free) < 4:
        free.append(buf)


def _pinned_buffer(n_bytes):
    """Pinned staging buffers are expensive to create (cudaHostAlloc); keep one per size.  A buffer
    that still has a copy in flight is waited for before it is handed out again."""
    entry = _PINNED.get(n_bytes)
    if entry is None:
        if len(_PINNED) > 16:
            for _, ev in _PINNED.values():
                if ev is not None:
                    ev.synchronize()
            _PINNED.clear()
        entry = [torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True), None]
        _PINNED[n_bytes] = entry
    if entry[1] is not None:
        entry[1].synchronize()
        entry[1] = None
    return entry[0]


def _mark_in_flight(buf):
    ev = torch.cuda.Event()
    ev.record()
    _PINNED[buf.numel()][1] = ev


_NP_DTYPES = {torch.float64: np.float64, torch.int32: np.int32, torch.uint8: np.uint8}
_TORCH_DTYPES = {np.float64: torch.float64, np.float32: torch.float32, np.int64: torch.int64, np.int32: torch.int32,
                 np.uint8: torch.uint8}
