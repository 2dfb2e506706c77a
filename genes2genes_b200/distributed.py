"""Multi-GPU plumbing: genes are independent (reference loop: Main.py:452-455), so each rank
aligns a contiguous block of ``gene_list`` and the only collective is one gather of fixed-size
per-gene result records to rank 0.  One process per GPU,
``torch.distributed`` with the NCCL backend (gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

# result arrays gathered per gene, in record order
RECORD_KEYS = ("opt_cost", "trends", "musigma", "align_len", "flags", "nnz", "align_str", "l_states")  # 8-, 4-, 1-byte types in turn so every view stays aligned


def shard_range(n_genes, rank, world):
    """Contiguous block [lo, hi) of genes owned by ``rank``.  Balanced split: block sizes differ by at
    most one and no rank is left empty while another holds two genes (with fewer genes than ranks the
    surplus ranks own an empty block, which every caller must accept)."""
    return n_genes * rank // world, n_genes * (rank + 1) // world


def max_shard(n_genes, world):
    """Largest block ``shard_range`` hands out: the fixed record count every rank sends to the gather."""
    return -(-n_genes // world) if n_genes else 0


def pack_records(tensors):
    """Concatenate the byte views of the result tensors (each [G_local, ...]) into one u8 buffer."""
    return torch.cat([tensors[k].contiguous().view(torch.uint8).reshape(-1) for k in RECORD_KEYS])


def unpack_records(buf, templates):
    """Inverse of ``pack_records``: ``templates[k]`` = (shape, dtype) of each array."""
    out, off = {}, 0
    for k in RECORD_KEYS:
        shape, dtype = templates[k]
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        out[k] = buf[off:off + n].view(dtype).reshape(shape)
        off += n
    return out


def record_templates(n_genes, n_bins):
    T, G = n_bins, n_genes
    return {"align_str": ((G, 2 * T), torch.uint8), "align_len": ((G,), torch.int32),
            "opt_cost": ((G,), torch.float64), "l_states": ((G, (T + 1) * (T + 1)), torch.uint8),
            "trends": ((G, 4, T), torch.float64), "musigma": ((G, 4, T), torch.float64),
            "flags": ((G,), torch.int32), "nnz": ((G, 2), torch.int32)}


def record_bytes(n_genes, n_bins):
    return sum(int(np.prod(s)) * torch.empty((), dtype=d).element_size()
               for s, d in record_templates(n_genes, n_bins).values())


class ResultGather:
    """Gathers every rank's result records to ``dst`` with a single collective.

    All ranks must hold the same number of genes per call (pad the last shard); buffers are
    allocated once so the gather can sit inside a timed loop.
    """

    def __init__(self, source, dst=0, group=None):
        self.source = source  # object with the RECORD_KEYS tensors as attributes (AlignmentEngine)
        self.dst, self.group = dst, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        # an engine keeps its records packed in this very layout: send that buffer as it is
        self.packed = getattr(source, "packed", None)
        probe = self.packed if self.packed is not None else pack_records({k: getattr(source, k) for k in RECORD_KEYS})
        self.send = probe if self.packed is not None else torch.empty_like(probe)
        self.recv = [torch.empty_like(probe) for _ in range(self.world)] if self.rank == dst else None

    def run(self):
        if self.packed is None:
            torch.cat([getattr(self.source, k).contiguous().view(torch.uint8).reshape(-1) for k in RECORD_KEYS],
                      out=self.send)
        dist.gather(self.send, self.recv, dst=self.dst, group=self.group)
        return self.recv

    def host_results(self, n_genes_per_rank, n_bins):
        """On ``dst``: list (one per rank) of dicts of numpy arrays."""
        if self.recv is None:
            return None
        tmpl = record_templates(n_genes_per_rank, n_bins)
        return [{k: v.cpu().numpy() for k, v in unpack_records(b, tmpl).items()} for b in self.recv]


class _DevicePointer:
    """Raw device memory as a torch tensor (CUDA array interface)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class PeerWindow:
    """Rank ``dst``'s receive window ``[world, slot_bytes]``, mapped into every rank of the box
    (``g2g_peer_window_*``): a rank's packed result records reach their slot with ONE asynchronous
    device-to-device copy over NVLink / NVSwitch, executed by a copy engine -- no kernel, so the gather of
    batch k never takes SMs from the persistent interpolation kernel of batch k+1 (an NCCL gather does).

    Collective constructor (every rank of the default group calls it); raises ``RuntimeError`` on every rank
    when any rank cannot map the window (no peer access, ranks on different hosts): callers then use the
    NCCL / gloo gather.  Ordering is the caller's: before ``push`` the window must be free (``dst`` has read
    the previous contents), after it a stream synchronise + barrier tells ``dst`` that every slot has landed."""

    def __init__(self, slot_bytes, device, dst=0):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world, self.dst = dist.get_rank(), dist.get_world_size(), dst
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.device = torch.device(device)
        self.ptr, self.owner = None, self.rank == dst
        handle = ctypes.create_string_buffer(64)
        ptr = ctypes.c_void_p()
        error = None
        payload = [None]
        with torch.cuda.device(self.device):
            if self.owner:
                try:
                    _lib.check(self.lib.g2g_peer_window_create(self.slot_bytes * self.world, ctypes.byref(ptr), handle),
                               "g2g_peer_window_create")
                    self.ptr = ptr.value
                    payload = [handle.raw]
                except Exception as exc:       # reported to every rank below
                    error = repr(exc)
            self._broadcast(payload, dst)
            if not self.owner:
                if payload[0] is None:
                    error = "the gathering rank could not create the window"
                else:
                    try:
                        _lib.check(self.lib.g2g_peer_window_open(payload[0], ctypes.byref(ptr)), "g2g_peer_window_open")
                        self.ptr = ptr.value
                    except Exception as exc:
                        error = repr(exc)
            errors = [None] * self.world
            dist.all_gather_object(errors, error)
        if any(e is not None for e in errors):
            self.close()
            raise RuntimeError("peer window unavailable: %s" % next(e for e in errors if e is not None))

    @staticmethod
    def _broadcast(payload, src):
        dist.broadcast_object_list(payload, src=src)

    def push(self, records, stream=None):
        """Copy this rank's packed records (u8 device tensor) into its slot, asynchronously on ``stream``
        (default: the current stream)."""
        from . import _lib
        n = records.numel() * records.element_size()
        if n > self.slot_bytes:
            raise ValueError("records (%d bytes) do not fit the window slot (%d)" % (n, self.slot_bytes))
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check(self.lib.g2g_peer_push(self.ptr + self.rank * self.slot_bytes, records.data_ptr(), n, st.cuda_stream),
                   "g2g_peer_push")

    def rows(self):
        """On ``dst``: the window as a ``[world, slot_bytes]`` u8 device tensor (a view, not a copy)."""
        if not self.owner:
            return None
        return torch.as_tensor(_DevicePointer(self.ptr, (self.world, self.slot_bytes)), device=self.device)

    def close(self):
        if self.ptr is not None:
            with torch.cuda.device(self.device):
                self.lib.g2g_peer_window_close(self.ptr, 1 if self.owner else 0)
            self.ptr = None

    def close_collective(self):
        """Every rank calls this: the peers unmap their views before the owner frees the memory."""
        if not self.owner:
            self.close()
        dist.barrier()
        if self.owner:
            self.close()


_windows = {}          # insertion-ordered: (slot bytes, device, world) -> PeerWindow or None
_MAX_WINDOWS = 2       # a process that aligns batches of many shapes keeps the two most recent windows


def peer_window(slot_bytes, device):
    """The process-wide window for records of ``slot_bytes`` (made on first use; ``None`` when the box cannot
    map one -- the answer is the same on every rank).  Collective on first use: every rank must ask for the
    same sequence of sizes (they do: the size is a function of the gene list and the bin count)."""
    key = (int(slot_bytes), str(device), dist.get_world_size())
    if key not in _windows:
        live = [k for k, w in _windows.items() if w is not None]
        if len(live) >= _MAX_WINDOWS:
            _windows.pop(live[0]).close_collective()
        try:
            _windows[key] = PeerWindow(slot_bytes, device)
        except RuntimeError:
            _windows[key] = None
    return _windows[key]
