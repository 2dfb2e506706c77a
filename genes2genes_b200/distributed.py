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
