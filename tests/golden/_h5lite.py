"""Minimal pure-Python HDF5 reader -- just enough for the two bundled tutorial
``.h5ad`` files of the reference (``notebooks/data/adata_{pam,lps}_local.h5ad``).

Test infrastructure only (used by ``make_golden.py`` in the build container;
``h5py``/``anndata`` are not installed there).  Handles: superblock v0,
symbol-table groups (B-tree v1 + local heap), v1 object headers with
continuation blocks, contiguous/compact/chunked(un-filtered, 1-D) layouts,
fixed-point / IEEE float / fixed strings / variable-length strings through the
global heap, and the attribute messages needed to read ``shape``.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Lite:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert self.b[8] == 0, "only superblock v0 supported"
        assert self.b[13] == 8 and self.b[14] == 8
        # root group symbol-table entry starts at byte 56
        self.root = self._obj(struct.unpack_from("<Q", self.b, 64)[0])

    # -- object headers ---------------------------------------------------
    def _messages(self, addr):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1, "only v1 object headers supported"
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:  # continuation
                    off, ln = struct.unpack_from("<QQ", b, body)
                    blocks.append((off, ln))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    def _obj(self, addr):
        msgs = self._messages(addr)
        o = {"addr": addr, "attrs": {}}
        for mtype, p, sz in msgs:
            if mtype == 0x11:
                o["btree"], o["heap"] = struct.unpack_from("<QQ", self.b, p)
            elif mtype == 0x01:
                o["shape"] = self._dataspace(p)
            elif mtype == 0x03:
                o["dtype"] = self._datatype(p)
            elif mtype == 0x08:
                o["layout"] = self._layout(p)
            elif mtype == 0x0B:
                o["filtered"] = True
            elif mtype == 0x0C:
                try:
                    k, v = self._attribute(p)
                    o["attrs"][k] = v
                except Exception:  # attributes we cannot decode are ignored
                    pass
        return o

    # -- messages -----------------------------------------------------------
    def _dataspace(self, p):
        b = self.b
        ver, rank, flags = struct.unpack_from("<BBB", b, p)
        if ver == 1:
            q = p + 8
        else:
            q = p + 4
        return tuple(struct.unpack_from("<%dQ" % rank, b, q)) if rank else ()

    def _datatype(self, p):
        b = self.b
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, p)
        cls = cv & 0x0F
        if cls == 0:
            signed = (b0 >> 3) & 1
            return ("int", np.dtype("<%s%d" % ("i" if signed else "u", size)))
        if cls == 1:
            return ("float", np.dtype("<f%d" % size))
        if cls == 3:
            return ("str", size)
        if cls == 9:
            return ("vlen", (b0 & 0x0F), self._datatype(p + 8))
        if cls == 8:  # enum (used for booleans); base type follows
            return ("enum", self._datatype(p + 8), size)
        raise NotImplementedError("datatype class %d" % cls)

    def _layout(self, p):
        b = self.b
        ver, cls = struct.unpack_from("<BB", b, p)
        assert ver == 3, "only layout v3 supported"
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", b, p + 2)
            return ("contiguous", addr, size)
        if cls == 0:
            (size,) = struct.unpack_from("<H", b, p + 2)
            return ("compact", p + 4, size)
        if cls == 2:
            rank = b[p + 2]
            (addr,) = struct.unpack_from("<Q", b, p + 3)
            dims = struct.unpack_from("<%dI" % rank, b, p + 11)
            return ("chunked", addr, dims)
        raise NotImplementedError

    def _attribute(self, p):
        b = self.b
        ver = b[p]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, p + 2)
            q = p + 8
            pad = lambda n: (n + 7) & ~7
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, p + 2)
            q = p + 8 + (1 if ver == 3 else 0)
            pad = lambda n: n
        else:
            raise NotImplementedError
        name = b[q:q + nsz].split(b"\0")[0].decode()
        q += pad(nsz)
        dt = self._datatype(q)
        q += pad(tsz)
        shape = self._dataspace(q)
        q += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        val = self._decode(dt, q, n)
        if not shape:
            val = val[0]
        return name, val

    # -- raw data -----------------------------------------------------------
    def _decode(self, dt, addr, n):
        b = self.b
        kind = dt[0]
        if kind in ("int", "float"):
            return np.frombuffer(b, dtype=dt[1], count=n, offset=addr).copy()
        if kind == "enum":
            return self._decode(dt[1], addr, n)
        if kind == "str":
            size = dt[1]
            return [b[addr + i * size: addr + (i + 1) * size].split(b"\0")[0].decode()
                    for i in range(n)]
        if kind == "vlen":
            out = []
            for i in range(n):
                ln, gaddr, idx = struct.unpack_from("<IQI", b, addr + 16 * i)
                out.append(self._gheap(gaddr, idx)[:ln].decode() if ln else "")
            return out
        raise NotImplementedError(kind)

    def _gheap(self, addr, idx):
        b = self.b
        assert b[addr:addr + 4] == b"GCOL"
        (csize,) = struct.unpack_from("<Q", b, addr + 8)
        p = addr + 16
        end = addr + csize
        while p + 16 <= end:
            oid, _rc, _, osz = struct.unpack_from("<HHIQ", b, p)
            if oid == idx:
                return b[p + 16:p + 16 + osz]
            if oid == 0:
                break
            p += 16 + ((osz + 7) & ~7)
        raise KeyError("global heap object %d" % idx)

    # -- groups ---------------------------------------------------------------
    def _heap_data(self, haddr):
        assert self.b[haddr:haddr + 4] == b"HEAP"
        return struct.unpack_from("<Q", self.b, haddr + 24)[0]

    def _children(self, o):
        b = self.b
        names = {}
        hdata = self._heap_data(o["heap"])

        def walk(node):
            assert b[node:node + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            p = node + 24
            for k in range(used):
                (child,) = struct.unpack_from("<Q", b, p + 8 + 16 * k)
                if level > 0:
                    walk(child)
                else:
                    assert b[child:child + 4] == b"SNOD"
                    (nsym,) = struct.unpack_from("<H", b, child + 6)
                    for s in range(nsym):
                        e = child + 8 + 40 * s
                        noff, oaddr = struct.unpack_from("<QQ", b, e)
                        nm = b[hdata + noff:].split(b"\0", 1)[0].decode()
                        names[nm] = oaddr

        walk(o["btree"])
        return names

    def keys(self, path="/"):
        return sorted(self._children(self._resolve(path)))

    def _resolve(self, path):
        o = self.root
        for part in [p for p in path.split("/") if p]:
            o = self._obj(self._children(o)[part])
        return o

    def attrs(self, path):
        return self._resolve(path)["attrs"]

    def read(self, path):
        o = self._resolve(path)
        assert not o.get("filtered"), "filtered datasets unsupported"
        shape, dt, lay = o["shape"], o["dtype"], o["layout"]
        n = int(np.prod(shape)) if shape else 1
        if lay[0] in ("contiguous", "compact"):
            val = self._decode(dt, lay[1], n)
        else:
            val = self._read_chunked(dt, lay, shape)
        if isinstance(val, np.ndarray):
            val = val.reshape(shape)
        return val

    def _read_chunked(self, dt, lay, shape):
        assert len(shape) == 1, "only 1-D chunked datasets supported"
        b = self.b
        _, addr, dims = lay
        csz = dims[0]
        parts = {}

        def walk(node):
            assert b[node:node + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            ksz = 8 + 8 * (len(shape) + 1)
            p = node + 24
            for k in range(used):
                kp = p + k * (ksz + 8)
                nbytes, _mask = struct.unpack_from("<II", b, kp)
                (off0,) = struct.unpack_from("<Q", b, kp + 8)
                (child,) = struct.unpack_from("<Q", b, kp + ksz)
                if level > 0:
                    walk(child)
                else:
                    parts[off0] = child

        walk(addr)
        out = []
        for off0 in sorted(parts):
            cnt = min(csz, shape[0] - off0)
            out.append((off0, self._decode(dt, parts[off0], cnt)))
        if isinstance(out[0][1], np.ndarray):
            return np.concatenate([v for _, v in out])
        return [x for _, v in out for x in v]


def read_h5ad_minimal(path):
    """Return (X_dense float32 [n_obs, n_var], time float64 [n_obs], obs_names, var_names)."""
    f = H5Lite(path)
    shape = tuple(int(v) for v in f.attrs("X")["shape"])
    data = f.read("X/data")
    indices = f.read("X/indices")
    indptr = f.read("X/indptr")
    enc = f.attrs("X").get("encoding-type", "csr_matrix")
    X = np.zeros(shape, dtype=data.dtype)
    if enc == "csr_matrix":
        for r in range(shape[0]):
            X[r, indices[indptr[r]:indptr[r + 1]]] = data[indptr[r]:indptr[r + 1]]
    else:
        for c in range(shape[1]):
            X[indices[indptr[c]:indptr[c + 1]], c] = data[indptr[c]:indptr[c + 1]]
    time = np.asarray(f.read("obs/time"), dtype=np.float64)
    obs_names = f.read("obs/" + f.attrs("obs").get("_index", "_index"))
    var_names = f.read("var/" + f.attrs("var").get("_index", "_index"))
    return X, time, list(obs_names), list(var_names)
