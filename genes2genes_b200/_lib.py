"""ctypes binding of ``libg2g.so`` (C ABI: ``include/g2g.h``).

The library is the product: there is no CPU fallback.  Importing this module without the built
shared object raises ``ImportError`` telling the user how to build it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libg2g.so")

c_i32, c_i64, c_f64, c_vp, c_sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t


class G2GError(RuntimeError):
    pass


class Layout(ctypes.Structure):
    _fields_ = [("n_groups", c_i32), ("bins_per_group", c_i32), ("row_floats", c_i32),
                ("genes_per_tile", c_i32), ("n_pad", c_i64), ("cells_per_item", c_i64), ("n_split", c_i64)]


class Side(ctypes.Structure):
    _fields_ = [("X", c_vp), ("n_cells", c_i64), ("ldx", c_i64), ("wtab", c_vp), ("pilot", c_vp),
                ("part_f64", c_vp), ("part_i32", c_vp), ("n_split", c_i64)]


class MmaLayout(ctypes.Structure):
    _fields_ = [("n_groups", c_i32), ("bins_per_group", c_i32), ("n_cols", c_i32), ("reserved", c_i32),
                ("n_pad", c_i64), ("cells_per_item", c_i64), ("n_split", c_i64), ("table_bytes", c_i64)]


class SparseSide(ctypes.Structure):
    _fields_ = [("col_ptr", c_vp), ("row_idx", c_vp), ("values", c_vp), ("n_cells", c_i64), ("wtab", c_vp),
                ("pilot", c_vp), ("part_f64", c_vp), ("part_i32", c_vp)]


# name -> (restype, argtypes); kept in one table so tests can check it against include/g2g.h
SIGNATURES = {
    "g2g_version": (c_i32, []),
    "g2g_last_error": (ctypes.c_char_p, []),
    "g2g_check_device": (c_i32, []),
    "g2g_layout": (c_i32, [c_i64, c_i32, ctypes.POINTER(Layout)]),
    "g2g_pack_rows": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "g2g_pack_csr": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "g2g_csr_to_csc_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "g2g_csr_to_csc": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "g2g_peer_window_create": (c_i32, [c_sz, ctypes.POINTER(c_vp), c_vp]),
    "g2g_peer_window_open": (c_i32, [c_vp, ctypes.POINTER(c_vp)]),
    "g2g_peer_window_close": (c_i32, [c_vp, c_i32]),
    "g2g_peer_push": (c_i32, [c_vp, c_vp, c_sz, c_vp]),
    "g2g_upload_block": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "g2g_scatter_rows": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "g2g_weights_workspace_bytes": (c_sz, [c_i64, c_i32]),
    "g2g_weights": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "g2g_pilot_mean": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "g2g_moments": (c_i32, [ctypes.POINTER(Side), c_i32, c_i64, c_i64, c_i32, c_vp]),
    "g2g_mma_layout": (c_i32, [c_i64, c_i32, ctypes.POINTER(MmaLayout)]),
    "g2g_weights_mma": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "g2g_moments_mma": (c_i32, [ctypes.POINTER(Side), c_i32, c_i64, c_i64, c_i32, c_vp, c_vp]),
    "g2g_moments_mma_ex": (c_i32, [ctypes.POINTER(Side), c_i32, c_i64, c_i64, c_i32, c_vp, c_i32, c_vp]),
    "g2g_moments_csc": (c_i32, [ctypes.POINTER(SparseSide), c_i32, c_i64, c_i64, c_i32, c_vp]),
    "g2g_fixup_workspace_bytes": (c_sz, [c_i64, c_i32]),
    "g2g_fixup_trends": (c_i32, [ctypes.POINTER(Side), c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32,
                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "g2g_cost_dp": (c_i32, [c_vp, c_i64, c_i32, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_f64, c_i32,
                            c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "g2g_fixup_cost_dp": (c_i32, [ctypes.POINTER(Side), c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32,
                                  ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_f64, c_i32,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "g2g_state_histogram": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "g2g_match_counts": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "g2g_strdist_workspace_bytes": (c_sz, [c_i64]),
    "g2g_levenshtein_matrix": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_sz, c_vp]),
    "g2g_hamming_matrix": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_sz, c_vp]),
    "g2g_de_statistics": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
}

_lib = None


def load():
    """Load ``libg2g.so`` once; raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "genes2genes_b200: %s is missing. Build it with `python -c \"import __graft_entry__ as g; "
            "g.build()\"` at the repository root (needs nvcc; sm_100a only, no CPU fallback)." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().g2g_last_error()
        raise G2GError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def mma_layout(n_cells, n_bins):
    out = MmaLayout()
    check(load().g2g_mma_layout(int(n_cells), int(n_bins), ctypes.byref(out)), "g2g_mma_layout")
    return out


def layout(n_cells, n_bins):
    out = Layout()
    check(load().g2g_layout(int(n_cells), int(n_bins), ctypes.byref(out)), "g2g_layout")
    return out
