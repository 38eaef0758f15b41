"""ctypes binding of the C ABI in include/feprogram_b200.h (libfeprogram_b200.so, built in-tree).

There is no CPU fallback: if the library cannot be loaded every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfeprogram_b200.so")

FE_OK = 0
FE_QUAD4, FE_QUAD9, FE_TRI3 = 4, 9, 3
FE_OP_REFERENCE, FE_OP_WEIGHTED, FE_OP_MASS, FE_OP_POISSON = 0, 1, 2, 3
FE_BC_ROWS, FE_BC_SYMMETRIC = 0, 1

_p = C.c_void_p
_i = C.c_int
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/feprogram_b200.h one to one
PROTOTYPES = {
    "fe_last_error": (None, [C.c_char_p, _sz]),
    "fe_version": (C.c_char_p, []),
    "fe_limits": (None, [C.POINTER(_i32)]),
    "fe_peek": (_i, [_p, _p, _i64, _p]),
    "fe_tables": (_i, [_i, _p, _p, _p, C.POINTER(_i32), C.POINTER(_i32)]),
    "fe_tags_to_conn": (_i, [_i, _i64, _p, _p, _p]),
    "fe_element_matrices": (_i, [_i, _i, _i64, _p, _p, _p, _p]),
    "fe_scan_workspace_bytes": (_sz, [_i64]),
    "fe_incidence_count": (_i, [_i, _i64, _i64, _p, _p, _p, _p]),
    "fe_incidence_fill": (_i, [_i, _i64, _i64, _p, _p, _p, _p, _p]),
    "fe_adjacency_count": (_i, [_i, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "fe_adjacency_fill": (_i, [_i, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "fe_csr_expand": (_i, [_i, _i64, _p, _p, _i, _p, _p, _p]),
    "fe_assemble": (_i, [_i, _i, _i64, _i64, _p, _p, _p, _p, _p, _p, _i32, _p, _p]),
    "fe_lattice_keys": (_i, [_i, _i64, _p, _p, _p, _p]),
    "fe_bbox": (_i, [_i64, _p, _p, _p, _p]),
    "fe_tile_form_count": (_i, [_i, _i64, _i64] + [_p] * 10),
    "fe_tile_form_sort": (_i, [_i, _i64] + [_p] * 8),
    "fe_tile_params": (None, [_i, C.POINTER(_i32)]),
    "fe_exclusive_scan_i32": (_i, [_p, _p, _i64, _p, _p]),
    "fe_morton_keys": (_i, [_i, _i64, _p, _p, _f64, _f64, _f64, _f64, _p, _p]),
    "fe_invert_permutation": (_i, [_i64, _p, _p, _p]),
    "fe_tile_owner": (_i, [_i64, _i64, _p, _p, _p, _p, _p, _p]),
    "fe_tile_layout": (_i, [_i, _i64, _i64] + [_p] * 13),
    "fe_tile_records": (_i, [_i, _i64, _p, _p, _p, _i64] + [_p] * 14),
    "fe_tile_conn": (_i, [_i, _i64] + [_p] * 7),
    "fe_assemble_tiled": (_i, [_i, _i, _i64] + [_p] * 9),
    "fe_tile_ws_params": (None, [_i, C.POINTER(_i32)]),
    "fe_tile_ws_build": (_i, [_i, _i64, _i64] + [_p] * 18),
    "fe_assemble_ws": (_i, [_i, _i, _i64] + [_p] * 7),
    "fe_bc_rhs": (_i, [_i64, _p, _i64, _p, _i64, _f64, _p, _p, _p]),
    "fe_apply_bc": (_i, [_i, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "fe_spmv": (_i, [_i64, _i64, _i, _p, _p, _p, _p, _p, _p, _p]),
    "fe_csr_diagonal": (_i, [_i64, _i, _p, _p, _p, _p, _p]),
    "fe_pcg_workspace_bytes": (_sz, [_i64]),
    "fe_pcg": (_i, [_i64, _i64, _i, _p, _p, _p, _p, _p, _p, _f64, _f64, _i64, _i32, _p, _p, _p, _p, _p]),
    "fe_spmv_nodal": (_i, [_i64, _p, _p, _p, _p, _p, _p, _p]),
    "fe_dist_available": (_i, []),
    "fe_dist_unique_id": (_i, [_p]),
    "fe_dist_comm_init": (_i, [_p, _i, _i, C.POINTER(C.c_void_p)]),
    "fe_dist_comm_destroy": (_i, [_p]),
    "fe_dist_halo_exchange": (_i, [_p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "fe_edge_load": (_i, [_i, _i64, _p, _p, _f64, _f64, _p, _p, _p]),
    "fe_gauss_gradients": (_i, [_i, _i64, _p, _p, _p, _p, _p]),
    "fe_nodal_strain": (_i, [_i, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "fe_dist_pcg_workspace_bytes": (_sz, [_i64]),
    "fe_dist_pcg": (_i, [_i64, _i64, _i64, _i, _p, _p, _p, _p, _p, _p, _f64, _f64, _i64, _i32, _p, _p, _p, _i32,
                         _p, _p, _p, _p, _p, _p, _p, _p]),
    "fe_dist_pcg_p2p": (_i, [_i64, _i64, _i, _p, _p, _p, _p, _p, _p, _f64, _f64, _i64, _i32, _p, _p, _p, _i32,
                             _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p, _p, _i64, _p, _i64, _p, _p, _i64, _p]),
}

_lib = None


class FeError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__("feprogram_b200 error %d: %s" % (code, message))
        self.code = code


def lib() -> C.CDLL:
    """Loads the CUDA library (once).  Raises if it has not been built -- no fallback path exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "feprogram_b200: %s is missing. Build it with `python -m feprogram_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().fe_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(code: int) -> None:
    if code != FE_OK:
        raise FeError(code, last_error())


def version() -> str:
    return lib().fe_version().decode()


def tile_params(elem_type: int):
    out = (_i32 * 6)()
    lib().fe_tile_params(elem_type, out)
    return dict(tile_elems=out[0], halo_cap=out[1], max_nodes=out[2], max_items=out[3], max_work=out[4],
                tile_rows=out[5])


def tile_ws_params(elem_type: int):
    out = (_i32 * 4)()
    lib().fe_tile_ws_params(elem_type, out)
    return dict(blob_cap=out[0], stage_cap=out[1], max_deg=out[2], max_nodes=out[3])


def limits():
    out = (_i32 * 4)()
    lib().fe_limits(out)
    return dict(max_adj=out[0], max_inc=out[1], max_nnpe=out[2], max_gauss=out[3])
