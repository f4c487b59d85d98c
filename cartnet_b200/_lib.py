"""ctypes binding of libcartnet_b200.so (the C ABI declared in include/cartnet_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised. Build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C cartnet_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcartnet_b200.so")

PREC_FP32, PREC_BF16, PREC_TF32, PREC_BF16X3 = 0, 1, 2, 3
NT_PAIR_MIN_ROWS = 32768      # include/cartnet_b200.h: CARTNET_NT_PAIR_MIN_ROWS
ACT_NONE, ACT_SILU, ACT_MUL_DSILU = 0, 1, 2

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmDesc(C.Structure):
    """Mirror of `cartnet_gemm_t`."""
    _fields_ = [
        ("prec", i32), ("M", i32), ("N", i32), ("K", i32),
        ("A", vp), ("lda", i64), ("B", vp), ("ldb", i64),
        ("bias", vp),
        ("gather0", vp), ("gidx0", vp), ("gather1", vp), ("gidx1", vp), ("ldg", i64),
        ("z_out", vp), ("ldz", i64),
        ("act", i32), ("_pad", i32),
        ("z_in", vp), ("ldzin", i64),
        ("resid", vp), ("ldr", i64),
        ("out_f32", vp), ("ldo", i64),
        ("out_t", vp), ("ldt", i64),
    ]


class CollateField(C.Structure):
    """Mirror of `cartnet_collate_field_t`."""
    _fields_ = [("src", vp), ("dst", vp), ("kind", i32), ("op", i32), ("elem_bytes", i32), ("_pad", i32)]


class LayerDesc(C.Structure):
    """Mirror of `cartnet_layer_t` (field order must match include/cartnet_b200.h)."""
    _PTRS1 = ["src32", "dst32", "row_ptr", "col_ptr", "perm_src", "dist", "x", "e", "x_t", "e_t",
              "G1", "A1", "bg1", "ba1", "G2", "A2", "bg2", "ba2", "bn1_w", "bn1_b", "bn2_w", "bn2_b",
              "bn1_rm", "bn1_rv", "bn2_rm", "bn2_rv",
              "W1n_t", "W1e_t", "G2_t", "A2_t", "W1nT_t", "W1eT_t", "G2T_t", "A2T_t", "b1",
              "P", "Z", "H", "g_t", "center", "bias_c", "hsum", "m", "s_t", "gn_t", "mean1", "var1", "mean2", "var2", "x_out", "e_out", "x_out_t", "e_out_t",
              "dx_out", "de_out", "dm", "ds_t", "dg_t", "dghat_t", "dZ", "dP", "sums1", "sums2", "dx_in", "de_in",
              "dG1", "dA1", "dbg1", "dba1", "dG2", "dA2", "dbg2", "dba2", "dbn1_w", "dbn1_b", "dbn2_w", "dbn2_b",
              "partial", "splitk"]
    _fields_ = ([("prec", i32), ("training", i32), ("use_envelope", i32), ("D", i32), ("num_nodes", i32), ("_pad0", i32),
                 ("num_edges", i64), ("radius", f32), ("eps", f32), ("momentum1", f32), ("momentum2", f32)]
                + [(n, vp) for n in _PTRS1] + [("splitk_bytes", i64)])


# name -> (restype, argtypes); every function listed here must be exported by the .so and
# declared in include/cartnet_b200.h (tests/test_abi.py checks both directions).
SIGNATURES = {
    "cartnet_version": (i32, []),
    "cartnet_last_error": (C.c_char_p, []),
    "cartnet_device_ok": (i32, [i32]),
    "cartnet_launch_count": (i64, []),
    "cartnet_nlist_reps": (i32, [vp, i32, f32, i32, vp, vp, vp]),
    "cartnet_nlist_count": (i32, [vp, vp, vp, vp, i32, f32, f32, vp, i32, vp, vp]),
    "cartnet_exclusive_scan_i32": (i32, [vp, i32, vp, vp]),
    "cartnet_nlist_fill": (i32, [vp, vp, vp, vp, i32, f32, f32, vp, i32, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "cartnet_nlist_cells_workspace": (i64, [i32, i32]),
    "cartnet_nlist_cells_build": (i32, [vp, vp, vp, vp, i32, i32, f32, vp, i32, vp, vp]),
    "cartnet_nlist_cells_count": (i32, [vp, vp, vp, vp, i32, i32, f32, f32, vp, i32, vp, vp, vp]),
    "cartnet_nlist_cells_fill": (i32, [vp, vp, vp, vp, i32, i32, f32, f32, vp, i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "cartnet_nlist_knn_mask": (i32, [vp, vp, i32, i32, f32, i32, vp, vp, vp, vp]),
    "cartnet_graph_split": (i32, [vp, i64, i32, vp, vp, vp, vp]),
    "cartnet_graph_csr": (i32, [vp, i64, i32, vp, vp, vp, vp]),
    "cartnet_edge_features": (i32, [vp, vp, vp, vp, i32, f32, i32, i64, vp, i32, i32, vp]),
    "cartnet_gemm": (i32, [C.POINTER(GemmDesc), vp]),
    "cartnet_gemm_colstats": (i32, [C.POINTER(GemmDesc), vp, vp, vp, vp, vp, f32, vp, vp]),
    "cartnet_gemm_tn_workspace": (i64, [i32, i32, i32, i64]),
    "cartnet_gemm_tn": (i32, [i32, i32, i32, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp]),
    "cartnet_gemm_tn_blocks": (i32, [i32, i32, i32, i64, vp, i64, vp, i64, C.POINTER(vp), i32, i64, vp, i64, vp]),
    "cartnet_colstats_workspace": (i64, [i32]),
    "cartnet_colstats": (i32, [vp, i32, i32, i64, i32, i64, vp, vp, vp, vp, vp, f32, vp, vp]),
    "cartnet_gate_center": (i32, [vp, i64, i64, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]),
    "cartnet_colsum": (i32, [vp, i32, i32, i64, i32, i64, vp, vp, vp]),
    "cartnet_edge_gate_aggregate": (i32, [vp, vp, vp, vp, vp, i32, i64, i32, vp, vp, vp, vp, f32, f32, i32, vp, vp, vp, i32, vp, vp]),
    "cartnet_node_update": (i32, [vp, vp, i32, i32, vp, vp, vp, vp, f32, vp, vp, i32, vp]),
    "cartnet_node_update_bwd_reduce": (i32, [vp, vp, i32, i32, vp, vp, vp, vp, f32, vp, vp, vp]),
    "cartnet_node_update_bwd_apply": (i32, [vp, vp, i32, i32, vp, vp, vp, vp, f32, vp, i32, vp, vp]),
    "cartnet_edge_gate_bwd_reduce": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, f32, i32, vp, vp, i32, vp, vp, vp, vp, f32, vp]),
    "cartnet_edge_gate_bwd_apply": (i32, [vp, vp, i64, i32, vp, vp, f32, vp, i32, vp, i32, vp, i32, vp]),
    "cartnet_segment_sum": (i32, [vp, i64, vp, vp, i32, i32, vp, i64, i32, i32, vp]),
    "cartnet_segment_sum_pair": (i32, [vp, i64, vp, vp, vp, i32, i32, vp, i64, i32, i32, vp]),
    "cartnet_dsilu_mul": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp]),
    "cartnet_cast_rows": (i32, [vp, i64, vp, i64, i64, i32, i32, vp]),
    "cartnet_uncast_rows": (i32, [vp, i64, vp, i64, i64, i32, i32, vp]),
    "cartnet_layer_splitk_bytes": (i64, [i32, i32, i32, i64]),
    "cartnet_layer_pack_weights": (i32, [C.POINTER(LayerDesc), vp]),
    "cartnet_layer_fwd": (i32, [C.POINTER(LayerDesc), vp]),
    "cartnet_layer_bwd": (i32, [C.POINTER(LayerDesc), vp]),
    "cartnet_cholesky_head_workspace": (i64, [i32, i32]),
    "cartnet_cholesky_head_fwd": (i32, [vp, i64, vp, vp, i32, i32, vp, vp, vp]),
    "cartnet_collate": (i32, [C.POINTER(CollateField), i32, vp, i32, vp, i64, vp, C.POINTER(i64), vp]),
    "cartnet_collate_close_csr": (i32, [vp, vp, i64, i64, vp]),
    "cartnet_loss_l1_mse": (i32, [vp, vp, i64, vp, vp]),
    "cartnet_loss_l1_mse_bwd": (i32, [vp, vp, i64, vp, vp, vp, vp]),
    "cartnet_cholesky_head_bwd": (i32, [vp, vp, i64, vp, vp, i32, i32, vp, i64, vp, vp, vp, vp]),
}

_lib = None


def load():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "cartnet_b200: %s not found -- the CUDA library has not been built "
            "(run __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().cartnet_last_error()
        raise RuntimeError("cartnet_b200.%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))
