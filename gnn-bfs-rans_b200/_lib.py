"""ctypes binding of libb2g.so (the C ABI declared in include/b2g.h).

The product path has NO fallback: if the shared library is missing or a call returns an error the
caller gets a RuntimeError (the reference re-wraps RuntimeError with layer context at
/root/reference/gnn_model.py:173-181)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2G_LIB") or os.path.join(_HERE, "libb2g.so")   # B2G_LIB: experiment builds only

i32, i64, u64, f32, vp, cp = C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_void_p, C.c_char_p

# name -> (restype, argtypes); must list every symbol include/b2g.h declares
SIGNATURES = {
    "b2g_version": (i32, []),
    "b2g_error_string": (cp, [i32]),
    "b2g_launch_count": (i64, []),
    "b2g_launch_count_reset": (None, []),
    "b2g_dropout_epoch_advance": (i32, [vp]),
    "b2g_dropout_epoch_set": (i32, [u64, vp]),
    "b2g_build_edge_index": (i32, [vp, vp, i64, i64, vp, vp]),
    "b2g_mask_to_map_workspace_bytes": (i64, [i64]),
    "b2g_mask_to_map": (i32, [vp, i64, vp, vp, vp, vp]),
    "b2g_build_graph_workspace_bytes": (i64, [i64, i64, i64]),
    "b2g_build_graph_count": (i32, [vp, vp, i64, i64, i32, vp, i64, i64, vp, vp, vp]),
    "b2g_build_graph_fill": (i32, [vp, vp, i64, i64, i32, vp, i64, i64, vp, i64, vp, vp]),
    "b2g_edge_attr": (i32, [vp, i64, vp, i64, vp, vp]),
    "b2g_csr_workspace_bytes": (i64, [i64, i64]),
    "b2g_csr_count": (i32, [vp, i64, i64, i32, i32, vp, vp, vp, vp]),
    "b2g_csr_fill": (i32, [vp, i64, i64, i32, i32, vp, i64, vp, vp, vp, vp, vp]),
    "b2g_csr_perm": (i32, [vp, vp, i64, vp, vp, vp]),
    "b2g_seg_sum": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp, vp, f32, vp, i32, vp]),
    "b2g_seg_sum_banded": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp, vp, f32, vp, i32, i64, vp]),
    "b2g_seg_sum_hinted": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp, vp, f32, vp, i32, i64, i64, vp]),
    "b2g_seg_sum_tuned": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp, vp, f32, vp, i32, i64, i64, i32, i32, i32, vp]),
    "b2g_gat_fwd": (i32, [vp, i64, vp, vp, i64, vp, i64, i64, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, f32, u64, i32, vp]),
    "b2g_gat_bwd_dst": (i32, [vp, i64, vp, vp, i64, vp, i64, i64, i32, i32, i32, i32, f32, vp, vp, vp, vp, f32, u64, vp, vp, vp, i64, vp]),
    "b2g_gat_bwd_src": (i32, [vp, i64, vp, vp, vp, i64, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, vp]),
    "b2g_gatz_supported": (i32, [i64, i32, i32, i32]),
    "b2g_gatw_gemm_supported": (i32, [i64, i32, i32, i32, i32]),
    "b2g_segw_gemm_supported": (i32, [i64, i32, i32, i32]),
    "b2g_segw_gemm": (i32, [vp, i64, vp, vp, vp, vp, f32, vp, i64, vp, i32, vp, i64, i64, i32, i32, i32, i64, vp]),
    "b2g_edge_rows_sl": (i32, [vp, i64, vp, vp, i64, vp, vp]),
    "b2g_gatw_gemm_ex": (i32, [vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, i64, i64, i32, i32, i32, i32, i64, vp]),
    "b2g_tz_alpha": (i32, [vp, i64, vp, i64, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, f32, u64, i64, i32, vp]),
    "b2g_batch_finalize": (i32, [vp, i64, vp, vp, i32, vp, i64, vp]),
    "b2g_wmse_workspace_bytes": (i64, []),
    "b2g_wmse_fwd": (i32, [vp, i64, vp, i64, i64, i32, vp, f32, vp, vp, vp, vp]),
    "b2g_wmse_bwd": (i32, [vp, i64, vp, i64, i64, i32, vp, vp, vp, i64, vp]),
    "b2g_adam_workspace_bytes": (i64, []),
    "b2g_clip_adam_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, vp, vp, vp]),
    "b2g_gat_alpha": (i32, [vp, i64, vp, vp, vp, i64, i64, i32, f32, f32, u64, vp, vp, vp, vp]),
    "b2g_linear_fwd_masked": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, vp, vp]),
    "b2g_gatw_gemm_sm": (i32, [vp, i64, vp, vp, vp, vp, i64, vp, f32, f32, u64, vp, vp, vp, i64, vp, vp, i64, i64, i64, i32, i32, i32, i32,
                                i64, vp]),
    "b2g_gatw_gemm": (i32, [vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, i64, i64, i32, i32, i32, i32, i64, vp]),
    "b2g_rowdot8": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp]),
    "b2g_gatz_fwd": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, f32, vp, vp, vp, vp, f32, u64, i64, vp]),
    "b2g_gatz_bwd_dst": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, f32, vp, vp, vp, vp, f32, u64, vp, vp, vp, i64, i32, vp, i64, i64, vp]),
    "b2g_gatz_bwd_src": (i32, [vp, i64, vp, vp, vp, i64, vp, i64, i32, i64, i32, i32, i32, vp, vp, vp, i64, vp]),
    "b2g_tz_fwd": (i32, [vp, i64, vp, vp, i64, vp, i64, i64, i32, i32, i32, vp, vp, vp, vp, f32, u64, i64, vp]),
    "b2g_tz_bwd_dst": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, vp, f32, u64, vp, vp, vp, i64, vp, i64, i64, vp]),
    "b2g_edge_dot4": (i32, [vp, i64, vp, vp, i64, i32, vp, vp]),
    "b2g_edge_wsum4": (i32, [vp, vp, vp, i64, i32, f32, u64, vp, i64, i32, vp]),
    "b2g_tconv_fwd": (i32, [vp, vp, vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, vp, f32, u64, vp]),
    "b2g_tconv_bwd_dst": (i32, [vp, vp, vp, i64, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, vp, f32, u64, vp, vp, vp, i64, vp]),
    "b2g_tconv_bwd_src": (i32, [vp, i64, vp, i64, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, vp]),
    "b2g_linear_workspace_bytes": (i64, [i64, i32, i32, i32, i32]),
    "b2g_linear_impl": (i32, [i64, i32, i32, i32, i32]),
    "b2g_linear_fwd": (i32, [vp, i64, vp, i64, vp, vp, vp, i64, vp, i64, i64, i32, i32, i32, i32, i32, i32, vp, vp]),
    "b2g_linear_dgrad": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, i32, i32, vp, vp]),
    "b2g_linear_wgrad": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp, vp]),
    "b2g_colsum_workspace_bytes": (i64, [i32]),
    "b2g_colsum": (i32, [vp, i64, i64, i32, i32, vp, vp, vp]),
    "b2g_mesh_num_cells": (i32, [vp, i64, vp, i64, vp, vp, vp]),
    "b2g_mesh_workspace_bytes": (i64, [i64, i64]),
    "b2g_mesh_cell_centers": (i32, [vp, i64, vp, i64, vp, i64, vp, vp, i64, i64, i64, vp, vp, vp, i64, vp]),
    "b2g_mesh_internal_cells": (i32, [vp, i64, vp, i64, i64, vp, vp, vp, vp]),
    "b2g_bn_workspace_bytes": (i64, [i32]),
    "b2g_bn_stats": (i32, [vp, i64, vp, i64, i64, i32, i32, f32, vp, vp, vp]),
    "b2g_bn_apply": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp, vp, i32, f32, u64, vp]),
    "b2g_bn_bwd_stats": (i32, [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, i32, f32, vp, vp, vp]),
    "b2g_bn_bwd_apply": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp, vp, i32, f32, i32, vp]),
    "b2g_rows_gather": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, vp]),
    "b2g_rows_scatter_add": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, vp]),
}

_lib = None


def load():
    """Load libb2g.so once; raise RuntimeError (never fall back) when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"b2g: CUDA extension {LIB_PATH} not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gnn-bfs-rans_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"b2g: {LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.b2g_version() != 100:
        raise RuntimeError("b2g: libb2g.so version mismatch; rebuild it")
    _lib = lib
    return lib


E_UNSUPPORTED = -5          # include/b2g.h B2G_E_UNSUPPORTED: "this build has no kernel for the request"


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b2g_error_string(int(rc)).decode()
        raise RuntimeError(f"b2g {what}: {msg} (code {rc})")


def launch_count() -> int:
    return int(load().b2g_launch_count())


def launch_count_reset() -> None:
    load().b2g_launch_count_reset()
