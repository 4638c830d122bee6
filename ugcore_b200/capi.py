"""ctypes prototypes of include/ug4b200.h (kernel-level ABI, ``dev``) and
include/ug4b200_solver.h (descriptor-level ABI, ``host``)."""
from __future__ import annotations

import ctypes as C
import os

_LIBDIR = os.environ.get("UG4B200_LIBDIR", os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib"))
DEV_SO = os.path.join(_LIBDIR, "libug4b200.so")
HOST_SO = os.path.join(_LIBDIR, "libug4b200_host.so")

for _p in (DEV_SO, HOST_SO):
    if not os.path.exists(_p):
        raise ImportError(
            f"{_p} is missing: build the native libraries first "
            "(python -c 'import __graft_entry__ as g; g.build()'); ugcore_b200 has no fallback path")

# RTLD_GLOBAL so that libug4b200_host.so resolves against the already loaded kernel library
dev = C.CDLL(DEV_SO, mode=os.RTLD_GLOBAL | os.RTLD_NOW)
host = C.CDLL(HOST_SO, mode=os.RTLD_GLOBAL | os.RTLD_NOW)

c_i64, c_int, c_dbl, c_vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
p_i64, p_int, p_dbl = C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double)


class Coef(C.Structure):
    _fields_ = [("dev", c_vp), ("host", c_dbl)]


class Fin(C.Structure):
    _fields_ = [("op", c_int), ("out", c_vp), ("out2", c_vp), ("a", c_vp), ("conv", c_vp)]


class ConvState(C.Structure):
    _fields_ = [("initial_defect", c_dbl), ("current_defect", c_dbl), ("last_defect", c_dbl),
                ("min_defect", c_dbl), ("rel_reduction", c_dbl), ("step", c_int), ("max_steps", c_int),
                ("done", c_int), ("status", c_int), ("history_cap", c_int), ("pad_", c_int), ("history", c_vp)]


class MatrixInfo(C.Structure):
    _fields_ = [("nrows", c_i64), ("ncols", c_i64), ("nnz", c_i64), ("padded_nnz", c_i64), ("num_slices", c_i64),
                ("device_bytes", c_i64), ("block", c_int), ("max_row_len", c_int), ("value_indexed", c_int),
                ("num_distinct_values", c_int), ("x_staged", c_int), ("x_staged_runs", c_int), ("x_staged_doubles", c_i64)]


class StreamPlan(C.Structure):
    _fields_ = [("num_slices", c_i64), ("padded_nnz", c_i64), ("max_row_len", c_int), ("num_distinct_values", c_int),
                ("max_column_window", c_i64), ("value_indexed", c_int), ("x_staged", c_int), ("x_staged_runs", c_int),
                ("x_staged_max_doubles", c_int), ("x_staged_doubles", c_i64)]


class SolverDesc(C.Structure):
    _fields_ = [("block", c_int), ("solver", c_int), ("precond", c_int), ("damp", c_dbl),
                ("max_steps", c_int), ("min_defect", c_dbl), ("rel_reduction", c_dbl),
                ("base_lev", c_int), ("top_lev", c_int), ("cycle", c_int), ("nu1", c_int), ("nu2", c_int),
                ("smoother", c_int), ("smoother_damp", c_dbl), ("base_solver", c_int), ("base_max_steps", c_int),
                ("base_min_defect", c_dbl), ("base_rel_reduction", c_dbl), ("flags", c_int), ("gather_lev", c_int),
                ("restart", c_int), ("ilu_beta", c_dbl), ("ilu_order", c_int)]


FIN_STORE, FIN_A_DIV_R, FIN_R_DIV_A, FIN_SQRT, FIN_CONV_START, FIN_CONV_UPDATE = range(6)
SMOOTH_ADD_IN, SMOOTH_JACOBI, SMOOTH_ADD_OUT, SMOOTH_SC_ZERO = 1, 2, 4, 8
MAT_DEFAULT, MAT_NO_COMPRESS, MAT_NO_XSTAGE = 0, 1, 2
FLAG_HOST_SCALARS, FLAG_NO_GRAPH, FLAG_NO_FUSED_JACOBI, FLAG_FINAL_LEVEL_DEFECT, FLAG_RAP, FLAG_DEVICE_BICGSTAB, FLAG_DEVICE_LINEAR = 1, 2, 4, 8, 16, 32, 64

# name -> (restype, argtypes); every symbol declared in include/ug4b200.h
DEV_API = {
    "ug4b200_ctx_create": (c_int, [c_int, c_vp, C.POINTER(c_vp)]),
    "ug4b200_ctx_destroy": (c_int, [c_vp]),
    "ug4b200_last_error": (C.c_char_p, [c_vp]),
    "ug4b200_sync": (c_int, [c_vp]),
    "ug4b200_stream": (c_vp, [c_vp]),
    "ug4b200_launch_count": (c_int, [c_vp, p_i64]),
    "ug4b200_set_guard": (c_int, [c_vp, c_vp]),
    "ug4b200_batch_enable": (c_int, [c_vp, c_int, c_i64]),
    "ug4b200_batch_flush": (c_int, [c_vp]),
    "ug4b200_batch_stats": (c_int, [c_vp, p_i64, C.POINTER(c_int)]),
    "ug4b200_graph_begin": (c_int, [c_vp]),
    "ug4b200_graph_end": (c_int, [c_vp, C.POINTER(c_vp)]),
    "ug4b200_graph_launch": (c_int, [c_vp, c_vp]),
    "ug4b200_graph_destroy": (c_int, [c_vp, c_vp]),
    "ug4b200_event_create": (c_int, [c_vp, C.POINTER(c_vp)]),
    "ug4b200_event_record": (c_int, [c_vp, c_vp]),
    "ug4b200_event_sync": (c_int, [c_vp, c_vp]),
    "ug4b200_event_elapsed_ms": (c_int, [c_vp, c_vp, c_vp, C.POINTER(C.c_float)]),
    "ug4b200_event_destroy": (c_int, [c_vp, c_vp]),
    "ug4b200_alloc": (c_int, [c_vp, C.c_size_t, C.POINTER(c_vp)]),
    "ug4b200_free": (c_int, [c_vp, c_vp]),
    "ug4b200_h2d": (c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "ug4b200_d2h": (c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "ug4b200_d2h_async": (c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "ug4b200_d2d": (c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "ug4b200_memset": (c_int, [c_vp, c_int, C.c_size_t]),
    "ug4b200_host_alloc": (c_int, [c_vp, C.c_size_t, C.POINTER(c_vp)]),
    "ug4b200_host_free": (c_int, [c_vp, c_vp]),
    "ug4b200_vec_set": (c_int, [c_vp, c_i64, c_vp, c_dbl]),
    "ug4b200_vec_copy": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "ug4b200_vec_scale": (c_int, [c_vp, c_i64, c_vp, c_dbl]),
    "ug4b200_vec_add": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "ug4b200_vec_sub": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "ug4b200_vec_scale_add2": (c_int, [c_vp, c_i64, c_vp, c_dbl, c_vp, c_dbl, c_vp]),
    "ug4b200_vec_scale_add3": (c_int, [c_vp, c_i64, c_vp, c_dbl, c_vp, c_dbl, c_vp, c_dbl, c_vp]),
    "ug4b200_vec_dot": (c_int, [c_vp, c_i64, c_vp, c_vp, p_dbl]),
    "ug4b200_vec_norm": (c_int, [c_vp, c_i64, c_vp, p_dbl]),
    "ug4b200_vec_gather": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "ug4b200_vec_scatter": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "ug4b200_vec_scatter_add": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "ug4b200_vec_scale_add2_ds": (c_int, [c_vp, c_i64, c_vp, Coef, c_vp, Coef, c_vp]),
    "ug4b200_vec_scale_add3_ds": (c_int, [c_vp, c_i64, c_vp, Coef, c_vp, Coef, c_vp, Coef, c_vp]),
    "ug4b200_vec_dot_ds": (c_int, [c_vp, c_i64, c_vp, c_vp, Fin]),
    "ug4b200_vec_scale_add2_norm_ds": (c_int, [c_vp, c_i64, c_vp, Coef, c_vp, Coef, c_vp, Fin]),
    "ug4b200_cg_update_ds": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, Fin]),
    "ug4b200_scalar_ratio_ds": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_scalar_ratio_conv_ds": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_scalar_fin_ds": (c_int, [c_vp, c_vp, Fin]),
    "ug4b200_conv_init": (c_int, [c_vp, c_vp, c_int, c_dbl, c_dbl, c_vp, c_int]),
    "ug4b200_matrix_upload_crs": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, C.POINTER(c_vp)]),
    "ug4b200_matrix_destroy": (c_int, [c_vp, c_vp]),
    "ug4b200_matrix_get_info": (c_int, [c_vp, C.POINTER(MatrixInfo)]),
    "ug4b200_host_stream_plan": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, C.POINTER(StreamPlan), c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_host_value_indexed_stream": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, p_int]),
    "ug4b200_matrix_axpy": (c_int, [c_vp, c_vp, c_vp, c_dbl, c_vp, c_dbl, c_vp, c_int]),
    "ug4b200_matrix_apply": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int]),
    "ug4b200_matrix_matmul_minus": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int]),
    "ug4b200_matrix_apply_ignore_zero_rows": (c_int, [c_vp, c_vp, c_vp, c_dbl, c_vp, c_int]),
    "ug4b200_matrix_apply_dot_ds": (c_int, [c_vp, c_vp, c_vp, c_vp, Fin]),
    "ug4b200_jacobi_prepare": (c_int, [c_vp, c_vp, c_dbl, c_int, c_vp]),
    "ug4b200_matrix_get_diag": (c_int, [c_vp, c_vp, c_vp]),
    "ug4b200_jacobi_invert_diag": (c_int, [c_vp, c_i64, c_int, c_dbl, c_int, c_vp, c_vp]),
    "ug4b200_jacobi_step": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "ug4b200_jacobi_step_add": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_jacobi_smooth_fused": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int]),
    "ug4b200_jacobi_smooth_fused_src": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int]),
    "ug4b200_restrict_jacobi_fused": (c_int, [c_vp, c_vp, c_vp, c_vp, c_dbl, c_vp, c_vp]),
    "ug4b200_color_greedy": (c_int, [c_i64, c_vp, c_vp, c_vp, p_int]),
    "ug4b200_color_check": (c_int, [c_i64, c_vp, c_vp, c_int, c_vp]),
    "ug4b200_gs_step": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_dbl, c_vp, c_vp]),
    "ug4b200_lu_apply": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_coarse_cg": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_dbl, c_dbl]),
    "ug4b200_comm_unique_id": (c_int, [c_vp]),
    "ug4b200_comm_init": (c_int, [c_vp, c_int, c_int, c_vp]),
    "ug4b200_comm_destroy": (c_int, [c_vp]),
    "ug4b200_allreduce_sum": (c_int, [c_vp, c_vp, c_int]),
    "ug4b200_interface_create": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_i64, C.POINTER(c_vp)]),
    "ug4b200_interface_destroy": (c_int, [c_vp, c_vp]),
    "ug4b200_additive_to_consistent": (c_int, [c_vp, c_vp, c_vp, c_int]),
    "ug4b200_additive_to_unique": (c_int, [c_vp, c_vp, c_vp, c_int]),
    "ug4b200_interface_arm": (c_int, [c_vp, c_vp, c_vp]),
    "ug4b200_matrix_apply_dot_allreduce_ds": (c_int, [c_vp, c_vp, c_vp, c_vp, Fin, c_vp]),
    "ug4b200_set_slaves_zero": (c_int, [c_vp, c_vp, c_vp, c_int]),
    "ug4b200_vec_dot_unique_ds": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "ug4b200_gather_create": (c_int, [c_vp, c_i64, c_i64, c_vp, c_int, C.POINTER(c_vp)]),
    "ug4b200_gather_commit": (c_int, [c_vp, c_vp]),
    "ug4b200_gather_sum": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_gather_destroy": (c_int, [c_vp, c_vp]),
    "ug4b200_p2p_window_create": (c_int, [c_vp, C.c_size_t, c_vp, C.POINTER(c_vp)]),
    "ug4b200_p2p_window_open": (c_int, [c_vp, c_int, c_int, c_vp]),
    "ug4b200_p2p_window_attach": (c_int, [c_vp, c_int, c_int, C.POINTER(c_vp)]),
    "ug4b200_p2p_window_destroy": (c_int, [c_vp]),
    "ug4b200_p2p_enabled": (c_int, [c_vp]),
    "ug4b200_p2p_check": (c_int, [c_vp]),
    "ug4b200_interface_commit": (c_int, [c_vp, c_vp]),
    "ug4b200_vec_dot_allreduce_ds": (c_int, [c_vp, c_i64, c_vp, c_vp, Fin, c_vp]),
}

HOST_API = {
    "ug4b200_host_init": (c_int, [c_int, c_vp]),
    "ug4b200_host_finalize": (c_int, []),
    "ug4b200_host_ctx": (c_vp, []),
    "ug4b200_host_last_error": (C.c_char_p, []),
    "ug4b200_host_comm_init": (c_int, [c_int, c_int, c_vp]),
    "ug4b200_solver_create": (c_int, [C.POINTER(SolverDesc), C.POINTER(c_vp)]),
    "ug4b200_solver_destroy": (c_int, [c_vp]),
    "ug4b200_solver_set_matrix": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "ug4b200_solver_set_level": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_solver_set_coloring": (c_int, [c_vp, c_int, c_i64, c_vp, c_int, c_vp]),
    "ug4b200_solver_set_surface_map": (c_int, [c_vp, c_i64, c_vp]),
    "ug4b200_solver_set_debug_dir": (c_int, [c_vp, C.c_char_p, c_vp, c_i64, c_int, c_int]),
    "ug4b200_solver_set_layouts": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_i64]),
    "ug4b200_solver_set_smoother_matrix": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp]),
    "ug4b200_host_vector_selftest": (c_int, [c_int, c_i64, C.c_uint, c_dbl, c_dbl, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_host_apply_transposed": (c_int, [c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_host_matrix_script": (c_int, [c_i64, c_vp, C.POINTER(c_vp)]),
    "ug4b200_host_matrix_isolated": (c_int, [c_vp, c_vp]),
    "ug4b200_host_ilu_factorize": (c_int, [c_i64, c_vp, c_vp, c_vp, c_dbl, c_dbl]),
    "ug4b200_host_ilu_factorize_block": (c_int, [c_int, c_i64, c_vp, c_vp, c_vp, c_dbl, c_dbl]),
    "ug4b200_host_level_sets": (c_int, [c_i64, c_vp, c_vp, c_int, c_vp, C.POINTER(c_int)]),
    "ug4b200_host_cuthill_mckee": (c_int, [c_i64, c_vp, c_vp, c_int, c_int, c_vp]),
    "ug4b200_solver_set_gathered_base": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "ug4b200_solver_set_gathered_level": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_solver_init": (c_int, [c_vp]),
    "ug4b200_solver_apply": (c_int, [c_vp, c_vp, c_vp]),
    "ug4b200_solver_apply_zero_guess": (c_int, [c_vp, c_vp, c_vp]),
    "ug4b200_solver_apply_device": (c_int, [c_vp, c_vp, c_vp]),
    "ug4b200_solver_steps": (c_int, [c_vp]),
    "ug4b200_solver_defect": (c_dbl, [c_vp]),
    "ug4b200_solver_history": (c_int, [c_vp, c_vp, c_int]),
    "ug4b200_solver_precond_apply": (c_int, [c_vp, c_vp, c_vp]),
    "ug4b200_solver_num_dofs": (c_i64, [c_vp]),
    "ug4b200_io_read_matrix": (c_int, [C.c_char_p, c_int, c_int, c_i64, C.POINTER(c_vp)]),
    "ug4b200_io_matrix_info": (c_int, [c_vp, p_i64, p_i64, p_i64, p_int, p_i64]),
    "ug4b200_io_matrix_export": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ug4b200_io_matrix_free": (None, [c_vp]),
    "ug4b200_io_write_matrix": (c_int, [C.c_char_p, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int]),
    "ug4b200_host_rap": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(c_vp)]),
    "ug4b200_io_vector_size": (c_int, [C.c_char_p, p_i64, p_int]),
    "ug4b200_io_read_vector": (c_int, [C.c_char_p, c_i64, c_vp, c_vp]),
    "ug4b200_io_write_vector": (c_int, [C.c_char_p, c_i64, c_vp, c_vp, c_int, c_int]),
}

for _name, (_res, _args) in DEV_API.items():
    _f = getattr(dev, _name)  # AttributeError here = header/library mismatch
    _f.restype, _f.argtypes = _res, _args
for _name, (_res, _args) in HOST_API.items():
    _f = getattr(host, _name)
    _f.restype, _f.argtypes = _res, _args


class UG4B200Error(RuntimeError):
    pass


def check(rc: int, ctx=None) -> int:
    """Raise on a non-zero C-ABI return code (the C++ layer turns it into UG_THROW)."""
    if rc != 0:
        msg = dev.ug4b200_last_error(ctx)
        raise UG4B200Error(f"ug4b200 error {rc}: {msg.decode() if msg else ''}")
    return rc


def check_host(rc: int) -> int:
    if rc < 0:
        raise UG4B200Error("ug4b200 host layer: " + host.ug4b200_host_last_error().decode())
    return rc
