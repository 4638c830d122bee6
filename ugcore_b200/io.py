"""Import / export of assembled matrices and vectors in ugcore's own dump formats.

ConnectionViewer ``.mat`` / ``.vec`` (ugbase/lib_algebra/common/connection_viewer_{output,input}.h)
and MatrixMarket ``.mtx`` (ugbase/lib_algebra/common/matrixio/matrix_io_mtx.{h,cpp}) are what a
ugcore installation writes when asked for its assembled objects: the
surface matrix, every GMG level matrix, P, R, right-hand sides and per-iteration residuals.
Reading them here lets a REAL UG4 assembly run through this solve path and be compared with
ugcore's own residual history.  The work is done by the host layer (csrc/host/matrix_io.h)
through the C ABI in include/ug4b200_solver.h; no device is involved.

Note: ugcore writes ``.mat`` values with the stream's default precision (6 significant digits) —
such a dump does not reproduce a residual history to 1e-10.  Dumps meant for parity must be
written with a raised precision (``precision=17`` here) or as MatrixMarket files with
``precision=16``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import check_host, host
from .problems import Crs

FORMAT_AUTO, FORMAT_CONNECTION_VIEWER, FORMAT_MATRIX_MARKET = 0, 1, 2


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def read_matrix(path: str, keep_zeros: bool = False, n_to: int = 0, fmt: int = FORMAT_AUTO):
    """Returns (Crs, positions[npos, 3] or None, dim).  ``n_to`` > 0: a ConnectionViewer file in the
    from / to form (rectangular P / R): rows 0..n_to-1, columns numbered behind them."""
    h = C.c_void_p()
    check_host(-abs(host.ug4b200_io_read_matrix(str(path).encode(), fmt, int(keep_zeros), int(n_to), C.byref(h))))
    try:
        nr, nc, nnz, npos = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        dim = C.c_int()
        check_host(-abs(host.ug4b200_io_matrix_info(h, C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(dim), C.byref(npos))))
        rowptr = np.zeros(nr.value + 1, np.int64)
        cols = np.zeros(nnz.value, np.int32)
        vals = np.zeros(nnz.value, np.float64)
        pos = np.zeros((npos.value, 3)) if npos.value else None
        check_host(-abs(host.ug4b200_io_matrix_export(h, _p(rowptr), _p(cols), _p(vals), _p(pos))))
    finally:
        host.ug4b200_io_matrix_free(h)
    return Crs(nrows=nr.value, ncols=nc.value, block=1, rowptr=rowptr, cols=cols, vals=vals), pos, dim.value


def write_matrix(path: str, A: Crs, positions=None, dim: int = 3, from_to: bool = False, precision: int = 0,
                 fmt: int = FORMAT_AUTO) -> None:
    """``positions``: [nrows, 3] (square ConnectionViewer file) or, with ``from_to``, [nrows + ncols, 3]
    (rows first).  ``precision`` 0 reproduces the reference writer byte for byte."""
    if A.block != 1:
        raise ValueError("scalar matrices only")
    pos = None if positions is None else np.ascontiguousarray(positions, np.float64)
    rowptr = np.ascontiguousarray(A.rowptr, np.int64)
    cols = np.ascontiguousarray(A.cols, np.int32)
    vals = np.ascontiguousarray(A.vals, np.float64)
    check_host(-abs(host.ug4b200_io_write_matrix(str(path).encode(), fmt, A.nrows, A.ncols, _p(rowptr), _p(cols), _p(vals),
                                                 _p(pos), dim, int(from_to), precision)))


def read_vector(path: str):
    """Returns (values, positions[n, 3], dim)."""
    n, dim = C.c_int64(), C.c_int()
    check_host(-abs(host.ug4b200_io_vector_size(str(path).encode(), C.byref(n), C.byref(dim))))
    v = np.zeros(n.value)
    pos = np.zeros((n.value, 3))
    check_host(-abs(host.ug4b200_io_read_vector(str(path).encode(), n.value, _p(v), _p(pos))))
    return v, pos, dim.value


def write_vector(path: str, values, positions=None, dim: int = 3, precision: int = 0) -> None:
    v = np.ascontiguousarray(values, np.float64)
    pos = None if positions is None else np.ascontiguousarray(positions, np.float64)
    check_host(-abs(host.ug4b200_io_write_vector(str(path).encode(), v.size, _p(v), _p(pos), dim, precision)))
