"""Synthetic problem hierarchies (ctypes binding of csrc/synth, host-only).

Stands in for what ugcore's DomainDiscretization/ApproximationSpace hand to the
solve path after assembly: level matrices, P, R, right-hand side.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_LIBDIR = os.environ.get("UG4B200_LIBDIR", os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib"))
_lib = None

POISSON, CONVDIFF, ELASTICITY = 0, 1, 2
ORDER_LEX, ORDER_HIER = 0, 1


class _Desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("base", C.c_int * 3), ("num_refs", C.c_int), ("base_lev", C.c_int),
        ("problem", C.c_int), ("order", C.c_int), ("eps", C.c_double), ("vel", C.c_double * 3),
        ("E", C.c_double), ("nu", C.c_double), ("part", C.c_int * 3), ("coord", C.c_int * 3),
    ]


class _Crs(C.Structure):
    _fields_ = [
        ("nrows", C.c_int64), ("ncols", C.c_int64), ("nnz", C.c_int64), ("block", C.c_int),
        ("rowptr", C.POINTER(C.c_int64)), ("cols", C.POINTER(C.c_int)), ("vals", C.POINTER(C.c_double)),
    ]


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_LIBDIR, "libug4synth.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(path)
        _lib.synth_last_error.restype = C.c_char_p
    return _lib


@dataclass
class Crs:
    """Host CRS matrix (views into generator-owned memory; keep the Problem alive)."""
    nrows: int
    ncols: int
    block: int
    rowptr: np.ndarray  # int64 [nrows+1]
    cols: np.ndarray    # int32 [nnz]
    vals: np.ndarray    # float64 [nnz*block*block], column-major inside a block
    owner: object = None  # keeps the generator (owner of the memory) alive

    @property
    def nnz(self):
        return int(self.cols.size)

    def to_scipy(self):
        import scipy.sparse as sp
        if self.block == 1:
            return sp.csr_matrix((self.vals, self.cols, self.rowptr), shape=(self.nrows, self.ncols))
        b = self.block
        data = self.vals.reshape(-1, b, b).transpose(0, 2, 1)  # column-major -> row-major blocks
        return sp.bsr_matrix((data, self.cols, self.rowptr), shape=(self.nrows * b, self.ncols * b)).tocsr()


class Problem:
    """A refined structured-grid hierarchy of one problem on one rank's sub-box."""

    def __init__(self, dim=3, num_refs=3, problem=POISSON, base=(1, 1, 1), base_lev=0, order=ORDER_LEX,
                 eps=1.0, vel=(1.0, 0.5, 0.25), E=1.0, nu=0.3, part=(1, 1, 1), coord=(0, 0, 0)):
        lib = _load()
        d = _Desc()
        d.dim, d.num_refs, d.base_lev, d.problem, d.order = dim, num_refs, base_lev, problem, order
        d.base = (C.c_int * 3)(*base)
        d.eps, d.E, d.nu = eps, E, nu
        d.vel = (C.c_double * 3)(*vel)
        d.part = (C.c_int * 3)(*part)
        d.coord = (C.c_int * 3)(*coord)
        self._p = C.c_void_p()
        if lib.synth_create(C.byref(d), C.byref(self._p)) != 0:
            raise ValueError(lib.synth_last_error().decode())
        self.dim, self.num_refs, self.base_lev, self.problem = dim, num_refs, base_lev, problem
        self.block = lib.synth_block(self._p)
        self.part, self.coord = tuple(part), tuple(coord)

    def __del__(self):
        try:
            if self._p:
                _load().synth_destroy(self._p)
                self._p = None
        except Exception:
            pass

    def _crs(self, fn, lev) -> Crs:
        c = _Crs()
        if fn(self._p, lev, C.byref(c)) != 0:
            raise ValueError(_load().synth_last_error().decode())
        as_arr = np.ctypeslib.as_array
        return Crs(c.nrows, c.ncols, c.block,
                   as_arr(c.rowptr, (c.nrows + 1,)),
                   as_arr(c.cols, (c.nnz,)) if c.nnz else np.zeros(0, np.int32),
                   as_arr(c.vals, (c.nnz * c.block * c.block,)) if c.nnz else np.zeros(0), self)

    def matrix(self, lev=None) -> Crs:
        return self._crs(_load().synth_level_matrix, self.num_refs if lev is None else lev)

    def prolongation(self, lev) -> Crs:
        return self._crs(_load().synth_prolongation, lev)

    def restriction(self, lev) -> Crs:
        return self._crs(_load().synth_restriction, lev)

    def _arr(self, fn, ctype, *lev):
        p = C.POINTER(ctype)()
        n = C.c_int64()
        if fn(self._p, *lev, C.byref(p), C.byref(n)) != 0:
            raise ValueError(_load().synth_last_error().decode())
        return np.ctypeslib.as_array(p, (n.value,))

    def rhs(self):
        return self._arr(_load().synth_rhs, C.c_double)

    def exact(self):
        return self._arr(_load().synth_exact, C.c_double)

    def dirichlet(self, lev=None):
        return self._arr(_load().synth_dirichlet, C.c_ubyte, self.num_refs if lev is None else lev)

    def global_ids(self, lev=None):
        return self._arr(_load().synth_global_ids, C.c_int64, self.num_refs if lev is None else lev)

    def dof_to_lex(self, lev=None):
        return self._arr(_load().synth_dof_to_lex, C.c_int64, self.num_refs if lev is None else lev)

    def dims(self, lev=None):
        d = (C.c_int * 3)()
        _load().synth_level_dims(self._p, self.num_refs if lev is None else lev, d)
        return tuple(d)

    @property
    def num_dofs(self):
        return self.matrix().nrows * self.block
