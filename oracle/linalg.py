"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Array primitives the reference's solvers call: SpMV `mul!`, `dot`, `norm`,
broadcast axpy-likes -- restated with the rounding sequence of the Julia dependencies
(SparseArrays / PartitionedArrays / LinearAlgebra; see oracle/csr_kernels.c header).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_THREADED = False  # True only for the CPU *timing* baseline


def build(force: bool = False) -> str:
    """gcc the C restatement (oracle/Makefile does the same)."""
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "csr_kernels.c")
    if force or (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", so, src, "-lm"]
        subprocess.check_call(cmd)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        i64, dbl = ctypes.c_int64, ctypes.c_double
        P = ctypes.c_void_p
        L.oracle_csr_spmv.argtypes = [i64, P, P, P, P, P]
        L.oracle_csr_spmv5.argtypes = [i64, P, P, P, P, P, dbl, dbl]
        for n in ("oracle_ew_mul", "oracle_ew_add", "oracle_ew_sub"):
            getattr(L, n).argtypes = [i64, P, P, P]
        L.oracle_ew_scale.argtypes = [i64, dbl, P, P]
        L.oracle_ew_axpy.argtypes = [i64, P, dbl, P, P]
        L.oracle_ew_axmy.argtypes = [i64, P, dbl, P, P]
        L.oracle_dot.argtypes = [i64, P, P]
        L.oracle_dot.restype = dbl
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


def set_threaded(on: bool, nthreads: int | None = None) -> int:
    """Timing-baseline mode: elementwise ops and reductions use the OpenMP C loops."""
    global _THREADED
    _THREADED = bool(on)
    L = lib()
    if nthreads:
        L.oracle_set_threads(int(nthreads))
    elif not on:
        L.oracle_set_threads(1)
    return L.oracle_max_threads()


def _p(a: np.ndarray):
    return a.ctypes.data


class CSR:
    """Local sparse matrix, CSR, fp64 values, int32 column ids (ascending within rows)."""

    def __init__(self, A):
        A = sp.csr_matrix(A)
        if not A.has_sorted_indices:
            A = A.copy()
            A.sort_indices()
        self.shape = A.shape
        self.rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.col = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.val = np.ascontiguousarray(A.data, dtype=np.float64)
        self.sp = A

    @property
    def nnz(self):
        return int(self.rowptr[-1])

    def diag(self):
        return np.asarray(self.sp.diagonal(), dtype=np.float64)

    def to_scipy(self):
        return self.sp


def mul(y: np.ndarray, A: CSR, x: np.ndarray):
    """mul!(y,A,x)"""
    assert y.shape[0] == A.shape[0] and x.shape[0] >= A.shape[1]
    assert y.flags.c_contiguous and x.flags.c_contiguous and y is not x
    lib().oracle_csr_spmv(A.shape[0], _p(A.rowptr), _p(A.col), _p(A.val), _p(x), _p(y))
    return y


def mul5(y: np.ndarray, A: CSR, x: np.ndarray, alpha: float, beta: float):
    """mul!(y,A,x,alpha,beta)"""
    lib().oracle_csr_spmv5(A.shape[0], _p(A.rowptr), _p(A.col), _p(A.val), _p(x), _p(y), float(alpha), float(beta))
    return y


def dot(a, b) -> float:
    if _THREADED:
        return float(lib().oracle_dot(a.shape[0], _p(a), _p(b)))
    return float(np.dot(a, b))


def norm(a) -> float:
    if _THREADED:
        return float(np.sqrt(lib().oracle_dot(a.shape[0], _p(a), _p(a))))
    return float(np.sqrt(np.dot(a, a)))


# elementwise broadcasts; out may alias inputs (purely elementwise)
def emul(z, a, b):  # z .= a .* b
    if _THREADED:
        lib().oracle_ew_mul(z.shape[0], _p(a), _p(b), _p(z))
    else:
        np.multiply(a, b, out=z)
    return z


def scale(z, s, a):  # z .= s .* a
    if _THREADED:
        lib().oracle_ew_scale(z.shape[0], float(s), _p(a), _p(z))
    else:
        np.multiply(a, s, out=z)
    return z


def add(z, a, b):  # z .= a .+ b
    if _THREADED:
        lib().oracle_ew_add(z.shape[0], _p(a), _p(b), _p(z))
    else:
        np.add(a, b, out=z)
    return z


def sub(z, a, b):  # z .= a .- b
    if _THREADED:
        lib().oracle_ew_sub(z.shape[0], _p(a), _p(b), _p(z))
    else:
        np.subtract(a, b, out=z)
    return z


def axpy(z, a, s, b):  # z .= a .+ s .* b
    if _THREADED:
        lib().oracle_ew_axpy(z.shape[0], _p(a), float(s), _p(b), _p(z))
    else:
        t = np.multiply(b, s)
        np.add(a, t, out=z)
    return z


def axmy(z, a, s, b):  # z .= a .- s .* b
    if _THREADED:
        lib().oracle_ew_axmy(z.shape[0], _p(a), float(s), _p(b), _p(z))
    else:
        t = np.multiply(b, s)
        np.subtract(a, t, out=z)
    return z
