"""Shared helpers: the same host arrays are handed to the CUDA path (C ABI) and to the oracle."""
import numpy as np
import scipy.sparse as sp

from oracle import linalg as ola
from oracle import solvers as OS


def csr_triplet(A):
    A = sp.csr_matrix(A)
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64)


def dev_matrix(gsb, ctx, A):
    return gsb.SparseMatrix.from_scipy(sp.csr_matrix(A), ctx)


def dev_vec(gsb, A, values=None, domain=True):
    v = gsb.allocate_in_domain(A) if domain else gsb.allocate_in_range(A)
    if values is not None:
        v.set(values)
    return v


def host_to_scipy(triplet, ncols):
    rp, c, v = triplet
    return sp.csr_matrix((v, c, rp), shape=(rp.shape[0] - 1, ncols))


def oracle_hierarchy(hh):
    """oracle CSR matrices of a single-rank synth.HostHierarchy"""
    mats = [ola.CSR(host_to_scipy(a, lp.n_own)) for a, lp in zip(hh.A, hh.levels)]
    P = [ola.CSR(host_to_scipy(p, hh.levels[l + 1].n_own)) for l, p in enumerate(hh.P)]
    R = [ola.CSR(host_to_scipy(r, hh.levels[l].n_own)) for l, r in enumerate(hh.R)]
    return mats, P, R


def rel_hist_diff(h_gpu, h_ref):
    """max_k |r_k^gpu - r_k^ref| / r_0^ref : difference of the RELATIVE residual histories."""
    n = min(len(h_gpu), len(h_ref))
    return float(np.max(np.abs(np.asarray(h_gpu[:n]) - np.asarray(h_ref[:n]))) / h_ref[0])
