"""Pins the oracle and the CUDA path to the REAL GridapSolvers.jl through fixtures exported by tools/export_reference.jl
(tests/golden/reference_*.npz: the reference's own assembled level matrices, transfer matrices, right-hand side and
`solver.log.residuals`).  The build container has no Julia, so the fixtures cannot be generated here: while there is none
every test below SKIPS (loudly) and DESIGN.md section 7 says "parity unpinned"; dropping the files in turns them on with
no code change.  Bar (BASELINE.json north_star): iteration count +-1, relative residual history within 1e-10."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import linalg as ola
from oracle import solvers as OS
from util import rel_hist_diff

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_*.npz")))
needs_fixture = pytest.mark.skipif(not GOLDEN, reason="no tests/golden/reference_*.npz: run tools/export_reference.jl where "
                                                      "GridapSolvers.jl is installed (parity stays UNPINNED until then)")


def _csc(d, name):
    shape = tuple(int(v) for v in d[name + "_shape"])
    return sp.csc_matrix((d[name + "_nzval"], d[name + "_rowval"] - 1, d[name + "_colptr"] - 1), shape=shape).tocsr()


def _load(path):
    d = np.load(path, allow_pickle=True)
    nlev = int(d["nlev"])
    mats = [_csc(d, f"A{l}") for l in range(1, nlev + 1)]
    P = [_csc(d, f"P{l}") for l in range(1, nlev)]
    R = [_csc(d, f"R{l}") for l in range(1, nlev)]
    return d, str(d["kind"]), mats, P, R


def _solver(S, kind, d, mats, P, R):
    tol = dict(maxiter=int(d["maxiter"]), atol=float(d["atol"]), rtol=float(d["rtol"]))
    if kind == "cg_jacobi":
        return S.CGSolver(S.JacobiLinearSolver(), **tol)
    assert kind == "gmg_pcg"
    n = len(mats)
    sm = [S.RichardsonSmoother(S.JacobiLinearSolver(), int(d["niter_smooth"]), float(d["omega"]))] * (n - 1)
    gmg = S.GMGLinearSolver(mats, P, R, pre_smoothers=sm, post_smoothers=sm, coarsest_solver=S.LUSolver(), maxiter=1,
                            mode="preconditioner", cycle_type="v_cycle")
    return S.CGSolver(gmg, **tol)


@needs_fixture
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_the_reference_history(path):
    d, kind, mats, P, R = _load(path)
    for p_, r_ in zip(P, R):  # what the device path assumes about mode=:residual transfers
        assert abs(r_ - p_.T).max() <= 1e-12 * abs(p_).max()
    om = [ola.CSR(m) for m in mats]
    s = _solver(OS, kind, d, om, [ola.CSR(p_) for p_ in P], [ola.CSR(r_) for r_ in R])
    x = np.zeros(mats[0].shape[1])
    OS.solve_(x, OS.numerical_setup(OS.symbolic_setup(s, om[0]), om[0]), np.asarray(d["b"], dtype=np.float64))
    ref = np.asarray(d["residuals"], dtype=np.float64)
    assert abs(s.log.num_iters - int(d["num_iters"])) <= 1
    n = min(s.log.num_iters, int(d["num_iters"])) + 1
    assert rel_hist_diff(s.log.history()[:n], ref[:n]) < 1e-10
    assert np.linalg.norm(x - d["x"]) <= 1e-6 * np.linalg.norm(d["x"])


@needs_fixture
@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_path_reproduces_the_reference_history(gsb, ctx, path):
    from util import dev_matrix, dev_vec

    d, kind, mats, P, R = _load(path)
    dm = [dev_matrix(gsb, ctx, m) for m in mats]
    s = _solver(gsb, kind, d, dm, [dev_matrix(gsb, ctx, p_) for p_ in P], [dev_matrix(gsb, ctx, r_) for r_ in R])
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, dm[0]), dm[0])
    x = dev_vec(gsb, dm[0])
    gsb.solve_(x, ns, dev_vec(gsb, dm[0], np.asarray(d["b"], dtype=np.float64)))
    ref = np.asarray(d["residuals"], dtype=np.float64)
    assert abs(s.log.num_iters - int(d["num_iters"])) <= 1
    n = min(s.log.num_iters, int(d["num_iters"])) + 1
    assert rel_hist_diff(s.log.history()[:n], ref[:n]) < 1e-10


@needs_fixture
@pytest.mark.parametrize("path", [p for p in GOLDEN if "parts" in os.path.basename(p)] or [None])
def test_index_maps_of_the_reference_partition(path):
    """own-first renumbering per part: own ids ascending in global order inside a part, every global dof owned once,
    ghosts of a part owned by another part -- the contract GridapSolversB200.jl (B200Matrix) and include/gsb200.h
    (gsb_plan_create) build on; and the Cartesian generator of this repository partitions the same mesh the same way"""
    if path is None:
        pytest.skip("no distributed fixture")
    d = np.load(path, allow_pickle=True)
    nparts, nlev = int(d["nparts"]), int(d["nlev"])
    for l in range(1, nlev + 1):
        n = int(d[f"A{l}_shape"][0])
        seen = np.zeros(n, dtype=np.int64)
        owner = np.zeros(n, dtype=np.int64)
        for p_ in range(1, nparts + 1):
            o2g = d[f"part{p_}_own_to_global_{l}"] - 1
            seen[o2g] += 1
            owner[o2g] = p_
        assert (seen == 1).all()
        for p_ in range(1, nparts + 1):
            g2g, gown = d[f"part{p_}_ghost_to_global_{l}"] - 1, d[f"part{p_}_ghost_owner_{l}"]
            assert (owner[g2g] == gown).all() and (gown != p_).all()


def test_fixture_layout_round_trip(tmp_path):
    """keeps the loader honest while no real fixture exists: a file with the exporter's layout (1-based CSC triplets,
    residuals, scalars) written from the oracle's own hierarchy loads back and drives the same comparison code"""
    from oracle import fem

    H = fem.poisson_hierarchy((8, 8, 8), 2)
    d = {"kind": "gmg_pcg", "nlev": 2, "nparts": 1, "rtol": 1e-8, "atol": 1e-14, "maxiter": 30, "niter_smooth": 10, "omega": 2.0 / 3.0,
         "b": H.systems[0].b}
    for name, M in (("A1", H.mats[0]), ("A2", H.mats[1]), ("P1", H.P[0]), ("R1", H.R[0])):
        C = sp.csc_matrix(M)
        C.sort_indices()
        d[name + "_colptr"], d[name + "_rowval"] = C.indptr.astype(np.int64) + 1, C.indices.astype(np.int64) + 1
        d[name + "_nzval"], d[name + "_shape"] = C.data, np.array(C.shape, dtype=np.int64)
    om = [ola.CSR(m) for m in H.mats]
    s = _solver(OS, "gmg_pcg", d, om, [ola.CSR(H.P[0])], [ola.CSR(H.R[0])])
    x = np.zeros(H.mats[0].shape[1])
    OS.solve_(x, OS.numerical_setup(OS.symbolic_setup(s, om[0]), om[0]), d["b"])
    d["x"], d["num_iters"], d["residuals"] = x, s.log.num_iters, s.log.history()
    path = tmp_path / "reference_selftest.npz"
    np.savez(path, **d)
    d2, kind, mats, P, R = _load(str(path))
    assert kind == "gmg_pcg" and len(mats) == 2 and abs(mats[0] - H.mats[0]).max() == 0.0 and abs(P[0] - H.P[0]).max() == 0.0
    s2 = _solver(OS, kind, d2, [ola.CSR(m) for m in mats], [ola.CSR(P[0])], [ola.CSR(R[0])])
    x2 = np.zeros(mats[0].shape[1])
    OS.solve_(x2, OS.numerical_setup(OS.symbolic_setup(s2, ola.CSR(mats[0])), ola.CSR(mats[0])), np.asarray(d2["b"]))
    assert s2.log.num_iters == int(d2["num_iters"]) and rel_hist_diff(s2.log.history(), np.asarray(d2["residuals"])) < 1e-12
