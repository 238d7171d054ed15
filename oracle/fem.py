"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Synthetic finite-element systems for the solve-phase oracle: genuine element-by-element
assembly (tensor-product Lagrange elements + Gauss quadrature) on uniform Cartesian meshes of
[0,1]^d, so that the matrices handed to the solvers have the structure the reference's tests
and drivers feed them:

  * Poisson Q1, Dirichlet on the whole boundary, manufactured u = x + y, f = -lap(u) = 0
    (reference: test/LinearSolvers/KrylovTests.jl:11-12,46-61, SmoothersTests.jl:12-29,
     GMGTests.jl:204-215; quadrature degree 2*order+1, KrylovTests.jl:50)
  * factor-2 nested hierarchy with RE-DISCRETISED coarse matrices
    (src/MultilevelTools/FESpaceHierarchies.jl:151-174) and nodal-interpolation prolongation
    P with zero Dirichlet values (src/MultilevelTools/GridTransferOperators.jl:391-401,
    226-230); restriction in mode=:residual is the dual projection == P^T
    (GridTransferOperators.jl:206-208,536-561)
  * vector-valued Q_p linear elasticity (test/Applications/Elasticity.jl:29-37 form) and the
    Q2-P1disc Stokes blocks (joss_paper/demo.jl:20-91) for the C4/C5-style parity cases.

DOF numbering: lexicographic over the (p*n+1)^d node grid, x fastest; vector fields are
node-major (all components of a node contiguous); free DOFs = nodes not on the Dirichlet part,
in lexicographic order.  For Q1 with dirichlet_tags="boundary" this coincides with Gridap's
free-DOF numbering (k-th interior vertex, SURVEY.md App. D, UNVERIFIED against a Julia run);
for Q2 Gridap numbers vertex/edge/face/interior DOFs separately, so Q2 index maps here are a
permutation of the reference's.  No reference-produced matrices exist in the reference tree
(no fixtures), so sparsity parity is "unpinned" (DESIGN.md).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

# ----------------------------------------------------------------------------- 1D pieces


def gauss_legendre_01(npts: int):
    x, w = np.polynomial.legendre.leggauss(npts)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_1d(order: int, xi: np.ndarray):
    """Equispaced Lagrange basis of given order on [0,1]: returns (N[q,a], dN[q,a])."""
    nodes = np.linspace(0.0, 1.0, order + 1)
    nq = len(xi)
    N = np.ones((nq, order + 1))
    dN = np.zeros((nq, order + 1))
    for a in range(order + 1):
        for b in range(order + 1):
            if b != a:
                N[:, a] *= (xi - nodes[b]) / (nodes[a] - nodes[b])
        for c in range(order + 1):
            if c == a:
                continue
            t = np.ones(nq) / (nodes[a] - nodes[c])
            for b in range(order + 1):
                if b != a and b != c:
                    t *= (xi - nodes[b]) / (nodes[a] - nodes[b])
            dN[:, a] += t
    return N, dN


def legendre_p1_1d(xi):
    """Monomial-type P1disc pieces on [0,1] are built in `pdisc_basis`."""
    raise NotImplementedError


# ----------------------------------------------------------------------------- tensor elements


def tensor_basis(order: int, d: int, h, nq: int):
    """Tensor-product Lagrange basis on a cell of size h (tuple).
    Returns phi[q, a], grad[q, a, dim], w[q]  (a lexicographic, x fastest)."""
    xi, wq = gauss_legendre_01(nq)
    N1, dN1 = lagrange_1d(order, xi)
    nb1 = order + 1
    nb, nQ = nb1**d, nq**d
    phi = np.ones((nQ, nb))
    grad = np.ones((nQ, nb, d))
    w = np.ones(nQ)
    # lexicographic with x fastest for both q and a
    qidx = list(itertools.product(*[range(nq)] * d))  # last index fastest -> reverse
    aidx = list(itertools.product(*[range(nb1)] * d))
    for Q, qt in enumerate(qidx):
        q = qt[::-1]  # q[0] is x
        for dim in range(d):
            w[Q] *= wq[q[dim]] * h[dim]
        for A, at in enumerate(aidx):
            a = at[::-1]
            v = 1.0
            for dim in range(d):
                v *= N1[q[dim], a[dim]]
            phi[Q, A] = v
            for g in range(d):
                t = 1.0
                for dim in range(d):
                    t *= (dN1[q[dim], a[dim]] / h[dim]) if dim == g else N1[q[dim], a[dim]]
                grad[Q, A, g] = t
    return phi, grad, w


@dataclass
class Grid:
    """Uniform Cartesian mesh of [0,1]^d with tensor Lagrange nodes of a given order."""

    nc: tuple  # cells per direction
    order: int = 1

    @property
    def d(self):
        return len(self.nc)

    @property
    def h(self):
        return tuple(1.0 / n for n in self.nc)

    @property
    def nn(self):  # nodes per direction
        return tuple(self.order * n + 1 for n in self.nc)

    @property
    def nnodes(self):
        return int(np.prod(self.nn))

    def node_multi_index(self):
        """(nnodes, d) integer grid index of each node, x fastest."""
        grids = np.meshgrid(*[np.arange(n) for n in self.nn], indexing="ij")
        # lexicographic x fastest -> flatten in Fortran order
        return np.stack([g.ravel(order="F") for g in grids], axis=1)

    def node_coords(self):
        mi = self.node_multi_index()
        return mi / (np.array(self.nn) - 1.0)

    def cell_conn(self):
        """(ncells, (order+1)^d) node ids, cells and local nodes lexicographic x fastest."""
        d, p = self.d, self.order
        cgr = np.meshgrid(*[np.arange(n) for n in self.nc], indexing="ij")
        cmi = np.stack([g.ravel(order="F") for g in cgr], axis=1)  # (ncells, d)
        strides = np.concatenate([[1], np.cumprod(self.nn)[:-1]]).astype(np.int64)
        loc = list(itertools.product(*[range(p + 1)] * d))
        conn = np.empty((cmi.shape[0], len(loc)), dtype=np.int64)
        for A, at in enumerate(loc):
            a = np.array(at[::-1])
            conn[:, A] = ((cmi * p + a) * strides).sum(axis=1)
        return conn

    def boundary_mask(self):
        mi = self.node_multi_index()
        nn = np.array(self.nn)
        return ((mi == 0) | (mi == nn - 1)).any(axis=1)


def _assemble(conn_r, conn_c, Ke, nrows, ncols):
    nb_r, nb_c = Ke.shape
    rows = np.repeat(conn_r, nb_c, axis=1).ravel()
    cols = np.tile(conn_c, (1, nb_r)).ravel()
    vals = np.tile(Ke.ravel(), conn_r.shape[0])
    A = sp.coo_matrix((vals, (rows, cols)), shape=(nrows, ncols)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def _vector_conn(conn, ncomp):
    """node-major vector dofs: dof = node*ncomp + comp; local order (node a, comp c) a-major."""
    return (conn[:, :, None] * ncomp + np.arange(ncomp)[None, None, :]).reshape(conn.shape[0], -1)


# ----------------------------------------------------------------------------- Poisson


@dataclass
class AffineSystem:
    A: sp.csr_matrix  # free x free
    b: np.ndarray
    M: sp.csr_matrix | None  # free x free mass matrix (for L2 errors)
    xstar: np.ndarray | None  # nodal interpolant of the manufactured solution at free dofs
    free: np.ndarray  # global dof ids of the free dofs
    grid: Grid
    ncomp: int = 1
    extra: dict = field(default_factory=dict)


def poisson(nc, order=1, sol=lambda X: X[:, 0] + X[:, 1]) -> AffineSystem:
    """-lap(u) = 0 with u = sol on the boundary (sol harmonic).  KrylovTests.jl:46-61."""
    grid = Grid(tuple(nc), order)
    d = grid.d
    phi, grad, w = tensor_basis(order, d, grid.h, order + 1)  # degree 2*order+1 exact
    Ke = np.einsum("q,qad,qbd->ab", w, grad, grad)
    Me = np.einsum("q,qa,qb->ab", w, phi, phi)
    conn = grid.cell_conn()
    n = grid.nnodes
    K = _assemble(conn, conn, Ke, n, n)
    Mfull = _assemble(conn, conn, Me, n, n)
    bnd = grid.boundary_mask()
    free = np.flatnonzero(~bnd)
    dirn = np.flatnonzero(bnd)
    X = grid.node_coords()
    ud = sol(X[dirn])
    A = K[free][:, free].tocsr()
    A.sort_indices()
    b = -(K[free][:, dirn] @ ud)
    M = Mfull[free][:, free].tocsr()
    M.sort_indices()
    return AffineSystem(A=A, b=np.asarray(b).ravel(), M=M, xstar=sol(X[free]), free=free, grid=grid)


def prolongation(coarse: Grid, fine: Grid, free_c: np.ndarray, free_f: np.ndarray, ncomp=1) -> sp.csr_matrix:
    """Nodal interpolation of coarse FE functions (zero Dirichlet values) at the fine nodes,
    restricted to free dofs: y = P x  (GridTransferOperators.jl:391-401 with dv_H = 0 :226-230).
    Built by evaluating the coarse 1D Lagrange bases at the fine node positions."""
    d, p = coarse.d, coarse.order
    assert fine.order == p and fine.d == d
    P1 = []
    for dim in range(d):
        nf, ncn = fine.nn[dim], coarse.nn[dim]
        xf = np.arange(nf) / (nf - 1.0)
        rows, cols, vals = [], [], []
        ncell = coarse.nc[dim]
        for i, x in enumerate(xf):
            c = min(int(np.floor(x * ncell + 1e-12)), ncell - 1)
            xi = x * ncell - c
            # exact dyadic coordinates for factor-2 nesting
            xi = round(xi * 2 * p) / (2 * p) if abs(xi * 2 * p - round(xi * 2 * p)) < 1e-9 else xi
            N, _ = lagrange_1d(p, np.array([xi]))
            for a in range(p + 1):
                if N[0, a] != 0.0:
                    rows.append(i)
                    cols.append(c * p + a)
                    vals.append(N[0, a])
        P1.append(sp.coo_matrix((vals, (rows, cols)), shape=(nf, ncn)).tocsr())
    # kron with x fastest: P = P_z (x) P_y (x) P_x
    P = P1[0]
    for dim in range(1, d):
        P = sp.kron(P1[dim], P, format="csr")
    if ncomp > 1:
        P = sp.kron(P, sp.identity(ncomp), format="csr")
    P = P[free_f][:, free_c].tocsr()
    P.sum_duplicates()
    P.eliminate_zeros()
    P.sort_indices()
    return P


@dataclass
class Hierarchy:
    systems: list  # AffineSystem per level (1 = finest)
    P: list  # P[l]: level l+1 (coarse) -> level l (fine)
    R: list  # R[l] = P[l]^T as CSR

    @property
    def mats(self):
        return [s.A for s in self.systems]


def poisson_hierarchy(nc_fine, nlevels, order=1) -> Hierarchy:
    """CartesianModelHierarchy-style factor-2 nested levels (ModelHierarchies.jl:80-148, nrefs=2),
    re-discretised level matrices (FESpaceHierarchies.jl:163-170), rhs on level 1 only."""
    systems, Ps, Rs = [], [], []
    nc = tuple(nc_fine)
    for lev in range(nlevels):
        systems.append(poisson(nc, order))
        if lev < nlevels - 1:
            assert all(n % 2 == 0 for n in nc), "factor-2 coarsening needs even cell counts"
            nc = tuple(n // 2 for n in nc)
    for lev in range(nlevels - 1):
        f, c = systems[lev], systems[lev + 1]
        P = prolongation(c.grid, f.grid, c.free, f.free)
        Ps.append(P)
        R = P.T.tocsr()
        R.sort_indices()
        Rs.append(R)
    return Hierarchy(systems, Ps, Rs)


def l2_error_sq(sys: AffineSystem, x: np.ndarray) -> float:
    """E = int (u_h - I_h u)^2 = (x-x*)^T M_ff (x-x*)  (KrylovTests.jl:21-25)."""
    e = x - sys.xstar
    return float(e @ (sys.M @ e))


# ----------------------------------------------------------------------------- elasticity (C4)


def elasticity(nc, order=2, lam=1.0, mu=1.0, body=(0.0, 0.0, -1.0)) -> AffineSystem:
    """Isotropic linear elasticity sigma = lam tr(eps) I + 2 mu eps, clamped on x=0, constant
    body force (bilinear form as test/Applications/Elasticity.jl:31-37; lam=mu=1 per SURVEY 8d)."""
    grid = Grid(tuple(nc), order)
    d = grid.d
    phi, grad, w = tensor_basis(order, d, grid.h, order + 1)
    nb = phi.shape[1]
    # vector basis (a, c): eps_ij = 0.5 (d_j phi_a delta_ic + d_i phi_a delta_jc)
    nv = nb * d
    Ke = np.zeros((nv, nv))
    # a(u,v) = int lam div u div v + 2 mu eps(u):eps(v)
    G = grad  # [q,a,g]
    for c1 in range(d):
        for c2 in range(d):
            blk = lam * np.einsum("q,qa,qb->ab", w, G[:, :, c1], G[:, :, c2])
            blk += mu * np.einsum("q,qa,qb->ab", w, G[:, :, c2], G[:, :, c1])
            if c1 == c2:
                blk += mu * np.einsum("q,qad,qbd->ab", w, G, G)
            Ke[c1::d, c2::d] = blk  # rows (a,c1) -> index a*d+c1
    Fe = np.zeros(nv)
    for c in range(d):
        Fe[c::d] = body[c] * (w @ phi)
    conn = _vector_conn(grid.cell_conn(), d)
    n = grid.nnodes * d
    K = _assemble(conn, conn, Ke, n, n)
    F = np.bincount(conn.ravel(), weights=np.tile(Fe, conn.shape[0]), minlength=n)
    mi = grid.node_multi_index()
    clamped = np.repeat(mi[:, 0] == 0, d)
    free = np.flatnonzero(~clamped)
    A = K[free][:, free].tocsr()
    A.sort_indices()
    return AffineSystem(A=A, b=F[free], M=None, xstar=None, free=free, grid=grid, ncomp=d)


def elasticity_hierarchy(nc_fine, nlevels, order=2, **kw) -> Hierarchy:
    systems, Ps, Rs = [], [], []
    nc = tuple(nc_fine)
    for lev in range(nlevels):
        systems.append(elasticity(nc, order, **kw))
        if lev < nlevels - 1:
            nc = tuple(n // 2 for n in nc)
    for lev in range(nlevels - 1):
        f, c = systems[lev], systems[lev + 1]
        P = prolongation(c.grid, f.grid, c.free, f.free, ncomp=f.ncomp)
        Ps.append(P)
        R = P.T.tocsr()
        R.sort_indices()
        Rs.append(R)
    return Hierarchy(systems, Ps, Rs)


# ----------------------------------------------------------------------------- Stokes (C5)


def stokes_cavity(nc, nlevels=1):
    """2D lid-driven cavity, Q2 velocity / P1-discontinuous pressure (joss_paper/demo.jl:20-91):
    blocks A (velocity Laplacian-type, int grad u : grad v), B (-(div v) p), pressure mass Mp,
    rhs from the lid Dirichlet data u=(1,0) on y=1.  Returns dict with block matrices and, when
    nlevels>1, the velocity-block hierarchy (A_l re-discretised, P_l nodal) for the GMG block."""
    out = {}
    levels = []
    ncl = tuple(nc)
    for lev in range(nlevels):
        grid = Grid(ncl, 2)
        d = 2
        phi, grad, w = tensor_basis(2, d, grid.h, 3)
        nb = phi.shape[1]
        Kes = np.einsum("q,qad,qbd->ab", w, grad, grad)
        Ke = np.zeros((nb * d, nb * d))
        for c in range(d):
            Ke[c::d, c::d] = Kes
        conn = _vector_conn(grid.cell_conn(), d)
        n = grid.nnodes * d
        K = _assemble(conn, conn, Ke, n, n)
        bnd = np.repeat(grid.boundary_mask(), d)
        free = np.flatnonzero(~bnd)
        lvl = dict(grid=grid, K=K, free=free, bnd=bnd, conn=conn, phi=phi, grad=grad, w=w)
        levels.append(lvl)
        if lev < nlevels - 1:
            ncl = tuple(n // 2 for n in ncl)
    L = levels[0]
    grid, K, free, conn, phi, grad, w = L["grid"], L["K"], L["free"], L["conn"], L["phi"], L["grad"], L["w"]
    d = 2
    ncell = conn.shape[0]
    # P1disc basis per cell: {1, xi-1/2, eta-1/2} in reference coordinates
    xi, _ = gauss_legendre_01(3)
    qx = np.tile(xi, 3)  # x fastest
    qy = np.repeat(xi, 3)
    psi = np.stack([np.ones(9), qx - 0.5, qy - 0.5], axis=1)  # [q, m]
    # B_e[m, (a,c)] = - int psi_m d_c phi_a
    Be = np.zeros((3, phi.shape[1] * d))
    for c in range(d):
        Be[:, c::d] = -np.einsum("q,qm,qa->ma", w, psi, grad[:, :, c])
    Mpe = np.einsum("q,qm,qn->mn", w, psi, psi)
    pconn = (np.arange(ncell)[:, None] * 3 + np.arange(3)[None, :]).astype(np.int64)
    npdof = ncell * 3
    B = _assemble(pconn, conn, Be, npdof, grid.nnodes * d)
    Mp = _assemble(pconn, pconn, Mpe, npdof, npdof)
    # Dirichlet data: u = (1,0) on the lid y=1 (excluding nothing: demo tags the top edge incl. corners)
    mi = grid.node_multi_index()
    ud = np.zeros(grid.nnodes * d)
    lid = mi[:, 1] == grid.nn[1] - 1
    ud[np.flatnonzero(lid) * d + 0] = 1.0
    dirn = np.flatnonzero(L["bnd"])
    A = K[free][:, free].tocsr()
    A.sort_indices()
    Bf = B[:, free].tocsr()
    Bf.sort_indices()
    fu = -(K[free][:, dirn] @ ud[dirn])
    fp = -(B[:, dirn] @ ud[dirn])
    out.update(A=A, B=Bf, Bt=Bf.T.tocsr(), Mp=Mp.tocsr(), fu=np.asarray(fu).ravel(), fp=np.asarray(fp).ravel(),
               free=free, grid=grid)
    out["Bt"].sort_indices()
    if nlevels > 1:
        mats = [A]
        frees = [free]
        for lvl in levels[1:]:
            Al = lvl["K"][lvl["free"]][:, lvl["free"]].tocsr()
            Al.sort_indices()
            mats.append(Al)
            frees.append(lvl["free"])
        Ps, Rs = [], []
        for l in range(nlevels - 1):
            P = prolongation(levels[l + 1]["grid"], levels[l]["grid"], frees[l + 1], frees[l], ncomp=d)
            Ps.append(P)
            R = P.T.tocsr()
            R.sort_indices()
            Rs.append(R)
        out.update(mats=mats, P=Ps, R=Rs)
    return out
