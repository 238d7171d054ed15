"""Host-side synthetic problem generator (stands in for the Gridap / GridapDistributed assembly
the Julia host does before the solve phase).  Not on the solve path.

Q1 Poisson on a uniform Cartesian mesh of [0,1]^d, Dirichlet everywhere, u = x + y
(test/LinearSolvers/KrylovTests.jl:11-12,46-61; GMGTests.jl:204-215), a factor-2 nested level
hierarchy with re-discretised level matrices (src/MultilevelTools/FESpaceHierarchies.jl:151-174)
and explicit nodal prolongations P / restrictions R = P^T (GridTransferOperators.jl:391-401,
536-561), row-partitioned over a px x py x pz box of ranks with PartitionedArrays-style own-first
local numbering and one ghost layer.

Ownership rule (unverified against a Julia run): cells are split in equal
blocks per direction; a node belongs to the highest part among the cells touching it, i.e. node
i of a direction belongs to part i // (n/p).  Ghosts are numbered after own dofs, sorted by
(owner rank, owner-local id); the reference's ghost order is assembly-dependent and not
recoverable without a Julia dump (DESIGN.md "index-map parity").
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib


def _p(a):
    return None if a is None else a.ctypes.data


def part_coords(rank, parts):
    c = []
    for p in parts:
        c.append(rank % p)
        rank //= p
    return tuple(c)


def part_rank(coords, parts):
    r, s = 0, 1
    for c, p in zip(coords, parts):
        r += c * s
        s *= p
    return r


def own_range(n, p, q):
    """own free nodes [lo,hi) of part q in a direction with n cells split over p parts."""
    assert n % p == 0, "cells per direction must divide evenly over the parts"
    lo = max(1, q * (n // p))
    hi = (q + 1) * (n // p) if q < p - 1 else n
    return lo, hi


@dataclass
class LevelPart:
    ncell: tuple
    parts: tuple
    rank: int
    olo: np.ndarray
    ohi: np.ndarray
    elo: np.ndarray
    ehi: np.ndarray
    ext_lid: np.ndarray  # int32 over the extended box, x fastest; >=0 local id, -2 Dirichlet
    n_own: int
    n_ghost: int
    ghost_owner: np.ndarray
    ghost_owner_lid: np.ndarray
    lengths: tuple = None  # domain edge lengths (default 1.0 each)
    nbr_snd: np.ndarray = field(default=None)
    snd_ptrs: np.ndarray = field(default=None)
    snd_ids: np.ndarray = field(default=None)
    nbr_rcv: np.ndarray = field(default=None)
    rcv_ptrs: np.ndarray = field(default=None)
    rcv_ids: np.ndarray = field(default=None)

    @property
    def d(self):
        return len(self.ncell)

    def own_offset_global(self):
        """first global id of this part's own dofs (own-first global numbering, part after part)."""
        off = 0
        for r in range(self.rank):
            c = part_coords(r, self.parts)
            n = 1
            for k in range(self.d):
                lo, hi = own_range(self.ncell[k], self.parts[k], c[k])
                n *= hi - lo
            off += n
        return off


def _box_lids(lo, hi, coords):
    """lexicographic (x fastest) id of node coords (N,d) inside box [lo,hi)."""
    lid = np.zeros(coords.shape[0], dtype=np.int64)
    stride = 1
    for k in range(coords.shape[1]):
        lid += (coords[:, k] - lo[k]) * stride
        stride *= hi[k] - lo[k]
    return lid


def _box_coords(lo, hi):
    axes = [np.arange(lo[k], hi[k], dtype=np.int64) for k in range(len(lo))]
    grids = np.meshgrid(*axes, indexing="ij")
    return np.stack([g.ravel(order="F") for g in grids], axis=1)


def make_level_part(ncell, parts, rank, lengths=None) -> LevelPart:
    ncell, parts = tuple(int(n) for n in ncell), tuple(int(p) for p in parts)
    d = len(ncell)
    pc = part_coords(rank, parts)
    olo = np.array([own_range(ncell[k], parts[k], pc[k])[0] for k in range(d)], dtype=np.int64)
    ohi = np.array([own_range(ncell[k], parts[k], pc[k])[1] for k in range(d)], dtype=np.int64)
    elo, ehi = olo - 1, ohi + 1
    X = _box_coords(elo, ehi)
    nvec = np.array(ncell, dtype=np.int64)
    dirichlet = ((X == 0) | (X == nvec)).any(axis=1)
    own = ((X >= olo) & (X < ohi)).all(axis=1)
    ext_lid = np.full(X.shape[0], -1, dtype=np.int64)
    ext_lid[dirichlet] = -2
    ext_lid[own] = _box_lids(olo, ohi, X[own])
    n_own = int(np.prod(ohi - olo))
    ghost = ~own & ~dirichlet
    G = X[ghost]
    # owner part of each ghost node
    per = np.array([ncell[k] // parts[k] for k in range(d)], dtype=np.int64)
    gq = np.minimum(G // per, np.array(parts) - 1)
    owner = np.zeros(G.shape[0], dtype=np.int64)
    stride = 1
    for k in range(d):
        owner += gq[:, k] * stride
        stride *= parts[k]
    owner_lid = np.zeros(G.shape[0], dtype=np.int64)
    for r in np.unique(owner):
        m = owner == r
        c = part_coords(int(r), parts)
        rlo = np.array([own_range(ncell[k], parts[k], c[k])[0] for k in range(d)])
        rhi = np.array([own_range(ncell[k], parts[k], c[k])[1] for k in range(d)])
        owner_lid[m] = _box_lids(rlo, rhi, G[m])
    order = np.lexsort((owner_lid, owner))
    gl = np.empty(G.shape[0], dtype=np.int64)
    gl[order] = n_own + np.arange(G.shape[0])
    ext_lid[ghost] = gl
    lp = LevelPart(ncell, parts, rank, olo, ohi, elo, ehi, ext_lid.astype(np.int32), n_own, int(G.shape[0]),
                   owner[order].astype(np.int32), owner_lid[order],
                   tuple(float(v) for v in (lengths if lengths is not None else (1.0,) * d)))
    # receive lists: ghosts grouped by owner (already sorted)
    nbr_rcv, counts = np.unique(lp.ghost_owner, return_counts=True)
    lp.nbr_rcv = nbr_rcv.astype(np.int32)
    lp.rcv_ptrs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    lp.rcv_ids = (n_own + np.arange(lp.n_ghost)).astype(np.int64)
    # send lists: my own nodes lying in each neighbour's extended box, ascending in my local id
    nbrs, ptrs, ids = [], [0], []
    offs = np.stack(np.meshgrid(*[[-1, 0, 1]] * d, indexing="ij"), axis=-1).reshape(-1, d)
    cand = []
    for o in offs:
        qc = np.array(pc) + o[::-1] if False else np.array(pc) + o
        if (o == 0).all() or (qc < 0).any() or (qc >= np.array(parts)).any():
            continue
        cand.append(part_rank(tuple(int(v) for v in qc), parts))
    for q in sorted(set(cand)):
        c = part_coords(q, parts)
        qlo = np.array([own_range(ncell[k], parts[k], c[k])[0] for k in range(d)]) - 1
        qhi = np.array([own_range(ncell[k], parts[k], c[k])[1] for k in range(d)]) + 1
        ilo, ihi = np.maximum(qlo, olo), np.minimum(qhi, ohi)
        if (ihi <= ilo).any():
            continue
        mine = np.sort(_box_lids(olo, ohi, _box_coords(ilo, ihi)))
        nbrs.append(q)
        ids.append(mine)
        ptrs.append(ptrs[-1] + mine.shape[0])
    lp.nbr_snd = np.array(nbrs, dtype=np.int32)
    lp.snd_ptrs = np.array(ptrs, dtype=np.int64)
    lp.snd_ids = (np.concatenate(ids) if ids else np.zeros(0)).astype(np.int64)
    return lp


def _two_pass(fn, n_rows, *args, with_b=False):
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    fn(*args, 0, _p(rowptr), None, None, *([None] if with_b else []))
    np.cumsum(rowptr, out=rowptr)
    nnz = int(rowptr[-1])
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    b = np.zeros(n_rows) if with_b else None
    fn(*args, 1, _p(rowptr), _p(col), _p(val), *([_p(b)] if with_b else []))
    return rowptr, col, val, b


def poisson_rows(lp: LevelPart):
    """(rowptr, col, val, b) of this part's rows of the Q1 Laplacian + Dirichlet lift of u=x+y."""
    S = _lib.synth()
    nc = np.array(lp.ncell, dtype=np.int64)
    L = np.array(lp.lengths, dtype=np.float64)
    return _two_pass(S.synth_poisson_rows, lp.n_own, lp.d, _p(nc), _p(L), _p(lp.elo), _p(lp.ehi), _p(lp.ext_lid), _p(lp.olo),
                     _p(lp.ohi), with_b=True)


def mass_rows(lp: LevelPart):
    S = _lib.synth()
    nc = np.array(lp.ncell, dtype=np.int64)
    L = np.array(lp.lengths, dtype=np.float64)
    return _two_pass(S.synth_mass_rows, lp.n_own, lp.d, _p(nc), _p(L), _p(lp.elo), _p(lp.ehi), _p(lp.ext_lid), _p(lp.olo),
                     _p(lp.ohi))[:3]


def prolong_rows(fine: LevelPart, coarse: LevelPart):
    S = _lib.synth()
    return _two_pass(S.synth_prolong_rows, fine.n_own, fine.d, _p(fine.olo), _p(fine.ohi), _p(coarse.elo), _p(coarse.ehi),
                     _p(coarse.ext_lid))[:3]


def restrict_rows(fine: LevelPart, coarse: LevelPart):
    S = _lib.synth()
    return _two_pass(S.synth_restrict_rows, coarse.n_own, coarse.d, _p(coarse.olo), _p(coarse.ohi), _p(fine.elo),
                     _p(fine.ehi), _p(fine.ext_lid))[:3]


def exact_solution(lp: LevelPart):
    """nodal values of u = x + y at the own dofs."""
    X = _box_coords(lp.olo, lp.ohi).astype(np.float64)
    h = np.array(lp.lengths, dtype=np.float64) / np.array(lp.ncell, dtype=np.float64)
    return X[:, 0] * h[0] + (X[:, 1] * h[1] if lp.d > 1 else 0.0)


def lexicographic_ids(lp: LevelPart):
    """global lexicographic free-dof id (the serial numbering) of every local dof (own, then ghost)."""
    X = _box_coords(lp.elo, lp.ehi)
    keep = lp.ext_lid >= 0
    ninner = np.array(lp.ncell, dtype=np.int64) - 1
    gid = _box_lids(np.ones(lp.d, dtype=np.int64), ninner + 1, X[keep])
    out = np.empty(lp.n_own + lp.n_ghost, dtype=np.int64)
    out[lp.ext_lid[keep]] = gid
    return out


def to_scipy(rowptr, col, val, ncols):
    import scipy.sparse as sp

    return sp.csr_matrix((val, col, rowptr), shape=(rowptr.shape[0] - 1, ncols))


def empty_level_part(ncell, parts, rank, lengths=None) -> LevelPart:
    """the part of a rank that does not hold this level (levels on fewer parts, HierarchicalArrays.jl:96-149)"""
    d = len(ncell)
    z = lambda: np.zeros(d, dtype=np.int64)
    i32, i64 = (lambda n: np.zeros(n, dtype=np.int32)), (lambda n: np.zeros(n, dtype=np.int64))
    return LevelPart(tuple(int(n) for n in ncell), tuple(int(q) for q in parts), rank, z(), z(), z(), z(), i32(0), 0, 0, i32(0), i64(0),
                     tuple(float(v) for v in (lengths if lengths is not None else (1.0,) * d)), i32(0), i64(1), i64(0), i32(0), i64(1), i64(0))


def level_part_or_empty(ncell, parts, rank, lengths=None) -> LevelPart:
    return make_level_part(ncell, parts, rank, lengths) if rank < int(np.prod(parts)) else empty_level_part(ncell, parts, rank, lengths)


def _owner_of(ncell, parts, X):
    """(owner rank, owner-local id) of the free nodes X (N,d) in the Cartesian partition `parts` of the mesh `ncell`"""
    d = len(ncell)
    per = np.array([ncell[k] // parts[k] for k in range(d)], dtype=np.int64)
    q = np.minimum(X // per, np.array(parts, dtype=np.int64) - 1)
    owner = np.zeros(X.shape[0], dtype=np.int64)
    stride = 1
    for k in range(d):
        owner += q[:, k] * stride
        stride *= parts[k]
    lid = np.zeros(X.shape[0], dtype=np.int64)
    for r in np.unique(owner):
        m = owner == r
        c = part_coords(int(r), parts)
        rlo = np.array([own_range(ncell[k], parts[k], c[k])[0] for k in range(d)])
        rhi = np.array([own_range(ncell[k], parts[k], c[k])[1] for k in range(d)])
        lid[m] = _box_lids(rlo, rhi, X[m])
    return owner, lid


def redistribution_lists(ncell, src_parts, dst_parts, rank):
    """Neighbour / id lists that move the own free-dof values of the mesh `ncell` from the partition `src_parts` to the
    partition `dst_parts` (part grids over the same ranks, a grid with fewer parts leaves the last ranks empty): the
    data of MultilevelTools' RedistributionOperator (GridTransferOperators.jl:447-532) for Cartesian partitions.
    Sender and receiver derive the same order independently: per neighbour, ascending SOURCE local id.
    Returns (n_src_own, n_dst_own, nbr_snd, snd_ptrs, snd_ids, nbr_rcv, rcv_ptrs, rcv_ids)."""
    ncell = tuple(int(n) for n in ncell)
    src, dst = level_part_or_empty(ncell, src_parts, rank), level_part_or_empty(ncell, dst_parts, rank)

    def grouped(owner, key):
        order = np.lexsort((key, owner))
        nbr, counts = np.unique(owner, return_counts=True)
        return nbr.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), order.astype(np.int64)

    e32, e64 = np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64)
    nbr_snd, snd_ptrs, snd_ids = e32, np.zeros(1, dtype=np.int64), e64
    nbr_rcv, rcv_ptrs, rcv_ids = e32, np.zeros(1, dtype=np.int64), e64
    if src.n_own:
        owner, _ = _owner_of(ncell, dst_parts, _box_coords(src.olo, src.ohi))
        nbr_snd, snd_ptrs, snd_ids = grouped(owner, np.arange(src.n_own))
    if dst.n_own:
        sowner, slid = _owner_of(ncell, src_parts, _box_coords(dst.olo, dst.ohi))
        nbr_rcv, rcv_ptrs, rcv_ids = grouped(sowner, slid)
    return src.n_own, dst.n_own, nbr_snd, snd_ptrs, snd_ids, nbr_rcv, rcv_ptrs, rcv_ids


@dataclass
class HostHierarchy:
    """Host (numpy) arrays of one rank's part of the level hierarchy."""

    levels: list  # LevelPart per level
    A: list  # (rowptr,col,val) per level
    P: list
    R: list
    b: np.ndarray
    # levels on fewer parts: per level boundary l (None = same parts on both sides) the coarse space of level l+1 in
    # the partition of level l's parts (columns of P[l] / rows of R[l]) and the two redistribution list sets
    coarse_red: list = None
    to_coarse: list = None
    to_fine: list = None


def _empty_rows():
    return np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float64)


def poisson_hierarchy_host(ncell_fine, nlevels, parts=None, rank=0, lengths=None, parts_per_level=None) -> HostHierarchy:
    """parts_per_level[l]: part grid of level l (default: `parts` on every level).  A level whose grid has fewer parts
    than ranks lives on the first ranks only (ModelHierarchy's np_per_level); the transfer operators across such a
    boundary act on the coarse space partitioned like the FINER level and come with redistribution lists."""
    d = len(ncell_fine)
    parts = tuple(parts) if parts is not None else (1,) * d
    ppl = [tuple(q) for q in parts_per_level] if parts_per_level is not None else [parts] * nlevels
    assert len(ppl) == nlevels and ppl[0] == parts
    levels, As = [], []
    nc = tuple(int(n) for n in ncell_fine)
    ncs = []
    b0 = None
    for l in range(nlevels):
        lp = level_part_or_empty(nc, ppl[l], rank, lengths)
        if lp.n_own:
            rowptr, col, val, b = poisson_rows(lp)
        else:
            (rowptr, col, val), b = _empty_rows(), np.zeros(0)
        if l == 0:
            b0 = b
        levels.append(lp)
        ncs.append(nc)
        As.append((rowptr, col, val))
        if l < nlevels - 1:
            assert all(n % 2 == 0 for n in nc)
            nc = tuple(n // 2 for n in nc)
    Ps, Rs, cred, tc, tf = [], [], [], [], []
    for l in range(nlevels - 1):
        fine = levels[l]
        if ppl[l + 1] == ppl[l]:
            coarse = levels[l + 1]
            cred.append(None); tc.append(None); tf.append(None)
        else:
            coarse = level_part_or_empty(ncs[l + 1], ppl[l], rank, lengths)  # coarse space on the finer level's parts
            cred.append(coarse)
            tc.append(redistribution_lists(ncs[l + 1], ppl[l], ppl[l + 1], rank))
            tf.append(redistribution_lists(ncs[l + 1], ppl[l + 1], ppl[l], rank))
        Ps.append(prolong_rows(fine, coarse) if fine.n_own else _empty_rows())
        Rs.append(restrict_rows(fine, coarse) if coarse.n_own else _empty_rows())
    return HostHierarchy(levels, As, Ps, Rs, b0, cred, tc, tf)


@dataclass
class DeviceHierarchy:
    ctx: object
    host: HostHierarchy
    plans: list
    A: list
    P: list
    R: list
    to_coarse: list = None  # RedistributionPlan per level boundary (None: same parts)
    to_fine: list = None
    plans_red: list = None

    @property
    def redist(self):
        """the `redist=` argument of GMGLinearSolver (None when no level is redistributed)"""
        if not self.to_coarse or all(t is None for t in self.to_coarse):
            return None
        return self.to_coarse, self.to_fine


def upload_hierarchy(ctx, hh: HostHierarchy) -> DeviceHierarchy:
    """numerical_setup-side upload: PSparseMatrix mirrors + exchange plans per level (+ redistribution plans where a
    level lives on fewer parts).  Every rank creates every plan, in the same order (plan creation is collective)."""
    from .api import ExchangePlan, RedistributionPlan, SparseMatrix

    def mkplan(lp):
        if ctx.nranks == 1:
            return None
        return ExchangePlan(ctx, lp.n_own, lp.n_ghost, lp.nbr_snd, lp.snd_ptrs, lp.snd_ids, lp.nbr_rcv, lp.rcv_ptrs, lp.rcv_ids)

    plans, As, Ps, Rs = [], [], [], []
    for lp, (rp, c, v) in zip(hh.levels, hh.A):
        plan = mkplan(lp)
        plans.append(plan)
        As.append(SparseMatrix(ctx, lp.n_own, lp.n_own, lp.n_ghost, rp, c, v, plan=plan))
    nb = len(hh.P)
    cred = hh.coarse_red or [None] * nb
    to_coarse, to_fine, plans_red = [None] * nb, [None] * nb, [None] * nb
    for l, ((rp, c, v), (rr, rc, rv)) in enumerate(zip(hh.P, hh.R)):
        f, co, pco = hh.levels[l], hh.levels[l + 1], plans[l + 1]
        if cred[l] is not None:
            co = cred[l]
            pco = plans_red[l] = mkplan(co)
            to_coarse[l] = RedistributionPlan(ctx, *hh.to_coarse[l])
            to_fine[l] = RedistributionPlan(ctx, *hh.to_fine[l])
        Ps.append(SparseMatrix(ctx, f.n_own, co.n_own, co.n_ghost, rp, c, v, plan=pco))
        Rs.append(SparseMatrix(ctx, co.n_own, f.n_own, f.n_ghost, rr, rc, rv, plan=plans[l]))
    return DeviceHierarchy(ctx, hh, plans, As, Ps, Rs, to_coarse, to_fine, plans_red)


# ------------------------------------------------------------------------------------------------
# Q_p tensor-product Lagrange problems on uniform Cartesian meshes (C4: 3D Q2 vector-valued elasticity,
# C5: 2D Q2-P1disc Stokes).  The element matrices are computed here (numpy), the assembled rows by the
# generic C generator synth_fe_rows (one element matrix, all cells congruent) -- no COO intermediate, so
# the full-size C4 system (64^3 cells, 6.4 M dofs, 1.24e9 non-zeros) is generated in seconds.
# Numbering: nodes lexicographic over the (p*n+1)^d node grid (x fastest), node-major vector dofs, free
# dofs = non-Dirichlet nodes in lexicographic order (for Q2 a permutation of Gridap's vertex/edge/face/
# interior numbering).  tests/ cross-check these generators against the oracle's own
# element-by-element assembly (oracle/fem.py).


def _gauss01(npts):
    x, w = np.polynomial.legendre.leggauss(npts)
    return 0.5 * (x + 1.0), 0.5 * w


def _lagrange_1d(order, xi):
    """values N[q,a] and derivatives dN[q,a] of the equispaced Lagrange basis of degree `order` on [0,1]"""
    nodes = np.linspace(0.0, 1.0, order + 1)
    xi = np.asarray(xi, dtype=np.float64)
    N = np.ones((xi.shape[0], order + 1))
    dN = np.zeros((xi.shape[0], order + 1))
    for a in range(order + 1):
        others = [b for b in range(order + 1) if b != a]
        den = np.prod([nodes[a] - nodes[b] for b in others])
        N[:, a] = np.prod([xi - nodes[b] for b in others], axis=0) / den
        for c in others:
            dN[:, a] += np.prod([xi - nodes[b] for b in others if b != c], axis=0) / den if len(others) > 1 else 1.0 / den
    return N, dN


def _tensor_basis(order, d, h, nq):
    """phi[q,a], grad[q,a,dim], w[q] of the Q_order basis on a cell of size h; q and a lexicographic, x fastest"""
    xi, wq = _gauss01(nq)
    N1, dN1 = _lagrange_1d(order, xi)
    p1 = order + 1
    qi = np.stack(np.meshgrid(*[np.arange(nq)] * d, indexing="ij"), axis=-1).reshape(-1, d)[:, ::-1]  # x fastest
    ai = np.stack(np.meshgrid(*[np.arange(p1)] * d, indexing="ij"), axis=-1).reshape(-1, d)[:, ::-1]
    # reorder so that the FIRST listed index varies fastest: build explicit lexicographic (x fastest) lists
    qi = np.array([[(Q // nq**k) % nq for k in range(d)] for Q in range(nq**d)])
    ai = np.array([[(A // p1**k) % p1 for k in range(d)] for A in range(p1**d)])
    w = np.ones(nq**d)
    for k in range(d):
        w *= wq[qi[:, k]] * h[k]
    phi = np.ones((nq**d, p1**d))
    grad = np.ones((nq**d, p1**d, d))
    for k in range(d):
        Nk = N1[qi[:, k]][:, ai[:, k]]
        dNk = dN1[qi[:, k]][:, ai[:, k]] / h[k]
        phi *= Nk
        for g in range(d):
            grad[:, :, g] *= dNk if g == k else Nk
    return phi, grad, w


def _node_grid(ncell, order):
    nn = tuple(order * n + 1 for n in ncell)
    grids = np.meshgrid(*[np.arange(n) for n in nn], indexing="ij")
    mi = np.stack([g.ravel(order="F") for g in grids], axis=1)  # (nnodes, d), x fastest
    return nn, mi


def fe_rows(ncell, order, ncomp, Ke, dirichlet_nodes, Fe=None, ud=None):
    """assembled rows (rowptr, col int32, val, b, n_free_dofs, free_node_ids) of the free dofs"""
    S = _lib.synth()
    d = len(ncell)
    nc = np.array(ncell, dtype=np.int64)
    free_nodes = np.flatnonzero(~dirichlet_nodes).astype(np.int64)
    node_free = np.full(dirichlet_nodes.shape[0], -1, dtype=np.int32)
    node_free[free_nodes] = np.arange(free_nodes.shape[0], dtype=np.int32)
    n = free_nodes.shape[0] * ncomp
    Ke = np.ascontiguousarray(Ke, dtype=np.float64)
    Fe = None if Fe is None else np.ascontiguousarray(Fe, dtype=np.float64)
    ud = None if ud is None else np.ascontiguousarray(ud, dtype=np.float64)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    args = (d, order, ncomp, _p(nc), _p(Ke), _p(Fe), _p(node_free), free_nodes.shape[0], _p(free_nodes), _p(ud))
    S.synth_fe_rows(*args, 0, _p(rowptr), None, None, None)
    np.cumsum(rowptr, out=rowptr)
    nnz = int(rowptr[-1])
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    b = np.zeros(n)
    S.synth_fe_rows(*args, 1, _p(rowptr), _p(col), _p(val), _p(b))
    return rowptr, col, val, b, n, free_nodes


def elasticity_element(h, order=2, lam=1.0, mu=1.0, body=(0.0, 0.0, -1.0)):
    """element matrix / load of a(u,v) = int lam div(u) div(v) + 2 mu eps(u):eps(v)  (form of
    test/Applications/Elasticity.jl:31-37; lam = mu = 1 per SURVEY.md 8d), node-major local dofs"""
    d = len(h)
    phi, G, w = _tensor_basis(order, d, h, order + 1)
    nb = phi.shape[1]
    Ke = np.zeros((nb * d, nb * d))
    GG = np.einsum("q,qad,qbd->ab", w, G, G)
    for c1 in range(d):
        for c2 in range(d):
            blk = lam * np.einsum("q,qa,qb->ab", w, G[:, :, c1], G[:, :, c2]) + mu * np.einsum("q,qa,qb->ab", w, G[:, :, c2], G[:, :, c1])
            if c1 == c2:
                blk = blk + mu * GG
            Ke[c1::d, c2::d] = blk
    Fe = np.zeros(nb * d)
    for c in range(d):
        Fe[c::d] = body[c] * (w @ phi)
    return Ke, Fe


def elasticity_rows(ncell, order=2, lam=1.0, mu=1.0, body=(0.0, 0.0, -1.0)):
    """C4 system: (rowptr, col, val, b, n) -- clamped on the face x = 0, constant body force"""
    d = len(ncell)
    h = tuple(1.0 / n for n in ncell)
    Ke, Fe = elasticity_element(h, order, lam, mu, body)
    nn, mi = _node_grid(ncell, order)
    rp, col, val, b, n, _ = fe_rows(ncell, order, d, Ke, mi[:, 0] == 0, Fe=Fe)
    return rp, col, val, b, n


def _prolong_1d(order, nc_coarse):
    """1D nodal interpolation from n cells to 2n cells of the degree-`order` Lagrange space (dense small rows)"""
    import scipy.sparse as sp

    nf, ncn = 2 * order * nc_coarse + 1, order * nc_coarse + 1
    rows, cols, vals = [], [], []
    for i in range(nf):
        c = min(i // (2 * order), nc_coarse - 1)
        xi = (i - c * 2 * order) / (2.0 * order)  # exact dyadic reference coordinate
        N, _ = _lagrange_1d(order, np.array([xi]))
        for a in range(order + 1):
            if N[0, a] != 0.0:
                rows.append(i)
                cols.append(c * order + a)
                vals.append(N[0, a])
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, ncn))


def fe_prolongation(ncell_coarse, order, ncomp, free_nodes_fine, free_nodes_coarse):
    """P (fine free dofs x coarse free dofs): nodal interpolation with zero Dirichlet values
    (src/MultilevelTools/GridTransferOperators.jl:391-401), CSR with ascending columns; R = P^T"""
    import scipy.sparse as sp

    P = _prolong_1d(order, ncell_coarse[0])
    for k in range(1, len(ncell_coarse)):
        P = sp.kron(_prolong_1d(order, ncell_coarse[k]), P, format="csr")  # x fastest
    P = P[free_nodes_fine][:, free_nodes_coarse]
    if ncomp > 1:
        P = sp.kron(P, sp.identity(ncomp), format="csr")
    P = sp.csr_matrix(P)
    P.eliminate_zeros()
    P.sort_indices()
    R = P.T.tocsr()
    R.sort_indices()
    return P, R


def _triplet(A):
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64)


@dataclass
class SerialLevel:
    """level descriptor of a single-part (serial) FE hierarchy: what upload_hierarchy needs"""
    n_own: int
    n_ghost: int = 0
    ncell: tuple = None


def elasticity_hierarchy_host(ncell_fine, nlevels, order=2, lam=1.0, mu=1.0, body=(0.0, 0.0, -1.0)) -> HostHierarchy:
    """C4 hierarchy: factor-2 nested meshes, re-discretised level matrices
    (src/MultilevelTools/FESpaceHierarchies.jl:151-174), nodal prolongations, R = P^T"""
    d = len(ncell_fine)
    nc = tuple(int(n) for n in ncell_fine)
    levels, As, frees, b0 = [], [], [], None
    for l in range(nlevels):
        h = tuple(1.0 / n for n in nc)
        Ke, Fe = elasticity_element(h, order, lam, mu, body)
        nn, mi = _node_grid(nc, order)
        rp, col, val, b, n, free_nodes = fe_rows(nc, order, d, Ke, mi[:, 0] == 0, Fe=Fe)
        if l == 0:
            b0 = b
        levels.append(SerialLevel(n, 0, nc))
        As.append((rp, col, val))
        frees.append(free_nodes)
        if l < nlevels - 1:
            assert all(n % 2 == 0 for n in nc), "factor-2 coarsening needs even cell counts"
            nc = tuple(n // 2 for n in nc)
    Ps, Rs = [], []
    for l in range(nlevels - 1):
        P, R = fe_prolongation(levels[l + 1].ncell, order, d, frees[l], frees[l + 1])
        Ps.append(_triplet(P))
        Rs.append(_triplet(R))
    return HostHierarchy(levels, As, Ps, Rs, b0)


def stokes_cavity_host(ncell, nlevels=1):
    """C5: 2D lid-driven cavity, Q2 velocity / P1-discontinuous pressure (joss_paper/demo.jl:20-91):
    A (int grad u : grad v), B (-(div v) p; pressure rows), Bt, pressure mass Mp, rhs from the lid data
    u = (1,0) on y = 1; with nlevels > 1 the velocity-block hierarchy (mats, P, R) for the GMG block.
    Everything as CSR triplets (rowptr int64, col int32, val)."""
    import scipy.sparse as sp

    d, order = 2, 2
    nc = tuple(int(n) for n in ncell)
    out = {}
    mats, frees, cells = [], [], []
    for l in range(nlevels):
        h = tuple(1.0 / n for n in nc)
        phi, G, w = _tensor_basis(order, d, h, 3)
        Ks = np.einsum("q,qad,qbd->ab", w, G, G)
        nb = phi.shape[1]
        Ke = np.zeros((nb * d, nb * d))
        for c in range(d):
            Ke[c::d, c::d] = Ks
        nn, mi = _node_grid(nc, order)
        bnd = ((mi == 0) | (mi == np.array(nn) - 1)).any(axis=1)
        ud = None
        if l == 0:
            ud = np.zeros((mi.shape[0], d))
            ud[mi[:, 1] == nn[1] - 1, 0] = 1.0
            ud = ud.ravel()
        rp, col, val, b, n, free_nodes = fe_rows(nc, order, d, Ke, bnd, ud=ud)
        mats.append((rp, col, val))
        frees.append(free_nodes)
        cells.append(nc)
        if l == 0:
            out.update(A=(rp, col, val), fu=b, n_u=n)
            # pressure blocks, cell by cell: P1disc basis {1, xi-1/2, eta-1/2} in reference coordinates
            xi, _ = _gauss01(3)
            qx, qy = np.tile(xi, 3), np.repeat(xi, 3)
            psi = np.stack([np.ones(9), qx - 0.5, qy - 0.5], axis=1)
            Be = np.zeros((3, nb * d))
            for c in range(d):
                Be[:, c::d] = -np.einsum("q,qm,qa->ma", w, psi, G[:, :, c])
            Mpe = np.einsum("q,qm,qn->mn", w, psi, psi)
            ncell_tot = nc[0] * nc[1]
            cx, cy = np.arange(ncell_tot) % nc[0], np.arange(ncell_tot) // nc[0]
            loc = np.arange(nb)
            conn = ((cy[:, None] * order + loc[None, :] // (order + 1)) * nn[0] + cx[:, None] * order + loc[None, :] % (order + 1))  # ascending
            vdof = (conn[:, :, None] * d + np.arange(d)[None, None, :]).reshape(ncell_tot, -1)  # global (node, comp) ids
            node_free = np.full(mi.shape[0], -1, dtype=np.int64)
            node_free[free_nodes] = np.arange(free_nodes.shape[0])
            fdof = node_free[conn][:, :, None] * d + np.arange(d)[None, None, :]
            fdof = np.where(node_free[conn][:, :, None] >= 0, fdof, -1).reshape(ncell_tot, -1)
            rows = np.repeat(np.arange(ncell_tot * 3).reshape(ncell_tot, 3), nb * d, axis=1).ravel()
            cols = np.tile(fdof, (1, 3)).ravel()
            vals = np.tile(Be.ravel(), ncell_tot)
            keep = cols >= 0
            B = sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(ncell_tot * 3, n))
            B.sort_indices()
            # fp = -B[:, dirichlet] ud
            gcols = np.tile(vdof, (1, 3)).ravel()
            fp = np.zeros(ncell_tot * 3)
            np.subtract.at(fp, rows[~keep], vals[~keep] * ud[gcols[~keep]])
            Mp = sp.kron(sp.identity(ncell_tot), Mpe, format="csr")
            Mp.sort_indices()
            Bt = B.T.tocsr()
            Bt.sort_indices()
            out.update(B=_triplet(B), Bt=_triplet(Bt), Mp=_triplet(Mp), fp=fp, n_p=ncell_tot * 3)
        if l < nlevels - 1:
            nc = tuple(n // 2 for n in nc)
    if nlevels > 1:
        Ps, Rs = [], []
        for l in range(nlevels - 1):
            P, R = fe_prolongation(cells[l + 1], order, d, frees[l], frees[l + 1])
            Ps.append(_triplet(P))
            Rs.append(_triplet(R))
        out.update(mats=mats, P=Ps, R=Rs)
    return out
