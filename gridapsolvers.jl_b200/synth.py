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


@dataclass
class HostHierarchy:
    """Host (numpy) arrays of one rank's part of the level hierarchy."""

    levels: list  # LevelPart per level
    A: list  # (rowptr,col,val) per level
    P: list
    R: list
    b: np.ndarray


def poisson_hierarchy_host(ncell_fine, nlevels, parts=None, rank=0, lengths=None) -> HostHierarchy:
    d = len(ncell_fine)
    parts = tuple(parts) if parts is not None else (1,) * d
    levels, As = [], []
    nc = tuple(int(n) for n in ncell_fine)
    b0 = None
    for l in range(nlevels):
        lp = make_level_part(nc, parts, rank, lengths)
        rowptr, col, val, b = poisson_rows(lp)
        if l == 0:
            b0 = b
        levels.append(lp)
        As.append((rowptr, col, val))
        if l < nlevels - 1:
            assert all(n % 2 == 0 for n in nc)
            nc = tuple(n // 2 for n in nc)
    Ps = [prolong_rows(levels[l], levels[l + 1]) for l in range(nlevels - 1)]
    Rs = [restrict_rows(levels[l], levels[l + 1]) for l in range(nlevels - 1)]
    return HostHierarchy(levels, As, Ps, Rs, b0)


@dataclass
class DeviceHierarchy:
    ctx: object
    host: HostHierarchy
    plans: list
    A: list
    P: list
    R: list


def upload_hierarchy(ctx, hh: HostHierarchy) -> DeviceHierarchy:
    """numerical_setup-side upload: PSparseMatrix mirrors + exchange plans per level."""
    from .api import ExchangePlan, SparseMatrix

    plans, As, Ps, Rs = [], [], [], []
    for lp, (rp, c, v) in zip(hh.levels, hh.A):
        plan = None
        if ctx.nranks > 1:
            plan = ExchangePlan(ctx, lp.n_own, lp.n_ghost, lp.nbr_snd, lp.snd_ptrs, lp.snd_ids, lp.nbr_rcv, lp.rcv_ptrs, lp.rcv_ids)
        plans.append(plan)
        As.append(SparseMatrix(ctx, lp.n_own, lp.n_own, lp.n_ghost, rp, c, v, plan=plan))
    for l, ((rp, c, v), (rr, rc, rv)) in enumerate(zip(hh.P, hh.R)):
        f, co = hh.levels[l], hh.levels[l + 1]
        Ps.append(SparseMatrix(ctx, f.n_own, co.n_own, co.n_ghost, rp, c, v, plan=plans[l + 1]))
        Rs.append(SparseMatrix(ctx, co.n_own, f.n_own, f.n_ghost, rr, rc, rv, plan=plans[l]))
    return DeviceHierarchy(ctx, hh, plans, As, Ps, Rs)
