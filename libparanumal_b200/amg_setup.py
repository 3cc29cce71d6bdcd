"""Host-side setup of the algebraic levels below degree 1 (parAlmond::parAlmond_t::AMGSetup,
libs/parAlmond/parAlmondAMGSetup.cpp:34-144) so that the MULTIGRID preconditioner can be built without the
reference at hand: strength of connection, MIS-2 aggregation, tentative + smoothed prolongator, Galerkin
product, rho(D^-1 A) estimates, dense coarse inverse.  Setup is host work in the reference too (C++ on
memory<T> arrays); the products cross the C ABI as CSR arrays (libp_csr_create / libp_parcsr_create) and the
apply path (V-cycle) runs in CUDA.

Design choice for P > 1 ranks: the hierarchy is built from the *global* degree-1 matrix with the reference's
one-rank algorithm, replicated on every rank, and then split into row blocks.  The reference aggregates
rank-locally, so its aggregates stop at rank boundaries; here they do not.  The global matrix is in the run's
DOF numbering (rank offset + gathered index), which is a permutation of the one-rank numbering, and the
random keys are drawn by index - so the aggregates of a P-rank run differ in detail from the one-rank ones
(iteration counts agree within +-1 in every test).  The one-rank hierarchy is pinned against the reference's
own dumps (tests/test_amg_setup_cpu.py); the distributed levels are pinned against scipy products of the same
global matrices (tests/multigpu_check.py).

Random numbers: the reference draws from glibc drand48 seeded by srand48(rank)
(libs/parAlmond/parAlmondKernels.cpp:56-58); Drand48 below restates that generator so the aggregates and
the Arnoldi start vectors are the reference's.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

_A48 = np.uint64(0x5DEECE66D)
_C48 = np.uint64(0xB)
_M48 = np.uint64((1 << 48) - 1)


class Drand48:
    """glibc srand48/drand48: X <- (0x5DEECE66D X + 0xB) mod 2^48, X0 = (seed << 16) | 0x330E, value X / 2^48."""

    def __init__(self, seed=0):
        self.x = np.uint64(((int(seed) & 0xFFFFFFFF) << 16) | 0x330E)

    def draw(self, n):
        n = int(n)
        if n == 0:
            return np.zeros(0)
        with np.errstate(over="ignore"):
            # X_k = a^k X_0 + c (1 + a + ... + a^(k-1)); uint64 wrap-around keeps the low 48 bits exact
            ak = np.cumprod(np.full(n, _A48, dtype=np.uint64))
            sk = np.empty(n, dtype=np.uint64)
            sk[0] = 1
            if n > 1:
                sk[1:] = np.cumsum(ak[:-1], dtype=np.uint64) + np.uint64(1)
            xs = (ak * self.x + _C48 * sk) & _M48
        self.x = xs[-1]
        return xs.astype(np.float64) * (1.0 / 281474976710656.0)


def arnoldi_rho(apply, v0, n_total, dot=None, k=10):
    """10-step Arnoldi estimate of rho(D^-1 A) (parCSR::rhoDinvA, parAlmondparCSR.cpp:365-447, and
    MGLevel::maxEigSmoothAx, ellipticPreconMultiGridLevel.cpp:393-473).  apply(v) = D^-1 A v on local rows
    (must return a new vector), v0 = start vector (numpy or torch), dot(a, b) = global inner product."""
    k = int(min(k, n_total))
    if dot is None:
        # einsum, not np.dot: a threaded BLAS-1 call on a busy host costs tens of ms for what is a 1 ms loop
        dot = lambda a, b: float(np.einsum("i,i->", a, b))
    H = np.zeros((k, k))
    V = [v0 * (1.0 / np.sqrt(dot(v0, v0)))]
    for j in range(k):
        w = apply(V[j])
        for i in range(j + 1):
            hij = dot(V[i], w)
            w = w - hij * V[i]
            H[i, j] = hij
        if j + 1 < k:
            nrm = np.sqrt(dot(w, w))
            H[j + 1, j] = nrm
            w = w * (1.0 / nrm)
        V.append(w)
    return float(np.max(np.abs(np.linalg.eigvals(H)))) if k else 0.0


def element_matrix_triplets(Nq, ggeo, wJ, D, lam, gid, threshold=1e-7):
    """Unassembled non-zeros of the continuous hex operator (BuildOperatorMatrixContinuousHex3D,
    solvers/elliptic/src/ellipticBuildOperatorMatrixContinuous.cpp:697-800): per element
    A_e = sum_ab D_a^T diag(G_ab) D_b + lambda diag(wJ) with D_r = I x I x D, D_s = I x D x I, D_t = D x I x I;
    rows/columns of masked nodes (gid < 0) and entries with |val| < threshold are dropped.
    ggeo [E][6][Np], wJ [E][Np], D [Nq][Nq], gid [E*Np] are torch tensors on one device; returns numpy
    (global row, global col, value) triplets."""
    import torch
    Np = Nq * Nq * Nq
    E = wJ.numel() // Np
    dev = ggeo.device
    D = D.reshape(Nq, Nq)
    I = torch.eye(Nq, dtype=torch.float64, device=dev)
    kron3 = lambda a, b, c: torch.kron(a, torch.kron(b, c))   # node index = x + y*Nq + z*Nq^2
    Dd = [kron3(I, I, D), kron3(I, D, I), kron3(D, I, I)]      # d/dr, d/ds, d/dt on the element nodes
    G = ggeo.reshape(E, 6, Np)
    comp = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 1): 3, (1, 2): 4, (2, 2): 5}
    gid = gid.reshape(E, Np).long()
    rows, cols, vals = [], [], []
    chunk = max(1, (1 << 24) // (Np * Np))
    for e0 in range(0, E, chunk):
        e1 = min(E, e0 + chunk)
        Ae = torch.zeros(e1 - e0, Np, Np, dtype=torch.float64, device=dev)
        for a in range(3):
            for b in range(3):
                g = G[e0:e1, comp[(min(a, b), max(a, b))]]
                Ae += torch.einsum("pn,ep,pm->enm", Dd[a], g, Dd[b])
        Ae += torch.diag_embed(wJ.reshape(E, Np)[e0:e1] * lam)
        ge = gid[e0:e1]
        keep = (ge[:, :, None] >= 0) & (ge[:, None, :] >= 0) & (Ae.abs() >= threshold)
        rows.append(ge[:, :, None].expand(-1, -1, Np)[keep].cpu().numpy())
        cols.append(ge[:, None, :].expand(-1, Np, -1)[keep].cpu().numpy())
        vals.append(Ae[keep].cpu().numpy())
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def strong_graph(A, theta, kind="SYMMETRIC"):
    """strongGraph (parAlmondStrongGraph.cpp:36-237): boolean CSR, diagonal always kept."""
    A = A.tocsr()
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    cols, vals = A.indices, A.data
    d = A.diagonal()
    if kind == "SYMMETRIC":
        keep = np.abs(vals) > theta * np.sqrt(np.abs(d[rows]) * np.abs(d[cols]))
    else:  # RUGESTUBEN
        sign = np.where(d >= 0, 1.0, -1.0)
        od = -sign[rows] * vals
        odm = np.where(rows == cols, 0.0, od)
        mx = np.zeros(n)
        np.maximum.at(mx, rows, odm)
        keep = od > theta * mx[rows]
    keep |= rows == cols
    C = sp.csr_matrix((np.ones(int(keep.sum()), dtype=np.int8), (rows[keep], cols[keep])), shape=A.shape)
    C.sort_indices()
    return C


def _lexmax_keys(ptr, cols, key):
    """per row of the graph: the largest key over the row's columns (the diagonal entry stands for the row itself)"""
    return np.maximum.reduceat(key[cols], ptr[:-1])


def form_aggregates(C, rng):
    """formAggregates (parAlmondFormAggregates.cpp:56-249): distance-2 maximal independent set on the strong
    graph, keyed by (#strong connections in the column + drand48, global id); returns FineToCoarse[N] and the
    number of aggregates.  Aggregates are numbered in ascending order of their root node.

    The reference compares (state, rand, id) triples lexicographically (customLess, parAlmondFormAggregates.cpp:34-48).
    (rand, id) is a fixed property of a node, so the triples are encoded as ONE integer key
    (state + 1) * n + rank_of(rand, id): one gather and one segmented maximum per sweep instead of three, with the
    identical order (ties in rand are broken by id in the rank)."""
    n = C.shape[0]
    ptr, cols = C.indptr, C.indices
    rands = rng.draw(n) + np.bincount(cols, minlength=n)
    ids = np.arange(n, dtype=np.int64)
    order = np.lexsort((ids, rands))            # ascending (rand, id)
    rank = np.empty(n, dtype=np.int64)
    rank[order] = ids
    states = np.zeros(n, dtype=np.int64)

    def encode(st, node):
        return (st + 1) * n + rank[node]

    def decode(key):
        return key // n - 1, order[key % n]     # (state, node that carries the winning (rand, id))

    while True:
        T = _lexmax_keys(ptr, cols, encode(states, ids))
        smax, imax = decode(_lexmax_keys(ptr, cols, T))
        und = states == 0
        new_mis = und & (imax == ids)
        states[new_mis] = 1
        states[und & ~new_mis & (smax == 1)] = -1
        if not np.any(states == 0):
            break
    roots = np.flatnonzero(states == 1)
    f2c = np.full(n, -1, dtype=np.int64)
    f2c[roots] = np.arange(roots.size)
    # first ring: adopt the aggregate of the strongest neighbour when that neighbour is a root
    T = _lexmax_keys(ptr, cols, encode(states, ids))
    Ts, Ti = decode(T)
    Tc = f2c[Ti]                       # aggregate of the winning node
    take = (f2c == -1) & (Ts == 1) & (Tc > -1)
    f2c1 = f2c.copy()
    f2c1[take] = Tc[take]
    # second ring, on the first-ring winners: when the strongest triple in the row carries state 1, its node is a root
    s2, i2 = decode(_lexmax_keys(ptr, cols, T))
    c2 = f2c[i2]
    take2 = (f2c1 == -1) & (s2 == 1) & (c2 > -1)
    f2c1[take2] = c2[take2]
    return f2c1, int(roots.size), roots


def tentative_prolongator(f2c, nagg, null):
    """tentativeProlongator (parAlmondTentativeProlongator.cpp:34-94): T[n, f2c[n]] = null[n], columns
    normalised; returns (T, coarse null vector)."""
    n = f2c.size
    cn = np.sqrt(np.bincount(f2c, weights=null * null, minlength=nagg))
    T = sp.csr_matrix((null / cn[f2c], (np.arange(n), f2c)), shape=(n, nagg))
    return T, cn


def setup_hierarchy(A, null, rng, coarse_target=1000, strength="SYMMETRIC", aggregation="SMOOTHED"):
    """parAlmond_t::AMGSetup on one (global) matrix.  Returns (levels, coarseA, coarseRho, coarseNull): levels =
    list of dicts {A, P, R, rho, roots} finest first (every level that has a prolongator), then the matrix of the
    exact solve, its rho and the null vector carried down to it (exactSolver_t::setup adds
    nullSpacePenalty * null null^T for all-Neumann problems, parAlmondCoarseExact.cpp:158-164)."""
    A = sp.csr_matrix(A)
    A.sort_indices()

    def rho_of(M):
        dinv = M.diagonal()
        dinv = np.where(dinv != 0.0, 1.0 / np.where(dinv != 0.0, dinv, 1.0), 0.0)
        return arnoldi_rho(lambda v: dinv * (M @ v), rng.draw(M.shape[0]), M.shape[0])

    rho = rho_of(A)
    levels = []
    n = A.shape[0]
    null = np.asarray(null, dtype=np.float64).copy()
    if n <= coarse_target:
        return levels, A, rho, null
    theta = 0.5 if strength == "RUGESTUBEN" else 0.08
    while True:
        C = strong_graph(A, theta, strength)
        f2c, nagg, roots = form_aggregates(C, rng)
        T, null = tentative_prolongator(f2c, nagg, null)
        if aggregation == "SMOOTHED":  # smoothProlongator: P = (I - omega D^-1 A) T, omega = (4/3)/rho
            omega = (4.0 / 3.0) / rho
            dinv = sp.diags(1.0 / A.diagonal())
            P = (T - omega * (dinv @ (A @ T))).tocsr()
        else:
            P = T
        P.sort_indices()
        R = P.T.tocsr()
        R.sort_indices()
        Ac = (R @ (A @ P)).tocsr()
        Ac.sort_indices()
        levels.append(dict(A=A, P=P, R=R, rho=rho, roots=roots))
        rho = rho_of(Ac)
        if strength == "SYMMETRIC":
            theta *= 0.5
        nc = Ac.shape[0]
        A = Ac
        if nc <= coarse_target or n < 2 * nc:
            return levels, A, rho, null
        n = nc


# --------------------------------------------------------------------------------------------- distribution
def split_rows(M, row_starts, col_starts, rank):
    """Row block of a global CSR matrix in the reference's parCSR form (parAlmondparCSR.hpp:43-92): local
    `diag` CSR + off-rank MCSR block with compressed rows and columns renumbered behind the local ones."""
    r0, r1 = int(row_starts[rank]), int(row_starts[rank + 1])
    c0, c1 = int(col_starts[rank]), int(col_starts[rank + 1])
    L = M[r0:r1].tocsr()
    L.sort_indices()
    rows = np.repeat(np.arange(r1 - r0), np.diff(L.indptr))
    loc = (L.indices >= c0) & (L.indices < c1)
    diag = sp.csr_matrix((L.data[loc], (rows[loc], L.indices[loc] - c0)), shape=(r1 - r0, c1 - c0))
    diag.sort_indices()
    orow, ocol, oval = rows[~loc], L.indices[~loc].astype(np.int64), L.data[~loc]
    colIds, inv = np.unique(ocol, return_inverse=True)
    nzr, rinv = np.unique(orow, return_inverse=True)
    order = np.lexsort((inv, rinv))
    counts = np.bincount(rinv, minlength=nzr.size)
    return dict(Nrows=r1 - r0, NlocalCols=c1 - c0,
                diag_rowStarts=diag.indptr.astype(np.int32), diag_cols=diag.indices.astype(np.int32),
                diag_vals=np.ascontiguousarray(diag.data, dtype=np.float64),
                offd_rows=nzr.astype(np.int32),
                offd_mRowStarts=np.concatenate([[0], np.cumsum(counts)]).astype(np.int32),
                offd_cols=((c1 - c0) + inv[order]).astype(np.int32),
                offd_vals=np.ascontiguousarray(oval[order], dtype=np.float64),
                offd_colIds=colIds.astype(np.int64), globalColStarts=np.asarray(col_starts, dtype=np.int64))


def coarse_partition(fine_starts, roots):
    """Aggregates are numbered by ascending root node, so 'an aggregate lives where its root lives' is a
    contiguous partition: starts[r] = number of roots below fine_starts[r]."""
    return np.searchsorted(roots, np.asarray(fine_starts), side="left").astype(np.int64)
