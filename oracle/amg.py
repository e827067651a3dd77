"""CPU restatement of the product's own preconditioner (TEST INFRASTRUCTURE).

The reference applies one hypre-BoomerAMG V-cycle on the block-diagonal matrix P
(/root/reference/src/CGx/KNPEMI/KNPEMIx_solver.py:269-273,386).  hypre is not
vendored/installable, and the north star asks for "block-Jacobi or smoothed
aggregation" on the GPU, so the product implements smoothed-aggregation AMG
(MIS(2) aggregation, Jacobi-smoothed constant prolongator, Galerkin coarse
operators, weighted-Jacobi/Chebyshev smoothing, dense coarsest solve).  This file
restates *that* algorithm in numpy/scipy so that (i) the GPU hierarchy can be
checked level by level and (ii) the CPU baseline solves with the same algorithm.
It is not a restatement of hypre; iteration counts are ours, not the reference's.
"""
import numpy as np
import scipy.sparse as sp


def _hash32(i):
    """Deterministic integer hash (same mixing as the C++/CUDA side)."""
    x = (i.astype(np.uint64) + np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(13))) * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x = x ^ (x >> np.uint64(16))
    return x.astype(np.int64)


def strength_graph(A, theta):
    """Symmetric strength of connection |a_ij| >= theta*sqrt(|a_ii a_jj|), no diagonal."""
    A = A.tocsr()
    d = np.abs(A.diagonal())
    C = A.tocoo()
    keep = (C.row != C.col) & (np.abs(C.data) >= theta * np.sqrt(d[C.row] * d[C.col])) & (C.data != 0.0)
    S = sp.csr_matrix((np.ones(keep.sum(), np.int8), (C.row[keep], C.col[keep])), shape=A.shape)
    S = ((S + S.T) > 0).astype(np.int8).tocsr()
    S.sort_indices()
    return S


def _nbr_max(S, key):
    """max of key over the closed neighbourhood of each node."""
    n = S.shape[0]
    out = key.copy()
    deg = np.diff(S.indptr)
    nz = deg > 0
    if S.nnz:
        red = np.maximum.reduceat(key[S.indices], S.indptr[:-1][nz])
        out[nz] = np.maximum(out[nz], red)
    return out


def mis2_aggregate(S):
    """MIS(2)-based aggregation (Bell/Dalton/Olson 2012). Returns (agg id per node, n_agg).
    Key = state * 2^52 + hash * 2^... packed into int64 so one max-propagation does the job."""
    n = S.shape[0]
    idx = np.arange(n, dtype=np.int64)
    pr = ((_hash32(idx) & np.int64(0x3FFFFFFF)) << np.int64(31)) | idx     # unique priority < 2^61
    state = np.zeros(n, np.int64)                      # 0 undecided, 1 in MIS, -1 removed
    BIG = np.int64(1) << np.int64(62)
    while True:
        und = state == 0
        if not und.any():
            break
        key = np.where(state == 1, BIG + pr, np.where(und, pr, np.int64(-1)))
        k1 = _nbr_max(S, key)
        k2 = _nbr_max(S, k1)
        new_mis = und & (k2 == key)
        state[new_mis] = 1
        # anything undecided within distance 2 of a MIS node is removed
        key = np.where(state == 1, BIG + pr, np.int64(-1))
        k2 = _nbr_max(S, _nbr_max(S, key))
        state[(state == 0) & (k2 >= BIG)] = -1
    roots = np.flatnonzero(state == 1)
    agg = np.full(n, -1, np.int64)
    agg[roots] = np.arange(roots.size)
    # distance-1 then distance-2 nodes join the neighbouring aggregate with the largest root priority
    for _ in range(2):
        key = np.where(agg >= 0, (pr[roots[np.maximum(agg, 0)]]), np.int64(-1))
        best = _nbr_max(S, key)
        join = (agg < 0) & (best >= 0)
        # recover aggregate id from the root priority (low 31 bits = root node index)
        root_node = best[join] & np.int64((1 << 31) - 1)
        agg[join] = agg[root_node]
    left = np.flatnonzero(agg < 0)                     # isolated nodes: singletons
    agg[left] = roots.size + np.arange(left.size)
    return agg, roots.size + left.size


def sa_level(A, theta=0.08, omega=4.0 / 3.0):
    """One smoothed-aggregation coarsening: returns (P, R, A_c, rho) with constant near-nullspace."""
    A = A.tocsr()
    n = A.shape[0]
    S = strength_graph(A, theta)
    agg, nagg = mis2_aggregate(S)
    # unnormalised tentative prolongator: the constant stays the near-nullspace vector on every level
    T = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nagg))
    dinv = 1.0 / A.diagonal()
    rho = float(np.max(np.abs(dinv) * np.asarray(np.abs(A).sum(axis=1)).ravel()))   # Gershgorin bound on rho(D^-1 A)
    P = (T - sp.diags((omega / rho) * dinv) @ (A @ T)).tocsr()
    P.sort_indices()
    R = P.T.tocsr()
    R.sort_indices()
    Ac = (R @ A @ P).tocsr()
    Ac.sort_indices()
    return P, R, Ac, rho


class SAAMG:
    """V(1,1)-cycle with weighted Jacobi (or Chebyshev) smoothing; dense inverse on the coarsest level."""

    def __init__(self, A, theta=0.08, max_levels=12, coarse_size=600, smoother="jacobi", cheb_deg=2):
        self.levels = []
        self.smoother, self.cheb_deg = smoother, cheb_deg
        A = A.tocsr()
        while A.shape[0] > coarse_size and len(self.levels) < max_levels - 1:
            P, R, Ac, rho = sa_level(A, theta)
            if Ac.shape[0] >= 0.8 * A.shape[0]:
                break
            self.levels.append(dict(A=A, P=P, R=R, dinv=1.0 / A.diagonal(), rho=rho))
            A = Ac
        self.Ac = A
        self.Ac_inv = np.linalg.inv(A.toarray())

    def _smooth(self, lv, x, b):
        A, dinv, rho = lv["A"], lv["dinv"], lv["rho"]
        if self.smoother == "jacobi":
            w = (4.0 / 3.0) / rho
            return x + w * dinv * (b - A @ x)
        # Chebyshev on [rho/30, 1.1 rho] for D^-1 A
        lmax, lmin = 1.1 * rho, rho / 30.0
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rhok = 1.0 / sigma
        r = dinv * (b - A @ x)
        dvec = r / theta
        x = x + dvec
        for _ in range(self.cheb_deg - 1):
            rhok1 = 1.0 / (2.0 * sigma - rhok)
            r = dinv * (b - A @ x)
            dvec = rhok1 * rhok * dvec + (2.0 * rhok1 / delta) * r
            x = x + dvec
            rhok = rhok1
        return x

    def vcycle(self, b, lvl=0):
        if lvl == len(self.levels):
            return self.Ac_inv @ b
        lv = self.levels[lvl]
        x = self._smooth(lv, np.zeros_like(b), b)
        rc = lv["R"] @ (b - lv["A"] @ x)
        x = x + lv["P"] @ self.vcycle(rc, lvl + 1)
        return self._smooth(lv, x, b)

    def __call__(self, b):
        return self.vcycle(b)

    def complexity(self):
        nnz = [lv["A"].nnz for lv in self.levels] + [self.Ac.nnz]
        return sum(nnz) / nnz[0], [lv["A"].shape[0] for lv in self.levels] + [self.Ac.shape[0]]
