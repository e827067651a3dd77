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


def drop_dirichlet_aggregates(A, agg, nagg):
    """Rows without a non-zero off-diagonal entry (what essential boundary conditions leave behind) leave the coarse space:
    agg = -1; the remaining aggregates keep their order.  (csrc/amg_setup.cpp::drop_dirichlet_aggregates)"""
    A = A.tocsr()
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    off = np.zeros(n, bool)
    off[rows[(A.indices != rows) & (A.data != 0.0)]] = True
    if off.all():
        return agg, nagg
    used = np.zeros(nagg, bool)
    used[agg[off]] = True
    remap = np.cumsum(used) - 1
    return np.where(off, remap[agg], -1), int(used.sum())


def sa_level(A, theta=0.08, omega=4.0 / 3.0, filtered=None, level=0):
    """One smoothed-aggregation coarsening: returns (P, R, A_c, rho) with constant near-nullspace.
    The prolongator is smoothed with the filtered matrix (strong off-diagonals, weak ones lumped into the diagonal)."""
    A = A.tocsr()
    n = A.shape[0]
    for attempt in range(4):                 # threshold halved until the strength graph has >= 3 edges per row
        S = strength_graph(A, theta * 0.5 ** attempt)
        if S.nnz >= 3.0 * n:
            break
    agg, nagg = mis2_aggregate(S)
    if level == 0:                           # boundary rows exist on the finest level only
        agg, nagg = drop_dirichlet_aggregates(A, agg, nagg)
    # unnormalised tentative prolongator: the constant stays the near-nullspace vector on every level
    inn = np.flatnonzero(agg >= 0)
    T = sp.csr_matrix((np.ones(inn.size), (inn, agg[inn])), shape=(n, nagg))
    dinv = 1.0 / A.diagonal()
    rho = float(np.max(np.abs(dinv) * np.asarray(np.abs(A).sum(axis=1)).ravel()))   # Gershgorin bound on rho(D^-1 A)
    keep = (S + sp.identity(n, dtype=np.int8, format="csr")).astype(float)
    AF = A.multiply(keep).tocsr()
    AF = (AF + sp.diags(np.asarray(A.sum(axis=1)).ravel() - np.asarray(AF.sum(axis=1)).ravel())).tocsr()
    if filtered is None:
        # the finest level (mesh edges with vanishing stiffness) and the dense Galerkin levels of 3D meshes
        filtered = level == 0 or A.nnz > 32.0 * n
    if not filtered:
        AF = A
    rhoF = float(np.max(np.abs(dinv) * np.asarray(np.abs(AF).sum(axis=1)).ravel()))
    P = (T - sp.diags((omega / rhoF) * dinv) @ (AF @ T)).tocsr()
    P.sort_indices()
    R = P.T.tocsr()
    R.sort_indices()
    Ac = (R @ A @ P).tocsr()
    Ac.sort_indices()
    return P, R, Ac, rho


class SAAMG:
    """V(1,1)-cycle with weighted Jacobi (or Chebyshev) smoothing; dense inverse on the coarsest level."""

    def __init__(self, A, theta=0.08, max_levels=12, coarse_size=600, smoother="jacobi", cheb_deg=2, filtered=None,
                 theta_decay=1.0, gamma=1, gamma_last=1 << 20, storage="float64"):
        """storage="float32": the cycle applies the level operators, transfers and the coarsest inverse rounded to single
        precision (products still accumulated in double) -- how the product stores its hierarchies on the device
        (csrc/solver.cu::to_f32); the setup and lv["A"] / Ac stay exact for the level-by-level comparisons."""
        self.levels = []
        self.storage = storage
        self.smoother, self.cheb_deg = smoother, cheb_deg
        self.gamma, self.gamma_last = gamma, gamma_last   # cycle index on levels 1..gamma_last (2: W-cycle)
        A = A.tocsr()
        while A.shape[0] > coarse_size and len(self.levels) < max_levels - 1:
            P, R, Ac, rho = sa_level(A, theta * theta_decay ** len(self.levels), filtered=filtered, level=len(self.levels))
            if Ac.shape[0] >= 0.8 * A.shape[0]:
                break
            self.levels.append(dict(A=A, P=P, R=R, dinv=1.0 / A.diagonal(), rho=rho))
            A = Ac
        self.Ac = A
        self.Ac_inv = np.linalg.inv(A.toarray())
        rnd = (lambda M: M) if storage == "float64" else self._rounded
        for lv in self.levels:
            lv["Aop"], lv["Pop"], lv["Rop"] = rnd(lv["A"]), rnd(lv["P"]), rnd(lv["R"])
        self.Ac_inv_op = self.Ac_inv if storage == "float64" else self.Ac_inv.astype(np.float32).astype(np.float64)

    @staticmethod
    def _rounded(M):
        M = M.copy()
        M.data = M.data.astype(np.float32).astype(np.float64)
        return M

    def _smooth(self, lv, x, b):
        A, dinv, rho = lv["Aop"], lv["dinv"], lv["rho"]
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
            return self.Ac_inv_op @ b
        lv = self.levels[lvl]
        x = self._smooth(lv, np.zeros_like(b), b)
        for _ in range(self.gamma if 1 <= lvl <= self.gamma_last else 1):
            rc = lv["Rop"] @ (b - lv["Aop"] @ x)
            x = x + lv["Pop"] @ self.vcycle(rc, lvl + 1)
        return self._smooth(lv, x, b)

    def __call__(self, b):
        return self.vcycle(b)

    def complexity(self):
        nnz = [lv["A"].nnz for lv in self.levels] + [self.Ac.nnz]
        return sum(nnz) / nnz[0], [lv["A"].shape[0] for lv in self.levels] + [self.Ac.shape[0]]


class SchurPC:
    """CPU restatement of the product's charge-conservation Schur preconditioner (TEST INFRASTRUCTURE; the
    reference has no counterpart -- it uses hypre on the block-diagonal P, KNPEMIx_solver.py:269-273).

    The potential row of `a` (KNPEMIx_problem.py:603-610) is the z_k-weighted sum of the ion rows (:598-600) minus
    sum_k z_k M c_k, so L A with L = [I 0; -Z I] has the block form [A_cc A_cphi; -Z M 0].  Its Schur complement
    Z M A_cc^-1 A_cphi behaves like sum_k (z_k^2 cbar_k/psi) M at high and like the phi block of P at low
    frequencies, hence  S~^-1 = (K_phi + (C_M/F) M_Gamma)^-1 + M_sigma^-1  (M_sigma lumped).  Application:
        v = L r ;  z_c = AMG_c(v_c) ;  t = v_phi + M (sum_k z_k z_ck) ;  z_phi = AMG_phi(t) + t / M_sigma.
    Everything is frozen at the state the oracle had when the object was built (the reference assembles P once)."""

    def __init__(self, o, theta=0.08, smoother="jacobi", exact=False, **amg_kw):
        p = o.p
        self.o = o
        ns = o.ns
        self.ic = np.concatenate([np.arange(o.base[s], o.base[s] + 3 * ns[s]) for s in range(2)])
        self.ip = np.concatenate([np.arange(o.base[s] + 3 * ns[s], o.base[s] + 4 * ns[s]) for s in range(2)])
        Pt = o.assemble_P(membrane_sign=+1.0).tocsr()
        Mm = o.assemble_P(membrane_sign=0.0, D_scale=0.0, bc_diag=0.0).tocsr()    # Dirichlet dofs: empty rows and columns
        self.M = [Mm[o.base[s]: o.base[s] + ns[s]][:, o.base[s]: o.base[s] + ns[s]].tocsr() for s in range(2)]
        self.msig = []
        for s in range(2):
            sigma = sum(p.z[k] ** 2 / p.psi * o.c[s][k][o.S[s]] for k in range(3))
            ms = o.lumped_mass(self.M[s])
            self.msig.append(np.where(ms != 0.0, sigma * ms, np.inf))
        Acc = Pt[self.ic][:, self.ic].tocsr()
        App = Pt[self.ip][:, self.ip].tocsr()
        if exact:
            import scipy.sparse.linalg as spla
            self.amg_c, self.amg_p = spla.splu(Acc.tocsc()).solve, spla.splu(App.tocsc()).solve
        else:
            amg_kw.setdefault("gamma", 2)
            amg_kw.setdefault("coarse_size", 2500)      # dense coarsest solve (inverted on the device in the product)
            self.amg_c = SAAMG(Acc, theta=theta, smoother=smoother, **amg_kw)
            self.amg_p = SAAMG(App, theta=theta, smoother=smoother, **amg_kw)
            if "gamma_last" not in amg_kw:              # W-cycle on all levels but the finest and the coarsest sparse one
                for a in (self.amg_c, self.amg_p):
                    a.gamma_last = max(1, len(a.levels) - 2)
        self.z = np.asarray(p.z, float)

    def __call__(self, r):
        o, ns = self.o, self.o.ns
        vc = r[self.ic]
        zc = self.amg_c(vc)
        t = []
        off = 0
        for s in range(2):
            rc = vc[off: off + 3 * ns[s]].reshape(3, ns[s])
            zz = zc[off: off + 3 * ns[s]].reshape(3, ns[s])
            rphi = r[o.base[s] + 3 * ns[s]: o.base[s] + 4 * ns[s]]
            t.append(rphi - self.z @ rc + self.M[s] @ (self.z @ zz))
            off += 3 * ns[s]
        t = np.concatenate(t)
        zp = self.amg_p(t) + t / np.concatenate(self.msig)
        out = np.empty_like(r)
        out[self.ic] = zc
        out[self.ip] = zp
        return out
