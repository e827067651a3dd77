"""Design experiment (CPU, oracle-side): the Schur preconditioner with rank-local hierarchies, simulated in one process.
V0 = all levels processor-local, V1 = rank-local aggregates with global Galerkin operators on every level, V2/V3 = mixed,
V5 = local hierarchies + global coarse correction on the composite-prolongator space.  Result (N = 128, 8 ranks): V0 343,
V1 87, V5 223-350 iterations against 30 for the undivided hierarchy -> the product builds one GLOBAL hierarchy per field
(field-parallel).  Usage: python multirank_pc_simulation.py 128 8 2 V0,V1
"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pc_experiment import *
import oracle.amg as oamg
part = importlib.import_module("knp-emi-cgx_b200.partition")
def build(o, owner, variant, coarse_size=600):
    pc = oamg.SchurPC(o)          # gives ic, ip, M, msig, z
    ns = o.ns
    Pt = o.assemble_P(membrane_sign=+1.0).tocsr()
    rk = np.concatenate([np.tile(owner[o.S[s]], 4) for s in range(2)])
    C = Pt.tocoo(); keep = rk[C.row] == rk[C.col]
    Ploc = sp.csr_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=Pt.shape)
    for name, idx in (("amg_c", pc.ic), ("amg_p", pc.ip)):
        Al, Ag = Ploc[idx][:, idx].tocsr(), Pt[idx][:, idx].tocsr()
        amg = oamg.SAAMG(Al, gamma=2, gamma_last=3, coarse_size=coarse_size)
        L = len(amg.levels)
        # composite global Galerkin operators with the rank-local prolongators
        Aglob = [Ag]
        for l in range(L): Aglob.append((amg.levels[l]["R"] @ Aglob[-1] @ amg.levels[l]["P"]).tocsr())
        def setop(l, A):
            amg.levels[l]["A"] = A; d = 1.0 / A.diagonal(); amg.levels[l]["dinv"] = d
            amg.levels[l]["rho"] = float(np.max(np.abs(d) * np.asarray(np.abs(A).sum(axis=1)).ravel()))
        if variant in ("V1",):
            for l in range(L): setop(l, Aglob[l])
        if variant in ("V2",): setop(0, Aglob[0])
        if variant in ("V1", "V2", "V3"):
            amg.Ac = Aglob[L]; amg.Ac_inv = np.linalg.inv(Aglob[L].toarray())
        print("   ", name, variant, "levels", [lv["A"].shape[0] for lv in amg.levels] + [amg.Ac.shape[0]])
        setattr(pc, name, amg)
    return pc
n = int(sys.argv[1]); R = int(sys.argv[2]); steps = int(sys.argv[3]); variants = sys.argv[4].split(",")
for variant in variants:
    o = make(n); owner = part.rcb_owner(o.mesh.x, R)
    x = o.pack(); its = []; pc = None
    for i in range(steps):
        o.t += o.p.dt; o.gate_update(); A, b = o.assemble(o.t); ns_ = o.nullspace()
        if i == 0: b = b - ns_ * (ns_ @ b)
        if pc is None: pc = build(o, owner, variant)
        x, k = o.solve_gmres(A, b, x, ns_, pc, 1e-9, maxit=600); o.unpack(x); its.append(k)
    print(R, "ranks", variant, "its", its, flush=True)

print("---- V5: local hierarchies + additive global coarse correction on the composite-prolongator space")
class AddCoarse:
    def __init__(self, amg, Ag, mult=False):
        self.amg = amg; Z = amg.levels[0]["P"]
        for l in range(1, len(amg.levels)): Z = (Z @ amg.levels[l]["P"]).tocsr()
        self.Z = Z; E = (Z.T @ Ag @ Z).toarray(); self.Einv = np.linalg.inv(E); self.Ag = Ag; self.mult = mult
        print("    coarse dim", E.shape[0], "nnz(Z)/row %.1f" % (Z.nnz / Z.shape[0]))
    def __call__(self, r):
        z = self.amg(r)
        if self.mult:   # multiplicative: coarse correction on the residual after the local cycle
            return z + self.Z @ (self.Einv @ (self.Z.T @ (r - self.Ag @ z)))
        return z + self.Z @ (self.Einv @ (self.Z.T @ r))
for mult in (False, True):
    o = make(n); owner = part.rcb_owner(o.mesh.x, R)
    x = o.pack(); its = []; pc = None
    for i in range(steps):
        o.t += o.p.dt; o.gate_update(); A, b = o.assemble(o.t); ns_ = o.nullspace()
        if i == 0: b = b - ns_ * (ns_ @ b)
        if pc is None:
            pc = build(o, owner, "V0")
            Pt = o.assemble_P(membrane_sign=+1.0).tocsr()
            pc.amg_c = AddCoarse(pc.amg_c, Pt[pc.ic][:, pc.ic].tocsr(), mult); pc.amg_p = AddCoarse(pc.amg_p, Pt[pc.ip][:, pc.ip].tocsr(), mult)
        x, k = o.solve_gmres(A, b, x, ns_, pc, 1e-9, maxit=600); o.unpack(x); its.append(k)
    print(R, "ranks V5", "multiplicative" if mult else "additive", "its", its, flush=True)
