"""Design experiment (CPU, oracle-side; lives under tests/ because only test infrastructure may import oracle/): GMRES iteration counts of candidate preconditioners on the perturbed C3
workload.  Not product code; its output motivates the Schur-complement preconditioner documented in DESIGN.md."""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cgx_b200 as kb
from oracle.fixtures import from_arrays
from oracle.knpemi import KNPEMIOracle, OracleParams
from oracle.amg import SAAMG

def make(n, cpd=8, dim=2, models=None):
    m = kb.mesh.cell_array_mesh(dim, n, cpd)
    om = from_arrays(dim, m.x, m.cells, m.cell_tags, m.intra_tags)
    it = tuple(m.intra_tags)
    p = OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,))
    o = KNPEMIOracle(om, p, models or [("NeuronalCT", None), ("HH", None), ("ATP", None)])
    X = om.x / 1e-6
    fac = 1 + 0.01 * np.sin(2 * np.pi * X[:, 0]) * np.sin(2 * np.pi * X[:, 1])
    for s in range(2):
        o.c[s] *= fac[None, :]
    dphi = 0.005 * np.cos(2 * np.pi * X[:, 0])
    o.phi_m += dphi
    o.phi[0] += dphi
    return o

def index_sets(o):
    ic, ip = [], []
    for s in range(2):
        ic.append(np.arange(o.base[s], o.base[s] + 3 * o.ns[s]))
        ip.append(np.arange(o.base[s] + 3 * o.ns[s], o.base[s] + 4 * o.ns[s]))
    return np.concatenate(ic), np.concatenate(ip)

def lumped_sigma_mass(o):
    """diag of sum_k z_k^2/psi * M[cbar_k], lumped, on the phi rows (intra then extra)."""
    p = o.p
    out = []
    for s in range(2):
        cells = o.cells_s[s]
        vol = o.geo[s]["vol"]
        d = o.mesh.gdim
        sig = sum(p.z[k] ** 2 / p.psi * o.c[s][k][cells].mean(axis=1) for k in range(3))
        diag = np.zeros(o.ns[s])
        np.add.at(diag, o.r[s][cells].ravel(), np.repeat(vol * sig / (d + 1), d + 1))
        out.append(diag)
    return np.concatenate(out)

def consistent_sigma_mass(o):
    p = o.p
    blocks = []
    for s in range(2):
        cells = o.cells_s[s]
        M = o.geo[s]["M"]
        sig = sum(p.z[k] ** 2 / p.psi * o.c[s][k][cells].mean(axis=1) for k in range(3))
        R = o.r[s][cells]
        rows = np.broadcast_to(R[:, :, None], M.shape).ravel(); cols = np.broadcast_to(R[:, None, :], M.shape).ravel()
        blocks.append(sp.coo_matrix(((sig[:, None, None] * M).ravel(), (rows, cols)), shape=(o.ns[s], o.ns[s])).tocsr())
    return sp.block_diag(blocks).tocsr()

def run(o, make_pinv, steps, rtol=1e-9, label=""):
    x = o.pack(); its = []; t0 = time.time()
    Pinv = None
    for i in range(steps):
        o.t += o.p.dt; o.gate_update()
        A, b = o.assemble(o.t); ns = o.nullspace()
        if i == 0:
            b = b - ns * (ns @ b)
        if Pinv is None:
            Pinv = make_pinv(o, A)
        x, k = o.solve_gmres(A, b, x, ns, Pinv, rtol); o.unpack(x); its.append(k)
    print(f"{label:28s} its {its}  ({time.time() - t0:.1f}s)", flush=True)
    return x

def bj_exact(o, A):
    lu = spla.splu(o.assemble_P().tocsc()); return lambda v: lu.solve(v)
def bj_amg(o, A):
    return SAAMG(o.assemble_P())

def tri(kind, solver, mass):
    def f(o, A):
        ic, ip = index_sets(o)
        A = A.tocsr()
        Acc = A[ic][:, ic].tocsc(); App = A[ip][:, ip].tocsc(); Apc = A[ip][:, ic].tocsr(); Acp = A[ic][:, ip].tocsr()
        if solver == "exact":
            # App is singular (constants): regularise by tiny shift for LU; nullspace is projected by GMRES anyway
            luc = spla.splu(Acc); lup = spla.splu((App + 1e-14 * sp.identity(App.shape[0]) * abs(App.diagonal()).max()).tocsc())
            sc, spp = luc.solve, lup.solve
        else:
            sc, spp = SAAMG(Acc.tocsr()), SAAMG(App.tocsr())
        if mass == "lumped":
            ml = lumped_sigma_mass(o); minv = lambda r: r / ml
        elif mass == "consistent":
            lum = spla.splu(consistent_sigma_mass(o).tocsc()); minv = lum.solve
        else:
            minv = lambda r: 0.0 * r
        def sinv(r):
            r = r - r.mean()
            return spp(r) + minv(r)
        def apply(v):
            z = np.zeros_like(v)
            if kind == "lower":
                zc = sc(v[ic]); zp = sinv(v[ip] - Apc @ zc)
            elif kind == "upper":
                zp = sinv(v[ip]); zc = sc(v[ic] - Acp @ zp)
            else:   # diag
                zc = sc(v[ic]); zp = sinv(v[ip])
            z[ic] = zc; z[ip] = zp
            return z
        return apply
    return f

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    which = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    cases = {
        "bj-exact": bj_exact, "bj-amg": bj_amg,
        "lower-exact-lumped": tri("lower", "exact", "lumped"), "upper-exact-lumped": tri("upper", "exact", "lumped"),
        "lower-exact-consistent": tri("lower", "exact", "consistent"), "lower-exact-nomass": tri("lower", "exact", "none"),
        "diag-exact-lumped": tri("diag", "exact", "lumped"),
        "lower-amg-lumped": tri("lower", "amg", "lumped"), "upper-amg-lumped": tri("upper", "amg", "lumped"),
    }
    ref = None
    for nm, f in cases.items():
        if which and nm not in which: continue
        x = run(make(n), f, steps, label=nm)
        if ref is None: ref = x
        else: print("    rel diff of final x vs first case:", np.linalg.norm(x - ref) / np.linalg.norm(ref))
