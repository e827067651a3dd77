"""CPU-only check of the edge-lane row kernel's inputs and arithmetic (csrc/topology.cpp tables + the closed-form P1 cell
entries of csrc/assembly.cu::rows_edge_kernel), restated in numpy and compared with the oracle's assembled matrix:
ion-ion blocks M + dt D_k K and, on rows away from the membrane, the ion-potential blocks (dt D_k z_k / psi) K[cbar_k].
Tolerance 1e-12 relative to the row's largest entry (north star)."""
import numpy as np
import pytest
from oracle.fixtures import unit_square, unit_cube, from_arrays
from oracle.knpemi import KNPEMIOracle, OracleParams
from conftest import MODELS_TEST


def emulate_edge_lanes(t, gdim, conc):
    """a_m, a_kk, X[3] per (dof, slot) exactly as the kernel forms them (same hit order, same closed forms, self slot
    from the row-sum identities).  conc[s][k] = concentrations per local dof of subdomain s."""
    adj, hit, meta, X = t["adjG"], t["hitG"], t["meta"], t["node_x"]
    W, G = adj.shape
    n_own = t["n_own_loc"][:2]
    n_loc = t["n_own_loc"][2:]
    out = np.zeros((W, G, 5))
    self_slot = (meta[:, 0] >> 8) & 255
    for s in range(2):
        w0, w1 = (0, n_own[0]) if s == 0 else (n_own[0], n_own[0] + n_own[1])
        A = adj[w0:w1]
        n = A.shape[0]
        Aq = np.where(A >= 0, A, 0)
        xs = X[(n_loc[0] if s else 0) + Aq]                      # (n, G, d)
        cs = np.stack([conc[s][k][Aq] for k in range(3)], -1)   # (n, G, 3)
        rows = np.arange(n)
        sf = self_slot[w0:w1]
        xp, cp = xs[rows, sf], cs[rows, sf]
        lane_ok = (A >= 0) & (np.arange(G)[None, :] != sf[:, None])
        acc = np.zeros((n, G, 5))
        nh = 2 if gdim == 2 else 8
        for h in range(nh):
            if gdim == 2:
                sl = (hit[w0:w1, :, 0] >> (8 * h)) & 255
                ok = lane_ok & (sl != 255)
                sl = np.where(ok, sl, 0)
                xr, cr_ = xs[rows[:, None], sl], cs[rows[:, None], sl]
                e1, e2 = xp[:, None, :] - xr, xs - xr
                cross = np.abs(e1[..., 0] * e2[..., 1] - e1[..., 1] * e2[..., 0])
                dot = (e1 * e2).sum(-1)
                with np.errstate(all="ignore"):
                    kab = -0.5 * dot * (1.0 / cross)
                mv = cross / 24.0
                cbar = (cp[:, None, :] + cs + cr_) * (1.0 / 3.0)
            else:
                code = (hit[w0:w1, :, h >> 1] >> (16 * (h & 1))) & 0xFFFF
                ok = lane_ok & (code != 0xFFFF)
                r_, s_ = np.where(ok, code & 255, 0), np.where(ok, code >> 8, 0)
                xr, xt = xs[rows[:, None], r_], xs[rows[:, None], s_]
                a, b, c = xs - xp[:, None, :], xr - xp[:, None, :], xt - xp[:, None, :]
                nq = np.cross(b, c)
                J = np.abs((a * nq).sum(-1))
                npv = np.cross(b - a, c - a)
                with np.errstate(all="ignore"):
                    kab = -(npv * nq).sum(-1) * (1.0 / (6.0 * J))
                mv = J / 120.0
                cbar = (cp[:, None, :] + cs + cs[rows[:, None], r_] + cs[rows[:, None], s_]) * 0.25
            acc[..., 0] += np.where(ok, mv, 0.0)
            acc[..., 1] += np.where(ok, kab, 0.0)
            acc[..., 2:] += np.where(ok[..., None], cbar * kab[..., None], 0.0)
        tot = acc.sum(axis=1)
        acc[rows, sf, 0] = tot[:, 0] * (2.0 / gdim)
        acc[rows, sf, 1:] = -tot[:, 1:]
        out[w0:w1] = acc
    return out


def _cells(kb, d, n, m, fill=0.5, shape=None):
    mm = kb.mesh.cell_array_mesh(d, n, m, fill=fill, shape=shape)
    om = from_arrays(d, mm.x, mm.cells, mm.cell_tags, mm.intra_tags)
    it = tuple(mm.intra_tags)
    return om, OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,))


CASES = {"square8": lambda kb: (unit_square(8), OracleParams()), "cube4": lambda kb: (unit_cube(4), OracleParams()),
         "cells2d": lambda kb: _cells(kb, 2, 24, 3), "cells3d": lambda kb: _cells(kb, 3, 8, 2),
         "plates3d": lambda kb: _cells(kb, 3, 16, 2, 0.75, {"plates": True, "thickness": 1, "pitch": 2, "spine": 1})}


@pytest.mark.parametrize("name", list(CASES))
def test_edge_lane_tables_reproduce_the_oracle_blocks(kb, name):
    om, p = CASES[name](kb)
    d = om.gdim
    o = KNPEMIOracle(om, p, MODELS_TEST)
    rng = np.random.default_rng(1)
    for s in range(2):
        o.c[s] *= 1 + 0.05 * rng.random(o.c[s].shape)
    A, _ = o.assemble(p.dt)
    qb, qw = kb.mesh.facet_quadrature(d)
    t = kb.lib.edge_tables_host(d, om.x, om.cells, om.cell_tags, p.intra_tags, p.extra_tag, om.mf_verts, om.mf_tags, qb, qw)
    assert t is not None
    assert t["adjG"].shape[0] == o.ns[0] + o.ns[1]
    conc = [[o.c[s][k][o.S[s]] for k in range(3)] for s in range(2)]
    acc = emulate_edge_lanes(t, d, conc)
    rowmax = np.abs(A).max(axis=1).toarray().ravel()
    on_membrane = np.zeros(om.x.shape[0], bool)
    on_membrane[o.mverts] = True
    worst = 0.0
    for s in range(2):
        w0 = 0 if s == 0 else o.ns[0]
        adj = t["adjG"][w0:w0 + o.ns[s]]
        pp, ee = np.nonzero(adj >= 0)
        vq, vp = o.S[s][adj[pp, ee]], o.S[s][pp]
        a = acc[w0 + pp, ee]
        for k in range(3):
            rk, ck = o.row(s, k, vp), o.row(s, k, vq)
            ref = np.asarray(A[rk, ck]).ravel()
            got = a[:, 0] + p.dt * p.D[k] * a[:, 1]
            worst = max(worst, (np.abs(got - ref) / rowmax[rk]).max())
            off = ~on_membrane[vp]                                 # membrane rows carry dS terms in this block
            if not off.any():
                continue
            ref2 = np.asarray(A[rk[off], o.row(s, 3, vq[off])]).ravel()
            got2 = (p.dt * p.D[k] * p.z[k] / p.psi) * a[off, 2 + k]
            worst = max(worst, (np.abs(got2 - ref2) / rowmax[rk[off]]).max())
    assert worst < 1e-12, worst
