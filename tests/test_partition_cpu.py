"""Multi-GPU host logic on the CPU: vertex partition, ghost layer, halo lists (simulated ranks in one process),
owner-computes assembly against the global oracle, and a real world_size-2 gloo run of the Comm shim + exchange."""
import os
import subprocess
import sys
import numpy as np
import pytest
import scipy.sparse as sp
from oracle.fixtures import from_arrays
from oracle.knpemi import KNPEMIOracle, OracleParams
from conftest import MODELS_TEST, ROOT


def _setup(kb, gdim, n, m, size):
    part = __import__("importlib").import_module("knp-emi-cgx_b200.partition")
    mesh = kb.mesh.cell_array_mesh(gdim, n, m)
    locs = [part.partition_mesh(mesh, r, size) for r in range(size)]
    lays = [part.Layout(part.local_dofmaps(l), l.n_owned) for l, _ in locs]
    reqs = []
    for (l, info), lay in zip(locs, lays):
        mine = part.ghost_requests(l, info, lay)
        reqs.append({r: {s: v[0] for s, v in d.items()} for r, d in mine.items()})
    lists = [part.build_halo_lists(l, info, lay, reqs) for (l, info), lay in zip(locs, lays)]
    return part, mesh, locs, lays, lists


@pytest.mark.parametrize("gdim,n,m,size", [(2, 24, 3, 2), (2, 24, 3, 3), (2, 32, 4, 4), (3, 8, 2, 2), (3, 8, 2, 8)])
def test_partition_invariants_and_halo_exchange(kb, gdim, n, m, size):
    part, mesh, locs, lays, lists = _setup(kb, gdim, n, m, size)
    nv = mesh.x.shape[0]
    owned = np.concatenate([info["l2g"][:l.n_owned] for l, info in locs])
    assert np.array_equal(np.sort(owned), np.arange(nv))                         # every vertex owned exactly once
    # every cell / membrane facet integrated exactly once
    cnt = np.zeros(mesh.cells.shape[0], int)
    key = {tuple(c): i for i, c in enumerate(mesh.cells.tolist())}
    for l, info in locs:
        gc = info["l2g"][l.cells]
        for c, o in zip(gc.tolist(), l.cell_owned.tolist()):
            cnt[key[tuple(c)]] += o
    assert (cnt == 1).all()
    assert sum(int(l.mf_owned.sum()) for l, _ in locs) == mesh.mf_verts.shape[0]
    # rows: union over ranks = global dof count
    o = KNPEMIOracle(from_arrays(gdim, mesh.x, mesh.cells, mesh.cell_tags, mesh.intra_tags),
                     OracleParams(intra_tags=tuple(mesh.intra_tags), extra_tag=1, membrane_tags=tuple(mesh.intra_tags),
                                  stimulus_tags=(2,)), MODELS_TEST)
    assert sum(lay.n_rows for lay in lays) == o.n
    # simulated halo exchange of a field that encodes (subdomain, field, global vertex)
    def field(s, f, gv):
        return 1000.0 * (4 * s + f) + gv + 0.5
    xs = []
    for (l, info), lay in zip(locs, lays):
        x = np.full(lay.n_cols, np.nan)
        for s in range(2):
            gv = info["l2g"][lay.node_vert[s]]
            for f in range(4):
                q = np.arange(lay.n_own[s])
                x[lay.col(s, f, q)] = field(s, f, gv[:lay.n_own[s]])
        xs.append(x)
    for r, (peers, sp_, sc, rp, rc) in enumerate(lists):
        for i, pr in enumerate(peers.tolist()):
            ppeers, psp, psc, _, _ = lists[pr]
            j = ppeers.tolist().index(r)
            buf = xs[pr][psc[psp[j]:psp[j + 1]]]                                  # what the peer packs for me
            assert buf.size == rp[i + 1] - rp[i]
            xs[r][rc[rp[i]:rp[i + 1]]] = buf
    for (l, info), lay, x in zip(locs, lays, xs):
        assert not np.isnan(x).any()                                             # every ghost column was filled
        for s in range(2):
            gv = info["l2g"][lay.node_vert[s]]
            for f in range(4):
                assert np.array_equal(x[lay.col(s, f, np.arange(lay.n_loc[s]))], field(s, f, gv))


def test_owner_computes_assembly_equals_global_rows(kb):
    """Assembling on the local mesh (owned vertices + one ghost-cell layer) reproduces the global rows of the owned
    dofs without any exchange: the claim behind the distributed row kernel."""
    part, mesh, locs, lays, _ = _setup(kb, 2, 24, 3, 3)
    it = tuple(mesh.intra_tags)
    P = OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,), scale_stimulus=False)
    og = KNPEMIOracle(from_arrays(2, mesh.x, mesh.cells, mesh.cell_tags, it), P, MODELS_TEST)
    rng = np.random.default_rng(0)
    for s in range(2):
        og.c[s] *= 1 + 0.05 * rng.random(og.c[s].shape)
    og.phi_m += 0.003 * rng.standard_normal(og.phi_m.shape)
    og.gates *= 1 + 0.1 * rng.random(og.gates.shape)
    Ag, bg = og.assemble(2 * P.dt)
    Ag = Ag.tocsr()
    for (l, info), lay in zip(locs, lays):
        l2g = info["l2g"]
        ol = KNPEMIOracle(from_arrays(2, l.x, l.cells, l.cell_tags, it), P, MODELS_TEST)
        for s in range(2):
            ol.c[s] = og.c[s][:, l2g].copy()
        ol.phi_m, ol.gates = og.phi_m[l2g].copy(), og.gates[:, l2g].copy()
        Al, bl = ol.assemble(2 * P.dt)
        # local oracle numbering (all local dofs, owned and ghost) -> global rows
        gmap = np.empty(ol.n, np.int64)
        for s in range(2):
            for f in range(4):
                lo = ol.base[s] + f * ol.ns[s]
                gmap[lo:lo + ol.ns[s]] = og.row(s, f, l2g[ol.S[s]])
        owned_rows = np.concatenate([ol.base[s] + f * ol.ns[s] + np.flatnonzero(ol.S[s] < l.n_owned)
                                     for s in range(2) for f in range(4)])
        C = Al[owned_rows].tocoo()
        Aloc = sp.csr_matrix((C.data, (gmap[owned_rows][C.row], gmap[C.col])), shape=Ag.shape)
        sel = sp.csr_matrix((np.ones(owned_rows.size), (gmap[owned_rows], gmap[owned_rows])), shape=Ag.shape)
        ref = sel @ Ag
        diff = abs(Aloc - ref)
        assert diff.max() <= 1e-13 * abs(Ag).max()
        assert np.abs(bl[owned_rows] - bg[gmap[owned_rows]]).max() <= 1e-13 * np.abs(bg).max()


def test_gloo_world_size_2():
    """Real two-process run (gloo): Comm shim collectives and a halo exchange through torch.distributed."""
    worker = os.path.join(os.path.dirname(__file__), "dist_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", worker],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("WORKER_OK") == 2


@pytest.mark.parametrize("gdim,n,m,size", [(2, 24, 3, 3), (3, 8, 2, 4)])
def test_native_local_patterns_tile_the_global_pattern(kb, gdim, n, m, size):
    """topology.cpp on every rank's local mesh (owned vertices + ghost-cell layer): the owned rows, mapped through the
    [owned | ghost tail] column layout back to global numbering, are exactly the rows of the global pattern -- the
    structure side of owner-computes assembly, checked on the host without a GPU."""
    part, mesh, locs, lays, _ = _setup(kb, gdim, n, m, size)
    it = tuple(mesh.intra_tags)
    og = KNPEMIOracle(from_arrays(gdim, mesh.x, mesh.cells, mesh.cell_tags, it),
                      OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,)), MODELS_TEST)
    Ag, _ = og.assemble(og.p.dt)
    Ag = Ag.tocsr()
    qb, qw = kb.mesh.facet_quadrature(gdim)
    seen = np.zeros(og.n, bool)
    for (l, info), lay in zip(locs, lays):
        ip, ix, vi, ve = kb.lib.pattern_host(gdim, l.x, l.cells, l.cell_tags, it, 1, l.mf_verts, l.mf_tags, qb, qw,
                                             n_owned_vertices=l.n_owned, cell_owned=l.cell_owned, mfacet_owned=l.mf_owned)
        assert np.array_equal(vi, lay.node_vert[0]) and np.array_equal(ve, lay.node_vert[1])
        assert ip.size - 1 == lay.n_rows
        l2g = info["l2g"]
        # local column -> global column
        gcol = np.empty(lay.n_cols, np.int64)
        for s in range(2):
            gv = l2g[lay.node_vert[s]]
            for f in range(4):
                gcol[lay.col(s, f, np.arange(lay.n_loc[s]))] = og.row(s, f, gv)
        for s in range(2):
            gv = l2g[lay.node_vert[s][:lay.n_own[s]]]
            for f in range(4):
                lrows = lay.rowbase[s] + f * lay.n_own[s] + np.arange(lay.n_own[s])
                grows = og.row(s, f, gv)
                for lr, gr in zip(lrows.tolist(), grows.tolist()):
                    mine = np.sort(gcol[ix[ip[lr]:ip[lr + 1]]])
                    ref = Ag.indices[Ag.indptr[gr]:Ag.indptr[gr + 1]]
                    assert np.array_equal(mine, ref), (s, f, lr, gr)
                    seen[gr] = True
    assert seen.all()


PLATES = {"plates": True, "thickness": 1, "pitch": 2, "spine": 1}


@pytest.mark.parametrize("gdim,n,m,size,shape", [(2, 24, 3, 2, None), (2, 24, 3, 4, None), (2, 32, 2, 8, None), (3, 8, 2, 2, None),
                                                 (3, 12, 2, 8, None), (3, 8, 2, 6, None), (3, 16, 2, 8, PLATES), (2, 32, 2, 4, PLATES)])
def test_local_slab_generator_equals_partition_of_the_global_mesh(kb, gdim, n, m, size, shape):
    """Multi-GPU runs of the structured tissue blocks generate only their own slab (mesh.cell_array_mesh_local); it must be
    exactly the local mesh partition_mesh cuts out of the global mesh under the same (block) vertex -> rank map."""
    part = __import__("importlib").import_module("knp-emi-cgx_b200.partition")
    g = kb.mesh.cell_array_mesh(gdim, n, m, fill=0.75 if shape else 0.5, shape=shape)
    own = kb.mesh.BlockOwner(gdim, n, size)
    owner = own(np.arange(g.x.shape[0]))
    assert np.bincount(owner, minlength=size).min() > 0
    for rank in range(size):
        a, ia = part.partition_mesh(g, rank, size, owner=owner)
        b, ib = kb.mesh.cell_array_mesh_local(gdim, n, m, rank, size, fill=0.75 if shape else 0.5, shape=shape)
        assert a.n_owned == b.n_owned and np.array_equal(ia["l2g"], ib["l2g"])
        assert np.allclose(a.x, b.x, rtol=0, atol=1e-22)
        assert np.array_equal(a.cells, b.cells) and np.array_equal(a.cell_tags, b.cell_tags)
        assert np.array_equal(a.cell_owned, b.cell_owned)
        assert np.array_equal(a.mf_verts, b.mf_verts) and np.array_equal(a.mf_tags, b.mf_tags)
        assert np.array_equal(a.mf_owned, b.mf_owned)
        gv = ia["l2g"]
        assert np.array_equal(ia["owner_of"](gv), ib["owner_of"](gv))


def test_tissue_like_plates_meet_the_c5_statistics(kb):
    """BASELINE C5: the plate-stack cells must give membrane facets / cells >= 0.2 and membrane
    vertices / vertices >= 0.5 (real tissue: 0.30 / 0.88); checked on a reduced block array with the C5 block size 32."""
    m = kb.mesh.cell_array_mesh(3, 64, 2, fill=0.875, shape=PLATES)
    nv, nc, nf = m.x.shape[0], m.cells.shape[0], m.mf_verts.shape[0]
    assert nf / nc >= 0.2, nf / nc
    assert np.unique(m.mf_verts).size / nv >= 0.5, np.unique(m.mf_verts).size / nv
