"""Mesh partitioning and ghost-column exchange lists for one-process-per-GPU runs.

Replaces the graph partitioner + GhostMode.shared_facet ghost layer that DOLFINx applies when the reference
reads its mesh in parallel (src/CGx/utils/mixed_dim_problem.py:21,649,666) and the PETSc VecScatter that moves
ghost values (src/CGx/KNPEMI/KNPEMIx_solver.py:439,458-468).

Design: vertices are partitioned (recursive coordinate bisection -- no METIS in this
image; for the structured fixtures it yields the block partition); a rank owns the rows of all 8 fields of its
vertices (so both restrictions of a membrane vertex live on one GPU) and keeps every cell that touches an owned
vertex, so that assembly is owner-computes with NO exchange beyond the halo of the previous solution.
"""
import numpy as np

from .mesh import Mesh


def rcb_owner(x, nparts):
    """Recursive coordinate bisection of the vertex cloud into `nparts` parts (any nparts >= 1);
    deterministic (stable sorts, ties broken by vertex id)."""
    owner = np.zeros(x.shape[0], np.int32)

    def split(ids, lo, hi):
        n = hi - lo
        if n == 1:
            owner[ids] = lo
            return
        left = n // 2
        pts = x[ids]
        ax = int(np.argmax(pts.max(0) - pts.min(0)))
        order = np.lexsort((ids, pts[:, ax]))
        cut = int(round(ids.size * left / n))
        split(ids[order[:cut]], lo, lo + left)
        split(ids[order[cut:]], lo + left, hi)

    split(np.arange(x.shape[0]), 0, nparts)
    return owner


def partition_mesh(m: Mesh, rank: int, size: int, owner=None):
    """Local mesh of `rank`: owned vertices first (ascending global id), then ghosts; every cell / membrane
    facet touching an owned vertex; ownership flags for functionals.  Returns (local Mesh, info dict).
    owner: vertex -> rank array (default: recursive coordinate bisection)."""
    if owner is None:
        owner = rcb_owner(m.x, size)
    nv = m.x.shape[0]
    mine = owner == rank
    cell_has = mine[m.cells].any(axis=1)
    lcells_g = m.cells[cell_has]
    used = np.zeros(nv, bool)
    used[lcells_g.ravel()] = True
    owned_ids = np.flatnonzero(mine & used)
    # owned vertices that touch no cell cannot exist in a conforming mesh, but keep them out of the dof maps
    ghost_ids = np.flatnonzero(used & ~mine)
    l2g = np.concatenate([owned_ids, ghost_ids])
    g2l = np.full(nv, -1, np.int64)
    g2l[l2g] = np.arange(l2g.size)
    lcells = g2l[lcells_g].astype(np.int32)
    ltags = m.cell_tags[cell_has]
    # a cell / facet is integrated by the rank that owns its lowest-numbered vertex
    cell_owned = (owner[lcells_g.min(axis=1)] == rank).astype(np.uint8)
    f_has = mine[m.mf_verts].any(axis=1) if m.mf_verts.size else np.zeros(0, bool)
    lfv_g = m.mf_verts[f_has]
    lfv = g2l[lfv_g].astype(np.int32)
    mf_owned = (owner[lfv_g.min(axis=1)] == rank).astype(np.uint8) if lfv_g.size else np.zeros(0, np.uint8)
    local = Mesh(m.gdim, m.x[l2g], lcells, ltags.astype(np.int32), m.intra_tags, m.extra_tag, lfv,
                 m.mf_tags[f_has].astype(np.int32), grid=m.grid, n_owned=int(owned_ids.size), cell_owned=cell_owned,
                 mf_owned=mf_owned, vert_global=l2g)
    if m.bc_verts is not None:
        lb = g2l[np.asarray(m.bc_verts, np.int64)]
        local.bc_verts = lb[lb >= 0].astype(np.int32)
    info = dict(owner_of=lambda gid, _o=owner: _o[gid], l2g=l2g, rank=rank, size=size)
    return local, info


def local_dofmaps(local: Mesh):
    """Restricted dof -> local vertex per subdomain, exactly as libknpemi_b200 numbers them
    (ascending local vertex id, which puts owned dofs first)."""
    is_in = np.isin(local.cell_tags, np.asarray(local.intra_tags))
    is_ex = local.cell_tags == local.extra_tag
    return (np.unique(local.cells[is_in].ravel()).astype(np.int32),
            np.unique(local.cells[is_ex].ravel()).astype(np.int32))


class Layout:
    """Column layout of include/knpemi_b200.h for given dof maps."""

    def __init__(self, node_vert, n_owned_vertices):
        self.node_vert = node_vert
        self.n_loc = [int(v.size) for v in node_vert]
        self.n_own = [int((v < n_owned_vertices).sum()) for v in node_vert]
        self.n_gh = [self.n_loc[s] - self.n_own[s] for s in range(2)]
        self.rowbase = [0, 4 * self.n_own[0]]
        self.n_rows = 4 * (self.n_own[0] + self.n_own[1])
        self.gbase = [0, 4 * self.n_gh[0]]
        self.n_cols = self.n_rows + 4 * (self.n_gh[0] + self.n_gh[1])

    def col(self, s, f, q):
        q = np.asarray(q)
        own = q < self.n_own[s]
        return np.where(own, self.rowbase[s] + f * self.n_own[s] + q,
                        self.n_rows + self.gbase[s] + f * self.n_gh[s] + (q - self.n_own[s]))


def ghost_requests(local: Mesh, info, lay: Layout):
    """For each owner rank: the (subdomain, global vertex ids) this rank needs, ascending global id."""
    owner_of, l2g = info["owner_of"], info["l2g"]
    req = {}
    for s in range(2):
        gh_nodes = np.arange(lay.n_own[s], lay.n_loc[s])
        gv = l2g[lay.node_vert[s][gh_nodes]]
        ow = np.asarray(owner_of(gv))
        for r in np.unique(ow):
            sel = ow == r
            req.setdefault(int(r), {})[s] = (gv[sel], gh_nodes[sel])
    return req


def build_halo_lists(local: Mesh, info, lay: Layout, all_requests):
    """all_requests[r] = {owner: {s: global vertex ids}} as gathered from every rank.
    Returns (peers, send_ptr, send_cols, recv_ptr, recv_cols) in the order [s][f][vertex] per peer."""
    rank, l2g = info["rank"], info["l2g"]
    g2node = []
    for s in range(2):
        d = {}
        gv = l2g[lay.node_vert[s]]
        g2node.append(dict(zip(gv.tolist(), range(gv.size))))
    mine = ghost_requests(local, info, lay)
    peers = sorted(set(mine.keys()) | {r for r, rq in enumerate(all_requests) if rank in rq and r != rank})
    send_ptr, recv_ptr, send_cols, recv_cols = [0], [0], [], []
    for pr in peers:
        # what `pr` wants from me
        want = all_requests[pr].get(rank, {})
        for s in range(2):
            if s in want:
                nodes = np.array([g2node[s][g] for g in want[s].tolist()], np.int64)
                assert (nodes < lay.n_own[s]).all(), "peer requested a dof this rank does not own"
                for f in range(4):
                    send_cols.append(lay.col(s, f, nodes))
        send_ptr.append(int(sum(a.size for a in send_cols)))
        # what I want from `pr`
        if pr in mine:
            for s in range(2):
                if s in mine[pr]:
                    for f in range(4):
                        recv_cols.append(lay.col(s, f, mine[pr][s][1]))
        recv_ptr.append(int(sum(a.size for a in recv_cols)))
    cat = lambda lst: np.concatenate(lst).astype(np.int32) if lst else np.zeros(0, np.int32)
    return (np.array(peers, np.int32), np.array(send_ptr, np.int64), cat(send_cols),
            np.array(recv_ptr, np.int64), cat(recv_cols))


def gather_requests(comm, local, info, lay):
    mine = ghost_requests(local, info, lay)
    payload = {r: {s: v[0] for s, v in d.items()} for r, d in mine.items()}
    return comm.allgather(payload)


def init_halo(problem, ctx):
    """Exchange ghost requests, build the send/recv column lists and hand them (plus a NCCL unique id
    broadcast from rank 0) to the device context."""
    from . import lib as _lib
    from .comm import MPI
    comm = problem.comm
    local, info = problem.mesh, problem.halo
    lay = Layout(problem._node_vert, local.n_owned)
    assert lay.n_rows == ctx.n_rows and lay.n_cols == ctx.n_cols
    allreq = gather_requests(comm, local, info, lay)
    peers, sp, sc, rp, rc = build_halo_lists(local, info, lay, allreq)
    uid = _lib.nccl_unique_id() if comm.rank == 0 else None
    uid = comm.bcast(uid, root=0)
    n_phi = int(round(comm.allreduce(float(lay.n_own[0] + lay.n_own[1]), op=MPI.SUM)))
    ctx.dist_init(comm.rank, comm.size, uid, n_phi, peers, sp, sc, rp, rc)
    problem.layout = lay
