// Row-distributed smoothed-aggregation hierarchy (setup time, host, one call per rank; collective).
//
// The reference's preconditioner (hypre BoomerAMG on P, KNPEMIx_solver.py:269-273) runs across all MPI ranks with a halo
// per level.  This is our counterpart: every level operator is split by rows over the ranks (a rank owns the coarse
// nodes = aggregates of its own fine nodes), columns are [owned | ghosts grouped by owning rank], and one packed halo
// exchange per level SpMV moves the ghost values.  Aggregates never cross ranks: the strength graph, MIS(2) and the
// prolongator smoother see only the owned x owned block (ghost couplings are lumped into the diagonal of the filtered
// matrix like weak connections), so P and R = P^T are block diagonal over the ranks and need no communication, while
// A_{l+1} = R A_l P is the exact Galerkin product of the GLOBAL operator -- the couplings across rank boundaries stay in
// every level operator and every smoother.  Once a level has no more than `repl_threshold` rows globally it is gathered
// onto every rank and the serial hierarchy (amg_setup.cpp) continues redundantly: no communication below that level.
//
// Setup communication per level: the P rows of the boundary nodes go to the neighbours (they are the ghost rows of the
// extended prolongator in A_l P), and the ghost coarse ids each rank references go back as the next level's send lists.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <numeric>
#include "amg_host.h"

namespace knp {

namespace {

template <class T>
void put(std::vector<char>& buf, const T* p, size_t n) {
  const char* c = reinterpret_cast<const char*>(p);
  buf.insert(buf.end(), c, c + n * sizeof(T));
}
template <class T>
void put1(std::vector<char>& buf, T v) {
  put(buf, &v, 1);
}
template <class T>
void get(const std::vector<char>& buf, size_t& at, T* p, size_t n) {
  memcpy(p, buf.data() + at, n * sizeof(T));
  at += n * sizeof(T);
}

int sum_i64(AmgComm& comm, int64_t* v, int n) {
  std::vector<double> d(v, v + n);
  KNP_TRY(comm.allreduce(d.data(), n, false));
  for (int i = 0; i < n; ++i) v[i] = (int64_t)std::llround(d[i]);
  return KNP_OK;
}

}  // namespace

int amg_dist_setup(AmgComm& comm, CsrHost&& A0, HaloHost&& halo0, std::vector<int32_t>&& ghost_owner0,
                   std::vector<int32_t>&& ghost_oidx0, double theta, int64_t repl_threshold, int max_levels,
                   DistHierarchyHost& out) {
  const int R = comm.size, me = comm.rank;
  const double omega = 4.0 / 3.0;
  out.levels.clear();
  DistLevelHost cur;
  cur.n_own = A0.n_rows;
  cur.n_ghost = A0.n_cols - A0.n_rows;
  cur.A = std::move(A0);
  cur.halo = std::move(halo0);
  cur.ghost_owner = std::move(ghost_owner0);
  cur.ghost_oidx = std::move(ghost_oidx0);
  KNP_CHECK(cur.n_ghost >= 0 && (int)cur.ghost_owner.size() == cur.n_ghost && (int)cur.ghost_oidx.size() == cur.n_ghost,
            "distributed AMG: inconsistent ghost description");

  while (true) {
    const int n = cur.n_own;
    int64_t glob[2] = {n, cur.A.nnz()};
    KNP_TRY(sum_i64(comm, glob, 2));
    const int64_t n_glob = glob[0], nnz_glob = glob[1];
    if (n_glob <= repl_threshold || (int)out.levels.size() >= max_levels - 1) break;
    // owned x owned block for the strength graph and the aggregation
    CsrHost Aloc;
    Aloc.n_rows = Aloc.n_cols = n;
    Aloc.indptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) {
      int k = 0;
      for (int j = cur.A.indptr[i]; j < cur.A.indptr[i + 1]; ++j) k += cur.A.indices[j] < n;
      Aloc.indptr[i + 1] = Aloc.indptr[i] + k;
    }
    Aloc.indices.resize(Aloc.indptr[n]);
    Aloc.vals.resize(Aloc.indptr[n]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      int pos = Aloc.indptr[i];
      for (int j = cur.A.indptr[i]; j < cur.A.indptr[i + 1]; ++j)
        if (cur.A.indices[j] < n) {
          Aloc.indices[pos] = cur.A.indices[j];
          Aloc.vals[pos++] = cur.A.vals[j];
        }
    }
    // strength threshold halved until the GLOBAL strength graph has >= 3 edges per row (same rule as the serial setup)
    Graph S;
    double theta_l = theta;
    for (int attempt = 0; attempt < 4; ++attempt, theta_l *= 0.5) {
      strength_graph(Aloc, theta_l, S);
      int64_t e = (int64_t)S.idx.size();
      KNP_TRY(sum_i64(comm, &e, 1));
      if ((double)e >= 3.0 * (double)n_glob) break;
    }
    std::vector<int32_t> agg;
    int nagg = n > 0 ? mis2_aggregate(S, agg) : 0;
    if (n > 0 && out.levels.empty()) nagg = drop_dirichlet_aggregates(cur.A, agg, nagg);   // finest level only (amg_setup.cpp)
    int64_t nagg_glob = nagg;
    KNP_TRY(sum_i64(comm, &nagg_glob, 1));
    if ((double)nagg_glob >= 0.8 * (double)n_glob) break;      // coarsening stalled: replicate this level
    const bool filtered = out.levels.empty() || (double)nnz_glob > 32.0 * (double)n_glob;
    std::vector<double> dinv;
    double bounds[2] = {0.0, 0.0};
    prolongator_bounds(cur.A, n, S, filtered, dinv, bounds[0], bounds[1]);
    KNP_TRY(comm.allreduce(bounds, 2, true));
    cur.rho = bounds[0];
    prolongator_build(cur.A, n, S, agg, nagg, filtered, dinv, omega / bounds[1], cur.P);
    std::vector<int32_t>().swap(agg);
    Aloc = CsrHost();
    // P rows of my boundary nodes -> the neighbours that hold them as ghosts
    const HaloHost& H = cur.halo;
    const int np = (int)H.peers.size();
    std::vector<std::vector<char>> sbuf(R), rbuf;
    for (int i = 0; i < np; ++i) {
      std::vector<char>& b = sbuf[H.peers[i]];
      for (int64_t k = H.send_ptr[i]; k < H.send_ptr[i + 1]; ++k) {
        const int row = H.send_idx[k];
        const int32_t len = cur.P.indptr[row + 1] - cur.P.indptr[row];
        put1(b, len);
        put(b, cur.P.indices.data() + cur.P.indptr[row], len);
        put(b, cur.P.vals.data() + cur.P.indptr[row], len);
      }
    }
    KNP_TRY(comm.alltoallv(sbuf, rbuf));
    // ghost rows of the extended prolongator, columns still in the owners' coarse numbering
    std::vector<int32_t> g_len(cur.n_ghost, 0), g_col;
    std::vector<double> g_val;
    std::vector<std::vector<int32_t>> want(np);                 // per peer: coarse ids referenced (sorted unique)
    for (int i = 0; i < np; ++i) {
      const std::vector<char>& b = rbuf[H.peers[i]];
      size_t at = 0;
      for (int64_t k = H.recv_ptr[i]; k < H.recv_ptr[i + 1]; ++k) {
        KNP_CHECK(at + sizeof(int32_t) <= b.size(), "distributed AMG: short prolongator message from rank %d", H.peers[i]);
        int32_t len;
        get(b, at, &len, 1);
        g_len[k] = len;
        const size_t o = g_col.size();
        g_col.resize(o + len);
        g_val.resize(o + len);
        get(b, at, g_col.data() + o, len);
        get(b, at, g_val.data() + o, len);
        want[i].insert(want[i].end(), g_col.begin() + o, g_col.end());
      }
      std::sort(want[i].begin(), want[i].end());
      want[i].erase(std::unique(want[i].begin(), want[i].end()), want[i].end());
    }
    // ghost coarse numbering: grouped by peer (ascending rank), ascending coarse id inside a peer
    std::vector<int64_t> cg_off(np + 1, 0);
    for (int i = 0; i < np; ++i) cg_off[i + 1] = cg_off[i] + (int64_t)want[i].size();
    const int ncg = (int)cg_off[np];
    CsrHost Pext;
    Pext.n_rows = n + cur.n_ghost;
    Pext.n_cols = nagg + ncg;
    Pext.indptr.resize(Pext.n_rows + 1);
    std::copy(cur.P.indptr.begin(), cur.P.indptr.end(), Pext.indptr.begin());
    Pext.indices = cur.P.indices;
    Pext.vals = cur.P.vals;
    {
      size_t at = 0;
      for (int i = 0; i < np; ++i)
        for (int64_t k = H.recv_ptr[i]; k < H.recv_ptr[i + 1]; ++k) {
          for (int t = 0; t < g_len[k]; ++t, ++at) {
            const int32_t* lo = want[i].data();
            const int32_t pos = (int32_t)(std::lower_bound(lo, lo + want[i].size(), g_col[at]) - lo);
            Pext.indices.push_back(nagg + (int32_t)cg_off[i] + pos);
            Pext.vals.push_back(g_val[at]);
          }
          Pext.indptr[n + k + 1] = (int32_t)Pext.indices.size();
        }
    }
    std::vector<int32_t>().swap(g_col);
    std::vector<double>().swap(g_val);
    // Galerkin product of the global operator, my rows: A_c = P^T (A [P ; P_ghost])
    CsrHost AP, Ac;
    transpose(cur.P, cur.R);
    spgemm(cur.A, Pext, AP);
    Pext = CsrHost();
    spgemm(cur.R, AP, Ac);
    AP = CsrHost();
    // next level: ghosts = the coarse ids referenced on the peers; its send lists = what the peers reference here
    DistLevelHost nxt;
    nxt.n_own = nagg;
    nxt.n_ghost = ncg;
    nxt.A = std::move(Ac);
    nxt.A.n_cols = nagg + ncg;
    nxt.ghost_owner.resize(ncg);
    nxt.ghost_oidx.resize(ncg);
    for (auto& b : sbuf) b.clear();
    for (int i = 0; i < np; ++i) {
      for (size_t t = 0; t < want[i].size(); ++t) {
        nxt.ghost_owner[cg_off[i] + t] = H.peers[i];
        nxt.ghost_oidx[cg_off[i] + t] = want[i][t];
      }
      put(sbuf[H.peers[i]], want[i].data(), want[i].size());
    }
    KNP_TRY(comm.alltoallv(sbuf, rbuf));
    nxt.halo.send_ptr.assign(1, 0);
    nxt.halo.recv_ptr.assign(1, 0);
    for (int i = 0; i < np; ++i) {
      const std::vector<char>& b = rbuf[H.peers[i]];
      const size_t ns = b.size() / sizeof(int32_t);
      if (ns == 0 && want[i].empty()) continue;
      nxt.halo.peers.push_back(H.peers[i]);
      const size_t o = nxt.halo.send_idx.size();
      nxt.halo.send_idx.resize(o + ns);
      memcpy(nxt.halo.send_idx.data() + o, b.data(), ns * sizeof(int32_t));
      for (size_t t = o; t < o + ns; ++t)
        KNP_CHECK(nxt.halo.send_idx[t] >= 0 && nxt.halo.send_idx[t] < nagg, "distributed AMG: rank %d requested coarse node %d of %d",
                  H.peers[i], nxt.halo.send_idx[t], nagg);
      nxt.halo.send_ptr.push_back((int64_t)nxt.halo.send_idx.size());
      nxt.halo.recv_ptr.push_back(nxt.halo.recv_ptr.back() + (int64_t)want[i].size());
    }
    out.levels.push_back(std::move(cur));
    cur = std::move(nxt);
  }

  // ---- replicate the current level on every rank ----
  {
    const int n = cur.n_own;
    std::vector<char> mine;
    put1(mine, (int64_t)n);
    std::vector<std::vector<char>> all;
    KNP_TRY(comm.allgatherv(mine, all));
    out.repl_off.assign(R + 1, 0);
    for (int r = 0; r < R; ++r) {
      int64_t c;
      size_t at = 0;
      get(all[r], at, &c, 1);
      out.repl_off[r + 1] = out.repl_off[r] + c;
    }
    KNP_CHECK(out.repl_off[R] < ((int64_t)1 << 31), "replicated AMG level too large");
    mine.clear();
    std::vector<int32_t> lens(n), gcols(cur.A.indices.size());
    for (int i = 0; i < n; ++i) lens[i] = cur.A.indptr[i + 1] - cur.A.indptr[i];
    for (size_t j = 0; j < gcols.size(); ++j) {
      const int c = cur.A.indices[j];
      gcols[j] = c < n ? (int32_t)(out.repl_off[me] + c)
                       : (int32_t)(out.repl_off[cur.ghost_owner[c - n]] + cur.ghost_oidx[c - n]);
    }
    put(mine, lens.data(), lens.size());
    put(mine, gcols.data(), gcols.size());
    put(mine, cur.A.vals.data(), cur.A.vals.size());
    KNP_TRY(comm.allgatherv(mine, all));
    CsrHost& G = out.Arepl;
    const int N = (int)out.repl_off[R];
    G.n_rows = G.n_cols = N;
    G.indptr.assign(1, 0);
    G.indices.clear();
    G.vals.clear();
    for (int r = 0; r < R; ++r) {
      const int nr = (int)(out.repl_off[r + 1] - out.repl_off[r]);
      size_t at = 0;
      std::vector<int32_t> len(nr);
      get(all[r], at, len.data(), nr);
      int64_t tot = 0;
      for (int v : len) tot += v;
      KNP_CHECK(at + (size_t)tot * 12 == all[r].size(), "distributed AMG: bad replication message from rank %d", r);
      const size_t o = G.indices.size();
      G.indices.resize(o + tot);
      G.vals.resize(o + tot);
      get(all[r], at, G.indices.data() + o, tot);
      get(all[r], at, G.vals.data() + o, tot);
      for (int i = 0; i < nr; ++i) G.indptr.push_back(G.indptr.back() + len[i]);
    }
    // rows sorted by global column (ghost columns interleave with owned ones in the global numbering)
#pragma omp parallel
    {
      std::vector<std::pair<int32_t, double>> row;
#pragma omp for schedule(dynamic, 1024)
      for (int i = 0; i < N; ++i) {
        const int a = G.indptr[i], b = G.indptr[i + 1];
        bool sorted = true;
        for (int j = a + 1; j < b; ++j) sorted = sorted && G.indices[j - 1] < G.indices[j];
        if (sorted) continue;
        row.resize(b - a);
        for (int j = a; j < b; ++j) row[j - a] = {G.indices[j], G.vals[j]};
        std::sort(row.begin(), row.end());
        for (int j = a; j < b; ++j) {
          G.indices[j] = row[j - a].first;
          G.vals[j] = row[j - a].second;
        }
      }
    }
  }
  return KNP_OK;
}

}  // namespace knp
