// Host-only harness for the CPU test tier: runs the row-distributed hierarchy setup (amg_dist.cpp) on R SIMULATED ranks
// (threads + mailboxes instead of NCCL) for a global matrix split by an owner array, and assembles the per-rank pieces back
// into global level operators and prolongators so that tests can check the Galerkin identity, the R = 1 equivalence with
// the serial setup and the convergence of the resulting cycle without a GPU.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include "amg_host.h"

namespace knp {

struct SimWorld {
  int R = 1;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long gen = 0;
  std::vector<std::vector<std::vector<char>>> box;   // box[src][dst]
  std::vector<std::vector<double>> red;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const long g = gen;
    if (++arrived == R) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

struct SimComm : AmgComm {
  SimWorld* w;
  int alltoallv(const std::vector<std::vector<char>>& send, std::vector<std::vector<char>>& recv) override {
    w->box[rank] = send;
    w->box[rank].resize(size);
    w->barrier();
    recv.assign(size, {});
    for (int q = 0; q < size; ++q) recv[q] = w->box[q][rank];
    w->barrier();
    return KNP_OK;
  }
  int allreduce(double* v, int n, bool take_max) override {
    w->red[rank].assign(v, v + n);
    w->barrier();
    for (int i = 0; i < n; ++i) {
      double a = w->red[0][i];
      for (int q = 1; q < size; ++q) a = take_max ? std::max(a, w->red[q][i]) : a + w->red[q][i];
      v[i] = a;
    }
    w->barrier();
    return KNP_OK;
  }
  int allgatherv(const std::vector<char>& mine, std::vector<std::vector<char>>& all) override {
    w->box[rank].assign(1, mine);
    w->barrier();
    all.assign(size, {});
    for (int q = 0; q < size; ++q) all[q] = w->box[q][0];
    w->barrier();
    return KNP_OK;
  }
};

// assembled result of the last knp_amg_dist_sim_host call
struct SimResult {
  std::vector<CsrHost> A, P;          // global level operators (dist levels + the replicated one) and prolongators
  std::vector<double> rho;
  std::vector<int32_t> perm0;         // global level-0 numbering: new index -> input index
};
static SimResult g_sim;

}  // namespace knp

using namespace knp;

extern "C" {

// Splits the n x n CSR matrix by owner[] (rank r owns the rows with owner == r, ascending), runs amg_dist_setup on nranks
// simulated ranks and keeps the assembled hierarchy for knp_amg_dist_sim_level.  n_levels counts the distributed levels
// plus the replicated one.
int knp_amg_dist_sim_host(int32_t nranks, int32_t n, const int32_t* indptr, const int32_t* indices, const double* vals,
                          const int32_t* owner, double theta, int64_t repl_threshold, int32_t* n_levels) {
  KNP_CHECK(nranks >= 1 && n > 0 && indptr && indices && vals && owner && n_levels, "bad arguments");
  const int R = nranks;
  // rank-ordered global numbering of level 0
  std::vector<int32_t> newid(n), start(R + 1, 0);
  for (int i = 0; i < n; ++i) {
    KNP_CHECK(owner[i] >= 0 && owner[i] < R, "owner out of range");
    ++start[owner[i] + 1];
  }
  for (int r = 0; r < R; ++r) start[r + 1] += start[r];
  {
    std::vector<int32_t> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) newid[i] = fill[owner[i]]++;
  }
  g_sim = SimResult();
  g_sim.perm0.resize(n);
  for (int i = 0; i < n; ++i) g_sim.perm0[newid[i]] = i;
  // per-rank inputs
  struct RankIn {
    CsrHost A;
    HaloHost halo;
    std::vector<int32_t> gown, goidx;
  };
  std::vector<RankIn> in(R);
  std::vector<std::vector<std::vector<int32_t>>> req(R, std::vector<std::vector<int32_t>>(R));   // req[r][q]: local ids on q that r needs
  for (int r = 0; r < R; ++r) {
    const int n_own = start[r + 1] - start[r];
    // ghosts: columns owned by other ranks, grouped by owner, ascending owner-local id
    std::vector<std::pair<int32_t, int32_t>> gh;   // (owner, owner-local id)
    for (int li = 0; li < n_own; ++li) {
      const int i = g_sim.perm0[start[r] + li];
      for (int j = indptr[i]; j < indptr[i + 1]; ++j)
        if (owner[indices[j]] != r) gh.push_back({owner[indices[j]], newid[indices[j]] - start[owner[indices[j]]]});
    }
    std::sort(gh.begin(), gh.end());
    gh.erase(std::unique(gh.begin(), gh.end()), gh.end());
    RankIn& I = in[r];
    I.A.n_rows = n_own;
    I.A.n_cols = n_own + (int)gh.size();
    I.A.indptr.assign(1, 0);
    for (int li = 0; li < n_own; ++li) {
      const int i = g_sim.perm0[start[r] + li];
      std::vector<std::pair<int32_t, double>> row;
      for (int j = indptr[i]; j < indptr[i + 1]; ++j) {
        const int c = indices[j], oc = owner[c];
        int32_t lc;
        if (oc == r) lc = newid[c] - start[r];
        else lc = n_own + (int32_t)(std::lower_bound(gh.begin(), gh.end(), std::make_pair((int32_t)oc, (int32_t)(newid[c] - start[oc]))) - gh.begin());
        row.push_back({lc, vals[j]});
      }
      std::sort(row.begin(), row.end());
      for (auto& e : row) {
        I.A.indices.push_back(e.first);
        I.A.vals.push_back(e.second);
      }
      I.A.indptr.push_back((int32_t)I.A.indices.size());
    }
    I.halo.recv_ptr.assign(1, 0);
    for (size_t k = 0; k < gh.size(); ++k) {
      I.gown.push_back(gh[k].first);
      I.goidx.push_back(gh[k].second);
      req[r][gh[k].first].push_back(gh[k].second);
    }
  }
  for (int r = 0; r < R; ++r) {
    HaloHost& H = in[r].halo;
    H.send_ptr.assign(1, 0);
    H.recv_ptr.assign(1, 0);
    for (int q = 0; q < R; ++q) {
      if (q == r || (req[r][q].empty() && req[q][r].empty())) continue;
      H.peers.push_back(q);
      H.send_idx.insert(H.send_idx.end(), req[q][r].begin(), req[q][r].end());
      H.send_ptr.push_back((int64_t)H.send_idx.size());
      H.recv_ptr.push_back(H.recv_ptr.back() + (int64_t)req[r][q].size());
    }
  }
  // run the ranks
  SimWorld world;
  world.R = R;
  world.box.assign(R, {});
  world.red.assign(R, {});
  std::vector<DistHierarchyHost> H(R);
  std::vector<int> rc(R, KNP_OK);
  std::vector<std::string> err(R);
  std::vector<std::thread> th;
  for (int r = 0; r < R; ++r)
    th.emplace_back([&, r] {
      SimComm comm;
      comm.rank = r;
      comm.size = R;
      comm.w = &world;
      rc[r] = amg_dist_setup(comm, std::move(in[r].A), std::move(in[r].halo), std::move(in[r].gown), std::move(in[r].goidx),
                             theta, repl_threshold, 16, H[r]);
      if (rc[r] != KNP_OK) err[r] = last_error();
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < R; ++r)
    if (rc[r] != KNP_OK) {
      set_error("simulated rank %d: %s", r, err[r].c_str());
      return rc[r];
    }
  // assemble the global levels
  const int L = (int)H[0].levels.size();
  for (int r = 1; r < R; ++r) KNP_CHECK((int)H[r].levels.size() == L, "ranks disagree on the number of distributed levels");
  std::vector<std::vector<int64_t>> off(L + 1, std::vector<int64_t>(R + 1, 0));
  for (int l = 0; l < L; ++l)
    for (int r = 0; r < R; ++r) off[l][r + 1] = off[l][r] + H[r].levels[l].n_own;
  off[L] = H[0].repl_off;
  for (int l = 0; l < L; ++l) {
    CsrHost A, P;
    A.n_rows = A.n_cols = (int)off[l][R];
    P.n_rows = (int)off[l][R];
    P.n_cols = (int)off[l + 1][R];
    A.indptr.assign(1, 0);
    P.indptr.assign(1, 0);
    for (int r = 0; r < R; ++r) {
      const DistLevelHost& D = H[r].levels[l];
      KNP_CHECK(off[l + 1][r + 1] - off[l + 1][r] == D.P.n_cols, "coarse sizes disagree");
      for (int i = 0; i < D.n_own; ++i) {
        std::vector<std::pair<int32_t, double>> row;
        for (int j = D.A.indptr[i]; j < D.A.indptr[i + 1]; ++j) {
          const int c = D.A.indices[j];
          const int64_t g = c < D.n_own ? off[l][r] + c : off[l][D.ghost_owner[c - D.n_own]] + D.ghost_oidx[c - D.n_own];
          row.push_back({(int32_t)g, D.A.vals[j]});
        }
        std::sort(row.begin(), row.end());
        for (auto& e : row) {
          A.indices.push_back(e.first);
          A.vals.push_back(e.second);
        }
        A.indptr.push_back((int32_t)A.indices.size());
        for (int j = D.P.indptr[i]; j < D.P.indptr[i + 1]; ++j) {
          P.indices.push_back((int32_t)(off[l + 1][r] + D.P.indices[j]));
          P.vals.push_back(D.P.vals[j]);
        }
        P.indptr.push_back((int32_t)P.indices.size());
      }
    }
    g_sim.A.push_back(std::move(A));
    g_sim.P.push_back(std::move(P));
    g_sim.rho.push_back(H[0].levels[l].rho);
  }
  for (int r = 1; r < R; ++r)
    KNP_CHECK(H[r].Arepl.indices == H[0].Arepl.indices && H[r].Arepl.vals == H[0].Arepl.vals, "replicated level differs between ranks");
  g_sim.A.push_back(H[0].Arepl);
  *n_levels = L + 1;
  return KNP_OK;
}

// which = 0: level operator A_l, 1: prolongator P_l (l < n_levels - 1).  Call with NULL arrays for the sizes first.
int knp_amg_dist_sim_level(int32_t level, int32_t which, int64_t* n_rows, int64_t* n_cols, int64_t* nnz, double* rho,
                           int32_t* indptr, int32_t* indices, double* vals) {
  const std::vector<CsrHost>& V = which == 0 ? g_sim.A : g_sim.P;
  KNP_CHECK(level >= 0 && level < (int)V.size(), "no such level");
  const CsrHost& M = V[level];
  if (n_rows) *n_rows = M.n_rows;
  if (n_cols) *n_cols = M.n_cols;
  if (nnz) *nnz = M.nnz();
  if (rho) *rho = level < (int)g_sim.rho.size() ? g_sim.rho[level] : 0.0;
  if (indptr) memcpy(indptr, M.indptr.data(), M.indptr.size() * sizeof(int32_t));
  if (indices) memcpy(indices, M.indices.data(), M.indices.size() * sizeof(int32_t));
  if (vals) memcpy(vals, M.vals.data(), M.vals.size() * sizeof(double));
  return KNP_OK;
}

int knp_amg_dist_sim_perm(int32_t* perm0) {
  KNP_CHECK(perm0, "NULL argument");
  memcpy(perm0, g_sim.perm0.data(), g_sim.perm0.size() * sizeof(int32_t));
  return KNP_OK;
}

}  // extern "C"
