// Host-side building blocks of the smoothed-aggregation setup, shared by the serial (amg_setup.cpp) and the
// row-distributed (amg_dist.cpp) hierarchy construction.
#pragma once
#include "common.cuh"

namespace knp {

const char* last_error();

struct Graph {
  int n = 0;
  std::vector<int32_t> ptr, idx;
};

void strength_graph(const CsrHost& A, double theta, Graph& S);
int mis2_aggregate(const Graph& S, std::vector<int32_t>& agg);
int drop_dirichlet_aggregates(const CsrHost& A, std::vector<int32_t>& agg, int nagg);
void spgemm(const CsrHost& A, const CsrHost& B, CsrHost& C);
void transpose(const CsrHost& A, CsrHost& At);
int dense_inverse(int n, std::vector<double>& M, std::vector<double>& inv);
void prolongator_bounds(const CsrHost& A, int n_own_cols, const Graph& S, bool filtered, std::vector<double>& dinv,
                        double& rho, double& rhoF);
void prolongator_build(const CsrHost& A, int n_own_cols, const Graph& S, const std::vector<int32_t>& agg, int nagg,
                       bool filtered, const std::vector<double>& dinv, double sc, CsrHost& P);

// ---- row-distributed hierarchy (amg_dist.cpp) -----------------------------------------------------------------------
// Communication needed by the distributed setup, abstracted so that the CPU test tier can run R simulated ranks in one
// process (threads + mailboxes) while the product talks NCCL.
struct AmgComm {
  int rank = 0, size = 1;
  virtual ~AmgComm() {}
  // send[q] goes to rank q (empty = nothing); recv[q] is what rank q sent here.  Collective.
  virtual int alltoallv(const std::vector<std::vector<char>>& send, std::vector<std::vector<char>>& recv) = 0;
  virtual int allreduce(double* v, int n, bool take_max) = 0;
  virtual int allgatherv(const std::vector<char>& mine, std::vector<std::vector<char>>& all) = 0;
};

// ghost exchange pattern of a vector laid out [owned | ghosts]: ghosts are grouped by peer (ascending rank) in the order
// the peer packs them, so a receive lands in place at n_own + recv_ptr[i]
struct HaloHost {
  std::vector<int32_t> peers;
  std::vector<int64_t> send_ptr, recv_ptr;   // per peer (size peers + 1)
  std::vector<int32_t> send_idx;             // owned indices to pack, grouped by peer
};

struct DistLevelHost {
  int n_own = 0, n_ghost = 0;
  CsrHost A;                                  // n_own x (n_own + n_ghost)
  CsrHost P, R;                               // rank-local: P is n_own x n_own(next level), R = P^T
  HaloHost halo;
  std::vector<int32_t> ghost_owner, ghost_oidx;   // per ghost: owning rank and its index there
  double rho = 2.0;
};

struct DistHierarchyHost {
  std::vector<DistLevelHost> levels;          // the distributed levels 0 .. L-1
  // level L is replicated: every rank holds the whole operator (rows ordered by rank) and runs the serial hierarchy on it
  CsrHost Arepl;
  std::vector<int64_t> repl_off;              // size + 1 row offsets of the ranks inside the replicated level
};

// Smoothed aggregation with rank-local aggregates: the strength graph, MIS(2) and the prolongator smoother ignore ghost
// columns (lumped into the diagonal), so P and R = P^T are block diagonal over the ranks, while every level operator is the
// exact Galerkin product R A P of the GLOBAL operator (ghost couplings kept on every level).
int amg_dist_setup(AmgComm& comm, CsrHost&& A0, HaloHost&& halo0, std::vector<int32_t>&& ghost_owner0,
                   std::vector<int32_t>&& ghost_oidx0, double theta, int64_t repl_threshold, int max_levels,
                   DistHierarchyHost& out);

}  // namespace knp
