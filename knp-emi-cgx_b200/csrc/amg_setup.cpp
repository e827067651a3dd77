// Smoothed-aggregation AMG hierarchy construction (setup time, host, OpenMP).
//
// The reference preconditions GMRES with one hypre BoomerAMG V-cycle on the block-diagonal matrix P,
// which is assembled ONCE (reassemble_P = False; KNPEMIx_solver.py:33-34,118-135,269-273,358-362,386).
// hypre is a third-party library that is not vendored in the reference; this is our own algorithm:
// MIS(2) aggregation with deterministic hashed priorities, constant tentative prolongator smoothed by one
// damped-Jacobi step (with the filtered matrix on dense Galerkin levels), Galerkin coarse operators, dense inverse on the coarsest level.  The V-cycle itself
// (all per-iteration work) runs on the GPU (solver.cu).  oracle/amg.py restates the same algorithm in
// numpy/scipy and tests compare the two hierarchies level by level.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include "common.cuh"
#include "amg_host.h"

namespace knp {

// KNP_AMG_TIMING=1 prints the setup phases per level (host profiling aid)
struct PhaseTimer {
  bool on = getenv("KNP_AMG_TIMING") && atoi(getenv("KNP_AMG_TIMING"));
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what, int level, int n) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "amg setup level %d (n = %d): %-10s %.3f s\n", level, n, what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

static inline int64_t hash32(int64_t i) {
  uint64_t x = ((uint64_t)i + 0x9E3779B9ull) & 0xFFFFFFFFull;
  x = ((x ^ (x >> 16)) * 0x85EBCA6Bull) & 0xFFFFFFFFull;
  x = ((x ^ (x >> 13)) * 0xC2B2AE35ull) & 0xFFFFFFFFull;
  x = x ^ (x >> 16);
  return (int64_t)x;
}

// symmetric strength graph |a_ij| >= theta sqrt(|a_ii a_jj|), i != j, a_ij != 0; symmetrised (S + S^T)
void strength_graph(const CsrHost& A, double theta, Graph& S) {
  const int n = A.n_rows;
  std::vector<double> d(n, 0.0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    double v = 0.0;
    for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j)
      if (A.indices[j] == i) v += A.vals[j];
    d[i] = std::fabs(v);
  }
  // directed strong edges
  std::vector<int32_t> cnt(n + 1, 0);
  std::vector<uint8_t> strong(A.indices.size(), 0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i)
    for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
      const int c = A.indices[j];
      const double v = A.vals[j];
      if (c != i && v != 0.0 && std::fabs(v) >= theta * std::sqrt(d[i] * d[c])) strong[j] = 1;
    }
  // symmetrise S + S^T.  Fast path (structurally symmetric pattern with sorted rows -- every matrix of this library):
  // entry (i,c) is strong if it or its transposed entry (c,i), found by binary search in row c, passes the test; rows
  // are independent.  A missing transposed entry switches to the general (serial) construction below.
  bool sym = true;
  std::vector<uint8_t> ssym(strong);          // written per row, `strong` is only read: no race
#pragma omp parallel for schedule(static) reduction(&& : sym)
  for (int i = 0; i < n; ++i)
    for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
      const int c = A.indices[j];
      if (c == i) continue;
      if (j > A.indptr[i] && A.indices[j - 1] >= c) sym = false;          // unsorted row
      const int32_t* rb = A.indices.data() + A.indptr[c];
      const int32_t* re = A.indices.data() + A.indptr[c + 1];
      const int32_t* it = std::lower_bound(rb, re, i);
      if (it == re || *it != i) {
        if (strong[j]) sym = false;                                        // a strong entry without a transposed partner
      } else if (strong[it - A.indices.data()]) {
        ssym[j] = 1;                                                       // strong through the transposed entry
      }
    }
  if (sym) {
    S.n = n;
    S.ptr.assign(n + 1, 0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      int k = 0;
      for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) k += ssym[j] != 0;
      cnt[i + 1] = k;
    }
    for (int i = 0; i < n; ++i) S.ptr[i + 1] = S.ptr[i] + cnt[i + 1];
    S.idx.resize(S.ptr[n]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      int pos = S.ptr[i];
      for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j)
        if (ssym[j]) S.idx[pos++] = A.indices[j];
    }
    return;
  }
  std::vector<uint8_t>().swap(ssym);
  // general case: collect (i,c) and (c,i)
  for (int i = 0; i < n; ++i)
    for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j)
      if (strong[j]) {
        ++cnt[i + 1];
        ++cnt[A.indices[j] + 1];
      }
  std::vector<int64_t> ptr(n + 1, 0);
  for (int i = 0; i < n; ++i) ptr[i + 1] = ptr[i] + cnt[i + 1];
  std::vector<int32_t> tmp(ptr[n]);
  {
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j)
        if (strong[j]) {
          const int c = A.indices[j];
          tmp[fill[i]++] = c;
          tmp[fill[c]++] = i;
        }
  }
  S.n = n;
  S.ptr.assign(n + 1, 0);
  std::vector<int32_t> ucnt(n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    std::sort(tmp.begin() + ptr[i], tmp.begin() + ptr[i + 1]);
    ucnt[i] = (int32_t)(std::unique(tmp.begin() + ptr[i], tmp.begin() + ptr[i + 1]) - (tmp.begin() + ptr[i]));
  }
  for (int i = 0; i < n; ++i) S.ptr[i + 1] = S.ptr[i] + ucnt[i];
  S.idx.resize(S.ptr[n]);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) std::copy(tmp.begin() + ptr[i], tmp.begin() + ptr[i] + ucnt[i], S.idx.begin() + S.ptr[i]);
}

static void nbr_max(const Graph& S, const std::vector<int64_t>& key, std::vector<int64_t>& out) {
  const int n = S.n;
  out.resize(n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    int64_t m = key[i];
    for (int j = S.ptr[i]; j < S.ptr[i + 1]; ++j) m = std::max(m, key[S.idx[j]]);
    out[i] = m;
  }
}

// MIS(2) aggregation; identical decisions to oracle/amg.py::mis2_aggregate
int mis2_aggregate(const Graph& S, std::vector<int32_t>& agg) {
  const int n = S.n;
  std::vector<int64_t> pr(n), key(n), k1, k2;
  std::vector<int8_t> state(n, 0);
  const int64_t BIG = (int64_t)1 << 62;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) pr[i] = ((hash32(i) & 0x3FFFFFFFll) << 31) | (int64_t)i;
  while (true) {
    bool any = false;
#pragma omp parallel for schedule(static) reduction(|| : any)
    for (int i = 0; i < n; ++i) {
      any = any || state[i] == 0;
      key[i] = state[i] == 1 ? BIG + pr[i] : (state[i] == 0 ? pr[i] : -1);
    }
    if (!any) break;
    nbr_max(S, key, k1);
    nbr_max(S, k1, k2);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i)
      if (state[i] == 0 && k2[i] == key[i]) state[i] = 1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) key[i] = state[i] == 1 ? BIG + pr[i] : -1;
    nbr_max(S, key, k1);
    nbr_max(S, k1, k2);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i)
      if (state[i] == 0 && k2[i] >= BIG) state[i] = -1;
  }
  agg.assign(n, -1);
  int nroots = 0;
  for (int i = 0; i < n; ++i)
    if (state[i] == 1) agg[i] = nroots++;
  // root priority per node (for joined nodes: priority of their aggregate's root)
  std::vector<int64_t> rootpr(n, -1);
  for (int i = 0; i < n; ++i)
    if (state[i] == 1) rootpr[i] = pr[i];
  for (int round = 0; round < 2; ++round) {
    nbr_max(S, rootpr, k1);
    std::vector<int32_t> newagg(agg);
    std::vector<int64_t> newpr(rootpr);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i)
      if (agg[i] < 0 && k1[i] >= 0) {
        const int root = (int)(k1[i] & (((int64_t)1 << 31) - 1));
        newagg[i] = agg[root];
        newpr[i] = k1[i];
      }
    agg.swap(newagg);
    rootpr.swap(newpr);
  }
  int nagg = nroots;
  for (int i = 0; i < n; ++i)
    if (agg[i] < 0) agg[i] = nagg++;
  return nagg;
}

// Dirichlet rows -- rows without a non-zero off-diagonal entry, as the boundary conditions leave them (knp_set_dirichlet) --
// are taken out of the coarse space: agg = -1, an empty row of the prolongator.  The level's Jacobi sweeps solve them, and
// their singleton aggregates would otherwise survive on every level (coarsening stalls at the number of boundary dofs).
// Applied on the finest level only: a coarse row that has lost its couplings is a whole connected component (the ion
// block of one biological cell, say) and keeps its singleton aggregate, so that the coarsest solve treats it exactly.
// Identical decisions to oracle/amg.py::drop_dirichlet_aggregates; a no-op for matrices without such rows.
int drop_dirichlet_aggregates(const CsrHost& A, std::vector<int32_t>& agg, int nagg) {
  const int n = A.n_rows;
  std::vector<uint8_t> dir(n, 0);
  bool any = false;
#pragma omp parallel for schedule(static) reduction(|| : any)
  for (int i = 0; i < n; ++i) {
    bool off = false;
    for (int j = A.indptr[i]; j < A.indptr[i + 1] && !off; ++j) off = A.indices[j] != i && A.vals[j] != 0.0;
    dir[i] = !off;
    any = any || !off;
  }
  if (!any) return nagg;
  std::vector<int32_t> remap(nagg, -1);
  for (int i = 0; i < n; ++i)
    if (!dir[i]) remap[agg[i]] = 0;
  int k = 0;
  for (int a = 0; a < nagg; ++a)
    if (remap[a] == 0) remap[a] = k++;
  for (int i = 0; i < n; ++i) agg[i] = dir[i] ? -1 : remap[agg[i]];
  return k;
}

// C = A * B (CSR, sorted columns out): two passes over the rows (count, then fill) with a dense marker / accumulator
// per thread -- no per-row allocations
void spgemm(const CsrHost& A, const CsrHost& B, CsrHost& C) {
  const int n = A.n_rows, mcols = B.n_cols;
  C.n_rows = n;
  C.n_cols = mcols;
  C.indptr.assign(n + 1, 0);
  std::vector<int32_t> cnt(n, 0);
#pragma omp parallel
  {
    std::vector<int32_t> mark(mcols, -1);
#pragma omp for schedule(dynamic, 2048)
    for (int i = 0; i < n; ++i) {
      int k = 0;
      for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
        const int r = A.indices[j];
        for (int l = B.indptr[r]; l < B.indptr[r + 1]; ++l) {
          const int c = B.indices[l];
          if (mark[c] != i) {
            mark[c] = i;
            ++k;
          }
        }
      }
      cnt[i] = k;
    }
  }
  for (int i = 0; i < n; ++i) C.indptr[i + 1] = C.indptr[i] + cnt[i];
  C.indices.resize(C.indptr[n]);
  C.vals.resize(C.indptr[n]);
#pragma omp parallel
  {
    std::vector<double> acc(mcols, 0.0);
    std::vector<int32_t> mark(mcols, -1);
#pragma omp for schedule(dynamic, 2048)
    for (int i = 0; i < n; ++i) {
      int32_t* list = C.indices.data() + C.indptr[i];
      int k = 0;
      for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
        const int r = A.indices[j];
        const double a = A.vals[j];
        for (int l = B.indptr[r]; l < B.indptr[r + 1]; ++l) {
          const int c = B.indices[l];
          if (mark[c] != i) {
            mark[c] = i;
            acc[c] = 0.0;
            list[k++] = c;
          }
          acc[c] += a * B.vals[l];
        }
      }
      std::sort(list, list + k);
      double* v = C.vals.data() + C.indptr[i];
      for (int t = 0; t < k; ++t) v[t] = acc[list[t]];
    }
  }
}

void transpose(const CsrHost& A, CsrHost& At) {
  const int n = A.n_rows, m = A.n_cols;
  At.n_rows = m;
  At.n_cols = n;
  At.indptr.assign(m + 1, 0);
  for (size_t j = 0; j < A.indices.size(); ++j) ++At.indptr[A.indices[j] + 1];
  for (int i = 0; i < m; ++i) At.indptr[i + 1] += At.indptr[i];
  At.indices.resize(A.indices.size());
  At.vals.resize(A.indices.size());
  std::vector<int32_t> fill(At.indptr.begin(), At.indptr.end() - 1);
  for (int i = 0; i < n; ++i)
    for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
      const int c = A.indices[j];
      At.indices[fill[c]] = i;
      At.vals[fill[c]] = A.vals[j];
      ++fill[c];
    }
}

// dense inverse by LU with partial pivoting (coarsest level only; n <= a few thousand)
int dense_inverse(int n, std::vector<double>& M, std::vector<double>& inv) {
  inv.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(M[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(M[(size_t)i * n + k]) > best) {
        best = std::fabs(M[(size_t)i * n + k]);
        piv = i;
      }
    if (best == 0.0) {
      set_error("AMG coarsest-level matrix is singular (column %d)", k);
      return KNP_E_INVALID;
    }
    if (piv != k)
      for (int j = 0; j < n; ++j) {
        std::swap(M[(size_t)k * n + j], M[(size_t)piv * n + j]);
        std::swap(inv[(size_t)k * n + j], inv[(size_t)piv * n + j]);
      }
    const double d = 1.0 / M[(size_t)k * n + k];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double f = M[(size_t)i * n + k] * d;
      if (f == 0.0) continue;
      for (int j = k; j < n; ++j) M[(size_t)i * n + j] -= f * M[(size_t)k * n + j];
      for (int j = 0; j < n; ++j) inv[(size_t)i * n + j] -= f * inv[(size_t)k * n + j];
    }
    for (int j = k; j < n; ++j) M[(size_t)k * n + j] *= d;
    for (int j = 0; j < n; ++j) inv[(size_t)k * n + j] *= d;
  }
  return KNP_OK;
}

// Gershgorin bounds on rho(D^-1 A) (used by the Jacobi smoother of the cycle) and on rho(D^-1 A_F), and D^-1, for the
// rows of A.  A may carry ghost columns (index >= n_own_cols, distributed levels): they count for rho and are lumped into
// the diagonal of A_F like weak connections.
void prolongator_bounds(const CsrHost& A, int n_own_cols, const Graph& S, bool filtered, std::vector<double>& dinv,
                        double& rho_out, double& rhoF_out) {
  const int n = A.n_rows;
  dinv.resize(n);
  double rho = 0.0, rhoF = 0.0;
#pragma omp parallel for schedule(static) reduction(max : rho, rhoF)
  for (int i = 0; i < n; ++i) {
    double d = 0.0, s = 0.0;
    for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
      if (A.indices[j] == i) d += A.vals[j];
      s += std::fabs(A.vals[j]);
    }
    dinv[i] = 1.0 / d;
    rho = std::max(rho, std::fabs(dinv[i]) * s);
    double diagF = 0.0, sabs = 0.0;
    int sp = S.ptr[i];
    const int se = S.ptr[i + 1];
    for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
      const int c = A.indices[j];
      const double v = A.vals[j];
      while (sp < se && S.idx[sp] < c) ++sp;
      const bool strong = c < n_own_cols && (!filtered || (sp < se && S.idx[sp] == c));
      if (c == i || !strong) diagF += v;
      else sabs += std::fabs(v);
    }
    rhoF = std::max(rhoF, std::fabs(dinv[i]) * (std::fabs(diagF) + sabs));
  }
  rho_out = rho;
  rhoF_out = rhoF;
}

// Prolongator smoothing with the FILTERED matrix A_F (strong off-diagonals only, weak ones -- and ghost columns -- lumped
// into the diagonal so that row sums are kept): P = T - sc D^-1 A_F T, T(i, agg[i]) = 1, sc = omega / rho_F.  Without the
// filter the Galerkin operators of 3D meshes fill in (hundreds of entries per row on level 2) and coarsening stalls.
// Two passes over the rows (count, then fill) with a dense marker / accumulator per thread.
void prolongator_build(const CsrHost& A, int n_own_cols, const Graph& S, const std::vector<int32_t>& agg, int nagg,
                       bool filtered, const std::vector<double>& dinv, double sc, CsrHost& P) {
  const int n = A.n_rows;
  P.n_rows = n;
  P.n_cols = nagg;
  P.indptr.assign(n + 1, 0);
  std::vector<int32_t> plen(n);
#pragma omp parallel
  {
    std::vector<int32_t> mark(nagg, -1);
#pragma omp for schedule(static)
    for (int i = 0; i < n; ++i) {
      int sp = S.ptr[i], k = 1;
      const int se = S.ptr[i + 1];
      if (agg[i] < 0) {                  // Dirichlet row: not interpolated
        plen[i] = 0;
        continue;
      }
      mark[agg[i]] = i;
      for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
        const int c = A.indices[j];
        while (sp < se && S.idx[sp] < c) ++sp;
        const bool strong = c < n_own_cols && agg[c] >= 0 && (!filtered || (sp < se && S.idx[sp] == c));
        if (c != i && strong && mark[agg[c]] != i) {
          mark[agg[c]] = i;
          ++k;
        }
      }
      plen[i] = k;
    }
  }
  for (int i = 0; i < n; ++i) P.indptr[i + 1] = P.indptr[i] + plen[i];
  P.indices.resize(P.indptr[n]);
  P.vals.resize(P.indptr[n]);
#pragma omp parallel
  {
    std::vector<int32_t> mark(nagg, -1);
    std::vector<double> acc(nagg, 0.0);
#pragma omp for schedule(static)
    for (int i = 0; i < n; ++i) {
      int32_t* list = P.indices.data() + P.indptr[i];
      int sp = S.ptr[i], k = 0;
      const int se = S.ptr[i + 1];
      if (agg[i] < 0) continue;
      mark[agg[i]] = i;
      acc[agg[i]] = 0.0;
      list[k++] = agg[i];
      double diagF = 0.0;
      for (int j = A.indptr[i]; j < A.indptr[i + 1]; ++j) {
        const int c = A.indices[j];
        const double v = A.vals[j];
        while (sp < se && S.idx[sp] < c) ++sp;
        const bool strong = c < n_own_cols && agg[c] >= 0 && (!filtered || (sp < se && S.idx[sp] == c));
        if (c == i || !strong) {
          diagF += v;
        } else {
          const int g = agg[c];
          if (mark[g] != i) {
            mark[g] = i;
            acc[g] = 0.0;
            list[k++] = g;
          }
          acc[g] += v;
        }
      }
      acc[agg[i]] += diagF;
      std::sort(list, list + k);
      double* pv = P.vals.data() + P.indptr[i];
      for (int t = 0; t < k; ++t) {
        double v = -(sc * dinv[i]) * acc[list[t]];
        if (list[t] == agg[i]) v += 1.0;
        pv[t] = v;
      }
    }
  }
}

int amg_setup_host(const CsrHost& A0, double theta, int coarse_size, int max_levels, std::vector<CsrHost>& As,
                   std::vector<CsrHost>& Ps, std::vector<CsrHost>& Rs, std::vector<double>& rhos,
                   std::vector<double>& coarse_inv, bool invert) {
  As.clear();
  Ps.clear();
  Rs.clear();
  rhos.clear();
  As.push_back(A0);
  const double omega = 4.0 / 3.0;
  PhaseTimer tm;
  while (As.back().n_rows > coarse_size && (int)As.size() < max_levels) {
    const CsrHost& A = As.back();
    const int n = A.n_rows;
    // strength threshold: halved until the strength graph has at least 3 edges per row on average -- Galerkin
    // operators of 3D meshes spread their weight over many small entries and would otherwise coarsen 2x per level
    Graph S;
    double theta_l = theta;
    for (int attempt = 0; attempt < 4; ++attempt, theta_l *= 0.5) {
      strength_graph(A, theta_l, S);
      if ((double)S.idx.size() >= 3.0 * n) break;
    }
    tm.lap("strength", (int)As.size() - 1, n);
    std::vector<int32_t> agg;
    int nagg = mis2_aggregate(S, agg);
    if (As.size() == 1) nagg = drop_dirichlet_aggregates(A, agg, nagg);      // boundary rows exist on the finest level only
    tm.lap("mis2", (int)As.size() - 1, n);
    if (nagg >= 0.8 * n) break;
    // Gershgorin bounds, D^-1 and the smoothed prolongator (prolongator_bounds / prolongator_build below).  The filter is
    // on for the finest level (mesh edges with vanishing stiffness) and for operators denser than 32 entries per row (the
    // Galerkin levels of 3D meshes); on sparse coarse levels the unfiltered smoother gives the better prolongator
    // (2D: 40 instead of 49 GMRES iterations at N = 512).
    const bool filtered = As.size() == 1 || (double)A.nnz() > 32.0 * n;
    std::vector<double> dinv;
    double rho = 0.0, rhoF = 0.0;
    prolongator_bounds(A, n, S, filtered, dinv, rho, rhoF);
    CsrHost P;
    prolongator_build(A, n, S, agg, nagg, filtered, dinv, omega / rhoF, P);
    tm.lap("prolong", (int)As.size() - 1, n);
    CsrHost R, AP, Ac;
    transpose(P, R);
    tm.lap("transpose", (int)As.size() - 1, n);
    spgemm(A, P, AP);
    tm.lap("A*P", (int)As.size() - 1, n);
    spgemm(R, AP, Ac);
    tm.lap("R*(AP)", (int)As.size() - 1, n);
    rhos.push_back(rho);
    Ps.push_back(std::move(P));
    Rs.push_back(std::move(R));
    As.push_back(std::move(Ac));
  }
  // dense inverse of the coarsest operator
  const CsrHost& Ac = As.back();
  const int nc = Ac.n_rows;
  if ((int64_t)nc * nc > (int64_t)64 * 1000 * 1000) {
    set_error("AMG coarsening stalled at %d unknowns; coarsest level too large for a dense solve", nc);
    return KNP_E_UNSUPPORTED;
  }
  std::vector<double> M((size_t)nc * nc, 0.0);
  for (int i = 0; i < nc; ++i)
    for (int j = Ac.indptr[i]; j < Ac.indptr[i + 1]; ++j) M[(size_t)i * nc + Ac.indices[j]] += Ac.vals[j];
  if (!invert) {            // the caller inverts the dense coarsest operator on the device
    coarse_inv.swap(M);
    return KNP_OK;
  }
  return dense_inverse(nc, M, coarse_inv);
}

}  // namespace knp
