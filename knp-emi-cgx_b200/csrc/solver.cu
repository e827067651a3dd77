// GPU Krylov solver: left-preconditioned restarted GMRES with classical Gram-Schmidt (two passes),
// preconditioned-norm convergence test relative to ||B b||, nullspace removal after every preconditioner
// application -- the semantics of the reference's PETSc KSP configuration
// (KNPEMIx_solver.py:212-214,276-280,324-333,386-389,435; PETSc defaults restated in SURVEY.md Appendix F).
// The preconditioner B is one V(1,1) cycle of our smoothed-aggregation hierarchy on P (amg_setup.cpp),
// Jacobi on P, or the identity.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "context.cuh"

namespace knp {

int ensure_workspace(knp_ctx* c, int restart) {
  const size_t ncols = (size_t)c->T.L.n_cols;
  if (c->ws_restart >= restart && c->V.p) return KNP_OK;
  c->ldv = (ncols + 31) / 32 * 32;
  KNP_TRY(c->V.alloc(c->ldv * (restart + 1)));
  KNP_CUDA(cudaMemset(c->V.p, 0, c->ldv * (restart + 1) * sizeof(double)));
  KNP_TRY(c->w.alloc(c->ldv));
  KNP_TRY(c->tmp.alloc(c->ldv));
  KNP_TRY(c->tmp2.alloc(c->ldv));
  KNP_TRY(c->colscale.alloc(c->ldv));
  KNP_CUDA(cudaMemset(c->tmp2.p, 0, c->ldv * sizeof(double)));
  KNP_CUDA(cudaMemset(c->w.p, 0, c->ldv * sizeof(double)));
  KNP_CUDA(cudaMemset(c->tmp.p, 0, c->ldv * sizeof(double)));
  KNP_TRY(c->partial.alloc((size_t)(restart + 2) * RED_BLOCKS));
  KNP_TRY(c->hdev.alloc(restart + 2));
  KNP_TRY(c->ydev.alloc(restart + 2));
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  KNP_CUDA(cudaMallocHost(&c->h_pinned, (restart + 2) * sizeof(double)));
  c->ws_restart = restart;
  return KNP_OK;
}

static int upload_csr(const CsrHost& h, CsrDev& d) {
  d.n_rows = h.n_rows;
  d.n_cols = h.n_cols;
  d.nnz = h.nnz();
  {
    std::vector<int32_t> ip(h.indptr);
    for (int k = 0; k < 4; ++k) ip.push_back(h.indptr.back());   // padding for the 16-byte TMA slices
    KNP_TRY(d.indptr.upload(ip));
  }
  KNP_TRY(d.indices.upload(h.indices));
  KNP_TRY(d.vals.upload(h.vals));
  std::vector<int32_t> blk;
  d.nblk = build_rowblocks(h.indptr.data(), h.n_rows, blk);
  if (d.nblk > 0) KNP_TRY(d.rowblk.upload(blk));
  else d.nblk = 0;
  return KNP_OK;
}

// ---- two-level additive Schwarz coarse space for multi-GPU runs -------------------------------------------------
// The processor-local AMG drops the couplings to ghost columns, which destroys the near-null constants of the phi
// blocks (K_w - (C_M/F) M_Gamma): the membrane-capacitor modes would then be left to GMRES alone (thousands of
// iterations).  One constant per (rank, field block) restores them: z += Z (Z^T P Z)^-1 Z^T r, with Z^T P Z formed
// from the *global* P (ghost couplings included) and inverted redundantly on every rank (8 nranks x nranks systems).
constexpr int CZ_BLOCKS = 64;

__global__ void __launch_bounds__(256) cz_partial_kernel(Layout L, const double* __restrict__ r, double* __restrict__ partial) {
  __shared__ double red[8];
  const int fb = blockIdx.y;                    // field block 0..7
  const int s = fb >> 2, f = fb & 3;
  const int lo = L.row(s, f, 0), n = L.n_own[s];
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int a = blockIdx.x * per, b = min(n, a + per);
  double acc = 0.0;
  for (int i = a + threadIdx.x; i < b; i += 256) acc += r[lo + i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if (lane == 0) red[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[fb * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void cz_final_kernel(int nranks, int rank, int nb, const double* __restrict__ partial, double* __restrict__ sums) {
  // sums[r*8 + fb]: zero except this rank's 8 entries
  const int t = threadIdx.x;
  if (t < 8 * nranks) {
    double v = 0.0;
    if (t / 8 == rank) {
      const int fb = t % 8;
      for (int i = 0; i < nb; ++i) v += partial[fb * nb + i];
    }
    sums[t] = v;
  }
}
__global__ void cz_add_kernel(Layout L, int nranks, const double* __restrict__ sums, const double* __restrict__ einv,
                              double* __restrict__ z) {
  __shared__ double y[8];
  if (threadIdx.x < 8) {
    double t = 0.0;
    for (int r = 0; r < nranks; ++r) t += einv[threadIdx.x * nranks + r] * sums[r * 8 + threadIdx.x];
    y[threadIdx.x] = t;
  }
  __syncthreads();
  const int n0 = 4 * L.n_own[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n_rows; i += gridDim.x * blockDim.x) {
    const int fb = i < n0 ? i / max(L.n_own[0], 1) : 4 + (i - n0) / max(L.n_own[1], 1);
    z[i] += y[fb];
  }
}

static int coarse_setup(knp_ctx* c, const std::vector<int32_t>& idx, const std::vector<double>& val) {
  c->cz_on = false;
  if (c->nranks <= 1) return KNP_OK;
  const Layout& L = c->T.L;
  const int n = L.n_rows, nr = c->nranks, np = (int)c->peers.size();
  std::vector<int32_t> ghost_owner((size_t)(L.n_cols - n), -1);
  for (int i = 0; i < np; ++i)
    for (int64_t k = c->recv_ptr[i]; k < c->recv_ptr[i + 1]; ++k) ghost_owner[c->h_recv_cols[k] - n] = c->peers[i];
  std::vector<double> E((size_t)8 * nr * nr, 0.0);      // E[(r', fb), r] flattened as ((r'*8 + fb) * nr + r)
  for (int s = 0; s < 2; ++s)
    for (int f = 0; f < 4; ++f)
      for (int p = 0; p < L.n_own[s]; ++p) {
        const int row = L.row(s, f, p);
        double* e = &E[((size_t)c->rank * 8 + 4 * s + f) * nr];
        for (int j = c->H.indptr_P[row]; j < c->H.indptr_P[row + 1]; ++j) {
          const int col = idx[j];
          const int rj = col < n ? c->rank : ghost_owner[col - n];
          if (rj < 0) {
            set_error("coarse space: ghost column %d has no owner in the halo lists", col);
            return KNP_E_INVALID;
          }
          e[rj] += val[j];
        }
      }
  DevBuf<double> dE;
  KNP_TRY(dE.upload(E));
  KNP_TRY(allreduce_sum(c, dE.p, (int)E.size(), c->stream));
  KNP_CUDA(cudaMemcpyAsync(E.data(), dE.p, E.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  // per field block: invert the nr x nr matrix, keep this rank's row
  std::vector<double> einv((size_t)8 * nr, 0.0);
  for (int fb = 0; fb < 8; ++fb) {
    std::vector<double> M((size_t)nr * nr), inv((size_t)nr * nr, 0.0);
    for (int a = 0; a < nr; ++a)
      for (int b = 0; b < nr; ++b) M[(size_t)a * nr + b] = E[((size_t)a * 8 + fb) * nr + b];
    for (int a = 0; a < nr; ++a) {
      bool empty = true;
      for (int b = 0; b < nr; ++b) empty = empty && M[(size_t)a * nr + b] == 0.0 && M[(size_t)b * nr + a] == 0.0;
      if (empty) M[(size_t)a * nr + a] = 1.0;      // rank without dofs of this field
      inv[(size_t)a * nr + a] = 1.0;
    }
    for (int k = 0; k < nr; ++k) {                 // Gauss-Jordan with partial pivoting
      int piv = k;
      for (int a = k + 1; a < nr; ++a)
        if (std::fabs(M[(size_t)a * nr + k]) > std::fabs(M[(size_t)piv * nr + k])) piv = a;
      if (M[(size_t)piv * nr + k] == 0.0) {
        set_error("coarse space matrix of field block %d is singular", fb);
        return KNP_E_INVALID;
      }
      for (int b = 0; b < nr; ++b) {
        std::swap(M[(size_t)k * nr + b], M[(size_t)piv * nr + b]);
        std::swap(inv[(size_t)k * nr + b], inv[(size_t)piv * nr + b]);
      }
      const double d = 1.0 / M[(size_t)k * nr + k];
      for (int b = 0; b < nr; ++b) {
        M[(size_t)k * nr + b] *= d;
        inv[(size_t)k * nr + b] *= d;
      }
      for (int a = 0; a < nr; ++a) {
        if (a == k) continue;
        const double fct = M[(size_t)a * nr + k];
        for (int b = 0; b < nr; ++b) {
          M[(size_t)a * nr + b] -= fct * M[(size_t)k * nr + b];
          inv[(size_t)a * nr + b] -= fct * inv[(size_t)k * nr + b];
        }
      }
    }
    for (int b = 0; b < nr; ++b) einv[(size_t)fb * nr + b] = inv[(size_t)c->rank * nr + b];
  }
  KNP_TRY(c->cz_einv.upload(einv));
  KNP_TRY(c->cz_sums.alloc((size_t)8 * nr));
  KNP_TRY(c->cz_partial.alloc((size_t)8 * CZ_BLOCKS));
  if (8 * nr > 1024) {
    set_error("coarse space supports at most 128 ranks");
    return KNP_E_UNSUPPORTED;
  }
  c->cz_on = true;
  return KNP_OK;
}

static int coarse_apply(knp_ctx* c, const double* r, double* z, cudaStream_t st) {
  if (!c->cz_on) return KNP_OK;
  const Layout& L = c->T.L;
  cz_partial_kernel<<<dim3(CZ_BLOCKS, 8), 256, 0, st>>>(L, r, c->cz_partial.p);
  KNP_LAUNCHED();
  cz_final_kernel<<<1, 1024, 0, st>>>(c->nranks, c->rank, CZ_BLOCKS, c->cz_partial.p, c->cz_sums.p);
  KNP_LAUNCHED();
  KNP_TRY(allreduce_sum(c, c->cz_sums.p, 8 * c->nranks, st));
  int grid = (L.n_rows + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  cz_add_kernel<<<grid, 256, 0, st>>>(L, c->nranks, c->cz_sums.p, c->cz_einv.p, z);
  KNP_LAUNCHED();
  return KNP_OK;
}

static int vcycle(Amg& M, int l, const double* bl, double* xout, cudaStream_t st);

// In-place Gauss-Jordan inversion of a dense SPD matrix on the device (no pivoting needed for SPD operators): the
// coarsest Galerkin operators of the Schur hierarchies have a few thousand unknowns, which removes the deepest,
// launch-bound levels from the cycle; the host inversion (with pivoting) stays for the indefinite blocks of P.
__global__ void gj_fetch_kernel(int n, int k, const double* __restrict__ A, double* __restrict__ fcol, double* __restrict__ prow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    fcol[i] = A[(size_t)i * n + k];
    prow[i] = A[(size_t)k * n + i];
  }
}
__global__ void gj_update_kernel(int n, int k, double* __restrict__ A, const double* __restrict__ fcol,
                                 const double* __restrict__ prow) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n) return;
  const double pr = (j == k ? 1.0 : prow[j]) / prow[k];
  const size_t at = (size_t)i * n + j;
  A[at] = i == k ? pr : (j == k ? 0.0 : A[at]) - fcol[i] * pr;
}
static int dense_inverse_device(int n, double* A, cudaStream_t st) {
  DevBuf<double> fcol, prow;
  KNP_TRY(fcol.alloc(n));
  KNP_TRY(prow.alloc(n));
  const dim3 grid((n + 255) / 256, n);
  for (int k = 0; k < n; ++k) {
    gj_fetch_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, k, A, fcol.p, prow.p);
    gj_update_kernel<<<grid, 256, 0, st>>>(n, k, A, fcol.p, prow.p);
  }
  KNP_CUDA(cudaGetLastError());
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

// builds the hierarchy of A0 on the host and uploads it; spd: coarsest operator (<= coarse_size unknowns) inverted on
// the device
static int build_amg(knp_ctx* c, const CsrHost& A0, std::unique_ptr<Amg>& out, int coarse_size = 600, bool spd = false) {
  std::vector<CsrHost> As, Ps, Rs;
  std::vector<double> rhos, cinv;
  KNP_TRY(amg_setup_host(A0, 0.08, coarse_size, 16, As, Ps, Rs, rhos, cinv, !spd));
  auto amg = std::make_unique<Amg>();
  const int nl = (int)Ps.size();
  for (int l = 0; l < nl; ++l) {
    auto* lv = new AmgLevelDev();
    amg->levels.push_back(lv);
    KNP_TRY(upload_csr(As[l], lv->A));
    KNP_TRY(upload_csr(Ps[l], lv->P));
    KNP_TRY(upload_csr(Rs[l], lv->R));
    lv->rho = rhos[l];
    const int nr = As[l].n_rows;
    KNP_TRY(lv->dinv.alloc(nr));
    KNP_TRY(lv->x.alloc(nr));
    KNP_TRY(lv->b.alloc(nr));
    KNP_TRY(lv->r.alloc(nr));
    KNP_TRY(launch_extract_dinv(nr, lv->A.indptr.p, lv->A.indices.p, lv->A.vals.p, lv->dinv.p, c->stream));
  }
  amg->n_coarse = As.back().n_rows;
  KNP_TRY(amg->coarse_inv.upload(cinv));
  if (spd) KNP_TRY(dense_inverse_device(amg->n_coarse, amg->coarse_inv.p, c->stream));
  KNP_TRY(amg->cb.alloc(amg->n_coarse));
  KNP_TRY(amg->cx.alloc(amg->n_coarse));
  amg->hostA = std::move(As);
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  out = std::move(amg);
  return KNP_OK;
}

// ---- charge-conservation Schur preconditioner (pc kind 3) ---------------------------------------------------------
// The potential row of `a` (KNPEMIx_problem.py:603-610) equals the z_k-weighted sum of the ion rows (:598-600) minus
// sum_k z_k M c_k (the membrane terms cancel too because sum_k alpha_k = 1), so with L = [I 0; -Z I] the system matrix
// becomes  L A = [A_cc A_cphi; -Z M 0].  Its Schur complement Z M A_cc^-1 A_cphi behaves like sum_k (z_k^2 c_k/psi) M at
// high and like the phi block K_phi + (C_M/F) M_Gamma at low frequencies, which gives
//      S~^-1 = (K_phi + (C_M/F) M_Gamma)^-1 + M_sigma^-1          (M_sigma lumped)
// and the block lower-triangular application
//      v = L r ;  z_c = AMG_c(v_c) ;  t = v_phi + M (sum_k z_k z_ck) ;  z_phi = AMG_phi(t) + t / M_sigma.
// Unlike the block-Jacobi form P of the reference (KNPEMIx_problem.py:657-744) no cancellation between the c and phi
// blocks has to be resolved by the inexact block solves: 20-30 GMRES iterations instead of 100-1500 on transient
// states (tests/experiments/pc_experiment.py).  oracle/amg.py::SchurPC restates it for the tests.
__global__ void schur_split_kernel(Layout L, double z0, double z1, double z2, const double* __restrict__ r,
                                   double* __restrict__ vc, double* __restrict__ t) {
  const int n0 = L.n_own[0], n1 = L.n_own[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += gridDim.x * blockDim.x) {
    const int s = i >= n0, p = s ? i - n0 : i, ns = s ? n1 : n0;
    const double* rs = r + L.rowbase[s];
    double* vs = vc + (s ? 3 * n0 : 0);
    const double a = rs[p], b = rs[ns + p], c = rs[2 * ns + p];
    vs[p] = a;
    vs[ns + p] = b;
    vs[2 * ns + p] = c;
    t[i] = rs[3 * ns + p] - ((z0 * a + z1 * b) + z2 * c);
  }
}
// q (full column layout, stored in the field-0 slots) = sum_k z_k z_ck
__global__ void schur_q_kernel(Layout L, double z0, double z1, double z2, const double* __restrict__ zc,
                               double* __restrict__ q) {
  const int n0 = L.n_own[0], n1 = L.n_own[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += gridDim.x * blockDim.x) {
    const int s = i >= n0, p = s ? i - n0 : i, ns = s ? n1 : n0;
    const double* zs = zc + (s ? 3 * n0 : 0);
    q[L.rowbase[s] + p] = (z0 * zs[p] + z1 * zs[ns + p]) + z2 * zs[2 * ns + p];
  }
}
// z (full layout) <- [z_c ; z_phi + t / M_sigma] ; rhs (optional, full layout) <- [v_c ; t]
__global__ void schur_merge_kernel(Layout L, const double* __restrict__ zc, const double* __restrict__ zp,
                                   const double* __restrict__ t, const double* __restrict__ msig_inv,
                                   const double* __restrict__ vc, double* __restrict__ z, double* __restrict__ rhs) {
  const int n0 = L.n_own[0], n1 = L.n_own[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += gridDim.x * blockDim.x) {
    const int s = i >= n0, p = s ? i - n0 : i, ns = s ? n1 : n0;
    const double* zs = zc + (s ? 3 * n0 : 0);
    double* out = z + L.rowbase[s];
    out[p] = zs[p];
    out[ns + p] = zs[ns + p];
    out[2 * ns + p] = zs[2 * ns + p];
    out[3 * ns + p] = zp[i] + t[i] * msig_inv[i];
    if (rhs) {
      const double* vs = vc + (s ? 3 * n0 : 0);
      double* ro = rhs + L.rowbase[s];
      ro[p] = vs[p];
      ro[ns + p] = vs[ns + p];
      ro[2 * ns + p] = vs[2 * ns + p];
      ro[3 * ns + p] = t[i];
    }
  }
}

// ---- multi-GPU: field-parallel hierarchies ---------------------------------------------------------------------
// Processor-local hierarchies (block Jacobi over the ranks) ruin the Schur form: the truncated ion solves are wrong
// near every rank boundary and GMRES needs 10-100x the iterations (2100 on 8 GPUs).  The eight diagonal blocks of the
// preconditioner are independent problems, so instead of cutting every block into nranks pieces, every block gets ONE
// global smoothed-aggregation hierarchy on ONE rank (each stage balanced by field size): the preconditioner is then the same operator as on a single GPU, independent of the partition.  Per
// application every rank ships its piece of each right-hand side to the field's owner and gets its piece of the result
// back (grouped ncclSend/ncclRecv over NVLink, 2 x 8 B per dof).
// Field -> rank: the two stages (ion fields, then potential fields) run one after the other, so each stage is balanced
// on its own: longest-processing-time-first over the global field sizes (ECS fields are ~3x the ICS ones).
void assign_field_owners(int nranks, const int64_t size_s[2], int owner[8]) {
  std::vector<int64_t> load(nranks, 0);
  auto least = [&](int exclude) {
    int best = -1;
    for (int r = 0; r < nranks; ++r)
      if (r != exclude && (best < 0 || load[r] < load[best])) best = r;
    return best < 0 ? 0 : best;
  };
  const int big = size_s[1] >= size_s[0] ? 1 : 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int s = pass == 0 ? big : 1 - big;
    for (int f = 0; f < 3; ++f) {
      const int r = least(-1);
      owner[4 * s + f] = r;
      load[r] += size_s[s];
    }
  }
  // potentials: the larger one on the rank with the least ion work, the other one on a different rank
  const int r_big = least(-1);
  owner[4 * big + 3] = r_big;
  owner[4 * (1 - big) + 3] = nranks > 1 ? least(r_big) : r_big;
}

static int schur_setup_fieldpar(knp_ctx* c, const std::vector<int32_t>& idx, const std::vector<double>& val) {
  const Layout& L = c->T.L;
  const int R = c->nranks, me = c->rank;
  cudaStream_t st = c->stream;
  knp_ctx::FieldPar& F = c->fp;
  const std::vector<int32_t>& ip = c->H.indptr_P;
  // owned node counts of every rank and nnz of every (rank, field) piece
  std::vector<double> tab((size_t)R * 10, 0.0);
  for (int s = 0; s < 2; ++s) tab[(size_t)me * 10 + s] = L.n_own[s];
  for (int s = 0; s < 2; ++s)
    for (int f = 0; f < 4; ++f)
      tab[(size_t)me * 10 + 2 + 4 * s + f] = L.n_own[s] ? (double)(ip[L.row(s, f, 0) + L.n_own[s]] - ip[L.row(s, f, 0)]) : 0.0;
  {
    DevBuf<double> d;
    KNP_TRY(d.upload(tab));
    KNP_TRY(allreduce_sum(c, d.p, (int)tab.size(), st));
    KNP_CUDA(cudaMemcpyAsync(tab.data(), d.p, tab.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    KNP_CUDA(cudaStreamSynchronize(st));
  }
  auto nown = [&](int r, int s) { return (int64_t)tab[(size_t)r * 10 + s]; };
  auto pnnz = [&](int r, int fld) { return (int64_t)tab[(size_t)r * 10 + 2 + fld]; };
  for (int s = 0; s < 2; ++s) {
    F.off[s].assign(R + 1, 0);
    for (int r = 0; r < R; ++r) F.off[s][r + 1] = F.off[s][r] + nown(r, s);
    KNP_CHECK(F.off[s][R] < ((int64_t)1 << 31), "global field too large for int32 columns");
  }
  // global node ids of the ghost columns: one halo exchange of an id vector (field-0 slots)
  std::vector<double> ids(L.n_cols, 0.0);
  for (int s = 0; s < 2; ++s)
    for (int p = 0; p < L.n_own[s]; ++p) ids[L.col(s, 0, p)] = (double)(F.off[s][me] + p);
  {
    DevBuf<double> d;
    KNP_TRY(d.upload(ids));
    KNP_TRY(halo_exchange(c, d.p, st));
    KNP_CUDA(cudaMemcpyAsync(ids.data(), d.p, ids.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    KNP_CUDA(cudaStreamSynchronize(st));
  }
  // my piece of every field: row lengths, GLOBAL columns, values
  int64_t c_size = 0, p_size = 0;
  {
    const int64_t size_s[2] = {F.off[0][R], F.off[1][R]};
    assign_field_owners(R, size_s, F.owner);
  }
  for (int s = 0; s < 2; ++s)
    for (int f = 0; f < 4; ++f) {
      const int fld = 4 * s + f;
      if (F.owner[fld] == me) {
        int64_t& acc = f < 3 ? c_size : p_size;
        F.base[fld] = acc;
        acc += F.off[s][R];
      }
    }
  struct Piece {
    std::vector<int32_t> len, col;
    std::vector<double> val;
  };
  std::vector<Piece> mine(8);
  for (int s = 0; s < 2; ++s)
    for (int f = 0; f < 4; ++f) {
      Piece& P = mine[4 * s + f];
      const int n_s = L.n_own[s];
      P.len.resize(n_s);
      for (int p = 0; p < n_s; ++p) {
        const int row = L.row(s, f, p);
        P.len[p] = ip[row + 1] - ip[row];
        for (int j = ip[row]; j < ip[row + 1]; ++j) {
          // column (s, f, q) -> global node id of q through the field-0 slot of the id vector
          const int cfull = idx[j];
          const int q = cfull < L.n_rows ? cfull - L.row(s, f, 0) : (cfull - L.n_rows - L.gbase[s] - f * L.n_gh[s]) + n_s;
          P.col.push_back((int32_t)ids[L.col(s, 0, q)]);
          P.val.push_back(val[j]);
        }
      }
    }
  // ship the pieces to the owners (device staging; one grouped exchange)
  std::vector<DevBuf<int32_t>> d_len(8), d_col(8);
  std::vector<DevBuf<double>> d_val(8);
  std::vector<std::vector<DevBuf<int32_t>>> r_len(8), r_col(8);
  std::vector<std::vector<DevBuf<double>>> r_val(8);
  std::vector<P2POp> ops;
  for (int fld = 0; fld < 8; ++fld) {
    const int s = fld >> 2, o = F.owner[fld];
    if (o != me) {
      KNP_TRY(d_len[fld].upload(mine[fld].len));
      KNP_TRY(d_col[fld].upload(mine[fld].col));
      KNP_TRY(d_val[fld].upload(mine[fld].val));
      ops.push_back({o, d_len[fld].p, mine[fld].len.size() * 4, true});
      ops.push_back({o, d_col[fld].p, mine[fld].col.size() * 4, true});
      ops.push_back({o, d_val[fld].p, mine[fld].val.size() * 8, true});
    } else {
      r_len[fld] = std::vector<DevBuf<int32_t>>(R);
      r_col[fld] = std::vector<DevBuf<int32_t>>(R);
      r_val[fld] = std::vector<DevBuf<double>>(R);
      for (int r = 0; r < R; ++r) {
        if (r == me) continue;
        KNP_TRY(r_len[fld][r].alloc((size_t)nown(r, s)));
        KNP_TRY(r_col[fld][r].alloc((size_t)pnnz(r, fld)));
        KNP_TRY(r_val[fld][r].alloc((size_t)pnnz(r, fld)));
        ops.push_back({r, r_len[fld][r].p, (size_t)nown(r, s) * 4, false});
        ops.push_back({r, r_col[fld][r].p, (size_t)pnnz(r, fld) * 4, false});
        ops.push_back({r, r_val[fld][r].p, (size_t)pnnz(r, fld) * 8, false});
      }
    }
  }
  KNP_TRY(p2p_exchange(c, ops, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  // merged global matrices of the fields this rank owns (block diagonal over the fields)
  CsrHost Gc, Gp;
  Gc.n_rows = Gc.n_cols = (int)c_size;
  Gp.n_rows = Gp.n_cols = (int)p_size;
  Gc.indptr.assign(1, 0);
  Gp.indptr.assign(1, 0);
  for (int pass = 0; pass < 2; ++pass)            // ion fields first, then the potential fields, each in base order
    for (int fld = 0; fld < 8; ++fld) {
      if (F.owner[fld] != me || ((fld & 3) < 3) != (pass == 0)) continue;
      CsrHost& G = pass == 0 ? Gc : Gp;
      const int s = fld >> 2;
      KNP_CHECK((int64_t)G.indptr.size() - 1 == F.base[fld], "field-parallel layout mismatch");
      for (int r = 0; r < R; ++r) {
        std::vector<int32_t> len, col;
        std::vector<double> v;
        if (r == me) {
          len.swap(mine[fld].len);
          col.swap(mine[fld].col);
          v.swap(mine[fld].val);
        } else {
          len.resize((size_t)nown(r, s));
          col.resize((size_t)pnnz(r, fld));
          v.resize((size_t)pnnz(r, fld));
          if (!len.empty()) KNP_CUDA(cudaMemcpy(len.data(), r_len[fld][r].p, len.size() * 4, cudaMemcpyDeviceToHost));
          if (!col.empty()) KNP_CUDA(cudaMemcpy(col.data(), r_col[fld][r].p, col.size() * 4, cudaMemcpyDeviceToHost));
          if (!v.empty()) KNP_CUDA(cudaMemcpy(v.data(), r_val[fld][r].p, v.size() * 8, cudaMemcpyDeviceToHost));
          r_len[fld][r].free();
          r_col[fld][r].free();
          r_val[fld][r].free();
        }
        size_t at = 0;
        for (size_t p = 0; p < len.size(); ++p) {
          // columns of one row sorted by global id (ghost columns interleave with owned ones)
          std::vector<std::pair<int32_t, double>> row(len[p]);
          for (int j = 0; j < len[p]; ++j, ++at) row[j] = {(int32_t)(col[at] + F.base[fld]), v[at]};
          std::sort(row.begin(), row.end());
          for (auto& e : row) {
            G.indices.push_back(e.first);
            G.vals.push_back(e.second);
          }
          G.indptr.push_back((int32_t)G.indices.size());
        }
      }
    }
  if (c_size > 0) KNP_TRY(build_amg(c, Gc, c->amg_c, 2500, true));
  if (p_size > 0) KNP_TRY(build_amg(c, Gp, c->amg_p, 2500, true));
  KNP_TRY(F.gc_in.alloc((size_t)c_size));
  KNP_TRY(F.gc_out.alloc((size_t)c_size));
  KNP_TRY(F.gp_in.alloc((size_t)p_size));
  KNP_TRY(F.gp_out.alloc((size_t)p_size));
  F.on = true;
  return KNP_OK;
}

// moves the ranks' pieces of the given fields to the owners' merged vectors (to_owner) or the results back
static int fieldpar_move(knp_ctx* c, bool ions, bool to_owner, double* local, double* merged, cudaStream_t st) {
  const Layout& L = c->T.L;
  knp_ctx::FieldPar& F = c->fp;
  const int R = c->nranks, me = c->rank, n0 = L.n_own[0];
  std::vector<P2POp> ops;
  for (int fld = 0; fld < 8; ++fld) {
    const int s = fld >> 2, f = fld & 3;
    if ((f < 3) != ions) continue;
    // my piece inside the compact local vector: ions [s=0: 3 n0 | s=1: 3 n1] field-major, potentials [n0 | n1]
    double* piece = ions ? local + (s ? 3 * n0 : 0) + (size_t)f * L.n_own[s] : local + (s ? n0 : 0);
    const size_t mine = (size_t)L.n_own[s] * 8;
    const int o = F.owner[fld];
    if (o != me) {
      ops.push_back({o, piece, mine, to_owner});
    } else {
      for (int r = 0; r < R; ++r) {
        double* at = merged + F.base[fld] + F.off[s][r];
        const size_t bytes = (size_t)(F.off[s][r + 1] - F.off[s][r]) * 8;
        if (r == me) {
          if (bytes) KNP_CUDA(cudaMemcpyAsync(to_owner ? at : piece, to_owner ? piece : at, bytes, cudaMemcpyDeviceToDevice, st));
        } else {
          ops.push_back({r, at, bytes, !to_owner});
        }
      }
    }
  }
  return p2p_exchange(c, ops, st);
}

static int schur_setup(knp_ctx* c) {
  const Layout& L = c->T.L;
  const int n = L.n_rows, n0 = L.n_own[0], n1 = L.n_own[1];
  KNP_CHECK(L.rowbase[0] == 0 && L.rowbase[1] == 4 * n0, "unexpected row layout");
  cudaStream_t st = c->stream;
  // P~: ion blocks M + dt D_k K, phi blocks K_phi + (C_M/F) M_Gamma (the sign the membrane term has in `a`)
  KParams kp = c->kp;
  kp.C_M = -c->kp.C_M;
  KNP_TRY(launch_rows(c->T, kp, 1, c->u.p, c->fe.p, c->P_vals.p, nullptr, c->H.max_deg, c->H.max_gdeg, st));
  c->P_assembled = true;
  // mass matrices: the same kernel with D = 0 leaves M in the ion blocks
  kp = c->kp;
  for (int k = 0; k < 3; ++k) kp.D[k] = 0.0;
  KNP_TRY(c->M_vals.alloc(c->H.nnz_P));
  KNP_TRY(launch_rows(c->T, kp, 1, c->u.p, c->fe.p, c->M_vals.p, nullptr, c->H.max_deg, c->H.max_gdeg, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  std::vector<int32_t> idx(c->H.nnz_P);
  std::vector<double> val(c->H.nnz_P), mval(c->H.nnz_P), u(L.n_cols);
  KNP_CUDA(cudaMemcpy(idx.data(), c->d_indices_P.p, idx.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  KNP_CUDA(cudaMemcpy(val.data(), c->P_vals.p, val.size() * sizeof(double), cudaMemcpyDeviceToHost));
  KNP_CUDA(cudaMemcpy(mval.data(), c->M_vals.p, mval.size() * sizeof(double), cudaMemcpyDeviceToHost));
  KNP_CUDA(cudaMemcpy(u.data(), c->u.p, u.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const std::vector<int32_t>& ip = c->H.indptr_P;
  // compact numbering: c part [s=0: 3 n0 | s=1: 3 n1], phi part [n0 | n1]; ghost columns are dropped (processor-local)
  auto cmap = [&](int i) -> int {   // full row/col -> compact index in its part, or -1 if it belongs to the other part
    if (i < 3 * n0) return i;
    if (i < 4 * n0) return -1;
    if (i < 4 * n0 + 3 * n1) return i - n0;
    return -1;
  };
  auto pmap = [&](int i) -> int {
    if (i < 3 * n0) return -1;
    if (i < 4 * n0) return i - 3 * n0;
    if (i < 4 * n0 + 3 * n1) return -1;
    return i - 3 * n0 - 3 * n1;
  };
  CsrHost Acc, App;
  Acc.n_rows = Acc.n_cols = 3 * (n0 + n1);
  App.n_rows = App.n_cols = n0 + n1;
  Acc.indptr.assign(1, 0);
  App.indptr.assign(1, 0);
  for (int i = 0; i < n && c->nranks == 1; ++i) {     // multi-GPU runs build field-parallel global hierarchies instead
    const bool isc = cmap(i) >= 0;
    CsrHost& M = isc ? Acc : App;
    for (int j = ip[i]; j < ip[i + 1]; ++j) {
      if (idx[j] >= n) continue;
      const int cc = isc ? cmap(idx[j]) : pmap(idx[j]);
      if (cc < 0) continue;
      M.indices.push_back(cc);
      M.vals.push_back(val[j]);
    }
    M.indptr.push_back((int32_t)M.indices.size());
  }
  // rows were visited in the order c(s=0), phi(s=0), c(s=1), phi(s=1) = ascending compact order in both parts
  c->fp.on = false;
  if (c->nranks > 1) {
    KNP_TRY(schur_setup_fieldpar(c, idx, val));
  } else {
    KNP_TRY(build_amg(c, Acc, c->amg_c, 2500, true));
    KNP_TRY(build_amg(c, App, c->amg_p, 2500, true));
  }
  // W-cycle on levels 1..3, V-cycle below: measured optimum on C3 (36 -> 15 iterations; deeper W recursion only adds
  // launch-bound visits of tiny levels)
  for (Amg* a : {c->amg_c.get(), c->amg_p.get()})
    if (a) {
      a->gamma = 2;
      a->gamma_last = 3;
      if (const char* e = getenv("KNP_W_LEVELS")) a->gamma_last = atoi(e);
    }
  // lumped M_sigma = (sum_k z_k^2 c_k / psi) at the node  x  row sum of the mass matrix
  std::vector<double> msig_inv((size_t)n0 + n1);
  const double* z = c->kp.z;
  for (int s = 0; s < 2; ++s)
    for (int p = 0; p < L.n_own[s]; ++p) {
      const int row = L.row(s, 0, p);
      double ms = 0.0;
      for (int j = ip[row]; j < ip[row + 1]; ++j) ms += mval[j];
      double sig = 0.0;
      for (int k = 0; k < 3; ++k) sig += z[k] * z[k] / c->kp.psi * u[L.col(s, k, p)];
      msig_inv[(size_t)(s ? n0 : 0) + p] = 1.0 / (sig * ms);
    }
  KNP_TRY(c->msig_inv.upload(msig_inv));
  // row blocks of the two mass-matrix row ranges (rows (s, 0, .) of the P pattern): TMA-staged SpMV on sub-ranges
  for (int s = 0; s < 2; ++s) {
    c->sch_nmblk[s] = 0;
    if (L.n_own[s] == 0) continue;
    std::vector<int32_t> blk;
    const int nb = build_rowblocks(ip.data() + L.row(s, 0, 0), L.n_own[s], blk);
    if (nb > 0 && L.row(s, 0, 0) % 4 == 0) {
      KNP_TRY(c->sch_mblk[s].upload(blk));
      c->sch_nmblk[s] = nb;
    }
  }
  KNP_TRY(c->sch_vc.alloc((size_t)3 * (n0 + n1)));
  KNP_TRY(c->sch_zc.alloc((size_t)3 * (n0 + n1)));
  KNP_TRY(c->sch_t.alloc((size_t)n0 + n1));
  KNP_TRY(c->sch_zp.alloc((size_t)n0 + n1));
  KNP_TRY(c->sch_q.alloc(L.n_cols));
  KNP_TRY(c->sch_rhs.alloc(L.n_rows));
  KNP_CUDA(cudaMemset(c->sch_q.p, 0, (size_t)L.n_cols * sizeof(double)));
  return KNP_OK;
}

// one cycle of a field owner's hierarchy, replayed from a CUDA graph after the first two calls (multi-GPU path)
static int vcycle_graphed(knp_ctx* c, Amg& M, const double* in, double* out, cudaStream_t st) {
  static const bool enabled = !(getenv("KNP_PC_GRAPH") && atoi(getenv("KNP_PC_GRAPH")) == 0);
  if (!enabled) return vcycle(M, 0, in, out, st);
  for (auto& g : c->cycle_graphs)
    if (g.amg == &M && g.in == in && g.out == out) {
      KNP_CUDA(cudaGraphLaunch(g.exec, st));
      g_kernel_launches += g.launches;
      return KNP_OK;
    }
  if (c->cycle_calls++ < 2 || c->cycle_graphs.size() >= 8) return vcycle(M, 0, in, out, st);   // warm-up first
  const unsigned long long l0 = g_kernel_launches;
  KNP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = vcycle(M, 0, in, out, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (rc != KNP_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == KNP_OK) set_error("CUDA graph capture of a cycle failed: %s", cudaGetErrorString(e));
    return rc != KNP_OK ? rc : KNP_E_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  KNP_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  c->cycle_graphs.push_back({&M, in, out, exec, g_kernel_launches - l0});
  KNP_CUDA(cudaGraphLaunch(exec, st));
  return KNP_OK;
}

static int schur_apply(knp_ctx* c, const double* r, double* zout, cudaStream_t st) {
  const Layout& L = c->T.L;
  const int n0 = L.n_own[0], n1 = L.n_own[1];
  const double* z = c->kp.z;
  int grid = (n0 + n1 + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  schur_split_kernel<<<grid, 256, 0, st>>>(L, z[0], z[1], z[2], r, c->sch_vc.p, c->sch_t.p);
  KNP_LAUNCHED();
  if (c->fp.on) {
    KNP_TRY(fieldpar_move(c, true, true, c->sch_vc.p, c->fp.gc_in.p, st));
    if (c->amg_c) KNP_TRY(vcycle_graphed(c, *c->amg_c, c->fp.gc_in.p, c->fp.gc_out.p, st));
    KNP_TRY(fieldpar_move(c, true, false, c->sch_zc.p, c->fp.gc_out.p, st));
  } else {
    KNP_TRY(vcycle(*c->amg_c, 0, c->sch_vc.p, c->sch_zc.p, st));
  }
  schur_q_kernel<<<grid, 256, 0, st>>>(L, z[0], z[1], z[2], c->sch_zc.p, c->sch_q.p);
  KNP_LAUNCHED();
  KNP_TRY(halo_exchange(c, c->sch_q.p, st));
  for (int s = 0; s < 2; ++s) {
    if (L.n_own[s] == 0) continue;
    const int row0 = L.row(s, 0, 0);
    const int64_t nnz_s = (int64_t)c->H.indptr_P[row0 + L.n_own[s]] - c->H.indptr_P[row0];
    double* tout = c->sch_t.p + (s ? n0 : 0);
    if (c->sch_nmblk[s] > 0 && ((uintptr_t)tout & 7u) == 0)
      KNP_TRY(launch_spmv_stream(c->sch_nmblk[s], c->sch_mblk[s].p, c->d_indptr_P.p + row0, c->d_indices_P.p, c->M_vals.p,
                                 c->sch_q.p, tout, EPI_ADD, nullptr, nullptr, 0.0, st, (double)nnz_s / L.n_own[s]));
    else
      KNP_TRY(launch_spmv(L.n_own[s], nnz_s, c->d_indptr_P.p + row0, c->d_indices_P.p, c->M_vals.p, c->sch_q.p, tout,
                          EPI_ADD, nullptr, nullptr, 0.0, st));
  }
  if (c->fp.on) {
    KNP_TRY(fieldpar_move(c, false, true, c->sch_t.p, c->fp.gp_in.p, st));
    if (c->amg_p) KNP_TRY(vcycle_graphed(c, *c->amg_p, c->fp.gp_in.p, c->fp.gp_out.p, st));
    KNP_TRY(fieldpar_move(c, false, false, c->sch_zp.p, c->fp.gp_out.p, st));
  } else {
    KNP_TRY(vcycle(*c->amg_p, 0, c->sch_t.p, c->sch_zp.p, st));
  }
  schur_merge_kernel<<<grid, 256, 0, st>>>(L, c->sch_zc.p, c->sch_zp.p, c->sch_t.p, c->msig_inv.p, c->sch_vc.p, zout,
                                           c->cz_on ? c->sch_rhs.p : nullptr);
  KNP_LAUNCHED();
  if (c->cz_on) KNP_TRY(coarse_apply(c, c->sch_rhs.p, zout, st));
  return KNP_OK;
}

void pc_graphs_clear(knp_ctx* c);

int pc_setup(knp_ctx* c, const knp_solve_opts* o) {
  const int n = c->T.L.n_rows;
  pc_graphs_clear(c);
  c->amg.reset();
  c->amg_c.reset();
  c->amg_p.reset();
  c->cz_on = false;
  c->pc_kind = o->pc;
  if (o->pc == 0) return KNP_OK;
  if (o->pc == 3) {
    KNP_CHECK(c->params_set, "knp_set_params must be called first");
    return schur_setup(c);
  }
  if (!c->P_assembled) {
    set_error("knp_pc_setup: assemble P first (knp_assemble_P)");
    return KNP_E_INVALID;
  }
  if (o->pc == 1) {
    KNP_TRY(c->pc_dinv.alloc(n));
    KNP_TRY(launch_extract_dinv(n, c->d_indptr_P.p, c->d_indices_P.p, c->P_vals.p, c->pc_dinv.p, c->stream));
    KNP_CUDA(cudaStreamSynchronize(c->stream));
    return KNP_OK;
  }
  if (o->pc != 2) {
    set_error("unknown preconditioner kind %d", o->pc);
    return KNP_E_INVALID;
  }
  // host copy of the owned-column part of P (processor-local block on multi-GPU runs)
  CsrHost P0;
  {
    std::vector<int32_t> idx(c->H.nnz_P);
    std::vector<double> val(c->H.nnz_P);
    KNP_CUDA(cudaStreamSynchronize(c->stream));
    KNP_CUDA(cudaMemcpy(idx.data(), c->d_indices_P.p, idx.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    KNP_CUDA(cudaMemcpy(val.data(), c->P_vals.p, val.size() * sizeof(double), cudaMemcpyDeviceToHost));
    P0.n_rows = n;
    P0.n_cols = n;
    P0.indptr.assign(n + 1, 0);
    P0.indices.reserve(idx.size());
    P0.vals.reserve(idx.size());
    for (int i = 0; i < n; ++i) {
      for (int j = c->H.indptr_P[i]; j < c->H.indptr_P[i + 1]; ++j)
        if (idx[j] < n) {
          P0.indices.push_back(idx[j]);
          P0.vals.push_back(val[j]);
        }
      P0.indptr[i + 1] = (int32_t)P0.indices.size();
    }
    KNP_TRY(coarse_setup(c, idx, val));
  }
  return build_amg(c, P0, c->amg);
}

// z = V-cycle(r); level-l right-hand side in bl, result written to xout (distinct from bl)
static CsrView view(const CsrDev& M) {
  return CsrView{M.n_rows, M.nnz, M.indptr.p, M.indices.p, M.vals.p, M.rowblk.p, M.nblk};
}

static int vcycle(Amg& M, int l, const double* bl, double* xout, cudaStream_t st) {
  const int nl = (int)M.levels.size();
  if (l == nl) return launch_dense_gemv(M.n_coarse, M.coarse_inv.p, bl, xout, st);
  AmgLevelDev& L = *M.levels[l];
  const int n = L.A.n_rows;
  const double w = (4.0 / 3.0) / L.rho;
  // pre-smooth from a zero initial guess
  KNP_TRY(launch_scale_dinv(n, w, L.dinv.p, bl, L.x.p, st));
  // coarse-grid correction; levels >= 1 repeat it `gamma` times (gamma = 2: W-cycle below the finest level, which
  // restores the two-level convergence rate of deep hierarchies at ~25 % extra cost because level 0 is visited once)
  const int reps = (l >= 1 && l <= M.gamma_last) ? M.gamma : 1;
  for (int rep = 0; rep < reps; ++rep) {
    // r = b - A x ; b_{l+1} = R r
    KNP_TRY(spmv(view(L.A), L.x.p, L.r.p, EPI_RESID, bl, nullptr, 0.0, st));
    // child right-hand side; the child's result goes into this level's r, which is free after the restriction
    double* bc = (l + 1 == nl) ? M.cb.p : M.levels[l + 1]->b.p;
    double* xc = L.r.p;
    KNP_TRY(spmv(view(L.R), L.r.p, bc, EPI_SET, nullptr, nullptr, 0.0, st));
    KNP_TRY(vcycle(M, l + 1, bc, xc, st));
    // x += P x_c
    KNP_TRY(spmv(view(L.P), xc, L.x.p, EPI_ADD, nullptr, nullptr, 0.0, st));
  }
  // post-smooth, out of place into xout
  KNP_TRY(spmv(view(L.A), L.x.p, xout, EPI_JACOBI, bl, L.dinv.p, w, st));
  return KNP_OK;
}

__global__ void dinv_mul_kernel(int n, const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ z) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) z[i] = dinv[i] * r[i];
}

void pc_graphs_clear(knp_ctx* c) {
  for (auto& g : c->pc_graphs) cudaGraphExecDestroy(g.exec);
  c->pc_graphs.clear();
  c->pc_applies = 0;
  for (auto& g : c->cycle_graphs) cudaGraphExecDestroy(g.exec);
  c->cycle_graphs.clear();
  c->cycle_calls = 0;
}

// The Schur application is ~200 small launches (W-cycle over two hierarchies): on one GPU it is captured once per
// (r, z) pointer pair into a CUDA graph and replayed (GMRES always applies it to the same two buffers).
static int schur_apply_graphed(knp_ctx* c, const double* r, double* z, cudaStream_t st) {
  static const bool enabled = !(getenv("KNP_PC_GRAPH") && atoi(getenv("KNP_PC_GRAPH")) == 0);
  if (!enabled || c->nranks > 1) return schur_apply(c, r, z, st);
  for (auto& g : c->pc_graphs)
    if (g.r == r && g.z == z) {
      KNP_CUDA(cudaGraphLaunch(g.exec, st));
      g_kernel_launches += g.launches;
      return KNP_OK;
    }
  if (c->pc_applies++ < 1 || c->pc_graphs.size() >= 8) return schur_apply(c, r, z, st);   // first call warms up attributes
  const unsigned long long l0 = g_kernel_launches;
  KNP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = schur_apply(c, r, z, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (rc != KNP_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == KNP_OK) set_error("CUDA graph capture of the preconditioner failed: %s", cudaGetErrorString(e));
    return rc != KNP_OK ? rc : KNP_E_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  KNP_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  c->pc_graphs.push_back({r, z, exec, g_kernel_launches - l0});
  KNP_CUDA(cudaGraphLaunch(exec, st));
  return KNP_OK;
}

int pc_apply(knp_ctx* c, const double* r, double* z, cudaStream_t st) {
  const int n = c->T.L.n_rows;
  if (c->pc_kind == 2 && c->amg) {
    KNP_TRY(vcycle(*c->amg, 0, r, z, st));
    return coarse_apply(c, r, z, st);
  }
  if (c->pc_kind == 3 && (c->fp.on || (c->amg_c && c->amg_p))) return schur_apply_graphed(c, r, z, st);
  if (c->pc_kind == 1) {
    int grid = (n + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    dinv_mul_kernel<<<grid, 256, 0, st>>>(n, c->pc_dinv.p, r, z);
    KNP_LAUNCHED();
    return KNP_OK;
  }
  KNP_CUDA(cudaMemcpyAsync(z, r, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return KNP_OK;
}

// x -= ns (ns . x): ns = normalised indicator of all phi_i and phi_e rows (KNPEMIx_solver.py:297-335)
static int project_nullspace(knp_ctx* c, double* x, cudaStream_t st) {
  const Layout& L = c->T.L;
  const int lo0 = L.row(0, 3, 0), hi0 = lo0 + L.n_own[0];
  const int lo1 = L.row(1, 3, 0), hi1 = lo1 + L.n_own[1];
  KNP_TRY(launch_range_sum(x, lo0, hi0, lo1, hi1, c->partial.p, c->hdev.p, st));
  KNP_TRY(allreduce_sum(c, c->hdev.p, 1, st));
  const double cnt = c->nranks > 1 ? (double)c->n_phi_global : (double)(L.n_own[0] + L.n_own[1]);
  return launch_range_shift(x, lo0, hi0, lo1, hi1, c->hdev.p, 1.0 / cnt, st);
}

int nullspace_remove(knp_ctx* c, double* x, cudaStream_t st) { return project_nullspace(c, x, st); }

// z = B v  (+ nullspace removal)
static int apply_B(knp_ctx* c, const knp_solve_opts* o, const double* v, double* z, cudaStream_t st) {
  KNP_TRY(pc_apply(c, v, z, st));
  if (o->project_nullspace) KNP_TRY(project_nullspace(c, z, st));
  return KNP_OK;
}

static int spmv_A(knp_ctx* c, const double* A_vals, double* x, double* y, int epi, const double* b, cudaStream_t st) {
  KNP_TRY(halo_exchange(c, x, st));
  const CsrView A{c->T.L.n_rows, c->H.nnz, c->d_indptr.p, c->d_indices.p, A_vals, c->d_rowblk_A.p, c->nblk_A};
  return spmv(A, x, y, epi, b, nullptr, 0.0, st);
}

// dots of w against V[0..m) plus ||w||^2 -> host (m+1 values)
static int dots_to_host(knp_ctx* c, int m, const double* w, double* host, cudaStream_t st) {
  const int n = c->T.L.n_rows;
  KNP_TRY(launch_multi_dot(n, m, c->V.p, c->ldv, w, c->partial.p, c->hdev.p, st));
  KNP_TRY(allreduce_sum(c, c->hdev.p, m + 1, st));
  KNP_CUDA(cudaMemcpyAsync(host, c->hdev.p, (m + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

int gmres_solve(knp_ctx* c, const double* A_vals, const double* b, double* x, const knp_solve_opts* o,
                knp_solve_info* info, cudaStream_t st) {
  const int n = c->T.L.n_rows;
  const int m = o->restart > 0 ? o->restart : 30;
  if (m > 62) {
    set_error("GMRES restart %d > 62 not supported", m);
    return KNP_E_INVALID;
  }
  KNP_TRY(ensure_workspace(c, m));
  double* hp = c->h_pinned;
  double* w = c->w.p;
  double* tmp = c->tmp.p;
  std::vector<double> H((size_t)(m + 1) * m, 0.0), g(m + 1, 0.0), cs(m, 0.0), sn(m, 0.0), y(m, 0.0);
  auto Hat = [&](int i, int j) -> double& { return H[(size_t)i * m + j]; };

  // optional per-field variable scaling x = D y (host fills D on the device once per solve)
  bool scaled = false;
  for (int f = 0; f < 8; ++f) scaled = scaled || o->field_scale[f] > 0.0;
  const double* D = nullptr;
  if (scaled) {
    std::vector<double> hs(c->ldv, 1.0);
    const Layout& L = c->T.L;
    for (int s = 0; s < 2; ++s)
      for (int f = 0; f < 4; ++f) {
        const double sc = o->field_scale[4 * s + f] > 0.0 ? o->field_scale[4 * s + f] : 1.0;
        for (int q = 0; q < L.n_loc[s]; ++q) hs[L.col(s, f, q)] = sc;
      }
    KNP_CUDA(cudaMemcpyAsync(c->colscale.p, hs.data(), c->ldv * sizeof(double), cudaMemcpyHostToDevice, st));
    KNP_CUDA(cudaStreamSynchronize(st));
    D = c->colscale.p;
  }
  // ||B b||  (in the scaled variables when D is set)
  KNP_TRY(apply_B(c, o, b, w, st));
  if (D) KNP_TRY(launch_pointwise(n, w, D, 1, w, st));
  KNP_TRY(dots_to_host(c, 0, w, hp, st));
  const double bnorm = std::sqrt(hp[0]);
  info->rnorm0 = bnorm;
  info->iterations = 0;
  info->converged = 0;
  info->rnorm = bnorm;
  if (!(bnorm == bnorm) || std::isinf(bnorm)) {
    set_error("GMRES: right-hand side is not finite");
    return KNP_E_NOCONV;
  }
  const double tol = o->rtol * bnorm;
  int its = 0;
  double prev_beta = -1.0;
  int stagn = 0;
  while (true) {
    // r = B (b - A x)
    KNP_TRY(spmv_A(c, A_vals, x, tmp, EPI_RESID, b, st));
    KNP_TRY(apply_B(c, o, tmp, w, st));
    if (D) KNP_TRY(launch_pointwise(n, w, D, 1, w, st));
    KNP_TRY(dots_to_host(c, 0, w, hp, st));
    const double beta = std::sqrt(hp[0]);
    info->rnorm = beta;
    if (!(beta == beta) || std::isinf(beta)) {
      set_error("GMRES: residual is not finite after %d iterations", its);
      info->iterations = its;
      return KNP_E_NOCONV;
    }
    if (beta <= tol || bnorm == 0.0) {
      info->converged = 1;
      break;
    }
    if (its >= o->max_it) break;
    if (o->refine > 0 && prev_beta > 0.0 && beta > 0.5 * prev_beta) {
      // "direct" mode: the true preconditioned residual stopped improving -> at the fp64 floor
      if (++stagn >= o->refine) {
        info->converged = 2;
        break;
      }
    }
    prev_beta = beta;
    // V0 = r / beta
    KNP_TRY(launch_axpby(n, 1.0 / beta, w, 0.0, c->V.p, st));
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int jdone = 0;
    bool done = false;
    for (int j = 0; j < m; ++j) {
      double* vj = c->V.p + (size_t)j * c->ldv;
      if (D) {
        KNP_TRY(launch_pointwise(n, vj, D, 0, c->tmp2.p, st));
        KNP_TRY(spmv_A(c, A_vals, c->tmp2.p, tmp, EPI_SET, nullptr, st));
      } else {
        KNP_TRY(spmv_A(c, A_vals, vj, tmp, EPI_SET, nullptr, st));
      }
      KNP_TRY(apply_B(c, o, tmp, w, st));
      if (D) KNP_TRY(launch_pointwise(n, w, D, 1, w, st));
      // classical Gram-Schmidt with refinement only if needed (DGKS criterion, PETSc's default
      // KSP_GMRES_CGS_REFINE_IFNEEDED): every pass is one fused multi-dot (+ ||w||^2) and one fused multi-axpy
      KNP_TRY(dots_to_host(c, j + 1, w, hp, st));
      double hsq = 0.0;
      for (int i = 0; i <= j; ++i) {
        Hat(i, j) = hp[i];
        hsq += hp[i] * hp[i];
      }
      const double before = hp[j + 1];
      KNP_TRY(launch_multi_axpy(n, j + 1, c->V.p, c->ldv, c->hdev.p, w, st));
      double nrm2 = before - hsq;
      if (!(nrm2 > 0.5 * before)) {
        KNP_TRY(dots_to_host(c, j + 1, w, hp, st));
        double h2sq = 0.0;
        for (int i = 0; i <= j; ++i) {
          Hat(i, j) += hp[i];
          h2sq += hp[i] * hp[i];
        }
        KNP_TRY(launch_multi_axpy(n, j + 1, c->V.p, c->ldv, c->hdev.p, w, st));
        nrm2 = hp[j + 1] - h2sq;
      }
      if (nrm2 < 0.0) nrm2 = 0.0;
      const double hn = std::sqrt(nrm2);
      Hat(j + 1, j) = hn;
      if (hn > 0.0) KNP_TRY(launch_axpby(n, 1.0 / hn, w, 0.0, c->V.p + (size_t)(j + 1) * c->ldv, st));
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * Hat(i, j) + sn[i] * Hat(i + 1, j);
        Hat(i + 1, j) = -sn[i] * Hat(i, j) + cs[i] * Hat(i + 1, j);
        Hat(i, j) = t;
      }
      const double den = std::hypot(Hat(j, j), Hat(j + 1, j));
      cs[j] = Hat(j, j) / den;
      sn[j] = Hat(j + 1, j) / den;
      Hat(j, j) = den;
      Hat(j + 1, j) = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      ++its;
      jdone = j + 1;
      info->rnorm = std::fabs(g[j + 1]);
      if (!(den == den)) {
        set_error("GMRES: breakdown (non-finite Hessenberg entry) at iteration %d", its);
        info->iterations = its;
        return KNP_E_NOCONV;
      }
      if (std::fabs(g[j + 1]) <= tol || its >= o->max_it || hn == 0.0) {
        done = std::fabs(g[j + 1]) <= tol;
        break;
      }
    }
    // y = H^-1 g ; x += V y
    for (int i = jdone - 1; i >= 0; --i) {
      double s = g[i];
      for (int k = i + 1; k < jdone; ++k) s -= Hat(i, k) * y[k];
      y[i] = s / Hat(i, i);
    }
    for (int i = 0; i < jdone; ++i) hp[i] = y[i];
    KNP_CUDA(cudaMemcpyAsync(c->ydev.p, hp, jdone * sizeof(double), cudaMemcpyHostToDevice, st));
    KNP_TRY(launch_update_x(n, jdone, c->V.p, c->ldv, c->ydev.p, x, D, st));
    KNP_CUDA(cudaStreamSynchronize(st));   // hp is reused by the next dots_to_host
    if (done && o->refine == 0) {
      info->converged = 1;
      break;
    }
    if (its >= o->max_it && !done) {
      // fall through to recompute the true residual once and exit
    }
  }
  info->iterations = its;
  if (o->zero_mean_solution) KNP_TRY(project_nullspace(c, x, st));
  KNP_TRY(halo_exchange(c, x, st));
  if (!info->converged) {
    set_error("GMRES did not converge: %d iterations, ||B r|| = %.3e, tol = %.3e", its, info->rnorm, tol);
    return KNP_E_NOCONV;
  }
  return KNP_OK;
}

}  // namespace knp
