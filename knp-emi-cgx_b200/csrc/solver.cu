// GPU Krylov solver: left-preconditioned restarted GMRES with classical Gram-Schmidt (two passes),
// preconditioned-norm convergence test relative to ||B b||, nullspace removal after every preconditioner
// application -- the semantics of the reference's PETSc KSP configuration
// (KNPEMIx_solver.py:212-214,276-280,324-333,386-389,435; PETSc defaults restated in SURVEY.md Appendix F).
// The preconditioner B is one V(1,1) cycle of our smoothed-aggregation hierarchy on P (amg_setup.cpp),
// Jacobi on P, or the identity.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
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
  // device memsets run on the legacy stream, asynchronously to the host and unordered with the context's non-blocking stream
  KNP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  KNP_TRY(c->partial.alloc((size_t)(restart + 2) * RED_BLOCKS));
  KNP_TRY(c->hdev.alloc(restart + 2));
  KNP_TRY(c->ydev.alloc(restart + 2));
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  KNP_CUDA(cudaMallocHost(&c->h_pinned, (restart + 2) * sizeof(double)));
  c->ws_restart = restart;
  return KNP_OK;
}

// KNP_AMG_TIMING=1: wall-clock phases of the preconditioner setup on stderr
struct SetupTimer {
  bool on = getenv("KNP_AMG_TIMING") && atoi(getenv("KNP_AMG_TIMING"));
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "pc setup: %-44s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

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

static int vcycle(Amg& M, int l, const double* bl, double* xout, cudaStream_t st);

// Single-precision STORAGE of the hierarchy operators (level matrices, prolongators, restrictions, dense coarsest inverse).
// One preconditioner application streams every operator of both hierarchies once or more (W-cycle) and is bound by that
// traffic; rounding the stored entries to float (relative perturbation 6e-8 of a preconditioner that is an approximation
// anyway) takes a third off the bytes per non-zero (12 -> 8).  All products are accumulated in double on double vectors, so
// the cycle remains a fixed linear operator and GMRES needs no flexible variant; the system matrix A and everything the
// solution is measured with stay in double.  KNP_AMG_F32=0 keeps double storage; the opt-in fused tail reads double too.
static bool amg_f32() {
  static const bool on = !(getenv("KNP_AMG_F32") && atoi(getenv("KNP_AMG_F32")) == 0) &&
                         !(getenv("KNP_FUSE_NNZ") && atoll(getenv("KNP_FUSE_NNZ")) > 0);
  return on;
}
static int to_f32(CsrDev& M, cudaStream_t st) {
  if (!amg_f32() || M.nnz == 0 || !M.vals.p) return KNP_OK;
  KNP_TRY(M.vals32.alloc((size_t)M.nnz));
  KNP_TRY(launch_to_f32(M.nnz, M.vals.p, M.vals32.p, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  M.vals.free();
  return KNP_OK;
}

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
  g_kernel_launches += 2ull * (unsigned long long)n;
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

static bool amg_setup_device_default(const knp_ctx* c) {
  // single-GPU runs build the hierarchy on the device (amg_device.cu: same decisions, bit-identical operators);
  // KNP_AMG_SETUP=host|device overrides, matrices the device form does not take (Dirichlet rows) fall back to the host
  static const char* where = getenv("KNP_AMG_SETUP");
  if (c->H.degree == 2) return false;       // P2 blocks: host setup (the device form has only been verified on P1 operators)
  return where ? !strcmp(where, "device") : c->nranks == 1;
}

// per-level work vectors, D^-1, single-precision storage; coarsest inverse (dense operator in cinv when spd, its inverse
// otherwise)
static int finish_amg(knp_ctx* c, std::unique_ptr<Amg>& amg, std::vector<double>& rhos, std::vector<double>& cinv, int n_coarse,
                      bool spd, int level0, std::unique_ptr<Amg>& out) {
  SetupTimer tm;
  for (size_t l = 0; l < amg->levels.size(); ++l) {
    AmgLevelDev* lv = amg->levels[l];
    lv->rho = rhos[l];
    const int nr = lv->A.n_rows;
    KNP_TRY(lv->dinv.alloc(nr));
    KNP_TRY(lv->x.alloc(nr));
    KNP_TRY(lv->b.alloc(nr));
    KNP_TRY(lv->r.alloc(nr));
    KNP_TRY(launch_extract_dinv(nr, lv->A.indptr.p, lv->A.indices.p, lv->A.vals.p, lv->dinv.p, c->stream));
    KNP_TRY(to_f32(lv->A, c->stream));
    KNP_TRY(to_f32(lv->P, c->stream));
    KNP_TRY(to_f32(lv->R, c->stream));
  }
  tm.lap("level vectors, D^-1, float32 storage");
  amg->n_coarse = n_coarse;
  KNP_TRY(amg->coarse_inv.upload(cinv));
  if (spd) KNP_TRY(dense_inverse_device(amg->n_coarse, amg->coarse_inv.p, c->stream));
  if (amg_f32() && amg->n_coarse > 0) {
    const int64_t nn = (int64_t)amg->n_coarse * amg->n_coarse;
    KNP_TRY(amg->coarse_inv32.alloc((size_t)nn));
    KNP_TRY(launch_to_f32(nn, amg->coarse_inv.p, amg->coarse_inv32.p, c->stream));
    KNP_CUDA(cudaStreamSynchronize(c->stream));
    amg->coarse_inv.free();
  }
  KNP_TRY(amg->cb.alloc(amg->n_coarse));
  KNP_TRY(amg->cx.alloc(amg->n_coarse));
  amg->level0 = level0;
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  tm.lap("coarsest inverse");
  out = std::move(amg);
  return KNP_OK;
}

// builds the hierarchy of A0 (host copy) and uploads it; spd: coarsest operator (<= coarse_size unknowns) inverted on
// the device
static int build_amg(knp_ctx* c, const CsrHost& A0, std::unique_ptr<Amg>& out, int coarse_size = 600, bool spd = false,
                     int level0 = 0, bool allow_device = true) {
  std::vector<CsrHost> As, Ps, Rs;
  std::vector<double> rhos, cinv;
  int used_device = 0;
  SetupTimer tm;
  if (allow_device && amg_setup_device_default(c))
    KNP_TRY(amg_setup_device(A0, 0.08, coarse_size, 16, As, Ps, Rs, rhos, cinv, c->stream, &used_device));
  if (used_device && !spd) {
    std::vector<double> dense;
    dense.swap(cinv);
    KNP_TRY(dense_inverse(As.back().n_rows, dense, cinv));
  }
  if (!used_device) KNP_TRY(amg_setup_host(A0, 0.08, coarse_size, 16, As, Ps, Rs, rhos, cinv, !spd));
  c->amg_setup_on_device = used_device;
  tm.lap(used_device ? "hierarchy (device setup, host in / out)" : "hierarchy (host setup)");
  auto amg = std::make_unique<Amg>();
  const int nl = (int)Ps.size();
  for (int l = 0; l < nl; ++l) {
    auto* lv = new AmgLevelDev();
    amg->levels.push_back(lv);
    KNP_TRY(upload_csr(As[l], lv->A));
    KNP_TRY(upload_csr(Ps[l], lv->P));
    KNP_TRY(upload_csr(Rs[l], lv->R));
  }
  tm.lap("upload of the levels, row blocks");
  const int n_coarse = As.back().n_rows;
  amg->hostA = std::move(As);
  return finish_amg(c, amg, rhos, cinv, n_coarse, spd, level0, out);
}

template <class T>
static void swap_buf(DevBuf<T>& a, DevBuf<T>& b) {
  std::swap(a.p, b.p);
  std::swap(a.n, b.n);
}

// device CSR of the setup -> operator of the cycle: padded row pointers, row blocks of the streaming SpMV; the arrays move.
// host_keep != nullptr receives the row pointers and -- when `full` -- the whole operator (inspection, knp_amg_level_host)
static int adopt_csr(DCsr& d, CsrDev& out, CsrHost* host_keep, bool full) {
  out.n_rows = d.n_rows;
  out.n_cols = d.n_cols;
  out.nnz = d.nnz;
  std::vector<int32_t> ip((size_t)d.n_rows + 1);
  KNP_CUDA(cudaMemcpy(ip.data(), d.indptr.p, ip.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (host_keep) {
    host_keep->n_rows = d.n_rows;
    host_keep->n_cols = d.n_cols;
    if (full) KNP_TRY(dcsr_download(d, *host_keep));
    else host_keep->indptr = ip;
  }
  std::vector<int32_t> blk;
  out.nblk = build_rowblocks(ip.data(), d.n_rows, blk);
  if (out.nblk > 0) KNP_TRY(out.rowblk.upload(blk));
  else out.nblk = 0;
  for (int k = 0; k < 4; ++k) ip.push_back(ip.back());           // padding for the 16-byte TMA slices
  KNP_TRY(out.indptr.upload(ip));
  swap_buf(out.indices, d.indices);
  swap_buf(out.vals, d.vals);
  return KNP_OK;
}

// The same for a matrix that already lives on the device: the hierarchy is built there and stays there (no host round trip
// of the operators; only row pointers and the small levels are copied for the row blocks / for inspection).  Falls back to the
// host setup when the device form does not take the matrix.
static int build_amg_dev(knp_ctx* c, std::unique_ptr<DCsr>& A0, std::unique_ptr<Amg>& out, int coarse_size, bool spd,
                         int level0 = 0) {
  DevHierarchy H;
  int used_device = 0;
  SetupTimer tm;
  KNP_TRY(amg_setup_device_core(A0, 0.08, coarse_size, 16, H, c->stream, &used_device));
  if (!used_device) {
    CsrHost h;
    KNP_TRY(dcsr_download(*A0, h));
    A0.reset();
    return build_amg(c, h, out, coarse_size, spd, level0, false);
  }
  c->amg_setup_on_device = 1;
  tm.lap("hierarchy (device setup, device resident)");
  auto amg = std::make_unique<Amg>();
  const int nl = (int)H.P.size();
  // host copies for inspection: every level but a large finest one (its row pointers only; KNP_AMG_KEEP_HOST=1 keeps all)
  static const bool keep_all = getenv("KNP_AMG_KEEP_HOST") && atoi(getenv("KNP_AMG_KEEP_HOST"));
  amg->hostA.assign(H.A.size(), CsrHost());
  KNP_TRY(dcsr_download(*H.A.back(), amg->hostA.back()));
  std::vector<double> cinv;
  {
    const CsrHost& Ac = amg->hostA.back();
    const int nc = Ac.n_rows;
    cinv.assign((size_t)nc * nc, 0.0);
    for (int i = 0; i < nc; ++i)
      for (int j = Ac.indptr[i]; j < Ac.indptr[i + 1]; ++j) cinv[(size_t)i * nc + Ac.indices[j]] += Ac.vals[j];
    if (!spd) {
      std::vector<double> dense;
      dense.swap(cinv);
      KNP_TRY(dense_inverse(nc, dense, cinv));
    }
  }
  const int n_coarse = H.A.back()->n_rows;
  for (int l = 0; l < nl; ++l) {
    auto* lv = new AmgLevelDev();
    amg->levels.push_back(lv);
    const bool full = keep_all || l > 0 || H.A[l]->n_rows <= 4000000;
    KNP_TRY(adopt_csr(*H.A[l], lv->A, &amg->hostA[l], full));
    KNP_TRY(adopt_csr(*H.P[l], lv->P, nullptr, false));
    KNP_TRY(adopt_csr(*H.R[l], lv->R, nullptr, false));
  }
  tm.lap("row blocks, host copies of the small levels");
  return finish_amg(c, amg, H.rhos, cinv, n_coarse, spd, level0, out);
}

// ---- fused cycle tail ------------------------------------------------------------------------------------------------
// Operation list of the sub-cycle below level `l` (the recursion of vcycle() written out; see linalg.cu::amg_tail_kernel).
static int tail_lanes(const CsrDev& M) {
  // few lanes per row = many rows in flight; every lane keeps four gathers in flight (linalg.cu::tail_row_sum)
  const double avg = M.n_rows > 0 ? (double)M.nnz / M.n_rows : 1.0;
  static const int force = getenv("KNP_TAIL_LANES") ? atoi(getenv("KNP_TAIL_LANES")) : 0;
  if (force) return force;
  return avg <= 8.0 ? 1 : avg <= 16.0 ? 2 : avg <= 32.0 ? 4 : 8;
}
static TailOp tail_spmv(const CsrDev& M, int epi, const double* x, double* out, const double* b, const double* dinv, double w,
                        double* out2 = nullptr) {
  TailOp o{};
  o.type = TAIL_SPMV;
  o.epi = epi;
  o.n = M.n_rows;
  o.lanes = tail_lanes(M);
  o.indptr = M.indptr.p;
  o.indices = M.indices.p;
  o.vals = M.vals.p;
  o.x = x;
  o.out = out;
  o.out2 = out2;
  o.b = b;
  o.dinv = dinv;
  o.w = w;
  return o;
}
static void emit_cycle(Amg& M, int l, const double* bl, double* xout, std::vector<TailOp>& ops) {
  const int nl = (int)M.levels.size();
  if (l == nl) {
    TailOp o{};
    o.type = TAIL_DENSE;
    o.n = M.n_coarse;
    o.vals = M.coarse_inv.p;
    o.x = bl;
    o.out = xout;
    ops.push_back(o);
    return;
  }
  AmgLevelDev& L = *M.levels[l];
  const double w = (4.0 / 3.0) / L.rho;
  const int lg = l + M.level0;
  const int reps = (lg >= 1 && lg <= M.gamma_last) ? M.gamma : 1;
  for (int rep = 0; rep < reps; ++rep) {
    // the first residual forms the pre-smoothed iterate x = w dinv b on the fly (no separate pass)
    if (rep == 0) ops.push_back(tail_spmv(L.A, EPI_RESID0, nullptr, L.r.p, bl, L.dinv.p, w, L.x.p));
    else ops.push_back(tail_spmv(L.A, EPI_RESID, L.x.p, L.r.p, bl, nullptr, 0.0));
    double* bc = (l + 1 == nl) ? M.cb.p : M.levels[l + 1]->b.p;
    ops.push_back(tail_spmv(L.R, EPI_SET, L.r.p, bc, nullptr, nullptr, 0.0));
    emit_cycle(M, l + 1, bc, L.r.p, ops);
    ops.push_back(tail_spmv(L.P, EPI_ADD, L.r.p, L.x.p, nullptr, nullptr, 0.0));
  }
  ops.push_back(tail_spmv(L.A, EPI_JACOBI, L.x.p, xout, bl, L.dinv.p, w));
}

// Chooses the first level whose operator is L2-resident and prebuilds the operation list of everything below it.  For
// fuse_from >= 1 the right-hand side / result pointers are the level buffers; a hierarchy that is small from level 0 on
// (test fixtures, the replicated tail of a distributed hierarchy) is fused for the caller's (in, out) pair.
static int prepare_tail(Amg& M, const double* in0, double* out0) {
  // opt-in (KNP_FUSE_NNZ = largest operator, in non-zeros, that goes into the fused tail): measured on C3 the fused tail is
  // correct but not faster than the graph-captured stream kernels, so the default is off
  static const int64_t fuse_nnz = getenv("KNP_FUSE_NNZ") ? atoll(getenv("KNP_FUSE_NNZ")) : 0;
  M.fuse_from = -1;
  M.tail_nops = 0;
  const int nl = (int)M.levels.size();
  if (fuse_nnz <= 0 || nl == 0) return KNP_OK;
  int from = -1;
  for (int l = 0; l < nl; ++l)
    if (M.levels[l]->A.nnz <= fuse_nnz) {
      from = l;
      break;
    }
  if (from < 0) return KNP_OK;
  if (from == 0 && (!in0 || !out0)) from = 1;
  if (from >= nl) return KNP_OK;
  std::vector<TailOp> ops;
  M.tail_in = from == 0 ? in0 : M.levels[from]->b.p;
  M.tail_out = from == 0 ? out0 : M.levels[from - 1]->r.p;
  emit_cycle(M, from, M.tail_in, M.tail_out, ops);
  std::vector<unsigned char> raw(ops.size() * sizeof(TailOp));
  memcpy(raw.data(), ops.data(), raw.size());
  KNP_TRY(M.tail_ops.upload(raw));
  if (!M.tail_bar.p) {
    KNP_TRY(M.tail_bar.alloc(2));
    KNP_CUDA(cudaMemset(M.tail_bar.p, 0, 2 * sizeof(unsigned)));
    KNP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  }
  M.tail_nops = (int)ops.size();
  M.fuse_from = from;
  return KNP_OK;
}

// ---- multi-GPU: row-distributed hierarchies (amg_dist.cpp) --------------------------------------------------------
// Every level operator is split by rows over the ranks like the system matrix itself; a level SpMV is preceded by one
// packed halo exchange of its input (in-place receives, dist.cu).  Below `repl_threshold` global rows the level is
// gathered onto every rank and the serial hierarchy continues redundantly, so the deep, latency-bound levels cost no
// communication at all.  The reference gets the same structure from hypre running across its MPI ranks
// (KNPEMIx_solver.py:269-273).
static int64_t repl_threshold() {
  static const int64_t v = getenv("KNP_AMG_REPL") ? atoll(getenv("KNP_AMG_REPL")) : 300000;
  return v;
}

static int build_dist_amg(knp_ctx* c, CsrHost&& A0, HaloHost&& halo0, std::vector<int32_t>&& gown, std::vector<int32_t>&& goidx,
                          std::unique_ptr<DistAmg>& out, int coarse_size, bool spd) {
  NcclAmgComm comm(c);
  DistHierarchyHost H;
  KNP_TRY(amg_dist_setup(comm, std::move(A0), std::move(halo0), std::move(gown), std::move(goidx), 0.08, repl_threshold(), 16, H));
  auto M = std::make_unique<DistAmg>();
  // one peer-visible arena for everything the neighbours write into: the [owned | ghost] iterate of every level and the
  // gathered right-hand side of the replicated level (sub-buffers 128-byte aligned)
  const int64_t ng = H.repl_off.back();
  auto pad16 = [](size_t n) { return (n + 15) / 16 * 16; };
  std::vector<size_t> xoff;
  size_t arena_n = 0;
  for (const DistLevelHost& h : H.levels) {
    xoff.push_back(arena_n);
    arena_n += pad16((size_t)h.n_own + h.n_ghost);
  }
  const size_t gb_off = arena_n;
  arena_n += pad16((size_t)ng) + 16;
  KNP_TRY(M->arena.alloc(arena_n));
  KNP_CUDA(cudaMemsetAsync(M->arena.p, 0, arena_n * sizeof(double), c->stream));
  M->gb = M->arena.p + gb_off;
  for (DistLevelHost& h : H.levels) {
    auto lv = std::make_unique<DistLevelDev>();
    lv->n_own = h.n_own;
    lv->n_ghost = h.n_ghost;
    lv->rho = h.rho;
    KNP_TRY(upload_csr(h.A, lv->A));
    KNP_TRY(upload_csr(h.P, lv->P));
    KNP_TRY(upload_csr(h.R, lv->R));
    KNP_TRY(halo_upload(h.halo, h.n_own, lv->halo));
    KNP_TRY(lv->dinv.alloc(h.n_own));
    lv->x = M->arena.p + xoff[M->levels.size()];
    KNP_TRY(lv->b.alloc(h.n_own));
    KNP_TRY(lv->r.alloc(h.n_own));
    KNP_TRY(launch_extract_dinv(h.n_own, lv->A.indptr.p, lv->A.indices.p, lv->A.vals.p, lv->dinv.p, c->stream));
    KNP_TRY(to_f32(lv->A, c->stream));
    KNP_TRY(to_f32(lv->P, c->stream));
    KNP_TRY(to_f32(lv->R, c->stream));
    {
      // direct NVLink exchange of this level: the neighbours store into the ghost tail of lv->x
      const int np = (int)h.halo.peers.size();
      std::vector<int64_t> sb(np), sc(np), off(np);
      for (int i = 0; i < np; ++i) {
        sb[i] = h.halo.send_ptr[i];
        sc[i] = h.halo.send_ptr[i + 1] - h.halo.send_ptr[i];
        off[i] = (int64_t)xoff[M->levels.size()] + h.n_own + h.halo.recv_ptr[i];
      }
      KNP_TRY(peer_link_create(c, h.halo.peers, sb, sc, M->arena.p, off, lv->halo.link));
      lv->halo.link_x = lv->x;
    }
    h.P = CsrHost();
    h.R = CsrHost();
    M->hostA.push_back(std::move(h.A));
    M->levels.push_back(std::move(lv));
  }
  M->off = H.repl_off;
  KNP_TRY(M->gx.alloc((size_t)ng));
  {
    // gather of the replicated level: every rank stores its piece of the right-hand side into every peer's gb
    std::vector<int32_t> peers;
    std::vector<int64_t> sb, sc, off;
    const int64_t mine = H.repl_off[c->rank + 1] - H.repl_off[c->rank];
    for (int r = 0; r < c->nranks; ++r)
      if (r != c->rank) {
        peers.push_back(r);
        sb.push_back(0);
        sc.push_back(mine);
        off.push_back((int64_t)gb_off + H.repl_off[r]);          // where rank r's piece lands in MY gb
      }
    KNP_TRY(peer_link_create(c, peers, sb, sc, M->arena.p, off, M->gather));
  }
  KNP_TRY(M->rb.alloc((size_t)(H.repl_off[c->rank + 1] - H.repl_off[c->rank]) + 1));
  KNP_TRY(build_amg(c, H.Arepl, M->tail, coarse_size, spd, (int)M->levels.size()));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  out = std::move(M);
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

// Level-0 operator of one part of the preconditioner on a multi-GPU run: this rank's rows in the part's compact numbering
// with the ghost columns KEPT (numbered in the order the main halo receives them, so that the part's own halo receives in
// place), the part's halo lists (the main lists filtered to the part) and, per ghost, the owner's compact index (obtained
// with one exchange of an index vector through the main halo).  map(i): full row / owned column -> compact index or -1;
// field_in_part(s, f): does field block (s, f) belong to the part.
template <class MapFn, class FieldFn>
static int dist_part(knp_ctx* c, const std::vector<int32_t>& ip, const std::vector<int32_t>& idx, const std::vector<double>& val,
                     const std::vector<double>& owner_index, MapFn map, FieldFn field_in_part, int n_part, CsrHost& A,
                     HaloHost& halo, std::vector<int32_t>& gown, std::vector<int32_t>& goidx) {
  const Layout& L = c->T.L;
  const int n = L.n_rows, np = (int)c->peers.size();
  const int ngh_full = L.n_cols - n;
  // field block of a full-layout ghost column
  auto ghost_in_part = [&](int col) -> bool {
    const int g = col - n;
    const int s = g >= L.gbase[1] ? 1 : 0;
    const int f = L.n_gh[s] > 0 ? (g - L.gbase[s]) / L.n_gh[s] : 0;
    return field_in_part(s, f);
  };
  std::vector<int32_t> ghost_id(ngh_full, -1);
  halo.peers.clear();
  halo.send_ptr.assign(1, 0);
  halo.recv_ptr.assign(1, 0);
  halo.send_idx.clear();
  gown.clear();
  goidx.clear();
  for (int i = 0; i < np; ++i) {
    for (int64_t k = c->send_ptr[i]; k < c->send_ptr[i + 1]; ++k) {
      const int cc = map(c->h_send_cols[k]);
      if (cc >= 0) halo.send_idx.push_back(cc);
    }
    for (int64_t k = c->recv_ptr[i]; k < c->recv_ptr[i + 1]; ++k) {
      const int col = c->h_recv_cols[k];
      if (!ghost_in_part(col)) continue;
      ghost_id[col - n] = (int32_t)gown.size();
      gown.push_back(c->peers[i]);
      goidx.push_back((int32_t)owner_index[col]);
    }
    halo.peers.push_back(c->peers[i]);
    halo.send_ptr.push_back((int64_t)halo.send_idx.size());
    halo.recv_ptr.push_back((int64_t)gown.size());
  }
  A.n_rows = n_part;
  A.n_cols = n_part + (int)gown.size();
  A.indptr.assign(1, 0);
  A.indices.clear();
  A.vals.clear();
  for (int i = 0; i < n; ++i) {
    if (map(i) < 0) continue;
    const size_t row0 = A.indices.size();
    bool sorted = true;
    for (int j = ip[i]; j < ip[i + 1]; ++j) {
      int cc;
      if (idx[j] < n) {
        cc = map(idx[j]);
      } else {
        cc = ghost_id[idx[j] - n];
        if (cc >= 0) cc += n_part;
      }
      if (cc < 0) continue;
      if (A.indices.size() > row0 && A.indices.back() >= cc) sorted = false;
      A.indices.push_back(cc);
      A.vals.push_back(val[j]);
    }
    if (!sorted) {                                   // ghost columns arrive in halo order, not ascending
      std::vector<std::pair<int32_t, double>> row(A.indices.size() - row0);
      for (size_t t = 0; t < row.size(); ++t) row[t] = {A.indices[row0 + t], A.vals[row0 + t]};
      std::sort(row.begin(), row.end());
      for (size_t t = 0; t < row.size(); ++t) {
        A.indices[row0 + t] = row[t].first;
        A.vals[row0 + t] = row[t].second;
      }
    }
    A.indptr.push_back((int32_t)A.indices.size());
  }
  KNP_CHECK((int)A.indptr.size() == n_part + 1, "distributed preconditioner part: row count mismatch");
  return KNP_OK;
}

// owner's compact index of every ghost column: x[i] = index(i) on owned rows, one main halo exchange
template <class IndexFn>
static int exchange_owner_index(knp_ctx* c, IndexFn index, std::vector<double>& out) {
  const Layout& L = c->T.L;
  out.assign(L.n_cols, -1.0);
  for (int i = 0; i < L.n_rows; ++i) out[i] = (double)index(i);
  DevBuf<double> d;
  KNP_TRY(d.upload(out));
  KNP_TRY(halo_exchange(c, d.p, c->stream));
  KNP_CUDA(cudaMemcpyAsync(out.data(), d.p, out.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  return KNP_OK;
}

// ---- single GPU: the ion / potential blocks of P~ and the lumped M_sigma are formed on the device ----
// compact numbering of a part: ion part [s=0: 3 n0 | s=1: 3 n1], potential part [n0 | n1]; -1: the row / column belongs to
// the other part
__device__ __forceinline__ int part_map(int i, int n0, int n1, int part) {
  if (part == 0) {
    if (i < 3 * n0) return i;
    if (i < 4 * n0) return -1;
    if (i < 4 * n0 + 3 * n1) return i - n0;
    return -1;
  }
  if (i < 3 * n0) return -1;
  if (i < 4 * n0) return i - 3 * n0;
  if (i < 4 * n0 + 3 * n1) return -1;
  return i - 3 * n0 - 3 * n1;
}
__global__ void part_count_kernel(int n, int n0, int n1, int part, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                                  int32_t* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = part_map(i, n0, n1, part);
  if (m < 0) return;
  int k = 0;
  for (int j = ip[i]; j < ip[i + 1]; ++j) k += ix[j] < n && part_map(ix[j], n0, n1, part) >= 0;
  cnt[m] = k;
}
__global__ void part_ptr_kernel(int m, const int32_t* __restrict__ cnt, int32_t* __restrict__ ptr) {
  // exclusive scan by one block (setup time, a few milliseconds): 1024 threads, each over a contiguous chunk
  __shared__ long long tot[1024];
  const int t = threadIdx.x;
  const long long chunk = ((long long)m + 1023) / 1024;
  const long long b = t * chunk, e = min((long long)m, b + chunk);
  long long s = 0;
  for (long long i = b; i < e; ++i) s += cnt[i];
  tot[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int k = 0; k < 1024; ++k) {
      const long long v = tot[k];
      tot[k] = run;
      run += v;
    }
    ptr[m] = (int32_t)run;
  }
  __syncthreads();
  long long run = tot[t];
  for (long long i = b; i < e; ++i) {
    ptr[i] = (int32_t)run;
    run += cnt[i];
  }
}
__global__ void part_fill_kernel(int n, int n0, int n1, int part, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                                 const double* __restrict__ val, const int32_t* __restrict__ optr, int32_t* __restrict__ oix,
                                 double* __restrict__ ov) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = part_map(i, n0, n1, part);
  if (m < 0) return;
  int pos = optr[m];
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    if (ix[j] >= n) continue;
    const int cc = part_map(ix[j], n0, n1, part);
    if (cc < 0) continue;
    oix[pos] = cc;
    ov[pos++] = val[j];
  }
}
// 1 / (M_sigma): (sum_k z_k^2 c_k / psi) at the node times the lumped mass of the node, in the host's operation order.
// Lumped mass: the row sum of the mass matrix (P1), or hrz x its diagonal entry (hrz > 0: P2, whose row sums vanish at the
// vertices in 2D and are negative in 3D -- HRZ lumping keeps the total mass with positive weights)
__global__ void msig_kernel(Layout L, const int32_t* __restrict__ ipP, const int32_t* __restrict__ ixP,
                            const double* __restrict__ mval, double hrz,
                            const double* __restrict__ u, double z0, double z1, double z2, double psi, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = L.n_own[0], n1 = L.n_own[1];
  if (t >= n0 + n1) return;
  const int s = t < n0 ? 0 : 1, p = s ? t - n0 : t;
  const int row = L.row(s, 0, p);
  double ms = 0.0;
  if (hrz > 0.0) {
    for (int j = ipP[row]; j < ipP[row + 1]; ++j)
      if (ixP[j] == row) ms = __dmul_rn(hrz, mval[j]);
  } else {
    for (int j = ipP[row]; j < ipP[row + 1]; ++j) ms = __dadd_rn(ms, mval[j]);
  }
  const double z[3] = {z0, z1, z2};
  double sig = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) sig = __dadd_rn(sig, __dmul_rn(__ddiv_rn(__dmul_rn(z[k], z[k]), psi), u[L.col(s, k, p)]));
  out[t] = ms != 0.0 ? __ddiv_rn(1.0, __dmul_rn(sig, ms)) : 0.0;
}
static int extract_part(knp_ctx* c, int part, std::unique_ptr<DCsr>& out) {
  const Layout& L = c->T.L;
  const int n = L.n_rows, n0 = L.n_own[0], n1 = L.n_own[1];
  const int m = part == 0 ? 3 * (n0 + n1) : n0 + n1;
  cudaStream_t st = c->stream;
  out = std::make_unique<DCsr>();
  out->n_rows = out->n_cols = m;
  DevBuf<int32_t> cnt;
  KNP_TRY(cnt.alloc((size_t)std::max(m, 1)));
  KNP_TRY(out->indptr.alloc((size_t)m + 1));
  part_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, n0, n1, part, c->d_indptr_P.p, c->d_indices_P.p, cnt.p);
  KNP_LAUNCHED();
  part_ptr_kernel<<<1, 1024, 0, st>>>(m, cnt.p, out->indptr.p);
  KNP_LAUNCHED();
  int32_t nnz = 0;
  KNP_CUDA(cudaMemcpyAsync(&nnz, out->indptr.p + m, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  out->nnz = nnz;
  KNP_TRY(out->indices.alloc((size_t)std::max(nnz, 1)));
  KNP_TRY(out->vals.alloc((size_t)std::max(nnz, 1)));
  part_fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, n0, n1, part, c->d_indptr_P.p, c->d_indices_P.p, c->P_vals.p,
                                                    out->indptr.p, out->indices.p, out->vals.p);
  KNP_LAUNCHED();
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

static int schur_setup(knp_ctx* c) {
  const Layout& L = c->T.L;
  const int n = L.n_rows, n0 = L.n_own[0], n1 = L.n_own[1];
  KNP_CHECK(L.rowbase[0] == 0 && L.rowbase[1] == 4 * n0, "unexpected row layout");
  cudaStream_t st = c->stream;
  SetupTimer tm;
  // P~: ion blocks M + dt D_k K, phi blocks K_phi + (C_M/F) M_Gamma (the sign the membrane term has in `a`)
  KParams kp = c->kp;
  kp.C_M = -c->kp.C_M;
  KNP_TRY(launch_rows(c->T, kp, 1, c->u.p, c->fe.p, c->P_vals.p, nullptr, c->H.max_deg, c->H.max_gdeg, st));
  // mass matrices: the same kernel with D = 0 leaves M in the ion blocks
  kp = c->kp;
  for (int k = 0; k < 3; ++k) kp.D[k] = 0.0;
  KNP_TRY(c->M_vals.alloc(c->H.nnz_P));
  KNP_TRY(launch_rows(c->T, kp, 1, c->u.p, c->fe.p, c->M_vals.p, nullptr, c->H.max_deg, c->H.max_gdeg, st));
  if (c->n_bc > 0) {
    // Dirichlet dofs are cut out of every operator of the preconditioner (identity rows in the blocks, empty rows and
    // columns in the mass matrices): with the boundary values in the initial guess their residual is zero throughout
    KNP_TRY(launch_bc_apply(c->n_bc_rows_P, c->bc_rows_P.p, c->d_indptr_P.p, c->d_indices_P.p, c->P_vals.p, nullptr,
                            c->n_bc, c->bc_cols.p, c->bc_vals.p, 1.0, st));
    KNP_TRY(launch_bc_apply(c->n_bc_rows_P, c->bc_rows_P.p, c->d_indptr_P.p, c->d_indices_P.p, c->M_vals.p, nullptr,
                            c->n_bc, c->bc_cols.p, c->bc_vals.p, 0.0, st));
  }
  KNP_CUDA(cudaStreamSynchronize(st));
  // one GPU: blocks, hierarchies and M_sigma are formed on the device; the operators never visit the host
  const bool dev_path = c->nranks == 1 && amg_setup_device_default(c);
  std::vector<int32_t> idx(dev_path ? 0 : c->H.nnz_P);
  std::vector<double> val(idx.size()), mval(idx.size()), u(dev_path ? 0 : L.n_cols);
  if (!dev_path) {
    KNP_CUDA(cudaMemcpy(idx.data(), c->d_indices_P.p, idx.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    KNP_CUDA(cudaMemcpy(val.data(), c->P_vals.p, val.size() * sizeof(double), cudaMemcpyDeviceToHost));
    KNP_CUDA(cudaMemcpy(mval.data(), c->M_vals.p, mval.size() * sizeof(double), cudaMemcpyDeviceToHost));
    KNP_CUDA(cudaMemcpy(u.data(), c->u.p, u.size() * sizeof(double), cudaMemcpyDeviceToHost));
  }
  const std::vector<int32_t>& ip = c->H.indptr_P;
  tm.lap("P~ and M assembly, copies to the host");
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
  for (int i = 0; i < n && c->nranks == 1 && !dev_path; ++i) {     // multi-GPU runs keep the ghost columns (dist_part below)
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
  tm.lap("ion / potential blocks extracted");
  if (c->nranks > 1) {
    std::vector<double> oidx;
    KNP_TRY(exchange_owner_index(c, [&](int i) { return cmap(i) >= 0 ? cmap(i) : pmap(i); }, oidx));
    HaloHost hc, hp;
    std::vector<int32_t> goc, gic, gop, gip;
    KNP_TRY(dist_part(c, ip, idx, val, oidx, cmap, [](int, int f) { return f < 3; }, 3 * (n0 + n1), Acc, hc, goc, gic));
    KNP_TRY(dist_part(c, ip, idx, val, oidx, pmap, [](int, int f) { return f == 3; }, n0 + n1, App, hp, gop, gip));
    KNP_TRY(build_dist_amg(c, std::move(Acc), std::move(hc), std::move(goc), std::move(gic), c->damg_c, 2500, true));
    KNP_TRY(build_dist_amg(c, std::move(App), std::move(hp), std::move(gop), std::move(gip), c->damg_p, 2500, true));
  } else if (dev_path) {
    std::unique_ptr<DCsr> dcc, dpp;
    KNP_TRY(extract_part(c, 0, dcc));
    KNP_TRY(extract_part(c, 1, dpp));
    tm.lap("ion / potential blocks extracted (device)");
    KNP_TRY(build_amg_dev(c, dcc, c->amg_c, 2500, true));
    int on_dev = c->amg_setup_on_device;
    KNP_TRY(build_amg_dev(c, dpp, c->amg_p, 2500, true));
    c->amg_setup_on_device = on_dev && c->amg_setup_on_device;
  } else {
    KNP_TRY(build_amg(c, Acc, c->amg_c, 2500, true));
    KNP_TRY(build_amg(c, App, c->amg_p, 2500, true));
  }
  // W-cycle on every level except the finest and the coarsest sparse one (level 0 is visited once, levels 1..depth-2 with
  // cycle index 2, the last sparse level and the dense coarsest solve as often as their parent): the W recursion restores
  // the two-level convergence rate of the deep hierarchies (C3: 36 -> 15 iterations), and stopping it one level early costs
  // no iteration but halves the launch-bound visits of the two smallest levels (C3: 2.18 -> 2.04 ms per application).
  // KNP_W_LEVELS overrides the last W level.
  auto last_w_level = [](int depth) {
    if (const char* e = getenv("KNP_W_LEVELS")) return atoi(e);
    return std::max(1, depth - 2);
  };
  for (Amg* a : {c->amg_c.get(), c->amg_p.get()})
    if (a) {
      a->gamma = 2;
      a->gamma_last = last_w_level((int)a->levels.size());
    }
  for (DistAmg* a : {c->damg_c.get(), c->damg_p.get()})
    if (a) {
      a->gamma = 2;
      a->gamma_last = last_w_level((int)a->levels.size() + (int)a->tail->levels.size());
      a->tail->gamma = 2;
      a->tail->gamma_last = a->gamma_last;
    }
  tm.lap("both hierarchies");
  // the P buffer now holds the sign-flipped Schur form, not the reference's block-Jacobi P: pc kinds 1 / 2 must re-assemble
  c->P_assembled = false;
  KNP_TRY(c->sch_vc.alloc((size_t)3 * (n0 + n1)));
  KNP_TRY(c->sch_zc.alloc((size_t)3 * (n0 + n1)));
  KNP_TRY(c->sch_t.alloc((size_t)n0 + n1));
  KNP_TRY(c->sch_zp.alloc((size_t)n0 + n1));
  if (c->amg_c) KNP_TRY(prepare_tail(*c->amg_c, c->sch_vc.p, c->sch_zc.p));
  if (c->amg_p) KNP_TRY(prepare_tail(*c->amg_p, c->sch_t.p, c->sch_zp.p));
  for (DistAmg* a : {c->damg_c.get(), c->damg_p.get()})
    if (a) KNP_TRY(prepare_tail(*a->tail, a->gb, a->gx.p));
  // lumped M_sigma = (sum_k z_k^2 c_k / psi) at the node  x  row sum of the mass matrix
  std::vector<double> msig_inv(dev_path ? 0 : (size_t)n0 + n1);
  const double* z = c->kp.z;
  const double hrz = c->H.degree == 2 ? c->H.p2.hrz : 0.0;        // P2: HRZ lumping (see msig_kernel)
  if (dev_path) {
    KNP_TRY(c->msig_inv.alloc((size_t)n0 + n1));
    if (n0 + n1 > 0) {
      msig_kernel<<<(n0 + n1 + 255) / 256, 256, 0, st>>>(L, c->d_indptr_P.p, c->d_indices_P.p, c->M_vals.p, hrz, c->u.p, z[0],
                                                         z[1], z[2], c->kp.psi, c->msig_inv.p);
      KNP_LAUNCHED();
    }
  }
  for (int s = 0; s < 2 && !dev_path; ++s)
    for (int p = 0; p < L.n_own[s]; ++p) {
      const int row = L.row(s, 0, p);
      double ms = 0.0;
      for (int j = ip[row]; j < ip[row + 1]; ++j) {
        if (hrz > 0.0) ms = idx[j] == row ? hrz * mval[j] : ms;
        else ms += mval[j];
      }
      double sig = 0.0;
      for (int k = 0; k < 3; ++k) sig += z[k] * z[k] / c->kp.psi * u[L.col(s, k, p)];
      msig_inv[(size_t)(s ? n0 : 0) + p] = ms != 0.0 ? 1.0 / (sig * ms) : 0.0;     // empty row: Dirichlet dof
    }
  if (!dev_path) KNP_TRY(c->msig_inv.upload(msig_inv));
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
  KNP_TRY(c->sch_q.alloc(L.n_cols));
  KNP_TRY(c->sch_rhs.alloc(L.n_rows));
  KNP_CUDA(cudaMemset(c->sch_q.p, 0, (size_t)L.n_cols * sizeof(double)));
  KNP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));      // legacy-stream memset vs the context's non-blocking stream
  tm.lap("lumped mass, row blocks of M, work vectors");
  return KNP_OK;
}

static int vcycle_dist(knp_ctx* c, DistAmg& M, int l, const double* bl, double* xout, cudaStream_t st);

static int schur_apply(knp_ctx* c, const double* r, double* zout, cudaStream_t st) {
  const Layout& L = c->T.L;
  const int n0 = L.n_own[0], n1 = L.n_own[1];
  const double* z = c->kp.z;
  int grid = (n0 + n1 + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  schur_split_kernel<<<grid, 256, 0, st>>>(L, z[0], z[1], z[2], r, c->sch_vc.p, c->sch_t.p);
  KNP_LAUNCHED();
  if (c->damg_c) KNP_TRY(vcycle_dist(c, *c->damg_c, 0, c->sch_vc.p, c->sch_zc.p, st));
  else KNP_TRY(vcycle(*c->amg_c, 0, c->sch_vc.p, c->sch_zc.p, st));
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
  if (c->damg_p) KNP_TRY(vcycle_dist(c, *c->damg_p, 0, c->sch_t.p, c->sch_zp.p, st));
  else KNP_TRY(vcycle(*c->amg_p, 0, c->sch_t.p, c->sch_zp.p, st));
  schur_merge_kernel<<<grid, 256, 0, st>>>(L, c->sch_zc.p, c->sch_zp.p, c->sch_t.p, c->msig_inv.p, c->sch_vc.p, zout, nullptr);
  KNP_LAUNCHED();
  return KNP_OK;
}

void pc_graphs_clear(knp_ctx* c);

int pc_setup(knp_ctx* c, const knp_solve_opts* o) {
  const int n = c->T.L.n_rows;
  pc_graphs_clear(c);
  c->amg.reset();
  c->amg_c.reset();
  c->amg_p.reset();
  c->damg.reset();
  c->damg_c.reset();
  c->damg_p.reset();
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
  // host copy of P: the whole matrix on one GPU, this rank's rows with their ghost columns on several
  std::vector<int32_t> idx(c->H.nnz_P);
  std::vector<double> val(c->H.nnz_P);
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  KNP_CUDA(cudaMemcpy(idx.data(), c->d_indices_P.p, idx.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  KNP_CUDA(cudaMemcpy(val.data(), c->P_vals.p, val.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CsrHost P0;
  if (c->nranks > 1) {
    std::vector<double> oidx;
    KNP_TRY(exchange_owner_index(c, [](int i) { return i; }, oidx));
    HaloHost h;
    std::vector<int32_t> go, gi;
    KNP_TRY(dist_part(c, c->H.indptr_P, idx, val, oidx, [](int i) { return i; }, [](int, int) { return true; }, n, P0, h, go, gi));
    KNP_TRY(build_dist_amg(c, std::move(P0), std::move(h), std::move(go), std::move(gi), c->damg, 600, false));
    return prepare_tail(*c->damg->tail, c->damg->gb, c->damg->gx.p);
  }
  P0.n_rows = n;
  P0.n_cols = n;
  P0.indptr = c->H.indptr_P;
  P0.indices.swap(idx);
  P0.vals.swap(val);
  KNP_TRY(build_amg(c, P0, c->amg));
  return prepare_tail(*c->amg, nullptr, nullptr);
}

// z = V-cycle(r); level-l right-hand side in bl, result written to xout (distinct from bl)
static CsrView view(const CsrDev& M) {
  return CsrView{M.n_rows, M.nnz, M.indptr.p, M.indices.p, M.vals.p, M.rowblk.p, M.nblk, M.vals32.p};
}

static int vcycle(Amg& M, int l, const double* bl, double* xout, cudaStream_t st) {
  const int nl = (int)M.levels.size();
  if (l == M.fuse_from && M.tail_nops > 0 && bl == M.tail_in && xout == M.tail_out)
    return launch_amg_tail(reinterpret_cast<const TailOp*>(M.tail_ops.p), M.tail_nops, M.tail_bar.p, st);
  if (l == nl)
    return M.coarse_inv32.p ? launch_dense_gemv(M.n_coarse, M.coarse_inv32.p, bl, xout, st)
                            : launch_dense_gemv(M.n_coarse, M.coarse_inv.p, bl, xout, st);
  AmgLevelDev& L = *M.levels[l];
  const int n = L.A.n_rows;
  const double w = (4.0 / 3.0) / L.rho;
  // pre-smooth from a zero initial guess
  KNP_TRY(launch_scale_dinv(n, w, L.dinv.p, bl, L.x.p, st));
  // coarse-grid correction; levels >= 1 repeat it `gamma` times (gamma = 2: W-cycle below the finest level, which
  // restores the two-level convergence rate of deep hierarchies at ~25 % extra cost because level 0 is visited once)
  const int lg = l + M.level0;        // level number inside a distributed hierarchy whose tail this is
  const int reps = (lg >= 1 && lg <= M.gamma_last) ? M.gamma : 1;
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

// the same cycle over the row-distributed levels: every SpMV with a level operator is preceded by the halo exchange of its
// input; P and R are rank-local.  At the replicated level the ranks' pieces of the right-hand side are gathered (grouped
// ncclSend/ncclRecv), every rank runs the serial tail and keeps its own piece of the result.
static int vcycle_dist(knp_ctx* c, DistAmg& M, int l, const double* bl, double* xout, cudaStream_t st) {
  const int nl = (int)M.levels.size();
  if (l == nl) {
    const int R = c->nranks, me = c->rank;
    const size_t mine = (size_t)(M.off[me + 1] - M.off[me]) * sizeof(double);
    if (M.gather.ready) {
      KNP_TRY(peer_push(c, M.gather, nullptr, bl, st));
    } else {
      std::vector<P2POp> ops;
      for (int r = 0; r < R; ++r) {
        if (r == me) continue;
        if (mine) ops.push_back({r, const_cast<double*>(bl), mine, true});
        const size_t theirs = (size_t)(M.off[r + 1] - M.off[r]) * sizeof(double);
        if (theirs) ops.push_back({r, M.gb + M.off[r], theirs, false});
      }
      KNP_TRY(p2p_exchange(c, ops, st));
    }
    if (mine) KNP_CUDA(cudaMemcpyAsync(M.gb + M.off[me], bl, mine, cudaMemcpyDeviceToDevice, st));
    KNP_TRY(vcycle(*M.tail, 0, M.gb, M.gx.p, st));
    if (mine) KNP_CUDA(cudaMemcpyAsync(xout, M.gx.p + M.off[me], mine, cudaMemcpyDeviceToDevice, st));
    return KNP_OK;
  }
  DistLevelDev& L = *M.levels[l];
  const int n = L.n_own;
  const double w = (4.0 / 3.0) / L.rho;
  KNP_TRY(launch_scale_dinv(n, w, L.dinv.p, bl, L.x, st));
  const int reps = (l >= 1 && l <= M.gamma_last) ? M.gamma : 1;
  for (int rep = 0; rep < reps; ++rep) {
    KNP_TRY(halo_exchange_inplace(c, L.halo, L.x, st));
    KNP_TRY(spmv(view(L.A), L.x, L.r.p, EPI_RESID, bl, nullptr, 0.0, st));
    double* bc = (l + 1 == nl) ? M.rb.p : M.levels[l + 1]->b.p;
    double* xc = L.r.p;
    KNP_TRY(spmv(view(L.R), L.r.p, bc, EPI_SET, nullptr, nullptr, 0.0, st));
    KNP_TRY(vcycle_dist(c, M, l + 1, bc, xc, st));
    KNP_TRY(spmv(view(L.P), xc, L.x, EPI_ADD, nullptr, nullptr, 0.0, st));
  }
  KNP_TRY(halo_exchange_inplace(c, L.halo, L.x, st));
  KNP_TRY(spmv(view(L.A), L.x, xout, EPI_JACOBI, bl, L.dinv.p, w, st));
  return KNP_OK;
}

__global__ void dinv_mul_kernel(int n, const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ z) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) z[i] = dinv[i] * r[i];
}

void pc_graphs_clear(knp_ctx* c) {
  for (auto& g : c->pc_graphs) cudaGraphExecDestroy(g.exec);
  c->pc_graphs.clear();
  c->pc_applies = 0;
}

// The Schur application is ~200 small launches (W-cycle over two hierarchies): on one GPU it is captured once per
// (r, z) pointer pair into a CUDA graph and replayed (GMRES always applies it to the same two buffers).
static int schur_apply_graphed(knp_ctx* c, const double* r, double* z, cudaStream_t st) {
  static const bool enabled = !(getenv("KNP_PC_GRAPH") && atoi(getenv("KNP_PC_GRAPH")) == 0);
  // multi-GPU: the application holds the peer-memory exchange kernels (or NCCL point-to-point groups), which capture like
  // any other kernel (measured on 4 GPUs: 97.3 -> 90.9 ms per step); KNP_PC_GRAPH_MULTI=0 switches the capture off
  static const bool multi = !(getenv("KNP_PC_GRAPH_MULTI") && atoi(getenv("KNP_PC_GRAPH_MULTI")) == 0);
  if (!enabled || (c->nranks > 1 && !multi)) return schur_apply(c, r, z, st);
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

// ---- algorithmic bytes of one preconditioner application (roofline denominator of bench.py) -----------------------
// Every operation of the cycle counted once with its minimal traffic: a CSR product moves 12 B per non-zero (value +
// column), 4 B per row pointer, its input and its output vector once, plus the epilogue operands (b, D^-1, the updated
// iterate); the dense coarsest solve reads the inverse once per visit; the Schur glue kernels read / write each of their
// vectors once.  Visits follow the cycle index (W on levels 1..gamma_last).
static double spmv_bytes(const CsrDev& M, int extra_vectors) {
  const double per_nnz = M.vals32.p ? 8.0 : 12.0;      // value (single- or double-precision storage) + column index
  return per_nnz * (double)M.nnz + 4.0 * M.n_rows + 8.0 * M.n_cols + 8.0 * M.n_rows + 8.0 * (double)extra_vectors * M.n_rows;
}
static double level_bytes(const CsrDev& A, const CsrDev& P, const CsrDev& R, int reps, double child) {
  double b = 24.0 * A.n_rows;                                                     // x = w D^-1 b
  b += reps * (spmv_bytes(A, 1) + spmv_bytes(R, 0) + child + spmv_bytes(P, 1));  // residual, restriction, child, x += P x_c
  return b + spmv_bytes(A, 2);                                                    // Jacobi sweep (b, D^-1)
}
static double amg_bytes(const Amg& M, int l) {
  const int nl = (int)M.levels.size();
  if (l == nl) return (M.coarse_inv32.p ? 4.0 : 8.0) * (double)M.n_coarse * M.n_coarse + 16.0 * M.n_coarse;
  const AmgLevelDev& L = *M.levels[l];
  const int lg = l + M.level0;
  const int reps = (lg >= 1 && lg <= M.gamma_last) ? M.gamma : 1;
  return level_bytes(L.A, L.P, L.R, reps, amg_bytes(M, l + 1));
}
static double damg_bytes(const DistAmg& M, int l) {
  if (l == (int)M.levels.size()) return 16.0 * (double)M.off.back() + amg_bytes(*M.tail, 0);   // gather + replicated tail
  const DistLevelDev& L = *M.levels[l];
  const int reps = (l >= 1 && l <= M.gamma_last) ? M.gamma : 1;
  return level_bytes(L.A, L.P, L.R, reps, damg_bytes(M, l + 1)) + 16.0 * (reps + 1) * L.n_ghost;   // halo: packed + received
}
double pc_bytes(const knp_ctx* c) {
  const Layout& L = c->T.L;
  const double n01 = (double)L.n_own[0] + L.n_own[1];
  if (c->pc_kind == 1) return 24.0 * L.n_rows;
  if (c->pc_kind == 2) return c->damg ? damg_bytes(*c->damg, 0) : (c->amg ? amg_bytes(*c->amg, 0) : 0.0);
  if (c->pc_kind != 3) return 16.0 * L.n_rows;
  double b = 64.0 * n01 + 32.0 * n01 + 80.0 * n01;                 // split (4 in, 4 out), q (3 in, 1 out), merge (6 in, 4 out)
  for (int s = 0; s < 2; ++s) {
    if (L.n_own[s] == 0) continue;
    const int row0 = L.row(s, 0, 0);
    const double nnz_s = (double)c->H.indptr_P[row0 + L.n_own[s]] - c->H.indptr_P[row0];
    b += 12.0 * nnz_s + 28.0 * L.n_own[s];                          // t += M q
  }
  b += c->damg_c ? damg_bytes(*c->damg_c, 0) : (c->amg_c ? amg_bytes(*c->amg_c, 0) : 0.0);
  b += c->damg_p ? damg_bytes(*c->damg_p, 0) : (c->amg_p ? amg_bytes(*c->amg_p, 0) : 0.0);
  return b;
}

int pc_apply(knp_ctx* c, const double* r, double* z, cudaStream_t st) {
  const int n = c->T.L.n_rows;
  if (c->pc_kind == 2 && c->damg) return vcycle_dist(c, *c->damg, 0, r, z, st);
  if (c->pc_kind == 2 && c->amg) return vcycle(*c->amg, 0, r, z, st);
  if (c->pc_kind == 3 && ((c->damg_c && c->damg_p) || (c->amg_c && c->amg_p))) return schur_apply_graphed(c, r, z, st);
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

// Optional breakdown of one solve by CUDA events (KNP_SOLVE_TIMING=1, stderr): where the time between the kernels goes.
struct SolveTimer {
  bool on;
  cudaStream_t st;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
  explicit SolveTimer(cudaStream_t s) : on(getenv("KNP_SOLVE_TIMING") && atoi(getenv("KNP_SOLVE_TIMING"))), st(s) {}
  template <class F>
  int run(int cat, F&& f) {
    if (!on) return f();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    const int rc = f();
    cudaEventRecord(b, st);
    spans.push_back({cat, {a, b}});
    return rc;
  }
  void report(int its) {
    if (!on) return;
    cudaStreamSynchronize(st);
    double t[4] = {0, 0, 0, 0};
    float first_last = 0.f;
    if (!spans.empty()) cudaEventElapsedTime(&first_last, spans.front().second.first, spans.back().second.second);
    double gap[4][4] = {};      // idle time between the end of a span of category a and the start of the next span (category b)
    const bool verbose = atoi(getenv("KNP_SOLVE_TIMING")) > 1;
    for (size_t i = 0; i + 1 < spans.size(); ++i) {
      float ms = 0.f, len = 0.f;
      cudaEventElapsedTime(&ms, spans[i].second.second, spans[i + 1].second.first);
      cudaEventElapsedTime(&len, spans[i].second.first, spans[i].second.second);
      gap[spans[i].first][spans[i + 1].first] += ms;
      if (verbose) fprintf(stderr, "  span %zu cat %d len %.3f ms, gap to next (cat %d) %.3f ms\n", i, spans[i].first, len, spans[i + 1].first, ms);
    }
    for (auto& s : spans) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s.second.first, s.second.second);
      t[s.first] += ms;
      cudaEventDestroy(s.second.first);
      cudaEventDestroy(s.second.second);
    }
    fprintf(stderr, "solve timing: %d iterations, A-SpMV %.2f ms, preconditioner (+ nullspace) %.2f ms, Gram-Schmidt %.2f ms, other %.2f ms; "
                    "first to last event %.2f ms; gaps: spmv->pc %.2f, pc->gs %.2f, gs->gs %.2f, gs->spmv %.2f, pc->spmv %.2f\n", its, t[0],
            t[1], t[2], t[3], first_last, gap[0][1], gap[1][2], gap[2][2], gap[2][0], gap[1][0]);
    spans.clear();
  }
};

int gmres_solve(knp_ctx* c, const double* A_vals, const double* b, double* x, const knp_solve_opts* o,
                knp_solve_info* info, cudaStream_t st) {
  const int n = c->T.L.n_rows;
  SolveTimer tm(st);
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
  KNP_TRY(tm.run(1, [&] { return apply_B(c, o, b, w, st); }));
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
    KNP_TRY(tm.run(0, [&] { return spmv_A(c, A_vals, x, tmp, EPI_RESID, b, st); }));
    KNP_TRY(tm.run(1, [&] { return apply_B(c, o, tmp, w, st); }));
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
        KNP_TRY(tm.run(0, [&] { return spmv_A(c, A_vals, vj, tmp, EPI_SET, nullptr, st); }));
      }
      KNP_TRY(tm.run(1, [&] { return apply_B(c, o, tmp, w, st); }));
      if (D) KNP_TRY(launch_pointwise(n, w, D, 1, w, st));
      // classical Gram-Schmidt with refinement only if needed (DGKS criterion, PETSc's default
      // KSP_GMRES_CGS_REFINE_IFNEEDED): every pass is one fused multi-dot (+ squared norm) and one fused multi-axpy.
      // The pass is applied to d = w - v_j = (B A - I) v_j, not to w: with a good preconditioner w is v_j plus a small
      // perturbation, so projecting w removes almost all of it and the DGKS test asks for a second pass in EVERY iteration
      // (measured: Gram-Schmidt was 29 % of the solve).  d spans the same Krylov space, H(:, j) = e_j + V^T d, the new
      // direction d - V V^T d equals w - V V^T w exactly, and the projection of d is free of that cancellation, so one
      // pass suffices.  d is formed on the fly inside both kernels (v_j is one of the rows they read anyway).
      KNP_TRY(tm.run(2, [&] {
        KNP_TRY(launch_multi_dot(n, j + 1, c->V.p, c->ldv, w, c->partial.p, c->hdev.p, st, j));
        KNP_TRY(allreduce_sum(c, c->hdev.p, j + 2, st));
        KNP_CUDA(cudaMemcpyAsync(hp, c->hdev.p, (j + 2) * sizeof(double), cudaMemcpyDeviceToHost, st));
        KNP_CUDA(cudaStreamSynchronize(st));
        return (int)KNP_OK;
      }));
      double hsq = 0.0;
      for (int i = 0; i <= j; ++i) {
        Hat(i, j) = hp[i] + (i == j ? 1.0 : 0.0);
        hsq += hp[i] * hp[i];
      }
      const double before = hp[j + 1];
      double nrm2 = before - hsq;
      double* vnext = c->V.p + (size_t)(j + 1) * c->ldv;
      bool normalized = false;
      // Second pass only if the first one cancelled more than a factor 10 (eta = 0.1): a pass amplifies the rounding error
      // of the orthogonality by ||d|| / ||d - V V^T d||, so up to that ratio the basis stays orthogonal to ~10 eps and the
      // Pythagorean norm is accurate to eps / eta^2.  (The classical DGKS constant 1/sqrt(2) asks for the second pass
      // whenever the projection removes more than 30 % -- with a good preconditioner that is every iteration; PETSc's own
      // default, KSP_GMRES_CGS_REFINE_NEVER, never refines.)
      static const double eta2 = getenv("KNP_GS_ETA2") ? atof(getenv("KNP_GS_ETA2")) : 0.01;
      if (tm.on && atoi(getenv("KNP_SOLVE_TIMING")) > 2) fprintf(stderr, "  gs j=%d ratio^2 = %.3e\n", j, nrm2 / before);
      if (nrm2 > eta2 * before) {
        // no refinement needed (the usual case): the norm after the projection is known from the Pythagorean identity, so
        // the projection and the normalisation of the next basis vector are ONE pass over w and V
        KNP_TRY(tm.run(2, [&] { return launch_multi_axpy_normalize(n, j + 1, c->V.p, c->ldv, c->hdev.p, w, vnext, 1.0 / std::sqrt(nrm2), st, j); }));
        normalized = true;
      } else {
        KNP_TRY(tm.run(2, [&] {
          KNP_TRY(launch_multi_axpy(n, j + 1, c->V.p, c->ldv, c->hdev.p, w, st, j));      // w <- d - V h
          return dots_to_host(c, j + 1, w, hp, st);
        }));
        double h2sq = 0.0;
        for (int i = 0; i <= j; ++i) {
          Hat(i, j) += hp[i];
          h2sq += hp[i] * hp[i];
        }
        KNP_TRY(tm.run(2, [&] { return launch_multi_axpy(n, j + 1, c->V.p, c->ldv, c->hdev.p, w, st); }));
        nrm2 = hp[j + 1] - h2sq;
      }
      if (nrm2 < 0.0) nrm2 = 0.0;
      const double hn = std::sqrt(nrm2);
      Hat(j + 1, j) = hn;
      if (hn > 0.0 && !normalized) KNP_TRY(launch_axpby(n, 1.0 / hn, w, 0.0, vnext, st));
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
    KNP_TRY(tm.run(3, [&] { return launch_update_x(n, jdone, c->V.p, c->ldv, c->ydev.p, x, D, st); }));
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
  tm.report(its);
  if (o->zero_mean_solution) KNP_TRY(project_nullspace(c, x, st));
  KNP_TRY(halo_exchange(c, x, st));
  if (c->nranks > 1) {
    KNP_CUDA(cudaStreamSynchronize(st));
    KNP_TRY(peer_error_check(c));
  }
  if (!info->converged) {
    set_error("GMRES did not converge: %d iterations, ||B r|| = %.3e, tol = %.3e", its, info->rnorm, tol);
    return KNP_E_NOCONV;
  }
  return KNP_OK;
}

// Preconditioned conjugate gradients (ksp_type "cg", KNPEMIx_solver.py:212 passes the type through to PETSc): same convergence
// test as the GMRES path (preconditioned norm ||B r|| relative to ||B b||, nonzero initial guess, nullspace removed after
// every preconditioner application).  The loop is device resident: alpha, beta, (r, z) and the norm history live in device
// memory (linalg.cu::cg_scalar_kernel), the vector updates read them from there, and the host looks at the state only every
// `KNP_CG_CHECK` (default 4) iterations; once the state is "converged" the update kernels are no-ops, so the iterate is the one
// of the converged step whatever the check interval.  Valid for symmetric positive definite operators and preconditioners
// only (the coupled KNP-EMI matrix is not symmetric: PETSc would run CG on it all the same, and so do we, reporting a
// breakdown when (p, A p) <= 0).
int cg_solve(knp_ctx* c, const double* A_vals, const double* b, double* x, const knp_solve_opts* o, knp_solve_info* info,
             cudaStream_t st) {
  const int n = c->T.L.n_rows;
  KNP_TRY(ensure_workspace(c, o->restart > 3 ? o->restart : 30));
  const int max_it = o->max_it > 0 ? o->max_it : 5000;
  if (c->cg_hist.n < (size_t)max_it + 2) KNP_TRY(c->cg_hist.alloc((size_t)max_it + 2));
  if (!c->cg_scal.p) KNP_TRY(c->cg_scal.alloc(8));
  static const int check = getenv("KNP_CG_CHECK") && atoi(getenv("KNP_CG_CHECK")) > 0 ? atoi(getenv("KNP_CG_CHECK")) : 4;
  double* hp = c->h_pinned;
  double* r = c->V.p;
  double* z = c->V.p + c->ldv;
  double* p = c->V.p + 2 * c->ldv;
  double* q = c->V.p + 3 * c->ldv;
  double* S = c->cg_scal.p;
  info->iterations = 0;
  info->converged = 0;
  // ||B b||
  KNP_TRY(apply_B(c, o, b, z, st));
  KNP_TRY(dots_to_host(c, 0, z, hp, st));
  const double bnorm = std::sqrt(hp[0]);
  info->rnorm0 = info->rnorm = bnorm;
  if (!(bnorm == bnorm) || std::isinf(bnorm)) {
    set_error("CG: right-hand side is not finite");
    return KNP_E_NOCONV;
  }
  const double tol = o->rtol * bnorm;
  {
    double init[8] = {0.0, 0.0, 0.0, 0.0, tol * tol, 0.0, 0.0, 0.0};
    for (int i = 0; i < 8; ++i) hp[i] = init[i];
    KNP_CUDA(cudaMemcpyAsync(S, hp, 8 * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  // r = b - A x ; z = B r ; p = z
  KNP_TRY(spmv_A(c, A_vals, x, r, EPI_RESID, b, st));
  KNP_TRY(apply_B(c, o, r, z, st));
  KNP_TRY(launch_multi_dot(n, 1, r, c->ldv, z, c->partial.p, c->hdev.p, st));
  KNP_TRY(allreduce_sum(c, c->hdev.p, 2, st));
  KNP_TRY(launch_cg_scalar(0, 0, c->hdev.p, S, c->cg_hist.p, st));
  KNP_CUDA(cudaMemcpyAsync(p, z, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  int it = 0, state = 0, it_state = 0;
  auto poll = [&]() -> int {
    KNP_CUDA(cudaMemcpyAsync(hp, S, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    KNP_CUDA(cudaStreamSynchronize(st));
    state = (int)hp[3];
    it_state = (int)hp[5];
    return KNP_OK;
  };
  KNP_TRY(poll());
  while (state == 0 && it < max_it) {
    const int stop = std::min(max_it, it + check);
    for (; it < stop; ++it) {
      KNP_TRY(spmv_A(c, A_vals, p, q, EPI_SET, nullptr, st));
      KNP_TRY(launch_multi_dot(n, 1, p, c->ldv, q, c->partial.p, c->hdev.p, st));
      KNP_TRY(allreduce_sum(c, c->hdev.p, 2, st));
      KNP_TRY(launch_cg_scalar(1, it + 1, c->hdev.p, S, c->cg_hist.p, st));
      KNP_TRY(launch_cg_xr(n, S, p, q, x, r, st));
      KNP_TRY(apply_B(c, o, r, z, st));
      KNP_TRY(launch_multi_dot(n, 1, r, c->ldv, z, c->partial.p, c->hdev.p, st));
      KNP_TRY(allreduce_sum(c, c->hdev.p, 2, st));
      KNP_TRY(launch_cg_scalar(2, it + 1, c->hdev.p, S, c->cg_hist.p, st));
      KNP_TRY(launch_cg_p(n, S, z, p, st));
    }
    KNP_TRY(poll());
  }
  const int its = state != 0 ? it_state : it;
  info->iterations = its;
  KNP_CUDA(cudaMemcpyAsync(hp, c->cg_hist.p + its, sizeof(double), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  info->rnorm = std::sqrt(hp[0]);
  if (o->zero_mean_solution) KNP_TRY(project_nullspace(c, x, st));
  KNP_TRY(halo_exchange(c, x, st));
  if (c->nranks > 1) {
    KNP_CUDA(cudaStreamSynchronize(st));
    KNP_TRY(peer_error_check(c));
  }
  if (state == 2) {
    set_error("CG: breakdown at iteration %d ((p, A p) <= 0 or a non-finite norm): operator or preconditioner not SPD", its);
    return KNP_E_NOCONV;
  }
  if (state == 1) {
    info->converged = 1;
    return KNP_OK;
  }
  set_error("CG did not converge: %d iterations, ||B r|| = %.3e, tol = %.3e", its, info->rnorm, tol);
  return KNP_E_NOCONV;
}

// ksp.solve dispatch on the Krylov type
int krylov_solve(knp_ctx* c, const double* A_vals, const double* b, double* x, const knp_solve_opts* o, knp_solve_info* info,
                 cudaStream_t st) {
  // essential boundary conditions: the initial guess takes the boundary values, so the residual of the constrained rows
  // (identity rows, b = g) is zero from the first Krylov vector on and the solution carries them exactly
  if (c->n_bc > 0) KNP_TRY(launch_bc_set(c->n_bc, c->bc_cols.p, c->bc_vals.p, c->T.L.n_rows, x, st));
  if (o->ksp_type != 0 && o->ksp_type != 1) {
    set_error("unknown ksp_type %d (0 = gmres, 1 = cg)", o->ksp_type);
    return KNP_E_INVALID;
  }
  const int rc = o->ksp_type == 1 ? cg_solve(c, A_vals, b, x, o, info, st) : gmres_solve(c, A_vals, b, x, o, info, st);
  // a pinned potential shares its vertex with unconstrained ion rows, whose residuals reach it through the row operation of
  // the Schur preconditioner: the iterate carries g only to solver tolerance there.  The other unknowns do not depend on it
  // (its column is zero), so the exact value is simply restored.
  if (c->n_bc > 0 && rc == KNP_OK) KNP_TRY(launch_bc_set(c->n_bc, c->bc_cols.p, c->bc_vals.p, c->T.L.n_rows, x, st));
  return rc;
}

}  // namespace knp
