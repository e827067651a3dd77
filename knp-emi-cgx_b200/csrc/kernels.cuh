// Kernel-side parameter block and launcher declarations.
#pragma once
#include "common.cuh"

namespace knp {


// Plain-old-data copy of the constants that enter the forms (passed by value to kernels).
struct KParams {
  double dt, F, C_M, psi, phi_rest;
  double z[3], D[3];
  double g_Na_bar, g_K_bar, g_leak[3], g_leak_g[3];
  double stim_lo[3], stim_hi[3];
  double K_e_init, K_i_g_init;
  int stim_dir[3];     // axes of the stimulus region (-1: unused; [0] = -1: no region)
  int ode_substeps, rush_larsen;
};

int launch_gate(const DevTopo& T, const KParams& P, const double* u, double* gates, cudaStream_t st);
int facet_ncomp(int gdim);
int launch_facets(const DevTopo& T, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim,
                  const double* u, const double* gates, double stim_fac, double* fe, cudaStream_t st);
int launch_rows(const DevTopo& T, const KParams& P, int mode, const double* u, const double* fe, double* vals,
                double* b, int max_deg, int max_gdeg, cudaStream_t st);
int launch_csr_indices(const DevTopo& T, int mode, int32_t* indices, cudaStream_t st);
int launch_l2_cells(int gdim, const Layout& L, int s, int field, int power, int n_cells, const int32_t* cell_nodes,
                    const int32_t* cell_tag, const int32_t* cell_owned, const double* node_x, int nodeoff,
                    const int32_t* tags, int n_tags, const double* u, double* partial, int n_partial,
                    cudaStream_t st);

int launch_probe(int n_out, const int32_t* ptr, const int32_t* col, const double* w, const double* u, double* out, cudaStream_t st);
int launch_stim_current(const DevTopo& T, const KParams& P, const int32_t* tag_stim, const int32_t* mf_owned,
                        const double* u, double stim_fac, double* partial, int n_partial, cudaStream_t st);

// ---- linalg.cu ----
enum SpmvEpi { EPI_SET = 0, EPI_RESID = 1, EPI_JACOBI = 2, EPI_ADD = 3, EPI_RESID0 = 4 };
// EPI_RESID0 (fused cycle tail only): the pre-smoothed iterate x = w dinv b is formed on the fly, out2 = x, out = b - A x
// out = epilogue(A x): SET: A x | RESID: b - A x | JACOBI: x + w*dinv*(b - A x) | ADD: out + A x
int launch_spmv(int n_rows, int64_t nnz, const int32_t* indptr, const int32_t* indices, const double* vals,
                const double* x, double* out, int epi, const double* b, const double* dinv, double w,
                cudaStream_t st);
// TMA-staged streaming SpMV over precomputed row blocks (falls back to launch_spmv when nblk <= 0)
int build_rowblocks(const int32_t* indptr_host, int n_rows, std::vector<int32_t>& blk);
int launch_spmv_stream(int nblk, const int32_t* rowblk, const int32_t* indptr, const int32_t* indices, const double* vals,
                       const double* x, double* out, int epi, const double* b, const double* dinv, double w, cudaStream_t st,
                       double avg_row = 16.0);
struct CsrView {
  int n_rows;
  int64_t nnz;
  const int32_t *indptr, *indices;
  const double* vals;
  const int32_t* rowblk;
  int nblk;
  const float* vals32 = nullptr;   // non-null: the values are stored in single precision (vals is ignored)
};
// out = epilogue(M x): streaming TMA kernel when row blocks exist and the arrays are 16-byte aligned, else CSR-vector
int spmv(const CsrView& M, const double* x, double* out, int epi, const double* b, const double* dinv, double w,
         cudaStream_t st);
int launch_scale_dinv(int n, double w, const double* dinv, const double* b, double* x, cudaStream_t st);   // x = w*dinv*b
int launch_extract_dinv(int n_rows, const int32_t* indptr, const int32_t* indices, const double* vals, double* dinv,
                        cudaStream_t st);
int launch_dense_gemv(int n, const double* Minv, const double* b, double* x, cudaStream_t st);
int launch_dense_gemv(int n, const float* Minv, const double* b, double* x, cudaStream_t st);
int launch_to_f32(int64_t n, const double* src, float* dst, cudaStream_t st);
// one operation of the fused cycle tail (linalg.cu::amg_tail_kernel)
enum TailType { TAIL_SPMV = 0, TAIL_DENSE = 1, TAIL_SCALE = 2 };
struct TailOp {
  int type, epi, n, lanes;
  const int32_t *indptr, *indices;
  const double* vals;        // CSR values, or the dense inverse
  const double* x;           // input vector
  double* out;
  double* out2;
  const double *b, *dinv;
  double w;
};
int launch_amg_tail(const TailOp* ops_dev, int nops, unsigned* bar, cudaStream_t st);
// Gram-Schmidt building blocks (deterministic two-stage reductions; no atomics)
constexpr int RED_BLOCKS = 1184;   // 8 x 148 SMs: full occupancy for the streaming reductions
// out[j] = sum_i V[j*ldv + i] * w[i], j < m ; out[m] = sum_i w[i]^2 ; partial is (m+1) x RED_BLOCKS scratch
// jsub >= 0: w is replaced on the fly by d = w - V[jsub]
int launch_multi_dot(int n, int m, const double* V, size_t ldv, const double* w, double* partial, double* out,
                     cudaStream_t st, int jsub = -1);
// w -= sum_j h[j] V_j   (h on device)
int launch_multi_axpy(int n, int m, const double* V, size_t ldv, const double* h, double* w, cudaStream_t st, int jsub = -1);
// the same, and vnext = w_new * inv_norm in the same pass (Gram-Schmidt projection + normalisation fused)
int launch_multi_axpy_normalize(int n, int m, const double* V, size_t ldv, const double* h, double* w, double* vnext,
                                double inv_norm, cudaStream_t st, int jsub = -1);
int launch_axpby(int n, double a, const double* x, double b, double* y, cudaStream_t st);              // y = a x + b y
int launch_scale_copy(int n, const double* alpha_dev, int invert, const double* x, double* y, cudaStream_t st);   // y = x*alpha or x/alpha
int launch_update_x(int n, int m, const double* V, size_t ldv, const double* y_dev, double* x, const double* scale,
                    cudaStream_t st);   // x += scale .* sum_j y[j] V_j   (scale may be NULL)
int launch_pointwise(int n, const double* a, const double* b, int divide, double* out, cudaStream_t st);   // out = a.*b or a./b
// sum over [lo0,hi0) U [lo1,hi1) of x -> out[0] (deterministic), and x[range] -= shift
int launch_range_sum(const double* x, int lo0, int hi0, int lo1, int hi1, double* partial, double* out, cudaStream_t st);
int launch_range_shift(double* x, int lo0, int hi0, int lo1, int hi1, const double* sum_dev, double inv_count,
                       cudaStream_t st);
int launch_reduce_partials(const double* partial, int n_partial, double* out, cudaStream_t st);
int launch_add_sparse(int n, const int32_t* rows, const double* vals, double* y, cudaStream_t st);   // y[rows[i]] += vals[i]
// Dirichlet conditions (knp_set_dirichlet): flags of the constrained columns, rows with a constrained entry, and the
// zero-rows-and-columns / lifting pass over those rows
int launch_bc_flags(int n_bc, const int32_t* bc_cols, uint8_t* flag, cudaStream_t st);
int launch_bc_touch(int n_rows, const int32_t* indptr, const int32_t* indices, const uint8_t* flag, uint8_t* touched,
                    cudaStream_t st);
int launch_bc_apply(int n_list, const int32_t* rows, const int32_t* indptr, const int32_t* indices, double* vals, double* b,
                    int n_bc, const int32_t* bc_cols, const double* bc_vals, double diag, cudaStream_t st);
int launch_bc_set(int n_bc, const int32_t* bc_cols, const double* bc_vals, int n_rows, double* x, cudaStream_t st);   // x[bc] = g (owned)
// preconditioned CG with device-resident scalars (solver.cu::cg_solve)
int launch_cg_scalar(int phase, int it, const double* dots, double* S, double* hist, cudaStream_t st);
int launch_cg_xr(int n, const double* S, const double* p, const double* q, double* x, double* r, cudaStream_t st);
int launch_cg_p(int n, const double* S, const double* z, double* p, cudaStream_t st);

}  // namespace knp
