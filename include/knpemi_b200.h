/* knpemi_b200.h -- C ABI of the B200-native KNP-EMI timestep library (libknpemi_b200.so).
 *
 * The reference (hherlyng/knp-emi-cgx, "CGx") has no FFI boundary of its own: its hot path is a
 * chain of Python calls into DOLFINx / multiphenicsx / PETSc.  Each entry point below states the
 * reference call site(s) it replaces (paths relative to /root/reference/src/CGx).  The Python
 * mirror of the reference classes (knp-emi-cgx_b200/{problem,solver,ionic_models}.py) binds these
 * with ctypes; INTEGRATION.md shows the stub a CGx maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative KNP_E_* code on error; knp_last_error()
 *     returns a thread-local message.
 *   - "host" pointers are plain CPU memory; "dev" pointers are CUDA device memory on the
 *     context's device.  Nothing here takes or returns a torch type.
 *   - one context per GPU; a context is not thread-safe; `stream` is a cudaStream_t passed as
 *     void* (NULL = the context's own stream).
 *   - unknown ordering is the reference's: field-major blocks
 *       [Na_i K_i Cl_i phi_i | Na_e K_e Cl_e phi_e], block (s,f) restricted to the vertices of
 *     subdomain s in ascending vertex order (multiphenicsx DofMapRestriction, owned first).
 *     Vectors in "column layout" have n_cols = n_rows + n_ghost_cols entries (ghost tail filled by
 *     the halo exchange); on a single GPU n_cols == n_rows.
 */
#ifndef KNPEMI_B200_H
#define KNPEMI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KNP_OK 0
#define KNP_E_INVALID (-1)   /* bad argument / inconsistent mesh */
#define KNP_E_CUDA (-2)      /* CUDA runtime error */
#define KNP_E_NOCONV (-3)    /* Krylov solver did not converge / produced non-finite values */
#define KNP_E_UNSUPPORTED (-4)
#define KNP_E_NCCL (-5)

typedef struct knp_ctx knp_ctx;

/* ---- membrane model flags (IonicModel subclasses, KNPEMI/KNPEMIx_ionic_model.py) ---- */
#define KNP_MODEL_PASSIVE     1u   /* PassiveModel            :77-91   */
#define KNP_MODEL_KIRNA       2u   /* KirNaKPumpModel         :93-222  */
#define KNP_MODEL_GLIAL_CT    4u   /* GlialCotransporters     :224-298 */
#define KNP_MODEL_NEURONAL_CT 8u   /* NeuronalCotransporters  :300-369 */
#define KNP_MODEL_ATP        16u   /* ATPPump                 :371-424 */
#define KNP_MODEL_HH         32u   /* HodgkinHuxley           :426-515 */

/* Local mesh of one rank (the whole mesh on a single GPU).  Vertices [0,n_owned_vertices) are owned
 * by this rank, the rest are ghosts; cells = every cell touching an owned vertex.  Replaces the
 * XDMF read + dS entity ordering of utils/mixed_dim_problem.py:634-733 as *input* to the path. */
typedef struct {
  int32_t gdim;                 /* 2 (triangles) or 3 (tetrahedra) */
  int64_t n_vertices;
  int64_t n_owned_vertices;
  const double* coords;         /* host, n_vertices x gdim, already scaled (mesh_conversion_factor) */
  int64_t n_cells;
  const int32_t* cell_verts;    /* host, n_cells x (gdim+1) */
  const int32_t* cell_tags;     /* host, n_cells */
  int32_t n_intra_tags;
  const int32_t* intra_tags;    /* host */
  int32_t extra_tag;
  int64_t n_mfacets;
  const int32_t* mfacet_verts;  /* host, n_mfacets x gdim : membrane facets (intra/extra interfaces) */
  const int32_t* mfacet_tags;   /* host, n_mfacets */
  const uint8_t* cell_owned;    /* host, n_cells or NULL (=all): this rank integrates functionals over the cell */
  const uint8_t* mfacet_owned;  /* host, n_mfacets or NULL (=all) */
  int32_t n_quad;               /* facet quadrature rule (degree 10 in the reference, :732-733) */
  const double* quad_bary;      /* host, n_quad x gdim barycentric points on the facet */
  const double* quad_w;         /* host, n_quad weights summing to 1 */
  int32_t degree;               /* element order, fem_order of utils/mixed_dim_problem.py:207-208: 0 or 1 = P1 (everything above
                                   as stated); 2 = P2 on ONE GPU: "vertices" are then the nodes of the P2 space -- the mesh
                                   vertices first, then one node per edge with the coordinates of its midpoint --, cell_verts
                                   lists per cell its gdim+1 vertices followed by its edge nodes in the order (0,1),(0,2),
                                   [(0,3),](1,2),[(1,3),(2,3)] of the local vertices (6 / 10 per cell), mfacet_verts per facet
                                   its gdim vertices followed by its edge nodes in the same order (3 / 6 per facet); the
                                   quadrature rule stays barycentric on the facet's gdim vertices */
} knp_mesh_desc;

typedef struct {
  int64_t n_rows, n_cols, nnz, nnz_P;
  int64_t n_own[2], n_loc[2];   /* restricted dofs per subdomain: owned, owned+ghost */
  int64_t n_mverts, n_mfacets;
  int64_t n_cells[2];
  int32_t max_deg, max_gdeg;
} knp_sizes;

/* ProblemKNPEMI.setup_constants (KNPEMI/KNPEMIx_problem.py:909-981) + the config keys of
 * utils/mixed_dim_problem.py:187-356 that enter the forms. */
typedef struct {
  double dt, F, R, T, C_M, phi_rest;
  double z[3], D[3];
  double g_Na_bar, g_K_bar, g_leak[3], g_leak_g[3];
  double g_syn_bar, a_syn, T_stim;
  int32_t scale_stimulus;
  int32_t stim_dir[3];          /* stimulus_region (utils/mixed_dim_problem.py:334-356): up to three axes (`multiple`), -1 = unused;
                                   stim_dir[0] = -1: no region */
  double stim_lo[3], stim_hi[3];  /* mask = prod_i [lo_i < x_{dir_i} < hi_i]  (KNPEMIx_ionic_model.py:558-587) */
  double K_e_init, K_i_g_init;  /* KirNaKPumpModel.E_K_init (:117) */
  int32_t ode_substeps;         /* HodgkinHuxley time_steps_ODE (:431) */
  int32_t rush_larsen;          /* use_Rush_Larsen (:430) */
  double stim_area;             /* global integral of mask over stimulus facets; <=0: computed locally */
} knp_params;

typedef struct {
  int32_t tag;
  uint32_t models;              /* OR of KNP_MODEL_* active on this membrane tag */
  int32_t stimulated;           /* tag in stimulus_tags (HH adds the synaptic Na current) */
} knp_tag_models;

/* KSP options of SolverKNPEMI.setup_solver (KNPEMI/KNPEMIx_solver.py:152-295). */
typedef struct {
  double rtol;                  /* ksp_rtol, preconditioned residual norm relative to ||B b|| */
  int32_t max_it;               /* ksp_max_it (5000) */
  int32_t restart;              /* GMRES restart (PETSc default 30) */
  int32_t pc;                   /* 0 none, 1 Jacobi(P), 2 smoothed-aggregation AMG V-cycle on P,
                                   3 charge-conservation Schur preconditioner (AMG on the ion and potential blocks) */
  int32_t project_nullspace;    /* remove the phi-constant nullspace after each PC apply (:324-333) */
  int32_t zero_mean_solution;   /* direct-solver convention: return the solution with ns^T x = 0 */
  int32_t refine;               /* "direct" mode: keep restarting until the true residual stagnates `refine` times */
  double field_scale[8];        /* >0: solve in variables scaled per field block (balances c ~ 1e2 against phi ~ 1e-2
                                   in the residual norm; used by the "direct" mode); all 0 = unscaled (PETSc semantics) */
  int32_t ksp_type;             /* 0 gmres (default), 1 cg: preconditioned conjugate gradients, device-resident loop
                                   (ksp_type is passed through to PETSc in the reference, KNPEMIx_solver.py:212) */
} knp_solve_opts;

typedef struct {
  int32_t iterations;
  int32_t converged;
  double rnorm0, rnorm;         /* preconditioned ||B b|| and final ||B r|| */
} knp_solve_info;

const char* knp_last_error(void);
int knp_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t knp_launch_count(void);

/* Build the restricted dof maps, CSR pattern (A and block-diagonal P) and all device-side gather maps.
 * Replaces: DofMapRestriction (KNPEMI/KNPEMIx_problem.py:85-94), create_matrix_block/create_vector_block
 * (KNPEMI/KNPEMIx_solver.py:157-161) and compute_integration_domains (utils/mixed_dim_problem.py:708-729). */
int knp_create(knp_ctx** out, const knp_mesh_desc* mesh, int device);
int knp_destroy(knp_ctx* ctx);
int knp_get_sizes(const knp_ctx* ctx, knp_sizes* out);
/* CSR structure of A (device views, borrowed) and host copies for parity checks. */
int knp_csr_dev(const knp_ctx* ctx, const int32_t** indptr, const int32_t** indices);
int knp_csr_host(const knp_ctx* ctx, int32_t* indptr, int32_t* indices);
int knp_csr_P_host(const knp_ctx* ctx, int32_t* indptr, int32_t* indices);
/* restricted dof -> local vertex id, per subdomain (length n_loc[s]) */
int knp_dofmap_host(const knp_ctx* ctx, int32_t* verts_intra, int32_t* verts_extra);
/* membrane vertex table (length n_mverts): local vertex id */
int knp_mverts_host(const knp_ctx* ctx, int32_t* verts);

int knp_set_params(knp_ctx* ctx, const knp_params* p, int32_t n_tags, const knp_tag_models* tags);
/* local integral of the stimulus mask over stimulated membrane facets owned by this rank
 * (HodgkinHuxley._add_stimulus, KNPEMI/KNPEMIx_ionic_model.py:591-598; the caller all-reduces). */
int knp_stimulus_area_local(knp_ctx* ctx, double* out);

/* State: u = previous solution in column layout (wh[*][*] restricted, KNPEMI/KNPEMIx_problem.py:51),
 * gates = n,m,h on membrane vertices (3 x n_mverts, KNPEMI/KNPEMIx_ionic_model.py:473-480). */
int knp_set_state(knp_ctx* ctx, const double* u_host, const double* gates_host);
int knp_get_state(knp_ctx* ctx, double* u_host, double* gates_host);
int knp_state_dev(knp_ctx* ctx, double** u_dev, double** gates_dev);
int knp_phi_m_host(knp_ctx* ctx, double* phi_m_host);     /* phi_i - phi_e on membrane vertices (:462-468) */

/* HodgkinHuxley.update_gating_variables (KNPEMI/KNPEMIx_ionic_model.py:605-671). */
int knp_gate_step(knp_ctx* ctx, void* stream);
/* SolverKNPEMI.assemble (KNPEMI/KNPEMIx_solver.py:104-116): A values (nnz) and b (n_rows) from the
 * current state at time t.  A_vals/b = NULL -> the context's own buffers. */
int knp_assemble(knp_ctx* ctx, double t, double* A_vals_dev, double* b_dev, void* stream);
/* assemble_preconditioner (KNPEMI/KNPEMIx_solver.py:118-135) for the block-Jacobi form
 * (KNPEMI/KNPEMIx_problem.py:717-738). */
int knp_assemble_P(knp_ctx* ctx, double* P_vals_dev, void* stream);
/* Time-independent source contribution to the right-hand side: b[rows[i]] += vals[i] in every knp_assemble / knp_step
 * (rows unique, owned).  Carries the ion-injection terms dt * (f_e, v)_{dx_e} of KNP-EMI (KNPEMI/KNPEMIx_problem.py:200-218,
 * 613-614), whose mass-matrix products the host forms once; n = 0 clears. */
int knp_set_source(knp_ctx* ctx, int32_t n, const int32_t* rows_host, const double* vals_host);
/* Essential boundary conditions, bcs = p.bcs of assemble_matrix_block / assemble_vector_block (KNPEMI/KNPEMIx_solver.py:113-116,
 * 123-126) for the conditions ProblemKNPEMI.setup_boundary_conditions builds (KNPEMI/KNPEMIx_problem.py:96-198: every field on
 * the exterior boundary, or phi_e pinned at one vertex): cols = the constrained dofs in the column layout -- owned AND ghost
 * columns of this rank -- and their values.  From then on knp_assemble / knp_step zero the rows and columns of these dofs
 * (diagonal 1) and lift the right-hand side (b_i -= sum_j A_ij g_j, b = g on constrained rows), knp_solve / knp_step start
 * from an initial guess that carries g, and knp_assemble_P and the preconditioner setups treat their matrices the same way.  The caller switches the nullspace
 * handling off (project_nullspace = zero_mean_solution = 0), as the reference does (:380,415).  n = 0 clears. */
int knp_set_dirichlet(knp_ctx* ctx, int32_t n, const int32_t* cols_host, const double* vals_host);
int knp_values_dev(knp_ctx* ctx, double** A_vals, double** b, double** P_vals, double** x);
/* y = A x with the context's CSR pattern (PETSc MatMult inside ksp.solve, :435). x in column layout. */
int knp_spmv(knp_ctx* ctx, const double* A_vals_dev, const double* x_dev, double* y_dev, void* stream);
/* KSP/PC setup (ksp.setOperators + ksp.setUp, :386-389): builds the AMG hierarchy from P. */
int knp_pc_setup(knp_ctx* ctx, const knp_solve_opts* opts);
/* z = B r : one preconditioner application (hypre V-cycle in the reference). */
int knp_pc_apply(knp_ctx* ctx, const double* r_dev, double* z_dev, void* stream);
/* Algorithmic bytes of ONE preconditioner application on this rank (every product / vector of the cycle counted once:
   12 B per non-zero, row pointers, input / output / epilogue vectors, the dense coarsest inverse per visit): the roofline
   denominator bench.py reports for the cycle that replaces hypre's (KNPEMIx_solver.py:269-273). */
int knp_pc_bytes(const knp_ctx* ctx, double* bytes);
/* ksp.solve(b, x) (:435) incl. nullspace handling (:297-335). x_dev in/out (column layout). */
int knp_solve(knp_ctx* ctx, const double* A_vals_dev, const double* b_dev, double* x_dev,
              const knp_solve_opts* opts, knp_solve_info* info, void* stream);
/* One pass of the time loop body (:365-468): t += dt, gate update (if any HH tag), assemble,
 * [nullspace.remove(b) on the first step], solve, u <- x.  Fully device resident. */
int knp_step(knp_ctx* ctx, const knp_solve_opts* opts, knp_solve_info* info, void* stream);
/* Same, through host buffers: H2D of (u, gates), step, D2H of (u, gates). Used for the e2e number. */
int knp_step_host(knp_ctx* ctx, double* u_host, double* gates_host, const knp_solve_opts* opts,
                  knp_solve_info* info);
int knp_set_time(knp_ctx* ctx, double t, int32_t step_index);
int knp_get_time(const knp_ctx* ctx, double* t, int32_t* step_index);
/* int u^2 dx over the owned share of cells of subdomain s with tag in tags (tests :45-51). */
int knp_l2_norm_sq(knp_ctx* ctx, int32_t subdomain, int32_t field, int32_t n_tags, const int32_t* tags,
                   double* out);
/* Conservation functionals of ProblemKNPEMI.print_conservation (KNPEMIx_problem.py:807-843): this rank's integral of
   u^power (power 1: ion amount, power 0: measure of the tagged cells, power 2 = knp_l2_norm_sq) over the owned cells of
   subdomain `subdomain` carrying one of `tags`; all-reduce the result over the ranks like the reference does. */
int knp_integral(knp_ctx* ctx, int32_t subdomain, int32_t field, int32_t power, int32_t n_tags, const int32_t* tags,
                 double* out);
/* this rank's part of the total stimulus current int stim_expr dS(stimulus_tags) at time t from the current state
   (SolverKNPEMI.init_png_data / save_png, KNPEMIx_solver.py:578-610; stim_expr: KNPEMIx_ionic_model.py:517-603) */
int knp_stimulus_current(knp_ctx* ctx, double t, double* out);
/* Point probes (scifem.evaluate_function in SolverKNPEMI.init_data / save_data, KNPEMI/KNPEMIx_solver.py:612-643): n_out sparse
   linear functionals out[i] = sum_{t in [ptr[i], ptr[i+1])} weights[t] * u[cols[t]] of the device state (columns in the
   column layout; the host mirror locates the containing cell and the barycentric weights once).  knp_probe_eval evaluates them
   on the device and copies the n_out values -- not the state -- to the host. */
int knp_probe_setup(knp_ctx* ctx, int32_t n_out, const int32_t* ptr, const int32_t* cols, const double* weights);
int knp_probe_eval(knp_ctx* ctx, double* out_host);
/* this rank's area of the membrane facets tagged `tag` (assemble_scalar(1*dS(tag)), KNPEMIx_problem.py:833-834) */
int knp_membrane_area(const knp_ctx* ctx, int32_t tag, double* out);
/* per-phase device timers of the last knp_step (ms): gate, facet, rows, solve, total */
int knp_last_timings(const knp_ctx* ctx, double* ms5);

/* plain copies on the context's stream, synchronous: kind 1 = host->device, 2 = device->host, 3 = device->device */
int knp_copy(knp_ctx* ctx, void* dst, const void* src, int64_t nbytes, int32_t kind);

/* AMG hierarchy inspection for level-by-level parity tests */
int knp_amg_num_levels(const knp_ctx* ctx);
/* levels of hierarchy `part`: pc 2 -> part 0 = the hierarchy on P; pc 3 -> part 0 = ion blocks, part 1 = potential blocks
   (levels are numbered part 0 first) */
int knp_amg_part_levels(const knp_ctx* ctx, int32_t part);
int knp_amg_level_sizes(const knp_ctx* ctx, int32_t level, int64_t* n, int64_t* nnz);
int knp_amg_level_host(const knp_ctx* ctx, int32_t level, int32_t* indptr, int32_t* indices, double* vals);
/* Host-only (no GPU): sizes, restricted dof maps (DofMapRestriction, KNPEMIx_problem.py:85-94) and the CSR pattern of A
   (create_matrix_block, KNPEMIx_solver.py:157) exactly as knp_create lays them out (owned rows; columns in the
   [owned | ghost tail] column layout).  Call once with NULL arrays for the sizes (n_own_loc4 = owned and local dof
   counts of the two subdomains), then with arrays of n_rows + 1, nnz, n_loc[0], n_loc[1] entries. */
int knp_pattern_host(const knp_mesh_desc* mesh, int64_t* n_rows, int64_t* nnz, int32_t* n_own_loc4, int32_t* indptr,
                     int32_t* indices, int32_t* dof_vert_i, int32_t* dof_vert_e);
/* Host-only (no GPU), TEST INFRASTRUCTURE: one assembly of the P2 element path (mesh->degree == 2) on the CPU with the very
   functions its kernels run per thread (csrc/p2.cuh), so that the test tier without a GPU checks the P2 tables and element
   math against the oracle.  mode 0: A values (nnz) and b (n_rows) at time t from (u, gates) as knp_assemble lays them out;
   mode 1: the preconditioner matrix P (nnz_P values; b, gates unused).  With scale_stimulus the stimulus area is taken
   from p->stim_area.  No product call reaches this function: knp_create / knp_assemble need a GPU. */
int knp_p2_emulate_host(const knp_mesh_desc* mesh, const knp_params* p, int32_t n_tags, const knp_tag_models* tags, double t,
                        int32_t mode, const double* u, const double* gates, double* vals, double* b);
/* Host-only (no GPU): the lane-group tables the edge-lane row kernel reads (csrc/topology.cpp): per (owned dof w, slot e)
   at (w << lgG) + e the neighbour node id (adjG, -1 beyond the degree) and the cells around the edge (w, neighbour) written
   as the adjacency slots of their other vertices (hitG: 1 word per entry in 2D, 4 words in 3D, unused bytes 0xFF); per dof
   metaG = {deg | self << 8 | gamma degree << 16, membrane vertex or -1}; node_x = coordinates per restricted local dof
   (intracellular dofs first).  Call once with NULL arrays for lgG / n_work / edge_ok (0: the mesh is served by the scan
   kernel, no tables). */
int knp_edge_tables_host(const knp_mesh_desc* mesh, int32_t* lgG, int64_t* n_work, int32_t* edge_ok, int32_t* adjG,
                         uint32_t* hitG, int32_t* metaG, double* node_x);
/* Host-only: the row blocks the TMA-staged SpMV (the MatMult of ksp.solve, KNPEMIx_solver.py:435) walks over --
   blocks4 = {first row (multiple of 4), rows, 4-aligned first non-zero, staged non-zeros} per block; n_blocks = -1 when a
   group of four rows exceeds the stage capacity (the CSR-vector kernel is used then). */
int knp_rowblocks_host(int32_t n_rows, const int32_t* indptr, int32_t max_blocks, int32_t* blocks4, int32_t* n_blocks);
/* Host-only (no GPU, not thread-safe): builds the smoothed-aggregation hierarchy of a CSR matrix with the setup code the
   preconditioners use (amg_setup.cpp; stands in for hypre's setup inside ksp.setUp, KNPEMIx_solver.py:386-389) and keeps
   it for inspection with knp_amg_host_level; used by the CPU test suite to compare with oracle/amg.py. */
int knp_amg_setup_host(int32_t n, const int32_t* indptr, const int32_t* indices, const double* vals, double theta,
                       int32_t coarse_size, int32_t* n_levels);
int knp_amg_host_level(int32_t level, int64_t* n, int64_t* nnz, int32_t* indptr, int32_t* indices, double* vals);
/* The same hierarchy built ON THE DEVICE (amg_device.cu) -- what knp_pc_setup runs on a single GPU, so that reassemble_P
   (KNPEMIx_solver.py:137-150,405-406) costs a fraction of a second; the level operators are kept for knp_amg_host_level and
   agree with knp_amg_setup_host bit for bit.  Needs a GPU (no CPU fallback); fails on matrices the device form hands to the
   host setup (Dirichlet rows, unsymmetric patterns). */
int knp_amg_setup_device(int32_t n, const int32_t* indptr, const int32_t* indices, const double* vals, double theta,
                         int32_t coarse_size, int32_t device, int32_t* n_levels);
/* 1 when the hierarchy of the last knp_pc_setup was built on the device, 0 when the host setup built it */
int knp_amg_setup_was_on_device(const knp_ctx* ctx);
/* Host-only (no GPU): the ROW-DISTRIBUTED hierarchy setup of multi-GPU runs (amg_dist.cpp; our counterpart of hypre running
   across the MPI ranks, KNPEMIx_solver.py:269-273) on `nranks` SIMULATED ranks: the matrix is split by owner[] (rows with
   owner == r belong to rank r), every rank runs the collective setup in its own thread, and the per-rank pieces are
   assembled into global level operators / prolongators for inspection.  n_levels = distributed levels + the replicated
   one.  knp_amg_dist_sim_level: which = 0 level operator, 1 prolongator to the next level; call with NULL arrays for the
   sizes first.  knp_amg_dist_sim_perm: level-0 numbering of the assembled operators (new index -> input index). */
int knp_amg_dist_sim_host(int32_t nranks, int32_t n, const int32_t* indptr, const int32_t* indices, const double* vals,
                          const int32_t* owner, double theta, int64_t repl_threshold, int32_t* n_levels);
int knp_amg_dist_sim_level(int32_t level, int32_t which, int64_t* n_rows, int64_t* n_cols, int64_t* nnz, double* rho,
                           int32_t* indptr, int32_t* indices, double* vals);
int knp_amg_dist_sim_perm(int32_t* perm0);

/* ---- multi-GPU: halo exchange of ghost columns + all-reduce over NCCL (one rank per GPU) ----
 * Replaces PETSc VecScatter/ghostUpdate and MPI_Allreduce inside KSP (KNPEMI/KNPEMIx_solver.py:435,439,458-468). */
int knp_nccl_unique_id(char* out128);
int knp_dist_init(knp_ctx* ctx, int32_t rank, int32_t nranks, const char* unique_id128,
                  int64_t n_phi_global,                               /* global count of phi_i + phi_e dofs */
                  int32_t n_peers, const int32_t* peers,
                  const int64_t* send_ptr, const int32_t* send_cols,   /* owned columns packed per peer */
                  const int64_t* recv_ptr, const int32_t* recv_cols);  /* ghost columns filled per peer */
/* 1 when ghost exchanges and all-reduces run as the library's own kernels over NVLink / NVSwitch peer memory (CUDA IPC
   mappings set up in knp_dist_init), 0 when they go through NCCL point-to-point / all-reduce (KNP_HALO=nccl or IPC
   unavailable). */
int knp_peer_direct(const knp_ctx* ctx);
int knp_halo_exchange(knp_ctx* ctx, double* x_dev, void* stream);
int knp_allreduce_sum(knp_ctx* ctx, double* buf_dev, int32_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif
