// extern "C" entry points of libknpemi_b200.so (see include/knpemi_b200.h for the reference call sites).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <dlfcn.h>
#include "context.cuh"

using namespace knp;

namespace knp {
int nullspace_remove(knp_ctx* c, double* x, cudaStream_t st);
void pc_graphs_clear(knp_ctx* c);
}

static cudaStream_t pick(knp_ctx* c, void* stream) { return stream ? (cudaStream_t)stream : c->stream; }

#define CTX_GUARD(c)                              \
  do {                                            \
    if (!(c)) {                                   \
      set_error("context is NULL");               \
      return KNP_E_INVALID;                       \
    }                                             \
    cudaError_t e_ = cudaSetDevice((c)->device);  \
    if (e_ != cudaSuccess) {                      \
      set_error("cudaSetDevice(%d): %s", (c)->device, cudaGetErrorString(e_)); \
      return KNP_E_CUDA;                          \
    }                                             \
  } while (0)

// knp_create for P2 elements (knp_mesh_desc::degree == 2, one GPU): the tables of topology_p2.cpp instead of the P1 ones;
// everything behind the assembly (CSR products, Krylov loop, preconditioners, boundary conditions) is element-agnostic.
static int create_p2(std::unique_ptr<knp_ctx>& c, const knp_mesh_desc* mesh, knp_ctx** out) {
  KNP_TRY(build_topology_p2(mesh, c->H));
  HostTopo& H = c->H;
  P2Host& Q = H.p2;
  KNP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev) KNP_CUDA(cudaEventCreate(&ev));
  auto& D = c->p2d;
  KNP_TRY(c->d_node_x.upload(H.node_x));
  KNP_TRY(D.adj_ptr.upload(Q.adj_ptr));
  KNP_TRY(D.gam_ptr.upload(Q.gam_ptr));
  KNP_TRY(D.inc_ptr.upload(Q.inc_ptr));
  KNP_TRY(D.inc_cell.upload(Q.inc_cell));
  KNP_TRY(D.inc_loc.upload(Q.inc_loc));
  KNP_TRY(D.inc_slots.upload(Q.inc_slots));
  KNP_TRY(D.minc_ptr.upload(Q.minc_ptr));
  KNP_TRY(D.minc_facet.upload(Q.minc_facet));
  KNP_TRY(D.minc_loc.upload(Q.minc_loc));
  KNP_TRY(D.minc_own.upload(Q.minc_own));
  KNP_TRY(D.minc_gam.upload(Q.minc_gam));
  KNP_TRY(D.cq_w.upload(Q.cq_w));
  KNP_TRY(D.cq_N.upload(Q.cq_N));
  KNP_TRY(D.cq_dN.upload(Q.cq_dN));
  KNP_TRY(D.fq_N.upload(Q.fq_N));
  KNP_TRY(D.fq_M.upload(Q.fq_M));
  KNP_TRY(c->d_qb.upload(Q.fq_b));
  KNP_TRY(c->d_qw.upload(Q.fq_w));
  KNP_TRY(c->d_mv_node0.upload(H.mv_node[0]));
  KNP_TRY(c->d_mv_node1.upload(H.mv_node[1]));
  KNP_TRY(c->d_mf_mv.upload(H.mf_mv));
  KNP_TRY(c->d_mf_tagidx.upload(H.mf_tagidx));
  KNP_TRY(c->d_mf_area.upload(H.mf_area));
  KNP_TRY(c->d_mf_owned.upload(H.mf_owned));
  {
    std::vector<int32_t> ip(H.indptr), ipP(H.indptr_P);
    for (int k = 0; k < 4; ++k) {   // padding for the 16-byte TMA slices of the streaming SpMV
      ip.push_back(H.indptr.back());
      ipP.push_back(H.indptr_P.back());
    }
    KNP_TRY(c->d_indptr.upload(ip));
    KNP_TRY(c->d_indptr_P.upload(ipP));
  }
  KNP_TRY(c->d_indices.upload(Q.indices));
  KNP_TRY(c->d_indices_P.upload(Q.indices_P));
  {
    std::vector<int32_t> blk;
    c->nblk_A = build_rowblocks(H.indptr.data(), H.L.n_rows, blk);
    if (c->nblk_A > 0) KNP_TRY(c->d_rowblk_A.upload(blk));
    else c->nblk_A = 0;
  }
  for (int s = 0; s < 2; ++s) {
    KNP_TRY(c->d_cell_nodes[s].upload(H.cell_nodes[s]));
    KNP_TRY(c->d_cell_tag[s].upload(H.cell_tag[s]));
    KNP_TRY(c->d_cell_owned[s].upload(H.cell_owned[s]));
  }
  P2View& V = c->p2v;
  V = p2_host_view(H);               // sizes and layout; every pointer is replaced by its device copy
  V.node_x = c->d_node_x.p;
  V.cell_nodes[0] = c->d_cell_nodes[0].p;
  V.cell_nodes[1] = c->d_cell_nodes[1].p;
  V.adj_ptr = D.adj_ptr.p;
  V.gam_ptr = D.gam_ptr.p;
  V.inc_ptr = D.inc_ptr.p;
  V.inc_cell = D.inc_cell.p;
  V.inc_loc = D.inc_loc.p;
  V.inc_slots = D.inc_slots.p;
  V.minc_ptr = D.minc_ptr.p;
  V.minc_facet = D.minc_facet.p;
  V.minc_loc = D.minc_loc.p;
  V.minc_own = D.minc_own.p;
  V.minc_gam = D.minc_gam.p;
  V.indptr = c->d_indptr.p;
  V.indptr_P = c->d_indptr_P.p;
  V.cq_w = D.cq_w.p;
  V.cq_N = D.cq_N.p;
  V.cq_dN = D.cq_dN.p;
  V.fq_b = c->d_qb.p;
  V.fq_w = c->d_qw.p;
  V.fq_N = D.fq_N.p;
  V.fq_M = D.fq_M.p;
  V.mv_node0 = c->d_mv_node0.p;
  V.mv_node1 = c->d_mv_node1.p;
  V.mf_mv = c->d_mf_mv.p;
  V.mf_tagidx = c->d_mf_tagidx.p;
  V.mf_area = c->d_mf_area.p;
  DevTopo& T = c->T;                 // the element-agnostic part of the P1 view (gate kernel, layout, patterns)
  T.gdim = H.gdim;
  T.L = H.L;
  T.n_work = H.n_work;
  T.n_mv = H.n_mv;
  T.n_mf = H.n_mf;
  T.nq = mesh->n_quad;
  T.node_x = c->d_node_x.p;
  T.mv_node0 = c->d_mv_node0.p;
  T.mv_node1 = c->d_mv_node1.p;
  T.mf_mv = c->d_mf_mv.p;
  T.mf_tagidx = c->d_mf_tagidx.p;
  T.mf_area = c->d_mf_area.p;
  T.indptr = c->d_indptr.p;
  T.indptr_P = c->d_indptr_P.p;
  T.qb = c->d_qb.p;
  T.qw = c->d_qw.p;
  T.max_inc = H.max_inc;
  T.p2 = &c->p2v;
  KNP_TRY(c->u.alloc(T.L.n_cols));
  KNP_TRY(c->gates.alloc((size_t)3 * T.n_mv));
  KNP_TRY(c->A_vals.alloc(H.nnz));
  KNP_TRY(c->P_vals.alloc(H.nnz_P));
  KNP_TRY(c->b.alloc(T.L.n_rows));
  KNP_TRY(c->fe.alloc((size_t)p2_facet_ncomp(H.gdim) * (T.n_mf > 0 ? T.n_mf : 1)));
  KNP_CUDA(cudaMemsetAsync(c->u.p, 0, (size_t)T.L.n_cols * sizeof(double), c->stream));
  KNP_TRY(c->fpartial.alloc(1024));
  KNP_TRY(c->fout.alloc(8));
  KNP_TRY(c->ftags.alloc(4096));
  c->n_phi_global = T.L.n_own[0] + T.L.n_own[1];
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  *out = c.release();
  return KNP_OK;
}

extern "C" {

const char* knp_last_error(void) { return knp::last_error(); }
int knp_version(void) { return 100; }
int64_t knp_launch_count(void) { return (int64_t)knp::g_kernel_launches; }

int knp_create(knp_ctx** out, const knp_mesh_desc* mesh, int device) {
  if (!out) {
    set_error("out is NULL");
    return KNP_E_INVALID;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s); libknpemi_b200 has no CPU fallback", cudaGetErrorString(e));
    return KNP_E_CUDA;
  }
  KNP_CHECK(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
  KNP_CUDA(cudaSetDevice(device));
  std::unique_ptr<knp_ctx> c(new knp_ctx());
  c->device = device;
  if (mesh && mesh->degree == 2) return create_p2(c, mesh, out);
  KNP_CHECK(!mesh || mesh->degree == 0 || mesh->degree == 1, "element degree %d is not supported (1 or 2)", mesh->degree);
  KNP_TRY(build_topology(mesh, c->H));
  HostTopo& H = c->H;
  KNP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev) KNP_CUDA(cudaEventCreate(&ev));
  // upload
  KNP_TRY(c->d_node_x.upload(H.node_x));
  KNP_TRY(c->d_adj_ptr.upload(H.adj_ptr));
  KNP_TRY(c->d_adj_idx.upload(H.adj_idx));
  KNP_TRY(c->d_inc_ptr.upload(H.inc_ptr));
  KNP_TRY(c->d_inc_slots.upload(H.inc_slots));
  KNP_TRY(c->d_self_slot.upload(H.self_slot));
  KNP_TRY(c->d_mv_of_node.upload(H.mv_of_node));
  KNP_TRY(c->d_gpre.upload(H.gpre));
  // lane-group tables of the edge-lane row kernel (default); KNP_ROWS=scan keeps the scan kernel for comparison, and a
  // mesh whose edge rings or degrees exceed the tables (HostTopo::edge_ok == 0) is served by it as well
  const bool use_edge = H.edge_ok && !(getenv("KNP_ROWS") && std::string(getenv("KNP_ROWS")) == "scan");
  if (use_edge) {
    KNP_TRY(c->d_adjG.upload(H.adjG));
    KNP_TRY(c->d_hitG.upload(H.hitG));
    KNP_TRY(c->d_metaG.upload(H.metaG));
  }
  KNP_TRY(c->d_mv_node0.upload(H.mv_node[0]));
  KNP_TRY(c->d_mv_node1.upload(H.mv_node[1]));
  KNP_TRY(c->d_mf_mv.upload(H.mf_mv));
  KNP_TRY(c->d_mf_tagidx.upload(H.mf_tagidx));
  KNP_TRY(c->d_mf_area.upload(H.mf_area));
  KNP_TRY(c->d_mf_owned.upload(H.mf_owned));
  KNP_TRY(c->d_gam_ptr.upload(H.gam_ptr));
  KNP_TRY(c->d_gam_mv.upload(H.gam_mv));
  KNP_TRY(c->d_minc_ptr.upload(H.minc_ptr));
  KNP_TRY(c->d_minc.upload(H.minc));
  {
    std::vector<int32_t> ip(H.indptr), ipP(H.indptr_P);
    for (int k = 0; k < 4; ++k) {   // padding for the 16-byte TMA slices of the streaming SpMV
      ip.push_back(H.indptr.back());
      ipP.push_back(H.indptr_P.back());
    }
    KNP_TRY(c->d_indptr.upload(ip));
    KNP_TRY(c->d_indptr_P.upload(ipP));
  }
  {
    std::vector<int32_t> blk;
    c->nblk_A = build_rowblocks(H.indptr.data(), H.L.n_rows, blk);
    if (c->nblk_A > 0) KNP_TRY(c->d_rowblk_A.upload(blk));
    else c->nblk_A = 0;
  }
  {
    std::vector<double> qb(mesh->quad_bary, mesh->quad_bary + (size_t)mesh->n_quad * mesh->gdim);
    std::vector<double> qw(mesh->quad_w, mesh->quad_w + mesh->n_quad);
    KNP_TRY(c->d_qb.upload(qb));
    KNP_TRY(c->d_qw.upload(qw));
  }
  for (int s = 0; s < 2; ++s) {
    KNP_TRY(c->d_cell_nodes[s].upload(H.cell_nodes[s]));
    KNP_TRY(c->d_cell_tag[s].upload(H.cell_tag[s]));
    KNP_TRY(c->d_cell_owned[s].upload(H.cell_owned[s]));
  }
  DevTopo& T = c->T;
  T.gdim = H.gdim;
  T.L = H.L;
  T.n_work = H.n_work;
  T.n_mv = H.n_mv;
  T.n_mf = H.n_mf;
  T.nq = mesh->n_quad;
  T.node_x = c->d_node_x.p;
  T.adj_ptr = c->d_adj_ptr.p;
  T.adj_idx = c->d_adj_idx.p;
  T.inc_ptr = c->d_inc_ptr.p;
  T.inc_slots = c->d_inc_slots.p;
  T.self_slot = c->d_self_slot.p;
  T.mv_of_node = c->d_mv_of_node.p;
  T.mv_node0 = c->d_mv_node0.p;
  T.mv_node1 = c->d_mv_node1.p;
  T.mf_mv = c->d_mf_mv.p;
  T.mf_tagidx = c->d_mf_tagidx.p;
  T.mf_area = c->d_mf_area.p;
  T.gam_ptr = c->d_gam_ptr.p;
  T.gam_mv = c->d_gam_mv.p;
  T.minc_ptr = c->d_minc_ptr.p;
  T.minc = c->d_minc.p;
  T.indptr = c->d_indptr.p;
  T.indptr_P = c->d_indptr_P.p;
  T.qb = c->d_qb.p;
  T.qw = c->d_qw.p;
  T.gpre = c->d_gpre.p;
  T.max_inc = H.max_inc;
  T.lgG = H.lgG;
  T.adjG = use_edge ? c->d_adjG.p : nullptr;
  T.hitG = use_edge ? c->d_hitG.p : nullptr;
  T.metaG = use_edge ? reinterpret_cast<const int2*>(c->d_metaG.p) : nullptr;
  // CSR column indices on the device
  KNP_TRY(c->d_indices.alloc(H.nnz));
  KNP_TRY(c->d_indices_P.alloc(H.nnz_P));
  KNP_TRY(launch_csr_indices(T, 0, c->d_indices.p, c->stream));
  KNP_TRY(launch_csr_indices(T, 1, c->d_indices_P.p, c->stream));
  // state and system storage
  KNP_TRY(c->u.alloc(T.L.n_cols));
  KNP_TRY(c->gates.alloc((size_t)3 * T.n_mv));
  KNP_TRY(c->A_vals.alloc(H.nnz));
  KNP_TRY(c->P_vals.alloc(H.nnz_P));
  KNP_TRY(c->b.alloc(T.L.n_rows));
  KNP_TRY(c->fe.alloc((size_t)facet_ncomp(H.gdim) * (T.n_mf > 0 ? T.n_mf : 1)));
  KNP_CUDA(cudaMemsetAsync(c->u.p, 0, (size_t)T.L.n_cols * sizeof(double), c->stream));
  KNP_TRY(c->fpartial.alloc(1024));
  KNP_TRY(c->fout.alloc(8));
  KNP_TRY(c->ftags.alloc(4096));
  c->n_phi_global = T.L.n_own[0] + T.L.n_own[1];
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  // the big host-side tables are only needed on the device from here on
  std::vector<uint32_t>().swap(H.inc_slots);
  std::vector<uint32_t>().swap(H.minc);
  std::vector<int32_t>().swap(H.adj_idx);
  std::vector<int32_t>().swap(H.adjG);
  std::vector<uint32_t>().swap(H.hitG);
  std::vector<int32_t>().swap(H.metaG);
  *out = c.release();
  return KNP_OK;
}

int knp_destroy(knp_ctx* c) {
  if (!c) return KNP_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  pc_graphs_clear(c);
  if (c->comm) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (h) {
      auto fn = (int (*)(ncclComm*))dlsym(h, "ncclCommDestroy");
      if (fn) fn(c->comm);
    }
  }
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return KNP_OK;
}

int knp_get_sizes(const knp_ctx* c, knp_sizes* o) {
  KNP_CHECK(c && o, "NULL argument");
  const Layout& L = c->T.L;
  o->n_rows = L.n_rows;
  o->n_cols = L.n_cols;
  o->nnz = c->H.nnz;
  o->nnz_P = c->H.nnz_P;
  for (int s = 0; s < 2; ++s) {
    o->n_own[s] = L.n_own[s];
    o->n_loc[s] = L.n_loc[s];
    o->n_cells[s] = (int64_t)c->H.cell_tag[s].size();
  }
  o->n_mverts = c->T.n_mv;
  o->n_mfacets = c->T.n_mf;
  o->max_deg = c->H.max_deg;
  o->max_gdeg = c->H.max_gdeg;
  return KNP_OK;
}

int knp_csr_dev(const knp_ctx* c, const int32_t** indptr, const int32_t** indices) {
  KNP_CHECK(c, "context is NULL");
  if (indptr) *indptr = c->d_indptr.p;
  if (indices) *indices = c->d_indices.p;
  return KNP_OK;
}

int knp_csr_host(const knp_ctx* c, int32_t* indptr, int32_t* indices) {
  KNP_CHECK(c, "context is NULL");
  KNP_CUDA(cudaSetDevice(c->device));
  if (indptr) memcpy(indptr, c->H.indptr.data(), c->H.indptr.size() * sizeof(int32_t));
  if (indices) KNP_CUDA(cudaMemcpy(indices, c->d_indices.p, c->H.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return KNP_OK;
}

int knp_csr_P_host(const knp_ctx* c, int32_t* indptr, int32_t* indices) {
  KNP_CHECK(c, "context is NULL");
  KNP_CUDA(cudaSetDevice(c->device));
  if (indptr) memcpy(indptr, c->H.indptr_P.data(), c->H.indptr_P.size() * sizeof(int32_t));
  if (indices) KNP_CUDA(cudaMemcpy(indices, c->d_indices_P.p, c->H.nnz_P * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return KNP_OK;
}

int knp_dofmap_host(const knp_ctx* c, int32_t* vi, int32_t* ve) {
  KNP_CHECK(c, "context is NULL");
  if (vi) memcpy(vi, c->H.node_vert[0].data(), c->H.node_vert[0].size() * sizeof(int32_t));
  if (ve) memcpy(ve, c->H.node_vert[1].data(), c->H.node_vert[1].size() * sizeof(int32_t));
  return KNP_OK;
}

int knp_mverts_host(const knp_ctx* c, int32_t* verts) {
  KNP_CHECK(c && verts, "NULL argument");
  memcpy(verts, c->H.mv_vert.data(), c->H.mv_vert.size() * sizeof(int32_t));
  return KNP_OK;
}

int knp_stimulus_area_local(knp_ctx* c, double* out) {
  KNP_CHECK(c && out, "NULL argument");
  KNP_CHECK(c->params_set, "knp_set_params must be called first");
  // setup-time integral of the stimulus mask over owned stimulated facets (host, fixed order)
  const HostTopo& H = c->H;
  const int d = H.gdim;
  const int fstride = H.degree == 2 ? H.p2.nt : d;          // nodes per membrane facet (its d vertices come first)
  std::vector<double> qb(c->d_qb.n), qw(c->d_qw.n);
  KNP_CUDA(cudaSetDevice(c->device));
  KNP_CUDA(cudaMemcpy(qb.data(), c->d_qb.p, qb.size() * sizeof(double), cudaMemcpyDeviceToHost));
  KNP_CUDA(cudaMemcpy(qw.data(), c->d_qw.p, qw.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const knp_params& p = c->params.p;
  double acc = 0.0;
  for (int f = 0; f < H.n_mf; ++f) {
    if (!H.mf_owned[f] || !c->params.tag_stim[H.mf_tagidx[f]]) continue;
    for (size_t q = 0; q < qw.size(); ++q) {
      double mask = 1.0;
      for (int i = 0; i < 3 && p.stim_dir[i] >= 0; ++i) {
        double xq = 0.0;
        for (int a = 0; a < d; ++a) {
          const int node = H.mv_node[0][H.mf_mv[(size_t)f * fstride + a]];
          xq += qb[q * d + a] * H.node_x[(size_t)node * d + p.stim_dir[i]];
        }
        mask *= (xq > p.stim_lo[i] && xq < p.stim_hi[i]) ? 1.0 : 0.0;
      }
      acc += H.mf_area[f] * qw[q] * mask;
    }
  }
  *out = acc;
  return KNP_OK;
}

int knp_set_params(knp_ctx* c, const knp_params* p, int32_t n_tags, const knp_tag_models* tags) {
  CTX_GUARD(c);
  KNP_CHECK(p, "params is NULL");
  KNP_CHECK(p->dt > 0 && p->F > 0 && p->R > 0 && p->T > 0, "dt, F, R, T must be positive");
  KNP_CHECK(p->ode_substeps >= 1, "ode_substeps must be >= 1");
  Params& P = c->params;
  P.p = *p;
  P.psi = p->R * p->T / p->F;
  P.n_tags = (int)c->H.mtags.size();
  P.any_hh = false;
  P.tag_models.clear();
  P.tag_stim.clear();
  std::vector<uint32_t> tm(P.n_tags > 0 ? P.n_tags : 1, 0u);
  std::vector<int32_t> ts(P.n_tags > 0 ? P.n_tags : 1, 0);
  for (int i = 0; i < P.n_tags; ++i) {
    bool found = false;
    for (int j = 0; j < n_tags; ++j)
      if (tags[j].tag == c->H.mtags[i]) {
        tm[i] = tags[j].models;
        ts[i] = tags[j].stimulated ? 1 : 0;
        found = true;
      }
    KNP_CHECK(found, "membrane tag %d present in the mesh has no ionic model (Mismatch between membrane tags and ionic models tags)",
              c->H.mtags[i]);
    P.tag_models.push_back(tm[i]);
    P.tag_stim.push_back(ts[i]);
    if (tm[i] & KNP_MODEL_HH) P.any_hh = true;
  }
  for (int j = 0; j < n_tags; ++j)
    if (tags[j].models & KNP_MODEL_HH) P.any_hh = true;
  KNP_TRY(c->d_tag_models.upload(tm));
  KNP_TRY(c->d_tag_stim.upload(ts));
  KParams& K = c->kp;
  K.dt = p->dt;
  K.F = p->F;
  K.C_M = p->C_M;
  K.psi = P.psi;
  K.phi_rest = p->phi_rest;
  for (int k = 0; k < 3; ++k) {
    K.z[k] = p->z[k];
    K.D[k] = p->D[k];
    K.g_leak[k] = p->g_leak[k];
    K.g_leak_g[k] = p->g_leak_g[k];
  }
  K.g_Na_bar = p->g_Na_bar;
  K.g_K_bar = p->g_K_bar;
  for (int i = 0; i < 3; ++i) {
    KNP_CHECK(p->stim_dir[i] < c->H.gdim, "stimulus_region direction %d on a %dD mesh", p->stim_dir[i], c->H.gdim);
    K.stim_dir[i] = p->stim_dir[i];
    K.stim_lo[i] = p->stim_lo[i];
    K.stim_hi[i] = p->stim_hi[i];
  }
  K.K_e_init = p->K_e_init;
  K.K_i_g_init = p->K_i_g_init;
  K.ode_substeps = p->ode_substeps;
  K.rush_larsen = p->rush_larsen;
  c->params_set = true;
  if (p->stim_area > 0.0) {
    P.stim_area = p->stim_area;
  } else {
    double a = 0.0;
    KNP_TRY(knp_stimulus_area_local(c, &a));
    P.stim_area = a;
  }
  return KNP_OK;
}

int knp_set_state(knp_ctx* c, const double* u_host, const double* gates_host) {
  CTX_GUARD(c);
  if (u_host)
    KNP_CUDA(cudaMemcpyAsync(c->u.p, u_host, (size_t)c->T.L.n_cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (gates_host && c->T.n_mv)
    KNP_CUDA(cudaMemcpyAsync(c->gates.p, gates_host, (size_t)3 * c->T.n_mv * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  return KNP_OK;
}

int knp_get_state(knp_ctx* c, double* u_host, double* gates_host) {
  CTX_GUARD(c);
  if (u_host)
    KNP_CUDA(cudaMemcpyAsync(u_host, c->u.p, (size_t)c->T.L.n_cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (gates_host && c->T.n_mv)
    KNP_CUDA(cudaMemcpyAsync(gates_host, c->gates.p, (size_t)3 * c->T.n_mv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  return KNP_OK;
}

int knp_state_dev(knp_ctx* c, double** u_dev, double** gates_dev) {
  KNP_CHECK(c, "context is NULL");
  if (u_dev) *u_dev = c->u.p;
  if (gates_dev) *gates_dev = c->gates.p;
  return KNP_OK;
}

int knp_phi_m_host(knp_ctx* c, double* out) {
  CTX_GUARD(c);
  KNP_CHECK(out, "NULL argument");
  std::vector<double> u(c->T.L.n_cols);
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  KNP_CUDA(cudaMemcpy(u.data(), c->u.p, u.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int g = 0; g < c->T.n_mv; ++g)
    out[g] = u[c->T.L.col(0, 3, c->H.mv_node[0][g])] - u[c->T.L.col(1, 3, c->H.mv_node[1][g])];
  return KNP_OK;
}

int knp_gate_step(knp_ctx* c, void* stream) {
  CTX_GUARD(c);
  KNP_CHECK(c->params_set, "knp_set_params must be called first");
  return launch_gate(c->T, c->kp, c->u.p, c->gates.p, pick(c, stream));
}

int knp_assemble(knp_ctx* c, double t, double* A_vals, double* b, void* stream) {
  CTX_GUARD(c);
  KNP_CHECK(c->params_set, "knp_set_params must be called first");
  cudaStream_t st = pick(c, stream);
  const knp_params& p = c->params.p;
  // HodgkinHuxley.update_t_mod (KNPEMIx_ionic_model.py:673-674) and the stimulus prefactor (:552,589,600)
  const double t_mod = std::fmod(t + 1e-12, p.T_stim);
  double stim_fac = p.g_syn_bar * std::exp(-t_mod / p.a_syn);
  if (p.scale_stimulus) stim_fac *= 1.0 / c->params.stim_area;
  KNP_TRY(launch_facets(c->T, c->kp, c->d_tag_models.p, c->d_tag_stim.p, c->u.p, c->gates.p, stim_fac, c->fe.p, st));
  KNP_CUDA(cudaEventRecord(c->ev[2], st));
  KNP_TRY(launch_rows(c->T, c->kp, 0, c->u.p, c->fe.p, A_vals ? A_vals : c->A_vals.p, b ? b : c->b.p, c->H.max_deg, c->H.max_gdeg, st));
  if (c->n_src > 0) KNP_TRY(launch_add_sparse(c->n_src, c->src_rows.p, c->src_vals.p, b ? b : c->b.p, st));   // :613-614
  if (c->n_bc > 0)      // bcs = p.bcs of assemble_matrix_block / assemble_vector_block (:113-116)
    KNP_TRY(launch_bc_apply(c->n_bc_rows_A, c->bc_rows_A.p, c->d_indptr.p, c->d_indices.p, A_vals ? A_vals : c->A_vals.p,
                            b ? b : c->b.p, c->n_bc, c->bc_cols.p, c->bc_vals.p, 1.0, st));
  return KNP_OK;
}

static int bc_touched_rows(knp_ctx* c, const int32_t* indptr, const int32_t* indices, const uint8_t* flag,
                           knp::DevBuf<int32_t>& out, int& n_out) {
  const int n = c->T.L.n_rows;
  knp::DevBuf<uint8_t> touched;
  KNP_TRY(touched.alloc(n));
  KNP_TRY(launch_bc_touch(n, indptr, indices, flag, touched.p, c->stream));
  std::vector<uint8_t> h(n);
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  if (n) KNP_CUDA(cudaMemcpy(h.data(), touched.p, n, cudaMemcpyDeviceToHost));
  std::vector<int32_t> list;
  for (int i = 0; i < n; ++i)
    if (h[i]) list.push_back(i);
  n_out = (int)list.size();
  return out.upload(list);
}

int knp_set_dirichlet(knp_ctx* c, int32_t n, const int32_t* cols, const double* vals) {
  CTX_GUARD(c);
  KNP_CHECK(n >= 0 && (n == 0 || (cols && vals)), "knp_set_dirichlet: invalid arguments");
  const Layout& L = c->T.L;
  std::vector<std::pair<int32_t, double>> bc(n);
  for (int i = 0; i < n; ++i) {
    KNP_CHECK(cols[i] >= 0 && cols[i] < L.n_cols, "knp_set_dirichlet: column out of range");
    bc[i] = {cols[i], vals[i]};
  }
  std::sort(bc.begin(), bc.end());
  for (int i = 1; i < n; ++i) KNP_CHECK(bc[i].first != bc[i - 1].first, "knp_set_dirichlet: duplicate column");
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  c->n_bc = c->n_bc_rows_A = c->n_bc_rows_P = 0;
  c->h_bc_rows.clear();
  if (n == 0) return KNP_OK;
  std::vector<int32_t> hc(n);
  std::vector<double> hv(n);
  for (int i = 0; i < n; ++i) {
    hc[i] = bc[i].first;
    hv[i] = bc[i].second;
    if (hc[i] < L.n_rows) c->h_bc_rows.push_back(hc[i]);
  }
  KNP_TRY(c->bc_cols.upload(hc));
  KNP_TRY(c->bc_vals.upload(hv));
  knp::DevBuf<uint8_t> flag;
  KNP_TRY(flag.alloc(L.n_cols));
  KNP_CUDA(cudaMemsetAsync(flag.p, 0, L.n_cols, c->stream));
  KNP_TRY(launch_bc_flags(n, c->bc_cols.p, flag.p, c->stream));
  KNP_TRY(bc_touched_rows(c, c->d_indptr.p, c->d_indices.p, flag.p, c->bc_rows_A, c->n_bc_rows_A));
  KNP_TRY(bc_touched_rows(c, c->d_indptr_P.p, c->d_indices_P.p, flag.p, c->bc_rows_P, c->n_bc_rows_P));
  c->n_bc = n;
  c->P_assembled = false;          // a preconditioner matrix assembled before carries no boundary rows
  return KNP_OK;
}

int knp_set_source(knp_ctx* c, int32_t n, const int32_t* rows, const double* vals) {
  CTX_GUARD(c);
  KNP_CHECK(n >= 0 && (n == 0 || (rows && vals)), "knp_set_source: invalid arguments");
  for (int i = 0; i < n; ++i) KNP_CHECK(rows[i] >= 0 && rows[i] < c->T.L.n_rows, "knp_set_source: row out of range");
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  c->n_src = 0;
  if (n > 0) {
    KNP_TRY(c->src_rows.upload(std::vector<int32_t>(rows, rows + n)));
    KNP_TRY(c->src_vals.upload(std::vector<double>(vals, vals + n)));
    c->n_src = n;
  }
  return KNP_OK;
}

int knp_assemble_P(knp_ctx* c, double* P_vals, void* stream) {
  CTX_GUARD(c);
  KNP_CHECK(c->params_set, "knp_set_params must be called first");
  KNP_TRY(launch_rows(c->T, c->kp, 1, c->u.p, c->fe.p, P_vals ? P_vals : c->P_vals.p, nullptr, c->H.max_deg, c->H.max_gdeg, pick(c, stream)));
  if (c->n_bc > 0)                 // assemble_matrix_block(p.P, bcs = p.bcs) (KNPEMIx_solver.py:125-126)
    KNP_TRY(launch_bc_apply(c->n_bc_rows_P, c->bc_rows_P.p, c->d_indptr_P.p, c->d_indices_P.p, P_vals ? P_vals : c->P_vals.p,
                            nullptr, c->n_bc, c->bc_cols.p, c->bc_vals.p, 1.0, pick(c, stream)));
  if (!P_vals) c->P_assembled = true;
  return KNP_OK;
}

int knp_values_dev(knp_ctx* c, double** A_vals, double** b, double** P_vals, double** x) {
  KNP_CHECK(c, "context is NULL");
  if (A_vals) *A_vals = c->A_vals.p;
  if (b) *b = c->b.p;
  if (P_vals) *P_vals = c->P_vals.p;
  if (x) *x = c->u.p;
  return KNP_OK;
}

int knp_spmv(knp_ctx* c, const double* A_vals, const double* x, double* y, void* stream) {
  CTX_GUARD(c);
  const CsrView A{c->T.L.n_rows, c->H.nnz, c->d_indptr.p, c->d_indices.p, A_vals ? A_vals : c->A_vals.p,
                  c->d_rowblk_A.p, c->nblk_A};
  return spmv(A, x, y, EPI_SET, nullptr, nullptr, 0.0, pick(c, stream));
}

int knp_pc_setup(knp_ctx* c, const knp_solve_opts* o) {
  CTX_GUARD(c);
  KNP_CHECK(o, "options are NULL");
  return pc_setup(c, o);
}

int knp_pc_apply(knp_ctx* c, const double* r, double* z, void* stream) {
  CTX_GUARD(c);
  KNP_CHECK(c->pc_kind >= 0, "knp_pc_setup must be called first");
  return pc_apply(c, r, z, pick(c, stream));
}

int knp_pc_bytes(const knp_ctx* c, double* bytes) {
  KNP_CHECK(c && bytes, "NULL argument");
  KNP_CHECK(c->pc_kind >= 0, "knp_pc_setup must be called first");
  *bytes = pc_bytes(c);
  return KNP_OK;
}

int knp_solve(knp_ctx* c, const double* A_vals, const double* b, double* x, const knp_solve_opts* o,
              knp_solve_info* info, void* stream) {
  CTX_GUARD(c);
  KNP_CHECK(o && info, "NULL argument");
  KNP_CHECK(c->pc_kind == o->pc, "knp_pc_setup must be called with the same preconditioner kind before knp_solve");
  return krylov_solve(c, A_vals ? A_vals : c->A_vals.p, b ? b : c->b.p, x ? x : c->u.p, o, info, pick(c, stream));
}

int knp_set_time(knp_ctx* c, double t, int32_t step_index) {
  KNP_CHECK(c, "context is NULL");
  c->t = t;
  c->step_index = step_index;
  return KNP_OK;
}
int knp_get_time(const knp_ctx* c, double* t, int32_t* step_index) {
  KNP_CHECK(c, "context is NULL");
  if (t) *t = c->t;
  if (step_index) *step_index = c->step_index;
  return KNP_OK;
}

int knp_step(knp_ctx* c, const knp_solve_opts* o, knp_solve_info* info, void* stream) {
  CTX_GUARD(c);
  KNP_CHECK(o && info, "NULL argument");
  KNP_CHECK(c->params_set, "knp_set_params must be called first");
  KNP_CHECK(c->pc_kind == o->pc, "knp_pc_setup must be called with the same preconditioner kind before knp_step");
  cudaStream_t st = pick(c, stream);
  KNP_TRY(ensure_workspace(c, o->restart > 0 ? o->restart : 30));
  c->t += c->params.p.dt;                      // KNPEMIx_solver.py:368
  c->step_index += 1;
  KNP_CUDA(cudaEventRecord(c->ev[0], st));
  if (c->params.any_hh) KNP_TRY(launch_gate(c->T, c->kp, c->u.p, c->gates.p, st));   // :395-399
  KNP_CUDA(cudaEventRecord(c->ev[1], st));
  KNP_TRY(knp_assemble(c, c->t, nullptr, nullptr, st));                              // :402-403 (records ev[2])
  KNP_CUDA(cudaEventRecord(c->ev[3], st));
  if (c->step_index == 1 && o->project_nullspace) KNP_TRY(nullspace_remove(c, c->b.p, st));   // :415-419,333
  int rc = krylov_solve(c, c->A_vals.p, c->b.p, c->u.p, o, info, st);                 // :435 ; u <- x (:451-468)
  KNP_CUDA(cudaEventRecord(c->ev[4], st));
  KNP_CUDA(cudaEventSynchronize(c->ev[4]));
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->last_ms[0] = ms;
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->last_ms[1] = ms;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->last_ms[2] = ms;
  cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]); c->last_ms[3] = ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[4]); c->last_ms[4] = ms;
  return rc;
}

int knp_step_host(knp_ctx* c, double* u_host, double* gates_host, const knp_solve_opts* o, knp_solve_info* info) {
  CTX_GUARD(c);
  KNP_CHECK(u_host, "u_host is NULL");
  cudaStream_t st = c->stream;
  KNP_CUDA(cudaMemcpyAsync(c->u.p, u_host, (size_t)c->T.L.n_cols * sizeof(double), cudaMemcpyHostToDevice, st));
  if (gates_host && c->T.n_mv)
    KNP_CUDA(cudaMemcpyAsync(c->gates.p, gates_host, (size_t)3 * c->T.n_mv * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = knp_step(c, o, info, st);
  if (rc != KNP_OK) return rc;
  KNP_CUDA(cudaMemcpyAsync(u_host, c->u.p, (size_t)c->T.L.n_cols * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (gates_host && c->T.n_mv)
    KNP_CUDA(cudaMemcpyAsync(gates_host, c->gates.p, (size_t)3 * c->T.n_mv * sizeof(double), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

int knp_last_timings(const knp_ctx* c, double* ms5) {
  KNP_CHECK(c && ms5, "NULL argument");
  for (int i = 0; i < 5; ++i) ms5[i] = c->last_ms[i];
  return KNP_OK;
}

static int cell_functional(knp_ctx* c, int32_t s, int32_t field, int32_t power, int32_t n_tags, const int32_t* tags,
                           double* out) {
  CTX_GUARD(c);
  KNP_CHECK(out && tags && n_tags > 0 && n_tags <= 4096, "bad tag list");
  KNP_CHECK((s == 0 || s == 1) && field >= 0 && field < 4 && power >= 0 && power <= 2, "bad subdomain/field/power");
  const int nc = (int)c->H.cell_tag[s].size();
  cudaStream_t st = c->stream;
  KNP_CUDA(cudaMemcpyAsync(c->ftags.p, tags, n_tags * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const int nb = 592;
  if (c->H.degree == 2)
    KNP_TRY(launch_l2_cells_p2(c->p2v, s, field, power, nc, c->d_cell_tag[s].p, c->d_cell_owned[s].p, c->ftags.p, n_tags,
                               c->u.p, c->fpartial.p, nb, st));
  else
    KNP_TRY(launch_l2_cells(c->T.gdim, c->T.L, s, field, power, nc, c->d_cell_nodes[s].p, c->d_cell_tag[s].p,
                            c->d_cell_owned[s].p, c->d_node_x.p, s ? c->T.L.n_loc[0] : 0, c->ftags.p, n_tags, c->u.p,
                            c->fpartial.p, nb, st));
  KNP_TRY(launch_reduce_partials(c->fpartial.p, nb, c->fout.p, st));
  KNP_CUDA(cudaMemcpyAsync(out, c->fout.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

int knp_l2_norm_sq(knp_ctx* c, int32_t s, int32_t field, int32_t n_tags, const int32_t* tags, double* out) {
  return cell_functional(c, s, field, 2, n_tags, tags, out);
}

int knp_integral(knp_ctx* c, int32_t s, int32_t field, int32_t power, int32_t n_tags, const int32_t* tags, double* out) {
  return cell_functional(c, s, field, power, n_tags, tags, out);
}

int knp_stimulus_current(knp_ctx* c, double t, double* out) {
  CTX_GUARD(c);
  KNP_CHECK(out, "NULL argument");
  KNP_CHECK(c->params_set, "knp_set_params must be called first");
  *out = 0.0;
  if (c->T.n_mf == 0 || !c->params.any_hh) return KNP_OK;
  const knp_params& p = c->params.p;
  const double t_mod = std::fmod(t + 1e-12, p.T_stim);            // HodgkinHuxley.update_t_mod, KNPEMIx_ionic_model.py:673-674
  double stim_fac = p.g_syn_bar * std::exp(-t_mod / p.a_syn);
  if (p.scale_stimulus) stim_fac *= 1.0 / c->params.stim_area;
  cudaStream_t st = c->stream;
  const int nb = 592;
  if (c->H.degree == 2)
    KNP_TRY(launch_stim_current_p2(c->p2v, c->kp, c->d_tag_stim.p, c->d_mf_owned.p, c->u.p, stim_fac, c->fpartial.p, nb, st));
  else
    KNP_TRY(launch_stim_current(c->T, c->kp, c->d_tag_stim.p, c->d_mf_owned.p, c->u.p, stim_fac, c->fpartial.p, nb, st));
  KNP_TRY(launch_reduce_partials(c->fpartial.p, nb, c->fout.p, st));
  KNP_CUDA(cudaMemcpyAsync(out, c->fout.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

int knp_probe_setup(knp_ctx* c, int32_t n_out, const int32_t* ptr, const int32_t* cols, const double* weights) {
  CTX_GUARD(c);
  KNP_CHECK(n_out >= 0 && (n_out == 0 || (ptr && cols && weights)), "bad probe tables");
  c->n_probe = 0;
  if (n_out == 0) return KNP_OK;
  const int nt = ptr[n_out];
  for (int t = 0; t < nt; ++t) KNP_CHECK(cols[t] >= 0 && cols[t] < c->T.L.n_cols, "probe column %d out of range", cols[t]);
  KNP_TRY(c->probe_ptr.upload(std::vector<int32_t>(ptr, ptr + n_out + 1)));
  KNP_TRY(c->probe_col.upload(std::vector<int32_t>(cols, cols + nt)));
  KNP_TRY(c->probe_w.upload(std::vector<double>(weights, weights + nt)));
  KNP_TRY(c->probe_out.alloc(n_out));
  c->n_probe = n_out;
  return KNP_OK;
}

int knp_probe_eval(knp_ctx* c, double* out_host) {
  CTX_GUARD(c);
  KNP_CHECK(out_host || c->n_probe == 0, "NULL argument");
  if (c->n_probe == 0) return KNP_OK;
  cudaStream_t st = c->stream;
  KNP_TRY(launch_probe(c->n_probe, c->probe_ptr.p, c->probe_col.p, c->probe_w.p, c->u.p, c->probe_out.p, st));
  KNP_CUDA(cudaMemcpyAsync(out_host, c->probe_out.p, c->n_probe * sizeof(double), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

int knp_membrane_area(const knp_ctx* c, int32_t tag, double* out) {
  KNP_CHECK(c && out, "NULL argument");
  const HostTopo& H = c->H;
  double acc = 0.0;
  for (int f = 0; f < H.n_mf; ++f)
    if (H.mf_owned[f] && H.mtags[H.mf_tagidx[f]] == tag) acc += H.mf_area[f];
  *out = acc;
  return KNP_OK;
}

int knp_copy(knp_ctx* c, void* dst, const void* src, int64_t nbytes, int32_t kind) {
  CTX_GUARD(c);
  KNP_CHECK(dst && src && nbytes >= 0 && kind >= 1 && kind <= 3, "bad knp_copy arguments");
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  KNP_CUDA(cudaMemcpyAsync(dst, src, (size_t)nbytes, k, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  return KNP_OK;
}

// Hierarchy inspection.  Levels are numbered over the hierarchies in use: pc kind 2 has one (on P); pc kind 3 has the
// ion-block hierarchy first and the potential-block hierarchy after it (knp_amg_part_levels tells where it starts).
static const CsrHost* amg_level(const knp_ctx* c, int level) {
  if (!c || level < 0) return nullptr;
  const Amg* parts[3] = {c->amg.get(), c->amg_c.get(), c->amg_p.get()};
  for (const Amg* a : parts) {
    if (!a) continue;
    if (level < (int)a->hostA.size()) return &a->hostA[level];
    level -= (int)a->hostA.size();
  }
  return nullptr;
}

int knp_amg_part_levels(const knp_ctx* c, int32_t part) {
  if (!c) return 0;
  if (c->amg) return part == 0 ? (int)c->amg->hostA.size() : 0;
  const Amg* a = part == 0 ? c->amg_c.get() : (part == 1 ? c->amg_p.get() : nullptr);
  return a ? (int)a->hostA.size() : 0;
}

int knp_amg_num_levels(const knp_ctx* c) { return knp_amg_part_levels(c, 0) + knp_amg_part_levels(c, 1); }

int knp_amg_level_sizes(const knp_ctx* c, int32_t level, int64_t* n, int64_t* nnz) {
  const CsrHost* A = amg_level(c, level);
  KNP_CHECK(A, "no such AMG level");
  if (n) *n = A->n_rows;
  if (nnz) *nnz = A->indptr.empty() ? 0 : (int64_t)A->indptr.back();
  return KNP_OK;
}

int knp_amg_level_host(const knp_ctx* c, int32_t level, int32_t* indptr, int32_t* indices, double* vals) {
  const CsrHost* A = amg_level(c, level);
  KNP_CHECK(A, "no such AMG level");
  KNP_CHECK(!(indices || vals) || A->indptr.empty() || (int64_t)A->indices.size() == (int64_t)A->indptr.back(),
            "the finest level of a hierarchy built on the device is kept on the host only up to 4 M rows "
            "(KNP_AMG_KEEP_HOST=1 keeps it)");
  if (indptr) memcpy(indptr, A->indptr.data(), A->indptr.size() * sizeof(int32_t));
  if (indices) memcpy(indices, A->indices.data(), A->indices.size() * sizeof(int32_t));
  if (vals) memcpy(vals, A->vals.data(), A->vals.size() * sizeof(double));
  return KNP_OK;
}

// ---- host-only structure builder (no GPU needed): restricted dof maps and the CSR pattern of A exactly as knp_create
//      lays them out (build_topology + the index rule of csr_indices_kernel), for the CPU test suite ----
int knp_pattern_host(const knp_mesh_desc* mesh, int64_t* n_rows, int64_t* nnz, int32_t* n_own_loc4, int32_t* indptr,
                     int32_t* indices, int32_t* dof_vert_i, int32_t* dof_vert_e) {
  HostTopo H;
  const bool p2 = mesh && mesh->degree == 2;
  KNP_TRY(p2 ? build_topology_p2(mesh, H) : build_topology(mesh, H));
  const Layout& L = H.L;
  if (n_rows) *n_rows = L.n_rows;
  if (nnz) *nnz = H.nnz;
  if (n_own_loc4) {
    n_own_loc4[0] = L.n_own[0];
    n_own_loc4[1] = L.n_own[1];
    n_own_loc4[2] = L.n_loc[0];
    n_own_loc4[3] = L.n_loc[1];
  }
  if (indptr) memcpy(indptr, H.indptr.data(), H.indptr.size() * sizeof(int32_t));
  if (dof_vert_i) memcpy(dof_vert_i, H.node_vert[0].data(), H.node_vert[0].size() * sizeof(int32_t));
  if (dof_vert_e) memcpy(dof_vert_e, H.node_vert[1].data(), H.node_vert[1].size() * sizeof(int32_t));
  if (indices && p2) {
    memcpy(indices, H.p2.indices.data(), H.p2.indices.size() * sizeof(int32_t));
  } else if (indices) {
    for (int w = 0; w < H.n_work; ++w) {
      const int s = w >= L.n_own[0] ? 1 : 0, p = w - (s ? L.n_own[0] : 0), o = 1 - s;
      const int a0 = H.adj_ptr[w], deg = H.adj_ptr[w + 1] - a0;
      const int g = H.mv_of_node[w];
      const int g0 = g >= 0 ? H.gam_ptr[g] : 0, gdeg = g >= 0 ? H.gam_ptr[g + 1] - g0 : 0;
      for (int f = 0; f < 4; ++f) {
        int pos = H.indptr[L.row(s, f, p)];
        if (s == 1)
          for (int e = 0; e < gdeg; ++e) indices[pos++] = L.col(o, 3, H.mv_node[o][H.gam_mv[g0 + e]]);
        for (int k = (f < 3 ? f : 0); k < (f < 3 ? f + 1 : 3); ++k)
          for (int e = 0; e < deg; ++e) indices[pos++] = L.col(s, k, H.adj_idx[a0 + e]);
        for (int e = 0; e < deg; ++e) indices[pos++] = L.col(s, 3, H.adj_idx[a0 + e]);
        if (s == 0)
          for (int e = 0; e < gdeg; ++e) indices[pos++] = L.col(o, 3, H.mv_node[o][H.gam_mv[g0 + e]]);
      }
    }
  }
  return KNP_OK;
}

// One assembly of the P2 element path ON THE CPU with the bodies the kernels run (p2.cuh): TEST INFRASTRUCTURE for the tier
// without a GPU (tables, slot maps and element math against the oracle); no product call reaches it.
int knp_p2_emulate_host(const knp_mesh_desc* mesh, const knp_params* p, int32_t n_tags, const knp_tag_models* tags, double t,
                        int32_t mode, const double* u, const double* gates, double* vals, double* b) {
  KNP_CHECK(mesh && mesh->degree == 2 && p && u && vals && (mode == 1 || (b && gates)), "knp_p2_emulate_host: bad arguments");
  HostTopo H;
  KNP_TRY(build_topology_p2(mesh, H));
  KParams K{};
  K.dt = p->dt;
  K.F = p->F;
  K.C_M = p->C_M;
  K.psi = p->R * p->T / p->F;
  K.phi_rest = p->phi_rest;
  for (int k = 0; k < 3; ++k) {
    K.z[k] = p->z[k];
    K.D[k] = p->D[k];
    K.g_leak[k] = p->g_leak[k];
    K.g_leak_g[k] = p->g_leak_g[k];
    K.stim_dir[k] = p->stim_dir[k];
    K.stim_lo[k] = p->stim_lo[k];
    K.stim_hi[k] = p->stim_hi[k];
  }
  K.g_Na_bar = p->g_Na_bar;
  K.g_K_bar = p->g_K_bar;
  K.K_e_init = p->K_e_init;
  K.K_i_g_init = p->K_i_g_init;
  K.ode_substeps = p->ode_substeps;
  K.rush_larsen = p->rush_larsen;
  std::vector<uint32_t> tm(H.mtags.size() + 1, 0u);
  std::vector<int32_t> ts(H.mtags.size() + 1, 0);
  for (size_t i = 0; i < H.mtags.size(); ++i)
    for (int j = 0; j < n_tags; ++j)
      if (tags[j].tag == H.mtags[i]) {
        tm[i] = tags[j].models;
        ts[i] = tags[j].stimulated ? 1 : 0;
      }
  const double t_mod = std::fmod(t + 1e-12, p->T_stim);
  double stim_fac = p->g_syn_bar * std::exp(-t_mod / p->a_syn);
  if (p->scale_stimulus) {
    KNP_CHECK(p->stim_area > 0.0, "knp_p2_emulate_host: pass the stimulus area in knp_params::stim_area");
    stim_fac *= 1.0 / p->stim_area;
  }
  return p2_emulate_host(H, K, tm.data(), ts.data(), stim_fac, mode, u, gates, vals, b);
}

// lane-group tables of the edge-lane row kernel, host only (CPU test tier: the tables and the closed-form cell entries are
// checked against the oracle's matrix without a GPU)
int knp_edge_tables_host(const knp_mesh_desc* mesh, int32_t* lgG, int64_t* n_work, int32_t* edge_ok, int32_t* adjG,
                         uint32_t* hitG, int32_t* metaG, double* node_x) {
  HostTopo H;
  KNP_CHECK(mesh && mesh->degree != 2, "the lane-group tables belong to the P1 row kernel");
  KNP_TRY(build_topology(mesh, H));
  if (lgG) *lgG = H.lgG;
  if (n_work) *n_work = H.n_work;
  if (edge_ok) *edge_ok = H.edge_ok;
  if (!H.edge_ok) return KNP_OK;
  if (adjG) memcpy(adjG, H.adjG.data(), H.adjG.size() * sizeof(int32_t));
  if (hitG) memcpy(hitG, H.hitG.data(), H.hitG.size() * sizeof(uint32_t));
  if (metaG) memcpy(metaG, H.metaG.data(), H.metaG.size() * sizeof(int32_t));
  if (node_x) memcpy(node_x, H.node_x.data(), H.node_x.size() * sizeof(double));
  return KNP_OK;
}

// ---- host-only views of two more setup decisions, for the CPU test tier ----
int knp_rowblocks_host(int32_t n_rows, const int32_t* indptr, int32_t max_blocks, int32_t* blocks4, int32_t* n_blocks) {
  KNP_CHECK(n_rows >= 0 && indptr && n_blocks, "bad arguments");
  std::vector<int32_t> blk;
  const int nb = build_rowblocks(indptr, n_rows, blk);
  *n_blocks = nb;
  if (nb > 0 && blocks4) {
    KNP_CHECK(nb <= max_blocks, "block buffer too small (%d blocks)", nb);
    memcpy(blocks4, blk.data(), blk.size() * sizeof(int32_t));
  }
  return KNP_OK;
}

// ---- host-only hierarchy builder (no GPU needed): lets the CPU test suite compare amg_setup.cpp with oracle/amg.py ----
static std::vector<CsrHost> g_host_levels;

int knp_amg_setup_host(int32_t n, const int32_t* indptr, const int32_t* indices, const double* vals, double theta,
                       int32_t coarse_size, int32_t* n_levels) {
  KNP_CHECK(n > 0 && indptr && indices && vals && n_levels, "bad arguments");
  CsrHost A0;
  A0.n_rows = A0.n_cols = n;
  A0.indptr.assign(indptr, indptr + n + 1);
  A0.indices.assign(indices, indices + indptr[n]);
  A0.vals.assign(vals, vals + indptr[n]);
  std::vector<CsrHost> Ps, Rs;
  std::vector<double> rhos, cinv;
  g_host_levels.clear();
  KNP_TRY(amg_setup_host(A0, theta, coarse_size, 16, g_host_levels, Ps, Rs, rhos, cinv, false));
  *n_levels = (int32_t)g_host_levels.size();
  return KNP_OK;
}

int knp_amg_setup_was_on_device(const knp_ctx* c) {
  return c ? c->amg_setup_on_device : 0;
}

int knp_amg_setup_device(int32_t n, const int32_t* indptr, const int32_t* indices, const double* vals, double theta,
                         int32_t coarse_size, int32_t device, int32_t* n_levels) {
  KNP_CHECK(n > 0 && indptr && indices && vals && n_levels, "bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0) {
    set_error("knp_amg_setup_device: no CUDA device %d available (this entry point has no CPU fallback; knp_amg_setup_host is the host form)", device);
    return KNP_E_CUDA;
  }
  KNP_CUDA(cudaSetDevice(device));
  CsrHost A0;
  A0.n_rows = A0.n_cols = n;
  A0.indptr.assign(indptr, indptr + n + 1);
  A0.indices.assign(indices, indices + indptr[n]);
  A0.vals.assign(vals, vals + indptr[n]);
  std::vector<CsrHost> Ps, Rs;
  std::vector<double> rhos, cinv;
  g_host_levels.clear();
  int used = 0;
  KNP_TRY(amg_setup_device(A0, theta, coarse_size, 16, g_host_levels, Ps, Rs, rhos, cinv, nullptr, &used));
  KNP_CHECK(used, "knp_amg_setup_device: the matrix needs the host setup (Dirichlet rows or an unsymmetric pattern)");
  *n_levels = (int32_t)g_host_levels.size();
  return KNP_OK;
}

int knp_amg_host_level(int32_t level, int64_t* n, int64_t* nnz, int32_t* indptr, int32_t* indices, double* vals) {
  KNP_CHECK(level >= 0 && level < (int)g_host_levels.size(), "no such level");
  const CsrHost& A = g_host_levels[level];
  if (n) *n = A.n_rows;
  if (nnz) *nnz = A.nnz();
  if (indptr) memcpy(indptr, A.indptr.data(), A.indptr.size() * sizeof(int32_t));
  if (indices) memcpy(indices, A.indices.data(), A.indices.size() * sizeof(int32_t));
  if (vals) memcpy(vals, A.vals.data(), A.vals.size() * sizeof(double));
  return KNP_OK;
}

}  // extern "C"
