// Internal definitions shared by the translation units of libknpemi_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "knpemi_b200.h"

namespace knp {

void set_error(const char* fmt, ...);

#define KNP_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      knp::set_error("%s:%d CUDA error %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return KNP_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

 // counts our own kernel launches (bench.py reports them as gpu_launches)
extern unsigned long long g_kernel_launches;
#define KNP_LAUNCHED()                      \
  do {                                      \
    KNP_CUDA(cudaGetLastError());           \
    ++knp::g_kernel_launches;               \
  } while (0)

#define KNP_CHECK(cond, ...)             \
  do {                                   \
    if (!(cond)) {                       \
      knp::set_error(__VA_ARGS__);       \
      return KNP_E_INVALID;              \
    }                                    \
  } while (0)

#define KNP_TRY(call)           \
  do {                          \
    int rc_ = (call);           \
    if (rc_ != KNP_OK) return rc_; \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    free();
    n = count;
    if (count == 0) return KNP_OK;
    KNP_CUDA(cudaMalloc(&p, count * sizeof(T)));
    return KNP_OK;
  }
  int upload(const std::vector<T>& h) {
    KNP_TRY(alloc(h.size()));
    if (!h.empty()) {
      // A synchronous copy from PAGEABLE host memory returns once the data sits in the driver's staging buffer; the DMA into
      // device memory is ordered on the legacy stream only, and the contexts' streams are non-blocking (no implicit ordering
      // with it).  Without the second call a kernel launched right after an upload can read the tail of the buffer before it
      // has landed (seen as a run-to-run varying coarse inverse / float32 operator copy in the host hierarchy setup).
      KNP_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
      KNP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
    return KNP_OK;
  }
  void free() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { free(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// Column layout: owned (s,f,p) -> rowbase[s] + f*n_own[s] + p ; ghost (s,f,q>=n_own[s]) ->
// n_rows + gbase[s] + f*n_gh[s] + (q - n_own[s]).
struct Layout {
  int n_own[2], n_loc[2], n_gh[2];
  int rowbase[2];
  int gbase[2];
  int n_rows, n_cols;
  __host__ __device__ __forceinline__ int col(int s, int f, int q) const {
    return q < n_own[s] ? rowbase[s] + f * n_own[s] + q
                        : n_rows + gbase[s] + f * n_gh[s] + (q - n_own[s]);
  }
  __host__ __device__ __forceinline__ int row(int s, int f, int p) const {
    return rowbase[s] + f * n_own[s] + p;
  }
};

// Host-side tables of the P2 element path (topology_p2.cpp; p2.cuh describes their use)
struct P2Host {
  int nloc = 0, nt = 0, nqc = 0;
  double hrz = 0.0;                           // 1 / trace of the reference mass matrix: HRZ lumping factor (Schur preconditioner)
  std::vector<int32_t> adj_ptr, adj_idx;      // per owned node: sorted subdomain-local node ids it shares a cell with (incl. itself)
  std::vector<int32_t> gam_ptr, gam_idx;      // per owned node: local node ids ACROSS the membrane it shares a facet with (sorted)
  std::vector<int32_t> inc_ptr, inc_cell;     // per owned node: incident cells, ascending (subdomain-local cell ids)
  std::vector<uint8_t> inc_loc;               // local index of the node in that cell
  std::vector<uint16_t> inc_slots;            // nloc per incidence: adjacency slot of every node of the cell
  std::vector<int32_t> minc_ptr, minc_facet;  // per owned node: incident membrane facets, ascending
  std::vector<uint8_t> minc_loc;              // local index of the node in that facet
  std::vector<uint16_t> minc_own, minc_gam;   // nt per incidence: adjacency slot (own side) / gamma slot (other side) of the facet's nodes
  std::vector<int32_t> indices, indices_P;    // CSR column indices of A and P
  std::vector<double> cq_w, cq_N, cq_dN;      // cell rule: weights, basis values [q][nloc], d/d lambda_m [q][nloc][gdim + 1]
  std::vector<double> fq_b, fq_w;             // the facet rule of the mesh descriptor: barycentrics [q][gdim], weights
  std::vector<double> fq_N, fq_M;             // trace basis at the facet rule's points [q][nt]; reference facet mass [nt][nt]
};
struct P2View;

// Host-built topology (topology.cpp)
struct HostTopo {
  int gdim = 0;
  int degree = 1;                         // 2: P2 elements -- "vertices" are nodes (vertices + edge nodes), tables in p2
  P2Host p2;
  Layout L{};
  int n_work = 0;                         // owned nodes (intra then extra)
  std::vector<double> node_x;             // [n_loc0+n_loc1][gdim], intra local nodes first
  std::vector<int32_t> node_vert[2];      // restricted dof -> local vertex
  std::vector<int32_t> adj_ptr, adj_idx;  // per owned node (unified index w), neighbours as subdomain-local node ids
  std::vector<int32_t> inc_ptr;           // per owned node
  std::vector<uint32_t> inc_slots;        // per (node, cell): adjacency slots of the cell's vertices (8 bit each)
  std::vector<int32_t> self_slot;         // per owned node: slot of itself in its adjacency
  // lane-group tables of the edge-lane row kernel (topology.cpp, end of build_topology)
  int lgG = 0, edge_ok = 0;
  std::vector<int32_t> adjG, metaG;
  std::vector<uint32_t> hitG;
  std::vector<int32_t> mv_of_node;        // per owned node: membrane vertex id or -1
  // membrane
  int n_mv = 0, n_mf = 0;
  std::vector<int32_t> mv_vert, mv_node[2];   // per membrane vertex
  std::vector<int32_t> mf_mv;                 // [n_mf][gdim-...]: d membrane-vertex ids per facet (d = gdim)
  std::vector<int32_t> mf_tagidx;             // index into the sorted unique membrane tag list
  std::vector<int32_t> mf_owned;              // this rank integrates functionals over the facet
  std::vector<double> mf_area;                // facet measure |F|
  std::vector<int32_t> mtags;                 // sorted unique membrane tags
  std::vector<int32_t> gam_ptr, gam_mv;       // per membrane vertex: sorted neighbouring membrane vertices (incl. self)
  std::vector<int32_t> minc_ptr;              // per membrane vertex: incident facets
  std::vector<uint32_t> minc;                 // 4 words per incidence: facet, a|slots_i<<8, slots_e, slots_gam
  // cells per subdomain (for functionals)
  std::vector<int32_t> cell_nodes[2];         // [(gdim+1) per cell] subdomain-local node ids
  std::vector<int32_t> cell_tag[2];
  std::vector<int32_t> cell_owned[2];         // 1 if the cell's lowest-id vertex is owned (integrate once)
  int max_deg = 0, max_gdeg = 0, max_inc = 0;
  // CSR sizes
  int64_t nnz = 0, nnz_P = 0;
  std::vector<int32_t> indptr, indptr_P;      // computed on host (n_rows+1)
  std::vector<int32_t> gpre;                  // per owned node: prefix of gamma degree (own subdomain)
};

int build_topology(const knp_mesh_desc* m, HostTopo& T);

// Device-side views handed to kernels (all pointers device)
struct DevTopo {
  int gdim;
  Layout L;
  int n_work;
  int n_mv, n_mf;
  int nq;
  const double* node_x;
  const int32_t *adj_ptr, *adj_idx, *inc_ptr, *self_slot, *mv_of_node;
  const uint32_t* inc_slots;
  const int32_t *mv_node0, *mv_node1, *mf_mv, *mf_tagidx, *gam_ptr, *gam_mv, *minc_ptr;
  const double* mf_area;
  const uint32_t* minc;
  const int32_t *indptr, *indptr_P;
  const double *qb, *qw;
  const int32_t* gpre;       // per owned node: prefix of the gamma degree
  int max_inc;               // largest number of cells incident to one owned node
  // lane-group tables of the edge-lane row kernel (nullptr: the scan kernel serves the mesh); index (w << lgG) + slot
  int lgG;
  const int32_t* adjG;       // neighbour node or -1
  const uint32_t* hitG;      // cells on the edge as slots of their other vertices (1 word in 2D, 4 words in 3D)
  const int2* metaG;         // {deg | self << 8 | gamma degree << 16, membrane vertex or -1}
  const P2View* p2;          // HOST pointer, non-null on a P2 context: the launchers hand the work to assembly_p2.cu
};

struct Params {
  knp_params p;
  double psi;
  double stim_area;
  int n_tags;
  std::vector<uint32_t> tag_models;   // per membrane tag present on this rank (index = position in HostTopo::mtags)
  std::vector<int> tag_stim;
  bool any_hh;
};

// ---- generic CSR level (AMG) ----
struct CsrDev {
  int n_rows = 0, n_cols = 0;
  int64_t nnz = 0;
  DevBuf<int32_t> indptr, indices, rowblk;
  DevBuf<double> vals;
  DevBuf<float> vals32;   // hierarchy operators are STORED in single precision (solver.cu::to_f32): the products are still
                          // accumulated in double, so the cycle stays a fixed linear operator; vals is freed after conversion
  int nblk = 0;   // row blocks of the streaming SpMV (0: use the CSR-vector kernel)
};

struct CsrHost {
  int n_rows = 0, n_cols = 0;
  std::vector<int32_t> indptr, indices;
  std::vector<double> vals;
  int64_t nnz() const { return (int64_t)indices.size(); }
};

struct AmgLevelDev {
  CsrDev A, P, R;
  DevBuf<double> dinv, x, b, r;
  double rho = 2.0;
};

struct Amg {
  std::vector<AmgLevelDev*> levels;   // levels[0].A is not owned (views the context's P) when external
  DevBuf<double> coarse_inv;          // dense n_c x n_c (row-major)
  DevBuf<float> coarse_inv32;         // the same in single-precision storage (used by the cycle when present)
  DevBuf<double> cb, cx;
  int n_coarse = 0;
  int gamma = 1;                      // cycle index on levels 1..gamma_last (1: V-cycle, 2: W-cycle below the finest level)
  int gamma_last = 1 << 20;
  int level0 = 0;                     // level number of levels[0] when this is the replicated tail of a distributed hierarchy
  // fused tail: levels >= fuse_from run as ONE persistent kernel over a prebuilt operation list (linalg.cu::amg_tail_kernel)
  int fuse_from = -1;                 // -1: not fused
  int tail_nops = 0;
  const double* tail_in = nullptr;    // right-hand side / result pointers the list was built for
  double* tail_out = nullptr;
  DevBuf<unsigned char> tail_ops;     // TailOp array (kernels.cuh)
  DevBuf<unsigned> tail_bar;
  std::vector<CsrHost> hostA;         // kept for inspection
  ~Amg() {
    for (auto* l : levels) delete l;
  }
};

int amg_setup_host(const CsrHost& A0, double theta, int coarse_size, int max_levels,
                   std::vector<CsrHost>& As, std::vector<CsrHost>& Ps, std::vector<CsrHost>& Rs,
                   std::vector<double>& rhos, std::vector<double>& coarse_inv, bool invert = true);
// ---- the same setup on the device (amg_device.cu) ----
struct DCsr {            // device CSR, 32-bit row pointers, double values
  int n_rows = 0, n_cols = 0;
  int64_t nnz = 0;
  DevBuf<int32_t> indptr, indices;
  DevBuf<double> vals;
};
struct DevHierarchy {    // A[l] (l = 0 .. L), P[l], R[l], rho[l] (l < L), all device resident
  std::vector<std::unique_ptr<DCsr>> A, P, R;
  std::vector<double> rhos;
};
int dcsr_upload(const CsrHost& H, DCsr& D);
int dcsr_download(const DCsr& D, CsrHost& H);
// Device-resident form: A0 is consumed (it becomes out.A[0]) only when *used_device = 1; *used_device = 0 means the matrix
// needs the host setup (Dirichlet rows on the finest level, unsymmetric pattern) and A0 / out are untouched.
int amg_setup_device_core(std::unique_ptr<DCsr>& A0, double theta, int coarse_size, int max_levels, DevHierarchy& out,
                          cudaStream_t st, int* used_device);
// Host-in / host-out form with the contract of amg_setup_host; coarse_dense = the dense coarsest OPERATOR (the caller
// inverts it); *used_device = 0: outputs untouched
int amg_setup_device(const CsrHost& A0, double theta, int coarse_size, int max_levels, std::vector<CsrHost>& As,
                     std::vector<CsrHost>& Ps, std::vector<CsrHost>& Rs, std::vector<double>& rhos,
                     std::vector<double>& coarse_dense, cudaStream_t st, int* used_device);

}  // namespace knp
