// The per-GPU context behind the opaque knp_ctx handle.
#pragma once
#include <memory>
#include "common.cuh"
#include "kernels.cuh"
#include "amg_host.h"
#include "p2.cuh"

struct ncclComm;

namespace knp {
// One peer of a direct NVLink exchange (dist.cu::peer_push_kernel): where this rank's data and flags go in the PEER's memory
// (pointers obtained through CUDA IPC) and where the peer's flags arrive here.
struct PushPeer {
  int64_t send_begin, send_count;
  double* remote_data;                 // peer memory: the slice of its ghost tail (or receive buffer) that belongs to this rank
  unsigned long long* remote_flags;    // peer memory: [0] = "I have consumed epoch e-1", [1] = "epoch e data is complete"
  unsigned long long* local_flags;     // own memory: the same two counters written by the peer
};
// A fixed exchange pattern executed by ONE kernel: pack + remote stores + flag handshake over NVLink / NVSwitch peer memory
struct PeerLink {
  bool ready = false;
  int np = 0;
  DevBuf<PushPeer> peers;
  DevBuf<unsigned long long> epoch;    // per peer: completed exchanges
};

// device side of HaloHost: ghosts are received in place at x + n_own + recv_ptr[i]
struct HaloDev {
  std::vector<int32_t> peers;
  std::vector<int64_t> send_ptr, recv_ptr;
  DevBuf<int32_t> send_idx;
  DevBuf<double> sbuf;
  int n_own = 0;
  PeerLink link;                       // direct peer-memory path (falls back to grouped ncclSend / ncclRecv when not ready)
  const double* link_x = nullptr;      // the vector the link was built for (its ghost tail is the remote target)
};

struct DistLevelDev {
  int n_own = 0, n_ghost = 0;
  CsrDev A, P, R;
  DevBuf<double> dinv, b, r;
  double* x = nullptr;                 // [owned | ghosts], lives in the hierarchy's peer-visible arena
  HaloDev halo;
  double rho = 2.0;
};

// Row-distributed hierarchy: levels split by rows over the ranks (halo exchange per level SpMV), then one level gathered
// onto every rank where the serial hierarchy `tail` continues redundantly (amg_dist.cpp).
struct DistAmg {
  std::vector<std::unique_ptr<DistLevelDev>> levels;
  std::unique_ptr<Amg> tail;
  std::vector<int64_t> off;            // row offsets of the ranks inside the replicated level
  DevBuf<double> arena;                // peer-visible (CUDA IPC) allocation holding every level's x and the gathered rhs gb
  double* gb = nullptr;                // replicated right-hand side (in the arena: the peers write their pieces into it)
  DevBuf<double> gx, rb;               // replicated solution; this rank's piece of the right-hand side
  PeerLink gather;                     // all-to-all of the rhs pieces at the replicated level
  int gamma = 1, gamma_last = 1 << 20;
  std::vector<CsrHost> hostA;          // this rank's rows of the distributed level operators (inspection)
};
}  // namespace knp

struct knp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  knp::HostTopo H;
  knp::DevTopo T{};
  // device copies of the topology
  knp::DevBuf<double> d_node_x, d_mf_area, d_qb, d_qw;
  knp::DevBuf<int32_t> d_adj_ptr, d_adj_idx, d_inc_ptr, d_self_slot, d_mv_of_node, d_gpre;
  knp::DevBuf<uint32_t> d_inc_slots, d_minc, d_hitG;
  knp::DevBuf<int32_t> d_adjG, d_metaG;
  knp::DevBuf<int32_t> d_mv_node0, d_mv_node1, d_mf_mv, d_mf_tagidx, d_gam_ptr, d_gam_mv, d_minc_ptr;
  knp::DevBuf<int32_t> d_indptr, d_indices, d_indptr_P, d_indices_P, d_rowblk_A;
  int nblk_A = 0;
  knp::DevBuf<int32_t> d_cell_nodes[2], d_cell_tag[2], d_cell_owned[2];
  knp::DevBuf<uint32_t> d_tag_models;
  knp::DevBuf<int32_t> d_tag_stim, d_mf_owned;
  // P2 element path (HostTopo::degree == 2): device copies of HostTopo::p2 and the view its kernels read (T.p2 points at it)
  struct P2Dev {
    knp::DevBuf<int32_t> adj_ptr, gam_ptr, inc_ptr, inc_cell, minc_ptr, minc_facet;
    knp::DevBuf<uint8_t> inc_loc, minc_loc;
    knp::DevBuf<uint16_t> inc_slots, minc_own, minc_gam;
    knp::DevBuf<double> cq_w, cq_N, cq_dN, fq_N, fq_M;
  } p2d;
  knp::P2View p2v{};
  // parameters
  knp::Params params{};
  knp::KParams kp{};
  bool params_set = false;
  // state and system
  knp::DevBuf<double> u, gates, A_vals, P_vals, b, fe;
  bool P_assembled = false;
  // time-independent source entries of the right-hand side (knp_set_source)
  knp::DevBuf<int32_t> src_rows;
  knp::DevBuf<double> src_vals;
  int n_src = 0;
  // Dirichlet conditions (knp_set_dirichlet): constrained columns (ascending) with their values, owned rows of A / of P that
  // hold a constrained entry
  knp::DevBuf<int32_t> bc_cols, bc_rows_A, bc_rows_P;
  knp::DevBuf<double> bc_vals;
  int n_bc = 0, n_bc_rows_A = 0, n_bc_rows_P = 0;
  std::vector<int32_t> h_bc_rows;     // constrained OWNED rows (host copy, for the Schur preconditioner setup)
  double t = 0.0;
  int step_index = 0;
  // Krylov workspace
  int ws_restart = 0;
  size_t ldv = 0;
  knp::DevBuf<double> V, w, tmp, tmp2, colscale, partial, hdev, ydev, pc_dinv;
  double* h_pinned = nullptr;
  knp::DevBuf<double> cg_scal, cg_hist;   // device-resident scalars and norm history of the CG loop
  // preconditioner
  int pc_kind = -1;
  int amg_setup_on_device = 0;          // the last hierarchy was built by amg_device.cu (1) or by the host setup (0)
  std::unique_ptr<knp::Amg> amg;
  // charge-conservation Schur preconditioner (pc kind 3): hierarchies of the ion and of the potential blocks
  std::unique_ptr<knp::Amg> amg_c, amg_p;
  // multi-GPU: the same hierarchies (and the one of pc kind 2) distributed by rows
  std::unique_ptr<knp::DistAmg> damg, damg_c, damg_p;
  knp::DevBuf<double> M_vals, msig_inv, sch_vc, sch_zc, sch_t, sch_zp, sch_q, sch_rhs;
  knp::DevBuf<int32_t> sch_mblk[2];   // row blocks of the mass-matrix rows (s, field 0) for the streaming SpMV
  int sch_nmblk[2] = {0, 0};
  // CUDA graphs of the preconditioner application, keyed by the (r, z) pointer pair (single-GPU runs)
  struct PcGraph {
    const double* r;
    double* z;
    cudaGraphExec_t exec;
    unsigned long long launches;
  };
  std::vector<PcGraph> pc_graphs;
  int pc_applies = 0;
  // distributed
  ncclComm* comm = nullptr;
  int rank = 0, nranks = 1;
  std::vector<int32_t> peers;
  std::vector<int64_t> send_ptr, recv_ptr;
  knp::DevBuf<int32_t> d_send_cols;
  knp::DevBuf<double> d_send_buf;
  int64_t n_phi_global = 0;   // global number of potential dofs (nullspace normalisation)
  std::vector<int32_t> h_recv_cols;   // ghost columns in receive order (host copy)
  std::vector<int32_t> h_send_cols;   // owned columns in send order (host copy)
  // direct NVLink exchanges: flag arena (peer-writable through CUDA IPC), the peers' arenas, opened IPC mappings
  knp::DevBuf<unsigned long long> flag_arena;
  int flag_slots_used = 0;
  std::vector<unsigned long long*> peer_flag_arena;    // per rank (nullptr: not mapped)
  std::vector<std::pair<std::string, void*>> ipc_open;  // (rank + handle bytes) -> mapped base pointer
  bool peer_direct = false;
  knp::PeerLink main_link;                              // main 8-field halo: pushes into the peers' receive buffers
  knp::PeerLink red_link;                               // all-reduce: every rank's partial sums into every peer's slot
  knp::DevBuf<double> red_slots;                        // nranks x RED_MAX doubles (peer-visible)
  // timers
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double last_ms[5] = {0, 0, 0, 0, 0};
  // point probes: sparse linear functionals of the state (knp_probe_setup / knp_probe_eval)
  knp::DevBuf<int32_t> probe_ptr, probe_col;
  knp::DevBuf<double> probe_w, probe_out;
  int n_probe = 0;
  // scratch for functionals
  knp::DevBuf<double> fpartial, fout;
  knp::DevBuf<int32_t> ftags;
};

namespace knp {
struct P2POp {
  int peer;
  void* buf;
  size_t bytes;
  bool send;
};
int p2p_exchange(knp_ctx* c, const std::vector<P2POp>& ops, cudaStream_t st);
int halo_exchange_inplace(knp_ctx* c, HaloDev& H, double* x, cudaStream_t st);
int halo_upload(const HaloHost& h, int n_own, HaloDev& d);
constexpr int RED_MAX = 64;            // doubles per rank in the peer-memory all-reduce
// Builds the direct peer-memory exchange for a pattern (collective over the ranks): `target_base` is the base of a cudaMalloc
// allocation of THIS rank that the peers write into, starting `target_offset[i]` doubles from the base for peer i.
int peer_link_create(knp_ctx* c, const std::vector<int32_t>& peers, const std::vector<int64_t>& send_begin,
                     const std::vector<int64_t>& send_count, void* target_base, const std::vector<int64_t>& target_offset,
                     PeerLink& out);
int peer_error_check(knp_ctx* c);
int peer_push(knp_ctx* c, PeerLink& L, const int32_t* send_idx, const double* x, cudaStream_t st);
struct NcclAmgComm : AmgComm {
  knp_ctx* c;
  explicit NcclAmgComm(knp_ctx* ctx) : c(ctx) {
    rank = ctx->rank;
    size = ctx->nranks;
  }
  int alltoallv(const std::vector<std::vector<char>>& send, std::vector<std::vector<char>>& recv) override;
  int allreduce(double* v, int n, bool take_max) override;
  int allgatherv(const std::vector<char>& mine, std::vector<std::vector<char>>& all) override;
};
int ensure_workspace(knp_ctx* c, int restart);
int pc_setup(knp_ctx* c, const knp_solve_opts* o);
int pc_apply(knp_ctx* c, const double* r, double* z, cudaStream_t st);
double pc_bytes(const knp_ctx* c);
int gmres_solve(knp_ctx* c, const double* A_vals, const double* b, double* x, const knp_solve_opts* o,
                knp_solve_info* info, cudaStream_t st);
int krylov_solve(knp_ctx* c, const double* A_vals, const double* b, double* x, const knp_solve_opts* o,
                 knp_solve_info* info, cudaStream_t st);
int halo_exchange(knp_ctx* c, double* x, cudaStream_t st);
int allreduce_sum(knp_ctx* c, double* buf, int n, cudaStream_t st);
const char* last_error();
}  // namespace knp
