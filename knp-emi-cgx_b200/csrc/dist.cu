// Multi-GPU plumbing: ghost-column halo exchange (grouped ncclSend/ncclRecv over NVLink) and all-reduce of
// the fused Gram-Schmidt blocks.  Replaces PETSc's VecScatter / ghostUpdate and MPI_Allreduce inside KSP
// (KNPEMIx_solver.py:435,439,458-468).  NCCL is bound lazily with dlopen so that the single-GPU library has
// no link-time dependency on it; the Python host passes a ncclUniqueId broadcast through torch.distributed.
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <string>
#include "context.cuh"

namespace knp {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static bool nccl_bind(NcclApi& api, std::string& err) {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    err = std::string("cannot load libnccl.so.2: ") + dlerror();
    return false;
  }
#define BIND(name)                                                        \
  api.name = (decltype(api.name))dlsym(h, "nccl" #name);                  \
  if (!api.name) {                                                        \
    err = "libnccl.so.2 lacks nccl" #name;                                \
    return false;                                                         \
  }
  BIND(GetUniqueId) BIND(CommInitRank) BIND(AllReduce) BIND(AllGather) BIND(Send) BIND(Recv) BIND(GroupStart) BIND(GroupEnd)
  BIND(GetErrorString) BIND(CommDestroy)
#undef BIND
  api.h = h;
  return true;
}

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] { nccl_bind(api, err); });
  if (!api.h) {
    set_error("%s", err.c_str());
    return nullptr;
  }
  return &api;
}

#define KNP_NCCL(call)                                                                 \
  do {                                                                                 \
    ncclResult_t r_ = (call);                                                          \
    if (r_ != ncclSuccess) {                                                           \
      set_error("%s:%d NCCL error %s", __FILE__, __LINE__, api->GetErrorString(r_));   \
      return KNP_E_NCCL;                                                               \
    }                                                                                  \
  } while (0)

__global__ void pack_kernel(int64_t n, const int32_t* __restrict__ cols, const double* __restrict__ x,
                            double* __restrict__ buf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    buf[i] = x[cols[i]];
}
__global__ void unpack_kernel(int64_t n, const int32_t* __restrict__ cols, const double* __restrict__ buf,
                              double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[cols[i]] = buf[i];
}

int halo_exchange(knp_ctx* c, double* x, cudaStream_t st) {
  if (c->nranks <= 1 || c->peers.empty()) return KNP_OK;
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  const int np = (int)c->peers.size();
  const int64_t ns = c->send_ptr[np], nr = c->recv_ptr[np];
  double* sbuf = c->d_send_buf.p;
  double* rbuf = c->d_send_buf.p + ns;
  if (c->main_link.ready) {
    KNP_TRY(peer_push(c, c->main_link, c->d_send_cols.p, x, st));
    if (nr > 0) {
      int grid = (int)((nr + 255) / 256);
      if (grid > 148 * 8) grid = 148 * 8;
      unpack_kernel<<<grid, 256, 0, st>>>(nr, c->d_send_cols.p + ns, rbuf, x);
      KNP_LAUNCHED();
    }
    return KNP_OK;
  }
  if (ns > 0) {
    int grid = (int)((ns + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    pack_kernel<<<grid, 256, 0, st>>>(ns, c->d_send_cols.p, x, sbuf);
    KNP_LAUNCHED();
  }
  KNP_NCCL(api->GroupStart());
  for (int i = 0; i < np; ++i) {
    const int64_t cs = c->send_ptr[i + 1] - c->send_ptr[i], cr = c->recv_ptr[i + 1] - c->recv_ptr[i];
    if (cs > 0) KNP_NCCL(api->Send(sbuf + c->send_ptr[i], (size_t)cs, ncclFloat64, c->peers[i], c->comm, st));
    if (cr > 0) KNP_NCCL(api->Recv(rbuf + c->recv_ptr[i], (size_t)cr, ncclFloat64, c->peers[i], c->comm, st));
  }
  KNP_NCCL(api->GroupEnd());
  if (nr > 0) {
    int grid = (int)((nr + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    unpack_kernel<<<grid, 256, 0, st>>>(nr, c->d_send_cols.p + ns, rbuf, x);
    KNP_LAUNCHED();
  }
  return KNP_OK;
}

// ---- direct NVLink / NVSwitch exchanges through peer memory (CUDA IPC) ----------------------------------------------
// One kernel per exchange instead of pack kernel + NCCL group (+ unpack): CTA i serves peer i.  It (1) tells the peer that
// this rank has finished reading what the peer sent last time (true by stream order: the consumer kernel precedes this
// launch), (2) waits for the peer's matching acknowledgement, (3) gathers its boundary values and stores them straight into
// the peer's ghost tail over NVLink, (4) publishes "epoch e complete" with a system-scope release and (5) waits for the
// peer's data of the same epoch.  All waits are bounded (PEER_SPIN_LIMIT cycles) so that a lost rank cannot hang the GPU:
// a timeout raises a flag that the host checks after the solve.
constexpr long long PEER_SPIN_LIMIT = 40000000000ll;     // ~20 s at 2 GHz

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_wait(const unsigned long long* flag, unsigned long long e, unsigned long long* err) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < e) {
    if (clock64() - t0 > PEER_SPIN_LIMIT) {
      *err = 1ull;
      break;
    }
  }
}

__global__ void __launch_bounds__(1024) peer_push_kernel(const PushPeer* __restrict__ peers, const int32_t* __restrict__ send_idx,
                                                         const double* __restrict__ x, unsigned long long* __restrict__ epoch,
                                                         int64_t count_override, unsigned long long* __restrict__ err) {
  const PushPeer P = peers[blockIdx.x];
  const unsigned long long e = epoch[blockIdx.x] + 1ull;
  const int64_t count = count_override >= 0 ? count_override : P.send_count;
  if (threadIdx.x == 0) {
    st_release_sys(P.remote_flags, e);
    peer_wait(P.local_flags, e, err);
  }
  __syncthreads();
  for (int64_t k = threadIdx.x; k < count; k += blockDim.x) {
    const int64_t sidx = P.send_begin + k;
    P.remote_data[k] = x[send_idx ? (int64_t)send_idx[sidx] : sidx];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release_sys(P.remote_flags + 1, e);
    peer_wait(P.local_flags + 1, e, err);
    epoch[blockIdx.x] = e;
  }
}

// out[j] = sum over the ranks in rank order of their j-th partial (own partial read from buf): identical on every rank
__global__ void peer_reduce_kernel(int n, int nranks, int me, const double* __restrict__ slots, double* __restrict__ buf) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int r = 0; r < nranks; ++r) s += r == me ? buf[j] : __ldcg(slots + (size_t)r * RED_MAX + j);
  buf[j] = s;
}

int peer_push(knp_ctx* c, PeerLink& L, const int32_t* send_idx, const double* x, cudaStream_t st) {
  if (L.np == 0) return KNP_OK;
  const int threads = 1024;
  peer_push_kernel<<<L.np, threads, 0, st>>>(L.peers.p, send_idx, x, L.epoch.p, -1, c->flag_arena.p);
  KNP_LAUNCHED();
  return KNP_OK;
}

static int ipc_open(knp_ctx* c, int rank, const char* handle64, void** out) {
  std::string key((const char*)&rank, sizeof(int));
  key.append(handle64, sizeof(cudaIpcMemHandle_t));
  for (auto& e : c->ipc_open)
    if (e.first == key) {
      *out = e.second;
      return KNP_OK;
    }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  KNP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->ipc_open.push_back({key, p});
  *out = p;
  return KNP_OK;
}

constexpr int FLAG_SLOTS = 8192;       // pairs of counters in the flag arena; slot 0 word 0 is the timeout flag

// Collective over ALL ranks (every rank calls it in the same order, also with an empty peer list).
int peer_link_create(knp_ctx* c, const std::vector<int32_t>& peers, const std::vector<int64_t>& send_begin,
                     const std::vector<int64_t>& send_count, void* target_base, const std::vector<int64_t>& target_offset,
                     PeerLink& out) {
  out.ready = false;
  out.np = 0;
  if (!c->peer_direct) return KNP_OK;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  const int np = (int)peers.size(), R = c->nranks;
  struct Msg {
    char handle[64];
    int64_t offset;
    int32_t slot, pad;
  };
  std::vector<std::vector<char>> send(R), recv;
  std::vector<int32_t> myslot(np);
  if (np > 0) {
    cudaIpcMemHandle_t h;
    KNP_CUDA(cudaIpcGetMemHandle(&h, target_base));
    for (int i = 0; i < np; ++i) {
      KNP_CHECK(c->flag_slots_used < FLAG_SLOTS, "peer links: flag arena exhausted");
      Msg m{};
      memcpy(m.handle, &h, 64);
      m.offset = target_offset[i];
      m.slot = myslot[i] = c->flag_slots_used++;
      send[peers[i]].resize(sizeof(Msg));
      memcpy(send[peers[i]].data(), &m, sizeof(Msg));
    }
  }
  NcclAmgComm comm(c);
  KNP_TRY(comm.alltoallv(send, recv));
  std::vector<PushPeer> pp(np);
  for (int i = 0; i < np; ++i) {
    const int q = peers[i];
    KNP_CHECK(recv[q].size() == sizeof(Msg), "peer links: rank %d did not answer rank %d (asymmetric exchange pattern)", q, c->rank);
    Msg m;
    memcpy(&m, recv[q].data(), sizeof(Msg));
    void* base = nullptr;
    KNP_TRY(ipc_open(c, q, m.handle, &base));
    pp[i].send_begin = send_begin[i];
    pp[i].send_count = send_count[i];
    pp[i].remote_data = reinterpret_cast<double*>(base) + m.offset;
    pp[i].remote_flags = c->peer_flag_arena[q] + 2 * (size_t)m.slot;
    pp[i].local_flags = c->flag_arena.p + 2 * (size_t)myslot[i];
  }
  if (np > 0) {
    KNP_TRY(out.peers.upload(pp));
    KNP_TRY(out.epoch.alloc(np));
    KNP_CUDA(cudaMemset(out.epoch.p, 0, np * sizeof(unsigned long long)));
    KNP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));    // legacy-stream memset vs the context's non-blocking stream
  }
  out.np = np;
  out.ready = true;
  return KNP_OK;
}

// flag arena + all-reduce slots + the main halo's link; decides collectively whether the direct path is usable
static int peer_direct_init(knp_ctx* c) {
  c->peer_direct = false;
  const char* mode = getenv("KNP_HALO");
  const bool want = !(mode && std::string(mode) == "nccl");
  const int R = c->nranks;
  NcclAmgComm comm(c);
  // every rank exports its flag arena; a rank that cannot (or does not want to) sends an empty handle
  std::vector<char> mine;
  if (want && c->flag_arena.alloc(2 * (size_t)FLAG_SLOTS) == KNP_OK &&
      cudaMemset(c->flag_arena.p, 0, 2 * (size_t)FLAG_SLOTS * sizeof(unsigned long long)) == cudaSuccess &&
      cudaStreamSynchronize(cudaStreamLegacy) == cudaSuccess) {      // the peers write flags as soon as they hold the handle
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->flag_arena.p) == cudaSuccess) mine.assign((const char*)&h, (const char*)&h + 64);
  }
  cudaGetLastError();
  std::vector<std::vector<char>> all;
  KNP_TRY(comm.allgatherv(mine, all));
  double ok = 1.0;
  c->peer_flag_arena.assign(R, nullptr);
  for (int r = 0; r < R && ok > 0.0; ++r) {
    if (all[r].size() != 64) {
      ok = 0.0;
      break;
    }
    if (r == c->rank) {
      c->peer_flag_arena[r] = c->flag_arena.p;
      continue;
    }
    void* p = nullptr;
    if (ipc_open(c, r, all[r].data(), &p) != KNP_OK) {
      ok = 0.0;
      cudaGetLastError();
      break;
    }
    c->peer_flag_arena[r] = reinterpret_cast<unsigned long long*>(p);
  }
  double agree[1] = {ok};
  KNP_TRY(comm.allreduce(agree, 1, false));
  if (agree[0] < (double)R - 0.5) {
    if (want && c->rank == 0) fprintf(stderr, "libknpemi_b200: CUDA IPC peer mapping unavailable, halo exchanges use NCCL point-to-point\n");
    return KNP_OK;
  }
  c->peer_direct = true;
  c->flag_slots_used = 1;            // slot 0 holds the timeout flag
  // main halo: the peers push into this rank's receive buffer (second half of d_send_buf)
  {
    const int np = (int)c->peers.size();
    const int64_t ns = c->send_ptr.empty() ? 0 : c->send_ptr[np];
    std::vector<int64_t> sb(np), sc(np), off(np);
    for (int i = 0; i < np; ++i) {
      sb[i] = c->send_ptr[i];
      sc[i] = c->send_ptr[i + 1] - c->send_ptr[i];
      off[i] = ns + c->recv_ptr[i];
    }
    KNP_TRY(peer_link_create(c, c->peers, sb, sc, c->d_send_buf.p, off, c->main_link));
  }
  // all-reduce: every rank owns nranks x RED_MAX slots; rank r writes its partial sums into slot r of every peer
  {
    KNP_TRY(c->red_slots.alloc((size_t)R * RED_MAX));
    KNP_CUDA(cudaMemset(c->red_slots.p, 0, (size_t)R * RED_MAX * sizeof(double)));
    KNP_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    std::vector<int32_t> peers;
    std::vector<int64_t> sb, sc, off;
    for (int r = 0; r < R; ++r)
      if (r != c->rank) {
        peers.push_back(r);
        sb.push_back(0);
        sc.push_back(RED_MAX);
        off.push_back((int64_t)r * RED_MAX);            // where rank r writes into MY array: slot r
      }
    KNP_TRY(peer_link_create(c, peers, sb, sc, c->red_slots.p, off, c->red_link));
  }
  return KNP_OK;
}

int peer_error_check(knp_ctx* c) {
  if (!c->peer_direct) return KNP_OK;
  unsigned long long flag = 0;
  KNP_CUDA(cudaMemcpy(&flag, c->flag_arena.p, sizeof(flag), cudaMemcpyDeviceToHost));
  if (flag) {
    set_error("a peer-memory exchange timed out: a rank of the job stopped participating");
    return KNP_E_NCCL;
  }
  return KNP_OK;
}

// Ghost exchange of a vector laid out [owned | ghosts grouped by peer in the order the peer packs them] (the layout of
// every level of the row-distributed hierarchies): one pack kernel, then grouped ncclSend / ncclRecv where every receive
// lands IN PLACE in the ghost tail -- no unpack pass.
int halo_exchange_inplace(knp_ctx* c, HaloDev& H, double* x, cudaStream_t st) {
  if (c->nranks <= 1 || H.peers.empty()) return KNP_OK;
  static const bool skip = getenv("KNP_HALO_SKIP") && atoi(getenv("KNP_HALO_SKIP"));   // timing experiments only: WRONG results
  if (skip) return KNP_OK;
  if (H.link.ready && x == H.link_x) return peer_push(c, H.link, H.send_idx.p, x, st);
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  const int np = (int)H.peers.size();
  const int64_t ns = H.send_ptr[np];
  if (ns > 0) {
    int grid = (int)((ns + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    pack_kernel<<<grid, 256, 0, st>>>(ns, H.send_idx.p, x, H.sbuf.p);
    KNP_LAUNCHED();
  }
  KNP_NCCL(api->GroupStart());
  for (int i = 0; i < np; ++i) {
    const int64_t cs = H.send_ptr[i + 1] - H.send_ptr[i], cr = H.recv_ptr[i + 1] - H.recv_ptr[i];
    if (cs > 0) KNP_NCCL(api->Send(H.sbuf.p + H.send_ptr[i], (size_t)cs, ncclFloat64, H.peers[i], c->comm, st));
    if (cr > 0) KNP_NCCL(api->Recv(x + H.n_own + H.recv_ptr[i], (size_t)cr, ncclFloat64, H.peers[i], c->comm, st));
  }
  KNP_NCCL(api->GroupEnd());
  return KNP_OK;
}

int halo_upload(const HaloHost& h, int n_own, HaloDev& d) {
  d.peers = h.peers;
  d.send_ptr = h.send_ptr;
  d.recv_ptr = h.recv_ptr;
  d.n_own = n_own;
  if (d.send_ptr.empty()) d.send_ptr.assign(1, 0);
  if (d.recv_ptr.empty()) d.recv_ptr.assign(1, 0);
  KNP_TRY(d.send_idx.upload(h.send_idx));
  KNP_TRY(d.sbuf.alloc(h.send_idx.size() + 1));
  return KNP_OK;
}

// ---- the setup communication of the row-distributed hierarchies over NCCL (amg_host.h::AmgComm) ----------------------
// Host payloads are staged through device buffers; these calls happen a few times per level at setup time only.
static int nccl_sizes_allgather(knp_ctx* c, NcclApi* api, const std::vector<int64_t>& mine, std::vector<int64_t>& all) {
  const int R = c->nranks, k = (int)mine.size();
  DevBuf<int64_t> ds, dr;
  KNP_TRY(ds.upload(mine));
  KNP_TRY(dr.alloc((size_t)R * k));
  KNP_NCCL(api->AllGather(ds.p, dr.p, (size_t)k, ncclInt64, c->comm, c->stream));
  all.resize((size_t)R * k);
  KNP_CUDA(cudaMemcpyAsync(all.data(), dr.p, all.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  return KNP_OK;
}

int NcclAmgComm::alltoallv(const std::vector<std::vector<char>>& send, std::vector<std::vector<char>>& recv) {
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  const int R = size;
  std::vector<int64_t> mine(R, 0), all;
  for (int q = 0; q < R && q < (int)send.size(); ++q) mine[q] = (int64_t)send[q].size();
  KNP_CHECK(mine[rank] == 0, "alltoallv: message to self");
  KNP_TRY(nccl_sizes_allgather(c, api, mine, all));
  std::vector<int64_t> soff(R + 1, 0), roff(R + 1, 0);
  for (int q = 0; q < R; ++q) {
    soff[q + 1] = soff[q] + ((mine[q] + 15) & ~(int64_t)15);
    roff[q + 1] = roff[q] + ((all[(size_t)q * R + rank] + 15) & ~(int64_t)15);
  }
  std::vector<char> hs((size_t)soff[R]);
  for (int q = 0; q < R; ++q)
    if (mine[q]) memcpy(hs.data() + soff[q], send[q].data(), (size_t)mine[q]);
  DevBuf<char> ds, dr;
  KNP_TRY(ds.upload(hs));
  KNP_TRY(dr.alloc((size_t)roff[R]));
  std::vector<P2POp> ops;
  for (int q = 0; q < R; ++q) {
    if (mine[q]) ops.push_back({q, ds.p + soff[q], (size_t)mine[q], true});
    const int64_t rb = all[(size_t)q * R + rank];
    if (rb) ops.push_back({q, dr.p + roff[q], (size_t)rb, false});
  }
  KNP_TRY(p2p_exchange(c, ops, c->stream));
  std::vector<char> hr((size_t)roff[R]);
  if (!hr.empty()) KNP_CUDA(cudaMemcpyAsync(hr.data(), dr.p, hr.size(), cudaMemcpyDeviceToHost, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  recv.assign(R, {});
  for (int q = 0; q < R; ++q) {
    const int64_t rb = all[(size_t)q * R + rank];
    if (rb) recv[q].assign(hr.begin() + roff[q], hr.begin() + roff[q] + rb);
  }
  return KNP_OK;
}

int NcclAmgComm::allreduce(double* v, int n, bool take_max) {
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  DevBuf<double> d;
  KNP_TRY(d.upload(std::vector<double>(v, v + n)));
  KNP_NCCL(api->AllReduce(d.p, d.p, (size_t)n, ncclFloat64, take_max ? ncclMax : ncclSum, c->comm, c->stream));
  KNP_CUDA(cudaMemcpyAsync(v, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  KNP_CUDA(cudaStreamSynchronize(c->stream));
  return KNP_OK;
}

int NcclAmgComm::allgatherv(const std::vector<char>& mine, std::vector<std::vector<char>>& all) {
  std::vector<std::vector<char>> send(size), recv;
  for (int q = 0; q < size; ++q)
    if (q != rank) send[q] = mine;
  KNP_TRY(alltoallv(send, recv));
  all = std::move(recv);
  all[rank] = mine;
  return KNP_OK;
}

// grouped point-to-point transfers of raw device bytes (setup and the field-parallel preconditioner); transfers between
// one pair of ranks are matched in the order of the list, so both sides must enumerate them identically
int p2p_exchange(knp_ctx* c, const std::vector<P2POp>& ops, cudaStream_t st) {
  if (ops.empty()) return KNP_OK;
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  KNP_NCCL(api->GroupStart());
  for (const P2POp& o : ops) {
    if (o.bytes == 0) continue;
    if (o.send) KNP_NCCL(api->Send(o.buf, o.bytes, ncclChar, o.peer, c->comm, st));
    else KNP_NCCL(api->Recv(o.buf, o.bytes, ncclChar, o.peer, c->comm, st));
  }
  KNP_NCCL(api->GroupEnd());
  return KNP_OK;
}

int allreduce_sum(knp_ctx* c, double* buf, int n, cudaStream_t st) {
  if (c->nranks <= 1) return KNP_OK;
  if (c->red_link.ready && n <= RED_MAX) {
    // every rank stores its partial sums into its slot on every peer, then sums the slots in rank order
    peer_push_kernel<<<c->red_link.np, 64, 0, st>>>(c->red_link.peers.p, nullptr, buf, c->red_link.epoch.p, n, c->flag_arena.p);
    KNP_LAUNCHED();
    peer_reduce_kernel<<<1, 64, 0, st>>>(n, c->nranks, c->rank, c->red_slots.p, buf);
    KNP_LAUNCHED();
    return KNP_OK;
  }
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  KNP_NCCL(api->AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, c->comm, st));
  return KNP_OK;
}

}  // namespace knp

using namespace knp;

extern "C" {

int knp_nccl_unique_id(char* out128) {
  KNP_CHECK(out128, "NULL argument");
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  KNP_NCCL(api->GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return KNP_OK;
}

int knp_dist_init(knp_ctx* c, int32_t rank, int32_t nranks, const char* unique_id128, int64_t n_phi_global,
                  int32_t n_peers, const int32_t* peers, const int64_t* send_ptr, const int32_t* send_cols,
                  const int64_t* recv_ptr, const int32_t* recv_cols) {
  KNP_CHECK(c, "context is NULL");
  KNP_CUDA(cudaSetDevice(c->device));
  KNP_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank/nranks");
  c->rank = rank;
  c->nranks = nranks;
  c->n_phi_global = n_phi_global;
  if (nranks == 1) return KNP_OK;
  KNP_CHECK(c->H.degree != 2, "P2 elements run on one GPU");
  KNP_CHECK(unique_id128 && n_peers >= 0 && send_ptr && recv_ptr && (n_peers == 0 || peers), "NULL argument");
  NcclApi* api = nccl_api();
  if (!api) return KNP_E_NCCL;
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  if (c->comm) {                       // re-initialisation: release the previous communicator
    api->CommDestroy(c->comm);
    c->comm = nullptr;
  }
  KNP_NCCL(api->CommInitRank(&c->comm, nranks, id, rank));
  c->peers.assign(peers, peers + n_peers);
  c->send_ptr.assign(send_ptr, send_ptr + n_peers + 1);
  c->recv_ptr.assign(recv_ptr, recv_ptr + n_peers + 1);
  const int64_t ns = c->send_ptr[n_peers], nr = c->recv_ptr[n_peers];
  const int ncols = c->T.L.n_cols, nrows = c->T.L.n_rows;
  std::vector<int32_t> cols((size_t)(ns + nr));
  for (int64_t i = 0; i < ns; ++i) {
    KNP_CHECK(send_cols[i] >= 0 && send_cols[i] < nrows, "send column %d is not an owned column", send_cols[i]);
    cols[i] = send_cols[i];
  }
  for (int64_t i = 0; i < nr; ++i) {
    KNP_CHECK(recv_cols[i] >= nrows && recv_cols[i] < ncols, "recv column %d is not a ghost column", recv_cols[i]);
    cols[ns + i] = recv_cols[i];
  }
  c->h_recv_cols.assign(recv_cols, recv_cols + nr);
  c->h_send_cols.assign(send_cols, send_cols + ns);
  KNP_TRY(c->d_send_cols.upload(cols));
  KNP_TRY(c->d_send_buf.alloc((size_t)(ns + nr) + 1));
  return peer_direct_init(c);
}

int knp_peer_direct(const knp_ctx* c) { return c && c->peer_direct ? 1 : 0; }

int knp_halo_exchange(knp_ctx* c, double* x_dev, void* stream) {
  KNP_CHECK(c, "context is NULL");
  KNP_CUDA(cudaSetDevice(c->device));
  return halo_exchange(c, x_dev ? x_dev : c->u.p, stream ? (cudaStream_t)stream : c->stream);
}

int knp_allreduce_sum(knp_ctx* c, double* buf_dev, int32_t n, void* stream) {
  KNP_CHECK(c && buf_dev, "NULL argument");
  KNP_CUDA(cudaSetDevice(c->device));
  return allreduce_sum(c, buf_dev, n, stream ? (cudaStream_t)stream : c->stream);
}

}  // extern "C"
