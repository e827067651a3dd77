// Multi-GPU plumbing: ghost-column halo exchange (grouped ncclSend/ncclRecv over NVLink) and all-reduce of
// the fused Gram-Schmidt blocks.  Replaces PETSc's VecScatter / ghostUpdate and MPI_Allreduce inside KSP
// (KNPEMIx_solver.py:435,439,458-468).  NCCL is bound lazily with dlopen so that the single-GPU library has
// no link-time dependency on it; the Python host passes a ncclUniqueId broadcast through torch.distributed.
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
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

// Ghost exchange of a vector laid out [owned | ghosts grouped by peer in the order the peer packs them] (the layout of
// every level of the row-distributed hierarchies): one pack kernel, then grouped ncclSend / ncclRecv where every receive
// lands IN PLACE in the ghost tail -- no unpack pass.
int halo_exchange_inplace(knp_ctx* c, HaloDev& H, double* x, cudaStream_t st) {
  if (c->nranks <= 1 || H.peers.empty()) return KNP_OK;
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
  return KNP_OK;
}

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
