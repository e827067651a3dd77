// Smoothed-aggregation hierarchy construction ON THE DEVICE (single-GPU runs).
//
// Same algorithm and the same DECISIONS as amg_setup.cpp (the host form, which stays the setup of the row-distributed
// hierarchies and the reference of the CPU test tier): strength graph with halved thresholds, MIS(2) aggregation with the
// hashed priorities, filtered prolongator smoothing, Galerkin products.  Every floating-point result is formed in the
// order the host forms it -- one thread walks a row sequentially, products and sums use the non-contracting intrinsics
// (the host code is compiled without FMA) -- so the two setups agree bit for bit (tests/test_gpu_parity.py::
// test_device_amg_setup_equals_host_setup).  Stands in for hypre's setup inside ksp.setUp (KNPEMIx_solver.py:386-389);
// with it `reassemble_P` (:137-150,405-406) costs a fraction of a second instead of the 7 s of the host setup on C3.
//
// Building blocks: a two-level exclusive scan, per-row flag / count / fill passes for the strength graph, max-propagation
// passes for MIS(2), a per-row prolongator pass that accumulates by aggregate inside the row's own scratch segment, a
// transpose by counting + per-row sort, and a row-wise SpGEMM with a per-row open-addressing hash table in a global scratch
// segment sized by the row's upper bound (count / compact / shell sort).  Rows with essential boundary conditions
// (drop_dirichlet_aggregates) and structurally unsymmetric matrices fall back to the host setup.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "context.cuh"

namespace knp {

#define DEV_LAUNCH(kernel, n, ...)                                                  \
  do {                                                                              \
    if ((n) > 0) {                                                                  \
      kernel<<<(unsigned)(((int64_t)(n) + 255) / 256), 256, 0, st>>>(__VA_ARGS__);  \
      KNP_LAUNCHED();                                                               \
    }                                                                               \
  } while (0)

// ------------------------------------------------------------------------------------------------ exclusive scan
constexpr int SCAN_ITEMS = 8;
template <typename Tin>
__global__ void scan_block_kernel(const Tin* __restrict__ in, int64_t* __restrict__ out, int64_t* __restrict__ block_sums,
                                  int64_t n) {
  __shared__ int64_t warp_tot[8];
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * SCAN_ITEMS;
  int64_t loc[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    loc[k] = base + k < n ? (int64_t)in[base + k] : 0;
    sum += loc[k];
  }
  int64_t incl = sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int64_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  int64_t woff = 0;
  for (int w = 0; w < wid; ++w) woff += warp_tot[w];
  int64_t run = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += loc[k];
  }
  if (threadIdx.x == blockDim.x - 1) block_sums[blockIdx.x] = woff + incl;
}
__global__ void scan_add_kernel(int64_t* __restrict__ out, const int64_t* __restrict__ block_off, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += block_off[i / (256 * SCAN_ITEMS)];
}
// out[0..n] = exclusive prefix sums of in[0..n) (out[n] = total, also returned on the host)
template <typename Tin>
static int exclusive_scan(const Tin* in, int64_t* out, int64_t n, int64_t* total, cudaStream_t st) {
  if (n == 0) {
    KNP_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), st));
    *total = 0;
    return KNP_OK;
  }
  const int64_t nb = (n + 256 * SCAN_ITEMS - 1) / (256 * SCAN_ITEMS);
  DevBuf<int64_t> bsum, boff;
  KNP_TRY(bsum.alloc(nb));
  KNP_TRY(boff.alloc(nb + 1));
  scan_block_kernel<Tin><<<(unsigned)nb, 256, 0, st>>>(in, out, bsum.p, n);
  KNP_LAUNCHED();
  int64_t tot = 0;
  if (nb > 1) {
    KNP_TRY(exclusive_scan<int64_t>(bsum.p, boff.p, nb, &tot, st));
    scan_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, boff.p, n);
    KNP_LAUNCHED();
  } else {
    KNP_CUDA(cudaMemcpyAsync(&tot, bsum.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    KNP_CUDA(cudaStreamSynchronize(st));
  }
  KNP_CUDA(cudaMemcpyAsync(out + n, &tot, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  *total = tot;
  return KNP_OK;
}

// ------------------------------------------------------------------------------------------------ small helpers
__global__ void narrow_ptr_kernel(int64_t n1, const int64_t* __restrict__ in, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1) out[i] = (int32_t)in[i];
}
__device__ __forceinline__ void shell_sort(int32_t* __restrict__ key, double* __restrict__ val, int n) {
  const int gaps[8] = {701, 301, 132, 57, 23, 10, 4, 1};
  for (int g = 0; g < 8; ++g) {
    const int gap = gaps[g];
    if (gap >= n && gap > 1) continue;
    for (int i = gap; i < n; ++i) {
      const int32_t k = key[i];
      const double v = val ? val[i] : 0.0;
      int j = i;
      while (j >= gap && key[j - gap] > k) {
        key[j] = key[j - gap];
        if (val) val[j] = val[j - gap];
        j -= gap;
      }
      key[j] = k;
      if (val) val[j] = v;
    }
  }
}
__device__ __forceinline__ int64_t hash32_dev(int64_t i) {
  unsigned long long x = ((unsigned long long)i + 0x9E3779B9ull) & 0xFFFFFFFFull;
  x = ((x ^ (x >> 16)) * 0x85EBCA6Bull) & 0xFFFFFFFFull;
  x = ((x ^ (x >> 13)) * 0xC2B2AE35ull) & 0xFFFFFFFFull;
  x = x ^ (x >> 16);
  return (int64_t)x;
}

// ------------------------------------------------------------------------------------------------ strength graph
__global__ void diag_abs_kernel(int n, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                                const double* __restrict__ v, double* __restrict__ d, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  bool off = false;
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    if (ix[j] == i) s = __dadd_rn(s, v[j]);
    else if (v[j] != 0.0) off = true;
  }
  d[i] = fabs(s);
  if (!off) flags[0] = 1;                       // a Dirichlet row: host setup (drop_dirichlet_aggregates)
}
__global__ void strong_kernel(int n, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                              const double* __restrict__ v, const double* __restrict__ d, double theta,
                              uint8_t* __restrict__ strong) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    const int c = ix[j];
    const double a = v[j];
    strong[j] = (c != i && a != 0.0 && fabs(a) >= __dmul_rn(theta, sqrt(__dmul_rn(d[i], d[c])))) ? 1 : 0;
  }
}
// S + S^T on a structurally symmetric pattern with sorted rows; flags[1] is raised otherwise (host setup then)
__global__ void sym_kernel(int n, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                           const uint8_t* __restrict__ strong, uint8_t* __restrict__ ssym, int32_t* __restrict__ cnt,
                           int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int k = 0;
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    const int c = ix[j];
    uint8_t s = strong[j];
    if (c != i) {
      if (j > ip[i] && ix[j - 1] >= c) flags[1] = 1;
      int lo = ip[c], hi = ip[c + 1];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ix[mid] < i) lo = mid + 1;
        else hi = mid;
      }
      if (lo == ip[c + 1] || ix[lo] != i) {
        if (s) flags[1] = 1;
      } else if (strong[lo]) {
        s = 1;
      }
    }
    ssym[j] = s;
    k += s != 0;
  }
  cnt[i] = k;
}
__global__ void graph_fill_kernel(int n, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                                  const uint8_t* __restrict__ ssym, const int64_t* __restrict__ sp, int32_t* __restrict__ sidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t pos = sp[i];
  for (int j = ip[i]; j < ip[i + 1]; ++j)
    if (ssym[j]) sidx[pos++] = ix[j];
}

// ------------------------------------------------------------------------------------------------ MIS(2) aggregation
constexpr long long MIS_BIG = 1ll << 62;
__global__ void mis_init_kernel(int n, long long* __restrict__ pr, int8_t* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pr[i] = ((hash32_dev(i) & 0x3FFFFFFFll) << 31) | (long long)i;
  state[i] = 0;
}
__global__ void mis_key_kernel(int n, const long long* __restrict__ pr, const int8_t* __restrict__ state, int undecided_too,
                               long long* __restrict__ key, int* __restrict__ any) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int8_t s = state[i];
  if (undecided_too && s == 0) any[0] = 1;
  key[i] = s == 1 ? MIS_BIG + pr[i] : ((s == 0 && undecided_too) ? pr[i] : -1);
}
__global__ void nbr_max_kernel(int n, const int64_t* __restrict__ sp, const int32_t* __restrict__ sidx,
                               const long long* __restrict__ key, long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long m = key[i];
  for (int64_t j = sp[i]; j < sp[i + 1]; ++j) m = max(m, key[sidx[j]]);
  out[i] = m;
}
__global__ void mis_select_kernel(int n, const long long* __restrict__ key, const long long* __restrict__ k2,
                                  int8_t* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && state[i] == 0 && k2[i] == key[i]) state[i] = 1;
}
__global__ void mis_remove_kernel(int n, const long long* __restrict__ k2, int8_t* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && state[i] == 0 && k2[i] >= MIS_BIG) state[i] = -1;
}
__global__ void flag_state_kernel(int n, const int8_t* __restrict__ state, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = state[i] == 1;
}
__global__ void roots_kernel(int n, const int8_t* __restrict__ state, const int64_t* __restrict__ scan,
                             const long long* __restrict__ pr, int32_t* __restrict__ agg, long long* __restrict__ rootpr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool root = state[i] == 1;
  agg[i] = root ? (int32_t)scan[i] : -1;
  rootpr[i] = root ? pr[i] : -1;
}
__global__ void join_kernel(int n, const int32_t* __restrict__ agg, const long long* __restrict__ rootpr,
                            const long long* __restrict__ k1, int32_t* __restrict__ newagg, long long* __restrict__ newpr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t a = agg[i];
  long long p = rootpr[i];
  if (a < 0 && k1[i] >= 0) {
    const int root = (int)(k1[i] & ((1ll << 31) - 1));
    a = agg[root];
    p = k1[i];
  }
  newagg[i] = a;
  newpr[i] = p;
}
__global__ void flag_left_kernel(int n, const int32_t* __restrict__ agg, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = agg[i] < 0;
}
__global__ void singles_kernel(int n, int nroots, const int64_t* __restrict__ scan, int32_t* __restrict__ agg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && agg[i] < 0) agg[i] = nroots + (int32_t)scan[i];
}

// ------------------------------------------------------------------------------------------------ prolongator
__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {      // v >= 0
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
__global__ void bounds_kernel(int n, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                              const double* __restrict__ v, const int64_t* __restrict__ sp, const int32_t* __restrict__ sidx,
                              int filtered, double* __restrict__ dinv, double* __restrict__ rho2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 0.0, s = 0.0;
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    if (ix[j] == i) d = __dadd_rn(d, v[j]);
    s = __dadd_rn(s, fabs(v[j]));
  }
  const double di = 1.0 / d;
  dinv[i] = di;
  double diagF = 0.0, sabs = 0.0;
  int64_t q = sp[i];
  const int64_t qe = sp[i + 1];
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    const int c = ix[j];
    while (q < qe && sidx[q] < c) ++q;
    const bool strong = !filtered || (q < qe && sidx[q] == c);
    if (c == i || !strong) diagF = __dadd_rn(diagF, v[j]);
    else sabs = __dadd_rn(sabs, fabs(v[j]));
  }
  atomic_max_pos(rho2, __dmul_rn(fabs(di), s));
  atomic_max_pos(rho2 + 1, __dmul_rn(fabs(di), __dadd_rn(fabs(diagF), sabs)));
}
// P = T - sc D^-1 A_F T, accumulated by aggregate inside the row's own scratch segment [ip[i], ip[i+1]) (one slot per
// matrix entry is enough), in the host's order: entries of the row ascending, then the lumped diagonal
__global__ void prolong_kernel(int n, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                               const double* __restrict__ v, const int64_t* __restrict__ sp, const int32_t* __restrict__ sidx,
                               const int32_t* __restrict__ agg, int filtered, const double* __restrict__ dinv, double sc,
                               int32_t* __restrict__ pcol, double* __restrict__ pval, int32_t* __restrict__ plen) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t* list = pcol + ip[i];
  double* acc = pval + ip[i];
  const int ai = agg[i];
  int k = 0;
  list[k] = ai;
  acc[k++] = 0.0;
  double diagF = 0.0;
  int64_t q = sp[i];
  const int64_t qe = sp[i + 1];
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    const int c = ix[j];
    const double a = v[j];
    while (q < qe && sidx[q] < c) ++q;
    const bool strong = !filtered || (q < qe && sidx[q] == c);
    if (c == i || !strong) {
      diagF = __dadd_rn(diagF, a);
    } else {
      const int g = agg[c];
      int t = 0;
      while (t < k && list[t] != g) ++t;
      if (t == k) {
        list[k] = g;
        acc[k++] = 0.0;
      }
      acc[t] = __dadd_rn(acc[t], a);
    }
  }
  acc[0] = __dadd_rn(acc[0], diagF);
  const double f = -__dmul_rn(sc, dinv[i]);
  for (int t = 0; t < k; ++t) {
    double p = __dmul_rn(f, acc[t]);
    if (list[t] == ai) p = __dadd_rn(p, 1.0);
    acc[t] = p;
  }
  shell_sort(list, acc, k);
  plen[i] = k;
}
__global__ void compact_rows_kernel(int n, const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ seg_ptr64,
                                    const int32_t* __restrict__ scol, const double* __restrict__ sval,
                                    const int64_t* __restrict__ optr, int32_t* __restrict__ ocol, double* __restrict__ oval) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t s0 = seg_ptr ? (int64_t)seg_ptr[i] : seg_ptr64[i];
  const int64_t o0 = optr[i], len = optr[i + 1] - o0;
  for (int64_t t = 0; t < len; ++t) {
    ocol[o0 + t] = scol[s0 + t];
    oval[o0 + t] = sval[s0 + t];
  }
}

// ------------------------------------------------------------------------------------------------ transpose
__global__ void col_count_kernel(int64_t nnz, const int32_t* __restrict__ ix, int32_t* __restrict__ cnt) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) atomicAdd(&cnt[ix[j]], 1);
}
__global__ void transpose_fill_kernel(int n, const int32_t* __restrict__ ip, const int32_t* __restrict__ ix,
                                      const double* __restrict__ v, const int64_t* __restrict__ tptr,
                                      int32_t* __restrict__ cursor, int32_t* __restrict__ tcol, double* __restrict__ tval) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int j = ip[i]; j < ip[i + 1]; ++j) {
    const int c = ix[j];
    const int64_t pos = tptr[c] + atomicAdd(&cursor[c], 1);
    tcol[pos] = i;
    tval[pos] = v[j];
  }
}
__global__ void sort_rows_kernel(int n, const int64_t* __restrict__ ptr, int32_t* __restrict__ col, double* __restrict__ val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  shell_sort(col + ptr[i], val + ptr[i], (int)(ptr[i + 1] - ptr[i]));
}

// ------------------------------------------------------------------------------------------------ SpGEMM
__global__ void spgemm_bound_kernel(int n, const int32_t* __restrict__ aip, const int32_t* __restrict__ aix,
                                    const int32_t* __restrict__ bip, int32_t* __restrict__ ub) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = 0;
  for (int j = aip[i]; j < aip[i + 1]; ++j) s += bip[aix[j] + 1] - bip[aix[j]];
  ub[i] = s;
}
// one thread per row: entries accumulate in the host's order (row of A ascending, row of B ascending) in an open-addressing
// table inside the row's scratch segment of `cap = upper bound` slots, then the used slots are packed to the front and sorted
__global__ void spgemm_hash_kernel(int n, const int32_t* __restrict__ aip, const int32_t* __restrict__ aix,
                                   const double* __restrict__ av, const int32_t* __restrict__ bip,
                                   const int32_t* __restrict__ bix, const double* __restrict__ bv,
                                   const int64_t* __restrict__ off, int32_t* __restrict__ scol, double* __restrict__ sval,
                                   int32_t* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t o = off[i];
  const unsigned cap = (unsigned)(off[i + 1] - o);
  int32_t* col = scol + o;
  double* val = sval + o;
  if (cap == 0) {
    cnt[i] = 0;
    return;
  }
  for (unsigned t = 0; t < cap; ++t) col[t] = -1;
  int k = 0;
  for (int j = aip[i]; j < aip[i + 1]; ++j) {
    const int r = aix[j];
    const double a = av[j];
    for (int l = bip[r]; l < bip[r + 1]; ++l) {
      const int c = bix[l];
      unsigned h = ((unsigned)c * 2654435761u) % cap;
      while (col[h] != -1 && col[h] != c) h = h + 1 == cap ? 0 : h + 1;
      if (col[h] == -1) {
        col[h] = c;
        val[h] = 0.0;
        ++k;
      }
      val[h] = __dadd_rn(val[h], __dmul_rn(a, bv[l]));
    }
  }
  // pack to the front (a slot is only ever moved towards lower indices), then sort by column
  unsigned w = 0;
  for (unsigned t = 0; t < cap; ++t)
    if (col[t] != -1) {
      const int32_t c = col[t];
      const double x = val[t];
      col[w] = c;
      val[w] = x;
      ++w;
    }
  shell_sort(col, val, k);
  cnt[i] = k;
}

static int narrow_ptr(const int64_t* p64, int64_t n1, DevBuf<int32_t>& out, cudaStream_t st) {
  KNP_TRY(out.alloc(n1));
  narrow_ptr_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(n1, p64, out.p);
  KNP_LAUNCHED();
  return KNP_OK;
}

static int spgemm_device(const DCsr& A, const DCsr& B, DCsr& C, cudaStream_t st) {
  const int n = A.n_rows;
  C.n_rows = n;
  C.n_cols = B.n_cols;
  DevBuf<int32_t> ub, cnt, scol;
  DevBuf<int64_t> off, cptr;
  DevBuf<double> sval;
  KNP_TRY(ub.alloc(n));
  KNP_TRY(cnt.alloc(n));
  KNP_TRY(off.alloc((size_t)n + 1));
  KNP_TRY(cptr.alloc((size_t)n + 1));
  DEV_LAUNCH(spgemm_bound_kernel, n, n, A.indptr.p, A.indices.p, B.indptr.p, ub.p);
  int64_t total = 0;
  KNP_TRY(exclusive_scan<int32_t>(ub.p, off.p, n, &total, st));
  KNP_TRY(scol.alloc((size_t)std::max<int64_t>(total, 1)));
  KNP_TRY(sval.alloc((size_t)std::max<int64_t>(total, 1)));
  DEV_LAUNCH(spgemm_hash_kernel, n, n, A.indptr.p, A.indices.p, A.vals.p, B.indptr.p, B.indices.p, B.vals.p, off.p, scol.p,
             sval.p, cnt.p);
  int64_t nnz = 0;
  KNP_TRY(exclusive_scan<int32_t>(cnt.p, cptr.p, n, &nnz, st));
  KNP_CHECK(nnz < ((int64_t)1 << 31), "device AMG setup: Galerkin product with %lld non-zeros exceeds 32-bit row pointers", (long long)nnz);
  C.nnz = nnz;
  KNP_TRY(C.indices.alloc((size_t)std::max<int64_t>(nnz, 1)));
  KNP_TRY(C.vals.alloc((size_t)std::max<int64_t>(nnz, 1)));
  DEV_LAUNCH(compact_rows_kernel, n, n, nullptr, off.p, scol.p, sval.p, cptr.p, C.indices.p, C.vals.p);
  KNP_TRY(narrow_ptr(cptr.p, (int64_t)n + 1, C.indptr, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

static int transpose_device(const DCsr& A, DCsr& T, cudaStream_t st) {
  const int n = A.n_rows, m = A.n_cols;
  T.n_rows = m;
  T.n_cols = n;
  T.nnz = A.nnz;
  DevBuf<int32_t> cnt;
  DevBuf<int64_t> tptr;
  KNP_TRY(cnt.alloc((size_t)std::max(m, 1)));
  KNP_TRY(tptr.alloc((size_t)m + 1));
  KNP_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)std::max(m, 1) * sizeof(int32_t), st));
  DEV_LAUNCH(col_count_kernel, A.nnz, A.nnz, A.indices.p, cnt.p);
  int64_t total = 0;
  KNP_TRY(exclusive_scan<int32_t>(cnt.p, tptr.p, m, &total, st));
  KNP_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)std::max(m, 1) * sizeof(int32_t), st));
  KNP_TRY(T.indices.alloc((size_t)std::max<int64_t>(A.nnz, 1)));
  KNP_TRY(T.vals.alloc((size_t)std::max<int64_t>(A.nnz, 1)));
  DEV_LAUNCH(transpose_fill_kernel, n, n, A.indptr.p, A.indices.p, A.vals.p, tptr.p, cnt.p, T.indices.p, T.vals.p);
  DEV_LAUNCH(sort_rows_kernel, m, m, tptr.p, T.indices.p, T.vals.p);
  KNP_TRY(narrow_ptr(tptr.p, (int64_t)m + 1, T.indptr, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

int dcsr_download(const DCsr& D, CsrHost& H) {
  H.n_rows = D.n_rows;
  H.n_cols = D.n_cols;
  H.indptr.resize((size_t)D.n_rows + 1);
  H.indices.resize((size_t)D.nnz);
  H.vals.resize((size_t)D.nnz);
  KNP_CUDA(cudaMemcpy(H.indptr.data(), D.indptr.p, H.indptr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (D.nnz) {
    KNP_CUDA(cudaMemcpy(H.indices.data(), D.indices.p, (size_t)D.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    KNP_CUDA(cudaMemcpy(H.vals.data(), D.vals.p, (size_t)D.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return KNP_OK;
}

// strength graph of A at threshold theta: S.ptr (64-bit), S.idx; returns the number of edges; flags: [0] Dirichlet row,
// [1] pattern not symmetric / rows not sorted
static int strength_device(const DCsr& A, double theta, DevBuf<int64_t>& sp, DevBuf<int32_t>& sidx, int64_t* edges,
                           int* hflags, cudaStream_t st) {
  const int n = A.n_rows;
  DevBuf<double> d;
  DevBuf<uint8_t> strong, ssym;
  DevBuf<int32_t> cnt;
  DevBuf<int> flags;
  KNP_TRY(d.alloc(n));
  KNP_TRY(strong.alloc((size_t)std::max<int64_t>(A.nnz, 1)));
  KNP_TRY(ssym.alloc((size_t)std::max<int64_t>(A.nnz, 1)));
  KNP_TRY(cnt.alloc(n));
  KNP_TRY(flags.alloc(2));
  KNP_CUDA(cudaMemsetAsync(flags.p, 0, 2 * sizeof(int), st));
  DEV_LAUNCH(diag_abs_kernel, n, n, A.indptr.p, A.indices.p, A.vals.p, d.p, flags.p);
  DEV_LAUNCH(strong_kernel, n, n, A.indptr.p, A.indices.p, A.vals.p, d.p, theta, strong.p);
  DEV_LAUNCH(sym_kernel, n, n, A.indptr.p, A.indices.p, strong.p, ssym.p, cnt.p, flags.p);
  KNP_TRY(sp.alloc((size_t)n + 1));
  KNP_TRY(exclusive_scan<int32_t>(cnt.p, sp.p, n, edges, st));
  KNP_TRY(sidx.alloc((size_t)std::max<int64_t>(*edges, 1)));
  DEV_LAUNCH(graph_fill_kernel, n, n, A.indptr.p, A.indices.p, ssym.p, sp.p, sidx.p);
  KNP_CUDA(cudaMemcpyAsync(hflags, flags.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  KNP_CUDA(cudaStreamSynchronize(st));
  return KNP_OK;
}

static int mis2_device(int n, const DevBuf<int64_t>& sp, const DevBuf<int32_t>& sidx, DevBuf<int32_t>& agg, int* nagg_out,
                       cudaStream_t st) {
  DevBuf<long long> pr, key, k1, k2, rootpr, newpr;
  DevBuf<int8_t> state;
  DevBuf<int32_t> flag, newagg;
  DevBuf<int64_t> scan;
  DevBuf<int> any;
  KNP_TRY(pr.alloc(n));
  KNP_TRY(key.alloc(n));
  KNP_TRY(k1.alloc(n));
  KNP_TRY(k2.alloc(n));
  KNP_TRY(rootpr.alloc(n));
  KNP_TRY(newpr.alloc(n));
  KNP_TRY(state.alloc(n));
  KNP_TRY(flag.alloc(n));
  KNP_TRY(newagg.alloc(n));
  KNP_TRY(agg.alloc(n));
  KNP_TRY(scan.alloc((size_t)n + 1));
  KNP_TRY(any.alloc(1));
  DEV_LAUNCH(mis_init_kernel, n, n, pr.p, state.p);
  for (int round = 0; round < 10000; ++round) {
    KNP_CUDA(cudaMemsetAsync(any.p, 0, sizeof(int), st));
    DEV_LAUNCH(mis_key_kernel, n, n, pr.p, state.p, 1, key.p, any.p);
    int h_any = 0;
    KNP_CUDA(cudaMemcpyAsync(&h_any, any.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    KNP_CUDA(cudaStreamSynchronize(st));
    if (!h_any) break;
    DEV_LAUNCH(nbr_max_kernel, n, n, sp.p, sidx.p, key.p, k1.p);
    DEV_LAUNCH(nbr_max_kernel, n, n, sp.p, sidx.p, k1.p, k2.p);
    DEV_LAUNCH(mis_select_kernel, n, n, key.p, k2.p, state.p);
    DEV_LAUNCH(mis_key_kernel, n, n, pr.p, state.p, 0, key.p, any.p);
    DEV_LAUNCH(nbr_max_kernel, n, n, sp.p, sidx.p, key.p, k1.p);
    DEV_LAUNCH(nbr_max_kernel, n, n, sp.p, sidx.p, k1.p, k2.p);
    DEV_LAUNCH(mis_remove_kernel, n, n, k2.p, state.p);
  }
  int64_t nroots = 0, nleft = 0;
  DEV_LAUNCH(flag_state_kernel, n, n, state.p, flag.p);
  KNP_TRY(exclusive_scan<int32_t>(flag.p, scan.p, n, &nroots, st));
  DEV_LAUNCH(roots_kernel, n, n, state.p, scan.p, pr.p, agg.p, rootpr.p);
  for (int round = 0; round < 2; ++round) {
    DEV_LAUNCH(nbr_max_kernel, n, n, sp.p, sidx.p, rootpr.p, k1.p);
    DEV_LAUNCH(join_kernel, n, n, agg.p, rootpr.p, k1.p, newagg.p, newpr.p);
    std::swap(agg.p, newagg.p);
    std::swap(rootpr.p, newpr.p);
  }
  DEV_LAUNCH(flag_left_kernel, n, n, agg.p, flag.p);
  KNP_TRY(exclusive_scan<int32_t>(flag.p, scan.p, n, &nleft, st));
  DEV_LAUNCH(singles_kernel, n, n, (int)nroots, scan.p, agg.p);
  KNP_CUDA(cudaStreamSynchronize(st));
  *nagg_out = (int)(nroots + nleft);
  return KNP_OK;
}

int dcsr_upload(const CsrHost& H, DCsr& D) {
  D.n_rows = H.n_rows;
  D.n_cols = H.n_cols;
  D.nnz = H.nnz();
  KNP_TRY(D.indptr.upload(H.indptr));
  KNP_TRY(D.indices.upload(H.indices));
  KNP_TRY(D.vals.upload(H.vals));
  return KNP_OK;
}

int amg_setup_device_core(std::unique_ptr<DCsr>& A0, double theta, int coarse_size, int max_levels, DevHierarchy& out,
                          cudaStream_t st, int* used_device) {
  *used_device = 0;
  const bool timing = getenv("KNP_AMG_TIMING") && atoi(getenv("KNP_AMG_TIMING"));
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what, int level, int n) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "amg device setup level %d (n = %d): %-12s %.3f s\n", level, n, what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  DevHierarchy H;
  const DCsr* cur = A0.get();
  const double omega = 4.0 / 3.0;
  int nlev = 1;
  while (cur->n_rows > coarse_size && nlev < max_levels) {
    const DCsr& A = *cur;
    const int n = A.n_rows;
    DevBuf<int64_t> sp;
    DevBuf<int32_t> sidx;
    double theta_l = theta;
    int64_t edges = 0;
    int hflags[2] = {0, 0};
    for (int attempt = 0; attempt < 4; ++attempt, theta_l *= 0.5) {
      KNP_TRY(strength_device(A, theta_l, sp, sidx, &edges, hflags, st));
      if ((hflags[0] && nlev == 1) || hflags[1]) return KNP_OK;      // host setup (Dirichlet rows live on level 0)
      if ((double)edges >= 3.0 * n) break;
    }
    lap("strength", nlev - 1, n);
    DevBuf<int32_t> agg;
    int nagg = 0;
    KNP_TRY(mis2_device(n, sp, sidx, agg, &nagg, st));
    lap("mis2", nlev - 1, n);
    if (nagg >= 0.8 * n) break;
    const bool filtered = nlev == 1 || (double)A.nnz > 32.0 * n;
    DevBuf<double> dinv, rho2;
    KNP_TRY(dinv.alloc(n));
    KNP_TRY(rho2.alloc(2));
    KNP_CUDA(cudaMemsetAsync(rho2.p, 0, 2 * sizeof(double), st));
    DEV_LAUNCH(bounds_kernel, n, n, A.indptr.p, A.indices.p, A.vals.p, sp.p, sidx.p, (int)filtered, dinv.p, rho2.p);
    double hrho[2];
    KNP_CUDA(cudaMemcpyAsync(hrho, rho2.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    KNP_CUDA(cudaStreamSynchronize(st));
    // prolongator: scratch segments = the rows of A
    auto P = std::make_unique<DCsr>();
    {
      DevBuf<int32_t> pcol, plen;
      DevBuf<double> pval;
      DevBuf<int64_t> pptr;
      KNP_TRY(pcol.alloc((size_t)A.nnz));
      KNP_TRY(pval.alloc((size_t)A.nnz));
      KNP_TRY(plen.alloc(n));
      KNP_TRY(pptr.alloc((size_t)n + 1));
      DEV_LAUNCH(prolong_kernel, n, n, A.indptr.p, A.indices.p, A.vals.p, sp.p, sidx.p, agg.p, (int)filtered, dinv.p,
                 omega / hrho[1], pcol.p, pval.p, plen.p);
      int64_t pnnz = 0;
      KNP_TRY(exclusive_scan<int32_t>(plen.p, pptr.p, n, &pnnz, st));
      P->n_rows = n;
      P->n_cols = nagg;
      P->nnz = pnnz;
      KNP_TRY(P->indices.alloc((size_t)pnnz));
      KNP_TRY(P->vals.alloc((size_t)pnnz));
      DEV_LAUNCH(compact_rows_kernel, n, n, A.indptr.p, nullptr, pcol.p, pval.p, pptr.p, P->indices.p, P->vals.p);
      KNP_TRY(narrow_ptr(pptr.p, (int64_t)n + 1, P->indptr, st));
      KNP_CUDA(cudaStreamSynchronize(st));
    }
    sp.free();
    sidx.free();
    agg.free();
    lap("prolongator", nlev - 1, n);
    auto R = std::make_unique<DCsr>();
    auto Ac = std::make_unique<DCsr>();
    DCsr AP;
    KNP_TRY(transpose_device(*P, *R, st));
    lap("transpose", nlev - 1, n);
    KNP_TRY(spgemm_device(A, *P, AP, st));
    lap("A*P", nlev - 1, n);
    KNP_TRY(spgemm_device(*R, AP, *Ac, st));
    lap("R*(AP)", nlev - 1, n);
    H.rhos.push_back(hrho[0]);
    H.P.push_back(std::move(P));
    H.R.push_back(std::move(R));
    H.A.push_back(std::move(Ac));               // H.A holds the levels 1.. until the setup is known to succeed
    cur = H.A.back().get();
    ++nlev;
  }
  const int nc = cur->n_rows;
  if ((int64_t)nc * nc > (int64_t)64 * 1000 * 1000) {
    set_error("AMG coarsening stalled at %d unknowns; coarsest level too large for a dense solve", nc);
    return KNP_E_UNSUPPORTED;
  }
  out.A.clear();
  out.A.push_back(std::move(A0));
  for (auto& a : H.A) out.A.push_back(std::move(a));
  out.P = std::move(H.P);
  out.R = std::move(H.R);
  out.rhos = std::move(H.rhos);
  *used_device = 1;
  return KNP_OK;
}

// dense form of a (small) CSR operator
static void dense_of(const CsrHost& Ac, std::vector<double>& M) {
  const int nc = Ac.n_rows;
  M.assign((size_t)nc * nc, 0.0);
  for (int i = 0; i < nc; ++i)
    for (int j = Ac.indptr[i]; j < Ac.indptr[i + 1]; ++j) M[(size_t)i * nc + Ac.indices[j]] += Ac.vals[j];
}

int amg_setup_device(const CsrHost& A0, double theta, int coarse_size, int max_levels, std::vector<CsrHost>& As,
                     std::vector<CsrHost>& Ps, std::vector<CsrHost>& Rs, std::vector<double>& rhos,
                     std::vector<double>& coarse_dense, cudaStream_t st, int* used_device) {
  auto d0 = std::make_unique<DCsr>();
  KNP_TRY(dcsr_upload(A0, *d0));
  DevHierarchy H;
  KNP_TRY(amg_setup_device_core(d0, theta, coarse_size, max_levels, H, st, used_device));
  if (!*used_device) return KNP_OK;
  As.assign(H.A.size(), CsrHost());
  Ps.assign(H.P.size(), CsrHost());
  Rs.assign(H.R.size(), CsrHost());
  As[0] = A0;
  for (size_t l = 1; l < H.A.size(); ++l) KNP_TRY(dcsr_download(*H.A[l], As[l]));
  for (size_t l = 0; l < H.P.size(); ++l) {
    KNP_TRY(dcsr_download(*H.P[l], Ps[l]));
    KNP_TRY(dcsr_download(*H.R[l], Rs[l]));
  }
  rhos = H.rhos;
  dense_of(As.back(), coarse_dense);
  return KNP_OK;
}

}  // namespace knp
