// Sparse / dense linear-algebra kernels of the Krylov solver (sm_100a).
//
//   spmv_kernel<LANES,EPI>  K5   CSR y = A x, LANES lanes per row, fused epilogues (residual, Jacobi sweep,
//                                prolongation add) -- replaces PETSc MatMult / hypre relax inside ksp.solve
//                                (KNPEMIx_solver.py:435)
//   multi_dot / multi_axpy  K6   classical Gram-Schmidt building blocks: all (j+1) inner products and ||w||^2
//                                in ONE pass over w and V, reduced in a fixed order (two-stage, no atomics)
//   range_sum / range_shift K8   nullspace projection x -= (ns.x) ns (KNPEMIx_solver.py:324-333)
#include <cstdlib>
#include "common.cuh"
#include "kernels.cuh"

namespace knp {

__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return (double)v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

static int grid_for(int n) {
  int grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  return grid < 1 ? 1 : grid;
}

template <int LANES, int EPI, typename VT>
__global__ void __launch_bounds__(256) spmv_kernel(int n_rows, const int32_t* __restrict__ indptr,
                                                   const int32_t* __restrict__ indices,
                                                   const VT* __restrict__ vals, const double* __restrict__ x,
                                                   double* __restrict__ out, const double* __restrict__ b,
                                                   const double* __restrict__ dinv, double w) {
  const int lane = threadIdx.x & (LANES - 1);
  constexpr int RPB = 256 / LANES;
  // block-uniform trip count so that the sub-warp shuffles are always convergent
  for (int base = blockIdx.x * RPB; base < n_rows; base += gridDim.x * RPB) {
    const int row = base + threadIdx.x / LANES;
    const bool valid = row < n_rows;
    double sum = 0.0;
    if (valid) {
      const int r0 = indptr[row], r1 = indptr[row + 1];
      for (int j = r0 + lane; j < r1; j += LANES) sum += ld_stream(vals + j) * __ldg(x + ld_stream(indices + j));
    }
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off, LANES);
    if (valid && lane == 0) {
      double r;
      if (EPI == EPI_SET) r = sum;
      else if (EPI == EPI_RESID) r = b[row] - sum;
      else if (EPI == EPI_JACOBI) r = x[row] + w * dinv[row] * (b[row] - sum);
      else r = out[row] + sum;
      out[row] = r;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Streaming SpMV: persistent CTAs walk over precomputed row blocks (<= SPMV_CAP non-zeros, <= 256 rows).  One elected
// thread moves the block's values and column indices into shared memory with two TMA 1-D bulk copies
// (cp.async.bulk ... mbarrier::complete_tx) into a 2-stage ring, so the next block streams in while the current one is
// processed.  Every thread then owns SPMV_CAP/256 consecutive-stride elements: the x gathers of a thread are independent
// (memory-level parallelism 8 instead of 1), products are written back in place, and each row is summed by one thread in
// a fixed order (bitwise reproducible, no atomics).
#ifndef KNP_SPMV_CAP
#define KNP_SPMV_CAP 1280
#endif
#ifndef KNP_SPMV_CTAS
#define KNP_SPMV_CTAS 6
#endif
constexpr int SPMV_CAP = KNP_SPMV_CAP;
constexpr int SPMV_THREADS = 256;
constexpr int SPMV_PAD = 8;
constexpr int SPMV_ROWS = 256;          // max rows per block (multiple of 4: the row-pointer slice is TMA-loaded too)

template <typename VT>
struct SpmvStage {
  VT v[SPMV_CAP + SPMV_PAD];
  int c[SPMV_CAP + SPMV_PAD];
  int rp[SPMV_ROWS + 8];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// blkinfo[blk] = {first row r0 (multiple of 4), number of rows, a4 = 4-aligned first non-zero, staged element count}
template <int EPI, int LANES, typename VT>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_stream_kernel(int nblk, const int4* __restrict__ blkinfo,
                                                                    const int32_t* __restrict__ indptr,
                                                                    const int32_t* __restrict__ indices,
                                                                    const VT* __restrict__ vals,
                                                                    const double* __restrict__ x, double* __restrict__ out,
                                                                    const double* __restrict__ b,
                                                                    const double* __restrict__ dinv, double w) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SpmvStage<VT>* st = reinterpret_cast<SpmvStage<VT>*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * sizeof(SpmvStage<VT>));
  __shared__ double rsum[SPMV_ROWS];
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](const int4 bi, int s) {   // thread 0 only: three bulk copies completing on one mbarrier
    const int ntma = bi.w & ~3;
    const int nrp = (bi.y + 1 + 3) & ~3;     // row pointers r0 .. r0+nrows, rounded up to 16 bytes (arrays are padded)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar[s], (uint32_t)ntma * (uint32_t)(sizeof(VT) + 4) + (uint32_t)nrp * 4u);
    tma_load_1d(st[s].rp, indptr + bi.x, (uint32_t)nrp * 4u, &bar[s]);
    if (ntma > 0) {
      tma_load_1d(st[s].v, vals + bi.z, (uint32_t)ntma * (uint32_t)sizeof(VT), &bar[s]);
      tma_load_1d(st[s].c, indices + bi.z, (uint32_t)ntma * 4u, &bar[s]);
    }
  };

  const int stride = gridDim.x;
  int blk = blockIdx.x;
  const int4 zero = make_int4(0, 0, 0, 0);
  int4 cur = blk < nblk ? blkinfo[blk] : zero;
  int4 nx1 = blk + stride < nblk ? blkinfo[blk + stride] : zero;
  if (tid == 0 && blk < nblk) issue(cur, 0);
  for (int it = 0; blk < nblk; blk += stride, ++it) {
    const int s = it & 1;
    // block descriptors are prefetched two iterations ahead so that neither the producer nor the consumers wait on them
    const int4 nx2 = blk + 2 * stride < nblk ? blkinfo[blk + 2 * stride] : zero;
    if (tid == 0 && blk + stride < nblk) issue(nx1, s ^ 1);
    const int r0 = cur.x, nrows = cur.y, a4 = cur.z, n = cur.w;
    const int ntma = n & ~3;
    VT* sv = st[s].v;
    int* sc = st[s].c;
    const int* rp = st[s].rp;
    // epilogue operands of "my" row (thread t <-> row r0 + t): coalesced, issued long before they are needed
    double eb = 0.0, ed = 0.0, ex = 0.0;
    if (tid < nrows) {
      const int r = r0 + tid;
      if (EPI == EPI_RESID || EPI == EPI_JACOBI) eb = b[r];
      if (EPI == EPI_JACOBI) {
        ed = dinv[r];
        ex = x[r];
      }
      if (EPI == EPI_ADD) eb = out[r];
    }
    // the (< 4) trailing elements that do not fill a 16-byte unit
    if (tid < n - ntma) {
      sv[ntma + tid] = vals[a4 + ntma + tid];
      sc[ntma + tid] = indices[a4 + ntma + tid];
    }
    mbar_wait(&bar[s], (uint32_t)((it >> 1) & 1));
    __syncthreads();
    // LANES lanes per row walk the staged row: lane l takes elements j0 + l, j0 + l + LANES, ... so that one warp-level
    // gather touches the same stencil column of 32 / LANES consecutive rows -- neighbouring rows have neighbouring
    // columns, i.e. the x gathers coalesce into a few cache lines.  Four gathers are in flight per lane; the partial
    // sums are combined in a fixed order (bitwise reproducible, no atomics).
    for (int i0 = 0; i0 < nrows; i0 += SPMV_THREADS / LANES) {
      const int i = i0 + tid / LANES;
      const int l = tid % LANES;
      double sum = 0.0;
      if (i < nrows) {
        const int j1 = rp[i + 1] - a4;
        int j = rp[i] - a4 + l;
        for (; j + 3 * LANES < j1; j += 4 * LANES) {
          const double x0 = __ldg(x + sc[j]), x1 = __ldg(x + sc[j + LANES]);
          const double x2 = __ldg(x + sc[j + 2 * LANES]), x3 = __ldg(x + sc[j + 3 * LANES]);
          sum += sv[j] * x0;
          sum += sv[j + LANES] * x1;
          sum += sv[j + 2 * LANES] * x2;
          sum += sv[j + 3 * LANES] * x3;
        }
        for (; j < j1; j += LANES) sum += sv[j] * __ldg(x + sc[j]);
      }
#pragma unroll
      for (int off = LANES >> 1; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off, LANES);
      if (l == 0 && i < nrows) rsum[i] = sum;
    }
    __syncthreads();
    if (tid < nrows) {
      const double sum = rsum[tid];
      double res;
      if (EPI == EPI_SET) res = sum;
      else if (EPI == EPI_RESID) res = eb - sum;
      else if (EPI == EPI_JACOBI) res = ex + w * ed * (eb - sum);
      else res = eb + sum;
      out[r0 + tid] = res;
    }
    __syncthreads();
    cur = nx1;
    nx1 = nx2;
  }
}

// greedy row blocks: first row a multiple of 4, <= SPMV_CAP staged non-zeros (counted from the 4-aligned start),
// <= SPMV_ROWS rows.  Returns the number of blocks or -1 when a group of 4 rows exceeds the stage capacity.
int build_rowblocks(const int32_t* indptr, int n_rows, std::vector<int32_t>& info) {
  info.clear();
  int r0 = 0;
  while (r0 < n_rows) {
    const int a4 = indptr[r0] & ~3;
    int r1 = r0;
    while (r1 < n_rows) {
      const int nxt = r1 + 4 < n_rows ? r1 + 4 : n_rows;
      if (nxt - r0 > SPMV_ROWS || indptr[nxt] - a4 > SPMV_CAP) break;
      r1 = nxt;
    }
    if (r1 == r0) return -1;
    info.push_back(r0);
    info.push_back(r1 - r0);
    info.push_back(a4);
    info.push_back(indptr[r1] - a4);
    r0 = r1;
  }
  return (int)(info.size() / 4);
}

template <int LANES, typename VT>
static int launch_spmv_stream_l(int nblk, const int32_t* blkinfo, const int32_t* indptr, const int32_t* indices,
                                const VT* vals, const double* x, double* out, int epi, const double* b,
                                const double* dinv, double w, cudaStream_t st) {
  const size_t smem = 2 * sizeof(SpmvStage<VT>) + 2 * sizeof(uint64_t);
  static bool configured = false;
  if (!configured) {
    KNP_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<EPI_SET, LANES, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KNP_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<EPI_RESID, LANES, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KNP_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<EPI_JACOBI, LANES, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KNP_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<EPI_ADD, LANES, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int grid = nblk < 148 * KNP_SPMV_CTAS ? nblk : 148 * KNP_SPMV_CTAS;   // persistent CTAs
  const int4* bi = reinterpret_cast<const int4*>(blkinfo);
  switch (epi) {
    case EPI_SET: spmv_stream_kernel<EPI_SET, LANES, VT><<<grid, SPMV_THREADS, smem, st>>>(nblk, bi, indptr, indices, vals, x, out, b, dinv, w); break;
    case EPI_RESID: spmv_stream_kernel<EPI_RESID, LANES, VT><<<grid, SPMV_THREADS, smem, st>>>(nblk, bi, indptr, indices, vals, x, out, b, dinv, w); break;
    case EPI_JACOBI: spmv_stream_kernel<EPI_JACOBI, LANES, VT><<<grid, SPMV_THREADS, smem, st>>>(nblk, bi, indptr, indices, vals, x, out, b, dinv, w); break;
    default: spmv_stream_kernel<EPI_ADD, LANES, VT><<<grid, SPMV_THREADS, smem, st>>>(nblk, bi, indptr, indices, vals, x, out, b, dinv, w); break;
  }
  KNP_LAUNCHED();
  return KNP_OK;
}

template <typename VT>
static int launch_spmv_stream_t(int nblk, const int32_t* blkinfo, const int32_t* indptr, const int32_t* indices, const VT* vals,
                                const double* x, double* out, int epi, const double* b, const double* dinv, double w,
                                cudaStream_t st, double avg_row) {
  if (nblk == 0) return KNP_OK;
  static const int force = getenv("KNP_SPMV_LANES") ? atoi(getenv("KNP_SPMV_LANES")) : 0;
  const int lanes = force ? force : (avg_row <= 10.0 ? 1 : avg_row <= 24.0 ? 2 : 4);
  switch (lanes) {
    case 1: return launch_spmv_stream_l<1, VT>(nblk, blkinfo, indptr, indices, vals, x, out, epi, b, dinv, w, st);
    case 2: return launch_spmv_stream_l<2, VT>(nblk, blkinfo, indptr, indices, vals, x, out, epi, b, dinv, w, st);
    case 4: return launch_spmv_stream_l<4, VT>(nblk, blkinfo, indptr, indices, vals, x, out, epi, b, dinv, w, st);
    default: return launch_spmv_stream_l<8, VT>(nblk, blkinfo, indptr, indices, vals, x, out, epi, b, dinv, w, st);
  }
}

int launch_spmv_stream(int nblk, const int32_t* blkinfo, const int32_t* indptr, const int32_t* indices, const double* vals,
                       const double* x, double* out, int epi, const double* b, const double* dinv, double w, cudaStream_t st,
                       double avg_row) {
  return launch_spmv_stream_t<double>(nblk, blkinfo, indptr, indices, vals, x, out, epi, b, dinv, w, st, avg_row);
}

template <int LANES, typename VT>
static int launch_spmv_l(int grid, int n_rows, const int32_t* indptr, const int32_t* indices, const VT* vals,
                         const double* x, double* out, int epi, const double* b, const double* dinv, double w,
                         cudaStream_t st) {
  switch (epi) {
    case EPI_SET: spmv_kernel<LANES, EPI_SET, VT><<<grid, 256, 0, st>>>(n_rows, indptr, indices, vals, x, out, b, dinv, w); break;
    case EPI_RESID: spmv_kernel<LANES, EPI_RESID, VT><<<grid, 256, 0, st>>>(n_rows, indptr, indices, vals, x, out, b, dinv, w); break;
    case EPI_JACOBI: spmv_kernel<LANES, EPI_JACOBI, VT><<<grid, 256, 0, st>>>(n_rows, indptr, indices, vals, x, out, b, dinv, w); break;
    default: spmv_kernel<LANES, EPI_ADD, VT><<<grid, 256, 0, st>>>(n_rows, indptr, indices, vals, x, out, b, dinv, w); break;
  }
  KNP_LAUNCHED();
  return KNP_OK;
}

template <typename VT>
static int launch_spmv_t(int n_rows, int64_t nnz, const int32_t* indptr, const int32_t* indices, const VT* vals,
                         const double* x, double* out, int epi, const double* b, const double* dinv, double w,
                         cudaStream_t st) {
  if (n_rows == 0) return KNP_OK;
  const double avg = (double)nnz / n_rows;
  int lanes = avg <= 3 ? 2 : avg <= 6 ? 4 : avg <= 12 ? 8 : avg <= 40 ? 16 : 32;
  const int rows_per_block = 256 / lanes;
  const int64_t want = ((int64_t)n_rows + rows_per_block - 1) / rows_per_block;
  const int grid = (int)(want < (int64_t)148 * 64 ? want : (int64_t)148 * 64);
  switch (lanes) {
    case 2: return launch_spmv_l<2, VT>(grid, n_rows, indptr, indices, vals, x, out, epi, b, dinv, w, st);
    case 4: return launch_spmv_l<4, VT>(grid, n_rows, indptr, indices, vals, x, out, epi, b, dinv, w, st);
    case 8: return launch_spmv_l<8, VT>(grid, n_rows, indptr, indices, vals, x, out, epi, b, dinv, w, st);
    case 16: return launch_spmv_l<16, VT>(grid, n_rows, indptr, indices, vals, x, out, epi, b, dinv, w, st);
    default: return launch_spmv_l<32, VT>(grid, n_rows, indptr, indices, vals, x, out, epi, b, dinv, w, st);
  }
}

int launch_spmv(int n_rows, int64_t nnz, const int32_t* indptr, const int32_t* indices, const double* vals,
                const double* x, double* out, int epi, const double* b, const double* dinv, double w,
                cudaStream_t st) {
  return launch_spmv_t<double>(n_rows, nnz, indptr, indices, vals, x, out, epi, b, dinv, w, st);
}

int spmv(const CsrView& M, const double* x, double* out, int epi, const double* b, const double* dinv, double w,
         cudaStream_t st) {
  const void* vp = M.vals32 ? (const void*)M.vals32 : (const void*)M.vals;
  const bool aligned = (((uintptr_t)vp | (uintptr_t)M.indices) & 15u) == 0;
  // the TMA-staged kernel pays off on long-enough rows and enough blocks to fill the persistent grid
  static const double min_avg = getenv("KNP_SPMV_STREAM_MIN_AVG") ? atof(getenv("KNP_SPMV_STREAM_MIN_AVG")) : 0.0;
  static const int min_blk = getenv("KNP_SPMV_STREAM_MIN_BLK") ? atoi(getenv("KNP_SPMV_STREAM_MIN_BLK")) : 1;
  const double avg = M.n_rows > 0 ? (double)M.nnz / M.n_rows : 0.0;
  const bool stream = M.nblk >= min_blk && M.rowblk && aligned && avg >= min_avg;
  if (M.vals32) {
    if (stream) return launch_spmv_stream_t<float>(M.nblk, M.rowblk, M.indptr, M.indices, M.vals32, x, out, epi, b, dinv, w, st, avg);
    return launch_spmv_t<float>(M.n_rows, M.nnz, M.indptr, M.indices, M.vals32, x, out, epi, b, dinv, w, st);
  }
  if (stream) return launch_spmv_stream(M.nblk, M.rowblk, M.indptr, M.indices, M.vals, x, out, epi, b, dinv, w, st, avg);
  return launch_spmv(M.n_rows, M.nnz, M.indptr, M.indices, M.vals, x, out, epi, b, dinv, w, st);
}

__global__ void to_f32_kernel(int64_t n, const double* __restrict__ src, float* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}
int launch_to_f32(int64_t n, const double* src, float* dst, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  const int64_t want = (n + 255) / 256;
  to_f32_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, st>>>(n, src, dst);
  KNP_LAUNCHED();
  return KNP_OK;
}

__global__ void scale_dinv_kernel(int n, double w, const double* __restrict__ dinv, const double* __restrict__ b,
                                  double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = w * dinv[i] * b[i];
}
int launch_scale_dinv(int n, double w, const double* dinv, const double* b, double* x, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  int grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  scale_dinv_kernel<<<grid, 256, 0, st>>>(n, w, dinv, b, x);
  KNP_LAUNCHED();
  return KNP_OK;
}

__global__ void extract_dinv_kernel(int n_rows, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                    const double* __restrict__ vals, double* __restrict__ dinv) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  double d = 0.0;
  for (int j = indptr[row]; j < indptr[row + 1]; ++j)
    if (indices[j] == row) d += vals[j];
  dinv[row] = 1.0 / d;
}
int launch_extract_dinv(int n_rows, const int32_t* indptr, const int32_t* indices, const double* vals, double* dinv,
                        cudaStream_t st) {
  if (n_rows == 0) return KNP_OK;
  extract_dinv_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(n_rows, indptr, indices, vals, dinv);
  KNP_LAUNCHED();
  return KNP_OK;
}

// x = Minv b, Minv dense row-major n x n (double or single-precision storage); one warp per row
template <typename VT>
__global__ void dense_gemv_kernel(int n, const VT* __restrict__ M, const double* __restrict__ b,
                                  double* __restrict__ x) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += M[(size_t)row * n + j] * b[j];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) x[row] = s;
}
int launch_dense_gemv(int n, const double* Minv, const double* b, double* x, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  dense_gemv_kernel<double><<<(n + 7) / 8, 256, 0, st>>>(n, Minv, b, x);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_dense_gemv(int n, const float* Minv, const double* b, double* x, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  dense_gemv_kernel<float><<<(n + 7) / 8, 256, 0, st>>>(n, Minv, b, x);
  KNP_LAUNCHED();
  return KNP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused tail of an AMG cycle.  Below the first one or two levels every operator of the hierarchy fits in L2 and a kernel
// launch per SpMV costs more than the SpMV (round-1 launch list: 100 of the 143 launches of one preconditioner
// application took < 12 us).  amg_tail_kernel is ONE persistent launch that executes the whole sub-cycle below a given
// level -- pre-smoothing, residuals, restrictions, the recursive W-cycle visits, the dense coarsest solve, prolongations,
// post-smoothing -- as a host-built list of operations separated by grid-wide barriers (sense-reversing counter in global
// memory; the grid is sized so that every CTA is resident).  Matrix data is read through the read-only path, vectors
// that other CTAs wrote earlier in the same launch through L2 (ld.global.cg).
__device__ __forceinline__ void tail_grid_barrier(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    unsigned prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
    if (prev == nblocks - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], 0;" ::"l"(bar) : "memory");
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
    } else {
      unsigned cur;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar + 1) : "memory");
      } while (cur == gen);
    }
  }
  __syncthreads();
}

constexpr int TAIL_THREADS = 1024;

// four independent (index -> x) gathers in flight per lane: these levels are L2-resident, so the loop is latency-bound
template <bool RESID0>
__device__ __forceinline__ double tail_row_sum(const TailOp& op, int r0, int r1, int lane, int lanes) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = r0 + lane;
  for (; j + 3 * lanes < r1; j += 4 * lanes) {
    const int c0 = __ldg(op.indices + j), c1 = __ldg(op.indices + j + lanes), c2 = __ldg(op.indices + j + 2 * lanes),
              c3 = __ldg(op.indices + j + 3 * lanes);
    const double v0 = __ldg(op.vals + j), v1 = __ldg(op.vals + j + lanes), v2 = __ldg(op.vals + j + 2 * lanes),
                 v3 = __ldg(op.vals + j + 3 * lanes);
    double x0, x1, x2, x3;
    if (RESID0) {
      x0 = op.w * __ldg(op.dinv + c0) * __ldcg(op.b + c0);
      x1 = op.w * __ldg(op.dinv + c1) * __ldcg(op.b + c1);
      x2 = op.w * __ldg(op.dinv + c2) * __ldcg(op.b + c2);
      x3 = op.w * __ldg(op.dinv + c3) * __ldcg(op.b + c3);
    } else {
      x0 = __ldcg(op.x + c0);
      x1 = __ldcg(op.x + c1);
      x2 = __ldcg(op.x + c2);
      x3 = __ldcg(op.x + c3);
    }
    s0 += v0 * x0;
    s1 += v1 * x1;
    s2 += v2 * x2;
    s3 += v3 * x3;
  }
  for (; j < r1; j += lanes) {
    const int c = __ldg(op.indices + j);
    s0 += __ldg(op.vals + j) * (RESID0 ? op.w * __ldg(op.dinv + c) * __ldcg(op.b + c) : __ldcg(op.x + c));
  }
  return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) amg_tail_kernel(const TailOp* __restrict__ ops, int nops, unsigned* bar) {
  const int tid = threadIdx.x;
  const int gtid = blockIdx.x * TAIL_THREADS + tid;
  const int gthreads = gridDim.x * TAIL_THREADS;
  for (int o = 0; o < nops; ++o) {
    const TailOp op = ops[o];
    if (op.type == TAIL_SPMV) {
      const int lanes = op.lanes;
      const int lane = tid & (lanes - 1);
      const int sub = gtid / lanes, nsub = gthreads / lanes;
      for (int base = 0; base < op.n; base += nsub) {          // uniform trip count: the sub-warp shuffles stay convergent
        const int row = base + sub;
        const bool valid = row < op.n;
        double sum = 0.0;
        if (valid) {
          const int r0 = __ldg(op.indptr + row), r1 = __ldg(op.indptr + row + 1);
          sum = op.epi == EPI_RESID0 ? tail_row_sum<true>(op, r0, r1, lane, lanes) : tail_row_sum<false>(op, r0, r1, lane, lanes);
        }
        for (int off = lanes >> 1; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off, lanes);
        if (valid && lane == 0) {
          double r;
          switch (op.epi) {
            case EPI_SET: r = sum; break;
            case EPI_RESID: r = __ldcg(op.b + row) - sum; break;
            case EPI_JACOBI: r = __ldcg(op.x + row) + op.w * __ldg(op.dinv + row) * (__ldcg(op.b + row) - sum); break;
            case EPI_ADD: r = __ldcg(op.out + row) + sum; break;
            default: {   // EPI_RESID0: x = w dinv b (written to out2), r = b - A x
              const double bi = __ldcg(op.b + row);
              op.out2[row] = op.w * __ldg(op.dinv + row) * bi;
              r = bi - sum;
            }
          }
          op.out[row] = r;
        }
      }
    } else if (op.type == TAIL_DENSE) {
      // x = Minv b, dense row-major n x n: one warp per row, eight loads in flight per lane
      const int lane = tid & 31;
      const int wrp = gtid >> 5, nw = gthreads >> 5;
      for (int row = wrp; row < op.n; row += nw) {
        const double* m = op.vals + (size_t)row * op.n;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int j = lane;
        for (; j + 224 < op.n; j += 256) {
          const double m0 = __ldg(m + j), m1 = __ldg(m + j + 32), m2 = __ldg(m + j + 64), m3 = __ldg(m + j + 96);
          const double m4 = __ldg(m + j + 128), m5 = __ldg(m + j + 160), m6 = __ldg(m + j + 192), m7 = __ldg(m + j + 224);
          a0 += m0 * __ldcg(op.x + j) + m4 * __ldcg(op.x + j + 128);
          a1 += m1 * __ldcg(op.x + j + 32) + m5 * __ldcg(op.x + j + 160);
          a2 += m2 * __ldcg(op.x + j + 64) + m6 * __ldcg(op.x + j + 192);
          a3 += m3 * __ldcg(op.x + j + 96) + m7 * __ldcg(op.x + j + 224);
        }
        for (; j < op.n; j += 32) a0 += __ldg(m + j) * __ldcg(op.x + j);
        double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if (lane == 0) op.out[row] = acc;
      }
    } else {   // TAIL_SCALE: out = w dinv b
      for (int i = gtid; i < op.n; i += gthreads) op.out[i] = op.w * __ldg(op.dinv + i) * __ldcg(op.b + i);
    }
    if (o + 1 < nops) {
      __threadfence();
      tail_grid_barrier(bar, gridDim.x);
    }
  }
}

int tail_grid_size() {
  static const int g = getenv("KNP_TAIL_GRID") ? atoi(getenv("KNP_TAIL_GRID")) : 148;
  return g < 1 ? 1 : (g > 148 ? 148 : g);      // one CTA per SM at most: every CTA is resident, the barrier cannot deadlock
}

int launch_amg_tail(const TailOp* ops_dev, int nops, unsigned* bar, cudaStream_t st) {
  if (nops == 0) return KNP_OK;
  amg_tail_kernel<<<tail_grid_size(), TAIL_THREADS, 0, st>>>(ops_dev, nops, bar);
  KNP_LAUNCHED();
  return KNP_OK;
}

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double block_reduce_256(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[i];
  }
  __syncthreads();
  return r;
}

constexpr int MD_CHUNK = 16;   // basis vectors handled per pass over w

// partial[(j)*RED_BLOCKS + blk] ; j in [j0, j0+cnt) for V rows, and slot m for ||w||^2 when with_norm.
// jsub >= 0: the dots are taken with d = w - V[jsub] formed on the fly (gmres_solve orthogonalises (B A - I) v_j, see there).
__global__ void __launch_bounds__(256) multi_dot_kernel(int n, int j0, int cnt, int m, int with_norm,
                                                        const double* __restrict__ V, size_t ldv,
                                                        const double* __restrict__ w, double* __restrict__ partial, int jsub) {
  __shared__ double red[8];
  double acc[MD_CHUNK + 1];
#pragma unroll
  for (int j = 0; j <= MD_CHUNK; ++j) acc[j] = 0.0;
  // Block b takes the 256-element chunks b, b + G, b + 2G, ... (G = gridDim.x, a compile-time constant of the launch): the
  // summation order is fixed whatever the launch schedule, and at any moment the grid works on ONE window of every row,
  // i.e. ~18 sequential DRAM streams instead of 18 per block (the former slab-per-block split kept 2e4 streams open and
  // reached 2.6 TB/s)
  const double* __restrict__ vsub = jsub >= 0 ? V + (size_t)jsub * ldv : nullptr;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const double wi = vsub ? w[i] - vsub[i] : w[i];
#pragma unroll
    for (int j = 0; j < MD_CHUNK; ++j)
      if (j < cnt) acc[j] += V[(size_t)(j0 + j) * ldv + i] * wi;
    acc[MD_CHUNK] += wi * wi;
  }
#pragma unroll
  for (int j = 0; j < MD_CHUNK; ++j) {
    if (j < cnt) {
      const double r = block_reduce_256(acc[j], red);
      if (threadIdx.x == 0) partial[(size_t)(j0 + j) * RED_BLOCKS + blockIdx.x] = r;
    }
  }
  if (with_norm) {
    const double r = block_reduce_256(acc[MD_CHUNK], red);
    if (threadIdx.x == 0) partial[(size_t)m * RED_BLOCKS + blockIdx.x] = r;
  }
}

// out[j] = sum_blk partial[j*RED_BLOCKS + blk], fixed order; one warp per j
__global__ void reduce_rows_kernel(int rows, int n_partial, const double* __restrict__ partial,
                                   double* __restrict__ out) {
  const int j = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (j >= rows) return;
  double s = 0.0;
  for (int i = lane; i < n_partial; i += 32) s += partial[(size_t)j * n_partial + i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) out[j] = s;
}

int launch_multi_dot(int n, int m, const double* V, size_t ldv, const double* w, double* partial, double* out,
                     cudaStream_t st, int jsub) {
  if (m == 0) {
    multi_dot_kernel<<<RED_BLOCKS, 256, 0, st>>>(n, 0, 0, 0, 1, V, ldv, w, partial, jsub);
    KNP_LAUNCHED();
  }
  for (int j0 = 0; j0 < m; j0 += MD_CHUNK) {
    const int cnt = m - j0 < MD_CHUNK ? m - j0 : MD_CHUNK;
    multi_dot_kernel<<<RED_BLOCKS, 256, 0, st>>>(n, j0, cnt, m, j0 == 0 ? 1 : 0, V, ldv, w, partial, jsub);
    KNP_LAUNCHED();
  }
  reduce_rows_kernel<<<(m + 1 + 7) / 8, 256, 0, st>>>(m + 1, RED_BLOCKS, partial, out);
  KNP_LAUNCHED();
  return KNP_OK;
}

// w += sign * (scale .*) sum_j h[j] V_j ; with vnext: vnext = w_new * inv_norm as well (the next Krylov basis vector in the
// same pass, when the norm after the projection is already known from ||w||^2 - sum h^2)
__global__ void __launch_bounds__(256) multi_axpy_kernel(int n, int m, const double* __restrict__ V, size_t ldv,
                                                         const double* __restrict__ h, double sign,
                                                         const double* __restrict__ scale, double* __restrict__ w,
                                                         double* __restrict__ vnext, double inv_norm, int jsub) {
  __shared__ double hs[64];
  if (threadIdx.x < m) hs[threadIdx.x] = h[threadIdx.x];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double acc = 0.0, sub = 0.0;
    for (int j = 0; j < m; ++j) {
      const double vj = V[(size_t)j * ldv + i];
      acc += hs[j] * vj;
      if (j == jsub) sub = vj;                    // jsub >= 0: the projected vector is d = w - V[jsub]
    }
    const double wi = (w[i] - sub) + sign * (scale ? scale[i] * acc : acc);
    w[i] = wi;
    if (vnext) vnext[i] = wi * inv_norm;
  }
}
int launch_multi_axpy(int n, int m, const double* V, size_t ldv, const double* h, double* w, cudaStream_t st, int jsub) {
  if (n == 0 || m == 0) return KNP_OK;
  multi_axpy_kernel<<<grid_for(n), 256, 0, st>>>(n, m, V, ldv, h, -1.0, nullptr, w, nullptr, 0.0, jsub);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_multi_axpy_normalize(int n, int m, const double* V, size_t ldv, const double* h, double* w, double* vnext,
                                double inv_norm, cudaStream_t st, int jsub) {
  if (n == 0 || m == 0) return KNP_OK;
  multi_axpy_kernel<<<grid_for(n), 256, 0, st>>>(n, m, V, ldv, h, -1.0, nullptr, w, vnext, inv_norm, jsub);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_update_x(int n, int m, const double* V, size_t ldv, const double* y_dev, double* x, const double* scale,
                    cudaStream_t st) {
  if (n == 0 || m == 0) return KNP_OK;
  multi_axpy_kernel<<<grid_for(n), 256, 0, st>>>(n, m, V, ldv, y_dev, 1.0, scale, x, nullptr, 0.0, -1);
  KNP_LAUNCHED();
  return KNP_OK;
}

__global__ void pointwise_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, int divide,
                                 double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = divide ? a[i] / b[i] : a[i] * b[i];
}
int launch_pointwise(int n, const double* a, const double* b, int divide, double* out, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  pointwise_kernel<<<grid_for(n), 256, 0, st>>>(n, a, b, divide, out);
  KNP_LAUNCHED();
  return KNP_OK;
}

__global__ void axpby_kernel(int n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    y[i] = a * x[i] + (b == 0.0 ? 0.0 : b * y[i]);
}
int launch_axpby(int n, double a, const double* x, double b, double* y, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  axpby_kernel<<<grid_for(n), 256, 0, st>>>(n, a, x, b, y);
  KNP_LAUNCHED();
  return KNP_OK;
}

__global__ void scale_copy_kernel(int n, const double* __restrict__ alpha, int invert, const double* __restrict__ x,
                                  double* __restrict__ y) {
  const double a = invert ? 1.0 / sqrt(alpha[0]) : alpha[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i];
}
// invert=1: y = x / sqrt(alpha[0])  (normalise by a squared norm held on the device)
int launch_scale_copy(int n, const double* alpha_dev, int invert, const double* x, double* y, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  scale_copy_kernel<<<grid_for(n), 256, 0, st>>>(n, alpha_dev, invert, x, y);
  KNP_LAUNCHED();
  return KNP_OK;
}

__global__ void __launch_bounds__(256) range_sum_kernel(const double* __restrict__ x, int lo0, int hi0, int lo1,
                                                        int hi1, double* __restrict__ partial) {
  __shared__ double red[8];
  const int n0 = hi0 - lo0, n = n0 + (hi1 - lo1);
  double acc = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) acc += x[i < n0 ? lo0 + i : lo1 + (i - n0)];
  const double r = block_reduce_256(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
int launch_range_sum(const double* x, int lo0, int hi0, int lo1, int hi1, double* partial, double* out,
                     cudaStream_t st) {
  range_sum_kernel<<<RED_BLOCKS, 256, 0, st>>>(x, lo0, hi0, lo1, hi1, partial);
  KNP_LAUNCHED();
  reduce_rows_kernel<<<1, 32, 0, st>>>(1, RED_BLOCKS, partial, out);
  KNP_LAUNCHED();
  return KNP_OK;
}
__global__ void range_shift_kernel(double* __restrict__ x, int lo0, int hi0, int lo1, int hi1,
                                   const double* __restrict__ sum, double inv_count) {
  const int n0 = hi0 - lo0, n = n0 + (hi1 - lo1);
  const double sh = sum[0] * inv_count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    x[i < n0 ? lo0 + i : lo1 + (i - n0)] -= sh;
}
int launch_range_shift(double* x, int lo0, int hi0, int lo1, int hi1, const double* sum_dev, double inv_count,
                       cudaStream_t st) {
  const int n = (hi0 - lo0) + (hi1 - lo1);
  if (n == 0) return KNP_OK;
  range_shift_kernel<<<grid_for(n), 256, 0, st>>>(x, lo0, hi0, lo1, hi1, sum_dev, inv_count);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_reduce_partials(const double* partial, int n_partial, double* out, cudaStream_t st) {
  reduce_rows_kernel<<<1, 32, 0, st>>>(1, n_partial, partial, out);
  KNP_LAUNCHED();
  return KNP_OK;
}

// y[rows[i]] += vals[i] (unique rows): time-independent source entries of the right-hand side (knp_set_source)
__global__ void add_sparse_kernel(int n, const int32_t* __restrict__ rows, const double* __restrict__ vals, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[rows[i]] += vals[i];
}
int launch_add_sparse(int n, const int32_t* rows, const double* vals, double* y, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  add_sparse_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, rows, vals, y);
  KNP_LAUNCHED();
  return KNP_OK;
}

// ------------------------------------------------------------------------------------------------ Dirichlet conditions
// Essential boundary conditions the way assemble_matrix_block / assemble_vector_block apply them (KNPEMIx_solver.py:113-116
// with bcs = p.bcs, KNPEMIx_problem.py:96-198): rows and columns of constrained dofs are zeroed, their diagonal set to `diag`,
// the right-hand side lifted, b_i -= sum_j A_ij g_j, and b = g on the constrained rows.  bc_cols (ascending, column layout:
// owned and ghost columns) / bc_vals hold the constrained dofs and their values; `rows` lists the owned rows with at least one
// constrained entry (found once by bc_touch_kernel), so the pass costs O(boundary), not O(nnz).  One thread per listed row
// walks it in ascending order: fixed summation order, every entry written by one thread.
__device__ __forceinline__ int bc_find(const int32_t* __restrict__ cols, int n, int c) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cols[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  return (lo < n && cols[lo] == c) ? lo : -1;
}
__global__ void bc_touch_kernel(int n_rows, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                const uint8_t* __restrict__ flag, uint8_t* __restrict__ touched) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  uint8_t t = flag[r];
  for (int k = indptr[r]; k < indptr[r + 1] && !t; ++k) t = flag[indices[k]];
  touched[r] = t;
}
__global__ void bc_flag_kernel(int n, const int32_t* __restrict__ cols, uint8_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[cols[i]] = 1;
}
__global__ void bc_apply_kernel(int n_list, const int32_t* __restrict__ rows, const int32_t* __restrict__ indptr,
                                const int32_t* __restrict__ indices, double* __restrict__ vals, double* __restrict__ b,
                                int n_bc, const int32_t* __restrict__ bc_cols, const double* __restrict__ bc_vals,
                                double diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const int r = rows[i];
  const int self = bc_find(bc_cols, n_bc, r);
  double lift = 0.0;
  for (int k = indptr[r]; k < indptr[r + 1]; ++k) {
    const int c = indices[k];
    if (self >= 0) {
      vals[k] = c == r ? diag : 0.0;
    } else {
      const int j = bc_find(bc_cols, n_bc, c);
      if (j >= 0) {
        lift += vals[k] * bc_vals[j];
        vals[k] = 0.0;
      }
    }
  }
  if (b) b[r] = self >= 0 ? bc_vals[self] : b[r] - lift;
}
__global__ void bc_set_kernel(int n, const int32_t* __restrict__ cols, const double* __restrict__ vals, int n_rows,
                              double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cols[i] < n_rows) x[cols[i]] = vals[i];
}
int launch_bc_set(int n_bc, const int32_t* bc_cols, const double* bc_vals, int n_rows, double* x, cudaStream_t st) {
  if (n_bc == 0) return KNP_OK;
  bc_set_kernel<<<(n_bc + 255) / 256, 256, 0, st>>>(n_bc, bc_cols, bc_vals, n_rows, x);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_bc_flags(int n_bc, const int32_t* bc_cols, uint8_t* flag, cudaStream_t st) {
  if (n_bc == 0) return KNP_OK;
  bc_flag_kernel<<<(n_bc + 255) / 256, 256, 0, st>>>(n_bc, bc_cols, flag);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_bc_touch(int n_rows, const int32_t* indptr, const int32_t* indices, const uint8_t* flag, uint8_t* touched,
                    cudaStream_t st) {
  if (n_rows == 0) return KNP_OK;
  bc_touch_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(n_rows, indptr, indices, flag, touched);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_bc_apply(int n_list, const int32_t* rows, const int32_t* indptr, const int32_t* indices, double* vals, double* b,
                    int n_bc, const int32_t* bc_cols, const double* bc_vals, double diag, cudaStream_t st) {
  if (n_list == 0) return KNP_OK;
  bc_apply_kernel<<<(n_list + 127) / 128, 128, 0, st>>>(n_list, rows, indptr, indices, vals, b, n_bc, bc_cols, bc_vals, diag);
  KNP_LAUNCHED();
  return KNP_OK;
}

// ------------------------------------------------------------------------------------------------ conjugate gradients
// Device-resident scalars of the preconditioned CG loop (solver.cu::cg_solve): S[0] = (r, z), S[1] = alpha, S[2] = beta,
// S[3] = state (0 running, 1 converged, 2 breakdown: (p, A p) <= 0 or non-finite), S[4] = tol^2, S[5] = iteration at which the
// state left 0.  Every update kernel is a no-op once the state is non-zero, so the iterate stays frozen at the converged
// step while the host catches up (it reads S back only every few iterations).
__global__ void cg_scalar_kernel(int phase, int it, const double* __restrict__ dots, double* __restrict__ S,
                                 double* __restrict__ hist) {
  if (phase == 0) {                       // after (r0, z0), ||z0||^2
    S[0] = dots[0];
    S[3] = 0.0;
    S[5] = 0.0;
    hist[0] = dots[1];
    if (!(dots[1] == dots[1])) S[3] = 2.0;
    else if (dots[1] <= S[4]) S[3] = 1.0;
  } else if (S[3] == 0.0) {
    if (phase == 1) {                     // after (p, A p)
      const double pq = dots[0];
      if (!(pq > 0.0)) {
        S[3] = 2.0;
        S[5] = (double)it;
      } else {
        S[1] = S[0] / pq;
      }
    } else {                              // after (r, z), ||z||^2 of the new residual
      S[2] = dots[0] / S[0];
      S[0] = dots[0];
      hist[it] = dots[1];
      if (!(dots[1] == dots[1])) {
        S[3] = 2.0;
        S[5] = (double)it;
      } else if (dots[1] <= S[4]) {
        S[3] = 1.0;
        S[5] = (double)it;
      }
    }
  }
}
int launch_cg_scalar(int phase, int it, const double* dots, double* S, double* hist, cudaStream_t st) {
  cg_scalar_kernel<<<1, 1, 0, st>>>(phase, it, dots, S, hist);
  KNP_LAUNCHED();
  return KNP_OK;
}
// x += alpha p ; r -= alpha q   (alpha = S[1])
__global__ void cg_xr_kernel(int n, const double* __restrict__ S, const double* __restrict__ p, const double* __restrict__ q,
                             double* __restrict__ x, double* __restrict__ r) {
  if (S[3] != 0.0) return;
  const double a = S[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] += a * p[i];
    r[i] -= a * q[i];
  }
}
// p = z + beta p   (beta = S[2])
__global__ void cg_p_kernel(int n, const double* __restrict__ S, const double* __restrict__ z, double* __restrict__ p) {
  if (S[3] != 0.0) return;
  const double b = S[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = z[i] + b * p[i];
}
int launch_cg_xr(int n, const double* S, const double* p, const double* q, double* x, double* r, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  cg_xr_kernel<<<grid_for(n), 256, 0, st>>>(n, S, p, q, x, r);
  KNP_LAUNCHED();
  return KNP_OK;
}
int launch_cg_p(int n, const double* S, const double* z, double* p, cudaStream_t st) {
  if (n == 0) return KNP_OK;
  cg_p_kernel<<<grid_for(n), 256, 0, st>>>(n, S, z, p);
  KNP_LAUNCHED();
  return KNP_OK;
}

}  // namespace knp
