// Kernels of the P2 element path (fem_order = 2): thin per-thread wrappers around the bodies of p2.cuh.
//   p2_facet_kernel   membrane-facet element tensors with the P2 trace basis (3 / 6 dofs per facet)
//   p2_rows_kernel    one thread per restricted dof: its four matrix rows and right-hand side entries
//   p2_l2_kernel      int u^power over tagged cells (L2 norms, ion amounts, measures)
//   p2_stim_kernel    total stimulus current
#include <cstdlib>
#include "p2.cuh"

#ifndef P2_DEFAULT_MINB
#define P2_DEFAULT_MINB 4
#endif

namespace knp {

template <int D>
__global__ void __launch_bounds__(64) p2_facet_kernel(P2View V, KParams P, const uint32_t* __restrict__ tag_models,
                                                      const int32_t* __restrict__ tag_stim, const double* __restrict__ u,
                                                      const double* __restrict__ gates, double stim_fac,
                                                      double* __restrict__ fe) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < V.n_mf) p2_facet_body<D>(V, P, tag_models, tag_stim, u, gates, stim_fac, fe, f);
}

// MINB = resident CTAs per SM the register allocation is capped for: 4 leaves 255 registers (the 80 accumulators of a 3D row
// stay in registers, 8 warps per SM), 6 caps at 168 (12 warps per SM, part of the accumulators in L1-backed local memory);
// KNP_P2_MINB selects, the default is the faster one measured on B200 (profiles/r02p_p2.md)
template <int D, int MODE, int MINB>
__global__ void __launch_bounds__(64, MINB) p2_rows_kernel(P2View V, P2Coef C, const double* __restrict__ u,
                                                           const double* __restrict__ fe, double* vals, double* bvec) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < V.n_work) p2_row_body<D, MODE>(V, C, u, fe, vals, bvec, w);
}

template <int D>
__global__ void __launch_bounds__(256) p2_l2_kernel(P2View V, int s, int field, int power, int n_cells,
                                                    const int32_t* __restrict__ cell_tag,
                                                    const int32_t* __restrict__ cell_owned,
                                                    const int32_t* __restrict__ tags, int n_tags,
                                                    const double* __restrict__ u, double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
    if (!cell_owned[c]) continue;
    const int t = cell_tag[c];
    bool hit = false;
    for (int i = 0; i < n_tags; ++i) hit |= (tags[i] == t);
    if (hit) acc += p2_cell_integral<D>(V, s, field, power, u, c);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

template <int D>
__global__ void __launch_bounds__(256) p2_stim_kernel(P2View V, KParams P, const int32_t* __restrict__ tag_stim,
                                                      const int32_t* __restrict__ mf_owned, const double* __restrict__ u,
                                                      double stim_fac, double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < V.n_mf; f += gridDim.x * blockDim.x) {
    if (!mf_owned[f] || tag_stim[V.mf_tagidx[f]] == 0) continue;
    acc += p2_facet_stim_current<D>(V, P, u, stim_fac, f);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

int launch_facets_p2(const P2View& V, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim, const double* u,
                     const double* gates, double stim_fac, double* fe, cudaStream_t st) {
  if (V.n_mf == 0) return KNP_OK;
  const int grid = (V.n_mf + 63) / 64;
  if (V.gdim == 2) p2_facet_kernel<2><<<grid, 64, 0, st>>>(V, P, tag_models, tag_stim, u, gates, stim_fac, fe);
  else p2_facet_kernel<3><<<grid, 64, 0, st>>>(V, P, tag_models, tag_stim, u, gates, stim_fac, fe);
  KNP_LAUNCHED();
  return KNP_OK;
}

int launch_rows_p2(const P2View& V, const KParams& P, int mode, const double* u, const double* fe, double* vals, double* b,
                   cudaStream_t st) {
  if (V.n_work == 0) return KNP_OK;
  const P2Coef C = p2_coef(P);
  const int grid = (V.n_work + 63) / 64;
  static const int minb = getenv("KNP_P2_MINB") ? atoi(getenv("KNP_P2_MINB")) : P2_DEFAULT_MINB;
  if (minb >= 6) {
    if (V.gdim == 2) {
      if (mode == 0) p2_rows_kernel<2, 0, 6><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
      else p2_rows_kernel<2, 1, 6><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
    } else {
      if (mode == 0) p2_rows_kernel<3, 0, 6><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
      else p2_rows_kernel<3, 1, 6><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
    }
  } else {
    if (V.gdim == 2) {
      if (mode == 0) p2_rows_kernel<2, 0, 4><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
      else p2_rows_kernel<2, 1, 4><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
    } else {
      if (mode == 0) p2_rows_kernel<3, 0, 4><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
      else p2_rows_kernel<3, 1, 4><<<grid, 64, 0, st>>>(V, C, u, fe, vals, b);
    }
  }
  KNP_LAUNCHED();
  return KNP_OK;
}

int launch_l2_cells_p2(const P2View& V, int s, int field, int power, int n_cells, const int32_t* cell_tag,
                       const int32_t* cell_owned, const int32_t* tags, int n_tags, const double* u, double* partial,
                       int n_partial, cudaStream_t st) {
  if (V.gdim == 2) p2_l2_kernel<2><<<n_partial, 256, 0, st>>>(V, s, field, power, n_cells, cell_tag, cell_owned, tags, n_tags, u, partial);
  else p2_l2_kernel<3><<<n_partial, 256, 0, st>>>(V, s, field, power, n_cells, cell_tag, cell_owned, tags, n_tags, u, partial);
  KNP_LAUNCHED();
  return KNP_OK;
}

int launch_stim_current_p2(const P2View& V, const KParams& P, const int32_t* tag_stim, const int32_t* mf_owned, const double* u,
                           double stim_fac, double* partial, int n_partial, cudaStream_t st) {
  if (V.gdim == 2) p2_stim_kernel<2><<<n_partial, 256, 0, st>>>(V, P, tag_stim, mf_owned, u, stim_fac, partial);
  else p2_stim_kernel<3><<<n_partial, 256, 0, st>>>(V, P, tag_stim, mf_owned, u, stim_fac, partial);
  KNP_LAUNCHED();
  return KNP_OK;
}

}  // namespace knp
