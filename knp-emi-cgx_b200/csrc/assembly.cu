// Per-timestep assembly kernels for sm_100a.
//
//   gate_kernel    K4  Rush-Larsen / forward-Euler gate ODE on membrane vertices
//                      (HodgkinHuxley.update_gating_variables, KNPEMIx_ionic_model.py:605-671)
//   facet_kernel   K3  membrane-facet element tensors: alpha_k, Nernst potentials, channel currents at
//                      the facet quadrature points (dS terms of KNPEMIx_problem.py:594-642 with the
//                      IonicModel._eval family), written to a facet-major SoA staging buffer
//   rows_kernel    K1/K2  one CTA per tile of restricted dofs ("nodes"), one thread per (node, adjacency slot):
//                      neighbour data staged once in shared memory, the P1 element rows recomputed from the
//                      coordinates per (node, cell), fixed-order accumulation in registers (no atomics), membrane
//                      facet rows added, finished CSR rows streamed out through a staging strip
//                      (every A value and b entry is written exactly once -> bitwise reproducible).
//   csr_indices_kernel   column indices of A / P from the node adjacency (setup)
//
// Design note: all ten (d+1)x(d+1) cell blocks of KNPEMIx_problem.py:598-605,633-634 are linear
// combinations of M^T, K^T and cbar_k K^T, and each block row shares the node's adjacency list, so the
// "cell -> nnz map" collapses to one byte per (node, cell, local vertex): the adjacency slot.
#include <algorithm>
#include <cstdlib>
#include <string>
#include "common.cuh"
#include "kernels.cuh"

namespace knp {

__device__ __forceinline__ int symidx(int a, int b, int D) {
  // a <= b ; row-major upper triangle
  return a * D - (a * (a - 1)) / 2 + (b - a);
}

// ------------------------------------------------------------------------------------------------ gates
__global__ void gate_kernel(DevTopo T, KParams P, const double* __restrict__ u, double* __restrict__ gates) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= T.n_mv) return;
  const double phim = u[T.L.col(0, 3, T.mv_node0[g])] - u[T.L.col(1, 3, T.mv_node1[g])];
  const double V = 1000.0 * (phim - P.phi_rest);
  double al[3], be[3];
  al[0] = 0.01e3 * (10.0 - V) / (exp((10.0 - V) / 10.0) - 1.0);
  be[0] = 0.125e3 * exp(-V / 80.0);
  al[1] = 0.1e3 * (25.0 - V) / (exp((25.0 - V) / 10.0) - 1.0);
  be[1] = 4.0e3 * exp(-V / 18.0);
  al[2] = 0.07e3 * exp(-V / 20.0);
  be[2] = 1.0e3 / (exp((30.0 - V) / 10.0) + 1.0);
  const double dt_ode = P.dt / P.ode_substeps;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double y = gates[(size_t)j * T.n_mv + g];
    if (P.rush_larsen) {
      const double tau = 1.0 / (al[j] + be[j]);
      const double yinf = al[j] * tau;
      const double yexp = exp(-dt_ode / tau);
      for (int it = 0; it < P.ode_substeps; ++it) y = yinf + (y - yinf) * yexp;
    } else {
      const double aa = al[j] * dt_ode, bb = be[j] * dt_ode;
      for (int it = 0; it < P.ode_substeps; ++it) y = y + (aa * (1.0 - y) - bb * y);
    }
    gates[(size_t)j * T.n_mv + g] = y;
  }
}

// ------------------------------------------------------------------------------------------------ facets
// Staging layout (component-major, facet fastest):
//   GA  : ((s*3+k)*NS + ab)            6*NS      NS = D(D+1)/2
//   bc  : 6*NS + (s*3+k)*D + a         6*D       already divided by F z_k
//   bphi: 6*NS + 6*D + a               D         already divided by F
template <int D>
__global__ void __launch_bounds__(128) facet_kernel(DevTopo T, KParams P, const uint32_t* __restrict__ tag_models,
                                                    const int32_t* __restrict__ tag_stim,
                                                    const double* __restrict__ u, const double* __restrict__ gates,
                                                    double stim_fac, double* __restrict__ fe) {
  constexpr int NS = D * (D + 1) / 2;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= T.n_mf) return;
  double ci[3][D], ce[3][D], pm[D], gn[D], gm[D], gh[D];
  int nodei[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    const int g = T.mf_mv[(size_t)f * D + a];
    const int qi = T.mv_node0[g], qe = T.mv_node1[g];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ci[k][a] = u[T.L.col(0, k, qi)];
      ce[k][a] = u[T.L.col(1, k, qe)];
    }
    pm[a] = u[T.L.col(0, 3, qi)] - u[T.L.col(1, 3, qe)];
    gn[a] = gates[g];
    gm[a] = gates[(size_t)T.n_mv + g];
    gh[a] = gates[(size_t)2 * T.n_mv + g];
    nodei[a] = qi;
  }
  const double area = T.mf_area[f];
  const int ti = T.mf_tagidx[f];
  const uint32_t models = tag_models[ti];
  const bool stim_on = tag_stim[ti] != 0;

  double GA[2][3][NS], bc[2][3][D], bphi[D];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int i = 0; i < NS; ++i) GA[s][k][i] = 0.0;
#pragma unroll
      for (int a = 0; a < D; ++a) bc[s][k][a] = 0.0;
    }
#pragma unroll
  for (int a = 0; a < D; ++a) bphi[a] = 0.0;

  const double psi = P.psi;
  for (int q = 0; q < T.nq; ++q) {
    double lam[D];
#pragma unroll
    for (int a = 0; a < D; ++a) lam[a] = T.qb[q * D + a];
    const double w = area * T.qw[q];
    double ciq[3], ceq[3], pmq = 0.0, nq = 0.0, mq = 0.0, hq = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ciq[k] = 0.0;
      ceq[k] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        ciq[k] += lam[a] * ci[k][a];
        ceq[k] += lam[a] * ce[k][a];
      }
      pmq += lam[a] * pm[a];
      nq += lam[a] * gn[a];
      mq += lam[a] * gm[a];
      hq += lam[a] * gh[a];
    }
    // alpha_{k,s} (KNPEMIx_problem.py:512-513,582-583)
    double al[2][3];
    {
      double di = 0.0, de = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        di += P.D[k] * P.z[k] * P.z[k] * ciq[k];
        de += P.D[k] * P.z[k] * P.z[k] * ceq[k];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        al[0][k] = P.D[k] * P.z[k] * P.z[k] * ciq[k] / di;
        al[1][k] = P.D[k] * P.z[k] * P.z[k] * ceq[k] / de;
      }
    }
    // Nernst potentials (KNPEMIx_problem.py:516)
    double E[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) E[k] = (psi / P.z[k]) * log(ceq[k] / ciq[k]);
    double I[3] = {0.0, 0.0, 0.0};
    if (models & KNP_MODEL_NEURONAL_CT) {      // KNPEMIx_ionic_model.py:342-369 (f_NKCC1 == 0, :50-75)
      const double I_KCC2 = 0.0068 * log((ciq[1] * ciq[2]) / (ceq[1] * ceq[2]));
      I[1] += I_KCC2;
      I[2] += -I_KCC2;
    }
    if (models & KNP_MODEL_HH) {               // :487-515 (+ stimulus :517-603)
      const double gNa = P.g_leak[0] + P.g_Na_bar * mq * mq * mq * hq;
      const double gK = P.g_leak[1] + P.g_K_bar * (nq * nq) * (nq * nq);
      double INa = gNa * (pmq - E[0]);
      if (stim_on) {
        // mask = prod_i [lo_i < x_{dir_i} < hi_i] at the quadrature point (:558-587, incl. `multiple` directions)
        double mask = 1.0;
        for (int i = 0; i < 3 && P.stim_dir[i] >= 0; ++i) {
          double xq = 0.0;
#pragma unroll
          for (int a = 0; a < D; ++a) xq += lam[a] * T.node_x[(size_t)nodei[a] * D + P.stim_dir[i]];
          mask *= (xq > P.stim_lo[i] && xq < P.stim_hi[i]) ? 1.0 : 0.0;
        }
        INa += mask * stim_fac * (pmq - E[0]);
      }
      I[0] += INa;
      I[1] += gK * (pmq - E[1]);
      I[2] += P.g_leak[2] * (pmq - E[2]);
    }
    if (models & KNP_MODEL_ATP) {              // :385-422
      const double p1 = 1.0 + 1.5 / ceq[1];
      const double p2 = 1.0 + 10.0 / ciq[0];
      const double I_ATP = 0.25 / ((p1 * p1) * (p2 * p2 * p2));
      I[0] += 3.0 * I_ATP;
      I[1] += -2.0 * I_ATP;
    }
    if (models & KNP_MODEL_GLIAL_CT) {         // :239-298 (f_NKCC1 == 0)
      const double I_KCC1 = (7e-2 * psi) * log((ciq[1] * ciq[2]) / (ceq[1] * ceq[2]));
      I[1] += I_KCC1;
      I[2] += -I_KCC1;
    }
    if (models & KNP_MODEL_KIRNA) {            // :117-222
      const double E_K_init = psi * log(P.K_e_init / P.K_i_g_init);
      const double rho = 1.1 * 1.12e-6;
      const double r = 10.0 / ciq[0];
      const double pump = (1.0 / (1.0 + r * sqrt(r))) * (1.0 / (1.0 + 1.5 / ceq[1])) * rho;
      const double A_ = 1.0 + exp(0.433);
      const double B_ = 1.0 + exp(-(0.1186 + E_K_init) / 0.0441);
      const double C_ = 1.0 + exp(((pmq - E[1]) + 0.0185) / 0.0425);
      const double D_ = 1.0 + exp(-(0.1186 + pmq) / 0.0441);
      const double f_kir = sqrt(ceq[1] / P.K_e_init) * A_ * B_ / (C_ * D_);
      I[0] += P.g_leak_g[0] * (pmq - E[0]) + 3.0 * P.z[0] * P.F * pump;
      I[1] += f_kir * P.g_leak_g[1] * (pmq - E[1]) - 2.0 * P.z[1] * P.F * pump;
      I[2] += P.g_leak_g[2] * (pmq - E[2]);
    }
    if (models & KNP_MODEL_PASSIVE) {          // :89-91
      I[0] += pmq;
      I[1] += pmq;
      I[2] += pmq;
    }
    const double Itot = (I[0] + I[1]) + I[2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double wa = w * al[s][k];
        const double rb = w * (P.dt * I[k] - al[s][k] * P.C_M * pmq);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          bc[s][k][a] += rb * lam[a];
#pragma unroll
          for (int b = a; b < D; ++b) GA[s][k][a * D - (a * (a - 1)) / 2 + (b - a)] += wa * (lam[a] * lam[b]);
        }
      }
    const double rp = w * (P.dt * Itot - P.C_M * pmq);
#pragma unroll
    for (int a = 0; a < D; ++a) bphi[a] += rp * lam[a];
  }
  const size_t nf = (size_t)T.n_mf;
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int i = 0; i < NS; ++i) fe[(size_t)((s * 3 + k) * NS + i) * nf + f] = GA[s][k][i];
      const double inv = 1.0 / (P.F * P.z[k]);
#pragma unroll
      for (int a = 0; a < D; ++a) fe[(size_t)(6 * NS + (s * 3 + k) * D + a) * nf + f] = bc[s][k][a] * inv;
    }
#pragma unroll
  for (int a = 0; a < D; ++a) fe[(size_t)(6 * NS + 6 * D + a) * nf + f] = bphi[a] / P.F;
}

// ------------------------------------------------------------------------------------------------ rows
template <int D>
struct CellGeom {
  double vol;
  double g[D + 1][D];
};

__device__ __forceinline__ void cell_geometry(const double (&x)[3][2], CellGeom<2>& G) {
  const double e1x = x[1][0] - x[0][0], e1y = x[1][1] - x[0][1];
  const double e2x = x[2][0] - x[0][0], e2y = x[2][1] - x[0][1];
  const double det = e1x * e2y - e1y * e2x;
  const double inv = 1.0 / det;
  G.vol = 0.5 * fabs(det);
  G.g[1][0] = e2y * inv;
  G.g[1][1] = -e2x * inv;
  G.g[2][0] = -e1y * inv;
  G.g[2][1] = e1x * inv;
  G.g[0][0] = -(G.g[1][0] + G.g[2][0]);
  G.g[0][1] = -(G.g[1][1] + G.g[2][1]);
}

__device__ __forceinline__ void cell_geometry(const double (&x)[4][3], CellGeom<3>& G) {
  double e[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) e[j][i] = x[j + 1][i] - x[0][i];
  // cross products
  const double c23x = e[1][1] * e[2][2] - e[1][2] * e[2][1];
  const double c23y = e[1][2] * e[2][0] - e[1][0] * e[2][2];
  const double c23z = e[1][0] * e[2][1] - e[1][1] * e[2][0];
  const double c31x = e[2][1] * e[0][2] - e[2][2] * e[0][1];
  const double c31y = e[2][2] * e[0][0] - e[2][0] * e[0][2];
  const double c31z = e[2][0] * e[0][1] - e[2][1] * e[0][0];
  const double c12x = e[0][1] * e[1][2] - e[0][2] * e[1][1];
  const double c12y = e[0][2] * e[1][0] - e[0][0] * e[1][2];
  const double c12z = e[0][0] * e[1][1] - e[0][1] * e[1][0];
  const double det = e[0][0] * c23x + e[0][1] * c23y + e[0][2] * c23z;
  const double inv = 1.0 / det;
  G.vol = fabs(det) / 6.0;
  G.g[1][0] = c23x * inv; G.g[1][1] = c23y * inv; G.g[1][2] = c23z * inv;
  G.g[2][0] = c31x * inv; G.g[2][1] = c31y * inv; G.g[2][2] = c31z * inv;
  G.g[3][0] = c12x * inv; G.g[3][1] = c12y * inv; G.g[3][2] = c12z * inv;
#pragma unroll
  for (int i = 0; i < 3; ++i) G.g[0][i] = -(G.g[1][i] + G.g[2][i] + G.g[3][i]);
}

// MODE 0: system matrix A and right-hand side b.   MODE 1: block-Jacobi preconditioner matrix P.
//
// One CTA owns a TILE of consecutive owned dofs ("nodes") of one subdomain.  A node is served by a group of G lanes
// (G = power of two >= the largest vertex degree): lane e of the group owns adjacency slot e, i.e. one column position
// of all ten block rows of that node, and (on membrane nodes) gamma slot e.
//   phase 1  gather: coordinates and concentrations of the lane's neighbour into shared memory (one gather per
//            (node, slot), all independent -> memory-level parallelism)
//   phase 2a one thread per (node, incident cell): P1 geometry from the staged coordinates, the cell's stiffness row
//            of that node, mass weight and cell-mean concentrations -> shared memory
//   phase 2b every lane walks its node's incident cells in ascending order (fixed order, no atomics -> bitwise
//            reproducible), picks the cells that contain its slot (byte-wise SIMD compare on the packed slots) and
//            accumulates mass, stiffness and cbar_k-weighted stiffness in registers; membrane nodes add the facet
//            tensors of facet_kernel the same way
//   phase 3  all ten block rows are formed from the five accumulators and dropped at their CSR-relative offsets of a
//            staging strip; the tile's rows of one field are ONE contiguous span of the CSR value array, which an
//            elected thread hands to the TMA engine (cp.async.bulk shared -> global), so the 8 B/nnz output stream
//            never passes through registers again; b_k = sum_e m_e c_k(e) is a fixed-order segmented sum.
// Every A value and b entry is written exactly once; index traffic is 1 byte per (node, cell, vertex).
constexpr int ROWS_THREADS = 256;
#ifndef ROWS_MIN_CTAS
#define ROWS_MIN_CTAS 4
#endif

struct RowsSmem {       // computed on the host
  int tile;             // nodes per CTA = ROWS_THREADS / G
  int lgG, lgI;         // log2 of the lanes per node and of the incidence slots per node
  int off_res, off_packed;
  int res_stride;       // doubles per node in the cell-result block, padded so that the lane groups of one warp hit
                        // different banks (stride mod 16 doubles = 4)
  int pk_stride;        // words per node in the packed-slot block (odd)
  int off_self;         // [tile][5] sums over all incident cells for the self slot
  int off_el;           // the tile's (node, slot) cell lists (bytes)
  int use_lists;
  int off_prod;         // products m_e c_k(e) for the right-hand side (inside the staging alias, after the rows)
  int off_rs;           // CSR row starts of the tile (not aliased)
  int total;
};

struct RowCoef {        // constants of the forms, folded on the host (KNPEMIx_problem.py:598-610,633-642)
  double dtD[3];        // dt D_k
  double cphi[3];       // dt D_k z_k / psi
  double cpp[3];        // dt D_k z_k^2 / psi
  double ck[3];         // dt z_k D_k
  double cmz[3];        // C_M / (F z_k)
  double cf;            // C_M / F
};

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

template <int D, int MODE, bool LISTS>
__global__ void __launch_bounds__(ROWS_THREADS, ROWS_MIN_CTAS) rows_kernel(DevTopo T, RowCoef C, const double* __restrict__ u,
                                                               const double* __restrict__ fe,
                                                               double* __restrict__ vals, double* __restrict__ bvec,
                                                               RowsSmem S, int ntile0) {
  constexpr int NV = D + 1;
  constexpr int NS = D * (D + 1) / 2;
  constexpr int NB = D + 3;                 // doubles per staged neighbour: coordinates + 3 concentrations
  constexpr int NR = NV + 4;                // doubles per (node, cell) result: stiffness row, mass weight, cbar[3]
  constexpr uint32_t VMASK = NV == 4 ? 0xFFFFFFFFu : 0x00FFFFFFu;
  constexpr uint32_t FMASK = D == 3 ? 0x00FFFFFFu : 0x0000FFFFu;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* nbr = reinterpret_cast<double*>(smraw);
  double* stg = nbr;                                             // alias: neighbours and cell results are dead by phase 3
  double* res = reinterpret_cast<double*>(smraw + S.off_res);
  uint32_t* packed = reinterpret_cast<uint32_t*>(smraw + S.off_packed);
  double* prod = reinterpret_cast<double*>(smraw + S.off_prod);
  int* rstart = reinterpret_cast<int*>(smraw + S.off_rs);        // [4][tile + 1] CSR row starts of the tile's rows
  double* selfacc = reinterpret_cast<double*>(smraw + S.off_self);   // [tile][5]: sums over all cells for the self slot
  uint8_t* el = smraw + S.off_el;                                // the tile's (node, slot) cell lists
  constexpr bool lists = LISTS;           // compile-time: the default (scan) kernel carries none of the list code

  const int tid = threadIdx.x;
  const int lgG = S.lgG, lgI = S.lgI, GI = 1 << lgI;
  const int tile = S.tile;
  const int s = blockIdx.x >= ntile0 ? 1 : 0;
  const int p0 = (blockIdx.x - (s ? ntile0 : 0)) * tile;
  const int n_own_s = T.L.n_own[s];
  const int nt = min(tile, n_own_s - p0);
  const int w0 = (s ? T.L.n_own[0] : 0) + p0;
  const int nodeoff = s ? T.L.n_loc[0] : 0;
  const int* __restrict__ iptr = MODE == 0 ? T.indptr : T.indptr_P;

  const int lw = tid >> lgG, e = tid & ((1 << lgG) - 1);
  const bool node_ok = lw < nt;
  int deg = 0, gdeg = 0, self = -1, g = -1, ninc = 0, i0 = 0, ecnt = 0;
  const int i_tile = T.inc_ptr[w0];
  if (node_ok) {
    const int w = w0 + lw;
    const int a0 = T.adj_ptr[w];
    deg = T.adj_ptr[w + 1] - a0;
    self = T.self_slot[w];
    g = T.mv_of_node[w];
    i0 = T.inc_ptr[w];
    ninc = T.inc_ptr[w + 1] - i0;
    if (lists && e < deg) ecnt = T.ecnt[a0 + e];
    if (MODE == 0) gdeg = T.gpre[w + 1] - T.gpre[w];
    if (e < 4) {
      rstart[e * (tile + 1) + lw] = iptr[T.L.row(s, e, p0 + lw)];
      if (lw == nt - 1) rstart[e * (tile + 1) + nt] = iptr[T.L.row(s, e, p0 + nt)];
    }
    // ---- phase 1: one gather per (node, slot) ----
    if (e < deg) {
      const int q = T.adj_idx[a0 + e];
      double* o = nbr + (size_t)tid * NB;
#pragma unroll
      for (int i = 0; i < D; ++i) o[i] = T.node_x[(size_t)(nodeoff + q) * D + i];
      const double* __restrict__ uc = q < n_own_s ? u + T.L.rowbase[s] + q : u + T.L.n_rows + T.L.gbase[s] + (q - n_own_s);
      const int fstride = q < n_own_s ? n_own_s : T.L.n_gh[s];
#pragma unroll
      for (int k = 0; k < 3; ++k) o[D + k] = uc[(size_t)k * fstride];
    }
  }
  const bool has_ent = node_ok && e < deg;
  __syncthreads();

  // ---- phase 2a: one thread per (node, incident cell); the loop bound is warp-uniform because the GI threads of a node
  //      also reduce the node's self-slot sums (all cells contribute to the self slot) with a shuffle butterfly ----
  for (int i = tid; i < (tile << lgI); i += ROWS_THREADS) {
    const int n = i >> lgI, j = i & (GI - 1);
    bool valid = n < nt;
    int ii0 = 0;
    if (valid) {
      ii0 = T.inc_ptr[w0 + n];
      valid = j < T.inc_ptr[w0 + n + 1] - ii0;
    }
    double sself[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (valid) {
      const uint32_t pk = T.inc_slots[ii0 + j];
      packed[(size_t)n * S.pk_stride + j] = pk;
      const int la = (__ffs(__vcmpeq4(pk, (uint32_t)T.self_slot[w0 + n] * 0x01010101u) & VMASK) - 1) >> 3;
      const double* nb = nbr + ((size_t)n << lgG) * NB;
      double x[NV][D], csum[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int b = 0; b < NV; ++b) {
        const double* v = nb + ((pk >> (8 * b)) & 255u) * NB;
#pragma unroll
        for (int d = 0; d < D; ++d) x[b][d] = v[d];
#pragma unroll
        for (int k = 0; k < 3; ++k) csum[k] += v[D + k];
      }
      CellGeom<D> G;
      cell_geometry(x, G);
      double gl[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double t = G.g[0][d];
#pragma unroll
        for (int a = 1; a < NV; ++a) t = (la == a) ? G.g[a][d] : t;
        gl[d] = t;
      }
      double* r = res + (size_t)n * S.res_stride + (size_t)j * NR;
      double kself = 0.0;
#pragma unroll
      for (int b = 0; b < NV; ++b) {
        double dot = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) dot += gl[d] * G.g[b][d];
        const double kab = G.vol * dot;
        r[b] = kab;
        kself = (b == la) ? kab : kself;
      }
      const double mv = G.vol * (1.0 / ((D + 1) * (D + 2)));
      r[NV] = mv;
      if (lists) {
        sself[0] = mv;
        sself[1] = kself;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double cb = csum[k] * (1.0 / NV);
        r[NV + 1 + k] = cb;
        if (lists) sself[2 + k] = cb * kself;
      }
    }
    if (lists) {
      for (int off = GI >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int q = 0; q < 5; ++q) sself[q] += __shfl_xor_sync(0xffffffffu, sself[q], off);
      }
      if (j == 0 && n < nt) {
#pragma unroll
        for (int q = 0; q < 5; ++q) selfacc[n * 5 + q] = sself[q];
      }
    }
  }
  if (lists) {      // the tile's cell lists: one contiguous byte range of the global table
    const int nbytes = (NV - 1) * (T.inc_ptr[w0 + nt] - i_tile);
    const uint8_t* __restrict__ src = T.elist + (size_t)(NV - 1) * i_tile;
    for (int i = tid; i < nbytes; i += ROWS_THREADS) el[i] = src[i];
  }
  __syncthreads();

  // ---- phase 2b: accumulate per (node, slot) in registers, cells in ascending order ----
  double a_m = 0.0, a_kk = 0.0, X[3] = {0.0, 0.0, 0.0}, kphi_m[3] = {0.0, 0.0, 0.0}, pp_m = 0.0;
  double bmem[4] = {0.0, 0.0, 0.0, 0.0}, ce[3] = {0.0, 0.0, 0.0};
  double ga[3] = {0.0, 0.0, 0.0}, g1 = 0.0;
  const bool is_self = has_ent && e == self;
  const bool has_gam = MODE == 0 && node_ok && e < gdeg;
  const uint32_t rep = (uint32_t)e * 0x01010101u;
  int eoff = 0;
  if (lists) {      // start of this lane's list inside its node's block: exclusive prefix sum of the list lengths
    int incl = ecnt;
    for (int off = 1; off < (1 << lgG); off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off, 1 << lgG);
      if (e >= off) incl += v;
    }
    eoff = incl - ecnt;
  }
  if (has_ent) {
#pragma unroll
    for (int k = 0; k < 3; ++k) ce[k] = nbr[(size_t)tid * NB + D + k];
    const double* rbase = res + (size_t)lw * S.res_stride;
    if (!lists) {
      const uint32_t* pk = packed + (size_t)lw * S.pk_stride;
      for (int j = 0; j < ninc; ++j) {
        const uint32_t m = __vcmpeq4(pk[j], rep) & VMASK;
        if (m) {
          const int b = (__ffs(m) - 1) >> 3;
          const double* r = rbase + j * NR;
          const double kab = r[b], mv = r[NV];
          a_m += is_self ? 2.0 * mv : mv;
          a_kk += kab;
#pragma unroll
          for (int k = 0; k < 3; ++k) X[k] += r[NV + 1 + k] * kab;
        }
      }
    } else if (is_self) {
      const double* sa = selfacc + lw * 5;
      a_m = 2.0 * sa[0];
      a_kk = sa[1];
#pragma unroll
      for (int k = 0; k < 3; ++k) X[k] = sa[2 + k];
    } else {
      const uint8_t* lst = el + (NV - 1) * (i0 - i_tile) + eoff;
      for (int t = 0; t < ecnt; ++t) {
        const int code = lst[t];
        const double* r = rbase + (code >> 2) * NR;
        const double kab = r[code & 3];
        a_m += r[NV];
        a_kk += kab;
#pragma unroll
        for (int k = 0; k < 3; ++k) X[k] += r[NV + 1 + k] * kab;
      }
    }
  }
  // membrane (dS) terms: KNPEMIx_problem.py:599,604,609-610,637-638,641-642 (P: :737-738); lane e serves adjacency
  // slot e and gamma slot e (couplings to the potential on the other side of the membrane)
  if (g >= 0 && (has_ent || has_gam)) {
    const size_t nf = (size_t)T.n_mf;
    const double sgn = s == 0 ? 1.0 : -1.0;
    const int m1 = T.minc_ptr[g + 1];
    for (int mi = T.minc_ptr[g]; mi < m1; ++mi) {
      const uint4 rec = reinterpret_cast<const uint4*>(T.minc)[mi];
      const int f = (int)rec.x;
      const int a = rec.y & 255u;
      const uint32_t ss = s == 0 ? (rec.y >> 8) : rec.z;
      const uint32_t ms = has_ent ? (__vcmpeq4(ss, rep) & FMASK) : 0u;
      const uint32_t mg = has_gam ? (__vcmpeq4(rec.w, rep) & FMASK) : 0u;
      if (ms) {
        const int b = (__ffs(ms) - 1) >> 3;
        const double G1 = T.mf_area[f] * ((a == b) ? 2.0 : 1.0) * (1.0 / (D * (D + 1)));
        if (MODE == 0) {
          const int ab = a <= b ? symidx(a, b, D) : symidx(b, a, D);
#pragma unroll
          for (int k = 0; k < 3; ++k) kphi_m[k] += C.cmz[k] * fe[(size_t)((s * 3 + k) * NS + ab) * nf + f];
          pp_m += C.cf * G1;
        } else {
          pp_m -= C.cf * G1;
        }
      }
      if (MODE == 0 && mg) {
        const int b = (__ffs(mg) - 1) >> 3;
        const int ab = a <= b ? symidx(a, b, D) : symidx(b, a, D);
#pragma unroll
        for (int k = 0; k < 3; ++k) ga[k] += C.cmz[k] * fe[(size_t)((s * 3 + k) * NS + ab) * nf + f];
        g1 += C.cf * (T.mf_area[f] * ((a == b) ? 2.0 : 1.0) * (1.0 / (D * (D + 1))));
      }
      if (MODE == 0 && is_self) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bmem[k] -= sgn * fe[(size_t)(6 * NS + (s * 3 + k) * D + a) * nf + f];
        bmem[3] -= sgn * fe[(size_t)(6 * NS + 6 * D + a) * nf + f];
      }
    }
  }
  __syncthreads();          // neighbour data and cell results are dead: the staging strip may overwrite them

  // ---- phase 3: form the ten block rows at their CSR-relative offsets of the staging strip ----
  int so[4], base[4], total[4];
  {
    int acc = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      base[f] = rstart[f * (tile + 1)];
      total[f] = rstart[f * (tile + 1) + nt] - base[f];
      so[f] = acc;                                   // even: 16-byte aligned start of the field's strip
      acc += (total[f] + 3) & ~1;                    // room for the phase shift (base & 1), rounded to even
    }
  }
  if (node_ok) {
    const int goff = s == 1 ? gdeg : 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int rsf = rstart[f * (tile + 1) + lw] - base[f];
      double* o = stg + so[f] + (base[f] & 1) + rsf;
      if (has_ent) {
        double* oe = o + goff + e;
        if (MODE == 0) {
          if (f < 3) {
            oe[0] = a_m + C.dtD[f] * a_kk;
            oe[deg] = C.cphi[f] * X[f] + kphi_m[f];
          } else {
            double pp = pp_m;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              oe[k * deg] = C.ck[k] * a_kk;
              pp += C.cpp[k] * X[k];
            }
            oe[3 * deg] = pp;
          }
        } else {
          if (f < 3) {
            oe[0] = a_m + C.dtD[f] * a_kk;
          } else {
            double pp = pp_m;
#pragma unroll
            for (int k = 0; k < 3; ++k) pp += C.cpp[k] * X[k];
            oe[0] = pp;
          }
        }
      }
      if (has_gam) o[(s == 1 ? 0 : (f < 3 ? 2 : 4) * deg) + e] = f < 3 ? -ga[f] : -g1;
    }
    if (MODE == 0 && has_ent && lgG > 5) {
#pragma unroll
      for (int k = 0; k < 3; ++k) prod[(size_t)tid * 3 + k] = a_m * ce[k];
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the TMA (async proxy)
  __syncthreads();

  // ---- copy-out: one TMA bulk store per field for the 16-byte aligned body, scalar head/tail ----
  if (tid == 0) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int sh = base[f] & 1;
      const int nbody = (total[f] - sh) & ~1;
      if (nbody > 0) bulk_store(vals + (size_t)base[f] + sh, stg + so[f] + 2 * sh, (uint32_t)nbody * 8u);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  } else if (tid >= 32 && tid < 36) {
#pragma unroll
    for (int f = 0; f < 4; ++f)
      if (f == tid - 32) {
        const int sh = base[f] & 1;
        if (sh && total[f] > 0) vals[(size_t)base[f]] = stg[so[f] + 1];
        if ((total[f] - sh) & 1) vals[(size_t)base[f] + total[f] - 1] = stg[so[f] + sh + total[f] - 1];
      }
  }
  // right-hand side: b_k = sum_e m_e c_k(e) (KNPEMIx_problem.py:613-614,641-642): segmented warp-shuffle reduction over
  // the node's lane group (fixed butterfly order -> reproducible); groups wider than a warp fall back to a serial sum
  if (MODE == 0) {
    if (lgG <= 5) {
      double bk[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) bk[k] = has_ent ? a_m * ce[k] : 0.0;
      for (int off = (1 << lgG) >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bk[k] += __shfl_xor_sync(0xffffffffu, bk[k], off);
      }
      if (is_self) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bvec[T.L.row(s, k, p0 + lw)] = bk[k] + bmem[k];
        bvec[T.L.row(s, 3, p0 + lw)] = bmem[3];
      }
    } else if (is_self) {
      const double* pr = prod + ((size_t)lw << lgG) * 3;
      double bk[3] = {0.0, 0.0, 0.0};
      for (int j = 0; j < deg; ++j) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bk[k] += pr[j * 3 + k];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) bvec[T.L.row(s, k, p0 + lw)] = bk[k] + bmem[k];
      bvec[T.L.row(s, 3, p0 + lw)] = bmem[3];
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ point probes
// out[i] = sum_{t in [ptr[i], ptr[i+1])} w[t] * u[col[t]]: point evaluation of P1 fields (scifem.evaluate_function in
// SolverKNPEMI.init_data / save_data, KNPEMIx_solver.py:612-643) with the containing cell and the barycentric weights found
// once on the host; one thread per output value, fixed summation order.
__global__ void probe_kernel(int n_out, const int32_t* __restrict__ ptr, const int32_t* __restrict__ col,
                             const double* __restrict__ w, const double* __restrict__ u, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  double s = 0.0;
  for (int t = ptr[i]; t < ptr[i + 1]; ++t) s += w[t] * u[col[t]];
  out[i] = s;
}
int launch_probe(int n_out, const int32_t* ptr, const int32_t* col, const double* w, const double* u, double* out, cudaStream_t st) {
  if (n_out == 0) return KNP_OK;
  probe_kernel<<<(n_out + 127) / 128, 128, 0, st>>>(n_out, ptr, col, w, u, out);
  KNP_LAUNCHED();
  return KNP_OK;
}

// ------------------------------------------------------------------------------------------------ CSR indices
template <int MODE>
__global__ void csr_indices_kernel(DevTopo T, int32_t* __restrict__ indices) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= T.n_work) return;
  const int s = w >= T.L.n_own[0] ? 1 : 0;
  const int p = w - (s ? T.L.n_own[0] : 0);
  const int o = 1 - s;
  const int a0 = T.adj_ptr[w], deg = T.adj_ptr[w + 1] - a0;
  const int g = T.mv_of_node[w];
  const int g0 = g >= 0 ? T.gam_ptr[g] : 0;
  const int gdeg = (MODE == 0 && g >= 0) ? T.gam_ptr[g + 1] - g0 : 0;
  const int32_t* mvo = o == 0 ? T.mv_node0 : T.mv_node1;
  const int* iptr = MODE == 0 ? T.indptr : T.indptr_P;
  for (int f = 0; f < 4; ++f) {
    int pos = iptr[T.L.row(s, f, p)];
    if (MODE == 1) {
      for (int e = 0; e < deg; ++e) indices[pos++] = T.L.col(s, f, T.adj_idx[a0 + e]);
      continue;
    }
    if (s == 1)
      for (int e = 0; e < gdeg; ++e) indices[pos++] = T.L.col(o, 3, mvo[T.gam_mv[g0 + e]]);
    if (f < 3) {
      for (int e = 0; e < deg; ++e) indices[pos++] = T.L.col(s, f, T.adj_idx[a0 + e]);
      for (int e = 0; e < deg; ++e) indices[pos++] = T.L.col(s, 3, T.adj_idx[a0 + e]);
    } else {
      for (int k = 0; k < 4; ++k)
        for (int e = 0; e < deg; ++e) indices[pos++] = T.L.col(s, k, T.adj_idx[a0 + e]);
    }
    if (s == 0)
      for (int e = 0; e < gdeg; ++e) indices[pos++] = T.L.col(o, 3, mvo[T.gam_mv[g0 + e]]);
  }
}

// ------------------------------------------------------------------------------------------------ functionals
// Cell functionals over tagged cells: power 2 = int u^2 (tests/KNPEMI/electric_potential_norms_direct_solver.py:45-51),
// power 1 = int u (ion amounts of ProblemKNPEMI.print_conservation, KNPEMIx_problem.py:807-843), power 0 = the measure of
// the tagged cells; per-block partial sums, reduced in a fixed order by reduce_partials_kernel (linalg.cu).
template <int D>
__global__ void __launch_bounds__(256) l2_cells_kernel(Layout L, int s, int field, int power, int n_cells,
                                                       const int32_t* __restrict__ cell_nodes,
                                                       const int32_t* __restrict__ cell_tag,
                                                       const int32_t* __restrict__ cell_owned,
                                                       const double* __restrict__ node_x, int nodeoff,
                                                       const int32_t* __restrict__ tags, int n_tags,
                                                       const double* __restrict__ u, double* __restrict__ partial) {
  constexpr int NV = D + 1;
  __shared__ double red[256];
  double acc = 0.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
    if (!cell_owned[c]) continue;
    const int t = cell_tag[c];
    bool hit = false;
    for (int i = 0; i < n_tags; ++i) hit |= (tags[i] == t);
    if (!hit) continue;
    double x[NV][D], uc[NV];
#pragma unroll
    for (int a = 0; a < NV; ++a) {
      const int q = cell_nodes[(size_t)c * NV + a];
#pragma unroll
      for (int i = 0; i < D; ++i) x[a][i] = node_x[(size_t)(nodeoff + q) * D + i];
      uc[a] = u[L.col(s, field, q)];
    }
    CellGeom<D> G;
    cell_geometry(x, G);
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int a = 0; a < NV; ++a) {
      s1 += uc[a];
      s2 += uc[a] * uc[a];
    }
    if (power == 2) acc += G.vol / ((D + 1) * (D + 2)) * (s2 + s1 * s1);
    else if (power == 1) acc += G.vol * (s1 * (1.0 / NV));
    else acc += G.vol;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// Total stimulus current int stim_expr dS(stimulus_tags) (SolverKNPEMI.init_png_data / save_png, KNPEMIx_solver.py:578-610,
// with stim_expr of HodgkinHuxley._add_stimulus, KNPEMIx_ionic_model.py:517-603): thread per owned stimulated facet,
// per-block partial sums in a fixed order.
template <int D>
__global__ void __launch_bounds__(256) stim_current_kernel(DevTopo T, KParams P, const int32_t* __restrict__ tag_stim,
                                                           const int32_t* __restrict__ mf_owned,
                                                           const double* __restrict__ u, double stim_fac,
                                                           double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < T.n_mf; f += gridDim.x * blockDim.x) {
    if (!mf_owned[f] || tag_stim[T.mf_tagidx[f]] == 0) continue;
    double ci[D], ce[D], pm[D];
    int nodei[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const int g = T.mf_mv[(size_t)f * D + a];
      const int qi = T.mv_node0[g], qe = T.mv_node1[g];
      ci[a] = u[T.L.col(0, 0, qi)];
      ce[a] = u[T.L.col(1, 0, qe)];
      pm[a] = u[T.L.col(0, 3, qi)] - u[T.L.col(1, 3, qe)];
      nodei[a] = qi;
    }
    const double area = T.mf_area[f];
    for (int q = 0; q < T.nq; ++q) {
      double ciq = 0.0, ceq = 0.0, pmq = 0.0;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const double lam = T.qb[q * D + a];
        ciq += lam * ci[a];
        ceq += lam * ce[a];
        pmq += lam * pm[a];
      }
      const double E_Na = (P.psi / P.z[0]) * log(ceq / ciq);
      double mask = 1.0;
      for (int i = 0; i < 3 && P.stim_dir[i] >= 0; ++i) {
        double xq = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) xq += T.qb[q * D + a] * T.node_x[(size_t)nodei[a] * D + P.stim_dir[i]];
        mask *= (xq > P.stim_lo[i] && xq < P.stim_hi[i]) ? 1.0 : 0.0;
      }
      acc += area * T.qw[q] * mask * stim_fac * (pmq - E_Na);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

int launch_stim_current(const DevTopo& T, const KParams& P, const int32_t* tag_stim, const int32_t* mf_owned,
                        const double* u, double stim_fac, double* partial, int n_partial, cudaStream_t st) {
  if (T.gdim == 2) stim_current_kernel<2><<<n_partial, 256, 0, st>>>(T, P, tag_stim, mf_owned, u, stim_fac, partial);
  else stim_current_kernel<3><<<n_partial, 256, 0, st>>>(T, P, tag_stim, mf_owned, u, stim_fac, partial);
  KNP_LAUNCHED();
  return KNP_OK;
}

// ------------------------------------------------------------------------------------------------ launchers
int launch_gate(const DevTopo& T, const KParams& P, const double* u, double* gates, cudaStream_t st) {
  if (T.n_mv == 0) return KNP_OK;
  gate_kernel<<<(T.n_mv + 127) / 128, 128, 0, st>>>(T, P, u, gates);
  KNP_LAUNCHED();
  return KNP_OK;
}

int facet_ncomp(int gdim) { return 6 * (gdim * (gdim + 1) / 2) + 7 * gdim; }

int launch_facets(const DevTopo& T, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim,
                  const double* u, const double* gates, double stim_fac, double* fe, cudaStream_t st) {
  if (T.n_mf == 0) return KNP_OK;
  const int grid = (T.n_mf + 127) / 128;
  if (T.gdim == 2)
    facet_kernel<2><<<grid, 128, 0, st>>>(T, P, tag_models, tag_stim, u, gates, stim_fac, fe);
  else
    facet_kernel<3><<<grid, 128, 0, st>>>(T, P, tag_models, tag_stim, u, gates, stim_fac, fe);
  KNP_LAUNCHED();
  return KNP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Row kernel, second form: ONE THREAD PER DOF over static, TMA-friendly geometry tables.
//
// The mesh does not move, so everything geometric is computed once at setup (geo_build_kernel) and stored in
// structure-of-arrays ELL tables whose leading dimension is the dof index (coalesced for a warp of 32 dofs):
//   adjE[e][w]      neighbour of slot e                       incE[j][w]   packed slots of incident cell j
//   geoK[j][b][w]   stiffness row of dof w in cell j          mslot/kslot[e][w]  assembled mass / stiffness per slot
// Per timestep only the concentration-weighted stiffness X_k[e] = sum_j cbar_k(j) K_j[b(e)] depends on the solution.
// A thread walks its dof's cells in ascending order (fixed order, no atomics), gathers the three concentrations of the
// cell's vertices, and adds cbar_k K into its PRIVATE column of a shared-memory table (dynamic slot index, conflict-free
// because the dof index is the fastest dimension).  The ten block rows are then formed field by field into a staging
// strip at their CSR-relative offsets and leave through TMA bulk stores (two strips, ping-pong, so a store overlaps the
// next field).  ~35 warp instructions per dof instead of ~245 for the lane-group kernel; the price is the static
// tables (2D: +0.34 kB per dof of reads on top of 0.67 kB of algorithmic traffic).
// MEASURED (B200, round 1): 2.54 ms on C3 and 5.96 ms on C4 against 1.75 / 3.17 ms for the lane-group kernel -- with
// 0.5-1.2 kB of shared memory per thread only 12 (2D) / 4 (3D) warps fit an SM and the dependent incE -> adjE -> u
// chain is latency-bound.  The kernel is therefore OPT-IN (KNP_ROWS=ell at context creation); it passes the same parity
// tests and is kept as the starting point for a software-pipelined version.

struct EllSmem {
  int max_deg, max_inc, Wp;
  int off_strip0, off_strip1;   // bytes
  int total;
};

template <int D>
__global__ void geo_build_kernel(DevTopo T, int Wp, int max_deg, int32_t* __restrict__ adjE, uint32_t* __restrict__ incE,
                                 double* __restrict__ geoK, double* __restrict__ mslot, double* __restrict__ kslot) {
  constexpr int NV = D + 1;
  constexpr uint32_t VMASK = NV == 4 ? 0xFFFFFFFFu : 0x00FFFFFFu;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= T.n_work) return;
  const int s = w >= T.L.n_own[0] ? 1 : 0;
  const int nodeoff = s ? T.L.n_loc[0] : 0;
  const int a0 = T.adj_ptr[w], deg = T.adj_ptr[w + 1] - a0;
  const int i0 = T.inc_ptr[w], ninc = T.inc_ptr[w + 1] - i0;
  const int self = T.self_slot[w];
  for (int e = 0; e < max_deg; ++e) {
    adjE[(size_t)e * Wp + w] = e < deg ? T.adj_idx[a0 + e] : -1;
    mslot[(size_t)e * Wp + w] = 0.0;
    kslot[(size_t)e * Wp + w] = 0.0;
  }
  for (int j = 0; j < T.max_inc; ++j) {
    if (j >= ninc) {
      incE[(size_t)j * Wp + w] = 0xFFFFFFFFu;
#pragma unroll
      for (int b = 0; b < NV; ++b) geoK[((size_t)j * NV + b) * Wp + w] = 0.0;
      continue;
    }
    const uint32_t pk = T.inc_slots[i0 + j];
    incE[(size_t)j * Wp + w] = pk;
    const int la = (__ffs(__vcmpeq4(pk, (uint32_t)self * 0x01010101u) & VMASK) - 1) >> 3;
    double x[NV][D];
#pragma unroll
    for (int b = 0; b < NV; ++b) {
      const int q = T.adj_idx[a0 + ((pk >> (8 * b)) & 255u)];
#pragma unroll
      for (int d = 0; d < D; ++d) x[b][d] = T.node_x[(size_t)(nodeoff + q) * D + d];
    }
    CellGeom<D> G;
    cell_geometry(x, G);
    double gl[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double t = G.g[0][d];
#pragma unroll
      for (int a = 1; a < NV; ++a) t = (la == a) ? G.g[a][d] : t;
      gl[d] = t;
    }
    const double mv = G.vol * (1.0 / ((D + 1) * (D + 2)));
#pragma unroll
    for (int b = 0; b < NV; ++b) {
      double dot = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) dot += gl[d] * G.g[b][d];
      const double kab = G.vol * dot;
      geoK[((size_t)j * NV + b) * Wp + w] = kab;
      const int e = (pk >> (8 * b)) & 255u;
      mslot[(size_t)e * Wp + w] += (b == la) ? 2.0 * mv : mv;
      kslot[(size_t)e * Wp + w] += kab;
    }
  }
}

template <int D, int MODE, int TB>
__global__ void __launch_bounds__(TB) rows_ell_kernel(DevTopo T, RowCoef C, const double* __restrict__ u,
                                                               const double* __restrict__ fe,
                                                               double* __restrict__ vals, double* __restrict__ bvec,
                                                               EllSmem S, int nb0) {
  constexpr int NV = D + 1;
  constexpr int NS = D * (D + 1) / 2;
  constexpr uint32_t FMASK = D == 3 ? 0x00FFFFFFu : 0x0000FFFFu;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* X = reinterpret_cast<double*>(smraw);                 // [3][max_deg][TB], thread-private columns
  double* strips[2] = {reinterpret_cast<double*>(smraw + S.off_strip0), reinterpret_cast<double*>(smraw + S.off_strip1)};
  __shared__ int sh_base[4], sh_end[4];
  const int tid = threadIdx.x;
  const int s = blockIdx.x >= nb0 ? 1 : 0;
  const int p0 = (blockIdx.x - (s ? nb0 : 0)) * TB;
  const int n_own_s = T.L.n_own[s], n_gh_s = T.L.n_gh[s];
  const int nt = min(TB, n_own_s - p0);
  const int p = p0 + tid;
  const bool active = tid < nt;
  const int w = (s ? T.L.n_own[0] : 0) + p;
  const int Wp = S.Wp, max_deg = S.max_deg;
  const int* __restrict__ iptr = MODE == 0 ? T.indptr : T.indptr_P;
  const double* __restrict__ u_own = u + T.L.rowbase[s];
  const double* __restrict__ u_gh = u + T.L.n_rows + T.L.gbase[s] - n_own_s;

  int deg = 0, gdeg = 0, g = -1, rs[4] = {0, 0, 0, 0};
  if (active) {
    deg = T.adj_ptr[w + 1] - T.adj_ptr[w];
    g = T.mv_of_node[w];
    if (MODE == 0) gdeg = T.gpre[w + 1] - T.gpre[w];
#pragma unroll
    for (int f = 0; f < 4; ++f) rs[f] = iptr[T.L.row(s, f, p)];
    if (tid == 0) {
#pragma unroll
      for (int f = 0; f < 4; ++f) sh_base[f] = rs[f];
    }
    if (tid == nt - 1) {
#pragma unroll
      for (int f = 0; f < 4; ++f) sh_end[f] = iptr[T.L.row(s, f, p) + 1];
    }
    for (int e = 0; e < deg; ++e) {
#pragma unroll
      for (int k = 0; k < 3; ++k) X[(size_t)(k * max_deg + e) * TB + tid] = 0.0;
    }
    // ---- X_k[e] += cbar_k(j) K_j[b]: cells in ascending order; the packed slots and the stiffness row of the next
    //      cell are fetched (read-only path) while the current one is processed ----
    const uint32_t* __restrict__ incE = T.incE;
    const int32_t* __restrict__ adjE = T.adjE;
    const double* __restrict__ geoK = T.geoK;
    uint32_t pk = __ldg(incE + w);
    double kb[NV];
#pragma unroll
    for (int b = 0; b < NV; ++b) kb[b] = __ldg(geoK + (size_t)b * Wp + w);
    for (int j = 0; j < S.max_inc && pk != 0xFFFFFFFFu; ++j) {
      uint32_t pk_n = 0xFFFFFFFFu;
      double kb_n[NV];
      if (j + 1 < S.max_inc) {
        pk_n = __ldg(incE + (size_t)(j + 1) * Wp + w);
#pragma unroll
        for (int b = 0; b < NV; ++b) kb_n[b] = __ldg(geoK + ((size_t)(j + 1) * NV + b) * Wp + w);
      }
      int sl[NV], q[NV];
#pragma unroll
      for (int b = 0; b < NV; ++b) {
        sl[b] = (pk >> (8 * b)) & 255u;
        q[b] = __ldg(adjE + (size_t)sl[b] * Wp + w);
      }
      double cv[3][NV];
#pragma unroll
      for (int b = 0; b < NV; ++b) {
        const double* __restrict__ uc = q[b] < n_own_s ? u_own + q[b] : u_gh + q[b];
        const int fs = q[b] < n_own_s ? n_own_s : n_gh_s;
#pragma unroll
        for (int k = 0; k < 3; ++k) cv[k][b] = __ldg(uc + (size_t)k * fs);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double cs = 0.0;
#pragma unroll
        for (int b = 0; b < NV; ++b) cs += cv[k][b];
        const double cb = cs * (1.0 / NV);
#pragma unroll
        for (int b = 0; b < NV; ++b) X[(size_t)(k * max_deg + sl[b]) * TB + tid] += cb * kb[b];
      }
      pk = pk_n;
#pragma unroll
      for (int b = 0; b < NV; ++b) kb[b] = kb_n[b];
    }
  }
  __syncthreads();

  const size_t nf = (size_t)T.n_mf;
  const double sgn = s == 0 ? 1.0 : -1.0;
  const int m0 = g >= 0 ? T.minc_ptr[g] : 0, m1 = g >= 0 ? T.minc_ptr[g + 1] : 0;
  const int goff = s == 1 ? gdeg : 0;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    double* strip = strips[f & 1];
    if (f >= 2) {                                   // the strip was last read by the bulk store of field f - 2
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
    }
    const int base = sh_base[f], total = sh_end[f] - base, sh = base & 1;
    if (active) {
      double* o = strip + sh + (rs[f] - base);
      double bsum = 0.0;
      double m_n = f < 3 ? __ldg(T.mslot + w) : 0.0, k_n = (f < 3 || MODE == 0) ? __ldg(T.kslot + w) : 0.0;
      int q_n = (f < 3 && MODE == 0) ? __ldg(T.adjE + w) : 0;
      for (int e = 0; e < deg; ++e) {
        double kphi_m = 0.0, pp_m = 0.0;
        if (g >= 0) {      // membrane (dS) terms: KNPEMIx_problem.py:599,604,609-610,637-638 (P: :737-738)
          const uint32_t rep = (uint32_t)e * 0x01010101u;
          for (int mi = m0; mi < m1; ++mi) {
            const uint4 rec = reinterpret_cast<const uint4*>(T.minc)[mi];
            const uint32_t ss = s == 0 ? (rec.y >> 8) : rec.z;
            const uint32_t ms = __vcmpeq4(ss, rep) & FMASK;
            if (ms) {
              const int fct = (int)rec.x, a = rec.y & 255u, b = (__ffs(ms) - 1) >> 3;
              if (f < 3) {
                if (MODE == 0) {
                  const int ab = a <= b ? symidx(a, b, D) : symidx(b, a, D);
                  kphi_m += C.cmz[f] * fe[(size_t)((s * 3 + f) * NS + ab) * nf + fct];
                }
              } else {
                const double G1 = T.mf_area[fct] * ((a == b) ? 2.0 : 1.0) * (1.0 / (D * (D + 1)));
                pp_m += (MODE == 0 ? C.cf : -C.cf) * G1;
              }
            }
          }
        }
        const double m = m_n, kk = k_n;
        const int q = q_n;
        if (e + 1 < deg) {
          const size_t at = (size_t)(e + 1) * Wp + w;
          if (f < 3) m_n = __ldg(T.mslot + at);
          if (f < 3 || MODE == 0) k_n = __ldg(T.kslot + at);
          if (f < 3 && MODE == 0) q_n = __ldg(T.adjE + at);
        }
        if (f < 3) {
          o[goff + e] = m + C.dtD[f] * kk;
          if (MODE == 0) {
            o[goff + deg + e] = C.cphi[f] * X[(size_t)(f * max_deg + e) * TB + tid] + kphi_m;
            bsum += m * (q < n_own_s ? __ldg(u_own + (size_t)f * n_own_s + q) : __ldg(u_gh + (size_t)f * n_gh_s + q));
          }
        } else {
          double pp = pp_m;
#pragma unroll
          for (int k = 0; k < 3; ++k) pp += C.cpp[k] * X[(size_t)(k * max_deg + e) * TB + tid];
          if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) o[goff + k * deg + e] = C.ck[k] * kk;
            o[goff + 3 * deg + e] = pp;
          } else {
            o[e] = pp;
          }
        }
      }
      if (MODE == 0) {
        double bm = 0.0;
        if (g >= 0) {
          // gamma entries (couplings to the potential on the other side of the membrane) and the membrane rhs
          double* og = o + (s == 1 ? 0 : (f < 3 ? 2 : 4) * deg);
          for (int eg = 0; eg < gdeg; ++eg) {
            const uint32_t rep = (uint32_t)eg * 0x01010101u;
            double acc = 0.0;
            for (int mi = m0; mi < m1; ++mi) {
              const uint4 rec = reinterpret_cast<const uint4*>(T.minc)[mi];
              const uint32_t mg = __vcmpeq4(rec.w, rep) & FMASK;
              if (mg) {
                const int fct = (int)rec.x, a = rec.y & 255u, b = (__ffs(mg) - 1) >> 3;
                if (f < 3) {
                  const int ab = a <= b ? symidx(a, b, D) : symidx(b, a, D);
                  acc += C.cmz[f] * fe[(size_t)((s * 3 + f) * NS + ab) * nf + fct];
                } else {
                  acc += C.cf * (T.mf_area[fct] * ((a == b) ? 2.0 : 1.0) * (1.0 / (D * (D + 1))));
                }
              }
            }
            og[eg] = -acc;
          }
          for (int mi = m0; mi < m1; ++mi) {
            const uint4 rec = reinterpret_cast<const uint4*>(T.minc)[mi];
            const int fct = (int)rec.x, a = rec.y & 255u;
            bm -= sgn * (f < 3 ? fe[(size_t)(6 * NS + (s * 3 + f) * D + a) * nf + fct] : fe[(size_t)(6 * NS + 6 * D + a) * nf + fct]);
          }
        }
        bvec[T.L.row(s, f, p)] = bsum + bm;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const int nbody = (total - sh) & ~1;
      if (nbody > 0) bulk_store(vals + (size_t)base + sh, strip + 2 * sh, (uint32_t)nbody * 8u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    } else if (tid == 32) {
      if (sh && total > 0) vals[(size_t)base] = strip[1];
      if ((total - sh) & 1) vals[(size_t)base + total - 1] = strip[sh + total - 1];
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Lanes per node (power of two >= the largest degree) and the shared-memory layout of the row kernel.
static int ceil_log2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

static RowsSmem rows_layout(int gdim, int mode, int max_deg, int max_gdeg, int max_inc) {
  RowsSmem S{};
  S.lgG = ceil_log2(std::max(std::max(max_deg, max_gdeg), 1));
  S.lgI = ceil_log2(std::max(max_inc, 1));
  S.tile = std::max(1, ROWS_THREADS >> S.lgG);
  const int NB = gdim + 3, NR = gdim + 1 + 4;
  const size_t nbr = (size_t)ROWS_THREADS * NB * 8;
  S.res_stride = (NR << S.lgI);
  while (S.res_stride % 16 != 4) ++S.res_stride;
  S.pk_stride = (1 << S.lgI) | 1;
  const size_t res = (size_t)S.tile * S.res_stride * 8;
  const size_t packed = (size_t)S.tile * S.pk_stride * 4;
  S.off_res = (int)nbr;
  S.off_packed = (int)(nbr + res);
  size_t work = (nbr + res + packed + 15) & ~(size_t)15;
  S.off_self = (int)work;
  work += (size_t)S.tile * 5 * 8;
  S.off_el = (int)work;
  work += ((size_t)S.tile * gdim * max_inc + 15) & ~(size_t)15;
  // staging strip: all four fields of the tile (+ phase shift and rounding per field), then the rhs products
  const size_t rows = (size_t)S.tile * (mode == 0 ? 10 * max_deg + 4 * max_gdeg : 4 * max_deg) + 16;
  S.off_prod = (int)(rows * 8);
  const size_t stage = rows * 8 + (mode == 0 ? (size_t)ROWS_THREADS * 3 * 8 : 0);
  S.off_rs = (int)((std::max(work, stage) + 15) & ~(size_t)15);
  S.total = S.off_rs + 4 * (S.tile + 1) * 4;
  return S;
}

template <int D, int MODE>
static int launch_rows_t(const DevTopo& T, const KParams& P, const double* u, const double* fe, double* vals,
                         double* b, int max_deg, int max_gdeg, cudaStream_t st) {
  RowsSmem S = rows_layout(D, MODE, max_deg, max_gdeg, T.max_inc);
  // list-driven phase 2b needs the lane groups (adjacency and incidence) inside one warp
  S.use_lists = (T.elist && T.ecnt && S.lgG <= 5 && S.lgI <= 5 && ((S.tile << S.lgI) % 32) == 0) ? 1 : 0;
  if (S.total > 227 * 1024 || (1 << S.lgG) > ROWS_THREADS) {
    set_error("vertex degree %d / valence %d too large for the row kernel (%d bytes of shared memory)", max_deg,
              T.max_inc, S.total);
    return KNP_E_UNSUPPORTED;
  }
  static int configured = 0;
  if (S.total > 48 * 1024 && S.total > configured) {
    KNP_CUDA(cudaFuncSetAttribute(rows_kernel<D, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total));
    KNP_CUDA(cudaFuncSetAttribute(rows_kernel<D, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total));
    configured = S.total;
  }
  RowCoef C;
  for (int k = 0; k < 3; ++k) {
    C.dtD[k] = P.dt * P.D[k];
    C.cphi[k] = P.dt * P.D[k] * P.z[k] / P.psi;
    C.cpp[k] = P.dt * P.D[k] * P.z[k] * P.z[k] / P.psi;
    C.ck[k] = P.dt * P.z[k] * P.D[k];
    C.cmz[k] = P.C_M / (P.F * P.z[k]);
  }
  C.cf = P.C_M / P.F;
  const int nt0 = (T.L.n_own[0] + S.tile - 1) / S.tile, nt1 = (T.L.n_own[1] + S.tile - 1) / S.tile;
  if (S.use_lists) rows_kernel<D, MODE, true><<<nt0 + nt1, ROWS_THREADS, S.total, st>>>(T, C, u, fe, vals, b, S, nt0);
  else rows_kernel<D, MODE, false><<<nt0 + nt1, ROWS_THREADS, S.total, st>>>(T, C, u, fe, vals, b, S, nt0);
  KNP_LAUNCHED();
  return KNP_OK;
}

static EllSmem ell_layout(int mode, int max_deg, int max_gdeg, int max_inc, int Wp, int ELL_THREADS) {
  EllSmem S{};
  S.max_deg = max_deg;
  S.max_inc = max_inc;
  S.Wp = Wp;
  const size_t x = (size_t)3 * max_deg * ELL_THREADS * 8;
  // strip 0 serves fields 0 and 2 (ion rows), strip 1 fields 1 and 3 (the potential rows are the long ones)
  const size_t ion = mode == 0 ? 2 * max_deg + max_gdeg : max_deg;
  const size_t pot = mode == 0 ? 4 * max_deg + max_gdeg : max_deg;
  const size_t s0 = ((size_t)ELL_THREADS * ion + 4) * 8, s1 = ((size_t)ELL_THREADS * std::max(ion, pot) + 4) * 8;
  S.off_strip0 = (int)x;
  S.off_strip1 = (int)(x + ((s0 + 15) & ~(size_t)15));
  S.total = S.off_strip1 + (int)((s1 + 15) & ~(size_t)15);
  return S;
}

// dofs per CTA: 128 when three CTAs fit an SM, else 64 (3D: the X table and the strips grow with the degree); 0 = use
// the lane-group kernel
static int ell_block(int mode, int max_deg, int max_gdeg, int max_inc, int Wp) {
  if (ell_layout(mode, max_deg, max_gdeg, max_inc, Wp, 128).total <= 75 * 1024) return 128;
  if (ell_layout(mode, max_deg, max_gdeg, max_inc, Wp, 64).total <= 110 * 1024) return 64;
  return 0;
}

template <int D, int MODE, int TB>
static int launch_rows_ell_t(const DevTopo& T, const KParams& P, const double* u, const double* fe, double* vals,
                             double* b, int max_deg, int max_gdeg, cudaStream_t st) {
  const EllSmem S = ell_layout(MODE, max_deg, max_gdeg, T.max_inc, T.Wp, TB);
  static int configured = 0;
  if (S.total > 48 * 1024 && S.total > configured) {
    KNP_CUDA(cudaFuncSetAttribute(rows_ell_kernel<D, MODE, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total));
    configured = S.total;
  }
  RowCoef C;
  for (int k = 0; k < 3; ++k) {
    C.dtD[k] = P.dt * P.D[k];
    C.cphi[k] = P.dt * P.D[k] * P.z[k] / P.psi;
    C.cpp[k] = P.dt * P.D[k] * P.z[k] * P.z[k] / P.psi;
    C.ck[k] = P.dt * P.z[k] * P.D[k];
    C.cmz[k] = P.C_M / (P.F * P.z[k]);
  }
  C.cf = P.C_M / P.F;
  const int nb0 = (T.L.n_own[0] + TB - 1) / TB, nb1 = (T.L.n_own[1] + TB - 1) / TB;
  rows_ell_kernel<D, MODE, TB><<<nb0 + nb1, TB, S.total, st>>>(T, C, u, fe, vals, b, S, nb0);
  KNP_LAUNCHED();
  return KNP_OK;
}

template <int D, int MODE>
static int launch_rows_ell(const DevTopo& T, const KParams& P, const double* u, const double* fe, double* vals,
                           double* b, int max_deg, int max_gdeg, int tb, cudaStream_t st) {
  return tb == 128 ? launch_rows_ell_t<D, MODE, 128>(T, P, u, fe, vals, b, max_deg, max_gdeg, st)
                   : launch_rows_ell_t<D, MODE, 64>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
}

// static geometry tables of the ELL row kernel (setup, once per context)
int build_static_geometry(DevTopo& T, int max_deg, int32_t* adjE, uint32_t* incE, double* geoK, double* mslot,
                          double* kslot, cudaStream_t st) {
  if (T.n_work == 0) return KNP_OK;
  const int grid = (T.n_work + 127) / 128;
  if (T.gdim == 2) geo_build_kernel<2><<<grid, 128, 0, st>>>(T, T.Wp, max_deg, adjE, incE, geoK, mslot, kslot);
  else geo_build_kernel<3><<<grid, 128, 0, st>>>(T, T.Wp, max_deg, adjE, incE, geoK, mslot, kslot);
  KNP_LAUNCHED();
  T.adjE = adjE;
  T.incE = incE;
  T.geoK = geoK;
  T.mslot = mslot;
  T.kslot = kslot;
  return KNP_OK;
}

static int rows_ell_block(const DevTopo& T, int mode, int max_deg, int max_gdeg) {
  if (!T.adjE) return 0;        // static tables exist only when the context was created with KNP_ROWS=ell
  return ell_block(mode, max_deg, max_gdeg, T.max_inc, T.Wp);
}

int launch_rows(const DevTopo& T, const KParams& P, int mode, const double* u, const double* fe, double* vals,
                double* b, int max_deg, int max_gdeg, cudaStream_t st) {
  if (T.n_work == 0) return KNP_OK;
  if (const int tb = rows_ell_block(T, mode, max_deg, max_gdeg)) {
    if (T.gdim == 2) return mode == 0 ? launch_rows_ell<2, 0>(T, P, u, fe, vals, b, max_deg, max_gdeg, tb, st)
                                      : launch_rows_ell<2, 1>(T, P, u, fe, vals, b, max_deg, max_gdeg, tb, st);
    return mode == 0 ? launch_rows_ell<3, 0>(T, P, u, fe, vals, b, max_deg, max_gdeg, tb, st)
                     : launch_rows_ell<3, 1>(T, P, u, fe, vals, b, max_deg, max_gdeg, tb, st);
  }
  if (T.gdim == 2) return mode == 0 ? launch_rows_t<2, 0>(T, P, u, fe, vals, b, max_deg, max_gdeg, st)
                                    : launch_rows_t<2, 1>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
  return mode == 0 ? launch_rows_t<3, 0>(T, P, u, fe, vals, b, max_deg, max_gdeg, st)
                   : launch_rows_t<3, 1>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
}

int launch_csr_indices(const DevTopo& T, int mode, int32_t* indices, cudaStream_t st) {
  if (T.n_work == 0) return KNP_OK;
  const int grid = (T.n_work + 127) / 128;
  if (mode == 0) csr_indices_kernel<0><<<grid, 128, 0, st>>>(T, indices);
  else csr_indices_kernel<1><<<grid, 128, 0, st>>>(T, indices);
  KNP_LAUNCHED();
  return KNP_OK;
}

int launch_l2_cells(int gdim, const Layout& L, int s, int field, int power, int n_cells, const int32_t* cell_nodes,
                    const int32_t* cell_tag, const int32_t* cell_owned, const double* node_x, int nodeoff,
                    const int32_t* tags, int n_tags, const double* u, double* partial, int n_partial,
                    cudaStream_t st) {
  if (gdim == 2)
    l2_cells_kernel<2><<<n_partial, 256, 0, st>>>(L, s, field, power, n_cells, cell_nodes, cell_tag, cell_owned, node_x,
                                                  nodeoff, tags, n_tags, u, partial);
  else
    l2_cells_kernel<3><<<n_partial, 256, 0, st>>>(L, s, field, power, n_cells, cell_nodes, cell_tag, cell_owned, node_x,
                                                  nodeoff, tags, n_tags, u, partial);
  KNP_LAUNCHED();
  return KNP_OK;
}

}  // namespace knp
