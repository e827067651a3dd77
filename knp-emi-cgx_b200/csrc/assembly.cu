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
#include "p2.cuh"

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
    // alpha_{k,s} (KNPEMIx_problem.py:512-513,582-583); one reciprocal per side instead of three divisions
    double al[2][3];
    {
      double di = 0.0, de = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        di += P.D[k] * P.z[k] * P.z[k] * ciq[k];
        de += P.D[k] * P.z[k] * P.z[k] * ceq[k];
      }
      const double idi = 1.0 / di, ide = 1.0 / de;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        al[0][k] = (P.D[k] * P.z[k] * P.z[k] * ciq[k]) * idi;
        al[1][k] = (P.D[k] * P.z[k] * P.z[k] * ceq[k]) * ide;
      }
    }
    // Nernst potentials (KNPEMIx_problem.py:516): lg[k] = log(c_e / c_i) is shared with the cotransporter currents, whose
    // argument (K_i Cl_i) / (K_e Cl_e) is exp(-(lg[1] + lg[2]))
    double E[3], lg[3], ici[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ici[k] = 1.0 / ciq[k];
      lg[k] = log(ceq[k] * ici[k]);
      E[k] = (psi / P.z[k]) * lg[k];
    }
    double I[3] = {0.0, 0.0, 0.0};
    if (models & KNP_MODEL_NEURONAL_CT) {      // KNPEMIx_ionic_model.py:342-369 (f_NKCC1 == 0, :50-75)
      const double I_KCC2 = -0.0068 * (lg[1] + lg[2]);
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
      const double p2 = 1.0 + 10.0 * ici[0];
      const double I_ATP = 0.25 / ((p1 * p1) * (p2 * p2 * p2));
      I[0] += 3.0 * I_ATP;
      I[1] += -2.0 * I_ATP;
    }
    if (models & KNP_MODEL_GLIAL_CT) {         // :239-298 (f_NKCC1 == 0)
      const double I_KCC1 = -(7e-2 * psi) * (lg[1] + lg[2]);
      I[1] += I_KCC1;
      I[2] += -I_KCC1;
    }
    if (models & KNP_MODEL_KIRNA) {            // :117-222
      const double E_K_init = psi * log(P.K_e_init / P.K_i_g_init);
      const double rho = 1.1 * 1.12e-6;
      const double r = 10.0 * ici[0];
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

template <int D, int MODE>
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
  int deg = 0, gdeg = 0, self = -1, g = -1, ninc = 0, i0 = 0;
  if (node_ok) {
    const int w = w0 + lw;
    const int a0 = T.adj_ptr[w];
    deg = T.adj_ptr[w + 1] - a0;
    self = T.self_slot[w];
    g = T.mv_of_node[w];
    i0 = T.inc_ptr[w];
    ninc = T.inc_ptr[w + 1] - i0;
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

  // ---- phase 2a: one thread per (node, incident cell) ----
  for (int i = tid; i < (tile << lgI); i += ROWS_THREADS) {
    const int n = i >> lgI, j = i & (GI - 1);
    bool valid = n < nt;
    int ii0 = 0;
    if (valid) {
      ii0 = T.inc_ptr[w0 + n];
      valid = j < T.inc_ptr[w0 + n + 1] - ii0;
    }
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
#pragma unroll
      for (int b = 0; b < NV; ++b) {
        double dot = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) dot += gl[d] * G.g[b][d];
        r[b] = G.vol * dot;
      }
      r[NV] = G.vol * (1.0 / ((D + 1) * (D + 2)));
#pragma unroll
      for (int k = 0; k < 3; ++k) r[NV + 1 + k] = csum[k] * (1.0 / NV);
    }
  }
  __syncthreads();

  // ---- phase 2b: accumulate per (node, slot) in registers, cells in ascending order ----
  double a_m = 0.0, a_kk = 0.0, X[3] = {0.0, 0.0, 0.0}, kphi_m[3] = {0.0, 0.0, 0.0}, pp_m = 0.0;
  double bmem[4] = {0.0, 0.0, 0.0, 0.0}, ce[3] = {0.0, 0.0, 0.0};
  double ga[3] = {0.0, 0.0, 0.0}, g1 = 0.0;
  const bool is_self = has_ent && e == self;
  const bool has_gam = MODE == 0 && node_ok && e < gdeg;
  const uint32_t rep = (uint32_t)e * 0x01010101u;
  if (has_ent) {
#pragma unroll
    for (int k = 0; k < 3; ++k) ce[k] = nbr[(size_t)tid * NB + D + k];
    const double* rbase = res + (size_t)lw * S.res_stride;
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
  if (T.p2) return launch_facets_p2(*T.p2, P, tag_models, tag_stim, u, gates, stim_fac, fe, st);
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
// Row kernel, edge-lane form (the default; the scan kernel above serves meshes whose tables do not fit).
//
// Same ownership as the scan kernel -- one CTA per tile of consecutive owned dofs, G = 2^LG lanes per dof, lane e owns
// adjacency slot e, i.e. one column position of all ten block rows -- but a lane computes its five sums itself from the
// cells around ITS edge (dof p, neighbour q): the setup (topology.cpp) stores per (dof, slot) the cells that contain the
// edge, each as the adjacency slots of the cell's other vertices, so the lane reads those vertices from the staged
// neighbour block and evaluates the P1 entries in closed form
//     2D  cell (p, q, r):      K_pq = -(p - r).(q - r) / (2 |(p - r) x (q - r)|),      |cell| = |(p - r) x (q - r)| / 2
//     3D  cell (p, q, r, s):   K_pq = -(n_p . n_q) / (6 |J|),  n_q = (r - p) x (s - p),  n_p = (r - q) x (s - q),
//                              J = (q - p) . n_q,                                      |cell| = |J| / 6
// (K_pq = |cell| grad(lambda_p) . grad(lambda_q), the entry the scan kernel forms from the barycentric gradients).  The
// sums of the self slot follow from the row-sum identities of the element matrices, K_pp = -sum_q K_pq per cell and
// M_pp = (2/d) sum_q M_pq, with one butterfly over the lane group.  Against the scan kernel this removes the per-(dof, cell)
// result block in shared memory, the byte-compare scan over all incident cells in every lane, one of the three CTA-wide
// barriers and one level of the index -> neighbour -> value load chain (the lane-group tables are read at
// (tile base << LG) + thread id: fully coalesced, no row-pointer lookup first).  Fixed summation order, no atomics:
// bitwise reproducible.  Membrane terms, row formation, TMA bulk stores and the right-hand side are those of the scan kernel.
#ifndef EDGE_MIN_CTAS
#define EDGE_MIN_CTAS 3
#endif
#ifndef EDGE_MIN_CTAS_3D
#define EDGE_MIN_CTAS_3D 2      // the 3D cell formulas plus the prefetched tables need ~100 registers
#endif
#ifndef EDGE_THREADS
#define EDGE_THREADS 256
#endif
constexpr int EDGE_NB = 6;      // doubles per staged neighbour: 2D {x, y, c0, c1, c2, -}, 3D {x, y, z, c0, c1, c2} (16-byte units)

// The kernel is PERSISTENT, software-pipelined and WARP-AUTONOMOUS.  A lane group never spans warps (G <= 32), so a warp
// owns a mini-tile of 32 / G consecutive dofs with its own neighbour buffers and staging strip in shared memory and
// needs no CTA-wide barrier at all: the warps of the grid walk over mini-tiles w, w + W, w + 2W, ... independently.  While
// mini-tile i is computed, the neighbour data of mini-tile i + 1 streams into the warp's second buffer with cp.async
// (LDGSTS: global -> shared without passing through registers) and the lane-group tables of mini-tile i + 2 are on their
// way into registers, so neither level of the index -> neighbour load chain is on the critical path, and the finished rows
// leave through per-warp TMA bulk stores.  (Round-2 profiles: the one-shot CTA form spent 35 % of its stall samples on
// those two loads; the persistent CTA form that hid them spent 22-28 % on the two CTA barriers per tile instead.)
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}

template <int D>
struct EdgeTables {      // what one lane holds of a tile before the tile is computed
  int q;                 // neighbour node of (dof, slot), -1: none
  uint32_t hit[D == 2 ? 1 : 4];
  int2 meta;
  int rs_a, rs_b;        // CSR row start of (field e, dof lw) for e < 4, and the end of the tile's last row (dof nt - 1)
};

template <int D, int MODE, int LG>
__global__ void __launch_bounds__(EDGE_THREADS, D == 3 ? EDGE_MIN_CTAS_3D : EDGE_MIN_CTAS) rows_edge_kernel(DevTopo T, RowCoef C, const double* __restrict__ u,
                                                                                const double* __restrict__ fe,
                                                                                double* __restrict__ vals, double* __restrict__ bvec,
                                                                                int ntile0, int ntiles, int stg_doubles) {
  constexpr int G = 1 << LG, TILE = 32 / G;                                      // dofs per warp (mini-tile)
  constexpr int NS = D * (D + 1) / 2;
  constexpr int NB = EDGE_NB;
  constexpr int HW = D == 2 ? 1 : 4;
  constexpr uint32_t FMASK = D == 3 ? 0x00FFFFFFu : 0x0000FFFFu;
  extern __shared__ __align__(16) unsigned char smraw[];
  const int wid = threadIdx.x >> 5, tid = threadIdx.x & 31;                      // warp in the CTA, lane
  // neighbour buffers of all warps first (compile-time stride: the compiler re-derives these addresses instead of holding
  // them in registers), then the staging strips
  double* nbr2 = reinterpret_cast<double*>(smraw) + wid * (2 * 32 * NB);         // [2][32][NB] of this warp
  double* stg = reinterpret_cast<double*>(smraw) + (EDGE_THREADS / 32) * (2 * 32 * NB) + (size_t)wid * stg_doubles;
  const int lw = tid >> LG, e = tid & (G - 1);
  const int* __restrict__ iptr = MODE == 0 ? T.indptr : T.indptr_P;
  const int stride = gridDim.x * (EDGE_THREADS / 32);                            // warps in the grid

  auto load_tables = [&](int t, EdgeTables<D>& E) {
    E.q = -1;
    E.meta = make_int2(0, -1);
    E.rs_a = E.rs_b = 0;
#pragma unroll
    for (int i = 0; i < HW; ++i) E.hit[i] = 0xFFFFFFFFu;
    if (t >= ntiles) return;
    const int s = t >= ntile0 ? 1 : 0;
    const int p0 = (t - (s ? ntile0 : 0)) * TILE;
    const int nt = min(TILE, T.L.n_own[s] - p0);
    const int w0 = (s ? T.L.n_own[0] : 0) + p0;
    if (lw < nt) {
      const size_t at = ((size_t)w0 << LG) + tid;
      E.q = T.adjG[at];
      E.meta = T.metaG[w0 + lw];
      if (D == 2) {
        E.hit[0] = T.hitG[at];
      } else {
        const uint4 h4 = reinterpret_cast<const uint4*>(T.hitG)[at];
        E.hit[0] = h4.x;
        E.hit[HW - 3 > 0 ? 1 : 0] = h4.y;
        E.hit[HW - 2 > 0 ? 2 : 0] = h4.z;
        E.hit[HW - 1 > 0 ? 3 : 0] = h4.w;
      }
      if (e < 4) {
        E.rs_a = iptr[T.L.row(s, e, p0 + lw)];
        if (lw == nt - 1) E.rs_b = iptr[T.L.row(s, e, p0 + nt)];
      }
    }
  };
  // neighbour (s, q) of tile t -> this lane's entry of the given buffer, asynchronously
  auto gather = [&](int t, int q, double* buf) {
    if (q >= 0) {
      const int s = t >= ntile0 ? 1 : 0;
      const int n_own_s = T.L.n_own[s];
      const int nodeoff = s ? T.L.n_loc[0] : 0;
      double* o = buf + (size_t)tid * NB;
      const double* __restrict__ xs = T.node_x + (size_t)(nodeoff + q) * D;
      const double* __restrict__ uc = q < n_own_s ? u + T.L.rowbase[s] + q : u + T.L.n_rows + T.L.gbase[s] + (q - n_own_s);
      const size_t fstride = q < n_own_s ? n_own_s : T.L.n_gh[s];
      if (D == 2) {
        cp_async16(o, xs);
      } else {
        cp_async8(o, xs);
        cp_async8(o + 1, xs + 1);
        cp_async8(o + 2, xs + 2);
      }
      cp_async8(o + D, uc);
      cp_async8(o + D + 1, uc + fstride);
      cp_async8(o + D + 2, uc + 2 * fstride);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int t = blockIdx.x * (EDGE_THREADS / 32) + wid;
  EdgeTables<D> cur, nxt;
  load_tables(t, cur);
  gather(t, cur.q, nbr2);
  load_tables(t + stride, nxt);

  for (int it = 0; t < ntiles; t += stride, ++it) {
    double* nbr = nbr2 + (size_t)(it & 1) * 32 * NB;
    const int s = t >= ntile0 ? 1 : 0;
    const int p0 = (t - (s ? ntile0 : 0)) * TILE;
    const int nt = min(TILE, T.L.n_own[s] - p0);
    const bool node_ok = lw < nt;
    asm volatile("cp.async.wait_group 0;" ::: "memory");                 // this lane's neighbour entry of the tile has landed
    if (tid < 4 && it > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging strip is free again
    __syncwarp();
    // ---- prefetch: neighbour data of the next mini-tile, tables of the one after it ----
    gather(t + stride, nxt.q, nbr2 + (size_t)((it + 1) & 1) * 32 * NB);
    EdgeTables<D> nn;
    load_tables(t + 2 * stride, nn);

    const int deg = cur.meta.x & 255, self = (cur.meta.x >> 8) & 255;
    const int gdeg = MODE == 0 ? (cur.meta.x >> 16) & 255 : 0;
    const int g = cur.meta.y;
    const bool has_ent = cur.q >= 0;

    // ---- phase 2: the lane's five sums over the cells around its edge, ascending cell order ----
    double a_m = 0.0, a_kk = 0.0, X[3] = {0.0, 0.0, 0.0}, kphi_m[3] = {0.0, 0.0, 0.0}, pp_m = 0.0;
    double bmem[4] = {0.0, 0.0, 0.0, 0.0};
    double ga[3] = {0.0, 0.0, 0.0}, g1 = 0.0;
    double ce[3] = {0.0, 0.0, 0.0};
    const bool is_self = has_ent && e == self;
    const bool has_gam = MODE == 0 && node_ok && e < gdeg;
    const uint32_t rep = (uint32_t)e * 0x01010101u;
    const double* grp = nbr + (size_t)(lw << LG) * NB;              // the dof's neighbour block
    if (has_ent) {
      const double2* qn = reinterpret_cast<const double2*>(nbr + (size_t)tid * NB);
      const double2* pn = reinterpret_cast<const double2*>(grp + (size_t)self * NB);
      if (D == 2) {
        const double2 xq = qn[0], q01 = qn[1];
        ce[0] = q01.x;
        ce[1] = q01.y;
        ce[2] = qn[2].x;
        if (!is_self) {
          const double2 xp = pn[0], c01 = pn[1];
          const double s0 = c01.x + ce[0], s1 = c01.y + ce[1], s2 = pn[2].x + ce[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t sl = (cur.hit[0] >> (8 * h)) & 255u;
            if (sl != 255u) {
              const double2* rn = reinterpret_cast<const double2*>(grp + (size_t)sl * NB);
              const double2 xr = rn[0], r01 = rn[1];
              const double r2 = rn[2].x;
              const double e1x = xp.x - xr.x, e1y = xp.y - xr.y, e2x = xq.x - xr.x, e2y = xq.y - xr.y;
              const double cr = fabs(e1x * e2y - e1y * e2x);
              const double dt = e1x * e2x + e1y * e2y;
              const double kab = -0.5 * dt * (1.0 / cr);
              a_m += cr * (1.0 / 24.0);
              a_kk += kab;
              X[0] += ((s0 + r01.x) * (1.0 / 3.0)) * kab;
              X[1] += ((s1 + r01.y) * (1.0 / 3.0)) * kab;
              X[2] += ((s2 + r2) * (1.0 / 3.0)) * kab;
            }
          }
        }
      } else {
        const double2 q01 = qn[0], q23 = qn[1], q45 = qn[2];          // {x, y} {z, c0} {c1, c2}
        ce[0] = q23.y;
        ce[1] = q45.x;
        ce[2] = q45.y;
        if (!is_self) {
          const double2 p01 = pn[0], p23 = pn[1], p45 = pn[2];
          const double ax = q01.x - p01.x, ay = q01.y - p01.y, az = q23.x - p23.x;
          const double s0 = p23.y + ce[0], s1 = p45.x + ce[1], s2 = p45.y + ce[2];
#pragma unroll 2
          for (int h = 0; h < 8; ++h) {
            const uint32_t wv = (h >> 1) == 0 ? cur.hit[0] : (h >> 1) == 1 ? cur.hit[HW - 3 > 0 ? 1 : 0]
                              : (h >> 1) == 2 ? cur.hit[HW - 2 > 0 ? 2 : 0] : cur.hit[HW - 1 > 0 ? 3 : 0];
            const uint32_t code = (wv >> (16 * (h & 1))) & 0xFFFFu;
            if (code == 0xFFFFu) break;
            const double2* rn = reinterpret_cast<const double2*>(grp + (size_t)(code & 255u) * NB);
            const double2* sn = reinterpret_cast<const double2*>(grp + (size_t)(code >> 8) * NB);
            const double2 r01 = rn[0], r23 = rn[1], r45 = rn[2];
            const double2 t01 = sn[0], t23 = sn[1], t45 = sn[2];
            const double bx = r01.x - p01.x, by = r01.y - p01.y, bz = r23.x - p23.x;
            const double cx = t01.x - p01.x, cy = t01.y - p01.y, cz = t23.x - p23.x;
            const double nqx = by * cz - bz * cy, nqy = bz * cx - bx * cz, nqz = bx * cy - by * cx;      // (r - p) x (s - p)
            const double J = fabs(ax * nqx + ay * nqy + az * nqz);
            const double ux = bx - ax, uy = by - ay, uz = bz - az, vx = cx - ax, vy = cy - ay, vz = cz - az;
            const double npx = uy * vz - uz * vy, npy = uz * vx - ux * vz, npz = ux * vy - uy * vx;      // (r - q) x (s - q)
            const double kab = -(npx * nqx + npy * nqy + npz * nqz) * (1.0 / (6.0 * J));
            a_m += J * (1.0 / 120.0);
            a_kk += kab;
            X[0] += ((s0 + r23.y + t23.y) * 0.25) * kab;
            X[1] += ((s1 + r45.x + t45.x) * 0.25) * kab;
            X[2] += ((s2 + r45.y + t45.y) * 0.25) * kab;
          }
        }
      }
    }
    {
      // self slot: K_pp = -sum_q K_pq, (cbar K)_pp = -sum_q (cbar K)_pq, and 2 sum_c |c|/((d+1)(d+2)) = (2/d) sum_q M_pq
      double ts[5] = {a_m, a_kk, X[0], X[1], X[2]};
#pragma unroll
      for (int off = G >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 5; ++i) ts[i] += __shfl_xor_sync(0xffffffffu, ts[i], off);
      }
      if (is_self) {
        a_m = ts[0] * (2.0 / D);
        a_kk = -ts[1];
#pragma unroll
        for (int k = 0; k < 3; ++k) X[k] = -ts[2 + k];
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

    // ---- phase 3: form the ten block rows at their CSR-relative offsets of the staging strip ----
    // CSR row starts of the mini-tile's rows travel by shuffle from the lanes that loaded them (lane (dof, field))
    int so[4], base[4], total[4], rsf4[4];
    {
      int acc = 0;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        base[f] = __shfl_sync(0xffffffffu, cur.rs_a, f);
        total[f] = __shfl_sync(0xffffffffu, cur.rs_b, ((nt - 1) << LG) + f) - base[f];
        rsf4[f] = __shfl_sync(0xffffffffu, cur.rs_a, (lw << LG) + f) - base[f];
        so[f] = acc;                                   // even: 16-byte aligned start of the field's strip
        acc += (total[f] + 3) & ~1;                    // room for the phase shift (base & 1), rounded to even
      }
    }
    if (node_ok) {
      const int goff = s == 1 ? gdeg : 0;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int rsf = rsf4[f];
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
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the TMA (async proxy)
    __syncwarp();

    // ---- copy-out: one TMA bulk store per field (threads 0..3) for the 16-byte aligned body, scalar head/tail ----
    if (tid < 4) {
      const int bf = tid == 0 ? base[0] : tid == 1 ? base[1] : tid == 2 ? base[2] : base[3];
      const int tf = tid == 0 ? total[0] : tid == 1 ? total[1] : tid == 2 ? total[2] : total[3];
      const int sof = tid == 0 ? so[0] : tid == 1 ? so[1] : tid == 2 ? so[2] : so[3];
      const int sh = bf & 1;
      const int nbody = (tf - sh) & ~1;
      if (nbody > 0) bulk_store(vals + (size_t)bf + sh, stg + sof + 2 * sh, (uint32_t)nbody * 8u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (sh && tf > 0) vals[(size_t)bf] = stg[sof + 1];
      if ((tf - sh) & 1) vals[(size_t)bf + tf - 1] = stg[sof + sh + tf - 1];
    }
    // right-hand side: b_k = sum_e m_e c_k(e) (KNPEMIx_problem.py:613-614,641-642): segmented warp-shuffle reduction over
    // the dof's lane group (fixed butterfly order -> reproducible)
    if (MODE == 0) {
      double bk[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) bk[k] = has_ent ? a_m * ce[k] : 0.0;
#pragma unroll
      for (int off = G >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bk[k] += __shfl_xor_sync(0xffffffffu, bk[k], off);
      }
      if (is_self) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bvec[T.L.row(s, k, p0 + lw)] = bk[k] + bmem[k];
        bvec[T.L.row(s, 3, p0 + lw)] = bmem[3];
      }
    }
    cur = nxt;
    nxt = nn;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (tid < 4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

static RowCoef make_coef(const KParams& P) {
  RowCoef C;
  for (int k = 0; k < 3; ++k) {
    C.dtD[k] = P.dt * P.D[k];
    C.cphi[k] = P.dt * P.D[k] * P.z[k] / P.psi;
    C.cpp[k] = P.dt * P.D[k] * P.z[k] * P.z[k] / P.psi;
    C.ck[k] = P.dt * P.z[k] * P.D[k];
    C.cmz[k] = P.C_M / (P.F * P.z[k]);
  }
  C.cf = P.C_M / P.F;
  return C;
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
  const size_t work = (nbr + res + packed + 15) & ~(size_t)15;
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
  if (S.total > 227 * 1024 || (1 << S.lgG) > ROWS_THREADS) {
    set_error("vertex degree %d / valence %d too large for the row kernel (%d bytes of shared memory)", max_deg,
              T.max_inc, S.total);
    return KNP_E_UNSUPPORTED;
  }
  static int configured = 0;
  if (S.total > 48 * 1024 && S.total > configured) {
    KNP_CUDA(cudaFuncSetAttribute(rows_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total));
    configured = S.total;
  }
  const RowCoef C = make_coef(P);
  const int nt0 = (T.L.n_own[0] + S.tile - 1) / S.tile, nt1 = (T.L.n_own[1] + S.tile - 1) / S.tile;
  rows_kernel<D, MODE><<<nt0 + nt1, ROWS_THREADS, S.total, st>>>(T, C, u, fe, vals, b, S, nt0);
  KNP_LAUNCHED();
  return KNP_OK;
}

template <int D, int MODE, int LG>
static int launch_rows_edge_t(const DevTopo& T, const KParams& P, const double* u, const double* fe, double* vals,
                              double* b, int max_deg, int max_gdeg, cudaStream_t st) {
  constexpr int TILE = 32 >> LG;                                                  // dofs per warp
  constexpr int WARPS = EDGE_THREADS / 32;
  // per warp: double-buffered neighbour block + staging strip of the mini-tile's rows (16-byte units)
  const int stg_doubles = (TILE * (MODE == 0 ? 10 * max_deg + 4 * max_gdeg : 4 * max_deg) + 16 + 1) & ~1;
  const int total = WARPS * (2 * 32 * EDGE_NB + stg_doubles) * 8;
  if (total > 227 * 1024) {
    set_error("vertex degree %d too large for the row kernel (%d bytes of shared memory)", max_deg, total);
    return KNP_E_UNSUPPORTED;
  }
  static int configured = -1, per_sm = 0;
  if (total > configured) {
    KNP_CUDA(cudaFuncSetAttribute(rows_edge_kernel<D, MODE, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, total));
    KNP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rows_edge_kernel<D, MODE, LG>, EDGE_THREADS, total));
    configured = total;
  }
  const int nt0 = (T.L.n_own[0] + TILE - 1) / TILE, nt1 = (T.L.n_own[1] + TILE - 1) / TILE;
  static const int sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  // persistent grid: every resident CTA slot of the device; the warps deal the mini-tiles round-robin
  const int grid = std::min((nt0 + nt1 + WARPS - 1) / WARPS, sms * std::max(per_sm, 1));
  rows_edge_kernel<D, MODE, LG><<<grid, EDGE_THREADS, total, st>>>(T, make_coef(P), u, fe, vals, b, nt0, nt0 + nt1, stg_doubles);
  KNP_LAUNCHED();
  return KNP_OK;
}

template <int D, int MODE>
static int launch_rows_edge(const DevTopo& T, const KParams& P, const double* u, const double* fe, double* vals, double* b,
                            int max_deg, int max_gdeg, cudaStream_t st) {
  switch (T.lgG) {
    case 2: return launch_rows_edge_t<D, MODE, 2>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
    case 3: return launch_rows_edge_t<D, MODE, 3>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
    case 4: return launch_rows_edge_t<D, MODE, 4>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
    case 5: return launch_rows_edge_t<D, MODE, 5>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
  }
  set_error("edge-lane row kernel: unsupported lane group 2^%d", T.lgG);
  return KNP_E_UNSUPPORTED;
}

int launch_rows(const DevTopo& T, const KParams& P, int mode, const double* u, const double* fe, double* vals,
                double* b, int max_deg, int max_gdeg, cudaStream_t st) {
  if (T.p2) return launch_rows_p2(*T.p2, P, mode, u, fe, vals, b, st);
  if (T.n_work == 0) return KNP_OK;
  if (T.adjG) {      // edge-lane kernel (tables exist: every edge ring fits, lane group <= one warp)
    if (T.gdim == 2) return mode == 0 ? launch_rows_edge<2, 0>(T, P, u, fe, vals, b, max_deg, max_gdeg, st)
                                      : launch_rows_edge<2, 1>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
    return mode == 0 ? launch_rows_edge<3, 0>(T, P, u, fe, vals, b, max_deg, max_gdeg, st)
                     : launch_rows_edge<3, 1>(T, P, u, fe, vals, b, max_deg, max_gdeg, st);
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
