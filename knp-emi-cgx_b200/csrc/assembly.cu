// Per-timestep assembly kernels for sm_100a.
//
//   gate_kernel    K4  Rush-Larsen / forward-Euler gate ODE on membrane vertices
//                      (HodgkinHuxley.update_gating_variables, KNPEMIx_ionic_model.py:605-671)
//   facet_kernel   K3  membrane-facet element tensors: alpha_k, Nernst potentials, channel currents at
//                      the facet quadrature points (dS terms of KNPEMIx_problem.py:594-642 with the
//                      IonicModel._eval family), written to a facet-major SoA staging buffer
//   rows_kernel    K1/K2  one thread per restricted dof ("node"): loops over its incident cells in a
//                      fixed order, recomputes the P1 element row from the vertex coordinates, accumulates
//                      per-adjacency-slot values in a private shared-memory strip (no atomics), adds the
//                      membrane-facet rows, and then each warp streams the finished CSR rows out
//                      (every A value and b entry is written exactly once -> bitwise reproducible).
//   csr_indices_kernel   column indices of A / P from the node adjacency (setup)
//
// Design note: all ten (d+1)x(d+1) cell blocks of KNPEMIx_problem.py:598-605,633-634 are linear
// combinations of M^T, K^T and cbar_k K^T, and each block row shares the node's adjacency list, so the
// "cell -> nnz map" collapses to one byte per (node, cell, local vertex): the adjacency slot.
#include <cstdlib>
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
  double ci[3][D], ce[3][D], pm[D], gn[D], gm[D], gh[D], xs[D];
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
    xs[a] = P.stim_dir >= 0 ? T.node_x[(size_t)qi * D + P.stim_dir] : 0.0;
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
    double ciq[3], ceq[3], pmq = 0.0, nq = 0.0, mq = 0.0, hq = 0.0, xq = 0.0;
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
      xq += lam[a] * xs[a];
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
        const double mask = (P.stim_dir < 0 || (xq > P.stim_lo && xq < P.stim_hi)) ? 1.0 : 0.0;
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
template <int D, int MODE>
__global__ void __launch_bounds__(ROWS_BLOCK) rows_kernel(DevTopo T, KParams P, const double* __restrict__ u,
                                                          const double* __restrict__ fe,
                                                          double* __restrict__ vals, double* __restrict__ bvec,
                                                          int stride, int group, int stage_len, int nb0) {
  constexpr int NV = D + 1;
  constexpr int NS = D * (D + 1) / 2;
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  // blocks [0, nb0) serve the intracellular dofs, the rest the extracellular ones, so that the rows a warp
  // produces for one field are one contiguous CSR span
  const int s = blockIdx.x >= nb0 ? 1 : 0;
  const int p = (blockIdx.x - (s ? nb0 : 0)) * blockDim.x + tid;
  const bool active = p < T.L.n_own[s];
  const int w = (s ? T.L.n_own[0] : 0) + p;
  double* acc = sm + (size_t)tid * stride;
  double* stg = sm + (size_t)blockDim.x * stride + (size_t)(tid >> 5) * stage_len;

  int deg = 0, gdeg = 0, g = -1;
  if (active) {
    const int a0 = T.adj_ptr[w];
    deg = T.adj_ptr[w + 1] - a0;
    g = T.mv_of_node[w];
    gdeg = g >= 0 ? T.gam_ptr[g + 1] - T.gam_ptr[g] : 0;
    const int nacc = 6 * deg + 4 * gdeg;
    for (int i = 0; i < nacc; ++i) acc[i] = 0.0;
    double* a_m = acc;
    double* a_kk = acc + deg;
    double* a_kphi = acc + 2 * deg;      // [3][deg]
    double* a_pp = acc + 5 * deg;
    double* a_ga = acc + 6 * deg;        // [3][gdeg]
    double* a_g1 = acc + 6 * deg + 3 * gdeg;
    const int self = T.self_slot[w];
    const int nodeoff = s ? T.L.n_loc[0] : 0;
    // column of (s, k, q): owned dofs are contiguous per field; ghosts live in the tail of the column layout
    const int n_own_s = T.L.n_own[s], n_gh_s = T.L.n_gh[s];
    const double* __restrict__ u_own = u + T.L.rowbase[s];
    const double* __restrict__ u_gh = u + T.L.n_rows + T.L.gbase[s] - n_own_s;
    double bk[3] = {0.0, 0.0, 0.0}, bp = 0.0;
    double cphi[3], cpp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      cphi[k] = P.dt * P.D[k] * P.z[k] / P.psi;
      cpp[k] = P.dt * P.D[k] * P.z[k] * P.z[k] / P.psi;
    }
    const double mfac = 1.0 / ((D + 1) * (D + 2));
    // ---- cell (dx) terms: KNPEMIx_problem.py:598,600,603,605,633-634 ----
    const int i1 = T.inc_ptr[w + 1];
    for (int inc = T.inc_ptr[w]; inc < i1; ++inc) {
      const uint32_t packed = T.inc_slots[inc];
      int sl[NV];
      double x[NV][D], c[3][NV];
      int la = 0;
#pragma unroll
      for (int b = 0; b < NV; ++b) {
        sl[b] = (packed >> (8 * b)) & 255u;
        if (sl[b] == self) la = b;
        const int q = T.adj_idx[a0 + sl[b]];
#pragma unroll
        for (int i = 0; i < D; ++i) x[b][i] = T.node_x[(size_t)(nodeoff + q) * D + i];
        if (q < n_own_s) {
#pragma unroll
          for (int k = 0; k < 3; ++k) c[k][b] = u_own[k * n_own_s + q];
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k) c[k][b] = u_gh[k * n_gh_s + q];
        }
      }
      CellGeom<D> G;
      cell_geometry(x, G);
      double cbar[3], csum[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < NV; ++b) t += c[k][b];
        csum[k] = t;
        cbar[k] = t * (1.0 / NV);
      }
      const double mv = G.vol * mfac;
      // gradient of this node's own basis function (runtime local index la -> register select)
      double gl[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double t = G.g[0][i];
#pragma unroll
        for (int a = 1; a < NV; ++a) t = (la == a) ? G.g[a][i] : t;
        gl[i] = t;
      }
#pragma unroll
      for (int b = 0; b < NV; ++b) {
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) dot += gl[i] * G.g[b][i];
        const double Kab = G.vol * dot;
        const int e = sl[b];
        a_m[e] += (b == la) ? 2.0 * mv : mv;
        a_kk[e] += Kab;
        double kp = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (MODE == 0) a_kphi[k * deg + e] += cphi[k] * cbar[k] * Kab;
          kp += cpp[k] * cbar[k] * Kab;
        }
        a_pp[e] += kp;
      }
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double cl = c[k][0];
#pragma unroll
          for (int a = 1; a < NV; ++a) cl = (la == a) ? c[k][a] : cl;
          bk[k] += mv * (csum[k] + cl);
        }
      }
    }
    // ---- membrane (dS) terms: KNPEMIx_problem.py:599,604,609-610,637-638,641-642 (P: :737-738) ----
    if (g >= 0) {
      const size_t nf = (size_t)T.n_mf;
      const double sgn = s == 0 ? 1.0 : -1.0;
      const double cf = P.C_M / P.F;
      const int m1 = T.minc_ptr[g + 1];
      for (int mi = T.minc_ptr[g]; mi < m1; ++mi) {
        const uint4 rec = reinterpret_cast<const uint4*>(T.minc)[mi];
        const int f = (int)rec.x;
        const int a = rec.y & 255u;
        const uint32_t ss = s == 0 ? (rec.y >> 8) : rec.z;
        const uint32_t gs = rec.w;
        const double area = T.mf_area[f];
#pragma unroll
        for (int b = 0; b < D; ++b) {
          const int es = (ss >> (8 * b)) & 255u;
          const int eg = (gs >> (8 * b)) & 255u;
          const double G1 = area * ((a == b) ? 2.0 : 1.0) / (D * (D + 1));
          if (MODE == 0) {
            const int ab = a <= b ? symidx(a, b, D) : symidx(b, a, D);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const double v = (P.C_M / (P.F * P.z[k])) * fe[(size_t)((s * 3 + k) * NS + ab) * nf + f];
              a_kphi[k * deg + es] += v;
              a_ga[k * gdeg + eg] += v;
            }
            a_pp[es] += cf * G1;
            a_g1[eg] += cf * G1;
          } else {
            a_pp[es] -= cf * G1;
          }
        }
        if (MODE == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) bk[k] -= sgn * fe[(size_t)(6 * NS + (s * 3 + k) * D + a) * nf + f];
          bp -= sgn * fe[(size_t)(6 * NS + 6 * D + a) * nf + f];
        }
      }
    }
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) bvec[T.L.row(s, k, p)] = bk[k];
      bvec[T.L.row(s, 3, p)] = bp;
    }
  }
  __syncwarp();
  // ---- output: each thread writes its finished row into the warp's staging strip at its CSR-relative offset, then
  //      the warp copies the contiguous span to global memory with fully coalesced stores (one pass per field) ----
  const int* __restrict__ iptr = MODE == 0 ? T.indptr : T.indptr_P;
  const double* a_m = acc;
  const double* a_kk = acc + deg;
  const double* a_kphi = acc + 2 * deg;
  const double* a_pp = acc + 5 * deg;
  const double* a_ga = acc + 6 * deg;
  const double* a_g1 = acc + 6 * deg + 3 * gdeg;
  const int gd = MODE == 0 ? gdeg : 0;
#pragma unroll 1
  for (int f = 0; f < 4; ++f) {
    const int nseg = MODE == 0 ? (f < 3 ? 2 : 4) : 1;
    const int rs = active ? iptr[T.L.row(s, f, p)] : 0;
    const int len = active ? nseg * deg + gd : 0;
#pragma unroll 1
    for (int g0 = 0; g0 < 32; g0 += group) {
      const int base = __shfl_sync(0xffffffffu, rs, g0);
      const bool mine = active && lane >= g0 && lane < g0 + group;
      if (mine) {
        double* o = stg + (rs - base);
        if (MODE == 0) {
          if (f < 3) {
            if (s == 1)
              for (int e = 0; e < gd; ++e) *o++ = -a_ga[f * gdeg + e];
            const double dk = P.dt * P.D[f];
            for (int e = 0; e < deg; ++e) *o++ = a_m[e] + dk * a_kk[e];
            for (int e = 0; e < deg; ++e) *o++ = a_kphi[f * deg + e];
            if (s == 0)
              for (int e = 0; e < gd; ++e) *o++ = -a_ga[f * gdeg + e];
          } else {
            if (s == 1)
              for (int e = 0; e < gd; ++e) *o++ = -a_g1[e];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const double ck = P.dt * P.z[k] * P.D[k];
              for (int e = 0; e < deg; ++e) *o++ = ck * a_kk[e];
            }
            for (int e = 0; e < deg; ++e) *o++ = a_pp[e];
            if (s == 0)
              for (int e = 0; e < gd; ++e) *o++ = -a_g1[e];
          }
        } else {
          if (f < 3) {
            const double dk = P.dt * P.D[f];
            for (int e = 0; e < deg; ++e) *o++ = a_m[e] + dk * a_kk[e];
          } else {
            for (int e = 0; e < deg; ++e) *o++ = a_pp[e];
          }
        }
      }
      const int total = __reduce_max_sync(0xffffffffu, mine ? rs + len - base : 0);
      __syncwarp();
      // coalesced copy-out, 128-bit where the global address allows it (the strip itself is 16-byte aligned)
      {
        const int head = base & 1;                      // vals + base is 16-byte aligned iff base is even
        if (head && lane == 0 && total > 0) vals[(size_t)base] = stg[0];
        const int npair = (total - head) >> 1;
        if (head == 0) {
          const double2* s2 = reinterpret_cast<const double2*>(stg);
          double2* g2 = reinterpret_cast<double2*>(vals + (size_t)base);
          for (int q = lane; q < npair; q += 32) g2[q] = s2[q];
        } else {
          double2* g2 = reinterpret_cast<double2*>(vals + (size_t)base + 1);
          for (int q = lane; q < npair; q += 32) g2[q] = make_double2(stg[2 * q + 1], stg[2 * q + 2]);
        }
        if (((total - head) & 1) && lane == 0) vals[(size_t)base + total - 1] = stg[total - 1];
      }
      __syncwarp();
    }
  }
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
// int u^2 over cells (tests/KNPEMI/electric_potential_norms_direct_solver.py:45-51): per-block partial sums,
// reduced in a fixed order by reduce_partials_kernel (linalg.cu).
template <int D>
__global__ void __launch_bounds__(256) l2_cells_kernel(Layout L, int s, int field, int n_cells,
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
    acc += G.vol / ((D + 1) * (D + 2)) * (s2 + s1 * s1);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
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

int rows_smem_stride(int max_deg, int max_gdeg) {
  int st = 6 * max_deg + 4 * max_gdeg;
  if (st % 2 == 0) ++st;   // odd stride (in doubles): conflict-free when all lanes touch the same offset
  return st;
}

// block size / staging group for the row kernel from the mesh's maximum degrees (shared-memory budget)
void rows_config(int max_deg, int max_gdeg, int mode, int& block, int& group, int& stage_len, size_t& smem) {
  const int stride = rows_smem_stride(max_deg, max_gdeg);
  const int maxrow = mode == 0 ? 4 * max_deg + max_gdeg : max_deg;
  block = 128;
  while (block > 32 && (size_t)block * stride * 8 > 64 * 1024) block >>= 1;
  const size_t acc = (size_t)block * stride * 8;
  const size_t budget = acc <= 60 * 1024 ? 74 * 1024 : (acc <= 100 * 1024 ? 112 * 1024 : 226 * 1024);
  group = 32;
  while (group > 1 && acc + (size_t)(block / 32) * group * maxrow * 8 > budget) group >>= 1;
  if (const char* e = getenv("KNP_ROWS_BLOCK")) block = atoi(e);
  if (const char* e = getenv("KNP_ROWS_GROUP")) group = atoi(e);
  stage_len = (group * maxrow + 1) & ~1;   // even: every warp's strip stays 16-byte aligned
  smem = (size_t)block * stride * 8 + (size_t)(block / 32) * stage_len * 8;
}

template <int D, int MODE>
static int launch_rows_t(const DevTopo& T, const KParams& P, const double* u, const double* fe, double* vals,
                         double* b, int stride, int max_deg, int max_gdeg, cudaStream_t st) {
  int block, group, stage_len;
  size_t smem;
  rows_config(max_deg, max_gdeg, MODE, block, group, stage_len, smem);
  if (smem > 227 * 1024) {
    set_error("vertex degree too large for the row kernel's shared-memory strip (%zu bytes)", smem);
    return KNP_E_UNSUPPORTED;
  }
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    KNP_CUDA(cudaFuncSetAttribute(rows_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int nb0 = (T.L.n_own[0] + block - 1) / block, nb1 = (T.L.n_own[1] + block - 1) / block;
  rows_kernel<D, MODE><<<nb0 + nb1, block, smem, st>>>(T, P, u, fe, vals, b, stride, group, stage_len, nb0);
  KNP_LAUNCHED();
  return KNP_OK;
}

int launch_rows(const DevTopo& T, const KParams& P, int mode, const double* u, const double* fe, double* vals,
                double* b, int max_deg, int max_gdeg, cudaStream_t st) {
  if (T.n_work == 0) return KNP_OK;
  const int stride = rows_smem_stride(max_deg, max_gdeg);
  if (T.gdim == 2) return mode == 0 ? launch_rows_t<2, 0>(T, P, u, fe, vals, b, stride, max_deg, max_gdeg, st)
                                    : launch_rows_t<2, 1>(T, P, u, fe, vals, b, stride, max_deg, max_gdeg, st);
  return mode == 0 ? launch_rows_t<3, 0>(T, P, u, fe, vals, b, stride, max_deg, max_gdeg, st)
                   : launch_rows_t<3, 1>(T, P, u, fe, vals, b, stride, max_deg, max_gdeg, st);
}

int launch_csr_indices(const DevTopo& T, int mode, int32_t* indices, cudaStream_t st) {
  if (T.n_work == 0) return KNP_OK;
  const int grid = (T.n_work + 127) / 128;
  if (mode == 0) csr_indices_kernel<0><<<grid, 128, 0, st>>>(T, indices);
  else csr_indices_kernel<1><<<grid, 128, 0, st>>>(T, indices);
  KNP_LAUNCHED();
  return KNP_OK;
}

int launch_l2_cells(int gdim, const Layout& L, int s, int field, int n_cells, const int32_t* cell_nodes,
                    const int32_t* cell_tag, const int32_t* cell_owned, const double* node_x, int nodeoff,
                    const int32_t* tags, int n_tags, const double* u, double* partial, int n_partial,
                    cudaStream_t st) {
  if (gdim == 2)
    l2_cells_kernel<2><<<n_partial, 256, 0, st>>>(L, s, field, n_cells, cell_nodes, cell_tag, cell_owned, node_x,
                                                  nodeoff, tags, n_tags, u, partial);
  else
    l2_cells_kernel<3><<<n_partial, 256, 0, st>>>(L, s, field, n_cells, cell_nodes, cell_tag, cell_owned, node_x,
                                                  nodeoff, tags, n_tags, u, partial);
  KNP_LAUNCHED();
  return KNP_OK;
}

}  // namespace knp
