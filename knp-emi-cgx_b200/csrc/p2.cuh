// P2 (fem_order = 2, utils/mixed_dim_problem.py:207-208, KNPEMIx_problem.py:38-42) element path: tables and the bodies of
// its kernels.  The forms are those of the P1 path (KNPEMIx_problem.py:594-642, preconditioner :717-738); what changes is the
// element: 6 / 10 dofs per cell (vertices + edge nodes), 3 / 6 per membrane facet, c_k-weighted stiffness integrands of degree
// 4 (no closed form in cell means), so the row kernel integrates with a degree-5 collapsed Gauss-Jacobi rule.
//
// Ownership is the P1 path's: one thread owns a restricted dof ("node") and with it the four matrix rows of that node (three
// ions and the potential), walks the node's incident cells and membrane facets in ascending order and adds each element row
// at precomputed slots (2 bytes per (node, cell, local dof): the position in the node's sorted adjacency).  Every entry is
// touched by one thread in a fixed order: no atomics, bitwise reproducible.  This is the plain owner-computes form, not a
// tuned one -- no shipped configuration of the reference uses order 2 and BASELINE.json benchmarks none.  Measured on B200
// (profiles/r02p_p2.md): 2.65 ms for 4.27 M unknowns in 2D, 4.05 ms for 1.20 M unknowns in 3D; the rows of the resident
// threads exceed the L2, so the kernel is bound by the latency of the read-modify-write of its own rows.
//
// The bodies are __host__ __device__: the kernels in assembly_p2.cu call them per thread, and knp_p2_emulate_host (api.cu)
// calls the same functions in a loop on the CPU so that the test tier without a GPU checks the tables and the element math
// against the oracle.  The emulation is test infrastructure: no product call reaches it.
#pragma once
#include <cmath>
#include "common.cuh"
#include "kernels.cuh"

#define KNP_HD __host__ __device__ __forceinline__

namespace knp {

// Device (or, in the emulation, host) views of the P2 tables
struct P2View {
  int gdim, nloc, nt, nqc, nqf;
  Layout L;
  int n_work, n_mf, n_mv;
  const double* node_x;                 // [n_loc0 + n_loc1][gdim], intracellular local nodes first
  const int32_t* cell_nodes[2];         // [nloc per cell] subdomain-local node ids, vertices first
  const int32_t *adj_ptr, *gam_ptr;     // per owned node: adjacency / membrane-coupling degree prefix
  const int32_t *inc_ptr, *inc_cell;
  const uint8_t* inc_loc;
  const uint16_t* inc_slots;
  const int32_t *minc_ptr, *minc_facet;
  const uint8_t* minc_loc;
  const uint16_t *minc_own, *minc_gam;
  const int32_t *indptr, *indptr_P;
  const double *cq_w, *cq_N, *cq_dN;    // cell rule: weights (sum 1), basis values [q][nloc], d/d lambda_m [q][nloc][gdim + 1]
  const double *fq_b, *fq_w, *fq_N;     // facet rule: barycentrics [q][gdim], weights (sum 1), trace basis values [q][nt]
  const double* fq_M;                   // reference facet mass matrix [nt][nt] (int N_a N_b / |F|)
  const int32_t *mv_node0, *mv_node1;   // membrane node -> local node id inside / outside
  const int32_t *mf_mv, *mf_tagidx;     // [nt per facet] membrane node ids; tag index
  const double* mf_area;
};

struct P2Coef {         // constants of the forms, folded on the host (KNPEMIx_problem.py:598-610,633-642)
  double dtD[3];        // dt D_k
  double cphi[3];       // dt D_k z_k / psi
  double cpp[3];        // dt D_k z_k^2 / psi
  double ck[3];         // dt z_k D_k
  double cmz[3];        // C_M / (F z_k)
  double cf;            // C_M / F
};

inline P2Coef p2_coef(const KParams& P) {
  P2Coef C;
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

KNP_HD int p2_sym(int a, int b, int n) {        // upper triangle, row-major, any order of (a, b)
  const int i = a < b ? a : b, j = a < b ? b : a;
  return i * n - (i * (i - 1)) / 2 + (j - i);
}
inline int p2_facet_ncomp(int gdim) {
  const int nt = gdim * (gdim + 1) / 2;
  return 6 * (nt * (nt + 1) / 2) + 7 * nt;
}

// volume and gradients of the barycentric coordinates of a simplex
KNP_HD void p2_geometry(const double (&x)[3][2], double& vol, double (&g)[3][2]) {
  const double e1x = x[1][0] - x[0][0], e1y = x[1][1] - x[0][1];
  const double e2x = x[2][0] - x[0][0], e2y = x[2][1] - x[0][1];
  const double det = e1x * e2y - e1y * e2x;
  const double inv = 1.0 / det;
  vol = 0.5 * fabs(det);
  g[1][0] = e2y * inv;
  g[1][1] = -e2x * inv;
  g[2][0] = -e1y * inv;
  g[2][1] = e1x * inv;
  g[0][0] = -(g[1][0] + g[2][0]);
  g[0][1] = -(g[1][1] + g[2][1]);
}
KNP_HD void p2_geometry(const double (&x)[4][3], double& vol, double (&g)[4][3]) {
  double e[3][3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) e[j][i] = x[j + 1][i] - x[0][i];
  const double c23[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2],
                         e[1][0] * e[2][1] - e[1][1] * e[2][0]};
  const double c31[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2],
                         e[2][0] * e[0][1] - e[2][1] * e[0][0]};
  const double c12[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2],
                         e[0][0] * e[1][1] - e[0][1] * e[1][0]};
  const double det = e[0][0] * c23[0] + e[0][1] * c23[1] + e[0][2] * c23[2];
  const double inv = 1.0 / det;
  vol = fabs(det) / 6.0;
  for (int i = 0; i < 3; ++i) {
    g[1][i] = c23[i] * inv;
    g[2][i] = c31[i] * inv;
    g[3][i] = c12[i] * inv;
    g[0][i] = -(g[1][i] + g[2][i] + g[3][i]);
  }
}

// Channel currents I_k of the IonicModel._eval family (KNPEMIx_ionic_model.py; the same expressions, in the same order, as
// facet_kernel of assembly.cu) at one quadrature point; stim = mask * g_syn_bar exp(-t/a_syn) [/ stimulus area], 0 when
// the facet's tag is not stimulated.
KNP_HD void p2_channel_currents(const KParams& P, uint32_t models, double stim, const double (&ciq)[3],
                                const double (&ceq)[3], double pmq, double nq, double mq, double hq, double (&I)[3]) {
  const double psi = P.psi;
  double E[3], lg[3], ici[3];
  for (int k = 0; k < 3; ++k) {
    ici[k] = 1.0 / ciq[k];
    lg[k] = log(ceq[k] * ici[k]);
    E[k] = (psi / P.z[k]) * lg[k];                       // KNPEMIx_problem.py:516
  }
  I[0] = I[1] = I[2] = 0.0;
  if (models & KNP_MODEL_NEURONAL_CT) {                  // :342-369 (f_NKCC1 == 0, :50-75)
    const double I_KCC2 = -0.0068 * (lg[1] + lg[2]);
    I[1] += I_KCC2;
    I[2] += -I_KCC2;
  }
  if (models & KNP_MODEL_HH) {                           // :487-515 (+ stimulus :517-603)
    const double gNa = P.g_leak[0] + P.g_Na_bar * mq * mq * mq * hq;
    const double gK = P.g_leak[1] + P.g_K_bar * (nq * nq) * (nq * nq);
    I[0] += gNa * (pmq - E[0]) + stim * (pmq - E[0]);
    I[1] += gK * (pmq - E[1]);
    I[2] += P.g_leak[2] * (pmq - E[2]);
  }
  if (models & KNP_MODEL_ATP) {                          // :385-422
    const double p1 = 1.0 + 1.5 / ceq[1];
    const double p2 = 1.0 + 10.0 * ici[0];
    const double I_ATP = 0.25 / ((p1 * p1) * (p2 * p2 * p2));
    I[0] += 3.0 * I_ATP;
    I[1] += -2.0 * I_ATP;
  }
  if (models & KNP_MODEL_GLIAL_CT) {                     // :239-298 (f_NKCC1 == 0)
    const double I_KCC1 = -(7e-2 * psi) * (lg[1] + lg[2]);
    I[1] += I_KCC1;
    I[2] += -I_KCC1;
  }
  if (models & KNP_MODEL_KIRNA) {                        // :117-222
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
  if (models & KNP_MODEL_PASSIVE) {                      // :89-91
    I[0] += pmq;
    I[1] += pmq;
    I[2] += pmq;
  }
}

// stimulus mask prod_i [lo_i < x_{dir_i} < hi_i] at a facet quadrature point (KNPEMIx_ionic_model.py:558-587); the facet is
// flat, so the point follows from its D vertices
template <int D>
KNP_HD double p2_stim_mask(const P2View& V, const KParams& P, const int (&nodei)[D], int q) {
  double mask = 1.0;
  for (int i = 0; i < 3 && P.stim_dir[i] >= 0; ++i) {
    double xq = 0.0;
    for (int a = 0; a < D; ++a) xq += V.fq_b[q * D + a] * V.node_x[(size_t)nodei[a] * D + P.stim_dir[i]];
    mask *= (xq > P.stim_lo[i] && xq < P.stim_hi[i]) ? 1.0 : 0.0;
  }
  return mask;
}

// Membrane-facet element tensors (dS terms of KNPEMIx_problem.py:594-642) of facet f into the facet-major staging buffer:
//   GA  : ((s*3+k)*NSF + ab)           6*NSF     NSF = NT (NT + 1) / 2
//   bc  : 6*NSF + (s*3+k)*NT + a       6*NT      already divided by F z_k
//   bphi: 6*NSF + 6*NT + a             NT        already divided by F
template <int D>
KNP_HD void p2_facet_body(const P2View& V, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim,
                          const double* u, const double* gates, double stim_fac, double* fe, int f) {
  constexpr int NT = D * (D + 1) / 2, NSF = NT * (NT + 1) / 2;
  double ci[3][NT], ce[3][NT], pm[NT], gn[NT], gm[NT], gh[NT];
  int nodei[D];
  for (int a = 0; a < NT; ++a) {
    const int g = V.mf_mv[(size_t)f * NT + a];
    const int qi = V.mv_node0[g], qe = V.mv_node1[g];
    for (int k = 0; k < 3; ++k) {
      ci[k][a] = u[V.L.col(0, k, qi)];
      ce[k][a] = u[V.L.col(1, k, qe)];
    }
    pm[a] = u[V.L.col(0, 3, qi)] - u[V.L.col(1, 3, qe)];
    gn[a] = gates[g];
    gm[a] = gates[(size_t)V.n_mv + g];
    gh[a] = gates[(size_t)2 * V.n_mv + g];
    if (a < D) nodei[a] = qi;
  }
  const double area = V.mf_area[f];
  const int ti = V.mf_tagidx[f];
  const uint32_t models = tag_models[ti];
  const bool stim_on = tag_stim[ti] != 0;
  double GA[2][3][NSF], bc[2][3][NT], bphi[NT];
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 3; ++k) {
      for (int i = 0; i < NSF; ++i) GA[s][k][i] = 0.0;
      for (int a = 0; a < NT; ++a) bc[s][k][a] = 0.0;
    }
  for (int a = 0; a < NT; ++a) bphi[a] = 0.0;
  for (int q = 0; q < V.nqf; ++q) {
    const double* N = V.fq_N + (size_t)q * NT;
    const double w = area * V.fq_w[q];
    double ciq[3] = {0.0, 0.0, 0.0}, ceq[3] = {0.0, 0.0, 0.0}, pmq = 0.0, nq = 0.0, mq = 0.0, hq = 0.0;
    for (int a = 0; a < NT; ++a) {
      for (int k = 0; k < 3; ++k) {
        ciq[k] += N[a] * ci[k][a];
        ceq[k] += N[a] * ce[k][a];
      }
      pmq += N[a] * pm[a];
      nq += N[a] * gn[a];
      mq += N[a] * gm[a];
      hq += N[a] * gh[a];
    }
    double al[2][3];                                     // alpha_{k,s} (KNPEMIx_problem.py:512-513,582-583)
    {
      double di = 0.0, de = 0.0;
      for (int k = 0; k < 3; ++k) {
        di += P.D[k] * P.z[k] * P.z[k] * ciq[k];
        de += P.D[k] * P.z[k] * P.z[k] * ceq[k];
      }
      const double idi = 1.0 / di, ide = 1.0 / de;
      for (int k = 0; k < 3; ++k) {
        al[0][k] = (P.D[k] * P.z[k] * P.z[k] * ciq[k]) * idi;
        al[1][k] = (P.D[k] * P.z[k] * P.z[k] * ceq[k]) * ide;
      }
    }
    const double stim = (stim_on && (models & KNP_MODEL_HH)) ? p2_stim_mask<D>(V, P, nodei, q) * stim_fac : 0.0;
    double I[3];
    p2_channel_currents(P, models, stim, ciq, ceq, pmq, nq, mq, hq, I);
    const double Itot = (I[0] + I[1]) + I[2];
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < 3; ++k) {
        const double wa = w * al[s][k];
        const double rb = w * (P.dt * I[k] - al[s][k] * P.C_M * pmq);
        for (int a = 0; a < NT; ++a) {
          bc[s][k][a] += rb * N[a];
          for (int b = a; b < NT; ++b) GA[s][k][a * NT - (a * (a - 1)) / 2 + (b - a)] += wa * (N[a] * N[b]);
        }
      }
    const double rp = w * (P.dt * Itot - P.C_M * pmq);
    for (int a = 0; a < NT; ++a) bphi[a] += rp * N[a];
  }
  const size_t nf = (size_t)V.n_mf;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 3; ++k) {
      for (int i = 0; i < NSF; ++i) fe[(size_t)((s * 3 + k) * NSF + i) * nf + f] = GA[s][k][i];
      const double inv = 1.0 / (P.F * P.z[k]);
      for (int a = 0; a < NT; ++a) fe[(size_t)(6 * NSF + (s * 3 + k) * NT + a) * nf + f] = bc[s][k][a] * inv;
    }
  for (int a = 0; a < NT; ++a) fe[(size_t)(6 * NSF + 6 * NT + a) * nf + f] = bphi[a] / P.F;
}

// The four matrix rows (and right-hand side entries) of owned node w.
// MODE 0: A and b (KNPEMIx_problem.py:598-642).  MODE 1: block-Jacobi preconditioner matrix P (:717-738).
// Row layout (columns ascending), deg = adjacency size, gdeg = nodes across the membrane:
//   MODE 0, ion row k   : [s = 1: phi_i gamma (gdeg)] [c_k (deg)] [phi_s (deg)] [s = 0: phi_e gamma (gdeg)]
//   MODE 0, potential   : [s = 1: phi_i gamma (gdeg)] [c_0 (deg)] [c_1 (deg)] [c_2 (deg)] [phi_s (deg)] [s = 0: phi_e gamma (gdeg)]
//   MODE 1, every row   : [own field (deg)]
template <int D, int MODE>
KNP_HD void p2_row_body(const P2View& V, const P2Coef& C, const double* u, const double* fe, double* vals, double* bvec,
                        int w) {
  constexpr int NV = D + 1, NL = (D + 1) * (D + 2) / 2, NT = D * (D + 1) / 2, NSF = NT * (NT + 1) / 2;
  const Layout& L = V.L;
  const int s = w >= L.n_own[0] ? 1 : 0;
  const int p = w - (s ? L.n_own[0] : 0);
  const int nodeoff = s ? L.n_loc[0] : 0;
  const int deg = V.adj_ptr[w + 1] - V.adj_ptr[w];
  const int gdeg = MODE == 0 ? V.gam_ptr[w + 1] - V.gam_ptr[w] : 0;
  const int32_t* iptr = MODE == 0 ? V.indptr : V.indptr_P;
  int rs[4];
  for (int f = 0; f < 4; ++f) {
    const int row = L.row(s, f, p);
    rs[f] = iptr[row];
    for (int j = rs[f]; j < iptr[row + 1]; ++j) vals[j] = 0.0;
  }
  const int goff = (MODE == 0 && s == 1) ? gdeg : 0;              // start of the own-subdomain segments
  const int gam_ion = s == 0 ? 2 * deg : 0, gam_phi = s == 0 ? 4 * deg : 0;
  double bk[3] = {0.0, 0.0, 0.0}, bp = 0.0;

  const int32_t* cells = V.cell_nodes[s];
  for (int ii = V.inc_ptr[w]; ii < V.inc_ptr[w + 1]; ++ii) {
    const int32_t* nodes = cells + (size_t)V.inc_cell[ii] * NL;
    const int a = V.inc_loc[ii];
    const uint16_t* slot = V.inc_slots + (size_t)ii * NL;
    double x[NV][D], g[NV][D], vol;
    for (int v = 0; v < NV; ++v)
      for (int i = 0; i < D; ++i) x[v][i] = V.node_x[(size_t)(nodeoff + nodes[v]) * D + i];
    p2_geometry(x, vol, g);
    double ck[3][NL];
    for (int b = 0; b < NL; ++b)
      for (int k = 0; k < 3; ++k) ck[k][b] = u[L.col(s, k, nodes[b])];
    double rM[NL], rK[NL], rW[3][NL];
    for (int b = 0; b < NL; ++b) {
      rM[b] = 0.0;
      rK[b] = 0.0;
      rW[0][b] = rW[1][b] = rW[2][b] = 0.0;
    }
    for (int q = 0; q < V.nqc; ++q) {
      const double* N = V.cq_N + (size_t)q * NL;
      const double* dN = V.cq_dN + (size_t)q * NL * NV;
      const double wq = vol * V.cq_w[q];
      double ga[D], cq[3] = {0.0, 0.0, 0.0};
      for (int i = 0; i < D; ++i) {
        double t = 0.0;
        for (int m = 0; m < NV; ++m) t += dN[a * NV + m] * g[m][i];
        ga[i] = t;
      }
      for (int b = 0; b < NL; ++b)
        for (int k = 0; k < 3; ++k) cq[k] += N[b] * ck[k][b];
      const double wNa = wq * N[a];
      for (int b = 0; b < NL; ++b) {
        double dot = 0.0;
        for (int i = 0; i < D; ++i) {
          double t = 0.0;
          for (int m = 0; m < NV; ++m) t += dN[b * NV + m] * g[m][i];
          dot += ga[i] * t;
        }
        const double wd = wq * dot;
        rK[b] += wd;
        for (int k = 0; k < 3; ++k) rW[k][b] += wd * cq[k];
        rM[b] += wNa * N[b];
      }
    }
    for (int b = 0; b < NL; ++b) {
      // the ten entries of column slot[b] are read together before the first one is written back: the loads are independent,
      // a load after a store to the same array would have to wait for it
      const int sl = slot[b];
      double* const vcc[3] = {vals + rs[0] + goff + sl, vals + rs[1] + goff + sl, vals + rs[2] + goff + sl};
      double* const vpp = vals + rs[3] + goff + (MODE == 0 ? 3 * deg : 0) + sl;
      double occ[3], ocp[3] = {0.0, 0.0, 0.0}, opc[3] = {0.0, 0.0, 0.0};
      for (int k = 0; k < 3; ++k) {
        occ[k] = *vcc[k];
        if (MODE == 0) {
          ocp[k] = vcc[k][deg];
          opc[k] = vals[rs[3] + goff + k * deg + sl];
        }
      }
      const double opp = *vpp;
      double pp = 0.0;
      for (int k = 0; k < 3; ++k) {
        *vcc[k] = occ[k] + (rM[b] + C.dtD[k] * rK[b]);
        pp += C.cpp[k] * rW[k][b];
        if (MODE == 0) {
          vcc[k][deg] = ocp[k] + C.cphi[k] * rW[k][b];
          vals[rs[3] + goff + k * deg + sl] = opc[k] + C.ck[k] * rK[b];
          bk[k] += rM[b] * ck[k][b];
        }
      }
      *vpp = opp + pp;
    }
  }
  // membrane (dS) terms: KNPEMIx_problem.py:599,604,609-610,637-638,641-642 (P: :737-738)
  const double sgn = s == 0 ? 1.0 : -1.0;
  const size_t nf = (size_t)V.n_mf;
  for (int mi = V.minc_ptr[w]; mi < V.minc_ptr[w + 1]; ++mi) {
    const int f = V.minc_facet[mi];
    const int a = V.minc_loc[mi];
    const uint16_t* so = V.minc_own + (size_t)mi * NT;
    const uint16_t* sg = V.minc_gam + (size_t)mi * NT;
    const double area = V.mf_area[f];
    for (int b = 0; b < NT; ++b) {
      const double G1 = C.cf * (area * V.fq_M[a * NT + b]);
      if (MODE == 0) {
        const int ab = p2_sym(a, b, NT);
        double ga[3], oo[3], og[3];
        for (int k = 0; k < 3; ++k) {          // reads first, then the writes (see above)
          ga[k] = C.cmz[k] * fe[(size_t)((s * 3 + k) * NSF + ab) * nf + f];
          oo[k] = vals[rs[k] + goff + deg + so[b]];
          og[k] = vals[rs[k] + gam_ion + sg[b]];
        }
        const double po = vals[rs[3] + goff + 3 * deg + so[b]], pg = vals[rs[3] + gam_phi + sg[b]];
        for (int k = 0; k < 3; ++k) {
          vals[rs[k] + goff + deg + so[b]] = oo[k] + ga[k];
          vals[rs[k] + gam_ion + sg[b]] = og[k] - ga[k];
        }
        vals[rs[3] + goff + 3 * deg + so[b]] = po + G1;
        vals[rs[3] + gam_phi + sg[b]] = pg - G1;
      } else {
        vals[rs[3] + so[b]] -= G1;
      }
    }
    if (MODE == 0) {
      for (int k = 0; k < 3; ++k) bk[k] -= sgn * fe[(size_t)(6 * NSF + (s * 3 + k) * NT + a) * nf + f];
      bp -= sgn * fe[(size_t)(6 * NSF + 6 * NT + a) * nf + f];
    }
  }
  if (MODE == 0) {
    for (int k = 0; k < 3; ++k) bvec[L.row(s, k, p)] = bk[k];
    bvec[L.row(s, 3, p)] = bp;
  }
}

// int u^power over one cell of subdomain s (power 0: the measure), P2 field `field`
template <int D>
KNP_HD double p2_cell_integral(const P2View& V, int s, int field, int power, const double* u, int c) {
  constexpr int NV = D + 1, NL = (D + 1) * (D + 2) / 2;
  const int32_t* nodes = V.cell_nodes[s] + (size_t)c * NL;
  const int nodeoff = s ? V.L.n_loc[0] : 0;
  double x[NV][D], g[NV][D], vol;
  for (int v = 0; v < NV; ++v)
    for (int i = 0; i < D; ++i) x[v][i] = V.node_x[(size_t)(nodeoff + nodes[v]) * D + i];
  p2_geometry(x, vol, g);
  if (power == 0) return vol;
  double uc[NL];
  for (int b = 0; b < NL; ++b) uc[b] = u[V.L.col(s, field, nodes[b])];
  double acc = 0.0;
  for (int q = 0; q < V.nqc; ++q) {
    const double* N = V.cq_N + (size_t)q * NL;
    double uq = 0.0;
    for (int b = 0; b < NL; ++b) uq += N[b] * uc[b];
    acc += V.cq_w[q] * (power == 2 ? uq * uq : uq);
  }
  return vol * acc;
}

// stimulus current density integrated over facet f (KNPEMIx_solver.py:578-610, KNPEMIx_ionic_model.py:517-603)
template <int D>
KNP_HD double p2_facet_stim_current(const P2View& V, const KParams& P, const double* u, double stim_fac, int f) {
  constexpr int NT = D * (D + 1) / 2;
  double ci[NT], ce[NT], pm[NT];
  int nodei[D];
  for (int a = 0; a < NT; ++a) {
    const int g = V.mf_mv[(size_t)f * NT + a];
    const int qi = V.mv_node0[g], qe = V.mv_node1[g];
    ci[a] = u[V.L.col(0, 0, qi)];
    ce[a] = u[V.L.col(1, 0, qe)];
    pm[a] = u[V.L.col(0, 3, qi)] - u[V.L.col(1, 3, qe)];
    if (a < D) nodei[a] = qi;
  }
  double acc = 0.0;
  for (int q = 0; q < V.nqf; ++q) {
    const double* N = V.fq_N + (size_t)q * NT;
    double ciq = 0.0, ceq = 0.0, pmq = 0.0;
    for (int a = 0; a < NT; ++a) {
      ciq += N[a] * ci[a];
      ceq += N[a] * ce[a];
      pmq += N[a] * pm[a];
    }
    const double E_Na = (P.psi / P.z[0]) * log(ceq / ciq);
    acc += V.mf_area[f] * V.fq_w[q] * p2_stim_mask<D>(V, P, nodei, q) * stim_fac * (pmq - E_Na);
  }
  return acc;
}

// assembly_p2.cu
int launch_facets_p2(const P2View& V, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim, const double* u,
                     const double* gates, double stim_fac, double* fe, cudaStream_t st);
int launch_rows_p2(const P2View& V, const KParams& P, int mode, const double* u, const double* fe, double* vals, double* b,
                   cudaStream_t st);
int launch_l2_cells_p2(const P2View& V, int s, int field, int power, int n_cells, const int32_t* cell_tag,
                       const int32_t* cell_owned, const int32_t* tags, int n_tags, const double* u, double* partial,
                       int n_partial, cudaStream_t st);
int launch_stim_current_p2(const P2View& V, const KParams& P, const int32_t* tag_stim, const int32_t* mf_owned, const double* u,
                           double stim_fac, double* partial, int n_partial, cudaStream_t st);
// host emulation of one assembly with the bodies above (test infrastructure; topology_p2.cpp)
int p2_emulate_host(const HostTopo& H, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim, double stim_fac,
                    int mode, const double* u, const double* gates, double* vals, double* b);
P2View p2_host_view(const HostTopo& H);
int build_topology_p2(const knp_mesh_desc* m, HostTopo& T);

}  // namespace knp
