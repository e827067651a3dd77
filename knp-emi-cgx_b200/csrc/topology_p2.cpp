// Host-side (setup-time) tables of the P2 element path: restricted dof sets over the nodes of the P2 space, node adjacency,
// node -> cell / node -> membrane-facet incidences with 2-byte slot maps, CSR patterns of A and P, quadrature and basis tables.
// Same role as topology.cpp for P1 (DofMapRestriction, KNPEMIx_problem.py:85-94, with ("Lagrange", fem_order) spaces, :38-42;
// create_matrix_block's pattern, KNPEMIx_solver.py:157; dS entity ordering, utils/mixed_dim_problem.py:708-729).  Also the
// host emulation of one assembly with the kernel bodies of p2.cuh (test infrastructure for the tier without a GPU).
#include <algorithm>
#include <cmath>
#include <numeric>
#include "p2.cuh"
#include "p2_quadrature.inc"

namespace knp {

// P2 Lagrange basis on a simplex with nv vertices at barycentric point lam: vertex functions lam_a (2 lam_a - 1), then
// 4 lam_i lam_j per edge (i < j, lexicographic); dN[a][m] = dN_a / dlam_m
static void p2_basis(int nv, const double* lam, double* N, double* dN) {
  const int nl = nv * (nv + 1) / 2;
  std::fill(dN, dN + (size_t)nl * nv, 0.0);
  for (int a = 0; a < nv; ++a) {
    N[a] = lam[a] * (2.0 * lam[a] - 1.0);
    dN[a * nv + a] = 4.0 * lam[a] - 1.0;
  }
  int e = nv;
  for (int i = 0; i < nv; ++i)
    for (int j = i + 1; j < nv; ++j, ++e) {
      N[e] = 4.0 * lam[i] * lam[j];
      dN[e * nv + i] = 4.0 * lam[j];
      dN[e * nv + j] = 4.0 * lam[i];
    }
}

int build_topology_p2(const knp_mesh_desc* m, HostTopo& T) {
  KNP_CHECK(m, "mesh descriptor is NULL");
  const int d = m->gdim;
  KNP_CHECK(d == 2 || d == 3, "gdim must be 2 or 3 (got %d)", d);
  const int nv = d + 1, NL = (d + 1) * (d + 2) / 2, NT = d * (d + 1) / 2;
  const int64_t NN = m->n_vertices, NC = m->n_cells, NF = m->n_mfacets;
  KNP_CHECK(NN > 0 && NC >= 0, "bad node/cell counts");
  KNP_CHECK(m->n_owned_vertices == NN, "P2 elements run on one GPU (every node must be owned)");
  KNP_CHECK(NN < (int64_t)1 << 31 && NC < (int64_t)1 << 31, "local mesh too large for int32 indices");
  KNP_CHECK(m->n_quad > 0 && m->n_quad <= 64 && m->quad_bary && m->quad_w, "facet quadrature rule missing (1..64 points)");
  T.gdim = d;
  T.degree = 2;
  P2Host& Q = T.p2;
  Q.nloc = NL;
  Q.nt = NT;

  std::vector<int32_t> itags(m->intra_tags, m->intra_tags + m->n_intra_tags);
  std::sort(itags.begin(), itags.end());
  KNP_CHECK(!std::binary_search(itags.begin(), itags.end(), m->extra_tag), "extra_tag is also listed as an intra tag");

  // ---- restricted node sets ----
  std::vector<int8_t> sub(NC);
  std::vector<uint8_t> mark[2];
  mark[0].assign(NN, 0);
  mark[1].assign(NN, 0);
  for (int64_t c = 0; c < NC; ++c) {
    const int t = m->cell_tags[c];
    const int s = (t == m->extra_tag) ? 1 : (std::binary_search(itags.begin(), itags.end(), t) ? 0 : -1);
    sub[c] = (int8_t)s;
    if (s < 0) continue;
    for (int a = 0; a < NL; ++a) {
      const int32_t v = m->cell_verts[c * NL + a];
      KNP_CHECK(v >= 0 && v < NN, "cell %lld references node %d out of range", (long long)c, v);
      mark[s][v] = 1;
    }
  }
  std::vector<int32_t> r[2];
  Layout& L = T.L;
  for (int s = 0; s < 2; ++s) {
    r[s].assign(NN, -1);
    T.node_vert[s].clear();
    for (int64_t v = 0; v < NN; ++v)
      if (mark[s][v]) {
        r[s][v] = (int32_t)T.node_vert[s].size();
        T.node_vert[s].push_back((int32_t)v);
      }
    L.n_loc[s] = L.n_own[s] = (int)T.node_vert[s].size();
    L.n_gh[s] = 0;
  }
  KNP_CHECK((int64_t)4 * (L.n_loc[0] + L.n_loc[1]) < ((int64_t)1 << 31), "too many unknowns for int32 columns");
  L.rowbase[0] = 0;
  L.rowbase[1] = 4 * L.n_own[0];
  L.n_rows = L.n_cols = 4 * (L.n_own[0] + L.n_own[1]);
  L.gbase[0] = L.gbase[1] = 0;
  const int nodeoff[2] = {0, L.n_loc[0]};
  const int workoff[2] = {0, L.n_own[0]};
  const int W = T.n_work = L.n_own[0] + L.n_own[1];
  T.node_x.resize((size_t)(L.n_loc[0] + L.n_loc[1]) * d);
  for (int s = 0; s < 2; ++s)
    for (int q = 0; q < L.n_loc[s]; ++q)
      for (int i = 0; i < d; ++i) T.node_x[(size_t)(nodeoff[s] + q) * d + i] = m->coords[(size_t)T.node_vert[s][q] * d + i];

  // ---- cells per subdomain ----
  for (int s = 0; s < 2; ++s) {
    T.cell_nodes[s].clear();
    T.cell_tag[s].clear();
    T.cell_owned[s].clear();
  }
  for (int64_t c = 0; c < NC; ++c) {
    const int s = sub[c];
    if (s < 0) continue;
    for (int a = 0; a < NL; ++a) T.cell_nodes[s].push_back(r[s][m->cell_verts[c * NL + a]]);
    T.cell_tag[s].push_back(m->cell_tags[c]);
    T.cell_owned[s].push_back(m->cell_owned ? m->cell_owned[c] : 1);
  }

  // ---- node -> cell incidence (cells ascending: fixed summation order) ----
  Q.inc_ptr.assign(W + 1, 0);
  for (int s = 0; s < 2; ++s)
    for (int32_t q : T.cell_nodes[s]) ++Q.inc_ptr[workoff[s] + q + 1];
  for (int w = 0; w < W; ++w) Q.inc_ptr[w + 1] += Q.inc_ptr[w];
  KNP_CHECK(Q.inc_ptr[W] >= 0, "incidence overflow");
  Q.inc_cell.resize(Q.inc_ptr[W]);
  Q.inc_loc.resize(Q.inc_ptr[W]);
  {
    std::vector<int32_t> fill(Q.inc_ptr.begin(), Q.inc_ptr.end() - 1);
    for (int s = 0; s < 2; ++s) {
      const auto& cn = T.cell_nodes[s];
      for (size_t c = 0; c < cn.size() / NL; ++c)
        for (int a = 0; a < NL; ++a) {
          const int k = fill[workoff[s] + cn[c * NL + a]]++;
          Q.inc_cell[k] = (int32_t)c;
          Q.inc_loc[k] = (uint8_t)a;
        }
    }
  }
  T.max_inc = 0;
  for (int w = 0; w < W; ++w) {
    KNP_CHECK(Q.inc_ptr[w + 1] > Q.inc_ptr[w], "dof %d has no incident cell", w);
    T.max_inc = std::max(T.max_inc, Q.inc_ptr[w + 1] - Q.inc_ptr[w]);
  }

  // ---- adjacency and cell slot maps ----
  Q.adj_ptr.assign(W + 1, 0);
  {
    std::vector<int32_t> deg(W);
#pragma omp parallel
    {
      std::vector<int32_t> tmp;
#pragma omp for schedule(static)
      for (int w = 0; w < W; ++w) {
        const auto& cn = T.cell_nodes[w >= workoff[1] ? 1 : 0];
        tmp.clear();
        for (int k = Q.inc_ptr[w]; k < Q.inc_ptr[w + 1]; ++k)
          for (int a = 0; a < NL; ++a) tmp.push_back(cn[(size_t)Q.inc_cell[k] * NL + a]);
        std::sort(tmp.begin(), tmp.end());
        deg[w] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
      }
    }
    int maxdeg = 0;
    for (int w = 0; w < W; ++w) {
      Q.adj_ptr[w + 1] = Q.adj_ptr[w] + deg[w];
      maxdeg = std::max(maxdeg, deg[w]);
    }
    KNP_CHECK(maxdeg < 65536, "node degree %d exceeds the 2-byte slot maps", maxdeg);
    T.max_deg = maxdeg;
  }
  Q.adj_idx.resize(Q.adj_ptr[W]);
  Q.inc_slots.resize((size_t)Q.inc_ptr[W] * NL);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(static)
    for (int w = 0; w < W; ++w) {
      const auto& cn = T.cell_nodes[w >= workoff[1] ? 1 : 0];
      tmp.clear();
      for (int k = Q.inc_ptr[w]; k < Q.inc_ptr[w + 1]; ++k)
        for (int a = 0; a < NL; ++a) tmp.push_back(cn[(size_t)Q.inc_cell[k] * NL + a]);
      std::sort(tmp.begin(), tmp.end());
      const int cnt = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
      int32_t* row = &Q.adj_idx[Q.adj_ptr[w]];
      std::copy(tmp.begin(), tmp.begin() + cnt, row);
      for (int k = Q.inc_ptr[w]; k < Q.inc_ptr[w + 1]; ++k)
        for (int a = 0; a < NL; ++a)
          Q.inc_slots[(size_t)k * NL + a] =
              (uint16_t)(std::lower_bound(row, row + cnt, cn[(size_t)Q.inc_cell[k] * NL + a]) - row);
    }
  }

  // ---- membrane nodes and facets ----
  T.n_mf = (int)NF;
  std::vector<int32_t> mvid(NN, -1);
  for (int64_t f = 0; f < NF; ++f)
    for (int a = 0; a < NT; ++a) {
      const int32_t v = m->mfacet_verts[f * NT + a];
      KNP_CHECK(v >= 0 && v < NN, "membrane facet %lld references node %d out of range", (long long)f, v);
      mvid[v] = 0;
    }
  T.mv_vert.clear();
  for (int64_t v = 0; v < NN; ++v)
    if (mvid[v] == 0) {
      mvid[v] = (int32_t)T.mv_vert.size();
      T.mv_vert.push_back((int32_t)v);
    }
  T.n_mv = (int)T.mv_vert.size();
  for (int s = 0; s < 2; ++s) {
    T.mv_node[s].resize(T.n_mv);
    for (int g = 0; g < T.n_mv; ++g) {
      T.mv_node[s][g] = r[s][T.mv_vert[g]];
      KNP_CHECK(T.mv_node[s][g] >= 0, "membrane node %d is not a node of an %s cell", T.mv_vert[g],
                s == 0 ? "intracellular" : "extracellular");
    }
  }
  T.mf_mv.resize((size_t)NF * NT);
  T.mtags.assign(m->mfacet_tags, m->mfacet_tags + NF);
  std::sort(T.mtags.begin(), T.mtags.end());
  T.mtags.erase(std::unique(T.mtags.begin(), T.mtags.end()), T.mtags.end());
  T.mf_tagidx.resize(NF);
  T.mf_owned.resize(NF);
  T.mf_area.resize(NF);
  for (int64_t f = 0; f < NF; ++f) {
    for (int a = 0; a < NT; ++a) T.mf_mv[f * NT + a] = mvid[m->mfacet_verts[f * NT + a]];
    const double* x0 = &m->coords[(size_t)m->mfacet_verts[f * NT + 0] * d];
    const double* x1 = &m->coords[(size_t)m->mfacet_verts[f * NT + 1] * d];
    if (d == 2) {
      T.mf_area[f] = std::sqrt((x1[0] - x0[0]) * (x1[0] - x0[0]) + (x1[1] - x0[1]) * (x1[1] - x0[1]));
    } else {
      const double* x2 = &m->coords[(size_t)m->mfacet_verts[f * NT + 2] * d];
      const double a0 = x1[0] - x0[0], a1 = x1[1] - x0[1], a2 = x1[2] - x0[2];
      const double b0 = x2[0] - x0[0], b1 = x2[1] - x0[1], b2 = x2[2] - x0[2];
      const double c0 = a1 * b2 - a2 * b1, c1 = a2 * b0 - a0 * b2, c2 = a0 * b1 - a1 * b0;
      T.mf_area[f] = 0.5 * std::sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    }
    T.mf_tagidx[f] = (int32_t)(std::lower_bound(T.mtags.begin(), T.mtags.end(), m->mfacet_tags[f]) - T.mtags.begin());
    T.mf_owned[f] = m->mfacet_owned ? m->mfacet_owned[f] : 1;
  }
  // node -> facet incidence (facets ascending), couplings across the membrane, facet slot maps
  Q.minc_ptr.assign(W + 1, 0);
  for (int64_t f = 0; f < NF; ++f)
    for (int a = 0; a < NT; ++a)
      for (int s = 0; s < 2; ++s) ++Q.minc_ptr[workoff[s] + r[s][m->mfacet_verts[f * NT + a]] + 1];
  for (int w = 0; w < W; ++w) Q.minc_ptr[w + 1] += Q.minc_ptr[w];
  Q.minc_facet.resize(Q.minc_ptr[W]);
  Q.minc_loc.resize(Q.minc_ptr[W]);
  {
    std::vector<int32_t> fill(Q.minc_ptr.begin(), Q.minc_ptr.end() - 1);
    for (int64_t f = 0; f < NF; ++f)
      for (int a = 0; a < NT; ++a)
        for (int s = 0; s < 2; ++s) {
          const int k = fill[workoff[s] + r[s][m->mfacet_verts[f * NT + a]]]++;
          Q.minc_facet[k] = (int32_t)f;
          Q.minc_loc[k] = (uint8_t)a;
        }
  }
  Q.gam_ptr.assign(W + 1, 0);
  Q.gam_idx.clear();
  Q.minc_own.resize((size_t)Q.minc_ptr[W] * NT);
  Q.minc_gam.resize((size_t)Q.minc_ptr[W] * NT);
  T.max_gdeg = 0;
  {
    std::vector<int32_t> tmp;
    for (int w = 0; w < W; ++w) {
      const int s = w >= workoff[1] ? 1 : 0, o = 1 - s;
      tmp.clear();
      for (int k = Q.minc_ptr[w]; k < Q.minc_ptr[w + 1]; ++k)
        for (int b = 0; b < NT; ++b) tmp.push_back(r[o][m->mfacet_verts[(size_t)Q.minc_facet[k] * NT + b]]);
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      const int32_t* adj = &Q.adj_idx[Q.adj_ptr[w]];
      const int deg = Q.adj_ptr[w + 1] - Q.adj_ptr[w];
      for (int k = Q.minc_ptr[w]; k < Q.minc_ptr[w + 1]; ++k)
        for (int b = 0; b < NT; ++b) {
          const int32_t v = m->mfacet_verts[(size_t)Q.minc_facet[k] * NT + b];
          const int32_t* it = std::lower_bound(adj, adj + deg, r[s][v]);
          KNP_CHECK(it != adj + deg && *it == r[s][v], "membrane facet %d is not a face of an %s cell", Q.minc_facet[k],
                    s == 0 ? "intracellular" : "extracellular");
          Q.minc_own[(size_t)k * NT + b] = (uint16_t)(it - adj);
          Q.minc_gam[(size_t)k * NT + b] = (uint16_t)(std::lower_bound(tmp.begin(), tmp.end(), r[o][v]) - tmp.begin());
        }
      Q.gam_idx.insert(Q.gam_idx.end(), tmp.begin(), tmp.end());
      Q.gam_ptr[w + 1] = (int32_t)Q.gam_idx.size();
      T.max_gdeg = std::max(T.max_gdeg, (int)tmp.size());
    }
  }

  // ---- CSR patterns: rows as described at p2_row_body ----
  T.indptr.assign(L.n_rows + 1, 0);
  T.indptr_P.assign(L.n_rows + 1, 0);
  int64_t nnz = 0, nnzP = 0;
  for (int s = 0; s < 2; ++s)
    for (int f = 0; f < 4; ++f)
      for (int p = 0; p < L.n_own[s]; ++p) {
        const int w = workoff[s] + p;
        const int deg = Q.adj_ptr[w + 1] - Q.adj_ptr[w], gdeg = Q.gam_ptr[w + 1] - Q.gam_ptr[w];
        nnz += (f < 3 ? 2 : 4) * (int64_t)deg + gdeg;
        nnzP += deg;
        KNP_CHECK(nnz < ((int64_t)1 << 31), "matrix has more than 2^31 non-zeros");
        T.indptr[L.row(s, f, p) + 1] = (int32_t)nnz;
        T.indptr_P[L.row(s, f, p) + 1] = (int32_t)nnzP;
      }
  T.nnz = nnz;
  T.nnz_P = nnzP;
  Q.indices.resize(nnz);
  Q.indices_P.resize(nnzP);
#pragma omp parallel for schedule(static)
  for (int w = 0; w < W; ++w) {
    const int s = w >= workoff[1] ? 1 : 0, p = w - workoff[s], o = 1 - s;
    const int32_t* adj = &Q.adj_idx[Q.adj_ptr[w]];
    const int deg = Q.adj_ptr[w + 1] - Q.adj_ptr[w];
    const int32_t* gam = Q.gam_idx.data() + Q.gam_ptr[w];
    const int gdeg = Q.gam_ptr[w + 1] - Q.gam_ptr[w];
    for (int f = 0; f < 4; ++f) {
      int pos = T.indptr[L.row(s, f, p)];
      if (s == 1)
        for (int e = 0; e < gdeg; ++e) Q.indices[pos++] = L.col(o, 3, gam[e]);
      for (int k = (f < 3 ? f : 0); k < (f < 3 ? f + 1 : 3); ++k)
        for (int e = 0; e < deg; ++e) Q.indices[pos++] = L.col(s, k, adj[e]);
      for (int e = 0; e < deg; ++e) Q.indices[pos++] = L.col(s, 3, adj[e]);
      if (s == 0)
        for (int e = 0; e < gdeg; ++e) Q.indices[pos++] = L.col(o, 3, gam[e]);
      int posP = T.indptr_P[L.row(s, f, p)];
      for (int e = 0; e < deg; ++e) Q.indices_P[posP++] = L.col(s, f, adj[e]);
    }
  }

  // ---- quadrature and basis tables ----
  const int nqc = d == 2 ? 9 : 27;
  Q.nqc = nqc;
  Q.cq_w.resize(nqc);
  Q.cq_N.resize((size_t)nqc * NL);
  Q.cq_dN.resize((size_t)nqc * NL * nv);
  for (int q = 0; q < nqc; ++q) {
    const double* row = d == 2 ? P2_RULE_2D[q] : P2_RULE_3D[q];
    Q.cq_w[q] = row[nv];
    p2_basis(nv, row, &Q.cq_N[(size_t)q * NL], &Q.cq_dN[(size_t)q * NL * nv]);
  }
  {
    double tr = 0.0;
    for (int q = 0; q < nqc; ++q)
      for (int a = 0; a < NL; ++a) tr += Q.cq_w[q] * Q.cq_N[(size_t)q * NL + a] * Q.cq_N[(size_t)q * NL + a];
    Q.hrz = 1.0 / tr;
  }
  const int nqf = m->n_quad;
  Q.fq_b.assign(m->quad_bary, m->quad_bary + (size_t)nqf * d);
  Q.fq_w.assign(m->quad_w, m->quad_w + nqf);
  Q.fq_N.resize((size_t)nqf * NT);
  Q.fq_M.assign((size_t)NT * NT, 0.0);
  {
    std::vector<double> dN((size_t)NT * d);
    for (int q = 0; q < nqf; ++q) {
      double* N = &Q.fq_N[(size_t)q * NT];
      p2_basis(d, &Q.fq_b[(size_t)q * d], N, dN.data());
      for (int a = 0; a < NT; ++a)
        for (int b = 0; b < NT; ++b) Q.fq_M[a * NT + b] += Q.fq_w[q] * N[a] * N[b];
    }
  }
  return KNP_OK;
}

P2View p2_host_view(const HostTopo& H) {
  const P2Host& Q = H.p2;
  P2View V{};
  V.gdim = H.gdim;
  V.nloc = Q.nloc;
  V.nt = Q.nt;
  V.nqc = Q.nqc;
  V.nqf = (int)Q.fq_w.size();
  V.L = H.L;
  V.n_work = H.n_work;
  V.n_mf = H.n_mf;
  V.n_mv = H.n_mv;
  V.node_x = H.node_x.data();
  V.cell_nodes[0] = H.cell_nodes[0].data();
  V.cell_nodes[1] = H.cell_nodes[1].data();
  V.adj_ptr = Q.adj_ptr.data();
  V.gam_ptr = Q.gam_ptr.data();
  V.inc_ptr = Q.inc_ptr.data();
  V.inc_cell = Q.inc_cell.data();
  V.inc_loc = Q.inc_loc.data();
  V.inc_slots = Q.inc_slots.data();
  V.minc_ptr = Q.minc_ptr.data();
  V.minc_facet = Q.minc_facet.data();
  V.minc_loc = Q.minc_loc.data();
  V.minc_own = Q.minc_own.data();
  V.minc_gam = Q.minc_gam.data();
  V.indptr = H.indptr.data();
  V.indptr_P = H.indptr_P.data();
  V.cq_w = Q.cq_w.data();
  V.cq_N = Q.cq_N.data();
  V.cq_dN = Q.cq_dN.data();
  V.fq_b = Q.fq_b.data();
  V.fq_w = Q.fq_w.data();
  V.fq_N = Q.fq_N.data();
  V.fq_M = Q.fq_M.data();
  V.mv_node0 = H.mv_node[0].data();
  V.mv_node1 = H.mv_node[1].data();
  V.mf_mv = H.mf_mv.data();
  V.mf_tagidx = H.mf_tagidx.data();
  V.mf_area = H.mf_area.data();
  return V;
}

// One assembly with the kernel bodies, on the CPU (TEST INFRASTRUCTURE: reached only through knp_p2_emulate_host)
template <int D>
static void emulate_t(const P2View& V, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim, double stim_fac,
                      int mode, const double* u, const double* gates, double* vals, double* b) {
  std::vector<double> fe((size_t)p2_facet_ncomp(D) * std::max(V.n_mf, 1));
  if (mode == 0)
    for (int f = 0; f < V.n_mf; ++f) p2_facet_body<D>(V, P, tag_models, tag_stim, u, gates, stim_fac, fe.data(), f);
  const P2Coef C = p2_coef(P);
  for (int w = 0; w < V.n_work; ++w) {
    if (mode == 0) p2_row_body<D, 0>(V, C, u, fe.data(), vals, b, w);
    else p2_row_body<D, 1>(V, C, u, fe.data(), vals, b, w);
  }
}

int p2_emulate_host(const HostTopo& H, const KParams& P, const uint32_t* tag_models, const int32_t* tag_stim, double stim_fac,
                    int mode, const double* u, const double* gates, double* vals, double* b) {
  const P2View V = p2_host_view(H);
  if (H.gdim == 2) emulate_t<2>(V, P, tag_models, tag_stim, stim_fac, mode, u, gates, vals, b);
  else emulate_t<3>(V, P, tag_models, tag_stim, stim_fac, mode, u, gates, vals, b);
  return KNP_OK;
}

}  // namespace knp
