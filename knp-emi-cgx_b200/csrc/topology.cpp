// Host-side (setup-time) construction of the restricted dof sets, node adjacency, node->cell incidence
// with packed adjacency slots, membrane tables and CSR row pointers.
//
// Replaces, as *structure* builders: multiphenicsx DofMapRestriction (KNPEMIx_problem.py:85-94),
// create_matrix_block's sparsity pattern (KNPEMIx_solver.py:157) and the dS integration-entity
// ordering (utils/mixed_dim_problem.py:708-729).  Nothing here runs per timestep.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <chrono>
#include <cstdlib>
#include <numeric>
#include "common.cuh"

namespace knp {

unsigned long long g_kernel_launches = 0;
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int build_topology(const knp_mesh_desc* m, HostTopo& T) {
  KNP_CHECK(m, "mesh descriptor is NULL");
  const int d = m->gdim;
  KNP_CHECK(d == 2 || d == 3, "gdim must be 2 or 3 (got %d)", d);
  const int nv = d + 1;
  const int64_t NV = m->n_vertices, NC = m->n_cells, NF = m->n_mfacets, NO = m->n_owned_vertices;
  KNP_CHECK(NV > 0 && NC >= 0 && NO >= 0 && NO <= NV, "bad vertex/cell counts");
  KNP_CHECK(NV < (int64_t)1 << 31 && NC < (int64_t)1 << 31, "local mesh too large for int32 indices");
  KNP_CHECK(m->n_quad > 0 && m->n_quad <= 64 && m->quad_bary && m->quad_w, "facet quadrature rule missing (1..64 points)");
  T.gdim = d;
  // optional phase timing on stderr (KNP_TOPO_TIMING=1)
  static const bool timing = getenv("KNP_TOPO_TIMING") && atoi(getenv("KNP_TOPO_TIMING"));
  auto t_prev = std::chrono::steady_clock::now();
  auto phase = [&](const char* name) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "build_topology: %-28s %.3f s\n", name, std::chrono::duration<double>(now - t_prev).count());
    t_prev = now;
  };

  std::vector<int32_t> itags(m->intra_tags, m->intra_tags + m->n_intra_tags);
  std::sort(itags.begin(), itags.end());
  KNP_CHECK(!std::binary_search(itags.begin(), itags.end(), m->extra_tag), "extra_tag is also listed as an intra tag");

  // ---- classify cells, restricted vertex sets ----
  std::vector<int8_t> sub(NC);
  std::vector<uint8_t> mark[2];
  mark[0].assign(NV, 0);
  mark[1].assign(NV, 0);
  int64_t ncell_s[2] = {0, 0};
  for (int64_t c = 0; c < NC; ++c) {
    int t = m->cell_tags[c];
    int s = (t == m->extra_tag) ? 1 : (std::binary_search(itags.begin(), itags.end(), t) ? 0 : -1);
    sub[c] = (int8_t)s;
    if (s < 0) continue;
    ++ncell_s[s];
    for (int a = 0; a < nv; ++a) {
      int32_t v = m->cell_verts[c * nv + a];
      KNP_CHECK(v >= 0 && v < NV, "cell %lld references vertex %d out of range", (long long)c, v);
      mark[s][v] = 1;
    }
  }
  std::vector<int32_t> r[2];
  Layout& L = T.L;
  for (int s = 0; s < 2; ++s) {
    r[s].assign(NV, -1);
    T.node_vert[s].clear();
    int own = 0;
    for (int64_t v = 0; v < NV; ++v)
      if (mark[s][v]) {
        r[s][v] = (int32_t)T.node_vert[s].size();
        T.node_vert[s].push_back((int32_t)v);
        if (v < NO) ++own;
      }
    L.n_loc[s] = (int)T.node_vert[s].size();
    L.n_own[s] = own;
    L.n_gh[s] = L.n_loc[s] - own;
  }
  KNP_CHECK((int64_t)4 * (L.n_loc[0] + L.n_loc[1]) < ((int64_t)1 << 31), "too many unknowns for int32 columns");
  L.rowbase[0] = 0;
  L.rowbase[1] = 4 * L.n_own[0];
  L.n_rows = 4 * (L.n_own[0] + L.n_own[1]);
  L.gbase[0] = 0;
  L.gbase[1] = 4 * L.n_gh[0];
  L.n_cols = L.n_rows + 4 * (L.n_gh[0] + L.n_gh[1]);
  const int nodeoff[2] = {0, L.n_loc[0]};
  const int workoff[2] = {0, L.n_own[0]};
  T.n_work = L.n_own[0] + L.n_own[1];

  T.node_x.resize((size_t)(L.n_loc[0] + L.n_loc[1]) * d);
  for (int s = 0; s < 2; ++s)
    for (int q = 0; q < L.n_loc[s]; ++q)
      for (int i = 0; i < d; ++i)
        T.node_x[(size_t)(nodeoff[s] + q) * d + i] = m->coords[(size_t)T.node_vert[s][q] * d + i];

  // ---- per-subdomain cell tables ----
  for (int s = 0; s < 2; ++s) {
    T.cell_nodes[s].clear();
    T.cell_nodes[s].reserve(ncell_s[s] * nv);
    T.cell_tag[s].clear();
    T.cell_owned[s].clear();
  }
  for (int64_t c = 0; c < NC; ++c) {
    int s = sub[c];
    if (s < 0) continue;
    for (int a = 0; a < nv; ++a) T.cell_nodes[s].push_back(r[s][m->cell_verts[c * nv + a]]);
    T.cell_tag[s].push_back(m->cell_tags[c]);
    T.cell_owned[s].push_back(m->cell_owned ? m->cell_owned[c] : 1);
  }

  phase("dof sets, cell tables");
  // ---- node -> cell incidence for owned nodes (cells in ascending order => fixed summation order) ----
  const int W = T.n_work;
  T.inc_ptr.assign(W + 1, 0);
  for (int s = 0; s < 2; ++s) {
    const auto& cn = T.cell_nodes[s];
    for (size_t k = 0; k < cn.size(); ++k)
      if (cn[k] < L.n_own[s]) ++T.inc_ptr[workoff[s] + cn[k] + 1];
  }
  for (int w = 0; w < W; ++w) T.inc_ptr[w + 1] += T.inc_ptr[w];
  KNP_CHECK(T.inc_ptr[W] >= 0, "incidence overflow");
  std::vector<int32_t> inc_cell(T.inc_ptr[W]);
  {
    std::vector<int32_t> fill(T.inc_ptr.begin(), T.inc_ptr.end() - 1);
    for (int s = 0; s < 2; ++s) {
      const auto& cn = T.cell_nodes[s];
      const size_t ncs = cn.size() / nv;
      for (size_t c = 0; c < ncs; ++c)
        for (int a = 0; a < nv; ++a) {
          int q = cn[c * nv + a];
          if (q < L.n_own[s]) inc_cell[fill[workoff[s] + q]++] = (int32_t)c;
        }
    }
  }
  // every owned node must touch a cell
  T.max_inc = 0;
  for (int w = 0; w < W; ++w) {
    KNP_CHECK(T.inc_ptr[w + 1] > T.inc_ptr[w], "owned dof %d has no incident cell", w);
    T.max_inc = std::max(T.max_inc, T.inc_ptr[w + 1] - T.inc_ptr[w]);
  }

  phase("incidence");
  // ---- adjacency (sorted unique subdomain-local node ids, includes the node itself) ----
  T.adj_ptr.assign(W + 1, 0);
  std::vector<int32_t> deg(W);
#pragma omp parallel for schedule(static)
  for (int w = 0; w < W; ++w) {
    const int s = w >= workoff[1] ? 1 : 0;
    const auto& cn = T.cell_nodes[s];
    int32_t tmp[1024];
    int cnt = 0;
    for (int k = T.inc_ptr[w]; k < T.inc_ptr[w + 1] && cnt + nv <= 1024; ++k)
      for (int a = 0; a < nv; ++a) tmp[cnt++] = cn[(size_t)inc_cell[k] * nv + a];
    std::sort(tmp, tmp + cnt);
    deg[w] = (int32_t)(std::unique(tmp, tmp + cnt) - tmp);
  }
  int maxdeg = 0;
  for (int w = 0; w < W; ++w) {
    T.adj_ptr[w + 1] = T.adj_ptr[w] + deg[w];
    maxdeg = std::max(maxdeg, deg[w]);
    KNP_CHECK((T.inc_ptr[w + 1] - T.inc_ptr[w]) * nv <= 1024, "vertex valence too large");
  }
  KNP_CHECK(maxdeg <= 255, "vertex degree %d exceeds 255", maxdeg);
  T.max_deg = maxdeg;
  T.adj_idx.resize(T.adj_ptr[W]);
  T.inc_slots.resize(inc_cell.size());
  T.self_slot.resize(W);
#pragma omp parallel for schedule(static)
  for (int w = 0; w < W; ++w) {
    const int s = w >= workoff[1] ? 1 : 0;
    const int q_self = w - workoff[s];
    const auto& cn = T.cell_nodes[s];
    int32_t tmp[1024];
    int cnt = 0;
    for (int k = T.inc_ptr[w]; k < T.inc_ptr[w + 1]; ++k)
      for (int a = 0; a < nv; ++a) tmp[cnt++] = cn[(size_t)inc_cell[k] * nv + a];
    std::sort(tmp, tmp + cnt);
    cnt = (int)(std::unique(tmp, tmp + cnt) - tmp);
    int32_t* row = &T.adj_idx[T.adj_ptr[w]];
    std::copy(tmp, tmp + cnt, row);
    T.self_slot[w] = (int32_t)(std::lower_bound(row, row + cnt, q_self) - row);
    for (int k = T.inc_ptr[w]; k < T.inc_ptr[w + 1]; ++k) {
      uint32_t packed = 0;
      for (int a = 0; a < nv; ++a) {
        int q = cn[(size_t)inc_cell[k] * nv + a];
        uint32_t sl = (uint32_t)(std::lower_bound(row, row + cnt, q) - row);
        packed |= sl << (8 * a);
      }
      T.inc_slots[k] = packed;
    }
  }

  phase("adjacency, packed slots");
  // ---- membrane ----
  T.n_mf = (int)NF;
  std::vector<int32_t> mvid(NV, -1);
  for (int64_t f = 0; f < NF; ++f)
    for (int a = 0; a < d; ++a) {
      int32_t v = m->mfacet_verts[f * d + a];
      KNP_CHECK(v >= 0 && v < NV, "membrane facet %lld references vertex %d out of range", (long long)f, v);
      mvid[v] = 0;
    }
  T.mv_vert.clear();
  for (int64_t v = 0; v < NV; ++v)
    if (mvid[v] == 0) {
      mvid[v] = (int32_t)T.mv_vert.size();
      T.mv_vert.push_back((int32_t)v);
    }
  T.n_mv = (int)T.mv_vert.size();
  for (int s = 0; s < 2; ++s) {
    T.mv_node[s].resize(T.n_mv);
    for (int g = 0; g < T.n_mv; ++g) {
      T.mv_node[s][g] = r[s][T.mv_vert[g]];
      KNP_CHECK(T.mv_node[s][g] >= 0, "membrane vertex %d is not a vertex of an %s cell", T.mv_vert[g],
                s == 0 ? "intracellular" : "extracellular");
    }
  }
  T.mf_mv.resize((size_t)NF * d);
  T.mtags.assign(m->mfacet_tags, m->mfacet_tags + NF);
  std::sort(T.mtags.begin(), T.mtags.end());
  T.mtags.erase(std::unique(T.mtags.begin(), T.mtags.end()), T.mtags.end());
  T.mf_tagidx.resize(NF);
  T.mf_owned.resize(NF);
  T.mf_area.resize(NF);
  for (int64_t f = 0; f < NF; ++f) {
    for (int a = 0; a < d; ++a) T.mf_mv[f * d + a] = mvid[m->mfacet_verts[f * d + a]];
    {
      const double* x0 = &m->coords[(size_t)m->mfacet_verts[f * d + 0] * d];
      const double* x1 = &m->coords[(size_t)m->mfacet_verts[f * d + 1] * d];
      if (d == 2) {
        T.mf_area[f] = std::sqrt((x1[0] - x0[0]) * (x1[0] - x0[0]) + (x1[1] - x0[1]) * (x1[1] - x0[1]));
      } else {
        const double* x2 = &m->coords[(size_t)m->mfacet_verts[f * d + 2] * d];
        double a0 = x1[0] - x0[0], a1 = x1[1] - x0[1], a2 = x1[2] - x0[2];
        double b0 = x2[0] - x0[0], b1 = x2[1] - x0[1], b2 = x2[2] - x0[2];
        double c0 = a1 * b2 - a2 * b1, c1 = a2 * b0 - a0 * b2, c2 = a0 * b1 - a1 * b0;
        T.mf_area[f] = 0.5 * std::sqrt(c0 * c0 + c1 * c1 + c2 * c2);
      }
    }
    T.mf_tagidx[f] = (int32_t)(std::lower_bound(T.mtags.begin(), T.mtags.end(), m->mfacet_tags[f]) - T.mtags.begin());
    T.mf_owned[f] = m->mfacet_owned ? m->mfacet_owned[f] : 1;
  }
  // membrane-vertex -> facets, gamma adjacency
  T.minc_ptr.assign(T.n_mv + 1, 0);
  for (size_t k = 0; k < T.mf_mv.size(); ++k) ++T.minc_ptr[T.mf_mv[k] + 1];
  for (int g = 0; g < T.n_mv; ++g) T.minc_ptr[g + 1] += T.minc_ptr[g];
  std::vector<int32_t> minc_f(T.minc_ptr[T.n_mv]), minc_a(T.minc_ptr[T.n_mv]);
  {
    std::vector<int32_t> fill(T.minc_ptr.begin(), T.minc_ptr.end() - 1);
    for (int64_t f = 0; f < NF; ++f)
      for (int a = 0; a < d; ++a) {
        int g = T.mf_mv[f * d + a];
        minc_f[fill[g]] = (int32_t)f;
        minc_a[fill[g]] = a;
        ++fill[g];
      }
  }
  T.gam_ptr.assign(T.n_mv + 1, 0);
  std::vector<std::vector<int32_t>> gam(T.n_mv);
  int maxg = 0;
  for (int g = 0; g < T.n_mv; ++g) {
    auto& lst = gam[g];
    for (int k = T.minc_ptr[g]; k < T.minc_ptr[g + 1]; ++k)
      for (int a = 0; a < d; ++a) lst.push_back(T.mf_mv[(size_t)minc_f[k] * d + a]);
    std::sort(lst.begin(), lst.end());
    lst.erase(std::unique(lst.begin(), lst.end()), lst.end());
    T.gam_ptr[g + 1] = T.gam_ptr[g] + (int32_t)lst.size();
    maxg = std::max(maxg, (int)lst.size());
  }
  KNP_CHECK(maxg <= 255, "membrane degree too large");
  T.max_gdeg = maxg;
  T.gam_mv.resize(T.gam_ptr[T.n_mv]);
  for (int g = 0; g < T.n_mv; ++g) std::copy(gam[g].begin(), gam[g].end(), T.gam_mv.begin() + T.gam_ptr[g]);

  T.mv_of_node.assign(W, -1);
  for (int g = 0; g < T.n_mv; ++g)
    for (int s = 0; s < 2; ++s)
      if (T.mv_node[s][g] < L.n_own[s]) T.mv_of_node[workoff[s] + T.mv_node[s][g]] = g;

  T.minc.assign((size_t)T.minc_ptr[T.n_mv] * 4, 0u);
  for (int g = 0; g < T.n_mv; ++g) {
    const bool owned = T.mv_vert[g] < NO;
    for (int k = T.minc_ptr[g]; k < T.minc_ptr[g + 1]; ++k) {
      const int f = minc_f[k];
      uint32_t w1 = (uint32_t)minc_a[k], w2 = 0, w3 = 0;
      if (owned) {
        for (int s = 0; s < 2; ++s) {
          const int w = workoff[s] + T.mv_node[s][g];
          const int32_t* row = &T.adj_idx[T.adj_ptr[w]];
          const int cnt = T.adj_ptr[w + 1] - T.adj_ptr[w];
          for (int b = 0; b < d; ++b) {
            int q = T.mv_node[s][T.mf_mv[(size_t)f * d + b]];
            const int32_t* it = std::lower_bound(row, row + cnt, q);
            KNP_CHECK(it != row + cnt && *it == q, "membrane facet %d is not a face of an %s cell at vertex %d", f,
                      s == 0 ? "intracellular" : "extracellular", T.mv_vert[g]);
            uint32_t sl = (uint32_t)(it - row);
            if (s == 0) w1 |= sl << (8 * (b + 1));
            else w2 |= sl << (8 * b);
          }
        }
        for (int b = 0; b < d; ++b) {
          int gb = T.mf_mv[(size_t)f * d + b];
          uint32_t sl = (uint32_t)(std::lower_bound(gam[g].begin(), gam[g].end(), gb) - gam[g].begin());
          w3 |= sl << (8 * b);
        }
      }
      T.minc[(size_t)k * 4 + 0] = (uint32_t)f;
      T.minc[(size_t)k * 4 + 1] = w1;
      T.minc[(size_t)k * 4 + 2] = w2;
      T.minc[(size_t)k * 4 + 3] = w3;
    }
  }

  phase("membrane tables");
  // ---- CSR row pointers of A and P ----
  T.gpre.assign(W + 1, 0);
  for (int w = 0; w < W; ++w) {
    int g = T.mv_of_node[w];
    T.gpre[w + 1] = T.gpre[w] + (g >= 0 ? T.gam_ptr[g + 1] - T.gam_ptr[g] : 0);
  }
  T.indptr.assign((size_t)L.n_rows + 1, 0);
  T.indptr_P.assign((size_t)L.n_rows + 1, 0);
  int64_t pos = 0, posP = 0;
  for (int s = 0; s < 2; ++s)
    for (int f = 0; f < 4; ++f)
      for (int p = 0; p < L.n_own[s]; ++p) {
        const int w = workoff[s] + p;
        const int dg = T.adj_ptr[w + 1] - T.adj_ptr[w];
        const int gd = T.gpre[w + 1] - T.gpre[w];
        const int row = L.row(s, f, p);
        T.indptr[row] = (int32_t)pos;
        T.indptr_P[row] = (int32_t)posP;
        pos += (f < 3 ? 2 : 4) * dg + gd;
        posP += dg;
      }
  KNP_CHECK(pos < ((int64_t)1 << 31), "nnz(A) = %lld does not fit int32 row pointers; shard the mesh over more GPUs",
            (long long)pos);
  T.indptr[L.n_rows] = (int32_t)pos;
  T.indptr_P[L.n_rows] = (int32_t)posP;
  T.nnz = pos;
  T.nnz_P = posP;

  phase("row pointers");
  // ---- lane-group tables of the edge-lane row kernel (assembly.cu::rows_edge_kernel) ----
  // A dof is served by G = 2^lgG lanes, lane e = adjacency slot e.  Per (dof, slot), at index (w << lgG) + e:
  //   adjG   the neighbour (subdomain-local node id), -1 beyond the degree
  //   hitG   the cells that contain the edge (dof, neighbour), in ascending cell order, each written as the adjacency
  //          slots of the cell's OTHER vertices: 2D one byte per cell (4 bytes = one word, at most 2 used), 3D two bytes
  //          per cell (8 cells = four words); unused entries are 0xFF bytes.  The self slot has no hits (its sums follow
  //          from the row-sum identities of the P1 element matrices).
  // Per dof: metaG = {deg | self << 8 | gamma degree << 16, membrane vertex or -1}.
  // edge_ok = 0 (ring of an edge longer than the table holds, or G > 32): the scan kernel serves the mesh.
  {
    int lg = 0;
    while ((1 << lg) < std::max(std::max(T.max_deg, T.max_gdeg), 4)) ++lg;
    T.lgG = lg;
    const int HW = d == 2 ? 1 : 4, HMAX = d == 2 ? 4 : 8;
    T.edge_ok = lg <= 5 ? 1 : 0;
    if (T.edge_ok) {
      const size_t NE = (size_t)W << lg;
      T.adjG.assign(NE, -1);
      T.hitG.assign(NE * HW, 0xFFFFFFFFu);
      T.metaG.assign((size_t)W * 2, 0);
      int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
      for (int w = 0; w < W; ++w) {
        const int a0 = T.adj_ptr[w], dg = T.adj_ptr[w + 1] - a0, self = T.self_slot[w];
        const int i0 = T.inc_ptr[w], ninc = T.inc_ptr[w + 1] - i0;
        const int gd = T.gpre[w + 1] - T.gpre[w];
        T.metaG[(size_t)w * 2 + 0] = dg | (self << 8) | (gd << 16);
        T.metaG[(size_t)w * 2 + 1] = T.mv_of_node[w];
        int cnt[256];
        for (int e = 0; e < dg; ++e) {
          T.adjG[((size_t)w << lg) + e] = T.adj_idx[a0 + e];
          cnt[e] = 0;
        }
        uint8_t* hb = reinterpret_cast<uint8_t*>(T.hitG.data());
        for (int j = 0; j < ninc; ++j) {
          const uint32_t pk = T.inc_slots[i0 + j];
          int sl[4];
          for (int b = 0; b < nv; ++b) sl[b] = (int)((pk >> (8 * b)) & 255u);
          for (int b = 0; b < nv; ++b) {
            const int e = sl[b];
            if (e == self) continue;
            if (cnt[e] >= HMAX) {
              ++bad;
              continue;
            }
            uint8_t* dst = hb + ((((size_t)w << lg) + e) * HW) * 4 + (size_t)cnt[e] * (d - 1);
            int k = 0;
            for (int c = 0; c < nv; ++c)
              if (c != b && sl[c] != self) dst[k++] = (uint8_t)sl[c];
            ++cnt[e];
          }
        }
      }
      if (bad) {
        T.edge_ok = 0;
        std::vector<int32_t>().swap(T.adjG);
        std::vector<uint32_t>().swap(T.hitG);
        std::vector<int32_t>().swap(T.metaG);
      }
    }
  }
  phase("lane-group tables");
  return KNP_OK;
}

}  // namespace knp
