// C-ABI entry points of libwae_b200 (see include/wae_b200.h): context, mesh, patterns,
// assembly, operator family (combine / SpMM).  LU, Arnoldi and Beyn live in lu_api.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <thread>

#include "lu.h"
#include "wae_internal.h"

void wae_launch_flame(wae_ctx* h, const int32_t* d_tets, int64_t n, const int32_t* d_rowpos, double* d_S, double* d_vol,
                      int64_t ref_tet, const double* x_ref, const double* n_ref, double fac, double* d_G,
                      const int32_t* d_colsrc, int64_t nrows, int ncols, double* d_out);

#define WAE_API_BEGIN \
  if (!h) return WAE_E_INVALID; \
  try {
#define WAE_API_END                       \
  }                                       \
  catch (const WaeError& e) {             \
    h->err = e.msg;                       \
    return e.code;                        \
  }                                       \
  catch (const std::bad_alloc&) {         \
    h->err = "host allocation failed";    \
    return WAE_E_NOMEM;                   \
  }                                       \
  catch (const std::exception& e) {       \
    h->err = e.what();                    \
    return WAE_E_INVALID;                 \
  }                                       \
  return WAE_OK;

extern "C" {

int32_t wae_create(wae_ctx** out, int32_t device, int32_t index_base) {
  if (!out || (index_base != 0 && index_base != 1)) return WAE_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return WAE_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return WAE_E_CUDA;
  wae_ctx* h = new wae_ctx();
  h->device = device;
  h->base = index_base;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  *out = h;
  return WAE_OK;
}

int32_t wae_destroy(wae_ctx* h) {
  if (!h) return WAE_E_INVALID;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->lus.clear();
  h->fams.clear();
  h->mats.clear();
  h->patterns.clear();
  for (int g = 0; g < 2; g++) {
    if (h->aux_stream[g]) cudaStreamDestroy(h->aux_stream[g]);
    if (h->ev_join[g]) cudaEventDestroy(h->ev_join[g]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return WAE_OK;
}

const char* wae_last_error(wae_ctx* h) { return h ? h->err.c_str() : "null context"; }

int32_t wae_set_stream(wae_ctx* h, void* s) {
  if (!h) return WAE_E_INVALID;
  h->stream = (cudaStream_t)s;
  return WAE_OK;
}

int32_t wae_sync(wae_ctx* h) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}

int64_t wae_launch_count(wae_ctx* h) { return h ? h->launches : -1; }

double wae_last_ms(wae_ctx* h, const char* phase) {
  if (!h || !phase) return -1.0;
  auto it = h->last_ms.find(phase);
  return it == h->last_ms.end() ? -1.0 : it->second;
}

// ---- mesh -------------------------------------------------------------------------------
int32_t wae_mesh_set(wae_ctx* h, int32_t order, int64_t n_pts, const double* xyz, int64_t n_tet, const uint32_t* tets,
                     int64_t n_tri, const uint32_t* tris, int64_t dim) {
  WAE_API_BEGIN
  if (order != 1 && order != 2) WAE_THROW(WAE_E_INVALID, "order must be 1 (:lin) or 2 (:quad)");
  if (n_pts <= 0 || !xyz || n_tet < 0 || n_tri < 0 || dim < n_pts) WAE_THROW(WAE_E_INVALID, "bad mesh sizes");
  CUDA_CHECK(cudaSetDevice(h->device));
  h->order = order;
  h->nloc = order == 1 ? 4 : 10;
  h->nloc3 = order == 1 ? 3 : 6;
  h->n_pts = n_pts; h->n_tet = n_tet; h->n_tri = n_tri; h->dim = dim;
  h->xyz.assign(xyz, xyz + 3 * n_pts);
  auto conv = [&](const uint32_t* src, int64_t n, std::vector<uint32_t>& dst) {
    dst.resize(n);
    for (int64_t i = 0; i < n; i++) {
      int64_t v = (int64_t)src[i] - h->base;
      if (v < 0 || v >= dim) WAE_THROW(WAE_E_INVALID, "element DOF index %lld out of range [0,%lld)", (long long)v, (long long)dim);
      dst[i] = (uint32_t)v;
    }
  };
  conv(tets, n_tet * h->nloc, h->tets);
  conv(tris, n_tri * h->nloc3, h->tris);
  for (int64_t e = 0; e < n_tet; e++)
    for (int k = 0; k < 4; k++)
      if (h->tets[e * h->nloc + k] >= (uint64_t)n_pts) WAE_THROW(WAE_E_INVALID, "tet %lld: vertex DOF >= n_pts", (long long)e);
  h->d_xyz.upload(h->xyz, h->stream);
  h->xyz_version++;
  h->d_tets.upload(h->tets, h->stream);
  h->d_tris.upload(h->tris, h->stream);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}

int32_t wae_mesh_update_points(wae_ctx* h, int64_t n_pts, const double* xyz) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  if (!h->order || n_pts != h->n_pts || !xyz) WAE_THROW(WAE_E_INVALID, "wae_mesh_update_points: point count differs from wae_mesh_set");
  h->xyz.assign(xyz, xyz + 3 * n_pts);
  CUDA_CHECK(cudaMemcpyAsync(h->d_xyz.p, xyz, 3 * n_pts * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->xyz_version++;
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}

// ---- patterns ---------------------------------------------------------------------------
static int new_pattern(wae_ctx* h) {
  h->patterns.emplace_back(new Pattern());
  return (int)h->patterns.size() - 1;
}

static void upload_pattern(wae_ctx* h, Pattern& P) {
  P.d_colptr.upload(P.colptr, h->stream);
  P.d_rowval.upload(P.rowval, h->stream);
}

int32_t wae_pattern_build(wae_ctx* h, int32_t elem_kind, int64_t n_elem, const int64_t* elem_ids, int32_t* pattern_id,
                          int64_t* nnz) {
  WAE_API_BEGIN
  if (!h->order) WAE_THROW(WAE_E_INVALID, "wae_mesh_set has not been called");
  if (elem_kind != 2 && elem_kind != 3) WAE_THROW(WAE_E_INVALID, "elem_kind must be 2 (triangles) or 3 (tetrahedra)");
  CUDA_CHECK(cudaSetDevice(h->device));
  int64_t total = elem_kind == 3 ? h->n_tet : h->n_tri;
  int id = new_pattern(h);
  Pattern& P = *h->patterns[id];
  P.elem_kind = elem_kind;
  if (elem_ids) {
    P.elems.resize(n_elem);
    for (int64_t i = 0; i < n_elem; i++) {
      int64_t e = elem_ids[i] - h->base;
      if (e < 0 || e >= total) WAE_THROW(WAE_E_INVALID, "element id %lld out of range", (long long)elem_ids[i]);
      P.elems[i] = e;
    }
  } else {
    P.elems.resize(total);
    std::iota(P.elems.begin(), P.elems.end(), 0);
  }
  const uint32_t* conn = elem_kind == 3 ? h->tets.data() : h->tris.data();
  int nloc = elem_kind == 3 ? h->nloc : h->nloc3;
  wae_build_pattern_from_elements(conn, nloc, P.elems, h->dim, P);
  upload_pattern(h, P);
  std::vector<int32_t> e32(P.elems.begin(), P.elems.end());
  P.d_elems.upload(e32, h->stream);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (pattern_id) *pattern_id = id;
  if (nnz) *nnz = P.nnz;
  WAE_API_END
}

int32_t wae_pattern_get(wae_ctx* h, int32_t pattern_id, int64_t* colptr, int64_t* rowval) {
  WAE_API_BEGIN
  Pattern& P = h->pat(pattern_id);
  if (colptr)
    for (int64_t j = 0; j <= P.dim; j++) colptr[j] = P.colptr[j] + h->base;
  if (rowval)
    for (int64_t k = 0; k < P.nnz; k++) rowval[k] = (int64_t)P.rowval[k] + h->base;
  WAE_API_END
}

static void ensure_slotmap(wae_ctx* h, Pattern& P) {
  if (P.slotmap_built) return;
  const uint32_t* conn = P.elem_kind == 3 ? h->tets.data() : h->tris.data();
  int nloc = P.elem_kind == 3 ? h->nloc : h->nloc3;
  std::vector<int32_t> sm;
  wae_build_slotmap(conn, nloc, P, sm);
  P.d_slotmap.upload(sm, h->stream);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  P.slotmap_built = true;
}

// *mat_id >= 0 on entry: overwrite that matrix in place (same pattern and type), else create a new one
// zero = false: the caller overwrites every nonzero (pair-program assembly) -- no memset pass
static int reuse_or_new_matrix(wae_ctx* h, const int32_t* mat_id, int pattern, bool is_complex, int64_t nnz, bool zero = true);

static int new_matrix(wae_ctx* h, int pattern, bool is_complex, int64_t nnz, bool zero = true) {
  h->mats.emplace_back(new Matrix());
  Matrix& M = *h->mats.back();
  M.pattern = pattern;
  M.is_complex = is_complex;
  M.d_val.alloc((size_t)nnz * (is_complex ? 2 : 1));
  if (zero) CUDA_CHECK(cudaMemsetAsync(M.d_val.p, 0, M.d_val.n * sizeof(double), h->stream));
  return (int)h->mats.size() - 1;
}

static int reuse_or_new_matrix(wae_ctx* h, const int32_t* mat_id, int pattern, bool is_complex, int64_t nnz, bool zero) {
  if (mat_id && *mat_id >= 0) {
    Matrix& M = h->mat(*mat_id);
    if (M.pattern != pattern || M.is_complex != is_complex) WAE_THROW(WAE_E_INVALID, "matrix %d cannot be reused: different pattern or type", *mat_id);
    if (zero) CUDA_CHECK(cudaMemsetAsync(M.d_val.p, 0, M.d_val.n * sizeof(double), h->stream));
    return *mat_id;
  }
  return new_matrix(h, pattern, is_complex, nnz, zero);
}

// M/K assembly kernel generation: 3 = star program (default for P2: 0.37 ms against 0.78 ms on the 64^3 box), 2 = owner-computes pair
// program (default for P1, where the two are within 5 % and the pair program is slightly ahead); WAE_ASM_GEN=2|3 forces one
static int asm_generation(const wae_ctx* h) {
  const char* env = getenv("WAE_ASM_GEN");
  if (env && (atoi(env) == 2 || atoi(env) == 3)) return atoi(env);
  return h->nloc == 10 ? 3 : 2;
}

static void upload_c(wae_ctx* h, Pattern& P, const double* c, int c_per_elem, DevBuf<double>& d_c) {
  if (!c) WAE_THROW(WAE_E_INVALID, "speed-of-sound array is NULL");
  int need = P.elem_kind == 3 ? 4 : 3;
  if (c_per_elem != 1 && c_per_elem != need) WAE_THROW(WAE_E_INVALID, "c_per_elem must be 1 or %d", need);
  d_c.upload(c, P.elems.size() * (size_t)c_per_elem, h->stream);
}

int32_t wae_assemble(wae_ctx* h, int32_t pattern_id, int32_t kind, const double* c, int32_t c_per_elem, double scale,
                     int32_t* mat_id) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  Pattern& P = h->pat(pattern_id);
  if (kind == WAE_OP_BOUNDARY ? P.elem_kind != 2 : P.elem_kind != 3)
    WAE_THROW(WAE_E_INVALID, "operator kind %d does not match the pattern's element kind", kind);
  if (kind < WAE_OP_MASS || kind > WAE_OP_BOUNDARY) WAE_THROW(WAE_E_INVALID, "unknown operator kind %d", kind);
  const bool use_gather = kind == WAE_OP_MASS && !getenv("WAE_FORCE_ATOMIC");
  const bool use_star = use_gather && asm_generation(h) == 3 && wae_ensure_star(h, P);
  if (use_star) {
  } else if (use_gather) wae_ensure_gather(h, P); else ensure_slotmap(h, P);
  DevBuf<double>& d_c = h->scratch_c;
  if (kind != WAE_OP_MASS) upload_c(h, P, c, c_per_elem, d_c);
  int id = reuse_or_new_matrix(h, mat_id, pattern_id, kind == WAE_OP_BOUNDARY, P.nnz, !use_gather);
  Matrix& M = *h->mats[id];
  PhaseTimer t(h, "assemble");
  if (kind == WAE_OP_MASS && use_star)
    wae_launch_assemble_star(h, P, nullptr, M.d_val.p, nullptr, scale);
  else if (kind == WAE_OP_MASS && use_gather)
    wae_launch_assemble_gather(h, P, nullptr, M.d_val.p, nullptr, scale);
  else if (kind == WAE_OP_MASS)
    wae_launch_assemble_atomic(h, P, 1, nullptr, 1, scale, M.d_val.p, nullptr);
  else if (kind == WAE_OP_STIFF) {
    if (scale != 1.0) WAE_THROW(WAE_E_INVALID, "scale != 1 is only supported for WAE_OP_MASS / WAE_OP_BOUNDARY");
    wae_launch_assemble_atomic(h, P, 2, d_c.p, c_per_elem, 1.0, nullptr, M.d_val.p);
  } else
    wae_launch_assemble_atomic(h, P, 0, d_c.p, c_per_elem, scale, M.d_val.p, nullptr);
  t.stop();
  M.symmetric = true;  // int phi_i phi_j, int grad phi_i . grad phi_j and the boundary mass are symmetric by construction
  if (mat_id) *mat_id = id;
  WAE_API_END
}

int32_t wae_assemble_mk(wae_ctx* h, int32_t pattern_id, const double* c, int32_t c_per_elem, int32_t* mass_id,
                        int32_t* stiff_id) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  Pattern& P = h->pat(pattern_id);
  if (P.elem_kind != 3) WAE_THROW(WAE_E_INVALID, "wae_assemble_mk needs a tetrahedral pattern");
  DevBuf<double>& d_c = h->scratch_c;
  upload_c(h, P, c, c_per_elem, d_c);
  const bool use_gather = c_per_elem == 1 && !getenv("WAE_FORCE_ATOMIC");
  int im = reuse_or_new_matrix(h, mass_id, pattern_id, false, P.nnz, !use_gather);
  int ik = reuse_or_new_matrix(h, stiff_id, pattern_id, false, P.nnz, !use_gather);
  if (use_gather && asm_generation(h) == 3 && wae_ensure_star(h, P)) {
    PhaseTimer t(h, "assemble");
    wae_launch_assemble_star(h, P, d_c.p, h->mats[im]->d_val.p, h->mats[ik]->d_val.p, 1.0);
    t.stop();
  } else if (use_gather) {
    wae_ensure_gather(h, P);
    PhaseTimer t(h, "assemble");
    wae_launch_assemble_gather(h, P, d_c.p, h->mats[im]->d_val.p, h->mats[ik]->d_val.p, 1.0);
    t.stop();
  } else {
    ensure_slotmap(h, P);
    PhaseTimer t(h, "assemble");
    wae_launch_assemble_atomic(h, P, 3, d_c.p, c_per_elem, 1.0, h->mats[im]->d_val.p, h->mats[ik]->d_val.p);
    t.stop();
  }
  h->mats[im]->symmetric = h->mats[ik]->symmetric = true;
  if (mass_id) *mass_id = im;
  if (stiff_id) *stiff_id = ik;
  WAE_API_END
}

int32_t wae_assemble_wallsrc(wae_ctx* h, int64_t n_tri, const int64_t* tri_ids, const double* c, int32_t c_per_elem, double* out) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  if (!h->order || n_tri < 0 || (n_tri && (!tri_ids || !c)) || !out || (c_per_elem != 1 && c_per_elem != 3))
    WAE_THROW(WAE_E_INVALID, "bad speaker-source arguments (c holds 1 or 3 values per triangle)");
  std::vector<int32_t> el(n_tri);
  for (int64_t i = 0; i < n_tri; i++) {
    const int64_t e = tri_ids[i] - h->base;
    if (e < 0 || e >= h->n_tri) WAE_THROW(WAE_E_INVALID, "speaker triangle %lld out of range", (long long)tri_ids[i]);
    el[i] = (int32_t)e;
  }
  DevBuf<int32_t> d_el;
  DevBuf<double> d_c, d_out;
  d_el.upload(el, h->stream);
  d_c.upload(c, (size_t)n_tri * c_per_elem, h->stream);
  d_out.alloc((size_t)2 * h->dim);
  CUDA_CHECK(cudaMemsetAsync(d_out.p, 0, (size_t)2 * h->dim * sizeof(double), h->stream));
  PhaseTimer t(h, "assemble");
  wae_launch_wallsrc(h, d_el.p, n_tri, d_c.p, c_per_elem, d_out.p);
  t.stop();
  CUDA_CHECK(cudaMemcpyAsync(out, d_out.p, (size_t)2 * h->dim * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}

int32_t wae_assemble_flame(wae_ctx* h, int64_t n_flame, const int64_t* flame_tets, int64_t ref_tet, const double* x_ref,
                           const double* n_ref, double nlocal, int32_t* pattern_id, int32_t* mat_id, int64_t* nnz) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  if (!h->order || n_flame <= 0 || !flame_tets || !x_ref || !n_ref) WAE_THROW(WAE_E_INVALID, "bad flame arguments");
  ref_tet -= h->base;
  if (ref_tet < 0 || ref_tet >= h->n_tet) WAE_THROW(WAE_E_INVALID, "reference tetrahedron out of range");
  const int nloc = h->nloc;
  std::vector<int32_t> ft(n_flame);
  std::vector<int32_t> rows;
  rows.reserve(n_flame * nloc);
  for (int64_t i = 0; i < n_flame; i++) {
    int64_t e = flame_tets[i] - h->base;
    if (e < 0 || e >= h->n_tet) WAE_THROW(WAE_E_INVALID, "flame tetrahedron out of range");
    ft[i] = (int32_t)e;
    for (int k = 0; k < nloc; k++) rows.push_back((int32_t)h->tets[e * nloc + k]);
  }
  std::vector<int32_t> urows(rows);
  std::sort(urows.begin(), urows.end());
  urows.erase(std::unique(urows.begin(), urows.end()), urows.end());
  std::vector<int32_t> rowpos(rows.size());
  for (size_t i = 0; i < rows.size(); i++)
    rowpos[i] = (int32_t)(std::lower_bound(urows.begin(), urows.end(), rows[i]) - urows.begin());
  // columns: DOFs of the reference tet, sorted; colsrc[c] = local index in the tet
  std::vector<std::pair<int32_t, int32_t>> cols;
  for (int k = 0; k < nloc; k++) cols.emplace_back((int32_t)h->tets[ref_tet * nloc + k], k);
  std::sort(cols.begin(), cols.end());
  const bool reuse = mat_id && *mat_id >= 0;
  int pid = reuse ? h->mat(*mat_id).pattern : new_pattern(h);
  Pattern& P = *h->patterns[pid];
  int64_t nr = (int64_t)urows.size();
  if (reuse && P.nnz != nr * nloc) WAE_THROW(WAE_E_INVALID, "flame matrix %d cannot be reused: different pattern", *mat_id);
  if (!reuse) {
  P.dim = h->dim;
  P.colptr.assign(h->dim + 1, 0);
  for (auto& cpair : cols) P.colptr[cpair.first + 1] = nr;
  for (int64_t j = 0; j < h->dim; j++) P.colptr[j + 1] += P.colptr[j];
  P.nnz = nr * nloc;
  P.rowval.resize(P.nnz);
  for (int c = 0; c < nloc; c++) std::copy(urows.begin(), urows.end(), P.rowval.begin() + (size_t)c * nr);
  upload_pattern(h, P);
  }
  std::vector<int32_t> colsrc(nloc);
  for (int c = 0; c < nloc; c++) colsrc[c] = cols[c].second;
  DevBuf<int32_t>&d_ft = h->scratch_i[0], &d_rowpos = h->scratch_i[1], &d_colsrc = h->scratch_i[2];
  DevBuf<double>&d_S = h->scratch_d[0], &d_G = h->scratch_d[1];
  d_ft.upload(ft, h->stream);
  d_rowpos.upload(rowpos, h->stream);
  d_colsrc.upload(colsrc, h->stream);
  d_S.reserve(nr + 1);
  d_G.reserve(nloc);
  CUDA_CHECK(cudaMemsetAsync(d_S.p, 0, (nr + 1) * sizeof(double), h->stream));
  int mid = reuse_or_new_matrix(h, mat_id, pid, false, P.nnz);
  PhaseTimer t(h, "assemble");
  wae_launch_flame(h, d_ft.p, n_flame, d_rowpos.p, d_S.p, d_S.p + nr, ref_tet, x_ref, n_ref, -nlocal, d_G.p, d_colsrc.p, nr,
                   nloc, h->mats[mid]->d_val.p);
  t.stop();
  {  // keep the rank-1 factors: Q = S (x) G
    Matrix& Q = *h->mats[mid];
    Q.rank1 = true;
    Q.r1_rows = urows;
    Q.r1_cols.resize(nloc);
    for (int c = 0; c < nloc; c++) Q.r1_cols[c] = cols[c].first;
    Q.d_r1_rows.upload(Q.r1_rows, h->stream);
    Q.d_r1_cols.upload(Q.r1_cols, h->stream);
    Q.d_r1_S.reserve(nr);
    Q.d_r1_G.reserve(nloc);
    CUDA_CHECK(cudaMemcpyAsync(Q.d_r1_S.p, d_S.p, nr * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    // G in the order of the sorted columns
    std::vector<double> Gh(nloc), Gs(nloc);
    CUDA_CHECK(cudaMemcpyAsync(Gh.data(), d_G.p, nloc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < nloc; c++) Gs[c] = Gh[colsrc[c]];
    Q.d_r1_G.upload(Gs, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
  if (pattern_id) *pattern_id = pid;
  if (mat_id) *mat_id = mid;
  if (nnz) *nnz = P.nnz;
  WAE_API_END
}

int32_t wae_assemble_bloch(wae_ctx* h, int32_t elem_kind, int64_t n_elem, const int64_t* elem_ids, int32_t kind, const double* c,
                           int32_t c_per_elem, double scale, int64_t dim_red, const int64_t* dof_new, const uint8_t* dof_flag,
                           int32_t n_class, int32_t* pattern_ids, int32_t* mat_ids) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  if (!h->order) WAE_THROW(WAE_E_INVALID, "wae_mesh_set has not been called");
  if (n_class != 1 && n_class != 3 && n_class != 6) WAE_THROW(WAE_E_INVALID, "n_class must be 1, 3 or 6");
  if (kind < WAE_OP_MASS || kind > WAE_OP_BOUNDARY || (kind == WAE_OP_BOUNDARY) != (elem_kind == 2)) WAE_THROW(WAE_E_INVALID, "operator kind does not match the element kind");
  if (!dof_new || !dof_flag || dim_red <= 0 || !pattern_ids || !mat_ids) WAE_THROW(WAE_E_INVALID, "bad Bloch arguments");
  const int64_t total = elem_kind == 3 ? h->n_tet : h->n_tri;
  Pattern work;  // carries the element list for the kernels
  work.elem_kind = elem_kind;
  if (elem_ids) {
    work.elems.resize(n_elem);
    for (int64_t i = 0; i < n_elem; i++) {
      int64_t e = elem_ids[i] - h->base;
      if (e < 0 || e >= total) WAE_THROW(WAE_E_INVALID, "element id out of range");
      work.elems[i] = e;
    }
  } else {
    work.elems.resize(total);
    std::iota(work.elems.begin(), work.elems.end(), 0);
  }
  std::vector<int64_t> dn(h->dim);
  for (int64_t d = 0; d < h->dim; d++) {
    dn[d] = dof_new[d] - h->base;
    if (dn[d] < 0 || dn[d] >= dim_red) WAE_THROW(WAE_E_INVALID, "folded DOF index out of range");
  }
  const uint32_t* conn = elem_kind == 3 ? h->tets.data() : h->tris.data();
  const int nloc = elem_kind == 3 ? h->nloc : h->nloc3;
  std::vector<Pattern> cls;
  std::vector<int32_t> slotmap;
  std::vector<int64_t> base;
  wae_build_bloch(conn, nloc, work.elems, dim_red, dn.data(), dof_flag, n_class, cls, slotmap, base);
  std::vector<int32_t> e32(work.elems.begin(), work.elems.end());
  work.d_elems.upload(e32, h->stream);
  work.d_slotmap.upload(slotmap, h->stream);
  work.slotmap_built = true;
  const bool cx = kind == WAE_OP_BOUNDARY;
  DevBuf<double> d_all, d_c;
  d_all.alloc((size_t)std::max<int64_t>(base[n_class], 1) * (cx ? 2 : 1));
  CUDA_CHECK(cudaMemsetAsync(d_all.p, 0, d_all.n * sizeof(double), h->stream));
  if (kind != WAE_OP_MASS) upload_c(h, work, c, c_per_elem, d_c);
  PhaseTimer t(h, "assemble");
  if (kind == WAE_OP_MASS)
    wae_launch_assemble_atomic(h, work, 1, nullptr, 1, scale, d_all.p, nullptr);
  else if (kind == WAE_OP_STIFF)
    wae_launch_assemble_atomic(h, work, 2, d_c.p, c_per_elem, 1.0, nullptr, d_all.p);
  else
    wae_launch_assemble_atomic(h, work, 0, d_c.p, c_per_elem, scale, d_all.p, nullptr);
  t.stop();
  for (int k = 0; k < n_class; k++) {
    int pid = new_pattern(h);
    Pattern& P = *h->patterns[pid];
    P.dim = dim_red;
    P.nnz = cls[k].nnz;
    P.colptr.swap(cls[k].colptr);
    P.rowval.swap(cls[k].rowval);
    upload_pattern(h, P);
    int mid = new_matrix(h, pid, cx, P.nnz);
    if (P.nnz)
      CUDA_CHECK(cudaMemcpyAsync(h->mats[mid]->d_val.p, d_all.p + (size_t)base[k] * (cx ? 2 : 1), (size_t)P.nnz * (cx ? 2 : 1) * sizeof(double),
                                 cudaMemcpyDeviceToDevice, h->stream));
    pattern_ids[k] = pid;
    mat_ids[k] = mid;
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}

int32_t wae_mat_info(wae_ctx* h, int32_t mat_id, int32_t* pattern_id, int32_t* is_complex, int64_t* nnz) {
  WAE_API_BEGIN
  Matrix& M = h->mat(mat_id);
  if (pattern_id) *pattern_id = M.pattern;
  if (is_complex) *is_complex = M.is_complex;
  if (nnz) *nnz = h->pat(M.pattern).nnz;
  WAE_API_END
}

int32_t wae_mat_get(wae_ctx* h, int32_t mat_id, double* out) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  Matrix& M = h->mat(mat_id);
  int64_t nnz = h->pat(M.pattern).nnz;
  if (M.is_complex) {
    CUDA_CHECK(cudaMemcpyAsync(out, M.d_val.p, 2 * nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  } else {
    std::vector<double> tmp(nnz);
    CUDA_CHECK(cudaMemcpyAsync(tmp.data(), M.d_val.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int64_t k = 0; k < nnz; k++) {
      out[2 * k] = tmp[k];
      out[2 * k + 1] = 0.0;
    }
  }
  WAE_API_END
}

int32_t wae_mat_set(wae_ctx* h, int64_t dim, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                    int32_t* pattern_id, int32_t* mat_id) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  if (dim <= 0 || !colptr || !rowval || !nzval) WAE_THROW(WAE_E_INVALID, "bad matrix arguments");
  int pid = new_pattern(h);
  Pattern& P = *h->patterns[pid];
  P.dim = dim;
  P.colptr.resize(dim + 1);
  for (int64_t j = 0; j <= dim; j++) P.colptr[j] = colptr[j] - h->base;
  P.nnz = P.colptr[dim];
  if (P.colptr[0] != 0 || P.nnz < 0 || P.nnz >= ((int64_t)1 << 31)) WAE_THROW(WAE_E_INVALID, "bad colptr");
  for (int64_t j = 0; j < dim; j++)  // monotone and inside [0, nnz]: the row loop below indexes rowval with it
    if (P.colptr[j] < 0 || P.colptr[j] > P.colptr[j + 1] || P.colptr[j + 1] > P.nnz) WAE_THROW(WAE_E_INVALID, "colptr must be non-decreasing (column %lld)", (long long)j);
  P.rowval.resize(P.nnz);
  for (int64_t j = 0; j < dim; j++)
    for (int64_t k = P.colptr[j]; k < P.colptr[j + 1]; k++) {
      int64_t r = rowval[k] - h->base;
      if (r < 0 || r >= dim || (k > P.colptr[j] && r <= P.rowval[k - 1])) WAE_THROW(WAE_E_INVALID, "row indices must be sorted and unique inside columns");
      P.rowval[k] = (int32_t)r;
    }
  upload_pattern(h, P);
  int mid = new_matrix(h, pid, true, P.nnz);
  {  // exact (bitwise) complex symmetry check A(i,j) == A(j,i)
    bool sym = true;
    for (int64_t j = 0; j < dim && sym; j++)
      for (int64_t k = P.colptr[j]; k < P.colptr[j + 1]; k++) {
        int32_t i = P.rowval[k];
        if (i == j) continue;
        const int32_t* b = P.rowval.data() + P.colptr[i];
        const int32_t* e = P.rowval.data() + P.colptr[i + 1];
        const int32_t* it = std::lower_bound(b, e, (int32_t)j);
        if (it == e || *it != j) { sym = false; break; }
        int64_t q = it - P.rowval.data();
        if (nzval[2 * k] != nzval[2 * q] || nzval[2 * k + 1] != nzval[2 * q + 1]) { sym = false; break; }
      }
    h->mats[mid]->symmetric = sym;
  }
  CUDA_CHECK(cudaMemcpyAsync(h->mats[mid]->d_val.p, nzval, 2 * P.nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (pattern_id) *pattern_id = pid;
  if (mat_id) *mat_id = mid;
  WAE_API_END
}

int32_t wae_mat_free(wae_ctx* h, int32_t mat_id) {
  WAE_API_BEGIN
  h->mat(mat_id);
  for (size_t f = 0; f < h->fams.size(); f++)
    if (h->fams[f])
      for (int m : h->fams[f]->mats)
        if (m == mat_id) WAE_THROW(WAE_E_INVALID, "matrix %d is still a term of family %d (wae_family_free first)", mat_id, (int)f);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->mats[mat_id].reset();
  WAE_API_END
}

// A pattern can go once no live matrix or family refers to it (its pair program, slot map and device index arrays go with it).
int32_t wae_pattern_free(wae_ctx* h, int32_t pattern_id) {
  WAE_API_BEGIN
  h->pat(pattern_id);
  for (auto& m : h->mats)
    if (m && m->pattern == pattern_id) WAE_THROW(WAE_E_INVALID, "pattern %d is still used by a matrix", pattern_id);
  for (auto& f : h->fams)
    if (f && f->pattern == pattern_id) WAE_THROW(WAE_E_INVALID, "pattern %d is still used by a family", pattern_id);
  CUDA_CHECK(cudaSetDevice(h->device));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->patterns[pattern_id].reset();
  WAE_API_END
}

// ---- family ---------------------------------------------------------------------------------
// Releases the value slots, term maps and staging buffers of a family.  Its matrices stay (they may serve other families); a union
// pattern that was merged for this family alone is released too.  LU handles analysed for the family must be freed first.
int32_t wae_family_free(wae_ctx* h, int32_t fam_id) {
  WAE_API_BEGIN
  Family& F = h->fam(fam_id);
  for (auto& l : h->lus)
    if (l && wae_lu_family_of(*l) == fam_id) WAE_THROW(WAE_E_INVALID, "family %d still has an LU handle (wae_lu_free first)", fam_id);
  CUDA_CHECK(cudaSetDevice(h->device));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  const int pid = F.pattern;
  const bool merged = F.owns_pattern;
  h->fams[fam_id].reset();
  bool used = false;
  for (auto& m : h->mats) used |= m && m->pattern == pid;
  for (auto& f : h->fams) used |= f && f->pattern == pid;
  if (merged && !used) h->patterns[pid].reset();
  WAE_API_END
}

int32_t wae_family_create(wae_ctx* h, int32_t n_terms, const int32_t* mat_ids, int32_t* fam_id, int64_t* nnz_union) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  if (n_terms <= 0 || !mat_ids) WAE_THROW(WAE_E_INVALID, "empty family");
  std::unique_ptr<Family> Fp(new Family());
  Family& F = *Fp;
  F.n_terms = n_terms;
  F.mats.assign(mat_ids, mat_ids + n_terms);
  int64_t dim = -1;
  std::vector<int> pats;
  for (int t = 0; t < n_terms; t++) {
    Matrix& M = h->mat(mat_ids[t]);
    Pattern& P = h->pat(M.pattern);
    if (dim < 0) dim = P.dim;
    if (P.dim != dim) WAE_THROW(WAE_E_INVALID, "terms have different dimensions");
    if (std::find(pats.begin(), pats.end(), M.pattern) == pats.end()) pats.push_back(M.pattern);
  }
  // union pattern: reuse the largest pattern if it contains all others, else merge column-wise
  int big = pats[0];
  for (int p : pats)
    if (h->pat(p).nnz > h->pat(big).nnz) big = p;
  auto contains = [&](const Pattern& A, const Pattern& B) {  // B subset of A ?
    for (int64_t j = 0; j < dim; j++) {
      const int32_t *a = A.rowval.data() + A.colptr[j], *ae = A.rowval.data() + A.colptr[j + 1];
      const int32_t *b = B.rowval.data() + B.colptr[j], *be = B.rowval.data() + B.colptr[j + 1];
      if (!std::includes(a, ae, b, be)) return false;
    }
    return true;
  };
  bool all_in = true;
  for (int p : pats)
    if (p != big && !contains(h->pat(big), h->pat(p))) all_in = false;
  int upid;
  if (all_in) {
    upid = big;
  } else {
    upid = new_pattern(h);
    F.owns_pattern = true;
    Pattern& U = *h->patterns[upid];
    U.dim = dim;
    U.colptr.assign(dim + 1, 0);
    // column-wise union of the term patterns, two passes (count, fill), both parallel over contiguous column ranges
    const unsigned nthr = dim < 4096 ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    for (int pass = 0; pass < 2; pass++) {
      std::vector<std::thread> th;
      for (unsigned q = 0; q < nthr; q++)
        th.emplace_back([&, q, pass]() {
          std::vector<int32_t> buf, tmp;
          for (int64_t j = dim * q / nthr; j < dim * (q + 1) / nthr; j++) {
            buf.clear();
            for (int p : pats) {
              const Pattern& A = h->pat(p);
              tmp.clear();
              std::set_union(buf.begin(), buf.end(), A.rowval.begin() + A.colptr[j], A.rowval.begin() + A.colptr[j + 1],
                             std::back_inserter(tmp));
              buf.swap(tmp);
            }
            if (pass == 0)
              U.colptr[j + 1] = (int64_t)buf.size();  // counts first, prefix sum below
            else
              std::copy(buf.begin(), buf.end(), U.rowval.begin() + U.colptr[j]);
          }
        });
      for (auto& x : th) x.join();
      if (pass == 0) {
        for (int64_t j = 0; j < dim; j++) U.colptr[j + 1] += U.colptr[j];
        U.nnz = U.colptr[dim];
        if (U.nnz >= ((int64_t)1 << 31)) WAE_THROW(WAE_E_INVALID, "union pattern too large");
        U.rowval.resize(U.nnz);
      }
    }
    upload_pattern(h, U);
  }
  F.pattern = upid;
  Pattern& U = h->pat(upid);
  F.identity.resize(n_terms);
  F.d_map.resize(n_terms);
  F.d_inv.resize(n_terms);
  F.map_of.assign(n_terms, -1);
  F.inv_of.assign(n_terms, -1);
  for (int t = 0; t < n_terms; t++) {
    Matrix& M = h->mat(mat_ids[t]);
    if (M.pattern == upid) {
      F.identity[t] = true;
      continue;
    }
    F.identity[t] = false;
    for (int t2 = 0; t2 < t; t2++)  // terms on the same pattern (M and K) share their maps
      if (!F.identity[t2] && h->mat(mat_ids[t2]).pattern == M.pattern) {
        F.map_of[t] = F.map_of[t2];
        F.inv_of[t] = F.inv_of[t2];
      }
    if (F.map_of[t] >= 0) continue;
    Pattern& A = h->pat(M.pattern);
    std::vector<int32_t> map(A.nnz);
    const unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto fill = [&](int64_t j0, int64_t j1) {
      for (int64_t j = j0; j < j1; j++) {
        const int32_t* ub = U.rowval.data() + U.colptr[j];
        const int32_t* ue = U.rowval.data() + U.colptr[j + 1];
        for (int64_t k = A.colptr[j]; k < A.colptr[j + 1]; k++) {
          ub = std::lower_bound(ub, ue, A.rowval[k]);
          map[k] = (int32_t)(ub - U.rowval.data());
        }
      }
    };
    if (A.nnz > (1 << 20)) {
      std::vector<std::thread> th;
      for (unsigned q = 0; q < nthr; q++) th.emplace_back(fill, dim * q / nthr, dim * (q + 1) / nthr);
      for (auto& x : th) x.join();
    } else
      fill(0, dim);
    F.d_map[t].upload(map, h->stream);
    F.map_of[t] = t;
    if (2 * A.nnz >= U.nnz) {  // a large term: inverse map for the fused gather pass of wae_combine
      std::vector<int32_t> inv(U.nnz, -1);
      for (int64_t k = 0; k < A.nnz; k++) inv[map[k]] = (int32_t)k;
      F.d_inv[t].upload(inv, h->stream);
      F.inv_of[t] = t;
    }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
  for (int s = 0; s < WAE_FAMILY_SLOTS; s++) F.slot[s].alloc(0);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->fams.emplace_back(std::move(Fp));
  if (fam_id) *fam_id = (int)h->fams.size() - 1;
  if (nnz_union) *nnz_union = U.nnz;
  WAE_API_END
}

int32_t wae_family_pattern_get(wae_ctx* h, int32_t fam_id, int64_t* colptr, int64_t* rowval) {
  if (!h) return WAE_E_INVALID;
  try {
    return wae_pattern_get(h, h->fam(fam_id).pattern, colptr, rowval);
  } catch (const WaeError& e) {
    h->err = e.msg;
    return e.code;
  }
}

int32_t wae_combine(wae_ctx* h, int32_t fam_id, const double* coeffs, int32_t slot) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  if (!coeffs) WAE_THROW(WAE_E_INVALID, "coeffs is NULL");
  PhaseTimer t(h, "combine");
  wae_combine_device(h, h->fam(fam_id), coeffs, slot);
  t.stop();
  WAE_API_END
}

int32_t wae_family_get(wae_ctx* h, int32_t fam_id, int32_t slot, double* out) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  Family& F = h->fam(fam_id);
  if (slot < 0 || slot >= WAE_FAMILY_SLOTS || !F.slot[slot].p) WAE_THROW(WAE_E_INVALID, "family slot %d is empty", slot);
  CUDA_CHECK(cudaMemcpyAsync(out, F.slot[slot].p, 2 * h->pat(F.pattern).nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}

int32_t wae_family_spmm(wae_ctx* h, int32_t fam_id, int32_t slot, int32_t trans, int32_t nrhs, const double* X, double* Y) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  Family& F = h->fam(fam_id);
  if (slot < 0 || slot >= WAE_FAMILY_SLOTS || !F.slot[slot].p) WAE_THROW(WAE_E_INVALID, "family slot %d is empty", slot);
  if (trans < 0 || trans > 2 || nrhs <= 0 || !X || !Y) WAE_THROW(WAE_E_INVALID, "bad spmm arguments");
  int64_t dim = h->pat(F.pattern).dim;
  DevBuf<double>&dX = F.d_io[0], &dY = F.d_io[1];
  dX.upload(X, 2 * dim * nrhs, h->stream);
  dY.reserve(2 * dim * nrhs);
  PhaseTimer t(h, "spmv");
  wae_spmm_device(h, F, slot, trans, nrhs, (const cplx*)dX.p, (cplx*)dY.p);
  t.stop();
  CUDA_CHECK(cudaMemcpyAsync(Y, dY.p, 2 * dim * nrhs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}


// Host-only diagnostic (no GPU, no context): build pattern + pair program of a tetrahedral mesh and replay the three passes of
// assemble_tet_pairs on the host with synthetic element entries; see include/wae_b200.h.
int32_t wae_pair_program_check(int32_t order, int64_t n_pts, const double* xyz, int64_t n_tet, const uint32_t* tets, int32_t slot_cap,
                               double* out) {
  try {
    const int nloc = order == 1 ? 4 : 10, nsym = nloc * (nloc + 1) / 2;
    int64_t dim = 0;
    for (int64_t k = 0; k < n_tet * nloc; k++) dim = std::max<int64_t>(dim, (int64_t)tets[k] + 1);
    Pattern P;
    P.elem_kind = 3;
    P.elems.resize(n_tet);
    std::iota(P.elems.begin(), P.elems.end(), 0);
    wae_build_pattern_from_elements(tets, nloc, P.elems, dim, P);
    GatherHost G;
    wae_build_gather(xyz, tets, nloc, P, slot_cap, G);
    auto val = [](int64_t e, int s) { return std::sin(1.0 + 0.37 * (double)e + 1.3 * s); };
    auto nz = [&](int32_t i, int32_t j) {
      return std::lower_bound(P.rowval.begin() + P.colptr[j], P.rowval.begin() + P.colptr[j + 1], i) - P.rowval.begin();
    };
    std::vector<double> ref(P.nnz, 0.0), got(P.nnz, 0.0);
    std::vector<int> written(P.nnz, 0);
    for (int64_t e = 0; e < n_tet; e++) {
      const uint32_t* d = tets + e * nloc;
      int s = 0;
      for (int a = 0; a < nloc; a++)
        for (int b = a; b < nloc; b++, s++) {
          ref[nz(d[a], d[b])] += val(e, s);
          if (a != b) ref[nz(d[b], d[a])] += val(e, s);
        }
    }
    int64_t bad = 0;
    const double unset = -1e300;
    for (int p = 0; p < G.n_patch; p++) {
      const int64_t* D = &G.desc[(size_t)p * 8];
      const int32_t* I = reinterpret_cast<const int32_t*>(D + 4);
      const int nt = I[0], nv = I[1], ng = I[2], nc = I[3];
      const uint8_t* B = G.blob.data() + D[0];
      if (D[0] % 16 || I[4] % 16 || (D[1] * 8) % 16) bad++;
      const uint16_t* lv = reinterpret_cast<const uint16_t*>(B);
      const int32_t* pt = reinterpret_cast<const int32_t*>(B + I[5]);
      const uint32_t* grp = reinterpret_cast<const uint32_t*>(B + I[6]);
      const uint8_t* cnt = B + I[7];
      const uint32_t* chunk = reinterpret_cast<const uint32_t*>(B + I[7] + 32 * (int64_t)ng);
      std::vector<double> slots(G.max_slots, unset);
      for (int t = 0; t < nt; t++) {  // element pass
        const int32_t e = pt[t];
        for (int a = 0; a < 4; a++)
          if (lv[4 * t + a] >= nv || G.gvtx[D[1] / 3 + lv[4 * t + a]] != tets[(int64_t)e * nloc + a]) bad++;
        const uint32_t* dp = &G.dest[((size_t)(D[2] + t / 32) * G.npk) * 32 + (t & 31)];
        for (int s = 0; s < nsym; s++) {
          const uint32_t w = dp[(s >> 1) * 32], sl = (s & 1) ? (w >> 16) : (w & 0xFFFFu);
          if (sl == 0xFFFFu) continue;
          if ((int)sl >= G.max_slots || slots[sl] != unset) { bad++; continue; }
          slots[sl] = val(e, s);
        }
      }
      for (int g = 0; g < ng; g++) {  // summation pass
        const int sb = (int)(grp[g] >> 8), niter = (int)(grp[g] & 255);
        for (int l = 0; l < 32; l++) {
          const int c = cnt[g * 32 + l];
          if (c > niter) bad++;
          double acc = 0.0;
          for (int k = 0; k < c; k++) {
            const double v = slots[sb + 32 * k + (l ^ (k & 7))];
            if (v == unset) bad++;
            acc += v;
          }
          if (c) slots[sb + l] = acc;
        }
      }
      for (int ch = 0; ch < nc; ch++) {  // store pass
        const uint32_t z0 = chunk[2 * ch], len = chunk[2 * ch + 1];
        if (len < 1 || len > 32 || (z0 >> 5) != ((z0 + len - 1) >> 5)) { bad++; continue; }
        for (uint32_t l = 0; l < len; l++) {
          got[z0 + l] = slots[G.res[(size_t)(D[3] + ch) * 32 + l]];
          written[z0 + l]++;
        }
      }
    }
    double err = 0.0;
    for (int64_t k = 0; k < P.nnz; k++) {
      err = std::max(err, std::fabs(got[k] - ref[k]));
      if (written[k] != 1) bad++;
    }
    out[0] = (double)P.nnz;
    out[1] = G.n_patch;
    out[2] = (double)G.n_staged;
    out[3] = (double)G.n_sources;
    out[4] = (double)G.n_pairs;
    out[5] = G.max_slots;
    out[6] = err;
    out[7] = (double)bad;
    return WAE_OK;
  } catch (...) {
    return WAE_E_INVALID;
  }
}
}  // extern "C"
