// Discrete-adjoint shape sensitivity (src/shape_sensitivity.jl:16-141) as ONE batched local re-assembly kernel.
//
// The reference moves every surface point by +-h along x, y, z, calls discretize() six times on a copy of the mesh whose
// domains are cut down to the simplices touching the point, and forms  sens = -v_adj^H (L_right(w0) - L_left(w0)) / (2h) v.
// Only the element matrices of the simplices touching the point differ between the two discretisations, so the same number is
//     - sum_{e touches p}  sum_ij conj(v_adj_i) (E_e(x+h) - E_e(x-h))_ij v_j / (2h)
// and no matrix is ever formed: one thread per (surface point, coordinate) walks the point's simplex list (CSR), evaluates the
// first-order element matrices (FEM.jl:435-441,704-710,1745-1765,2283-2311,2429-2431,2442-2448) at the two perturbed positions
// and contracts the difference with the two eigenvectors.  One launch per descriptor entry; the flame term keeps the reference's
// behaviour of normalising with the volume of the *reduced* flame domain (shape_sensitivity.jl:50-69 + Helmholtz.jl:325).
// The reference's discretize() call takes no `order`, i.e. this path is first-order (4-node tetrahedra, 3-node triangles) only.
// Unit-cell meshes (shape_sensitivity.jl:84-118, b = :b with params[b] = 1): points move along their local cylindrical basis,
// a point of the Bloch reference plane moves together with its image, DOFs are folded as blochify does (Bloch.jl:4-66) and an
// entry that couples an image DOF j with a plain DOF i carries exp(+i 2 pi / DOS) (the other way round: its conjugate).  Entries
// that touch an axis DOF are dropped: at b = 1 their class scalar delta(b) vanishes (Helmholtz.jl:95-100); the diagonal penalty
// term (1 - delta(b)) D (Helmholtz.jl:551-568) acts on the axis DOFs only, where both eigenvectors vanish, and is not evaluated.
//
// The per-thread function is __host__ __device__: wae_shape_sens_check replays it on the host (CPU tests, no GPU).
#include <cuda_runtime.h>

#include <cmath>

#include "wae_internal.h"

#define WAE_HD __host__ __device__ __forceinline__

struct SensArgs {
  const double* xyz;       // 3 x n_pts
  const uint32_t* conn;    // tetrahedra (kinds mass / stiff / flame) or triangles (boundary), `stride` DOFs per element, vertices first
  int stride;
  const int64_t* points;   // n_sp moved points
  const int64_t* partner;  // n_sp or NULL: second point moved at the same time (image of a Bloch-plane point), -1 = none
  int cylindrical;         // 0: move along x, y, z; 1: along the local cylindrical basis (r, phi, z) of every moved point
  const int32_t* dof_new;  // NULL = identity, else folded DOF of every mesh point (Bloch unit cell)
  const uint8_t* dof_flag; // bit 0: image of the Bloch plane, bit 1: axis DOF
  double ph_re, ph_im;     // exp(+i b 2 pi / DOS) at b = 1
  const int64_t* ptr;      // n_sp + 1: simplex list of every moved point
  const int64_t* elems;    // element ids
  const double* c;         // speed of sound per list entry (c_per_elem values each), unused for mass / flame
  int c_per_elem;
  const cplx* v;           // direct eigenvector (normalised by the caller)
  const cplx* va;          // adjoint eigenvector (normalised by the caller)
  double step;             // h
  double coef_re, coef_im; // scalar of the term at w0
  int kind;
  int64_t ref_tet;         // flame: reference tetrahedron
  double n_ref[3];
  double nl;               // flame: (gamma-1)/rho*nglobal  (divided by the reduced flame volume in the kernel)
};

WAE_HD cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
WAE_HD cplx cconj(cplx a) { return make_double2(a.x, -a.y); }

// CooTrafo of a tetrahedron (FEM.jl:2-21): gradients of the barycentric coordinates and |det J|
WAE_HD void sens_tet_geom(const double X[4][3], double G[4][3], double& adet) {
  double a[3][3];
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < 3; r++) a[r][k] = X[k][r] - X[3][r];
  double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  double id = 1.0 / det;
  G[0][0] = c00 * id;
  G[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  G[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  G[1][0] = c01 * id;
  G[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  G[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  G[2][0] = c02 * id;
  G[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  G[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
  for (int d = 0; d < 3; d++) G[3][d] = -(G[0][d] + G[1][d] + G[2][d]);
  adet = fabs(det);
}

// twice the area of a triangle = |det| of its CooTrafo (FEM.jl:14-17)
WAE_HD double sens_tri_adet(const double X[3][3]) {
  double e1[3], e2[3];
  for (int r = 0; r < 3; r++) {
    e1[r] = X[0][r] - X[2][r];
    e2[r] = X[1][r] - X[2][r];
  }
  double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
  return sqrt(nx * nx + ny * ny + nz * nz);
}

// Positions of one vertex in the two perturbed meshes, exactly as the reference forms them: x + h d, then (x + h d) - 2 h d with
// d = unit vector crd (shape_sensitivity.jl:107,116) or column crd of get_cylindrics(x) (:96-98,112-114,361-370).
// Returns false if the vertex is not one of the moved points.
WAE_HD bool sens_move(const SensArgs& a, int64_t vtx, int64_t p, int64_t q, int crd, const double x[3], double xr[3], double xl[3]) {
  if (vtx != p && vtx != q) return false;
  double d[3] = {0.0, 0.0, 0.0};
  if (!a.cylindrical) {
    d[crd] = 1.0;
  } else if (crd == 2) {
    d[2] = 1.0;
  } else {
    const double nr = sqrt(x[0] * x[0] + x[1] * x[1]);
    const double rx = x[0] / nr, ry = x[1] / nr;
    if (crd == 0) {
      d[0] = rx;
      d[1] = ry;
    } else {  // e_z x e_r
      d[0] = -ry;
      d[1] = rx;
    }
  }
  for (int r = 0; r < 3; r++) {
    xr[r] = x[r] + a.step * d[r];
    xl[r] = xr[r] - 2 * a.step * d[r];
  }
  return true;
}

// conj(va_i) v_j of the element's vertices with the Bloch folding, class phases and the axis rule applied
template <int N>
WAE_HD void sens_weights(const SensArgs& a, const uint32_t* d, cplx w[N][N]) {
  cplx vv[N], vc[N];
  int img[N], axis[N];
  for (int k = 0; k < N; k++) {
    const int64_t dof = a.dof_new ? a.dof_new[d[k]] : (int64_t)d[k];
    const int fl = a.dof_flag ? a.dof_flag[d[k]] : 0;
    img[k] = fl & 1;
    axis[k] = (fl >> 1) & 1;
    vv[k] = a.v[dof];
    vc[k] = cconj(a.va[dof]);
  }
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      cplx t = cmul(vc[i], vv[j]);
      if (img[j] && !img[i]) t = cmul(t, make_double2(a.ph_re, a.ph_im));
      if (img[i] && !img[j]) t = cmul(t, make_double2(a.ph_re, -a.ph_im));
      if (axis[i] || axis[j]) t = make_double2(0.0, 0.0);
      w[i][j] = t;
    }
}

// sum_ij conj(va_i) E_ij v_j for the element matrix of one tetrahedron; kind mass: E = |det| (1+delta_ij)/120,
// stiff: E = -cfac |det| grad(l_i).grad(l_j)
WAE_HD cplx sens_tet_form(const double X[4][3], int kind, double cfac, const cplx w[4][4]) {
  double G[4][3], adet;
  sens_tet_geom(X, G, adet);
  double re = 0, im = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double e;
      if (kind == WAE_SENS_MASS)
        e = adet * (i == j ? 2.0 : 1.0) / 120.0;
      else
        e = -cfac * adet * (G[i][0] * G[j][0] + G[i][1] * G[j][1] + G[i][2] * G[j][2]);
      re += e * w[i][j].x;
      im += e * w[i][j].y;
    }
  return make_double2(re, im);
}

WAE_HD cplx sens_tri_form(const double X[3][3], const double* c, int c_per_elem, const cplx w[3][3]) {
  double adet = sens_tri_adet(X);
  double re = 0, im = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double m;
      if (c_per_elem == 1) {
        m = c[0] * (i == j ? 2.0 : 1.0) / 24.0;
      } else {  // int c phi_i phi_j, c linear (FEM.jl:469-483): 1/20 (i=j=k), 1/60 (two indices equal), 1/120 (all distinct)
        m = 0;
        for (int k = 0; k < 3; k++) {
          int eq = (i == j) + (i == k) + (j == k);
          m += c[k] * (eq == 3 ? 1.0 / 20.0 : eq == 1 ? 1.0 / 60.0 : 1.0 / 120.0);
        }
      }
      double e = m * adet;
      re += e * w[i][j].x;
      im += e * w[i][j].y;
    }
  return make_double2(im, -re);  // times -i (Helmholtz.jl:459)
}

// contribution of one descriptor entry to sens[crd, s]
WAE_HD cplx shape_sens_item(const SensArgs& a, int64_t s, int crd) {
  const int64_t p = a.points[s];
  const int64_t q = a.partner ? a.partner[s] : -1;
  cplx acc = make_double2(0.0, 0.0);
  if (a.kind == WAE_SENS_MASS || a.kind == WAE_SENS_STIFF) {
    for (int64_t it = a.ptr[s]; it < a.ptr[s + 1]; it++) {
      const uint32_t* d = a.conn + (size_t)a.elems[it] * a.stride;
      bool moved = false;
      double Xr[4][3], Xl[4][3];
      for (int k = 0; k < 4; k++) {
        double x[3];
        for (int r = 0; r < 3; r++) x[r] = Xr[k][r] = Xl[k][r] = a.xyz[3 * (size_t)d[k] + r];
        moved |= sens_move(a, (int64_t)d[k], p, q, crd, x, Xr[k], Xl[k]);
      }
      if (!moved) continue;  // the element contains no moved point: identical in both meshes
      cplx w[4][4];
      sens_weights<4>(a, d, w);
      double cfac = 0;
      if (a.kind == WAE_SENS_STIFF) {
        if (a.c_per_elem == 1) {
          cfac = a.c[it] * a.c[it] / 6.0;
        } else {  // FEM.jl:2283-2311: sum_{k<=l} c_k c_l / 60
          const double* cv = a.c + 4 * it;
          for (int k = 0; k < 4; k++)
            for (int l = k; l < 4; l++) cfac += cv[k] * cv[l];
          cfac /= 60.0;
        }
      }
      cplx fr = sens_tet_form(Xr, a.kind, cfac, w);
      cplx fl = sens_tet_form(Xl, a.kind, cfac, w);
      acc.x += fr.x - fl.x;
      acc.y += fr.y - fl.y;
    }
  } else if (a.kind == WAE_SENS_BOUNDARY) {
    for (int64_t it = a.ptr[s]; it < a.ptr[s + 1]; it++) {
      const uint32_t* d = a.conn + (size_t)a.elems[it] * a.stride;
      bool moved = false;
      double Xr[3][3], Xl[3][3];
      for (int k = 0; k < 3; k++) {
        double x[3];
        for (int r = 0; r < 3; r++) x[r] = Xr[k][r] = Xl[k][r] = a.xyz[3 * (size_t)d[k] + r];
        moved |= sens_move(a, (int64_t)d[k], p, q, crd, x, Xr[k], Xl[k]);
      }
      if (!moved) continue;
      cplx w[3][3];
      sens_weights<3>(a, d, w);
      const double* cv = a.c + (size_t)a.c_per_elem * it;
      cplx fr = sens_tri_form(Xr, cv, a.c_per_elem, w);
      cplx fl = sens_tri_form(Xl, cv, a.c_per_elem, w);
      acc.x += fr.x - fl.x;
      acc.y += fr.y - fl.y;
    }
  } else {  // flame: Q = S (x) G with S over the flame tetrahedra touching the point, G = -nl/V_reduced * grad(l_j).n_ref on ref_tet
    if (a.ptr[s] == a.ptr[s + 1]) return acc;  // reduced flame domain empty: no Q entries at all
    cplx A[2] = {make_double2(0, 0), make_double2(0, 0)};
    double V[2] = {0, 0};
    for (int64_t it = a.ptr[s]; it < a.ptr[s + 1]; it++) {
      const uint32_t* d = a.conn + (size_t)a.elems[it] * a.stride;
      double X[2][4][3];
      cplx sv = make_double2(0, 0);
      for (int k = 0; k < 4; k++) {
        double x[3];
        for (int r = 0; r < 3; r++) x[r] = X[0][k][r] = X[1][k][r] = a.xyz[3 * (size_t)d[k] + r];
        sens_move(a, (int64_t)d[k], p, q, crd, x, X[0][k], X[1][k]);
        cplx t = cconj(a.va[d[k]]);
        sv.x += t.x;
        sv.y += t.y;
      }
      for (int sgn = 0; sgn < 2; sgn++) {
        double G[4][3], adet;
        sens_tet_geom(X[sgn], G, adet);
        A[sgn].x += adet / 24.0 * sv.x;  // S_i = |det|/24 (FEM.jl:2429-2431)
        A[sgn].y += adet / 24.0 * sv.y;
        V[sgn] += adet / 6.0;            // compute_size! of the reduced domain (Meshutils.jl:757-780)
      }
    }
    const uint32_t* d = a.conn + (size_t)a.ref_tet * a.stride;
    double X[2][4][3];
    for (int k = 0; k < 4; k++) {
      double x[3];
      for (int r = 0; r < 3; r++) x[r] = X[0][k][r] = X[1][k][r] = a.xyz[3 * (size_t)d[k] + r];
      sens_move(a, (int64_t)d[k], p, q, crd, x, X[0][k], X[1][k]);
    }
    cplx qv[2];
    for (int sgn = 0; sgn < 2; sgn++) {
      double G[4][3], adet;
      sens_tet_geom(X[sgn], G, adet);
      cplx B = make_double2(0, 0);
      for (int j = 0; j < 4; j++) {  // grad(l_j).n_ref (FEM.jl:2442-2448)
        double g = G[j][0] * a.n_ref[0] + G[j][1] * a.n_ref[1] + G[j][2] * a.n_ref[2];
        B.x += g * a.v[d[j]].x;
        B.y += g * a.v[d[j]].y;
      }
      double f = -a.nl / V[sgn];
      qv[sgn] = cmul(A[sgn], B);
      qv[sgn].x *= f;
      qv[sgn].y *= f;
    }
    acc.x = qv[0].x - qv[1].x;
    acc.y = qv[0].y - qv[1].y;
  }
  cplx r = cmul(make_double2(a.coef_re, a.coef_im), acc);
  const double f = -1.0 / (2 * a.step);
  return make_double2(r.x * f, r.y * f);
}

__global__ void __launch_bounds__(128) shape_sens_kernel(SensArgs a, int64_t n_sp, cplx* __restrict__ sens) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n_sp) return;
  cplx r = shape_sens_item(a, t / 3, (int)(t % 3));
  sens[t].x += r.x;  // sens is 3 x n_sp column-major: entry (crd, s) at 3 s + crd == t
  sens[t].y += r.y;
}

// ---- state of one begin / add* / end sequence (owned by the context) ------------------------------------------------------
struct WaeShapeSens {
  int64_t n_sp = 0;
  double step = 0;
  DevBuf<int64_t> d_points, d_partner, d_ptr, d_elems;
  DevBuf<double> d_c;
  DevBuf<cplx> d_v, d_va, d_sens;
  DevBuf<int32_t> d_dof_new;
  DevBuf<uint8_t> d_dof_flag;
  bool has_partner = false, folded = false;
  int cylindrical = 0;
  double ph_re = 1.0, ph_im = 0.0;
  bool open = false;
};

static void check_lists(int64_t n_sp, const int64_t* ptr, const int64_t* elems, int64_t n_elem_total, int base) {
  if (!ptr || ptr[0] != 0) WAE_THROW(WAE_E_INVALID, "shape sensitivity: ptr[0] must be 0");
  for (int64_t s = 0; s < n_sp; s++)
    if (ptr[s + 1] < ptr[s]) WAE_THROW(WAE_E_INVALID, "shape sensitivity: ptr must be non-decreasing");
  if (ptr[n_sp] > 0 && !elems) WAE_THROW(WAE_E_INVALID, "shape sensitivity: missing element list");
  for (int64_t k = 0; k < ptr[n_sp]; k++)
    if (elems[k] - base < 0 || elems[k] - base >= n_elem_total) WAE_THROW(WAE_E_INVALID, "shape sensitivity: element id %lld out of range", (long long)elems[k]);
}

extern "C" {

int32_t wae_shape_sens_begin(wae_ctx* h, int64_t n_sp, const int64_t* points, const int64_t* partner, double step, int32_t cylindrical,
                             int64_t vdim, const double* v, const double* v_adj, const int64_t* dof_new, const uint8_t* dof_flag,
                             const double* phase) {
  if (!h) return WAE_E_INVALID;
  try {
    CUDA_CHECK(cudaSetDevice(h->device));
    if (h->order != 1 || h->dim != h->n_pts) WAE_THROW(WAE_E_INVALID, "shape sensitivity needs a first-order mesh (wae_mesh_set with order 1): the reference's path is :lin only");
    if (n_sp <= 0 || !points || !v || !v_adj || !(step > 0)) WAE_THROW(WAE_E_INVALID, "shape sensitivity: bad arguments");
    if ((dof_new != nullptr) != (dof_flag != nullptr) || (dof_new && !phase)) WAE_THROW(WAE_E_INVALID, "shape sensitivity: dof_new, dof_flag and phase go together");
    if (vdim != (dof_new ? vdim : h->dim) || vdim <= 0 || vdim > h->dim) WAE_THROW(WAE_E_INVALID, "shape sensitivity: eigenvector length does not match the (folded) dimension");
    if (!h->shape) h->shape = std::make_shared<WaeShapeSens>();
    WaeShapeSens& S = *h->shape;
    S.open = false;
    std::vector<int64_t> pts(n_sp), par(n_sp, -1);
    for (int64_t s = 0; s < n_sp; s++) {
      pts[s] = points[s] - h->base;
      if (pts[s] < 0 || pts[s] >= h->n_pts) WAE_THROW(WAE_E_INVALID, "shape sensitivity: point %lld out of range", (long long)points[s]);
      if (partner && partner[s] - h->base >= 0) {
        par[s] = partner[s] - h->base;
        if (par[s] >= h->n_pts) WAE_THROW(WAE_E_INVALID, "shape sensitivity: partner point %lld out of range", (long long)partner[s]);
      }
    }
    S.folded = dof_new != nullptr;
    if (S.folded) {
      std::vector<int32_t> dn(h->n_pts);
      for (int64_t i = 0; i < h->n_pts; i++) {
        const int64_t d = dof_new[i] - h->base;
        if (d < 0 || d >= vdim) WAE_THROW(WAE_E_INVALID, "shape sensitivity: folded DOF %lld out of range", (long long)dof_new[i]);
        dn[i] = (int32_t)d;
      }
      S.d_dof_new.upload(dn, h->stream);
      S.d_dof_flag.upload(dof_flag, h->n_pts, h->stream);
      S.ph_re = phase[0];
      S.ph_im = phase[1];
      CUDA_CHECK(cudaStreamSynchronize(h->stream));  // dn is a local
    }
    S.n_sp = n_sp;
    S.step = step;
    S.cylindrical = cylindrical ? 1 : 0;
    S.has_partner = partner != nullptr;
    S.d_points.upload(pts, h->stream);
    S.d_partner.upload(par, h->stream);
    S.d_v.upload((const cplx*)v, vdim, h->stream);
    S.d_va.upload((const cplx*)v_adj, vdim, h->stream);
    S.d_sens.reserve(3 * n_sp);
    CUDA_CHECK(cudaMemsetAsync(S.d_sens.p, 0, 3 * n_sp * sizeof(cplx), h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));  // pts, par are locals
    S.open = true;
    return WAE_OK;
  } catch (const WaeError& e) {
    h->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    return WAE_E_INVALID;
  }
}

int32_t wae_shape_sens_add(wae_ctx* h, int32_t kind, const int64_t* ptr, const int64_t* elems, const double* c, int32_t c_per_elem,
                           const double* coef, int64_t ref_tet, const double* n_ref, double nl) {
  if (!h) return WAE_E_INVALID;
  try {
    CUDA_CHECK(cudaSetDevice(h->device));
    if (!h->shape || !h->shape->open) WAE_THROW(WAE_E_INVALID, "wae_shape_sens_add without wae_shape_sens_begin");
    WaeShapeSens& S = *h->shape;
    if (kind < WAE_SENS_MASS || kind > WAE_SENS_FLAME || !coef) WAE_THROW(WAE_E_INVALID, "shape sensitivity: unknown term kind");
    const bool tri = kind == WAE_SENS_BOUNDARY;
    check_lists(S.n_sp, ptr, elems, tri ? h->n_tri : h->n_tet, h->base);
    const int64_t n_items = ptr[S.n_sp];
    const int want_c = kind == WAE_SENS_STIFF ? 4 : 3;
    if (kind == WAE_SENS_STIFF || tri) {
      if (n_items && (!c || (c_per_elem != 1 && c_per_elem != want_c))) WAE_THROW(WAE_E_INVALID, "shape sensitivity: c must hold 1 or %d values per list entry", want_c);
    }
    SensArgs a{};
    if (kind == WAE_SENS_FLAME) {
      if (S.folded) WAE_THROW(WAE_E_INVALID, "shape sensitivity: flame terms on Bloch-folded meshes are not supported");
      if (!n_ref || ref_tet - h->base < 0 || ref_tet - h->base >= h->n_tet) WAE_THROW(WAE_E_INVALID, "shape sensitivity: bad flame reference");
      a.ref_tet = ref_tet - h->base;
      for (int r = 0; r < 3; r++) a.n_ref[r] = n_ref[r];
      a.nl = nl;
    }
    if (n_items == 0) return WAE_OK;
    std::vector<int64_t> el(n_items);
    for (int64_t k = 0; k < n_items; k++) el[k] = elems[k] - h->base;
    S.d_ptr.upload(ptr, S.n_sp + 1, h->stream);
    S.d_elems.upload(el, h->stream);
    if (kind == WAE_SENS_STIFF || tri) S.d_c.upload(c, (size_t)n_items * c_per_elem, h->stream);
    a.xyz = h->d_xyz.p;
    a.conn = tri ? h->d_tris.p : h->d_tets.p;
    a.stride = tri ? h->nloc3 : h->nloc;
    a.points = S.d_points.p;
    a.partner = S.has_partner ? S.d_partner.p : nullptr;
    a.cylindrical = S.cylindrical;
    a.dof_new = S.folded ? S.d_dof_new.p : nullptr;
    a.dof_flag = S.folded ? S.d_dof_flag.p : nullptr;
    a.ph_re = S.ph_re;
    a.ph_im = S.ph_im;
    a.ptr = S.d_ptr.p;
    a.elems = S.d_elems.p;
    a.c = S.d_c.p;
    a.c_per_elem = c_per_elem;
    a.v = S.d_v.p;
    a.va = S.d_va.p;
    a.step = S.step;
    a.coef_re = coef[0];
    a.coef_im = coef[1];
    a.kind = kind;
    PhaseTimer tm(h, "shape_sens");
    const int64_t nthr = 3 * S.n_sp;
    shape_sens_kernel<<<(unsigned)((nthr + 127) / 128), 128, 0, h->stream>>>(a, S.n_sp, S.d_sens.p);
    CUDA_CHECK(cudaGetLastError());
    h->launches++;
    tm.stop();  // synchronises: the host staging vectors may go out of scope
    return WAE_OK;
  } catch (const WaeError& e) {
    h->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    return WAE_E_INVALID;
  }
}

int32_t wae_shape_sens_end(wae_ctx* h, double* sens) {
  if (!h) return WAE_E_INVALID;
  try {
    CUDA_CHECK(cudaSetDevice(h->device));
    if (!h->shape || !h->shape->open || !sens) WAE_THROW(WAE_E_INVALID, "wae_shape_sens_end without wae_shape_sens_begin");
    WaeShapeSens& S = *h->shape;
    CUDA_CHECK(cudaMemcpyAsync(sens, S.d_sens.p, 3 * S.n_sp * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    S.open = false;
    return WAE_OK;
  } catch (const WaeError& e) {
    h->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    return WAE_E_INVALID;
  }
}

// Host-only diagnostic (no GPU, no context): the per-thread function of shape_sens_kernel evaluated in a plain loop, 0-based
// indices, first-order connectivity (4 x n_tet, 3 x n_tri); accumulates into sens (3 x n_sp complex).  See include/wae_b200.h.
int32_t wae_shape_sens_check(int64_t n_pts, const double* xyz, int64_t n_tet, const uint32_t* tets, int64_t n_tri, const uint32_t* tris,
                             int64_t n_sp, const int64_t* points, const int64_t* partner, double step, int32_t cylindrical, const double* v,
                             const double* v_adj, const int32_t* dof_new, const uint8_t* dof_flag, const double* phase, int32_t kind,
                             const int64_t* ptr, const int64_t* elems, const double* c, int32_t c_per_elem, const double* coef,
                             int64_t ref_tet, const double* n_ref, double nl, double* sens) {
  try {
    if (n_pts <= 0 || !xyz || !points || !v || !v_adj || !coef || !sens || kind < WAE_SENS_MASS || kind > WAE_SENS_FLAME) return WAE_E_INVALID;
    const bool tri = kind == WAE_SENS_BOUNDARY;
    check_lists(n_sp, ptr, elems, tri ? n_tri : n_tet, 0);
    SensArgs a{};
    a.xyz = xyz;
    a.conn = tri ? tris : tets;
    a.stride = tri ? 3 : 4;
    a.points = points;
    a.partner = partner;
    a.cylindrical = cylindrical ? 1 : 0;
    a.dof_new = dof_new;
    a.dof_flag = dof_new ? dof_flag : nullptr;
    a.ph_re = phase ? phase[0] : 1.0;
    a.ph_im = phase ? phase[1] : 0.0;
    a.ptr = ptr;
    a.elems = elems;
    a.c = c;
    a.c_per_elem = c_per_elem;
    a.v = (const cplx*)v;
    a.va = (const cplx*)v_adj;
    a.step = step;
    a.coef_re = coef[0];
    a.coef_im = coef[1];
    a.kind = kind;
    if (kind == WAE_SENS_FLAME) {
      if (!n_ref || ref_tet < 0 || ref_tet >= n_tet) return WAE_E_INVALID;
      a.ref_tet = ref_tet;
      for (int r = 0; r < 3; r++) a.n_ref[r] = n_ref[r];
      a.nl = nl;
    }
    cplx* out = (cplx*)sens;
    for (int64_t t = 0; t < 3 * n_sp; t++) {
      cplx r = shape_sens_item(a, t / 3, (int)(t % 3));
      out[t].x += r.x;
      out[t].y += r.y;
    }
    return WAE_OK;
  } catch (...) {
    return WAE_E_INVALID;
  }
}

}  // extern "C"
