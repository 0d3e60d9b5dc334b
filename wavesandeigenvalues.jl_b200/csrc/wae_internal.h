// Internal data structures of libwae_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/wae_b200.h"

typedef double2 cplx;  // interleaved complex fp64 (x = re, y = im)

struct WaeError {
  int code;
  std::string msg;
};

#define WAE_THROW(code, ...)                         \
  do {                                               \
    char _b[512];                                    \
    snprintf(_b, sizeof(_b), __VA_ARGS__);           \
    throw WaeError{(code), std::string(_b)};         \
  } while (0)

#define CUDA_CHECK(x)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (x);                                                                    \
    if (_e != cudaSuccess)                                                                   \
      WAE_THROW(WAE_E_CUDA, "%s failed at %s:%d: %s", #x, __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

// RAII device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) return;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e != cudaSuccess) WAE_THROW(WAE_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    n = count;
  }
  void reserve(size_t count) {
    if (n < count) alloc(count);
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (n < count) alloc(count);
    if (count) CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& v, cudaStream_t s) { upload(v.data(), v.size(), s); }
};

// CSC pattern, 0-based, row indices sorted inside columns.
struct Pattern {
  int64_t dim = 0, nnz = 0;
  std::vector<int64_t> colptr;  // dim+1
  std::vector<int32_t> rowval;  // nnz
  DevBuf<int64_t> d_colptr;
  DevBuf<int32_t> d_rowval;
  // element list the pattern was built from (kind 3 tets / 2 tris / 0 user)
  int elem_kind = 0;
  std::vector<int64_t> elems;   // 0-based element ids
  DevBuf<int32_t> d_elems;
  // owner-computes pair program (assembly): built lazily, see assembly_symbolic.cpp
  struct Gather {
    int n_patch = 0, max_slots = 0, max_blob = 0, max_nv = 0;
    DevBuf<int64_t> d_desc;               // 64-byte patch descriptors
    DevBuf<uint8_t> d_blob;               // per patch: local vertex numbers, element ids, group words, counts, store chunks
    DevBuf<uint32_t> d_gvtx;              // patch-local vertex -> mesh vertex
    DevBuf<double> d_pxyz;                // vertex coordinates in patch order (3 per local vertex), refreshed when the points change
    DevBuf<uint32_t> d_dest;              // per 32-element block: npk x 32 words = two 16-bit shared-memory slots per (element, local pair)
    DevBuf<uint16_t> d_res;               // per chunk lane: slot holding the sum of that nonzero
    int64_t n_pairs = 0, n_sources = 0, n_staged = 0, n_pv = 0;
    uint64_t xyz_version = 0;             // h->xyz_version the pxyz copy was gathered from
    bool built = false;
  } gather;
  // star program (assembly generation 3): built lazily, see assembly_star_symbolic.cpp
  struct Star {
    int n_patch = 0, max_nt = 0, max_nv = 0, max_rows = 0, max_ng = 0, max_a = 0, max_b = 0;
    int64_t max_smem = 0;
    DevBuf<int64_t> d_desc;               // 64-byte patch descriptors
    DevBuf<uint8_t> d_blob;               // blob A per patch: local vertex numbers, element ids, group headers, lane counts, star sources
    DevBuf<uint8_t> d_blobB;              // blob B per patch: store chunks, codes (group << 9 | lane << 4 | role per owned nonzero)
    DevBuf<uint32_t> d_gvtx;              // patch-local vertex -> mesh vertex
    DevBuf<double> d_pxyz;                // vertex coordinates in patch order
    int64_t n_staged = 0, n_entities = 0, n_sources = 0, n_chunks = 0, n_pv = 0, program_bytes = 0;
    uint64_t xyz_version = 0;
    int64_t budget = 0;                   // shared-memory budget the program was cut for
    bool built = false, failed = false;
  } star;
  // scatter map for the atomic (first-generation) kernels: n_loc^2 slots per element
  DevBuf<int32_t> d_slotmap;
  bool slotmap_built = false;
};

struct Matrix {
  int pattern = -1;
  bool is_complex = false;
  DevBuf<double> d_val;  // nnz (real) or 2*nnz (complex, interleaved)
  // structure hints for the solver: complex-symmetric operators (M, K, C by construction) allow an LDL^T-type
  // factorisation; a flame operator is the rank-1 matrix S (x) G and is handled by the Sherman-Morrison-Woodbury formula
  bool symmetric = false;
  bool rank1 = false;
  DevBuf<double> d_r1_S, d_r1_G;          // rank-1 factors (real): S over r1_rows, G over r1_cols
  std::vector<int32_t> r1_rows, r1_cols;  // DOF ids (cols in the order of d_r1_G)
  DevBuf<int32_t> d_r1_rows, d_r1_cols;
};

struct Family {
  int n_terms = 0;
  std::vector<int> mats;
  int pattern = -1;                        // union pattern id
  bool owns_pattern = false;               // the union pattern was merged for this family (else it is a term's own pattern)
  std::vector<bool> identity;              // term pattern == union pattern
  std::vector<DevBuf<int32_t>> d_map;      // term nz -> union nz (empty if identity); terms on the same pattern share the map of the first one
  std::vector<int> map_of;                 // term -> index into d_map (the term that owns the map)
  std::vector<DevBuf<int32_t>> d_inv;      // large sub-pattern terms: union nz -> term nz (-1: absent), so that they join the fused gather pass
  std::vector<int> inv_of;                 // term -> index into d_inv (owner term) or -1
  DevBuf<double> slot[WAE_FAMILY_SLOTS];   // complex values, 2*nnz doubles each
  std::vector<double> slot_coeffs[WAE_FAMILY_SLOTS];  // term coefficients of the last wae_combine into each slot
  DevBuf<double> d_io[2];                  // cached staging buffers of wae_family_spmm (grow-only)
  DevBuf<int32_t> d_tr_perm;               // CSC->CSR permutation for transposed SpMM (lazy)
  DevBuf<int64_t> d_rowptr;
  DevBuf<int32_t> d_colidx;
  bool tr_built = false;
};

struct LuSolver;  // lu_symbolic.h
int wae_lu_family_of(const LuSolver& S);  // lu_api.cu (LuSolver is incomplete here)
struct WaeShapeSens;  // shape_sens.cu

struct wae_ctx {
  int device = 0;
  int base = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  std::map<std::string, double> last_ms;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 148;
  // two auxiliary streams + fork / join events: the numeric LU runs the fronts of a tree level in two groups, so that the latency-bound
  // diagonal-block / panel steps of one group run under the GEMMs of the other (created on first use, destroyed with the context)
  cudaStream_t aux_stream[2] = {nullptr, nullptr};
  cudaStream_t cap_stream = nullptr;  // stream the sweep graphs of the solves are captured on (never executes anything)
  // device scratch of the assembly entry points (speed of sound, flame work arrays): kept between calls -- every cudaMalloc / cudaFree is a
  // device-wide synchronisation under the driver's allocation lock, a dozen of them per re-assembly cost more than the kernels
  DevBuf<double> scratch_c, scratch_d[2];
  DevBuf<int32_t> scratch_i[3];
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};

  // mesh
  int order = 0, nloc = 0, nloc3 = 0;
  int64_t n_pts = 0, n_tet = 0, n_tri = 0, dim = 0;
  std::vector<double> xyz;        // 3*n_pts
  std::vector<uint32_t> tets;     // nloc*n_tet, 0-based
  std::vector<uint32_t> tris;     // nloc3*n_tri, 0-based
  DevBuf<double> d_xyz;
  uint64_t xyz_version = 1;       // bumped whenever d_xyz changes (patch-ordered coordinate copies are refreshed lazily)
  DevBuf<uint32_t> d_tets, d_tris;

  std::vector<std::unique_ptr<Pattern>> patterns;
  std::vector<std::unique_ptr<Matrix>> mats;
  std::vector<std::unique_ptr<Family>> fams;
  std::vector<std::shared_ptr<LuSolver>> lus;  // shared_ptr: LuSolver is incomplete here
  std::shared_ptr<WaeShapeSens> shape;         // state of a wae_shape_sens_begin/add/end sequence

  Pattern& pat(int id) {
    if (id < 0 || id >= (int)patterns.size() || !patterns[id]) WAE_THROW(WAE_E_INVALID, "unknown pattern id %d", id);
    return *patterns[id];
  }
  Matrix& mat(int id) {
    if (id < 0 || id >= (int)mats.size() || !mats[id]) WAE_THROW(WAE_E_INVALID, "unknown matrix id %d", id);
    return *mats[id];
  }
  Family& fam(int id) {
    if (id < 0 || id >= (int)fams.size() || !fams[id]) WAE_THROW(WAE_E_INVALID, "unknown family id %d", id);
    return *fams[id];
  }
};

// timing helper: records ms of a phase on the context stream
struct PhaseTimer {
  wae_ctx* h;
  const char* name;
  PhaseTimer(wae_ctx* h_, const char* n) : h(h_), name(n) { cudaEventRecord(h->ev0, h->stream); }
  void stop() {
    cudaEventRecord(h->ev1, h->stream);
    cudaEventSynchronize(h->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->last_ms[name] = ms;
  }
};

// ---- symbolic (host) -------------------------------------------------------------
void wae_build_pattern_from_elements(const uint32_t* conn, int nloc, const std::vector<int64_t>& elems,
                                     int64_t dim, Pattern& P);
void wae_build_slotmap(const uint32_t* conn, int nloc, const Pattern& P, std::vector<int32_t>& slotmap);
// std::vector whose resize() leaves trivially constructible elements uninitialised (multi-GB program arrays that are fully
// overwritten right after: no zero-fill pass)
template <class T>
struct default_init_alloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = default_init_alloc<U>;
  };
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    if constexpr (sizeof...(A) == 0)
      ::new ((void*)p) U;
    else
      ::new ((void*)p) U(std::forward<A>(a)...);
  }
};
template <class T>
using raw_vector = std::vector<T, default_init_alloc<T>>;

struct GatherHost {
  std::vector<int64_t> desc;     // 8 words per patch: blob offset (bytes), pxyz offset (doubles), first 32-element block, first chunk,
                                 // then 8 x int32: nt, nv, ng, nc, blob bytes, offsets of the tets / grp / cnt sections
  raw_vector<uint8_t> blob;
  raw_vector<uint32_t> gvtx, dest;
  raw_vector<uint16_t> res;
  int n_patch = 0, max_slots = 0, max_blob = 0, max_nv = 0, npk = 0;
  int64_t n_pairs = 0, n_sources = 0, n_staged = 0;
};
void wae_build_gather(const double* xyz, const uint32_t* conn, int nloc, const Pattern& P, int slot_cap, GatherHost& G);
struct StarHost {
  std::vector<int64_t> desc;     // one StarDesc (64 bytes) per patch, see assembly_star.h
  raw_vector<uint8_t> blob, blobB;  // blob A (geometry + star pass) and blob B (store pass) of all patches
  raw_vector<uint32_t> gvtx;
  int nloc = 0, n_patch = 0, max_nt = 0, max_nv = 0, max_rows = 0, max_ng = 0, max_a = 0, max_b = 0;
  int64_t max_smem = 0, n_staged = 0, n_entities = 0, n_sources = 0, n_chunks = 0;
};
int wae_star_max_staged(int nloc);
void wae_build_star(const double* xyz, const uint32_t* conn, int nloc, const Pattern& P, int64_t smem_budget, StarHost& G);
bool wae_ensure_star(wae_ctx* h, Pattern& P);  // false: no star program for this pattern (the caller falls back to the pair program)
// Morton rank of the elements, DOF -> incident elements, DOF owner order (shared by the pair and the star program)
struct OwnerOrder {
  std::vector<int64_t> nptr;   // DOF -> incident elements (positions in P.elems), CSR
  std::vector<int32_t> nadj;
  std::vector<int32_t> rank;   // element -> Morton rank
  std::vector<int32_t> order;  // position -> DOF (owner order)
  std::vector<int32_t> pos;    // DOF -> position, -1 if untouched
  int max_inc = 1;             // largest number of elements around one DOF
};
void wae_build_owner_order(const double* xyz, const uint32_t* conn, int nloc, const Pattern& P, OwnerOrder& O);
void wae_ensure_gather(wae_ctx* h, Pattern& P);
void wae_build_bloch(const uint32_t* conn, int nloc, const std::vector<int64_t>& elems, int64_t dim_red, const int64_t* dof_new,
                     const uint8_t* dof_flag, int n_class, std::vector<Pattern>& P, std::vector<int32_t>& slotmap,
                     std::vector<int64_t>& class_base);

// ---- kernels launchers (device) ----------------------------------------------------
void wae_launch_assemble_atomic(wae_ctx* h, Pattern& P, int kind, const double* d_c, int c_per_elem,
                                double scale, double* d_out_a, double* d_out_b);
void wae_launch_assemble_gather(wae_ctx* h, Pattern& P, const double* d_c, double* d_mass, double* d_stiff,
                                double mass_scale);
void wae_launch_assemble_star(wae_ctx* h, Pattern& P, const double* d_c, double* d_mass, double* d_stiff, double mass_scale);
void wae_launch_wallsrc(wae_ctx* h, const int32_t* d_elems, int64_t n, const double* d_c, int c_per_elem, double* d_out);
void wae_combine_device(wae_ctx* h, Family& F, const double* coeffs_host, int slot);
void wae_spmm_device(wae_ctx* h, Family& F, int slot, int trans, int nrhs, const cplx* X, cplx* Y);
void wae_spmm_values(wae_ctx* h, Family& F, const cplx* val, int trans, int nrhs, const cplx* X, cplx* Y);
void wae_values_to_csr(wae_ctx* h, Family& F, const cplx* val, cplx* val_csr);
void wae_spmm_values_csr(wae_ctx* h, Family& F, const cplx* val_csr, int nrhs, const cplx* X, cplx* Y);
void wae_family_ensure_csr(wae_ctx* h, Family& F);
void wae_axpy_term(wae_ctx* h, Family& F, int t, double cr, double ci, cplx* out);
