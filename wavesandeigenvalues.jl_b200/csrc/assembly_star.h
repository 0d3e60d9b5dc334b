// Layout of a patch of the star program in global / shared memory, shared by the host builder (assembly_star_symbolic.cpp), the
// kernel and its host replay (assembly_star.cu).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define WAE_STAR_HD __host__ __device__
#else
#define WAE_STAR_HD
#endif

// 64-byte patch descriptor
struct StarDesc {
  long long blobA, pxyz, blobB, reserved;  // byte offset of blob A, offset (doubles) of the patch's coordinates, byte offset of blob B
  int nt, nv, ng, nc, nsrc, ncode, bytesA, bytesB;
};

// blob A (geometry + star pass): [ lvtx: nt x 4 u16 | tets: nt i32 | grp: ng x (u32 first source, u32 row | type << 16 | iterations << 24)
//                                  | cnt: ng x 32 u8 | src: nsrc u16 ]
// blob B (store pass):           [ chunk: nc x (u32 first nonzero, u32 length | code offset << 6) | code: ncode u16 ]
// every section starts on a 16-byte boundary; both blobs are multiples of 16 bytes (one bulk copy each)
struct StarBlob {
  int o_tets, o_grp, o_cnt, o_src, bytesA, o_code, bytesB;
  WAE_STAR_HD static int pad16(int x) { return (x + 15) & ~15; }
  WAE_STAR_HD StarBlob(int nt, int ng, int nsrc, int nc, int ncode) {
    o_tets = pad16(8 * nt);
    o_grp = o_tets + pad16(4 * nt);
    o_cnt = o_grp + pad16(8 * ng);
    o_src = o_cnt + 32 * ng;
    bytesA = o_src + pad16(2 * nsrc);
    o_code = pad16(8 * nc);
    bytesB = o_code + pad16(2 * ncode);
  }
};

// ---- record blocks ---------------------------------------------------------------------------------------------------------------
// A group of 32 sub-simplices of one type owns a block of rows of 32 doubles (entry-major: row r holds entry r of the 32 simplices):
//   P2  vertex   [K0 | M]                          edge  [K0..K3 | M(v,v) | M(v,e) | M(e,e)]
//       face     [K0..K5 | M(v,e) | M(e,e')]       tet   [K0..K2 | M(e,e')]
//   P1  vertex   [K0 | M]                          edge  [K0 | M]
// K_j = role j of the simplex (fem_gen.h: wae_p*_star_*), M = (mass coefficient of the role class) * sum of |det| over the star.
// Vertex stars are split over WAE_STAR_VSPLIT lanes (a vertex of a Kuhn mesh has 24 elements around it, an edge 4-8): lanes 4 v .. 4 v + 3
// sum a quarter of the star each, the partial sums are combined as (p0 + p1) + (p2 + p3) and lane 4 v owns the record.
#define WAE_STAR_VSPLIT 4
// Row stride of a record block in doubles.  33, not 32: the entries of ONE simplex (its roles sit in the same lane of consecutive rows)
// then fall into different shared-memory banks -- a column holds up to three nonzeros of the same simplex, which the store pass reads in
// one warp step.
#ifndef WAE_STAR_RS
#define WAE_STAR_RS 33
#endif
WAE_STAR_HD inline int star_rows(int nloc, int type) { return nloc == 4 ? 2 : (type == 0 ? 2 : type == 1 ? 7 : type == 2 ? 8 : 4); }
// roles: P2 0 vertex | 1..4 edge | 5..10 face | 11..13 tet;  P1 0 vertex | 1 edge
WAE_STAR_HD inline int star_krow(int nloc, int role) { return nloc == 4 ? 0 : (int)((0x21054321032100ULL >> (4 * role)) & 7); }
WAE_STAR_HD inline int star_mrow(int nloc, int role) { return nloc == 4 ? 1 : (int)((0x33377766665541ULL >> (4 * role)) & 7); }

// The shared-memory words (8 bytes each, relative to the start of the gram blocks) one source word makes its lane read in the star pass, in
// the order of the kernel's loads (star_sums in assembly_star.cu; the determinant comes last here, the order is irrelevant for the banks).
// Used by the bank-conflict simulator of the host replay and by the builder, which orders the sources of a lane so that the lanes of a
// half-warp read different bank pairs in the same iteration.
WAE_STAR_HD inline int star_load_words(int nloc, int type, unsigned w, int* word) {
  const int ia = w & 3, ib = (w >> 2) & 3, ic = (w >> 4) & 3;
  const int t = type < 3 ? (int)(w >> (2 * (type + 1))) : (int)w;
  int idx[7], n = 0;
  if (type == 0) idx[n++] = 5 * ia;
  else if (type == 1 && nloc == 4) idx[n++] = 4 * ia + ib;
  else if (type == 1) { idx[n++] = 5 * ia; idx[n++] = 5 * ib; idx[n++] = 4 * ia + ib; }
  else if (type == 2) { idx[n++] = 5 * ia; idx[n++] = 5 * ib; idx[n++] = 5 * ic; idx[n++] = 4 * ia + ib; idx[n++] = 4 * ia + ic; idx[n++] = 4 * ib + ic; }
  else { idx[n++] = 1; idx[n++] = 2; idx[n++] = 3; idx[n++] = 6; idx[n++] = 7; idx[n++] = 11; }
  idx[n++] = 16;
  for (int j = 0; j < n; j++) word[j] = t * 17 + idx[j];
  return n;
}

// ---- shared memory of one CTA: gram blocks of the staged elements (17 doubles each) | record rows | blob A | coordinates | blob B ------
struct StarLayout {
  int off_rec, off_a, off_px, off_b;
  long long total;
  WAE_STAR_HD static long long pad(long long x) { return (x + 15) & ~15LL; }
  WAE_STAR_HD StarLayout(int max_nt, int max_rows, int max_a, int max_b, int max_nv) {
    long long o = pad((long long)max_nt * 17 * 8);
    off_rec = (int)o;
    o += pad((long long)max_rows * WAE_STAR_RS * 8);
    off_a = (int)o;
    o += pad(max_a);
    off_px = (int)o;
    o += pad((long long)max_nv * 24);
    off_b = (int)o;
    o += pad(max_b);
    total = o;
  }
};
#define WAE_STAR_STATIC_SMEM 512  // descriptor ring + mbarriers (static shared memory of the kernel), rounded up
