// mc_core.h -- per-cell arithmetic of the marching-cubes kernels (csrc/mcubes.cu), written once for the device and,
// for the CPU test harness tests/mc_host_harness.cpp (test infrastructure: it runs exactly this code cell by cell on
// the host so that the indexing, the table decoding and the float operation order can be checked against the oracle
// without a GPU), for the host.  Not a CPU fallback: nothing in the library calls the host instantiation.
#pragma once
#include <math.h>
#include "mc_table.h"

#ifdef __CUDACC__
#define MC_FN __device__ __forceinline__
#define MC_TABLE_QUAL __constant__
#define MC_LOAD(p) __ldg(p)
#define MC_ADD(a, b) __fadd_rn((a), (b))
#define MC_SUB(a, b) __fsub_rn((a), (b))
#define MC_MUL(a, b) __fmul_rn((a), (b))
#define MC_DIV(a, b) __fdiv_rn((a), (b))
#else   // host harness: compile with -ffp-contract=off so that every operation rounds once
#define MC_FN static inline
#define MC_TABLE_QUAL static const
#define MC_LOAD(p) (*(p))
#define MC_ADD(a, b) ((a) + (b))
#define MC_SUB(a, b) ((a) - (b))
#define MC_MUL(a, b) ((a) * (b))
#define MC_DIV(a, b) ((a) / (b))
#endif

MC_TABLE_QUAL unsigned long long c_mc_tri[256] = {SGNN_MC_TABLE};
// corners in cube-index bit order (p010 p110 p100 p000 p011 p111 p101 p001) as (sx, sy, sz); edges as corner pairs in the
// order the reference interpolates them
MC_TABLE_QUAL int c_corner[8][3] = {{0, 1, 0}, {1, 1, 0}, {1, 0, 0}, {0, 0, 0}, {0, 1, 1}, {1, 1, 1}, {1, 0, 1}, {0, 0, 1}};
MC_TABLE_QUAL int c_edge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

struct McArgs {
  const float* tsdf; int n0, n1, n2;
  float iso, trunc, thresh;
};

// get_voxel (:72-105): value, and whether it is observed and inside the truncation band
MC_FN bool mc_voxel(const McArgs& a, int x, int y, int z, float* v) {
  if (z < 0 || z >= a.n0 || y < 0 || y >= a.n1 || x < 0 || x >= a.n2) return false;
  const float d = MC_LOAD(a.tsdf + ((size_t)z * a.n1 + y) * a.n2 + x);
  *v = d;
  return d != -INFINITY && fabsf(d) < a.trunc;
}

// trilerp (:107-131) at the corner (sx,sy,sz) of cell (x,y,z): the 2x2x2 voxels starting at (x-1+sx, ...), weights
// 0.5*0.5*0.5 each, accumulated in the reference's order 000,100,010,001,110,011,101,111 (x,y,z offsets)
MC_FN bool mc_corner(const McArgs& a, int x, int y, int z, int sx, int sy, int sz, float* out) {
  const int bx = x - 1 + sx, by = y - 1 + sy, bz = z - 1 + sz;
  const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
  float dist = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v;
    if (!mc_voxel(a, bx + off[i][0], by + off[i][1], bz + off[i][2], &v)) return false;
    dist = MC_ADD(dist, MC_MUL(0.125f, v));
  }
  *out = dist;
  return true;
}

// corner values + cube index of a cell; false = the reference emits nothing for it
MC_FN bool mc_cell(const McArgs& a, int x, int y, int z, float (&dc)[8], unsigned* cube) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (!mc_corner(a, x, y, z, c_corner[c][0], c_corner[c][1], c_corner[c][2], &dc[c])) return false;
  unsigned idx = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (dc[c] < a.iso) idx |= 1u << c;
  for (int k = 0; k < 8; ++k)
    for (int l = 0; l < 8; ++l) {
      if (MC_MUL(dc[k], dc[l]) < 0.0f) {
        if (MC_ADD(fabsf(dc[k]), fabsf(dc[l])) > a.thresh) return false;
      } else {
        if (fabsf(MC_SUB(dc[k], dc[l])) > a.thresh) return false;
      }
    }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (fabsf(dc[c]) > a.thresh) return false;
  *cube = idx;
  return true;
}

MC_FN int mc_tri_vertices(unsigned long long w) {   // nibbles before the 0xF terminator
  int n = 0;
  while (n < 16 && ((w >> (4 * n)) & 0xF) != 0xF) ++n;
  return n;
}

// triangles the reference emits for cell (x,y,z)
MC_FN int mc_cell_count(const McArgs& a, int x, int y, int z) {
  float dc[8];
  unsigned cube;
  if (!mc_cell(a, x, y, z, dc, &cube)) return 0;
  return mc_tri_vertices(c_mc_tri[cube]) / 3;
}

// vertexInterp (:133-154), operand order kept, no contraction
MC_FN void mc_interp(float iso, const float (&p1)[3], const float (&p2)[3], float d1, float d2, float* out) {
  if (fabsf(MC_SUB(iso, d1)) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  if (fabsf(MC_SUB(iso, d2)) < 0.00001f) { out[0] = p2[0]; out[1] = p2[1]; out[2] = p2[2]; return; }
  if (fabsf(MC_SUB(d1, d2)) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  const float mu = MC_DIV(MC_SUB(iso, d1), MC_SUB(d2, d1));
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = MC_ADD(p1[i], MC_MUL(mu, MC_SUB(p2[i], p1[i])));
}

// the n_tri triangles of cell (x,y,z), 9 floats each, to dst
MC_FN void mc_cell_emit(const McArgs& a, int x, int y, int z, int n_tri, float* dst) {
  float dc[8];
  unsigned cube;
  if (!mc_cell(a, x, y, z, dc, &cube)) return;   // cannot happen when n_tri came from mc_cell_count
  const unsigned long long w = c_mc_tri[cube];
  for (int v = 0; v < 3 * n_tri; ++v) {
    const int e = (int)((w >> (4 * v)) & 0xF);
    const int ca = c_edge[e][0], cb = c_edge[e][1];
    float p1[3], p2[3];
    p1[0] = (float)x + (c_corner[ca][0] ? 0.5f : -0.5f); p1[1] = (float)y + (c_corner[ca][1] ? 0.5f : -0.5f);
    p1[2] = (float)z + (c_corner[ca][2] ? 0.5f : -0.5f);
    p2[0] = (float)x + (c_corner[cb][0] ? 0.5f : -0.5f); p2[1] = (float)y + (c_corner[cb][1] ? 0.5f : -0.5f);
    p2[2] = (float)z + (c_corner[cb][2] ? 0.5f : -0.5f);
    mc_interp(a.iso, p1, p2, dc[ca], dc[cb], dst + 3 * v);
  }
}
