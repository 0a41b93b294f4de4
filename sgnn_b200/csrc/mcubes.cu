// mcubes.cu -- SURVEY 8(f4): the mesh extraction SG-NN runs on the predicted dense TSDF after the forward pass
// (reference torch/marching_cubes/marching_cubes.cpp:459-517 run_marching_cubes, called from data_util.py:270-284).
// The reference walks the dense grid in a triple loop on one CPU core (extract_isosurface_at_position :156-262:
// 8 corner values = trilinear samples at the cell corners, each the average of 2x2x2 voxels; cube index; vertices by
// linear interpolation on the 12 edges; triangles from the table), then merges vertices closer than 1e-5 through a hash
// grid in first-come order (merge_close_vertices :359-455, approx) and drops degenerate / duplicate faces.
//
// Here the grid walk -- all of the floating-point work and 99 % of the reference's time -- is two CUDA kernels
// (count triangles per cell -> exclusive scan -> emit), one thread per cell, producing the triangle soup in exactly the
// reference's order (cells z,y,x raster, table order) and with exactly its bits: every float operation is written with
// a round-to-nearest intrinsic in the reference's operand order so that nvcc cannot contract multiplies and adds into
// FMAs.  The first-come vertex merge is inherently sequential and order defining; it stays on the host
// (sgnn_mc_merge_host, O(vertices), an open-addressing hash grid).
//
// Conventions (restated from the reference): corner pXYZ = cell + (+-0.5 x, +-0.5 y, +-0.5 z); cube-index bit order
// p010 p110 p100 p000 p011 p111 p101 p001; edge e joins corners kEdge[e][0] -> kEdge[e][1] (interpolation order).
#include <math.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "mc_table.h"

namespace {

__constant__ unsigned long long c_mc_tri[256] = {SGNN_MC_TABLE};
const unsigned long long h_mc_tri[256] = {SGNN_MC_TABLE};

__constant__ int c_corner[8][3] = {{0, 1, 0}, {1, 1, 0}, {1, 0, 0}, {0, 0, 0}, {0, 1, 1}, {1, 1, 1}, {1, 0, 1}, {0, 0, 1}};
__constant__ int c_edge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

struct McArgs {
  const float* tsdf; int n0, n1, n2;
  float iso, trunc, thresh;
};

// get_voxel (:72-105): value, and whether it is observed and inside the truncation band
__device__ __forceinline__ bool mc_voxel(const McArgs& a, int x, int y, int z, float* v) {
  if (z < 0 || z >= a.n0 || y < 0 || y >= a.n1 || x < 0 || x >= a.n2) return false;
  const float d = __ldg(a.tsdf + ((size_t)z * a.n1 + y) * a.n2 + x);
  *v = d;
  return d != -INFINITY && fabsf(d) < a.trunc;
}

// trilerp (:107-131) at the corner (sx,sy,sz) of cell (x,y,z): the 2x2x2 voxels starting at (x-1+sx, ...), weights
// 0.5*0.5*0.5 each, accumulated in the reference's order 000,100,010,001,110,011,101,111 (x,y,z offsets)
__device__ __forceinline__ bool mc_corner(const McArgs& a, int x, int y, int z, int sx, int sy, int sz, float* out) {
  const int bx = x - 1 + sx, by = y - 1 + sy, bz = z - 1 + sz;
  const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
  float dist = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v;
    if (!mc_voxel(a, bx + off[i][0], by + off[i][1], bz + off[i][2], &v)) return false;
    dist = __fadd_rn(dist, __fmul_rn(0.125f, v));
  }
  *out = dist;
  return true;
}

// corner values + cube index of a cell; false = the reference emits nothing for it
__device__ __forceinline__ bool mc_cell(const McArgs& a, int x, int y, int z, float (&dc)[8], unsigned* cube) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (!mc_corner(a, x, y, z, c_corner[c][0], c_corner[c][1], c_corner[c][2], &dc[c])) return false;
  unsigned idx = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (dc[c] < a.iso) idx |= 1u << c;
  for (int k = 0; k < 8; ++k)
    for (int l = 0; l < 8; ++l) {
      if (__fmul_rn(dc[k], dc[l]) < 0.0f) {
        if (__fadd_rn(fabsf(dc[k]), fabsf(dc[l])) > a.thresh) return false;
      } else {
        if (fabsf(__fsub_rn(dc[k], dc[l])) > a.thresh) return false;
      }
    }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (fabsf(dc[c]) > a.thresh) return false;
  *cube = idx;
  return true;
}

__device__ __forceinline__ int mc_tri_vertices(unsigned long long w) {   // nibbles before the 0xF terminator
  int n = 0;
  while (n < 16 && ((w >> (4 * n)) & 0xF) != 0xF) ++n;
  return n;
}

__global__ void mc_count_kernel(McArgs a, int* __restrict__ counts) {
  const long long total = (long long)a.n0 * a.n1 * a.n2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % a.n2), y = (int)((i / a.n2) % a.n1), z = (int)(i / ((long long)a.n1 * a.n2));
    float dc[8];
    unsigned cube;
    int n = 0;
    if (mc_cell(a, x, y, z, dc, &cube)) n = mc_tri_vertices(c_mc_tri[cube]) / 3;
    counts[i] = n;
  }
}

// vertexInterp (:133-154), operand order kept, no contraction
__device__ __forceinline__ void mc_interp(float iso, const float (&p1)[3], const float (&p2)[3], float d1, float d2,
                                          float* out) {
  if (fabsf(__fsub_rn(iso, d1)) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  if (fabsf(__fsub_rn(iso, d2)) < 0.00001f) { out[0] = p2[0]; out[1] = p2[1]; out[2] = p2[2]; return; }
  if (fabsf(__fsub_rn(d1, d2)) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  const float mu = __fdiv_rn(__fsub_rn(iso, d1), __fsub_rn(d2, d1));
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = __fadd_rn(p1[i], __fmul_rn(mu, __fsub_rn(p2[i], p1[i])));
}

__global__ void mc_emit_kernel(McArgs a, const int* __restrict__ offs, float* __restrict__ tris) {
  const long long total = (long long)a.n0 * a.n1 * a.n2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int first = offs[i], n_tri = offs[i + 1] - first;
    if (n_tri == 0) continue;
    const int x = (int)(i % a.n2), y = (int)((i / a.n2) % a.n1), z = (int)(i / ((long long)a.n1 * a.n2));
    float dc[8];
    unsigned cube;
    if (!mc_cell(a, x, y, z, dc, &cube)) continue;   // cannot happen: the count kernel saw the same cell
    const unsigned long long w = c_mc_tri[cube];
    float* dst = tris + (size_t)first * 9;
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
}

McArgs mc_args(const float* tsdf, int n0, int n1, int n2, float iso, float trunc, float thresh) {
  McArgs a;
  a.tsdf = tsdf; a.n0 = n0; a.n1 = n1; a.n2 = n2; a.iso = iso; a.trunc = trunc; a.thresh = thresh;
  return a;
}

// ---- host side of merge_close_vertices(1e-5, approx): open-addressing hash grid over the quantised coordinates
struct QCell { int x, y, z; unsigned id; bool used; };

inline size_t qhash(int x, int y, int z) {
  unsigned long long h = (unsigned long long)(unsigned)x * 0x9E3779B185EBCA87ull;
  h ^= (unsigned long long)(unsigned)y * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= (unsigned long long)(unsigned)z * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
  return (size_t)(h ^ (h >> 29));
}

inline int sgn_f(float v) { return (0.0f < v) - (v < 0.0f); }

}  // namespace

extern "C" int sgnn_mc_count(const float* tsdf, int32_t n0, int32_t n1, int32_t n2, float isovalue, float truncation,
                             float thresh, int32_t* offs, void* scratch, size_t scratch_bytes, void* stream) {
  if (!tsdf || !offs || n0 <= 0 || n1 <= 0 || n2 <= 0) return SGNN_E_INVALID;
  const long long total = (long long)n0 * n1 * n2;
  if (total >= 0x7fffffffLL / 5) return SGNN_E_TOO_LARGE;          // <= 5 triangles per cell, int32 offsets
  const size_t need = sgnn_scan_scratch_bytes(total) + (size_t)total * 4 + 256;
  if (!scratch || scratch_bytes < need) return SGNN_E_NOMEM;
  int* counts = (int*)scratch;
  void* scan_scratch = (char*)scratch + (((size_t)total * 4 + 255) & ~(size_t)255);
  cudaStream_t st = (cudaStream_t)stream;
  mc_count_kernel<<<sgnn_blocks(total, 256), 256, 0, st>>>(mc_args(tsdf, n0, n1, n2, isovalue, truncation, thresh), counts);
  SGNN_CHECK_LAUNCH();
  return sgnn_scan_exclusive(counts, SCAN_I32, offs, total, scan_scratch, scratch_bytes - (((size_t)total * 4 + 255) & ~(size_t)255), st);
}

extern "C" size_t sgnn_mc_scratch_bytes(int32_t n0, int32_t n1, int32_t n2) {
  const long long total = (long long)n0 * n1 * n2;
  return sgnn_scan_scratch_bytes(total) + (size_t)total * 4 + 512;
}

extern "C" int sgnn_mc_emit(const float* tsdf, int32_t n0, int32_t n1, int32_t n2, float isovalue, float truncation,
                            float thresh, const int32_t* offs, float* tris, void* stream) {
  if (!tsdf || !offs || !tris || n0 <= 0 || n1 <= 0 || n2 <= 0) return SGNN_E_INVALID;
  const long long total = (long long)n0 * n1 * n2;
  mc_emit_kernel<<<sgnn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
      mc_args(tsdf, n0, n1, n2, isovalue, truncation, thresh), offs, tris);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_mc_merge_host(const float* tris, int64_t n_tri, float* verts, int32_t* faces, int64_t* n_verts,
                                  int64_t* n_faces) {
  if (n_tri < 0 || (n_tri > 0 && (!tris || !verts || !faces)) || !n_verts || !n_faces) return SGNN_E_INVALID;
  if (n_tri * 3 >= 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  const size_t nv = (size_t)n_tri * 3;
  size_t cap = 64;
  while (cap < nv * 2 + 16) cap <<= 1;
  std::vector<QCell> table(cap);
  for (auto& c : table) c.used = false;
  auto find = [&](int x, int y, int z) -> const QCell* {
    size_t h = qhash(x, y, z) & (cap - 1);
    while (table[h].used) {
      if (table[h].x == x && table[h].y == y && table[h].z == z) return &table[h];
      h = (h + 1) & (cap - 1);
    }
    return nullptr;
  };
  std::vector<unsigned> look(nv);
  unsigned cnt = 0;
  const float q = 0.00001f;
  for (size_t v = 0; v < nv; ++v) {
    const float px = tris[3 * v], py = tris[3 * v + 1], pz = tris[3 * v + 2];
    const int cx = (int)(px / q + 0.5f * sgn_f(px)), cy = (int)(py / q + 0.5f * sgn_f(py)), cz = (int)(pz / q + 0.5f * sgn_f(pz));
    const QCell* hit = nullptr;
    for (int i = -1; i <= 1 && !hit; ++i)
      for (int j = -1; j <= 1 && !hit; ++j)
        for (int k = -1; k <= 1 && !hit; ++k) hit = find(cx + i, cy + j, cz + k);
    if (hit) {
      look[v] = hit->id;
    } else {
      size_t h = qhash(cx, cy, cz) & (cap - 1);
      while (table[h].used) h = (h + 1) & (cap - 1);
      table[h].x = cx; table[h].y = cy; table[h].z = cz; table[h].id = cnt; table[h].used = true;
      verts[3 * (size_t)cnt] = px; verts[3 * (size_t)cnt + 1] = py; verts[3 * (size_t)cnt + 2] = pz;
      look[v] = cnt++;
    }
  }
  // faces: degenerate ones out, then the first of every vertex set (original winding kept)
  size_t fcap = 64;
  while (fcap < (size_t)n_tri * 2 + 16) fcap <<= 1;
  std::vector<unsigned> fkeys(fcap * 3);
  std::vector<unsigned char> fused(fcap, 0);
  int64_t nf = 0;
  for (int64_t t = 0; t < n_tri; ++t) {
    const unsigned a = look[3 * t], b = look[3 * t + 1], c = look[3 * t + 2];
    if (a == b || a == c || b == c) continue;
    unsigned s0 = a, s1 = b, s2 = c, tmp;
    if (s0 > s1) { tmp = s0; s0 = s1; s1 = tmp; }
    if (s1 > s2) { tmp = s1; s1 = s2; s2 = tmp; }
    if (s0 > s1) { tmp = s0; s0 = s1; s1 = tmp; }
    size_t h = qhash((int)s0, (int)s1, (int)s2) & (fcap - 1);
    bool dup = false;
    while (fused[h]) {
      if (fkeys[3 * h] == s0 && fkeys[3 * h + 1] == s1 && fkeys[3 * h + 2] == s2) { dup = true; break; }
      h = (h + 1) & (fcap - 1);
    }
    if (dup) continue;
    fused[h] = 1; fkeys[3 * h] = s0; fkeys[3 * h + 1] = s1; fkeys[3 * h + 2] = s2;
    faces[3 * nf] = (int32_t)a; faces[3 * nf + 1] = (int32_t)b; faces[3 * nf + 2] = (int32_t)c;
    ++nf;
  }
  *n_verts = cnt;
  *n_faces = nf;
  return SGNN_OK;
}

// The triangulation table the kernels use (tests compare it with the one recovered from the reference).
extern "C" int sgnn_mc_table(uint64_t* out256) {
  if (!out256) return SGNN_E_INVALID;
  memcpy(out256, h_mc_tri, sizeof(h_mc_tri));
  return SGNN_OK;
}
