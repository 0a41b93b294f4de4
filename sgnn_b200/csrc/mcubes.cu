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
// The per-cell arithmetic lives in mc_core.h (shared with the CPU test harness).
// Conventions (restated from the reference): corner pXYZ = cell + (+-0.5 x, +-0.5 y, +-0.5 z); cube-index bit order
// p010 p110 p100 p000 p011 p111 p101 p001; edge e joins corners kEdge[e][0] -> kEdge[e][1] (interpolation order).
#include <math.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "mc_core.h"

namespace {

const unsigned long long h_mc_tri[256] = {SGNN_MC_TABLE};

__global__ void mc_count_kernel(McArgs a, int* __restrict__ counts) {
  const long long total = (long long)a.n0 * a.n1 * a.n2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % a.n2), y = (int)((i / a.n2) % a.n1), z = (int)(i / ((long long)a.n1 * a.n2));
    counts[i] = mc_cell_count(a, x, y, z);
  }
}

__global__ void mc_emit_kernel(McArgs a, const int* __restrict__ offs, float* __restrict__ tris) {
  const long long total = (long long)a.n0 * a.n1 * a.n2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int first = offs[i], n_tri = offs[i + 1] - first;
    if (n_tri == 0) continue;
    const int x = (int)(i % a.n2), y = (int)((i / a.n2) % a.n1), z = (int)(i / ((long long)a.n1 * a.n2));
    mc_cell_emit(a, x, y, z, n_tri, tris + (size_t)first * 9);
  }
}

// the cell (linear z,y,x index) every emitted triangle came from: the reference colours a triangle's three vertices with
// the colour of that cell's voxel (marching_cubes.cpp:228-231,255-257)
__global__ void mc_tri_cells_kernel(const int* __restrict__ offs, long long total, int* __restrict__ tri_cell) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int first = offs[i], last = offs[i + 1];
    for (int t = first; t < last; ++t) tri_cell[t] = (int)i;
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

extern "C" int sgnn_mc_tri_cells(const int32_t* offs, int64_t n_cells, int32_t* tri_cell, void* stream) {
  if (!offs || !tri_cell || n_cells <= 0) return SGNN_E_INVALID;
  mc_tri_cells_kernel<<<sgnn_blocks(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(offs, n_cells, tri_cell);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_mc_merge_host(const float* tris, int64_t n_tri, float* verts, int32_t* faces, int64_t* n_verts,
                                  int64_t* n_faces) {
  return sgnn_mc_merge_host_src(tris, n_tri, verts, faces, nullptr, n_verts, n_faces);
}

extern "C" int sgnn_mc_merge_host_src(const float* tris, int64_t n_tri, float* verts, int32_t* faces, int32_t* vert_src,
                                      int64_t* n_verts, int64_t* n_faces) {
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
      if (vert_src) vert_src[cnt] = (int32_t)v;
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
