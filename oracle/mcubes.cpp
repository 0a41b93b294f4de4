// oracle/mcubes.cpp -- ORACLE (TEST INFRASTRUCTURE ONLY; never linked into or called from the product path).
// CPU restatement of the mesh extraction SG-NN runs after the forward pass:
//   reference torch/marching_cubes/marching_cubes.cpp  run_marching_cubes (:480-517)
//     -> run_marching_cubes_internal (:459-478)  cells in (z, y, x) raster order
//     -> extract_isosurface_at_position (:156-262), trilerp (:107-131), get_voxel (:72-105), vertexInterp (:133-154)
//     -> merge_close_vertices(thresh 1e-5, approx) (:359-455) with hasNearestNeighborApprox (:343-355)
//     -> remove_degenerate_faces (:301-323), remove_duplicate_faces (:266-300)
// Pinned by tests/test_oracle_mcubes.py: equal, bit for bit in the vertices and exactly in the faces, to the outputs of
// the REAL reference (oracle/_ref/marching_cubes_cpp.so, compiled in place by oracle/build_ref.py) on the committed
// fixtures tests/golden/mc_ref.npz and, in the build container, on random volumes.
// The triangulation table is data recovered from the reference by probing (tests/golden/make_mc_golden.py), passed in.
// Compile with -ffp-contract=off: every float operation below is meant to round once, as in the reference build.
#include <cmath>
#include <cstdint>
#include <limits>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

struct V3 { float x, y, z; };
struct I3 {
  int x, y, z;
  bool operator==(const I3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct I3Hash {
  size_t operator()(const I3& v) const {
    return ((size_t)(uint32_t)v.x * 0x9E3779B97F4A7C15ull) ^ ((size_t)(uint32_t)v.y * 0xC2B2AE3D27D4EB4Full) ^ ((size_t)(uint32_t)v.z * 0x165667B19E3779F9ull);
  }
};

struct Volume {
  const float* d; int n0, n1, n2; float trunc;
  // get_voxel: value + "observed and inside the truncation band"
  bool voxel(int x, int y, int z, float* out) const {
    if (z < 0 || z >= n0 || y < 0 || y >= n1 || x < 0 || x >= n2) return false;
    const float v = d[((size_t)z * n1 + y) * n2 + x];
    *out = v;
    return v != -std::numeric_limits<float>::infinity() && std::fabs(v) < trunc;
  }
  // trilerp at a cell corner: corner = cell index + (sx, sy, sz) * 1 - 0.5, i.e. the average of the 2x2x2 voxels
  // starting at (x - 1 + sx, ...), summed in the reference's order with its weights (all 0.5 -> 0.125 each)
  bool corner(int x, int y, int z, int sx, int sy, int sz, float* out) const {
    const int bx = x - 1 + sx, by = y - 1 + sy, bz = z - 1 + sz;
    static const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
    const float w = (0.5f * 0.5f) * 0.5f;
    float dist = 0.0f;
    for (int i = 0; i < 8; ++i) {
      float v;
      if (!voxel(bx + off[i][0], by + off[i][1], bz + off[i][2], &v)) return false;
      dist += w * v;
    }
    *out = dist;
    return true;
  }
};

// corners in cube-index bit order and the 12 edges as (first, second) corner in the order the reference interpolates
const int kCorner[8][3] = {{0, 1, 0}, {1, 1, 0}, {1, 0, 0}, {0, 0, 0}, {0, 1, 1}, {1, 1, 1}, {1, 0, 1}, {0, 0, 1}};
const int kEdge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

V3 interp(float iso, const V3& p1, const V3& p2, float d1, float d2) {
  if (std::fabs(iso - d1) < 0.00001f) return p1;
  if (std::fabs(iso - d2) < 0.00001f) return p2;
  if (std::fabs(d1 - d2) < 0.00001f) return p1;
  const float mu = (iso - d1) / (d2 - d1);
  V3 r;
  r.x = p1.x + mu * (p2.x - p1.x);
  r.y = p1.y + mu * (p2.y - p1.y);
  r.z = p1.z + mu * (p2.z - p1.z);
  return r;
}

int sgn(float v) { return (0.0f < v) - (v < 0.0f); }

std::vector<float> g_verts;
std::vector<int> g_faces;
std::vector<float> g_soup;   // triangle soup before the merge, 9 floats per triangle (tests feed it to the product's merge)

}  // namespace

extern "C" int mc_run(const float* tsdf, int n0, int n1, int n2, float iso, float trunc, float thresh,
                      const signed char* tri_table) {
  Volume vol{tsdf, n0, n1, n2, trunc};
  std::vector<V3> tri;   // triangle soup, 3 vertices per triangle
  for (int z = 0; z < n0; ++z)
    for (int y = 0; y < n1; ++y)
      for (int x = 0; x < n2; ++x) {
        float dc[8];
        V3 pc[8];
        bool ok = true;
        // the reference evaluates the corners in the order 000,100,010,001,110,011,101,111; validity is all-or-nothing
        for (int c = 0; c < 8 && ok; ++c) {
          const int* s = kCorner[c];
          ok = vol.corner(x, y, z, s[0], s[1], s[2], &dc[c]);
          pc[c] = V3{(float)x + (s[0] ? 0.5f : -0.5f), (float)y + (s[1] ? 0.5f : -0.5f), (float)z + (s[2] ? 0.5f : -0.5f)};
        }
        if (!ok) continue;
        unsigned cube = 0;
        for (int c = 0; c < 8; ++c)
          if (dc[c] < iso) cube |= 1u << c;
        bool skip = false;
        for (int k = 0; k < 8 && !skip; ++k)
          for (int l = 0; l < 8 && !skip; ++l) {
            if (dc[k] * dc[l] < 0.0f) skip = std::fabs(dc[k]) + std::fabs(dc[l]) > thresh;
            else skip = std::fabs(dc[k] - dc[l]) > thresh;
          }
        for (int c = 0; c < 8 && !skip; ++c) skip = std::fabs(dc[c]) > thresh;
        if (skip) continue;
        const signed char* t = tri_table + cube * 16;
        for (int i = 0; i < 16 && t[i] >= 0; ++i) {
          const int e = t[i];
          tri.push_back(interp(iso, pc[kEdge[e][0]], pc[kEdge[e][1]], dc[kEdge[e][0]], dc[kEdge[e][1]]));
        }
      }
  g_soup.clear();
  for (const V3& p : tri) { g_soup.push_back(p.x); g_soup.push_back(p.y); g_soup.push_back(p.z); }
  // merge_close_vertices(thresh = 1e-5, approx): first vertex seen in a 3x3x3 neighbourhood of quantised cells wins
  const float q = 0.00001f;
  std::unordered_map<I3, unsigned, I3Hash> grid;
  grid.reserve(tri.size() * 2);
  std::vector<unsigned> look(tri.size());
  g_verts.clear();
  unsigned cnt = 0;
  for (size_t v = 0; v < tri.size(); ++v) {
    const V3& p = tri[v];
    const I3 c{(int)(p.x / q + 0.5f * sgn(p.x)), (int)(p.y / q + 0.5f * sgn(p.y)), (int)(p.z / q + 0.5f * sgn(p.z))};
    unsigned nn = 0xffffffffu;
    for (int i = -1; i <= 1 && nn == 0xffffffffu; ++i)
      for (int j = -1; j <= 1 && nn == 0xffffffffu; ++j)
        for (int k = -1; k <= 1 && nn == 0xffffffffu; ++k) {
          auto it = grid.find(I3{c.x + i, c.y + j, c.z + k});
          if (it != grid.end()) nn = it->second;
        }
    if (nn == 0xffffffffu) {
      grid[c] = cnt;
      g_verts.push_back(p.x); g_verts.push_back(p.y); g_verts.push_back(p.z);
      look[v] = cnt++;
    } else {
      look[v] = nn;
    }
  }
  // faces: drop degenerate ones, then duplicates (same vertex set, first occurrence kept, original winding)
  g_faces.clear();
  struct Key { unsigned a, b, c; bool operator==(const Key& o) const { return a == o.a && b == o.b && c == o.c; } };
  struct KeyHash { size_t operator()(const Key& k) const { return ((size_t)k.a * 73856093u) ^ ((size_t)k.b * 19349669u) ^ ((size_t)k.c * 83492791u); } };
  std::unordered_set<Key, KeyHash> seen;
  for (size_t f = 0; f + 2 < tri.size(); f += 3) {
    const unsigned a = look[f], b = look[f + 1], c = look[f + 2];
    if (a == b || a == c || b == c) continue;
    unsigned s0 = a, s1 = b, s2 = c;
    if (s0 > s1) std::swap(s0, s1);
    if (s1 > s2) std::swap(s1, s2);
    if (s0 > s1) std::swap(s0, s1);
    if (!seen.insert(Key{s0, s1, s2}).second) continue;
    g_faces.push_back((int)a); g_faces.push_back((int)b); g_faces.push_back((int)c);
  }
  return 0;
}

extern "C" void mc_counts(int* n_verts, int* n_faces) {
  *n_verts = (int)(g_verts.size() / 3);
  *n_faces = (int)(g_faces.size() / 3);
}

extern "C" void mc_copy(float* verts, int* faces) {
  for (size_t i = 0; i < g_verts.size(); ++i) verts[i] = g_verts[i];
  for (size_t i = 0; i < g_faces.size(); ++i) faces[i] = g_faces[i];
}

extern "C" int mc_soup_triangles(void) { return (int)(g_soup.size() / 9); }
extern "C" void mc_soup_copy(float* tris) {
  for (size_t i = 0; i < g_soup.size(); ++i) tris[i] = g_soup[i];
}
