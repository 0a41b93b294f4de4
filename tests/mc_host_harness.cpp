// tests/mc_host_harness.cpp -- TEST INFRASTRUCTURE (not part of the product, not a CPU fallback): runs the per-cell code
// of the marching-cubes kernels (sgnn_b200/csrc/mc_core.h, the very functions mc_count_kernel / mc_emit_kernel call)
// cell by cell on the host, so that tests can check indexing, table decoding and float operation order against the
// oracle without a GPU.  Built by tests/test_mesh.py with  g++ -O2 -ffp-contract=off.
#include <stddef.h>
#include <vector>
#include "../sgnn_b200/csrc/mc_core.h"

static std::vector<float> g_tris;
static std::vector<int> g_cells;   // == mc_tri_cells_kernel: the cell of every triangle

extern "C" int mch_run(const float* tsdf, int n0, int n1, int n2, float iso, float trunc, float thresh) {
  McArgs a;
  a.tsdf = tsdf; a.n0 = n0; a.n1 = n1; a.n2 = n2; a.iso = iso; a.trunc = trunc; a.thresh = thresh;
  const long long total = (long long)n0 * n1 * n2;
  std::vector<int> offs(total + 1, 0);
  for (long long i = 0; i < total; ++i) {                      // == mc_count_kernel + exclusive scan
    const int x = (int)(i % n2), y = (int)((i / n2) % n1), z = (int)(i / ((long long)n1 * n2));
    offs[i + 1] = offs[i] + mc_cell_count(a, x, y, z);
  }
  g_tris.assign((size_t)offs[total] * 9, 0.f);
  g_cells.assign((size_t)offs[total], 0);
  for (long long i = 0; i < total; ++i) {                      // == mc_emit_kernel
    const int n_tri = offs[i + 1] - offs[i];
    if (!n_tri) continue;
    const int x = (int)(i % n2), y = (int)((i / n2) % n1), z = (int)(i / ((long long)n1 * n2));
    mc_cell_emit(a, x, y, z, n_tri, g_tris.data() + (size_t)offs[i] * 9);
    for (int t = offs[i]; t < offs[i + 1]; ++t) g_cells[t] = (int)i;
  }
  return offs[total];
}

extern "C" void mch_copy(float* tris) {
  for (size_t i = 0; i < g_tris.size(); ++i) tris[i] = g_tris[i];
}

extern "C" void mch_cells(int* cells) {
  for (size_t i = 0; i < g_cells.size(); ++i) cells[i] = g_cells[i];
}
