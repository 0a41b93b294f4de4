// common.cuh -- shared device helpers for libsgnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sgnn_b200.h"

extern int g_sgnn_last_cuda_error;
extern long long g_sgnn_launches;  // kernels launched by this library (bench.py's gpu_launches evidence)

#define SGNN_CHECK_LAUNCH()                                   \
  do {                                                        \
    ++g_sgnn_launches;                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) {                                 \
      g_sgnn_last_cuda_error = (int)e__;                      \
      return SGNN_E_CUDA;                                     \
    }                                                         \
  } while (0)

#define SGNN_CUDA(call)                                       \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) {                                 \
      g_sgnn_last_cuda_error = (int)e__;                      \
      return SGNN_E_CUDA;                                     \
    }                                                         \
  } while (0)

static inline int sgnn_blocks(int64_t n, int threads, int64_t cap = (int64_t)148 * 64) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// Device view of SgnnGrid (passed by value to kernels).
struct GridView {
  int nb, d0, d1, d2, wx;
  long long n_words;
  const unsigned long long* mask;
  const int* prefix;
  const int* row_of_rank;
};

static inline GridView make_view(const SgnnGrid* g) {
  GridView v;
  v.nb = g->nb; v.d0 = g->d0; v.d1 = g->d1; v.d2 = g->d2; v.wx = g->wx;
  v.n_words = g->n_words;
  v.mask = (const unsigned long long*)g->mask;
  v.prefix = g->prefix;
  v.row_of_rank = g->row_of_rank;
  return v;
}

// word index of the x-row holding (b,z,y,x); caller guarantees in-extent coordinates
__device__ __forceinline__ long long grid_word(const GridView& g, int b, int z, int y, int x) {
  return (((long long)b * g.d0 + z) * g.d1 + y) * g.wx + (x >> 6);
}

// row id of an in-extent cell, or -1
__device__ __forceinline__ int grid_row(const GridView& g, int b, int z, int y, int x) {
  long long w = grid_word(g, b, z, y, x);
  unsigned long long m = __ldg(g.mask + w);
  unsigned long long bit = 1ull << (x & 63);
  if (!(m & bit)) return -1;
  int rank = __ldg(g.prefix + w) + __popcll(m & (bit - 1));
  return g.row_of_rank ? __ldg(g.row_of_rank + rank) : rank;
}

__device__ __forceinline__ int grid_row_checked(const GridView& g, int b, int z, int y, int x) {
  if ((unsigned)b >= (unsigned)g.nb || (unsigned)z >= (unsigned)g.d0 ||
      (unsigned)y >= (unsigned)g.d1 || (unsigned)x >= (unsigned)g.d2)
    return -1;
  return grid_row(g, b, z, y, x);
}

// The 27 submanifold neighbours of site c = (z,y,x,b) (offset id k = (dz+1)*9 + (dy+1)*3 + (dx+1), -1 = absent or outside the
// extent): shared by the rulebook kernels (grid.cu) and the fused rulebook + tile-plan kernel (conv_ur.cu).  The three
// x-neighbours of a (z+dz, y+dy) row live in one mask word (two when x sits on a word edge): one word + one prefix load and the
// three bits are tested in registers.
__device__ __forceinline__ void rulebook_probe27(const GridView& g, int4 c, int (&idx)[27]) {
  const int z = c.x, y = c.y, x = c.z, b = c.w;
  const bool inb = (unsigned)z < (unsigned)g.d0 && (unsigned)y < (unsigned)g.d1 &&
                   (unsigned)x < (unsigned)g.d2 && (unsigned)b < (unsigned)g.nb;
  const bool edge = ((x & 63) == 0) || ((x & 63) == 63);
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int zz = z + dz, yy = y + dy;
      const int k0 = (dz + 1) * 9 + (dy + 1) * 3;
      int r0 = -1, r1 = -1, r2 = -1;
      if (inb && (unsigned)zz < (unsigned)g.d0 && (unsigned)yy < (unsigned)g.d1) {
        if (!edge) {
          long long w = grid_word(g, b, zz, yy, x);
          unsigned long long m = __ldg(g.mask + w);
          unsigned sh = (x & 63) - 1;
          unsigned bits = (unsigned)(m >> sh) & 7u;
          if (bits) {
            int base = __ldg(g.prefix + w) + __popcll(m & ((1ull << sh) - 1));
            int rk0 = base, rk1 = base + (bits & 1), rk2 = rk1 + ((bits >> 1) & 1);
            if (g.row_of_rank) {
              if (bits & 1) r0 = __ldg(g.row_of_rank + rk0);
              if (bits & 2) r1 = __ldg(g.row_of_rank + rk1);
              if (bits & 4) r2 = __ldg(g.row_of_rank + rk2);
            } else {
              if (bits & 1) r0 = rk0;
              if (bits & 2) r1 = rk1;
              if (bits & 4) r2 = rk2;
            }
          }
        } else {
          if (x - 1 >= 0) r0 = grid_row(g, b, zz, yy, x - 1);
          r1 = grid_row(g, b, zz, yy, x);
          if (x + 1 < g.d2) r2 = grid_row(g, b, zz, yy, x + 1);
        }
      }
      idx[k0] = r0; idx[k0 + 1] = r1; idx[k0 + 2] = r2;
    }
  }
}

// literal fp32 restatement of `nn.Sigmoid()(x) > 0.5` (model.py:233,322)
__device__ __forceinline__ bool sigmoid_gt_half(float x) {
  float s = 1.0f / (1.0f + expf(-x));
  return s > 0.5f;
}

// scan.cu
int sgnn_scan_exclusive(const void* in, int mode, int* out, int64_t n, void* scratch,
                        size_t scratch_bytes, cudaStream_t st);
enum { SCAN_I32 = 0, SCAN_POPC64 = 1, SCAN_U8 = 2 };
