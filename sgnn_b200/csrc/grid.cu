// grid.cu -- active-site sets and rulebooks.
//
// B200-first replacement for scn's per-sample google::dense_hash_map: the
// active set of a resolution is a bitmask over the bounded (batch,z,y,x) extent -- 1 bit per
// cell, 64 x-cells per word -- with an exclusive popcount prefix per word.  For the workloads
// of BASELINE.json the whole structure is L2 resident (32 blocks of 64^3: 1 MiB mask + 0.5 MiB
// prefix; 763 blocks: 25 MiB + 12.5 MiB, L2 is 126 MB), so a neighbour probe is an L2 hit on a
// word shared with the x-neighbours instead of a random 32-byte DRAM sector per hash probe.
#include "common.cuh"

// ------------------------------------------------------------------ build
__global__ void grid_set_bits_kernel(GridView g, unsigned long long* __restrict__ mask,
                                     const void* __restrict__ coords, int is_i64, long long n,
                                     int* __restrict__ coords_out, int* __restrict__ status) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int z, y, x, b;
    bool ok = true;
    if (is_i64) {
      const longlong2* p = reinterpret_cast<const longlong2*>(coords) + 2 * i;
      longlong2 a = __ldg(p), c = __ldg(p + 1);
      ok = (unsigned long long)a.x < (unsigned long long)g.d0 &&
           (unsigned long long)a.y < (unsigned long long)g.d1 &&
           (unsigned long long)c.x < (unsigned long long)g.d2 &&
           (unsigned long long)c.y < (unsigned long long)g.nb;
      z = (int)a.x; y = (int)a.y; x = (int)c.x; b = (int)c.y;
      if (!ok) { z = y = x = b = -1; }
    } else {
      int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
      z = c.x; y = c.y; x = c.z; b = c.w;
      ok = (unsigned)z < (unsigned)g.d0 && (unsigned)y < (unsigned)g.d1 &&
           (unsigned)x < (unsigned)g.d2 && (unsigned)b < (unsigned)g.nb;
    }
    if (coords_out) reinterpret_cast<int4*>(coords_out)[i] = make_int4(z, y, x, b);
    if (ok) {
      atomicOr(mask + grid_word(g, b, z, y, x), 1ull << (x & 63));
    } else if (status) {
      *status = 1;
    }
  }
}

__global__ void grid_fill_rank_kernel(GridView g, const int* __restrict__ coords, long long n,
                                      int* __restrict__ row_of_rank) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
    if ((unsigned)c.x >= (unsigned)g.d0 || (unsigned)c.y >= (unsigned)g.d1 ||
        (unsigned)c.z >= (unsigned)g.d2 || (unsigned)c.w >= (unsigned)g.nb)
      continue;
    long long w = grid_word(g, c.w, c.x, c.y, c.z);
    unsigned long long m = g.mask[w];
    int rank = g.prefix[w] + __popcll(m & ((1ull << (c.z & 63)) - 1));
    atomicMax(row_of_rank + rank, (int)i);  // duplicates: the later row owns the cell
  }
}

extern "C" int sgnn_grid_build(const SgnnGrid* g, const void* coords, int coords_i64, int64_t n,
                               int32_t* coords_i32_out, int32_t* status, void* scratch,
                               size_t scratch_bytes, void* stream) {
  if (!g || !g->mask || !g->prefix || n < 0 || (n > 0 && !coords)) return SGNN_E_INVALID;
  if (g->wx != (g->d2 + 63) / 64 ||
      g->n_words != (int64_t)g->nb * g->d0 * g->d1 * g->wx)
    return SGNN_E_INVALID;
  if (n > 0x7fffffffLL || g->n_words > 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  if (g->row_of_rank && !coords_i32_out && n > 0) return SGNN_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  GridView v = make_view(g);
  SGNN_CUDA(cudaMemsetAsync(g->mask, 0, (size_t)g->n_words * 8, st));
  if (n > 0) {
    grid_set_bits_kernel<<<sgnn_blocks(n, 256), 256, 0, st>>>(
        v, (unsigned long long*)g->mask, coords, coords_i64, (long long)n, coords_i32_out, status);
    SGNN_CHECK_LAUNCH();
  }
  int rc = sgnn_scan_exclusive(g->mask, SCAN_POPC64, g->prefix, g->n_words, scratch,
                               scratch_bytes, st);
  if (rc) return rc;
  if (g->row_of_rank && n > 0) {
    SGNN_CUDA(cudaMemsetAsync(g->row_of_rank, 0xff, (size_t)n * 4, st));
    grid_fill_rank_kernel<<<sgnn_blocks(n, 256), 256, 0, st>>>(v, coords_i32_out, (long long)n,
                                                               g->row_of_rank);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// ---------------------------------------------------------------- coarsen
// keep the even-position bits of (v | v>>1) and pack them into the low 32 bits
__device__ __forceinline__ unsigned pair_or_compact(unsigned long long v) {
  unsigned long long t = (v | (v >> 1)) & 0x5555555555555555ull;
  t = (t | (t >> 1)) & 0x3333333333333333ull;
  t = (t | (t >> 2)) & 0x0f0f0f0f0f0f0f0full;
  t = (t | (t >> 4)) & 0x00ff00ff00ff00ffull;
  t = (t | (t >> 8)) & 0x0000ffff0000ffffull;
  t = (t | (t >> 16)) & 0x00000000ffffffffull;
  return (unsigned)t;
}

__global__ void grid_coarsen_kernel(GridView f, GridView c, unsigned long long* __restrict__ cmask) {
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < c.n_words;
       w += (long long)gridDim.x * blockDim.x) {
    int wc = (int)(w % c.wx);
    long long r = w / c.wx;
    int Y = (int)(r % c.d1); r /= c.d1;
    int Z = (int)(r % c.d0);
    int b = (int)(r / c.d0);
    unsigned long long f0 = 0, f1 = 0;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        int z = 2 * Z + dz, y = 2 * Y + dy;
        if (z < f.d0 && y < f.d1) {
          long long fw = (((long long)b * f.d0 + z) * f.d1 + y) * f.wx + 2 * wc;
          f0 |= __ldg(f.mask + fw);
          if (2 * wc + 1 < f.wx) f1 |= __ldg(f.mask + fw + 1);
        }
      }
    unsigned long long out = (unsigned long long)pair_or_compact(f0) |
                             ((unsigned long long)pair_or_compact(f1) << 32);
    int rem = c.d2 - wc * 64;  // valid coarse cells in this word
    if (rem < 64) out &= (1ull << rem) - 1;
    cmask[w] = out;
  }
}

extern "C" int sgnn_grid_coarsen(const SgnnGrid* fine, const SgnnGrid* coarse, void* scratch,
                                 size_t scratch_bytes, void* stream) {
  if (!fine || !coarse || !fine->mask || !coarse->mask || !coarse->prefix) return SGNN_E_INVALID;
  if (coarse->nb != fine->nb || coarse->d0 > (fine->d0 + 1) / 2 || coarse->d1 > (fine->d1 + 1) / 2 ||
      coarse->d2 > (fine->d2 + 1) / 2 || coarse->wx != (coarse->d2 + 63) / 64 ||
      coarse->n_words != (int64_t)coarse->nb * coarse->d0 * coarse->d1 * coarse->wx)
    return SGNN_E_INVALID;
  if (coarse->n_words > 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (coarse->n_words > 0) {
    grid_coarsen_kernel<<<sgnn_blocks(coarse->n_words, 256), 256, 0, st>>>(
        make_view(fine), make_view(coarse), (unsigned long long*)coarse->mask);
    SGNN_CHECK_LAUNCH();
  }
  return sgnn_scan_exclusive(coarse->mask, SCAN_POPC64, coarse->prefix, coarse->n_words, scratch,
                             scratch_bytes, st);
}

// -------------------------------------------------------------- enumerate
__global__ void grid_enumerate_kernel(GridView g, int* __restrict__ coords_out) {
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < g.n_words;
       w += (long long)gridDim.x * blockDim.x) {
    unsigned long long m = __ldg(g.mask + w);
    if (!m) continue;
    int wc = (int)(w % g.wx);
    long long r = w / g.wx;
    int y = (int)(r % g.d1); r /= g.d1;
    int z = (int)(r % g.d0);
    int b = (int)(r / g.d0);
    int rank = __ldg(g.prefix + w);
    while (m) {
      int bit = __ffsll((long long)m) - 1;
      m &= m - 1;
      int row = g.row_of_rank ? g.row_of_rank[rank] : rank;
      if (row >= 0) reinterpret_cast<int4*>(coords_out)[row] = make_int4(z, y, wc * 64 + bit, b);
      ++rank;
    }
  }
}

extern "C" int sgnn_grid_enumerate(const SgnnGrid* g, int32_t* coords_out, void* stream) {
  if (!g || !g->mask || !g->prefix || !coords_out) return SGNN_E_INVALID;
  if (g->n_words > 0) {
    grid_enumerate_kernel<<<sgnn_blocks(g->n_words, 128), 128, 0, (cudaStream_t)stream>>>(
        make_view(g), coords_out);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// ------------------------------------------- coarse set, second half, in one kernel
// Coordinates of a raster-ordered coarse set (grid_enumerate) AND its filter-2 stride-2 rulebook against the fine set
// (rulebook_strided + the memset of its children table), from the coarse side: a thread walks the set bits of one byte of a
// coarse mask word; for each coarse site it writes the coordinates, probes its 8 fine cells (two x-neighbours share a fine mask word),
// writes children[k][row] for all 8 (-1 = absent: no memset needed) and parent[fine_row] = row * 8 + k for the present ones.
// Every fine site whose coordinates halve into the coarse extent is visited exactly once (fine rows are unique cells), so
// `parent` is fully written when the fine extents are even; otherwise the caller presets it to -1.
__global__ void __launch_bounds__(128)
grid_coarse_build_kernel(GridView c, GridView f, int* __restrict__ ccoords, int* __restrict__ parent,
                         int* __restrict__ children, long long n_coarse) {
  // a thread takes one BYTE of a coarse mask word (8 threads per word: the large levels have ~1 site per word, the dense
  // small ones up to 64 -- a whole word per thread left most of the chip idle there)
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < c.n_words * 8;
       t += (long long)gridDim.x * blockDim.x) {
    const long long w = t >> 3;
    const int g8 = (int)(t & 7) * 8;
    const unsigned long long mw = __ldg(c.mask + w);
    unsigned m = (unsigned)(mw >> g8) & 0xffu;
    if (!m) continue;
    const int wc = (int)(w % c.wx);
    long long r = w / c.wx;
    const int y = (int)(r % c.d1); r /= c.d1;
    const int z = (int)(r % c.d0);
    const int b = (int)(r / c.d0);
    int row = __ldg(c.prefix + w) + __popcll(mw & ((1ull << g8) - 1));
    while (m) {
      const int bit = g8 + __ffs((int)m) - 1;
      m &= m - 1;
      const int x = wc * 64 + bit;
      reinterpret_cast<int4*>(ccoords)[row] = make_int4(z, y, x, b);
#pragma unroll
      for (int kz = 0; kz < 2; ++kz)
#pragma unroll
        for (int ky = 0; ky < 2; ++ky) {
          const int fz = 2 * z + kz, fy = 2 * y + ky, fx = 2 * x;
          int r0 = -1, r1 = -1;
          if (fz < f.d0 && fy < f.d1 && fx < f.d2) {
            const long long fw = grid_word(f, b, fz, fy, fx);            // fx even: fx and fx + 1 sit in the same word
            const unsigned long long fm = __ldg(f.mask + fw);
            const unsigned two = (unsigned)(fm >> (fx & 63)) & 3u;
            if (two) {
              const int base = __ldg(f.prefix + fw) + __popcll(fm & ((1ull << (fx & 63)) - 1));
              if (two & 1) r0 = f.row_of_rank ? __ldg(f.row_of_rank + base) : base;
              if ((two & 2) && fx + 1 < f.d2) {
                const int rk = base + (two & 1);
                r1 = f.row_of_rank ? __ldg(f.row_of_rank + rk) : rk;
              }
            }
          }
          const int k = (kz << 2) | (ky << 1);
          children[(long long)k * n_coarse + row] = r0;
          children[(long long)(k + 1) * n_coarse + row] = r1;
          if (r0 >= 0) parent[r0] = row * 8 + k;
          if (r1 >= 0) parent[r1] = row * 8 + k + 1;
        }
      ++row;
    }
  }
}

extern "C" int sgnn_grid_coarse_build(const SgnnGrid* fine, const SgnnGrid* coarse, int64_t n_fine, int64_t n_coarse,
                                      int32_t* coarse_coords, int32_t* parent, int32_t* children, void* stream) {
  if (!fine || !coarse || !fine->mask || !fine->prefix || !coarse->mask || !coarse->prefix || n_fine < 0 || n_coarse < 0 ||
      coarse->row_of_rank || (n_fine > 0 && !parent) || (n_coarse > 0 && (!coarse_coords || !children)))
    return SGNN_E_INVALID;
  if (n_coarse > 0x0fffffffLL) return SGNN_E_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  // fine sites outside the coarse extent (odd fine extents: scn drops parents >= (S-2)/2+1) keep parent -1
  const bool covered = coarse->d0 * 2 >= fine->d0 && coarse->d1 * 2 >= fine->d1 && coarse->d2 * 2 >= fine->d2;
  if (n_fine > 0 && (!covered || n_coarse == 0)) SGNN_CUDA(cudaMemsetAsync(parent, 0xff, (size_t)n_fine * 4, st));
  if (n_coarse > 0) {
    grid_coarse_build_kernel<<<sgnn_blocks(coarse->n_words * 8, 128), 128, 0, st>>>(make_view(coarse), make_view(fine),
                                                                              coarse_coords, parent, children, (long long)n_coarse);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// ----------------------------------------------------------------- lookup
__global__ void grid_lookup_kernel(GridView g, const int* __restrict__ coords, long long n,
                                   int shift, int* __restrict__ rows) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
    int r = -1;
    if (c.x >= 0 && c.y >= 0 && c.z >= 0)
      r = grid_row_checked(g, c.w, c.x >> shift, c.y >> shift, c.z >> shift);
    rows[i] = r;
  }
}

extern "C" int sgnn_grid_lookup(const SgnnGrid* g, const int32_t* coords, int64_t n, int shift,
                                int32_t* rows_out, void* stream) {
  if (!g || !g->mask || !g->prefix || n < 0 || (n > 0 && (!coords || !rows_out)) || shift < 0 ||
      shift > 30)
    return SGNN_E_INVALID;
  if (n > 0) {
    grid_lookup_kernel<<<sgnn_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
        make_view(g), coords, (long long)n, shift, rows_out);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// --------------------------------------------------- submanifold rulebook
// One thread per output site.  The 27 probes of a site touch 9 x-rows; each row's 3 cells sit in
// one mask word (two when x is on a word edge), so the per-row word + prefix are loaded once and
// the three bits are tested in registers.  Stores are k-major: consecutive sites -> coalesced.
// COMPACT = false: dense k-major table nbr[k][i] (-1 = absent).  COMPACT = true: only the PRESENT offsets of a site, in
// ascending k (the centre included), packed (k << 27 | row) into slots[s][i], s < cnt[i] -- for site sets whose rows have few
// neighbours (the 5 %-occupancy encoder input: 2.3 of 27) the dense table is 108 B/site of which 9 B are rules.
template <bool COMPACT>
__global__ void __launch_bounds__(256)
rulebook_submanifold_kernel(GridView g, const int* __restrict__ coords, long long n,
                            int* __restrict__ nbr, unsigned char* __restrict__ cnt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + i);
    int idx[27];
    rulebook_probe27(g, c, idx);
    int filled = 0;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      if (COMPACT) {
        if (idx[k] >= 0) { nbr[(long long)filled * n + i] = (int)(((unsigned)k << 27) | (unsigned)idx[k]); ++filled; }
      } else {
        nbr[(long long)k * n + i] = idx[k];
      }
    }
    if (COMPACT) cnt[i] = (unsigned char)filled;
  }
}

extern "C" int sgnn_rulebook_submanifold(const SgnnGrid* g, const int32_t* coords, int64_t n,
                                         int32_t* nbr, void* stream) {
  if (!g || !g->mask || !g->prefix || n < 0 || (n > 0 && (!coords || !nbr))) return SGNN_E_INVALID;
  if (n * 27 > 0x7fffffff00LL) return SGNN_E_TOO_LARGE;
  if (n > 0) {
    rulebook_submanifold_kernel<false><<<sgnn_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
        make_view(g), coords, (long long)n, nbr, nullptr);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

extern "C" int sgnn_rulebook_submanifold_compact(const SgnnGrid* g, const int32_t* coords, int64_t n,
                                                 int32_t* slots, uint8_t* cnt, void* stream) {
  if (!g || !g->mask || !g->prefix || n < 0 || (n > 0 && (!coords || !slots || !cnt))) return SGNN_E_INVALID;
  if (n >= (1LL << 27)) return SGNN_E_TOO_LARGE;          // row index shares a word with the 5-bit offset id
  if (n > 0) {
    rulebook_submanifold_kernel<true><<<sgnn_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
        make_view(g), coords, (long long)n, slots, cnt);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// -------------------------------------------------------- strided rulebook
__global__ void rulebook_strided_kernel(GridView c, const int* __restrict__ fine_coords,
                                        long long n_fine, int* __restrict__ parent,
                                        int* __restrict__ children, long long n_coarse) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_fine;
       i += (long long)gridDim.x * blockDim.x) {
    int4 f = __ldg(reinterpret_cast<const int4*>(fine_coords) + i);
    int p = -1;
    if (f.x >= 0 && f.y >= 0 && f.z >= 0) {
      int row = grid_row_checked(c, f.w, f.x >> 1, f.y >> 1, f.z >> 1);
      if (row >= 0) {
        int k = ((f.x & 1) << 2) | ((f.y & 1) << 1) | (f.z & 1);
        p = row * 8 + k;
        if (children && row < n_coarse) children[(long long)k * n_coarse + row] = (int)i;
      }
    }
    parent[i] = p;
  }
}

extern "C" int sgnn_rulebook_strided(const SgnnGrid* coarse, const int32_t* fine_coords,
                                     int64_t n_fine, int32_t* parent, int32_t* children,
                                     int64_t n_coarse, void* stream) {
  if (!coarse || !coarse->mask || !coarse->prefix || n_fine < 0 || n_coarse < 0 ||
      (n_fine > 0 && (!fine_coords || !parent)))
    return SGNN_E_INVALID;
  if (n_coarse > 0x0fffffffLL) return SGNN_E_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (children && n_coarse > 0)
    SGNN_CUDA(cudaMemsetAsync(children, 0xff, (size_t)n_coarse * 8 * 4, st));
  if (n_fine > 0) {
    rulebook_strided_kernel<<<sgnn_blocks(n_fine, 256), 256, 0, st>>>(
        make_view(coarse), fine_coords, (long long)n_fine, parent, children, (long long)n_coarse);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// --------------------------------------------------------------- concat_skip
// model.py:338-355 without the two dense int64 indicator volumes: one grid probe per target row.
__global__ void concat_skip_kernel(GridView g, const float* __restrict__ src, int ld_src, int c,
                                   const int* __restrict__ coords, long long n,
                                   float* __restrict__ dst, int ld_dst, int col0) {
  // c threads (rounded to a power-of-two group) per row
  int gsz = 1;
  while (gsz < c && gsz < 32) gsz <<= 1;
  const int per_blk = blockDim.x / gsz;
  const int sub = threadIdx.x / gsz, ch = threadIdx.x % gsz;
  for (long long i = (long long)blockIdx.x * per_blk + sub; i < n;
       i += (long long)gridDim.x * per_blk) {
    int4 p = __ldg(reinterpret_cast<const int4*>(coords) + i);
    int row = -1;
    if (p.x >= 0 && p.y >= 0 && p.z >= 0) row = grid_row_checked(g, p.w, p.x, p.y, p.z);
    for (int j = ch; j < c; j += gsz)
      dst[i * ld_dst + col0 + j] = row >= 0 ? __ldg(src + (long long)row * ld_src + j) : 0.0f;
  }
}

extern "C" int sgnn_concat_skip(const SgnnGrid* g, const float* src, int32_t ld_src, int32_t c,
                                const int32_t* coords, int64_t n, float* dst, int32_t ld_dst,
                                int32_t col0, void* stream) {
  if (!g || !g->mask || !g->prefix || n < 0 || c <= 0 || col0 < 0 || col0 + c > ld_dst ||
      (n > 0 && (!coords || !dst)))
    return SGNN_E_INVALID;
  if (n > 0) {
    int gsz = 1;
    while (gsz < c && gsz < 32) gsz <<= 1;
    int per_blk = 256 / gsz;
    concat_skip_kernel<<<sgnn_blocks(n, per_blk), 256, 0, (cudaStream_t)stream>>>(
        make_view(g), src, ld_src, c, coords, (long long)n, dst, ld_dst, col0);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// ------------------------------------------------------------ misc
__global__ void coords_to_i64_kernel(const int* __restrict__ in, long long n4,
                                     long long* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i];
}

extern "C" int sgnn_coords_to_i64(const int32_t* in, int64_t n, int64_t* out, void* stream) {
  if (n < 0 || (n > 0 && (!in || !out))) return SGNN_E_INVALID;
  if (n > 0) {
    coords_to_i64_kernel<<<sgnn_blocks(n * 4, 256), 256, 0, (cudaStream_t)stream>>>(
        in, (long long)n * 4, (long long*)out);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}
