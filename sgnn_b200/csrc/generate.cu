// generate.cu -- the generative (coarse-to-fine) steps of SG-NN as device kernels:
// occupancy heads, the literal `sigmoid(x) > 0.5` mask, and STABLE compaction of the kept candidates
// (flags -> exclusive scan -> ordered write), replacing the host-synchronising boolean-mask indexing of
// model.py:238-246 and :322-335.  Candidate order is preserved so output rows match the reference order.
#include "common.cuh"

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t sgnn_compact_scratch_bytes(int64_t n_items) {
  int64_t n = n_items < 1 ? 1 : n_items;
  return align256((size_t)n) + align256((size_t)(n + 1) * 4) + sgnn_scan_scratch_bytes(n);
}

struct CompactScratch {
  unsigned char* flags;
  int* offs;
  void* scan;
  size_t scan_bytes;
};

static int carve(void* scratch, size_t bytes, int64_t n, CompactScratch* cs) {
  if (!scratch || bytes < sgnn_compact_scratch_bytes(n)) return SGNN_E_INVALID;
  int64_t m = n < 1 ? 1 : n;
  char* p = (char*)scratch;
  cs->flags = (unsigned char*)p; p += align256((size_t)m);
  cs->offs = (int*)p; p += align256((size_t)(m + 1) * 4);
  cs->scan = p;
  cs->scan_bytes = bytes - (size_t)(p - (char*)scratch);
  return SGNN_OK;
}

// ------------------------------------------------------------ a8: dense_coarse_to_sparse
__global__ void d2s_flag_kernel(const float* __restrict__ dense_out, int nb, long long vol,
                                unsigned char* __restrict__ flags, float* __restrict__ cand_out) {
  const long long total = (long long)nb * vol;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / vol, cell = i % vol;
    const float occ = dense_out[(b * 2 + 0) * vol + cell];
    const float sdf = dense_out[(b * 2 + 1) * vol + cell];
    flags[i] = sigmoid_gt_half(occ) ? 1 : 0;
    if (cand_out) reinterpret_cast<float2*>(cand_out)[i] = make_float2(occ, sdf);
  }
}

__global__ void d2s_write_kernel(const float* __restrict__ dense_feats, const float* __restrict__ dense_out, int nb,
                                 int c, int d0, int d1, int d2, const unsigned char* __restrict__ flags,
                                 const int* __restrict__ offs, int* __restrict__ locs, float* __restrict__ feats,
                                 int ld, int* __restrict__ count) {
  const long long vol = (long long)d0 * d1 * d2;
  const long long total = (long long)nb * vol;
  if (count && blockIdx.x == 0 && threadIdx.x == 0) *count = offs[total];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (!flags[i]) continue;
    const long long b = i / vol, cell = i % vol;
    const int x = (int)(cell % d2), y = (int)((cell / d2) % d1), z = (int)(cell / ((long long)d1 * d2));
    const int pos = offs[i];
    reinterpret_cast<int4*>(locs)[pos] = make_int4(z, y, x, (int)b);
    float* f = feats + (long long)pos * ld;
    f[0] = dense_out[(b * 2 + 0) * vol + cell];
    f[1] = dense_out[(b * 2 + 1) * vol + cell];
    for (int ch = 0; ch < c; ++ch) f[2 + ch] = dense_feats[(b * c + ch) * vol + cell];
    for (int ch = c + 2; ch < ld; ++ch) f[ch] = 0.f;      // spare columns for the skip join
  }
}

extern "C" int sgnn_dense_to_sparse(const float* dense_feats, const float* dense_out, int32_t nb, int32_t c,
                                    int32_t d0, int32_t d1, int32_t d2, int32_t* locs, float* feats,
                                    int32_t ld_feats, float* cand_out, int32_t* count, void* scratch,
                                    size_t scratch_bytes, void* stream) {
  if (nb < 0 || c < 0 || d0 < 0 || d1 < 0 || d2 < 0 || !count || ld_feats < c + 2) return SGNN_E_INVALID;
  const long long total = (long long)nb * d0 * d1 * d2;
  if (total > 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (total == 0) {
    SGNN_CUDA(cudaMemsetAsync(count, 0, 4, st));
    return SGNN_OK;
  }
  if (!dense_feats || !dense_out || !locs || !feats) return SGNN_E_INVALID;
  CompactScratch cs;
  int rc = carve(scratch, scratch_bytes, total, &cs);
  if (rc) return rc;
  d2s_flag_kernel<<<sgnn_blocks(total, 256), 256, 0, st>>>(dense_out, nb, (long long)d0 * d1 * d2, cs.flags,
                                                           cand_out);
  SGNN_CHECK_LAUNCH();
  rc = sgnn_scan_exclusive(cs.flags, SCAN_U8, cs.offs, total, cs.scan, cs.scan_bytes, st);
  if (rc) return rc;
  d2s_write_kernel<<<sgnn_blocks(total, 256), 256, 0, st>>>(dense_feats, dense_out, nb, c, d0, d1, d2, cs.flags,
                                                            cs.offs, locs, feats, ld_feats, count);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// ------------------------------------------------------------ a9: heads + mask + compaction
__global__ void heads_flag_kernel(const float* __restrict__ x, int ld_x, int c, const float* __restrict__ w_occ,
                                  const float* __restrict__ b_occ, const float* __restrict__ w_sdf,
                                  const float* __restrict__ b_sdf, long long n_cand,
                                  unsigned char* __restrict__ flags, float* __restrict__ cand_out) {
  const bool vec4 = (c & 3) == 0 && (ld_x & 3) == 0 && ((size_t)x & 15) == 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_cand;
       i += (long long)gridDim.x * blockDim.x) {
    const float* xr = x + i * ld_x;
    float occ = 0.f, sdf = 0.f;
    if (vec4) {                       // 16-byte row loads (c, ld_x multiples of 4, x 16-byte aligned); same fmaf order
      for (int ch = 0; ch < c; ch += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + ch);
        occ = fmaf(v.x, __ldg(w_occ + ch), occ);         sdf = fmaf(v.x, __ldg(w_sdf + ch), sdf);
        occ = fmaf(v.y, __ldg(w_occ + ch + 1), occ);     sdf = fmaf(v.y, __ldg(w_sdf + ch + 1), sdf);
        occ = fmaf(v.z, __ldg(w_occ + ch + 2), occ);     sdf = fmaf(v.z, __ldg(w_sdf + ch + 2), sdf);
        occ = fmaf(v.w, __ldg(w_occ + ch + 3), occ);     sdf = fmaf(v.w, __ldg(w_sdf + ch + 3), sdf);
      }
    } else {
      for (int ch = 0; ch < c; ++ch) {
        const float v = xr[ch];
        occ = fmaf(v, __ldg(w_occ + ch), occ);
        sdf = fmaf(v, __ldg(w_sdf + ch), sdf);
      }
    }
    occ += __ldg(b_occ);
    sdf += __ldg(b_sdf);
    flags[i] = sigmoid_gt_half(occ) ? 1 : 0;
    reinterpret_cast<float2*>(cand_out)[i] = make_float2(occ, sdf);
  }
}

// 8 lanes per candidate row: the kept row (c features + occ + sdf + zeroed spare columns) is written as 16-byte /
// 4-byte pieces by neighbouring lanes instead of one thread striding through a 112-byte row.
// JOIN: the skip join of the next level (concat_skip, model.py:338-355: the features of the encoder site at the kept
// child's coordinates, zeros where there is none) is written with the row -- columns [c+2, c+2+c_skip) -- instead of by a
// second kernel that re-reads every coordinate and re-walks every row (sgnn_concat_skip).
struct SkipJoin { GridView g; const float* f; int ld, c; };
template <bool JOIN>
__global__ void heads_write_kernel(const float* __restrict__ x, int ld_x, int c, const float* __restrict__ cand_out,
                                   const int* __restrict__ parent_coords, long long n_cand,
                                   const unsigned char* __restrict__ flags, const int* __restrict__ offs,
                                   int* __restrict__ locs, float* __restrict__ feats, int ld,
                                   int* __restrict__ count, SkipJoin sk) {
  if (count && blockIdx.x == 0 && threadIdx.x == 0) *count = offs[n_cand];
  const int sub = threadIdx.x & 7;
  const bool st4 = (ld & 3) == 0 && ((size_t)feats & 15) == 0;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3; i < n_cand;
       i += ((long long)gridDim.x * blockDim.x) >> 3) {
    if (!flags[i]) continue;
    const int pos = offs[i];
    const int4 p = __ldg(reinterpret_cast<const int4*>(parent_coords) + (i >> 3));
    const int ch8 = (int)(i & 7);
    const int4 q = make_int4(2 * p.x + ((ch8 >> 2) & 1), 2 * p.y + ((ch8 >> 1) & 1), 2 * p.z + (ch8 & 1), p.w);
    if (sub == 0) reinterpret_cast<int4*>(locs)[pos] = q;
    const float* srow = nullptr;
    if (JOIN) {
      const int r = grid_row_checked(sk.g, q.w, q.x, q.y, q.z);      // 8 lanes, one address: a broadcast
      if (r >= 0) srow = sk.f + (long long)r * sk.ld;
    }
    float* f = feats + (long long)pos * ld;
    const float* xr = x + i * ld_x;
    auto value = [&](int ch) -> float {
      if (ch < c) return xr[ch];
      if (ch == c) return cand_out[2 * i];
      if (ch == c + 1) return cand_out[2 * i + 1];
      if (JOIN && srow && ch < c + 2 + sk.c) return __ldg(srow + (ch - c - 2));
      return 0.f;                                            // spare columns (skip join of the next level / padding)
    };
    if (st4) {                                               // rows of ld % 4 == 0 floats, 16-byte aligned: one float4 per lane
      for (int q = sub * 4; q < ld; q += 32)
        *reinterpret_cast<float4*>(f + q) = make_float4(value(q), value(q + 1), value(q + 2), value(q + 3));
    } else {
      for (int ch = sub; ch < ld; ch += 8) f[ch] = value(ch);
    }
  }
}

extern "C" int sgnn_heads_compact(const float* x, int32_t ld_x, int32_t c, const float* w_occ, const float* b_occ,
                                  const float* w_sdf, const float* b_sdf, const int32_t* parent_coords,
                                  int64_t n_parent, float* cand_out, int32_t* locs, float* feats,
                                  int32_t ld_feats, int32_t* count, void* scratch, size_t scratch_bytes,
                                  void* stream) {
  if (n_parent < 0 || c <= 0 || !count || ld_feats < c + 2 || !w_occ || !b_occ || !w_sdf || !b_sdf)
    return SGNN_E_INVALID;
  const long long n_cand = (long long)n_parent * 8;
  if (n_cand > 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_cand == 0) {
    SGNN_CUDA(cudaMemsetAsync(count, 0, 4, st));
    return SGNN_OK;
  }
  if (!x || !parent_coords || !cand_out || !locs || !feats) return SGNN_E_INVALID;
  CompactScratch cs;
  int rc = carve(scratch, scratch_bytes, n_cand, &cs);
  if (rc) return rc;
  heads_flag_kernel<<<sgnn_blocks(n_cand, 256), 256, 0, st>>>(x, ld_x, c, w_occ, b_occ, w_sdf, b_sdf, n_cand,
                                                              cs.flags, cand_out);
  SGNN_CHECK_LAUNCH();
  rc = sgnn_scan_exclusive(cs.flags, SCAN_U8, cs.offs, n_cand, cs.scan, cs.scan_bytes, st);
  if (rc) return rc;
  heads_write_kernel<false><<<sgnn_blocks(n_cand * 8, 256), 256, 0, st>>>(x, ld_x, c, cand_out, parent_coords, n_cand,
                                                                      cs.flags, cs.offs, locs, feats, ld_feats, count, SkipJoin());
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}


// ------------------------------------------------------------ two-phase variants (count known before the write)
// Phase 1 evaluates the heads / mask and scans; offs[n] is the kept count.  The caller reads it (one 4-byte D2H),
// allocates exactly that many rows and calls phase 2.  Used by the native generator (generator.cu).
extern "C" int sgnn_heads_flags(const float* x, int32_t ld_x, int32_t c, const float* w_occ, const float* b_occ,
                                const float* w_sdf, const float* b_sdf, int64_t n_cand, float* cand_out,
                                uint8_t* flags, int32_t* offs, void* scratch, size_t scratch_bytes, void* stream) {
  if (n_cand < 0 || c <= 0 || !offs || !w_occ || !b_occ || !w_sdf || !b_sdf) return SGNN_E_INVALID;
  if (n_cand > 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_cand > 0) {
    if (!x || !cand_out || !flags) return SGNN_E_INVALID;
    heads_flag_kernel<<<sgnn_blocks(n_cand, 256), 256, 0, st>>>(x, ld_x, c, w_occ, b_occ, w_sdf, b_sdf, n_cand, flags,
                                                                cand_out);
    SGNN_CHECK_LAUNCH();
  }
  return sgnn_scan_exclusive(flags, SCAN_U8, offs, n_cand, scratch, scratch_bytes, st);
}

extern "C" int sgnn_heads_write(const float* x, int32_t ld_x, int32_t c, const float* cand_out,
                                const int32_t* parent_coords, int64_t n_cand, const uint8_t* flags,
                                const int32_t* offs, int32_t* locs, float* feats, int32_t ld_feats, void* stream) {
  if (n_cand < 0 || c <= 0 || ld_feats < c + 2) return SGNN_E_INVALID;
  if (n_cand == 0) return SGNN_OK;
  if (!x || !cand_out || !parent_coords || !flags || !offs || !locs || !feats) return SGNN_E_INVALID;
  heads_write_kernel<false><<<sgnn_blocks(n_cand * 8, 256), 256, 0, (cudaStream_t)stream>>>(
      x, ld_x, c, cand_out, parent_coords, n_cand, flags, offs, locs, feats, ld_feats, nullptr, SkipJoin());
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_heads_write_join(const float* x, int32_t ld_x, int32_t c, const float* cand_out,
                                     const int32_t* parent_coords, int64_t n_cand, const uint8_t* flags,
                                     const int32_t* offs, int32_t* locs, float* feats, int32_t ld_feats,
                                     const SgnnGrid* skip_grid, const float* skip_feats, int32_t ld_skip, int32_t c_skip,
                                     void* stream) {
  if (n_cand < 0 || c <= 0 || c_skip <= 0 || ld_feats < c + 2 + c_skip) return SGNN_E_INVALID;
  if (n_cand == 0) return SGNN_OK;
  if (!x || !cand_out || !parent_coords || !flags || !offs || !locs || !feats) return SGNN_E_INVALID;
  if (!skip_grid || !skip_grid->mask || !skip_grid->prefix || !skip_feats) return SGNN_E_INVALID;
  SkipJoin sk;
  sk.g = make_view(skip_grid); sk.f = skip_feats; sk.ld = ld_skip; sk.c = c_skip;
  heads_write_kernel<true><<<sgnn_blocks(n_cand * 8, 256), 256, 0, (cudaStream_t)stream>>>(
      x, ld_x, c, cand_out, parent_coords, n_cand, flags, offs, locs, feats, ld_feats, nullptr, sk);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_dense_flags(const float* dense_out, int32_t nb, int64_t vol, float* cand_out, uint8_t* flags,
                                int32_t* offs, void* scratch, size_t scratch_bytes, void* stream) {
  if (nb < 0 || vol < 0 || !offs) return SGNN_E_INVALID;
  const long long total = (long long)nb * vol;
  if (total > 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (total > 0) {
    if (!dense_out || !flags) return SGNN_E_INVALID;
    d2s_flag_kernel<<<sgnn_blocks(total, 256), 256, 0, st>>>(dense_out, nb, vol, flags, cand_out);
    SGNN_CHECK_LAUNCH();
  }
  return sgnn_scan_exclusive(flags, SCAN_U8, offs, total, scratch, scratch_bytes, st);
}

extern "C" int sgnn_dense_write(const float* dense_feats, const float* dense_out, int32_t nb, int32_t c, int32_t d0,
                                int32_t d1, int32_t d2, const uint8_t* flags, const int32_t* offs, int32_t* locs,
                                float* feats, int32_t ld_feats, void* stream) {
  if (nb < 0 || c < 0 || ld_feats < c + 2) return SGNN_E_INVALID;
  const long long total = (long long)nb * d0 * d1 * d2;
  if (total == 0) return SGNN_OK;
  if (!dense_feats || !dense_out || !flags || !offs || !locs || !feats) return SGNN_E_INVALID;
  d2s_write_kernel<<<sgnn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(dense_feats, dense_out, nb, c, d0, d1, d2,
                                                                             flags, offs, locs, feats, ld_feats,
                                                                             nullptr);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// ------------------------------------------------------------ candidate coordinates (model.py:192-207)
__global__ void children_coords_kernel(const int* __restrict__ parent_coords, long long n_cand,
                                       int* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_cand;
       i += (long long)gridDim.x * blockDim.x) {
    const int4 p = __ldg(reinterpret_cast<const int4*>(parent_coords) + (i >> 3));
    const int c = (int)(i & 7);
    reinterpret_cast<int4*>(out)[i] =
        make_int4(2 * p.x + ((c >> 2) & 1), 2 * p.y + ((c >> 1) & 1), 2 * p.z + (c & 1), p.w);
  }
}

extern "C" int sgnn_children_coords(const int32_t* parent_coords, int64_t n_parent, int32_t* out, void* stream) {
  if (n_parent < 0 || (n_parent > 0 && (!parent_coords || !out))) return SGNN_E_INVALID;
  if (n_parent * 8 > 0x7fffffffLL) return SGNN_E_TOO_LARGE;
  if (n_parent > 0) {
    children_coords_kernel<<<sgnn_blocks(n_parent * 8, 256), 256, 0, (cudaStream_t)stream>>>(
        parent_coords, (long long)n_parent * 8, out);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}

// ------------------------------------------------------------ ABI bookkeeping
int g_sgnn_last_cuda_error = 0;
long long g_sgnn_launches = 0;

extern "C" int64_t sgnn_launch_count(void) { return (int64_t)g_sgnn_launches; }

extern "C" int sgnn_version(void) { return SGNN_VERSION; }

extern "C" int sgnn_last_cuda_error(void) { return g_sgnn_last_cuda_error; }

extern "C" const char* sgnn_error_string(int code) {
  switch (code) {
    case SGNN_OK: return "ok";
    case SGNN_E_INVALID: return "invalid argument";
    case SGNN_E_CUDA: return "CUDA runtime error (see sgnn_last_cuda_error)";
    case SGNN_E_TOO_LARGE: return "extent or row count exceeds 2^31-1 indexing";
    case SGNN_E_UNSUPPORTED: return "unsupported configuration";
    case SGNN_E_ALIGN: return "pointer or leading dimension violates the 16-byte alignment rule";
    case SGNN_E_NOMEM: return "workspace arena too small (SgnnGeneratorOut.arena_needed has the size to retry with)";
    default: return "unknown sgnn error code";
  }
}


// ------------------------------------------------------------ output export (one launch for every result tensor)
// The generator's results live in its arena (valid until the next pass); the reference returns fresh tensors and int64
// coordinates (model.py:247,336,380,416).  Formatting them tensor by tensor from the host mirror cost ~50 framework calls
// and ~0.9 ms of host time per pass -- more than the GPU had queued, so the device idled.  One kernel walks a list of
// segments instead.
struct ExportSegs { SgnnExportSeg s[SGNN_EXPORT_MAX_SEGS]; long long start[SGNN_EXPORT_MAX_SEGS + 1]; int n; };

__device__ __forceinline__ void export_store4(void* dst, long long row, int to64, int z, int y, int x, int b) {
  if (to64) {
    longlong2* d = reinterpret_cast<longlong2*>(dst) + 2 * row;
    d[0] = make_longlong2(z, y);
    d[1] = make_longlong2(x, b);
  } else {
    reinterpret_cast<int4*>(dst)[row] = make_int4(z, y, x, b);
  }
}

__global__ void __launch_bounds__(256)
export_kernel(ExportSegs e) {
  const long long total = e.start[e.n];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int k = 0;
    while (k + 1 < e.n && t >= e.start[k + 1]) ++k;
    const SgnnExportSeg& g = e.s[k];
    const long long i = t - e.start[k];
    switch (g.kind) {
      case SGNN_EXPORT_COPY32:
        reinterpret_cast<int*>(g.dst)[i] = reinterpret_cast<const int*>(g.src)[i];
        break;
      case SGNN_EXPORT_COORDS: {          // [n,4] int32 rows -> int32 / int64 rows
        const int4 c = __ldg(reinterpret_cast<const int4*>(g.src) + i);
        export_store4(g.dst, i, g.to_i64, c.x, c.y, c.z, c.w);
        break;
      }
      case SGNN_EXPORT_CHILDREN: {        // row i = child (i & 7) of parent row i >> 3 (model.py:192-207)
        const int4 p = __ldg(reinterpret_cast<const int4*>(g.src) + (i >> 3));
        const int c = (int)(i & 7);
        export_store4(g.dst, i, g.to_i64, 2 * p.x + ((c >> 2) & 1), 2 * p.y + ((c >> 1) & 1), 2 * p.z + (c & 1), p.w);
        break;
      }
      default: {                          // SGNN_EXPORT_DENSE_CELLS: all cells of [nb, d0, d1, d2], batch-major raster (model.py:319-321)
        const int d0 = g.aux[1], d1 = g.aux[2], d2 = g.aux[3];
        const int x = (int)(i % d2), y = (int)((i / d2) % d1), z = (int)((i / ((long long)d1 * d2)) % d0);
        const int b = (int)(i / ((long long)d0 * d1 * d2));
        export_store4(g.dst, i, g.to_i64, z, y, x, b);
      }
    }
  }
}

extern "C" int sgnn_export(const SgnnExportSeg* segs, int32_t n_segs, void* stream) {
  if (n_segs < 0 || n_segs > SGNN_EXPORT_MAX_SEGS || (n_segs > 0 && !segs)) return SGNN_E_INVALID;
  ExportSegs e;
  e.n = 0;
  e.start[0] = 0;
  for (int i = 0; i < n_segs; ++i) {
    const SgnnExportSeg& g = segs[i];
    if (g.n < 0 || g.kind < 0 || g.kind > SGNN_EXPORT_DENSE_CELLS) return SGNN_E_INVALID;
    if (g.n == 0) continue;
    if (!g.dst || (g.kind != SGNN_EXPORT_DENSE_CELLS && !g.src)) return SGNN_E_INVALID;
    if (g.kind == SGNN_EXPORT_DENSE_CELLS && (g.aux[1] <= 0 || g.aux[2] <= 0 || g.aux[3] <= 0)) return SGNN_E_INVALID;
    e.s[e.n] = g;
    e.start[e.n + 1] = e.start[e.n] + g.n;
    ++e.n;
  }
  if (e.n == 0) return SGNN_OK;
  export_kernel<<<sgnn_blocks(e.start[e.n], 256), 256, 0, (cudaStream_t)stream>>>(e);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}
