// conv.cu -- sparse convolution forward, output-stationary gather -> small-GEMM (SURVEY §8 rows a3, a4,
// a6, a9).  fp32 FFMA path with a SPECIFIED summation order (k ascending, then ci ascending, one fmaf
// chain per output element starting from +0) so that results are bit-reproducible and equal to
// oracle/o3.c.  No scatter and no atomics: the rulebook is a neighbour table nbr[k][row].
//
// Tile: SGNN_CONV_TILE output rows x Cout per CTA.  Thread (sg, cg) owns rows {sg + 32 t, t<4} and output
// channels [4cg, 4cg+4): 16 accumulators.  Per stage (filter offset k, 16-channel slice of Cin) the CTA
// gathers the 128 neighbour row slices into shared memory with cp.async (16-byte chunks, zero rows for
// absent neighbours), double buffered against the FFMA loop.  The filter bank [K][Cin][Cout] is staged in
// shared memory once per CTA.  Row stride of the X tile is 20 words so the 8 distinct rows a warp reads with
// one LDS.128 fall on disjoint banks.
#include "common.cuh"

#define TM SGNN_CONV_TILE
#define TS 4
#define NSG (TM / TS)  // 32 site groups
#define CHUNK 16       // input channels per stage
#define XS (CHUNK + 4) // X tile row stride (floats)


struct ConvParams {
  const float* in;
  int ld_in;
  const int* nbr;
  long long nbr_stride;
  int K;
  int child_mode;
  const float* weight;
  int cin, cout;
  long long n_out;
  const float* residual;
  int ld_res;
  float* out_a; int ld_a; int relu_a; const float* scale_a; const float* shift_a;
  float* out_b; int ld_b; int relu_b; const float* scale_b; const float* shift_b;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// neighbour row feeding output row j at filter offset k
__device__ __forceinline__ int conv_src_row(const ConvParams& p, long long j, int k) {
  if (!p.child_mode) return __ldg(p.nbr + (long long)k * p.nbr_stride + j);
  // child c of parent row j>>3; offset (dz,dy,dx) lands in parent offset floor((c_a + d_a) / 2) per axis
  const int c = (int)(j & 7);
  const int dz = k / 9 - 1, dy = (k / 3) % 3 - 1, dx = k % 3 - 1;
  const int pz = (((c >> 2) & 1) + dz + 2) / 2 - 1;
  const int py = (((c >> 1) & 1) + dy + 2) / 2 - 1;
  const int px = ((c & 1) + dx + 2) / 2 - 1;
  const int kp = (pz + 1) * 9 + (py + 1) * 3 + (px + 1);
  return __ldg(p.nbr + (long long)kp * p.nbr_stride + (j >> 3));
}

// Shared epilogue: residual add, two output slots with optional affine + relu (see SgnnEpilogue).
template <int COUT>
__device__ __forceinline__ void conv_epilogue_rows(const ConvParams& p, const float (&acc)[TS][4],
                                                   const long long (&rows)[TS], int cg);

template <int COUT>
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, const float (&acc)[TS][4], long long tile_base,
                                              int sg, int cg) {
  long long rows[TS];
#pragma unroll
  for (int t = 0; t < TS; ++t) rows[t] = tile_base + sg + NSG * t;
  conv_epilogue_rows<COUT>(p, acc, rows, cg);
}

template <int COUT>
__device__ __forceinline__ void conv_epilogue_rows(const ConvParams& p, const float (&acc)[TS][4],
                                                   const long long (&rows)[TS], int cg) {
  float4 sa = make_float4(1.f, 1.f, 1.f, 1.f), ta = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa, tb = ta;
  if (p.out_a && p.scale_a) {
    sa = __ldg(reinterpret_cast<const float4*>(p.scale_a) + cg);
    ta = __ldg(reinterpret_cast<const float4*>(p.shift_a) + cg);
  }
  if (p.out_b && p.scale_b) {
    sb = __ldg(reinterpret_cast<const float4*>(p.scale_b) + cg);
    tb = __ldg(reinterpret_cast<const float4*>(p.shift_b) + cg);
  }
#pragma unroll
  for (int t = 0; t < TS; ++t) {
    const long long j = rows[t];
    if (j >= p.n_out) continue;
    float4 v = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    if (p.residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + j * p.ld_res) + cg);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (p.out_a) {
      float4 y = v;
      if (p.scale_a) {
        y.x = fmaf(v.x, sa.x, ta.x); y.y = fmaf(v.y, sa.y, ta.y);
        y.z = fmaf(v.z, sa.z, ta.z); y.w = fmaf(v.w, sa.w, ta.w);
      }
      if (p.relu_a) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
      reinterpret_cast<float4*>(p.out_a + j * p.ld_a)[cg] = y;
    }
    if (p.out_b) {
      float4 y = v;
      if (p.scale_b) {
        y.x = fmaf(v.x, sb.x, tb.x); y.y = fmaf(v.y, sb.y, tb.y);
        y.z = fmaf(v.z, sb.z, tb.z); y.w = fmaf(v.w, sb.w, tb.w);
      }
      if (p.relu_b) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
      reinterpret_cast<float4*>(p.out_b + j * p.ld_b)[cg] = y;
    }
  }
}

template <int COUT, int VEC>
__global__ void __launch_bounds__(NSG * (COUT / 4))
conv_gather_f32_kernel(ConvParams p) {
  constexpr int NCG = COUT / 4;
  constexpr int NT = NSG * NCG;
  extern __shared__ __align__(16) float smem[];
  const int cinp = (p.cin + 3) & ~3;
  float* W_s = smem;                            // [K][cinp][COUT]
  float* X_s = smem + (size_t)p.K * cinp * COUT;  // [2][TM][XS]
  const int tid = threadIdx.x;
  const int sg = tid / NCG, cg = tid % NCG;
  const long long tile_base = (long long)blockIdx.x * TM;
  const int nchunks = (cinp + CHUNK - 1) / CHUNK;
  const int S = p.K * nchunks;

  // ---- stage the filter bank (zero rows for the channel padding)
  for (int idx = tid; idx < p.K * cinp * NCG; idx += NT) {
    int c4 = idx % NCG;
    int ci = (idx / NCG) % cinp;
    int k = idx / (NCG * cinp);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ci < p.cin)
      w = __ldg(reinterpret_cast<const float4*>(p.weight + ((size_t)k * p.cin + ci) * COUT) + c4);
    reinterpret_cast<float4*>(W_s + ((size_t)k * cinp + ci) * COUT)[c4] = w;
  }

  float acc[TS][4];
#pragma unroll
  for (int t = 0; t < TS; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;

  // gather of stage s into buffer s&1; returns whether this thread copied any live row
  auto issue = [&](int s) -> int {
    const int k = s / nchunks, c0 = (s % nchunks) * CHUNK;
    const int cw = min(CHUNK, cinp - c0);  // multiple of 4
    float* X = X_s + (size_t)(s & 1) * TM * XS;
    int live = 0;
    if (VEC == 4) {
      const int cpr = cw >> 2;  // 16-byte chunks per row
      for (int idx = tid; idx < TM * cpr; idx += NT) {
        const int site = idx / cpr, ch = idx % cpr;
        const long long j = tile_base + site;
        int r = -1;
        if (j < p.n_out) r = conv_src_row(p, j, k);
        float* dst = X + site * XS + ch * 4;
        const int cb = c0 + ch * 4;
        if (r >= 0) {
          live = 1;
          const float* src = p.in + (long long)r * p.ld_in + cb;
          if (cb + 4 <= p.cin) {
            cp_async16(dst, src);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (cb + e < p.cin) cp_async4(dst + e, src + e);
              else dst[e] = 0.f;
            }
          }
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
      for (int idx = tid; idx < TM * cw; idx += NT) {
        const int site = idx / cw, e = idx % cw;
        const long long j = tile_base + site;
        int r = -1;
        if (j < p.n_out) r = conv_src_row(p, j, k);
        float* dst = X + site * XS + e;
        if (r >= 0 && c0 + e < p.cin) {
          live = 1;
          cp_async4(dst, p.in + (long long)r * p.ld_in + c0 + e);
        } else {
          *dst = 0.f;
        }
      }
    }
    return live;
  };

  int live_cur = issue(0);
  cp_async_commit();
  for (int s = 0; s < S; ++s) {
    int live_next = 0;
    if (s + 1 < S) {
      live_next = issue(s + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const int any = __syncthreads_or(live_cur);
    if (any) {
      const int k = s / nchunks, c0 = (s % nchunks) * CHUNK;
      const int cw = min(CHUNK, cinp - c0);
      const float* X = X_s + (size_t)(s & 1) * TM * XS;
      const float* Wk = W_s + ((size_t)k * cinp + c0) * COUT + cg * 4;
      for (int c4 = 0; c4 < cw; c4 += 4) {
        float4 xv[TS];
#pragma unroll
        for (int t = 0; t < TS; ++t)
          xv[t] = *reinterpret_cast<const float4*>(X + (sg + NSG * t) * XS + c4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 w = *reinterpret_cast<const float4*>(Wk + (size_t)(c4 + e) * COUT);
#pragma unroll
          for (int t = 0; t < TS; ++t) {
            const float x = e == 0 ? xv[t].x : e == 1 ? xv[t].y : e == 2 ? xv[t].z : xv[t].w;
            acc[t][0] = fmaf(x, w.x, acc[t][0]);
            acc[t][1] = fmaf(x, w.y, acc[t][1]);
            acc[t][2] = fmaf(x, w.z, acc[t][2]);
            acc[t][3] = fmaf(x, w.w, acc[t][3]);
          }
        }
      }
    }
    __syncthreads();
    live_cur = live_next;
  }

  conv_epilogue<COUT>(p, acc, tile_base, sg, cg);
}

// ------------------------------------------------------------------------------------------------
// Child-mode kernel (generative upsampling, model.py:192-207,224-225): the n1 convolution over
// the 8 children of every site.  A tile is 16 parents = 128 outputs.  The 27 parent-neighbour rows of each
// parent are staged in shared memory ONCE (16 x 27 rows) and serve all 8 children x 27 offsets -- 8x less
// gather traffic than treating the 128 children as independent rows.  Only the 3 KB filter slice of the
// current offset streams (double buffered).  Thread (g, w, cg): parents {g, g+8}, children {w, w+4},
// channels [4cg, 4cg+4): the 8 lane groups of a warp read 8 different parents of the same (child, offset),
// i.e. 8 rows at stride 27*XS -- conflict free.  Arithmetic order identical to the generic kernels.
template <int CINP>
struct TileCfg {
  static constexpr int XS2 = (CINP % 8 == 4) ? CINP : CINP + 4;  // row stride == 4 (mod 8) words: conflict-free LDS.128
};

template <int COUT, int CINP>
__global__ void __launch_bounds__(NSG * (COUT / 4), 2)
conv_child_f32_kernel(ConvParams p) {
  constexpr int NCG = COUT / 4;
  constexpr int NT = NSG * NCG;
  constexpr int PT = TM / 8;  // parents per tile
  constexpr int XS2 = TileCfg<CINP>::XS2;
  constexpr int CPR = CINP / 4;
  constexpr int WT = CINP * COUT;
  extern __shared__ __align__(16) float smem[];
  float* X_s = smem;                             // [PT][27][XS2]
  float* W_s = X_s + PT * 27 * XS2;              // [2][CINP][COUT]
  int* idx_s = reinterpret_cast<int*>(W_s + 2 * WT);  // [27][PT]
  __shared__ unsigned live_kp_s;
  __shared__ signed char kp_tab[8][27];

  const int tid = threadIdx.x;
  const int sg = tid / NCG, cg = tid % NCG;
  const int g = sg & 7, w = sg >> 3;
  const long long parent_base = (long long)blockIdx.x * PT;
  const long long n_parent = p.n_out >> 3;

  if (tid == 0) live_kp_s = 0u;
  for (int e = tid; e < 8 * 27; e += NT) {
    const int c = e / 27, k = e % 27;
    const int dz = k / 9 - 1, dy = (k / 3) % 3 - 1, dx = k % 3 - 1;
    const int pz = (((c >> 2) & 1) + dz + 2) / 2 - 1;
    const int py = (((c >> 1) & 1) + dy + 2) / 2 - 1;
    const int px = ((c & 1) + dx + 2) / 2 - 1;
    kp_tab[c][k] = (signed char)((pz + 1) * 9 + (py + 1) * 3 + (px + 1));
  }
  for (int b = 0; b < 2; ++b)
    for (int q = p.cin * COUT + tid; q < WT; q += NT) W_s[b * WT + q] = 0.f;
  __syncthreads();
  unsigned my_live = 0u;
  for (int e = tid; e < 27 * PT; e += NT) {
    const int kp = e / PT, pl = e % PT;
    const long long pr = parent_base + pl;
    const int r = pr < n_parent ? __ldg(p.nbr + (long long)kp * p.nbr_stride + pr) : -1;
    idx_s[e] = r;
    if (r >= 0) my_live |= 1u << kp;
  }
  my_live = __reduce_or_sync(0xffffffffu, my_live);
  if ((tid & 31) == 0 && my_live) atomicOr(&live_kp_s, my_live);
  __syncthreads();
  // stage all parent-neighbour rows of the tile
  for (int pi = tid; pi < 27 * PT * CPR; pi += NT) {
    const int ch = pi % CPR, e = pi / CPR;
    const int kp = e / PT, pl = e % PT;
    const int r = idx_s[e];
    float* dst = X_s + (pl * 27 + kp) * XS2 + ch * 4;
    if (r >= 0) {
      const float* src = p.in + (long long)r * p.ld_in + ch * 4;
      if (ch * 4 + 4 <= p.cin) {
        cp_async16(dst, src);
      } else {
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          if (ch * 4 + e2 < p.cin) cp_async4(dst + e2, src + e2);
          else dst[e2] = 0.f;
        }
      }
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // offsets with at least one live (child, parent neighbour) in the tile
  const unsigned live_kp = live_kp_s;
  unsigned live_k = 0u;
  for (int k = 0; k < 27; ++k) {
    bool any = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) any = any || ((live_kp >> kp_tab[c][k]) & 1u);
    if (any) live_k |= 1u << k;
  }
  auto issue_w = [&](int k, int buf) {
    const float* Wk = p.weight + (size_t)k * p.cin * COUT;
    float* Wd = W_s + buf * WT;
    for (int q = tid; q < p.cin * NCG; q += NT) cp_async16(Wd + q * 4, Wk + q * 4);
  };
  unsigned issue_mask = live_k;
  const int S = __popc(live_k);
  if (issue_mask) {
    issue_w(__ffs(issue_mask) - 1, 0);
    issue_mask &= issue_mask - 1;
  }
  cp_async_commit();  // group 0: all X rows + first filter slice

  float acc[TS][4];
#pragma unroll
  for (int t = 0; t < TS; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;

  unsigned comp_mask = live_k;
  for (int s = 0; s < S; ++s) {
    cp_async_wait<0>();
    __syncthreads();
    if (issue_mask) {
      issue_w(__ffs(issue_mask) - 1, (s + 1) & 1);
      issue_mask &= issue_mask - 1;
    }
    cp_async_commit();
    const int k = __ffs(comp_mask) - 1;
    comp_mask &= comp_mask - 1;
    const float* Wd = W_s + (s & 1) * WT + cg * 4;
    const float* xr[TS];
#pragma unroll
    for (int t = 0; t < TS; ++t) {
      const int pl = g + 8 * (t & 1), c = w + 4 * (t >> 1);
      xr[t] = X_s + (pl * 27 + kp_tab[c][k]) * XS2;
    }
#pragma unroll
    for (int c4 = 0; c4 < CINP; c4 += 4) {
      float4 xv[TS];
#pragma unroll
      for (int t = 0; t < TS; ++t) xv[t] = *reinterpret_cast<const float4*>(xr[t] + c4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 wv = *reinterpret_cast<const float4*>(Wd + (c4 + e) * COUT);
#pragma unroll
        for (int t = 0; t < TS; ++t) {
          const float x = e == 0 ? xv[t].x : e == 1 ? xv[t].y : e == 2 ? xv[t].z : xv[t].w;
          acc[t][0] = fmaf(x, wv.x, acc[t][0]);
          acc[t][1] = fmaf(x, wv.y, acc[t][1]);
          acc[t][2] = fmaf(x, wv.z, acc[t][2]);
          acc[t][3] = fmaf(x, wv.w, acc[t][3]);
        }
      }
    }
  }
  cp_async_wait<0>();
  long long rows[TS];
#pragma unroll
  for (int t = 0; t < TS; ++t) rows[t] = (parent_base + g + 8 * (t & 1)) * 8 + w + 4 * (t >> 1);
  conv_epilogue_rows<COUT>(p, acc, rows, cg);
}

template <int COUT, int CINP>
static int launch_child(const ConvParams& p, cudaStream_t st) {
  constexpr int NT = NSG * (COUT / 4);
  constexpr int PT = TM / 8;
  const size_t smem = ((size_t)PT * 27 * TileCfg<CINP>::XS2 + 2 * CINP * COUT) * sizeof(float) + 27 * PT * sizeof(int);
  const long long tiles = ((p.n_out >> 3) + PT - 1) / PT;
  if (tiles > 0x7fffffff) return SGNN_E_TOO_LARGE;
  static bool attr_set = false;
  if (!attr_set) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_child_f32_kernel<COUT, CINP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_set = true;
  }
  conv_child_f32_kernel<COUT, CINP><<<(int)tiles, NT, smem, st>>>(p);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// ------------------------------------------------------------------------------------------------
// v4 "row-owner" kernel (default).  History, with the ncu evidence under profiles/:
//   v2 (4x4 register tile, weights re-read by every lane group through LDS.128): shared-memory pipe saturated,
//      FFMA pipe 39 %;  v3 (weights in __constant__, LDCU -> FFMA R,R,UR,R): no smem traffic for weights but every
//      stage touches a fresh 1-3 KB constant slice -> constant-cache MISS latency chain, ~3.3 us per stage, a
//      90 us floor per launch.
// v4: each thread owns S whole output rows (all COUT accumulators in registers).  A stage = (filter offset k,
// 16-channel slice of Cin).  Per stage a WARP gathers its 32*S neighbour-row slices cooperatively (CPR lanes per
// row, coalesced) into the owners' private slots of a shared-memory ring and its own copy of the W[k] slice
// (<= 1 KB); the owner then reads its rows with conflict-free LDS.128 (XOR chunk swizzle) and the weights with
// warp-UNIFORM LDS.128 (one broadcast wavefront).  Only __syncwarp is needed -- no block barrier, warps run
// decoupled.  Neighbour indices are prefetched one stage ahead (coalesced).  Arithmetic unchanged: k ascending,
// ci ascending, one fmaf chain per output element from +0.
template <int CH>
struct RoCfg {
  static constexpr int CPR = CH / 4;  // 16-byte chunks per row slice
  static constexpr int TZ = (CPR % 8 == 0) ? 3 : (CPR % 4 == 0) ? 2 : (CPR % 2 == 0) ? 1 : 0;
};
// chunk swizzle of the slice owned by lane l (slices are CPR chunks apart; for even CPR XOR the low TZ bits)
template <int CH>
__device__ __forceinline__ int ro_swz(int l) {
  return (l >> (3 - RoCfg<CH>::TZ)) & ((1 << RoCfg<CH>::TZ) - 1);
}

template <int COUT, int S, int W>   // one compute step over a W-channel slice (W multiple of 4, compile time)
__device__ __forceinline__ void ro_compute(float (&acc)[S][COUT], const float* __restrict__ X, int xs_stride,
                                           int my_sw, const float* __restrict__ Wst) {
#pragma unroll
  for (int ch = 0; ch < W / 4; ++ch) {
    float4 xv[S];
#pragma unroll
    for (int s = 0; s < S; ++s) xv[s] = *reinterpret_cast<const float4*>(X + (size_t)s * xs_stride + ((ch * 4) ^ my_sw));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
      for (int co = 0; co < COUT; co += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(Wst + (ch * 4 + e) * COUT + co);  // warp-uniform
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float x = e == 0 ? xv[s].x : e == 1 ? xv[s].y : e == 2 ? xv[s].z : xv[s].w;
          acc[s][co] = fmaf(x, wv.x, acc[s][co]);
          acc[s][co + 1] = fmaf(x, wv.y, acc[s][co + 1]);
          acc[s][co + 2] = fmaf(x, wv.z, acc[s][co + 2]);
          acc[s][co + 3] = fmaf(x, wv.w, acc[s][co + 3]);
        }
      }
    }
  }
}

// epilogue of one output row held entirely by one thread: residual add, two slots with optional affine + relu
template <int COUT>
__device__ __forceinline__ void ro_epilogue_row(const ConvParams& p, const float (&acc)[COUT], long long j) {
  if (j >= p.n_out) return;
#pragma unroll
  for (int c = 0; c < COUT; c += 4) {
    float4 v = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    if (p.residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + j * p.ld_res + c));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (p.out_a) {
      float4 y = v;
      if (p.scale_a) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale_a + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift_a + c));
        y.x = fmaf(v.x, sc.x, sh.x); y.y = fmaf(v.y, sc.y, sh.y); y.z = fmaf(v.z, sc.z, sh.z); y.w = fmaf(v.w, sc.w, sh.w);
      }
      if (p.relu_a) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
      *reinterpret_cast<float4*>(p.out_a + j * p.ld_a + c) = y;
    }
    if (p.out_b) {
      float4 y = v;
      if (p.scale_b) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale_b + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift_b + c));
        y.x = fmaf(v.x, sc.x, sh.x); y.y = fmaf(v.y, sc.y, sh.y); y.z = fmaf(v.z, sc.z, sh.z); y.w = fmaf(v.w, sc.w, sh.w);
      }
      if (p.relu_b) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
      *reinterpret_cast<float4*>(p.out_b + j * p.ld_b + c) = y;
    }
  }
}

// CIN: exact input channels; CH: channels per stage (CINP when CIN <= 16, else 16); S rows per thread; NS ring depth
template <int COUT, int CIN, int CH, int S, int NS, bool VEC, bool CHILD>
__global__ void __launch_bounds__(128)
conv_ro_kernel(ConvParams p) {
  constexpr bool vec = VEC;
  constexpr int CINP = (CIN + 3) & ~3;
  constexpr int NSUB = (CINP + CH - 1) / CH;          // stages per filter offset
  constexpr int TAIL = CINP - (NSUB - 1) * CH;         // channels (padded to 4) of the last slice
  constexpr int CPR = CH / 4;
  constexpr int XSL = S * 128 * CH;                    // floats: X slots of one stage (whole CTA)
  constexpr int WSL = 4 * CH * COUT;                   // floats: per-warp W slices of one stage
  extern __shared__ __align__(16) float smem[];        // [NS][ X: [S][128][CH] | W: [4 warps][CH][COUT] ]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wbase = tid & ~31;
  const long long row0 = (long long)blockIdx.x * (128 * S) + tid;
  const int my_sw = ro_swz<CH>(lane) << 2;

  float acc[S][COUT];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[s][c] = 0.f;

  int ir[S];
  auto load_idx = [&](int k) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const long long j = row0 + 128 * s;
      if (CHILD) ir[s] = j < p.n_out ? conv_src_row(p, j, k) : -1;
      else ir[s] = j < p.n_out ? __ldg(p.nbr + (long long)k * p.nbr_stride + j) : -1;
    }
  };
  // gather of stage (k, sub) into ring slot `buf`; ir[] holds the row indices of offset k
  auto issue = [&](int k, int sub, int buf) {
    float* Xb = smem + (size_t)buf * (XSL + WSL);
    const int c0 = sub * CH;
    if (vec) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float* slot = Xb + ((size_t)s * 128 + wbase) * CH;
#pragma unroll
        for (int it = 0; it < CPR; ++it) {
          const int q = lane + 32 * it;
          const int rl = q / CPR, ch = q % CPR;
          const int r = __shfl_sync(0xffffffffu, ir[s], rl);
          float* dst = slot + rl * CH + ((ch ^ ro_swz<CH>(rl)) * 4);
          const int cb = c0 + ch * 4;
          if (r >= 0 && cb < CIN) {
            const float* src = p.in + (long long)r * p.ld_in + cb;
            if (cb + 4 <= CIN) {
              cp_async16(dst, src);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (cb + e < CIN) cp_async4(dst + e, src + e);
                else dst[e] = 0.f;
              }
            }
          } else {
            *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float* dst = Xb + ((size_t)s * 128 + tid) * CH;
        const int r = ir[s];
        const int sw = ro_swz<CH>(lane);
#pragma unroll
        for (int e = 0; e < CH; ++e) {
          float* d = dst + (((e >> 2) ^ sw) << 2) + (e & 3);
          if (r >= 0 && c0 + e < CIN) cp_async4(d, p.in + (long long)r * p.ld_in + c0 + e);
          else *d = 0.f;
        }
      }
    }
    // this warp's copy of the filter slice W[k][c0 .. c0+CH) (contiguous in global memory)
    float* Wd = Xb + XSL + warp * (CH * COUT);
    const int wn = min(CH, CIN - c0) * COUT / 4;  // 16-byte chunks (COUT % 4 == 0)
    const float* Wk = p.weight + ((size_t)k * CIN + c0) * COUT;
    for (int q = lane; q < wn; q += 32) cp_async16(Wd + q * 4, Wk + q * 4);
    if (CINP != CIN && sub == NSUB - 1)   // weight rows of the channel padding (ring slots are reused: zero each time)
      for (int q = (CIN - (NSUB - 1) * CH) * COUT + lane; q < TAIL * COUT; q += 32) Wd[q] = 0.f;
  };

  const int total = p.K * NSUB;
  int ik = 0, isub = 0, issued = 0;   // next stage to issue
  load_idx(0);
#pragma unroll
  for (int i = 0; i < NS - 1; ++i) {
    if (issued < total) {
      issue(ik, isub, i);
      ++issued;
      if (++isub == NSUB) { isub = 0; ++ik; if (ik < p.K) load_idx(ik); }
    }
    cp_async_commit();
  }
  int st = 0, fill = NS - 1, csub = 0;
  for (int t = 0; t < total; ++t) {
    cp_async_wait<NS - 2>();
    __syncwarp();  // stage t landed for every lane of this warp; all lanes are done with the slot refilled below
    if (issued < total) {
      issue(ik, isub, fill);
      ++issued;
      if (++isub == NSUB) { isub = 0; ++ik; if (ik < p.K) load_idx(ik); }
    }
    cp_async_commit();
    const float* Xb = smem + (size_t)st * (XSL + WSL);
    const float* X = Xb + (size_t)tid * CH;
    const float* Wst = Xb + XSL + warp * (CH * COUT);
    if (csub < NSUB - 1) ro_compute<COUT, S, CH>(acc, X, 128 * CH, my_sw, Wst);
    else ro_compute<COUT, S, TAIL>(acc, X, 128 * CH, my_sw, Wst);
    if (++csub == NSUB) csub = 0;
    st = st + 1 == NS ? 0 : st + 1;
    fill = fill + 1 == NS ? 0 : fill + 1;
  }
  cp_async_wait<0>();

  // ---- epilogue (per row: all COUT channels are in this thread)
#pragma unroll
  for (int s = 0; s < S; ++s) ro_epilogue_row<COUT>(p, acc[s], row0 + 128 * s);
}

template <int COUT, int CIN, int S, int CH>
static int launch_ro(const ConvParams& p, bool vec, cudaStream_t st) {
  constexpr int STAGE = (S * 128 * CH + 4 * CH * COUT) * 4;
  constexpr int NS = (3 * STAGE <= 60 * 1024) ? 3 : 2;
  const size_t smem = (size_t)NS * STAGE;
  const long long tiles = (p.n_out + 128 * S - 1) / (128 * S);
  if (tiles > 0x7fffffff) return SGNN_E_TOO_LARGE;
  static bool attr_set = false;
  if (!attr_set) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_ro_kernel<COUT, CIN, CH, S, NS, true, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGNN_CUDA(cudaFuncSetAttribute(conv_ro_kernel<COUT, CIN, CH, S, NS, false, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGNN_CUDA(cudaFuncSetAttribute(conv_ro_kernel<COUT, CIN, CH, S, NS, true, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (p.child_mode) {
    if (!vec) return SGNN_E_ALIGN;
    conv_ro_kernel<COUT, CIN, CH, S, NS, true, true><<<(int)tiles, 128, smem, st>>>(p);
  } else if (vec) {
    conv_ro_kernel<COUT, CIN, CH, S, NS, true, false><<<(int)tiles, 128, smem, st>>>(p);
  } else {
    conv_ro_kernel<COUT, CIN, CH, S, NS, false, false><<<(int)tiles, 128, smem, st>>>(p);
  }
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// exact (cin, cout) pairs of the SG-NN channel plan
static int dispatch_ro(const ConvParams& p, bool vec, cudaStream_t st, bool* handled) {
  *handled = true;
  // Rows per thread by launch size: S=4 amortises the weight reads best but makes 512-row CTAs; a launch that cannot
  // fill the chip (148 SMs x 3 CTAs) with those is latency bound (ncu: 45-80 us floors on the coarse levels), so
  // mid-size launches use S=2 and small ones S=1 (4x more, 4x shorter CTAs).
  if (p.child_mode && !vec) { *handled = false; return SGNN_OK; }
  const long long t4 = 300000, t2 = 90000;
  const int sz = p.n_out >= t4 ? 4 : (p.n_out >= t2 ? 2 : 1);
#define SGNN_RO_CASE(CO, CI, SS, CHH) \
  if (p.cout == CO && p.cin == CI) return launch_ro<CO, CI, SS, CHH>(p, vec, st);
#define SGNN_RO_SIZED(CO, CI, CHH)              \
  if (p.cout == CO && p.cin == CI) {            \
    if (sz == 4) return launch_ro<CO, CI, 4, CHH>(p, vec, st); \
    if (sz == 2) return launch_ro<CO, CI, 2, CHH>(p, vec, st); \
    return launch_ro<CO, CI, 1, CHH>(p, vec, st);              \
  }
  SGNN_RO_SIZED(8, 1, 4)
  SGNN_RO_SIZED(8, 8, 8)
  SGNN_RO_SIZED(12, 8, 8)
  SGNN_RO_SIZED(12, 12, 12)
  SGNN_RO_SIZED(16, 12, 12)
  SGNN_RO_SIZED(16, 16, 16)
  {   // wide inputs: one stage per offset (whole padded row)
    if (sz >= 2) {
      SGNN_RO_CASE(16, 26, 2, 28)
      SGNN_RO_CASE(16, 30, 2, 32)
      SGNN_RO_CASE(16, 34, 2, 36)
      SGNN_RO_CASE(16, 48, 2, 24)
    } else {
      SGNN_RO_CASE(16, 26, 1, 28)
      SGNN_RO_CASE(16, 30, 1, 32)
      SGNN_RO_CASE(16, 34, 1, 36)
      SGNN_RO_CASE(16, 48, 1, 24)
    }
  }
#undef SGNN_RO_SIZED
#undef SGNN_RO_CASE
  *handled = false;
  return SGNN_OK;
}

// Generic fallback: any Cout, one thread per (row, co).  Same summation order.
__global__ void conv_gather_f32_generic_kernel(ConvParams p) {
  const long long total = p.n_out * p.cout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long j = idx / p.cout;
    const int co = (int)(idx % p.cout);
    float acc = 0.f;
    for (int k = 0; k < p.K; ++k) {
      const int r = conv_src_row(p, j, k);
      if (r < 0) continue;
      const float* x = p.in + (long long)r * p.ld_in;
      const float* w = p.weight + (size_t)k * p.cin * p.cout + co;
      for (int ci = 0; ci < p.cin; ++ci) acc = fmaf(__ldg(x + ci), __ldg(w + (size_t)ci * p.cout), acc);
    }
    float v = acc;
    if (p.residual) v += p.residual[j * p.ld_res + co];
    if (p.out_a) {
      float y = v;
      if (p.scale_a) y = fmaf(v, p.scale_a[co], p.shift_a[co]);
      if (p.relu_a) y = fmaxf(y, 0.f);
      p.out_a[j * p.ld_a + co] = y;
    }
    if (p.out_b) {
      float y = v;
      if (p.scale_b) y = fmaf(v, p.scale_b[co], p.shift_b[co]);
      if (p.relu_b) y = fmaxf(y, 0.f);
      p.out_b[j * p.ld_b + co] = y;
    }
  }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <int COUT>
static int launch_conv(const ConvParams& p, bool vec, cudaStream_t st) {
  constexpr int NT = NSG * (COUT / 4);
  const int cinp = (p.cin + 3) & ~3;
  const size_t smem = ((size_t)p.K * cinp * COUT + 2 * (size_t)TM * XS) * sizeof(float);
  if (smem > 227 * 1024) return SGNN_E_UNSUPPORTED;
  const long long tiles = (p.n_out + TM - 1) / TM;
  if (tiles > 0x7fffffff) return SGNN_E_TOO_LARGE;
  if (vec) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_gather_f32_kernel<COUT, 4>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_gather_f32_kernel<COUT, 4><<<(int)tiles, NT, smem, st>>>(p);
  } else {
    SGNN_CUDA(cudaFuncSetAttribute(conv_gather_f32_kernel<COUT, 1>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_gather_f32_kernel<COUT, 1><<<(int)tiles, NT, smem, st>>>(p);
  }
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

static int check_epilogue(const SgnnEpilogue& e, bool need_vec) {
  if (!e.out) return SGNN_OK;
  if ((e.scale == nullptr) != (e.shift == nullptr)) return SGNN_E_INVALID;
  if (need_vec && (!aligned16(e.out) || (e.ld & 3) || (e.scale && (!aligned16(e.scale) || !aligned16(e.shift)))))
    return SGNN_E_ALIGN;
  return SGNN_OK;
}

int sgnn_conv_tc_bf16(const void* in, int ld_in, const int* tbl, long long tbl_stride, int K, int mode, const void* w,
                      int cin, int cout, long long n_out, const SgnnEpilogue* ep, cudaStream_t st);  // conv_tc.cu

int sgnn_conv_forward_rowlane(const SgnnConvArgs* a, cudaStream_t st);   // conv_sp.cu

// K = 27 launches up to this many rows are latency bound on the staged row-owner pipeline (27 stages of cp.async -> wait ->
// compute: a flat 25-27 us even at 256 rows); the lane = (row, channel group) kernel of conv_sp.cu takes them (G taps of
// loads in flight per thread).  Same fmaf chain, same bits.  Measured (scratch/sp_bench.py, B200): 2048 rows 16->16: 18.4 vs
// 23.5 us through ctypes; at 16 k rows and beyond, and for K = 8 at every size, the row-owner kernel wins (its weights are
// warp-uniform broadcasts, the row-lane kernel re-reads them per row: 25.5 vs 23.6 us at 16 k rows, 142 vs 80 us at 126 k).
#define SGNN_ROWLANE_MAX_ROWS 4096

extern "C" int sgnn_conv_forward(const SgnnConvArgs* a, void* stream) {
  if (!a || a->n_out < 0 || a->cin <= 0 || a->cout <= 0 || !a->weight) return SGNN_E_INVALID;
  if (a->dtype == SGNN_BF16) {   // tcgen05 path (conv_tc.cu): bf16 features/weights/outputs, fp32 accumulation in TMEM
    if (a->child_mode || a->residual || a->b.out) return SGNN_E_UNSUPPORTED;
    return sgnn_conv_tc_bf16(a->in, a->ld_in, a->nbr, a->nbr_stride, a->K, 0, a->weight, a->cin, a->cout, a->n_out,
                             &a->a, (cudaStream_t)stream);
  }
  if (a->dtype != SGNN_F32) return SGNN_E_UNSUPPORTED;
  if (a->K != 27 && a->K != 8) return SGNN_E_UNSUPPORTED;
  if (a->child_mode && a->K != 27) return SGNN_E_INVALID;
  if (a->n_out == 0) return SGNN_OK;
  if (!a->a.out && !a->b.out) return SGNN_E_INVALID;
  if (!a->in || !a->nbr) return SGNN_E_INVALID;
  if (a->cin > 64) return SGNN_E_UNSUPPORTED;
  ConvParams p;
  p.in = (const float*)a->in; p.ld_in = a->ld_in;
  p.nbr = a->nbr; p.nbr_stride = a->nbr_stride; p.K = a->K; p.child_mode = a->child_mode;
  p.weight = (const float*)a->weight; p.cin = a->cin; p.cout = a->cout; p.n_out = a->n_out;
  p.residual = (const float*)a->residual; p.ld_res = a->ld_res;
  p.out_a = (float*)a->a.out; p.ld_a = a->a.ld; p.relu_a = a->a.relu; p.scale_a = a->a.scale; p.shift_a = a->a.shift;
  p.out_b = (float*)a->b.out; p.ld_b = a->b.ld; p.relu_b = a->b.relu; p.scale_b = a->b.scale; p.shift_b = a->b.shift;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tiled = (a->cout == 4 || a->cout == 8 || a->cout == 12 || a->cout == 16);
  int rc;
  if ((rc = check_epilogue(a->a, tiled))) return rc;
  if ((rc = check_epilogue(a->b, tiled))) return rc;
  if (tiled) {
    if (!aligned16(a->weight)) return SGNN_E_ALIGN;
    if (a->residual && (!aligned16(a->residual) || (a->ld_res & 3))) return SGNN_E_ALIGN;
    const bool vec = aligned16(a->in) && (a->ld_in & 3) == 0;
    if (!p.child_mode && !(a->flags & SGNN_CONV_NO_ROWLANE) &&
        a->K == 27 && (a->n_out <= SGNN_ROWLANE_MAX_ROWS || (a->flags & SGNN_CONV_ROWLANE))) {
      rc = sgnn_conv_forward_rowlane(a, st);
      if (rc != SGNN_E_UNSUPPORTED) return rc;
    }
    {
      bool handled = false;
      if (vec && p.child_mode && p.cout == 16 && p.cin == 48 && (p.n_out & 7) == 0) return launch_child<16, 48>(p, st);
      rc = dispatch_ro(p, vec, st, &handled);
      if (handled) return rc;
    }
    switch (a->cout) {
      case 4: return launch_conv<4>(p, vec, st);
      case 8: return launch_conv<8>(p, vec, st);
      case 12: return launch_conv<12>(p, vec, st);
      default: return launch_conv<16>(p, vec, st);
    }
  }
  conv_gather_f32_generic_kernel<<<sgnn_blocks(a->n_out * a->cout, 256), 256, 0, st>>>(p);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// ------------------------------------------------------------ deconvolution (filter 2, stride 2)
// every fine row has exactly one coarse parent: out[i] = in[parent>>3] @ W[parent&7]
__global__ void deconv_f32_kernel(const float* __restrict__ in, int ld_in, const int* __restrict__ parent,
                                  const float* __restrict__ weight, int cin, int cout, long long n,
                                  float* out, int ld, int relu, const float* scale, const float* shift) {
  extern __shared__ float W_s[];  // [8][cin][cout]
  for (int i = threadIdx.x; i < 8 * cin * cout; i += blockDim.x) W_s[i] = weight[i];
  __syncthreads();
  const long long total = n * cout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / cout;
    const int co = (int)(idx % cout);
    const int pk = __ldg(parent + i);
    float acc = 0.f;
    if (pk >= 0) {
      const float* x = in + (long long)(pk >> 3) * ld_in;
      const float* w = W_s + (size_t)(pk & 7) * cin * cout + co;
      for (int ci = 0; ci < cin; ++ci) acc = fmaf(__ldg(x + ci), w[(size_t)ci * cout], acc);
    }
    float y = acc;
    if (scale) y = fmaf(acc, scale[co], shift[co]);
    if (relu) y = fmaxf(y, 0.f);
    out[i * ld + co] = y;
  }
}

extern "C" int sgnn_deconv_forward(const void* in, int32_t ld_in, int32_t dtype, const int32_t* parent,
                                   const void* weight, int32_t cin, int32_t cout, int64_t n_fine,
                                   const SgnnEpilogue* ep, void* stream) {
  if (n_fine < 0 || cin <= 0 || cout <= 0 || !ep || !ep->out || !weight) return SGNN_E_INVALID;
  if (dtype == SGNN_BF16)
    return sgnn_conv_tc_bf16(in, ld_in, parent, 0, 8, 1, weight, cin, cout, n_fine, ep, (cudaStream_t)stream);
  if (dtype != SGNN_F32) return SGNN_E_UNSUPPORTED;
  if ((ep->scale == nullptr) != (ep->shift == nullptr)) return SGNN_E_INVALID;
  if (n_fine == 0) return SGNN_OK;
  if (!in || !parent) return SGNN_E_INVALID;
  const size_t smem = (size_t)8 * cin * cout * 4;
  if (smem > 200 * 1024) return SGNN_E_UNSUPPORTED;
  SGNN_CUDA(cudaFuncSetAttribute(deconv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  deconv_f32_kernel<<<sgnn_blocks(n_fine * cout, 256, 148 * 8), 256, smem, (cudaStream_t)stream>>>(
      (const float*)in, ld_in, parent, (const float*)weight, cin, cout, (long long)n_fine,
      (float*)ep->out, ep->ld, ep->relu, ep->scale, ep->shift);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// ------------------------------------------------------------ unpooling
__global__ void unpool_kernel(const float* __restrict__ in, int ld_in, const int* __restrict__ parent, int c,
                              long long n, float* out, int ld, int relu, const float* scale, const float* shift) {
  const long long total = n * c;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / c;
    const int ch = (int)(idx % c);
    const int pk = __ldg(parent + i);
    float v = pk >= 0 ? __ldg(in + (long long)(pk >> 3) * ld_in + ch) : 0.f;
    if (scale) v = fmaf(v, scale[ch], shift[ch]);
    if (relu) v = fmaxf(v, 0.f);
    out[i * ld + ch] = v;
  }
}

// 16-byte form (c, both leading dimensions and all pointers multiples of 4 floats / 16 bytes): one thread per (row, 4 channels),
// no 64-bit division per element -- the outer unpool of the finest level moves 690 k rows x 32 channels.
template <int Q>   // Q = c / 4 (compile time: 4 -> 16 channels, 8 -> 32 channels)
__global__ void unpool_vec4_kernel(const float* __restrict__ in, int ld_in, const int* __restrict__ parent, long long n,
                                   float* out, int ld, int relu, const float* scale, const float* shift) {
  const int q = threadIdx.x % Q;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) { sc = __ldg(reinterpret_cast<const float4*>(scale) + q); sh = __ldg(reinterpret_cast<const float4*>(shift) + q); }
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / Q; i < n; i += ((long long)gridDim.x * blockDim.x) / Q) {
    const int pk = __ldg(parent + i);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pk >= 0) v = __ldg(reinterpret_cast<const float4*>(in + (long long)(pk >> 3) * ld_in) + q);
    if (scale) { v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w); }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(out + i * ld)[q] = v;
  }
}

extern "C" int sgnn_unpool(const float* in, int32_t ld_in, const int32_t* parent, int32_t c, int64_t n_fine,
                           const SgnnEpilogue* ep, void* stream) {
  if (n_fine < 0 || c <= 0 || !ep || !ep->out) return SGNN_E_INVALID;
  if ((ep->scale == nullptr) != (ep->shift == nullptr)) return SGNN_E_INVALID;
  if (n_fine == 0) return SGNN_OK;
  if (!in || !parent) return SGNN_E_INVALID;
  const bool v4 = (c == 16 || c == 32) && (ld_in & 3) == 0 && (ep->ld & 3) == 0 && aligned16(in) && aligned16(ep->out) &&
                  (!ep->scale || (aligned16(ep->scale) && aligned16(ep->shift)));
  if (v4) {
    const int blocks = sgnn_blocks(n_fine * (c / 4), 256);
    if (c == 16) unpool_vec4_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, ld_in, parent, (long long)n_fine, (float*)ep->out,
                                                                             ep->ld, ep->relu, ep->scale, ep->shift);
    else unpool_vec4_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, ld_in, parent, (long long)n_fine, (float*)ep->out,
                                                                        ep->ld, ep->relu, ep->scale, ep->shift);
    SGNN_CHECK_LAUNCH();
    return SGNN_OK;
  }
  unpool_kernel<<<sgnn_blocks(n_fine * c, 256), 256, 0, (cudaStream_t)stream>>>(
      in, ld_in, parent, c, (long long)n_fine, (float*)ep->out, ep->ld, ep->relu, ep->scale, ep->shift);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}
