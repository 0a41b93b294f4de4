// conv_ur.cu -- submanifold 3^3 convolution on tcgen05 with every DISTINCT input row of a tile staged once by TMA.
// SURVEY §8 rows a2 / a3 (reference torch/model.py:32,38,40,179,186,254 and the FullyConvolutionalNet blocks);
// the north_star's "TMA staging of the filter bank and per-offset input slices into shared memory".
//
// Why.  ncu on the round-1 kernels (profiles/r01_conv_tc32_*): every tensor-core generation moved 27 x 64..96 B per
// output row from L2 in ~300 us -- the L2->SM gather was the common wall, and with the generator's row order a 128-row
// tile touches only ~1.6 distinct input rows per output row (scratch/tile_stats.py: median 203, max 302 per tile).
//
// Two kernels.
//  (1) tile_plan_kernel, once per site set (rulebook time): for every 128-row tile the sorted list of distinct input
//      rows its 27 filter offsets touch (`urows`, by a shared-memory BITMAP over the tile's row-id span + popcount
//      ranks -- no hash, no sort) and the 27 x 128 table of 16-bit LOCAL indices into that list (`lidx`, 0xFFFF =
//      absent).  54 B of index per row instead of the 108 B of the k-major neighbour table, and consecutive output
//      rows get consecutive local indices (conflict-free shared-memory reads).
//  (2) conv_ur_kernel, one persistent CTA per SM, 14 warps in four roles:
//        loader   (warp 13)   TMA: the prepared filter bank (one cp.async.bulk), the tile's lidx block (one bulk copy)
//                             and its distinct input rows (one 64/128-byte cp.async.bulk per row) into a ring of
//                             64-row chunks -- mbarrier complete_tx, no registers, runs up to a whole tile ahead;
//        producers (warps 0-7) cut each landed row ONCE into the three bf16 planes (shared memory, chunk-major so that
//                             consecutive rows are consecutive 16-byte units), then, tap by tap, thread r moves row
//                             lidx[k][r]'s 96 x Q bytes shared memory -> TENSOR MEMORY (6Q LDS.128 + 3Q tcgen05.st);
//                             absent taps read a zero row (no predicates); the two warp groups take alternate taps;
//        MMA      (warp 12)   6Q tcgen05.mma (A in TMEM, B = resident filter bank) per tap into 7 main + 1 correction
//                             accumulators (tc32_common.cuh), double-buffered accumulators (2 x 128 TMEM columns);
//        epilogue (warps 8-11) tcgen05.ld, round-to-nearest sum of the partial accumulators, residual, two affine+ReLU
//                             slots -- overlaps the next tile's taps.
//      A tile with more distinct rows than the staging area takes several passes over the taps (same accumulators);
//      a tile the plan could not describe (row-id span or distinct count over its caps) runs in DIRECT mode (rows
//      gathered from global memory per tap, split in registers) -- correct for any input, never seen in the generator.
// Arithmetic and MMA order are those of conv_tc32_kernel (exact 3-way bf16 split, tc32_common.cuh).
#include "ur_common.cuh"

unsigned long long* g_ur_diag_host = nullptr;   // mapped pinned host memory, 128 words (sgnn_debug_ur_diag reads it)

namespace {

// ---------------------------------------------------------------------------------------------------------------
// PROBE = false: the neighbour table exists, read it.  PROBE = true (sgnn_rulebook_submanifold_plan): rulebook and plan in ONE
// kernel -- thread = row probes its 27 neighbours in the site grid (rulebook_probe27), stores them k-major into `nbr_out`
// (the table the direct mode, the child-mode kernel and the FFMA fallbacks read) and goes straight on to the plan: the table
// is written once and not read back (27 x 4 B per row less traffic and one launch less per site set).
template <bool PROBE>
__global__ void __launch_bounds__(128, 8)
tile_plan_kernel(const int* __restrict__ nbr, long long nbr_stride, long long n_rows, long long n_tiles,
                 int* __restrict__ ucount, int* __restrict__ urows, unsigned short* __restrict__ lidx,
                 GridView g, const int* __restrict__ coords, int* __restrict__ nbr_out) {
  __shared__ unsigned bm[UR_BM_WORDS];
  __shared__ unsigned short pre[UR_BM_WORDS];    // ranks <= 27 * 128
  __shared__ int red_min[4], red_max[4], warp_sum[4];
  __shared__ int s_min, s_max;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long j = tile * 128 + tid;
    int idx[27];
    int mn = 0x7fffffff, mx = -1;
    if (PROBE) {
      if (j < n_rows) {
        rulebook_probe27(g, __ldg(reinterpret_cast<const int4*>(coords) + j), idx);
#pragma unroll
        for (int k = 0; k < 27; ++k) nbr_out[(long long)k * nbr_stride + j] = idx[k];
      } else {
#pragma unroll
        for (int k = 0; k < 27; ++k) idx[k] = -1;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 27; ++k) idx[k] = j < n_rows ? __ldg(nbr + (long long)k * nbr_stride + j) : -1;
    }
#pragma unroll
    for (int k = 0; k < 27; ++k)
      if (idx[k] >= 0) { mn = min(mn, idx[k]); mx = max(mx, idx[k]); }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red_min[warp] = mn; red_max[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
      s_min = min(min(red_min[0], red_min[1]), min(red_min[2], red_min[3]));
      s_max = max(max(red_max[0], red_max[1]), max(red_max[2], red_max[3]));
    }
    __syncthreads();
    const int lo = s_min, hi = s_max;
    unsigned short* my_lidx = lidx + tile * (27 * 128);
    if (hi < 0) {                                   // no neighbour at all (cannot happen for rows < n_rows)
      if (tid == 0) ucount[tile] = 0;
#pragma unroll
      for (int k = 0; k < 27; ++k) my_lidx[k * 128 + tid] = 0xffff;
      __syncthreads();
      continue;
    }
    const long long span = (long long)hi - lo + 1;
    if (span > (long long)UR_BM_WORDS * 32) {       // direct mode
      if (tid == 0) ucount[tile] = -1;
      __syncthreads();
      continue;
    }
    const int nw = (int)((span + 31) >> 5);
    for (int w = tid; w < nw; w += 128) bm[w] = 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 27; ++k)
      if (idx[k] >= 0) {
        const int o = idx[k] - lo;
        atomicOr(&bm[o >> 5], 1u << (o & 31));
      }
    __syncthreads();
    // exclusive scan of the word popcounts: thread t owns words [t*per, (t+1)*per)
    const int per = (nw + 127) >> 7;
    int local = 0;
    for (int i = 0; i < per; ++i) {
      const int w = tid * per + i;
      if (w < nw) local += __popc(bm[w]);
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    int base = incl - local;
    for (int w = 0; w < warp; ++w) base += warp_sum[w];
    const int total = warp_sum[0] + warp_sum[1] + warp_sum[2] + warp_sum[3];
    for (int i = 0; i < per; ++i) {
      const int w = tid * per + i;
      if (w < nw) { pre[w] = (unsigned short)base; base += __popc(bm[w]); }
    }
    __syncthreads();
    if (total > UR_PLAN_CAP) {                      // direct mode
      if (tid == 0) ucount[tile] = -1;
      __syncthreads();
      continue;
    }
    if (tid == 0) ucount[tile] = total;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      unsigned short l = 0xffff;
      if (idx[k] >= 0) {
        const int o = idx[k] - lo;
        l = (unsigned short)(pre[o >> 5] + __popc(bm[o >> 5] & ((1u << (o & 31)) - 1u)));
      }
      my_lidx[k * 128 + tid] = l;
    }
    int* my_rows = urows + tile * UR_PLAN_CAP;
    for (int w = tid; w < nw; w += 128) {
      unsigned bits = bm[w];
      int at = pre[w];
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        my_rows[at++] = lo + w * 32 + b;
      }
    }
    __syncthreads();                                // bm / pre are rebuilt by the next tile
  }
}

// A work ITEM = KG consecutive filter offsets of one (tile, pass); the items of a CTA are numbered G = tp * NI + i (tp = running
// tile-pass count); item G uses A stage G % NST of tensor memory and is produced by warp group G % UR_NG.
template <int Q>
struct UrCfg {
  static constexpr int US = Q == 1 ? 512 : 320;          // distinct rows staged per pass
  static constexpr int PL_BUFS = Q == 1 ? 2 : 1;         // plane buffers
  static constexpr int NRING = Q == 1 ? 8 : 6;           // landing ring, chunks of UR_CHUNK rows
  static constexpr int KG = Q == 1 ? 3 : 1;              // filter offsets per item (3 groups x 2 offsets measured slower: 192 vs 155 us)
  static constexpr int NI = (27 + KG - 1) / KG;          // items per pass
  static constexpr int ST_COLS = KG * Q * 24;            // TMEM columns of an A stage
  // A stages (3 x 72 / 5 x 48 columns).  Stage, use count and producer group of an item are functions of the RUNNING item
  // count G = tp * NI + i (stage G % NST, use G / NST, group G % UR_NG): consecutive uses of a stage are then exactly NST items
  // apart also across pass boundaries, and no waiter can be two mbarrier phases away from its barrier.  (Numbering the
  // stages per pass, i % NST with NI % NST != 0, let a group lap the other at a pass boundary and alias the parity wait: a
  // rare deadlock, reproduced by scratch/ur_protocol_sim.py.)
  static constexpr int NST = 256 / ST_COLS;
  static constexpr int ROWB = 64 * Q;                    // bytes of a landed fp32 row
  static constexpr int NARR = 6 * Q;                     // 16-byte plane arrays: [slice][plane][half]
  static constexpr int ASTR = (((US + 1) * 16 + 127) / 128) * 128 + 64;   // array stride: odd multiple of 64 B
  static constexpr int BANK = 27 * Q * 3 * T32_BBLK;
  static constexpr int PLANES = PL_BUFS * NARR * ASTR;
  static constexpr int RING = NRING * UR_CHUNK * ROWB;
  static constexpr int SMEM = BANK + PLANES + RING + 2 * UR_LIDX_BYTES;
};

#define UR_NG 2                                  // producer warp groups (4 warps each)
#define UR_NPW (4 * UR_NG)                        // producer warps
#define UR_NPT (128 * UR_NG)                      // producer threads
#define UR_THREADS (UR_NPT + 128 + 32 + 64)       // + 4 epilogue warps, the MMA warp, 2 loader warps
#define UR_NBAR (1 + 2 + 2 + 2 + 2 + 8 + 8 + 8 + 8)   // w_full, lidx f/e, acc f/e, ring f/e (<= 8), stage f/e (<= 8)

template <int Q>
__global__ void __launch_bounds__(UR_THREADS, 1)
conv_ur_kernel(Tc32Params p, PlanView plan, long long n_tiles) {
  using C = UrCfg<Q>;
  static_assert(C::NRING <= 8 && C::NST <= 8, "barrier table");
  static_assert(C::NST >= UR_NG && C::NST * C::ST_COLS <= 256, "stage protocol");   // NST >= groups: no parity aliasing
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bars[UR_NBAR];
  __shared__ unsigned tmem_ptr_s;
  const unsigned sm_a = smem_u32(sm);
  const unsigned bank_a = sm_a;                                   // [27][Q][3][512]
  const unsigned planes_a = bank_a + C::BANK;                     // [PL_BUFS][NARR][ASTR]
  const unsigned ring_a = planes_a + C::PLANES;                   // [NRING][UR_CHUNK][ROWB]
  const unsigned lidx_a = ring_a + C::RING;                       // [2][27][128] u16
  const unsigned bar_a = smem_u32(bars);
  const unsigned w_full = bar_a, lidx_full = bar_a + 8, lidx_empty = bar_a + 24, acc_full = bar_a + 40, acc_empty = bar_a + 56,
                 ring_full = bar_a + 72, ring_empty = bar_a + 136, st_full = bar_a + 200, st_empty = bar_a + 264;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 32) {
    mb_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mb_init(lidx_full + 8 * i, 1);
      mb_init(lidx_empty + 8 * i, UR_NPW);
      mb_init(acc_full + 8 * i, 1);
      mb_init(acc_empty + 8 * i, 4);
    }
    for (int i = 0; i < C::NRING; ++i) {
      mb_init(ring_full + 8 * i, 1);
      mb_init(ring_empty + 8 * i, UR_NPW);
    }
    for (int i = 0; i < C::NST; ++i) {
      mb_init(st_full + 8 * i, 4);
      mb_init(st_empty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  // zero rows of every plane array (index US): what an absent tap reads
  for (int i = tid; i < C::PL_BUFS * C::NARR; i += UR_THREADS)
    *reinterpret_cast<uint4*>(sm + C::BANK + (size_t)i * C::ASTR + C::US * 16) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;

  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  constexpr int n_main = 7;
  // bytes of a row the loader may copy: the row may be shorter than the 16 Q channels the kernel works on
  const unsigned row_bytes = (unsigned)(min(16 * Q, p.ld_in) * 4);
  // rows whose stride in memory equals the copied bytes are packed at that stride in the ring, so that a run of consecutive
  // row ids is ONE bulk copy; otherwise (a column view of wider rows) every row is its own copy at the full stride
  const bool coalesce = (unsigned)p.ld_in * 4u == row_bytes;
  const unsigned rstride = coalesce ? row_bytes : (unsigned)C::ROWB;

  if (warp < UR_NPW) {
    // ---------------------------------------------------------------------------------- producers
    constexpr int ROLE_ID = 0;
    const int g = warp >> 2;                            // warp group
    const int r = (warp & 3) * 32 + lane;               // output row inside the tile == TMEM lane
    const int pt = tid;                                 // 0..UR_NPT-1: piece index of the split
    const unsigned lane_base = tmem + ((unsigned)((warp & 3) * 32) << 16) + 256u;
    unsigned ring_it = 0, tp = 0;
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int u = __ldg(plan.ucount + tile);
      const int lb = (int)(tl & 1);
      const bool direct = u < 0;
      const int npass = direct ? 1 : max(1, (u + C::US - 1) / C::US);
      mb_wait(lidx_full + 8 * lb, (unsigned)((tl >> 1) & 1));
      const unsigned my_lidx = lidx_a + (unsigned)(lb * UR_LIDX_BYTES + r * 2);
      const long long j = tile * T32_M + r;
      for (int pass = 0; pass < npass; ++pass, ++tp) {
        const unsigned pl_a = planes_a + (C::PL_BUFS == 2 ? (tp & 1u) * (unsigned)(C::NARR * C::ASTR) : 0u);
        if (!direct) {
          if (C::PL_BUFS == 1) bar_sync(2, UR_NPT);     // every producer has left the previous tap loop
          const int rows_this = min(C::US, u - pass * C::US);
          const int nch = (rows_this + UR_CHUNK - 1) / UR_CHUNK;
          unsigned char* pl = sm + (pl_a - sm_a);
          for (int c = 0; c < nch; ++c, ++ring_it) {
            const unsigned slot = ring_it % C::NRING;
            mb_wait(ring_full + 8 * slot, (ring_it / C::NRING) & 1u);
            const unsigned char* src = sm + (ring_a - sm_a) + (size_t)slot * UR_CHUNK * C::ROWB;
            for (int piece = pt; piece < 256 * Q; piece += UR_NPT) {   // [row in chunk][16-byte piece of the row]
              const int rr = piece / (4 * Q), c4 = piece % (4 * Q);
              float4 v = *reinterpret_cast<const float4*>(src + (size_t)rr * rstride + c4 * 16);
              const int ch = 4 * c4;
              if (ch + 0 >= p.cin) v.x = 0.f;
              if (ch + 1 >= p.cin) v.y = 0.f;
              if (ch + 2 >= p.cin) v.z = 0.f;
              if (ch + 3 >= p.cin) v.w = 0.f;
              uint2 h, m, l;
              split2(v.x, v.y, h.x, m.x, l.x);
              split2(v.z, v.w, h.y, m.y, l.y);
              const int row = c * UR_CHUNK + rr;
              if (row < rows_this) {
                unsigned char* d = pl + (size_t)((c4 >> 2) * 6 + ((c4 >> 1) & 1)) * C::ASTR + row * 16 + (c4 & 1) * 8;
                *reinterpret_cast<uint2*>(d) = h;
                *reinterpret_cast<uint2*>(d + 2 * C::ASTR) = m;
                *reinterpret_cast<uint2*>(d + 4 * C::ASTR) = l;
              }
            }
            __syncwarp();
            if (lane == 0) mb_arrive(ring_empty + 8 * slot);
          }
          bar_sync(1, UR_NPT);                          // planes of this pass complete
        }
        const unsigned base = (unsigned)(pass * C::US);
        if (!direct) {
#pragma unroll 1
          for (int i = 0; i < C::NI; ++i) {
            const unsigned G = tp * (unsigned)C::NI + (unsigned)i;          // running item count
            if (G % UR_NG != (unsigned)g) continue;
            const unsigned s = G % C::NST, n = G / C::NST;                  // A stage, uses of it before this one
            unsigned rg[C::KG][Q][3][8];
            unsigned la[C::KG];
#pragma unroll
            for (int kk = 0; kk < C::KG; ++kk)
              if (i * C::KG + kk < 27) la[kk] = lds16(my_lidx + (unsigned)((i * C::KG + kk) * 256));
#pragma unroll
            for (int kk = 0; kk < C::KG; ++kk)
              if (i * C::KG + kk < 27) {
                const unsigned lw = la[kk] - base;                            // absent (0xFFFF) / other pass: out of range
                const unsigned here = lw < (unsigned)C::US ? 1u : 0u;
                const unsigned a = pl_a + lw * 16;
#pragma unroll
                for (int q = 0; q < Q; ++q)
#pragma unroll
                  for (int x = 0; x < 3; ++x) {
                    const uint4 lo4 = lds128_if(a + (unsigned)((q * 6 + x * 2) * C::ASTR), here);
                    const uint4 hi4 = lds128_if(a + (unsigned)((q * 6 + x * 2 + 1) * C::ASTR), here);
                    rg[kk][q][x][0] = lo4.x; rg[kk][q][x][1] = lo4.y; rg[kk][q][x][2] = lo4.z; rg[kk][q][x][3] = lo4.w;
                    rg[kk][q][x][4] = hi4.x; rg[kk][q][x][5] = hi4.y; rg[kk][q][x][6] = hi4.z; rg[kk][q][x][7] = hi4.w;
                  }
              }
            if (n > 0) mb_wait(st_empty + 8 * s, (n - 1) & 1u);                     // the MMAs that read this A stage completed
            asm volatile("tcgen05.fence::after_thread_sync;" ::);
            const unsigned a_stage = lane_base + s * (unsigned)C::ST_COLS;
#pragma unroll
            for (int kk = 0; kk < C::KG; ++kk)
              if (i * C::KG + kk < 27) {
#pragma unroll
                for (int q = 0; q < Q; ++q)
#pragma unroll
                  for (int x = 0; x < 3; ++x) tmem_st8(a_stage + (unsigned)((kk * Q + q) * 24 + x * 8), rg[kk][q][x]);
              }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::);
            __syncwarp();
            if (lane == 0) mb_arrive(st_full + 8 * s);
          }
        } else {
          // DIRECT mode: the tile's rows straight from global memory, split in registers (same item protocol)
#pragma unroll 1
          for (int i = 0; i < C::NI; ++i) {
            const unsigned G = tp * (unsigned)C::NI + (unsigned)i;
            if (G % UR_NG != (unsigned)g) continue;
            const unsigned s = G % C::NST, n = G / C::NST;
            if (n > 0) mb_wait(st_empty + 8 * s, (n - 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::);
            const unsigned a_stage = lane_base + s * (unsigned)C::ST_COLS;
#pragma unroll 1
            for (int kk = 0; kk < C::KG; ++kk) {
              const int k = i * C::KG + kk;
              if (k >= 27) break;
              const int idx = j < p.n_rows ? __ldg(p.nbr + (long long)k * p.nbr_stride + j) : -1;
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                float x0[8], x1[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { x0[e] = 0.f; x1[e] = 0.f; }
                if (idx >= 0) {
                  const float* src = p.in + (long long)idx * p.ld_in;
                  load8<false>(src, 16 * q, p.cin, x0);
                  load8<false>(src, 16 * q + 8, p.cin, x1);
                }
                split16_tmem(x0, x1, a_stage + (unsigned)((kk * Q + q) * 24));
              }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::);
            __syncwarp();
            if (lane == 0) mb_arrive(st_full + 8 * s);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mb_arrive(lidx_empty + 8 * lb);
    }
  } else if (warp < UR_NPW + 4) {
    // ---------------------------------------------------------------------------------- epilogue
    constexpr int ROLE_ID = 1;
    const int qd = warp & 3;
    const unsigned lane_base = tmem + ((unsigned)(qd * 32) << 16);
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const int ab = (int)(tl & 1);
      mb_wait(acc_full + 8 * ab, (unsigned)((tl >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      unsigned v[16], vc[16];
      const unsigned acc = lane_base + (unsigned)(ab * T32_COLS);
      tmem_ld16(acc + T32_CORR, v);
      for (int a = n_main - 1; a >= 0; --a) {
        tmem_ld16(acc + 16u * (unsigned)a, vc);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(vc[c]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      __syncwarp();
      if (lane == 0) mb_arrive(acc_empty + 8 * ab);
      const long long j = (blockIdx.x + tl * gridDim.x) * T32_M + qd * 32 + lane;
      if (j < p.n_rows) epilogue_row16(p, v, j);
    }
  } else if (warp == UR_NPW + 4) {
    // ---------------------------------------------------------------------------------- MMA issuer
    constexpr int ROLE_ID = 2;
    mb_wait(w_full, 0u);
    const unsigned long long bdesc0 = umma_desc(bank_a);          // + 32 per 512-byte B block
    unsigned tp = 0;
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int u = __ldg(plan.ucount + tile);
      const int npass = u < 0 ? 1 : max(1, (u + C::US - 1) / C::US);
      const int ab = (int)(tl & 1);
      if (tl >= 2) mb_wait(acc_empty + 8 * ab, (unsigned)(((tl >> 1) - 1) & 1));   // epilogue of tile tl-2 drained
      const unsigned acc = tmem + (unsigned)(ab * T32_COLS);
      for (int pass = 0; pass < npass; ++pass, ++tp) {
        const bool first = pass == 0;
#pragma unroll 1
        for (int i = 0; i < C::NI; ++i) {
          const unsigned G = tp * (unsigned)C::NI + (unsigned)i;
          const unsigned s = G % C::NST, n = G / C::NST;
          mb_wait(st_full + 8 * s, n & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::);
          if (elect_one()) {
            const unsigned a_stage = tmem + 256u + s * (unsigned)C::ST_COLS;
#pragma unroll
            for (int kk = 0; kk < C::KG; ++kk) {
              const int k = i * C::KG + kk;
              if (k < 27) {
#pragma unroll
                for (int qc = 0; qc < Q; ++qc) {
                  const unsigned a = a_stage + (unsigned)((kk * Q + qc) * 24);           // planes at +0, +8, +16 columns
                  const unsigned long long b = bdesc0 + (unsigned long long)((k * Q + qc) * 3 * 32);   // planes at +32, +64
                  const bool acc_corr = !(first && k == 0 && qc == 0);
                  const bool acc_main = !(first && (k & 3) == 0 && qc == 0);
                  mma_ts(acc + T32_CORR, a + 16u, b, acc_corr);                 // x2 w0
                  mma_ts(acc + T32_CORR, a + 8u, b + 32, true);                 // x1 w1
                  mma_ts(acc + T32_CORR, a, b + 64, true);                      // x0 w2
                  mma_ts(acc + T32_CORR, a + 8u, b, true);                      // x1 w0
                  mma_ts(acc + T32_CORR, a, b + 32, true);                      // x0 w1
                  mma_ts(acc + 16u * (unsigned)(k >> 2), a, b, acc_main);       // x0 w0
                }
              }
            }
            mma_commit_a(st_empty + 8 * s);
            if (pass == npass - 1 && i == C::NI - 1) mma_commit_a(acc_full + 8 * ab);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------------------------------------------------------------------------- loaders (TMA), the last two warps
    constexpr int ROLE_ID = 3;
    const unsigned ldr = (unsigned)(warp - (UR_NPW + 5));           // two loader warps take alternate chunks; warp 13 also the rest
    if (lane == 0 && ldr == 0) {
      mb_expect_tx(w_full, (unsigned)C::BANK);
      bulk_g2s(bank_a, p.wsplit, (unsigned)C::BANK, w_full);
    }
    unsigned ring_it = 0;
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int u = __ldg(plan.ucount + tile);
      const int lb = (int)(tl & 1);
      if (ldr == 0 && tl >= 2) mb_wait(lidx_empty + 8 * lb, (unsigned)(((tl >> 1) - 1) & 1));
      if (lane == 0 && ldr == 0) {
        mb_expect_tx(lidx_full + 8 * lb, (unsigned)UR_LIDX_BYTES);
        bulk_g2s(lidx_a + lb * UR_LIDX_BYTES, plan.lidx + tile * (27 * 128), (unsigned)UR_LIDX_BYTES, lidx_full + 8 * lb);
      }
      if (u <= 0) continue;
      const int* rows = plan.urows + tile * UR_PLAN_CAP;
      const int npass = (u + C::US - 1) / C::US;
      for (int pass = 0; pass < npass; ++pass) {
        const int rows_this = min(C::US, u - pass * C::US);
        const int nch = (rows_this + UR_CHUNK - 1) / UR_CHUNK;
        for (int c = 0; c < nch; ++c, ++ring_it) {
          if ((ring_it & 1u) != ldr) continue;
          const unsigned slot = ring_it % C::NRING;
          const int cnt = min(UR_CHUNK, rows_this - c * UR_CHUNK);
          const int first = pass * C::US + c * UR_CHUNK;
          const int id[2] = {lane < cnt ? __ldg(rows + first + lane) : -1, lane + 32 < cnt ? __ldg(rows + first + lane + 32) : -1};
          if (ring_it >= C::NRING) mb_wait(ring_empty + 8 * slot, (ring_it / C::NRING - 1) & 1u);
          if (lane == 0) mb_expect_tx(ring_full + 8 * slot, (unsigned)cnt * row_bytes);
          __syncwarp();
          const unsigned dst = ring_a + slot * (unsigned)(UR_CHUNK * C::ROWB);
          // consecutive row ids are consecutive in memory when the row stride equals the copied bytes: one bulk copy per RUN
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cnt_h = min(32, max(0, cnt - 32 * h));
            const int prev = __shfl_up_sync(0xffffffffu, id[h], 1);
            const bool start = lane < cnt_h && (lane == 0 || !coalesce || id[h] != prev + 1);
            const unsigned mask = __ballot_sync(0xffffffffu, start);
            if (start) {
              const unsigned higher = mask & ~((2u << lane) - 1u);
              const int end = higher ? __ffs(higher) - 1 : cnt_h;
              bulk_g2s(dst + (unsigned)(32 * h + lane) * rstride, p.in + (long long)id[h] * p.ld_in,
                       (unsigned)(end - lane) * row_bytes, ring_full + 8 * slot);
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}


template <int Q>
int launch_ur(const Tc32Params& p, const PlanView& plan, cudaStream_t st) {
  using C = UrCfg<Q>;
  { const int rc = ur_diag_init(); if (rc) return rc; }
  int dev = 0;
  SGNN_CUDA(cudaGetDevice(&dev));
  static bool attr_set[64] = {};
  if (dev < 0 || dev >= 64) return SGNN_E_INVALID;
  if (!attr_set[dev]) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_ur_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SGNN_CUDA(cudaFuncSetAttribute(conv_ur_kernel<Q>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set[dev] = true;
  }
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = 148;
  if (grid > tiles) grid = tiles;
  conv_ur_kernel<Q><<<(int)grid, UR_THREADS, C::SMEM, st>>>(p, plan, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

}  // namespace

// Diagnostic record of a watchdog trap in conv_ur_kernel (128 words; word 0 = recording block + 1, then 4 words per warp: line<<32|parity, tid<<32|barrier offset, barrier state).
extern "C" int sgnn_debug_ur_diag(uint64_t* out64) {
  if (!out64) return SGNN_E_INVALID;
  for (int i = 0; i < 128; ++i) out64[i] = g_ur_diag_host ? g_ur_diag_host[i] : 0;
  return SGNN_OK;
}

extern "C" size_t sgnn_tile_plan_bytes(int64_t n_rows) {
  if (n_rows <= 0) return 256;
  const size_t tiles = (size_t)((n_rows + 127) / 128);
  return plan_align(tiles * 4) + plan_align(tiles * UR_PLAN_CAP * 4) + plan_align(tiles * UR_LIDX_BYTES);
}

extern "C" int sgnn_tile_plan_build(const int32_t* nbr, int64_t nbr_stride, int64_t n_rows, void* plan, size_t plan_bytes,
                                    void* stream) {
  if (n_rows < 0 || (n_rows > 0 && (!nbr || !plan))) return SGNN_E_INVALID;
  if (n_rows == 0) return SGNN_OK;
  if (plan_bytes < sgnn_tile_plan_bytes(n_rows)) return SGNN_E_NOMEM;
  if (!al(plan, 256)) return SGNN_E_ALIGN;
  const long long tiles = (n_rows + 127) / 128;
  PlanView v = plan_view(plan, tiles);
  long long grid = tiles < 148 * 9 ? tiles : 148 * 9;
  tile_plan_kernel<false><<<(int)grid, 128, 0, (cudaStream_t)stream>>>(nbr, nbr_stride, n_rows, tiles, (int*)v.ucount, (int*)v.urows,
                                                                        (unsigned short*)v.lidx, GridView(), nullptr, nullptr);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_rulebook_submanifold_plan(const SgnnGrid* g, const int32_t* coords, int64_t n, int32_t* nbr, void* plan,
                                              size_t plan_bytes, void* stream) {
  if (!g || !g->mask || !g->prefix || n < 0 || (n > 0 && (!coords || !nbr || !plan))) return SGNN_E_INVALID;
  if (n == 0) return SGNN_OK;
  if (n * 27 > 0x7fffffff00LL) return SGNN_E_TOO_LARGE;
  if (plan_bytes < sgnn_tile_plan_bytes(n)) return SGNN_E_NOMEM;
  if (!al(plan, 256)) return SGNN_E_ALIGN;
  const long long tiles = (n + 127) / 128;
  PlanView v = plan_view(plan, tiles);
  long long grid = tiles < 148 * 8 ? tiles : 148 * 8;
  tile_plan_kernel<true><<<(int)grid, 128, 0, (cudaStream_t)stream>>>(nullptr, n, n, tiles, (int*)v.ucount, (int*)v.urows,
                                                                       (unsigned short*)v.lidx, make_view(g), coords, nbr);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_conv_forward_tc32_ur(const SgnnConvArgs* a, const void* plan, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  if (!a || a->n_out < 0 || a->cin <= 0 || !a->weight) return SGNN_E_INVALID;
  if (a->dtype != SGNN_F32 || (a->cout != 16 && a->cout != 12 && a->cout != 8) || a->cin > 32 || a->K != 27 || a->child_mode)
    return SGNN_E_UNSUPPORTED;
  if (a->n_out == 0) return SGNN_OK;
  if (!a->a.out && !a->b.out) return SGNN_E_INVALID;
  if (!a->in || !a->nbr || !workspace || !plan) return SGNN_E_INVALID;
  if (workspace_bytes < sgnn_conv_tc32_workspace_bytes(a->K, a->cin, 0)) return SGNN_E_NOMEM;
  const SgnnEpilogue* eps[2] = {&a->a, &a->b};
  for (int i = 0; i < 2; ++i) {
    const SgnnEpilogue& e = *eps[i];
    if (!e.out) continue;
    if ((e.scale == nullptr) != (e.shift == nullptr)) return SGNN_E_INVALID;
    if (!al(e.out, 16) || (e.ld & 3) || (e.scale && (!al(e.scale, 16) || !al(e.shift, 16)))) return SGNN_E_ALIGN;
  }
  if (!al(a->in, 16) || (a->ld_in & 3) || a->ld_in < a->cin || !al(workspace, 16) || !al(plan, 256)) return SGNN_E_ALIGN;
  if (a->residual && (!al(a->residual, 16) || (a->ld_res & 3))) return SGNN_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int Q = (a->cin + 15) / 16;
  Tc32Params p;
  p.in = (const float*)a->in; p.ld_in = a->ld_in; p.cin = a->cin; p.cout = a->cout;
  p.nbr = a->nbr; p.nbr_stride = a->nbr_stride; p.K = a->K;
  p.wsplit = (const unsigned char*)workspace;
  p.planes = nullptr; p.n_in = a->n_in;
  p.n_rows = a->n_out;
  p.residual = (const float*)a->residual; p.ld_res = a->ld_res;
  p.out_a = (float*)a->a.out; p.ld_a = a->a.ld; p.relu_a = a->a.relu; p.scale_a = a->a.scale; p.shift_a = a->a.shift;
  p.out_b = (float*)a->b.out; p.ld_b = a->b.ld; p.relu_b = a->b.relu; p.scale_b = a->b.scale; p.shift_b = a->b.shift;
  if (!(a->flags & SGNN_CONV_PREPARED)) {
    const int total = a->K * Q * 256;
    tc32_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float*)a->weight, a->K, a->cin, a->cout, Q, (unsigned char*)workspace);
    SGNN_CHECK_LAUNCH();
  }
  const PlanView v = plan_view(plan, (a->n_out + 127) / 128);
  return Q == 1 ? launch_ur<1>(p, v, st) : launch_ur<2>(p, v, st);
}
