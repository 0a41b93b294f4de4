// conv_urc.cu -- generative upsampling (child-mode 3^3 convolution, reference torch/model.py:192-207,224-225: the 8
// children of every kept site, features replicated x8, then SubmanifoldConvolution 48 -> 16) on tcgen05 with every
// DISTINCT parent row of a tile staged once by TMA.  SURVEY §8 row a9.
//
// Arithmetic as conv_tc32_child_ws_kernel (conv_tc32.cu): because all 8 children of a kept site exist,
//   out[8p+c] = sum_e x[nbr(p,e)] W'_c[e],   W'_c[e] = sum_{d: floor((c+d)/2) = e} W[d]   (64 (e,c) pairs, pre-summed in fp64)
// on an exact 3-way bf16 split of features and filters.  What changes is the data flow -- the round-1 kernel gathered
// 11 present parent offsets per parent from L2 for 1.6 distinct rows (6.9x amplification, profiles/r01_*child*):
//   * the tile plan of the PARENT site set (conv_ur.cu: sorted distinct rows + 16-bit local indices per 128-parent tile) is
//     shared with the parents' own unique-row convolutions;
//   * loader warp 0: TMA (cp.async.bulk, one per run of consecutive rows) of the tile's distinct parent rows (192 B each) into
//     a ring of 64-row chunks, and of its index block; loader warp 1: TMA of the pre-summed filters of each round (1, 2 or 4
//     children x 4.6 KB) into a 16-slot ring that runs ~7 rounds ahead of the MMAs -- the 295 KB bank does not fit in shared
//     memory, and fetched round by round in lock-step with the 3 A stages the L2 latency was exposed (240 us -> see DESIGN);
//   * producers (8 warps): split each landed row once into bf16 planes, then per ROUND (parent offset e; the centre offset,
//     read by all 8 children, takes two rounds) move row lidx[e][r] shared memory -> tensor memory (18 LDS.128 + 9 tcgen05.st);
//   * MMA warp: children of a round that are consecutive in the accumulator (z-major child index) share ONE tcgen05.mma of
//     N = 16 x run length (the filter blocks of a round are stored child-minor so that a run is contiguous): 792 MMAs per
//     tile-pass instead of 1152.  The two centre rounds come FIRST and overwrite the accumulators (they touch all 8
//     children), every later round accumulates;
//   * epilogue (4 warps): 8 children x (main + correction) accumulators -> 512 contiguous bytes per parent.
// Barrier protocol, watchdog and the multi-pass / DIRECT fallbacks as in conv_ur.cu.
#include "ur_common.cuh"

namespace {

struct UrcRound {
  int e, np, nruns, w_off;               // parent offset, children in the round, MMA runs, first (e,c) pair of the round
  int child[4];                          // children, ascending
  int run_c0[4], run_j0[4], run_len[4];  // runs of consecutive children: first child, its index in the round, length
  int slot, dep;                         // first 4608-byte slot in the filter ring; rounds back to the last item whose slots it reuses
};
#define URC_ROUNDS 28
__constant__ UrcRound c_rounds[URC_ROUNDS];

#define URC_WSLOTS 16                    // filter ring: 16 slots of one (e,c) pair; 64 pairs per pass = 4 laps exactly

// Round order: the two centre rounds (children 0-3, 4-7), the 6 face offsets (4 children each), the 12 edge offsets (2), the 8
// corner offsets (1).  Sizes descend, so a round's slots never straddle the end of the ring, and the ring position of a round
// is the same in every pass.
void build_rounds(UrcRound* r) {
  int order[URC_ROUNDS], n = 0;
  order[n++] = 13; order[n++] = 13;
  for (int want = 4; want >= 1; want >>= 1)
    for (int e = 0; e < 27; ++e) {
      if (e == 13) continue;
      int np = 0;
      for (int c = 0; c < 8; ++c) np += child_uses(c, e) ? 1 : 0;
      if (np == want) order[n++] = e;
    }
  int w_off = 0;
  for (int i = 0; i < URC_ROUNDS; ++i) {
    UrcRound& R = r[i];
    R.e = order[i];
    R.np = 0;
    for (int c = 0; c < 8; ++c) {
      if (!child_uses(c, R.e)) continue;
      if (R.e == 13 && (c >> 2) != i) continue;    // centre: children 0-3 in round 0, 4-7 in round 1
      R.child[R.np++] = c;
    }
    R.nruns = 0;
    for (int j = 0; j < R.np; ++j) {
      if (j > 0 && R.child[j] == R.child[j - 1] + 1) { ++R.run_len[R.nruns - 1]; continue; }
      R.run_c0[R.nruns] = R.child[j]; R.run_j0[R.nruns] = j; R.run_len[R.nruns] = 1; ++R.nruns;
    }
    R.w_off = w_off;
    R.slot = w_off % URC_WSLOTS;
    w_off += R.np;
  }
  // dep: item G may overwrite its slots once item G - dep (the LAST earlier item touching any of them) has been consumed
  for (int i = 0; i < URC_ROUNDS; ++i) {
    int dep = 0;
    for (int back = 1; back <= URC_ROUNDS && !dep; ++back) {
      const UrcRound& P = r[((i - back) % URC_ROUNDS + URC_ROUNDS) % URC_ROUNDS];
      if (P.slot < r[i].slot + r[i].np && r[i].slot < P.slot + P.np) dep = back;
    }
    r[i].dep = dep;
  }
}

#define URC_PAIR_BYTES (3 * 3 * T32_BBLK)   // 4608: one (e, c) pair = 3 slices x 3 planes x 512 B

// pre-summed child filters in ROUND order: [round][slice][plane][child of the round][512 B]
__global__ void urc_prep_kernel(const float* __restrict__ w, int cin, unsigned char* __restrict__ out) {
  const int total = 64 * 3 * 256;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx & 15, cil = (idx >> 4) & 15, pq = idx >> 8, qc = pq % 3, pair = pq / 3;
    int r = 0;
    while (r + 1 < URC_ROUNDS && c_rounds[r + 1].w_off <= pair) ++r;
    const UrcRound& R = c_rounds[r];
    const int j = pair - R.w_off, c = R.child[j], e = R.e;
    const int ci = qc * 16 + cil;
    double s = 0.0;
    if (ci < cin)
      for (int d = 0; d < 27; ++d)
        if (child_parent_offset(c, d) == e) s += (double)__ldg(w + ((size_t)d * cin + ci) * 16 + co);
    // planes of this (slice, child): plane stride inside the round = np * 512
    unsigned char* base = out + (size_t)R.w_off * URC_PAIR_BYTES + (size_t)(qc * 3 * R.np + j) * T32_BBLK;
    const float v = (float)s;
    const unsigned u = __float_as_uint(v);
    const float r1 = v - __uint_as_float(u & 0xffff0000u);
    const unsigned u1 = __float_as_uint(r1);
    const float r2 = r1 - __uint_as_float(u1 & 0xffff0000u);
    const int off = (co >> 3) * 256 + (cil >> 3) * 128 + (co & 7) * 16 + (cil & 7) * 2;
    *reinterpret_cast<unsigned short*>(base + off) = (unsigned short)(u >> 16);
    *reinterpret_cast<unsigned short*>(base + (size_t)R.np * T32_BBLK + off) = (unsigned short)(u1 >> 16);
    *reinterpret_cast<unsigned short*>(base + (size_t)2 * R.np * T32_BBLK + off) = (unsigned short)(__float_as_uint(r2) >> 16);
  }
}

struct UrcCfg {
  static constexpr int Q = 3;
  static constexpr int US = 256;                         // distinct parent rows staged per pass
  static constexpr int NRING = 4;                        // landing ring, chunks of UR_CHUNK rows
  static constexpr int NST = 3;                          // A stages in tensor memory
  static constexpr int ST_COLS = Q * 24;                 // 72
  static constexpr int ROWB = 64 * Q;                    // 192
  static constexpr int NARR = 6 * Q;                     // 18
  static constexpr int ASTR = (((US + 1) * 16 + 127) / 128) * 128 + 64;
  static constexpr int PLANES = NARR * ASTR;
  static constexpr int RING = NRING * UR_CHUNK * ROWB;
  static constexpr int WRING = URC_WSLOTS * URC_PAIR_BYTES;   // filter ring
  static constexpr int SMEM = PLANES + RING + WRING + 2 * UR_LIDX_BYTES;
};

#define URC_NG 2
#define URC_NPW (4 * URC_NG)
#define URC_NPT (128 * URC_NG)
#define URC_THREADS (URC_NPT + 128 + 32 + 64)
#define URC_NWB 16                       // filter barriers: one pair per item in flight (items run <= 12 ahead)
#define URC_NBAR (2 + 2 + 1 + 1 + 4 + 4 + 3 + 3 + 2 * URC_NWB)   // lidx f/e, acc f/e, ring f/e, stage f/e, filter f/e
#define URC_IDESC(N) ((1u << 4) | (1u << 7) | (1u << 10) | (((unsigned)(N) >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void mma_ts_n(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned idesc, bool acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc ? 1u : 0u));
}

__global__ void __launch_bounds__(URC_THREADS, 1)
conv_urc_kernel(Tc32Params p, PlanView plan, long long n_tiles) {
  using C = UrcCfg;
  constexpr int Q = C::Q;
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bars[URC_NBAR];
  __shared__ unsigned tmem_ptr_s;
  const unsigned sm_a = smem_u32(sm);
  const unsigned planes_a = sm_a;                                 // [NARR][ASTR]
  const unsigned ring_a = planes_a + C::PLANES;                   // [NRING][UR_CHUNK][ROWB]
  const unsigned wst_a = ring_a + C::RING;                        // [URC_WSLOTS][4608]
  const unsigned lidx_a = wst_a + C::WRING;                       // [2][27][128] u16
  const unsigned bar_a = smem_u32(bars);
  const unsigned lidx_full = bar_a, lidx_empty = bar_a + 16, acc_full = bar_a + 32, acc_empty = bar_a + 40, ring_full = bar_a + 48,
                 ring_empty = bar_a + 80, st_full = bar_a + 112, st_empty = bar_a + 136, w_full = bar_a + 160,
                 w_empty = bar_a + 160 + 8 * URC_NWB;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) {
      mb_init(lidx_full + 8 * i, 1);
      mb_init(lidx_empty + 8 * i, URC_NPW);
    }
    mb_init(acc_full, 1);
    mb_init(acc_empty, 4);
    for (int i = 0; i < C::NRING; ++i) {
      mb_init(ring_full + 8 * i, 1);
      mb_init(ring_empty + 8 * i, URC_NPW);
    }
    for (int i = 0; i < C::NST; ++i) {
      mb_init(st_full + 8 * i, 4);
      mb_init(st_empty + 8 * i, 1);
    }
    for (int i = 0; i < URC_NWB; ++i) {
      mb_init(w_full + 8 * i, 1);
      mb_init(w_empty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;            // columns [0,128) main, [128,256) correction accumulators, [256,472) A stages

  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const unsigned row_bytes = (unsigned)(min(16 * Q, p.ld_in) * 4);
  const bool coalesce = (unsigned)p.ld_in * 4u == row_bytes;
  const unsigned rstride = coalesce ? row_bytes : (unsigned)C::ROWB;

  if (warp < URC_NPW) {
    // ---------------------------------------------------------------------------------- producers
    constexpr int ROLE_ID = 0;
    const int g = warp >> 2;
    const int r = (warp & 3) * 32 + lane;               // parent row inside the tile == TMEM lane
    const int pt = tid;
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
        if (!direct) {
          bar_sync(2, URC_NPT);                         // every producer has left the previous round loop (single plane buffer)
          const int rows_this = min(C::US, u - pass * C::US);
          const int nch = (rows_this + UR_CHUNK - 1) / UR_CHUNK;
          for (int c = 0; c < nch; ++c, ++ring_it) {
            const unsigned slot = ring_it % C::NRING;
            mb_wait(ring_full + 8 * slot, (ring_it / C::NRING) & 1u);
            const unsigned char* src = sm + (ring_a - sm_a) + (size_t)slot * UR_CHUNK * C::ROWB;
            for (int piece = pt; piece < 256 * Q; piece += URC_NPT) {   // [row in chunk][16-byte piece of the row]
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
                unsigned char* d = sm + (size_t)((c4 >> 2) * 6 + ((c4 >> 1) & 1)) * C::ASTR + row * 16 + (c4 & 1) * 8;
                *reinterpret_cast<uint2*>(d) = h;
                *reinterpret_cast<uint2*>(d + 2 * C::ASTR) = m;
                *reinterpret_cast<uint2*>(d + 4 * C::ASTR) = l;
              }
            }
            __syncwarp();
            if (lane == 0) mb_arrive(ring_empty + 8 * slot);
          }
          bar_sync(1, URC_NPT);                         // planes of this pass complete
        }
        const unsigned base = (unsigned)(pass * C::US);
#pragma unroll 1
        for (int rd = 0; rd < URC_ROUNDS; ++rd) {
          const unsigned G = tp * (unsigned)URC_ROUNDS + (unsigned)rd;
          if (G % URC_NG != (unsigned)g) continue;
          const unsigned s = G % C::NST, n = G / C::NST;
          const int e = c_rounds[rd].e;
          unsigned rg[Q][3][8];
          if (!direct) {
            const unsigned lw = lds16(my_lidx + (unsigned)(e * 256)) - base;   // absent (0xFFFF) / other pass: out of range
            const unsigned here = lw < (unsigned)C::US ? 1u : 0u;
            const unsigned a = planes_a + lw * 16;
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
              for (int x = 0; x < 3; ++x) {
                const uint4 lo4 = lds128_if(a + (unsigned)((q * 6 + x * 2) * C::ASTR), here);
                const uint4 hi4 = lds128_if(a + (unsigned)((q * 6 + x * 2 + 1) * C::ASTR), here);
                rg[q][x][0] = lo4.x; rg[q][x][1] = lo4.y; rg[q][x][2] = lo4.z; rg[q][x][3] = lo4.w;
                rg[q][x][4] = hi4.x; rg[q][x][5] = hi4.y; rg[q][x][6] = hi4.z; rg[q][x][7] = hi4.w;
              }
          } else {
            const int idx = j < p.n_rows ? __ldg(p.nbr + (long long)e * p.nbr_stride + j) : -1;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              float x0[8], x1[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) { x0[i] = 0.f; x1[i] = 0.f; }
              if (idx >= 0) {
                const float* src = p.in + (long long)idx * p.ld_in;
                load8<false>(src, 16 * q, p.cin, x0);
                load8<false>(src, 16 * q + 8, p.cin, x1);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                split2(x0[2 * i], x0[2 * i + 1], rg[q][0][i], rg[q][1][i], rg[q][2][i]);
                split2(x1[2 * i], x1[2 * i + 1], rg[q][0][4 + i], rg[q][1][4 + i], rg[q][2][4 + i]);
              }
            }
          }
          if (n > 0) mb_wait(st_empty + 8 * s, (n - 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::);
          const unsigned a_stage = lane_base + s * (unsigned)C::ST_COLS;
#pragma unroll
          for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int x = 0; x < 3; ++x) tmem_st8(a_stage + (unsigned)(q * 24 + x * 8), rg[q][x]);
          asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::);
          __syncwarp();
          if (lane == 0) mb_arrive(st_full + 8 * s);
        }
      }
      __syncwarp();
      if (lane == 0) mb_arrive(lidx_empty + 8 * lb);
    }
  } else if (warp < URC_NPW + 4) {
    // ---------------------------------------------------------------------------------- epilogue
    constexpr int ROLE_ID = 1;
    const int qd = warp & 3;
    const unsigned lane_base = tmem + ((unsigned)(qd * 32) << 16);
    for (long long tl = 0; tl < my_tiles; ++tl) {
      mb_wait(acc_full, (unsigned)(tl & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      const long long pj = (blockIdx.x + tl * gridDim.x) * T32_M + qd * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        unsigned v[16], vc[16];
        tmem_ld16(lane_base + 16u * c, v);
        tmem_ld16(lane_base + 128u + 16u * c, vc);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(vc[q]));
        if (pj < p.n_rows) epilogue_row16(p, v, pj * 8 + c);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      __syncwarp();
      if (lane == 0) mb_arrive(acc_empty);            // the accumulators may be overwritten by the next tile
    }
  } else if (warp == URC_NPW + 4) {
    // ---------------------------------------------------------------------------------- MMA issuer
    constexpr int ROLE_ID = 2;
    unsigned tp = 0;
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int u = __ldg(plan.ucount + tile);
      const int npass = u < 0 ? 1 : max(1, (u + C::US - 1) / C::US);
      if (tl >= 1) mb_wait(acc_empty, (unsigned)((tl - 1) & 1));     // epilogue of the previous tile drained
      for (int pass = 0; pass < npass; ++pass, ++tp) {
#pragma unroll 1
        for (int rd = 0; rd < URC_ROUNDS; ++rd) {
          const unsigned G = tp * (unsigned)URC_ROUNDS + (unsigned)rd;
          const unsigned s = G % C::NST, n = G / C::NST;
          const unsigned wb = G % URC_NWB, wn = G / URC_NWB;
          mb_wait(st_full + 8 * s, n & 1u);
          mb_wait(w_full + 8 * wb, wn & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::);
          if (elect_one()) {
            const unsigned a_stage = tmem + 256u + s * (unsigned)C::ST_COLS;
            const unsigned long long wdesc = umma_desc(wst_a + (unsigned)(c_rounds[rd].slot * URC_PAIR_BYTES));
            const int np = c_rounds[rd].np, nruns = c_rounds[rd].nruns;
            const bool fresh = pass == 0 && rd < 2;               // the centre rounds of the first pass overwrite
            for (int ri = 0; ri < nruns; ++ri) {
              const int c0 = c_rounds[rd].run_c0[ri], j0 = c_rounds[rd].run_j0[ri], L = c_rounds[rd].run_len[ri];
              const unsigned idesc = URC_IDESC(16 * L);
              const unsigned d_main = tmem + 16u * (unsigned)c0, d_corr = d_main + 128u;
#pragma unroll
              for (int qc = 0; qc < Q; ++qc) {
                const unsigned a = a_stage + (unsigned)(qc * 24);                                    // planes at +0, +8, +16
                const unsigned long long b0 = wdesc + (unsigned long long)(((qc * 3 + 0) * np + j0) * 32);   // 512 B = 32 units
                const unsigned long long b1 = wdesc + (unsigned long long)(((qc * 3 + 1) * np + j0) * 32);
                const unsigned long long b2 = wdesc + (unsigned long long)(((qc * 3 + 2) * np + j0) * 32);
                const bool acc = !(fresh && qc == 0);
                mma_ts_n(d_corr, a + 16u, b0, idesc, acc);      // x2 w0
                mma_ts_n(d_corr, a + 8u, b1, idesc, true);      // x1 w1
                mma_ts_n(d_corr, a, b2, idesc, true);           // x0 w2
                mma_ts_n(d_corr, a + 8u, b0, idesc, true);      // x1 w0
                mma_ts_n(d_corr, a, b1, idesc, true);           // x0 w1
                mma_ts_n(d_main, a, b0, idesc, acc);            // x0 w0
              }
            }
            mma_commit_a(st_empty + 8 * s);
            mma_commit_a(w_empty + 8 * wb);
            if (pass == npass - 1 && rd == URC_ROUNDS - 1) mma_commit_a(acc_full);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == URC_NPW + 5) {
    // ---------------------------------------------------------------------------------- loader 0 (TMA): index blocks + parent rows
    constexpr int ROLE_ID = 3;
    unsigned ring_it = 0;
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int u = __ldg(plan.ucount + tile);
      const int lb = (int)(tl & 1);
      if (tl >= 2) mb_wait(lidx_empty + 8 * lb, (unsigned)(((tl >> 1) - 1) & 1));
      if (lane == 0) {
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
          const unsigned slot = ring_it % C::NRING;
          const int cnt = min(UR_CHUNK, rows_this - c * UR_CHUNK);
          const int first = pass * C::US + c * UR_CHUNK;
          const int id[2] = {lane < cnt ? __ldg(rows + first + lane) : -1, lane + 32 < cnt ? __ldg(rows + first + lane + 32) : -1};
          if (ring_it >= C::NRING) mb_wait(ring_empty + 8 * slot, (ring_it / C::NRING - 1) & 1u);
          if (lane == 0) mb_expect_tx(ring_full + 8 * slot, (unsigned)cnt * row_bytes);
          __syncwarp();
          const unsigned dst = ring_a + slot * (unsigned)(UR_CHUNK * C::ROWB);
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
  } else {
    // ---------------------------------------------------------------------------------- loader 1 (TMA): the round's filters
    constexpr int ROLE_ID = 4;
    unsigned tp = 0;
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int u = __ldg(plan.ucount + tile);
      const int npass = u < 0 ? 1 : max(1, (u + C::US - 1) / C::US);
      for (int pass = 0; pass < npass; ++pass, ++tp) {
#pragma unroll 1
        for (int rd = 0; rd < URC_ROUNDS; ++rd) {
          const unsigned G = tp * (unsigned)URC_ROUNDS + (unsigned)rd;
          const unsigned dep = (unsigned)c_rounds[rd].dep;
          if (G >= dep) {                                          // the last item that used these slots has been consumed
            const unsigned X = G - dep;
            mb_wait(w_empty + 8 * (X % URC_NWB), (X / URC_NWB) & 1u);
          }
          if (lane == 0) {
            const unsigned bytes = (unsigned)(c_rounds[rd].np * URC_PAIR_BYTES);
            const unsigned wb = G % URC_NWB;
            mb_expect_tx(w_full + 8 * wb, bytes);
            bulk_g2s(wst_a + (unsigned)(c_rounds[rd].slot * URC_PAIR_BYTES), p.wsplit + (size_t)c_rounds[rd].w_off * URC_PAIR_BYTES, bytes,
                     w_full + 8 * wb);
          }
          __syncwarp();
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int rounds_init() {          // constant memory is per device
  static bool done[64] = {};
  int dev = 0;
  SGNN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SGNN_E_INVALID;
  if (done[dev]) return SGNN_OK;
  UrcRound r[URC_ROUNDS];
  build_rounds(r);
  SGNN_CUDA(cudaMemcpyToSymbol(c_rounds, r, sizeof(r)));
  done[dev] = true;
  return SGNN_OK;
}

}  // namespace

// Prepared filter bank of the child-mode unique-row kernel (same size as the round-1 layout, different order).
extern "C" int sgnn_conv_urc_prepare(const void* weight, int32_t cin, void* workspace, size_t workspace_bytes, void* stream) {
  if (!weight || !workspace || cin != 48) return SGNN_E_INVALID;
  if (workspace_bytes < (size_t)64 * URC_PAIR_BYTES) return SGNN_E_NOMEM;
  if (!al(workspace, 16)) return SGNN_E_ALIGN;
  { const int rc = rounds_init(); if (rc) return rc; }
  urc_prep_kernel<<<96, 512, 0, (cudaStream_t)stream>>>((const float*)weight, cin, (unsigned char*)workspace);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_conv_forward_tc32_urc(const SgnnConvArgs* a, const void* plan, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  if (!a || a->n_out < 0 || !a->weight) return SGNN_E_INVALID;
  if (a->dtype != SGNN_F32 || a->cout != 16 || a->cin != 48 || a->K != 27 || !a->child_mode || a->residual || (a->n_out & 7))
    return SGNN_E_UNSUPPORTED;
  if (a->n_out == 0) return SGNN_OK;
  if (!a->a.out && !a->b.out) return SGNN_E_INVALID;
  if (!a->in || !a->nbr || !workspace || !plan) return SGNN_E_INVALID;
  if (workspace_bytes < (size_t)64 * URC_PAIR_BYTES) return SGNN_E_NOMEM;
  const SgnnEpilogue* eps[2] = {&a->a, &a->b};
  for (int i = 0; i < 2; ++i) {
    const SgnnEpilogue& e = *eps[i];
    if (!e.out) continue;
    if ((e.scale == nullptr) != (e.shift == nullptr)) return SGNN_E_INVALID;
    if (!al(e.out, 16) || (e.ld & 3) || (e.scale && (!al(e.scale, 16) || !al(e.shift, 16)))) return SGNN_E_ALIGN;
  }
  if (!al(a->in, 16) || (a->ld_in & 3) || a->ld_in < a->cin || !al(workspace, 16) || !al(plan, 256)) return SGNN_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  { const int rc = rounds_init(); if (rc) return rc; }
  { const int rc = ur_diag_init(); if (rc) return rc; }
  Tc32Params p;
  p.in = (const float*)a->in; p.ld_in = a->ld_in; p.cin = a->cin; p.cout = 16;
  p.nbr = a->nbr; p.nbr_stride = a->nbr_stride; p.K = a->K;
  p.wsplit = (const unsigned char*)workspace;
  p.planes = nullptr; p.n_in = a->n_in;
  p.n_rows = a->n_out / 8;
  p.residual = nullptr; p.ld_res = 0;
  p.out_a = (float*)a->a.out; p.ld_a = a->a.ld; p.relu_a = a->a.relu; p.scale_a = a->a.scale; p.shift_a = a->a.shift;
  p.out_b = (float*)a->b.out; p.ld_b = a->b.ld; p.relu_b = a->b.relu; p.scale_b = a->b.scale; p.shift_b = a->b.shift;
  if (!(a->flags & SGNN_CONV_PREPARED)) {
    urc_prep_kernel<<<96, 512, 0, st>>>((const float*)a->weight, a->cin, (unsigned char*)workspace);
    SGNN_CHECK_LAUNCH();
  }
  int dev = 0;
  SGNN_CUDA(cudaGetDevice(&dev));
  static bool attr_set[64] = {};
  if (dev < 0 || dev >= 64) return SGNN_E_INVALID;
  if (!attr_set[dev]) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_urc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UrcCfg::SMEM));
    SGNN_CUDA(cudaFuncSetAttribute(conv_urc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set[dev] = true;
  }
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = 148;
  if (grid > tiles) grid = tiles;
  const PlanView v = plan_view(plan, tiles);
  conv_urc_kernel<<<(int)grid, URC_THREADS, UrcCfg::SMEM, st>>>(p, v, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}
