// conv_tc32.cu -- fp32 sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM) by a 3-way bf16 split,
// SURVEY §8 rows a3 / a4 / a9 for the fp32 generator (BASELINE.json configs[1]); the north_star's "tcgen05 tiles only
// for the per-offset dense (Nactive x Cin).(Cin x Cout) contraction".
//
// Arithmetic.  Every fp32 value v is cut EXACTLY into three bf16 pieces v = v0 + v1 + v2 (8 + 8 + 8 significant bits,
// by truncation: v0 = top 16 bits of v, v1 = top 16 bits of v - v0, v2 = v - v0 - v1, all differences exact in fp32).
// The product x*w is evaluated as the six partial products x_i*w_j with i + j <= 2 (each exact in the fp32 accumulator;
// the three dropped ones are below 2^-24 |x w|), accumulated in fp32 in TMEM (leading products and corrections in
// separate accumulators, added in the epilogue).  The result has fp32 accuracy but the
// tensor-core summation order is not the fmaf chain of conv.cu / oracle/o3.c, so this path is NOT bit-identical to the
// FFMA path: it is selected explicitly (SGNN_GEN_TC32 / sgnn_conv_forward_tc32) and tested to 4e-6 of sum |x||w|.
//
// Layout.  Everything the tensor core reads is in the canonical K-major / no-swizzle core-matrix layout validated by
// conv_tc.cu:  row r, 8-element (16-byte) K chunk c  ->  (r/8)*256 + c*128 + (r%8)*16   (SBO = 256 B, LBO = 128 B).
// An "A block" is 128 rows x 16 bf16 (4096 B), a "B block" 16 (Cout) x 16 (Cin slice) bf16 (512 B).  A filter offset
// with Cin = 16 Q channels uses Q x 3 A blocks (Q slices x 3 split planes), Q x 3 B blocks and 6 Q MMAs
// (tcgen05.mma.cta_group::1.kind::f16, M128 N16 K16).
//
// Regular kernel (submanifold K = 27, strided K = 8): CTA = 128 threads = the 128 output rows of a tile; thread t
// gathers ITS row's neighbour at every filter offset of the current group (one index, 32-byte vector loads), cuts it
// into the three planes in registers and writes 16-byte chunks (8 consecutive lanes = 8 consecutive rows = 128
// contiguous bytes: conflict free).  The rows of the NEXT work item are prefetched into registers while the MMAs of
// the current one run; several CTAs per SM cover the rest of the latency.  Thread 0 issues the MMAs, commit ->
// mbarrier; epilogue tcgen05.ld 32x32b.x16: thread t owns accumulator row t (residual, two affine+ReLU slots).
//
// Child-mode kernel (generative upsampling, model.py:192-207,224-225).  All 8 children of every parent exist, so for
// child position c and filter offset d the neighbour is a child of parent-neighbour e = floor((c+d)/2) and
//   out[8p+c] = sum_d x[p+e(c,d)] W[d] = sum_e x[p+e] W'_c[e],   W'_c[e] = sum_{d: e(c,d)=e} W[d]:
// a 2x2x2 stencil per child over PARENT rows -- 64 (e,c) pairs instead of 216 (d,c) pairs.  A tile is 128 parents
// (M = 128), the accumulators 2 x [128 x (8 children x 16)] fp32 = 256 TMEM columns; per parent offset e the 128 neighbour
// rows (48 channels) are gathered ONCE and multiplied with the pre-summed filter of every child that uses e.
#include "tc32_common.cuh"

namespace {


// pre-summed child filters W'_c[e], pairs ordered by e then c: [64 pairs][3 slices][3 planes][512 B]
__global__ void tc32_prep_child_kernel(const float* __restrict__ w, int cin, unsigned char* __restrict__ out) {
  const int total = 64 * 3 * 256;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx & 15, cil = (idx >> 4) & 15, pq = idx >> 8, qc = pq % 3, pair = pq / 3;
    int e = 0, c = 0, cnt = 0;
    bool found = false;
    for (int ee = 0; ee < 27 && !found; ++ee)
      for (int cc = 0; cc < 8; ++cc)
        if (child_uses(cc, ee)) {
          if (cnt == pair) { e = ee; c = cc; found = true; break; }
          ++cnt;
        }
    const int ci = qc * 16 + cil;
    double s = 0.0;
    if (ci < cin)
      for (int d = 0; d < 27; ++d)
        if (child_parent_offset(c, d) == e) s += (double)__ldg(w + ((size_t)d * cin + ci) * 16 + co);
    store_w_split(out + (size_t)pq * 3 * T32_BBLK, co, cil, (float)s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int Q, int KG, bool A32>
__global__ void __launch_bounds__(128)
conv_tc32_kernel(Tc32Params p, long long n_tiles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  constexpr int A_OFF = Q * 3 * T32_ABLK;   // bytes of A per filter offset
  constexpr int B_OFF = Q * 3 * T32_BBLK;   // bytes of B per filter offset
  unsigned char* As = sm;                   // [KG][Q][3][4096]
  unsigned char* Bs = sm + KG * A_OFF;      // [KG][Q][3][512]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_ptr_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(T32_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;
  unsigned phase = 0;

  const int ngroups = (p.K + KG - 1) / KG;
  const int n_main = (p.K + 3) >> 2;   // main accumulators in use (filter offset k accumulates into column 16 (k >> 2))
  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_items = my_tiles * ngroups;

  float x[KG][2 * Q][8];
  int idx[KG], idx_next[KG];
  const int row_off = (tid >> 3) * 256 + (tid & 7) * 16;   // this thread's row inside an A block

  auto load_idx = [&](long long item, int (&dst)[KG]) {
    const long long tile = blockIdx.x + (item / ngroups) * gridDim.x;
    const int k0 = (int)(item % ngroups) * KG;
    const long long j = tile * T32_M + tid;
#pragma unroll
    for (int kk = 0; kk < KG; ++kk)
      dst[kk] = (k0 + kk < p.K && j < p.n_rows) ? __ldg(p.nbr + (long long)(k0 + kk) * p.nbr_stride + j) : -1;
  };
  auto load_rows = [&]() {
#pragma unroll
    for (int kk = 0; kk < KG; ++kk)
      if (idx[kk] >= 0) {
        const float* src = p.in + (long long)idx[kk] * p.ld_in;
#pragma unroll
        for (int u = 0; u < 2 * Q; ++u) load8<A32>(src, 8 * u, p.cin, x[kk][u]);
      }
  };

  if (n_items > 0) {
    load_idx(0, idx);
    load_rows();
    if (n_items > 1) load_idx(1, idx_next);
  }
  for (long long it = 0; it < n_items; ++it) {
    const long long tile = blockIdx.x + (it / ngroups) * gridDim.x;
    const int g = (int)(it % ngroups);
    const int k0 = g * KG, kg = min(KG, p.K - k0);
    // ---- stage item `it`: copy the prepared filter slices (asynchronously, under the conversion below), split the
    // prefetched rows into the three planes
    {
      const unsigned char* wsrc = p.wsplit + (size_t)k0 * B_OFF;
      for (int i = tid; i < kg * (B_OFF / 16); i += 128) cp16(Bs + i * 16, wsrc + i * 16);
      asm volatile("cp.async.commit_group;\n" ::);
    }
#pragma unroll
    for (int kk = 0; kk < KG; ++kk)
      if (kk < kg) {
#pragma unroll
        for (int u = 0; u < 2 * Q; ++u) {
          unsigned char* dst = As + kk * A_OFF + (u >> 1) * (3 * T32_ABLK) + (u & 1) * 128 + row_off;
          if (idx[kk] >= 0) split8_store(x[kk][u], dst, T32_ABLK);
          else zero_store(dst, T32_ABLK);
        }
      }
    asm volatile("cp.async.wait_group 0;\n" ::);
    asm volatile("fence.proxy.async.shared::cta;" ::);   // generic-proxy writes -> visible to the tensor-core proxy
    __syncthreads();                                      // item staged; previous epilogue's TMEM loads retired
    if (warp == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      if (elect_one()) {
        for (int kk = 0; kk < kg; ++kk) {
          const int k = k0 + kk;
#pragma unroll
          for (int qc = 0; qc < Q; ++qc)
            mma_split6(tmem + 16u * (unsigned)(k >> 2), tmem + T32_CORR, smem_u32(As + kk * A_OFF + qc * 3 * T32_ABLK),
                       smem_u32(Bs + kk * B_OFF + qc * 3 * T32_BBLK), ((k & 3) == 0 && qc == 0) ? 0u : 1u,
                       (k == 0 && qc == 0) ? 0u : 1u);
        }
        mma_commit(&mbar);
      }
      __syncwarp();
    }
    // ---- prefetch the rows of item it+1 (and the indices of it+2) while the tensor core works
    if (it + 1 < n_items) {
#pragma unroll
      for (int kk = 0; kk < KG; ++kk) idx[kk] = idx_next[kk];
      load_rows();
      if (it + 2 < n_items) load_idx(it + 2, idx_next);
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    if (g == ngroups - 1) {
      // sum of the partial accumulators in fp32 round-to-nearest (unbiased), corrections first
      unsigned v[16], vc[16];
      const unsigned lane_base = tmem + ((unsigned)(warp * 32) << 16);
      tmem_ld16(lane_base + T32_CORR, v);
      for (int a = n_main - 1; a >= 0; --a) {
        tmem_ld16(lane_base + 16u * (unsigned)a, vc);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(vc[c]));
      }
      const long long j = tile * T32_M + tid;
      if (j < p.n_rows) epilogue_row16(p, v, j);
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(T32_COLS));
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised version of the regular kernel.  Warps 0-3 (thread t = output row t of the tile) are PRODUCERS:
// they gather + split the neighbour rows of work item i into stage i % 2 of a two-stage shared-memory ring and
// arrive on full[stage]; warp 4 is the MMA ISSUER: it waits for full[stage], issues the item's tcgen05.mma's and
// commits them to empty[stage] (the stage may be refilled when the tensor core has read it) and, after the last item
// of a tile, to acc_full[buf].  The accumulators are double buffered in TMEM (2 x 128 columns), so the producers run
// the epilogue of tile t (wait acc_full, tcgen05.ld, sum of the partial accumulators, stores, arrive acc_empty) one
// item into tile t+1 while the tensor core already works on t+1.  Nobody waits for the tensor core in the steady
// state: conversion, MMA issue and epilogue overlap inside one CTA.
template <int Q, int KG, bool A32>
__global__ void __launch_bounds__(160)
conv_tc32_ws_kernel(Tc32Params p, long long n_tiles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  constexpr int A_OFF = Q * 3 * T32_ABLK;
  constexpr int B_OFF = Q * 3 * T32_BBLK;
  constexpr int STAGE = KG * (A_OFF + B_OFF);   // [A: KG x Q x 3 x 4096 | B: KG x Q x 3 x 512]
  __shared__ __align__(8) unsigned long long full[2], empty[2], acc_full[2], acc_empty[2];
  __shared__ unsigned tmem_ptr_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(2 * T32_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 128);
      mbar_init(&empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;

  const int ngroups = (p.K + KG - 1) / KG;
  const int n_main = (p.K + 3) >> 2;
  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_items = my_tiles * ngroups;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers + epilogue
    float x[KG][2 * Q][8];
    int idx[KG], idx_next[KG];
    const int row_off = (tid >> 3) * 256 + (tid & 7) * 16;
    auto load_idx = [&](long long item, int (&dst)[KG]) {
      const long long tile = blockIdx.x + (item / ngroups) * gridDim.x;
      const int k0 = (int)(item % ngroups) * KG;
      const long long j = tile * T32_M + tid;
#pragma unroll
      for (int kk = 0; kk < KG; ++kk)
        dst[kk] = (k0 + kk < p.K && j < p.n_rows) ? __ldg(p.nbr + (long long)(k0 + kk) * p.nbr_stride + j) : -1;
    };
    auto load_rows = [&]() {
#pragma unroll
      for (int kk = 0; kk < KG; ++kk)
        if (idx[kk] >= 0) {
          const float* src = p.in + (long long)idx[kk] * p.ld_in;
#pragma unroll
          for (int u = 0; u < 2 * Q; ++u) load8<A32>(src, 8 * u, p.cin, x[kk][u]);
        }
    };
    auto epilogue = [&](long long tl) {   // tl: index of the tile in this CTA's sequence
      const int ab = (int)(tl & 1);
      mbar_wait(&acc_full[ab], (unsigned)((tl >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      unsigned v[16], vc[16];
      const unsigned lane_base = tmem + (unsigned)(ab * T32_COLS) + ((unsigned)(warp * 32) << 16);
      tmem_ld16(lane_base + T32_CORR, v);
      for (int a = n_main - 1; a >= 0; --a) {
        tmem_ld16(lane_base + 16u * (unsigned)a, vc);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(vc[c]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      mbar_arrive(&acc_empty[ab]);        // the accumulator buffer may be overwritten by tile tl + 2
      const long long j = (blockIdx.x + tl * gridDim.x) * T32_M + tid;
      if (j < p.n_rows) epilogue_row16(p, v, j);
    };

    if (n_items > 0) {
      load_idx(0, idx);
      load_rows();
      if (n_items > 1) load_idx(1, idx_next);
    }
    for (long long it = 0; it < n_items; ++it) {
      const long long tl = it / ngroups;
      const int g = (int)(it % ngroups);
      const int k0 = g * KG, kg = min(KG, p.K - k0);
      const int s = (int)(it & 1);
      const long long u = it >> 1;
      if (u > 0) mbar_wait(&empty[s], (unsigned)((u - 1) & 1));   // the MMAs that read this stage last have completed
      unsigned char* As = sm + s * STAGE;
      unsigned char* Bs = As + KG * A_OFF;
      {
        const unsigned char* wsrc = p.wsplit + (size_t)k0 * B_OFF;
        for (int i = tid; i < kg * (B_OFF / 16); i += 128) cp16(Bs + i * 16, wsrc + i * 16);
        asm volatile("cp.async.commit_group;\n" ::);
      }
#pragma unroll
      for (int kk = 0; kk < KG; ++kk)
        if (kk < kg) {
#pragma unroll
          for (int uu = 0; uu < 2 * Q; ++uu) {
            unsigned char* dst = As + kk * A_OFF + (uu >> 1) * (3 * T32_ABLK) + (uu & 1) * 128 + row_off;
            if (idx[kk] >= 0) split8_store(x[kk][uu], dst, T32_ABLK);
            else zero_store(dst, T32_ABLK);
          }
        }
      asm volatile("cp.async.wait_group 0;\n" ::);
      asm volatile("fence.proxy.async.shared::cta;" ::);
      mbar_arrive(&full[s]);
      if (it + 1 < n_items) {
#pragma unroll
        for (int kk = 0; kk < KG; ++kk) idx[kk] = idx_next[kk];
        load_rows();
        if (it + 2 < n_items) load_idx(it + 2, idx_next);
      }
      if (g == 0 && tl > 0) epilogue(tl - 1);   // previous tile, one item late: its MMAs are done or nearly so
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    for (long long it = 0; it < n_items; ++it) {
      const long long tl = it / ngroups;
      const int g = (int)(it % ngroups);
      const int k0 = g * KG, kg = min(KG, p.K - k0);
      const int s = (int)(it & 1);
      const int ab = (int)(tl & 1);
      mbar_wait(&full[s], (unsigned)((it >> 1) & 1));
      if (g == 0 && tl >= 2) mbar_wait(&acc_empty[ab], (unsigned)(((tl >> 1) - 1) & 1));   // epilogue of tile tl-2 drained
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      if (elect_one()) {
        const unsigned char* As = sm + s * STAGE;
        const unsigned char* Bs = As + KG * A_OFF;
        const unsigned acc = tmem + (unsigned)(ab * T32_COLS);
        for (int kk = 0; kk < kg; ++kk) {
          const int k = k0 + kk;
#pragma unroll
          for (int qc = 0; qc < Q; ++qc)
            mma_split6(acc + 16u * (unsigned)(k >> 2), acc + T32_CORR, smem_u32(As + kk * A_OFF + qc * 3 * T32_ABLK),
                       smem_u32(Bs + kk * B_OFF + qc * 3 * T32_BBLK), ((k & 3) == 0 && qc == 0) ? 0u : 1u,
                       (k == 0 && qc == 0) ? 0u : 1u);
        }
        mma_commit(&empty[s]);
        if (g == ngroups - 1) mma_commit(&acc_full[ab]);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * T32_COLS));
}

// ---------------------------------------------------------------------------------------------------------------
// v3: the A operand goes through TENSOR MEMORY instead of shared memory (tcgen05.mma with A in TMEM, "TS" form).
// ncu on v2 (profiles/r01_tc32_*): the L1/shared data pipe was the limiter at 79 % -- 578 wavefronts per work item for
// the st.shared of the split planes (64 B per wavefront) plus 669 for the row gathers, and the tensor core re-read the
// same bytes from shared memory.  Here thread t (= output row t = TMEM lane t) writes its row's split planes with
// tcgen05.st.32x32b.x8 (16 bf16 = 8 packed 32-bit columns per plane and 16-channel slice): no shared-memory stores,
// no generic->async proxy fence, no bank conflicts, 256 B/clk TMEM write port.  The whole prepared filter bank is
// resident in shared memory (loaded once per persistent CTA).  Roles as in the warp-specialised kernel: warps 0-3
// produce (rows of item i+1 are in flight while item i is converted: two register sets) and run the epilogue one item
// into the next tile, warp 4 issues the MMAs; TMEM columns: [0,128) accumulators (7 main + correction),
// [128,256) two A stages of 64 columns.
#define T32_ASTAGE_COLS 64u

template <int Q, int KG, bool A32>
__global__ void __launch_bounds__(160)
conv_tc32_tm_kernel(Tc32Params p, long long n_tiles) {
  static_assert(KG * Q * 24 <= 64, "A stage must fit its 64 TMEM columns");
  extern __shared__ __align__(1024) unsigned char sm[];   // prepared filter bank [K][Q][3][512 B]
  constexpr int B_OFF = Q * 3 * T32_BBLK;
  __shared__ __align__(8) unsigned long long full[2], empty[2], acc_full, acc_empty;
  __shared__ unsigned tmem_ptr_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 128);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  // filter bank: plain 16-byte copies (already in the canonical layout), once per CTA
  for (int i = tid; i < p.K * (B_OFF / 16); i += 160) cp16(sm + i * 16, p.wsplit + (size_t)i * 16);
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  asm volatile("fence.proxy.async.shared::cta;" ::);
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;

  const int ngroups = (p.K + KG - 1) / KG;
  const int n_main = (p.K + 3) >> 2;
  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_items = my_tiles * ngroups;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers + epilogue
    float x0[KG][2 * Q][8], x1[KG][2 * Q][8];
    int idxn[KG];                                   // neighbour rows of the NEXT item to load
    const unsigned lane_base = tmem + ((unsigned)(warp * 32) << 16);
    auto load_idx = [&](long long item) {
      const long long tile = blockIdx.x + (item / ngroups) * gridDim.x;
      const int k0 = (int)(item % ngroups) * KG;
      const long long j = tile * T32_M + tid;
#pragma unroll
      for (int kk = 0; kk < KG; ++kk)
        idxn[kk] = (k0 + kk < p.K && j < p.n_rows) ? __ldg(p.nbr + (long long)(k0 + kk) * p.nbr_stride + j) : -1;
    };
    auto load_rows = [&](float (&x)[KG][2 * Q][8]) {   // rows idxn[] -> x (zeros for absent neighbours)
#pragma unroll
      for (int kk = 0; kk < KG; ++kk) {
        if (idxn[kk] >= 0) {
          const float* src = p.in + (long long)idxn[kk] * p.ld_in;
#pragma unroll
          for (int u = 0; u < 2 * Q; ++u) load8<A32>(src, 8 * u, p.cin, x[kk][u]);
        } else {
#pragma unroll
          for (int u = 0; u < 2 * Q; ++u)
#pragma unroll
            for (int e = 0; e < 8; ++e) x[kk][u][e] = 0.f;
        }
      }
    };
    auto epilogue = [&](long long tl) {
      mbar_wait(&acc_full, (unsigned)(tl & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      unsigned v[16], vc[16];
      tmem_ld16(lane_base + T32_CORR, v);
      for (int a = n_main - 1; a >= 0; --a) {
        tmem_ld16(lane_base + 16u * (unsigned)a, vc);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(vc[c]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      mbar_arrive(&acc_empty);
      const long long j = (blockIdx.x + tl * gridDim.x) * T32_M + tid;
      if (j < p.n_rows) epilogue_row16(p, v, j);
    };
    // one pipeline step: rows of item it+1 -> xn (in flight), item `it` (held in xc) -> TMEM stage it % 2
    auto step = [&](long long it, float (&xc)[KG][2 * Q][8], float (&xn)[KG][2 * Q][8]) {
      const long long tl = it / ngroups;
      const int g = (int)(it % ngroups);
      const int kg = min(KG, p.K - g * KG);
      const int s = (int)(it & 1);
      const long long u = it >> 1;
      if (it + 1 < n_items) {
        load_rows(xn);
        if (it + 2 < n_items) load_idx(it + 2);
      }
      if (u > 0) mbar_wait(&empty[s], (unsigned)((u - 1) & 1));   // the MMAs that read this A stage have completed
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      const unsigned a_stage = lane_base + 128u + (unsigned)s * T32_ASTAGE_COLS;
#pragma unroll
      for (int kk = 0; kk < KG; ++kk)
        if (kk < kg) {
#pragma unroll
          for (int qc = 0; qc < Q; ++qc)
            split16_tmem(xc[kk][2 * qc], xc[kk][2 * qc + 1], a_stage + (unsigned)((kk * Q + qc) * 24));
        }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      mbar_arrive(&full[s]);
      if (g == 0 && tl > 0) epilogue(tl - 1);
    };

    if (n_items > 0) {
      load_idx(0);
      load_rows(x0);
      if (n_items > 1) load_idx(1);
    }
    for (long long it = 0; it < n_items; it += 2) {
      step(it, x0, x1);
      if (it + 1 < n_items) step(it + 1, x1, x0);
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    for (long long it = 0; it < n_items; ++it) {
      const long long tl = it / ngroups;
      const int g = (int)(it % ngroups);
      const int k0 = g * KG, kg = min(KG, p.K - k0);
      const int s = (int)(it & 1);
      mbar_wait(&full[s], (unsigned)((it >> 1) & 1));
      if (g == 0 && tl >= 1) mbar_wait(&acc_empty, (unsigned)((tl - 1) & 1));   // epilogue of the previous tile drained
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      if (elect_one()) {
        const unsigned a_stage = tmem + 128u + (unsigned)s * T32_ASTAGE_COLS;
        for (int kk = 0; kk < kg; ++kk) {
          const int k = k0 + kk;
#pragma unroll
          for (int qc = 0; qc < Q; ++qc) {
            const unsigned a = a_stage + (unsigned)((kk * Q + qc) * 24);          // planes at +0, +8, +16 columns
            const unsigned b = smem_u32(sm + (size_t)(k * Q + qc) * 3 * T32_BBLK); // planes at +0, +512, +1024 bytes
            const unsigned first_corr = (k == 0 && qc == 0) ? 0u : 1u;
            const unsigned first_main = ((k & 3) == 0 && qc == 0) ? 0u : 1u;
            mma_bf16_ts(tmem + T32_CORR, a + 16u, umma_desc(b), first_corr);                    // x2 w0
            mma_bf16_ts(tmem + T32_CORR, a + 8u, umma_desc(b + T32_BBLK), 1u);                  // x1 w1
            mma_bf16_ts(tmem + T32_CORR, a, umma_desc(b + 2 * T32_BBLK), 1u);                   // x0 w2
            mma_bf16_ts(tmem + T32_CORR, a + 8u, umma_desc(b), 1u);                             // x1 w0
            mma_bf16_ts(tmem + T32_CORR, a, umma_desc(b + T32_BBLK), 1u);                       // x0 w1
            mma_bf16_ts(tmem + 16u * (unsigned)(k >> 2), a, umma_desc(b), first_main);          // x0 w0
          }
        }
        mma_commit(&empty[s]);
        if (g == ngroups - 1) mma_commit(&acc_full);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---------------------------------------------------------------------------------------------------------------
// v4: as v3, but the input rows are split ONCE per layer instead of once per (row, tap): tc32_split_rows_kernel writes
// the three bf16 planes of the layer's input as separate [n_in][16] arrays (plane-major, so the 32-byte rows of
// raster-adjacent sites are contiguous), and the producers move 3 x 32 bytes per (row, tap, 16-channel slice) from
// global memory to tensor memory (ld.global.v8 -> tcgen05.st.x8): ~20 instructions where v2/v3 spend ~300 on the
// conversion, no shared-memory traffic at all for the A operand.
__global__ void tc32_split_rows_kernel(const float* __restrict__ in, int ld_in, int cin, long long n_in, int Q,
                                       unsigned char* __restrict__ planes) {
  const long long total = n_in * Q;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx % n_in;
    const int qc = (int)(idx / n_in);
    float x0[8], x1[8];
    load8<false>(in + row * ld_in, 16 * qc, cin, x0);
    load8<false>(in + row * ld_in, 16 * qc + 8, cin, x1);
    uint4 h[2], m[2], l[2];
    split2(x0[0], x0[1], h[0].x, m[0].x, l[0].x); split2(x0[2], x0[3], h[0].y, m[0].y, l[0].y);
    split2(x0[4], x0[5], h[0].z, m[0].z, l[0].z); split2(x0[6], x0[7], h[0].w, m[0].w, l[0].w);
    split2(x1[0], x1[1], h[1].x, m[1].x, l[1].x); split2(x1[2], x1[3], h[1].y, m[1].y, l[1].y);
    split2(x1[4], x1[5], h[1].z, m[1].z, l[1].z); split2(x1[6], x1[7], h[1].w, m[1].w, l[1].w);
    uint4* d0 = reinterpret_cast<uint4*>(planes + ((size_t)(0 * Q + qc) * (size_t)n_in + (size_t)row) * 32);
    uint4* d1 = reinterpret_cast<uint4*>(planes + ((size_t)(1 * Q + qc) * (size_t)n_in + (size_t)row) * 32);
    uint4* d2 = reinterpret_cast<uint4*>(planes + ((size_t)(2 * Q + qc) * (size_t)n_in + (size_t)row) * 32);
    d0[0] = h[0]; d0[1] = h[1];
    d1[0] = m[0]; d1[1] = m[1];
    d2[0] = l[0]; d2[1] = l[1];
  }
}


template <int Q, int KG>
__global__ void __launch_bounds__(160)
conv_tc32_pm_kernel(Tc32Params p, long long n_tiles) {
  static_assert(KG * Q * 24 <= 64, "A stage must fit its 64 TMEM columns");
  extern __shared__ __align__(1024) unsigned char sm[];   // prepared filter bank [K][Q][3][512 B]
  constexpr int B_OFF = Q * 3 * T32_BBLK;
  __shared__ __align__(8) unsigned long long full[2], empty[2], acc_full, acc_empty;
  __shared__ unsigned tmem_ptr_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 128);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  // filter bank: plain 16-byte copies (already in the canonical layout), once per CTA
  for (int i = tid; i < p.K * (B_OFF / 16); i += 160) cp16(sm + i * 16, p.wsplit + (size_t)i * 16);
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  asm volatile("fence.proxy.async.shared::cta;" ::);
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;

  const int ngroups = (p.K + KG - 1) / KG;
  const int n_main = (p.K + 3) >> 2;
  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_items = my_tiles * ngroups;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers + epilogue
    unsigned x0[KG][Q][3][8], x1[KG][Q][3][8];   // packed bf16 pairs, straight from the pre-split planes
    int idxn[KG];                                   // neighbour rows of the NEXT item to load
    const unsigned lane_base = tmem + ((unsigned)(warp * 32) << 16);
    auto load_idx = [&](long long item) {
      const long long tile = blockIdx.x + (item / ngroups) * gridDim.x;
      const int k0 = (int)(item % ngroups) * KG;
      const long long j = tile * T32_M + tid;
#pragma unroll
      for (int kk = 0; kk < KG; ++kk)
        idxn[kk] = (k0 + kk < p.K && j < p.n_rows) ? __ldg(p.nbr + (long long)(k0 + kk) * p.nbr_stride + j) : -1;
    };
    auto load_rows = [&](unsigned (&x)[KG][Q][3][8]) {   // rows idxn[] of every plane -> x (zeros for absent neighbours)
#pragma unroll
      for (int kk = 0; kk < KG; ++kk) {
        if (idxn[kk] >= 0) {
#pragma unroll
          for (int qc = 0; qc < Q; ++qc)
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
              const unsigned char* src = p.planes + ((size_t)(pl * Q + qc) * (size_t)p.n_in + (size_t)idxn[kk]) * 32;
              asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(x[kk][qc][pl][0]), "=r"(x[kk][qc][pl][1]), "=r"(x[kk][qc][pl][2]), "=r"(x[kk][qc][pl][3]),
                             "=r"(x[kk][qc][pl][4]), "=r"(x[kk][qc][pl][5]), "=r"(x[kk][qc][pl][6]), "=r"(x[kk][qc][pl][7])
                           : "l"(src));
            }
        } else {
#pragma unroll
          for (int qc = 0; qc < Q; ++qc)
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
              for (int e = 0; e < 8; ++e) x[kk][qc][pl][e] = 0u;
        }
      }
    };
    auto epilogue = [&](long long tl) {
      mbar_wait(&acc_full, (unsigned)(tl & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      unsigned v[16], vc[16];
      tmem_ld16(lane_base + T32_CORR, v);
      for (int a = n_main - 1; a >= 0; --a) {
        tmem_ld16(lane_base + 16u * (unsigned)a, vc);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(vc[c]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      mbar_arrive(&acc_empty);
      const long long j = (blockIdx.x + tl * gridDim.x) * T32_M + tid;
      if (j < p.n_rows) epilogue_row16(p, v, j);
    };
    // one pipeline step: rows of item it+1 -> xn (in flight), item `it` (held in xc) -> TMEM stage it % 2
    auto step = [&](long long it, unsigned (&xc)[KG][Q][3][8], unsigned (&xn)[KG][Q][3][8]) {
      const long long tl = it / ngroups;
      const int g = (int)(it % ngroups);
      const int kg = min(KG, p.K - g * KG);
      const int s = (int)(it & 1);
      const long long u = it >> 1;
      if (it + 1 < n_items) {
        load_rows(xn);
        if (it + 2 < n_items) load_idx(it + 2);
      }
      if (u > 0) mbar_wait(&empty[s], (unsigned)((u - 1) & 1));   // the MMAs that read this A stage have completed
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      const unsigned a_stage = lane_base + 128u + (unsigned)s * T32_ASTAGE_COLS;
#pragma unroll
      for (int kk = 0; kk < KG; ++kk)
        if (kk < kg) {
#pragma unroll
          for (int qc = 0; qc < Q; ++qc)
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) tmem_st8(a_stage + (unsigned)((kk * Q + qc) * 24 + pl * 8), xc[kk][qc][pl]);
        }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      mbar_arrive(&full[s]);
      if (g == 0 && tl > 0) epilogue(tl - 1);
    };

    if (n_items > 0) {
      load_idx(0);
      load_rows(x0);
      if (n_items > 1) load_idx(1);
    }
    for (long long it = 0; it < n_items; it += 2) {
      step(it, x0, x1);
      if (it + 1 < n_items) step(it + 1, x1, x0);
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    for (long long it = 0; it < n_items; ++it) {
      const long long tl = it / ngroups;
      const int g = (int)(it % ngroups);
      const int k0 = g * KG, kg = min(KG, p.K - k0);
      const int s = (int)(it & 1);
      mbar_wait(&full[s], (unsigned)((it >> 1) & 1));
      if (g == 0 && tl >= 1) mbar_wait(&acc_empty, (unsigned)((tl - 1) & 1));   // epilogue of the previous tile drained
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      if (elect_one()) {
        const unsigned a_stage = tmem + 128u + (unsigned)s * T32_ASTAGE_COLS;
        for (int kk = 0; kk < kg; ++kk) {
          const int k = k0 + kk;
#pragma unroll
          for (int qc = 0; qc < Q; ++qc) {
            const unsigned a = a_stage + (unsigned)((kk * Q + qc) * 24);          // planes at +0, +8, +16 columns
            const unsigned b = smem_u32(sm + (size_t)(k * Q + qc) * 3 * T32_BBLK); // planes at +0, +512, +1024 bytes
            const unsigned first_corr = (k == 0 && qc == 0) ? 0u : 1u;
            const unsigned first_main = ((k & 3) == 0 && qc == 0) ? 0u : 1u;
            mma_bf16_ts(tmem + T32_CORR, a + 16u, umma_desc(b), first_corr);                    // x2 w0
            mma_bf16_ts(tmem + T32_CORR, a + 8u, umma_desc(b + T32_BBLK), 1u);                  // x1 w1
            mma_bf16_ts(tmem + T32_CORR, a, umma_desc(b + 2 * T32_BBLK), 1u);                   // x0 w2
            mma_bf16_ts(tmem + T32_CORR, a + 8u, umma_desc(b), 1u);                             // x1 w0
            mma_bf16_ts(tmem + T32_CORR, a, umma_desc(b + T32_BBLK), 1u);                       // x0 w1
            mma_bf16_ts(tmem + 16u * (unsigned)(k >> 2), a, umma_desc(b), first_main);          // x0 w0
          }
        }
        mma_commit(&empty[s]);
        if (g == ngroups - 1) mma_commit(&acc_full);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---------------------------------------------------------------------------------------------------------------
// v5 (EXPERIMENTAL, hook 28, not yet run on a GPU): stage every DISTINCT input row of a tile once.
// scratch/reuse_model.py: with the generator's row order a 128-row tile of the fine levels issues 14 present taps per
// output row but touches only 1.6 distinct input rows per output row -- the kernels above read (and convert) every row
// ~9 times, and the L2->SM gather is their common wall.  Per tile this kernel
//   1. loads the 27 x 128 neighbour indices and de-duplicates them in a shared-memory hash table (atomicCAS on the row
//      id; a second pass numbers the occupied slots = local ids, one list entry per distinct row);
//   2. gathers each distinct row ONCE (64 B), splits it into the three bf16 planes and parks the 96 bytes in shared
//      memory (U_CAP = 512 rows; rows past the cap -- not seen on the fine levels -- take a slow per-tap global path);
//   3. for every tap moves the 128 rows' planes shared memory -> tensor memory (6 LDS.128 + 3 tcgen05.st per row) for
//      the MMA warp, exactly as the pre-split kernel does from global memory.
// 16-channel inputs only (Q = 1), 3^3 submanifold convolutions (the strided rulebook has no reuse).
#define UR_CAP 512
#define UR_HASH 1024
__device__ __forceinline__ void bar_producers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int KG>
__global__ void __launch_bounds__(160)
conv_tc32_ur_kernel(Tc32Params p, long long n_tiles) {
  static_assert(KG * 24 <= 64, "A stage must fit its 64 TMEM columns");
  extern __shared__ __align__(1024) unsigned char sm[];
  constexpr int B_OFF = 3 * T32_BBLK;
  unsigned char* bank = sm;                                               // [K][3][512]
  unsigned char* rows = sm + 27 * B_OFF;                                  // [UR_CAP][3][32 B]
  int* hkey = reinterpret_cast<int*>(rows + UR_CAP * 96);                 // [UR_HASH] row id or -1
  int* hval = hkey + UR_HASH;                                             // [UR_HASH] local id
  int* list = hval + UR_HASH;                                             // [UR_CAP] row id of local id
  __shared__ __align__(8) unsigned long long full[2], empty[2], acc_full, acc_empty;
  __shared__ unsigned tmem_ptr_s;
  __shared__ int n_unique;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 128);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  for (int i = tid; i < 27 * (B_OFF / 16); i += 160) cp16(bank + i * 16, p.wsplit + (size_t)i * 16);
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  asm volatile("fence.proxy.async.shared::cta;" ::);
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;

  constexpr int K = 27;                       // 3^3 submanifold only: lets the tap loop unroll, indices stay in registers
  constexpr int ngroups = (K + KG - 1) / KG;
  constexpr int n_main = (K + 3) >> 2;
  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers + epilogue
    const unsigned lane_base = tmem + ((unsigned)(warp * 32) << 16);
    auto epilogue = [&](long long tl) {
      mbar_wait(&acc_full, (unsigned)(tl & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      unsigned v[16], vc[16];
      tmem_ld16(lane_base + T32_CORR, v);
      for (int a = n_main - 1; a >= 0; --a) {
        tmem_ld16(lane_base + 16u * (unsigned)a, vc);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(vc[c]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
      mbar_arrive(&acc_empty);
      const long long j = (blockIdx.x + tl * gridDim.x) * T32_M + tid;
      if (j < p.n_rows) epilogue_row16(p, v, j);
    };

    long long it = 0;                                   // running work-item counter (stage / parity bookkeeping)
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long j = (blockIdx.x + tl * gridDim.x) * T32_M + tid;
      // ---- 1. neighbour indices of this thread's row, hash slots of the present ones
      int idx[27];
#pragma unroll
      for (int k = 0; k < 27; ++k)
        idx[k] = j < p.n_rows ? __ldg(p.nbr + (long long)k * p.nbr_stride + j) : -1;
      for (int i = tid; i < UR_HASH; i += 128) hkey[i] = -1;
      if (tid == 0) n_unique = 0;
      bar_producers();
      unsigned short slot[27];
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        slot[k] = 0xffff;
        if (idx[k] >= 0) {
          unsigned h = ((unsigned)idx[k] * 2654435761u) >> 22;          // 10 bits
          slot[k] = 0xfffe;                                             // "not in the table": fetched per tap from global
          for (int tries = 0; tries < 32; ++tries) {                    // bounded: a tile may have > UR_HASH distinct rows
            const int old = atomicCAS(&hkey[h], -1, idx[k]);
            if (old == -1 || old == idx[k]) { slot[k] = (unsigned short)h; break; }
            h = (h + 1) & (UR_HASH - 1);
          }
        }
      }
      bar_producers();
      // ---- number the occupied slots (local ids) and list the distinct rows
      for (int i = tid; i < UR_HASH; i += 128) {
        const int key = hkey[i];
        if (key >= 0) {
          const int lid = atomicAdd(&n_unique, 1);
          hval[i] = lid;
          if (lid < UR_CAP) list[lid] = key;
        }
      }
      bar_producers();
      const int nu = min(n_unique, UR_CAP);
      // ---- 2. every distinct row once: gather, split, park the three planes
      for (int q = tid; q < nu; q += 128) {
        const float* src = p.in + (long long)list[q] * p.ld_in;
        float a0[8], a1[8];
        load8<false>(src, 0, p.cin, a0);
        load8<false>(src, 8, p.cin, a1);
        uint4 h0, h1, m0, m1, l0, l1;
        split2(a0[0], a0[1], h0.x, m0.x, l0.x); split2(a0[2], a0[3], h0.y, m0.y, l0.y);
        split2(a0[4], a0[5], h0.z, m0.z, l0.z); split2(a0[6], a0[7], h0.w, m0.w, l0.w);
        split2(a1[0], a1[1], h1.x, m1.x, l1.x); split2(a1[2], a1[3], h1.y, m1.y, l1.y);
        split2(a1[4], a1[5], h1.z, m1.z, l1.z); split2(a1[6], a1[7], h1.w, m1.w, l1.w);
        uint4* dst = reinterpret_cast<uint4*>(rows + q * 96);
        dst[0] = h0; dst[1] = h1; dst[2] = m0; dst[3] = m1; dst[4] = l0; dst[5] = l1;
      }
      bar_producers();
      // ---- 3. tap by tap: shared memory -> tensor memory
#pragma unroll
      for (int g = 0; g < ngroups; ++g, ++it) {
        const int s = (int)(it & 1);
        const long long u = it >> 1;
        if (u > 0) mbar_wait(&empty[s], (unsigned)((u - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::);
        const unsigned a_stage = lane_base + 128u + (unsigned)s * T32_ASTAGE_COLS;
#pragma unroll
        for (int kk = 0; kk < KG; ++kk) {
          if (g * KG + kk < K) {
            const int my_idx = idx[g * KG + kk];
            const unsigned my_slot = slot[g * KG + kk];
            unsigned r[3][8];
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
              for (int e = 0; e < 8; ++e) r[pl][e] = 0u;
            if (my_idx >= 0) {
              const int lid = my_slot < UR_HASH ? hval[my_slot] : UR_CAP;
              if (lid < UR_CAP) {
                const uint4* src = reinterpret_cast<const uint4*>(rows + lid * 96);
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                  const uint4 lo = src[2 * pl], hi = src[2 * pl + 1];
                  r[pl][0] = lo.x; r[pl][1] = lo.y; r[pl][2] = lo.z; r[pl][3] = lo.w;
                  r[pl][4] = hi.x; r[pl][5] = hi.y; r[pl][6] = hi.z; r[pl][7] = hi.w;
                }
              } else {                                   // past the cap: this row was not staged, split it here
                const float* src = p.in + (long long)my_idx * p.ld_in;
                float a0[8], a1[8];
                load8<false>(src, 0, p.cin, a0);
                load8<false>(src, 8, p.cin, a1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  split2(a0[2 * i], a0[2 * i + 1], r[0][i], r[1][i], r[2][i]);
                  split2(a1[2 * i], a1[2 * i + 1], r[0][4 + i], r[1][4 + i], r[2][4 + i]);
                }
              }
            }
            __syncwarp();
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) tmem_st8(a_stage + (unsigned)(kk * 24 + pl * 8), r[pl]);
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::);
        mbar_arrive(&full[s]);
        if (g == 0 && tl > 0) epilogue(tl - 1);
      }
      bar_producers();    // every producer is done reading `rows` / the hash before the next tile rebuilds them
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    const long long n_items = my_tiles * ngroups;
    for (long long it = 0; it < n_items; ++it) {
      const long long tl = it / ngroups;
      const int g = (int)(it % ngroups);
      const int k0 = g * KG, kg = min(KG, K - k0);
      const int s = (int)(it & 1);
      mbar_wait(&full[s], (unsigned)((it >> 1) & 1));
      if (g == 0 && tl >= 1) mbar_wait(&acc_empty, (unsigned)((tl - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      if (elect_one()) {
        const unsigned a_stage = tmem + 128u + (unsigned)s * T32_ASTAGE_COLS;
        for (int kk = 0; kk < kg; ++kk) {
          const int k = k0 + kk;
          const unsigned a = a_stage + (unsigned)(kk * 24);
          const unsigned b = smem_u32(bank + (size_t)k * 3 * T32_BBLK);
          const unsigned first_corr = k == 0 ? 0u : 1u;
          const unsigned first_main = (k & 3) == 0 ? 0u : 1u;
          mma_bf16_ts(tmem + T32_CORR, a + 16u, umma_desc(b), first_corr);
          mma_bf16_ts(tmem + T32_CORR, a + 8u, umma_desc(b + T32_BBLK), 1u);
          mma_bf16_ts(tmem + T32_CORR, a, umma_desc(b + 2 * T32_BBLK), 1u);
          mma_bf16_ts(tmem + T32_CORR, a + 8u, umma_desc(b), 1u);
          mma_bf16_ts(tmem + T32_CORR, a, umma_desc(b + T32_BBLK), 1u);
          mma_bf16_ts(tmem + 16u * (unsigned)(k >> 2), a, umma_desc(b), first_main);
        }
        mma_commit(&empty[s]);
        if (g == ngroups - 1) mma_commit(&acc_full);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---------------------------------------------------------------------------------------------------------------
// child mode: Cin = 48 (Q = 3), Cout = 16; p.n_rows = parent rows, output row 8 p + c
__global__ void __launch_bounds__(128)
conv_tc32_child_kernel(Tc32Params p, long long n_tiles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  constexpr int Q = 3;
  constexpr int A_BYTES = Q * 3 * T32_ABLK;     // 36864
  constexpr int PAIR_BYTES = Q * 3 * T32_BBLK;  // 4608
  unsigned char* As = sm;                       // [Q][3][4096]
  unsigned char* Bs = sm + A_BYTES;             // [<= 8 pairs][Q][3][512]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_ptr_s;
  __shared__ int pair_start[28];

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
    int cnt = 0;
    for (int e = 0; e < 27; ++e) {
      pair_start[e] = cnt;
      for (int c = 0; c < 8; ++c) cnt += child_uses(c, e) ? 1 : 0;
    }
    pair_start[27] = cnt;   // 64
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;
  unsigned phase = 0;

  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_items = my_tiles * 27;
  float x[2 * Q][8];
  int idx = -1, idx_next = -1;
  const int row_off = (tid >> 3) * 256 + (tid & 7) * 16;

  auto load_idx = [&](long long item) -> int {
    const long long tile = blockIdx.x + (item / 27) * gridDim.x;
    const int e = (int)(item % 27);
    const long long j = tile * T32_M + tid;
    return j < p.n_rows ? __ldg(p.nbr + (long long)e * p.nbr_stride + j) : -1;
  };
  auto load_rows = [&]() {
    if (idx >= 0) {
      const float* src = p.in + (long long)idx * p.ld_in;
#pragma unroll
      for (int u = 0; u < 2 * Q; ++u) load8<true>(src, 8 * u, p.cin, x[u]);
    }
  };

  if (n_items > 0) {
    idx = load_idx(0);
    load_rows();
    if (n_items > 1) idx_next = load_idx(1);
  }
  unsigned written = 0;   // (warp 0) children whose accumulator columns hold data of the current tile
  for (long long it = 0; it < n_items; ++it) {
    const long long tile = blockIdx.x + (it / 27) * gridDim.x;
    const int e = (int)(it % 27);
    const int ps = pair_start[e], np = pair_start[e + 1] - ps;
    {   // pre-summed filters of the children that read parent offset e: asynchronous, lands under the conversion below
      const unsigned char* wsrc = p.wsplit + (size_t)ps * PAIR_BYTES;
      for (int i = tid; i < np * (PAIR_BYTES / 16); i += 128) cp16(Bs + i * 16, wsrc + i * 16);
      asm volatile("cp.async.commit_group;\n" ::);
    }
#pragma unroll
    for (int u = 0; u < 2 * Q; ++u) {
      unsigned char* dst = As + (u >> 1) * (3 * T32_ABLK) + (u & 1) * 128 + row_off;
      if (idx >= 0) split8_store(x[u], dst, T32_ABLK);
      else zero_store(dst, T32_ABLK);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    asm volatile("fence.proxy.async.shared::cta;" ::);
    __syncthreads();
    if (warp == 0) {   // all lanes track `written`; one elected lane issues
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      if (e == 0) written = 0;
      int slot = 0;
      for (int c = 0; c < 8; ++c) {
        if (!child_uses(c, e)) continue;
        const unsigned d = tmem + 16u * c;
        const unsigned had = (written >> c) & 1u;
        if (elect_one()) {
#pragma unroll
          for (int qc = 0; qc < Q; ++qc)
            mma_split6(d, d + 128u, smem_u32(As + qc * 3 * T32_ABLK), smem_u32(Bs + slot * PAIR_BYTES + qc * 3 * T32_BBLK),
                       (qc == 0 && !had) ? 0u : 1u, (qc == 0 && !had) ? 0u : 1u);
        }
        written |= 1u << c;
        ++slot;
      }
      if (elect_one()) mma_commit(&mbar);
      __syncwarp();
    }
    if (it + 1 < n_items) {
      idx = idx_next;
      load_rows();
      if (it + 2 < n_items) idx_next = load_idx(it + 2);
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    if (e == 26) {
      const long long pj = tile * T32_M + tid;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        unsigned v[16], vc[16];
        tmem_ld16(tmem + ((unsigned)(warp * 32) << 16) + 16u * c, v);
        tmem_ld16(tmem + ((unsigned)(warp * 32) << 16) + 128u + 16u * c, vc);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(vc[q]));
        if (pj < p.n_rows) epilogue_row16(p, v, pj * 8 + c);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised child-mode kernel.  ncu on the single-role version above: 2 CTAs x 4 warps per SM (TMEM: 256
// accumulator columns per CTA), issue slots 18 % busy, long-scoreboard stalls dominant -- every item serialises row
// latency, conversion, barrier, MMA issue (43 MMAs on average, by one thread) and the commit round trip.  Here warps
// 0-3 only produce: the rows of item i+1 are in flight (second register set) while item i is split into stage i % 2
// of a two-stage ring, and warp 4 issues the MMAs and commits them to the stage's `empty` barrier.  A work item is a
// ROUND = (parent offset e, <= 4 children): the centre offset, read by all 8 children, takes two rounds, so a stage's
// filter slot holds 4 pairs (18 KB) and two stages fit twice per SM.
__device__ __forceinline__ void child_round(int r, int& e, int& half) {
  e = r <= 13 ? r : r - 1;
  half = r == 14 ? 1 : 0;
}

__global__ void __launch_bounds__(160)
conv_tc32_child_ws_kernel(Tc32Params p, long long n_tiles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  constexpr int Q = 3;
  constexpr int A_BYTES = Q * 3 * T32_ABLK;     // 36864
  constexpr int PAIR_BYTES = Q * 3 * T32_BBLK;  // 4608
  constexpr int STAGE = A_BYTES + 4 * PAIR_BYTES;
  constexpr int ROUNDS = 28;
  __shared__ __align__(8) unsigned long long full[2], empty[2], acc_full, acc_empty;
  __shared__ unsigned tmem_ptr_s;
  __shared__ int pair_start[28];

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 128);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
    int cnt = 0;
    for (int e = 0; e < 27; ++e) {
      pair_start[e] = cnt;
      for (int c = 0; c < 8; ++c) cnt += child_uses(c, e) ? 1 : 0;
    }
    pair_start[27] = cnt;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;

  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_items = my_tiles * ROUNDS;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers + epilogue
    float x0[2 * Q][8], x1[2 * Q][8];
    int idxn = -1;                                   // neighbour row of the NEXT item to load
    const int row_off = (tid >> 3) * 256 + (tid & 7) * 16;
    const unsigned lane_base = tmem + ((unsigned)(warp * 32) << 16);
    auto load_idx = [&](long long item) {
      int e, half;
      child_round((int)(item % ROUNDS), e, half);
      const long long j = (blockIdx.x + (item / ROUNDS) * gridDim.x) * T32_M + tid;
      idxn = j < p.n_rows ? __ldg(p.nbr + (long long)e * p.nbr_stride + j) : -1;
    };
    auto load_rows = [&](float (&x)[2 * Q][8]) {
      if (idxn >= 0) {
        const float* src = p.in + (long long)idxn * p.ld_in;
#pragma unroll
        for (int u = 0; u < 2 * Q; ++u) load8<true>(src, 8 * u, p.cin, x[u]);
      } else {
#pragma unroll
        for (int u = 0; u < 2 * Q; ++u)
#pragma unroll
          for (int q = 0; q < 8; ++q) x[u][q] = 0.f;
      }
    };
    auto epilogue = [&](long long tl) {
      mbar_wait(&acc_full, (unsigned)(tl & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      const long long pj = (blockIdx.x + tl * gridDim.x) * T32_M + tid;
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
      mbar_arrive(&acc_empty);
    };
    auto step = [&](long long it, float (&xc)[2 * Q][8], float (&xn)[2 * Q][8]) {
      const long long tl = it / ROUNDS;
      const int r = (int)(it % ROUNDS);
      int e, half;
      child_round(r, e, half);
      const int s = (int)(it & 1);
      const long long u = it >> 1;
      if (it + 1 < n_items) {
        load_rows(xn);
        if (it + 2 < n_items) load_idx(it + 2);
      }
      if (u > 0) mbar_wait(&empty[s], (unsigned)((u - 1) & 1));
      unsigned char* As = sm + s * STAGE;
      unsigned char* Bs = As + A_BYTES;
      const int ps = pair_start[e] + 4 * half;
      const int np = min(4, pair_start[e + 1] - ps);
      {
        const unsigned char* wsrc = p.wsplit + (size_t)ps * PAIR_BYTES;
        for (int i = tid; i < np * (PAIR_BYTES / 16); i += 128) cp16(Bs + i * 16, wsrc + i * 16);
        asm volatile("cp.async.commit_group;\n" ::);
      }
#pragma unroll
      for (int uu = 0; uu < 2 * Q; ++uu)
        split8_store(xc[uu], As + (uu >> 1) * (3 * T32_ABLK) + (uu & 1) * 128 + row_off, T32_ABLK);
      asm volatile("cp.async.wait_group 0;\n" ::);
      asm volatile("fence.proxy.async.shared::cta;" ::);
      mbar_arrive(&full[s]);
      if (r == 0 && tl > 0) epilogue(tl - 1);
    };

    if (n_items > 0) {
      load_idx(0);
      load_rows(x0);
      if (n_items > 1) load_idx(1);
    }
    for (long long it = 0; it < n_items; it += 2) {
      step(it, x0, x1);
      if (it + 1 < n_items) step(it + 1, x1, x0);
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    unsigned written = 0;
    for (long long it = 0; it < n_items; ++it) {
      const long long tl = it / ROUNDS;
      const int r = (int)(it % ROUNDS);
      int e, half;
      child_round(r, e, half);
      const int s = (int)(it & 1);
      mbar_wait(&full[s], (unsigned)((it >> 1) & 1));
      if (r == 0 && tl >= 1) mbar_wait(&acc_empty, (unsigned)((tl - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      if (r == 0) written = 0;
      {   // every lane walks the child list (uniform); one elected lane issues
        const unsigned char* As = sm + s * STAGE;
        const unsigned char* Bs = As + A_BYTES;
        int seen = 0, slot = 0;
        for (int c = 0; c < 8; ++c) {
          if (!child_uses(c, e)) continue;
          if (seen++ < 4 * half) continue;      // second round of the centre offset: children 4..7
          if (slot == 4) break;
          const unsigned d = tmem + 16u * c;
          const unsigned had = (written >> c) & 1u;
          if (elect_one()) {
#pragma unroll
            for (int qc = 0; qc < Q; ++qc)
              mma_split6(d, d + 128u, smem_u32(As + qc * 3 * T32_ABLK), smem_u32(Bs + slot * PAIR_BYTES + qc * 3 * T32_BBLK),
                         (qc == 0 && !had) ? 0u : 1u, (qc == 0 && !had) ? 0u : 1u);
          }
          written |= 1u << c;
          ++slot;
        }
        if (elect_one()) {
          mma_commit(&empty[s]);
          if (r == ROUNDS - 1) mma_commit(&acc_full);
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

bool al(const void* p, uintptr_t a) { return ((uintptr_t)p & (a - 1)) == 0; }

// CTAs one SM holds: shared memory (227 KB usable, 1 KB reserved per CTA), registers (64 K, allocated
// per warp in units of 8 per thread) and TMEM columns.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor answered 1 for
// these kernels -- it assumes the default shared-memory carve-out -- which left three quarters of each SM idle.)
int resident_ctas(const void* fn, size_t dyn_smem, int tmem_limit, int threads = 128) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) { g_sgnn_last_cuda_error = (int)cudaGetLastError(); return -1; }
  const int by_smem = (int)((227 * 1024) / (dyn_smem + fa.sharedSizeBytes + 1024));
  const int regs = (fa.numRegs + 7) & ~7;
  const int by_regs = 65536 / (regs * threads);
  int n = by_smem < by_regs ? by_smem : by_regs;
  if (n > tmem_limit) n = tmem_limit;
  return n < 1 ? 1 : n;
}

template <int Q, int KG, bool A32>
int launch_regular(const Tc32Params& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)KG * Q * 3 * (T32_ABLK + T32_BBLK);
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_kernel<Q, KG, A32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_kernel<Q, KG, A32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ctas_per_sm = resident_ctas((const void*)conv_tc32_kernel<Q, KG, A32>, smem, 512 / T32_COLS);
    if (ctas_per_sm < 0) return SGNN_E_CUDA;
  }
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = (long long)148 * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  conv_tc32_kernel<Q, KG, A32><<<(int)grid, 128, smem, st>>>(p, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

template <int Q, int KG, bool A32>
int launch_ws(const Tc32Params& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)2 * KG * Q * 3 * (T32_ABLK + T32_BBLK);
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_ws_kernel<Q, KG, A32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_ws_kernel<Q, KG, A32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ctas_per_sm = resident_ctas((const void*)conv_tc32_ws_kernel<Q, KG, A32>, smem, 512 / (2 * T32_COLS), 160);
    if (ctas_per_sm < 0) return SGNN_E_CUDA;
  }
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = (long long)148 * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  conv_tc32_ws_kernel<Q, KG, A32><<<(int)grid, 160, smem, st>>>(p, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// Shared-memory carve-out of the TMEM-operand kernels (percent of the 228 KB).  They need only the filter bank in shared
// memory (2 CTAs x 42 KB), so a smaller carve-out leaves the L1 data cache room for the distinct rows of a tile, which the
// 27 filter offsets re-read ~9 times (scratch/reuse_model.py).  A/B knob: SGNN_TC32_TM_CARVEOUT (default 100).
int tm_carveout() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SGNN_TC32_TM_CARVEOUT");
    v = e ? atoi(e) : 100;
    if (v < 40 || v > 100) v = 100;
  }
  return v;
}

template <int Q, int KG, bool A32>
int launch_tm(const Tc32Params& p, cudaStream_t st) {
  const size_t smem = (size_t)p.K * Q * 3 * T32_BBLK;
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_tm_kernel<Q, KG, A32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 27 * Q * 3 * T32_BBLK));
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_tm_kernel<Q, KG, A32>, cudaFuncAttributePreferredSharedMemoryCarveout, tm_carveout()));
    ctas_per_sm = resident_ctas((const void*)conv_tc32_tm_kernel<Q, KG, A32>, (size_t)27 * Q * 3 * T32_BBLK, 2, 160);
    if (ctas_per_sm < 0) return SGNN_E_CUDA;
  }
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = (long long)148 * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  conv_tc32_tm_kernel<Q, KG, A32><<<(int)grid, 160, smem, st>>>(p, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

template <int Q, int KG>
int launch_pm(const Tc32Params& p, const float* in, int ld_in, cudaStream_t st) {
  const size_t smem = (size_t)p.K * Q * 3 * T32_BBLK;
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_pm_kernel<Q, KG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 27 * Q * 3 * T32_BBLK));
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_pm_kernel<Q, KG>, cudaFuncAttributePreferredSharedMemoryCarveout, tm_carveout()));
    ctas_per_sm = resident_ctas((const void*)conv_tc32_pm_kernel<Q, KG>, (size_t)27 * Q * 3 * T32_BBLK, 2, 160);
    if (ctas_per_sm < 0) return SGNN_E_CUDA;
  }
  tc32_split_rows_kernel<<<sgnn_blocks(p.n_in * Q, 256), 256, 0, st>>>(in, ld_in, p.cin, p.n_in, Q, (unsigned char*)p.planes);
  SGNN_CHECK_LAUNCH();
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = (long long)148 * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  conv_tc32_pm_kernel<Q, KG><<<(int)grid, 160, smem, st>>>(p, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

int launch_ur(const Tc32Params& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)27 * 3 * T32_BBLK + (size_t)UR_CAP * 96 + (size_t)(2 * UR_HASH + UR_CAP) * 4;
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_ur_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_ur_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ctas_per_sm = resident_ctas((const void*)conv_tc32_ur_kernel<2>, smem, 2, 160);
    if (ctas_per_sm < 0) return SGNN_E_CUDA;
  }
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = (long long)148 * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  conv_tc32_ur_kernel<2><<<(int)grid, 160, smem, st>>>(p, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

size_t tc32_weight_bytes(int K, int cin, int child_mode) {
  const int Q = (cin + 15) / 16;
  const size_t b = child_mode ? (size_t)64 * 3 * 3 * T32_BBLK : (size_t)K * Q * 3 * T32_BBLK;
  return (b + 255) & ~(size_t)255;
}

}  // namespace

extern "C" size_t sgnn_conv_tc32_workspace_bytes_rows(int32_t K, int32_t cin, int32_t child_mode, int64_t n_in) {
  const int Q = (cin + 15) / 16;
  size_t b = tc32_weight_bytes(K, cin, child_mode);
  if (!child_mode && Q <= 2 && n_in > 0) b += (size_t)3 * Q * (size_t)n_in * 32;
  return b;
}

extern "C" size_t sgnn_conv_tc32_workspace_bytes(int32_t K, int32_t cin, int32_t child_mode) {
  const int Q = (cin + 15) / 16;
  return child_mode ? (size_t)64 * 3 * 3 * T32_BBLK : (size_t)K * Q * 3 * T32_BBLK;
}

extern "C" int sgnn_conv_tc32_prepare(const void* weight, int32_t K, int32_t cin, int32_t cout, int32_t child_mode, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  if (!weight || !workspace || cin <= 0 || cin > 48) return SGNN_E_INVALID;
  if ((K != 27 && K != 8) || (cout != 16 && cout != 12 && cout != 8)) return SGNN_E_UNSUPPORTED;
  if (child_mode && (K != 27 || cin != 48 || cout != 16)) return SGNN_E_UNSUPPORTED;
  if (workspace_bytes < sgnn_conv_tc32_workspace_bytes(K, cin, child_mode)) return SGNN_E_NOMEM;
  if (!al(workspace, 16)) return SGNN_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (child_mode) {
    tc32_prep_child_kernel<<<96, 512, 0, st>>>((const float*)weight, cin, (unsigned char*)workspace);
  } else {
    const int Q = (cin + 15) / 16, total = K * Q * 256;
    tc32_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float*)weight, K, cin, cout, Q, (unsigned char*)workspace);
  }
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_conv_forward_tc32(const SgnnConvArgs* a, void* workspace, size_t workspace_bytes, void* stream) {
  if (!a || a->n_out < 0 || a->cin <= 0 || !a->weight) return SGNN_E_INVALID;
  if (a->dtype != SGNN_F32 || a->cout != 16 || a->cin > 48) return SGNN_E_UNSUPPORTED;
  if (a->K != 27 && a->K != 8) return SGNN_E_UNSUPPORTED;
  if (a->child_mode && (a->K != 27 || a->cin != 48 || a->residual || (a->n_out & 7))) return SGNN_E_UNSUPPORTED;
  if (a->n_out == 0) return SGNN_OK;
  if (!a->a.out && !a->b.out) return SGNN_E_INVALID;
  if (!a->in || !a->nbr || !workspace) return SGNN_E_INVALID;
  if (workspace_bytes < sgnn_conv_tc32_workspace_bytes(a->K, a->cin, a->child_mode)) return SGNN_E_NOMEM;
  const SgnnEpilogue* eps[2] = {&a->a, &a->b};
  for (int i = 0; i < 2; ++i) {
    const SgnnEpilogue& e = *eps[i];
    if (!e.out) continue;
    if ((e.scale == nullptr) != (e.shift == nullptr)) return SGNN_E_INVALID;
    if (!al(e.out, 16) || (e.ld & 3) || (e.scale && (!al(e.scale, 16) || !al(e.shift, 16)))) return SGNN_E_ALIGN;
  }
  if (!al(a->in, 16) || (a->ld_in & 3) || a->ld_in < a->cin || !al(workspace, 16)) return SGNN_E_ALIGN;
  if (a->residual && (!al(a->residual, 16) || (a->ld_res & 3))) return SGNN_E_ALIGN;
  const bool a32 = al(a->in, 32) && (a->ld_in & 7) == 0;
  if (a->child_mode && !a32) return SGNN_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int Q = (a->cin + 15) / 16;
  Tc32Params p;
  p.in = (const float*)a->in; p.ld_in = a->ld_in; p.cin = a->cin; p.cout = 16;
  p.nbr = a->nbr; p.nbr_stride = a->nbr_stride; p.K = a->K;
  p.wsplit = (const unsigned char*)workspace;
  p.planes = nullptr; p.n_in = a->n_in;
  p.n_rows = a->child_mode ? a->n_out / 8 : a->n_out;
  p.residual = (const float*)a->residual; p.ld_res = a->ld_res;
  p.out_a = (float*)a->a.out; p.ld_a = a->a.ld; p.relu_a = a->a.relu; p.scale_a = a->a.scale; p.shift_a = a->a.shift;
  p.out_b = (float*)a->b.out; p.ld_b = a->b.ld; p.relu_b = a->b.relu; p.scale_b = a->b.scale; p.shift_b = a->b.shift;
  const bool prepared = (a->flags & SGNN_CONV_PREPARED) != 0;
  if (a->child_mode) {
    if (!prepared) {
      tc32_prep_child_kernel<<<96, 512, 0, st>>>((const float*)a->weight, a->cin, (unsigned char*)workspace);
      SGNN_CHECK_LAUNCH();
    }
    if (g_sgnn_conv_impl != 25) {   // default: warp-specialised kernel; 25 = single-role kernel (A/B)
      constexpr size_t smem_ws = (size_t)2 * (3 * 3 * T32_ABLK + 4 * 3 * 3 * T32_BBLK);
      static int ctas_ws = 0;
      if (!ctas_ws) {
        SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_child_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
        SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_child_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctas_ws = resident_ctas((const void*)conv_tc32_child_ws_kernel, smem_ws, 2, 160);
        if (ctas_ws < 0) return SGNN_E_CUDA;
      }
      const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
      long long grid = (long long)148 * ctas_ws;
      if (grid > tiles) grid = tiles;
      conv_tc32_child_ws_kernel<<<(int)grid, 160, smem_ws, st>>>(p, tiles);
      SGNN_CHECK_LAUNCH();
      return SGNN_OK;
    }
    constexpr size_t smem = (size_t)3 * 3 * T32_ABLK + (size_t)8 * 3 * 3 * T32_BBLK;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
      SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_child_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_child_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      ctas_per_sm = resident_ctas((const void*)conv_tc32_child_kernel, smem, 2);   // 256 TMEM columns each
      if (ctas_per_sm < 0) return SGNN_E_CUDA;
    }
    const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
    long long grid = (long long)148 * ctas_per_sm;
    if (grid > tiles) grid = tiles;
    conv_tc32_child_kernel<<<(int)grid, 128, smem, st>>>(p, tiles);
    SGNN_CHECK_LAUNCH();
    return SGNN_OK;
  }
  if (!prepared) {
    const int total = a->K * Q * 256;
    tc32_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float*)a->weight, a->K, a->cin, 16, Q, (unsigned char*)workspace);
    SGNN_CHECK_LAUNCH();
  }
  if (g_sgnn_conv_impl == 28 && Q == 1 && a->K == 27) return launch_ur(p, st);   // EXPERIMENTAL: distinct rows staged once per tile
  if (g_sgnn_conv_impl == 27 && Q <= 2 && a->n_in > 0 &&
      workspace_bytes >= sgnn_conv_tc32_workspace_bytes_rows(a->K, a->cin, 0, a->n_in)) {   // A/B: pre-split planes -> TMEM (v4)
    p.planes = (const unsigned char*)workspace + tc32_weight_bytes(a->K, a->cin, 0);
    if (Q == 1) return launch_pm<1, 2>(p, (const float*)a->in, a->ld_in, st);
    return launch_pm<2, 1>(p, (const float*)a->in, a->ld_in, st);
  }
  if (g_sgnn_conv_impl == 24 && Q <= 2) {    // A/B: A operand through tensor memory (v3)
    if (Q == 1) return a32 ? launch_tm<1, 2, true>(p, st) : launch_tm<1, 2, false>(p, st);
    return a32 ? launch_tm<2, 1, true>(p, st) : launch_tm<2, 1, false>(p, st);
  }
  if (g_sgnn_conv_impl == 23) {              // A/B: warp-specialised kernel (producers / MMA issuer, double-buffered TMEM)
    if (Q == 1) return a32 ? launch_ws<1, 3, true>(p, st) : launch_ws<1, 3, false>(p, st);
    if (Q == 2) return a32 ? launch_ws<2, 1, true>(p, st) : launch_ws<2, 1, false>(p, st);
    return a32 ? launch_ws<3, 1, true>(p, st) : launch_ws<3, 1, false>(p, st);
  }
  const bool kg1 = g_sgnn_conv_impl == 21;   // A/B: one filter offset per work item for 16-channel inputs
  if (Q == 1) {
    if (kg1) return a32 ? launch_regular<1, 1, true>(p, st) : launch_regular<1, 1, false>(p, st);
    return a32 ? launch_regular<1, 3, true>(p, st) : launch_regular<1, 3, false>(p, st);
  }
  if (Q == 2) {
    if (g_sgnn_conv_impl == 26) return a32 ? launch_regular<2, 1, true>(p, st) : launch_regular<2, 1, false>(p, st);   // A/B
    return a32 ? launch_regular<2, 2, true>(p, st) : launch_regular<2, 2, false>(p, st);
  }
  return a32 ? launch_regular<3, 1, true>(p, st) : launch_regular<3, 1, false>(p, st);
}
