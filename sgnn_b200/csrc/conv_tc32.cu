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
  int dev = 0;
  SGNN_CUDA(cudaGetDevice(&dev));
  static int ctas_dev[64] = {};       // function attributes and occupancy are per device
  if (dev < 0 || dev >= 64) return SGNN_E_INVALID;
  if (!ctas_dev[dev]) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_kernel<Q, KG, A32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_kernel<Q, KG, A32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ctas_dev[dev] = resident_ctas((const void*)conv_tc32_kernel<Q, KG, A32>, smem, 512 / T32_COLS);
    if (ctas_dev[dev] < 0) { ctas_dev[dev] = 0; return SGNN_E_CUDA; }
  }
  const int ctas_per_sm = ctas_dev[dev];
  const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
  long long grid = (long long)148 * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  conv_tc32_kernel<Q, KG, A32><<<(int)grid, 128, smem, st>>>(p, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

size_t tc32_weight_bytes(int K, int cin, int child_mode) {
  const int Q = (cin + 15) / 16;
  const size_t b = child_mode ? (size_t)64 * 3 * 3 * T32_BBLK : (size_t)K * Q * 3 * T32_BBLK;
  return (b + 255) & ~(size_t)255;
}

}  // namespace

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
    constexpr size_t smem_ws = (size_t)2 * (3 * 3 * T32_ABLK + 4 * 3 * 3 * T32_BBLK);
    int dev = 0;
    SGNN_CUDA(cudaGetDevice(&dev));
    static int ctas_ws[64] = {};
    if (dev < 0 || dev >= 64) return SGNN_E_INVALID;
    if (!ctas_ws[dev]) {
      SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_child_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
      SGNN_CUDA(cudaFuncSetAttribute(conv_tc32_child_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      ctas_ws[dev] = resident_ctas((const void*)conv_tc32_child_ws_kernel, smem_ws, 2, 160);
      if (ctas_ws[dev] < 0) { ctas_ws[dev] = 0; return SGNN_E_CUDA; }
    }
    const long long tiles = (p.n_rows + T32_M - 1) / T32_M;
    long long grid = (long long)148 * ctas_ws[dev];
    if (grid > tiles) grid = tiles;
    conv_tc32_child_ws_kernel<<<(int)grid, 160, smem_ws, st>>>(p, tiles);
    SGNN_CHECK_LAUNCH();
    return SGNN_OK;
  }
  if (!prepared) {
    const int total = a->K * Q * 256;
    tc32_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float*)a->weight, a->K, a->cin, 16, Q, (unsigned char*)workspace);
    SGNN_CHECK_LAUNCH();
  }
  if (Q == 1) return a32 ? launch_regular<1, 3, true>(p, st) : launch_regular<1, 3, false>(p, st);
  if (Q == 2) return a32 ? launch_regular<2, 2, true>(p, st) : launch_regular<2, 2, false>(p, st);
  return a32 ? launch_regular<3, 1, true>(p, st) : launch_regular<3, 1, false>(p, st);
}
