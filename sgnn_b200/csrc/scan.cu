// scan.cu -- exclusive prefix sums (int32) used by the grid build (popcount ranks) and by the
// stable mask compactions (model.py:238,242,325,335).  Hierarchical: 2048 items per CTA, CTA
// totals scanned recursively, offsets added back.  out has n+1 entries; out[n] = total.
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int scan_load(const void* in, int mode, long long i) {
  if (mode == SCAN_I32) return ((const int*)in)[i];
  if (mode == SCAN_POPC64) return __popcll(((const unsigned long long*)in)[i]);
  return (int)((const unsigned char*)in)[i];
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_kernel(const void* __restrict__ in, int mode, int* __restrict__ out, long long n,
                  int* __restrict__ sums) {
  __shared__ int warp_tot[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    long long idx = base + i;
    v[i] = idx < n ? scan_load(in, mode, idx) : 0;
    tsum += v[i];
  }
  // inclusive warp scan of thread sums
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = w;  // inclusive
  }
  __syncthreads();
  int excl = inc - tsum + (wid ? warp_tot[wid - 1] : 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    long long idx = base + i;
    if (idx < n) out[idx] = excl;
    excl += v[i];
  }
  if (threadIdx.x == SCAN_THREADS - 1) {
    sums[blockIdx.x] = excl;
    if (gridDim.x == 1) out[n] = excl;
  }
}

// Small inputs (the per-level grids and candidate lists of a 32-block batch): ONE CTA instead of the 3-launch
// hierarchy -- these scans sit on the critical path between two host reads.  1024 threads, IT contiguous items per
// thread, one sweep: thread sums -> warp scans -> scan of the 32 warp totals -> write (2 barriers; the earlier
// 256-thread version walked up to 16 tiles serially with 4 barriers each, 11 us per call, 20 calls per pass).
#define SCAN_SINGLE_THREADS 1024
#define SCAN_SINGLE_MAX (SCAN_SINGLE_THREADS * 32)
template <int IT>
__global__ void __launch_bounds__(SCAN_SINGLE_THREADS)
scan_single_kernel(const void* __restrict__ in, int mode, int* __restrict__ out, long long n) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long base = (long long)threadIdx.x * IT;
  int v[IT];
  int tsum = 0;
#pragma unroll
  for (int i = 0; i < IT; ++i) {
    const long long idx = base + i;
    v[i] = idx < n ? scan_load(in, mode, idx) : 0;
    tsum += v[i];
  }
  int inc = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    warp_tot[lane] = w;  // inclusive
  }
  __syncthreads();
  int excl = inc - tsum + (wid ? warp_tot[wid - 1] : 0);
#pragma unroll
  for (int i = 0; i < IT; ++i) {
    const long long idx = base + i;
    if (idx < n) out[idx] = excl;
    excl += v[i];
  }
  if (threadIdx.x == SCAN_SINGLE_THREADS - 1) out[n] = excl;   // items past n are zeros: the last thread holds the total
}

// Single-pass chained scan (decoupled look-back): one launch for any n.  Tiles take their index from a global counter (so a
// tile never waits on one that has not started), publish (state, value) in ONE 64-bit word -- state 1 = the tile's own sum,
// 2 = inclusive prefix through the tile -- and warp 0 walks the predecessors' words 32 at a time until it meets an inclusive
// prefix.  Replaces the three-launch hierarchy (block scans, scan of the block sums, add-back: 14 us on the critical path,
// ~25 scans per generator pass).  ctrl[0] = tile counter, ctrl[1 + t] = status of tile t; zeroed before the launch.
#define SCAN_STATE(w) ((unsigned)((w) >> 32))
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(const void* __restrict__ in, int mode, int* __restrict__ out, long long n,
                     unsigned long long* __restrict__ ctrl, int n_tiles) {
  __shared__ int warp_tot[SCAN_THREADS / 32];
  __shared__ int s_tile, s_excl;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ctrl, 1ull);
  __syncthreads();
  const int tile = s_tile;
  volatile unsigned long long* status = ctrl + 1;
  const long long base = (long long)tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    long long idx = base + i;
    v[i] = idx < n ? scan_load(in, mode, idx) : 0;
    tsum += v[i];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = w;  // inclusive
    const int aggregate = __shfl_sync(0xffffffffu, w, SCAN_THREADS / 32 - 1);
    if (lane == 0 && tile > 0) status[tile] = (1ull << 32) | (unsigned)aggregate;
    int excl = 0;
    int idx = tile - 1;
    while (idx >= 0) {
      const int j = idx - lane;
      unsigned long long sw = (2ull << 32);            // before tile 0: inclusive prefix 0
      if (j >= 0) {
        do { sw = status[j]; } while (SCAN_STATE(sw) == 0);
      }
      const unsigned incl = __ballot_sync(0xffffffffu, SCAN_STATE(sw) == 2);
      const int first = incl ? __ffs(incl) - 1 : 32;
      int val = lane <= first ? (int)(unsigned)sw : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
      excl += val;
      if (incl) break;
      idx -= 32;
    }
    if (lane == 0) {
      status[tile] = (2ull << 32) | (unsigned)(excl + aggregate);
      s_excl = excl;
      if (tile == n_tiles - 1) out[n] = excl + aggregate;
    }
  }
  __syncthreads();
  int excl = s_excl + inc - tsum + (wid ? warp_tot[wid - 1] : 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    long long idx = base + i;
    if (idx < n) out[idx] = excl;
    excl += v[i];
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int* __restrict__ out, long long n, const int* __restrict__ sums_ex, int nb) {
  const int off = sums_ex[blockIdx.x];
  const long long base = (long long)blockIdx.x * SCAN_TILE;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS) {
    long long idx = base + i;
    if (idx < n) out[idx] += off;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums_ex[nb];
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" size_t sgnn_scan_scratch_bytes(int64_t n_items) {
  size_t total = 256;
  int64_t n = n_items < 1 ? 1 : n_items;
  while (true) {
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    total += align_up((size_t)nb * 4, 256) + align_up((size_t)(nb + 1) * 4, 256);
    if (nb <= 1) break;
    n = nb;
  }
  return total;
}

int sgnn_scan_exclusive(const void* in, int mode, int* out, int64_t n, void* scratch,
                        size_t scratch_bytes, cudaStream_t st) {
  if (n < 0 || !out || (!in && n > 0) || !scratch) return SGNN_E_INVALID;
  if (scratch_bytes < sgnn_scan_scratch_bytes(n)) return SGNN_E_INVALID;
  // one 1024-thread CTA up to 8 items per thread (9 us); beyond that the single sweep is latency bound (16 k-32 k items: 37 us
  // measured) and the three-kernel hierarchy (14 us) wins
  if (n <= SCAN_SINGLE_THREADS * 8) {
    const int per = (int)((n + SCAN_SINGLE_THREADS - 1) / SCAN_SINGLE_THREADS);
    if (per <= 4) scan_single_kernel<4><<<1, SCAN_SINGLE_THREADS, 0, st>>>(in, mode, out, (long long)n);
    else if (per <= 8) scan_single_kernel<8><<<1, SCAN_SINGLE_THREADS, 0, st>>>(in, mode, out, (long long)n);
    else if (per <= 16) scan_single_kernel<16><<<1, SCAN_SINGLE_THREADS, 0, st>>>(in, mode, out, (long long)n);
    else scan_single_kernel<32><<<1, SCAN_SINGLE_THREADS, 0, st>>>(in, mode, out, (long long)n);
    SGNN_CHECK_LAUNCH();
    return SGNN_OK;
  }
  int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb < 1) nb = 1;
  if (nb > 0x7fffffff) return SGNN_E_TOO_LARGE;
  if (((uintptr_t)scratch & 7) == 0 && (size_t)(nb + 1) * 8 <= scratch_bytes) {
    if (cudaMemsetAsync(scratch, 0, (size_t)(nb + 1) * 8, st) != cudaSuccess) return SGNN_E_CUDA;
    scan_lookback_kernel<<<(int)nb, SCAN_THREADS, 0, st>>>(in, mode, out, (long long)n, (unsigned long long*)scratch, (int)nb);
    SGNN_CHECK_LAUNCH();
    return SGNN_OK;
  }
  int* sums = (int*)scratch;
  scan_block_kernel<<<(int)nb, SCAN_THREADS, 0, st>>>(in, mode, out, (long long)n, sums);
  SGNN_CHECK_LAUNCH();
  if (nb > 1) {
    size_t used = align_up((size_t)nb * 4, 256);
    int* sums_ex = (int*)((char*)scratch + used);
    used += align_up((size_t)(nb + 1) * 4, 256);
    int rc = sgnn_scan_exclusive(sums, SCAN_I32, sums_ex, nb, (char*)scratch + used,
                                 scratch_bytes - used, st);
    if (rc) return rc;
    scan_add_kernel<<<(int)nb, SCAN_THREADS, 0, st>>>(out, (long long)n, sums_ex, (int)nb);
    SGNN_CHECK_LAUNCH();
  }
  return SGNN_OK;
}
