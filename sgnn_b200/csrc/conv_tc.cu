// conv_tc.cu -- bf16 sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), SURVEY §8 rows a3/a4 for
// BASELINE.json configs[2] (bf16 features, C = 16 stride-2 Convolution + Deconvolution).  The fp32 generator path
// stays on the bit-reproducible FFMA kernels (conv.cu); this is the north_star's "tcgen05 tiles only for the
// per-offset dense (Nactive x Cin).(Cin x Cout) contraction".
//
// Work item = (tile of 128 output rows (UMMA M = 128), group of <= 9 filter offsets); N = Cout = 16, and per offset k
// one tcgen05.mma.cta_group::1.kind::f16 with K = Cin = 16 (bf16 UMMA_K).  CTAs are persistent (a few per SM) and
// double buffer the A tiles, so the gather of the next item overlaps the MMAs + epilogue of the current one.  Per item:
//   1. all 128 threads gather the neighbour rows (32 B each) with cp.async straight into the canonical
//      K-major / no-swizzle core-matrix layout the UMMA shared-memory descriptor expects
//        row r, 16-byte chunk c  ->  (r/8)*256 + c*128 + (r%8)*16      (SBO = 256 B, LBO = 128 B)
//      (absent neighbours -> zero rows), and stage W[k] transposed to [Cout][Cin] in the same layout;
//   2. fence.proxy.async + barrier; ONE thread issues the MMAs (D in TMEM, accumulate over k) and commits to an mbarrier;
//   3. everybody waits on the mbarrier (MMAs done => shared memory reusable).
// Epilogue: tcgen05.ld 32x32b.x16 (thread t of warp w owns accumulator row 32w+t), optional affine + ReLU, bf16 store.
// Deconvolution (filter 2, stride 2) is the same kernel: offset k contributes in[parent>>3] for the rows whose
// parent&7 == k and a zero row otherwise.
#include <cuda_bf16.h>
#include "common.cuh"

#define TC_M 128
#define TC_KG 9          // filter offsets per shared-memory group
#define TC_A_BYTES 4096  // 128 rows x 16 bf16
#define TC_B_BYTES 512   // 16 x 16 bf16

struct TcParams {
  const __nv_bfloat16* in; int ld_in;
  const int* tbl;            // mode 0: nbr [K][stride];  mode 1: parent [n_out]
  long long tbl_stride;
  int K, mode;
  const __nv_bfloat16* w;    // [K][16][16]  (Cin, Cout)
  long long n_out;
  __nv_bfloat16* out; int ld_out;
  const float* scale; const float* shift; int relu;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout_type = SWIZZLE_NONE [61,64)
  return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
}

// InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
#define TC_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void cp16(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(g));
}

// Persistent, software-pipelined version: a CTA walks work items (tile, offset group) with stride gridDim.x.  The
// whole filter bank is staged once per CTA; A tiles are double buffered, so the cp.async gather of item j+1 is in
// flight while the MMAs of item j complete and its epilogue (TMEM -> registers -> bf16 rows) runs.
__global__ void __launch_bounds__(128)
conv_tc_bf16_kernel(TcParams p, int kg_max, long long n_tiles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int ngroups = (p.K + kg_max - 1) / kg_max;
  unsigned char* B = sm;                                   // [K][512]
  unsigned char* A0 = sm + ((p.K * TC_B_BYTES + 1023) & ~1023);
  const int a_buf = kg_max * TC_A_BYTES;                   // A0 + buf * a_buf : [kg_max][4096]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_ptr_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_s)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  // filter bank once per CTA: W[k][ci][co] -> element (n = co, kdim = ci) of the canonical K-major layout
  for (int idx = tid; idx < p.K * 256; idx += 128) {
    const int kk = idx >> 8, e = idx & 255, ci = e >> 4, co = e & 15;
    *reinterpret_cast<__nv_bfloat16*>(B + kk * TC_B_BYTES + (co >> 3) * 256 + (ci >> 3) * 128 + (co & 7) * 16 + (ci & 7) * 2) =
        p.w[(size_t)kk * 256 + e];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const unsigned tmem = tmem_ptr_s;
  unsigned phase = 0;

  // gather of work item (tile, group g) into A buffer `buf`.  Thread t copies chunk c = t&1 of rows {t>>1, 64 + (t>>1)} for
  // every offset of the group: ALL neighbour indices are loaded first (independent loads in flight together), then the
  // dependent cp.async copies are issued -- a naive loop exposes one index-load latency per copied chunk.
  auto gather = [&](long long tile, int g, int buf) {
    const long long tile_base = tile * TC_M;
    const int k0 = g * kg_max, kg = min(kg_max, p.K - k0);
    unsigned char* A = A0 + buf * a_buf;
    const int c = tid & 1;
    int src[TC_KG][2];
    if (p.mode == 0) {
#pragma unroll
      for (int kk = 0; kk < TC_KG; ++kk)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long j = tile_base + (tid >> 1) + 64 * h;
          src[kk][h] = (kk < kg && j < p.n_out) ? __ldg(p.tbl + (long long)(k0 + kk) * p.tbl_stride + j) : -1;
        }
    } else {
      int pk[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long j = tile_base + (tid >> 1) + 64 * h;
        pk[h] = j < p.n_out ? __ldg(p.tbl + j) : -1;
      }
#pragma unroll
      for (int kk = 0; kk < TC_KG; ++kk)
#pragma unroll
        for (int h = 0; h < 2; ++h) src[kk][h] = (kk < kg && pk[h] >= 0 && (pk[h] & 7) == k0 + kk) ? (pk[h] >> 3) : -1;
    }
#pragma unroll
    for (int kk = 0; kk < TC_KG; ++kk)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (kk < kg) {
          const int r = (tid >> 1) + 64 * h;
          unsigned char* dst = A + kk * TC_A_BYTES + (r >> 3) * 256 + c * 128 + (r & 7) * 16;
          if (src[kk][h] >= 0) cp16(dst, p.in + (long long)src[kk][h] * p.ld_in + c * 8);
          else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
  };

  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_items = my_tiles * ngroups;
  if (n_items > 0) gather(blockIdx.x, 0, 0);
  asm volatile("cp.async.commit_group;\n" ::);
  for (long long it = 0; it < n_items; ++it) {
    const long long tile = blockIdx.x + (it / ngroups) * gridDim.x;
    const int g = (int)(it % ngroups);
    const int buf = (int)(it & 1);
    asm volatile("cp.async.wait_group 0;\n" ::);
    asm volatile("fence.proxy.async.shared::cta;" ::);   // generic-proxy writes -> visible to the tensor-core (async) proxy
    __syncthreads();                                      // item `it` staged; previous epilogue's TMEM loads retired
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
      const int k0 = g * kg_max, kg = min(kg_max, p.K - k0);
      for (int kk = 0; kk < kg; ++kk) {
        const unsigned long long da = umma_desc(smem_u32(A0 + buf * a_buf + kk * TC_A_BYTES));
        const unsigned long long db = umma_desc(smem_u32(B + (k0 + kk) * TC_B_BYTES));
        const unsigned accumulate = (k0 + kk) > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                   : "memory");
    }
    // prefetch the next item into the other buffer (its previous reader, item it-1, completed: we waited on its commit)
    if (it + 1 < n_items) {
      const long long nt = blockIdx.x + ((it + 1) / ngroups) * gridDim.x;
      gather(nt, (int)((it + 1) % ngroups), buf ^ 1);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    // wait until the MMAs of this item have completed
    {
      const unsigned addr = smem_u32(&mbar);
      asm volatile(
          "{\n\t"
          ".reg .pred P1;\n\t"
          "TC_WAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
          "@P1 bra TC_DONE;\n\t"
          "bra TC_WAIT;\n\t"
          "TC_DONE:\n\t"
          "}\n" ::"r"(addr), "r"(phase), "r"(0x989680)
          : "memory");
    }
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    if (g == ngroups - 1) {
      // ---- epilogue of the tile: TMEM -> registers -> bf16 rows
      unsigned v[16];
      const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      const long long j = tile * TC_M + warp * 32 + lane;
      if (j < p.n_out) {
        __nv_bfloat16 o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float y = __uint_as_float(v[c]);
          if (p.scale) y = fmaf(y, __ldg(p.scale + c), __ldg(p.shift + c));
          if (p.relu) y = fmaxf(y, 0.f);
          o[c] = __float2bfloat16_rn(y);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + j * p.ld_out);
        dst[0] = *reinterpret_cast<uint4*>(&o[0]);
        dst[1] = *reinterpret_cast<uint4*>(&o[8]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::);
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// shared launcher: mode 0 = neighbour table [K][stride], mode 1 = parent table (deconvolution)
int sgnn_conv_tc_bf16(const void* in, int ld_in, const int* tbl, long long tbl_stride, int K, int mode,
                      const void* w, int cin, int cout, long long n_out, const SgnnEpilogue* ep, cudaStream_t st) {
  if (cin != 16 || cout != 16) return SGNN_E_UNSUPPORTED;   // SG-NN's bf16 configuration (BASELINE configs[2])
  if (K < 1 || K > 27) return SGNN_E_UNSUPPORTED;
  if (!ep || !ep->out || (ep->scale == nullptr) != (ep->shift == nullptr)) return SGNN_E_INVALID;
  if (n_out == 0) return SGNN_OK;
  if (!in || !tbl || !w) return SGNN_E_INVALID;
  if (!al16(in) || (ld_in & 7) || !al16(ep->out) || (ep->ld & 7)) return SGNN_E_ALIGN;
  TcParams p;
  p.in = (const __nv_bfloat16*)in; p.ld_in = ld_in; p.tbl = tbl; p.tbl_stride = tbl_stride; p.K = K; p.mode = mode;
  p.w = (const __nv_bfloat16*)w; p.n_out = n_out; p.out = (__nv_bfloat16*)ep->out; p.ld_out = ep->ld;
  p.scale = ep->scale; p.shift = ep->shift; p.relu = ep->relu;
  const int kg_max = K < TC_KG ? K : TC_KG;
  const size_t smem = (((size_t)K * TC_B_BYTES + 1023) & ~(size_t)1023) + 2 * (size_t)kg_max * TC_A_BYTES + 1024;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    SGNN_CUDA(cudaFuncSetAttribute(conv_tc_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const long long tiles = (n_out + TC_M - 1) / TC_M;
  const int per_sm = (int)((220 * 1024) / (smem + 2048));
  long long grid = (long long)148 * (per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm));
  if (grid > tiles) grid = tiles;
  conv_tc_bf16_kernel<<<(int)grid, 128, smem, st>>>(p, kg_max, tiles);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}
