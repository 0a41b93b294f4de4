// ur_common.cuh -- shared by the unique-row tensor-core convolutions (conv_ur.cu: regular, conv_urc.cu: child mode):
// tile-plan view, mbarrier / TMA / shared-memory helpers on 32-bit shared-window addresses, the barrier watchdog.
#pragma once
#include "tc32_common.cuh"

#define UR_PLAN_CAP 512          // distinct rows per tile the plan can list
#define UR_BM_WORDS 4096         // bitmap words of the plan kernel: row-id span <= 131072 per tile
#define UR_LIDX_BYTES (27 * 128 * 2)
#define UR_CHUNK 64              // rows per ring chunk

// watchdog record (mapped pinned host memory, 128 words), defined in conv_ur.cu; each translation unit points its own
// device-side copy of the pointer at it (ur_diag_init)
extern unsigned long long* g_ur_diag_host;

namespace {

struct PlanView {
  const int* ucount;             // [tiles]  distinct rows, or -1: direct mode
  const int* urows;              // [tiles][UR_PLAN_CAP]
  const unsigned short* lidx;    // [tiles][27][128]
};

size_t plan_align(size_t b) { return (b + 255) & ~(size_t)255; }

// ---------------------------------------------------------------------------------------------------------------
// Barrier / TMA helpers on 32-bit shared-window addresses (computed once: going through generic pointers makes the
// compiler rebuild the window address from SR_CgaCtaId around every use).
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_init(unsigned a, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mb_arrive(unsigned a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
// Watchdog: a wait that does not complete within 3 s (%globaltimer) records where it stood in the
// mapped host buffer `g_ur_diag` and traps -- a protocol error must fail loudly, not hang the device.
static __device__ unsigned long long* g_ur_diag_dev = nullptr;
__device__ __noinline__ void mb_timeout(unsigned a, unsigned parity, unsigned site, unsigned bar0, bool fatal) {
  unsigned long long* d = g_ur_diag_dev;
  if (d) {
    const unsigned long long me = (unsigned long long)blockIdx.x + 1;
    const unsigned long long owner = atomicCAS(d, 0ull, me);
    if (owner == 0ull || owner == me) {          // one block records: 4 words per warp
      unsigned long long* w = d + 4 + 4 * (threadIdx.x >> 5);
      unsigned long long state;
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(state) : "r"(a));
      w[0] = ((unsigned long long)site << 32) | parity;
      w[1] = ((unsigned long long)threadIdx.x << 32) | (a - bar0);
      w[2] = state;
      __threadfence_system();
    }
  }
  if (fatal) { __threadfence_system(); __trap(); }
}
__device__ __forceinline__ bool mb_try(unsigned a, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}
// slow path of a wait, out of line so that the unrolled role loops stay small
// one out-of-line copy per warp role (ROLE_ID of the calling scope), so that profiler samples of the waits separate by role
template <int ROLE>
__device__ __noinline__ void mb_wait_slow(unsigned a, unsigned parity, unsigned site, unsigned bar0) {
  unsigned long long t0 = 0;
  bool noted = false;
  while (!mb_try(a, parity)) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    else if (now - t0 > 3000000000ull) mb_timeout(a, parity, site, bar0, true);
    else if (now - t0 > 1500000000ull && !noted) { noted = true; mb_timeout(a, parity, site, bar0, false); }
  }
}
template <int ROLE>
__device__ __forceinline__ void mb_wait_site(unsigned a, unsigned parity, unsigned site, unsigned bar0) {
  if (!mb_try(a, parity)) mb_wait_slow<ROLE>(a, parity, site, bar0);
}
#define mb_wait(a, parity) mb_wait_site<ROLE_ID>((a), (parity), (unsigned)__LINE__, bar_a)
__device__ __forceinline__ void mma_commit_a(unsigned a) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a) : "memory");
}
// TMA, non-tensor form: `bytes` (multiple of 16) global -> shared, completion on the mbarrier's transaction count
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// predicated form: lanes with pred == 0 issue no shared-memory request (fewer bank conflicts than reading a common zero row)
__device__ __forceinline__ uint4 lds128_if(unsigned a, unsigned pred) {
  uint4 v;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "mov.b32 %0, 0;\n\t"
      "mov.b32 %1, 0;\n\t"
      "mov.b32 %2, 0;\n\t"
      "mov.b32 %3, 0;\n\t"
      "@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t"
      "}\n"
      : "=&r"(v.x), "=&r"(v.y), "=&r"(v.z), "=&r"(v.w)
      : "r"(a), "r"(pred));
  return v;
}
__device__ __forceinline__ unsigned lds16(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void mma_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, bool acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(T32_IDESC), "r"(acc ? 1u : 0u));
}

bool al(const void* p, uintptr_t a) { return ((uintptr_t)p & (a - 1)) == 0; }

PlanView plan_view(const void* plan, long long tiles) {
  PlanView v;
  const unsigned char* b = (const unsigned char*)plan;
  v.ucount = (const int*)b;
  b += plan_align((size_t)tiles * 4);
  v.urows = (const int*)b;
  b += plan_align((size_t)tiles * UR_PLAN_CAP * 4);
  v.lidx = (const unsigned short*)b;
  return v;
}

// allocates the watchdog record on first use and points this translation unit's device pointer at it
int ur_diag_init() {
  static bool done_dev[64] = {};
  int dev = 0;
  SGNN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SGNN_E_INVALID;
  bool& done = done_dev[dev];
  if (done) return SGNN_OK;
  if (!g_ur_diag_host) {
    SGNN_CUDA(cudaHostAlloc((void**)&g_ur_diag_host, 128 * 8, cudaHostAllocMapped));
    for (int i = 0; i < 128; ++i) g_ur_diag_host[i] = 0;
  }
  unsigned long long* dptr = nullptr;
  SGNN_CUDA(cudaHostGetDevicePointer((void**)&dptr, g_ur_diag_host, 0));
  SGNN_CUDA(cudaMemcpyToSymbol(g_ur_diag_dev, &dptr, sizeof(dptr)));
  done = true;
  return SGNN_OK;
}

}  // namespace
