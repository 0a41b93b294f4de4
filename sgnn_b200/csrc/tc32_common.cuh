// tc32_common.cuh -- device helpers shared by the tcgen05 fp32 convolution kernels (conv_tc32.cu, conv_ur.cu):
// descriptors, MMA issue, mbarrier / TMEM access, the exact 3-way bf16 split, epilogue, filter-bank preparation.
#pragma once
#include <stdlib.h>
#include "common.cuh"

#define T32_M 128
#define T32_ABLK 4096
#define T32_BBLK 512
#define T32_COLS 128   // TMEM columns of the regular kernel: 7 main accumulators (16 each) + the correction accumulator
#define T32_CORR 112u


struct Tc32Params {
  const float* in; int ld_in; int cin; int cout;   // cout: 16, or 8 / 12 (unique-row kernel: accumulator columns >= cout are zero)
  const int* nbr; long long nbr_stride; int K;
  const unsigned char* wsplit;
  const unsigned char* planes;   // pre-split input rows [3][Q][n_in][16] bf16 (conv_tc32_pm_kernel), else NULL
  long long n_in;
  long long n_rows;     // output rows (regular) / parent rows (child mode)
  const float* residual; int ld_res;
  float* out_a; int ld_a; int relu_a; const float* scale_a; const float* shift_a;
  float* out_b; int ld_b; int relu_b; const float* scale_b; const float* shift_b;
};

namespace {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30) = 128 B, SBO>>4 [32,46) = 256 B, version 1 [46,48), no swizzle
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
  return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
}
// InstrDescriptor: D fp32 (1<<4), A/B bf16 (1<<7, 1<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
#define T32_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void cp16(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(g));
}

__device__ __forceinline__ void mma_bf16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(T32_IDESC), "r"(acc));
}

// One lane of a CONVERGED warp.  Issuing tcgen05.mma under `if (lane == 0)` makes ptxas wrap every UTCHMMA in an
// ELECT / BRA.U.ANY serialisation loop (12 instructions + a branch per MMA); under elect.sync it emits straight-line
// uniform-datapath code.
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mma_commit(unsigned long long* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, unsigned phase) {
  const unsigned addr = smem_u32(mbar);
  unsigned ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(phase), "r"(0x989680)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// exact 3-way bf16 split of two fp32 values; element a goes to the low half (lower address), b to the high half
__device__ __forceinline__ void split2(float a, float b, unsigned& p0, unsigned& p1, unsigned& p2) {
  const unsigned ua = __float_as_uint(a), ub = __float_as_uint(b);
  const float ra = a - __uint_as_float(ua & 0xffff0000u), rb = b - __uint_as_float(ub & 0xffff0000u);
  const unsigned va = __float_as_uint(ra), vb = __float_as_uint(rb);
  const float sa = ra - __uint_as_float(va & 0xffff0000u), sb = rb - __uint_as_float(vb & 0xffff0000u);
  p0 = __byte_perm(ua, ub, 0x7632);
  p1 = __byte_perm(va, vb, 0x7632);
  p2 = __byte_perm(__float_as_uint(sa), __float_as_uint(sb), 0x7632);
}

// 8 consecutive channels of one row -> one 16-byte chunk in each of the three planes
__device__ __forceinline__ void split8_store(const float (&x)[8], unsigned char* dst0, int plane_stride) {
  uint4 h, m, l;
  split2(x[0], x[1], h.x, m.x, l.x);
  split2(x[2], x[3], h.y, m.y, l.y);
  split2(x[4], x[5], h.z, m.z, l.z);
  split2(x[6], x[7], h.w, m.w, l.w);
  *reinterpret_cast<uint4*>(dst0) = h;
  *reinterpret_cast<uint4*>(dst0 + plane_stride) = m;
  *reinterpret_cast<uint4*>(dst0 + 2 * plane_stride) = l;
}

__device__ __forceinline__ void zero_store(unsigned char* dst0, int plane_stride) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(dst0) = z;
  *reinterpret_cast<uint4*>(dst0 + plane_stride) = z;
  *reinterpret_cast<uint4*>(dst0 + 2 * plane_stride) = z;
}

// 8 channels [c0, c0+8) of row `src` (channels >= cin read as 0).  A32: 32-byte aligned rows, one 256-bit load.
template <bool A32>
__device__ __forceinline__ void load8(const float* __restrict__ src, int c0, int cin, float (&x)[8]) {
  if (A32) {
    if (c0 + 8 <= cin) {
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7])
                   : "l"(src + c0));
      return;
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int cb = c0 + 4 * h;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cb < cin) {   // ld_in is a multiple of 4 >= cin, so the 16-byte load stays inside the row
      v = __ldg(reinterpret_cast<const float4*>(src + cb));
      if (cb + 1 >= cin) v.y = 0.f;
      if (cb + 2 >= cin) v.z = 0.f;
      if (cb + 3 >= cin) v.w = 0.f;
    }
    x[4 * h] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w;
  }
}

// accumulator row -> residual add, two output slots with optional affine + relu (SgnnEpilogue semantics)
__device__ __forceinline__ void epilogue_row16(const Tc32Params& p, const unsigned (&v)[16], long long j) {
#pragma unroll
  for (int c = 0; c < 16; c += 4) {
    if (c >= p.cout) break;
    float4 a = make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]),
                           __uint_as_float(v[c + 3]));
    if (p.residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + j * p.ld_res + c));
      a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    }
    if (p.out_a) {
      float4 y = a;
      if (p.scale_a) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale_a + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift_a + c));
        y.x = fmaf(a.x, sc.x, sh.x); y.y = fmaf(a.y, sc.y, sh.y); y.z = fmaf(a.z, sc.z, sh.z); y.w = fmaf(a.w, sc.w, sh.w);
      }
      if (p.relu_a) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
      *reinterpret_cast<float4*>(p.out_a + j * p.ld_a + c) = y;
    }
    if (p.out_b) {
      float4 y = a;
      if (p.scale_b) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale_b + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift_b + c));
        y.x = fmaf(a.x, sc.x, sh.x); y.y = fmaf(a.y, sc.y, sh.y); y.z = fmaf(a.z, sc.z, sh.z); y.w = fmaf(a.w, sc.w, sh.w);
      }
      if (p.relu_b) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
      *reinterpret_cast<float4*>(p.out_b + j * p.ld_b + c) = y;
    }
  }
}

// the six partial products of one (A slice, B slice) pair: planes (i, j), i + j <= 2.  The leading product x0*w0 goes
// to the MAIN accumulator, the five correction products (2^-8 .. 2^-16 of it) to a separate CORRECTION accumulator;
// the epilogue adds the two.  The tensor core rounds each accumulate step on its own (not round-to-nearest-even), so
// keeping the small terms out of the main accumulator leaves it one rounding per (offset, slice) instead of six.
__device__ __forceinline__ void mma_split6(unsigned tmem_main, unsigned tmem_corr, unsigned a_base, unsigned b_base,
                                           unsigned acc_main, unsigned acc_corr) {
  const int pi[5] = {2, 1, 0, 1, 0};
  const int pj[5] = {0, 1, 2, 0, 1};
#pragma unroll
  for (int t = 0; t < 5; ++t)
    mma_bf16(tmem_corr, umma_desc(a_base + pi[t] * T32_ABLK), umma_desc(b_base + pj[t] * T32_BBLK), t == 0 ? acc_corr : 1u);
  mma_bf16(tmem_main, umma_desc(a_base), umma_desc(b_base), acc_main);
}

// ---------------------------------------------------------------------------------------------------------------
// filter preparation: fp32 [K][cin][16] -> split planes in the canonical B layout, [K][Q][3][512 B]
__device__ __forceinline__ void store_w_split(unsigned char* blk3, int co, int cil, float w) {
  const unsigned u = __float_as_uint(w);
  const float r = w - __uint_as_float(u & 0xffff0000u);
  const unsigned v = __float_as_uint(r);
  const float s = r - __uint_as_float(v & 0xffff0000u);
  const int off = (co >> 3) * 256 + (cil >> 3) * 128 + (co & 7) * 16 + (cil & 7) * 2;
  *reinterpret_cast<unsigned short*>(blk3 + off) = (unsigned short)(u >> 16);
  *reinterpret_cast<unsigned short*>(blk3 + T32_BBLK + off) = (unsigned short)(v >> 16);
  *reinterpret_cast<unsigned short*>(blk3 + 2 * T32_BBLK + off) = (unsigned short)(__float_as_uint(s) >> 16);
}

__global__ void tc32_prep_kernel(const float* __restrict__ w, int K, int cin, int cout, int Q, unsigned char* __restrict__ out) {
  const int total = K * Q * 256;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx & 15, cil = (idx >> 4) & 15, kq = idx >> 8, qc = kq % Q, k = kq / Q;
    const int ci = qc * 16 + cil;
    const float v = (ci < cin && co < cout) ? __ldg(w + ((size_t)k * cin + ci) * cout + co) : 0.f;
    store_w_split(out + (size_t)kq * 3 * T32_BBLK, co, cil, v);
  }
}

__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}


__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void mma_bf16_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(T32_IDESC), "r"(acc));
}

// 16 channels (two 8-channel units) of one row -> the three planes, 8 packed columns each
__device__ __forceinline__ void split16_tmem(const float (&x0)[8], const float (&x1)[8], unsigned taddr) {
  unsigned h[8], m[8], l[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    split2(x0[2 * i], x0[2 * i + 1], h[i], m[i], l[i]);
    split2(x1[2 * i], x1[2 * i + 1], h[4 + i], m[4 + i], l[4 + i]);
  }
  tmem_st8(taddr, h);
  tmem_st8(taddr + 8u, m);
  tmem_st8(taddr + 16u, l);
}


// child mode: parent offset reached from child c (z-major bit order) by filter offset d -- same arithmetic as
// conv_src_row() in conv.cu
__host__ __device__ __forceinline__ int child_parent_offset(int c, int d) {
  const int dz = d / 9 - 1, dy = (d / 3) % 3 - 1, dx = d % 3 - 1;
  const int pz = (((c >> 2) & 1) + dz + 2) / 2 - 1;
  const int py = (((c >> 1) & 1) + dy + 2) / 2 - 1;
  const int px = ((c & 1) + dx + 2) / 2 - 1;
  return (pz + 1) * 9 + (py + 1) * 3 + (px + 1);
}
// does child c read parent offset e at all?  (axis value -1 needs child bit 0, +1 needs child bit 1)
__host__ __device__ __forceinline__ bool child_uses(int c, int e) {
  const int ez = e / 9 - 1, ey = (e / 3) % 3 - 1, ex = e % 3 - 1;
  const int cz = (c >> 2) & 1, cy = (c >> 1) & 1, cx = c & 1;
  return (ez == 0 || ez == 2 * cz - 1) && (ey == 0 || ey == 2 * cy - 1) && (ex == 0 || ex == 2 * cx - 1);
}

}  // namespace
