// conv_sp.cu -- SubmanifoldConvolution forward over a COMPACT rulebook (SURVEY 8 rows a2 + a3; reference model.py:32,38,40
// on the encoder's input level).  The dense k-major neighbour table costs 27 index slots per site and the row-owner kernel
// (conv.cu) walks all 27 filter offsets of every row; at the 5 % occupancy of the encoder input a site has 2.3 present
// offsets, so 92 % of that walk -- and of the 108 B/site of table traffic -- is for nothing (round-1 VERDICT: 45 MB of table
// for 7.6 MB of rules, 60-70 us per 420 k-row layer).  Here the rulebook kernel (grid.cu, COMPACT) emits, per site, only the
// present offsets in ascending k, packed (k << 27 | input row) into slots[s][site], s < cnt[site], and this kernel visits
// exactly those.
//
// Mapping: lane = (row r = lane / 4 of the warp's 8 rows, channel group cg = lane % 4 owning COUT / 4 output channels).
// The 4 lanes of a row read the same input row (one L1 request, broadcast) and COUT/4 consecutive filter columns each from
// the shared-memory copy of the whole filter bank (27 x Cin x Cout fp32, <= 28 KB; k-stride padded so that rows of a warp
// sitting on different offsets fall on different banks).  CTAs are persistent (the bank is staged once per CTA).
//
// The same lane mapping also serves DENSE tables (sgnn_conv_forward's nbr[k][row], -1 = absent) for K = 27 layers of a few
// hundred to a few thousand rows, where the row-owner kernel's staged pipeline is pure latency.  There the taps are taken G at a
// time -- G indices, then G rows loaded back to back (absent taps load row 0 and multiply by +0, exactly what the row-owner
// kernel's zero rows do) -- so a thread has G independent loads in flight instead of one.  (Measured and NOT used: the same
// for the filter-2 stride-2 convolutions and for large launches -- the row-owner kernel's warp-uniform weight reads win there.)
//
// Arithmetic: per output element ONE fmaf chain from +0 over (k ascending, ci ascending) of the present offsets -- the order
// of oracle/o3.c, which skips absent offsets too -- so results are bit-identical to sgnn_conv_forward.
#include "common.cuh"

namespace {

struct SpParams {
  const float* in; int ld_in;
  const int* slots; const unsigned char* cnt; long long stride;
  const float* weight;
  long long n_out;
  const float* residual; int ld_res;
  float* out_a; int ld_a; int relu_a; const float* scale_a; const float* shift_a;
  float* out_b; int ld_b; int relu_b; const float* scale_b; const float* shift_b;
};

template <int COUT, int CIN>
struct SpCfg {
  static constexpr int CPL = COUT / 4;                       // output channels per lane
  static constexpr int WST = CIN * COUT + (COUT == 8 ? 8 : 16);   // floats per filter offset in shared memory
  static constexpr int SMEM = 27 * WST * 4;
};

template <int CPL>
__device__ __forceinline__ void sp_store(float* dst, const float (&v)[CPL]) {
  if constexpr (CPL == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (CPL == 2) *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
  else {
#pragma unroll
    for (int j = 0; j < CPL; ++j) dst[j] = v[j];
  }
}

template <int CPL, int COUT>
__device__ __forceinline__ void sp_fma_row(float (&acc)[CPL], float x, const float* __restrict__ w) {
  float wv[CPL];
  if constexpr (CPL == 4) {
    const float4 t = *reinterpret_cast<const float4*>(w);
    wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
  } else if constexpr (CPL == 2) {
    const float2 t = *reinterpret_cast<const float2*>(w);
    wv[0] = t.x; wv[1] = t.y;
  } else {
#pragma unroll
    for (int j = 0; j < CPL; ++j) wv[j] = w[j];
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = fmaf(x, wv[j], acc[j]);
}

// Loads as volatile asm: ptxas otherwise sinks every load next to its first use (index -> row -> fmaf, one dependent round
// trip per tap); volatile statements keep their program order, so the G index loads and then the G row loads of a tap group
// are issued back to back.
__device__ __forceinline__ int sp_ld_i32(const int* q) {
  int v;
  asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(q));
  return v;
}
template <int COUT, int CIN, bool VEC>
__device__ __forceinline__ void sp_load_row(float (&x)[CIN], const float* __restrict__ xr) {
  if constexpr (VEC && (CIN & 3) == 0) {
#pragma unroll
    for (int q = 0; q < CIN / 4; ++q)
      asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(x[4 * q]), "=f"(x[4 * q + 1]), "=f"(x[4 * q + 2]), "=f"(x[4 * q + 3]) : "l"(xr + 4 * q));
  } else if constexpr (VEC && (CIN & 1) == 0) {
#pragma unroll
    for (int q = 0; q < CIN / 2; ++q)
      asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(x[2 * q]), "=f"(x[2 * q + 1]) : "l"(xr + 2 * q));
  } else {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) asm volatile("ld.global.f32 %0, [%1];" : "=f"(x[ci]) : "l"(xr + ci));
  }
}

// K = 0: compact rulebook (slots / cnt).  K = 27: dense table p.slots[k][row] (-1 = absent), taps taken G at a time.
template <int COUT, int CIN, bool VEC, int K>
__global__ void __launch_bounds__(256)
conv_sp_kernel(SpParams p) {
  using Cfg = SpCfg<COUT, CIN>;
  constexpr int CPL = Cfg::CPL, WST = Cfg::WST;
  constexpr int NK = K == 0 ? 27 : K;
  extern __shared__ __align__(16) float Wsm[];
  for (int q = threadIdx.x; q < NK * CIN * COUT; q += 256) {
    const int k = q / (CIN * COUT), r = q - k * (CIN * COUT);
    Wsm[k * WST + r] = __ldg(p.weight + q);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (lane & 3) * CPL;
  const long long tiles = (p.n_out + 63) >> 6;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row = (tile << 6) + warp * 8 + (lane >> 2);
    const bool valid = row < p.n_out;
    float acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
    if constexpr (K == 0) {
      const int c = valid ? (int)__ldg(p.cnt + row) : 0;
      unsigned e = c > 0 ? (unsigned)__ldg(p.slots + row) : 0u;
      for (int s = 0; s < c; ++s) {
        const int k = (int)(e >> 27);
        const float* xr = p.in + (long long)(e & 0x7ffffffu) * p.ld_in;
        if (s + 1 < c) e = (unsigned)__ldg(p.slots + (long long)(s + 1) * p.stride + row);
        float x[CIN];
        sp_load_row<COUT, CIN, VEC>(x, xr);
        const float* w = Wsm + k * WST + c0;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) sp_fma_row<CPL, COUT>(acc, x[ci], w + ci * COUT);
      }
    } else {
      constexpr int G = CIN <= 8 ? 9 : (CIN <= 16 ? 3 : 1);
      static_assert(K % G == 0, "tap groups");
      // rows past the end walk row 0's taps (valid memory) and store nothing: the warp stays converged for the fences below
      const long long rr = valid ? row : 0;
#pragma unroll 1
      for (int k0 = 0; k0 < K; k0 += G) {
        int id[G];
#pragma unroll
        for (int g = 0; g < G; ++g) id[g] = sp_ld_i32(p.slots + (long long)(k0 + g) * p.stride + rr);
        asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");   // ptxas may not sink a load below a warp barrier: all G index loads are in flight together ...
        float x[G][CIN];
#pragma unroll
        for (int g = 0; g < G; ++g) sp_load_row<COUT, CIN, VEC>(x[g], p.in + (long long)(id[g] < 0 ? 0 : id[g]) * p.ld_in);
        asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");   // ... and so are the G row loads
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float* w = Wsm + (k0 + g) * WST + c0;
          const bool present = id[g] >= 0;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) sp_fma_row<CPL, COUT>(acc, present ? x[g][ci] : 0.f, w + ci * COUT);
        }
      }
    }
    if (!valid) continue;
    // epilogue: residual add, two slots with optional affine + relu (SgnnEpilogue)
    float v[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) v[j] = acc[j];
    if (p.residual) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] += __ldg(p.residual + row * p.ld_res + c0 + j);
    }
    if (p.out_a) {
      float y[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        y[j] = p.scale_a ? fmaf(v[j], __ldg(p.scale_a + c0 + j), __ldg(p.shift_a + c0 + j)) : v[j];
        if (p.relu_a) y[j] = fmaxf(y[j], 0.f);
      }
      sp_store<CPL>(p.out_a + row * p.ld_a + c0, y);
    }
    if (p.out_b) {
      float y[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        y[j] = p.scale_b ? fmaf(v[j], __ldg(p.scale_b + c0 + j), __ldg(p.shift_b + c0 + j)) : v[j];
        if (p.relu_b) y[j] = fmaxf(y[j], 0.f);
      }
      sp_store<CPL>(p.out_b + row * p.ld_b + c0, y);
    }
  }
}

template <int COUT, int CIN, int K>
int launch_sp(const SpParams& p, bool vec, cudaStream_t st) {
  using Cfg = SpCfg<COUT, CIN>;
  const long long tiles = (p.n_out + 63) >> 6;
  const int grid = (int)(tiles < 148 * 8 ? tiles : 148 * 8);
  if (Cfg::SMEM > 48 * 1024) {     // per device: the opt-in is a per-context function attribute
    static bool done[64] = {false};
    int dev = 0;
    SGNN_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !done[dev]) {
      SGNN_CUDA(cudaFuncSetAttribute(conv_sp_kernel<COUT, CIN, true, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
      SGNN_CUDA(cudaFuncSetAttribute(conv_sp_kernel<COUT, CIN, false, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
      done[dev] = true;
    }
  }
  if (vec) conv_sp_kernel<COUT, CIN, true, K><<<grid, 256, Cfg::SMEM, st>>>(p);
  else conv_sp_kernel<COUT, CIN, false, K><<<grid, 256, Cfg::SMEM, st>>>(p);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

bool al16(const void* q) { return ((uintptr_t)q & 15) == 0; }

}  // namespace

static int sp_params(const SgnnConvArgs* a, SpParams* p, bool* vec) {
  if (!a || a->n_out < 0 || !a->weight) return SGNN_E_INVALID;
  if (a->dtype != SGNN_F32 || a->child_mode) return SGNN_E_UNSUPPORTED;
  if (a->n_out == 0) return SGNN_OK;
  if (!a->in || (!a->a.out && !a->b.out) || a->nbr_stride < a->n_out) return SGNN_E_INVALID;
  const SgnnEpilogue* eps[2] = {&a->a, &a->b};
  for (const SgnnEpilogue* e : eps) {
    if (!e->out) continue;
    if ((e->scale == nullptr) != (e->shift == nullptr)) return SGNN_E_INVALID;
    if (!al16(e->out) || (e->ld & 3)) return SGNN_E_ALIGN;
  }
  if (a->residual && (a->ld_res & 3)) return SGNN_E_ALIGN;
  p->in = (const float*)a->in; p->ld_in = a->ld_in; p->stride = a->nbr_stride;
  p->weight = (const float*)a->weight; p->n_out = a->n_out;
  p->residual = (const float*)a->residual; p->ld_res = a->ld_res;
  p->out_a = (float*)a->a.out; p->ld_a = a->a.ld; p->relu_a = a->a.relu; p->scale_a = a->a.scale; p->shift_a = a->a.shift;
  p->out_b = (float*)a->b.out; p->ld_b = a->b.ld; p->relu_b = a->b.relu; p->scale_b = a->b.scale; p->shift_b = a->b.shift;
  *vec = al16(a->in) && (a->ld_in & 3) == 0;
  return SGNN_OK;
}

extern "C" int sgnn_conv_forward_compact(const SgnnConvArgs* a, const int32_t* slots, const uint8_t* cnt, void* stream) {
  SpParams p;
  bool vec = false;
  int rc = sp_params(a, &p, &vec);
  if (rc || a->n_out == 0) return rc;
  if (a->K != 27) return SGNN_E_UNSUPPORTED;
  if (!slots || !cnt) return SGNN_E_INVALID;
  if (a->n_out >= (1LL << 27)) return SGNN_E_TOO_LARGE;
  p.slots = slots; p.cnt = cnt;
  cudaStream_t st = (cudaStream_t)stream;
#define SP_CASE(CO, CI) if (a->cout == CO && a->cin == CI) return launch_sp<CO, CI, 0>(p, vec, st);
  SP_CASE(8, 1) SP_CASE(8, 8) SP_CASE(12, 8) SP_CASE(12, 12) SP_CASE(16, 12) SP_CASE(16, 16)
#undef SP_CASE
  return SGNN_E_UNSUPPORTED;
}

// sgnn_conv_forward's dense-table form on the lane = (row, channel group) kernel: the latency-bound launches (see the
// header).  Internal: generator.cu and sgnn_conv_forward route to it; same bits as the row-owner kernel.
int sgnn_conv_forward_rowlane(const SgnnConvArgs* a, cudaStream_t st) {
  SpParams p;
  bool vec = false;
  int rc = sp_params(a, &p, &vec);
  if (rc || a->n_out == 0) return rc;
  if (!a->nbr) return SGNN_E_INVALID;
  p.slots = a->nbr; p.cnt = nullptr;
#define SP_CASE(CO, CI, KK) if (a->cout == CO && a->cin == CI && a->K == KK) return launch_sp<CO, CI, KK>(p, vec, st);
  SP_CASE(8, 1, 27) SP_CASE(8, 8, 27) SP_CASE(12, 8, 27) SP_CASE(12, 12, 27) SP_CASE(16, 12, 27) SP_CASE(16, 16, 27)
#undef SP_CASE
  return SGNN_E_UNSUPPORTED;
}
