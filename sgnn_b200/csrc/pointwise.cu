// pointwise.cu -- per-row kernels: BatchNormReLU (eval), AddTable, JoinTable slot copy, nn.Linear heads,
// SparseToDense.  All HBM-bound; rows are short (<= 48 floats) so the
// mapping is one thread per element with the channel index fastest (coalesced across a row).
#include "common.cuh"

__global__ void affine_relu_kernel(const float* __restrict__ x, int ld_x, float* __restrict__ y, int ld_y,
                                   long long n, int c, const float* __restrict__ scale,
                                   const float* __restrict__ shift, int relu) {
  const long long total = n * c;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / c;
    const int ch = (int)(idx % c);
    float v = x[i * ld_x + ch];
    if (scale) v = fmaf(v, __ldg(scale + ch), __ldg(shift + ch));
    if (relu) v = fmaxf(v, 0.f);
    y[i * ld_y + ch] = v;
  }
}

extern "C" int sgnn_affine_relu(const float* x, int32_t ld_x, float* y, int32_t ld_y, int64_t n, int32_t c,
                                const float* scale, const float* shift, int32_t relu, void* stream) {
  if (n < 0 || c <= 0 || (scale == nullptr) != (shift == nullptr)) return SGNN_E_INVALID;
  if (n == 0) return SGNN_OK;
  if (!x || !y) return SGNN_E_INVALID;
  affine_relu_kernel<<<sgnn_blocks(n * c, 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, y, ld_y, (long long)n, c,
                                                                                scale, shift, relu);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

__global__ void add_rows_kernel(const float* __restrict__ a, int ld_a, const float* __restrict__ b, int ld_b,
                                float* __restrict__ y, int ld_y, long long n, int c) {
  const long long total = n * c;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / c;
    const int ch = (int)(idx % c);
    y[i * ld_y + ch] = a[i * ld_a + ch] + b[i * ld_b + ch];
  }
}

extern "C" int sgnn_add_rows(const float* a, int32_t ld_a, const float* b, int32_t ld_b, float* y, int32_t ld_y,
                             int64_t n, int32_t c, void* stream) {
  if (n < 0 || c <= 0) return SGNN_E_INVALID;
  if (n == 0) return SGNN_OK;
  if (!a || !b || !y) return SGNN_E_INVALID;
  add_rows_kernel<<<sgnn_blocks(n * c, 256), 256, 0, (cudaStream_t)stream>>>(a, ld_a, b, ld_b, y, ld_y,
                                                                             (long long)n, c);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                                 long long n, int c) {
  const long long total = n * c;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / c;
    const int ch = (int)(idx % c);
    dst[i * ld_dst + ch] = src[i * ld_src + ch];
  }
}

extern "C" int sgnn_copy_cols(const float* src, int32_t ld_src, float* dst, int32_t ld_dst, int64_t n, int32_t c,
                              void* stream) {
  if (n < 0 || c <= 0) return SGNN_E_INVALID;
  if (n == 0) return SGNN_OK;
  if (!src || !dst) return SGNN_E_INVALID;
  copy_cols_kernel<<<sgnn_blocks(n * c, 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst, ld_dst,
                                                                              (long long)n, c);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// y[i][o] = (fma chain over c ascending of x[i][c]*w[o][c], from +0) + b[o]
__global__ void linear_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ w,
                              const float* __restrict__ b, float* __restrict__ y, int ld_y, long long n, int cin,
                              int cout) {
  const long long total = n * cout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / cout;
    const int o = (int)(idx % cout);
    const float* xr = x + i * ld_x;
    const float* wr = w + (size_t)o * cin;
    float acc = 0.f;
    for (int c = 0; c < cin; ++c) acc = fmaf(xr[c], __ldg(wr + c), acc);
    if (b) acc += __ldg(b + o);
    y[i * ld_y + o] = acc;
  }
}

// cout == 1 (TSDF head, model.py:258,271): 4 lanes per row read the row coalesced and hand the SAME sequential fmaf
// chain from lane to lane (lane q continues with channels [q*cin/4, (q+1)*cin/4)) -- identical bits, 4x fewer
// uncoalesced 192-byte row walks.
__global__ void linear1_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ y, int ld_y, long long n, int cin) {
  const int q = threadIdx.x & 3;
  const int per = cin >> 2;
  const bool vec4 = (per & 3) == 0 && (ld_x & 3) == 0 && ((size_t)x & 15) == 0;
  for (long long i0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 2; ; i0 += ((long long)gridDim.x * blockDim.x) >> 2) {
    const long long base = i0 - (i0 & 7);   // rows handled by this warp iteration: keep the warp converged for shuffles
    if (base >= n) break;
    const bool live = i0 < n;
    const float* xr = x + (live ? i0 : 0) * ld_x + q * per;
    float v[16];
    if (vec4) {                       // per, ld_x multiples of 4 and x 16-byte aligned: 16-byte loads of the lane's quarter row
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && c < per) t = *reinterpret_cast<const float4*>(xr + c);
        v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = (live && c < per) ? xr[c] : 0.f;
    }
    float acc = 0.f;
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      if (q == step) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < per) acc = fmaf(v[c], __ldg(w + step * per + c), acc);
      }
      const float nxt = __shfl_sync(0xffffffffu, acc, (threadIdx.x & 28) | step, 32);
      if (q == step + 1) acc = nxt;
      if (step == 3 && q == 0) acc = nxt;
    }
    if (live && q == 0) y[i0 * ld_y] = b ? acc + __ldg(b) : acc;
  }
}

extern "C" int sgnn_linear(const float* x, int32_t ld_x, const float* w, const float* b, float* y, int32_t ld_y,
                           int64_t n, int32_t cin, int32_t cout, void* stream) {
  if (n < 0 || cin <= 0 || cout <= 0 || !w) return SGNN_E_INVALID;
  if (n == 0) return SGNN_OK;
  if (!x || !y) return SGNN_E_INVALID;
  if (cout == 1 && (cin & 3) == 0 && cin <= 64) {
    linear1_kernel<<<sgnn_blocks(n * 4, 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, w, b, y, ld_y, (long long)n, cin);
    SGNN_CHECK_LAUNCH();
    return SGNN_OK;
  }
  linear_kernel<<<sgnn_blocks(n * cout, 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, w, b, y, ld_y, (long long)n,
                                                                              cin, cout);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

__global__ void sparse_to_dense_kernel(const float* __restrict__ feats, int ld, const int* __restrict__ coords,
                                       long long n, int c, float* __restrict__ dense, int nb, int d0, int d1,
                                       int d2) {
  const long long total = n * c;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / c;
    const int ch = (int)(idx % c);
    const int4 p = __ldg(reinterpret_cast<const int4*>(coords) + i);
    if ((unsigned)p.x >= (unsigned)d0 || (unsigned)p.y >= (unsigned)d1 || (unsigned)p.z >= (unsigned)d2 ||
        (unsigned)p.w >= (unsigned)nb)
      continue;
    dense[((((long long)p.w * c + ch) * d0 + p.x) * d1 + p.y) * d2 + p.z] = feats[i * ld + ch];
  }
}

extern "C" int sgnn_sparse_to_dense(const float* feats, int32_t ld, const int32_t* coords, int64_t n, int32_t c,
                                    float* dense, int32_t nb, int32_t d0, int32_t d1, int32_t d2, void* stream) {
  if (n < 0 || c <= 0 || nb < 0 || d0 < 0 || d1 < 0 || d2 < 0 || !dense) return SGNN_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)nb * c * d0 * d1 * d2 * 4;
  if (bytes) SGNN_CUDA(cudaMemsetAsync(dense, 0, bytes, st));
  if (n == 0 || bytes == 0) return SGNN_OK;
  if (!feats || !coords) return SGNN_E_INVALID;
  sparse_to_dense_kernel<<<sgnn_blocks(n * c, 256), 256, 0, st>>>(feats, ld, coords, (long long)n, c, dense, nb, d0,
                                                                  d1, d2);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

// ------------------------------------------------------------ measured FFMA ceiling (bench.py roofline denominator)
// 3-register FFMAs, 16 independent chains per thread, operands in registers: what the fp32 pipe of this part
// actually sustains (the fp32 convolution is FFMA bound, not HBM bound -- DESIGN.md §5).
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  float x = a + threadIdx.x * 1e-9f, y = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
      x += 1e-9f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

extern "C" int sgnn_debug_ffma_peak(int iters, double* tflops, void* stream) {
  if (!tflops || iters <= 0) return SGNN_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  float* d = nullptr;
  SGNN_CUDA(cudaMalloc(&d, 4));
  cudaEvent_t e0, e1;
  SGNN_CUDA(cudaEventCreate(&e0));
  SGNN_CUDA(cudaEventCreate(&e1));
  const int blocks = 148 * 8;
  ffma_peak_kernel<<<blocks, 256, 0, st>>>(d, 16, 1.0001f, 0.5f);   // warm-up
  SGNN_CUDA(cudaEventRecord(e0, st));
  ffma_peak_kernel<<<blocks, 256, 0, st>>>(d, iters, 1.0001f, 0.5f);
  SGNN_CHECK_LAUNCH();
  SGNN_CUDA(cudaEventRecord(e1, st));
  SGNN_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  SGNN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = 2.0 * 128.0 * (double)iters * blocks * 256 / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return SGNN_OK;
}
