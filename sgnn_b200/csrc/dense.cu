// dense.cu -- the coarse dense U-Net of the SG-NN encoder (reference torch/model.py:89-136,152-166; SURVEY §8
// row a12) as direct convolutions with a SPECIFIED summation order, so the whole generator is bit-reproducible
// (cuDNN picks algorithms per shape/batch and cost ~3.8 ms per step for 32 8^3 volumes -- 17 % of the step).
// The volumes are tiny (8^3 .. 2^3 per block): one thread per output element, weights broadcast within a warp
// (consecutive threads = consecutive x of the same output channel), inputs served by L1.
//   conv3d : out[b,co,o] = sum_{ci asc} sum_{kz,ky,kx asc} in[b,ci,o*s-p+k] * w[co,ci,k]        (nn.Conv3d)
//   convT3d: out[b,co,o] = sum_{ci asc} sum_{kz,ky,kx asc, (o+p-k)%s==0} in[b,ci,(o+p-k)/s] * w[ci,co,k] (nn.ConvTranspose3d)
// one fmaf chain from +0, then optional y = fmaf(y, scale[co], shift[co]) (eval BatchNorm3d folded on the host)
// and relu.  The input may be the channel concatenation of two tensors (torch.cat of model.py:156,160).
#include "common.cuh"


struct DenseArgs {
  const float* in0; int c0;
  const float* in1; int c1;
  int nb, d0, d1, d2;      // input extent
  int o0, o1, o2;          // output extent
  const float* w; int cout;
  int ks, stride, pad;
  const float* scale; const float* shift; int relu;
  float* out;
};

__device__ __forceinline__ const float* dense_chan(const DenseArgs& a, int b, int ci, long long vol) {
  return ci < a.c0 ? a.in0 + ((long long)b * a.c0 + ci) * vol
                   : a.in1 + ((long long)b * a.c1 + (ci - a.c0)) * vol;
}

__global__ void __launch_bounds__(128)
dense_conv3d_kernel(DenseArgs a) {
  const long long ovol = (long long)a.o0 * a.o1 * a.o2, ivol = (long long)a.d0 * a.d1 * a.d2;
  const long long total = (long long)a.nb * a.cout * ovol;
  const int cin = a.c0 + a.c1, k3 = a.ks * a.ks * a.ks;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % a.o2), y = (int)((idx / a.o2) % a.o1), z = (int)((idx / ((long long)a.o1 * a.o2)) % a.o0);
    const int co = (int)((idx / ovol) % a.cout), b = (int)(idx / (ovol * a.cout));
    const int z0 = z * a.stride - a.pad, y0 = y * a.stride - a.pad, x0 = x * a.stride - a.pad;
    float acc = 0.f;
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = dense_chan(a, b, ci, ivol);
      const float* wk = a.w + ((long long)co * cin + ci) * k3;
      for (int kz = 0; kz < a.ks; ++kz) {
        const int iz = z0 + kz;
        if ((unsigned)iz >= (unsigned)a.d0) continue;
        for (int ky = 0; ky < a.ks; ++ky) {
          const int iy = y0 + ky;
          if ((unsigned)iy >= (unsigned)a.d1) continue;
          const float* row = src + ((long long)iz * a.d1 + iy) * a.d2;
          const float* wr = wk + (kz * a.ks + ky) * a.ks;
          for (int kx = 0; kx < a.ks; ++kx) {
            const int ix = x0 + kx;
            if ((unsigned)ix < (unsigned)a.d2) acc = fmaf(__ldg(row + ix), __ldg(wr + kx), acc);
          }
        }
      }
    }
    float v = acc;
    if (a.scale) v = fmaf(v, __ldg(a.scale + co), __ldg(a.shift + co));
    if (a.relu) v = fmaxf(v, 0.f);
    a.out[idx] = v;
  }
}

// Specialisations of the direct convolution for the two filter sizes of the SG-NN coarse U-Net (4^3 stride 2 pad 1
// and 1^3): taps fully unrolled, all loads of an input channel issued before its fmaf chain so that load latency
// overlaps (the generic kernel interleaves a bounds test, a load and a dependent fmaf per tap: ~170 cycles per MAC on
// the 8192-output layer).  Same order: ci ascending, then kz,ky,kx ascending over the in-range taps.
template <int KS>
__global__ void __launch_bounds__(64)
dense_conv3d_fast_kernel(DenseArgs a) {
  const long long ovol = (long long)a.o0 * a.o1 * a.o2, ivol = (long long)a.d0 * a.d1 * a.d2;
  const long long total = (long long)a.nb * a.cout * ovol;
  const int cin = a.c0 + a.c1;
  constexpr int K3 = KS * KS * KS;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % a.o2), y = (int)((idx / a.o2) % a.o1), z = (int)((idx / ((long long)a.o1 * a.o2)) % a.o0);
    const int co = (int)((idx / ovol) % a.cout), b = (int)(idx / (ovol * a.cout));
    const int z0 = z * a.stride - a.pad, y0 = y * a.stride - a.pad, x0 = x * a.stride - a.pad;
    int off[K3];
    bool ok[K3];
#pragma unroll
    for (int t = 0; t < K3; ++t) {
      const int kz = t / (KS * KS), ky = (t / KS) % KS, kx = t % KS;
      const int iz = z0 + kz, iy = y0 + ky, ix = x0 + kx;
      ok[t] = (unsigned)iz < (unsigned)a.d0 && (unsigned)iy < (unsigned)a.d1 && (unsigned)ix < (unsigned)a.d2;
      off[t] = (iz * a.d1 + iy) * a.d2 + ix;
    }
    float acc = 0.f;
    if (KS == 1) {
      // 1^3 layers (bottleneck, final, heads): one load pair per channel -- taken 8 channels at a time so that 16 loads are in
      // flight instead of a dependent round trip per channel (26 us for the 28 -> 16 layer on 8^3 x 32 before)
      for (int c0 = 0; c0 < cin; c0 += 8) {
        float xv[8], wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int ci = c0 + u;
          const bool in = ci < cin && ok[0];
          xv[u] = in ? __ldg(dense_chan(a, b, in ? ci : 0, ivol) + off[0]) : 0.f;
          wv[u] = ci < cin ? __ldg(a.w + (long long)co * cin + ci) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < cin && ok[0]) acc = fmaf(xv[u], wv[u], acc);
      }
    } else {
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = dense_chan(a, b, ci, ivol);
      const float* wk = a.w + ((long long)co * cin + ci) * K3;
      float xv[K3], wv[K3];
#pragma unroll
      for (int t = 0; t < K3; ++t) {
        xv[t] = ok[t] ? __ldg(src + off[t]) : 0.f;
        wv[t] = __ldg(wk + t);
      }
#pragma unroll
      for (int t = 0; t < K3; ++t)
        if (ok[t]) acc = fmaf(xv[t], wv[t], acc);
    }
    }
    float v = acc;
    if (a.scale) v = fmaf(v, __ldg(a.scale + co), __ldg(a.shift + co));
    if (a.relu) v = fmaxf(v, 0.f);
    a.out[idx] = v;
  }
}

// The two strided encoder layers (nn.Conv3d k4 s2 p1: 16 -> 24 on 8^3, 24 -> 32 on 4^3 per sample) with both operands in
// shared memory.  One thread per output element as above (the fmaf chain of an output is sequential by definition: ci
// ascending, then kz,ky,kx over the in-range taps), a CTA = one sample x COG output channels:
//  * the sample's whole input is staged once, de-interleaved into the 8 PARITY PLANES of (z,y,x): output (z,y,x) reads input
//    2z+kz-1, so for a fixed tap every thread reads the same plane at its own (z,y,x) + a constant -- consecutive positions
//    are consecutive words: conflict-free, where the stride-2 reads of a plain layout collide 4-way;
//  * the filters of the CTA's channels stream through in chunks of CC input channels, laid out [ci][tap][co]: a warp reads
//    one word (same channel) or 4 consecutive words -- a broadcast;
//  * out-of-range taps are selected to +0 instead of branched around (fmaf(+0, w, acc) == acc), so the 64 taps of a channel
//    are 128 independent shared-memory loads followed by the chain.
// (A first shared-memory version without the parity planes and with a branch per tap was 2x SLOWER than the global-load
// kernel above: profiles/r02_ab_runs.txt.)
__global__ void __launch_bounds__(256, 2)
dense_conv3d_k4s2p1_pp_kernel(DenseArgs a, int cog, int cc, int spc) {
  extern __shared__ __align__(16) float dcs[];
  const int o0 = a.o0, o1 = a.o1, o2 = a.o2;
  const int ovol = o0 * o1 * o2, ivol = 8 * ovol, cin = a.c0 + a.c1;
  float* xin = dcs;                         // [spc samples][cin][8 planes][ovol]
  float* wsm = dcs + spc * cin * ivol;      // [cc][64][cog]
  // a CTA = spc samples x cog output channels (small volumes: several samples share one copy of the filters)
  const int b0 = blockIdx.x * spc, co0 = blockIdx.y * cog, tid = threadIdx.x, nthr = blockDim.x;
  const int ns = min(spc, a.nb - b0);
  // staging loops: 8 independent loads in flight per thread (one load -> one dependent shared-memory store per iteration
  // serialises a full memory round trip per element: measured 41 / 107 us per layer, most of it here)
  const int per_s = cin * ivol;
  for (int i0 = tid; i0 < ns * per_s; i0 += 8 * nthr) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * nthr;
      const int sl = i / per_s, q = i - sl * per_s;
      const int ci = q / ivol;
      v[u] = i < ns * per_s ? __ldg(dense_chan(a, b0 + sl, ci, ivol) + (q - ci * ivol)) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * nthr;
      if (i >= ns * per_s) break;
      const int sl = i / per_s, q = i - sl * per_s;
      const int ci = q / ivol, r = q - ci * ivol;
      const int ix = r % a.d2, iy = (r / a.d2) % a.d1, iz = r / (a.d1 * a.d2);
      const int plane = ((iz & 1) << 2) | ((iy & 1) << 1) | (ix & 1);
      xin[sl * per_s + ci * ivol + plane * ovol + ((iz >> 1) * o1 + (iy >> 1)) * o2 + (ix >> 1)] = v[u];
    }
  }
  const int sl = tid / (cog * ovol), rem = tid - sl * (cog * ovol);   // blockDim = spc * cog * ovol
  const int col = rem / ovol, pos = rem - col * ovol;
  const bool live = sl < ns;
  const int x = pos % o2, y = (pos / o2) % o1, z = pos / (o1 * o2);
  // tap k of an axis: input 2o + k - 1 = plane parity (k + 1) & 1, half-coordinate o + dk with dk = -1, 0, 0, +1
  bool vz[4], vy[4], vx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dk = k == 0 ? -1 : (k == 3 ? 1 : 0);
    vz[k] = (unsigned)(z + dk) < (unsigned)o0;
    vy[k] = (unsigned)(y + dk) < (unsigned)o1;
    vx[k] = (unsigned)(x + dk) < (unsigned)o2;
  }
  float acc = 0.f;
  for (int c0 = 0; c0 < cin; c0 += cc) {
    const int nc = min(cc, cin - c0);
    __syncthreads();                         // previous chunk consumed (and, the first time, the input staged)
    for (int i0 = tid; i0 < nc * 64 * cog; i0 += 8 * nthr) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * nthr;
        const int t = i & 63, cl = (i >> 6) % nc, j = (i >> 6) / nc;
        v[u] = i < nc * 64 * cog ? __ldg(a.w + ((long long)(co0 + j) * cin + c0 + cl) * 64 + t) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * nthr;
        if (i >= nc * 64 * cog) break;
        const int t = i & 63, cl = (i >> 6) % nc, j = (i >> 6) / nc;
        wsm[(cl * 64 + t) * cog + j] = v[u];
      }
    }
    __syncthreads();
    for (int cl = 0; cl < nc; ++cl) {
      const float* xc = xin + (live ? sl : 0) * per_s + (c0 + cl) * ivol + pos;
      const float* wc = wsm + cl * 64 * cog + col;
#pragma unroll
      for (int half = 0; half < 2; ++half) {       // 32 taps at a time: 64 loads in flight, <= 128 registers (2 CTAs per SM)
        float xv[32], wv[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const int t = half * 32 + u;
          const int kz = t >> 4, ky = (t >> 2) & 3, kx = t & 3;
          const int dz = kz == 0 ? -1 : (kz == 3 ? 1 : 0), dy = ky == 0 ? -1 : (ky == 3 ? 1 : 0), dx = kx == 0 ? -1 : (kx == 3 ? 1 : 0);
          const int plane = (((kz + 1) & 1) << 2) | (((ky + 1) & 1) << 1) | ((kx + 1) & 1);
          const bool ok = vz[kz] && vy[ky] && vx[kx];
          const int off = plane * ovol + (dz * o1 + dy) * o2 + dx;
          xv[u] = ok ? xc[off] : 0.f;
          wv[u] = wc[t * cog];
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) acc = fmaf(xv[u], wv[u], acc);
      }
    }
  }
  if (!live) return;
  const int co = co0 + col;
  float v = acc;
  if (a.scale) v = fmaf(v, __ldg(a.scale + co), __ldg(a.shift + co));
  if (a.relu) v = fmaxf(v, 0.f);
  a.out[((long long)(b0 + sl) * a.cout + co) * ovol + pos] = v;
}

// Transposed convolution: per axis only the taps k == (o + pad) (mod stride) reach an input cell; they are
// enumerated once per output element (at most 4 per axis), so the channel loop runs over live taps only.
// Order of the live taps is kz, ky, kx ascending -- the order of the full loop with the dead taps removed.
__global__ void __launch_bounds__(128)
dense_convT3d_kernel(DenseArgs a) {
  const long long ovol = (long long)a.o0 * a.o1 * a.o2, ivol = (long long)a.d0 * a.d1 * a.d2;
  const long long total = (long long)a.nb * a.cout * ovol;
  const int cin = a.c0 + a.c1, k2 = a.ks * a.ks, k3 = k2 * a.ks;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % a.o2), y = (int)((idx / a.o2) % a.o1), z = (int)((idx / ((long long)a.o1 * a.o2)) % a.o0);
    const int co = (int)((idx / ovol) % a.cout), b = (int)(idx / (ovol * a.cout));
    int kzs[4], izs[4], kys[4], iys[4], kxs[4], ixs[4];
    int nz = 0, ny = 0, nx = 0;
    for (int k = 0; k < a.ks; ++k) {
      int t = z + a.pad - k;
      if (t >= 0 && t % a.stride == 0 && t / a.stride < a.d0 && nz < 4) { kzs[nz] = k; izs[nz++] = t / a.stride; }
      t = y + a.pad - k;
      if (t >= 0 && t % a.stride == 0 && t / a.stride < a.d1 && ny < 4) { kys[ny] = k; iys[ny++] = t / a.stride; }
      t = x + a.pad - k;
      if (t >= 0 && t % a.stride == 0 && t / a.stride < a.d2 && nx < 4) { kxs[nx] = k; ixs[nx++] = t / a.stride; }
    }
    float acc = 0.f;
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = dense_chan(a, b, ci, ivol);
      const float* wk = a.w + ((long long)ci * a.cout + co) * k3;
#pragma unroll 2
      for (int iz = 0; iz < nz; ++iz) {
#pragma unroll 2
        for (int iy = 0; iy < ny; ++iy) {
          const float* row = src + ((long long)izs[iz] * a.d1 + iys[iy]) * a.d2;
          const float* wr = wk + kzs[iz] * k2 + kys[iy] * a.ks;
#pragma unroll 2
          for (int ix = 0; ix < nx; ++ix) acc = fmaf(__ldg(row + ixs[ix]), __ldg(wr + kxs[ix]), acc);
        }
      }
    }
    float v = acc;
    if (a.scale) v = fmaf(v, __ldg(a.scale + co), __ldg(a.shift + co));
    if (a.relu) v = fmaxf(v, 0.f);
    a.out[idx] = v;
  }
}

// Fast path of the two decoder layers (model.py:112,121): kernel 4, stride 2, padding 1 -> exactly two taps per
// axis, k in {(o+1)&1, ((o+1)&1)+2}, everything in registers.  Same tap order as the general kernel.
__global__ void __launch_bounds__(128)
dense_convT3d_k4s2p1_kernel(DenseArgs a) {
  const long long ovol = (long long)a.o0 * a.o1 * a.o2, ivol = (long long)a.d0 * a.d1 * a.d2;
  const long long total = (long long)a.nb * a.cout * ovol;
  const int cin = a.c0 + a.c1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % a.o2), y = (int)((idx / a.o2) % a.o1), z = (int)((idx / ((long long)a.o1 * a.o2)) % a.o0);
    const int co = (int)((idx / ovol) % a.cout), b = (int)(idx / (ovol * a.cout));
    const int kz0 = (z + 1) & 1, ky0 = (y + 1) & 1, kx0 = (x + 1) & 1;
    int off[8], wof[8];
    bool ok[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int kz = kz0 + 2 * (t >> 2), ky = ky0 + 2 * ((t >> 1) & 1), kx = kx0 + 2 * (t & 1);
      const int tz = z + 1 - kz, ty = y + 1 - ky, tx = x + 1 - kx;
      const int iz = tz >> 1, iy = ty >> 1, ix = tx >> 1;
      ok[t] = tz >= 0 && ty >= 0 && tx >= 0 && iz < a.d0 && iy < a.d1 && ix < a.d2;
      off[t] = (iz * a.d1 + iy) * a.d2 + ix;
      wof[t] = (kz * 4 + ky) * 4 + kx;
    }
    float acc = 0.f;
#pragma unroll 4
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = dense_chan(a, b, ci, ivol);
      const float* wk = a.w + ((long long)ci * a.cout + co) * 64;
      float xv[8], wv[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        xv[t] = ok[t] ? __ldg(src + off[t]) : 0.f;
        wv[t] = __ldg(wk + wof[t]);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (ok[t]) acc = fmaf(xv[t], wv[t], acc);
    }
    float v = acc;
    if (a.scale) v = fmaf(v, __ldg(a.scale + co), __ldg(a.shift + co));
    if (a.relu) v = fmaxf(v, 0.f);
    a.out[idx] = v;
  }
}

// Register-tiled version of the k4 s2 p1 transposed convolution (the two decoder layers are 60 % of the dense
// U-Net's time: the kernel above issues 16 loads per 8 fmaf).  An output cell with parities (z&1, y&1, x&1) reads
// the same 8 filter taps as every other cell of that parity class, and the cells of a class map 1:1 to input cells
// q: (z,y,x) = 2q + parity.  A CTA owns one (parity class, tile of 4 output channels): its cin x 8 taps x 4 filter
// values sit in shared memory ([ci][tap][4], read back as one warp-uniform LDS.128 per tap); a thread owns one
// (batch, q) and 4 accumulators -- 8 input loads + 8 LDS.128 per 32 fmaf.  Per output element the order is unchanged:
// ci ascending, then the in-range taps in kz,ky,kx order, one fmaf chain from +0.
#define DCT_CO 4
__global__ void __launch_bounds__(128)
dense_convT3d_k4s2p1_tile_kernel(DenseArgs a) {
  extern __shared__ __align__(16) float dct_w[];   // [cin][8][DCT_CO]
  const int cin = a.c0 + a.c1;
  const int cls = blockIdx.y & 7, co0 = (blockIdx.y >> 3) * DCT_CO;
  const int pz = (cls >> 2) & 1, py = (cls >> 1) & 1, px = cls & 1;          // output parities of this class
  const int kz0 = (pz + 1) & 1, ky0 = (py + 1) & 1, kx0 = (px + 1) & 1;
  for (int i = threadIdx.x; i < cin * 8 * DCT_CO; i += blockDim.x) {
    const int j = i % DCT_CO, t = (i / DCT_CO) & 7, ci = i / (8 * DCT_CO);
    const int kz = kz0 + 2 * (t >> 2), ky = ky0 + 2 * ((t >> 1) & 1), kx = kx0 + 2 * (t & 1);
    dct_w[i] = __ldg(a.w + ((long long)ci * a.cout + co0 + j) * 64 + (kz * 4 + ky) * 4 + kx);
  }
  __syncthreads();
  const long long ivol = (long long)a.d0 * a.d1 * a.d2, ovol = (long long)a.o0 * a.o1 * a.o2;
  const long long total = (long long)a.nb * ivol;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int qx = (int)(idx % a.d2), qy = (int)((idx / a.d2) % a.d1), qz = (int)((idx / ((long long)a.d1 * a.d2)) % a.d0);
    const int b = (int)(idx / ivol);
    const int z = 2 * qz + pz, y = 2 * qy + py, x = 2 * qx + px;
    int off[8];
    bool ok[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int kz = kz0 + 2 * (t >> 2), ky = ky0 + 2 * ((t >> 1) & 1), kx = kx0 + 2 * (t & 1);
      const int tz = z + 1 - kz, ty = y + 1 - ky, tx = x + 1 - kx;
      const int iz = tz >> 1, iy = ty >> 1, ix = tx >> 1;
      ok[t] = tz >= 0 && ty >= 0 && tx >= 0 && iz < a.d0 && iy < a.d1 && ix < a.d2;
      off[t] = (iz * a.d1 + iy) * a.d2 + ix;
    }
    float acc[DCT_CO];
#pragma unroll
    for (int j = 0; j < DCT_CO; ++j) acc[j] = 0.f;
#pragma unroll 2
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = dense_chan(a, b, ci, ivol);
      float xv[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) xv[t] = ok[t] ? __ldg(src + off[t]) : 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float4 wv = *reinterpret_cast<const float4*>(dct_w + (ci * 8 + t) * DCT_CO);
        if (ok[t]) {
          acc[0] = fmaf(xv[t], wv.x, acc[0]);
          acc[1] = fmaf(xv[t], wv.y, acc[1]);
          acc[2] = fmaf(xv[t], wv.z, acc[2]);
          acc[3] = fmaf(xv[t], wv.w, acc[3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < DCT_CO; ++j) {
      const int co = co0 + j;
      float v = acc[j];
      if (a.scale) v = fmaf(v, __ldg(a.scale + co), __ldg(a.shift + co));
      if (a.relu) v = fmaxf(v, 0.f);
      a.out[((long long)b * a.cout + co) * ovol + ((long long)z * a.o1 + y) * a.o2 + x] = v;
    }
  }
}

static int dense_launch(bool transposed, const float* in0, int c0, const float* in1, int c1, int nb, int d0, int d1,
                        int d2, const float* w, int cout, int ks, int stride, int pad, const float* scale,
                        const float* shift, int relu, float* out, void* stream) {
  if (nb < 0 || c0 <= 0 || c1 < 0 || d0 <= 0 || d1 <= 0 || d2 <= 0 || cout <= 0 || ks <= 0 || stride <= 0 ||
      pad < 0 || !in0 || (c1 > 0 && !in1) || !w || !out || (scale == nullptr) != (shift == nullptr))
    return SGNN_E_INVALID;
  DenseArgs a;
  a.in0 = in0; a.c0 = c0; a.in1 = in1; a.c1 = c1; a.nb = nb; a.d0 = d0; a.d1 = d1; a.d2 = d2;
  if (!transposed) {
    a.o0 = (d0 + 2 * pad - ks) / stride + 1; a.o1 = (d1 + 2 * pad - ks) / stride + 1; a.o2 = (d2 + 2 * pad - ks) / stride + 1;
  } else {
    a.o0 = (d0 - 1) * stride - 2 * pad + ks; a.o1 = (d1 - 1) * stride - 2 * pad + ks; a.o2 = (d2 - 1) * stride - 2 * pad + ks;
  }
  if (a.o0 <= 0 || a.o1 <= 0 || a.o2 <= 0) return SGNN_E_INVALID;
  if (transposed && (ks + stride - 1) / stride > 4) return SGNN_E_UNSUPPORTED;
  a.w = w; a.cout = cout; a.ks = ks; a.stride = stride; a.pad = pad; a.scale = scale; a.shift = shift; a.relu = relu;
  a.out = out;
  const long long total = (long long)nb * cout * a.o0 * a.o1 * a.o2;
  if (total == 0) return SGNN_OK;
  const int blocks = sgnn_blocks(total, 128, (int64_t)148 * 32);
  const int ovol_i = a.o0 * a.o1 * a.o2;
  int cog = 0;
  if (!transposed && ks == 4 && stride == 2 && pad == 1 && !(d0 & 1) && !(d1 & 1) && !(d2 & 1) && ovol_i <= 256 && nb <= 65535)
    for (int g = 256 / ovol_i > 8 ? 8 : 256 / ovol_i; g >= 1; --g)          // output channels per CTA (<= 8, <= 256 threads)
      if (cout % g == 0) { cog = g; break; }
  int spc = cog ? 256 / (cog * ovol_i) : 0;                              // samples per CTA: fill 256 threads
  if (spc > nb) spc = nb;
  if (spc < 1) spc = 1;
  const int dcc = cog ? (64 / cog > 0 ? 64 / cog : 1) : 0;              // <= 16 KB of filters per chunk
  const size_t dsm = cog ? ((size_t)spc * (c0 + c1) * 8 * ovol_i + (size_t)dcc * 64 * cog) * 4 : 0;
  if (cog && dsm <= 48 * 1024) {
    dim3 grid((unsigned)((nb + spc - 1) / spc), (unsigned)(cout / cog));
    dense_conv3d_k4s2p1_pp_kernel<<<grid, spc * cog * ovol_i, dsm, (cudaStream_t)stream>>>(a, cog, dcc, spc);
  } else if (!transposed && (ks == 4 || ks == 1)) {
    const int fb = sgnn_blocks(total, 64, (int64_t)148 * 64);
    if (ks == 4) dense_conv3d_fast_kernel<4><<<fb, 64, 0, (cudaStream_t)stream>>>(a);
    else dense_conv3d_fast_kernel<1><<<fb, 64, 0, (cudaStream_t)stream>>>(a);
  } else if (transposed && ks == 4 && stride == 2 && pad == 1 && cout % DCT_CO == 0 && (c0 + c1) * 8 * DCT_CO * 4 <= 48 * 1024 &&
             true) {
    // outputs = 8 parity classes x (nb x input cells): the extents are exactly twice the input's
    const long long cells = (long long)nb * d0 * d1 * d2;
    dim3 grid((unsigned)sgnn_blocks(cells, 128, 4096), (unsigned)(8 * (cout / DCT_CO)));
    dense_convT3d_k4s2p1_tile_kernel<<<grid, 128, (size_t)(c0 + c1) * 8 * DCT_CO * 4, (cudaStream_t)stream>>>(a);
  } else if (transposed && ks == 4 && stride == 2 && pad == 1) dense_convT3d_k4s2p1_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(a);
  else if (transposed) dense_convT3d_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(a);
  else dense_conv3d_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(a);
  SGNN_CHECK_LAUNCH();
  return SGNN_OK;
}

extern "C" int sgnn_dense_conv3d(const float* in0, int32_t c0, const float* in1, int32_t c1, int32_t nb, int32_t d0,
                                 int32_t d1, int32_t d2, const float* w, int32_t cout, int32_t ksize, int32_t stride,
                                 int32_t pad, const float* scale, const float* shift, int32_t relu, float* out,
                                 void* stream) {
  return dense_launch(false, in0, c0, in1, c1, nb, d0, d1, d2, w, cout, ksize, stride, pad, scale, shift, relu, out,
                      stream);
}

extern "C" int sgnn_dense_convT3d(const float* in0, int32_t c0, const float* in1, int32_t c1, int32_t nb, int32_t d0,
                                  int32_t d1, int32_t d2, const float* w, int32_t cout, int32_t ksize, int32_t stride,
                                  int32_t pad, const float* scale, const float* shift, int32_t relu, float* out,
                                  void* stream) {
  return dense_launch(true, in0, c0, in1, c1, nb, d0, d1, d2, w, cout, ksize, stride, pad, scale, shift, relu, out,
                      stream);
}
