/* ORACLE O3 -- TEST INFRASTRUCTURE ONLY (never linked into or called from the product path).
 *
 * Plain-C, fixed-summation-order restatement of the floating-point part of the sparse path
 * (SURVEY App. A.4-A.8, §7 step 4).  scn's CPU path computes out[o] += in[i] @ W[k] per filter offset
 * with a BLAS mm whose internal order is unspecified; this oracle PINS one order -- k ascending, then
 * ci ascending, a single fmaf chain per output element starting from +0 -- which the sm_100a kernels in
 * sgnn_b200/csrc/conv.cu reproduce bit for bit.  PARITY UNPINNED against upstream SparseConvNet itself
 * (not available offline); O3 is checked against oracle O2 (oracle/sparseconvnet, torch mm order) to
 * 1e-5 and O2 against the dense conv3d identity O1 (oracle/dense_equiv.py).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libo3.so oracle/o3.c -lm
 */
#include <math.h>
#include <stdint.h>

static int src_row(const int32_t* nbr, int64_t stride, int child_mode, int64_t j, int k) {
  if (!child_mode) return nbr[(int64_t)k * stride + j];
  int c = (int)(j & 7);
  int dz = k / 9 - 1, dy = (k / 3) % 3 - 1, dx = k % 3 - 1;
  int pz = (((c >> 2) & 1) + dz + 2) / 2 - 1;
  int py = (((c >> 1) & 1) + dy + 2) / 2 - 1;
  int px = ((c & 1) + dx + 2) / 2 - 1;
  int kp = (pz + 1) * 9 + (py + 1) * 3 + (px + 1);
  return nbr[(int64_t)kp * stride + (j >> 3)];
}

/* out[j] = sum_k in[nbr[k][j]] @ W[k] (+residual) ; optional y = fmaf(y, scale, shift) ; optional relu
 * (model.py:32,38,40,44,179,186,254; upstream *_updateOutput) */
void o3_conv(const float* in, int ld_in, const int32_t* nbr, int64_t nbr_stride, int K, int child_mode,
             const float* W, int cin, int cout, int64_t n_out, const float* residual, int ld_res,
             float* out, int ld_out, const float* scale, const float* shift, int relu) {
  for (int64_t j = 0; j < n_out; ++j) {
    for (int co = 0; co < cout; ++co) {
      float acc = 0.0f;
      for (int k = 0; k < K; ++k) {
        int r = src_row(nbr, nbr_stride, child_mode, j, k);
        if (r < 0) continue;
        const float* x = in + (int64_t)r * ld_in;
        const float* w = W + (int64_t)k * cin * cout + co;
        for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], w[(int64_t)ci * cout], acc);
      }
      float v = acc;
      if (residual) v = v + residual[j * ld_res + co];
      if (scale) v = fmaf(v, scale[co], shift[co]);
      if (relu) v = v > 0.0f ? v : 0.0f;
      out[j * ld_out + co] = v;
    }
  }
}

/* scn.Deconvolution filter 2 stride 2: out[i] = in[parent>>3] @ W[parent&7] */
void o3_deconv(const float* in, int ld_in, const int32_t* parent, const float* W, int cin, int cout,
               int64_t n, float* out, int ld_out) {
  for (int64_t i = 0; i < n; ++i)
    for (int co = 0; co < cout; ++co) {
      float acc = 0.0f;
      int pk = parent[i];
      if (pk >= 0) {
        const float* x = in + (int64_t)(pk >> 3) * ld_in;
        const float* w = W + (int64_t)(pk & 7) * cin * cout + co;
        for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], w[(int64_t)ci * cout], acc);
      }
      out[i * ld_out + co] = acc;
    }
}

/* scn.BatchNormReLU eval with host-folded constants (App. A.8) */
void o3_affine_relu(const float* x, int ld_x, float* y, int ld_y, int64_t n, int c, const float* scale,
                    const float* shift, int relu) {
  for (int64_t i = 0; i < n; ++i)
    for (int ch = 0; ch < c; ++ch) {
      float v = x[i * ld_x + ch];
      if (scale) v = fmaf(v, scale[ch], shift[ch]);
      if (relu) v = v > 0.0f ? v : 0.0f;
      y[i * ld_y + ch] = v;
    }
}

/* nn.Linear heads (model.py:230-231,271) */
void o3_linear(const float* x, int ld_x, const float* w, const float* b, float* y, int ld_y, int64_t n,
               int cin, int cout) {
  for (int64_t i = 0; i < n; ++i)
    for (int o = 0; o < cout; ++o) {
      float acc = 0.0f;
      for (int c = 0; c < cin; ++c) acc = fmaf(x[i * ld_x + c], w[(int64_t)o * cin + c], acc);
      if (b) acc = acc + b[o];
      y[i * ld_y + o] = acc;
    }
}

/* literal fp32 `sigmoid(x) > 0.5` (model.py:233,322) */
void o3_sigmoid_gt_half(const float* x, int64_t n, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    float s = 1.0f / (1.0f + expf(-x[i]));
    out[i] = s > 0.5f ? 1 : 0;
  }
}

/* Dense coarse U-Net layers (model.py:89-136,152-166), NCDHW, fixed order: ci asc, then kz,ky,kx asc.
 * transposed == 0: nn.Conv3d, weight [cout][cin][k][k][k]; transposed == 1: nn.ConvTranspose3d, weight
 * [cin][cout][k][k][k].  Optional folded BatchNorm3d (scale, shift) and relu. */
void o3_dense_conv(int transposed, const float* in, int cin, int nb, int d0, int d1, int d2, const float* w,
                   int cout, int ks, int stride, int pad, const float* scale, const float* shift, int relu,
                   float* out, int o0, int o1, int o2) {
  int64_t ivol = (int64_t)d0 * d1 * d2, ovol = (int64_t)o0 * o1 * o2;
  int k3 = ks * ks * ks;
  for (int b = 0; b < nb; ++b)
    for (int co = 0; co < cout; ++co)
      for (int z = 0; z < o0; ++z)
        for (int y = 0; y < o1; ++y)
          for (int x = 0; x < o2; ++x) {
            float acc = 0.0f;
            for (int ci = 0; ci < cin; ++ci) {
              const float* src = in + ((int64_t)b * cin + ci) * ivol;
              const float* wk = transposed ? w + ((int64_t)ci * cout + co) * k3 : w + ((int64_t)co * cin + ci) * k3;
              for (int kz = 0; kz < ks; ++kz)
                for (int ky = 0; ky < ks; ++ky)
                  for (int kx = 0; kx < ks; ++kx) {
                    int iz, iy, ix;
                    if (!transposed) {
                      iz = z * stride - pad + kz; iy = y * stride - pad + ky; ix = x * stride - pad + kx;
                    } else {
                      int tz = z + pad - kz, ty = y + pad - ky, tx = x + pad - kx;
                      if (tz < 0 || ty < 0 || tx < 0 || tz % stride || ty % stride || tx % stride) continue;
                      iz = tz / stride; iy = ty / stride; ix = tx / stride;
                    }
                    if (iz < 0 || iy < 0 || ix < 0 || iz >= d0 || iy >= d1 || ix >= d2) continue;
                    acc = fmaf(src[((int64_t)iz * d1 + iy) * d2 + ix], wk[(kz * ks + ky) * ks + kx], acc);
                  }
            }
            float v = acc;
            if (scale) v = fmaf(v, scale[co], shift[co]);
            if (relu) v = v > 0.0f ? v : 0.0f;
            out[(((int64_t)b * cout + co) * o0 + z) * o1 * o2 + (int64_t)y * o2 + x] = v;
          }
}
