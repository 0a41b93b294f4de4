"""Structured probes of the tc32 tensor-core convolution (run on the GPU box; prints error patterns that localise a
layout / descriptor / pipeline mistake from one run).  Scratch tool, not part of the product or the tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import o3                                   # noqa: E402
import sgnn_b200.engine as E                # noqa: E402
from helpers import random_coords, nbr_table  # noqa: E402


def report(tag, got, want):
    got = got.cpu()
    err = (got - want).abs()
    bad = ~torch.isfinite(got)
    scale = float(want.abs().max()) + 1e-30
    print('%-44s max|err| %.3e  rel %.2e  nonfinite %d' % (tag, float(err[~bad].max()) if (~bad).any() else -1,
                                                            float(err[~bad].max()) / scale if (~bad).any() else -1,
                                                            int(bad.sum())))
    if float(err.nan_to_num(1e9).max()) / scale > 1e-4:
        e = err.nan_to_num(1e9)
        n = e.shape[0]
        print('   per-column max :', ' '.join('%.1e' % v for v in e.max(0).values.tolist()))
        rows = torch.arange(n)
        print('   by row%8       :', ' '.join('%.1e' % float(e[rows % 8 == i].max()) for i in range(8)))
        print('   by (row//8)%16 :', ' '.join('%.1e' % float(e[(rows // 8) % 16 == i].max()) if ((rows // 8) % 16 == i).any()
                                              else '-' for i in range(16)))
        print('   by row//128    :', ' '.join('%.1e' % float(e[rows // 128 == i].max()) for i in range(min(8, (n + 127) // 128))))
        r = int(e.max(1).values.argmax())
        print('   worst row %d got  :' % r, ' '.join('%+.4f' % v for v in got[r].tolist()))
        print('   worst row %d want :' % r, ' '.join('%+.4f' % v for v in want[r].tolist()))


def run(x, nbr, w, n_out, **kw):
    out = torch.full((n_out, 16), float('nan'), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), n_out, out, tc32=True, **kw)
    torch.cuda.synchronize()
    return out


def main():
    from sgnn_b200._lib import lib
    lib.sgnn_debug_set_conv_impl(int(os.environ.get('SGNN_DIAG_IMPL', '0')))
    rng = np.random.default_rng(0)
    for n in (128, 100, 300):
        x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
        ident = torch.full((27, n), -1, dtype=torch.int32)
        ident[13] = torch.arange(n, dtype=torch.int32)
        w = torch.zeros((27, 16, 16))
        w[13] = torch.eye(16)
        report('n=%d centre tap, W=I (out == x)' % n, run(x, ident, w, n), x)
        w[13] = torch.from_numpy(rng.standard_normal((16, 16)).astype(np.float32))
        report('n=%d centre tap, random W' % n, run(x, ident, w, n), o3.conv(x, ident, w, n))
        w1 = torch.zeros((27, 16, 16))
        w1[13, 3, 5] = 1.0
        report('n=%d centre tap, W[3,5]=1 (out[:,5] == x[:,3])' % n, run(x, ident, w1, n), o3.conv(x, ident, w1, n))
        allk = torch.arange(n, dtype=torch.int32).repeat(27, 1)
        wr = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.1).astype(np.float32))
        report('n=%d all 27 taps -> same row, random W' % n, run(x, allk, wr, n), o3.conv(x, allk, wr, n))
        perm = torch.stack([torch.from_numpy(rng.permutation(n).astype(np.int32)) for _ in range(27)])
        report('n=%d random gathers, random W' % n, run(x, perm, wr, n), o3.conv(x, perm, wr, n))
    for cin in (12, 26, 30, 34, 48):
        n = 200
        ld = (cin + 3) // 4 * 4
        xb = torch.zeros((n, ld))
        xb[:, :cin] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
        perm = torch.stack([torch.from_numpy(rng.permutation(n).astype(np.int32)) for _ in range(27)])
        perm[rng.random((27, n)) < 0.3] = -1
        wr = torch.from_numpy((rng.standard_normal((27, cin, 16)) * 0.1).astype(np.float32))
        out = torch.full((n, 16), float('nan'), device='cuda')
        E.conv(xb.cuda()[:, :cin], perm.cuda(), wr.cuda(), n, out, tc32=True)
        report('cin=%d random gathers with holes' % cin, out, o3.conv(xb[:, :cin], perm, wr, n))
    # K = 8 (strided)
    n, nc = 300, 90
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    ch = torch.from_numpy(rng.integers(-1, n, (8, nc)).astype(np.int32))
    w8 = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.2).astype(np.float32))
    report('K=8 strided', run(x, ch, w8, nc), o3.conv(x, ch, w8, nc))
    # child mode
    c = random_coords(rng, 2, (7, 6, 9), 0.4)
    n = c.shape[0]
    nbr = torch.from_numpy(nbr_table(c))
    x48 = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    wc = torch.zeros((27, 48, 16))
    wc[13, :16, :] = torch.eye(16)
    report('child: centre tap W=I on ch 0..15 (out[8p+c] == x[p])', run(x48, nbr, wc, 8 * n, child_mode=True),
           o3.conv(x48, nbr, wc, 8 * n, child_mode=True))
    wc = torch.from_numpy((rng.standard_normal((27, 48, 16)) * 0.05).astype(np.float32))
    got = run(x48, nbr, wc, 8 * n, child_mode=True)
    want = o3.conv(x48, nbr, wc, 8 * n, child_mode=True)
    report('child: random W, %d parents' % n, got, want)
    e = (got.cpu() - want).abs().max(1).values
    print('   child err by child index:', ' '.join('%.1e' % float(e[i::8].max()) for i in range(8)))
    # accuracy statistics on a bigger case: error in units of fp32 ulp of the accumulated magnitude
    c = random_coords(rng, 4, (24, 24, 24), 0.3)
    n = c.shape[0]
    nbr = torch.from_numpy(nbr_table(c))
    x = torch.from_numpy(np.abs(rng.standard_normal((n, 16))).astype(np.float32))          # post-ReLU like: no cancellation in x
    w = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.1).astype(np.float32))
    want64 = None
    try:
        xx, ww = x.double(), w.double()
        want64 = torch.zeros((n, 16), dtype=torch.float64)
        for k in range(27):
            r = nbr[k].long()
            m = r >= 0
            want64[m] += xx[r[m]] @ ww[k]
    except Exception as ex:      # pragma: no cover
        print('fp64 reference failed', ex)
    got = run(x, nbr, w, n).cpu().double()
    ffma = torch.empty((n, 16), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, ffma)
    mag = o3.conv(x.abs(), nbr, w.abs(), n).double()
    if want64 is not None:
        print('accuracy vs fp64 (max err / sum|x||w|): tc32 %.2e   ffma %.2e   (n=%d)' %
              (float(((got - want64).abs() / mag).max()), float(((ffma.cpu().double() - want64).abs() / mag).max()), n))
        print('mean signed err / sum|x||w|:           tc32 %+.2e   ffma %+.2e' %
              (float(((got - want64) / mag).mean()), float(((ffma.cpu().double() - want64) / mag).mean())))


if __name__ == '__main__':
    main()
