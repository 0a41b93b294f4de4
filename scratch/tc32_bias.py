"""Where does the tc32 generator deviate from the golden fixtures?  (1) kernel-level signed error toward zero
(round-toward-zero accumulation shows up as a coherent shrinkage), (2) per-level deviations of exact / tc32 vs the
golden fixture and vs each other, (3) the fused Python path with only the regular or only the child-mode
convolutions routed to tc32.  Scratch tool."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sgnn_b200                            # noqa: E402
import sgnn_b200.engine as E                # noqa: E402
from sgnn_b200.synth import fill_parameters  # noqa: E402
from helpers import random_coords, nbr_table  # noqa: E402

ONES = np.ones(5, dtype=np.float32)


def ref64(x, nbr, w, n_out, child=False):
    xx, ww = x.double(), w.double()
    out = torch.zeros((n_out, w.shape[2]), dtype=torch.float64)
    if not child:
        for k in range(w.shape[0]):
            r = nbr[k].long()
            m = r >= 0
            out[m] += xx[r[m]] @ ww[k]
        return out
    n = nbr.shape[1]
    for c in range(8):
        for d in range(27):
            dz, dy, dx = d // 9 - 1, (d // 3) % 3 - 1, d % 3 - 1
            pz = (((c >> 2) & 1) + dz + 2) // 2 - 1
            py = (((c >> 1) & 1) + dy + 2) // 2 - 1
            px = ((c & 1) + dx + 2) // 2 - 1
            e = (pz + 1) * 9 + (py + 1) * 3 + (px + 1)
            r = nbr[e].long()
            m = r >= 0
            rows = torch.arange(n)[m] * 8 + c
            out[rows] += xx[r[m]] @ ww[d]
    return out


def shrink(tag, got, want):
    got = got.cpu().double()
    big = want.abs() > 0.3 * want.abs().mean()
    rel = ((got - want) * torch.sign(want) / want.abs())[big]
    print('%-34s signed rel err toward +inf of |y|: mean %+.3e  std %.3e  max|rel| %.3e' %
          (tag, float(rel.mean()), float(rel.std()), float(rel.abs().max())))


def kernel_bias():
    rng = np.random.default_rng(1)
    c = random_coords(rng, 4, (24, 24, 24), 0.5)
    n = c.shape[0]
    nbr = torch.from_numpy(nbr_table(c))
    for cin in (16, 48):
        x = torch.from_numpy(np.abs(rng.standard_normal((n, cin))).astype(np.float32))
        w = torch.from_numpy((np.abs(rng.standard_normal((27, cin, 16))) * 0.1).astype(np.float32))   # all positive: no cancellation
        want = ref64(x, nbr, w, n)
        for tc in (False, True):
            out = torch.empty((n, 16), device='cuda')
            E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, out, tc32=tc)
            shrink('cin=%d positive data %s' % (cin, 'tc32' if tc else 'ffma'), out, want)
        w = torch.from_numpy((rng.standard_normal((27, cin, 16)) * 0.1).astype(np.float32))
        want = ref64(x, nbr, w, n)
        for tc in (False, True):
            out = torch.empty((n, 16), device='cuda')
            E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, out, tc32=tc)
            shrink('cin=%d signed weights %s' % (cin, 'tc32' if tc else 'ffma'), out, want)
    m = 4000
    x = torch.from_numpy(np.abs(rng.standard_normal((n, 48))).astype(np.float32))
    w = torch.from_numpy((np.abs(rng.standard_normal((27, 48, 16))) * 0.05).astype(np.float32))
    sub = nbr[:, :m].contiguous()
    want = ref64(x, sub, w, 8 * m, child=True)
    for tc in (False, True):
        out = torch.empty((8 * m, 16), device='cuda')
        E.conv(x.cuda(), sub.cuda(), w.cuda(), 8 * m, out, child_mode=True, tc32=tc)
        shrink('child positive data %s' % ('tc32' if tc else 'ffma'), out, want)


def model(dims, seed, mode):
    m = sgnn_b200.GenModel(8, list(dims), 1, 16, 16, 4, True, True, 1, 1)
    fill_parameters(m, seed)
    m.conv_mode = mode
    from sgnn_b200._lib import lib
    lib.sgnn_debug_set_tc32_min_rows(0)
    return m.cuda().eval()


def levels_dev(tag, out, g):
    (ol, os_), lv = out
    s = []
    for i, l in enumerate(lv):
        if isinstance(l[1], list):
            s.append('L%d: -' % i)
            continue
        same = l[0].shape[0] == g['cand%d_locs' % i].shape[0]
        s.append('L%d: %.2e' % (i, float(np.abs(l[1].cpu().numpy() - g['cand%d' % i]).max())) if same else 'L%d: shape!' % i)
    same = ol.shape[0] == g['out_locs'].shape[0]
    s.append('sdf: %.2e' % float(np.abs(os_.cpu().numpy() - g['out_sdf']).max()) if same else 'sdf: shape!')
    print('%-30s %s' % (tag, '  '.join(s)))


def fixtures():
    orig_conv = E.conv
    for name in ('b2_s32', 'ragged', 'b1_s64'):
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'sgnn_ref_%s.npz' % name))
        locs = torch.from_numpy(g['in_locs'].astype(np.int64))
        feats = torch.from_numpy(g['in_feats']).cuda()
        print('fixture', name, 'margin', float(g['margin']))
        me = model(g['dims'], int(g['param_seed']), 'exact')
        mt = model(g['dims'], int(g['param_seed']), 'tc32')
        a = me([locs, feats], ONES)
        b = mt([locs, feats], ONES)
        levels_dev('exact  vs golden', a, g)
        levels_dev('tc32   vs golden', b, g)
        for i, (x, y) in enumerate(zip(a[1], b[1])):
            if not isinstance(x[1], list) and x[1].shape == y[1].shape:
                print('   tc32 vs exact L%d: %.2e' % (i, float((x[1] - y[1]).abs().max())))
        for which in ('regular', 'child', 'all'):
            def patched(x, nbr, weight, n_out, out_a, child_mode=False, **kw):
                K, cin, cout = weight.shape
                use = cout == 16 and 12 <= cin <= 48 and x.dtype == torch.float32 and \
                    ((child_mode and which in ('child', 'all')) or (not child_mode and which in ('regular', 'all')))
                if use:
                    try:
                        return orig_conv(x, nbr, weight, n_out, out_a, child_mode=child_mode, tc32=True, **kw)
                    except RuntimeError:
                        pass
                return orig_conv(x, nbr, weight, n_out, out_a, child_mode=child_mode, **kw)
            E.conv = patched
            try:
                levels_dev('fused tc32[%s] vs golden' % which, me.forward_fused([locs, feats], ONES), g)
            finally:
                E.conv = orig_conv


if __name__ == '__main__':
    kernel_bias()
    fixtures()
