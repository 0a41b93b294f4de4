"""-m gpu: the fp32 tensor-core convolution path (csrc/conv_tc32.cu: tcgen05 on an exact 3-way bf16 split of fp32
features and filters).  It has fp32 ACCURACY but not the fixed fmaf order of the FFMA kernels, so the comparison with
oracle O3 is by tolerance: |got - want| <= 4e-6 * sum_k |x||w| (a few fp32 ulps of the accumulated magnitude; the
tensor core adds the six exact partial products in its own order).  The whole generator in conv_mode='tc32' is checked
against the golden fixtures of the unmodified reference model.py: occupancy coordinates exactly equal, TSDF <= 1e-3."""
import os

import numpy as np
import pytest
import torch

import o3
from conftest import GOLDEN
from helpers import random_coords, nbr_table, coarse_sets

pytestmark = pytest.mark.gpu
ONES = np.ones(5, dtype=np.float32)


def _E():
    import sgnn_b200.engine as E
    return E


def _bound(x, nbr, w, n_out, child_mode=False):
    """sum over the rules of |x| @ |w|: the magnitude the rounding errors scale with."""
    return o3.conv(x.abs(), nbr, w.abs(), n_out, child_mode=child_mode)


def _check(got, want, bound, rel=4e-6):
    err = (got.cpu() - want).abs()
    tol = rel * bound + 1e-7
    assert bool((err <= tol).all()), 'max err / tol = %.2f (max err %.3e)' % (float((err / tol).max()), float(err.max()))


@pytest.mark.parametrize('cin', [16, 12, 26, 30, 34, 48])
@pytest.mark.parametrize('nb,dims,occ', [(2, (12, 10, 14), 0.35), (1, (5, 5, 5), 0.9), (3, (20, 18, 22), 0.12)])
def test_tc32_submanifold(cin, nb, dims, occ):
    E = _E()
    rng = np.random.default_rng(cin * 7 + dims[0])
    c = random_coords(rng, nb, dims, occ)
    n = c.shape[0]
    ld = (cin + 3) // 4 * 4
    xbuf = torch.zeros((n, ld), dtype=torch.float32)
    xbuf[:, :cin] = torch.from_numpy((rng.standard_normal((n, cin)) * np.exp(rng.uniform(-3, 3, (n, 1)))).astype(np.float32))
    if ld > cin:
        xbuf[:, cin:] = float('nan')                          # padding must never reach the sum
    w = torch.from_numpy((rng.standard_normal((27, cin, 16)) * 0.1).astype(np.float32))
    nbr = torch.from_numpy(nbr_table(c))
    x = xbuf[:, :cin]
    want = o3.conv(x, nbr, w, n)
    out = torch.full((n, 16), float('nan'), device='cuda')
    E.conv(xbuf.cuda()[:, :cin], nbr.cuda(), w.cuda(), n, out, tc32=True)
    _check(out, want, _bound(x, nbr, w, n))


def test_tc32_epilogues_residual_dual_slot_views():
    E = _E()
    rng = np.random.default_rng(5)
    c = random_coords(rng, 2, (12, 10, 14), 0.35)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    r = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.1).astype(np.float32))
    sa, ta = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    sb, tb = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    nbr = torch.from_numpy(nbr_table(c))
    wa = o3.conv(x, nbr, w, n, residual=r, scale=sa, shift=ta, relu=True)
    wb = o3.conv(x, nbr, w, n, residual=r, scale=sb, shift=tb, relu=False)
    bound = _bound(x, nbr, w, n) + r.abs()
    wide = torch.full((n, 48), -7.0, device='cuda')           # JoinTable slot: columns 16..31 of a 48-wide row
    ob = torch.empty((n, 16), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, wide[:, 16:32], residual=r.cuda(), scale_a=sa.cuda(),
           shift_a=ta.cuda(), relu_a=True, out_b=ob, scale_b=sb.cuda(), shift_b=tb.cuda(), relu_b=False, tc32=True)
    _check(wide[:, 16:32], wa, bound * sa.abs() + 1e-6)
    _check(ob, wb, bound * sb.abs() + 1e-6)
    assert (wide[:, :16] == -7).all() and (wide[:, 32:] == -7).all()


@pytest.mark.parametrize('dims', [(12, 10, 14), (9, 7, 5)])
def test_tc32_strided(dims):
    E = _E()
    rng = np.random.default_rng(3)
    c = random_coords(rng, 2, dims, 0.4)
    n = c.shape[0]
    cc, parent, children, cd = coarse_sets(c, dims)
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.2).astype(np.float32))
    ch = torch.from_numpy(children)
    want = o3.conv(x, ch, w, cc.shape[0])
    out = torch.empty((cc.shape[0], 16), device='cuda')
    s, t = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    ob = torch.empty((cc.shape[0], 16), device='cuda')
    E.conv(x.cuda(), ch.cuda(), w.cuda(), cc.shape[0], out, out_b=ob, scale_b=s.cuda(), shift_b=t.cuda(), relu_b=True,
           tc32=True)
    b = _bound(x, ch, w, cc.shape[0])
    _check(out, want, b)
    _check(ob, o3.conv(x, ch, w, cc.shape[0], scale=s, shift=t, relu=True), b * s.abs() + 1e-6)


@pytest.mark.parametrize('dims,occ', [((7, 6, 9), 0.4), ((16, 16, 16), 0.15), ((3, 3, 3), 1.0)])
def test_tc32_child_mode(dims, occ):
    """Generative upsampling: the 27 offsets of each child collapse onto 8 parent neighbours with pre-summed filters."""
    E = _E()
    rng = np.random.default_rng(9)
    c = random_coords(rng, 2, dims, occ)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 48, 16)) * 0.05).astype(np.float32))
    s, t = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(x, nbr, w, 8 * n, child_mode=True, scale=s, shift=t, relu=True)
    out = torch.full((8 * n, 16), float('nan'), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), 8 * n, out, child_mode=True, scale_a=s.cuda(), shift_a=t.cuda(), relu_a=True,
           tc32=True)
    _check(out, want, _bound(x, nbr, w, 8 * n, child_mode=True) * s.abs() + 1e-6)


def test_tc32_large_multi_tile_persistent():
    """More tiles than resident CTAs: exercises the persistent loop, the register prefetch and the TMEM reuse."""
    E = _E()
    rng = np.random.default_rng(11)
    c = random_coords(rng, 40, (32, 32, 32), 0.2)
    n = c.shape[0]
    assert n > 148 * 8 * 128
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.1).astype(np.float32))
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(x, nbr, w, n)
    out = torch.empty((n, 16), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, out, tc32=True)
    _check(out, want, _bound(x, nbr, w, n))
    # child mode on a subset large enough for several tiles per CTA
    m = 148 * 3 * 128 * 2 + 77
    x48 = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    w48 = torch.from_numpy((rng.standard_normal((27, 48, 16)) * 0.05).astype(np.float32))
    nb_sub = nbr[:, :m].contiguous()
    wantc = o3.conv(x48, nb_sub, w48, 8 * m, child_mode=True)
    outc = torch.empty((8 * m, 16), device='cuda')
    E.conv(x48.cuda(), nb_sub.cuda(), w48.cuda(), 8 * m, outc, child_mode=True, tc32=True)
    _check(outc, wantc, _bound(x48, nb_sub, w48, 8 * m, child_mode=True))


def test_tc32_rejects_unsupported_shapes():
    E = _E()
    from sgnn_b200._lib import SgnnError
    x = torch.zeros((4, 8), device='cuda')
    nbr = torch.full((27, 4), -1, dtype=torch.int32, device='cuda')
    with pytest.raises(SgnnError):
        E.conv(x, nbr, torch.zeros((27, 8, 8), device='cuda'), 4, torch.empty((4, 8), device='cuda'), tc32=True)


def _model(dims, seed, mode):
    import sgnn_b200
    from sgnn_b200.synth import fill_parameters
    m = sgnn_b200.GenModel(8, list(dims), 1, 16, 16, 4, True, True, 1, 1)
    fill_parameters(m, seed)
    m.conv_mode = mode
    m.tc32_min_rows = m.ur_min_rows = 1          # tests: every eligible layer on the tensor cores, every site set planned
    return m.cuda().eval()


@pytest.mark.parametrize('name', ['b2_s32', 'ragged', 'b1_s64'])
def test_tc32_generator_against_golden_fixture(name):
    g = np.load(os.path.join(GOLDEN, 'sgnn_ref_%s.npz' % name))
    m = _model(g['dims'], int(g['param_seed']), 'tc32')
    locs = torch.from_numpy(g['in_locs'].astype(np.int64))
    feats = torch.from_numpy(g['in_feats']).cuda()
    (out_locs, out_sdf), levels = m([locs, feats], ONES)
    assert float(g['margin']) > 2e-5          # no oracle logit within 2e-5 of the threshold: coordinates must be equal
    for i, l in enumerate(levels):
        assert np.array_equal(l[0].cpu().numpy(), g['cand%d_locs' % i].astype(np.int64)), 'level %d candidates' % i
        assert np.abs(l[1].cpu().numpy() - g['cand%d' % i]).max() <= 1e-4
        assert int((torch.sigmoid(l[1][:, 0]) > 0.5).sum()) == int(g['kept'][i])
    assert np.array_equal(out_locs.cpu().numpy(), g['out_locs'].astype(np.int64))
    assert np.abs(out_sdf.cpu().numpy() - g['out_sdf']).max() <= 1e-3


def test_tc32_generator_vs_exact_mode_baseline_config():
    """BASELINE.json configs[1] (32 x 64^3 @5 %): tensor-core mode against the bit-reproducible FFMA mode of the same
    engine.  Candidate sets can only differ where an exact-mode logit sits within 1e-5 of the threshold."""
    from sgnn_b200.synth import synthetic_batch
    locs, feats = synthetic_batch(32, 64, 0.05)
    a = _model((64, 64, 64), 0, 'exact')([locs.cuda(), feats.cuda()], ONES)
    b = _model((64, 64, 64), 0, 'tc32')([locs.cuda(), feats.cuda()], ONES)
    from helpers import compare_generator_outputs
    compare_generator_outputs(a, b, margin=1e-5, tol_logit=1e-4, tol_sdf=1e-3, tag='configs[1] tc32 vs exact')
