"""-m gpu: the unique-row tensor-core convolution (csrc/conv_ur.cu): tile plan (integer work: exact) and the TMA-staged
tcgen05 convolution against oracle O3 (same tolerance as tests/test_gpu_tc32.py: 4e-6 of sum |x||w|).  Covers the
multi-pass path (more distinct rows in a tile than the staging area holds) and DIRECT mode (tiles the plan cannot
describe: row-id span or distinct count over its caps)."""
import numpy as np
import pytest
import torch

import o3
from helpers import random_coords, nbr_table

pytestmark = pytest.mark.gpu

PLAN_CAP = 512          # UR_PLAN_CAP
SPAN_CAP = 4096 * 32    # UR_BM_WORDS * 32


def _E():
    import sgnn_b200.engine as E
    return E


def _bound(x, nbr, w, n_out):
    return o3.conv(x.abs(), nbr, w.abs(), n_out)


def _check(got, want, bound, rel=4e-6):
    err = (got.cpu() - want).abs()
    tol = rel * bound + 1e-7
    assert bool((err <= tol).all()), 'max err / tol = %.2f (max err %.3e)' % (float((err / tol).max()), float(err.max()))


def _parse_plan(plan, n_rows):
    """White-box view of the opaque plan buffer (layout of conv_ur.cu::plan_view)."""
    al = lambda b: (b + 255) & ~255
    tiles = (n_rows + 127) // 128
    raw = plan.cpu().numpy()
    o = 0
    ucount = raw[o:o + tiles * 4].view(np.int32); o += al(tiles * 4)
    urows = raw[o:o + tiles * PLAN_CAP * 4].view(np.int32).reshape(tiles, PLAN_CAP); o += al(tiles * PLAN_CAP * 4)
    lidx = raw[o:o + tiles * 27 * 128 * 2].view(np.uint16).reshape(tiles, 27, 128)
    return ucount, urows, lidx


def _check_plan(nbr, n, plan):
    ucount, urows, lidx = _parse_plan(plan, n)
    nbr = nbr.numpy() if isinstance(nbr, torch.Tensor) else nbr
    direct = 0
    for t in range((n + 127) // 128):
        blk = nbr[:, t * 128:(t + 1) * 128]
        present = blk[blk >= 0]
        u = np.unique(present)
        span = int(u.max() - u.min() + 1) if u.size else 0
        if span > SPAN_CAP or u.size > PLAN_CAP:
            assert ucount[t] == -1
            direct += 1
            continue
        assert ucount[t] == u.size, (t, ucount[t], u.size)
        assert np.array_equal(urows[t, :u.size], u)                       # sorted, distinct
        want = np.full((27, 128), 0xFFFF, dtype=np.uint16)
        w = blk.shape[1]
        pos = np.searchsorted(u, np.where(blk >= 0, blk, u[0] if u.size else 0))
        want[:, :w] = np.where(blk >= 0, pos, 0xFFFF).astype(np.uint16)
        assert np.array_equal(lidx[t], want)
    return direct


@pytest.mark.parametrize('nb,dims,occ', [(2, (12, 10, 14), 0.35), (1, (5, 5, 5), 0.9), (3, (20, 18, 22), 0.12),
                                         (1, (40, 40, 40), 0.5)])
def test_tile_plan_exact(nb, dims, occ):
    E = _E()
    rng = np.random.default_rng(dims[0])
    c = random_coords(rng, nb, dims, occ)
    n = c.shape[0]
    nbr = torch.from_numpy(nbr_table(c))
    plan = E.tile_plan(nbr.cuda(), n)
    torch.cuda.synchronize()
    _check_plan(nbr, n, plan)


@pytest.mark.parametrize('cin', [16, 12, 26, 30, 32])
@pytest.mark.parametrize('nb,dims,occ', [(2, (12, 10, 14), 0.35), (1, (5, 5, 5), 0.9), (3, (20, 18, 22), 0.12)])
def test_ur_submanifold(cin, nb, dims, occ):
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
    nbr_d = nbr.cuda()
    plan = E.tile_plan(nbr_d, n)
    E.conv(xbuf.cuda()[:, :cin], nbr_d, w.cuda(), n, out, plan=plan)
    _check(out, want, _bound(x, nbr, w, n))


@pytest.mark.parametrize('cin,cout', [(8, 8), (8, 12), (12, 12), (16, 8)])
def test_ur_narrow_outputs(cin, cout):
    """The encoder's 8- and 12-channel layers: accumulator columns >= cout are zero and never stored."""
    E = _E()
    rng = np.random.default_rng(cin * 31 + cout)
    c = random_coords(rng, 3, (20, 18, 22), 0.12)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    r = torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32))
    s, t = torch.rand(cout) + 0.5, torch.rand(cout) - 0.5
    nbr = torch.from_numpy(nbr_table(c))
    want_a = o3.conv(x, nbr, w, n, residual=r)
    want_b = o3.conv(x, nbr, w, n, residual=r, scale=s, shift=t, relu=True)
    wide = torch.full((n, cout + 8), -3.0, device='cuda')          # guard columns around the output slot
    ob = torch.empty((n, cout), device='cuda')
    nbr_d = nbr.cuda()
    E.conv(x.cuda(), nbr_d, w.cuda(), n, wide[:, 4:4 + cout], residual=r.cuda(), out_b=ob, scale_b=s.cuda(), shift_b=t.cuda(),
           relu_b=True, plan=E.tile_plan(nbr_d, n))
    bound = _bound(x, nbr, w, n) + r.abs()
    _check(wide[:, 4:4 + cout], want_a, bound)
    _check(ob, want_b, bound * s.abs() + 1e-6)
    assert (wide[:, :4] == -3).all() and (wide[:, 4 + cout:] == -3).all()


def test_ur_wide_rows_padded_ld():
    """Joined feature rows as the generator lays them out: cin 26 in 32-float rows (128-byte TMA copies), junk in the pad."""
    E = _E()
    rng = np.random.default_rng(2)
    c = random_coords(rng, 2, (16, 16, 16), 0.3)
    n = c.shape[0]
    xbuf = torch.full((n, 32), float('nan'), dtype=torch.float32)
    xbuf[:, :26] = torch.from_numpy(rng.standard_normal((n, 26)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 26, 16)) * 0.1).astype(np.float32))
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(xbuf[:, :26], nbr, w, n)
    out = torch.empty((n, 16), device='cuda')
    nbr_d = nbr.cuda()
    E.conv(xbuf.cuda()[:, :26], nbr_d, w.cuda(), n, out, plan=E.tile_plan(nbr_d, n))
    _check(out, want, _bound(xbuf[:, :26], nbr, w, n))


def test_ur_input_is_column_view_of_wider_rows():
    """x = 16 columns of 48-float rows: the row stride exceeds the copied bytes, so every row is its own bulk copy."""
    E = _E()
    rng = np.random.default_rng(8)
    c = random_coords(rng, 2, (16, 16, 16), 0.3)
    n = c.shape[0]
    wide = torch.full((n, 48), float('nan'), dtype=torch.float32)
    wide[:, 16:32] = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.1).astype(np.float32))
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(wide[:, 16:32], nbr, w, n)
    out = torch.empty((n, 16), device='cuda')
    nbr_d = nbr.cuda()
    E.conv(wide.cuda()[:, 16:32], nbr_d, w.cuda(), n, out, plan=E.tile_plan(nbr_d, n))
    _check(out, want, _bound(wide[:, 16:32], nbr, w, n))


def test_ur_epilogues_residual_dual_slot_views():
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
    wide = torch.full((n, 48), -7.0, device='cuda')
    ob = torch.empty((n, 16), device='cuda')
    nbr_d = nbr.cuda()
    E.conv(x.cuda(), nbr_d, w.cuda(), n, wide[:, 16:32], residual=r.cuda(), scale_a=sa.cuda(), shift_a=ta.cuda(),
           relu_a=True, out_b=ob, scale_b=sb.cuda(), shift_b=tb.cuda(), relu_b=False, plan=E.tile_plan(nbr_d, n))
    _check(wide[:, 16:32], wa, bound * sa.abs() + 1e-6)
    _check(ob, wb, bound * sb.abs() + 1e-6)
    assert (wide[:, :16] == -7).all() and (wide[:, 32:] == -7).all()


@pytest.mark.parametrize('cin', [16, 30])
def test_ur_large_persistent_multi_tile(cin):
    """Many tiles per CTA: the TMA ring, the double-buffered planes / lidx / accumulators and every mbarrier wrap around."""
    E = _E()
    rng = np.random.default_rng(11 + cin)
    c = random_coords(rng, 24, (32, 32, 32), 0.2)
    n = c.shape[0]
    assert n > 148 * 6 * 128
    ld = (cin + 7) // 8 * 8
    xbuf = torch.zeros((n, ld), dtype=torch.float32)
    xbuf[:, :cin] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, cin, 16)) * 0.1).astype(np.float32))
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(xbuf[:, :cin], nbr, w, n)
    out = torch.empty((n, 16), device='cuda')
    nbr_d = nbr.cuda()
    plan = E.tile_plan(nbr_d, n)
    E.conv(xbuf.cuda()[:, :cin], nbr_d, w.cuda(), n, out, plan=plan)
    _check(out, want, _bound(xbuf[:, :cin], nbr, w, n))
    # and again on the same plan (a plan is shared by every convolution of its site set)
    out2 = torch.empty((n, 16), device='cuda')
    E.conv(xbuf.cuda()[:, :cin], nbr_d, w.cuda(), n, out2, plan=plan)
    assert torch.equal(out, out2)


def _shuffled(rng, nb, dims, occ):
    """Site set in a random row order: every tile's neighbours are scattered over the whole row range."""
    c = random_coords(rng, nb, dims, occ)
    return np.ascontiguousarray(c[rng.permutation(c.shape[0])])


@pytest.mark.parametrize('cin', [16, 26])
def test_ur_multi_pass_and_plan_cap(cin):
    """Dense, shuffled rows: ~27 distinct rows per output row -> tiles with 320 < U <= 512 take two passes for 32-channel
    inputs, tiles with U > 512 run in direct mode."""
    E = _E()
    rng = np.random.default_rng(cin)
    c = _shuffled(rng, 1, (14, 14, 14), 0.2)
    n = c.shape[0]
    nbr = torch.from_numpy(nbr_table(c))
    nbr_d = nbr.cuda()
    plan = E.tile_plan(nbr_d, n)
    torch.cuda.synchronize()
    _check_plan(nbr, n, plan)
    ucount, _, _ = _parse_plan(plan, n)
    assert (ucount > 320).any() and (ucount >= 0).all(), ucount
    ld = (cin + 7) // 8 * 8
    xbuf = torch.zeros((n, ld), dtype=torch.float32)
    xbuf[:, :cin] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, cin, 16)) * 0.1).astype(np.float32))
    want = o3.conv(xbuf[:, :cin], nbr, w, n)
    out = torch.empty((n, 16), device='cuda')
    E.conv(xbuf.cuda()[:, :cin], nbr_d, w.cuda(), n, out, plan=plan)
    _check(out, want, _bound(xbuf[:, :cin], nbr, w, n))


@pytest.mark.parametrize('cin', [16, 26])
def test_ur_direct_mode(cin):
    """Shuffled dense rows (U > 512 per tile) and a row-id span over the bitmap cap: every tile in direct mode, mixed with
    planned tiles in a second, locally ordered part of the same site set."""
    E = _E()
    rng = np.random.default_rng(100 + cin)
    a = _shuffled(rng, 1, (16, 16, 16), 0.6)            # ~2400 rows, ~17 present taps: U ~ 128 * 10 > 512
    b = random_coords(rng, 1, (16, 16, 16), 0.3)
    b[:, 3] = 1
    c = np.ascontiguousarray(np.concatenate([a, b]))
    n = c.shape[0]
    nbr = torch.from_numpy(nbr_table(c))
    nbr_d = nbr.cuda()
    plan = E.tile_plan(nbr_d, n)
    torch.cuda.synchronize()
    direct = _check_plan(nbr, n, plan)
    ucount, _, _ = _parse_plan(plan, n)
    assert direct > 0 and (ucount >= 0).any()
    ld = (cin + 7) // 8 * 8
    xbuf = torch.zeros((n, ld), dtype=torch.float32)
    xbuf[:, :cin] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, cin, 16)) * 0.1).astype(np.float32))
    want = o3.conv(xbuf[:, :cin], nbr, w, n)
    out = torch.empty((n, 16), device='cuda')
    E.conv(xbuf.cuda()[:, :cin], nbr_d, w.cuda(), n, out, plan=plan)
    _check(out, want, _bound(xbuf[:, :cin], nbr, w, n))


def test_ur_span_cap_direct():
    """Two far-apart row ranges referenced from one tile (span > 131072 rows) -> direct mode by the span rule."""
    E = _E()
    rng = np.random.default_rng(3)
    n = SPAN_CAP + 4096
    nbr = np.full((27, n), -1, dtype=np.int32)
    nbr[13] = np.arange(n)
    nbr[0, :64] = n - 1 - np.arange(64)                  # first tile reaches to the far end
    nbr[26, 200:300] = np.arange(100)
    nbr_t = torch.from_numpy(nbr)
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.1).astype(np.float32))
    nbr_d = nbr_t.cuda()
    plan = E.tile_plan(nbr_d, n)
    torch.cuda.synchronize()
    ucount, _, _ = _parse_plan(plan, n)
    assert ucount[0] == -1 and (ucount[1:] >= 0).all()
    want = o3.conv(x, nbr_t, w, n)
    out = torch.empty((n, 16), device='cuda')
    E.conv(x.cuda(), nbr_d, w.cuda(), n, out, plan=plan)
    _check(out, want, _bound(x, nbr_t, w, n))


def test_ur_rejects_unsupported_shapes():
    E = _E()
    from sgnn_b200._lib import SgnnError
    nbr = torch.full((27, 4), -1, dtype=torch.int32, device='cuda')
    plan = E.tile_plan(nbr, 4)
    with pytest.raises(SgnnError):
        E.conv(torch.zeros((4, 48), device='cuda'), nbr, torch.zeros((27, 48, 16), device='cuda'), 4,
               torch.empty((4, 16), device='cuda'), plan=plan)
    with pytest.raises(SgnnError):
        E.conv(torch.zeros((4, 16), device='cuda'), nbr, torch.zeros((27, 16, 4), device='cuda'), 4,
               torch.empty((4, 4), device='cuda'), plan=plan)


# ------------------------------------------------------------------------------------------------ child mode (conv_urc.cu)
def _child_case(E, rng, c, scale=True):
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 48, 16)) * 0.05).astype(np.float32))
    s, t = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(x, nbr, w, 8 * n, child_mode=True, scale=s, shift=t, relu=True)
    out = torch.full((8 * n, 16), float('nan'), device='cuda')
    nbr_d = nbr.cuda()
    plan = E.tile_plan(nbr_d, n)
    E.conv(x.cuda(), nbr_d, w.cuda(), 8 * n, out, child_mode=True, scale_a=s.cuda(), shift_a=t.cuda(), relu_a=True, plan=plan)
    bound = o3.conv(x.abs(), nbr, w.abs(), 8 * n, child_mode=True) * s.abs() + 1e-6
    _check(out, want, bound)
    return plan, n


@pytest.mark.parametrize('dims,occ', [((7, 6, 9), 0.4), ((16, 16, 16), 0.15), ((3, 3, 3), 1.0), ((24, 20, 28), 0.3)])
def test_urc_child_mode(dims, occ):
    """Generative upsampling on the unique-row kernel: centre rounds first, children of a round merged into N = 16 x run MMAs."""
    E = _E()
    rng = np.random.default_rng(9 + dims[0])
    _child_case(E, rng, random_coords(rng, 2, dims, occ))


def test_urc_large_persistent_multi_tile():
    """Several tiles per CTA: filter ring, row ring, index double buffer and the single accumulator set all wrap around."""
    E = _E()
    rng = np.random.default_rng(21)
    c = random_coords(rng, 12, (32, 32, 32), 0.12)
    assert c.shape[0] > 148 * 2 * 128
    _child_case(E, rng, c)


def test_urc_multi_pass_and_direct():
    """Shuffled parents: tiles with more than 256 distinct rows take two passes, tiles over the plan cap run in direct mode."""
    E = _E()
    rng = np.random.default_rng(33)
    a = _shuffled(rng, 1, (14, 14, 14), 0.2)               # U ~ 390 per tile: two passes
    b = _shuffled(rng, 1, (16, 16, 16), 0.6)               # U > 512: direct mode
    b[:, 3] = 1
    d = random_coords(rng, 1, (16, 16, 16), 0.3)           # planned, single pass
    d[:, 3] = 2
    c = np.ascontiguousarray(np.concatenate([a, b, d]))
    plan, n = _child_case(E, rng, c)
    ucount, _, _ = _parse_plan(plan, n)
    assert (ucount > 256).any() and (ucount == -1).any() and ((ucount >= 0) & (ucount <= 256)).any()


def test_fused_rulebook_and_plan_equals_the_two_calls():
    """sgnn_rulebook_submanifold_plan: the same neighbour table, and a plan that drives the unique-row convolution to the same
    bits as the plan built from the table by sgnn_tile_plan_build (caller-ordered rows, ragged last tile, x on word edges)."""
    import sgnn_b200.engine as E
    from helpers import random_coords
    rng = np.random.default_rng(9)
    for nb, dims, occ in [(2, (20, 24, 130), 0.3), (1, (33, 17, 64), 0.7)]:
        c = random_coords(rng, nb, dims, occ)
        c = c[rng.permutation(c.shape[0])]
        g = E.build_grid(torch.from_numpy(c).cuda(), nb, dims)
        n = g.n
        nbr = E.rulebook_submanifold(g)
        nbr2, plan2 = E.rulebook_submanifold_plan(g)
        assert torch.equal(nbr, nbr2)
        plan = E.tile_plan(nbr, n).clone()
        x = torch.randn((n, 16), device='cuda')
        w = torch.randn((27, 16, 16), device='cuda') * 0.1
        a = torch.empty((n, 16), device='cuda')
        b = torch.empty((n, 16), device='cuda')
        E.conv(x, nbr, w, n, a, plan=plan)
        E.conv(x, nbr2, w, n, b, plan=plan2)
        assert torch.equal(a, b)
        tiles = (n + 127) // 128
        assert torch.equal(plan[:tiles * 4], plan2[:tiles * 4])                 # distinct-row counts per tile
