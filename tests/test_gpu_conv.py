"""-m gpu: floating-point kernels through the C ABI.  fp32 results are compared BIT-EXACT with oracle O3
(fixed fmaf order) and to 1e-4 with the dense-conv identity O1."""
import numpy as np
import pytest
import torch

import o3
import dense_equiv as o1
from helpers import random_coords, nbr_table, coarse_sets

pytestmark = pytest.mark.gpu

PAIRS = [(1, 8), (8, 8), (8, 12), (12, 12), (12, 16), (16, 16), (34, 16), (30, 16), (26, 16), (48, 16), (18, 4)]


def _E():
    import sgnn_b200.engine as E
    return E


def _site_case(seed, nb=2, dims=(12, 10, 14), occ=0.35):
    rng = np.random.default_rng(seed)
    c = random_coords(rng, nb, dims, occ)
    return rng, c, dims, nb


@pytest.mark.parametrize('cin,cout', PAIRS)
@pytest.mark.parametrize('pad', [0, 1])
def test_submanifold_conv_bit_exact(cin, cout, pad):
    E = _E()
    rng, c, dims, nb = _site_case(cin * 31 + cout)
    n = c.shape[0]
    ld = (cin + 3) // 4 * 4 if pad else cin                   # padded rows -> 16-byte gather path
    xbuf = torch.zeros((n, ld), dtype=torch.float32)
    xbuf[:, :cin] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    if ld > cin:
        xbuf[:, cin:] = float('nan')                          # padding must never be read into the sum
    w = torch.from_numpy((rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32))
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(xbuf[:, :cin], nbr, w, n)
    xd = xbuf.cuda()
    out = torch.empty((n, cout), dtype=torch.float32, device='cuda')
    E.conv(xd[:, :cin], nbr.cuda(), w.cuda(), n, out)
    assert torch.equal(out.cpu(), want)
    dense = o1.submanifold_conv(torch.from_numpy(c), xbuf[:, :cin].contiguous(), w, nb, dims)
    assert torch.allclose(out.cpu(), dense, atol=1e-4, rtol=1e-4)


def test_conv_epilogues_residual_dual_slot_views():
    E = _E()
    rng, c, dims, nb = _site_case(5)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    r = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.1).astype(np.float32))
    sa, ta = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    sb, tb = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    nbr = torch.from_numpy(nbr_table(c))
    wa = o3.conv(x, nbr, w, n, residual=r, scale=sa, shift=ta, relu=True)
    wb = o3.conv(x, nbr, w, n, residual=r, scale=sb, shift=tb, relu=False)
    wide = torch.full((n, 48), -7.0, device='cuda')           # JoinTable slot: columns 16..31 of a 48-wide row
    ob = torch.empty((n, 16), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, wide[:, 16:32], residual=r.cuda(), scale_a=sa.cuda(),
           shift_a=ta.cuda(), relu_a=True, out_b=ob, scale_b=sb.cuda(), shift_b=tb.cuda(), relu_b=False)
    assert torch.equal(wide[:, 16:32].cpu(), wa) and torch.equal(ob.cpu(), wb)
    assert (wide[:, :16] == -7).all() and (wide[:, 32:] == -7).all()
    raw = torch.empty((n, 16), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, raw)
    assert torch.equal(raw.cpu(), o3.conv(x, nbr, w, n))


@pytest.mark.parametrize('c_', [8, 12, 16])
def test_strided_conv_deconv_unpool(c_):
    E = _E()
    rng, c, dims, nb = _site_case(c_, dims=(12, 10, 14))
    n = c.shape[0]
    cc, parent, children, cd = coarse_sets(c, dims)
    x = torch.from_numpy(rng.standard_normal((n, c_)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((8, c_, c_)) * 0.2).astype(np.float32))
    want = o3.conv(x, torch.from_numpy(children), w, cc.shape[0])
    out = torch.empty((cc.shape[0], c_), device='cuda')
    E.conv(x.cuda(), torch.from_numpy(children).cuda(), w.cuda(), cc.shape[0], out)
    assert torch.equal(out.cpu(), want)
    assert torch.allclose(want, o1.strided_conv(torch.from_numpy(c), x, w, nb, dims, torch.from_numpy(cc)), atol=1e-4)
    # deconvolution back to the fine set (BASELINE.json configs[2] op pair)
    wd = torch.from_numpy((rng.standard_normal((8, c_, c_)) * 0.2).astype(np.float32))
    dz = torch.empty((n, c_), device='cuda')
    E.deconv(out, torch.from_numpy(parent).cuda(), wd.cuda(), dz)
    assert torch.equal(dz.cpu(), o3.deconv(want, torch.from_numpy(parent), wd))
    assert torch.allclose(dz.cpu(), o1.strided_deconv(torch.from_numpy(cc), want, wd, nb, cd, torch.from_numpy(c)),
                          atol=1e-4)
    up = torch.empty((n, c_), device='cuda')
    E.unpool(out, torch.from_numpy(parent).cuda(), up)
    assert torch.equal(up.cpu(), o1.unpool(torch.from_numpy(cc), want, nb, cd, torch.from_numpy(c)))
    s, t = torch.rand(c_) + 0.5, torch.rand(c_) - 0.5
    E.unpool(out, torch.from_numpy(parent).cuda(), up, scale=s.cuda(), shift=t.cuda(), relu=True)
    assert torch.equal(up.cpu(), o3.affine_relu(o1.unpool(torch.from_numpy(cc), want, nb, cd, torch.from_numpy(c)), s, t))


def test_child_mode_conv_bit_exact():
    E = _E()
    rng, c, dims, nb = _site_case(9, dims=(7, 6, 9), occ=0.4)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 48, 16)) * 0.05).astype(np.float32))
    s, t = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    nbr = torch.from_numpy(nbr_table(c))
    want = o3.conv(x, nbr, w, 8 * n, child_mode=True, scale=s, shift=t, relu=True)
    out = torch.empty((8 * n, 16), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), 8 * n, out, child_mode=True, scale_a=s.cuda(), shift_a=t.cuda(), relu_a=True)
    assert torch.equal(out.cpu(), want)


def test_generic_cout_and_empty():
    E = _E()
    rng, c, dims, nb = _site_case(3)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 5)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, 5, 7)) * 0.1).astype(np.float32))
    nbr = torch.from_numpy(nbr_table(c))
    out = torch.empty((n, 7), device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w.cuda(), n, out)
    assert torch.equal(out.cpu(), o3.conv(x, nbr, w, n))
    E.conv(x.cuda()[:0], nbr.cuda()[:, :0], w.cuda(), 0, out[:0])          # n_out == 0 is a no-op


def test_pointwise_kernels():
    E = _E()
    rng = np.random.default_rng(0)
    n = 1000
    x = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    s, t = torch.rand(48) + 0.5, torch.rand(48) - 0.5
    y = torch.empty((n, 48), device='cuda')
    E.affine_relu(x.cuda(), y, s.cuda(), t.cuda())
    assert torch.equal(y.cpu(), o3.affine_relu(x, s, t))
    b = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    E.add_rows(x.cuda(), b.cuda(), y)
    assert torch.equal(y.cpu(), x + b)
    wide = torch.zeros((n, 64), device='cuda')
    E.copy_cols(x.cuda(), wide[:, 8:56])
    assert torch.equal(wide[:, 8:56].cpu(), x) and float(wide[:, :8].abs().sum()) == 0
    lin = torch.nn.Linear(48, 1)
    o = torch.empty((n, 1), device='cuda')
    E.linear(x.cuda(), lin.weight.detach().cuda(), lin.bias.detach().cuda(), o)
    assert torch.equal(o.cpu(), o3.linear(x, lin.weight.detach(), lin.bias.detach()))


def test_sparse_to_dense_and_back():
    E = _E()
    rng, c, dims, nb = _site_case(4, nb=3, dims=(8, 8, 8), occ=0.3)
    n = c.shape[0]
    f = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    d = E.sparse_to_dense(f.cuda(), torch.from_numpy(c.astype(np.int32)).cuda(), nb, list(dims))
    assert torch.equal(d.cpu(), o1.densify(torch.from_numpy(c), f, nb, dims))
    occsdf = torch.from_numpy(rng.standard_normal((nb, 2, 8, 8, 8)).astype(np.float32))
    occsdf[0, 0, 0, 0, :4] = torch.tensor([0.0, 5e-8, 1.3e-7, -1e-9])      # threshold edge
    locs, feats, cand, m = E.dense_to_sparse(d, occsdf.cuda(), ld_feats=36)
    keep = o3.sigmoid_gt_half(occsdf[:, 0].reshape(-1))
    assert torch.equal(keep, torch.sigmoid(occsdf[:, 0].reshape(-1)) > 0.5)
    assert m == int(keep.sum())
    zz, yy, xx = torch.meshgrid(torch.arange(8), torch.arange(8), torch.arange(8), indexing='ij')
    cell = torch.stack([zz, yy, xx], -1).view(-1, 3)
    all_locs = torch.cat([cell.repeat(nb, 1), torch.arange(nb).repeat_interleave(512).view(-1, 1)], 1)
    assert torch.equal(locs.cpu().long(), all_locs[keep])
    allf = torch.cat([occsdf.permute(0, 2, 3, 4, 1).reshape(-1, 2), d.cpu().permute(0, 2, 3, 4, 1).reshape(-1, 16)], 1)
    assert torch.equal(feats[:, :18].cpu(), allf[keep]) and float(feats[:, 18:].abs().sum()) == 0
    assert torch.equal(cand.cpu(), allf[:, :2])


def test_heads_compact_and_children_and_skip():
    E = _E()
    rng, c, dims, nb = _site_case(6, dims=(6, 6, 6), occ=0.4)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((8 * n, 16)).astype(np.float32))
    lo, ls = torch.nn.Linear(16, 1), torch.nn.Linear(16, 1)
    pc = torch.from_numpy(c.astype(np.int32)).cuda()
    locs, feats, cand, m = E.heads_compact(x.cuda(), lo.weight.detach().view(-1).cuda(), lo.bias.detach().cuda(),
                                           ls.weight.detach().view(-1).cuda(), ls.bias.detach().cuda(), pc, ld_feats=36)
    occ = o3.linear(x, lo.weight.detach(), lo.bias.detach())
    sdf = o3.linear(x, ls.weight.detach(), ls.bias.detach())
    assert torch.equal(cand.cpu(), torch.cat([occ, sdf], 1))
    keep = o3.sigmoid_gt_half(occ.view(-1))
    kids = (c[:, None, :] * np.array([2, 2, 2, 1]) +
            np.array([[z, y, x_, 0] for z in (0, 1) for y in (0, 1) for x_ in (0, 1)])[None]).reshape(-1, 4)
    assert np.array_equal(E.children_coords(pc).cpu().numpy(), kids)
    assert m == int(keep.sum())
    assert np.array_equal(locs.cpu().numpy(), kids[keep.numpy()])
    assert torch.equal(feats[:, :18].cpu(), torch.cat([x, occ, sdf], 1)[keep])
    # concat_skip: join an "encoder" set onto the kept children by coordinate
    enc_c = random_coords(np.random.default_rng(8), nb, (12, 12, 12), 0.5)
    enc_f = torch.from_numpy(rng.standard_normal((enc_c.shape[0], 12)).astype(np.float32))
    g = E.build_grid(torch.from_numpy(enc_c).cuda(), nb, (12, 12, 12))
    E.concat_skip(g, enc_f.cuda(), locs, feats, 18)
    import sparseconvnet as o2
    rows = o2._SiteSet(torch.from_numpy(enc_c)).lookup(kids[keep.numpy()])
    want = torch.where(torch.from_numpy(rows >= 0).view(-1, 1), enc_f[np.maximum(rows, 0)], torch.zeros(1))
    assert torch.equal(feats[:, 18:30].cpu(), want)


def test_error_codes():
    E = _E()
    from sgnn_b200._lib import SgnnError
    x = torch.zeros((10, 16), device='cuda')
    nbr = torch.full((27, 10), -1, dtype=torch.int32, device='cuda')
    w = torch.zeros((27, 16, 16), device='cuda')
    bad = torch.zeros((10, 17), device='cuda')[:, 1:]                # misaligned output rows
    with pytest.raises(SgnnError, match='alignment'):
        E.conv(x, nbr, w, 10, bad)
    with pytest.raises(SgnnError, match='unsupported'):
        E.conv(x, nbr[:5], w[:5].contiguous(), 10, torch.empty((10, 16), device='cuda'))
    with pytest.raises(RuntimeError):
        E.conv(x.cpu(), nbr, w, 10, torch.empty((10, 16), device='cuda'))


@pytest.mark.parametrize('transposed,c0,c1,cout,k,s,p,dims', [
    (False, 16, 0, 24, 4, 2, 1, (8, 8, 8)), (False, 24, 0, 32, 4, 2, 1, (4, 4, 4)), (False, 32, 0, 32, 1, 1, 0, (2, 2, 2)),
    (True, 32, 32, 32, 4, 2, 1, (2, 2, 2)), (True, 32, 24, 28, 4, 2, 1, (4, 6, 4)), (False, 28, 0, 16, 1, 1, 0, (8, 4, 8)),
    (False, 16, 0, 2, 1, 1, 0, (8, 8, 8)), (True, 8, 0, 4, 3, 1, 1, (4, 4, 4)), (True, 6, 2, 5, 2, 2, 0, (3, 3, 3)),
    (True, 5, 0, 3, 4, 1, 0, (3, 2, 3)), (False, 7, 3, 5, 3, 2, 1, (7, 5, 6)),
])
def test_dense_unet_layers_bit_exact(transposed, c0, c1, cout, k, s, p, dims):
    """a12: the coarse dense U-Net layers (model.py:89-136) -- bit-exact vs O3, 1e-5 vs torch."""
    E = _E()
    torch.manual_seed(c0 + cout)
    nb = 3
    x0 = torch.randn(nb, c0, *dims)
    x1 = torch.randn(nb, c1, *dims) if c1 else None
    cin = c0 + c1
    w = torch.randn((cin, cout, k, k, k) if transposed else (cout, cin, k, k, k)) * 0.05
    sc, sh = torch.rand(cout) + 0.5, torch.rand(cout) - 0.5
    xcat = torch.cat([x0, x1], 1) if c1 else x0
    want = o3.dense_conv(xcat, w, k, s, p, sc, sh, True, transposed)
    got = E.dense_conv(x0.cuda(), x1.cuda() if c1 else None, w.cuda(), cout, k, s, p, sc.cuda(), sh.cuda(), True,
                       transposed)
    assert torch.equal(got.cpu(), want)
    fn = torch.nn.functional.conv_transpose3d if transposed else torch.nn.functional.conv3d
    ref = torch.relu(fn(xcat, w, stride=s, padding=p) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))
    assert torch.allclose(got.cpu(), ref, atol=2e-5, rtol=1e-5)
    raw = E.dense_conv(x0.cuda(), x1.cuda() if c1 else None, w.cuda(), cout, k, s, p, transposed=transposed)
    assert torch.equal(raw.cpu(), o3.dense_conv(xcat, w, k, s, p, transposed=transposed))


def test_dispatched_conv_kernels_bit_identical_to_o3():
    """Whatever kernel the dispatcher picks (row-owner for the SG-NN channel plan, child-mode kernel, runtime-shape kernel for
    other channel counts), the fmaf chains are those of O3: bit-identical results."""
    E = _E()
    rng, c, dims, nb = _site_case(77, dims=(9, 8, 11), occ=0.45)
    n = c.shape[0]
    nbr = torch.from_numpy(nbr_table(c))
    if True:
        for cin, cout, child in [(16, 16, False), (34, 16, False), (26, 16, False), (48, 16, True), (8, 12, False),
                                 (7, 8, False), (20, 16, False), (5, 4, False)]:
            ld = (cin + 3) // 4 * 4
            xb = torch.zeros((n, ld))
            xb[:, :cin] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
            w = torch.from_numpy((rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32))
            r = torch.from_numpy(rng.standard_normal(((8 if child else 1) * n, cout)).astype(np.float32))
            s_, t_ = torch.rand(cout) + 0.5, torch.rand(cout) - 0.5
            no = 8 * n if child else n
            want = o3.conv(xb[:, :cin], nbr, w, no, child_mode=child, residual=r, scale=s_, shift=t_, relu=True)
            out = torch.empty((no, cout), device='cuda')
            E.conv(xb.cuda()[:, :cin], nbr.cuda(), w.cuda(), no, out, child_mode=child, residual=r.cuda(),
                   scale_a=s_.cuda(), shift_a=t_.cuda(), relu_a=True)
            assert torch.equal(out.cpu(), want), (cin, cout, child)


# ------------------------------------------------------------------ compact rulebook + conv_sp.cu (encoder input level)
def _decode_compact(slots, cnt):
    """(slots [27,n], cnt [n]) -> dense [27,n] table with -1 for absent; asserts ascending k within a row."""
    slots, cnt = slots.cpu().numpy().astype(np.uint32), cnt.cpu().numpy()
    n = cnt.shape[0]
    dense = np.full((27, n), -1, dtype=np.int32)
    last_k = np.full(n, -1, dtype=np.int64)
    for s in range(int(cnt.max()) if n else 0):
        live = np.nonzero(cnt > s)[0]
        e = slots[s, live]
        k, r = (e >> 27).astype(np.int64), (e & 0x7ffffff).astype(np.int32)
        assert (k > last_k[live]).all(), 'offsets of a row must ascend'
        last_k[live] = k
        dense[k, live] = r
    return dense


@pytest.mark.parametrize('nb,dims,occ', [(2, (12, 10, 14), 0.35), (3, (20, 70, 130), 0.05), (1, (9, 9, 64), 0.9)])
def test_compact_rulebook_equals_dense_table(nb, dims, occ):
    E = _E()
    rng = np.random.default_rng(nb * 7 + dims[2])
    c = random_coords(rng, nb, dims, occ)
    c = c[rng.permutation(c.shape[0])]                        # caller order (row_of_rank in play), x on word edges for 130
    g = E.build_grid(torch.from_numpy(c).cuda(), nb, dims)
    nbr = E.rulebook_submanifold(g)
    slots, cnt = E.rulebook_submanifold_compact(g)
    assert np.array_equal(nbr.cpu().numpy(), nbr_table(c))
    assert np.array_equal(_decode_compact(slots, cnt), nbr.cpu().numpy())
    assert np.array_equal(cnt.cpu().numpy(), (nbr.cpu().numpy() >= 0).sum(0).astype(np.uint8))


@pytest.mark.parametrize('cin,cout', [(1, 8), (8, 8), (8, 12), (12, 12), (12, 16), (16, 16)])
@pytest.mark.parametrize('occ', [0.05, 0.5])
def test_compact_conv_bit_identical_to_dense_table_conv_and_oracle(cin, cout, occ):
    E = _E()
    rng = np.random.default_rng(cin * 17 + cout)
    nb, dims = 2, (14, 12, 70)
    c = random_coords(rng, nb, dims, occ)
    c = c[rng.permutation(c.shape[0])]
    n = c.shape[0]
    ld = (cin + 3) // 4 * 4
    xbuf = torch.full((n, ld), float('nan'))
    xbuf[:, :cin] = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32))
    r = torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32))
    sa, ta = torch.rand(cout) + 0.5, torch.rand(cout) - 0.5
    g = E.build_grid(torch.from_numpy(c).cuda(), nb, dims)
    nbr = E.rulebook_submanifold(g)
    comp = E.rulebook_submanifold_compact(g)
    xd = xbuf.cuda()
    # plain
    a = torch.empty((n, cout), device='cuda')
    b = torch.empty((n, cout), device='cuda')
    E.conv(xd[:, :cin], nbr, w.cuda(), n, a)
    E.conv(xd[:, :cin], None, w.cuda(), n, b, compact=comp)
    assert torch.equal(a, b) and torch.equal(b.cpu(), o3.conv(xbuf[:, :cin], nbr.cpu(), w, n))
    # residual + two epilogue slots, one of them a column view of a wider buffer
    wide = torch.full((n, 48), -7.0, device='cuda')
    raw = torch.empty((n, cout), device='cuda')
    E.conv(xd[:, :cin], None, w.cuda(), n, wide[:, 16:16 + cout], residual=r.cuda(), scale_a=sa.cuda(), shift_a=ta.cuda(),
           relu_a=True, out_b=raw, compact=comp)
    assert torch.equal(wide[:, 16:16 + cout].cpu(), o3.conv(xbuf[:, :cin], nbr.cpu(), w, n, residual=r, scale=sa, shift=ta, relu=True))
    assert torch.equal(raw.cpu(), o3.conv(xbuf[:, :cin], nbr.cpu(), w, n, residual=r))
    assert (wide[:, :16] == -7).all() and (wide[:, 16 + cout:] == -7).all()


def test_generator_same_bits_with_compact_and_dense_rules():
    """The native generator with the compact rulebook on the encoder input level (default) and with the dense table (A/B flag):
    identical outputs, both convolution modes."""
    import sgnn_b200
    from sgnn_b200.synth import fill_parameters, synthetic_batch
    m = sgnn_b200.GenModel(8, [32, 32, 32], 1, 16, 16, 4, True, True, 1, 1)
    fill_parameters(m, 0)
    m = m.cuda().eval()
    locs, feats = synthetic_batch(3, [32, 32, 32], 0.08)
    ones = np.ones(5, dtype=np.float32)
    for mode in ('exact', 'tc32'):
        m.conv_mode = mode
        m.dense_rules = False
        (la, sa), lva = m([locs.cuda(), feats.cuda()], ones)
        m.dense_rules = True
        (lb, sb), lvb = m([locs.cuda(), feats.cuda()], ones)
        assert torch.equal(la, lb) and torch.equal(sa, sb)
        for x, y in zip(lva, lvb):
            assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])
        m.dense_rules = False


def test_generator_same_bits_with_and_without_the_side_stream():
    """Small levels build their coarse site sets on a side stream under the level's first convolutions (generator.cu fork_side /
    join_side; GenModel.overlap_max_rows, -1 = never).  Same bits as the single-stream order, repeatedly (a missing
    cross-stream dependency would show as a difference on some repetition)."""
    import sgnn_b200
    from sgnn_b200.synth import fill_parameters, synthetic_batch
    m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
    fill_parameters(m, 0)
    m = m.cuda().eval()
    ones = np.ones(5, dtype=np.float32)
    batches = [synthetic_batch(8, 64, 0.05, first=8 * i) for i in range(3)]
    for mode in ('tc32', 'exact'):
        m.conv_mode = mode
        m.overlap_max_rows = -1
        want = [m([l.cuda(), f.cuda(), 8], ones) for l, f in batches]
        for setting in (0, 1 << 40):                      # default threshold; every level
            m.overlap_max_rows = setting
            for rep in range(3):
                for (l, f), ((wl, ws), wlv) in zip(batches, want):
                    (gl, gs), glv = m([l.cuda(), f.cuda(), 8], ones)
                    assert torch.equal(wl, gl) and torch.equal(ws, gs), (mode, setting, rep)
                    for x, y in zip(wlv, glv):
                        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]), (mode, setting, rep)
