"""-m gpu: active-site grid + rulebooks (integer work: bit-exact against oracle O2's site index)."""
import numpy as np
import pytest
import torch

from helpers import random_coords, nbr_table, coarse_sets

pytestmark = pytest.mark.gpu

CASES = [
    # nb, dims, occ, empty, shuffle
    (1, (8, 8, 8), 0.3, (), False),
    (2, (16, 12, 10), 0.15, (), True),
    (3, (8, 16, 8), 0.4, (1,), True),
    (1, (6, 6, 200), 0.2, (), True),        # 4 mask words per x-row, neighbours across word edges
    (2, (4, 4, 4), 1.0, (), False),         # dense, every border
    (2, (9, 7, 65), 0.3, (0,), True),       # odd extents, first sample empty
    (4, (64, 64, 64), 0.05, (), False),     # BASELINE block shape
]


def _E():
    import sgnn_b200.engine as E
    return E


@pytest.mark.parametrize('nb,dims,occ,empty,shuffle', CASES)
@pytest.mark.parametrize('i64', [True, False])
def test_grid_and_submanifold_rulebook(nb, dims, occ, empty, shuffle, i64):
    E = _E()
    rng = np.random.default_rng(1)
    c = random_coords(rng, nb, dims, occ, empty)
    if shuffle:
        c = c[rng.permutation(c.shape[0])]
    ct = torch.from_numpy(c if i64 else c.astype(np.int32)).cuda()
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    g = E.build_grid(ct, nb, dims, status=status)
    assert int(status.item()) == 0
    assert int(g.prefix[g.n_words].item()) == c.shape[0]
    assert np.array_equal(g.coords.cpu().numpy(), c)
    nbr = E.rulebook_submanifold(g).cpu().numpy()
    assert np.array_equal(nbr, nbr_table(c))
    # lookups: present rows map to themselves, absent cells to -1
    rows = E.grid_lookup(g, g.coords).cpu().numpy()
    assert np.array_equal(rows, np.arange(c.shape[0]))
    far = g.coords.clone()
    far[:, 2] += dims[2]
    assert (E.grid_lookup(g, far).cpu().numpy() == -1).all()


@pytest.mark.parametrize('nb,dims,occ,empty,shuffle', CASES)
def test_coarsen_and_strided_rulebook(nb, dims, occ, empty, shuffle):
    E = _E()
    rng = np.random.default_rng(2)
    c = random_coords(rng, nb, dims, occ, empty)
    if shuffle:
        c = c[rng.permutation(c.shape[0])]
    cc, parent, children, cd = coarse_sets(c, dims)
    g = E.build_grid(torch.from_numpy(c).cuda(), nb, dims)
    cg = E.coarsen(g, dims_cap=cd)
    assert cg.n == cc.shape[0]
    assert np.array_equal(cg.coords.cpu().numpy(), cc)          # raster order
    p, ch = E.rulebook_strided(g, cg)
    assert np.array_equal(p.cpu().numpy(), parent)
    assert np.array_equal(ch.cpu().numpy(), children)
    # the same three tables from the one-kernel form (sgnn_grid_coarse_build: coordinates + strided rulebook, coarse side)
    fc, fp, fch = E.coarse_build(g, cg)
    assert np.array_equal(fc.cpu().numpy(), cc) and np.array_equal(fp.cpu().numpy(), parent)
    assert np.array_equal(fch.cpu().numpy(), children)
    # shifted lookup = parent row
    rows = E.grid_lookup(cg, g.coords, shift=1).cpu().numpy()
    assert np.array_equal(rows, np.where(parent >= 0, parent >> 3, -1))
    # second level
    if cg.n:
        cc2, parent2, children2, cd2 = coarse_sets(cc, cd)
        cg2 = E.coarsen(cg, dims_cap=cd2)
        assert np.array_equal(cg2.coords.cpu().numpy(), cc2)
        assert np.array_equal(E.rulebook_submanifold(cg2).cpu().numpy(), nbr_table(cc2))


def test_duplicates_out_of_range_and_empty():
    E = _E()
    c = torch.tensor([[1, 1, 1, 0], [2, 2, 2, 0], [1, 1, 1, 0], [1, 1, 2, 0]], dtype=torch.int64).cuda()
    g = E.build_grid(c, 1, (4, 4, 4))
    nbr = E.rulebook_submanifold(g).cpu().numpy()
    assert nbr[13].tolist() == [2, 1, 2, 3]          # later duplicate owns the cell
    assert nbr[14].tolist()[0] == 3 and nbr[12].tolist()[3] == 2
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    bad = torch.tensor([[1, 1, 1, 0], [9, 0, 0, 0], [0, 0, 0, 3], [-1, 0, 0, 0]], dtype=torch.int64).cuda()
    g = E.build_grid(bad, 1, (4, 4, 4), status=status)
    assert int(status.item()) == 1 and int(g.prefix[g.n_words].item()) == 1
    e = E.build_grid(torch.zeros((0, 4), dtype=torch.int64, device='cuda'), 1, (4, 4, 4))
    assert e.n == 0 and int(e.prefix[e.n_words].item()) == 0
    assert E.rulebook_submanifold(e).shape == (27, 0)
    ce = E.coarsen(e)
    assert ce.n == 0


def test_rulebook_properties_at_baseline_size():
    """32 blocks of 64^3 @5 % (BASELINE.json configs[1]): symmetry  nbr[k][j]=i <=> nbr[26-k][i]=j."""
    E = _E()
    from sgnn_b200.synth import synthetic_batch
    locs, _ = synthetic_batch(32, 64, 0.05)
    g = E.build_grid(locs.cuda(), 32, (64, 64, 64))
    nbr = E.rulebook_submanifold(g)
    n = g.n
    ar = torch.arange(n, device='cuda', dtype=torch.int32)
    assert torch.equal(nbr[13], ar)
    for k in range(13):
        j = torch.nonzero(nbr[k] >= 0).view(-1)
        i = nbr[k][j].long()
        assert torch.equal(nbr[26 - k][i].long(), j)
    r = int((nbr >= 0).sum().item())
    assert 2.0 * n < r < 2.7 * n                      # 1 + 26 * 0.05 rules per site


def test_rulebook_microbench_size_properties():
    """BASELINE.json configs[4]: ~10 M active sites (763 blocks of 64^3 @5 %), 3^3 submanifold rulebook.
    Size-independent properties: centre offset is the identity, the table is symmetric, rule count ~ 1 + 26*0.05."""
    E = _E()
    nb = 763
    g = torch.Generator(device='cuda').manual_seed(1234)
    mask = torch.rand((nb, 64, 64, 64), device='cuda', generator=g) < 0.05
    coords = torch.nonzero(mask)[:, [1, 2, 3, 0]].contiguous()         # (z,y,x,b), batch-major raster order
    n = coords.shape[0]
    assert 9.5e6 < n < 10.5e6
    del mask
    grid = E.build_grid(coords, nb, (64, 64, 64))
    assert int(grid.prefix[grid.n_words].item()) == n
    nbr = E.rulebook_submanifold(grid)
    assert torch.equal(nbr[13], torch.arange(n, device='cuda', dtype=torch.int32))
    for k in (0, 4, 12):
        j = torch.nonzero(nbr[k] >= 0).view(-1)
        i = nbr[k][j].long()
        assert torch.equal(nbr[26 - k][i].long(), j)
    r = int((nbr >= 0).sum().item())
    assert 2.2 * n < r < 2.4 * n
    # coarse set: every site has exactly one parent, children tables invert the parent table
    cg = E.coarsen(grid)
    parent, children = E.rulebook_strided(grid, cg)
    assert int((parent < 0).sum().item()) == 0
    k = (parent & 7).long()
    row = (parent >> 3).long()
    assert torch.equal(children[k, row].long(), torch.arange(n, device='cuda'))


def test_config2_block_128_conv_deconv_pair():
    """BASELINE.json configs[2] geometry: one 128^3 block @3 %, C=16 stride-2 Convolution then Deconvolution back to
    the fine set, fp32 on the exact kernels (the bf16 tcgen05 pair of the same geometry is tests/test_gpu_tc.py).  Checked against
    the dense identities."""
    import dense_equiv as o1
    E = _E()
    rng = np.random.default_rng(1234)
    mask = rng.random((128, 128, 128)) < 0.03
    c = np.ascontiguousarray(np.concatenate([np.argwhere(mask), np.zeros((int(mask.sum()), 1), dtype=np.int64)], 1))
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32))
    wc = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.2).astype(np.float32))
    wd = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.2).astype(np.float32))
    g = E.build_grid(torch.from_numpy(c).cuda(), 1, (128, 128, 128))
    cg = E.coarsen(g)
    parent, children = E.rulebook_strided(g, cg)
    y = torch.empty((cg.n, 16), device='cuda')
    E.conv(x.cuda(), children, wc.cuda(), cg.n, y)
    z = torch.empty((n, 16), device='cuda')
    E.deconv(y, parent, wd.cuda(), z)
    cc = cg.coords.cpu().long()
    want_y = o1.strided_conv(torch.from_numpy(c), x, wc, 1, (128, 128, 128), cc)
    assert torch.allclose(y.cpu(), want_y, atol=1e-4, rtol=1e-4)
    want_z = o1.strided_deconv(cc, want_y, wd, 1, (64, 64, 64), torch.from_numpy(c))
    assert torch.allclose(z.cpu(), want_z, atol=1e-4, rtol=1e-4)
