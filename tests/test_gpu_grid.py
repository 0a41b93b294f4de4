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
