"""Oracle O2 (restated scn) against oracle O1 (dense conv3d identities): pins offset order, weight layout,
cross-correlation, stride-2 rules, deconvolution / unpooling, batching, empty samples, volume borders."""
import numpy as np
import pytest
import torch

import sparseconvnet as o2
import dense_equiv as o1
from helpers import random_coords

CASES = [
    # nb, dims, occ, empty samples, cin, cout
    (1, (8, 8, 8), 0.3, (), 1, 8),
    (2, (16, 12, 10), 0.15, (), 8, 12),
    (3, (8, 16, 8), 0.4, (1,), 12, 16),
    (1, (6, 6, 70), 0.2, (), 16, 16),       # x extent > 64 (two mask words per row on the device side)
    (2, (4, 4, 4), 1.0, (), 5, 7),          # fully dense incl. all borders
]


def _tensor(nb, dims, occ, empty, cin, seed):
    rng = np.random.default_rng(seed)
    c = torch.from_numpy(random_coords(rng, nb, dims, occ, empty))
    f = torch.from_numpy(rng.standard_normal((c.shape[0], cin)).astype(np.float32))
    t = o2.InputLayer(3, list(dims), mode=0)([c, f, nb])
    return c, f, t


@pytest.mark.parametrize('nb,dims,occ,empty,cin,cout', CASES)
def test_submanifold_matches_dense(nb, dims, occ, empty, cin, cout):
    c, f, t = _tensor(nb, dims, occ, empty, cin, 1)
    conv = o2.SubmanifoldConvolution(3, cin, cout, 3, False)
    with torch.no_grad():
        got = conv(t).features
        want = o1.submanifold_conv(c, f, conv.weight, nb, dims)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize('nb,dims,occ,empty,cin,cout', CASES)
def test_strided_deconv_unpool_match_dense(nb, dims, occ, empty, cin, cout):
    c, f, t = _tensor(nb, dims, occ, empty, cin, 2)
    conv = o2.Convolution(3, cin, cout, 2, 2, False)
    with torch.no_grad():
        y = conv(t)
        cc = y.metadata.getSpatialLocations(y.spatial_size)
        want = o1.strided_conv(c, f, conv.weight, nb, dims, cc)
        assert torch.allclose(y.features, want, atol=2e-5, rtol=1e-5)
        # every parent of an in-range fine site is active exactly once
        cd = [(d - 2) // 2 + 1 for d in dims]
        par = c.clone()
        par[:, :3] //= 2
        ok = (par[:, 0] < cd[0]) & (par[:, 1] < cd[1]) & (par[:, 2] < cd[2])
        assert set(map(tuple, par[ok].tolist())) == set(map(tuple, cc.tolist()))
        assert len(set(map(tuple, cc.tolist()))) == cc.shape[0]
        if all(d % 2 == 0 for d in dims):
            dec = o2.Deconvolution(3, cout, cin, 2, 2, False)
            z = dec(y)
            wantz = o1.strided_deconv(cc, y.features, dec.weight, nb, cd, c)
            assert torch.allclose(z.features, wantz, atol=2e-5, rtol=1e-5)
            u = o2.UnPooling(3, 2, 2)(y)
            wantu = o1.unpool(cc, y.features, nb, cd, c)
            assert torch.equal(u.features, wantu)


def test_batchnorm_relu_eval_and_train():
    torch.manual_seed(0)
    c, f, t = _tensor(2, (8, 8, 8), 0.3, (), 12, 3)
    bn = o2.BatchNormReLU(12)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.3, 0.3)
        bn.running_mean.uniform_(-0.3, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
    bn.eval()
    ref = torch.nn.functional.batch_norm(f, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, 1e-4)
    assert torch.allclose(bn(t).features, ref.clamp_min(0), atol=1e-6)
    bn.train()
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    y = bn(t).features
    ref = torch.nn.functional.batch_norm(f, None, None, bn.weight, bn.bias, True, 0.0, 1e-4).clamp_min(0)
    assert torch.allclose(y, ref, atol=1e-5)
    assert torch.allclose(bn.running_mean, 0.9 * rm0 + 0.1 * f.mean(0), atol=1e-6)
    assert torch.allclose(bn.running_var, 0.9 * rv0 + 0.1 * f.var(0, unbiased=True), atol=1e-5)


def test_sparse_to_dense_and_fcn_width():
    c, f, t = _tensor(2, (8, 8, 8), 0.3, (), 16, 4)
    d = o2.SparseToDense(3, 16)(t)
    assert d.shape == (2, 16, 8, 8, 8)
    assert torch.equal(o1.densify(c, f, 2, (8, 8, 8)), d)
    fcn = o2.FullyConvolutionalNet(3, reps=1, nPlanes=[16, 16, 16], residual_blocks=True).eval()
    with torch.no_grad():
        y = fcn(t)
    assert y.features.shape == (c.shape[0], 48)          # model.py:181,256,258


def test_input_layer_mode0_duplicates_and_empty():
    c = torch.tensor([[1, 1, 1, 0], [2, 2, 2, 0], [1, 1, 1, 0]])
    f = torch.arange(3.).view(3, 1)
    t = o2.InputLayer(3, [4, 4, 4], mode=0)([c, f])
    rules = t.metadata.getSubmanifoldRuleBook(t.spatial_size, 3)
    centre = rules[13]
    # the later duplicate (row 2) owns the cell: rows 0 and 2 both read row 2 at the centre offset
    assert dict(zip(centre[1].tolist(), centre[0].tolist())) == {0: 2, 1: 1, 2: 2}
    e = o2.InputLayer(3, [4, 4, 4], mode=0)([torch.zeros((0, 4), dtype=torch.long), torch.zeros((0, 1))])
    assert o2.SubmanifoldConvolution(3, 1, 8, 3, False)(e).features.shape == (0, 8)
