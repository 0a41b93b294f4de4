"""Oracle O3 (fixed-order C, the bit-exact target of the CUDA kernels) against oracle O2 (torch mm order)."""
import numpy as np
import pytest
import torch

import o3
import sparseconvnet as o2
from helpers import random_coords, nbr_table, coarse_sets

PAIRS = [(1, 8), (8, 8), (8, 12), (12, 12), (12, 16), (16, 16), (34, 16), (30, 16), (26, 16), (48, 16)]


@pytest.mark.parametrize('cin,cout', PAIRS)
def test_o3_conv_matches_o2(cin, cout):
    rng = np.random.default_rng(cin * 100 + cout)
    dims = (10, 9, 12)
    c = random_coords(rng, 2, dims, 0.35)
    f = torch.from_numpy(rng.standard_normal((c.shape[0], cin)).astype(np.float32))
    t = o2.InputLayer(3, list(dims), mode=0)([torch.from_numpy(c), f])
    conv = o2.SubmanifoldConvolution(3, cin, cout, 3, False)
    with torch.no_grad():
        want = conv(t).features
    got = o3.conv(f, torch.from_numpy(nbr_table(c)), conv.weight.detach(), c.shape[0])
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def test_o3_strided_deconv_linear_bn():
    rng = np.random.default_rng(7)
    dims = (8, 12, 10)
    c = random_coords(rng, 2, dims, 0.3)
    f = torch.from_numpy(rng.standard_normal((c.shape[0], 16)).astype(np.float32))
    t = o2.InputLayer(3, list(dims), mode=0)([torch.from_numpy(c), f])
    conv = o2.Convolution(3, 16, 16, 2, 2, False)
    dec = o2.Deconvolution(3, 16, 16, 2, 2, False)
    with torch.no_grad():
        y = conv(t)
        z = dec(y)
    cc, parent, children, cd = coarse_sets(c, dims)
    # O2 numbers coarse rows by first touch, the engine by raster order: compare as coordinate-keyed rows
    o2c = y.metadata.getSpatialLocations(y.spatial_size).numpy()
    key = lambda a: ((a[:, 3] * 64 + a[:, 0]) * 64 + a[:, 1]) * 64 + a[:, 2]
    perm = np.argsort(key(o2c))
    assert np.array_equal(key(o2c)[perm], key(cc))
    got = o3.conv(f, torch.from_numpy(children), conv.weight.detach(), cc.shape[0])
    assert torch.allclose(got, y.features[perm], atol=1e-5, rtol=1e-5)
    gotz = o3.deconv(got, torch.from_numpy(parent), dec.weight.detach())
    assert torch.allclose(gotz, z.features, atol=1e-5, rtol=1e-5)
    lin = torch.nn.Linear(16, 1)
    with torch.no_grad():
        assert torch.allclose(o3.linear(f, lin.weight, lin.bias), lin(f), atol=1e-5)
    s, b = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    assert torch.allclose(o3.affine_relu(f, s, b), (f * s + b).clamp_min(0), atol=1e-6)


def test_o3_child_mode_equals_replicated_children():
    """model.py:192-207,224-225: SMC over all 8 children with x8 replicated parent features."""
    rng = np.random.default_rng(11)
    dims = (6, 5, 7)
    c = random_coords(rng, 2, dims, 0.4)
    n = c.shape[0]
    f = torch.from_numpy(rng.standard_normal((n, 48)).astype(np.float32))
    kids = (c[:, None, :] * np.array([2, 2, 2, 1]) +
            np.array([[z, y, x, 0] for z in (0, 1) for y in (0, 1) for x in (0, 1)])[None]).reshape(-1, 4)
    w = torch.from_numpy(rng.standard_normal((27, 48, 16)).astype(np.float32) * 0.05)
    want = o3.conv(f.repeat_interleave(8, 0), torch.from_numpy(nbr_table(kids)), w, 8 * n)
    # child-level neighbour rows are 8*parent_row + child: integer-divide to parent rows
    got = o3.conv(f, torch.from_numpy(nbr_table(c)), w, 8 * n, child_mode=True)
    assert torch.equal(got, want)


def test_sigmoid_threshold_literal():
    x = torch.tensor([0.0, 1e-8, 5e-8, 8.9e-8, 1.2e-7, 1e-6, -1e-6, 3.0, -3.0], dtype=torch.float32)
    assert torch.equal(o3.sigmoid_gt_half(x), torch.sigmoid(x) > 0.5)
    xs = torch.from_numpy(np.random.default_rng(0).standard_normal(100000).astype(np.float32) * 1e-7)
    assert torch.equal(o3.sigmoid_gt_half(xs), torch.sigmoid(xs) > 0.5)


@pytest.mark.parametrize('transposed,cin,cout,k,s,p,dims', [
    (False, 16, 24, 4, 2, 1, (8, 8, 8)), (False, 24, 32, 4, 2, 1, (4, 4, 4)), (False, 32, 32, 1, 1, 0, (2, 2, 2)),
    (True, 64, 32, 4, 2, 1, (2, 2, 2)), (True, 56, 28, 4, 2, 1, (4, 6, 4)), (False, 28, 16, 1, 1, 0, (8, 4, 8)),
])
def test_o3_dense_layers_match_torch(transposed, cin, cout, k, s, p, dims):
    torch.manual_seed(cin + cout)
    x = torch.randn(2, cin, *dims)
    if transposed:
        m = torch.nn.ConvTranspose3d(cin, cout, k, s, p, bias=False)
    else:
        m = torch.nn.Conv3d(cin, cout, k, s, p, bias=False)
    bn = torch.nn.BatchNorm3d(cout).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5), bn.bias.uniform_(-0.3, 0.3)
        bn.running_mean.uniform_(-0.3, 0.3), bn.running_var.uniform_(0.5, 1.5)
        want = torch.relu(bn(m(x)))
        inv = (bn.running_var + bn.eps).pow(-0.5)
        scale = inv * bn.weight
        shift = bn.bias - bn.running_mean * scale
        got = o3.dense_conv(x, m.weight, k, s, p, scale, shift, True, transposed)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)
