"""-m gpu: the bf16 tcgen05/TMEM convolution path (BASELINE.json configs[2]: bf16 features, C = 16 stride-2
Convolution + Deconvolution; also 3^3 submanifold).  Products of bf16 values are exact in fp32, so the reference
is the fp32 fixed-order oracle O3 evaluated on the bf16-rounded inputs; the tensor-core accumulation order and the
final bf16 rounding are covered by a tolerance of one bf16 ulp of the row's magnitude."""
import numpy as np
import pytest
import torch

import o3
from helpers import random_coords, nbr_table, coarse_sets

pytestmark = pytest.mark.gpu


def _E():
    import sgnn_b200.engine as E
    return E


def _close(got_bf16, want_f32):
    got = got_bf16.float().cpu()
    tol = 2.0 ** -7 * want_f32.abs().amax(1, keepdim=True).clamp_min(1e-3) + 1e-3
    assert bool(((got - want_f32).abs() <= tol).all()), float((got - want_f32).abs().max())


@pytest.mark.parametrize('nb,dims,occ', [(1, (12, 10, 14), 0.35), (2, (16, 16, 16), 0.2), (1, (5, 5, 5), 0.9)])
def test_tcgen05_submanifold_and_strided(nb, dims, occ):
    E = _E()
    rng = np.random.default_rng(5)
    c = random_coords(rng, nb, dims, occ)
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32)).bfloat16()
    w27 = torch.from_numpy((rng.standard_normal((27, 16, 16)) * 0.2).astype(np.float32)).bfloat16()
    nbr = torch.from_numpy(nbr_table(c))
    out = torch.empty((n, 16), dtype=torch.bfloat16, device='cuda')
    E.conv(x.cuda(), nbr.cuda(), w27.cuda(), n, out)
    _close(out, o3.conv(x.float(), nbr, w27.float(), n))
    # epilogue: affine + relu in fp32 before the bf16 store
    s, t = torch.rand(16) + 0.5, torch.rand(16) - 0.5
    E.conv(x.cuda(), nbr.cuda(), w27.cuda(), n, out, scale_a=s.cuda(), shift_a=t.cuda(), relu_a=True)
    _close(out, o3.conv(x.float(), nbr, w27.float(), n, scale=s, shift=t, relu=True))
    # stride-2 convolution and deconvolution back (configs[2] operator pair)
    cc, parent, children, cd = coarse_sets(c, dims)
    w8 = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.3).astype(np.float32)).bfloat16()
    y = torch.empty((cc.shape[0], 16), dtype=torch.bfloat16, device='cuda')
    E.conv(x.cuda(), torch.from_numpy(children).cuda(), w8.cuda(), cc.shape[0], y)
    want_y = o3.conv(x.float(), torch.from_numpy(children), w8.float(), cc.shape[0])
    _close(y, want_y)
    wd = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.3).astype(np.float32)).bfloat16()
    z = torch.empty((n, 16), dtype=torch.bfloat16, device='cuda')
    E.deconv(y, torch.from_numpy(parent).cuda(), wd.cuda(), z)
    _close(z, o3.deconv(y.float().cpu(), torch.from_numpy(parent), wd.float()))


def test_tcgen05_config2_block():
    """BASELINE.json configs[2]: one 128^3 block @3 %, bf16, C=16 Convolution(k2,s2) + Deconvolution(k2,s2)."""
    E = _E()
    rng = np.random.default_rng(1234)
    mask = rng.random((128, 128, 128)) < 0.03
    c = np.ascontiguousarray(np.concatenate([np.argwhere(mask), np.zeros((int(mask.sum()), 1), dtype=np.int64)], 1))
    n = c.shape[0]
    x = torch.from_numpy(rng.standard_normal((n, 16)).astype(np.float32)).bfloat16()
    wc = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.2).astype(np.float32)).bfloat16()
    wd = torch.from_numpy((rng.standard_normal((8, 16, 16)) * 0.2).astype(np.float32)).bfloat16()
    g = E.build_grid(torch.from_numpy(c).cuda(), 1, (128, 128, 128))
    cg = E.coarsen(g)
    parent, children = E.rulebook_strided(g, cg)
    y = torch.empty((cg.n, 16), dtype=torch.bfloat16, device='cuda')
    E.conv(x.cuda(), children, wc.cuda(), cg.n, y)
    z = torch.empty((n, 16), dtype=torch.bfloat16, device='cuda')
    E.deconv(y, parent, wd.cuda(), z)
    want_y = o3.conv(x.float(), children.cpu(), wc.float(), cg.n)
    _close(y, want_y)
    _close(z, o3.deconv(y.float().cpu(), parent.cpu(), wd.float()))
