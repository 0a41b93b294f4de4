"""Oracle for SURVEY 8(f4) (marching cubes after the forward pass, data_util.py:270-284): oracle/mcubes.cpp, the CPU
restatement of the reference's run_marching_cubes, against
  * tests/golden/mc_ref.npz -- outputs of the REAL reference (torch/marching_cubes/marching_cubes.cpp compiled in place
    into oracle/_ref/ by oracle/build_ref.py; generator: tests/golden/make_mc_golden.py), and
  * the real reference run live on random volumes, wherever oracle/_ref/marching_cubes_cpp.so exists.
Vertices are compared bit for bit, faces exactly."""
import os

import numpy as np
import pytest
import torch

import build_ref
import mcubes
from conftest import GOLDEN


def _same(v, f, V, F):
    assert v.shape == V.shape and f.shape == F.shape
    assert np.array_equal(v.view(np.uint32), V.view(np.uint32)), 'vertex bits differ'
    assert np.array_equal(f, F)


@pytest.mark.parametrize('name', ['sphere', 'blobs', 'noise', 'plane'])
def test_restatement_equals_reference_fixture(name):
    g = np.load(os.path.join(GOLDEN, 'mc_ref.npz'))
    v, f = mcubes.marching_cubes(g['case_%s_tsdf' % name])
    _same(v, f, g['case_%s_verts' % name], g['case_%s_faces' % name])
    assert f.min() >= 0 and f.max() < v.shape[0]


def test_recovered_table_is_a_marching_cubes_table():
    t = mcubes.tri_table()
    assert t.shape == (256, 16) and t.dtype == np.int8
    assert (t[0] == -1).all() and (t[255] == -1).all()
    n = (t >= 0).sum(1)
    assert (n % 3 == 0).all() and n.max() <= 15
    assert int((n > 0).sum()) == 252            # the reference skips two more configurations (its edge-mask == 255 test)
    for c in range(256):                        # every triangle vertex sits on an edge whose two corners differ in sign
        edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
        for e in t[c][t[c] >= 0]:
            a, b = edges[e]
            assert ((c >> a) & 1) != ((c >> b) & 1)


def test_empty_and_unobserved_volumes():
    v, f = mcubes.marching_cubes(np.full((6, 6, 6), 1.5, dtype=np.float32))
    assert v.shape == (0, 3) and f.shape == (0, 3)
    d = np.full((6, 6, 6), -np.inf, dtype=np.float32)
    v, f = mcubes.marching_cubes(d)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v, f = mcubes.marching_cubes(np.zeros((2, 2, 2), dtype=np.float32))      # no cell has eight valid corners
    assert v.shape == (0, 3) and f.shape == (0, 3)


@pytest.mark.skipif(not os.path.exists(build_ref.mc_path()), reason='oracle/_ref/marching_cubes_cpp.so not built')
@pytest.mark.parametrize('seed,dims', [(0, (12, 12, 12)), (1, (9, 17, 13)), (2, (24, 20, 28)), (3, (5, 6, 7))])
def test_restatement_equals_live_reference(seed, dims):
    mc = build_ref.load_marching_cubes()
    rng = np.random.default_rng(seed)
    n = rng.standard_normal(dims)
    for ax in range(3):
        n = (np.roll(n, 1, ax) + n + np.roll(n, -1, ax)) / 3
    d = (3.4 * n / np.abs(n).max()).astype(np.float32)          # some voxels beyond the truncation: holes
    d[rng.random(dims) < 0.02] = -np.inf                        # unobserved voxels
    d[rng.random(dims) < 0.02] = 0.0                            # exact isovalue hits
    t = torch.from_numpy(d)
    col = torch.ones(dims + (3,), dtype=torch.uint8) * 220
    V, _, F = mc.run_marching_cubes(t, col, 0.0, 3.0, 10.0)
    v, f = mcubes.marching_cubes(d)
    _same(v, f, V.numpy().astype(np.float32), F.numpy().astype(np.int32))
