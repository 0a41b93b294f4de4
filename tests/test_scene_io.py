""".sdf / .knw readers and scene preparation (CPU)."""
import os
import struct
import sys

import numpy as np
import pytest

from conftest import REFERENCE
from sgnn_b200 import scene_io


def _write_reference_style(path, locs_xyz, sdf_m, dims_xyz, vs, w2g):
    """Byte-level writer following datagen VoxelGrid.h:120-158 with struct.pack (independent of scene_io.save_scene)."""
    with open(path, 'wb') as f:
        f.write(struct.pack('QQQ', *dims_xyz))
        f.write(struct.pack('f', vs))
        f.write(struct.pack('f' * 16, *w2g.reshape(-1)))
        f.write(struct.pack('Q', locs_xyz.shape[0]))
        f.write(struct.pack('I' * locs_xyz.size, *locs_xyz.reshape(-1)))
        f.write(struct.pack('f' * sdf_m.size, *sdf_m))


def test_sdf_roundtrip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    n, dims_xyz, vs = 500, (40, 30, 20), 0.02
    locs_xyz = np.stack([rng.integers(0, d, n) for d in dims_xyz], 1).astype(np.uint32)
    sdf_m = (rng.uniform(-5, 5, n) * vs).astype(np.float32)
    w2g = rng.standard_normal((4, 4)).astype(np.float32)
    p = str(tmp_path / 'a.sdf')
    _write_reference_style(p, locs_xyz, sdf_m, dims_xyz, vs, w2g)
    (locs, sdf), dims, got_w2g = scene_io.load_scene(p)
    assert dims == [20, 30, 40]                                   # [dimz, dimy, dimx]
    assert np.array_equal(locs, locs_xyz[:, ::-1].astype(np.int32))   # flipped to (z,y,x)
    assert np.allclose(sdf, sdf_m / np.float32(vs)) and np.array_equal(got_w2g, w2g)
    q = str(tmp_path / 'b.sdf')
    scene_io.save_scene(q, locs, sdf, dims, vs, w2g)
    assert open(p, 'rb').read()[:100] == open(q, 'rb').read()[:100]
    (l2, s2), d2, _ = scene_io.load_scene(q)
    assert np.array_equal(l2, locs) and np.allclose(s2, sdf, atol=1e-6) and d2 == dims
    with open(str(tmp_path / 'c.knw'), 'wb') as f:
        f.write(open(p, 'rb').read()[:8 * 3 + 4 + 64])
        f.write(bytes(range(256)) * (40 * 30 * 20 // 256) + bytes(40 * 30 * 20 % 256))
    k = scene_io.load_scene_known(str(tmp_path / 'c.knw'))
    assert k.shape == (20, 30, 40) and k[0, 0, 5] == 5
    with open(str(tmp_path / 'bad.sdf'), 'wb') as f:
        f.write(open(p, 'rb').read()[:200])
    with pytest.raises(IOError):
        scene_io.load_scene(str(tmp_path / 'bad.sdf'))


def test_prepare_scene_padding_truncation_height_cap():
    rng = np.random.default_rng(1)
    dims = [150, 70, 33]
    locs = np.stack([rng.integers(0, d, 2000) for d in dims], 1).astype(np.int32)
    sdf = rng.uniform(-6, 6, 2000).astype(np.float32)
    coords, feats, pdims = scene_io.prepare_scene(locs, sdf, dims)
    assert pdims == [128, 96, 64]                                 # height capped at 128, padded to multiples of 32
    assert coords.shape[1] == 4 and int(coords[:, 0].max()) < 128 and float(feats.abs().max()) < 3.0
    keep = (locs[:, 0] < 128) & (np.abs(sdf) < 3.0)
    assert coords.shape[0] == int(keep.sum()) and np.array_equal(coords[:, :3].numpy(), locs[keep])


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='reference tree only exists in the build container')
def test_matches_reference_reader(tmp_path):
    """The reference's data_util.py imports plyfile / its compiled marching-cubes extension (absent here), so its
    load_scene source is exec'ed stand-alone."""
    src = open(os.path.join(REFERENCE, 'data_util.py')).read()
    start = src.index('def load_scene(file):')
    end = src.index('def load_scene_known(file):')
    ns = {'np': np, 'struct': struct}
    exec(src[start:end], ns)
    rng = np.random.default_rng(2)
    locs_xyz = np.stack([rng.integers(0, d, 300) for d in (16, 24, 32)], 1).astype(np.uint32)
    sdf_m = rng.uniform(-0.1, 0.1, 300).astype(np.float32)
    p = str(tmp_path / 'r.sdf')
    _write_reference_style(p, locs_xyz, sdf_m, (16, 24, 32), 0.02, np.eye(4, dtype=np.float32))
    (rl, rs), rd, rw = ns['load_scene'](p)
    (ml, ms), md, mw = scene_io.load_scene(p)
    assert np.array_equal(rl, ml) and np.array_equal(rs, ms) and list(rd) == md and np.array_equal(rw, mw)
