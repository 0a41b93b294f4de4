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


def _reference_modules():
    """The reference's data_util.py / scene_dataloader.py imported in place with stand-ins for the modules this image
    lacks (plyfile, the compiled marching-cubes extension): neither is touched by the readers or the data set."""
    import importlib.util
    import types
    saved = {k: sys.modules.get(k) for k in ('plyfile', 'marching_cubes', 'marching_cubes.marching_cubes', 'data_util')}
    try:
        sys.modules['plyfile'] = types.ModuleType('plyfile')
        pkg = types.ModuleType('marching_cubes')
        pkg.marching_cubes = types.ModuleType('marching_cubes.marching_cubes')
        sys.modules['marching_cubes'], sys.modules['marching_cubes.marching_cubes'] = pkg, pkg.marching_cubes
        mods = []
        for name in ('data_util', 'scene_dataloader'):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, name + '.py'))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = m
            spec.loader.exec_module(m)
            mods.append(m)
        return mods
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.pop('scene_dataloader', None)


def _synthetic_chunk(rng, dims_zyx):
    def sparse(d, n):
        flat = rng.choice(d[0] * d[1] * d[2], size=n, replace=False)
        locs = np.stack(np.unravel_index(flat, d), 1).astype(np.int32)
        return locs, rng.uniform(-4, 4, n).astype(np.float32)
    inp, tgt = sparse(dims_zyx, 700), sparse(dims_zyx, 900)
    hier = [sparse([v // f for v in dims_zyx], 200 // f) for f in (2, 4, 8)]
    known = rng.integers(0, 4, dims_zyx).astype(np.uint8)
    return inp, tgt, known, hier


def _same(a, b):
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if a is None or isinstance(a, str):
        return a == b
    a, b = (t.numpy() if hasattr(t, 'numpy') else np.asarray(t) for t in (a, b))
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


def test_sdfs_chunk_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    dims = [32, 48, 64]
    inp, tgt, known, hier = _synthetic_chunk(rng, dims)
    p = str(tmp_path / 'c__0__.sdfs')
    scene_io.save_train_file(p, inp[0], inp[1], tgt[0], tgt[1], known, hier, dims, 0.02)
    (il, iv), target, d, w2g, k, h = scene_io.load_train_file(p)
    assert d == dims and np.array_equal(il, inp[0]) and np.allclose(iv, inp[1], atol=1e-5)
    assert target.shape == tuple(dims) and np.isneginf(target).sum() == target.size - 900
    assert np.allclose(target[tgt[0][:, 0], tgt[0][:, 1], tgt[0][:, 2]], tgt[1], atol=1e-5)
    assert np.array_equal(k, known) and [g.shape for g in h] == [(4, 6, 8), (8, 12, 16), (16, 24, 32)]
    with open(str(tmp_path / 'bad.sdfs'), 'wb') as f:
        f.write(open(p, 'rb').read()[:5000])
    with pytest.raises(IOError):
        scene_io.load_train_file(str(tmp_path / 'bad.sdfs'))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='reference tree only exists in the build container')
def test_train_chunks_dataset_and_collate_match_reference(tmp_path):
    """.sdfs reader, SceneDataset (chunk mode) and collate against the reference's own functions on the same files
    (data_util.py:63-109, scene_dataloader.py:13-120): every field of every sample and of the collated batch equal."""
    import torch
    du, sd = _reference_modules()
    rng = np.random.default_rng(6)
    files = []
    for i in range(3):
        inp, tgt, known, hier = _synthetic_chunk(rng, [64, 64, 64])
        files.append(str(tmp_path / ('chunk%d__0__.sdfs' % i)))
        scene_io.save_train_file(files[-1], inp[0], inp[1], tgt[0], tgt[1], known, hier, [64, 64, 64], 0.02,
                                 rng.standard_normal((4, 4)).astype(np.float32))
    assert _same(list(du.load_train_file(files[0])), list(scene_io.load_train_file(files[0])))
    for levels in (4, 3):
        ref_ds = sd.SceneDataset(files + ['/nonexistent.sdfs'], [64, 64, 64], 3.0, levels, 128, 0, '')
        my_ds = scene_io.SceneDataset(files + ['/nonexistent.sdfs'], [64, 64, 64], 3.0, levels, 128, 0, '')
        assert len(ref_ds) == len(my_ds) == 3
        rs, ms = [ref_ds[i] for i in range(3)], [my_ds[i] for i in range(3)]
        for r, m in zip(rs, ms):
            assert sorted(r) == sorted(m) and all(_same(r[k], m[k]) for k in r), [k for k in r if not _same(r[k], m[k])]
        rb, mb = sd.collate(rs), scene_io.collate(ms)
        assert sorted(rb) == sorted(mb) and all(_same(rb[k], mb[k]) for k in rb), [k for k in rb if not _same(rb[k], mb[k])]
        assert mb['input'][0].dtype == torch.int64 and int(mb['input'][0][:, 3].max()) == 2


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='reference tree only exists in the build container')
@pytest.mark.parametrize('max_height', [128, 40])
def test_scene_pairs_dataset_matches_reference(tmp_path, max_height):
    """Scene mode (input .sdf + target .sdf/.knw, test_scene.py:117): padding, height cap, truncation mask."""
    du, sd = _reference_modules()
    rng = np.random.default_rng(7)
    dims = [50, 70, 33]
    os.makedirs(str(tmp_path / 'in')), os.makedirs(str(tmp_path / 'tgt'))
    for i in range(2):
        for sub, n in (('in', 1500), ('tgt', 2500)):
            locs = np.stack([rng.integers(0, d, n) for d in dims], 1).astype(np.int32)
            locs = np.unique(locs, axis=0)
            scene_io.save_scene(str(tmp_path / sub / ('s%d.sdf' % i)), locs, rng.uniform(-6, 6, locs.shape[0]).astype(np.float32), dims)
        with open(str(tmp_path / 'tgt' / ('s%d.knw' % i)), 'wb') as f:
            f.write(open(str(tmp_path / 'tgt' / ('s%d.sdf' % i)), 'rb').read()[:8 * 3 + 4 + 64])
            f.write(rng.integers(0, 3, dims[0] * dims[1] * dims[2]).astype(np.uint8).tobytes())
    files = [str(tmp_path / 'in' / ('s%d.sdf' % i)) for i in range(2)]
    ref_ds = sd.SceneDataset(files, [128, 64, 64], 3.0, 4, max_height, 0, str(tmp_path / 'tgt'))
    my_ds = scene_io.SceneDataset(files, [128, 64, 64], 3.0, 4, max_height, 0, str(tmp_path / 'tgt'))
    assert len(ref_ds) == len(my_ds) == 2
    for i in range(2):
        r, m = ref_ds[i], my_ds[i]
        assert all(_same(r[k], m[k]) for k in r), [k for k in r if not _same(r[k], m[k])]
        b_r, b_m = sd.collate([r]), scene_io.collate([m])
        assert all(_same(b_r[k], b_m[k]) for k in b_r)
    assert tuple(my_ds[0]['sdf'].shape[1:]) == ((64 if max_height == 40 else 64), 96, 64)


def test_save_predictions_scope_is_the_test_scene_call():
    """sgnn_b200.scene.save_predictions mirrors the call test_scene.py:98 makes (targets None, per-level occupancy None);
    the train.py-only forms are refused loudly instead of silently writing something else."""
    import torch
    from sgnn_b200 import scene
    locs = torch.zeros((1, 4), dtype=torch.long)
    feats = torch.zeros((1, 1))
    for kw in ({'target_for_sdf': torch.zeros(1, 1, 2, 2, 2)}, {'target_for_occs': [None]}, {'output_occs': [None]}):
        args = dict(target_for_sdf=None, target_for_occs=None, output_occs=None)
        args.update(kw)
        with pytest.raises(ValueError):
            scene.save_predictions('/tmp/unused', ['x'], [locs, feats], args['target_for_sdf'], args['target_for_occs'], [None],
                                   args['output_occs'], None, 3.0)
