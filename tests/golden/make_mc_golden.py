"""Golden vectors for SURVEY 8(f4) (marching cubes after the forward pass, data_util.py:270-284), produced by the REAL
reference: oracle/_ref/marching_cubes_cpp.so is /root/reference/torch/marching_cubes/marching_cubes.cpp compiled in place
(oracle/build_ref.py).  Run in the build container (needs /root/reference):

    python tests/golden/make_mc_golden.py

Writes tests/golden/mc_ref.npz:
  tri_table [256,16] int8   the triangulation of every cube configuration, RECOVERED by probing the reference with 3^3
                            volumes whose eight corner averages have the signs of the configuration (only the centre
                            cell of a 3^3 volume has eight valid corners) and mapping the returned vertices back to
                            cube edges -- including the reference's own quirks (configurations it skips);
  case_<name>_{tsdf,verts,faces}   inputs and outputs of run_marching_cubes(tsdf, 220-grey, isovalue 0, truncation 3,
                            thresh 10) -- the arguments of data_util.py:270 -- on small volumes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import build_ref  # noqa: E402

# corner naming of the reference: pXYZ = cell centre + (+-0.5 x, +-0.5 y, +-0.5 z); bit order of the cube index
CORNER_BITS = [(0, 1, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0), (0, 1, 1), (1, 1, 1), (1, 0, 1), (0, 0, 1)]     # (sx, sy, sz)
# cube edges in the reference's vertlist order: pairs of corners (sx, sy, sz)
EDGES = [((0, 1, 0), (1, 1, 0)), ((1, 1, 0), (1, 0, 0)), ((1, 0, 0), (0, 0, 0)), ((0, 0, 0), (0, 1, 0)),
         ((0, 1, 1), (1, 1, 1)), ((1, 1, 1), (1, 0, 1)), ((1, 0, 1), (0, 0, 1)), ((0, 0, 1), (0, 1, 1)),
         ((0, 1, 0), (0, 1, 1)), ((1, 1, 0), (1, 1, 1)), ((1, 0, 0), (1, 0, 1)), ((0, 0, 0), (0, 0, 1))]


def run(mc, tsdf):
    t = torch.from_numpy(np.ascontiguousarray(tsdf, dtype=np.float32))
    col = torch.ones(t.shape[0], t.shape[1], t.shape[2], 3, dtype=torch.uint8) * 220
    v, c, f = mc.run_marching_cubes(t, col, 0.0, 3.0, 10.0)
    return v.numpy().astype(np.float32), f.numpy().astype(np.int32)


def probe_volume(config):
    """3^3 volume whose centre cell has corner averages -0.1 (bit set: inside) / +0.2 (outside)."""
    a = np.zeros((8, 27))
    s = np.zeros(8)
    for bit, (sx, sy, sz) in enumerate(CORNER_BITS):
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    a[bit, ((sz + dz) * 3 + (sy + dy)) * 3 + (sx + dx)] = 0.125
        s[bit] = -0.1 if (config >> bit) & 1 else 0.2
    v = np.linalg.pinv(a) @ s
    assert np.abs(v).max() < 2.9, np.abs(v).max()
    return v.reshape(3, 3, 3).astype(np.float32)


def edge_of(p):
    """Cube edge (reference numbering) a vertex of the centre cell [0.5,1.5]^3 lies on."""
    on = [abs(p[i] - 0.5) < 1e-4 or abs(p[i] - 1.5) < 1e-4 for i in range(3)]
    assert sum(on) == 2, p
    for e, (ca, cb) in enumerate(EDGES):
        ok = True
        for i in range(3):
            lo, hi = 0.5 + min(ca[i], cb[i]), 0.5 + max(ca[i], cb[i])
            if ca[i] == cb[i]:
                ok &= abs(p[i] - lo) < 1e-4
            else:
                ok &= lo - 1e-4 < p[i] < hi + 1e-4
        if ok:
            return e
    raise AssertionError(p)


def main():
    build_ref.build_marching_cubes()
    mc = build_ref.load_marching_cubes()
    assert mc is not None, 'needs /root/reference (build container)'
    tri = -np.ones((256, 16), dtype=np.int8)
    for config in range(256):
        verts, faces = run(mc, probe_volume(config))
        flat = [edge_of(verts[i]) for f in faces for i in f]
        assert len(flat) <= 15
        tri[config, :len(flat)] = flat
    out = {'tri_table': tri}
    rng = np.random.default_rng(7)
    z, y, x = np.mgrid[0:20, 0:20, 0:20].astype(np.float32)
    out['case_sphere_tsdf'] = np.clip(np.sqrt((x - 9.3) ** 2 + (y - 10.1) ** 2 + (z - 9.7) ** 2) - 6.3, -3, 3)
    z, y, x = np.mgrid[0:14, 0:22, 0:18].astype(np.float32)
    d = np.minimum(np.sqrt((x - 5) ** 2 + (y - 7) ** 2 + (z - 6) ** 2) - 3.6, np.sqrt((x - 12) ** 2 + (y - 14) ** 2 + (z - 7) ** 2) - 4.4)
    d = np.clip(d, -3, 3).astype(np.float32)
    d[:, 10:12, :] = -np.inf                                   # unobserved slab: no surface may cross it
    out['case_blobs_tsdf'] = d
    n = rng.standard_normal((16, 16, 16))
    for ax in range(3):                                        # cheap smoothing: box filter twice per axis
        for _ in range(2):
            n = (np.roll(n, 1, ax) + n + np.roll(n, -1, ax)) / 3
    out['case_noise_tsdf'] = (2.5 * n / np.abs(n).max()).astype(np.float32)
    z, y, x = np.mgrid[0:8, 0:9, 0:16].astype(np.float32)
    out['case_plane_tsdf'] = np.clip(x - 7.5, -3, 3)           # corner averages hit the isovalue exactly
    for name in ('sphere', 'blobs', 'noise', 'plane'):
        v, f = run(mc, out['case_%s_tsdf' % name])
        out['case_%s_verts' % name], out['case_%s_faces' % name] = v, f
        print(name, out['case_%s_tsdf' % name].shape, v.shape, f.shape)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'mc_ref.npz'), **out)
    print('configurations the reference triangulates:', int((tri[:, 0] >= 0).sum()), 'of 256')


if __name__ == '__main__':
    main()
