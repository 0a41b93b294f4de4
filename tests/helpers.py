"""Shared test helpers (CPU side): reference tables built from oracle O2's site index."""
import numpy as np
import torch

import sparseconvnet as o2          # oracle/sparseconvnet (conftest puts oracle/ on sys.path)

OFFS27 = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


def random_coords(rng, nb, dims, occ, empty=()):
    cs = []
    for b in range(nb):
        if b in empty:
            continue
        m = rng.random(tuple(dims)) < occ
        c = np.argwhere(m)
        cs.append(np.concatenate([c, np.full((c.shape[0], 1), b)], 1))
    if not cs:
        return np.zeros((0, 4), dtype=np.int64)
    return np.ascontiguousarray(np.concatenate(cs).astype(np.int64))


def nbr_table(coords):
    """[27, n] int32 neighbour table from oracle O2's site index (App. A.3 order)."""
    ss = o2._SiteSet(torch.from_numpy(coords))
    out = np.empty((27, coords.shape[0]), dtype=np.int32)
    for k, (dz, dy, dx) in enumerate(OFFS27):
        q = coords.copy()
        q[:, 0] += dz
        q[:, 1] += dy
        q[:, 2] += dx
        out[k] = ss.lookup(q)
    return out


def coarse_sets(coords, dims):
    """Raster-ordered coarse coords, parent (row*8+k) and children [8, nc] for filter 2 stride 2."""
    cd = [(d - 2) // 2 + 1 if d >= 2 else 0 for d in dims]
    q = coords.copy()
    q[:, :3] >>= 1
    ok = (q[:, 0] < cd[0]) & (q[:, 1] < cd[1]) & (q[:, 2] < cd[2])
    k = ((coords[:, 0] & 1) << 2) | ((coords[:, 1] & 1) << 1) | (coords[:, 2] & 1)
    key = ((q[:, 3] * max(cd[0], 1) + q[:, 0]) * max(cd[1], 1) + q[:, 1]) * max(cd[2], 1) + q[:, 2]
    uk, inv = np.unique(key[ok], return_inverse=True)
    nc = uk.shape[0]
    ccoords = np.zeros((nc, 4), dtype=np.int64)
    ccoords[inv] = q[ok]
    parent = np.full(coords.shape[0], -1, dtype=np.int32)
    parent[ok] = inv * 8 + k[ok]
    children = np.full((8, nc), -1, dtype=np.int32)
    children[k[ok], inv] = np.nonzero(ok)[0]
    return ccoords, parent, children, cd


def sorted_rows(locs, *vals):
    """Sort rows by (b,z,y,x) for set comparison; locs numpy [n,4]."""
    locs = np.asarray(locs).astype(np.int64)
    keys = [locs[:, i] for i in range(locs.shape[1])]
    order = np.lexsort(tuple(keys[:3][::-1]) + ((keys[3],) if len(keys) > 3 else ()))
    return (locs[order],) + tuple(np.asarray(v)[order] for v in vals)
