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


def compare_generator_outputs(want, got, margin=1e-5, tol_logit=1e-4, tol_sdf=1e-3, max_flips_per_level=2, tag=''):
    """Margin-aware comparison of two generator results  ((out_locs, out_sdf), levels), block by block.

    Blocks (batch indices) are independent.  At every level the candidate coordinates of the blocks that have not diverged
    must be EQUAL, in order, and the (occ, sdf) values within tol_logit.  A mask flip is legal only where the reference
    logit is within `margin` of the threshold; it is COUNTED (bounded by max_flips_per_level, printed) and only the block
    it happened in is excluded from the later levels -- every other block keeps being compared to the end, including the
    final coordinates and the TSDF head (tol_sdf).  Returns (flips per level, diverged blocks)."""
    (wl, ws), wlv = want
    (gl, gs), glv = got
    cpu = lambda t: t.detach().cpu() if isinstance(t, torch.Tensor) else t
    diverged = set()
    flips_per_level = []

    def live(locs):
        locs = cpu(locs)
        if not diverged:
            return torch.ones(locs.shape[0], dtype=torch.bool)
        return ~torch.isin(locs[:, 3], torch.tensor(sorted(diverged), dtype=locs.dtype))
    assert len(wlv) == len(glv)
    for i, (w, g) in enumerate(zip(wlv, glv)):
        if isinstance(w[0], list) or isinstance(g[0], list):          # empty level (model.py:211)
            assert isinstance(w[0], list) and isinstance(g[0], list), 'level %d: one side empty' % i
            flips_per_level.append(0)
            continue
        w0, w1, g0, g1 = cpu(w[0]), cpu(w[1]), cpu(g[0]), cpu(g[1])
        kw, kg = live(w0), live(g0)
        assert torch.equal(w0[kw], g0[kg]), '%s candidate coordinates at level %d (blocks %s excluded)' % (tag, i, sorted(diverged))
        lw, lg = w1[kw], g1[kg]
        err = float((lw - lg).abs().max()) if lw.numel() else 0.0
        assert err <= tol_logit, '%s level %d logits differ by %.3e' % (tag, i, err)
        fl = (torch.sigmoid(lw[:, 0]) > 0.5) != (torch.sigmoid(lg[:, 0]) > 0.5)
        n = int(fl.sum())
        flips_per_level.append(n)
        if n:
            assert bool((lw[:, 0][fl].abs() < margin).all()), '%s illegal mask flip at level %d (|logit| >= %g)' % (tag, i, margin)
            assert n <= max_flips_per_level, '%s %d legal flips at level %d' % (tag, n, i)
            diverged |= set(int(b) for b in w0[kw][fl][:, 3].tolist())
    if isinstance(wl, list) or isinstance(gl, list):
        assert isinstance(wl, list) and isinstance(gl, list)
    else:
        wl, ws, gl, gs = cpu(wl), cpu(ws), cpu(gl), cpu(gs)
        kw, kg = live(wl), live(gl)
        assert torch.equal(wl[kw], gl[kg]), '%s final coordinates' % tag
        err = float((ws[kw] - gs[kg]).abs().max()) if int(kw.sum()) else 0.0
        assert err <= tol_sdf, '%s TSDF differs by %.3e' % (tag, err)
        print('%s parity: flips per level %s, diverged blocks %s, TSDF max err %.2e over %d voxels' % (
            tag, flips_per_level, sorted(diverged), err, int(kw.sum())))
    return flips_per_level, diverged


def compare_teacher_forced(oracle, locs, feats, got, margin=1e-5, tol_logit=1e-4, rtol_logit=1e-5, tol_sdf=1e-3, max_flips=8,
                           tag=''):
    """Whole-input parity with NOTHING excluded (for batch-1 scenes, where one flipped mask bit would otherwise take the
    only block out of the comparison): the oracle generator is run with the device pass's keep decisions forced on it
    (OracleGenModel.forward(..., forced_keep=...)), so both sides refine the same sites at every level.  Checked at every
    level over all candidates: coordinates EQUAL in order, (occ, sdf) within tol_logit + rtol_logit * |value| (whole scenes
    reach |logit| ~ 50, where 1e-4 absolute would be 0.3 ulp-scale of the fp32 sums behind it), and the oracle's OWN decision
    sigmoid(occ) > 0.5 equal to the device's except where the oracle logit is within `margin` of the threshold -- those
    legal flips are counted, bounded (max_flips in total) and printed.  Then the final coordinates (equal) and the TSDF
    head (tol_sdf).  Returns the flips per level."""
    (gl, gs), glv = got
    cpu = lambda t: t.detach().cpu() if isinstance(t, torch.Tensor) else t
    forced = []
    for g in glv:
        forced.append(torch.zeros(0, dtype=torch.bool) if isinstance(g[0], list) else torch.sigmoid(cpu(g[1])[:, 0]) > 0.5)
    with torch.no_grad():
        (wl, ws), wlv = oracle(locs, feats, forced_keep=forced)
    flips = []
    for i, (w, g) in enumerate(zip(wlv, glv)):
        if isinstance(w[0], list) or isinstance(g[0], list):
            assert isinstance(w[0], list) and isinstance(g[0], list), 'level %d: one side empty' % i
            flips.append(0)
            continue
        w0, w1, g0, g1 = cpu(w[0]), cpu(w[1]), cpu(g[0]), cpu(g[1])
        assert torch.equal(w0, g0), '%s candidate coordinates at level %d' % (tag, i)
        excess = (w1 - g1).abs() - (tol_logit + rtol_logit * w1.abs())
        worst = int(excess.argmax())
        assert float(excess.max()) <= 0, '%s level %d: |%.6f - %.6f| = %.3e beyond %.0e + %.0e * |value|' % (
            tag, i, float(w1.reshape(-1)[worst]), float(g1.reshape(-1)[worst]),
            float((w1 - g1).abs().reshape(-1)[worst]), tol_logit, rtol_logit)
        fl = (torch.sigmoid(w1[:, 0]) > 0.5) != forced[i]
        flips.append(int(fl.sum()))
        assert bool((w1[:, 0][fl].abs() < margin).all()), '%s illegal mask flip at level %d (|logit| >= %g)' % (tag, i, margin)
    assert sum(flips) <= max_flips, '%s %s legal flips' % (tag, flips)
    if isinstance(wl, list) or isinstance(gl, list) or len(wl) == 0:
        assert len(wl) == 0 and len(gl) == 0
        err, n = 0.0, 0
    else:
        wl, ws, gl, gs = cpu(wl), cpu(ws), cpu(gl), cpu(gs)
        assert torch.equal(wl, gl), '%s final coordinates' % tag
        err, n = float((ws - gs).abs().max()), int(wl.shape[0])
        assert err <= tol_sdf, '%s TSDF differs by %.3e' % (tag, err)
    print('%s teacher-forced parity: legal flips per level %s, TSDF max err %.2e over %d voxels' % (tag, flips, err, n))
    return flips
