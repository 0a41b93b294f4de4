"""Synthetic inputs and deterministic parameters: shared by bench.py, the parity tests and
the golden-fixture generator.  No RNG-version dependence for the parameters (integer hash -> fp32)."""
import numpy as np
import torch

def _hash_uniform(n, seed):
    """Platform-independent uniform [-1, 1) floats: integer hash -> 24-bit mantissa (exact in fp32)."""
    i = np.arange(n, dtype=np.uint64) + np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    i ^= i >> np.uint64(33)
    i *= np.uint64(0xFF51AFD7ED558CCD)
    i ^= i >> np.uint64(33)
    i *= np.uint64(0xC4CEB9FE1A85EC53)
    i ^= i >> np.uint64(33)
    return ((i >> np.uint64(40)).astype(np.float64) / float(1 << 23) - 1.0).astype(np.float32)


def fill_parameters(model, seed=0):
    """Deterministic, non-trivial parameters for ANY module tree with the SG-NN state_dict layout (used for
    the golden fixtures, the parity tests and the benchmark; no RNG-version dependence)."""
    sd = model.state_dict()
    with torch.no_grad():
        for idx, (name, t) in enumerate(sd.items()):
            if t.dtype not in (torch.float32, torch.float64):
                continue            # num_batches_tracked
            u = torch.from_numpy(_hash_uniform(t.numel(), seed * 1000 + idx + 1)).view(t.shape)
            leaf = name.rsplit('.', 1)[-1]
            if leaf == 'running_var':
                v = 1.0 + 0.5 * u
            elif leaf == 'running_mean':
                v = 0.2 * u
            elif t.dim() == 1 and leaf == 'weight':      # BatchNorm gamma
                v = 1.0 + 0.5 * u
            elif leaf == 'bias':
                v = 0.2 * u
            else:
                fan_in = t.numel() // t.shape[-1] if t.dim() == 3 else (t[0].numel() if t.dim() > 1 else 1)
                if t.dim() == 5 and 'decode_dense' in name:   # ConvTranspose3d [in, out, k,k,k]
                    fan_in = t.shape[0] * 8
                v = u * float(np.sqrt(6.0 / max(fan_in, 1)))
            t.copy_(v.to(t.dtype))
    model.load_state_dict(sd)
    return model


def synthetic_block(b, size=64, occ=0.05, seed0=1234):
    """SURVEY §8(d) synthetic input: block b -> (coords int64 [n,4] raster order, feats fp32 [n,1])."""
    size = [size] * 3 if np.isscalar(size) else list(size)
    rng = np.random.default_rng(seed0 + b)
    mask = rng.random(tuple(size)) < occ
    c = np.argwhere(mask)
    f = rng.uniform(-3, 3, (c.shape[0], 1)).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([c, np.full((c.shape[0], 1), b, dtype=np.int64)], 1).astype(np.int64)), f


def synthetic_batch(blocks, size=64, occ=0.05, seed0=1234, first=0):
    cs, fs = [], []
    for i in range(blocks):
        c, f = synthetic_block(first + i, size, occ, seed0)
        c[:, 3] = i
        cs.append(c)
        fs.append(f)
    return (torch.from_numpy(np.ascontiguousarray(np.concatenate(cs))),
            torch.from_numpy(np.ascontiguousarray(np.concatenate(fs))))


def synthetic_scene(dims_zyx=(128, 320, 256), seed=0, truncation=3.0):
    """A room-like whole scene: floor, four walls with a doorway, a table slab and a few spheres; the
    signed distance to the nearest surface in voxels (+ a little measurement noise), kept where |sdf| < truncation.
    -> (locs int32 [n,3] (z,y,x) raster order, sdf float32 [n]).  128 x 320 x 256 gives ~1.3 M sites."""
    d0, d1, d2 = (int(v) for v in dims_zyx)
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(d0, dtype=np.float32), np.arange(d1, dtype=np.float32),
                          np.arange(d2, dtype=np.float32), indexing='ij', sparse=True)
    big = np.float32(1e3)
    sd = np.broadcast_to(z - np.float32(4.3), (d0, d1, d2)).copy()                             # floor
    sd = np.minimum(sd, np.where((y > d1 * 0.4) & (y < d1 * 0.5) & (z < d0 * 0.6), big, x - np.float32(5.6)))   # wall + door
    sd = np.minimum(sd, np.float32(d2 - 6.2) - x)
    sd = np.minimum(sd, y - np.float32(4.9))
    sd = np.minimum(sd, np.float32(d1 - 5.4) - y)
    slab = np.maximum(np.maximum(np.abs(z - d0 * 0.35) - 2.2, np.abs(y - d1 * 0.55) - d1 * 0.12), np.abs(x - d2 * 0.5) - d2 * 0.2)
    sd = np.minimum(sd, slab.astype(np.float32))
    for _ in range(6):
        c = rng.uniform([10, 30, 30], [d0 * 0.5, d1 - 30, d2 - 30]).astype(np.float32)
        r = np.float32(rng.uniform(6, 14))
        sd = np.minimum(sd, np.sqrt((z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2) - r)
    keep = np.abs(sd) < truncation
    locs = np.argwhere(keep).astype(np.int32)
    vals = sd[keep].astype(np.float32) + rng.normal(0, 0.02, int(keep.sum())).astype(np.float32)
    return locs, np.clip(vals, -truncation + 1e-3, truncation - 1e-3).astype(np.float32)
