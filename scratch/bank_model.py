"""Shared-memory wavefront model of the unique-row convolution's tap gather: per quarter-warp (8 consecutive output rows)
and filter offset, the LDS.128 of the 8 lanes needs max-multiplicity(unit mod 8) wavefronts (same address = broadcast).
Compares local-index -> 16-byte-unit mappings.  CPU only; scratch tool."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from genmodel import OracleGenModel
from sgnn_b200.synth import fill_parameters, synthetic_batch
from helpers import nbr_table

US = 512
PRED = False
def wavefronts(units):           # units [8] ints (unit index), -1 absent -> zero row at unit US (or predicated off)
    u = units[units >= 0] if PRED else np.where(units >= 0, units, US)
    if u.size == 0:
        return 0
    uniq = np.unique(u)
    return np.bincount(uniq % 8, minlength=8).max()

def analyse(tag, coords, maps):
    nbr = nbr_table(coords); n = coords.shape[0]
    tot = {m: 0 for m in maps}; ideal = 0
    for t0 in range(0, n - 127, 128):
        blk = nbr[:, t0:t0 + 128]
        u = np.unique(blk[blk >= 0])
        l = np.where(blk >= 0, np.searchsorted(u, np.maximum(blk, 0)), -1)
        for k in range(27):
            for q in range(16):
                lanes = l[k, 8 * q:8 * q + 8]
                ideal += 1
                for name, f in maps.items():
                    tot[name] += wavefronts(np.where(lanes >= 0, f(lanes), -1))
    print(tag, 'quarter-phases', ideal, {k: round(v / ideal, 3) for k, v in tot.items()})

maps = {
    'identity': lambda l: l,
    'xor_hi3': lambda l: l ^ ((l >> 3) & 7),
    'xor_hi6': lambda l: l ^ ((l >> 3) & 7) ^ ((l >> 6) & 7),
    'mul5': lambda l: (l * 5) & 1023,
}
locs, feats = synthetic_batch(2, 64, 0.05)
m = OracleGenModel(); fill_parameters(m, 0); m.eval()
with torch.no_grad():
    (out_locs, out_sdf), levels = m(locs, feats)
analyse('surface rows', out_locs.numpy()[:20000], maps)
PRED = True
analyse('surface rows, absent lanes predicated off', out_locs.numpy()[:20000], {'identity': maps['identity']})
PRED = False
c = levels[2][0].numpy(); kept = (torch.sigmoid(levels[2][1][:, 0]) > 0.5).numpy()
analyse('level 2 kept', c[kept], maps)
