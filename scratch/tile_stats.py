"""Per 128-row tile: distinct input rows (U) and row-id span of the 27-neighbour table, for the tile-plan prepass
(bitmap de-duplication) and the staging capacity of the unique-row convolution.  CPU only; scratch tool."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from genmodel import OracleGenModel
from sgnn_b200.synth import fill_parameters, synthetic_batch
from helpers import nbr_table

def analyse(tag, coords, tile=128):
    nbr = nbr_table(coords)
    n = coords.shape[0]
    us, spans, runs = [], [], []
    for t0 in range(0, n, tile):
        blk = nbr[:, t0:t0 + tile]
        present = blk[blk >= 0]
        u = np.unique(present)
        us.append(u.size); spans.append(int(u.max() - u.min() + 1) if u.size else 0)
        runs.append(int((np.diff(u) != 1).sum()) + 1 if u.size else 0)
    us = np.array(us); spans = np.array(spans); runs = np.array(runs)
    q = lambda a: ' '.join('%d' % np.percentile(a, p) for p in (50, 90, 99, 100))
    print('%-30s rows %7d tiles %5d  U p50/90/99/max: %s   span p50/90/99/max: %s  runs p50/90/99/max: %s  frac U>256: %.3f >384: %.3f >512: %.3f' % (
        tag, n, len(us), q(us), q(spans), q(runs), (us > 256).mean(), (us > 384).mean(), (us > 512).mean()))

def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    locs, feats = synthetic_batch(nb, 64, 0.05)
    m = OracleGenModel(); fill_parameters(m, 0); m.eval()
    with torch.no_grad():
        (out_locs, out_sdf), levels = m(locs, feats)
    analyse('input level (5 % iid)', locs.numpy())
    for i, l in enumerate(levels[1:], 1):
        c = l[0].numpy()
        kept = (torch.sigmoid(l[1][:, 0]) > 0.5).numpy()
        analyse('level %d kept' % i, c[kept])
    analyse('surface rows', out_locs.numpy())
main()
