"""How much of the convolution's gather is redundant?  For the row order the generator actually produces (kept candidates:
parent order x 8 z-major children, hierarchically from the 8^3 coarse grid), count per 128-row tile the DISTINCT input rows
its 27 filter offsets touch against the number of (row, offset) reads the kernels issue today.  CPU only (oracle generator);
feeds the round-2 plan in DESIGN.md section 7.  Scratch tool."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from genmodel import OracleGenModel                      # noqa: E402
from sgnn_b200.synth import fill_parameters, synthetic_batch   # noqa: E402
from helpers import nbr_table                            # noqa: E402


def analyse(tag, coords, tile=128):
    nbr = nbr_table(coords)                              # [27, n]
    n = coords.shape[0]
    reads = uniq = 0
    lines64 = 0
    for t0 in range(0, n, tile):
        blk = nbr[:, t0:t0 + tile]
        present = blk[blk >= 0]
        reads += present.size
        u = np.unique(present)
        uniq += u.size
        lines64 += np.unique(u // 2).size                # 128-byte lines holding the distinct 64-byte rows
    print('%-34s rows %7d  taps present %.1f/27  reads %9d  distinct per tile %9d  amplification %.2fx  '
          '(128-B lines: %d, %.2f rows/line)' % (tag, n, reads / n, reads, uniq, reads / max(uniq, 1), lines64, uniq / max(lines64, 1)))


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    locs, feats = synthetic_batch(nb, 64, 0.05)
    m = OracleGenModel()
    fill_parameters(m, 0)
    m.eval()
    with torch.no_grad():
        (out_locs, out_sdf), levels = m(locs, feats)
    analyse('input level (5 % iid)', locs.numpy())
    for i, l in enumerate(levels[1:], 1):
        c = l[0].numpy()
        kept = (torch.sigmoid(l[1][:, 0]) > 0.5).numpy()
        analyse('level %d candidates' % i, c)
        analyse('level %d kept (next rows; child-conv parents)' % i, c[kept])
    analyse('surface level rows', out_locs.numpy())
    # the same surface rows in raster order, for comparison
    o = out_locs.numpy()
    order = np.lexsort((o[:, 2], o[:, 1], o[:, 0], o[:, 3]))
    analyse('surface level rows, raster order', o[order])


if __name__ == '__main__':
    main()

