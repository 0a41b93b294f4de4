"""N generator passes over BASELINE configs[1] (32 x 64^3 @5 %) in a given conv mode -- the workload for ncu captures
(bench.py adds ledger / profiling / e2e passes that make launch counting awkward).  Scratch tool."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgnn_b200                                        # noqa: E402
from sgnn_b200.synth import fill_parameters, synthetic_batch   # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'exact'
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
fill_parameters(m, 0)
m = m.cuda().eval()
m.conv_mode = mode
locs, feats = synthetic_batch(32, 64, 0.05)
locs, feats = locs.cuda(), feats.cuda()
ones = np.ones(5, dtype=np.float32)
for _ in range(passes):
    out = m([locs, feats, 32], ones)
torch.cuda.synchronize()
print('mode', mode, 'passes', passes, 'out voxels', int(out[0][0].shape[0]))
