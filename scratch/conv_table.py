"""Per-convolution CUDA-event times of one generator pass (SGNN_GEN_PROFILE) at BASELINE configs[1], for A/B runs of the
convolution kernels: python scratch/conv_table.py [ur_min_rows] [tc32_min_rows].  Scratch tool."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgnn_b200                                        # noqa: E402
from sgnn_b200._lib import lib                          # noqa: E402
from sgnn_b200.synth import fill_parameters, synthetic_batch   # noqa: E402

ur_min = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tc_min = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
fill_parameters(m, 0)
m = m.cuda().eval()
m.conv_mode = 'tc32'
m.ur_min_rows, m.tc32_min_rows = ur_min, tc_min
locs, feats = synthetic_batch(32, 64, 0.05)
locs, feats = locs.cuda(), feats.cuda()
ones = np.ones(5, dtype=np.float32)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for _ in range(3):
    out = m([locs, feats, 32], ones)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(10):
    flush.fill_(i)
    e0.record()
    out = m([locs, feats, 32], ones)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print('ur_min_rows %d tc32_min_rows %d: ms/pass median %.3f min %.3f  out voxels %d' % (
    ur_min, tc_min, float(np.median(ts)), min(ts), int(out[0][0].shape[0])))
m._native.profile = True
flush.fill_(1)
out = m([locs, feats, 32], ones)
torch.cuda.synchronize()
rec6, rms = (C.c_int64 * 6)(), C.c_float(0)
tot = 0.0
print('   n_out cin  co  K child tc |   us')
for ci in range(int(m._native.last.n_conv)):
    if lib.sgnn_generator_profile_entry(ci, rec6, C.byref(rms)) != 0:
        break
    tot += rms.value * 1e3
    print('%8d %3d %3d %2d %5d %2d | %7.1f' % (rec6[0], rec6[1], rec6[2], rec6[3], rec6[4], rec6[5], rms.value * 1e3))
print('TOTAL conv us %.1f' % tot)
