"""In-situ phase times of the generator pass at BASELINE configs[1] (SGNN_GEN_PHASES: one CUDA event per phase boundary;
launch gaps, memsets and host reads are inside the phase they belong to).  Mean of N passes over rotating inputs, L2 flushed
before each.  -> gpurun_out/<tag>_phases.txt"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgnn_b200
from sgnn_b200.synth import fill_parameters, synthetic_batch

mode = sys.argv[1] if len(sys.argv) > 1 else 'tc32'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
fill_parameters(m, 0)
m = m.cuda().eval()
m.conv_mode = mode
ones = np.ones(5, dtype=np.float32)
sets = [synthetic_batch(32, 64, 0.05, first=32 * s) for s in range(4)]
sets = [(l.cuda(), f.cuda()) for l, f in sets]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for i in range(3):
    m([sets[i % 4][0], sets[i % 4][1], 32], ones)
m._native.phases = True
acc, order = {}, []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for i in range(N):
    flush.fill_(i)
    e0.record()
    m([sets[i % 4][0], sets[i % 4][1], 32], ones)
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
    for name, ms in m._native.phase_table():
        if name not in acc:
            acc[name] = 0.0
            order.append(name)
        acc[name] += ms
print('# phases of one generator pass, %s mode, mean of %d passes; whole call incl. output formatting: %.3f ms' % (mode, N, tot / N))
s = 0.0
for name in order:
    print('%-52s %8.1f us' % (name, acc[name] / N * 1e3))
    s += acc[name] / N
print('%-52s %8.1f us' % ('SUM of phases (C-ABI call, first to last mark)', s * 1e3))
