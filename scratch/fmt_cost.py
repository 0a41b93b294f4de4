"""How much of the step is output formatting?  K steps of (a) GenModel.__call__ (reference-shaped outputs: int64 coordinates,
clones out of the arena), (b) the native generator call alone (raw arena views), same inputs, L2 flush per step."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgnn_b200
from sgnn_b200.synth import fill_parameters, synthetic_batch
m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
fill_parameters(m, 0)
m = m.cuda().eval()
ones = np.ones(5, dtype=np.float32)
sets = [synthetic_batch(32, 64, 0.05, first=32 * s) for s in range(4)]
sets = [(l.cuda(), f.cuda()) for l, f in sets]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
K = 20
def run(fn, flushing=True):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(K):
        if flushing:
            flush.fill_(i & 255)
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K, (time.perf_counter() - t0) * 1e3 / K
full = lambda i: m([sets[i % 4][0], sets[i % 4][1], 32], ones)
g = m._native if m._native is not None else None
full(0)
g = m._native
raw = lambda i: g.forward(sets[i % 4][0], sets[i % 4][1], want_cand_locs=True, nb=32)
raw_nolocs = lambda i: g.forward(sets[i % 4][0], sets[i % 4][1], want_cand_locs=False, nb=32)
print('GenModel call (formatted outputs)      %.3f ms device, %.3f ms wall' % run(full))
print('native call, raw arena views            %.3f ms device, %.3f ms wall' % run(raw))
print('native call, no candidate coordinates   %.3f ms device, %.3f ms wall' % run(raw_nolocs))
print('GenModel call, no L2 flush              %.3f ms device, %.3f ms wall' % run(full, False))
