import sys, time, torch
sys.path.insert(0, '.')
import sgnn_b200
from sgnn_b200.synth import fill_parameters, synthetic_batch
from sgnn_b200.streaming import StreamingRunner
m = sgnn_b200.GenModel(8, [64, 64, 64], 1, 16, 16, 4, True, True, 1, 1)
fill_parameters(m, 0); m = m.cuda().eval()
ONES = [1.0] * 4
host = []
for s in range(4):
    l, f = synthetic_batch(32, 64, 0.05, seed0=100 + s)
    host.append((l.pin_memory(), f.pin_memory()))
K = 20
def run(r, k):
    prev = None
    r.submit(host[0][0], host[0][1], 32)
    for i in range(k):
        if i + 1 < k:
            r.submit(*host[(i + 1) % 4], 32)
        t = r.step(ONES)
        if prev is not None: r.result(prev)
        prev = t
    r.result(prev)
for oi, oo in [(False, False), (True, False), (False, True), (True, True)]:
    with torch.no_grad():
        r = StreamingRunner(m, 2, oi, oo)
        run(r, 3); torch.cuda.synchronize()
        t0 = time.perf_counter(); run(r, K); torch.cuda.synchronize()
        print('overlap_in', oi, 'overlap_out', oo, 'ms/step', (time.perf_counter() - t0) * 1e3 / K, flush=True)
# direct device-resident
res = [(l.cuda(), f.cuda()) for l, f in host]
with torch.no_grad():
    for i in range(3): m([*res[i % 4], 32], ONES)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K): m([*res[i % 4], 32], ONES)
    torch.cuda.synchronize(); print('resident ms/step', (time.perf_counter() - t0) * 1e3 / K)
