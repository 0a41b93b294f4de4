"""Isolated timing of the unique-row kernels on the generator's own site sets (level-3 kept sites as parents of the child
convolution; the final surface rows for the regular convolution).  Scratch tool."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgnn_b200
import sgnn_b200.engine as E
from sgnn_b200.synth import fill_parameters, synthetic_batch
m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
fill_parameters(m, 0); m = m.cuda().eval(); m.conv_mode = 'exact'
locs, feats = synthetic_batch(32, 64, 0.05)
ones = np.ones(5, dtype=np.float32)
(out_locs, out_sdf), levels = m([locs.cuda(), feats.cuda(), 32], ones)
def timeit(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for i in range(reps + 2):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts[2:]))
which = sys.argv[1] if len(sys.argv) > 1 else 'child'
g = torch.Generator(device='cuda'); g.manual_seed(1)
if which == 'child':
    l = levels[3]
    kept = torch.sigmoid(l[1][:, 0]) > 0.5
    par = l[0][kept].contiguous()                        # parents of the surface level = kept level-3 candidates
    # the child conv of refinement h=2 runs on the kept level-2 candidates
    l2 = levels[2]; par = l2[0][torch.sigmoid(l2[1][:, 0]) > 0.5].contiguous()
    grid = E.build_grid(par, 32, (32, 32, 32))
    nbr = E.rulebook_submanifold(grid)
    n = par.shape[0]
    plan = E.tile_plan(nbr, n)
    x = torch.randn((n, 48), device='cuda', generator=g)
    w = torch.randn((27, 48, 16), device='cuda', generator=g) * 0.05
    out = torch.empty((8 * n, 16), device='cuda')
    s, t = torch.rand(16, device='cuda') + 0.5, torch.rand(16, device='cuda') - 0.5
    us = timeit(lambda: E.conv(x, nbr, w, 8 * n, out, child_mode=True, scale_a=s, shift_a=t, relu_a=True, plan=plan))
    print('child parents %d candidates %d: %.1f us (incl. filter preparation ~3 us)' % (n, 8 * n, us))
else:
    cin = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rows = out_locs.contiguous()
    grid = E.build_grid(rows, 32, (64, 64, 64))
    nbr = E.rulebook_submanifold(grid)
    n = rows.shape[0]
    plan = E.tile_plan(nbr, n)
    ld = (cin + 7) // 8 * 8
    xb = torch.randn((n, ld), device='cuda', generator=g)
    w = torch.randn((27, cin, 16), device='cuda', generator=g) * 0.1
    out = torch.empty((n, 16), device='cuda')
    us = timeit(lambda: E.conv(xb[:, :cin], nbr, w, n, out, plan=plan))
    print('regular rows %d cin %d: %.1f us (incl. filter preparation ~3 us)' % (n, cin, us))

