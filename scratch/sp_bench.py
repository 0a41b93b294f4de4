"""A/B of sgnn_conv_forward's two dense-table kernels (flags 2 = lane-per-(row, channel group) kernel of conv_sp.cu, 4 = row-owner
kernel of conv.cu) and of the compact-rulebook kernel, on the generator's own encoder site sets (32 x 64^3 @5 %).
CUDA events around 20 back-to-back launches after 3 warm-ups, 256 MiB L2 flush before each timed batch."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sgnn_b200.engine as E
from sgnn_b200.synth import synthetic_batch

dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    flush.fill_(1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def run(tag, x, tbl, w, n_out, **kw):
    out = torch.empty((n_out, w.shape[2]), device=dev)
    res = []
    for flags in (2, 4):
        res.append(timeit(lambda: E.conv(x, tbl, w, n_out, out, flags=flags, **kw)))
    print('%-34s rows %8d  rowlane %7.1f us   row-owner %7.1f us' % (tag, n_out, res[0], res[1]), flush=True)


def main():
    blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    locs, _ = synthetic_batch(blocks, 64, 0.05)
    g0 = E.build_grid(locs.to(dev), blocks, (64, 64, 64))
    grids = [g0]
    for _ in range(5):
        grids.append(E.coarsen(grids[-1]))
    chans = [8, 12, 16, 16, 16, 16]
    rng = np.random.default_rng(0)
    for l, g in enumerate(grids[:-1]):
        c = chans[l]
        n = g.n
        x = torch.randn((n, c), device=dev)
        w = torch.randn((27, c, c), device=dev) * 0.1
        nbr = E.rulebook_submanifold(g)
        run('L%d submanifold %d->%d K=27' % (l, c, c), x, nbr, w, n)
        if l == 0:
            comp = E.rulebook_submanifold_compact(g)
            out = torch.empty((n, c), device=dev)
            t = timeit(lambda: E.conv(x, None, w, n, out, compact=comp))
            print('%-34s rows %8d  compact %7.1f us   (%.2f taps/row)' % ('L0 compact rulebook', n, t, float(comp[1].float().mean())), flush=True)
            tr = timeit(lambda: E.rulebook_submanifold(g))
            tc = timeit(lambda: E.rulebook_submanifold_compact(g))
            print('rulebook dense %.1f us   compact %.1f us' % (tr, tc))
        cg = grids[l + 1]
        parent, children = E.rulebook_strided(g, cg)
        w8 = torch.randn((8, c, c), device=dev) * 0.1
        run('L%d->L%d stride-2 %d->%d K=8' % (l, l + 1, c, c), x, children, w8, cg.n)


def wide():
    """first convolution of a refinement level (Cin = 34 / 30 / 26 joined rows, ld = 40 / 32): row-lane vs row-owner"""
    for size, occ, cin in ((16, 0.11, 34), (32, 0.035, 34), (32, 0.15, 30)):
        locs, _ = synthetic_batch(32, size, occ)
        g = E.build_grid(locs.to(dev), 32, (size, size, size))
        nbr = E.rulebook_submanifold(g)
        ld = (cin + 7) // 8 * 8
        x = torch.randn((g.n, ld), device=dev)[:, :cin]
        w = torch.randn((27, cin, 16), device=dev) * 0.1
        run('%d^3 @%.3f  %d->16 K=27' % (size, occ, cin), x, nbr, w, g.n)


if __name__ == '__main__':
    wide()
    main()
