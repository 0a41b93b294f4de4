"""Randomised interleaving simulation of conv_ur_kernel's mbarrier protocol (stage / accumulator barriers), with the real
parity semantics of mbarrier.try_wait.parity (a waiter two phases late blocks forever).  CPU only; scratch tool."""
import random, sys

class MBar:
    def __init__(self, count): self.count, self.pending, self.phase = count, count, 0
    def arrive(self):
        self.pending -= 1
        if self.pending == 0: self.pending = self.count; self.phase += 1
    def test(self, parity): return (self.phase & 1) != parity     # true iff the phase with that parity has completed

NG = int(sys.argv[3]) if len(sys.argv) > 3 else 2
def sim(Q, tiles_passes, seed):
    KG = 3 if Q == 1 else 1
    NI = (27 + KG - 1) // KG
    ST = KG * Q * 24
    NST = NST_OVERRIDE or 256 // ST
    upp = lambda s: (NI - s + NST - 1) // NST
    rnd = random.Random(seed)
    st_full = [MBar(4) for _ in range(NST)]; st_empty = [MBar(1) for _ in range(NST)]
    acc_full = [MBar(1), MBar(1)]; acc_empty = [MBar(4), MBar(4)]
    def producer(g):
        tp = 0
        for tl, npass in enumerate(tiles_passes):
            for p in range(npass):
                for i in range(NI):
                    G = tp * NI + i
                    if G % NG != g: continue
                    s, n = (G % NST, G // NST) if not PER_PASS else (i % NST, tp * upp(i % NST) + i // NST)
                    if n > 0:
                        while not st_empty[s].test((n - 1) & 1): yield ('st_empty', s, n)
                        if st_empty[s].phase < n: print('  ALIAS: producer g%d passes st_empty[%d] for use %d at phase %d (tp %d item %d)' % (g, s, n, st_empty[s].phase, tp, i))
                    yield None
                    st_full[s].arrive()
                tp += 1
    def mma():
        tp = 0
        for tl, npass in enumerate(tiles_passes):
            ab = tl & 1
            if tl >= 2:
                while not acc_empty[ab].test(((tl >> 1) - 1) & 1): yield ('acc_empty', ab, tl)
            for p in range(npass):
                for i in range(NI):
                    G = tp * NI + i
                    s, n = (G % NST, G // NST) if not PER_PASS else (i % NST, tp * upp(i % NST) + i // NST)
                    while not st_full[s].test(n & 1): yield ('st_full', s, n)
                    if st_full[s].phase < n + 1: print('  ALIAS: mma passes st_full[%d] for use %d at phase %d' % (s, n, st_full[s].phase))
                    yield None
                    st_empty[s].arrive()
                    if p == npass - 1 and i == NI - 1: acc_full[ab].arrive()
                tp += 1
    def epi():
        for tl, npass in enumerate(tiles_passes):
            ab = tl & 1
            while not acc_full[ab].test((tl >> 1) & 1): yield ('acc_full', ab, tl)
            yield None
            acc_empty[ab].arrive()
    # each producer warp is its own agent (4 per group), 4 epilogue warps
    agents = [producer(g) for g in range(NG) for _ in range(4)] + [mma()] + [epi() for _ in range(4)]
    names = ['p%d.%d' % (g, w) for g in range(NG) for w in range(4)] + ['mma'] + ['epi%d' % w for w in range(4)]
    alive = list(range(len(agents)))
    last = {}
    idle = 0
    while alive:
        k = rnd.choice(alive)
        try:
            r = next(agents[k])
            last[k] = r
            idle = idle + 1 if r is not None else 0
            if idle > 20000:
                return 'DEADLOCK ' + str({names[a]: last.get(a) for a in alive})
        except StopIteration:
            alive.remove(k)
    return 'ok'

NST_OVERRIDE = int(sys.argv[1]) if len(sys.argv) > 1 else 0
PER_PASS = len(sys.argv) > 2      # the buggy per-pass stage numbering
for Q in (1, 2):
    for seed in range(200):
        rnd = random.Random(seed)
        tp = [rnd.choice([1, 1, 1, 2]) for _ in range(12)]
        r = sim(Q, tp, seed)
        if r != 'ok':
            print('Q', Q, 'seed', seed, tp, r); break
    else:
        print('Q', Q, 'all ok')
