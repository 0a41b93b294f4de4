"""Attribute ncu pc-sampling of the unique-row kernels to warp roles: the slow-path waits are one out-of-line function per
role (mb_wait_slow<ROLE>); everything else is attributed by the role branch the instruction lies in.  Scratch tool."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]; li = sys.argv[2] if len(sys.argv) > 2 else '0'
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', li, '--launch-count', '1', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[h]; ix = {x: i for i, x in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']]) for r in data)
# find function boundaries by RET / EXIT: print cumulative samples in windows split at 'RET' instructions
seg, acc, start = [], 0, 0
for i, r in enumerate(data):
    acc += int(r[ix['# Samples']])
    s = r[ix['Source']]
    if 'RET.' in s or ' EXIT' in s and 'P0' not in s:
        seg.append((start, i, acc)); acc = 0; start = i + 1
seg.append((start, len(data) - 1, acc))
print('total', tot)
for a, b, n in seg:
    if n:
        top = sorted(range(a, b + 1), key=lambda i: -int(data[i][ix['# Samples']]))[:3]
        print('instr %5d-%5d samples %6d  top: %s' % (a, b, n, '; '.join('%s(%s)' % (data[i][ix['Source']].strip()[:40], data[i][ix['# Samples']]) for i in top)))
