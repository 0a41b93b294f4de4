#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.
  python profiles/summarize.py rep <file.ncu-rep> <out.txt>       key metrics per captured kernel
  python profiles/summarize.py launches <launches.csv> <out.txt> [forwards-per-run kernel-name count]
  python profiles/summarize.py dram <metrics.csv> <out.json> [last-n]  DRAM bytes + duration per captured launch (bench.py's
                                                                  roofline.traffic reads profiles/conv_dram_traffic.json)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def rep(path, out):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none summary of %s (one column per captured launch)\n' % path)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write('%-86s %-16s %s\n' % (k, units[i], ' | '.join(r[i][:44] for r in rows[2:])))


def launches(path, out, per=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        name = re.sub(r'\(.*', '', r['Kernel Name'])[:70]
        t = float(r['Metric Value'].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    nf = 1.0
    if per:
        nf = agg[per[0]][0] / float(per[1])
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none launch list of `python bench.py --steps 1 '
                '--warmup 3` (%s), %d launches, %.1f generator passes; cold-cache serialised times: compare SHARES\n'
                % (path, len(rows), nf))
        f.write('%-72s %9s %14s %7s\n' % ('kernel', 'launches', 'us per pass', 'share'))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-72s %9.1f %14.1f %6.1f%%\n' % (k, n / nf, t / 1e3 / nf, 100 * t / tot))
        f.write('%-72s %9s %14.1f\n' % ('TOTAL', '', tot / 1e3 / nf))


def dram(path, out, last=0):
    import json
    lines = [l for l in open(path) if not l.startswith('==')]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by.setdefault(r['ID'], {'kernel': re.sub(r'\(.*', '', r['Kernel Name'])})
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r['Metric Unit'], 1)
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', '')) * scale
    launches_ = [{'kernel': d['kernel'], 'dram_read_bytes': d.get('dram__bytes_read.sum', 0.0),
                  'dram_write_bytes': d.get('dram__bytes_write.sum', 0.0), 'us': d.get('gpu__time_duration.sum', 0.0)}
                 for d in by.values()]
    if last:            # a first pass that outgrew the arena is re-run by the host mirror: keep the complete pass only
        launches_ = launches_[-last:]
    tot = sum(l['dram_read_bytes'] + l['dram_write_bytes'] for l in launches_)
    # stamp: sha256 over the CUDA sources the capture was made with (bench.py quotes the file only for the same sources)
    import glob
    import hashlib
    import os
    h = hashlib.sha256()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for fn in sorted(glob.glob(os.path.join(root, 'sgnn_b200', 'csrc', '*.cu*'))):
        h.update(open(fn, 'rb').read())
    with open(out, 'w') as f:
        json.dump({'kernel_source_sha256': h.hexdigest(), 'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control '
                             'none over the %d convolution launches of one generator pass (cold cache: ncu flushes L2 '
                             'before every launch)' % len(launches_),
                   'n_launches': len(launches_), 'dram_bytes_total': tot,
                   'dram_bytes_per_launch': tot / max(len(launches_), 1), 'launches': launches_}, f, indent=1)


if __name__ == '__main__':
    if sys.argv[1] == 'rep':
        rep(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == 'dram':
        dram(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4:6] if len(sys.argv) > 5 else None)
