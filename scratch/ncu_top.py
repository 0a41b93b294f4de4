"""Summarise an ncu report: per launch key metrics (raw page) and the top stalled SASS lines of one launch (source page).
usage: python scratch/ncu_top.py report.ncu-rep [launch_index_for_source] [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__inst_executed_op_tma_ld.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_sample_count']
for li, r in enumerate(rows[2:]):
    print('--- launch', li)
    for k in keys:
        if k in ix:
            print('   %-75s %s %s' % (k, r[ix[k]][:80], rows[1][ix[k]]))
if len(sys.argv) > 2:
    li = int(sys.argv[2]); ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(li), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    for i, r in enumerate(rows):
        if r and r[0] == 'Address':
            h = i; break
    hdr = rows[h]; ix = {x: i for i, x in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[ix['# Samples']]) for r in data)
    print('total samples', tot, 'instructions', len(data))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:ntop]
    for i in sorted(top):
        r = data[i]
        print('%5d smp %5s exe %8s long %5s short %5s wait %4s bar %4s  %s' % (i, r[ix['# Samples']], r[ix['Instructions Executed']],
              r[ix['stall_long_sb']], r[ix['stall_short_sb']], r[ix['stall_wait']], r[ix['stall_barrier']], r[ix['Source']].strip()[:100]))
