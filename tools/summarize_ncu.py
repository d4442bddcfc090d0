"""Summarise an .ncu-rep into profiles/<name>.summary.txt (key raw metrics + hottest instructions by stall samples)."""
import csv, io, subprocess, sys, os, json

rep = sys.argv[1]
name = os.path.splitext(os.path.basename(rep))[0]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', name + '.summary.txt')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
lines = ['# ncu summary of %s' % os.path.basename(rep), '']
summary = {}
for i, h in enumerate(hdr):
    if h in want:
        lines.append('%-70s %-14s %s' % (h, units[i], vals[i]))
        summary[h] = vals[i]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    isrc, isamp, iex = sh.index('Source'), sh.index('# Samples'), sh.index('Instructions Executed')
    stall_cols = [(n, sh.index(n)) for n in sh if n.startswith('stall_') and '(' not in n]
    data = [r for r in srows[2:] if len(r) > isamp and r[isamp].isdigit()]
    tot = sum(int(r[isamp]) for r in data) or 1
    agg = {n: sum(int(r[i]) for r in data) for n, i in stall_cols}
    lines += ['', '# warp-stall samples by reason (all instructions): total %d' % tot]
    for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        lines.append('  %-26s %8d  %5.1f%%' % (n, v, 100.0 * v / tot))
    lines += ['', '# hottest instructions (samples, %, executions, SASS)']
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:25]:
        lines.append('  %7d %5.1f%% ex=%-10s %s' % (int(r[isamp]), 100.0 * int(r[isamp]) / tot, r[iex], r[isrc].strip()[:90]))
open(out, 'w').write('\n'.join(lines) + '\n')
print('wrote', out)
