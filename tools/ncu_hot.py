"""Per-instruction warp-stall samples of one kernel from an ncu report (source page), in program order for an address
range or as a top list.  usage: ncu_hot.py report.ncu-rep [--range lo hi (hex)] [--top N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if 'Address' in r and 'Source' in r)
data = rows[rows.index(hdr) + 1:]
ia, isrc, iss, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
tot = sum(int(r[iss] or 0) for r in data)
base = int(data[0][ia], 16) if data else 0
if '--range' in sys.argv:
    k = sys.argv.index('--range'); lo, hi = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
    acc = 0
    for r in data:
        a = int(r[ia], 16) - base
        if lo <= a <= hi:
            acc += int(r[iss] or 0)
            print('%05x %6s %5.1f%% %9s  %s' % (a, r[iss], 100.0 * int(r[iss] or 0) / max(tot, 1), r[iex], r[isrc][:100]))
    print('range samples %d of %d (%.1f%%)' % (acc, tot, 100.0 * acc / max(tot, 1)))
else:
    n = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 25
    print('total samples', tot)
    for r in sorted(data, key=lambda r: -int(r[iss] or 0))[:n]:
        print('%05x %6s %5.1f%% %9s  %s' % (int(r[ia], 16) - base, r[iss], 100.0 * int(r[iss]) / max(tot, 1), r[iex], r[isrc][:100]))
