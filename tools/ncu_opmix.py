"""Dynamic opcode mix + per-source-line issue/stall aggregation of one kernel in an .ncu-rep.
usage: python tools/ncu_opmix.py rep.ncu-rep warps_times_iters"""
import csv, collections, re, subprocess, sys, io
rep = sys.argv[1]; W = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
if rows[0][0] == 'Kernel Name': rows = rows[1:]
h = rows[0]
ci = h.index('Source'); ce = h.index('Instructions Executed'); cs = h.index('# Samples')
agg = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r in rows[1:]:
    try: n = int(r[ce]); sm = int(r[cs])
    except Exception: continue
    s = re.sub(r'^@!?U?P\d+\s+', '', r[ci].strip())
    op = s.split()[0].split('.')[0]
    agg[op] += n; smp[op] += sm; tot += n; tots += sm
print('total warp-instructions', tot, ' per unit', tot / W, ' samples', tots)
for k, v in agg.most_common(32):
    print(f'{k:10s} {v:10d} {100*v/tot:5.1f}%  per unit {v/W:7.1f}   samples {100*smp[k]/tots:5.1f}%')
