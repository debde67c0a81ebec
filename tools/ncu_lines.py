"""profiling helper: per-CUDA-source-line executed warp instructions of one kernel from an .ncu-rep
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex | launch number> [min_pct]
(the report must have been taken with --import-source on; the source text comes from the report itself)"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
sel = ['--kernel-id', ':::' + rx] if rx.isdigit() else ['--kernel-name', f'regex:{rx}']
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'] + sel,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, lines, seen_fn = None, [], 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No': hdr = r; ie = r.index('Instructions Executed'); continue
    if r[0].isdigit() and len(r) > ie and r[2] == '-' and r[ie].isdigit():
        lines.append((fname, int(r[0]), r[1], int(r[ie])))
# several kernel instances are concatenated: keep the first occurrence of each (file,line)
agg = {}
for f, l, s, n in lines:
    agg.setdefault((f, l), [s, n])
tot = sum(v[1] for v in agg.values())
print('total warp instructions', tot)
for (f, l), (s, n) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if n >= tot * minpct / 100: print(f'{n:11d} {100*n/tot:5.1f}%  {f}:{l:<4d} {s.strip()[:120]}')
