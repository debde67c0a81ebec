"""profiling helper (not a test): dynamic opcode mix of one kernel from an ncu report with the source page
usage: python tools/ncu_opmix.py REPORT.ncu-rep KERNEL_REGEX|LAUNCH_NUMBER [top]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sel = ['--kernel-id', ':::' + rx] if rx.isdigit() else ['--kernel-name', 'regex:' + rx]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'] + sel,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
ci = rows[hdr].index('Instructions Executed')
mix, total = collections.Counter(), 0
for r in rows[hdr + 1:]:
    if len(r) <= ci or not r[0].startswith('0x'):
        continue
    ins = r[1].split()
    op = ins[1] if ins[0].startswith('@') else ins[0]
    op = op.split('.')[0].rstrip(';')
    n = int(r[ci]); mix[op] += n; total += n
print(f'{rows[0][1]}: {total / 1e6:.1f} M warp-instructions')
for op, n in mix.most_common(top):
    print(f'  {op:12s} {n / 1e6:8.2f} M  {100 * n / total:5.1f} %')
