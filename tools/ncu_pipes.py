"""profiling helper: one line of pipe utilisation per kernel of an .ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'us'), ('smsp__inst_executed.sum', 'inst'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu%'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64%'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smemwf%'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'),
        ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB')]
idx = [h.index(w) if w in h else -1 for w, _ in want]
print(' '.join(f'{n:>9s}' for _, n in want))
for row in r[2:]:
    cells = []
    for (w, n), i in zip(want, idx):
        v = row[i] if i >= 0 else '-'
        if n == 'kernel': v = v.split('(')[0][-24:]
        else:
            try: v = f'{float(v):.1f}' if float(v) < 1e6 else f'{float(v)/1e6:.1f}M'
            except ValueError: pass
        cells.append(f'{v:>9s}')
    print(' '.join(cells))
