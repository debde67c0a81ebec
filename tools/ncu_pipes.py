"""profiling helper: one line of pipe utilisation / traffic per kernel of an .ncu-rep (ncu --set full capture)
usage: python tools/ncu_pipes.py <report.ncu-rep>"""
import csv, subprocess, sys, io
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, units = r[0], r[1]
SCALE = {'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'byte': 1e-6, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6}
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'ms'), ('smsp__inst_executed.sum', 'Minst'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu%'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64%'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'),
        ('dram__bytes_read.sum', 'dramR_MB'), ('dram__bytes_write.sum', 'dramW_MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%')]
idx = [h.index(w) if w in h else -1 for w, _ in want]
print(' '.join(f'{n:>9s}' for _, n in want))
for row in r[2:]:
    cells = []
    for (w, n), i in zip(want, idx):
        v = row[i] if i >= 0 else '-'
        if n == 'kernel':
            v = v.split('(')[0].replace('void ', '')[-26:]
            cells.append(f'{v:>26s}')
            continue
        try:
            f = float(v) * SCALE.get(units[i], 1.0)
            if n == 'Minst': f /= 1e6
            v = f'{f:.3f}' if n in ('ms', 'dramR_MB', 'dramW_MB') else f'{f:.1f}'
        except ValueError:
            pass
        cells.append(f'{v:>9s}')
    print(' '.join(cells))
