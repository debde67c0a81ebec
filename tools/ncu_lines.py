"""profiling helper (not a test): warp-instructions executed per source line of one kernel.
Joins the SASS page of an ncu report (per-instruction execution counts) with `nvdisasm -g` line info of the library
the report was taken from.
usage: python tools/ncu_lines.py REPORT.ncu-rep LAUNCH_NUMBER LIB.so MANGLED_SUBSTRING [top [SOURCE_DIR]]
(SOURCE_DIR: where the sources of that build are, default ../csrc next to the library)"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, launch, lib, sym = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
srcdir = sys.argv[6] if len(sys.argv) > 6 else os.path.join(os.path.dirname(os.path.abspath(lib)), '..', 'csrc')
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-id', ':::' + launch],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
ci = rows[hdr].index('Instructions Executed')
ins = [(int(r[0], 16), r[1].strip(), int(r[ci])) for r in rows[hdr + 1:] if r and r[0].startswith('0x')]
base = ins[0][0]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and sym in l)
line_of, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('//-----'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
per, total = collections.Counter(), 0
for addr, text, n in ins:
    per[line_of.get(addr - base)] += n; total += n
print(f'{rows[0][1]}: {total / 1e6:.1f} M warp-instructions (source-page count)')
src = {}
for (key, n) in per.most_common(top):
    if key is None:
        print(f'{n / 1e6:8.2f} M {100 * n / total:5.1f} %  <no line>'); continue
    f, ln = key
    if f not in src:
        p = os.path.join(srcdir, f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    ln2 = ln + int(os.environ.get('LINE_SHIFT', '0'))       # sources edited above since the build: shift the look-up
    text = src[f][ln2 - 1].strip() if 0 <= ln2 - 1 < len(src[f]) else ''
    print(f'{n / 1e6:8.2f} M {100 * n / total:5.1f} %  {f}:{ln}  {text[:110]}')
