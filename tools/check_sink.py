"""driver check (GPU box): PNG mode vs --raw-sink on the same synthetic workflow; the streams must hold identical frames"""
import os, subprocess, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import numpy as np, cv2
wf, n = '/tmp/wf_sink', int(sys.argv[1]) if len(sys.argv) > 1 else 48
subprocess.check_call([sys.executable, os.path.join(ROOT, 'tools', 'make_workflow.py'), wf, str(n), '1080', '1920'], stdout=subprocess.DEVNULL)
drv = os.path.join(ROOT, 'video-stereo-converter_b200', 'sbs_generator.py')
common = [sys.executable, drv, wf, '--no-interactive', '--gpus', '1', '--slots', '12', '--io-threads', '16']
t0 = time.time(); subprocess.check_call(common, stdout=subprocess.DEVNULL); t_png = time.time() - t0
t0 = time.time(); subprocess.check_call(common + ['--raw-sink', wf + '/sbs.rgb'], stdout=subprocess.DEVNULL); t_raw = time.time() - t0
meta = json.load(open(wf + '/sbs.rgb.json'))
raw = np.memmap(wf + '/sbs.rgb', np.uint8, 'r').reshape(meta['frames'], meta['height'], meta['width'], 3)
bad = 0
for i, fn in enumerate(meta['frame_numbers']):
    png = cv2.cvtColor(cv2.imread(f'{wf}/sbs/sbs_{fn}.png', cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    bad += int(not np.array_equal(png, raw[i]))
print(f'frames {n}: png mode {n / t_png:.1f} img/s (whole process {t_png:.1f} s), raw sink {n / t_raw:.1f} img/s ({t_raw:.1f} s), mismatching frames {bad}')
