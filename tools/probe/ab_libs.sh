#!/bin/bash
# profiling helper: A/B of library builds on one box.  usage: ab_libs.sh OUTTAG lib1.so lib2.so ...
# (paths relative to video-stereo-converter_b200/lib); the first library also runs the GPU test-suite when TESTS=1
cd "$(dirname "$0")/../.."
tag=$1; shift
mkdir -p gpurun_out
if [ "${TESTS:-1}" = 1 ]; then
  VSC_B200_LIB=$PWD/video-stereo-converter_b200/lib/$1 timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
  echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
fi
for lib in "$@"; do
  export VSC_B200_LIB=$PWD/video-stereo-converter_b200/lib/$lib
  echo "== $lib"
  timeout 120 python tools/kernel_times.py 2>&1 | tee gpurun_out/${tag}_${lib%.so}_kt.txt | head -12
  timeout 400 python bench.py --steps ${STEPS:-6} --warmup 3 --no-8k --no-driver --no-cpu-baseline ${BENCH_EXTRA} > gpurun_out/${tag}_${lib%.so}_bench.json 2> gpurun_out/${tag}_${lib%.so}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_${lib%.so}_bench.json').read().strip().splitlines()[-1])
    w4 = d.get('workloads', {}).get('4k', {})
    print('bench', d['value'], d['e2e']['value'], '4k', w4.get('value'), w4.get('e2e', {}).get('value') if isinstance(w4.get('e2e'), dict) else w4.get('e2e'))
except Exception as e:
    print('bench failed', e)
PY
done
