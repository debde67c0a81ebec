#!/bin/bash
# profiling helper: tests + A/B of the experimental builds + one ncu capture of the wide kernels (after the plain run)
cd "$(dirname "$0")/../.."
bash tools/probe/ab_libs.sh "$@"
unset VSC_B200_LIB
python tools/prof_one_frame.py 1080 1920 3 1 > gpurun_out/plain_w3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"depth_front|warp_kernel|bilateral|backend|telea_prepare|lanczos|normalize" -s 9 -c 9 -o gpurun_out/prof_r02_wide_v3 -f python tools/prof_one_frame.py 1080 1920 3 1 > gpurun_out/ncu_w3.log 2>&1
echo "ncu rc=$?"
