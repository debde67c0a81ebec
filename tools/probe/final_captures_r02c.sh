# round-2 evidence, last pass (after the bilateral epilogue / odd-row staging changes): tests, the two bench lines,
# the launch list and the wide-kernel ncu capture.  Every ncu pass follows a plain run of the same command.
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r02c_final.log 2>&1; tail -2 gpurun_out/t_r02c_final.log
timeout 900 python bench.py > gpurun_out/bench_r02c_n1.json 2> gpurun_out/bench_r02c_n1.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02c_n1_steps20.json 2> gpurun_out/bench_r02c_n1_steps20.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver --no-4k --no-8k > gpurun_out/plain_launches_c.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02c_1080p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver --no-4k --no-8k > gpurun_out/ncu_l2c.log 2>&1
python tools/prof_one_frame.py 1080 1920 3 1 > gpurun_out/plain_r02cw.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"depth_front|warp_kernel|bilateral|backend|telea_prepare|lanczos|normalize" -s 9 -c 9 -o gpurun_out/prof_r02c_wide -f python tools/prof_one_frame.py 1080 1920 3 1 > gpurun_out/ncu_r02cw.log 2>&1
python tools/prof_one_frame.py 2160 3840 3 1 > gpurun_out/plain_r02ck.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:"depth_front|warp_kernel|bilateral|backend|telea_prepare|telea_march" -s 7 -c 7 -o gpurun_out/prof_r02c_4k -f python tools/prof_one_frame.py 2160 3840 3 1 > gpurun_out/ncu_r02ck.log 2>&1
python tools/kernel_times.py > gpurun_out/kernel_times_r02c.txt 2>&1; head -7 gpurun_out/kernel_times_r02c.txt
