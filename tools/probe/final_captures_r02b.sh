# round-2 evidence after the instruction cuts: run on a B200 box from the repo root
# (gpurun -- 'bash tools/probe/final_captures_r02b.sh').  Every ncu pass follows a plain run of the same command.
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r02b_final.log 2>&1; tail -2 gpurun_out/t_r02b_final.log
timeout 900 python bench.py > gpurun_out/bench_r02b_n1.json 2> gpurun_out/bench_r02b_n1.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b_n1_steps20.json 2> gpurun_out/bench_r02b_n1_steps20.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver --no-4k --no-8k > gpurun_out/plain_launches_b.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02b_1080p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver --no-4k --no-8k > gpurun_out/ncu_l2b.log 2>&1
python tools/prof_one_frame.py 1080 1920 3 1 > gpurun_out/plain_r02bw.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"depth_front|warp_kernel|bilateral|backend|telea_prepare|lanczos|normalize" -s 9 -c 9 -o gpurun_out/prof_r02b_wide -f python tools/prof_one_frame.py 1080 1920 3 1 > gpurun_out/ncu_r02bw.log 2>&1
python tools/prof_one_frame.py 1080 1920 3 4 > gpurun_out/plain_r02bm.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"telea_march_kernel" -s 2 -c 1 -o gpurun_out/prof_r02b_march -f python tools/prof_one_frame.py 1080 1920 3 4 > gpurun_out/ncu_r02bm.log 2>&1
python tools/prof_one_frame.py 2160 3840 3 1 > gpurun_out/plain_r02bk.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:"depth_front|warp_kernel|bilateral|backend|telea_prepare|telea_march" -s 7 -c 7 -o gpurun_out/prof_r02b_4k -f python tools/prof_one_frame.py 2160 3840 3 1 > gpurun_out/ncu_r02bk.log 2>&1
python tools/soak_determinism.py 4 > gpurun_out/soak_r02b.txt 2>&1; tail -1 gpurun_out/soak_r02b.txt
python tools/kernel_times.py > gpurun_out/kernel_times_r02b.txt 2>&1
