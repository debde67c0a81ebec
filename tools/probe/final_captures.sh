timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"depth_front|warp_kernel|bilateral|backend|telea_prepare" -s 6 -c 6 -o gpurun_out/prof_r01_wide -f python tools/prof_one_frame.py 1080 1920 3 1 > gpurun_out/ncu_w.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"telea_cluster_kernel" -s 2 -c 1 -o gpurun_out/prof_r01_march -f python tools/prof_one_frame.py 1080 1920 3 4 > gpurun_out/ncu_m.log 2>&1
VSC_BENCH_WORKLOAD=4k timeout 600 python bench.py --slots 13 --batch 104 --steps 3 --no-cpu-baseline > gpurun_out/bench_r01_4k.json 2> gpurun_out/bench_4k.err
tail -c 300 gpurun_out/bench_4k.err
