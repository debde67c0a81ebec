# round-2 evidence: run on a B200 box from the repo root (gpurun -- 'bash tools/probe/final_captures_r02.sh')
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r02_final.log 2>&1; tail -2 gpurun_out/t_r02_final.log
timeout 900 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver --no-4k > gpurun_out/plain_launches.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_1080p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-driver --no-4k > gpurun_out/ncu_l2.log 2>&1
for S in 12 14; do timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-driver --slots-4k $S > gpurun_out/bench_r02_4k_s$S.json 2> gpurun_out/bench_r02_4k_s$S.err; done
