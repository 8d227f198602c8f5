set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json
timeout 600 python bench.py --workload config3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_eval_kernel -s 4 -c 1 -o gpurun_out/prof_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_eval_kernel -s 4 -c 1 -o gpurun_out/prof_c3 python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out
