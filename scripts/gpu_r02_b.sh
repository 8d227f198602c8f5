# scripts/gpu_r02_b.sh: parity suite + bench + launch lists + full captures (obstacle kernel config3, eval kernel config5)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/r02_launches_config3.csv python bench.py --workload config3 --steps 4 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_l3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_obstacle_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_obstacle_config3 python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_o3.log 2>&1
FRX_BENCH_C5_V=56 timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_eval_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_eval_config5s python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_e5.log 2>&1
ls -la gpurun_out | head -40
