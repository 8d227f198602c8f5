# scripts/gpu_r02_q.sh: what goes into profiles/ for the round -- full GPU suite, default bench line, ncu launch lists of the
# bench command (config 5 headline, config 3), full captures of the obstacle kernel (1/8-size config 5, config 3)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_config5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_l5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/r02_launches_config3.csv python bench.py --workload config3 --steps 4 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_l3.log 2>&1
bash scripts/gpu_prof_obs3.sh
bash scripts/gpu_prof_obs5.sh
ls -la gpurun_out | head -40
