# scripts/gpu_final.sh: the numbers and captures that go into profiles/ (run under gpurun, 1 GPU)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r01_bench_config2.json 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --workload config3 > gpurun_out/r01_bench_config3.json 2> gpurun_out/bench_c3.err
timeout 600 python bench.py --workload config4 > gpurun_out/r01_bench_config4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_config5.json 2> gpurun_out/bench_c5.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r01_launches_config2.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
for wl in config2 config3; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_eval_kernel -s 4 -c 1 -f -o gpurun_out/r01_prof_$wl python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$wl.log 2>&1
done
ls -la gpurun_out | head -40
