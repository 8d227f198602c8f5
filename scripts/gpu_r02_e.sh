# scripts/gpu_r02_e.sh: eval-kernel occupancy variants (config2 / config3 / a 1.25M-row config5 shard) + the full default bench
set -x
mkdir -p gpurun_out
bash scripts/gpu_sweep2.sh "1" config2 config3 > /dev/null 2>&1; cp gpurun_out/sweep2.txt gpurun_out/r02_sweep_variants.txt
for lib in build/var/libfrx_*.so; do FRX_BENCH_C5_V=56 FRX_LIB=$PWD/$lib timeout 300 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib config5/8', d['kernels_ms'])" >> gpurun_out/r02_sweep_variants.txt; done
cat gpurun_out/r02_sweep_variants.txt
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
