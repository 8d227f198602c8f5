# scripts/gpu_prof.sh <tag> [workloads]: full ncu capture of the eval kernel for each workload (default config2 config3)
set -x
tag=$1; shift
WL="${@:-config2 config3}"
mkdir -p gpurun_out
for wl in $WL; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_${tag}_${wl} python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${tag}_${wl}.log 2>&1
done
ls -la gpurun_out
