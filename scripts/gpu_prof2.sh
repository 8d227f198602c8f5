mkdir -p gpurun_out
for s in 1 4; do
FRX_SEG=$s timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_seg${s}_config2 python bench.py --workload config2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_seg${s}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
