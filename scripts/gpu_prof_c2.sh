mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_seg1_config2 python bench.py --workload config2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
