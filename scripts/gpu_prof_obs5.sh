# one full ncu capture of frx_obstacle_kernel on a 1/8-size config 5 (1.25 M rows, 50 obstacles, 51 samples)
mkdir -p gpurun_out
FRX_BENCH_C5_V=56 timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_obstacle_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_obstacle_config5s python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_o5.log 2>&1
ls -la gpurun_out/*.ncu-rep
