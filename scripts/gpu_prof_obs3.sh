# one full ncu capture of frx_obstacle_kernel (and the finish kernel) on config 3 (200,000 rows, 20 obstacles, 31 samples)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_obstacle -s 6 -c 2 -f -o gpurun_out/r02_prof_obstacle_config3 python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_o3.log 2>&1
ls -la gpurun_out/*.ncu-rep
