mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 9 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload config3 --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -o '"void frx[^"]*\|"frx_[a-z_]*[^"]*\|"ns","[0-9]*"' gpurun_out/launches_c3.csv | paste - - | head -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_obstacle_kernel -s 4 -c 1 -f -o gpurun_out/prof_obsB python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
