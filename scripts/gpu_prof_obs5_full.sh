# one ncu capture of frx_obstacle_kernel<0,512,0> on the FULL config 5 (10,012,800 rows, 50 obstacles, 51 samples)
mkdir -p gpurun_out
timeout 200 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --clock-control none --import-source on -k regex:frx_obstacle_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_obstacle_config5_full python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_o5f.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_o5f.log; ls -la gpurun_out/r02_prof_obstacle_config5_full.ncu-rep
