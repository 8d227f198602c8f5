# scripts/gpu_r02_k.sh: pipelined prediction step; block shape A/B (2 x 256 threads, 84 KB staged | 1 x 512, everything staged)
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
for v in base v512; do
  lib=$PWD/frenetix_motion_planner_b200/libfrx_b200.so; [ $v = v512 ] && lib=$PWD/frenetix_motion_planner_b200/libfrx_b200_v512.so
  for kb in 0 84 160; do
    for wl in config5 config3; do
      FRX_LIB=$lib FRX_OBS_STAGE_KB=$kb timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-also --steps 10 > gpurun_out/sweep_${v}_${wl}_${kb}.json 2> gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
    done
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*_*_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), d["roofline"].get("kernel"), round(d["roofline"].get("kernel_ms", 0), 4))
    except Exception as e:
        print(f, "unreadable", e)
PY
