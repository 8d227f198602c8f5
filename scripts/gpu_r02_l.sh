# scripts/gpu_r02_l.sh: LDS record path; parity, block shape A/B, default bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
for wide in 0 1; do
  for wl in config5 config3; do
    FRX_OBS_WIDE=$wide timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-also --steps 10 > gpurun_out/sweep_wide${wide}_${wl}.json 2> gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_wide*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), d["roofline"].get("kernel"), round(d["roofline"].get("kernel_ms", 0), 4))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
