# scripts/gpu_r02_i.sh: prediction records staged in shared memory -- parity, default bench, A/B against the L1 path
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
FRX_OBS_STAGE=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_nostage.json 2> gpurun_out/bench_nostage.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench.json", "gpurun_out/r02_bench_nostage.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), [(a["config"]["workload"], a["ms_per_step"], a.get("roofline", {}).get("kernel_ms")) for a in d.get("also", []) if isinstance(a, dict)])
PY
