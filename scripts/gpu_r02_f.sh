# scripts/gpu_r02_f.sh: full GPU suite (incl. hypothesis properties), planner-level latency, sanitizer runs
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python - > gpurun_out/planner_e2e.json 2> gpurun_out/planner_e2e.err <<'PY'
import json, bench
print(json.dumps(bench.planner_e2e(0), indent=1))
PY
cat gpurun_out/planner_e2e.json; tail -3 gpurun_out/planner_e2e.err
bash scripts/gpu_sanitize.sh
