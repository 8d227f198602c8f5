# scripts/gpu_r02_g.sh: new parity tests (collision probability, wall cull, winner record) + memcheck
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_planner.py tests/test_reference_dropin.py tests/test_gpu_exchange.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
SEL="test_device_matches_oracle_on_golden_inputs or test_split_obstacle_kernel_equals_fused_pass or test_step_chunked_obstacle_pass or test_collision_probability or test_static_boxes_fused"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frontend.py tests/test_gpu_exchange.py -m gpu -q -x -k "$SEL or frontend or two_contexts" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r02_sanitizer_memcheck.log | tail -4
timeout 300 python - <<'PY'
import json, bench
print(json.dumps(bench.planner_e2e(0)["config1"], indent=None))
PY
