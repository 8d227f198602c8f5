# scripts/gpu_sanitize.sh: compute-sanitizer memcheck + racecheck over a slice of the golden parity suite
set -x
mkdir -p gpurun_out
SEL="test_device_matches_oracle_on_golden_inputs or test_split_obstacle_kernel_equals_fused_pass or test_step_chunked_obstacle_pass or test_every_lanes_per_candidate"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frontend.py tests/test_gpu_exchange.py -m gpu -q -x -k "$SEL or frontend or two_contexts" \
      > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r02_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r02_sanitizer_$tool.log | tail -4
done
