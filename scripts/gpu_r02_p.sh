# scripts/gpu_r02_p.sh: ticket / round-robin dealing as two instances -- parity (shape x staging x dealing), A/B, default bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
bash scripts/gpu_r02_o.sh
unset FRX_OBS_TICKET
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()}, {k: (round(v["ms_per_step"], 4), round(v["kernels_ms"]["frx_obstacle_kernel"], 4)) for k, v in d.get("also", {}).items()})
PY
