# scripts/gpu_r02_r.sh: deferred plans launch the eval-kernel instance without the fused obstacle pass -- parity slice + bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -k "split or chunk or block_shapes or scale or probability or fallback or golden_inputs" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_r.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_r.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()}, {k: (round(v["ms_per_step"], 4), {a: round(b, 4) for a, b in v["kernels_ms"].items()}) for k, v in d.get("also", {}).items()})
PY
