# scripts/gpu_r02_final.sh: the round's closing pass -- full GPU suite, smoke, default bench line, reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/bench_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()}, d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"])
r = json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1])
print(r["value"], r["ms_per_step"], r["cpu_baseline"])
PY
