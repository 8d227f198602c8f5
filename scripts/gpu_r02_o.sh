set -x
mkdir -p gpurun_out
for st in 0 1; do export FRX_OBS_TICKET=$((1-st))
  timeout 600 python bench.py --no-cpu-baseline --no-also --steps 10 > gpurun_out/st${st}_full.json 2> gpurun_out/st.err
  FRX_BENCH_C5_V=56 timeout 600 python bench.py --no-cpu-baseline --no-also --steps 10 > gpurun_out/st${st}_eighth.json 2>> gpurun_out/st.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/st?_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()})
PY
