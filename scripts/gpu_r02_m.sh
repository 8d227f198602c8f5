# scripts/gpu_r02_m.sh: new shape/staging test, then (gpurun --gpus N) the scaling bench at N = 1 and N GPUs
set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exchange.py -m gpu -q -k "block_shapes or exchange or two" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_m.log; tail -6 gpurun_out/pytest_m.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 rc=$?"
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/r02_scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "n$n rc=$?"; tail -3 gpurun_out/scale_n$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_scale_n?.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"] / 1e6, 1), "M/s", d["selected"]["row"], {k: round(v["ms_per_step"], 4) for k, v in d["also"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
