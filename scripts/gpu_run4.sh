mkdir -p gpurun_out
for s in 1 2 4; do
FRX_SEG=$s timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_seg$s.log 2>&1; echo "seg $s pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_seg$s.log
done
: > gpurun_out/segsweep.txt
for wl in config2 config3; do for s in 1 2 4; do
FRX_SEG=$s timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl seg $s', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'frac=%.3f' % d['roofline']['frac'], 'value=%.3e' % d['value'], 'e2e=%.3e' % d['e2e']['value'], 'sel=', d.get('selected'))
" >> gpurun_out/segsweep.txt
done; done
cat gpurun_out/segsweep.txt
