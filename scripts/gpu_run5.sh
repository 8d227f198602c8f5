mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
: > gpurun_out/zc.txt
for zc in 1 0; do for wl in config2 config3; do for s in 1 4; do
FRX_ZEROCOPY=$zc FRX_SEG=$s timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl zerocopy $zc seg $s', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'frac=%.3f' % d['roofline']['frac'], 'value=%.3e' % d['value'], 'e2e=%.3e' % d['e2e']['value'], 'sel=', d.get('selected'))
" >> gpurun_out/zc.txt
done; done; done
cat gpurun_out/zc.txt
