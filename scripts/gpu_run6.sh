mkdir -p gpurun_out
: > gpurun_out/zc.txt
for zc in 1 0; do for wl in config2; do for s in 1 4; do
FRX_ZEROCOPY=$zc FRX_SEG=$s timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl zerocopy $zc seg $s', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'e2e=%.3e' % d['e2e']['value'], {k: round(v, 4) for k, v in d['e2e'].items() if k.endswith('ms') or k.endswith('step')})
" >> gpurun_out/zc.txt
done; done; done
cat gpurun_out/zc.txt
