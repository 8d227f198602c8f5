mkdir -p gpurun_out; : > gpurun_out/scale.txt
for s in 1 2 4 8 16; do
FRX_BENCH_SCALE=$s timeout 300 python bench.py --workload config2 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('scale $s', 'rows', d['config']['rows_total'], 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'frac=%.3f' % d['roofline']['frac'], 'value=%.3e' % d['value'])
" >> gpurun_out/scale.txt
done
cat gpurun_out/scale.txt
