#!/bin/bash
# scripts/gpu_sweep.sh [workloads...] : bench every tuning build under build/var (one line per build and workload)
mkdir -p gpurun_out
WL="${@:-config2 config3}"
out=gpurun_out/sweep.txt; : > $out
for lib in build/var/libfrx_*.so; do
  for wl in $WL; do
    FRX_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', '$wl', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'frac=%.3f' % d['roofline']['frac'], 'value=%.3e' % d['value'], 'e2e=%.3e' % d['e2e']['value'], 'sel=', d.get('selected', {}).get('row'), 'clk=', d.get('clocks', {}).get('sm_mhz'))
" >> $out 2>&1 || echo "$lib $wl FAILED" >> $out
  done
done
cat $out
