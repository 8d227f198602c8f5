# scripts/gpu_sweep2.sh "<segs>" <workloads...>: every tuning build x every FRX_SEG
mkdir -p gpurun_out
SEGS="$1"; shift
WL="${@:-config2}"
out=gpurun_out/sweep2.txt; : > $out
for lib in build/var/libfrx_*.so; do for wl in $WL; do for s in $SEGS; do
    FRX_SEG=$s FRX_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-also 2> gpurun_out/sweep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', '$wl', 'seg $s', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'frac=%.3f' % d['roofline']['frac'], 'value=%.3e' % d['value'], 'sel=', d.get('selected', {}).get('row'))
" >> $out 2>&1 || echo "$lib $wl $s FAILED" >> $out
done; done; done
cat $out
