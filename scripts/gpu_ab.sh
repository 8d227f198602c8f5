# scripts/gpu_ab.sh libA libB [workloads]: bit-equality digest (scripts/ab_hash.py) and timing of two tuning builds
mkdir -p gpurun_out
A=$1; B=$2; shift 2; WL="${@:-config2}"
FRX_LIB=$PWD/$A python scripts/ab_hash.py > gpurun_out/hashA.txt 2>gpurun_out/hash.err
FRX_LIB=$PWD/$B python scripts/ab_hash.py > gpurun_out/hashB.txt 2>>gpurun_out/hash.err
diff gpurun_out/hashA.txt gpurun_out/hashB.txt > /dev/null && echo "A/B IDENTICAL ($(wc -l < gpurun_out/hashA.txt) cases)" || { echo "A/B DIFFER"; diff gpurun_out/hashA.txt gpurun_out/hashB.txt | head -5; }
for wl in $WL; do for lib in $A $B; do
FRX_LIB=$PWD/$lib python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', '$wl', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'frac=%.3f' % d['roofline']['frac'], 'value=%.3e' % d['value'], 'sel=', d.get('selected', {}).get('row'))
"; done; done
