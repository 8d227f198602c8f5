mkdir -p gpurun_out
for lib in build/var/libfrx_W6_C2.so build/var/libfrx_W6_C2_OPT_FASTSTEP0.so; do FRX_LIB=$PWD/$lib python scripts/ab_hash.py > gpurun_out/hash_$(basename $lib).txt 2>gpurun_out/hash.err; done
diff gpurun_out/hash_libfrx_W6_C2.so.txt gpurun_out/hash_libfrx_W6_C2_OPT_FASTSTEP0.so.txt && echo "A/B IDENTICAL" || echo "A/B DIFFER"
head -3 gpurun_out/hash_libfrx_W6_C2.so.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_sweep2.sh "1" config2 config3
