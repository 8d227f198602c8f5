mkdir -p gpurun_out
nvidia-smi -L | head -3
for wl in config2 config5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 5 --warmup 3 > gpurun_out/bench2_$wl.json 2> gpurun_out/bench2_$wl.err
echo "rc=$?"; tail -c 900 gpurun_out/bench2_$wl.json; tail -2 gpurun_out/bench2_$wl.err
done
