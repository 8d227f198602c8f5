# scripts/gpu_r02_d.sh (gpurun --gpus 2): exchange tests on two GPUs, bench at N = 1 and N = 2 (both transports)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_exchange.py -m gpu -q > gpurun_out/pytest_exchange.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_exchange.log; tail -6 gpurun_out/pytest_exchange.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_scale_n2.json 2> gpurun_out/scale_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/scale_n2.err
FRX_BENCH_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r02_scale_n2_nccl.json 2> gpurun_out/scale_n2_nccl.err; echo "n2 nccl rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 3 > gpurun_out/r02_scale_n2_ref.json 2> gpurun_out/scale_n2_ref.err; echo "ref rc=$?"
ls -la gpurun_out | head
